"""Host-side mirror of the reference's layer API (src/ProbabilisticLayers.py), backed by the
sm_100a kernels in libpl_b200.so.

Same class names, constructor / forward signatures, parameter names (``weight.loc``,
``weight.logscale``, ``bias.loc``, ``bias.logscale``), initialisation and side-effect attributes
(``.kl_div``, ``.entropy``) as the reference, and the same MC-first tensor convention
(README.md:12): dim 0 = Monte-Carlo sample, dim 1 = batch.

What differs, by design (SURVEY.md appendix B / hard part H4):
  * sampled weights never exist in HBM.  ``layer.sampled_w``, ``layer.sampled_b`` and
    ``layer.weight.eps`` are lazy properties that re-materialise the last forward's draw on demand.
  * eps comes from a counter-based Philox stream (``eps_source='philox'``, default).  For parity
    work ``eps_source='torch_cpu'`` draws ``torch.randn`` on the CPU generator in the reference's
    order (weight then bias, src/ProbabilisticLayers.py:925-926) and injects it, and
    ``layer.inject_eps(eps_w, eps_b)`` injects arbitrary tensors.
  * there is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F
from torch.distributions import Normal

from . import _cabi
from . import functional as PF

# ------------------------------------------------------------------------------------------------
# global sampling configuration
# ------------------------------------------------------------------------------------------------
_STATE = {
    "seed": 0x5EEDB200,          # base Philox key; combined with a per-layer call counter
    "precision": "auto",         # see set_precision
    "eps_source": "philox",      # 'philox' | 'torch_cpu'
    "mc_global_offset": 0,       # first global MC index of this rank (MC sharding)
    "next_layer_id": 0,
    "seed_dev": None,            # int64[1] CUDA tensor added to every seed on the device (CUDA-graph replays)
}
_GOLDEN = 0x9E3779B97F4A7C15


def manual_seed(seed: int):
    """Base seed of the on-chip eps stream (all ranks of an MC-sharded job must agree)."""
    _STATE["seed"] = int(seed) & 0xFFFFFFFFFFFFFFFF


_PRECISION_NAMES = ("fp32", "tf32", "bf16", "auto", "auto_tf32", "bf16_fast", "tf32_fast", "auto_fast")


def set_precision(precision: str):
    """Arithmetic of the fused layers.
      'fp32'       exact FFMA kernels (any shape)
      'bf16'       tcgen05 bf16 operands, fp32 accumulate, mean and noise in SEPARATE operands
                   (Y = X.bf16(mu) + X.bf16(sigma*eps)): the Monte-Carlo noise survives however small sigma is
      'tf32'       the same with tf32 operands read straight from the fp32 tensors (primary-tolerance mode, rel 2e-3)
      'bf16_fast' / 'tf32_fast'   one rounded operand round(mu + sigma*eps), half the MMAs: only faithful once
                   sigma >~ ulp(mu) (NOT at the reference's initialisation, src/ProbabilisticLayers.py:881)
      'auto' / 'auto_tf32' / 'auto_fast'   the bf16 / tf32 / bf16_fast kernels where the layer shape tiles, fp32 elsewhere
    A forced tensor-core mode also falls back to fp32 for layers the tensor-core kernels cannot tile (odd widths,
    per-MC parameters); BayesConv2d has no tf32 kernel and runs 'tf32*' requests in fp32."""
    if precision not in _PRECISION_NAMES:
        raise ValueError(f"precision must be one of {_PRECISION_NAMES}, got {precision!r}")
    _STATE["precision"] = precision


def set_eps_source(source: str):
    if source not in ("philox", "torch_cpu"):
        raise ValueError(f"eps_source must be 'philox' or 'torch_cpu', got {source!r}")
    _STATE["eps_source"] = source


def set_seed_device_offset(t):
    """``t``: int64[1] CUDA tensor (or None).  While set, the kernels draw eps with key ``seed + t[0]`` (mod 2^64):
    a captured CUDA graph replays fixed launch parameters, so the step-dependent part of the seed is advanced on
    the device (train.ELBOTrainStep.enable_cuda_graph).  The lazy ``.eps`` / ``.sampled_w`` views do not follow it."""
    if t is not None and not (t.is_cuda and t.dtype == torch.int64 and t.numel() == 1):
        raise ValueError("seed offset must be an int64[1] CUDA tensor")
    _STATE["seed_dev"] = t


def set_mc_shard(mc_global_offset: int):
    """This rank owns global MC indices [offset, offset + local MC)."""
    _STATE["mc_global_offset"] = int(mc_global_offset)


def _new_layer_id():
    i = _STATE["next_layer_id"]
    _STATE["next_layer_id"] += 1
    return i


def _softplus_inv(x):
    return torch.log(torch.exp(x) - 1)


# ------------------------------------------------------------------------------------------------
# priors
# ------------------------------------------------------------------------------------------------
class GaussianPrior:
    """N(0, scale) prior (reference: src/ProbabilisticLayers.py:121-129)."""

    def __init__(self, scale):
        assert scale > 0.
        self.scale = scale

    def dist(self):
        return Normal(0, self.scale)


class LaplacePrior(torch.nn.Module):
    """Per-element Gaussian prior whose variance is an EMA of 1/grad^2 of ``module.weight.loc``
    (reference: src/ProbabilisticLayers.py:131-159).  The KL kernel consumes ``prior_scale()`` as
    its per-element prior-scale operand."""

    def __init__(self, module, clamp=False):
        super().__init__()
        self.scale = torch.ones_like(module.weight.loc.data)
        module.weight.loc.register_hook(self._save_grad)
        self.clamp = clamp
        self.step = 0
        self.beta = 0.99

    def _save_grad(self, grad):
        self.step += 1
        correction = 1 - self.beta ** self.step
        if self.scale.device != grad.device:
            self.scale = self.scale.to(grad.device)
        inv_sq = (1 / grad.detach() ** 2) / (correction + 1e-8)
        self.scale.mul_(self.beta).add_(inv_sq, alpha=1 - self.beta)

    def prior_scale(self):
        if self.clamp:
            return torch.clamp(self.scale ** 0.5, min=1.0)
        return self.scale ** 0.5 + 1e-8

    def dist(self):
        return Normal(0, self.prior_scale())


# ------------------------------------------------------------------------------------------------
# variational posterior
# ------------------------------------------------------------------------------------------------
class VariationalNormal(torch.nn.Module):
    """Factorised Gaussian posterior with parameters ``loc`` and ``logscale``
    (sigma = softplus(logscale)); reference: src/ProbabilisticLayers.py:811-841."""

    def __init__(self, loc, scale):
        super().__init__()
        assert loc.shape == scale.shape
        self.loc = torch.nn.Parameter(loc)
        self.logscale = torch.nn.Parameter(_softplus_inv(scale))
        self._eps = None
        self._eps_fn = None

    @property
    def batch_shape(self):
        return self.loc.shape

    def dist(self):
        return Normal(self.loc, F.softplus(self.logscale))

    @property
    def eps(self):
        """eps of the last draw, [MC, *loc.shape]; re-materialised lazily after a fused forward."""
        if self._eps is None and self._eps_fn is not None:
            self._eps = self._eps_fn()
        return self._eps

    @eps.setter
    def eps(self, value):
        self._eps = value
        self._eps_fn = None

    def _set_lazy_eps(self, fn):
        self._eps = None
        self._eps_fn = fn

    def rsample(self, sample_shape):
        """Eager compatibility path (materialises eps and the samples like the reference does);
        the fused layers do not call it."""
        shape = torch.Size(sample_shape) + self.loc.shape
        self.eps = torch.randn(shape, dtype=self.loc.dtype, device=self.loc.device)
        return self.loc + self.eps * F.softplus(self.logscale)


# ------------------------------------------------------------------------------------------------
# shared machinery of the fused layers
# ------------------------------------------------------------------------------------------------
class _FusedBayesLayer(torch.nn.Module):
    def _init_sampling(self):
        self.layer_id = _new_layer_id()
        self._calls = 0
        self._injected = None
        self._last = None
        self.eps_source: Optional[str] = None      # None -> global setting
        self.precision: Optional[str] = None       # None -> global setting
        # False: kl_div is computed without an autograd graph; the caller (train.ELBOTrainStep)
        # streams the analytic KL gradient into .grad with pl_kl_bwd(accumulate=1)
        self.kl_autograd = True
        # True: forward does not compute kl_div at all (train.ELBOTrainStep's fused optimiser pass
        # produces both the KL gradient and the KL value); kl_div reads as a zero scalar
        self.kl_skip = False

    def inject_eps(self, eps_w, eps_b=None):
        """Use these eps tensors ([MC,*weight.shape], [MC,*bias.shape]) for the next forward(s);
        ``inject_eps(None)`` returns to the configured source."""
        self._injected = None if eps_w is None else (eps_w, eps_b)

    def _precision_code(self, shape_ok_for_tc: bool, tc_possible: bool = None, conv: bool = False):
        """Resolves the configured precision for this call.  ``shape_ok_for_tc``: the tensor-core kernels pay off
        (the 'auto*' modes); ``tc_possible``: they can run at all (forced modes fall back to fp32 otherwise)."""
        p = self.precision or _STATE["precision"]
        if tc_possible is None:
            tc_possible = shape_ok_for_tc
        if p.startswith("auto"):
            p = {"auto": "bf16", "auto_tf32": "tf32", "auto_fast": "bf16_fast"}[p] if shape_ok_for_tc else "fp32"
        elif p != "fp32" and not tc_possible:
            p = "fp32"
        if conv and p.startswith("tf32"):
            p = "fp32"          # no tf32 conv kernel: keep the tolerance class (rel 2e-3) rather than drop to bf16
        return _cabi.PRECISIONS[p]

    def _draw(self, mc, device, tc_ok, tc_possible=None, conv=False):
        """Returns (eps_w, eps_b, SampleSpec) for this call."""
        self._calls += 1
        seed = (_STATE["seed"] + self._calls * _GOLDEN) & 0xFFFFFFFFFFFFFFFF
        spec = PF.SampleSpec(seed=seed, layer_id=self.layer_id,
                             mc_global_offset=_STATE["mc_global_offset"],
                             precision=self._precision_code(tc_ok, tc_possible, conv))
        if _STATE["seed_dev"] is not None:
            if self._injected is not None or (self.eps_source or _STATE["eps_source"]) != "philox":
                raise RuntimeError("a device-side seed offset (CUDA-graph mode) needs the on-chip Philox sampler")
            spec.seed_dev = _STATE["seed_dev"].data_ptr()
        eps_w = eps_b = None
        if self._injected is not None:
            eps_w, eps_b = self._injected
            eps_w = eps_w.to(device=device, dtype=torch.float32)
            eps_b = eps_b.to(device=device, dtype=torch.float32) if eps_b is not None else None
            if self.bias is not None and eps_b is None:
                raise ValueError("inject_eps: layer has a bias, eps_b is required")
        elif (self.eps_source or _STATE["eps_source"]) == "torch_cpu":
            # the reference's draw order: weight eps, then bias eps (ProbabilisticLayers.py:925-926)
            eps_w = torch.randn((mc,) + tuple(self.weight.loc.shape)).to(device)
            if self.bias is not None:
                eps_b = torch.randn((mc,) + tuple(self.bias.loc.shape)).to(device)
        return eps_w, eps_b, spec

    def _remember(self, mc, eps_w, eps_b, spec):
        """Wire the lazy ``weight.eps`` / ``bias.eps`` / ``sampled_w`` / ``sampled_b`` views."""
        self._last = (mc, eps_w, eps_b, spec)
        dev = self.weight.loc.device
        wshape = tuple(self.weight.loc.shape)
        rows, cols = wshape[0], int(math.prod(wshape[1:]))

        def w_eps():
            if eps_w is not None:
                return eps_w
            return PF.eps_dump(spec.seed, 2 * spec.layer_id, spec.mc_global_offset, mc, rows, cols,
                               dev).reshape((mc,) + wshape)

        self.weight._set_lazy_eps(w_eps)
        if self.bias is not None:
            n = self.bias.loc.numel()

            def b_eps():
                if eps_b is not None:
                    return eps_b
                return PF.eps_dump(spec.seed, 2 * spec.layer_id + 1, spec.mc_global_offset, mc, 1,
                                   n, dev).reshape(mc, n)

            self.bias._set_lazy_eps(b_eps)

    @property
    def sampled_w(self):
        """W[mc] = loc + softplus(logscale) * eps[mc] of the last forward (debug view)."""
        return self.weight.loc + self.weight.eps * F.softplus(self.weight.logscale)

    @property
    def sampled_b(self):
        if self.bias is None:
            return None
        return self.bias.loc + self.bias.eps * F.softplus(self.bias.logscale)


# ------------------------------------------------------------------------------------------------
# BayesLinear
# ------------------------------------------------------------------------------------------------
class BayesLinear(_FusedBayesLayer):
    """Bayes-by-backprop linear layer over [MC, B, in] inputs
    (reference: src/ProbabilisticLayers.py:844-961)."""

    def __init__(self, in_features, out_features, num_MC=None, prior=1., bias=True):
        super().__init__()
        self.dim_input = in_features
        self.dim_output = out_features
        self.num_MC = num_MC
        self.mu_init_std = torch.sqrt(torch.scalar_tensor(2 / (in_features + out_features)))
        self.logsigma_init_std = 0.001
        # same RNG consumption as the reference ctor (one normal_ per loc), so that a seeded
        # construction yields an identical state_dict
        self.weight = VariationalNormal(
            torch.empty(in_features, out_features).normal_(0., float(self.mu_init_std)),
            torch.full((in_features, out_features), self.logsigma_init_std))
        if bias:
            self.bias = VariationalNormal(
                torch.empty(out_features).normal_(0., float(self.mu_init_std)),
                torch.full((out_features,), self.logsigma_init_std))
        else:
            self.bias = None
        if prior == 'laplace':
            self.prior = LaplacePrior(module=self)
        elif prior == 'laplace_clamp':
            self.prior = LaplacePrior(module=self, clamp=True)
        else:
            self.prior = GaussianPrior(scale=float(prior))
        self._init_sampling()
        self.reset_parameters(scale_offset=0)

    def reset_parameters(self, scale_offset=0):
        # weight layout is [in, out]; kaiming_uniform_ on that layout gives U(+-1/sqrt(out))
        torch.nn.init.kaiming_uniform_(self.weight.loc.data, a=math.sqrt(5))
        self.weight.logscale.data.fill_(
            float(_softplus_inv(self.mu_init_std / self.weight.loc.shape[1]) + scale_offset))
        if self.bias is not None:
            fan_in, _ = torch.nn.init._calculate_fan_in_and_fan_out(self.weight.loc.data)
            bound = 1 / math.sqrt(fan_in)
            torch.nn.init.uniform_(self.bias.loc.data, -bound, bound)
            self.bias.logscale.data.fill_(float(_softplus_inv(self.mu_init_std) + scale_offset))

    def _tc_ok(self, mc, batch):
        # what the tcgen05 kernels tile without waste: 16-byte aligned rows, at least one full
        # 128-wide tile in every GEMM dimension
        return (self.dim_input % 8 == 0 and self.dim_output % 8 == 0 and batch >= 128
                and self.dim_input >= 128 and self.dim_output >= 128)

    def _tc_possible(self):
        # hard limits of the tcgen05 kernels: 16-byte aligned bf16 rows in both weight dimensions
        return self.dim_input % 8 == 0 and self.dim_output % 8 == 0 and self.dim_input >= 8 and self.dim_output >= 8

    def _narrow_head_pad(self, batch):
        """Columns to pad a NARROW output to for the tensor-core kernels (0: not applicable).  A classifier head such as
        C2's 1200 -> 10 is latency bound in the fp32 kernels (0.29 ms of a 0.97 ms step); padded to a multiple of 8 columns
        the tcgen05 kernels run it in one 128-wide tile.  The eps counter is indexed by (row, column), so the first
        ``out_features`` columns draw exactly the eps of the unpadded layer."""
        p = self.precision or _STATE["precision"]
        if p not in ("auto", "auto_fast", "auto_tf32") or self.dim_output >= 128 or self._injected is not None:
            return 0
        if (self.eps_source or _STATE["eps_source"]) != "philox" or isinstance(self.prior, LaplacePrior):
            return 0
        if self.dim_input % 8 or self.dim_input < 256 or batch < 128:
            return 0
        return -(-self.dim_output // 8) * 8

    def _forward_padded_head(self, x, o_pad):
        """Tensor-core forward of a narrow layer through zero-padded parameter views: sigma = 0 in the padded columns
        (rho = -100), outputs and gradients of those columns are sliced away by autograd."""
        mc, pad = x.shape[0], o_pad - self.dim_output
        w_mu = F.pad(self.weight.loc, (0, pad))
        w_rho = F.pad(self.weight.logscale, (0, pad), value=-100.0)
        b_mu = None if self.bias is None else F.pad(self.bias.loc, (0, pad))
        b_rho = None if self.bias is None else F.pad(self.bias.logscale, (0, pad), value=-100.0)
        self._calls += 1
        seed = (_STATE["seed"] + self._calls * _GOLDEN) & 0xFFFFFFFFFFFFFFFF
        base = {"auto": "bf16", "auto_fast": "bf16_fast", "auto_tf32": "tf32"}[self.precision or _STATE["precision"]]
        spec = PF.SampleSpec(seed=seed, layer_id=self.layer_id, mc_global_offset=_STATE["mc_global_offset"],
                             precision=_cabi.PRECISIONS[base])
        if _STATE["seed_dev"] is not None:
            spec.seed_dev = _STATE["seed_dev"].data_ptr()
        out = PF.bayes_linear(x, w_mu, w_rho, b_mu, b_rho, None, None, spec)
        # lazy eps / sampled_w views of the REAL columns (same counter: the unpadded spec dumps the same numbers)
        self._remember(mc, None, None, spec)
        return out[..., :self.dim_output]

    def forward(self, x: torch.Tensor, prior=None, stochastic=True):
        assert x.dim() == 3, f"Input tensor not of shape [N_MC, BatchSize, Features] but is {x.shape=}"
        mc = x.shape[0]
        o_pad = self._narrow_head_pad(x.shape[1]) if x.is_cuda else 0
        if o_pad:
            out = self._forward_padded_head(x, o_pad)
            eps_w = eps_b = None
        else:
            eps_w, eps_b, spec = self._draw(mc, x.device, self._tc_ok(mc, x.shape[1]), self._tc_possible())
            out = PF.bayes_linear(x, self.weight.loc, self.weight.logscale,
                                  None if self.bias is None else self.bias.loc,
                                  None if self.bias is None else self.bias.logscale,
                                  eps_w, eps_b, spec)
            self._remember(mc, eps_w, eps_b, spec)
        if isinstance(self.prior, LaplacePrior):
            pse = self.prior.prior_scale().to(x.device)
            self.kl_div, self.entropy = PF.kl_and_entropy(self.weight.loc, self.weight.logscale,
                                                          1.0, pse)
        elif self.kl_skip:
            self.kl_div = self.entropy = torch.zeros((), device=x.device)
        elif self.kl_autograd:
            self.kl_div, self.entropy = PF.kl_and_entropy(self.weight.loc, self.weight.logscale,
                                                          self.prior.scale)
        else:
            with torch.no_grad():
                self.kl_div, self.entropy = PF.kl_and_entropy(
                    self.weight.loc.detach(), self.weight.logscale.detach(), self.prior.scale)
        return out

    def __repr__(self):
        return f'BayesLinear(in_features={self.dim_input}, out_features={self.dim_output}'


# ------------------------------------------------------------------------------------------------
# BayesConv2d
# ------------------------------------------------------------------------------------------------
class BayesConv2d(_FusedBayesLayer):
    """Bayes-by-backprop conv layer over [MC, B, C, H, W] inputs; per-MC filters
    (reference: src/ProbabilisticLayers.py:963-1061).  Returns a contiguous [MC,B,Cout,H',W']
    tensor (the reference returns a permuted view of the same logical shape)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 num_MC=1, prior_scale=1):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = kernel_size
        self.stride = stride
        self.padding = padding
        self.dilation = dilation
        self.epsilon_sigma = 1.0
        self.num_MC = num_MC
        self.prior_scale = prior_scale
        self.mu_init_std = torch.scalar_tensor(1. / math.sqrt(in_channels * kernel_size ** 2))
        self.logsigma_init_std = 0.001
        wshape = (out_channels, in_channels, kernel_size, kernel_size)
        self.weight = VariationalNormal(torch.empty(wshape).normal_(0, float(self.mu_init_std)),
                                        torch.full(wshape, self.logsigma_init_std))
        self.bias = VariationalNormal(torch.empty(out_channels).normal_(0., float(self.mu_init_std)),
                                      torch.full((out_channels,), self.logsigma_init_std))
        self._init_sampling()
        self.reset_parameters(scale_offset=0.)

    def reset_parameters(self, scale_offset=0):
        torch.nn.init.kaiming_uniform_(self.weight.loc.data, a=math.sqrt(5))
        self.weight.logscale.data.fill_(
            float(_softplus_inv((self.mu_init_std + scale_offset) / self.weight.loc.shape[1])))
        fan_in, _ = torch.nn.init._calculate_fan_in_and_fan_out(self.weight.loc.data)
        bound = 1 / math.sqrt(fan_in)
        torch.nn.init.uniform_(self.bias.loc.data, -bound, bound)
        self.bias.logscale.data.fill_(float(_softplus_inv(self.mu_init_std + scale_offset)))

    def forward(self, x: torch.Tensor, stochastic=True, prior=None):
        assert x.dim() == 5, f'Input tensor not of shape [num_MC, batch_size in_channels, height, width], but {x.shape=}'
        assert x.shape[0] == self.num_MC, f'Input.shape={x.shape} with [{x.shape[0]=}]!=[{self.num_MC=}]'
        mc = x.shape[0]
        # tensor cores pay off once the per-sample GEMM has a few full tiles
        k = self.kernel_size
        ho = (x.shape[3] + 2 * self.padding - self.dilation * (k - 1) - 1) // self.stride + 1
        wo = (x.shape[4] + 2 * self.padding - self.dilation * (k - 1) - 1) // self.stride + 1
        tc_ok = (x.shape[1] * ho * wo >= 512 and self.out_channels >= 32
                 and self.in_channels * k * k >= 32)
        eps_w, eps_b, spec = self._draw(mc, x.device, tc_ok, True, conv=True)
        out = PF.bayes_conv2d(x, self.weight.loc, self.weight.logscale, self.bias.loc,
                              self.bias.logscale, self.kernel_size, self.stride, self.padding,
                              self.dilation, eps_w, eps_b, spec)
        self._remember(mc, eps_w, eps_b, spec)
        if self.kl_skip:
            self.kl_div = torch.zeros((), device=x.device)
        elif self.kl_autograd:
            self.kl_div, _ = PF.kl_and_entropy(self.weight.loc, self.weight.logscale, self.prior_scale)
        else:
            with torch.no_grad():
                self.kl_div, _ = PF.kl_and_entropy(self.weight.loc.detach(),
                                                   self.weight.logscale.detach(), self.prior_scale)
        return out


# ------------------------------------------------------------------------------------------------
# SIVI last layer
# ------------------------------------------------------------------------------------------------
class SIVILayer(torch.nn.Module):
    """Semi-implicit VI layer: a hyper-network maps noise to per-MC (mu, sigma)
    (reference: src/ProbabilisticLayers.py:257-335).  The hyper-net (three tiny Linear layers) stays
    in torch; sampling + contraction + their gradients run in the fused kernel's per-MC-parameter
    mode.  Like the reference, the bias broadcast is only defined for out_features == 1."""

    def init_weights(self, module):
        if type(module) == torch.nn.Linear:
            torch.nn.init.orthogonal_(module.weight, gain=1.)
            module.bias.data.normal_(0., 0.01)

    def __init__(self, in_features, out_features, _dim_noise_input):
        super().__init__()
        self.dim_input = in_features
        self.dim_output = out_features
        self.dim_noise_input = _dim_noise_input
        self.dim_output_params = in_features * out_features * 2 + out_features * 2
        self.num_hidden = min(self.dim_output_params,
                              int((_dim_noise_input + self.dim_output_params) / 2))
        self.prior_sigma = torch.scalar_tensor(1.0)
        self.sivi_net = torch.nn.Sequential(
            torch.nn.Linear(self.dim_noise_input, self.num_hidden), torch.nn.Tanh(),
            torch.nn.Linear(self.num_hidden, self.num_hidden), torch.nn.Tanh(),
            torch.nn.Linear(self.num_hidden, self.dim_output_params))
        self.sivi_net.apply(self.init_weights)
        self.layer_id = _new_layer_id()
        self._calls = 0
        self._injected = None
        self.eps_source: Optional[str] = None

    def inject_noise(self, noise, eps_w, eps_b):
        self._injected = None if noise is None else (noise, eps_w, eps_b)

    def forward(self, x):
        assert x.dim() == 3, 'Input tensor not of shape [N_MC, BatchSize, Features]'
        if self.dim_output != 1:
            raise NotImplementedError("SIVILayer: the reference's bias broadcast is only defined "
                                      "for out_features == 1 (src/ProbabilisticLayers.py:326)")
        num_MC = x.shape[0]
        dev = x.device
        self._calls += 1
        eps_w = eps_b = None
        if self._injected is not None:
            noise, eps_w, eps_b = (t.to(device=dev, dtype=torch.float32) for t in self._injected)
        elif (self.eps_source or _STATE["eps_source"]) == "torch_cpu":
            # reference draw order: noise, eps_w, eps_b
            noise = torch.randn(num_MC, self.dim_noise_input).to(dev)
            eps_w = torch.randn(num_MC, self.dim_input, self.dim_output).to(dev)
            eps_b = torch.randn(num_MC, self.dim_output).to(dev)
        else:
            seed = (_STATE["seed"] + self._calls * _GOLDEN) & 0xFFFFFFFFFFFFFFFF
            # hyper-net noise from the same counter stream (bias stream id + 2^16 keeps it disjoint), drawn on the
            # device: with a device-side seed offset the draw can be captured in a CUDA graph
            sd = _STATE["seed_dev"]
            noise = PF.eps_dump(seed, 2 * self.layer_id + (1 << 16), _STATE["mc_global_offset"],
                                num_MC, 1, self.dim_noise_input, dev,
                                seed_dev=sd.data_ptr() if sd is not None else None).reshape(num_MC, -1)
        sivi_out = self.sivi_net(noise)
        nw = self.dim_input * self.dim_output
        w_mu, w_logsigma = sivi_out[:, :nw], sivi_out[:, nw:2 * nw]
        b_mu, b_logsigma = sivi_out[:, 2 * nw:2 * nw + self.dim_output], sivi_out[:, 2 * nw + self.dim_output:]
        self.w_mu = w_mu.reshape(num_MC, self.dim_input, self.dim_output)
        self.w_std = F.softplus(w_logsigma.reshape(num_MC, self.dim_input, self.dim_output))
        self.b_mu = b_mu
        self.b_std = F.softplus(b_logsigma)
        seed = (_STATE["seed"] + self._calls * _GOLDEN) & 0xFFFFFFFFFFFFFFFF
        spec = PF.SampleSpec(seed=seed, layer_id=self.layer_id,
                             mc_global_offset=_STATE["mc_global_offset"],
                             precision=_cabi.PREC_FP32)
        if _STATE["seed_dev"] is not None:
            if eps_w is not None:
                raise RuntimeError("a device-side seed offset (CUDA-graph mode) needs the on-chip Philox sampler")
            spec.seed_dev = _STATE["seed_dev"].data_ptr()
        out = PF.bayes_linear(x, self.w_mu, self.w_std, self.b_mu, self.b_std, eps_w, eps_b, spec,
                              param_mode=_cabi.PARAM_PER_MC_SIGMA)
        # KL(N(mu_mc, std_mc) || N(0,1)).mean(0).sum(), weights only (:330-333): [MC*I*O] elements,
        # tiny -- stays in torch so that it differentiates through the hyper-net
        kl_w = 0.5 * (self.w_std ** 2 + self.w_mu ** 2 - 1.0) - torch.log(self.w_std)
        self.kl_div = kl_w.mean(dim=0).sum()
        return out


# ------------------------------------------------------------------------------------------------
# layers around the hot path (SURVEY.md 8f "next" rows keep torch ops underneath for now)
# ------------------------------------------------------------------------------------------------
class MC_ExpansionLayer(torch.nn.Module):
    """[B, ...] -> [MC, B, ...] (reference: src/ProbabilisticLayers.py:765-785).  Returns a
    stride-0 expanded view instead of ``repeat``'s MC-fold copy; the fused layers read it with an
    MC stride of 0, so the duplicated input never exists in HBM."""

    def __init__(self, num_MC=1, input_dim=2):
        super().__init__()
        self.num_MC = num_MC
        self.input_dim = input_dim

    def forward(self, x):
        if x.dim() == self.input_dim:
            return x.unsqueeze(0).expand(self.num_MC, *x.shape)
        if x.dim() == self.input_dim + 1:
            return x
        raise ValueError(f"Input.dim()={x.dim()}, but should be either {self.input_dim} and expanded or {self.input_dim+1}")


class MC_MaxPool2d(torch.nn.Module):
    """Max-pool each [H, W] plane of a [MC, B, C, H, W] tensor
    (reference: src/ProbabilisticLayers.py:207-236; pooling is per plane, so folding MC and B is
    equivalent to the reference's fold of MC into channels without its two permute copies)."""

    def __init__(self, kernel_size=2, stride=2, num_MC=None):
        super().__init__()
        self.kernel_size = kernel_size
        self.stride = stride

    def forward(self, x, num_MC=None):
        assert x.dim() == 5, 'Input tensor does not have dimensionality of 5 [num_MC, Batch_Size, Ci, H, W]'
        if num_MC != None:
            assert isinstance(num_MC, int)
            assert x.shape[0] == num_MC
        mc, b = x.shape[0], x.shape[1]
        out = F.max_pool2d(x.reshape(mc * b, *x.shape[2:]), kernel_size=self.kernel_size,
                           stride=self.stride)
        return out.reshape(mc, b, *out.shape[1:])


class _MCBatchNorm(torch.nn.modules.batchnorm._BatchNorm):
    def _check_input_dim(self, input):
        return

    def _factor(self):
        factor = 0.0 if self.momentum is None else self.momentum
        if self.training and self.track_running_stats and self.num_batches_tracked is not None:
            self.num_batches_tracked += 1
            factor = (1.0 / float(self.num_batches_tracked)) if self.momentum is None else self.momentum
        return factor


class MC_BatchNorm1D(_MCBatchNorm):
    """BatchNorm over [MC, B, F] with statistics pooled over MC*B
    (reference: src/ProbabilisticLayers.py:699-734)."""

    def forward(self, x):
        assert x.dim() == 3, 'Input tensor does not have dimensionality of 5 [num_MC, Batch_Size, Features]'
        assert self.num_features == x.shape[-1]
        mc, bs, feat = x.shape
        out = F.batch_norm(x.reshape(mc * bs, feat), self.running_mean, self.running_var,
                           self.weight, self.bias, self.training or not self.track_running_stats,
                           self._factor(), self.eps)
        return out.reshape(mc, bs, feat)


class MC_BatchNorm2D(_MCBatchNorm):
    """BatchNorm over [MC, B, C, H, W] with per-channel statistics pooled over MC, B, H, W jointly
    (reference: src/ProbabilisticLayers.py:736-762)."""

    def forward(self, x):
        assert x.dim() == 5, 'Input tensor does not have dimensionality of 5 [num_MC, Batch_Size, Features]'
        assert self.num_features == x.shape[2]
        mc, bs = x.shape[0], x.shape[1]
        out = F.batch_norm(x.reshape(mc * bs, *x.shape[2:]), self.running_mean, self.running_var,
                           self.weight, self.bias, self.training or not self.track_running_stats,
                           self._factor(), self.eps)
        return out.reshape(x.shape)


class BayesAdaptiveInit_FlattenAndLinear(torch.nn.Module):
    """flatten(2,-1) then a lazily created BayesLinear
    (reference: src/ProbabilisticLayers.py:787-808)."""

    def __init__(self, num_hidden):
        super().__init__()
        self.adaptively_initialized = False
        self.num_hidden = num_hidden

    def forward(self, x):
        x = x.flatten(2, -1)
        if not self.adaptively_initialized:
            self.adaptively_initialized = True
            self.linear = BayesLinear(x.shape[-1], self.num_hidden).to(x.device)
        return self.linear(x)


class BayesianNeuralNetwork(torch.nn.Module):
    """Base class kept for import compatibility (reference: src/ProbabilisticLayers.py:1063-1091)."""

    def __init__(self, num_MC):
        super().__init__()
        self.num_MC = num_MC

    def forward(self, x):
        return x.unsqueeze(0).expand(self.num_MC, *x.shape)

    def collect_kl_div(self):
        raise NotImplementedError
