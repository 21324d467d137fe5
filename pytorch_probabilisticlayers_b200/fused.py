"""Opt-in fusion of a BayesLinear -> LeakyReLU -> BayesLinear -> ... chain.

The reference assembles its MLPs as ``Sequential(BayesLinear, LeakyReLU, BayesLinear, ...)``
(experiments/BNN.py:212-226).  Run module by module, every hidden activation makes four extra trips
through HBM in fp32 (y write, LeakyReLU read+write, bf16 cast read) and the same again backwards.
``FusedBayesMLP`` owns the whole chain: the GEMM epilogue applies the activation and writes the bf16
operand the next layer's TMA loads directly, and the dX epilogue applies the activation derivative and
writes the bf16 dY of the previous layer (C ABI: the ``x_lp / y_lp / dy_lp / dx_lp`` fields of
``pl_linear_args``).  Same arithmetic as the unfused bf16 path (identical bf16 MMA operands); only the
bias gradients see bf16-rounded dY.  ``fuse_bayes_mlp(net)`` rewrites a Sequential in place-compatible
fashion: the BayesLinear modules (and their parameters / state_dict keys) are reused, not copied.
"""
from __future__ import annotations

import ctypes as C
from typing import List

import torch

from . import _cabi
from . import functional as PF
from .layers import BayesLinear, LaplacePrior


class _FusedMLPFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, slope, specs, *params):
        lib = _cabi.lib()
        nl = len(specs)
        xb, x_stride, mc = PF._mc_view(x, 2)
        batch = x.shape[1]
        dev = x.device
        stream = _cabi.current_stream(dev)
        acts: List[torch.Tensor] = []       # bf16 activated outputs of the hidden layers
        states = []
        y = None
        with torch.cuda.device(dev):
            for li in range(nl):
                w_mu, w_rho, b_mu, b_rho = params[4 * li:4 * li + 4]
                fin, fout = w_mu.shape
                last = li == nl - 1
                a = _cabi.LinearArgs()
                a.mc, a.batch, a.in_features, a.out_features = mc, batch, fin, fout
                a.precision, a.param_mode = specs[li].precision, _cabi.PARAM_SHARED_RHO
                if li == 0:
                    a.x, a.x_mc_stride = _cabi.ptr(xb), x_stride
                else:
                    a.x_lp, a.x_mc_stride = _cabi.ptr(acts[-1]), batch * fin
                a.w_mu, a.w_rho = _cabi.ptr(w_mu), _cabi.ptr(w_rho)
                a.b_mu, a.b_rho = _cabi.ptr(b_mu), _cabi.ptr(b_rho)
                sp = specs[li]
                a.seed, a.layer_id, a.mc_global_offset = sp.seed, sp.layer_id, sp.mc_global_offset
                a.seed_dev = sp.seed_dev or None
                a.w_lp, a.w_lp_half = sp.w_lp or None, sp.w_lp_half
                if last:
                    y = torch.empty((mc, batch, fout), dtype=torch.float32, device=dev)
                    a.y = _cabi.ptr(y)
                else:
                    act = torch.empty((mc, batch, fout), dtype=torch.bfloat16, device=dev)
                    a.y_lp, a.y_act, a.y_act_slope = _cabi.ptr(act), 1, float(slope)
                state = torch.empty(int(lib.pl_linear_state_bytes(C.byref(a))), dtype=torch.uint8, device=dev)
                a.fwd_state, a.fwd_state_bytes = _cabi.ptr(state), state.numel()
                ws = PF._workspace(lib.pl_linear_workspace_bytes(C.byref(a)), dev)
                a.workspace, a.workspace_bytes = _cabi.ptr(ws), ws.numel()
                a.stream = stream
                PF._timed(f"linear_fwd[{mc}x{batch}x{fin}x{fout}]",
                          lambda: _cabi.check(lib.pl_linear_fwd(C.byref(a)), "pl_linear_fwd"),
                          ("flops", 2.0 * mc * batch * fin * fout))
                states.append(state)
                if not last:
                    acts.append(act)
        ctx.save_for_backward(*params, *acts, *states)
        ctx.meta = (specs, float(slope), x_stride, mc, batch, nl)
        return y

    @staticmethod
    def backward(ctx, dy):
        specs, slope, x_stride, mc, batch, nl = ctx.meta
        saved = ctx.saved_tensors
        params = saved[:4 * nl]
        acts = saved[4 * nl:4 * nl + nl - 1]
        states = saved[4 * nl + nl - 1:]
        lib = _cabi.lib()
        dy = dy.contiguous()
        dev = dy.device
        stream = _cabi.current_stream(dev)
        grads = [None] * (4 * nl)
        dy_lp = None
        dx0 = None                                      # gradient of the chain's input, when something upstream needs it
        with torch.cuda.device(dev):
            for li in reversed(range(nl)):
                w_mu, w_rho, b_mu, b_rho = params[4 * li:4 * li + 4]
                fin, fout = w_mu.shape
                a = _cabi.LinearArgs()
                a.mc, a.batch, a.in_features, a.out_features = mc, batch, fin, fout
                a.precision, a.param_mode = specs[li].precision, _cabi.PARAM_SHARED_RHO
                a.w_mu, a.w_rho = _cabi.ptr(w_mu), _cabi.ptr(w_rho)
                a.b_mu, a.b_rho = _cabi.ptr(b_mu), _cabi.ptr(b_rho)
                sp = specs[li]
                a.seed, a.layer_id, a.mc_global_offset = sp.seed, sp.layer_id, sp.mc_global_offset
                a.seed_dev = sp.seed_dev or None
                a.w_lp, a.w_lp_half, a.dw_rho_raw = sp.w_lp or None, sp.w_lp_half, sp.dw_rho_raw
                a.fwd_state, a.fwd_state_bytes = _cabi.ptr(states[li]), states[li].numel()
                if li == 0:
                    a.x_mc_stride = x_stride            # layer 0 reads the bf16 x kept in its state
                else:
                    a.x_lp, a.x_mc_stride = _cabi.ptr(acts[li - 1]), batch * fin
                if li == nl - 1:
                    a.dy = _cabi.ptr(dy)
                else:
                    a.dy_lp = _cabi.ptr(dy_lp)
                if sp.grad_out is not None:         # the trainer's flat gradient buffer: written in place, nothing for
                    dw_mu, dw_rho, db_mu, db_rho = sp.grad_out      # autograd to accumulate (saves a 12 B/parameter pass)
                else:
                    dw_mu, dw_rho = torch.empty_like(w_mu), torch.empty_like(w_rho)
                    db_mu, db_rho = torch.empty_like(b_mu), torch.empty_like(b_rho)
                a.dw_mu, a.dw_rho, a.db_mu, a.db_rho = (_cabi.ptr(t) for t in (dw_mu, dw_rho, db_mu, db_rho))
                sc = sp.dw_scatter

                def set_scatter(on):
                    if sc is None:
                        return
                    a.dw_scatter_world = sc["world"] if on else 0
                    if on:                      # the weight gradients go straight to their owners' staging rows
                        a.dw_scatter_peers = C.cast(sc["peers"], C.c_void_p)
                        a.dw_scatter_mu_off, a.dw_scatter_rho_off = sc["mu_off"], sc["rho_off"]
                        a.dw_scatter_chunk, a.dw_scatter_rank, a.dw_scatter_bf16 = sc["chunk"], sc["rank"], sc["bf16"]
                        a.dw_mu = a.dw_rho = None
                set_scatter(True)
                dx_lp = None
                if li > 0:                              # dX * act'(x) as the bf16 dY of layer li-1
                    dx_lp = torch.empty((mc, batch, fin), dtype=torch.bfloat16, device=dev)
                    a.dx_lp, a.x_act_slope = _cabi.ptr(dx_lp), slope
                elif ctx.needs_input_grad[0]:           # the chain is not the first thing in the net
                    dx0 = torch.empty((mc, batch, fin), dtype=torch.float32, device=dev)
                    a.dx = _cabi.ptr(dx0)
                ws = PF._workspace(lib.pl_linear_workspace_bytes(C.byref(a)), dev)
                a.workspace, a.workspace_bytes = _cabi.ptr(ws), ws.numel()
                a.stream = stream
                shape = f"[{mc}x{batch}x{fin}x{fout}]"
                flops = 2.0 * mc * batch * fin * fout
                if PF._PROFILE is None:
                    _cabi.check(lib.pl_linear_bwd(C.byref(a)), "pl_linear_bwd")
                else:                                   # one call per output group, timed separately
                    outs = (a.dx_lp, a.dw_mu, a.dw_rho, a.db_mu, a.db_rho, a.dx)
                    for tag, keep in (("dx", (0, 5)), ("dw", (1, 2)), ("db", (3, 4))):
                        sel = [o if i in keep else None for i, o in enumerate(outs)]
                        set_scatter(tag == "dw")
                        if all(v is None for v in sel) and not (tag == "dw" and sc is not None):
                            continue
                        a.dx_lp, a.dw_mu, a.dw_rho, a.db_mu, a.db_rho, a.dx = sel
                        PF._timed(f"linear_bwd_{tag}{shape}",
                                  lambda: _cabi.check(lib.pl_linear_bwd(C.byref(a)), "pl_linear_bwd"),
                                  ("flops", flops) if tag != "db" else ("bytes", 2.0 * mc * batch * fout))
                    a.dx_lp, a.dw_mu, a.dw_rho, a.db_mu, a.db_rho, a.dx = outs
                if sp.grad_out is None:
                    grads[4 * li:4 * li + 4] = [dw_mu, dw_rho, db_mu, db_rho]
                elif sp.after_bwd is not None:          # e.g. start this layer's gradient transfer on a side stream
                    sp.after_bwd()
                dy_lp = dx_lp
        return (dx0, None, None, *grads)


class FusedBayesMLP(torch.nn.Module):
    """A chain of BayesLinear layers with LeakyReLU(negative_slope) between them, run as one fused
    forward/backward.  Input [MC, B, in] (or an MC-expanded view), output [MC, B, out] fp32."""

    def __init__(self, layers: List[BayesLinear], negative_slope: float = 0.01):
        super().__init__()
        for l in layers:
            if not isinstance(l, BayesLinear) or l.bias is None:
                raise TypeError("FusedBayesMLP needs BayesLinear layers with a bias")
        for a, b in zip(layers[:-1], layers[1:]):
            if a.dim_output != b.dim_input:
                raise ValueError("layer widths do not chain")
        self.layers = torch.nn.ModuleList(layers)
        self.negative_slope = float(negative_slope)

    @staticmethod
    def supported(layers, batch) -> bool:
        """The chain hands bf16 activations from layer to layer, so every layer must resolve to a bf16 tensor-core
        precision ('auto', 'auto_fast', 'bf16', 'bf16_fast'): a user who configured fp32 / tf32 is never downgraded."""
        return all(l.bias is not None and l._tc_ok(0, batch) and l._injected is None
                   and not isinstance(l.prior, LaplacePrior)
                   and l._precision_code(True, True) in _cabi.BF16_CLASS for l in layers)

    def forward(self, x):
        assert x.dim() == 3, f"Input tensor not of shape [N_MC, BatchSize, Features] but is {x.shape=}"
        mc = x.shape[0]
        specs, params = [], []
        for l in self.layers:
            eps_w, eps_b, spec = l._draw(mc, x.device, True)
            if eps_w is not None:
                raise RuntimeError("FusedBayesMLP runs the on-chip Philox sampler only; use the unfused "
                                   "layers for injected eps")
            if spec.precision not in _cabi.BF16_CLASS:
                raise RuntimeError("FusedBayesMLP chains bf16 activations: set_precision('auto' | 'bf16' | '*_fast') "
                                   "or run the layers unfused")
            sh = getattr(l, "_shard", None)
            if sh is not None:                          # set by train.ELBOTrainStep (flat gradients / peer-sharded step)
                spec.w_lp, spec.w_lp_half, spec.dw_rho_raw = sh.get("w_lp", 0), sh.get("w_lp_half", 0), sh.get("dw_rho_raw", 0)
                spec.grad_out = sh.get("grad_out")
                spec.after_bwd = sh.get("after_bwd")
                spec.dw_scatter = sh.get("dw_scatter")
            specs.append(spec)
            l._remember(mc, None, None, spec)
            params += [l.weight.loc, l.weight.logscale, l.bias.loc, l.bias.logscale]
        out = _FusedMLPFn.apply(x, self.negative_slope, specs, *params)
        for l in self.layers:                           # kl_div / entropy exactly as the layer would
            if l.kl_skip:
                l.kl_div = l.entropy = torch.zeros((), device=x.device)
            elif l.kl_autograd:
                l.kl_div, l.entropy = PF.kl_and_entropy(l.weight.loc, l.weight.logscale, l.prior.scale)
            else:
                with torch.no_grad():
                    l.kl_div, l.entropy = PF.kl_and_entropy(l.weight.loc.detach(),
                                                            l.weight.logscale.detach(), l.prior.scale)
        return out


def fuse_bayes_mlp(net: torch.nn.Sequential, batch: int) -> torch.nn.Sequential:
    """Returns a Sequential in which every maximal run BayesLinear (LeakyReLU BayesLinear)+ whose
    layers the tensor-core path tiles is replaced by one FusedBayesMLP over the SAME layer objects."""
    mods = list(net)
    out, i = [], 0
    while i < len(mods):
        if isinstance(mods[i], BayesLinear):
            run, slope, j = [mods[i]], None, i + 1
            # grow the run only over layers the tensor-core path tiles: a narrow head (C2's 1200 -> 10) must not
            # keep the wide layers in front of it from chaining
            while (FusedBayesMLP.supported(run, batch)
                   and j + 1 < len(mods) and isinstance(mods[j], torch.nn.LeakyReLU)
                   and isinstance(mods[j + 1], BayesLinear)
                   and FusedBayesMLP.supported([mods[j + 1]], batch)
                   and (slope is None or slope == mods[j].negative_slope)):
                slope = mods[j].negative_slope
                run.append(mods[j + 1])
                j += 2
            if len(run) > 1 and FusedBayesMLP.supported(run, batch):
                out.append(FusedBayesMLP(run, slope))
                i = j
                continue
        out.append(mods[i])
        i += 1
    return torch.nn.Sequential(*out)
