"""MC_BatchNorm2D -> LeakyReLU -> MC_MaxPool2d(2, 2) as one fused module (csrc/bn_pool.cu).

The reference's conv blocks (experiments/BNN.py:232-247) run batch norm, activation and pooling as three
modules with several permute / contiguous copies each (src/ProbabilisticLayers.py:736-762, :207-236); on a
B200 their ATen / cuDNN kernels take about half of the C3 step.  ``fuse_bn_act_pool(net)`` replaces each such
triple by a FusedBNActPool over the SAME MC_BatchNorm2D object (parameters, running statistics and
``state_dict`` keys of the batch norm are unchanged), opt-in like ``fused.fuse_bayes_mlp``.
"""
import ctypes as C

import torch

from . import _cabi
from . import functional as PF
from .layers import MC_BatchNorm2D, MC_MaxPool2d


class _BnActPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, mean, rstd, slope, batch_stats):
        lib = _cabi.lib()
        mc, b, c, h, w = x.shape
        xc = x.contiguous()
        y = torch.empty((mc, b, c, h // 2, w // 2), dtype=torch.float32, device=x.device)
        idx = torch.empty((mc, b, c, h // 2, w // 2), dtype=torch.uint8, device=x.device)
        a = _cabi.BnPoolArgs()
        a.n, a.channels, a.h, a.w = mc * b, c, h, w
        a.slope, a.batch_stats = float(slope), int(batch_stats)
        a.x, a.gamma, a.beta = _cabi.ptr(xc), _cabi.ptr(gamma), _cabi.ptr(beta)
        a.mean, a.rstd = _cabi.ptr(mean), _cabi.ptr(rstd)
        a.y, a.idx = _cabi.ptr(y), _cabi.ptr(idx)
        a.stream = _cabi.current_stream(x.device)
        with torch.cuda.device(x.device):
            PF._timed(f"bn_act_pool_fwd[{mc * b}x{c}x{h}x{w}]",
                      lambda: _cabi.check(lib.pl_bn_act_pool_fwd(C.byref(a)), "pl_bn_act_pool_fwd"),
                      ("bytes", 4.0 * xc.numel() * (1 + 0.25) + idx.numel()))
        ctx.save_for_backward(xc, gamma, beta, mean, rstd, idx)
        ctx.meta = (float(slope), int(batch_stats))
        ctx.mark_non_differentiable(idx)
        return y

    @staticmethod
    def backward(ctx, dy):
        xc, gamma, beta, mean, rstd, idx = ctx.saved_tensors
        slope, batch_stats = ctx.meta
        lib = _cabi.lib()
        mc, b, c, h, w = xc.shape
        dyc = dy.contiguous()
        dx = torch.empty_like(xc) if ctx.needs_input_grad[0] else None
        dgamma = torch.empty_like(gamma) if ctx.needs_input_grad[1] else None
        dbeta = torch.empty_like(beta) if ctx.needs_input_grad[2] else None
        ws = PF._workspace(lib.pl_bn_act_pool_workspace_bytes(c), xc.device)
        a = _cabi.BnPoolArgs()
        a.n, a.channels, a.h, a.w = mc * b, c, h, w
        a.slope, a.batch_stats = slope, batch_stats
        a.x, a.gamma, a.beta = _cabi.ptr(xc), _cabi.ptr(gamma), _cabi.ptr(beta)
        a.mean, a.rstd, a.idx = _cabi.ptr(mean), _cabi.ptr(rstd), _cabi.ptr(idx)
        a.dy, a.dx, a.dgamma, a.dbeta = _cabi.ptr(dyc), _cabi.ptr(dx), _cabi.ptr(dgamma), _cabi.ptr(dbeta)
        a.workspace, a.workspace_bytes = _cabi.ptr(ws), ws.numel()
        a.stream = _cabi.current_stream(xc.device)
        with torch.cuda.device(xc.device):
            PF._timed(f"bn_act_pool_bwd[{mc * b}x{c}x{h}x{w}]",
                      lambda: _cabi.check(lib.pl_bn_act_pool_bwd(C.byref(a)), "pl_bn_act_pool_bwd"),
                      ("bytes", 4.0 * xc.numel() * 3 + 4.0 * dyc.numel() * 2))
        return dx, dgamma, dbeta, None, None, None, None


class FusedBNActPool(torch.nn.Module):
    """``pool(leaky_relu(bn(x)))`` for x [MC, B, C, H, W] with H, W even, in one pass each way."""

    def __init__(self, bn: MC_BatchNorm2D, negative_slope: float = 0.01, pool: MC_MaxPool2d = None):
        super().__init__()
        self.bn = bn
        self.negative_slope = float(negative_slope)
        self.pool = pool if pool is not None else MC_MaxPool2d(2, 2)

    def forward(self, x):
        assert x.dim() == 5, 'Input tensor does not have dimensionality of 5 [num_MC, Batch_Size, Features]'
        bn = self.bn
        assert bn.num_features == x.shape[2]
        mc, b, c, h, w = x.shape
        if not (x.is_cuda and x.dtype == torch.float32 and h % 2 == 0 and w % 2 == 0 and h >= 2 and w >= 2):
            # odd planes / CPU: the three modules as they are (they are torch ops in either case)
            return self.pool(torch.nn.functional.leaky_relu(bn(x), self.negative_slope))
        lib = _cabi.lib()
        xc = x.contiguous()
        use_batch = bn.training or not bn.track_running_stats
        factor = bn._factor()
        if use_batch:
            sums = torch.empty(2 * c, dtype=torch.float64, device=x.device)
            with torch.cuda.device(x.device):
                PF._timed(f"bn_stats[{mc * b}x{c}x{h}x{w}]", lambda: _cabi.check(
                    lib.pl_bn_stats(_cabi.ptr(xc), mc * b, c, h * w, _cabi.ptr(sums),
                                    _cabi.current_stream(x.device)), "pl_bn_stats"), ("bytes", 4.0 * xc.numel()))
            cnt = mc * b * h * w
            with torch.no_grad():
                mean64 = sums[0::2] / cnt
                var64 = (sums[1::2] / cnt - mean64 * mean64).clamp_(min=0.0)
                mean = mean64.float()
                rstd = torch.rsqrt(var64 + bn.eps).float()
                if bn.training and bn.track_running_stats and bn.running_mean is not None:
                    bn.running_mean.mul_(1.0 - factor).add_(mean, alpha=factor)
                    unbiased = var64 * (cnt / max(cnt - 1, 1))
                    bn.running_var.mul_(1.0 - factor).add_(unbiased.float(), alpha=factor)
        else:
            mean = bn.running_mean
            rstd = torch.rsqrt(bn.running_var + bn.eps)
        gamma = bn.weight if bn.weight is not None else torch.ones(c, device=x.device)
        beta = bn.bias if bn.bias is not None else torch.zeros(c, device=x.device)
        return _BnActPoolFn.apply(xc, gamma, beta, mean.contiguous(), rstd.contiguous(), self.negative_slope,
                                  use_batch)


def fuse_bn_act_pool(net: torch.nn.Sequential) -> torch.nn.Sequential:
    """Returns a Sequential in which every MC_BatchNorm2D, LeakyReLU, MC_MaxPool2d(2, 2) triple is one
    FusedBNActPool over the same batch-norm object."""
    mods = list(net)
    out, i = [], 0
    while i < len(mods):
        if (i + 2 < len(mods) and isinstance(mods[i], MC_BatchNorm2D) and isinstance(mods[i + 1], torch.nn.LeakyReLU)
                and isinstance(mods[i + 2], MC_MaxPool2d) and mods[i + 2].kernel_size == 2 and mods[i + 2].stride == 2):
            out.append(FusedBNActPool(mods[i], mods[i + 1].negative_slope, mods[i + 2]))
            i += 3
        else:
            out.append(mods[i])
            i += 1
    return torch.nn.Sequential(*out)
