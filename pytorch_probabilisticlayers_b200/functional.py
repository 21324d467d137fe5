"""torch.autograd.Functions that hand raw device pointers to the C-ABI kernels.

These are the only places where the Python host layer touches arithmetic; everything numeric
runs in libpl_b200.so.  Reference call sites replaced: src/ProbabilisticLayers.py:925-935
(rsample + baddbmm/bmm), :955-956 and :1059 (KL / entropy), :1032-1057 (grouped conv),
:321-333 (SIVI contraction + KL).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import torch

from . import _cabi


@dataclass
class SampleSpec:
    """How eps is obtained for one layer call."""
    seed: int = 0
    layer_id: int = 0
    mc_global_offset: int = 0          # first global MC index owned by this rank
    precision: int = _cabi.PREC_FP32
    seed_dev: int = 0                  # device address of a uint64 added to ``seed`` by the kernels (CUDA graphs)
    # MC-sharded training over peer memory (train.ELBOTrainStep, shard.PeerShard); used by the fused chain only
    w_lp: int = 0                      # device address of the weight in operand form (0: pack per call)
    w_lp_half: int = 0
    dw_rho_raw: int = 0                # drho leaves the dW kernel as sum dW*eps (the shard owner applies sigmoid(rho))
    grad_out: Optional[tuple] = None   # (dw_mu, dw_rho, db_mu, db_rho) tensors the backward writes straight into
    after_bwd: Optional[object] = None  # callable run right after this layer's backward kernels are enqueued
    dw_scatter: Optional[dict] = None   # reduce-scatter fused into the dW epilogue: peers (ctypes array), mu_off, rho_off,
                                        # chunk, rank, world, bf16 (pl_linear_args.dw_scatter_*)


# ---- optional per-call CUDA-event timing (bench.py's roofline leg) -----------------------------
_PROFILE = None
_WORK = {}     # tag -> ("flops" | "bytes", algorithmic amount per call), filled while profiling


def profile_start():
    """Record a CUDA-event pair around every C-ABI call until profile_stop()."""
    global _PROFILE
    _PROFILE = {}


def profile_stop():
    """-> {tag: (calls, total_ms)}; synchronises."""
    global _PROFILE
    prof, _PROFILE = _PROFILE, None
    torch.cuda.synchronize()
    return {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in (prof or {}).items()}


_NVTX = __import__("os").environ.get("PL_NVTX", "0") == "1"    # PL_NVTX=1: one NVTX range per C-ABI call (ncu --nvtx)


def _timed(tag, fn, work=None):
    if _NVTX:
        torch.cuda.nvtx.range_push(tag)
        try:
            return _timed_inner(tag, fn, work)
        finally:
            torch.cuda.nvtx.range_pop()
    return _timed_inner(tag, fn, work)


def _timed_inner(tag, fn, work=None):
    if _PROFILE is None:
        return fn()
    if work is not None:
        _WORK[tag] = work
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    r = fn()
    b.record()
    _PROFILE.setdefault(tag, []).append((a, b))
    return r


_KL_WS = {}


def _kl_workspace(device):
    """Zero-initialised, self-resetting ticket workspace of pl_kl_fwd, one per (device, stream)."""
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream)
    ws = _KL_WS.get(key)
    if ws is None:
        ws = torch.zeros(int(_cabi.lib().pl_kl_workspace_bytes()), dtype=torch.uint8, device=device)
        _KL_WS[key] = ws
    return ws


_CE_WS = {}


def _ce_workspace(device):
    """Zero-initialised, self-resetting ticket workspace of pl_mc_cross_entropy, one per (device, stream)."""
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream)
    ws = _CE_WS.get(key)
    if ws is None:
        ws = torch.zeros(int(_cabi.lib().pl_mc_cross_entropy_workspace_bytes()), dtype=torch.uint8, device=device)
        _CE_WS[key] = ws
    return ws


def mc_cross_entropy_raw(pred, target, grad_scale=None):
    """One pass over the [MC,B,C] logits: returns (out, dpred) with out = device tensor
    [sum of -log p[target], number of argmax hits] and dpred = (softmax - onehot) * grad_scale
    (None when grad_scale is None).  Reference: src/ProbabilisticLayers.py:32-43, :46-90."""
    _cabi.require_cuda_f32(pred, "pred")
    if pred.dim() != 3 or target.dim() != 1 or target.shape[0] != pred.shape[1]:
        raise ValueError(f"pred must be [MC,B,C] and target [B], got {tuple(pred.shape)} and {tuple(target.shape)}")
    lib = _cabi.lib()
    pc = pred.contiguous()
    tc = target.to(device=pred.device, dtype=torch.int64).contiguous()
    mc, batch, classes = pc.shape
    out = torch.empty(2, dtype=torch.float32, device=pred.device)
    dpred = torch.empty_like(pc) if grad_scale is not None else None
    ws = _ce_workspace(pred.device)
    a = _cabi.CeArgs()
    a.mc, a.batch, a.classes = mc, batch, classes
    a.grad_scale = float(grad_scale) if grad_scale is not None else 0.0
    a.pred, a.target, a.dpred, a.out = _cabi.ptr(pc), _cabi.ptr(tc), _cabi.ptr(dpred), _cabi.ptr(out)
    a.workspace, a.workspace_bytes = _cabi.ptr(ws), ws.numel()
    a.stream = _cabi.current_stream(pred.device)
    with torch.cuda.device(pred.device):
        _timed(f"mc_cross_entropy[{mc}x{batch}x{classes}]",
               lambda: _cabi.check(lib.pl_mc_cross_entropy(C.byref(a)), "pl_mc_cross_entropy"),
               ("bytes", 4.0 * pc.numel() * (2 if dpred is not None else 1)))
    return out, dpred


class _MCCrossEntropyFn(torch.autograd.Function):
    """loss = num_samples * mean over (MC,B) of CE; the gradient is produced by the same kernel pass."""

    @staticmethod
    def forward(ctx, pred, target, num_samples):
        rows = pred.shape[0] * pred.shape[1]
        scale = float(num_samples) / max(rows, 1)
        out, dpred = mc_cross_entropy_raw(pred, target, scale if ctx.needs_input_grad[0] else None)
        ctx.save_for_backward(dpred)
        ctx.mark_non_differentiable(out)
        return out[0] * scale, out

    @staticmethod
    def backward(ctx, g_loss, _g_out):
        (dpred,) = ctx.saved_tensors
        return dpred * g_loss, None, None


def mc_cross_entropy(pred, target, num_samples):
    """Returns (loss, out) with out[1] = number of correct argmax predictions (for MC_Accuracy without a
    second pass over the logits)."""
    return _MCCrossEntropyFn.apply(pred, target, num_samples)


def _workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def _mc_view(x, feature_dims):
    """Returns (tensor to pass, mc_stride in elements, MC).  An MC-expanded view (stride 0 on
    dim 0, what our MC_ExpansionLayer produces) is passed as a single sample with stride 0."""
    mc = x.shape[0]
    if x.stride(0) == 0 and mc > 1 and x[0].numel() > 0:
        base = x[0].contiguous()
        return base, 0, mc
    xc = x.contiguous()
    per = 1
    for d in xc.shape[1:]:
        per *= d
    return xc, per, mc


class _BayesLinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w_mu, w_rho, b_mu, b_rho, eps_w, eps_b, spec: SampleSpec, param_mode: int):
        for t, n in ((x, "x"), (w_mu, "weight.loc"), (w_rho, "weight.logscale")):
            _cabi.require_cuda_f32(t, n)
        lib = _cabi.lib()
        xb, x_stride, mc = _mc_view(x, 2)
        batch, fin = x.shape[1], x.shape[2]
        fout = w_mu.shape[-1]
        w_mu_c, w_rho_c = w_mu.contiguous(), w_rho.contiguous()
        b_mu_c = b_mu.contiguous() if b_mu is not None else None
        b_rho_c = b_rho.contiguous() if b_rho is not None else None
        eps_w_c = eps_w.contiguous() if eps_w is not None else None
        eps_b_c = eps_b.contiguous() if eps_b is not None else None
        y = torch.empty((mc, batch, fout), dtype=torch.float32, device=x.device)
        a = _cabi.LinearArgs()
        a.mc, a.batch, a.in_features, a.out_features = mc, batch, fin, fout
        a.precision, a.param_mode = spec.precision, param_mode
        a.x, a.x_mc_stride = _cabi.ptr(xb), x_stride
        a.w_mu, a.w_rho = _cabi.ptr(w_mu_c), _cabi.ptr(w_rho_c)
        a.b_mu, a.b_rho = _cabi.ptr(b_mu_c), _cabi.ptr(b_rho_c)
        a.seed, a.layer_id, a.mc_global_offset = spec.seed, spec.layer_id, spec.mc_global_offset
        a.seed_dev = spec.seed_dev or None
        a.eps_w, a.eps_b = _cabi.ptr(eps_w_c), _cabi.ptr(eps_b_c)
        a.y = _cabi.ptr(y)
        # tensor-core path: sigma and the bf16 copy of x are kept for the backward instead of fp32 x
        state = None
        nstate = lib.pl_linear_state_bytes(C.byref(a))
        if nstate and any(ctx.needs_input_grad[:5]):
            state = torch.empty(int(nstate), dtype=torch.uint8, device=x.device)
            a.fwd_state, a.fwd_state_bytes = _cabi.ptr(state), state.numel()
        ws = _workspace(lib.pl_linear_workspace_bytes(C.byref(a)), x.device)
        a.workspace, a.workspace_bytes = _cabi.ptr(ws), ws.numel()
        a.stream = _cabi.current_stream(x.device)
        with torch.cuda.device(x.device):
            _timed(f"linear_fwd[{mc}x{batch}x{fin}x{fout}]",
                   lambda: _cabi.check(lib.pl_linear_fwd(C.byref(a)), "pl_linear_fwd"),
                   ("flops", 2.0 * mc * batch * fin * fout))
        keep_x = state is None or spec.precision not in _cabi.BF16_CLASS   # bf16: the state holds x as bf16
        ctx.save_for_backward(xb if keep_x else None, w_mu_c, w_rho_c, b_mu_c, b_rho_c,
                              eps_w_c, eps_b_c, state)
        ctx.meta = (spec, param_mode, x_stride, mc, batch, fin, fout, tuple(x.shape))
        return y

    @staticmethod
    def backward(ctx, dy):
        xb, w_mu, w_rho, b_mu, b_rho, eps_w, eps_b, state = ctx.saved_tensors
        spec, param_mode, x_stride, mc, batch, fin, fout, x_shape = ctx.meta
        lib = _cabi.lib()
        dy = dy.contiguous()
        dev = dy.device
        need_dx = ctx.needs_input_grad[0]
        need_w = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        need_b = b_mu is not None and (ctx.needs_input_grad[3] or ctx.needs_input_grad[4])
        per_mc = param_mode == _cabi.PARAM_PER_MC_SIGMA
        dx = torch.empty((mc, batch, fin), dtype=torch.float32, device=dev) if need_dx else None
        wshape = (mc, fin, fout) if per_mc else (fin, fout)
        bshape = (mc, fout) if per_mc else (fout,)
        dw_mu = torch.empty(wshape, dtype=torch.float32, device=dev) if need_w else None
        dw_rho = torch.empty(wshape, dtype=torch.float32, device=dev) if need_w else None
        db_mu = torch.empty(bshape, dtype=torch.float32, device=dev) if need_b else None
        db_rho = torch.empty(bshape, dtype=torch.float32, device=dev) if need_b else None
        a = _cabi.LinearArgs()
        a.mc, a.batch, a.in_features, a.out_features = mc, batch, fin, fout
        a.precision, a.param_mode = spec.precision, param_mode
        a.x, a.x_mc_stride = _cabi.ptr(xb), x_stride
        a.w_mu, a.w_rho = _cabi.ptr(w_mu), _cabi.ptr(w_rho)
        a.b_mu, a.b_rho = _cabi.ptr(b_mu), _cabi.ptr(b_rho)
        a.seed, a.layer_id, a.mc_global_offset = spec.seed, spec.layer_id, spec.mc_global_offset
        a.seed_dev = spec.seed_dev or None
        a.eps_w, a.eps_b = _cabi.ptr(eps_w), _cabi.ptr(eps_b)
        a.dy, a.dx = _cabi.ptr(dy), _cabi.ptr(dx)
        a.dw_mu, a.dw_rho = _cabi.ptr(dw_mu), _cabi.ptr(dw_rho)
        a.db_mu, a.db_rho = _cabi.ptr(db_mu), _cabi.ptr(db_rho)
        if state is not None:
            a.fwd_state, a.fwd_state_bytes = _cabi.ptr(state), state.numel()
        ws = _workspace(lib.pl_linear_workspace_bytes(C.byref(a)), dev)
        a.workspace, a.workspace_bytes = _cabi.ptr(ws), ws.numel()
        a.stream = _cabi.current_stream(dev)
        shape = f"[{mc}x{batch}x{fin}x{fout}]"
        with torch.cuda.device(dev):
            if _PROFILE is None:
                _cabi.check(lib.pl_linear_bwd(C.byref(a)), "pl_linear_bwd")
            else:
                # same kernels, one C-ABI call per output group so each can be timed on its own
                outs = (a.dx, a.dw_mu, a.dw_rho, a.db_mu, a.db_rho)
                for tag, keep in (("dx", (0,)), ("dw", (1, 2)), ("db", (3, 4))):
                    sel = [o if i in keep else None for i, o in enumerate(outs)]
                    if all(v is None for v in sel):
                        continue
                    a.dx, a.dw_mu, a.dw_rho, a.db_mu, a.db_rho = sel
                    _timed(f"linear_bwd_{tag}{shape}",
                           lambda: _cabi.check(lib.pl_linear_bwd(C.byref(a)), "pl_linear_bwd"),
                           ("flops", 2.0 * mc * batch * fin * fout) if tag != "db"
                           else ("bytes", 4.0 * mc * batch * fout))
                a.dx, a.dw_mu, a.dw_rho, a.db_mu, a.db_rho = outs
        return dx, dw_mu, dw_rho, db_mu, db_rho, None, None, None, None


def bayes_linear(x, w_mu, w_rho, b_mu=None, b_rho=None, eps_w=None, eps_b=None,
                 spec: Optional[SampleSpec] = None, param_mode: int = _cabi.PARAM_SHARED_RHO):
    """Y[mc] = X[mc] @ (mu + sigma*eps[mc]) + (b_mu + b_sigma*eps_b[mc])."""
    return _BayesLinearFn.apply(x, w_mu, w_rho, b_mu, b_rho, eps_w, eps_b, spec or SampleSpec(),
                                param_mode)


class _BayesConv2dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w_mu, w_rho, b_mu, b_rho, eps_w, eps_b, spec: SampleSpec, hyper):
        for t, n in ((x, "x"), (w_mu, "weight.loc"), (w_rho, "weight.logscale")):
            _cabi.require_cuda_f32(t, n)
        lib = _cabi.lib()
        k, stride, padding, dilation = hyper
        xb, x_stride, mc = _mc_view(x, 4)
        _, batch, cin, h, w = x.shape
        cout = w_mu.shape[0]
        ho = (h + 2 * padding - dilation * (k - 1) - 1) // stride + 1
        wo = (w + 2 * padding - dilation * (k - 1) - 1) // stride + 1
        tens = [t.contiguous() if t is not None else None
                for t in (w_mu, w_rho, b_mu, b_rho, eps_w, eps_b)]
        y = torch.empty((mc, batch, cout, ho, wo), dtype=torch.float32, device=x.device)
        a = _cabi.Conv2dArgs()
        a.mc, a.batch, a.cin, a.cout, a.h, a.w = mc, batch, cin, cout, h, w
        a.k, a.stride, a.padding, a.dilation = k, stride, padding, dilation
        a.precision = spec.precision
        a.x, a.x_mc_stride = _cabi.ptr(xb), x_stride
        a.w_mu, a.w_rho, a.b_mu, a.b_rho = (_cabi.ptr(t) for t in tens[:4])
        a.seed, a.layer_id, a.mc_global_offset = spec.seed, spec.layer_id, spec.mc_global_offset
        a.seed_dev = spec.seed_dev or None
        a.eps_w, a.eps_b = _cabi.ptr(tens[4]), _cabi.ptr(tens[5])
        a.y = _cabi.ptr(y)
        state = None
        nstate = lib.pl_conv2d_state_bytes(C.byref(a))
        if nstate and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]):
            state = torch.empty(int(nstate), dtype=torch.uint8, device=x.device)   # bf16 im2col matrix
            a.fwd_state, a.fwd_state_bytes = _cabi.ptr(state), state.numel()
        ws = _workspace(lib.pl_conv2d_workspace_bytes(C.byref(a)), x.device)
        a.workspace, a.workspace_bytes = _cabi.ptr(ws), ws.numel()
        a.stream = _cabi.current_stream(x.device)
        with torch.cuda.device(x.device):
            _timed(f"conv_fwd[{mc}x{batch}x{cin}x{h}x{w}->{cout},k{k}]",
                   lambda: _cabi.check(lib.pl_conv2d_fwd(C.byref(a)), "pl_conv2d_fwd"),
                   ("flops", 2.0 * mc * batch * ho * wo * cout * cin * k * k))
        ctx.save_for_backward(xb, *tens, state)
        ctx.meta = (spec, hyper, x_stride, (mc, batch, cin, cout, h, w), a.precision)
        return y

    @staticmethod
    def backward(ctx, dy):
        xb, w_mu, w_rho, b_mu, b_rho, eps_w, eps_b, state = ctx.saved_tensors
        spec, hyper, x_stride, dims, precision = ctx.meta
        mc, batch, cin, cout, h, w = dims
        k, stride, padding, dilation = hyper
        lib = _cabi.lib()
        dy = dy.contiguous()
        dev = dy.device
        need_dx = ctx.needs_input_grad[0]
        need_w = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        need_b = ctx.needs_input_grad[3] or ctx.needs_input_grad[4]
        dx = torch.empty((mc, batch, cin, h, w), dtype=torch.float32, device=dev) if need_dx else None
        dw_mu = torch.empty_like(w_mu) if need_w else None
        dw_rho = torch.empty_like(w_mu) if need_w else None
        db_mu = torch.empty_like(b_mu) if need_b else None
        db_rho = torch.empty_like(b_mu) if need_b else None
        a = _cabi.Conv2dArgs()
        a.mc, a.batch, a.cin, a.cout, a.h, a.w = mc, batch, cin, cout, h, w
        a.k, a.stride, a.padding, a.dilation = k, stride, padding, dilation
        a.precision = precision
        a.x, a.x_mc_stride = _cabi.ptr(xb), x_stride
        a.w_mu, a.w_rho, a.b_mu, a.b_rho = (_cabi.ptr(t) for t in (w_mu, w_rho, b_mu, b_rho))
        a.seed, a.layer_id, a.mc_global_offset = spec.seed, spec.layer_id, spec.mc_global_offset
        a.seed_dev = spec.seed_dev or None
        a.eps_w, a.eps_b = _cabi.ptr(eps_w), _cabi.ptr(eps_b)
        a.dy, a.dx = _cabi.ptr(dy), _cabi.ptr(dx)
        a.dw_mu, a.dw_rho, a.db_mu, a.db_rho = (_cabi.ptr(t) for t in (dw_mu, dw_rho, db_mu, db_rho))
        if state is not None:
            a.fwd_state, a.fwd_state_bytes = _cabi.ptr(state), state.numel()
        ws = _workspace(lib.pl_conv2d_workspace_bytes(C.byref(a)), dev)
        a.workspace, a.workspace_bytes = _cabi.ptr(ws), ws.numel()
        a.stream = _cabi.current_stream(dev)
        with torch.cuda.device(dev):
            ho = (h + 2 * padding - dilation * (k - 1) - 1) // stride + 1
            wo = (w + 2 * padding - dilation * (k - 1) - 1) // stride + 1
            _timed(f"conv_bwd[{mc}x{batch}x{cin}x{h}x{w}->{cout},k{k}]",
                   lambda: _cabi.check(lib.pl_conv2d_bwd(C.byref(a)), "pl_conv2d_bwd"),
                   ("flops", 2.0 * mc * batch * ho * wo * cout * cin * k * k * (2 if need_dx else 1)))
        return dx, dw_mu, dw_rho, db_mu, db_rho, None, None, None, None


def bayes_conv2d(x, w_mu, w_rho, b_mu, b_rho, kernel_size, stride=1, padding=0, dilation=1,
                 eps_w=None, eps_b=None, spec: Optional[SampleSpec] = None):
    return _BayesConv2dFn.apply(x, w_mu, w_rho, b_mu, b_rho, eps_w, eps_b, spec or SampleSpec(),
                                (int(kernel_size), int(stride), int(padding), int(dilation)))


class _KLFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, rho, prior_scale, prior_scale_elem):
        _cabi.require_cuda_f32(mu, "weight.loc")
        _cabi.require_cuda_f32(rho, "weight.logscale")
        lib = _cabi.lib()
        mu_c, rho_c = mu.contiguous(), rho.contiguous()
        pse = prior_scale_elem.contiguous() if prior_scale_elem is not None else None
        out = torch.empty(2, dtype=torch.float32, device=mu.device)
        ws = _kl_workspace(mu.device)
        with torch.cuda.device(mu.device):
            _timed(f"kl_fwd[{mu_c.numel()}]", lambda: _cabi.check(
                lib.pl_kl_fwd(_cabi.ptr(mu_c), _cabi.ptr(rho_c), mu_c.numel(), float(prior_scale),
                              _cabi.ptr(pse), _cabi.ptr(out), _cabi.ptr(ws), ws.numel(),
                              _cabi.current_stream(mu.device)), "pl_kl_fwd"),
                   ("bytes", 8.0 * mu_c.numel()))
        ctx.save_for_backward(mu_c, rho_c, pse)
        ctx.prior_scale = float(prior_scale)
        return out[0], out[1]

    @staticmethod
    def backward(ctx, g_kl, g_ent):
        mu, rho, pse = ctx.saved_tensors
        lib = _cabi.lib()
        dmu = torch.empty_like(mu)
        drho = torch.empty_like(rho)
        g_kl = g_kl.contiguous().float()
        g_ent = g_ent.contiguous().float()
        with torch.cuda.device(mu.device):
            _cabi.check(lib.pl_kl_bwd(_cabi.ptr(mu), _cabi.ptr(rho), mu.numel(), ctx.prior_scale,
                                      _cabi.ptr(pse), _cabi.ptr(g_kl), _cabi.ptr(g_ent), 1.0,
                                      _cabi.ptr(dmu), _cabi.ptr(drho), 0,
                                      _cabi.current_stream(mu.device)), "pl_kl_bwd")
        return dmu, drho, None, None


def kl_and_entropy(mu, rho, prior_scale=1.0, prior_scale_elem=None):
    """(kl_div, entropy) of N(mu, softplus(rho)) against N(0, prior_scale), summed over elements."""
    return _KLFn.apply(mu, rho, prior_scale, prior_scale_elem)


def eps_dump(seed, stream_id, mc_global_offset, num_mc, rows, cols, device, raw=False, seed_dev=None):
    """The on-chip eps stream, materialised (test bridge for src/ProbabilisticLayers.py:835).  ``seed_dev``: device
    address of a uint64 added to the seed on the device (CUDA-graph replays)."""
    lib = _cabi.lib()
    if raw:
        out = torch.empty((num_mc, rows, 4 * ((cols + 7) // 8)), dtype=torch.int32, device=device)
    else:
        out = torch.empty((num_mc, rows, cols), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        _cabi.check(lib.pl_eps_dump_dev(int(seed), seed_dev or None, int(stream_id), int(mc_global_offset), num_mc, rows,
                                        cols, _cabi.ptr(out), 1 if raw else 0,
                                        _cabi.current_stream(out.device)), "pl_eps_dump")
    return out
