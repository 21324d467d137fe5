"""MC-sharded ELBO train step (the harness that stands in for the reference's Lightning loop).

Semantics follow the reference's ``training_step`` (experiments/BNN.py:287-313,
experiments/RegressionBNN.py:109-120, experiments/SIVILastLayer.py:287-296):

    loss = LL(pred, y) / num_samples + sum_{BayesLinear, SIVILayer} kl_div / num_samples

with Adam (BNN.py:349).  Sharding: the Monte-Carlo samples are independent given
(mu, rho), so rank r of N owns global samples [r*MC/N, (r+1)*MC/N), holds a full parameter replica
and sees the full batch.  The cross-entropy ELBO decomposes over samples, so the only exchange is
an all-reduce(sum) of one flat fp32 buffer [all gradients | LL, KL, ACC scalars] per step; KL and
its gradient are replica-identical and are pre-scaled by 1/N instead of being reduced separately.
The KL gradient is streamed straight into the flat gradient buffer by ``pl_kl_bwd(accumulate=1)``
(no autograd temporaries).  With ``overlap=True`` (default) the flat buffer is reduced in per-layer
slices: a layer's slice is handed to NCCL (async, on NCCL's own stream) from a
post-accumulate-grad hook as soon as that layer's backward kernels have been enqueued, so the
transfer of layer L runs under the backward of layer L-1 and only the first layer's slice is
exposed (SURVEY.md hard part H3).

``exchange='peer'`` (the default on CUDA when every rank can map every other rank's memory) replaces the
all-reduce + replicated optimiser pass by the NCCL-free sharded step of shard.py / csrc/comm.cu: reduce-scatter by
peer stores, Adam + KL on the rank's 1/N shard, all-gather of the bf16 operand the forward consumes, two device-side
barriers; the whole step is then a single stream of this library's kernels (CUDA-graph capturable at any N).
"""
from __future__ import annotations

import math
from typing import Iterable, Optional

import torch
import torch.distributed as dist

from . import _cabi, layers
from . import functional as PF
from .layers import BayesConv2d, BayesLinear, SIVILayer
from .functional import mc_cross_entropy_raw
from .loss import MC_Accuracy, MC_CrossEntropyLoss

_GOLDEN_SIGNED = layers._GOLDEN - (1 << 64)     # the per-call seed increment as a signed 64-bit value


def collect_kl_div(net, kinds=(BayesLinear, SIVILayer)):
    """Sum of ``kl_div`` over the layer kinds the reference's collect_kl_div matches
    (experiments/BNN.py:269-276: BayesLinear only -- conv KL never reaches the reference's loss)."""
    total = None
    for m in net.modules():
        if isinstance(m, tuple(kinds)) and hasattr(m, "kl_div"):
            total = m.kl_div if total is None else total + m.kl_div
    return total


class ELBOTrainStep:
    def __init__(self, net: torch.nn.Module, criterion, num_samples: int, lr: float = 1e-3,
                 kl_kinds: Iterable = (BayesLinear, SIVILayer), with_accuracy: bool = True,
                 group: Optional[dist.ProcessGroup] = None, mc_coupled_loss: bool = False,
                 overlap: bool = True, fused_optimizer: bool = True,
                 betas=(0.9, 0.999), eps: float = 1e-8, exchange: str = "auto", grad_wire: str = "fp32"):
        self.net, self.criterion, self.num_samples = net, criterion, int(num_samples)
        self.kl_kinds = tuple(kl_kinds)
        self.with_accuracy = with_accuracy
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if self.world > 1:
            # MC_NLL / MC_BatchNorm couple samples across MC, LaplacePrior's EMA hooks see the local pre-reduction
            # gradient (SURVEY.md 8e caveats): such nets only run unsharded ("replicas only")
            from .loss import MC_NLL
            coupled = [type(m).__name__ for m in net.modules()
                       if isinstance(m, (layers._MCBatchNorm, layers.LaplacePrior))]
            if mc_coupled_loss or isinstance(criterion, MC_NLL) or coupled:
                raise ValueError("MC sharding over %d ranks needs samples that are independent given (mu, rho); this "
                                 "configuration couples them (%s): run it unsharded (replicas only)"
                                 % (self.world, ", ".join(["criterion"] * bool(mc_coupled_loss or isinstance(criterion, MC_NLL))
                                                          + coupled)))
        self.params = [p for p in net.parameters() if p.requires_grad]
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n + 4, dtype=torch.float32, device=dev)   # grads | LL KL ACC pad
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.stats = self.flat[n:n + 4]
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        # fused optimiser pass (CUDA): Adam + analytic KL gradient + KL value in one kernel per
        # tensor (pl_adam_kl_step); replaces torch.optim.Adam and the KL forward/backward kernels
        self.fused_opt = fused_optimizer and dev.type == "cuda"
        self.opt = None if self.fused_opt else torch.optim.Adam(self.params, lr, betas=betas, eps=eps,
                                                                fused=dev.type == "cuda")
        # fused KL gradient: layers compute kl_div without an autograd graph
        self.kl_layers = [m for m in net.modules()
                          if isinstance(m, (BayesLinear, BayesConv2d)) and isinstance(m, self.kl_kinds)]
        for m in net.modules():
            # kl_div of layers outside kl_kinds (conv in the reference's experiments) never reaches
            # the loss: no graph needed for them either
            if isinstance(m, (BayesLinear, BayesConv2d)):
                m.kl_autograd = False
                if m not in self.kl_layers:
                    m.kl_skip = True
        if self.fused_opt:
            self.adam_m = torch.zeros(n, dtype=torch.float32, device=dev)
            self.adam_v = torch.zeros(n, dtype=torch.float32, device=dev)
            self.adam_t = 0
            self.kl_acc = torch.zeros(1, dtype=torch.float64, device=dev)
            kinds = {}
            self.kl_const = 0.0
            for m in self.kl_layers:
                gaussian = not (isinstance(m, BayesLinear) and isinstance(m.prior, layers.LaplacePrior))
                if not gaussian:
                    continue                      # per-element prior: stays on the autograd path
                prior = float(m.prior.scale if isinstance(m, BayesLinear) else m.prior_scale)
                kinds[id(m.weight.loc)] = (1, prior)
                kinds[id(m.weight.logscale)] = (2, prior)
                self.kl_const += m.weight.loc.numel() * (math.log(prior) - 0.5)
                m.kl_skip = True
            self._adam_plan = []
            off = 0
            for p in self.params:
                kind, prior = kinds.get(id(p), (0, 1.0))
                self._adam_plan.append((p, off, p.numel(), kind, prior))
                off += p.numel()
        # ---- gradient exchange for world > 1: 'peer' (shard.PeerShard) or 'nccl' (flat all-reduce) ----
        self.shard = None
        self.exchange = None if self.world == 1 else "nccl"
        if grad_wire not in ("fp32", "bf16"):
            raise ValueError("grad_wire must be 'fp32' or 'bf16'")
        # peer exchange only: format of the partial gradients on the wire / in the owners' staging.  'bf16' halves the
        # reduce-scatter bytes; every partial (a sum over MC/N samples) is rounded once to bf16 (rel 2^-9) and the owner sums
        # the N partials in fp32 -- inside the tolerance class of the bf16 compute modes, outside the 1e-5 N-invariance of
        # the default
        self.grad_wire = grad_wire
        if exchange not in ("auto", "peer", "nccl"):
            raise ValueError("exchange must be 'auto', 'peer' or 'nccl'")
        if self.world > 1 and exchange in ("auto", "peer") and self.fused_opt:
            try:
                self._setup_peer_exchange()
                self.exchange = "peer"
            except Exception as e:           # no peer mapping on this machine (or an unsupported net): NCCL path
                if exchange == "peer":
                    raise
                self.exchange_fallback_reason = f"{type(e).__name__}: {e}"[:300]
                self.shard = None
        if self.world == 1 and self.fused_opt:
            # single GPU: the same kernel as ONE multi-tensor launch over all parameter tensors (segment table) instead of
            # one launch per tensor; the gradient buffer is the shard arena's staging area, so nothing is copied
            try:
                self._setup_peer_exchange()
            except Exception as e:
                self.exchange_fallback_reason = f"{type(e).__name__}: {e}"[:300]
                self.shard = None
            self._direct_grads()
        # per-layer slices of the flat buffer for the overlapped all-reduce
        self.overlap = overlap and self.world > 1 and self.shard is None
        self._handles = []
        self._graph = None         # CUDA-graph mode (enable_cuda_graph)
        self._capturing = False
        self._slices = []          # (start, end, layer or None)
        self._inv = 1.0 / (self.num_samples * self.world)
        if self.overlap:
            start_of = {}
            off = 0
            for p in self.params:
                start_of[id(p)] = off
                off += p.numel()
            covered = []
            self._uncovered = []
            for m in self.kl_layers:
                ps = [p for p in m.parameters() if p.requires_grad]
                lo = min(start_of[id(p)] for p in ps)
                hi = max(start_of[id(p)] + p.numel() for p in ps)
                if hi - lo != sum(p.numel() for p in ps):
                    self._uncovered.append(m)     # not contiguous in the flat buffer: leave to the tail
                    continue
                state = {"left": len(ps), "n": len(ps), "lo": lo, "hi": hi, "layer": m}
                covered.append((lo, hi))
                for p in ps:
                    p.register_post_accumulate_grad_hook(self._make_hook(state))
            # everything not covered by a layer slice (+ the stats tail) goes in the final reduce
            covered.sort()
            cur, rest = 0, []
            for lo, hi in covered:
                if lo > cur:
                    rest.append((cur, lo))
                cur = hi
            rest.append((cur, n + 4))
            self._rest = rest

    # ---------------------------------------------------------------------------------------------------------
    def _fused_chains(self):
        from .fused import FusedBayesMLP
        return [m for m in self.net.modules() if isinstance(m, FusedBayesMLP)]

    def _direct_grads(self):
        """Fused chains write their parameter gradients straight into the flat buffer (no autograd accumulate pass)."""
        for chain in self._fused_chains():
            for l in chain.layers:
                ps = (l.weight.loc, l.weight.logscale, l.bias.loc, l.bias.logscale)
                if all(p.requires_grad and p.grad is not None for p in ps):
                    sh = getattr(l, "_shard", None) or {}
                    sh["grad_out"] = tuple(p.grad for p in ps)
                    l._shard = sh

    def _setup_peer_exchange(self, overlap_push=True, scatter_from_dw=True):
        """Builds the shard plan (shard.ShardPlan), moves parameters and gradients to its layout, maps the peers."""
        from . import shard as S
        if self.world > 1 and any(isinstance(m, SIVILayer) for m in self.net.modules()):
            raise RuntimeError("SIVILayer's KL needs autograd through its hyper-network: NCCL exchange")
        index = {id(p): i for i, p in enumerate(self.params)}
        chained = {}
        for chain in self._fused_chains():
            for l in chain.layers:
                chained[id(l)] = l._precision_code(True, True)
        segs, used = [], set()
        for m in self.kl_layers:
            gaussian = not (isinstance(m, BayesLinear) and isinstance(m.prior, layers.LaplacePrior))
            if not gaussian:
                if self.world > 1:
                    raise RuntimeError("LaplacePrior needs the autograd KL path: NCCL exchange")
                continue                          # per-element prior: plain Adam tensors, KL through autograd
            prior = float(m.prior.scale if isinstance(m, BayesLinear) else m.prior_scale)
            prec = chained.get(id(m))
            gather = {_cabi.PREC_BF16: 1, _cabi.PREC_BF16_FAST: 2}.get(prec, 0)
            if m.weight.loc.numel() % 8 or self.world == 1:
                gather = 0                        # single GPU: the layers pack their operand per call as before
            i1, i2 = index[id(m.weight.loc)], index[id(m.weight.logscale)]
            segs.append(S.SegSpec(kind=1, numel=m.weight.loc.numel(), p_index=i1, p2_index=i2, gather=gather,
                                  rho_raw=1 if gather else 0, prior_scale=prior))
            used.update((i1, i2))
        for i, p in enumerate(self.params):
            if i not in used:
                segs.append(S.SegSpec(kind=0, numel=p.numel(), p_index=i))
        if len(segs) > S.MAX_SEGS:
            raise RuntimeError(f"{len(segs)} optimiser segments > {S.MAX_SEGS}")
        plan = S.ShardPlan(self.world, [p.numel() for p in self.params], segs)
        dev = self.params[0].device
        self.shard = S.PeerShard(plan, self.params, self.group, wire_bf16=self.grad_wire == "bf16")
        self.shard_out = torch.zeros(4, dtype=torch.float32, device=dev)
        if self.world == 1:
            # the gradients live directly in the (single-row) staging area the optimiser kernel reads: no push
            lo = min(s.g_off for s in segs)
            self.flat = self.shard.arena[lo:plan.stats_off].view(torch.float32)
            for s in segs:
                self.params[s.p_index].grad = self.shard.staging_view(s).view_as(self.params[s.p_index])
                if s.kind == 1:
                    self.params[s.p2_index].grad = self.shard.staging_view(s, True).view_as(self.params[s.p2_index])
            self.stats = torch.zeros(4, dtype=torch.float32, device=dev)
            self.adam_m = self.adam_v = None          # the shard owns the optimiser state
            return
        # flat gradient buffer in the arena's parameter layout (+ the 4 statistics)
        n4 = plan.param_bytes // 4
        self.flat = torch.zeros(n4 + 4, dtype=torch.float32, device=dev)
        for p, off in zip(self.params, plan.param_offs):
            p.grad = self.flat[off // 4:off // 4 + p.numel()].view_as(p)
        self.stats = self.flat[n4:n4 + 4]
        by_index = {s.p_index: (k, s) for k, s in enumerate(segs) if s.kind == 1}
        # a chained layer's weight gradients leave for their owners on a side stream as soon as its backward kernels are
        # enqueued: the transfer of layer L runs under the backward of layer L-1 (the push blocks fit next to the one
        # tcgen05 CTA per SM); only the first layer's transfer and the small tensors stay on the critical path
        self._side = torch.cuda.Stream(device=dev)
        self._pushed = []                  # unit ranges already handed to the side stream in this step
        self._scattered = []               # unit ranges the dW kernels deliver themselves (never pushed)
        # per-layer optimiser pass + operand gather on the side stream as well: implemented and parity-tested, but measured
        # zero-sum at N = 8 (2.22 vs 2.18 ms/step): the gather competes with the backward kernels for the same HBM / NVLink
        # ingress, so it stays off
        self._overlap_adam = False
        self._adam_done = []
        for chain in self._fused_chains():
            for l in chain.layers:
                k, s = by_index.get(index[id(l.weight.loc)], (None, None))
                sh = {}
                if s is not None and s.gather:
                    ptr, half = self.shard.lp_pointers(s)
                    sh.update(w_lp=ptr, w_lp_half=half, dw_rho_raw=1)
                tiles = -(-l.dim_input // 128) * -(-l.dim_output // 128)
                import os
                if os.environ.get("PL_SHARD_NO_DW_SCATTER", "0") == "1":     # A/B switch: separate push kernel
                    scatter_from_dw = False
                if s is not None and s.gather and scatter_from_dw and tiles >= 148:
                    # reduce-scatter fused into the dW epilogue: the tile goes straight into the owners' staging rows, no
                    # push kernel (and no local gradient buffer traffic) for this weight
                    import ctypes as C
                    peers = (C.c_void_p * 8)(*[int(p_) for p_ in self.shard.peers])
                    sh["dw_scatter"] = dict(peers=peers, mu_off=s.g_off, rho_off=s.g2_off, chunk=s.chunk, rank=self.rank,
                                            world=self.world, bf16=1 if self.shard.wire_bf16 else 0)
                    self._scattered.append(self.shard.seg_units(k))
                    if os.environ.get("PL_SHARD_OVERLAP_ADAM", "0") == "1":     # experiment switch, see _make_adam_hook
                        sh["after_bwd"] = self._make_adam_hook(k)
                elif s is not None and overlap_push:
                    sh["after_bwd"] = self._make_push_hook(self.shard.seg_units(k), k)
                l._shard = sh
        self._direct_grads()
        self.shard.initial_gather()
        torch.cuda.synchronize(dev)
        if self.shard.status() != 0:
            raise RuntimeError("peer barrier timed out during set-up")

    def _make_push_hook(self, units, seg_index):
        """Side-stream work of one chained layer, issued right after its backward kernels: push its gradients to their
        owners, barrier (channel 1), then the owners' optimiser pass + operand gather of THIS layer -- all under the
        backward of the layers before it.  (Nobody reads this layer's operands again before the next forward, and every
        rank has finished its backward of the layer when it has pushed.)"""
        def hook():
            main = torch.cuda.current_stream(self.flat.device)
            ev = torch.cuda.Event()
            ev.record(main)
            self._side.wait_event(ev)
            with torch.cuda.stream(self._side):
                self.shard.push(self.flat, None, units)
                if self._overlap_adam:
                    self.shard.barrier(1)
                    self.shard.adam(1.0 / self.num_samples, self.lr, self.betas, self.eps, self.adam_t, self._sched_cur,
                                    segs=(seg_index, seg_index + 1), zero_kl=False)
                    self._adam_done.append(seg_index)
            self._pushed.append(units)
        return hook

    def _make_adam_hook(self, seg_index):
        """Experiment (PL_SHARD_OVERLAP_ADAM=1): a scattered layer's barrier + optimiser pass + operand gather on the side
        stream, under the backward of the layers before it."""
        def hook():
            main = torch.cuda.current_stream(self.flat.device)
            ev = torch.cuda.Event()
            ev.record(main)
            self._side.wait_event(ev)
            with torch.cuda.stream(self._side):
                self.shard.barrier(1)
                self.shard.adam(1.0 / self.num_samples, self.lr, self.betas, self.eps, self.adam_t, self._sched_cur,
                                segs=(seg_index, seg_index + 1), zero_kl=False)
            self._adam_done.append(seg_index)
            self._side_used = True
        return hook

    def _begin_opt_step(self):
        """Start-of-step state of the shard optimiser (the per-layer passes may run before the backward has finished)."""
        self.adam_t += 1
        self.shard.kl_acc.zero_()
        self._sched_cur = None
        self._adam_done = []
        if self._capturing:
            self._t_dev.add_(1.0)
            bc1 = 1.0 - torch.pow(torch.full_like(self._t_dev, self.betas[0]), self._t_dev)
            bc2 = 1.0 - torch.pow(torch.full_like(self._t_dev, self.betas[1]), self._t_dev)
            self._sched.copy_(torch.cat([self.lr / bc1, torch.rsqrt(bc2)]).float())
            self._sched_cur = self._sched

    def sync_full_params(self):
        """Peer exchange keeps the fp32 values of operand-gathered weights current on their owners only; this brings
        them to every rank (checkpoints, evaluation, tests).  No-op otherwise."""
        if self.shard is not None:
            self.shard.sync_full_params(self.params)

    def _kl_backward(self, m):
        """Stream the analytic KL gradient of layer m (scaled 1/(N*num_samples)) into its .grad."""
        if self.fused_opt or m.kl_div.requires_grad:   # fused optimiser pass / autograd handles it
            return
        lib = _cabi.lib()
        w = m.weight
        prior = m.prior.scale if isinstance(m, BayesLinear) else m.prior_scale
        stream = _cabi.current_stream(w.loc.device)
        PF._timed(f"kl_bwd_acc[{w.loc.numel()}]", lambda: _cabi.check(
            lib.pl_kl_bwd(w.loc.data_ptr(), w.logscale.data_ptr(), w.loc.numel(), float(prior),
                          None, None, None, self._inv, w.loc.grad.data_ptr(),
                          w.logscale.grad.data_ptr(), 1, stream), "pl_kl_bwd"),
                  ("bytes", 24.0 * w.loc.numel()))

    def _make_hook(self, state):
        def hook(_param):
            state["left"] -= 1
            if state["left"] == 0:
                state["left"] = state["n"]
                self._kl_backward(state["layer"])
                self._handles.append((dist.all_reduce(self.flat[state["lo"]:state["hi"]],
                                                      op=dist.ReduceOp.SUM, group=self.group,
                                                      async_op=True), state["lo"], state["hi"]))
        return hook

    def set_shard(self, total_mc: int):
        """Local MC count for a global ``total_mc`` and registers this rank's global MC offset."""
        assert total_mc % self.world == 0, "MC must divide evenly across ranks"
        local = total_mc // self.world
        layers.set_mc_shard(self.rank * local)
        return local

    def step(self, x, y):
        """One optimiser step; returns a device tensor [LL, KL, ACC, loss] (global values)."""
        if self._graph is not None:
            self._gx.copy_(x, non_blocking=True)
            self._gy.copy_(y, non_blocking=True)
            self._graph.replay()
            self.adam_t += 1
            return self._gout.clone()
        return self._eager_step(x, y)

    def enable_cuda_graph(self, x, y, warmup: int = 3):
        """Capture one whole step (forward, loss, backward, all-reduce, optimiser) in a CUDA graph and replay it from
        now on: for small configurations the step is bound by Python / launch overhead, not by the GPU.
        A graph replays fixed launch parameters, so what changes from step to step moves to device memory: the
        per-step part of the eps seed (``layers.set_seed_device_offset``; the kernels draw with seed + offset) and
        Adam's bias-correction scalars (``pl_adam_kl_step_dev``), both advanced by small kernels inside the graph.
        ``x`` / ``y`` fix the input shapes; later ``step(x, y)`` calls copy into static buffers.  Runs ``warmup``
        ordinary steps first (they count as training steps)."""
        dev = self.flat.device
        if not (self.fused_opt and dev.type == "cuda"):
            raise RuntimeError("CUDA-graph mode needs the fused optimiser pass on a CUDA device")
        self._gx, self._gy = x.clone(), y.clone()
        # Layers keep tensors of their last forward (kl_div, SIVILayer's w_mu / w_std ...).  Those keep the previous
        # iteration's autograd graph -- and with it the AccumulateGrad nodes of the leaf parameters, which remember the
        # stream they were created on -- alive.  Detach them so that the warm-up below recreates the nodes on the capture
        # stream (otherwise autograd synchronises the capture with the legacy default stream: illegal during capture).
        for m in self.net.modules():
            for name, val in list(vars(m).items()):
                if isinstance(val, torch.Tensor) and val.grad_fn is not None:
                    setattr(m, name, val.detach())
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(int(warmup), 1)):
                self._eager_step(self._gx, self._gy)
        torch.cuda.current_stream(dev).wait_stream(side)
        self._seed_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self._t_dev = torch.full((1,), float(self.adam_t), dtype=torch.float64, device=dev)
        self._sched = torch.zeros(2, dtype=torch.float32, device=dev)
        layers.set_seed_device_offset(self._seed_dev)
        graph = torch.cuda.CUDAGraph()
        self._capturing = True
        torch.cuda.synchronize(dev)
        try:
            # capture on the stream the warm-up ran on: layers that keep last-forward tensors (SIVILayer's w_mu / kl_div,
            # hence the AccumulateGrad nodes of its hyper-net) must not tie the capture to the legacy default stream
            with torch.cuda.graph(graph, stream=side):
                self._gout = self._eager_step(self._gx, self._gy)
                # advance the device-side step state for the NEXT replay
                self._seed_dev.add_(_GOLDEN_SIGNED)
        finally:
            self._capturing = False
            layers.set_seed_device_offset(None)
        self.adam_t -= 1            # the capture pass enqueued nothing: it was not a step
        self._graph = graph
        return self

    def _eager_step(self, x, y):
        self.flat.zero_()
        if self.shard is not None:
            self._begin_opt_step()
        pred = self.net(x)
        inv = 1.0 / (self.num_samples * self.world)
        # cross entropy on CUDA: value, accuracy and d(loss)/d(pred) from one pass over the logits (csrc/ce.cu)
        fused_ce = isinstance(self.criterion, MC_CrossEntropyLoss) and pred.is_cuda and pred.dtype == torch.float32
        if fused_ce:
            rows = max(pred.shape[0] * pred.shape[1], 1)
            scale = self.criterion.num_samples / rows * inv
            ce_out, dpred = mc_cross_entropy_raw(pred.detach(), y, scale)
            ll = ce_out[0] * scale
        else:
            ll = self.criterion(pred, y) * inv
        kl_graph = None      # KL terms that still need autograd (SIVI hyper-net, Laplace prior)
        kl_value = torch.zeros((), device=pred.device)
        for m in self.net.modules():
            if isinstance(m, self.kl_kinds) and hasattr(m, "kl_div"):
                if m.kl_div.requires_grad:
                    kl_graph = m.kl_div if kl_graph is None else kl_graph + m.kl_div
                kl_value = kl_value + m.kl_div.detach()
        with torch.no_grad():               # before backward: the tail slice is reduced last
            self.stats[0] = ll.detach()
            self.stats[1] = kl_value * inv
            if self.with_accuracy:
                if fused_ce:
                    self.stats[2] = ce_out[1] / (rows * self.world)
                else:
                    self.stats[2] = MC_Accuracy(pred.detach(), y)[0] / self.world
        # overlap mode: hooks launch per-layer KL grad + all-reduce during the backward
        if fused_ce:
            if kl_graph is None:
                pred.backward(dpred)
            else:
                torch.autograd.backward([pred, kl_graph * inv], [dpred, None])
        else:
            loss = ll if kl_graph is None else ll + kl_graph * inv
            loss.backward()
        if self.shard is not None:
            return self._peer_finish()
        if not self.overlap:
            for m in self.kl_layers:
                self._kl_backward(m)
            if self.world > 1:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        else:
            for m in self._uncovered:
                self._kl_backward(m)
            for lo, hi in self._rest:
                self._handles.append((dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM,
                                                      group=self.group, async_op=True), lo, hi))
            if not self.fused_opt:
                for h, _, _ in self._handles:
                    h.wait()
                self._handles.clear()
        if self.fused_opt:
            self._fused_step()      # waits slice by slice: a layer's optimiser pass runs under the later slices' reduce
        else:
            self.opt.step()
        out = self.stats.clone()
        out[3] = out[0] + out[1]
        return out

    def _peer_finish(self):
        """Gradient reduce-scatter (peer stores), sharded Adam + KL, operand all-gather (peer stores): csrc/comm.cu."""
        sh = self.shard
        sched = self._sched_cur
        if self.world == 1:
            # one multi-tensor launch: Adam + KL gradient + KL value over every parameter tensor
            PF._timed("adam_kl_multi", lambda: sh.adam(1.0 / self.num_samples, self.lr, self.betas, self.eps, self.adam_t, sched,
                                                       zero_kl=False),
                      ("bytes", 28.0 * sh.plan.state_elems))
            self.stats[1] += ((sh.kl_acc[0] + self.kl_const) / self.num_samples).float()
            out = self.stats.clone()
            out[3] = out[0] + out[1]
            return out
        # what the side stream has not taken already (+ the statistics), then join it
        done = sorted(self._pushed + self._scattered)
        side_used = bool(self._pushed) or getattr(self, "_side_used", False)
        self._side_used = False
        self._pushed = []
        cur, rest = 0, []
        for lo, hi in done:
            if lo > cur:
                rest.append((cur, lo))
            cur = max(cur, hi)
        total = int(sh.ctx.push_units)
        if cur < total:
            rest.append((cur, total))
        if not rest:
            rest = [(total, total)]

        def push_rest():
            for i, r in enumerate(rest):
                sh.push(self.flat, self.stats if i == 0 else None, r)
            if side_used:
                ev = torch.cuda.Event()
                ev.record(self._side)
                torch.cuda.current_stream(self.flat.device).wait_event(ev)
        PF._timed("shard_push", push_rest, ("bytes", 4.0 * (self.flat.numel() - 4)))
        PF._timed("shard_barrier", sh.barrier)
        # the segments whose optimiser pass did not already run on the side stream
        done = set(self._adam_done)
        todo, cur = [], None
        for k in range(len(sh.plan.segs)):
            if k in done:
                if cur is not None:
                    todo.append((cur, k))
                    cur = None
            elif cur is None:
                cur = k
        if cur is not None:
            todo.append((cur, len(sh.plan.segs)))

        def adam_rest():
            for r in todo:
                sh.adam(1.0 / self.num_samples, self.lr, self.betas, self.eps, self.adam_t, sched, segs=r, zero_kl=False)
        PF._timed("shard_adam", adam_rest, ("bytes", 4.0 * (self.flat.numel() - 4) + 28.0 * sh.plan.state_elems))
        PF._timed("shard_finish", lambda: sh.finish(self.kl_const, 1.0 / self.num_samples, self.shard_out))
        return self.shard_out.clone()

    def _fused_step(self):
        """Adam + KL gradient + KL value, one streaming kernel per parameter tensor."""
        lib = _cabi.lib()
        self.adam_t += 1
        self.kl_acc.zero_()
        dev = self.flat.device
        stream = _cabi.current_stream(dev)
        kl_scale = 1.0 / self.num_samples       # gradients are already summed over ranks
        if self._capturing:                     # step-dependent scalars from device memory (see enable_cuda_graph)
            self._t_dev.add_(1.0)
            bc1 = 1.0 - torch.pow(torch.full_like(self._t_dev, self.betas[0]), self._t_dev)
            bc2 = 1.0 - torch.pow(torch.full_like(self._t_dev, self.betas[1]), self._t_dev)
            self._sched.copy_(torch.cat([self.lr / bc1, torch.rsqrt(bc2)]).float())
        # all-reduce slices complete in launch order; update each tensor as soon as ITS slice has arrived
        pending, waited = self._handles, 0

        def slice_of(off):
            for i, (_, lo, hi) in enumerate(pending):
                if lo <= off < hi:
                    return i
            return -1
        plan = sorted(self._adam_plan, key=lambda e: slice_of(e[1])) if pending else self._adam_plan
        with torch.cuda.device(dev):
            for p, off, n, kind, prior in plan:
                k = slice_of(off) if pending else -1
                while waited <= k:
                    pending[waited][0].wait()
                    waited += 1
                if self._capturing:
                    _cabi.check(lib.pl_adam_kl_step_dev(
                        p.data_ptr(), self.flat[off:off + n].data_ptr(), self.adam_m[off:off + n].data_ptr(),
                        self.adam_v[off:off + n].data_ptr(), n, kind, prior, kl_scale, self.betas[0], self.betas[1],
                        self.eps, self._sched.data_ptr(), self.kl_acc.data_ptr(), stream), "pl_adam_kl_step_dev")
                    continue
                PF._timed(f"adam_kl[{n}]", lambda: _cabi.check(
                    lib.pl_adam_kl_step(p.data_ptr(), self.flat[off:off + n].data_ptr(),
                                        self.adam_m[off:off + n].data_ptr(),
                                        self.adam_v[off:off + n].data_ptr(), n, kind, prior, kl_scale,
                                        self.lr, self.betas[0], self.betas[1], self.eps, self.adam_t,
                                        self.kl_acc.data_ptr(), stream), "pl_adam_kl_step"),
                          ("bytes", 28.0 * n))
        while waited < len(pending):            # the tail slice carries the LL / ACC statistics
            pending[waited][0].wait()
            waited += 1
        pending.clear()
        # KL value of the pre-update parameters (replica-identical: not part of the all-reduce)
        self.stats[1] += ((self.kl_acc[0] + self.kl_const) / self.num_samples).float()
