"""Host side of the NCCL-free MC-sharded step (csrc/comm.cu, ``pl_shard_*`` in include/pl_b200.h).

Reference being replaced for the N-GPU step: Lightning DDP's gradient all-reduce followed by a replicated
``torch.optim.Adam.step`` (experiments/BNN.py:349, :399-400).  Here every rank owns 1/N of every parameter tensor:

  * the local gradients are scattered into the owners' staging areas by peer stores over NVLink (reduce-scatter),
  * the owner sums the N partials in rank order, adds the analytic KL gradient, runs Adam on its shard (the optimiser
    state exists only for the shard) and writes the updated parameters straight into every rank's memory in the form the
    forward consumes: the bf16 {mu | sigma} operand of the tcgen05 kernels for the wide weights (4 B/weight instead of
    two fp32 tensors), fp32 for the small tensors (all-gather),
  * two device-side barriers per step; no NCCL call, no host synchronisation: the whole step is one stream of kernels
    and can be captured in a CUDA graph.

``ShardPlan`` is pure Python (unit-tested on CPU); ``PeerShard`` owns the arena, the CUDA-IPC mappings and the C structs.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import List, Optional

import torch
import torch.distributed as dist

from . import _cabi

MAX_RANKS, MAX_SEGS = 8, 48


class ShardSeg(C.Structure):
    _fields_ = [("kind", C.c_int32), ("gather", C.c_int32), ("rho_raw", C.c_int32), ("prior_scale", C.c_float),
                ("numel", C.c_int64), ("chunk", C.c_int64), ("p_off", C.c_int64), ("p2_off", C.c_int64),
                ("g_off", C.c_int64), ("g2_off", C.c_int64), ("s_off", C.c_int64), ("lp_off", C.c_int64),
                ("lp_half", C.c_int64)]


class ShardCtx(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("num_segs", C.c_int32), ("wire_bf16", C.c_int32),
                ("peers", C.c_void_p * MAX_RANKS), ("segs", C.c_void_p), ("mc_base", C.c_void_p),
                ("stats_off", C.c_int64), ("kl_off", C.c_int64), ("flags_off", C.c_int64),
                ("push_units", C.c_int64), ("adam_units", C.c_int64),
                ("push_prefix", C.c_int64 * (2 * MAX_SEGS + 1)), ("adam_prefix", C.c_int64 * (MAX_SEGS + 1))]


def _align(v, a=256):
    return (v + a - 1) // a * a


@dataclass
class SegSpec:
    """One optimiser segment: a plain tensor, or the (mean, rho) pair of one weight."""
    kind: int                      # 0 plain, 1 pair
    numel: int
    p_index: int                   # index of the (mean) tensor in the trainer's parameter list
    p2_index: int = -1             # pair: index of the rho tensor
    gather: int = 0                # 0 fp32, 1 bf16 split operand, 2 bf16 packed operand
    rho_raw: int = 0
    prior_scale: float = 1.0
    # filled by ShardPlan
    chunk: int = 0
    p_off: int = 0
    p2_off: int = 0
    g_off: int = 0
    g2_off: int = 0
    s_off: int = 0
    lp_off: int = 0
    lp_half: int = 0


@dataclass
class ShardPlan:
    """Arena layout + segment table for ``world`` ranks.  ``param_numels[i]`` = numel of parameter i (trainer order, the
    order of the flat gradient buffer); ``segs`` reference parameters by index.  Every parameter must be covered once."""
    world: int
    param_numels: List[int]
    segs: List[SegSpec]
    param_offs: List[int] = field(default_factory=list)     # BYTE offset of parameter i in the arena's fp32 area
    param_bytes: int = 0
    arena_bytes: int = 0
    state_elems: int = 0
    stats_off: int = 0
    kl_off: int = 0
    flags_off: int = 0

    def __post_init__(self):
        assert 1 <= self.world <= MAX_RANKS and len(self.segs) <= MAX_SEGS
        covered = sorted([s.p_index for s in self.segs] + [s.p2_index for s in self.segs if s.kind == 1])
        assert covered == list(range(len(self.param_numels))), "every parameter must belong to exactly one segment"
        off = 0
        self.param_offs = []
        for n in self.param_numels:          # fp32 parameter area: tensors 32-byte aligned (float4 pairs never straddle)
            self.param_offs.append(off)
            off += _align(n * 4, 32)
        self.param_bytes = off
        off = _align(off)
        s_off = 0
        for s in self.segs:
            s.chunk = (math.ceil(s.numel / self.world) + 7) // 8 * 8
            s.p_off = self.param_offs[s.p_index]
            s.g_off = off
            off += _align(self.world * s.chunk * 4)
            s.s_off = s_off
            s_off += s.chunk
            if s.kind == 1:
                assert self.param_numels[s.p2_index] == s.numel
                s.p2_off = self.param_offs[s.p2_index]
                s.g2_off = off
                off += _align(self.world * s.chunk * 4)
                s_off += s.chunk
        self.state_elems = s_off
        for s in self.segs:
            if s.gather:
                assert s.kind == 1 and s.numel % 8 == 0
                s.lp_off = off
                s.lp_half = s.numel * 2 if s.gather == 1 else 0
                off += _align(s.numel * 4)
        self.stats_off = off
        off += _align(2 * MAX_RANKS * 16)   # two slots (alternating by step) x 8 ranks x 4 floats
        self.kl_off = off
        off += _align(2 * MAX_RANKS * 8)
        self.flags_off = off
        off += _align(4 * 64)               # 4 barrier channels x 16 words
        self.arena_bytes = off

    # 8-element units ------------------------------------------------------------------------------------------
    def push_prefix(self):
        pre = [0]
        for s in self.segs:
            units = (s.numel + 7) // 8
            pre.append(pre[-1] + units)
            pre.append(pre[-1] + (units if s.kind == 1 else 0))
        return pre

    def owned(self, s: SegSpec, rank: int):
        """Number of elements of segment s owned by ``rank``."""
        return max(0, min(s.chunk, s.numel - rank * s.chunk))

    def adam_prefix(self, rank: int):
        pre = [0]
        for s in self.segs:
            pre.append(pre[-1] + (self.owned(s, rank) + 7) // 8)
        return pre


def emulate_step(plan: ShardPlan, params, grads_per_rank, m, v, lr, betas, eps, step, kl_scale):
    """Pure-torch CPU restatement of push + adam of csrc/comm.cu for ONE logical parameter set (tests): ``params`` is the
    list of fp32 tensors (updated in place), ``grads_per_rank[r][i]`` rank r's local gradient of parameter i, ``m`` / ``v``
    the per-rank shard states (lists of 1-D tensors of plan.state_elems).  Returns the KL value partial of every rank."""
    world = plan.world
    b1, b2 = betas
    step_size = lr / (1 - b1 ** step)
    inv_sqrt_bc2 = 1.0 / math.sqrt(1 - b2 ** step)
    kls = [0.0] * world
    for s in plan.segs:
        tensors = [(s.p_index, 1 if s.kind == 1 else 0, 0)] + ([(s.p2_index, 2, s.chunk)] if s.kind == 1 else [])
        for pi, role, soff in tensors:
            flat = params[pi].view(-1)
            for r in range(world):
                lo, n = r * s.chunk, plan.owned(s, r)
                if n == 0:
                    continue
                g = sum(grads_per_rank[q][pi].reshape(-1)[lo:lo + n] for q in range(world))
                w = flat[lo:lo + n]
                ip2 = 1.0 / s.prior_scale ** 2
                if role == 1:
                    g = g + kl_scale * ip2 * w
                    kls[r] += float((0.5 * ip2 * w.double() ** 2).sum())
                elif role == 2:
                    sigma, sig = torch.nn.functional.softplus(w), torch.sigmoid(w)
                    if s.rho_raw:
                        g = g * sig
                    g = g + kl_scale * sig * (sigma * ip2 - 1 / sigma)
                    kls[r] += float((0.5 * ip2 * sigma.double() ** 2 - sigma.double().log()).sum())
                ms = m[r][s.s_off + soff:s.s_off + soff + n]
                vs = v[r][s.s_off + soff:s.s_off + soff + n]
                ms.mul_(b1).add_(g, alpha=1 - b1)
                vs.mul_(b2).addcmul_(g, g, value=1 - b2)
                w.sub_(step_size * ms / (vs.sqrt() * inv_sqrt_bc2 + eps))
    return kls


class PeerShard:
    """Arena + peer mappings + C context of one rank.  ``params``: the trainer's parameters in flat-gradient order."""

    def __init__(self, plan: ShardPlan, params, group=None, wire_bf16=False, use_symmetric_memory=True):
        self.plan = plan
        self.wire_bf16 = bool(wire_bf16) and plan.world > 1
        self.group = group
        multi = plan.world > 1
        self.rank = dist.get_rank(group) if multi else 0
        self.world = dist.get_world_size(group) if multi else 1
        assert self.world == plan.world
        self.lib = _cabi.lib()
        dev = params[0].device
        self.device = dev
        self.transport = "local"
        self._symm = None
        with torch.cuda.device(dev):
            self.arena = None
            import os
            if multi and use_symmetric_memory and os.environ.get("PL_SHARD_NO_SYMM", "0") != "1":
                # torch's symmetric memory: peer mappings + (on NVSwitch systems) an NVLS multicast mapping of the arenas
                try:
                    import torch.distributed._symmetric_memory as symm
                    arena = symm.empty(plan.arena_bytes, dtype=torch.uint8, device=dev)
                    arena.zero_()
                    torch.cuda.synchronize(dev)
                    name = (group or dist.group.WORLD).group_name
                    self._symm = symm.rendezvous(arena, name)
                    self.arena = arena
                    self.transport = "symmetric_memory" + ("+multicast" if self._symm.multicast_ptr else "")
                except Exception as e:      # older driver / no fabric support: CUDA IPC below
                    self._symm = None
                    self.symm_fallback_reason = f"{type(e).__name__}: {e}"[:200]
            if self.arena is None:
                self.arena = torch.zeros(plan.arena_bytes, dtype=torch.uint8, device=dev)
                if multi:
                    self.transport = "cuda_ipc"
            # parameters move into the arena's fp32 area (same values): the owners' optimiser pass writes the updated
            # small tensors straight into every rank's copy
            for p, off in zip(params, plan.param_offs):
                view = self.arena[off:off + p.numel() * 4].view(torch.float32).view(p.shape)
                view.copy_(p.data)
                p.data = view
            torch.cuda.synchronize(dev)
            self._opened = []
            peers = []
            everyone = [None]
            if multi and self._symm is not None:
                everyone = []
                peers = [int(p) for p in self._symm.buffer_ptrs]
                assert peers[self.rank] == self.arena.data_ptr()
            elif multi:
                handle = (C.c_uint8 * 64)()
                offset = C.c_int64(0)
                _cabi.check(self.lib.pl_ipc_export(C.c_void_p(self.arena.data_ptr()), handle, C.byref(offset)),
                            "pl_ipc_export")
                mine = (bytes(handle), int(offset.value))
                everyone = [None] * self.world
                dist.all_gather_object(everyone, mine, group=group)
            for r, item in enumerate(everyone):
                if r == self.rank:
                    peers.append(self.arena.data_ptr())
                    continue
                h, off = item
                out = C.c_void_p()
                buf = (C.c_uint8 * 64).from_buffer_copy(h)
                _cabi.check(self.lib.pl_ipc_open(buf, C.c_int64(off), C.byref(out)), "pl_ipc_open")
                peers.append(out.value)
                self._opened.append((out.value, off))
            self.peers = peers
            segs = (ShardSeg * max(len(plan.segs), 1))()
            for i, s in enumerate(plan.segs):
                segs[i] = ShardSeg(s.kind, s.gather, s.rho_raw, float(s.prior_scale), s.numel, s.chunk, s.p_off, s.p2_off,
                                   s.g_off, s.g2_off, s.s_off, s.lp_off, s.lp_half)
            raw = bytes(segs)
            self.segs_dev = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
            ctx = ShardCtx()
            ctx.rank, ctx.world, ctx.num_segs = self.rank, self.world, len(plan.segs)
            ctx.wire_bf16 = 1 if self.wire_bf16 else 0
            for r, ptr in enumerate(peers):
                ctx.peers[r] = ptr
            ctx.segs = self.segs_dev.data_ptr()
            ctx.mc_base = int(self._symm.multicast_ptr) if (self._symm is not None and self._symm.multicast_ptr) else None
            ctx.stats_off, ctx.kl_off, ctx.flags_off = plan.stats_off, plan.kl_off, plan.flags_off
            pp, ap = plan.push_prefix(), plan.adam_prefix(self.rank)
            for i, vv in enumerate(pp):
                ctx.push_prefix[i] = vv
            for i, vv in enumerate(ap):
                ctx.adam_prefix[i] = vv
            ctx.push_units, ctx.adam_units = pp[-1], ap[-1]
            self.ctx = ctx
            self.m = torch.zeros(max(plan.state_elems, 8), dtype=torch.float32, device=dev)
            self.v = torch.zeros(max(plan.state_elems, 8), dtype=torch.float32, device=dev)
            self.kl_acc = torch.zeros(1, dtype=torch.float64, device=dev)
        if multi:
            dist.barrier(group=group)  # every rank has mapped every arena before anyone stores into a peer

    def _stream(self):
        return _cabi.current_stream(self.device)

    def staging_view(self, seg: SegSpec, second=False):
        """fp32 view of this rank's own staging row of a segment (world == 1: the gradient buffer itself)."""
        off = (seg.g2_off if second else seg.g_off) + self.rank * seg.chunk * 4
        return self.arena[off:off + seg.numel * 4].view(torch.float32)

    def lp_pointers(self, seg: SegSpec):
        """(device pointer of the operand-form weight in OUR arena, byte distance to its sigma half)."""
        return self.arena.data_ptr() + seg.lp_off, seg.lp_half

    def push(self, grads_flat, stats4, units=None):
        """Scatter the local gradients to their owners; ``units`` = (begin, end) restricts it to a range of push units."""
        u0, u1 = units if units is not None else (0, int(self.ctx.push_units))
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.pl_shard_push_range(C.byref(self.ctx), C.c_void_p(grads_flat.data_ptr()),
                                                     C.c_void_p(stats4.data_ptr() if stats4 is not None else None),
                                                     C.c_int64(u0), C.c_int64(u1), C.c_void_p(self._stream())),
                        "pl_shard_push")

    def seg_units(self, seg_index):
        """Push-unit range of segment ``seg_index`` (both tensors of a pair)."""
        pp = self.plan.push_prefix()
        return pp[2 * seg_index], pp[2 * seg_index + 2]

    def barrier(self, channel=0):
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.pl_shard_barrier(C.byref(self.ctx), C.c_int32(channel), C.c_void_p(self._stream())),
                        "pl_shard_barrier")

    def adam(self, kl_scale, lr, betas, eps, step, sched_dev=None, with_kl=True, segs=None, zero_kl=True):
        """Optimiser pass + gather over the segments ``segs`` = (begin, end) (default: all).  ``zero_kl``: clear the KL
        accumulator first (once per step when the pass is issued in several pieces)."""
        s0, s1 = segs if segs is not None else (0, len(self.plan.segs))
        with torch.cuda.device(self.device):
            if zero_kl:
                self.kl_acc.zero_()
            _cabi.check(self.lib.pl_shard_adam_range(
                C.byref(self.ctx), C.c_int32(s0), C.c_int32(s1), C.c_void_p(self.m.data_ptr()),
                C.c_void_p(self.v.data_ptr()), C.c_float(kl_scale),
                C.c_float(lr), C.c_float(betas[0]), C.c_float(betas[1]), C.c_float(eps), C.c_int32(step),
                C.c_void_p(sched_dev.data_ptr() if sched_dev is not None else None),
                C.c_void_p(self.kl_acc.data_ptr() if with_kl else None), C.c_void_p(self._stream())), "pl_shard_adam")

    def finish(self, kl_const, inv_ns, stats_out):
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.pl_shard_finish(C.byref(self.ctx), C.c_void_p(self.kl_acc.data_ptr()), C.c_double(kl_const),
                                                 C.c_double(inv_ns), C.c_void_p(stats_out.data_ptr()),
                                                 C.c_void_p(self._stream())), "pl_shard_finish")

    def initial_gather(self):
        """Fills every rank's operand areas (and fp32 copies) from the owners' initial parameters: one optimiser pass with
        zero gradients and lr = 0 (parameters unchanged), then the optimiser state is cleared again."""
        self.barrier()
        self.adam(0.0, 0.0, (0.9, 0.999), 1e-8, 1, with_kl=False)
        self.barrier()
        self.m.zero_()
        self.v.zero_()

    def status(self):
        """0 = fine, 1 = a barrier timed out (a peer never arrived)."""
        flags = self.arena[self.plan.flags_off:self.plan.flags_off + 8].view(torch.int32)
        return int(flags[1].item())

    def sync_full_params(self, params):
        """Brings the fp32 values of the operand-gathered weights (kept current only on their owners) to every rank:
        for checkpoints / evaluation, not part of the step."""
        for s in self.plan.segs:
            if not s.gather:
                continue
            for pi in (s.p_index, s.p2_index):
                flat = params[pi].data.view(-1)
                mine = torch.zeros(s.chunk, dtype=torch.float32, device=self.device)
                n = self.plan.owned(s, self.rank)
                if n:
                    mine[:n] = flat[self.rank * s.chunk:self.rank * s.chunk + n]
                full = torch.empty(self.world * s.chunk, dtype=torch.float32, device=self.device)
                dist.all_gather_into_tensor(full, mine, group=self.group)
                flat.copy_(full[:s.numel])

    def close(self):
        for ptr, off in self._opened:
            self.lib.pl_ipc_close(C.c_void_p(ptr), C.c_int64(off))
        self._opened = []
