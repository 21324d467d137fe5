"""Host logic of the NCCL-free sharded step (pytorch_probabilisticlayers_b200/shard.py) on CPU:
  * ShardPlan: every element of every tensor has exactly one owner, staging / operand areas do not overlap, the unit
    prefix arrays the kernels index with are consistent, struct mirrors have the sizes the C header declares;
  * emulate_step (the pure-torch restatement of csrc/comm.cu's push + adam) equals dense torch.optim.Adam with the
    analytic KL gradient added, for 1, 2, 3 and 8 ranks;
  * world_size-2 gloo run: two ranks with different local gradients exchange them shard-wise (all_gather standing in for
    the peer stores), each updates only its shard, the gathered parameters equal the dense single-process update."""
import ctypes as C
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pytorch_probabilisticlayers_b200 import shard as S


def _toy_plan(world):
    numels = [24 * 16, 24 * 16, 16, 16, 16 * 10, 16 * 10, 10, 10, 7]
    segs = [S.SegSpec(1, 24 * 16, 0, 1, gather=1, rho_raw=1, prior_scale=0.7),
            S.SegSpec(1, 16 * 10, 4, 5, gather=0, rho_raw=0, prior_scale=1.3),
            S.SegSpec(0, 16, 2), S.SegSpec(0, 16, 3), S.SegSpec(0, 10, 6), S.SegSpec(0, 10, 7), S.SegSpec(0, 7, 8)]
    return S.ShardPlan(world, numels, segs), numels


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_plan_partitions_every_tensor_once(world):
    plan, numels = _toy_plan(world)
    for s in plan.segs:
        assert s.chunk % 8 == 0 and s.chunk * world >= s.numel
        assert sum(plan.owned(s, r) for r in range(world)) == s.numel
    # byte ranges: parameter area, stagings, operand areas, stats, kl, flags are disjoint and inside the arena
    spans = [(0, plan.param_bytes)]
    for s in plan.segs:
        spans.append((s.g_off, s.g_off + world * s.chunk * 4))
        if s.kind == 1:
            spans.append((s.g2_off, s.g2_off + world * s.chunk * 4))
        if s.gather:
            spans.append((s.lp_off, s.lp_off + s.numel * 4))
            assert s.lp_half == s.numel * 2
    spans += [(plan.stats_off, plan.stats_off + 2 * 8 * 16), (plan.kl_off, plan.kl_off + 2 * 8 * 8),
              (plan.flags_off, plan.flags_off + 4 * 64)]
    spans.sort()
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 <= b0
    assert spans[-1][1] <= plan.arena_bytes
    assert all(off % 32 == 0 for off in plan.param_offs)
    pp = plan.push_prefix()
    assert len(pp) == 2 * len(plan.segs) + 1 and pp[-1] == sum((n + 7) // 8 for n in numels)
    for r in range(world):
        ap = plan.adam_prefix(r)
        assert ap[-1] == sum((plan.owned(s, r) + 7) // 8 for s in plan.segs)
    assert plan.state_elems == sum(s.chunk * (2 if s.kind == 1 else 1) for s in plan.segs)


def test_struct_mirrors_match_the_header():
    """ctypes mirrors against sizeof() of the structs in the built library (no GPU needed)."""
    from pytorch_probabilisticlayers_b200 import _cabi
    lib = _cabi.lib()
    assert C.sizeof(S.ShardSeg) == 16 + 9 * 8
    assert C.sizeof(S.ShardCtx) == 16 + 8 * 8 + 16 + 5 * 8 + (2 * 48 + 1) * 8 + (48 + 1) * 8
    for which, mirror in enumerate((_cabi.LinearArgs, _cabi.Conv2dArgs, _cabi.CeArgs, _cabi.BnPoolArgs, S.ShardSeg, S.ShardCtx)):
        assert lib.pl_struct_size(which) == C.sizeof(mirror), (which, mirror.__name__)


def _dense_reference(plan, params, grads_sum, lr, steps, kl_scale):
    """torch.optim.Adam on the summed gradients + the analytic KL gradient (what N = 1 does)."""
    ps = [p.clone().requires_grad_(True) for p in params]
    opt = torch.optim.Adam(ps, lr)
    for t in range(steps):
        for i, p in enumerate(ps):
            p.grad = grads_sum[t][i].clone()
        for s in plan.segs:
            if s.kind != 1:
                continue
            mu, rho = ps[s.p_index], ps[s.p2_index]
            ip2 = 1.0 / s.prior_scale ** 2
            sig, sigma = torch.sigmoid(rho.detach()), torch.nn.functional.softplus(rho.detach())
            if s.rho_raw:
                rho.grad = rho.grad * sig
            mu.grad = mu.grad + kl_scale * ip2 * mu.detach()
            rho.grad = rho.grad + kl_scale * sig * (sigma * ip2 - 1 / sigma)
        opt.step()
    return [p.detach() for p in ps]


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_emulated_sharded_adam_equals_dense_adam(world):
    plan, numels = _toy_plan(world)
    g = torch.Generator().manual_seed(world)
    params = [torch.randn(n, generator=g) * 0.3 for n in numels]
    for s in plan.segs:
        if s.kind == 1:
            params[s.p2_index].uniform_(-4, -1, generator=g)
    steps, lr, kl_scale = 3, 1e-2, 1e-3
    grads = [[[torch.randn(n, generator=g) for n in numels] for _ in range(world)] for _ in range(steps)]
    want = _dense_reference(plan, params, [[sum(grads[t][r][i] for r in range(world)) for i in range(len(numels))]
                                          for t in range(steps)], lr, steps, kl_scale)
    got = [p.clone() for p in params]
    m = [torch.zeros(plan.state_elems) for _ in range(world)]
    v = [torch.zeros(plan.state_elems) for _ in range(world)]
    for t in range(steps):
        S.emulate_step(plan, got, grads[t], m, v, lr, (0.9, 0.999), 1e-8, t + 1, kl_scale)
    for a, b in zip(got, want):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_rank(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    plan, numels = _toy_plan(world)
    g = torch.Generator().manual_seed(11)
    params = [torch.randn(n, generator=g) * 0.3 for n in numels]
    for s in plan.segs:
        if s.kind == 1:
            params[s.p2_index].uniform_(-4, -1, generator=g)
    m, v = torch.zeros(plan.state_elems), torch.zeros(plan.state_elems)
    for t in range(3):
        gl = torch.Generator().manual_seed(100 * t + rank)                  # this rank's local gradients
        local = [torch.randn(n, generator=gl) for n in numels]
        # reduce-scatter stand-in: every rank learns every rank's partials, but only uses them for ITS shard
        staged = [None] * world
        dist.all_gather_object(staged, local)
        mine = [p.clone() for p in params]
        ms = [torch.zeros(plan.state_elems) for _ in range(world)]
        vs = [torch.zeros(plan.state_elems) for _ in range(world)]
        ms[rank], vs[rank] = m, v
        S.emulate_step(plan, mine, staged, ms, vs, 1e-2, (0.9, 0.999), 1e-8, t + 1, 1e-3)
        # all-gather stand-in: only the owner's chunk of every tensor is authoritative
        for s in plan.segs:
            for pi in ([s.p_index] + ([s.p2_index] if s.kind == 1 else [])):
                chunk = torch.zeros(s.chunk)
                n = plan.owned(s, rank)
                chunk[:n] = mine[pi][rank * s.chunk:rank * s.chunk + n]
                full = [torch.zeros(s.chunk) for _ in range(world)]
                dist.all_gather(full, chunk)
                params[pi] = torch.cat(full)[:s.numel].clone()
    np.save(os.path.join(out_dir, f"r{rank}.npy"), torch.cat(params).numpy())
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_step_equals_dense(tmp_path):
    d = str(tmp_path)
    mp.spawn(_gloo_rank, args=(2, _free_port(), d), nprocs=2, join=True)
    r0, r1 = np.load(os.path.join(d, "r0.npy")), np.load(os.path.join(d, "r1.npy"))
    assert np.array_equal(r0, r1)
    plan, numels = _toy_plan(2)
    g = torch.Generator().manual_seed(11)
    params = [torch.randn(n, generator=g) * 0.3 for n in numels]
    for s in plan.segs:
        if s.kind == 1:
            params[s.p2_index].uniform_(-4, -1, generator=g)
    sums = []
    for t in range(3):
        per_rank = [[torch.randn(n, generator=torch.Generator().manual_seed(100 * t + r)) for n in numels] for r in range(2)]
        # generators must be consumed sequentially per rank as in _gloo_rank
        per_rank = []
        for r in range(2):
            gl = torch.Generator().manual_seed(100 * t + r)
            per_rank.append([torch.randn(n, generator=gl) for n in numels])
        sums.append([per_rank[0][i] + per_rank[1][i] for i in range(len(numels))])
    want = torch.cat(_dense_reference(plan, params, sums, 1e-2, 3, 1e-3)).numpy()
    assert np.allclose(r0, want, rtol=1e-5, atol=1e-6)
