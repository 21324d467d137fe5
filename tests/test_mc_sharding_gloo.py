"""world_size-2 gloo test (CPU) of the MC-sharded train step's host logic: shard offsets, the
1/N scaling of likelihood and KL, the single flat all-reduce, identical replicas after the step.

The CUDA layers cannot run here, so a stand-in layer with the same contract (per-MC eps indexed by
GLOBAL MC index, ``kl_div`` attribute) is built from the oracle's torch port; the trainer code under
test is the product's ``ELBOTrainStep``."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _StandInBayesLinear(torch.nn.Module):
    """BayesLinear semantics in plain torch with eps looked up by global MC index."""

    def __init__(self, i, o, eps_w, eps_b):
        super().__init__()
        g = torch.Generator().manual_seed(i * 131 + o)
        self.w_mu = torch.nn.Parameter(torch.randn(i, o, generator=g) * 0.2)
        self.w_rho = torch.nn.Parameter(torch.full((i, o), -2.0))
        self.b_mu = torch.nn.Parameter(torch.zeros(o))
        self.b_rho = torch.nn.Parameter(torch.full((o,), -2.0))
        self.eps_w, self.eps_b = eps_w, eps_b      # [MC_total, ...]

    def forward(self, x):
        from pytorch_probabilisticlayers_b200 import layers
        lo = layers._STATE["mc_global_offset"]
        mc = x.shape[0]
        sp = torch.nn.functional.softplus
        w = self.w_mu + sp(self.w_rho) * self.eps_w[lo:lo + mc]
        b = self.b_mu + sp(self.b_rho) * self.eps_b[lo:lo + mc]
        s = sp(self.w_rho)
        self.kl_div = (0.5 * (s ** 2 + self.w_mu ** 2 - 1) - torch.log(s)).sum()
        return torch.baddbmm(b.unsqueeze(1), x, w)


def _build(mc_total, mc_local):
    from pytorch_probabilisticlayers_b200 import MC_ExpansionLayer
    g = torch.Generator().manual_seed(7)
    e = lambda *s: torch.randn(*s, generator=g)
    return torch.nn.Sequential(
        MC_ExpansionLayer(mc_local, 2),
        _StandInBayesLinear(6, 8, e(mc_total, 6, 8), e(mc_total, 8)), torch.nn.LeakyReLU(),
        _StandInBayesLinear(8, 4, e(mc_total, 8, 4), e(mc_total, 4)))


def _run_rank(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    from pytorch_probabilisticlayers_b200 import MC_CrossEntropyLoss
    from pytorch_probabilisticlayers_b200.train import ELBOTrainStep
    mc_total, ns = 8, 500
    net = _build(mc_total, mc_total // world)
    tr = ELBOTrainStep(net, MC_CrossEntropyLoss(ns), ns, lr=1e-2,
                       kl_kinds=(_StandInBayesLinear,))
    assert tr.set_shard(mc_total) == mc_total // world
    g = torch.Generator().manual_seed(3)
    x = torch.randn(5, 6, generator=g)
    y = torch.randint(0, 4, (5,), generator=g)
    stats = [tr.step(x, y).clone() for _ in range(3)]
    flat = torch.cat([p.detach().flatten() for p in net.parameters()])
    np.save(os.path.join(out_dir, f"w{world}_r{rank}.npy"),
            torch.cat([flat, torch.stack(stats).flatten()]).numpy())
    if world > 1:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_mc_sharding_equals_single_rank(tmp_path):
    d = str(tmp_path)
    _run_rank(0, 1, _free_port(), d)
    mp.spawn(_run_rank, args=(2, _free_port(), d), nprocs=2, join=True)
    single = np.load(os.path.join(d, "w1_r0.npy"))
    r0 = np.load(os.path.join(d, "w2_r0.npy"))
    r1 = np.load(os.path.join(d, "w2_r1.npy"))
    assert np.array_equal(r0, r1)                       # replicas stay identical
    assert np.allclose(r0, single, rtol=2e-5, atol=1e-6)  # == unsharded run (reduction order only)


def test_mc_coupled_losses_refuse_sharding():
    from pytorch_probabilisticlayers_b200.train import ELBOTrainStep
    net = torch.nn.Linear(2, 2)
    ELBOTrainStep(net, None, 10, mc_coupled_loss=True)   # fine unsharded
