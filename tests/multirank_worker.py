"""Worker of tests/test_gpu_multirank.py: a small C4-shaped MC-sharded training run (fused bf16 chain + fp32 head + fused
CE + sharded / fused optimiser) on WORLD_SIZE GPUs; rank 0 saves the per-step statistics and the final parameters."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--exchange", default="auto")
    ap.add_argument("--graph", type=int, default=0)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--precision", default="auto")
    ap.add_argument("--grad-wire", default="fp32")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import pytorch_probabilisticlayers_b200 as plb
    from pytorch_probabilisticlayers_b200.fused import fuse_bayes_mlp
    from pytorch_probabilisticlayers_b200.train import ELBOTrainStep
    plb.set_precision(args.precision)
    plb.manual_seed(4711)
    torch.manual_seed(99)
    # every layer has >= 148 weight tiles of 128x128, so no dW kernel takes its atomic split path (whose summation order --
    # hence the last bit of the gradients -- changes from run to run and gets amplified by the bf16 chain over the steps)
    dims, mc_total, b, ns = [1664, 1536, 1664, 1536], 8, 256, 2000
    mods = [plb.MC_ExpansionLayer(mc_total // world, 2)]
    for li in range(3):
        mods.append(plb.BayesLinear(dims[li], dims[li + 1]))
        if li < 2:
            mods.append(torch.nn.LeakyReLU())
    net = torch.nn.Sequential(*mods).to(dev)
    with torch.no_grad():       # trained-size sigma so that the noise matters in every mode
        for m in net.modules():
            if isinstance(m, plb.BayesLinear):
                m.weight.logscale.fill_(-5.0)
    net = fuse_bayes_mlp(net, b)
    tr = ELBOTrainStep(net, plb.MC_CrossEntropyLoss(ns), ns, lr=1e-3, exchange=args.exchange, grad_wire=args.grad_wire)
    assert tr.set_shard(mc_total) == mc_total // world
    x = torch.randn(b, dims[0], device=dev)
    y = torch.randint(0, dims[-1], (b,), device=dev)
    stats = []
    if args.graph:
        for _ in range(2):
            stats.append(tr.step(x, y).tolist())
        tr.enable_cuda_graph(x, y, warmup=1)
        stats.append(None)                                    # the warm-up step inside enable_cuda_graph
        for _ in range(args.steps - 3):
            stats.append(tr.step(x, y).tolist())
    else:
        for _ in range(args.steps):
            stats.append(tr.step(x, y).tolist())
    torch.cuda.synchronize()
    tr.sync_full_params()
    status = tr.shard.status() if tr.shard is not None else 0
    if rank == 0:
        flat = torch.cat([p.detach().flatten() for p in net.parameters()]).cpu().numpy()
        np.savez(args.out, stats=np.array([s if s is not None else [np.nan] * 4 for s in stats], np.float64), params=flat,
                 exchange=str(tr.exchange), status=status, transport=str(getattr(tr.shard, "transport", "")),
                 fallback=str(getattr(tr, "exchange_fallback_reason", "")))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
