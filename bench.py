#!/usr/bin/env python
"""Headline benchmark: MC-sample*examples / second of one ELBO train step
(forward + likelihood + KL + backward [+ gradient exchange] + Adam) -- BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c2|c3|c1|c5] [--precision ...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's own CPU path (oracle/_ref, unmodified), host cores

Workloads (synthetic data, reference initialisation):
  c4  wide Bayesian MLP 4096-4096-4096-1000, MC=64, batch 512, num_samples 100000  (default: the
      config north_star's targets are quoted on; fits one GPU; MC-sharded over N GPUs -> strong scaling)
  c2  MNIST-shaped MLP 784-1200-1200-10, MC=32, batch 256, num_samples 55000
  c3  CIFAR-shaped Bayesian ConvNet, MC=16, batch 128 (single GPU: MC_BatchNorm2D couples the samples)
  c1  1-D regression MLP 1-49-51-52-1, MC=10, batch 50, MC_NLL (launch-latency bound: reports us/step)
  c5  SIVI last layer on a deterministic feature extractor, MC=128, batch 100, MC_NLL

One JSON line on stdout (rank 0).  `value` = device-resident inputs, CUDA-event timed, max over
ranks.  `e2e` = the same step through the public API with pinned HOST inputs copied in and the
loss read back every step.  `roofline` = the dominant kernel, timed live with CUDA events (`frac` against the
sustained measured peak, `frac_burst` against the burst one).  Secondary objects (N=1, rank 0 only; each
bounded to a few seconds): `tf32` (the same step in the primary-tolerance TF32 mode), `gpu_eager_reference`
(the UNMODIFIED reference run eagerly on this GPU, allow_tf32 off / on: the real bar), `cpu_baseline`
(the unmodified reference on the host cores, bounded MC sample).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    "c4": dict(dims=[4096, 4096, 4096, 1000], mc=64, batch=512, num_samples=100000, lr=1e-3, cpu_mc=8,
               name="C4 wide Bayesian MLP 4096-4096-4096-1000, MC=64, batch 512"),
    "c2": dict(dims=[784, 1200, 1200, 10], mc=32, batch=256, num_samples=55000, lr=1e-3, cpu_mc=8,
               name="C2 MNIST-shaped Bayesian MLP 784-1200-1200-10, MC=32, batch 256"),
    # CIFAR-shaped Bayesian ConvNet (experiments/BNN.py:228-253); single GPU only: MC_BatchNorm2D
    # pools statistics over MC
    "c3": dict(conv=True, in_dims=(3, 32, 32), dims=[3 * 32 * 32, 10], mc=16, batch=128, cpu_mc=2,
               num_samples=50000, lr=1e-3, flops=1.8334e12,
               name="C3 CIFAR-shaped Bayesian ConvNet 3-96-128-256-128 (k5,p2)+BN+pool, 512-200-10, "
                    "MC=16, batch 128"),
    # regression configs: MC_NLL couples the samples -> single GPU ("replicas only"), latency bound
    "c1": dict(regression="mlp", dims=[1, 49, 51, 52, 1], mc=10, batch=50, num_samples=1000, lr=1e-2, cpu_mc=10,
               name="C1 1-D regression Bayesian MLP 1-49-51-52-1 (ReLU), MC=10, batch 50, MC_NLL"),
    "c5": dict(regression="sivi", dims=[1, 50, 50, 1], mc=128, batch=100, num_samples=1000, lr=1e-2, cpu_mc=128,
               name="C5 SIVI last layer: Linear(1,50)-Linear(50,50)-SIVILayer(50,1,5), MC=128, batch 100, MC_NLL"),
}
METRIC = "mc_samples_x_examples_per_sec_per_train_step"
UNIT = "MC-sample*examples/s"


def workload_config(wl):
    """The `config` object: identical for this repo's arm and the reference arm (the driver compares them)."""
    return {"workload": wl["name"], "mc_total": wl["mc"], "batch": wl["batch"], "num_samples": wl["num_samples"],
            "l2": "per-step working set (parameters + activations >> 126 MB) exceeds L2; no explicit flush"
                  if not wl.get("regression") else "latency-bound toy config (fits L2); us/step is the figure of merit"}


def step_flops(dims, mc, batch):
    """Algorithmic FLOPs of one step: 2*MC*B*I*O x (fwd, dW, dX; no dX for
    the first layer)."""
    total = 0
    for li in range(len(dims) - 1):
        total += 2 * mc * batch * dims[li] * dims[li + 1] * (2 if li == 0 else 3)
    return total


# DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernels at the
# C4 shapes, from the committed `ncu --set full` captures under profiles/ (file named per entry)
NCU_TRAFFIC_BYTES = {
    "linear_fwd[64x512x4096x4096]": 1.293e9 + 0.510e9,      # profiles/r01_ncu_full_fwd_v3_summary.json
    "linear_bwd_dx[64x512x4096x4096]": 2.619e9 + 0.257e9,   # profiles/r01_ncu_full_final_summary.json
    "linear_bwd_dw[64x512x4096x4096]": 2.281e9 + 0.130e9,   # profiles/r01_ncu_full_dw_v3_summary.json
}
_traffic_override = os.path.join(ROOT, "profiles", "ncu_traffic.json")     # refreshed captures of later rounds
if os.path.exists(_traffic_override):
    with open(_traffic_override) as _f:
        NCU_TRAFFIC_BYTES.update({k: v for k, v in json.load(_f).items() if not k.startswith("_")})


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=float(p["hbm_gbs"]), tf_burst=float(p["bf16_tflops"]),
                    tf_sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0,
                source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own modules (oracle/_ref, unmodified) or, if they did not travel, the
# op-for-op port (oracle/torch_port.py); on the host cores, or eagerly on a CUDA device
# ------------------------------------------------------------------------------------------------
def reference_throughput(wl, mc_sample, steps, warmup, device="cpu", allow_tf32=None):
    """-> (value, ms_per_step, kind, threads).  Times `steps` train steps of the reference after `warmup`."""
    import torch
    from oracle import make_ref
    use_ref = make_ref.available()
    cores = os.cpu_count() or 1
    if device == "cpu":
        torch.set_num_threads(cores)
    dev = torch.device(device)
    if allow_tf32 is not None:
        torch.backends.cuda.matmul.allow_tf32 = bool(allow_tf32)
        torch.backends.cudnn.allow_tf32 = bool(allow_tf32)
    torch.manual_seed(1234)
    ns, batch = wl["num_samples"], wl["batch"]
    if use_ref:
        from oracle import ref_harness as rh
        if wl.get("regression"):
            net = rh.build_regression_mlp(mc_sample) if wl["regression"] == "mlp" else rh.SIVILastLayerNet(mc_sample)
            crit, step_fn = rh.nll_criterion(ns), rh.train_step_regression
        else:
            net = rh.build_convnet(mc_sample, wl["in_dims"]) if wl.get("conv") else rh.build_mlp(wl["dims"], mc_sample)
            crit, step_fn = rh.criterion(ns), rh.train_step
    else:
        from oracle import torch_port as tp
        if wl.get("regression"):
            raise RuntimeError("the regression configs need oracle/_ref (python oracle/make_ref.py)")
        net = tp.build_convnet(mc_sample, wl["in_dims"]) if wl.get("conv") else tp.build_mlp(wl["dims"], mc_sample)
        crit = tp.MCCrossEntropyPort(ns)

        def step_fn(net, crit, opt, x, y, ns):
            return tp.train_step(net, crit, opt, x, y, ns)
    net = net.to(dev)
    opt = torch.optim.Adam(net.parameters(), wl["lr"])
    if wl.get("regression"):
        x, y = torch.randn(batch, 1, device=dev), torch.randn(batch, 1, device=dev)
    elif wl.get("conv"):
        x, y = torch.randn(batch, *wl["in_dims"], device=dev), torch.randint(0, wl["dims"][-1], (batch,), device=dev)
    else:
        x, y = torch.randn(batch, wl["dims"][0], device=dev), torch.randint(0, wl["dims"][-1], (batch,), device=dev)

    def sync():
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)
    for _ in range(warmup):
        step_fn(net, crit, opt, x, y, ns)
    sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        step_fn(net, crit, opt, x, y, ns)
    sync()
    t = (time.perf_counter() - t0) / steps
    return mc_sample * batch / t, t * 1e3, ("reference" if use_ref else "port"), cores


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    mc_sample = min(wl["mc"], wl["cpu_mc"])
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    value, ms, kind, cores = reference_throughput(wl, mc_sample, steps, warmup)
    sample = (f"MC={mc_sample} of {wl['mc']} per step (throughput in MC-sample*examples/s is MC-independent on the CPU "
              f"path: every op is O(MC)), full batch {wl['batch']}, {steps} timed steps after {warmup} warm-up, "
              f"torch CPU fp32, {cores} threads")
    what = ("UNMODIFIED reference modules (oracle/_ref = copy of src/ProbabilisticLayers*.py) under "
            "oracle/ref_harness.py" if kind == "reference" else "oracle/torch_port.py (op-for-op port; oracle/_ref absent)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(wl),
        "impl_config": {"mc_sample": mc_sample, "what": what},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "20", "-i", str(index)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f:
            parts = [c.strip() for c in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        self.f.close()
        os.unlink(self.f.name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def build_net(plb, wl, mc_local, dev, precision, fuse, world=1):
    import torch
    dims, batch = wl["dims"], wl["batch"]
    if wl.get("regression") == "mlp":       # experiments/RegressionBNN.py:73-81
        mods = [plb.MC_ExpansionLayer(num_MC=mc_local, input_dim=2)]
        for li in range(len(dims) - 1):
            mods.append(plb.BayesLinear(dims[li], dims[li + 1]))
            if li + 2 < len(dims):
                mods.append(torch.nn.ReLU())
        return torch.nn.Sequential(*mods).to(dev)
    if wl.get("regression") == "sivi":      # experiments/SIVILastLayer.py
        return torch.nn.Sequential(torch.nn.Linear(1, 50), torch.nn.LeakyReLU(), torch.nn.Linear(50, 50),
                                   torch.nn.LeakyReLU(), plb.MC_ExpansionLayer(num_MC=mc_local, input_dim=2),
                                   plb.SIVILayer(50, 1, 5)).to(dev)
    if wl.get("conv"):
        assert world == 1, "C3 couples MC samples through MC_BatchNorm2D: single GPU (replicas only)"
        chans = [wl["in_dims"][0], 96, 128, 256, 128]
        mods = [plb.MC_ExpansionLayer(num_MC=mc_local, input_dim=4)]
        for li in range(4):
            mods += [plb.BayesConv2d(chans[li], chans[li + 1], 5, stride=1, padding=2, num_MC=mc_local),
                     plb.MC_BatchNorm2D(chans[li + 1]), torch.nn.LeakyReLU(), plb.MC_MaxPool2d(2, 2)]
        mods += [plb.BayesAdaptiveInit_FlattenAndLinear(200), torch.nn.LeakyReLU(),
                 plb.BayesLinear(200, dims[-1])]
        net = torch.nn.Sequential(*mods).to(dev)
        with torch.no_grad():      # the reference's dry run (experiments/BNN.py:253), on the device
            net(torch.randn(1, *wl["in_dims"], device=dev))
        if fuse:
            # same modules and parameters; each batch-norm / LeakyReLU / max-pool triple runs as one pass
            from pytorch_probabilisticlayers_b200.bnpool import fuse_bn_act_pool
            net = fuse_bn_act_pool(net)
        return net
    mods = [plb.MC_ExpansionLayer(num_MC=mc_local, input_dim=2)]
    for li in range(len(dims) - 1):
        mods.append(plb.BayesLinear(dims[li], dims[li + 1]))
        if li + 2 < len(dims):
            mods.append(torch.nn.LeakyReLU())
    net = torch.nn.Sequential(*mods).to(dev)
    if fuse and precision in FUSABLE:
        # same layers and parameters; the Linear-LeakyReLU-Linear runs hand bf16 activations from
        # GEMM epilogue to the next TMA load instead of four fp32 passes per hidden layer
        from pytorch_probabilisticlayers_b200.fused import fuse_bayes_mlp
        net = fuse_bayes_mlp(net, batch)
    return net


FUSABLE = ("auto", "bf16", "bf16_fast", "auto_fast")
DTYPES = {"fp32": "f32", "bf16": "bf16", "auto": "bf16", "tf32": "tf32", "auto_tf32": "tf32", "bf16_fast": "bf16",
          "auto_fast": "bf16", "tf32_fast": "tf32"}


def make_inputs(torch, wl, pin=True):
    batch, dims = wl["batch"], wl["dims"]
    if wl.get("regression"):
        x, y = torch.randn(batch, 1), torch.randn(batch, 1)
    elif wl.get("conv"):
        x, y = torch.randn(batch, *wl["in_dims"]), torch.randint(0, dims[-1], (batch,))
    else:
        x, y = torch.randn(batch, dims[0]), torch.randint(0, dims[-1], (batch,))
    return (x.pin_memory(), y.pin_memory()) if pin else (x, y)


GRAD_WIRE = ["fp32"]


def make_trainer(plb, wl, net, **kw):
    kw.setdefault("grad_wire", GRAD_WIRE[0])
    from pytorch_probabilisticlayers_b200.train import ELBOTrainStep
    ns = wl["num_samples"]
    crit = plb.MC_NLL(ns) if wl.get("regression") else plb.MC_CrossEntropyLoss(ns)
    return ELBOTrainStep(net, crit, ns, lr=wl["lr"], with_accuracy=not wl.get("regression"), **kw)


def run_b200(args, wl):
    import torch
    import torch.distributed as dist

    import pytorch_probabilisticlayers_b200 as plb
    from pytorch_probabilisticlayers_b200 import functional as PF

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout at communicator creation; keep stdout = one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    lib = plb._cabi.lib()
    assert lib.pl_device_supported() == 1, "bench needs an sm_100 (B200) device"

    plb.set_precision(args.precision)
    plb.manual_seed(20261019)
    torch.manual_seed(1234)                      # identical replicas + identical batch on all ranks
    dims, mc_total, batch, ns = wl["dims"], wl["mc"], wl["batch"], wl["num_samples"]
    if wl.get("regression") or wl.get("conv"):
        assert world == 1, "this workload couples the MC samples (MC_NLL / MC_BatchNorm2D): single GPU, replicas only"
    assert mc_total % world == 0
    mc_local = mc_total // world

    net = build_net(plb, wl, mc_local, dev, args.precision, args.fuse, world)
    trainer = make_trainer(plb, wl, net)
    assert trainer.set_shard(mc_total) == mc_local

    x_host, y_host = make_inputs(torch, wl)
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = {}

    def timed(fn, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        host_ms[fn.__name__] = (time.perf_counter() - t0) * 1e3 / steps   # host time to ENQUEUE a step
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def step_resident():
        return trainer.step(x_dev, y_dev)

    # e2e: every step's inputs come from pinned host memory and its [LL, KL, ACC, loss] is read back to the host.  The
    # loader is double-buffered like any input pipeline: the H2D copy of step k+1 is issued on a copy stream before step k
    # is launched, so it runs under step k's kernels; one H2D copy per step stays inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [(torch.empty_like(x_dev), torch.empty_like(y_dev)) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    e2e_i = [0]

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            bufs[i % 2][0].copy_(x_host, non_blocking=True)
            bufs[i % 2][1].copy_(y_host, non_blocking=True)
            ready[i % 2].record(copy_stream)

    def step_e2e():
        i = e2e_i[0]
        if i == 0:
            prefetch(0)
        torch.cuda.current_stream(dev).wait_event(ready[i % 2])
        prefetch(i + 1)                         # buffer (i+1)%2 was last read by step i-1, which the .cpu() below completed
        e2e_i[0] = i + 1
        xs, ys = bufs[i % 2]
        return trainer.step(xs, ys).cpu()        # D2H read of [LL, KL, ACC, loss] every step

    # clocks are sampled from before the warm-up to the end of the timed region (steps are ~10 ms:
    # nvidia-smi needs the warm-up to spin up)
    clocks = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        stats = step_resident()
    torch.cuda.synchronize()

    def roofline_leg():
        """Per-call CUDA events over a few steps (same stream)."""
        n = max(2, min(args.steps, 5))
        barrier()
        PF.profile_start()
        for _ in range(n):
            trainer._eager_step(x_dev, y_dev) if trainer._graph is not None else step_resident()
        return PF.profile_stop(), n

    # The step is replayed from a CUDA graph by default (--no-cuda-graph for the eager path): the regression configs and
    # C2 are bound by launch overhead, the N>1 peer step is one stream of this library's kernels (no NCCL inside), and
    # for the GPU-bound C4 / C3 the replay removes the launch gap after every end-to-end step's host read-back.
    # (An NCCL-exchange multi-GPU step is not captured.)
    can_graph = world == 1 or trainer.exchange == "peer"
    use_graph = (args.cuda_graph if args.cuda_graph is not None else True) and can_graph
    prof = None
    graph_error = None
    launches_per_step = None
    eager_us = None
    if use_graph:
        assert world == 1 or trainer.exchange == "peer", "--cuda-graph at N > 1 needs the peer exchange (no NCCL in the step)"
        # a replayed graph hides the individual launches: take the roofline leg (and the launch count) first
        prof, prof_steps = roofline_leg()
        l1 = lib.pl_launch_count()
        step_resident()
        launches_per_step = lib.pl_launch_count() - l1      # the graph replays exactly these launches
        eager_us = timed(step_resident, args.steps) / args.steps * 1e3
        try:
            trainer.enable_cuda_graph(x_dev, y_dev)
        except Exception as e:            # a net the capture cannot handle: keep the eager step, say so
            if world > 1:
                raise                     # never continue with ranks in different modes
            use_graph = False
            graph_error = f"{type(e).__name__}: {e}"[:200]
    torch.cuda.synchronize()

    # ---- headline: device-resident inputs ----
    l0 = lib.pl_launch_count()
    ms_total = timed(step_resident, args.steps)
    launches = (lib.pl_launch_count() - l0) if launches_per_step is None else launches_per_step * args.steps
    clock_info = clocks.stop() if clocks else None
    ms_step = ms_total / args.steps
    value = mc_total * batch / (ms_step * 1e-3)

    # ---- e2e: host inputs, loss read back ----
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps) / args.steps
    e2e_value = mc_total * batch / (ms_e2e * 1e-3)

    # ---- roofline leg AFTER the timed region: the kernels are timed in the same sustained (power-capped) state as
    # the step itself, which is what the sustained-peak denominator describes ----
    if prof is None:
        prof, prof_steps = roofline_leg()
    final_stats = stats.tolist()

    # ---- N > 1 only, report-only: the same step with the per-GPU share held at the full MC (weak scaling:
    # total MC = MC x N); the headline above is the strong-scaling figure BASELINE.json's config names ----
    weak = None
    exchange = getattr(trainer, "exchange", None)
    transport = getattr(trainer.shard, "transport", None) if trainer.shard is not None else None
    # 0 = every device-side barrier of the run completed; 1 = one timed out (a peer never arrived: the numbers are void)
    peer_status = trainer.shard.status() if (trainer.shard is not None and world > 1) else None
    assert not peer_status, "a cross-GPU barrier timed out"
    if world > 1 and not wl.get("conv") and not args.no_weak:
        if trainer.shard is not None:
            torch.cuda.synchronize()
            dist.barrier()
            trainer.shard.close()
        del trainer, net
        torch.cuda.empty_cache()
        net_w = build_net(plb, wl, mc_total, dev, args.precision, args.fuse, world)
        trainer_w = make_trainer(plb, wl, net_w)
        assert trainer_w.set_shard(mc_total * world) == mc_total
        if use_graph and trainer_w.exchange == "peer":
            trainer_w.enable_cuda_graph(x_dev, y_dev)

        def step_weak():
            return trainer_w.step(x_dev, y_dev)
        for _ in range(3):
            step_weak()
        ms_w = timed(step_weak, args.steps) / args.steps
        weak = {"mc_total": mc_total * world, "mc_per_gpu": mc_total, "ms_per_step": ms_w,
                "value": mc_total * world * batch / (ms_w * 1e-3), "unit": UNIT,
                "note": "report-only: per-GPU MC fixed at the config's MC, total MC grows with N"}
        net = net_w            # same layer shapes: the KL micro-benchmark below only needs the parameters

    # ---- N > 1, peer exchange: the same step with the OTHER gradient wire format, for the record ----
    other_wire = None
    if world > 1 and exchange == "peer" and not wl.get("conv"):
        try:
            ow = "fp32" if args.grad_wire == "bf16" else "bf16"
            torch.manual_seed(1234)
            net_o = build_net(plb, wl, mc_local, dev, args.precision, args.fuse, world)
            GRAD_WIRE[0] = ow
            trainer_o = make_trainer(plb, wl, net_o)
            GRAD_WIRE[0] = args.grad_wire
            assert trainer_o.set_shard(mc_total) == mc_local

            def step_other():
                return trainer_o.step(x_dev, y_dev)
            for _ in range(3):
                step_other()
            if use_graph:
                trainer_o.enable_cuda_graph(x_dev, y_dev)
            ms_o = timed(step_other, args.steps) / args.steps
            other_wire = {"grad_wire": ow, "ms_per_step": ms_o, "value": mc_total * batch / (ms_o * 1e-3), "unit": UNIT}
            torch.cuda.synchronize()
            dist.barrier()
            trainer_o.shard.close()
            del trainer_o, net_o
        except Exception as e:
            other_wire = {"error": f"{type(e).__name__}: {e}"[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()

    def roofline_of(prof, prof_steps, tf32=False, precision=None):
        precision = precision or args.precision
        per_call = {k: v[1] / v[0] for k, v in prof.items()}                  # ms per call
        per_step = {k: v[1] / prof_steps for k, v in prof.items()}            # ms per step
        top = max(per_step, key=per_step.get)
        kind, amount = PF._WORK.get(top, ("flops", 0.0))
        if kind == "flops":
            achieved = amount / (per_call[top] * 1e-3) / 1e12
            r = {"kernel": top, "bound": "tensor", "achieved": achieved, "peak": pk["tf_sustained"],
                 "unit": "TFLOP/s", "frac": achieved / pk["tf_sustained"],
                 "peak_burst": pk["tf_burst"], "frac_burst": achieved / pk["tf_burst"],
                 "traffic": None if tf32 else NCU_TRAFFIC_BYTES.get(top),
                 "peak_source": pk["source"] + ": dense bf16 torch.matmul, sustained (4 s back to back) and burst (best of 10)"
                                + ("; the TF32 tensor peak is half of it" if tf32 else ""),
                 "ms_per_launch": per_call[top], "algorithmic_flops_per_launch": amount,
                 "note": "C-ABI call = fused tcgen05 kernel + its HBM-bound pre-passes"}
            # the split-operand forward / dX kernels execute the mean MMA and the noise MMA: twice the algorithmic FLOPs
            split = DTYPES[precision] in ("bf16", "tf32") and not precision.endswith("fast") \
                and not top.startswith("linear_bwd_dw") and top.startswith("linear_")
            if split and not tf32:
                # layer 1 reads an MC-broadcast input: its mean MMA is hoisted (executed once, not per sample)
                calls = prof[top][0] / prof_steps
                first = f"linear_fwd[{mc_total // world}x{batch}x{dims[0]}x{dims[1]}]"
                mult = (2 * calls - 1) / calls if (top == first and not wl.get("conv")) else 2.0
                r["executed_flops_per_launch"] = mult * amount
                r["frac_executed"] = mult * achieved / pk["tf_sustained"]
                r["note"] += (f"; split operand: the tensor pipe executes the mean MMA and the noise MMA, {mult:.2f}x the "
                              "algorithmic FLOPs averaged over this tag's calls (the MC-broadcast first layer's mean MMA is "
                              "hoisted out of the per-sample launches); frac counts the algorithmic FLOPs only")
        else:
            achieved = amount / (per_call[top] * 1e-3) / 1e9
            r = {"kernel": top, "bound": "hbm", "achieved": achieved, "peak": pk["hbm"],
                 "unit": "GB/s", "frac": achieved / pk["hbm"], "frac_burst": achieved / pk["hbm"], "traffic": None,
                 "peak_source": pk["source"], "ms_per_launch": per_call[top],
                 "algorithmic_bytes_per_launch": amount}
        # every tensor-core call of the step against both peaks
        calls = {}
        for k in per_call:
            kd, am = PF._WORK.get(k, (None, 0.0))
            if kd == "flops" and am > 0:
                tf = am / (per_call[k] * 1e-3) / 1e12
                calls[k] = {"ms": round(per_call[k], 4), "tflops": round(tf, 1),
                            "frac": round(tf / pk["tf_sustained"], 3), "frac_burst": round(tf / pk["tf_burst"], 3)}
        r["tensor_calls"] = calls
        return r, per_step

    roofline, per_step = roofline_of(prof, prof_steps)

    # secondary: the KL reduction against the HBM roofline (north_star target (c)): the kernel
    # launched back to back over the two 4096x4096 layers' (mu, rho) in turn (2 x 134 MB > L2, so
    # every launch streams from HBM), CUDA events around the batch; once on the live parameters
    # (fresh-init sigma ~1e-5: the small-sigma path) and once on copies with sigma ~ 0.5 (general path)
    kl_info = None
    big = [m for m in net.modules() if isinstance(m, plb.BayesLinear)]
    big = [m for m in big if m.weight.loc.numel() == max(b.weight.loc.numel() for b in big)]
    if big and not wl.get("regression"):
        reps = 100
        out2 = torch.empty(2, device=dev)
        ws = torch.zeros(int(lib.pl_kl_workspace_bytes()), dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        n_ = big[0].weight.loc.numel()

        def kl_leg(ptrs):
            def kl_call(i):          # straight through the C ABI: no Python-side autograd overhead
                mu_p, rho_p, n_el = ptrs[i % len(ptrs)]
                rc = lib.pl_kl_fwd(mu_p, rho_p, n_el, 1.0, None, out2.data_ptr(), ws.data_ptr(),
                                   ws.numel(), stream)
                assert rc == 0
            for i in range(20):
                kl_call(i)
            torch.cuda.synchronize()
            batches = []
            for _ in range(5):       # 5 batches of `reps` back-to-back launches: median (and best) batch
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for i in range(reps):
                    kl_call(i)
                b.record()
                torch.cuda.synchronize()
                batches.append(a.elapsed_time(b) * 1e3 / reps)
            us = statistics.median(batches)
            kl_best.append(min(batches))
            return us, 8.0 * n_ / (us * 1e-6) / 1e9
        kl_best = []
        us, gbs = kl_leg([(m.weight.loc.data_ptr(), m.weight.logscale.data_ptr(), n_) for m in big])
        wide = [(m.weight.loc.detach().clone(), torch.empty_like(m.weight.logscale).uniform_(-1.0, 0.5)) for m in big]
        us_w, gbs_w = kl_leg([(a_.data_ptr(), b_.data_ptr(), n_) for a_, b_ in wide])
        del wide
        kl_info = {"kernel": f"kl_fwd_kernel[{n_}]", "us_per_launch": us, "achieved_gbs": gbs,
                   "frac_of_hbm_peak": gbs / pk["hbm"], "algorithmic_bytes_per_launch": 8 * n_,
                   "best_batch_us_per_launch": kl_best[0], "best_batch_frac": 8.0 * n_ / (kl_best[0] * 1e-6) / 1e9 / pk["hbm"],
                   "general_sigma_path": {"us_per_launch": us_w, "achieved_gbs": gbs_w, "frac_of_hbm_peak": gbs_w / pk["hbm"],
                                          "note": "rho ~ U(-1, 0.5), sigma 0.3-1: the general (log + softplus) path"},
                   "note": f"median of 5 batches of {reps} back-to-back single-launch reductions alternating over "
                           f"{len(big)} layers (working set {len(big) * 8 * n_ / 1e6:.0f} MB > L2); first figure: "
                           f"live parameters (fresh-init sigma ~1e-5, small-sigma path)"}

    flops = wl.get("flops") or step_flops(dims, mc_total, batch)

    # ---- secondary legs (single GPU): other precision modes of the same step, the reference eagerly on this GPU,
    # the reference on the host cores ----
    def side_leg(precision, steps):
        """ms/step + roofline of the same workload under another precision mode (own net / trainer)."""
        try:
            plb.set_precision(precision)
            torch.manual_seed(1234)
            net2 = build_net(plb, wl, mc_total, dev, precision, args.fuse)
            tr2 = make_trainer(plb, wl, net2)
            tr2.set_shard(mc_total)
            for _ in range(3):
                tr2.step(x_dev, y_dev)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(steps):
                st2 = tr2.step(x_dev, y_dev)
            b.record()
            torch.cuda.synchronize()
            ms2 = a.elapsed_time(b) / steps
            PF.profile_start()
            for _ in range(2):
                tr2.step(x_dev, y_dev)
            r2, _ = roofline_of(PF.profile_stop(), 2, tf32=precision.startswith("tf32") or precision == "auto_tf32",
                                precision=precision)
            out = {"precision": precision, "ms_per_step": ms2, "value": mc_total * batch / (ms2 * 1e-3), "unit": UNIT,
                   "steps": steps, "step_tflops": flops / (ms2 * 1e-3) / 1e12, "roofline": r2,
                   "final_stats": [float(v) for v in st2.tolist()]}
            del tr2, net2
            torch.cuda.empty_cache()
            return out
        except Exception as e:  # a secondary leg must never take the headline line down
            return {"precision": precision, "error": f"{type(e).__name__}: {e}"[:300]}
        finally:
            plb.set_precision(args.precision)

    side = {}
    eager_ref = None
    cpu = None
    if world == 1 and not args.no_side_legs:
        n_side = max(3, min(args.steps, 10))
        if not wl.get("regression") and not wl.get("conv"):
            for prec in ("auto_tf32", "auto_fast"):
                if prec != args.precision:
                    side[prec] = side_leg(prec, n_side)
        # the reference itself, eagerly on this GPU: the bar a PyTorch user starts from
        del trainer
        torch.cuda.empty_cache()
        eager_ref = {}
        for name, tf in (("fp32", False), ("allow_tf32", True)):
            mc_try = mc_total
            while mc_try >= 1:
                try:
                    v, ms, kind, _ = reference_throughput(wl, mc_try, max(3, min(args.steps, 10)), 3, device=str(dev),
                                                          allow_tf32=tf)
                    eager_ref[name] = {"value": v, "unit": UNIT, "ms_per_step": ms * mc_total / mc_try, "mc": mc_try,
                                       "ms_per_step_at_mc": ms, "kind": kind}
                    break
                except torch.cuda.OutOfMemoryError:
                    torch.cuda.empty_cache()
                    mc_try //= 2
                except Exception as e:
                    eager_ref[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
                    break
            torch.cuda.empty_cache()
        torch.backends.cuda.matmul.allow_tf32 = False
        eager_ref["note"] = ("UNMODIFIED reference modules (oracle/_ref) under oracle/ref_harness.py, device = this GPU, "
                             "eager PyTorch (cuBLAS bmm + materialised eps / W of shape [MC,I,O]); ms_per_step is scaled "
                             "to the full MC when a smaller MC had to be used (memory)")
        if not args.no_cpu_baseline:
            mc_sample = min(wl["mc"], wl["cpu_mc"])
            try:
                v, ms, kind, cores = reference_throughput(wl, mc_sample, 3, 1)
                cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "ms_per_step_sample": ms,
                       "sample": f"MC={mc_sample} of {wl['mc']}, full batch {wl['batch']}, 3 timed steps "
                                 f"after 1 warm-up, torch CPU fp32, {cores} threads"}
            except Exception as e:
                cpu = {"error": f"{type(e).__name__}: {e}"[:300]}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None,
        "dtype": DTYPES[args.precision],
        "data": "synthetic",
        "config": workload_config(wl),
        "impl_config": {"precision": args.precision, "mc_per_gpu": mc_local,
                        "fused_chain": bool(args.fuse and not wl.get("conv") and not wl.get("regression")
                                            and args.precision in FUSABLE),
                        "cuda_graph": bool(use_graph), "cuda_graph_error": graph_error,
                        "parallelism": f"mc-shard x{world}" if world > 1 else "single",
                        "grad_exchange": exchange, "grad_wire": args.grad_wire if exchange == "peer" else None,
                        "peer_transport": transport if world > 1 else None, "peer_barrier_status": peer_status},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": x_host.numel() * x_host.element_size() + y_host.numel() * y_host.element_size(),
                "d2h_bytes_per_step": 16,
                "loader": "double-buffered: the pinned-host -> device copy of step k+1 is issued on a copy stream when "
                          "step k is launched; the result of every step is read back before the next one starts"},
        "gpu_launches": int(launches),
        "us_per_step": ms_step * 1e3,
        "eager_us_per_step": eager_us,
        "host_enqueue_ms_per_step": round(host_ms.get("step_resident", 0.0), 3),
        "clocks": clock_info,
        "roofline": roofline,
        "kl_roofline": kl_info,
        "step_tflops": flops / (ms_step * 1e-3) / 1e12 * (1.0 / world),
        "kernel_ms_per_step": {k: round(v, 4) for k, v in sorted(per_step.items(), key=lambda kv: -kv[1])},
        "tf32": side.get("auto_tf32"),
        "bf16_fast": side.get("auto_fast"),
        "gpu_eager_reference": eager_ref,
        "cpu_baseline": cpu,
        "weak_scaling": weak,
        "other_wire": other_wire,
        "final_stats": {"LL": final_stats[0], "KL": final_stats[1], "ACC": final_stats[2],
                        "loss": final_stats[3]},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="c4")
    ap.add_argument("--precision", choices=sorted(DTYPES), default="auto",
                    help="auto = bf16 tcgen05 kernels (noise-preserving split operand) wherever the layer shape tiles, "
                         "fp32 kernels for tiny layers; *_fast = single rounded operand (opt-in)")
    ap.add_argument("--mc", type=int, default=0,
                    help="developer option: override the workload's MC (e.g. 8 = the per-GPU share at N=8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side-legs", action="store_true",
                    help="skip the secondary legs (tf32 / bf16_fast / gpu_eager_reference / cpu_baseline)")
    ap.add_argument("--no-weak", action="store_true", help="N>1: skip the report-only weak-scaling leg")
    ap.add_argument("--grad-wire", choices=["auto", "fp32", "bf16"], default="auto",
                    help="N>1, peer exchange: format of the partial gradients on the wire.  bf16 halves the reduce-scatter "
                         "bytes (every partial rounded once to bf16, summed in fp32 by the owner; gated by "
                         "tests/test_gpu_multirank.py); auto = bf16 when the layers compute with bf16 operands, else fp32.  "
                         "The line also carries the other format's timing (`other_wire`).")
    ap.add_argument("--cuda-graph", dest="cuda_graph", action="store_true", default=None,
                    help="replay the whole step from a CUDA graph (default for the launch-bound c1 / c5)")
    ap.add_argument("--no-cuda-graph", dest="cuda_graph", action="store_false")
    ap.add_argument("--no-fuse", dest="fuse", action="store_false",
                    help="run BayesLinear / LeakyReLU module by module instead of as a fused chain")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.grad_wire == "auto":
        args.grad_wire = "bf16" if DTYPES[args.precision] == "bf16" else "fp32"
    GRAD_WIRE[0] = args.grad_wire
    if args.mc:
        wl["mc"] = args.mc
        wl["name"] = wl.get("name", args.workload) + f" [MC overridden to {args.mc}]"
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_b200(args, wl)


if __name__ == "__main__":
    main()
