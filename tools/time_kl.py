"""Manual GPU probe: pl_kl_fwd on two 4096x4096 (mu, rho) pairs in turn (268 MB working set > L2), us per launch."""
import sys
import torch
sys.path.insert(0, ".")
import pytorch_probabilisticlayers_b200 as p
lib = p._cabi.lib()
dev = torch.device("cuda")
n = 4096 * 4096
pairs = [(torch.randn(n, device=dev) * 0.01, torch.full((n,), -12.0, device=dev)) for _ in range(2)]
wide = [(torch.randn(n, device=dev) * 0.01, torch.empty(n, device=dev).uniform_(-1, 0.5)) for _ in range(2)]
out2 = torch.empty(2, device=dev)
ws = torch.zeros(int(lib.pl_kl_workspace_bytes()), dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
for name, ps in (("fresh-init sigma", pairs), ("sigma ~0.5", wide)):
    best = 1e9
    for trial in range(5):
        for i in range(8):
            lib.pl_kl_fwd(ps[i % 2][0].data_ptr(), ps[i % 2][1].data_ptr(), n, 1.0, None, out2.data_ptr(), ws.data_ptr(), ws.numel(), st)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(100):
            lib.pl_kl_fwd(ps[i % 2][0].data_ptr(), ps[i % 2][1].data_ptr(), n, 1.0, None, out2.data_ptr(), ws.data_ptr(), ws.numel(), st)
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 10)
    print(f"{name:18s} {best:.2f} us per launch = {8 * n / best / 1e3:.0f} GB/s = {8 * n / best / 1e3 / 6553:.3f} of 6553")
