"""Manual GPU probe: forward time of one C4-sized BayesLinear for per-MC and MC-broadcast inputs in the split / fast modes."""
import sys
import torch
sys.path.insert(0, ".")
import pytorch_probabilisticlayers_b200 as p

mc, b, i, o = 64, 512, 4096, 4096
p.manual_seed(7)
layer = p.BayesLinear(i, o).cuda()
xm = torch.randn(mc, b, i, device="cuda")
xb = torch.randn(b, i, device="cuda").unsqueeze(0).expand(mc, b, i)
for prec in sys.argv[1:] or ["bf16", "bf16_fast"]:
    layer.precision = prec
    for name, x in (("per-mc", xm), ("broadcast", xb)):
        with torch.no_grad():
            for _ in range(3):
                layer(x)
            torch.cuda.synchronize()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                layer(x)
            e.record()
            torch.cuda.synchronize()
        print(f"{prec:10s} {name:10s} {a.elapsed_time(e) / 10:.3f} ms per forward call")
