"""Manual GPU probe: a few C4-sized fused forward/backward calls of one BayesLinear (for ncu captures)."""
import sys
import torch
sys.path.insert(0, ".")
import pytorch_probabilisticlayers_b200 as p

mc, b, i, o = 64, 512, 4096, 4096
p.manual_seed(7)
layer = p.BayesLinear(i, o).cuda()
layer.precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
x = torch.randn(mc, b, i, device="cuda", requires_grad=True)
for _ in range(3):
    y = layer(x)
    y.square().mean().backward()
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
