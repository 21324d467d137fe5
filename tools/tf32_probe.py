"""Manual GPU probe (run from the repo root: python tools/...): TF32 tensor-core path vs the fp32 path."""
import sys
import torch
sys.path.insert(0, ".")
import pytorch_probabilisticlayers_b200 as p

torch.manual_seed(0)
p.manual_seed(5)
mc, b, i, o = 1, 128, 32, 128
layer = p.BayesLinear(i, o, bias=False).cuda()
with torch.no_grad():
    layer.weight.logscale.fill_(-20.0)       # W == mu: isolates the operand plumbing
x = torch.randn(mc, b, i, device="cuda")
outs = {}
for prec in ("fp32", "tf32"):
    layer.precision = prec
    layer._calls = 0
    with torch.no_grad():
        outs[prec] = layer(x)[0]
ref, got = outs["fp32"], outs["tf32"]
print("max rel err", float((got - ref).abs().max() / ref.abs().max()))
# which reference row does each output row resemble?
for r in (0, 1, 2, 3, 8, 9, 64, 127):
    d = (ref - got[r][None, :]).abs().sum(1)
    print("out row", r, "closest ref row", int(d.argmin()), "dist", float(d.min()), "self dist", float(d[r]))
# column structure
for c in (0, 1, 4, 32, 33, 127):
    d = (ref - got[:, c][:, None]).abs().sum(0)
    print("out col", c, "closest ref col", int(d.argmin()), "dist", float(d.min()), "self dist", float(d[c]))

print("---- backward ----")
res = {}
for prec in ("fp32", "tf32"):
    layer.precision = prec
    layer._calls = 0
    layer.zero_grad()
    xx = x.clone().requires_grad_(True)
    y = layer(xx)
    gy = torch.ones_like(y) * torch.arange(o, device="cuda").float()[None, None, :] / o
    (y * gy).sum().backward()
    res[prec] = (xx.grad.clone(), layer.weight.loc.grad.clone())
for k, n in enumerate(("dx", "dw_mu")):
    a, r = res["tf32"][k], res["fp32"][k]
    print(n, "rel err", float((a - r).abs().max() / r.abs().max()), "tf32 absmax", float(a.abs().max()), "ref absmax", float(r.abs().max()))
