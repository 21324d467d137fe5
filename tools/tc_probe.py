"""Manual GPU probe (run from the repo root: python tools/...): bf16 tensor-core path vs the fp32 path on the same Philox eps."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import pytorch_probabilisticlayers_b200 as p


def run(mc, b, i, o, bias=True, expand=False, inject=False, seed=1):
    torch.manual_seed(seed)
    p.manual_seed(100 + seed)
    layer = p.BayesLinear(i, o, bias=bias).cuda()
    with torch.no_grad():
        layer.weight.logscale.uniform_(-5, -1)
        if bias:
            layer.bias.logscale.uniform_(-3, 0)
    if expand:
        x0 = torch.randn(b, i, device="cuda")
        x = x0.unsqueeze(0).expand(mc, b, i)
    else:
        x = torch.randn(mc, b, i, device="cuda")
    x = x.clone().requires_grad_(True) if not expand else x
    gy = torch.randn(mc, b, o, device="cuda")
    if inject:
        layer.inject_eps(torch.randn(mc, i, o, device="cuda"), torch.randn(mc, o, device="cuda") if bias else None)
    res = {}
    for prec in ("fp32", "bf16"):
        layer.precision = prec
        layer._calls = 0
        layer.zero_grad()
        if x.requires_grad and x.grad is not None:
            x.grad = None
        out = layer(x)
        (out * gy).sum().backward()
        torch.cuda.synchronize()
        res[prec] = (out.detach().clone(), x.grad.clone() if x.requires_grad else None,
                     layer.weight.loc.grad.clone(), layer.weight.logscale.grad.clone())
    def rel(a, b_):
        return float((a - b_).abs().max() / b_.abs().max())
    names = ["out", "dx", "dw_mu", "dw_rho"]
    errs = {n: (rel(res["bf16"][k], res["fp32"][k]) if res["fp32"][k] is not None else None)
            for k, n in enumerate(names)}
    print(f"mc={mc} b={b} i={i} o={o} bias={bias} expand={expand} inject={inject}: {errs}", flush=True)


if __name__ == "__main__":
    run(2, 128, 128, 128, inject=True)
    run(2, 128, 128, 128)
    run(1, 128, 64, 128, bias=False)
    run(3, 512, 256, 384)
    run(2, 256, 192, 256)
    run(2, 200, 192, 200)
    run(2, 300, 200, 136)
    run(1, 640, 128, 128)
    run(4, 512, 512, 256, expand=True)
    run(8, 512, 1024, 1024)
