"""Manual GPU probe: per-layer gradients of the small conv net of tests/test_gpu_train_modes.py under two precisions."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import pytorch_probabilisticlayers_b200 as plb

def run(prec):
    plb.set_precision(prec)
    torch.manual_seed(2); plb.manual_seed(77)
    mc, b = 2, 8
    net = torch.nn.Sequential(plb.MC_ExpansionLayer(mc, 4), plb.BayesConv2d(8, 32, 3, padding=1, num_MC=mc), torch.nn.LeakyReLU(),
                              plb.MC_MaxPool2d(2, 2), plb.BayesAdaptiveInit_FlattenAndLinear(136), torch.nn.LeakyReLU(),
                              plb.BayesLinear(136, 10)).cuda()
    x = torch.randn(b, 8, 16, 16, device="cuda")
    with torch.no_grad():
        net(x)
    k = 0
    for m in net.modules():
        if isinstance(m, (plb.BayesLinear, plb.BayesConv2d)):
            m.layer_id = 4000 + k; k += 1; m._calls = 0
            with torch.no_grad():
                m.weight.logscale.fill_(-4.0)
    torch.manual_seed(3)
    lin = [m for m in net.modules() if isinstance(m, plb.BayesLinear)]
    with torch.no_grad():
        for l in lin:
            l.weight.loc.copy_(torch.randn_like(l.weight.loc) * 0.05)
    acts = {}
    def keep(n):
        def hook(mod, i, o):
            if o.requires_grad:
                o.retain_grad()
            acts[n] = o
        return hook
    hooks = [m.register_forward_hook(keep(n)) for n, m in enumerate(net)]
    y = net(x)
    y.square().sum().backward()
    out = {f"act{n}": a.detach().cpu().numpy() for n, a in acts.items()}
    out.update({f"gact{n}": a.grad.cpu().numpy() for n, a in acts.items() if a.grad is not None})
    out["conv_dw"] = net[1].weight.loc.grad.cpu().numpy()
    out["precs"] = [getattr(m, "_last", [0, 0, 0, None])[3].precision for m in net.modules() if isinstance(m, (plb.BayesLinear, plb.BayesConv2d))]
    return out

ref = run("fp32")
for prec in sys.argv[1:]:
    got = run(prec)
    print(prec, "precisions", got["precs"])
    for k in sorted(ref):
        if k == "precs":
            continue
        a, b = got[k].astype(np.float64), ref[k].astype(np.float64)
        print(f"  {k:8s} rel_err(max) {np.abs(a-b).max()/max(np.abs(b).max(),1e-30):.3e}  rel_l2 {np.linalg.norm(a-b)/max(np.linalg.norm(b),1e-30):.3e}")
