"""train.ELBOTrainStep on the regression configurations (C1: MC_NLL over a BayesLinear MLP, C5: SIVILayer last layer) against
the reference's golden trajectories, their CUDA-graph replays, and every set_precision value on a conv net.

Reference: experiments/RegressionBNN.py:109-120, experiments/SIVILastLayer.py:287-296 (training_step), goldens from
tests/golden/make_golden.py (train_c1_regression, train_c5_sivi)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def plb():
    import pytorch_probabilisticlayers_b200 as p
    assert torch.cuda.is_available()
    p.set_precision("fp32")
    p.set_eps_source("philox")
    return p


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_elbo_train_step_mc_nll_matches_c1_golden(plb):
    """ELBOTrainStep (fused Adam + KL pass) with MC_NLL on the C1 net, the reference's eps injected step by step."""
    from pytorch_probabilisticlayers_b200.train import ELBOTrainStep
    g = load_golden("train_c1_regression")
    dims = [int(d) for d in g["dims"]]
    mc, ns, steps = int(g["mc"]), int(g["num_samples"]), int(g["steps"])
    mods = [plb.MC_ExpansionLayer(num_MC=mc, input_dim=2)]
    for li in range(len(dims) - 1):
        mods.append(plb.BayesLinear(dims[li], dims[li + 1]))
        if li + 2 < len(dims):
            mods.append(torch.nn.ReLU())
    net = torch.nn.Sequential(*mods).cuda()
    layers = [m for m in net if isinstance(m, plb.BayesLinear)]
    with torch.no_grad():
        for li, l in enumerate(layers):
            l.weight.loc.copy_(dev(g[f"init_w_mu{li}"]))
            l.weight.logscale.copy_(dev(g[f"init_w_rho{li}"]))
            l.bias.loc.copy_(dev(g[f"init_b_mu{li}"]))
            l.bias.logscale.copy_(dev(g[f"init_b_rho{li}"]))
    tr = ELBOTrainStep(net, plb.MC_NLL(ns), ns, lr=float(g["lr"]), with_accuracy=False)
    x, y = dev(g["x"]), dev(g["y"])
    for s in range(steps):
        for li, l in enumerate(layers):
            l.inject_eps(dev(g[f"s{s}_eps_w{li}"]), dev(g[f"s{s}_eps_b{li}"]))
        st = tr.step(x, y).tolist()
        assert abs(st[0] - float(g[f"s{s}_LL"])) / abs(float(g[f"s{s}_LL"])) < 2e-4, s
        assert abs(st[1] - float(g[f"s{s}_KL"].item())) / abs(float(g[f"s{s}_KL"].item())) < 1e-5, s
    for li, l in enumerate(layers):
        assert rel_err(l.weight.loc.detach().cpu().numpy(), g[f"final_w_mu{li}"]) < 1e-3
        assert rel_err(l.weight.logscale.detach().cpu().numpy(), g[f"final_w_rho{li}"]) < 1e-3


def _c5_net(plb, g, mc):
    net = torch.nn.Sequential(torch.nn.Linear(1, 50), torch.nn.LeakyReLU(), torch.nn.Linear(50, 50), torch.nn.LeakyReLU(),
                              plb.MC_ExpansionLayer(num_MC=mc, input_dim=2), plb.SIVILayer(50, 1, 5)).cuda()
    if g is not None:
        net[0].load_state_dict({k[5:]: dev(v) for k, v in g.items() if k.startswith("fc1__")})
        net[2].load_state_dict({k[5:]: dev(v) for k, v in g.items() if k.startswith("fc2__")})
        net[5].load_state_dict({k[6:].replace("__", "."): dev(v) for k, v in g.items() if k.startswith("sivi__")})
    return net


def test_elbo_train_step_sivi_matches_c5_golden(plb):
    """ELBOTrainStep with the SIVI last layer (hyper-net KL through autograd, plain-Adam tensors in the fused pass)."""
    from pytorch_probabilisticlayers_b200.train import ELBOTrainStep
    g = load_golden("train_c5_sivi")
    mc, ns, steps = int(g["mc"]), int(g["num_samples"]), int(g["steps"])
    net = _c5_net(plb, g, mc)
    tr = ELBOTrainStep(net, plb.MC_NLL(ns), ns, lr=1e-2, with_accuracy=False)
    x, y = dev(g["x"]), dev(g["y"])
    for s in range(steps):
        net[5].inject_noise(dev(g[f"s{s}_noise"]), dev(g[f"s{s}_eps_w"]), dev(g[f"s{s}_eps_b"]))
        st = tr.step(x, y).tolist()
        assert abs(st[0] - float(g[f"s{s}_LL"])) / abs(float(g[f"s{s}_LL"])) < 5e-4, s
        assert abs(st[1] - float(g[f"s{s}_KL"])) / abs(float(g[f"s{s}_KL"])) < 1e-4, s
    assert rel_err(net[0].weight.detach().cpu().numpy(), g["final_fc1_weight"]) < 2e-3
    assert rel_err(net[5].sivi_net[4].bias.detach().cpu().numpy(), g["final_sivi_last_bias"]) < 2e-3


@pytest.mark.parametrize("config", ["c1", "c5"])
def test_cuda_graph_replay_of_regression_configs(plb, config):
    """The launch-bound configs replayed from a CUDA graph (SIVILayer's hyper-net noise is drawn on the device with the
    device-side seed offset) follow the eager trajectory."""
    from pytorch_probabilisticlayers_b200.train import ELBOTrainStep
    results = []
    for graph in (False, True):
        torch.manual_seed(11)
        plb.manual_seed(123)
        if config == "c1":
            mc, b, ns = 10, 50, 1000
            dims = [1, 49, 51, 52, 1]
            mods = [plb.MC_ExpansionLayer(num_MC=mc, input_dim=2)]
            for li in range(4):
                mods.append(plb.BayesLinear(dims[li], dims[li + 1]))
                if li < 3:
                    mods.append(torch.nn.ReLU())
            net = torch.nn.Sequential(*mods).cuda()
        else:
            mc, b, ns = 128, 100, 1000
            net = _c5_net(plb, None, mc)
        k = 0
        for m in net.modules():
            if isinstance(m, (plb.BayesLinear, plb.SIVILayer)):
                m.layer_id = 3000 + k                               # same eps streams in both runs
                k += 1
        tr = ELBOTrainStep(net, plb.MC_NLL(ns), ns, lr=1e-2, with_accuracy=False)
        x = torch.linspace(-1, 1, b, device="cuda").reshape(b, 1)
        y = torch.sin(3 * x) + 0.1 * torch.randn(b, 1, device="cuda")
        if graph:                       # two eager steps on the default stream first (as bench.py does), then capture
            for _ in range(2):
                tr.step(x, y)
            tr.enable_cuda_graph(x, y, warmup=1)
            stats = [tr.step(x, y).tolist() for _ in range(5)]
        else:
            stats = [tr.step(x, y).tolist() for _ in range(8)][3:]
        torch.cuda.synchronize()
        results.append(stats)
    s0, s1 = results
    assert len({tuple(s) for s in s1}) == 5                          # fresh draws every replay
    for a, c in zip(s0, s1):
        assert abs(a[3] - c[3]) <= 2e-3 * max(abs(a[3]), 1e-6), (a, c)


@pytest.mark.parametrize("precision", ["fp32", "tf32", "bf16", "auto", "auto_tf32", "bf16_fast", "tf32_fast", "auto_fast"])
def test_conv_and_narrow_layers_run_under_every_precision(plb, precision):
    """Every set_precision value runs a conv net with a narrow head forward + backward (conv has no tf32 kernel and takes
    the fp32 path; forced tensor-core modes fall back to fp32 for layers that do not tile) and agrees with fp32."""
    outs = {}
    for prec in ("fp32", precision):
        plb.set_precision(prec)
        try:
            torch.manual_seed(2)
            plb.manual_seed(77)
            mc, b = 2, 8
            # smooth activations and no pooling: a LeakyReLU / max-pool kink next to zero turns a 1e-3 forward difference
            # into an O(1) difference of single gradient elements, which says nothing about the kernels
            net = torch.nn.Sequential(plb.MC_ExpansionLayer(mc, 4), plb.BayesConv2d(8, 32, 3, padding=1, num_MC=mc),
                                      torch.nn.Tanh(), plb.BayesAdaptiveInit_FlattenAndLinear(136), torch.nn.Tanh(),
                                      plb.BayesLinear(136, 10)).cuda()
            x = torch.randn(b, 8, 16, 16, device="cuda")
            with torch.no_grad():
                net(x)                                               # adaptive init creates the hidden BayesLinear
            k = 0
            for m in net.modules():
                if isinstance(m, (plb.BayesLinear, plb.BayesConv2d)):
                    m.layer_id = 4000 + k                            # same eps streams in both runs
                    k += 1
                    m._calls = 0
                    with torch.no_grad():
                        m.weight.logscale.fill_(-4.0)
            torch.manual_seed(3)
            lin = [m for m in net.modules() if isinstance(m, plb.BayesLinear)]
            with torch.no_grad():
                for l in lin:
                    l.weight.loc.copy_(torch.randn_like(l.weight.loc) * 0.05)
            y = net(x)
            y.square().sum().backward()
            outs[prec] = (y.detach().cpu().numpy(), net[1].weight.loc.grad.cpu().numpy())
        finally:
            plb.set_precision("fp32")
    tol = 2e-2 if "bf16" in precision or precision in ("auto", "auto_fast") else 2e-3
    assert rel_err(outs[precision][0], outs["fp32"][0]) < tol
    # the conv weight gradient sees the error of three stacked layers and of d(y^2)/dy = 2y: 5x the per-layer tolerance
    assert rel_err(outs[precision][1], outs["fp32"][1]) < 5 * tol
