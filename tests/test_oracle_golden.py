"""Pin the CPU oracle (oracle/) against outputs of the unmodified reference (tests/golden)."""
import numpy as np
import pytest

from conftest import load_golden, rel_err
from oracle import closed_form as cf
from oracle import philox

TOL = 2e-5        # reference is fp32 on CPU; the oracle computes in fp64


@pytest.mark.parametrize("name", ["linear_small", "linear_nobias", "linear_odd", "linear_expand",
                                  "linear_tile"])
def test_linear_matches_reference(name):
    g = load_golden(name)
    bias = "b_mu" in g
    out = cf.linear_fwd(g["x"], g["w_mu"], g["w_rho"], g["eps_w"],
                        g.get("b_mu"), g.get("b_rho"), g.get("eps_b"))
    assert rel_err(out, g["out"]) < TOL
    assert rel_err(cf.rsample(g["w_mu"], g["w_rho"], g["eps_w"]), g["sampled_w"]) < TOL
    b = cf.linear_bwd(g["gy"], g["x"], g["w_mu"], g["w_rho"], g["eps_w"],
                      g.get("b_rho"), g.get("eps_b"))
    assert rel_err(b["dx"], g["dx"]) < TOL
    assert rel_err(b["dw_mu"], g["dw_mu"]) < TOL
    assert rel_err(b["dw_rho"], g["dw_rho"]) < TOL
    if bias:
        assert rel_err(b["db_mu"], g["db_mu"]) < TOL
        assert rel_err(b["db_rho"], g["db_rho"]) < TOL
    kl, ent = cf.kl_fwd(g["w_mu"], g["w_rho"], float(g["prior"]))
    assert abs(kl - g["kl"]) / abs(g["kl"]) < 1e-5
    assert abs(ent - g["entropy"]) / abs(g["entropy"]) < 1e-5
    dkm, dkr = cf.kl_bwd(g["w_mu"], g["w_rho"], float(g["prior"]))
    assert rel_err(dkm, g["dkl_mu"]) < TOL
    assert rel_err(dkr, g["dkl_rho"]) < TOL
    assert rel_err(cf.entropy_bwd(g["w_rho"]), g["dent_rho"]) < TOL


@pytest.mark.parametrize("name", ["conv_small", "conv_strided", "conv_k5"])
def test_conv_matches_reference(name):
    g = load_golden(name)
    k, s, p, d = (int(v) for v in g["hyper"])
    out = cf.conv2d_fwd(g["x"], g["w_mu"], g["w_rho"], g["eps_w"], g["b_mu"], g["b_rho"],
                        g["eps_b"], s, p, d)
    assert out.shape == g["out"].shape
    assert rel_err(out, g["out"]) < TOL
    b = cf.conv2d_bwd(g["gy"], g["x"], g["w_mu"], g["w_rho"], g["eps_w"], g["b_rho"], g["eps_b"],
                      s, p, d)
    for key in ("dx", "dw_mu", "dw_rho", "db_mu", "db_rho"):
        assert rel_err(b[key], g[key]) < TOL, key
    kl, _ = cf.kl_fwd(g["w_mu"], g["w_rho"], float(g["prior"]))
    assert abs(kl - g["kl"]) / abs(g["kl"]) < 1e-5


def test_conv_output_is_permuted_view_in_reference():
    # SURVEY 8c pin 4: the reference returns [MC,B,C,H,W] as a permuted view of [B,MC,C,H,W]
    g = load_golden("conv_small")
    mc, b, c, h, w = g["out"].shape
    assert tuple(g["out_strides"]) == (c * h * w, mc * c * h * w, h * w, w, 1)


def test_losses_match_reference():
    g = load_golden("losses")
    loss, dpred = cf.mc_cross_entropy(g["ce_pred"], g["ce_target"], 1000)
    assert abs(loss - g["ce_loss"]) / abs(g["ce_loss"]) < 1e-5
    assert rel_err(dpred, g["ce_dpred"]) < TOL
    assert abs(cf.mc_accuracy(g["ce_pred"], g["ce_target"]) - float(g["ce_acc"][0])) < 1e-6
    loss, dpred = cf.mc_nll(g["nll_pred"], g["nll_target"], 500)
    assert abs(loss - g["nll_loss"]) / abs(g["nll_loss"]) < 1e-5
    assert rel_err(dpred, g["nll_dpred"]) < 5e-5


def test_sivi_matches_reference():
    g = load_golden("sivi_small")
    out = cf.sivi_linear_fwd(g["x"], g["w_mu"], g["w_std"], g["eps_w"], g["b_mu"], g["b_std"],
                             g["eps_b"])
    assert rel_err(out, g["out"]) < TOL
    b = cf.sivi_linear_bwd(g["gy"], g["x"], g["w_mu"], g["w_std"], g["eps_w"], g["eps_b"])
    for key in ("dx", "dw_mu", "dw_std", "db_mu", "db_std"):
        assert rel_err(b[key], g[key]) < TOL, key
    assert abs(cf.sivi_kl(g["w_mu"], g["w_std"]) - g["kl"]) / abs(g["kl"]) < 1e-5


@pytest.mark.parametrize("name,slope,loss", [("train_c1_regression", 0.0, "nll"),
                                             ("train_mlp_ce", 0.01, "ce")])
def test_elbo_step_matches_reference(name, slope, loss):
    """First optimiser step of the reference train loop: pred, LL, KL, loss and all grads."""
    g = load_golden(name)
    nl = len(g["dims"]) - 1
    params = [dict(w_mu=g[f"init_w_mu{i}"], w_rho=g[f"init_w_rho{i}"], b_mu=g[f"init_b_mu{i}"],
                   b_rho=g[f"init_b_rho{i}"]) for i in range(nl)]
    eps = [dict(w=g[f"s0_eps_w{i}"], b=g[f"s0_eps_b{i}"]) for i in range(nl)]
    r = cf.mlp_elbo_step(g["x"], g["y"], params, eps, int(g["num_samples"]), 1.0, slope, loss)
    assert rel_err(r["pred"], g["s0_pred"]) < TOL
    assert abs(r["LL"] - g["s0_LL"]) / abs(g["s0_LL"]) < 2e-5
    assert abs(r["KL"] - g["s0_KL"].item()) / abs(g["s0_KL"].item()) < 1e-5
    for i in range(nl):
        for k in ("w_mu", "w_rho", "b_mu", "b_rho"):
            assert rel_err(r["grads"][i][k], g[f"s0_g_{k}{i}"]) < 2e-4, (i, k)


def test_init_statistics():
    # SURVEY 8c pin 6: BayesLinear(784,1200) fresh init
    g = load_golden("init_stats")
    sigma_w = cf.softplus(g["lin_w_rho0"])
    sigma_b = cf.softplus(g["lin_b_rho0"])
    assert abs(sigma_w - np.sqrt(2 / (784 + 1200)) / 1200) / sigma_w < 1e-3
    assert abs(sigma_b - np.sqrt(2 / (784 + 1200))) / sigma_b < 1e-4
    assert g["lin_w_mu_absmax"] <= 1 / np.sqrt(1200) + 1e-7
    assert abs(g["lin_kl"] - 9.4455e6) / 9.4455e6 < 1e-2
    assert abs(g["lin_entropy"] + 8.5808e6) / 8.5808e6 < 1e-2


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
         (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for ctr, key, want in kat:
        got = philox.philox4x32_10(*ctr, *key)
        assert tuple(int(x) for x in got) == want


def test_philox_normals_are_standard_and_tiling_invariant():
    e = philox.eps_normal(1234, 0, 0, 4, 128, 130)
    assert abs(e.mean()) < 0.02 and abs(e.std() - 1) < 0.02
    assert abs((e ** 4).mean() - 3.0) < 0.15
    # column/row/MC sub-blocks are the same numbers (counter is positional)
    sub = philox.eps_normal(1234, 0, 2, 2, 128, 64)
    assert np.array_equal(sub, e[2:4, :, :64])
    # different stream / seed decorrelates
    other = philox.eps_normal(1234, 1, 0, 4, 128, 130)
    assert abs(np.corrcoef(e.ravel(), other.ravel())[0, 1]) < 0.02


# ----------------------------------------------------------------------------------------------
# torch_port (the CPU baseline that bench.py times) must be the reference, op for op
# ----------------------------------------------------------------------------------------------
def test_torch_port_layer_reproduces_reference_draws_and_outputs():
    import torch
    from oracle import torch_port as tp
    g = load_golden("linear_small")           # generated with torch.manual_seed(11 + 1) before forward
    layer = tp.BayesLinearPort(5, 4, prior=0.7)
    with torch.no_grad():
        layer.weight.loc.copy_(torch.from_numpy(g["w_mu"]))
        layer.weight.logscale.copy_(torch.from_numpy(g["w_rho"]))
        layer.bias.loc.copy_(torch.from_numpy(g["b_mu"]))
        layer.bias.logscale.copy_(torch.from_numpy(g["b_rho"]))
    torch.manual_seed(12)
    out = layer(torch.from_numpy(g["x"]))
    assert np.array_equal(layer.weight.eps.numpy(), g["eps_w"])
    assert np.array_equal(layer.bias.eps.numpy(), g["eps_b"])
    assert np.allclose(out.detach().numpy(), g["out"], rtol=1e-6, atol=1e-6)
    assert abs(layer.kl_div.item() - g["kl"]) / abs(g["kl"]) < 1e-6

    gc = load_golden("conv_k5")
    conv = tp.BayesConv2dPort(3, 6, 5, stride=1, padding=2, dilation=1, num_MC=2, prior_scale=1.3)
    with torch.no_grad():
        conv.weight.loc.copy_(torch.from_numpy(gc["w_mu"]))
        conv.weight.logscale.copy_(torch.from_numpy(gc["w_rho"]))
        conv.bias.loc.copy_(torch.from_numpy(gc["b_mu"]))
        conv.bias.logscale.copy_(torch.from_numpy(gc["b_rho"]))
    torch.manual_seed(24)
    out = conv(torch.from_numpy(gc["x"]))
    assert np.allclose(out.detach().numpy(), gc["out"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name,act,loss", [("train_c1_regression", "relu", "nll"),
                                           ("train_mlp_ce", "leaky", "ce")])
def test_torch_port_train_step_reproduces_reference_trajectory(name, act, loss):
    import torch
    from oracle import torch_port as tp
    g = load_golden(name)
    seed = 31 if loss == "nll" else 32
    dims = [int(d) for d in g["dims"]]
    mc, ns = int(g["mc"]), int(g["num_samples"])
    net = tp.build_mlp(dims, mc, torch.nn.ReLU if act == "relu" else torch.nn.LeakyReLU)
    layers = [m for m in net if isinstance(m, tp.BayesLinearPort)]
    with torch.no_grad():
        for li, l in enumerate(layers):
            l.weight.loc.copy_(torch.from_numpy(g[f"init_w_mu{li}"]))
            l.weight.logscale.copy_(torch.from_numpy(g[f"init_w_rho{li}"]))
            l.bias.loc.copy_(torch.from_numpy(g[f"init_b_mu{li}"]))
            l.bias.logscale.copy_(torch.from_numpy(g[f"init_b_rho{li}"]))
    crit = tp.MCNLLPort(ns) if loss == "nll" else tp.MCCrossEntropyPort(ns)
    opt = torch.optim.Adam(net.parameters(), float(g["lr"]))
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    for s in range(int(g["steps"])):
        torch.manual_seed(seed + 100 + s)
        r = tp.train_step(net, crit, opt, x, y, ns, with_accuracy=(loss == "ce"))
        assert abs(r["loss"] - float(g[f"s{s}_loss"])) / abs(float(g[f"s{s}_loss"])) < 1e-5, s
    for li, l in enumerate(layers):
        assert np.allclose(l.weight.loc.detach().numpy(), g[f"final_w_mu{li}"], rtol=1e-4, atol=1e-6)


def test_torch_port_convnet_reproduces_reference_trajectory():
    """The 'cbnn' block structure at toy size: conv -> MC batch norm -> LeakyReLU -> pool (x2) -> flatten/linear ->
    linear, three Adam steps; KL from the BayesLinear layers only (experiments/BNN.py:269-276)."""
    import torch
    from oracle import torch_port as tp
    g = load_golden("train_c3_convnet")
    chans, hw, mc, ns = [int(c) for c in g["chans"]], int(g["hw"]), int(g["mc"]), int(g["num_samples"])
    net = tp.build_convnet(mc, (chans[0], hw, hw), int(g["hidden"]), int(g["classes"]), chans=chans)
    layers = [m for m in net if isinstance(m, tp.BayesConv2dPort)] + [net[-3].linear, net[-1]]
    with torch.no_grad():
        for li, l in enumerate(layers):
            l.weight.loc.copy_(torch.from_numpy(g[f"init_w_mu{li}"]))
            l.weight.logscale.copy_(torch.from_numpy(g[f"init_w_rho{li}"]))
            l.bias.loc.copy_(torch.from_numpy(g[f"init_b_mu{li}"]))
            l.bias.logscale.copy_(torch.from_numpy(g[f"init_b_rho{li}"]))
    crit = tp.MCCrossEntropyPort(ns)
    opt = torch.optim.Adam(net.parameters(), float(g["lr"]))
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    for s in range(int(g["steps"])):
        torch.manual_seed(71 + 100 + s)
        r = tp.train_step(net, crit, opt, x, y, ns)
        assert abs(r["loss"] - float(g[f"s{s}_loss"])) / abs(float(g[f"s{s}_loss"])) < 1e-5, s
    for li, l in enumerate(layers):
        assert np.allclose(l.weight.loc.detach().numpy(), g[f"final_w_mu{li}"], rtol=1e-4, atol=1e-6)
    bns = [m for m in net if isinstance(m, tp.MCBatchNorm2DPort)]
    for bi, bn in enumerate(bns):     # (running statistics also carry the reference's random dry-run batch: not compared)
        assert np.allclose(bn.weight.detach().numpy(), g[f"final_bn_w{bi}"], rtol=1e-5, atol=1e-7)
        assert np.allclose(bn.bias.detach().numpy(), g[f"final_bn_b{bi}"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("name,clamp", [("laplace_prior", False), ("laplace_prior_clamp", True)])
def test_laplace_prior_kl_and_ema_restatement(name, clamp):
    """LaplacePrior (src/ProbabilisticLayers.py:131-159): per-element prior scale sqrt(EMA)+1e-8 (or clamped at
    1), EMA of 1/grad^2 of weight.loc with bias correction, KL against it -- the oracle's closed forms against
    the reference's own values."""
    g = load_golden(name)
    ema = np.ones_like(g["w_mu"], dtype=np.float64)
    for s in range(int(g["steps"])):
        ps = np.maximum(np.sqrt(ema), 1.0) if clamp else np.sqrt(ema) + 1e-8
        assert rel_err(ps, g[f"s{s}_prior_scale_in"]) < 1e-6
        kl, _ = cf.kl_fwd(g["w_mu"], g["w_rho"], ps)
        assert abs(kl - float(g[f"s{s}_kl"])) / abs(float(g[f"s{s}_kl"])) < 1e-5
        # total gradient of weight.loc = data term + 0.01 * dKL/dmu, which is what the tensor hook sees
        data = cf.linear_bwd(g["gy"], g["x"], g["w_mu"], g["w_rho"], g[f"s{s}_eps_w"], g["b_rho"], g[f"s{s}_eps_b"])
        dmu_kl, drho_kl = cf.kl_bwd(g["w_mu"], g["w_rho"], ps)
        gmu = data["dw_mu"] + 0.01 * dmu_kl
        assert rel_err(gmu, g[f"s{s}_g_w_mu"]) < 1e-4
        assert rel_err(data["dw_rho"] + 0.01 * drho_kl, g[f"s{s}_g_w_rho"]) < 1e-4
        corr = 1.0 - 0.99 ** (s + 1)
        ema = 0.99 * ema + 0.01 * (1.0 / g[f"s{s}_g_w_mu"].astype(np.float64) ** 2) / (corr + 1e-8)
        assert rel_err(ema, g[f"s{s}_ema"]) < 1e-4


def test_eps_stream_statistics_oracle():
    """Statistical gate of the 8-normals-per-call sampler contract on the CPU restatement (2.1e6 draws); the GPU test
    runs the same checks on 1.7e7 draws of pl_eps_dump."""
    from sampler_stats import check_normal_stream
    from oracle import philox
    check_normal_stream(philox.eps_normal(0xC0FFEE1234, 6, 3, 4, 512, 1024), "oracle")
