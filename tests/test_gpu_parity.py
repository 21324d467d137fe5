"""Parity of the CUDA path (through the C ABI) against the reference's golden vectors and the CPU
oracle.  Everything here needs a B200: run with ``pytest -m gpu`` under gpurun.

Tolerances: fp32 path rel 2e-5 (it is plain fp32 FFMA arithmetic, well
inside the 2e-3 bar); bf16 tensor-core path rel 2e-2; KL sums rel 1e-5.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import closed_form as cf
from oracle import philox

pytestmark = pytest.mark.gpu

FP32_TOL = 2e-5
KL_TOL = 1e-5


@pytest.fixture(scope="module")
def plb():
    import pytorch_probabilisticlayers_b200 as p
    assert torch.cuda.is_available(), "these tests need a GPU"
    assert p._cabi.lib().pl_device_supported() == 1, "not an sm_100 device"
    p.set_precision("fp32")
    p.set_eps_source("philox")
    return p


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def load_linear(p, g, bias=True):
    i, o = g["w_mu"].shape
    layer = p.BayesLinear(i, o, bias=bias, prior=float(g["prior"])).cuda()
    with torch.no_grad():
        layer.weight.loc.copy_(dev(g["w_mu"]))
        layer.weight.logscale.copy_(dev(g["w_rho"]))
        if bias:
            layer.bias.loc.copy_(dev(g["b_mu"]))
            layer.bias.logscale.copy_(dev(g["b_rho"]))
    return layer


# ---------------------------------------------------------------------------------------------
# eps stream
# ---------------------------------------------------------------------------------------------
def test_eps_stream_matches_oracle(plb):
    from pytorch_probabilisticlayers_b200.functional import eps_dump
    seed, stream, mc0, mc, rows, cols = 0x123456789ABCDEF, 7, 5, 3, 37, 50
    raw = eps_dump(seed, stream, mc0, mc, rows, cols, "cuda", raw=True).cpu().numpy().view(np.uint32)
    want = philox.raw_words(seed, stream, mc0, mc, rows, cols)
    assert np.array_equal(raw, want)                      # integer stream: bit exact
    eps = eps_dump(seed, stream, mc0, mc, rows, cols, "cuda").cpu().numpy()
    ref = philox.eps_normal(seed, stream, mc0, mc, rows, cols)
    assert np.abs(eps - ref).max() < 5e-5                 # MUFU lg2/sin/cos approximations
    big = eps_dump(99, 0, 0, 8, 512, 512, "cuda")
    assert abs(big.mean().item()) < 5e-3 and abs(big.std().item() - 1) < 5e-3


def test_eps_stream_statistics(plb):
    """Statistical gate of the sampler on 1.7e7 draws of the on-chip stream: moments, KS vs N(0,1),
    chi-square on |z| bins to the tail, correlation within a Box-Muller pair, across the 8 normals of one Philox call,
    across column groups, rows and MC samples (tests/sampler_stats.py)."""
    from sampler_stats import check_normal_stream
    from pytorch_probabilisticlayers_b200.functional import eps_dump
    eps = eps_dump(0xC0FFEE1234, 6, 3, 8, 1024, 2048, "cuda").cpu().numpy()
    check_normal_stream(eps, "pl_eps_dump")


# ---------------------------------------------------------------------------------------------
# KL / entropy
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["linear_small", "linear_odd", "linear_tile", "conv_k5"])
def test_kl_matches_reference(plb, name):
    from pytorch_probabilisticlayers_b200.functional import kl_and_entropy
    g = load_golden(name)
    mu = dev(g["w_mu"]).requires_grad_(True)
    rho = dev(g["w_rho"]).requires_grad_(True)
    kl, ent = kl_and_entropy(mu, rho, float(g["prior"]))
    assert abs(kl.item() - g["kl"]) / abs(g["kl"]) < KL_TOL
    if "entropy" in g:
        assert abs(ent.item() - g["entropy"]) / abs(g["entropy"]) < KL_TOL
    kl.backward()
    assert rel_err(mu.grad.cpu().numpy(), g["dkl_mu"]) < FP32_TOL
    assert rel_err(rho.grad.cpu().numpy(), g["dkl_rho"]) < FP32_TOL
    if "dent_rho" in g:
        mu.grad = rho.grad = None
        kl, ent = kl_and_entropy(mu, rho, float(g["prior"]))
        ent.backward()
        assert rel_err(rho.grad.cpu().numpy(), g["dent_rho"]) < FP32_TOL
        assert float(mu.grad.abs().max()) == 0.0


@pytest.mark.parametrize("n", [1, 3, 1025, 1_000_003, 16_777_216])
def test_kl_large_and_ragged_vs_fp64_oracle(plb, n):
    from pytorch_probabilisticlayers_b200.functional import kl_and_entropy
    rs = np.random.RandomState(n % 1000)
    mu = (rs.uniform(-1, 1, n) / 64).astype(np.float32)
    rho = rs.uniform(-13.0, 2.0, n).astype(np.float32)       # sigma from 2e-6 to 2.1
    want_kl, want_ent = cf.kl_fwd(mu, rho, 0.5)
    tmu, trho = dev(mu).requires_grad_(True), dev(rho).requires_grad_(True)
    kl, ent = kl_and_entropy(tmu, trho, 0.5)
    assert abs(kl.item() - want_kl) / abs(want_kl) < KL_TOL
    assert abs(ent.item() - want_ent) / abs(want_ent) < KL_TOL
    (2.0 * kl + 3.0 * ent).backward()
    dm, dr = cf.kl_bwd(mu, rho, 0.5)
    dr = 2.0 * dr + 3.0 * cf.entropy_bwd(rho)
    assert rel_err(tmu.grad.cpu().numpy(), 2.0 * dm) < FP32_TOL
    assert rel_err(trho.grad.cpu().numpy(), dr) < FP32_TOL
    # unaligned views take the scalar path and must agree
    if n > 8:
        kl2, _ = kl_and_entropy(tmu.detach()[1:], trho.detach()[1:], 0.5)
        want2, _ = cf.kl_fwd(mu[1:], rho[1:], 0.5)
        assert abs(kl2.item() - want2) / abs(want2) < KL_TOL


def test_kl_per_element_prior(plb):
    from pytorch_probabilisticlayers_b200.functional import kl_and_entropy
    rs = np.random.RandomState(3)
    n = 4099
    mu = rs.normal(size=n).astype(np.float32) * 0.1
    rho = rs.uniform(-6, 1, n).astype(np.float32)
    ps = rs.uniform(0.5, 2.0, n).astype(np.float32)
    want, _ = cf.kl_fwd(mu, rho, ps)
    tmu, trho = dev(mu).requires_grad_(True), dev(rho).requires_grad_(True)
    kl, _ = kl_and_entropy(tmu, trho, 1.0, dev(ps))
    assert abs(kl.item() - want) / abs(want) < KL_TOL
    kl.backward()
    dm, dr = cf.kl_bwd(mu, rho, ps)
    assert rel_err(tmu.grad.cpu().numpy(), dm) < FP32_TOL
    assert rel_err(trho.grad.cpu().numpy(), dr) < FP32_TOL


# ---------------------------------------------------------------------------------------------
# BayesLinear, injected eps = the reference's own torch.randn draws
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["linear_small", "linear_nobias", "linear_odd", "linear_expand",
                                  "linear_tile"])
def test_linear_injected_eps_matches_reference(plb, name):
    g = load_golden(name)
    bias = "b_mu" in g
    layer = load_linear(plb, g, bias)
    layer.inject_eps(dev(g["eps_w"]), dev(g["eps_b"]) if bias else None)
    x = dev(g["x"]).requires_grad_(True)
    out = layer(x)
    assert rel_err(out.detach().cpu().numpy(), g["out"]) < FP32_TOL
    assert abs(layer.kl_div.item() - g["kl"]) / abs(g["kl"]) < KL_TOL
    assert abs(layer.entropy.item() - g["entropy"]) / abs(g["entropy"]) < KL_TOL
    (out * dev(g["gy"])).sum().backward()
    assert rel_err(x.grad.cpu().numpy(), g["dx"]) < FP32_TOL
    assert rel_err(layer.weight.loc.grad.cpu().numpy(), g["dw_mu"]) < FP32_TOL
    assert rel_err(layer.weight.logscale.grad.cpu().numpy(), g["dw_rho"]) < FP32_TOL
    if bias:
        assert rel_err(layer.bias.loc.grad.cpu().numpy(), g["db_mu"]) < FP32_TOL
        assert rel_err(layer.bias.logscale.grad.cpu().numpy(), g["db_rho"]) < FP32_TOL
    # lazy debug views reproduce the reference's eager attributes
    assert rel_err(layer.sampled_w.detach().cpu().numpy(), g["sampled_w"]) < FP32_TOL


def test_linear_kl_gradient_adds_to_likelihood_gradient(plb):
    """loss = sum(out*gy) + c*kl_div: autograd must add the two paths like the reference."""
    g = load_golden("linear_small")
    layer = load_linear(plb, g)
    layer.inject_eps(dev(g["eps_w"]), dev(g["eps_b"]))
    out = layer(dev(g["x"]))
    ((out * dev(g["gy"])).sum() + 0.25 * layer.kl_div).backward()
    assert rel_err(layer.weight.loc.grad.cpu().numpy(), g["dw_mu"] + 0.25 * g["dkl_mu"]) < FP32_TOL
    assert rel_err(layer.weight.logscale.grad.cpu().numpy(), g["dw_rho"] + 0.25 * g["dkl_rho"]) < FP32_TOL
    assert layer.bias.loc.grad is not None


def test_linear_philox_mode_matches_oracle_through_eps_bridge(plb):
    """Production path: eps synthesised on chip.  Dump the same stream, feed it to the oracle."""
    plb.manual_seed(2024)
    torch.manual_seed(0)
    mc, b, i, o = 6, 19, 45, 38
    layer = plb.BayesLinear(i, o).cuda()
    with torch.no_grad():
        layer.weight.logscale.uniform_(-4, 0)
        layer.bias.logscale.uniform_(-4, 0)
    x = torch.randn(mc, b, i, device="cuda", requires_grad=True)
    gy = torch.randn(mc, b, o, device="cuda")
    out = layer(x)
    (out * gy).sum().backward()
    eps_w = layer.weight.eps.cpu().numpy()
    eps_b = layer.bias.eps.cpu().numpy()
    assert eps_w.shape == (mc, i, o) and eps_b.shape == (mc, o)
    assert abs(eps_w.std() - 1) < 0.05
    n = lambda t: t.detach().cpu().numpy()
    want = cf.linear_fwd(n(x), n(layer.weight.loc), n(layer.weight.logscale), eps_w,
                         n(layer.bias.loc), n(layer.bias.logscale), eps_b)
    assert rel_err(n(out), want) < FP32_TOL
    bw = cf.linear_bwd(n(gy), n(x), n(layer.weight.loc), n(layer.weight.logscale), eps_w,
                       n(layer.bias.logscale), eps_b)
    assert rel_err(n(x.grad), bw["dx"]) < FP32_TOL
    assert rel_err(n(layer.weight.loc.grad), bw["dw_mu"]) < FP32_TOL
    assert rel_err(n(layer.weight.logscale.grad), bw["dw_rho"]) < FP32_TOL
    assert rel_err(n(layer.bias.loc.grad), bw["db_mu"]) < FP32_TOL
    assert rel_err(n(layer.bias.logscale.grad), bw["db_rho"]) < FP32_TOL
    # a second forward draws fresh noise
    out2 = layer(x.detach())
    assert (out2 - out.detach()).abs().max().item() > 1e-4


def test_linear_mc_shards_reproduce_the_unsharded_result(plb):
    """Rank r owning global MC indices [r*MC/N, (r+1)*MC/N) sees exactly the unsharded samples."""
    plb.manual_seed(77)
    torch.manual_seed(1)
    mc, b, i, o = 8, 16, 24, 20
    x = torch.randn(mc, b, i, device="cuda")
    gy = torch.randn(mc, b, o, device="cuda")
    base = plb.BayesLinear(i, o).cuda()
    with torch.no_grad():
        base.weight.logscale.uniform_(-3, 0)
    sd = base.state_dict()

    def run(lo, hi):
        layer = plb.BayesLinear(i, o).cuda()
        layer.load_state_dict(sd)
        layer.layer_id = base.layer_id
        plb.set_mc_shard(lo)
        out = layer(x[lo:hi])
        (out * gy[lo:hi]).sum().backward()
        plb.set_mc_shard(0)
        return out.detach(), [p.grad.clone() for p in layer.parameters()]

    full_out, full_g = run(0, mc)
    parts = [run(0, 4), run(4, 8)]
    assert torch.equal(torch.cat([p[0] for p in parts]), full_out)
    for k, gfull in enumerate(full_g):
        gsum = parts[0][1][k] + parts[1][1][k]
        assert rel_err(gsum.cpu().numpy(), gfull.cpu().numpy()) < 1e-5


def test_expansion_layer_stride0_equals_materialised_repeat(plb):
    plb.manual_seed(5)
    torch.manual_seed(2)
    layer = plb.BayesLinear(12, 9).cuda()
    x = torch.randn(7, 12, device="cuda")
    xe = plb.MC_ExpansionLayer(num_MC=5, input_dim=2)(x)
    assert xe.shape == (5, 7, 12) and xe.stride(0) == 0
    layer._calls = 0
    a = layer(xe)
    layer._calls = 0
    b = layer(xe.contiguous())
    assert torch.equal(a, b)


# ---------------------------------------------------------------------------------------------
# BayesConv2d
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["conv_small", "conv_strided", "conv_k5"])
def test_conv_injected_eps_matches_reference(plb, name):
    g = load_golden(name)
    k, s, pad, d = (int(v) for v in g["hyper"])
    mc = g["x"].shape[0]
    cout, cin = g["w_mu"].shape[:2]
    layer = plb.BayesConv2d(cin, cout, k, stride=s, padding=pad, dilation=d, num_MC=mc,
                            prior_scale=float(g["prior"])).cuda()
    with torch.no_grad():
        layer.weight.loc.copy_(dev(g["w_mu"]))
        layer.weight.logscale.copy_(dev(g["w_rho"]))
        layer.bias.loc.copy_(dev(g["b_mu"]))
        layer.bias.logscale.copy_(dev(g["b_rho"]))
    layer.inject_eps(dev(g["eps_w"]), dev(g["eps_b"]))
    x = dev(g["x"]).requires_grad_(True)
    out = layer(x)
    assert out.shape == g["out"].shape
    assert rel_err(out.detach().cpu().numpy(), g["out"]) < FP32_TOL
    assert abs(layer.kl_div.item() - g["kl"]) / abs(g["kl"]) < KL_TOL
    (out * dev(g["gy"])).sum().backward()
    assert rel_err(x.grad.cpu().numpy(), g["dx"]) < FP32_TOL
    assert rel_err(layer.weight.loc.grad.cpu().numpy(), g["dw_mu"]) < FP32_TOL
    assert rel_err(layer.weight.logscale.grad.cpu().numpy(), g["dw_rho"]) < FP32_TOL
    assert rel_err(layer.bias.loc.grad.cpu().numpy(), g["db_mu"]) < FP32_TOL
    assert rel_err(layer.bias.logscale.grad.cpu().numpy(), g["db_rho"]) < FP32_TOL


def test_conv_philox_mode_matches_oracle_through_eps_bridge(plb):
    plb.manual_seed(31)
    torch.manual_seed(3)
    mc, b, cin, cout, h, w, k = 3, 2, 5, 7, 9, 10, 3
    layer = plb.BayesConv2d(cin, cout, k, stride=1, padding=1, num_MC=mc).cuda()
    with torch.no_grad():
        layer.weight.logscale.uniform_(-4, 0)
        layer.bias.logscale.uniform_(-4, 0)
    x = torch.randn(mc, b, cin, h, w, device="cuda", requires_grad=True)
    out = layer(x)
    gy = torch.randn_like(out)
    (out * gy).sum().backward()
    n = lambda t: t.detach().cpu().numpy()
    eps_w, eps_b = n(layer.weight.eps), n(layer.bias.eps)
    want = cf.conv2d_fwd(n(x), n(layer.weight.loc), n(layer.weight.logscale), eps_w,
                         n(layer.bias.loc), n(layer.bias.logscale), eps_b, 1, 1, 1)
    assert rel_err(n(out), want) < FP32_TOL
    bw = cf.conv2d_bwd(n(gy), n(x), n(layer.weight.loc), n(layer.weight.logscale), eps_w,
                       n(layer.bias.logscale), eps_b, 1, 1, 1)
    assert rel_err(n(x.grad), bw["dx"]) < FP32_TOL
    assert rel_err(n(layer.weight.loc.grad), bw["dw_mu"]) < FP32_TOL
    assert rel_err(n(layer.weight.logscale.grad), bw["dw_rho"]) < FP32_TOL
    assert rel_err(n(layer.bias.logscale.grad), bw["db_rho"]) < FP32_TOL


# ---------------------------------------------------------------------------------------------
# SIVI
# ---------------------------------------------------------------------------------------------
def test_sivi_matches_reference(plb):
    g = load_golden("sivi_small")
    mc, b, i = g["x"].shape
    layer = plb.SIVILayer(i, 1, 5).cuda()
    sd = {k[4:].replace("__", "."): dev(v) for k, v in g.items() if k.startswith("sd__")}
    layer.load_state_dict(sd)
    layer.inject_noise(dev(g["noise"]), dev(g["eps_w"]), dev(g["eps_b"]))
    x = dev(g["x"]).requires_grad_(True)
    out = layer(x)
    assert rel_err(layer.w_mu.detach().cpu().numpy(), g["w_mu"]) < FP32_TOL
    assert rel_err(layer.w_std.detach().cpu().numpy(), g["w_std"]) < FP32_TOL
    assert rel_err(out.detach().cpu().numpy(), g["out"]) < FP32_TOL
    assert abs(layer.kl_div.item() - g["kl"]) / abs(g["kl"]) < KL_TOL
    layer.w_mu.retain_grad()
    layer.w_std.retain_grad()
    layer.b_std.retain_grad()
    (out * dev(g["gy"])).sum().backward()
    assert rel_err(x.grad.cpu().numpy(), g["dx"]) < FP32_TOL
    assert rel_err(layer.w_mu.grad.cpu().numpy(), g["dw_mu"]) < FP32_TOL
    assert rel_err(layer.w_std.grad.cpu().numpy(), g["dw_std"]) < FP32_TOL
    assert rel_err(layer.b_std.grad.cpu().numpy(), g["db_std"]) < FP32_TOL


# ---------------------------------------------------------------------------------------------
# layers around the path + losses
# ---------------------------------------------------------------------------------------------
def test_aux_layers_and_losses_match_reference(plb):
    g = load_golden("aux_layers")
    x5 = dev(g["x5"])
    bn2 = plb.MC_BatchNorm2D(6).cuda()
    assert rel_err(bn2(x5).cpu().detach().numpy(), g["bn2"]) < 1e-4
    assert rel_err(bn2.running_var.cpu().numpy(), g["bn2_rv"]) < 1e-5
    assert rel_err(plb.MC_MaxPool2d(2, 2)(x5).cpu().numpy(), g["maxpool"]) == 0.0
    bn1 = plb.MC_BatchNorm1D(7).cuda()
    assert rel_err(bn1(dev(g["x3"])).cpu().detach().numpy(), g["bn1"]) < 1e-4
    ex = plb.MC_ExpansionLayer(num_MC=4, input_dim=2)(dev(g["x3"][0]))
    assert np.array_equal(ex.cpu().numpy(), g["expand"])
    l = load_golden("losses")
    pred = dev(l["ce_pred"]).requires_grad_(True)
    ce = plb.MC_CrossEntropyLoss(1000)(pred, dev(l["ce_target"]))
    ce.backward()
    assert abs(ce.item() - l["ce_loss"]) / abs(l["ce_loss"]) < 1e-5
    assert rel_err(pred.grad.cpu().numpy(), l["ce_dpred"]) < 1e-5
    assert abs(plb.MC_Accuracy(pred, dev(l["ce_target"])).item() - float(l["ce_acc"][0])) < 1e-6
    pr = dev(l["nll_pred"]).requires_grad_(True)
    nll = plb.MC_NLL(500)(pr, dev(l["nll_target"]))
    nll.backward()
    assert abs(nll.item() - l["nll_loss"]) / abs(l["nll_loss"]) < 1e-5
    assert rel_err(pr.grad.cpu().numpy(), l["nll_dpred"]) < 5e-5


@pytest.mark.parametrize("mc,b,c", [(3, 5, 10), (2, 7, 7), (4, 16, 33), (2, 64, 1000), (1, 3, 4100), (8, 300, 12)])
def test_fused_cross_entropy_matches_fp64_oracle(plb, mc, b, c):
    """pl_mc_cross_entropy: loss, accuracy count and gradient from one pass, against the fp64 restatement of
    MC_CrossEntropyLoss / MC_Accuracy (src/ProbabilisticLayers.py:32-90); deterministic across launches."""
    from pytorch_probabilisticlayers_b200 import functional as PF
    rng = np.random.default_rng(mc * 100 + c)
    pred = (rng.standard_normal((mc, b, c)) * 3).astype(np.float32)
    target = rng.integers(0, c, size=b)
    pred[0, 0, :] = 0.25                       # an all-ties row: argmax must be index 0
    ns = 1234
    want_loss, want_d = cf.mc_cross_entropy(pred, target, ns)
    want_hits = float((pred.argmax(-1) == target[None, :]).sum())
    scale = ns / (mc * b)
    out, d = PF.mc_cross_entropy_raw(dev(pred), torch.from_numpy(target).cuda(), scale)
    out2, d2 = PF.mc_cross_entropy_raw(dev(pred), torch.from_numpy(target).cuda(), scale)
    assert torch.equal(out, out2) and torch.equal(d, d2)
    assert abs(out[0].item() * scale - want_loss) / abs(want_loss) < 1e-5
    assert out[1].item() == want_hits
    assert rel_err(d.cpu().numpy(), want_d) < 1e-5
    # module path with autograd and a non-unit upstream gradient
    p = dev(pred).requires_grad_(True)
    crit = plb.MC_CrossEntropyLoss(ns)
    (crit(p, torch.from_numpy(target).cuda()) * 0.5).backward()
    assert rel_err(p.grad.cpu().numpy(), 0.5 * want_d) < 1e-5
    assert crit.last_correct.item() == want_hits
    # value only (no gradient requested)
    out3, d3 = PF.mc_cross_entropy_raw(dev(pred), torch.from_numpy(target).cuda(), None)
    assert d3 is None and torch.equal(out3, out)


@pytest.mark.parametrize("name,prior", [("laplace_prior", "laplace"), ("laplace_prior_clamp", "laplace_clamp")])
def test_laplace_prior_matches_reference(plb, name, prior):
    """BayesLinear(prior='laplace' | 'laplace_clamp'): KL against the per-element prior (pl_kl_fwd / pl_kl_bwd with
    the prior-scale operand) and the EMA the gradient hook maintains, step by step against the reference."""
    g = load_golden(name)
    i, o = g["w_mu"].shape
    layer = plb.BayesLinear(i, o, prior=prior).cuda()
    with torch.no_grad():
        layer.weight.loc.copy_(dev(g["w_mu"]))
        layer.weight.logscale.copy_(dev(g["w_rho"]))
        layer.bias.loc.copy_(dev(g["b_mu"]))
        layer.bias.logscale.copy_(dev(g["b_rho"]))
    x, gy = dev(g["x"]), dev(g["gy"])
    for s in range(int(g["steps"])):
        layer.zero_grad()
        layer.inject_eps(dev(g[f"s{s}_eps_w"]), dev(g[f"s{s}_eps_b"]))
        out = layer(x)
        assert rel_err(out.detach().cpu().numpy(), g[f"s{s}_out"]) < 2e-5
        assert abs(layer.kl_div.item() - float(g[f"s{s}_kl"])) / abs(float(g[f"s{s}_kl"])) < 1e-5
        ((out * gy).sum() + 0.01 * layer.kl_div).backward()
        assert rel_err(layer.weight.loc.grad.cpu().numpy(), g[f"s{s}_g_w_mu"]) < 1e-4
        assert rel_err(layer.weight.logscale.grad.cpu().numpy(), g[f"s{s}_g_w_rho"]) < 1e-4
        assert rel_err(layer.prior.scale.cpu().numpy(), g[f"s{s}_ema"]) < 1e-3


# ---------------------------------------------------------------------------------------------
# end-to-end: the reference's train loop, step by step, eps injected
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,act,loss", [("train_c1_regression", torch.nn.ReLU, "nll"),
                                           ("train_mlp_ce", torch.nn.LeakyReLU, "ce")])
def test_training_trajectory_matches_reference(plb, name, act, loss):
    g = load_golden(name)
    dims = [int(d) for d in g["dims"]]
    mc, ns, steps = int(g["mc"]), int(g["num_samples"]), int(g["steps"])
    mods = [plb.MC_ExpansionLayer(num_MC=mc, input_dim=2)]
    for li in range(len(dims) - 1):
        mods.append(plb.BayesLinear(dims[li], dims[li + 1]))
        if li + 2 < len(dims):
            mods.append(act())
    net = torch.nn.Sequential(*mods).cuda()
    layers = [m for m in net if isinstance(m, plb.BayesLinear)]
    with torch.no_grad():
        for li, l in enumerate(layers):
            l.weight.loc.copy_(dev(g[f"init_w_mu{li}"]))
            l.weight.logscale.copy_(dev(g[f"init_w_rho{li}"]))
            l.bias.loc.copy_(dev(g[f"init_b_mu{li}"]))
            l.bias.logscale.copy_(dev(g[f"init_b_rho{li}"]))
    crit = plb.MC_NLL(ns) if loss == "nll" else plb.MC_CrossEntropyLoss(ns)
    opt = torch.optim.Adam(net.parameters(), float(g["lr"]))
    x, y = dev(g["x"]), dev(g["y"])
    for s in range(steps):
        for li, l in enumerate(layers):
            l.inject_eps(dev(g[f"s{s}_eps_w{li}"]), dev(g[f"s{s}_eps_b{li}"]))
        opt.zero_grad()
        pred = net(x)
        ll = crit(pred, y) / ns
        kl = sum(l.kl_div for l in layers) / ns
        total = ll + kl
        total.backward()
        assert rel_err(pred.detach().cpu().numpy(), g[f"s{s}_pred"]) < 2e-4, s
        assert abs(ll.item() - g[f"s{s}_LL"]) / abs(g[f"s{s}_LL"]) < 2e-4, s
        assert abs(kl.item() - g[f"s{s}_KL"].item()) / abs(g[f"s{s}_KL"].item()) < KL_TOL, s
        if s == 0:
            for li, l in enumerate(layers):
                assert rel_err(l.weight.loc.grad.cpu().numpy(), g[f"s0_g_w_mu{li}"]) < 2e-4
                assert rel_err(l.weight.logscale.grad.cpu().numpy(), g[f"s0_g_w_rho{li}"]) < 2e-4
                assert rel_err(l.bias.logscale.grad.cpu().numpy(), g[f"s0_g_b_rho{li}"]) < 2e-4
        opt.step()
    for li, l in enumerate(layers):
        assert rel_err(l.weight.loc.detach().cpu().numpy(), g[f"final_w_mu{li}"]) < 1e-3
        assert rel_err(l.weight.logscale.detach().cpu().numpy(), g[f"final_w_rho{li}"]) < 1e-3


@pytest.mark.parametrize("mc,b,c,h,w,train", [(3, 4, 6, 8, 12, True), (2, 3, 5, 4, 4, True), (1, 2, 96, 32, 32, True),
                                              (2, 2, 7, 6, 10, False)])
def test_fused_bn_act_pool_equals_module_chain(plb, mc, b, c, h, w, train):
    """bnpool.FusedBNActPool against MC_BatchNorm2D -> LeakyReLU -> MC_MaxPool2d(2,2) (the reference's three
    modules, src/ProbabilisticLayers.py:736-762, :207-236): outputs, gradients of x / weight / bias and the
    running statistics, in training and in eval mode."""
    from pytorch_probabilisticlayers_b200.bnpool import FusedBNActPool
    torch.manual_seed(mc * 100 + c)
    x0 = torch.randn(mc, b, c, h, w, device="cuda") * 1.7 + 0.3
    gy = torch.randn(mc, b, c, h // 2, w // 2, device="cuda")
    res = {}
    for kind in ("chain", "fused"):
        bn = plb.MC_BatchNorm2D(c).cuda()
        with torch.no_grad():
            bn.weight.copy_(torch.linspace(0.5, 1.5, c))
            bn.bias.copy_(torch.linspace(-0.3, 0.4, c))
            bn.running_mean.copy_(torch.linspace(-0.2, 0.2, c))
            bn.running_var.copy_(torch.linspace(0.8, 1.6, c))
        bn.train(train)
        x = x0.clone().requires_grad_(True)
        if kind == "chain":
            y = plb.MC_MaxPool2d(2, 2)(torch.nn.functional.leaky_relu(bn(x), 0.01))
        else:
            y = FusedBNActPool(bn, 0.01)(x)
        (y * gy).sum().backward()
        res[kind] = [t.detach().cpu().numpy() for t in (y, x.grad, bn.weight.grad, bn.bias.grad, bn.running_mean,
                                                        bn.running_var)]
    for a, r, tol in zip(res["fused"], res["chain"], (2e-6, 2e-4, 2e-4, 2e-4, 1e-5, 1e-5)):
        assert rel_err(a, r) < tol


def test_convnet_training_trajectory_matches_reference(plb, fuse_blocks=False):
    """The reference's 'cbnn' block structure (experiments/BNN.py:228-253) at toy size, three Adam steps with the
    reference's own eps injected: BayesConv2d / MC_BatchNorm2D / MC_MaxPool2d / flatten+linear end to end."""
    g = load_golden("train_c3_convnet")
    chans, hw, mc, ns = [int(c) for c in g["chans"]], int(g["hw"]), int(g["mc"]), int(g["num_samples"])
    mods = [plb.MC_ExpansionLayer(num_MC=mc, input_dim=4)]
    for ci in range(len(chans) - 1):
        mods += [plb.BayesConv2d(chans[ci], chans[ci + 1], kernel_size=5, padding=2, stride=1, num_MC=mc),
                 plb.MC_BatchNorm2D(chans[ci + 1]), torch.nn.LeakyReLU(), plb.MC_MaxPool2d(kernel_size=2, stride=2)]
    mods += [plb.BayesAdaptiveInit_FlattenAndLinear(int(g["hidden"])), torch.nn.LeakyReLU(),
             plb.BayesLinear(int(g["hidden"]), int(g["classes"]))]
    net = torch.nn.Sequential(*mods).cuda()
    with torch.no_grad():
        net(torch.randn(1, chans[0], hw, hw, device="cuda"))       # dry run creates the flatten layer
    convs = [m for m in net if isinstance(m, plb.BayesConv2d)]
    lins = [net[-3].linear, net[-1]]
    layers = convs + lins
    bns = [m for m in net if isinstance(m, plb.MC_BatchNorm2D)]
    if fuse_blocks:
        from pytorch_probabilisticlayers_b200.bnpool import FusedBNActPool, fuse_bn_act_pool
        net = fuse_bn_act_pool(net)
        assert sum(isinstance(m, FusedBNActPool) for m in net) == len(bns)
    with torch.no_grad():
        for li, l in enumerate(layers):
            l.weight.loc.copy_(dev(g[f"init_w_mu{li}"]))
            l.weight.logscale.copy_(dev(g[f"init_w_rho{li}"]))
            l.bias.loc.copy_(dev(g[f"init_b_mu{li}"]))
            l.bias.logscale.copy_(dev(g[f"init_b_rho{li}"]))
        for bn in bns:                                              # the dry run moved the running statistics
            bn.reset_running_stats()
    # the reference's dry run also updated ITS running stats once (momentum 0.1) before training
    crit = plb.MC_CrossEntropyLoss(ns)
    opt = torch.optim.Adam(net.parameters(), float(g["lr"]))
    x, y = dev(g["x"]), dev(g["y"])
    net.train()
    for s in range(int(g["steps"])):
        for li, l in enumerate(layers):
            l.inject_eps(dev(g[f"s{s}_eps_w{li}"]), dev(g[f"s{s}_eps_b{li}"]))
        opt.zero_grad()
        pred = net(x)
        total = crit(pred, y) / ns + sum(l.kl_div for l in lins) / ns
        total.backward()
        assert rel_err(pred.detach().cpu().numpy(), g[f"s{s}_pred"]) < 5e-4, s
        assert abs(total.item() - float(g[f"s{s}_loss"])) / abs(float(g[f"s{s}_loss"])) < 2e-4, s
        if s == 0:
            for li, l in enumerate(layers):
                assert rel_err(l.weight.loc.grad.cpu().numpy(), g[f"s0_g_w_mu{li}"]) < 1e-3, li
                assert rel_err(l.weight.logscale.grad.cpu().numpy(), g[f"s0_g_w_rho{li}"]) < 1e-3, li
        opt.step()
    for li, l in enumerate(layers):
        assert rel_err(l.weight.loc.detach().cpu().numpy(), g[f"final_w_mu{li}"]) < 2e-3, li
    for bi, bn in enumerate(bns):
        assert rel_err(bn.weight.detach().cpu().numpy(), g[f"final_bn_w{bi}"]) < 1e-3


def test_convnet_trajectory_with_fused_bn_act_pool_blocks(plb):
    test_convnet_training_trajectory_matches_reference(plb, fuse_blocks=True)


# ---------------------------------------------------------------------------------------------
# error behaviour
# ---------------------------------------------------------------------------------------------
def test_error_behaviour(plb):
    layer = plb.BayesLinear(4, 3).cuda()
    with pytest.raises(AssertionError):
        layer(torch.zeros(2, 4, device="cuda"))                     # x.dim() != 3, as the reference
    conv = plb.BayesConv2d(2, 3, 3, num_MC=4).cuda()
    with pytest.raises(AssertionError):
        conv(torch.zeros(2, 1, 2, 5, 5, device="cuda"))             # x.shape[0] != num_MC
    with pytest.raises(RuntimeError):
        plb.BayesLinear(4, 3)(torch.zeros(1, 2, 4))                 # CPU: no fallback
    # empty batch / empty MC shard do not crash
    out = layer(torch.zeros(3, 0, 4, device="cuda"))
    assert out.shape == (3, 0, 3)


# ---------------------------------------------------------------------------------------------
# the product's train step (ELBOTrainStep: fused Adam + KL pass) against the reference trajectory
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fused", [True, False])
def test_elbo_train_step_matches_reference_trajectory(plb, fused):
    from pytorch_probabilisticlayers_b200.train import ELBOTrainStep
    g = load_golden("train_mlp_ce")
    dims = [int(d) for d in g["dims"]]
    mc, ns, steps = int(g["mc"]), int(g["num_samples"]), int(g["steps"])
    mods = [plb.MC_ExpansionLayer(num_MC=mc, input_dim=2)]
    for li in range(len(dims) - 1):
        mods.append(plb.BayesLinear(dims[li], dims[li + 1]))
        if li + 2 < len(dims):
            mods.append(torch.nn.LeakyReLU())
    net = torch.nn.Sequential(*mods).cuda()
    layers = [m for m in net if isinstance(m, plb.BayesLinear)]
    with torch.no_grad():
        for li, l in enumerate(layers):
            l.weight.loc.copy_(dev(g[f"init_w_mu{li}"]))
            l.weight.logscale.copy_(dev(g[f"init_w_rho{li}"]))
            l.bias.loc.copy_(dev(g[f"init_b_mu{li}"]))
            l.bias.logscale.copy_(dev(g[f"init_b_rho{li}"]))
    tr = ELBOTrainStep(net, plb.MC_CrossEntropyLoss(ns), ns, lr=float(g["lr"]), fused_optimizer=fused)
    x, y = dev(g["x"]), dev(g["y"])
    for s in range(steps):
        for li, l in enumerate(layers):
            l.inject_eps(dev(g[f"s{s}_eps_w{li}"]), dev(g[f"s{s}_eps_b{li}"]))
        ll, kl, acc, loss = tr.step(x, y).tolist()
        assert abs(ll - float(g[f"s{s}_LL"])) / abs(float(g[f"s{s}_LL"])) < 2e-4, s
        assert abs(kl - g[f"s{s}_KL"].item()) / abs(g[f"s{s}_KL"].item()) < 1e-5, s
        assert abs(loss - float(g[f"s{s}_loss"])) / abs(float(g[f"s{s}_loss"])) < 2e-4, s
    for li, l in enumerate(layers):
        assert rel_err(l.weight.loc.detach().cpu().numpy(), g[f"final_w_mu{li}"]) < 1e-3
        assert rel_err(l.weight.logscale.detach().cpu().numpy(), g[f"final_w_rho{li}"]) < 1e-3
        assert rel_err(l.bias.logscale.detach().cpu().numpy(), g[f"final_b_rho{li}"]) < 1e-3


def test_sivi_last_layer_training_matches_reference_trajectory(plb):
    """Config 5 (experiments/SIVILastLayer.py): deterministic feature extractor + SIVILayer(50,1,5),
    MC=128, batch 100, MC_NLL + SIVI KL, 3 Adam steps with the reference's draws injected."""
    g = load_golden("train_c5_sivi")
    mc, ns, steps = int(g["mc"]), int(g["num_samples"]), int(g["steps"])
    fc1, fc2 = torch.nn.Linear(1, 50).cuda(), torch.nn.Linear(50, 50).cuda()
    sivi = plb.SIVILayer(50, 1, 5).cuda()
    fc1.load_state_dict({k[5:]: dev(v) for k, v in g.items() if k.startswith("fc1__")})
    fc2.load_state_dict({k[5:]: dev(v) for k, v in g.items() if k.startswith("fc2__")})
    sivi.load_state_dict({k[6:].replace("__", "."): dev(v) for k, v in g.items() if k.startswith("sivi__")})
    params = list(fc1.parameters()) + list(fc2.parameters()) + list(sivi.parameters())
    opt = torch.optim.Adam(params, 1e-2)
    crit = plb.MC_NLL(ns)
    x, y = dev(g["x"]), dev(g["y"])
    F = torch.nn.functional
    for s in range(steps):
        opt.zero_grad()
        h = F.leaky_relu(fc2(F.leaky_relu(fc1(x))))
        h = h.unsqueeze(0).expand(mc, *h.shape)
        sivi.inject_noise(dev(g[f"s{s}_noise"]), dev(g[f"s{s}_eps_w"]), dev(g[f"s{s}_eps_b"]))
        pred = sivi(h)
        ll = crit(pred, y) / ns
        kl = sivi.kl_div / ns
        (ll + kl).backward()
        assert rel_err(pred.detach().cpu().numpy(), g[f"s{s}_pred"]) < 2e-4, s
        assert abs(ll.item() - float(g[f"s{s}_LL"])) / abs(float(g[f"s{s}_LL"])) < 5e-4, s
        assert abs(kl.item() - float(g[f"s{s}_KL"])) / abs(float(g[f"s{s}_KL"])) < 1e-4, s
        opt.step()
    assert rel_err(fc1.weight.detach().cpu().numpy(), g["final_fc1_weight"]) < 2e-3
    assert rel_err(sivi.sivi_net[4].bias.detach().cpu().numpy(), g["final_sivi_last_bias"]) < 2e-3
