"""bf16 tensor-core path (tcgen05/TMEM/TMA kernels) against the reference's golden vectors, the
fp64 oracle and the fp32 path.  Tolerance: rel 2e-2 (north_star's bf16 mode)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import closed_form as cf

pytestmark = pytest.mark.gpu

BF16_TOL = 2e-2
TF32_TOL = 2e-3      # north_star: rel 2e-3 at TF32 / fp32 accumulate


@pytest.fixture(scope="module")
def plb():
    import pytorch_probabilisticlayers_b200 as p
    assert torch.cuda.is_available()
    p.set_precision("fp32")
    p.set_eps_source("philox")
    return p


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_tc_injected_eps_matches_reference_golden(plb):
    """The reference's own torch.randn draws through the tcgen05 kernels (128^3 tile, MC=2)."""
    g = load_golden("linear_tile")
    layer = plb.BayesLinear(128, 128, prior=float(g["prior"])).cuda()
    layer.precision = "bf16"
    with torch.no_grad():
        layer.weight.loc.copy_(dev(g["w_mu"]))
        layer.weight.logscale.copy_(dev(g["w_rho"]))
        layer.bias.loc.copy_(dev(g["b_mu"]))
        layer.bias.logscale.copy_(dev(g["b_rho"]))
    layer.inject_eps(dev(g["eps_w"]), dev(g["eps_b"]))
    x = dev(g["x"]).requires_grad_(True)
    out = layer(x)
    assert rel_err(out.detach().cpu().numpy(), g["out"]) < BF16_TOL
    assert abs(layer.kl_div.item() - g["kl"]) / abs(g["kl"]) < 1e-5
    (out * dev(g["gy"])).sum().backward()
    assert rel_err(x.grad.cpu().numpy(), g["dx"]) < BF16_TOL
    assert rel_err(layer.weight.loc.grad.cpu().numpy(), g["dw_mu"]) < BF16_TOL
    assert rel_err(layer.weight.logscale.grad.cpu().numpy(), g["dw_rho"]) < BF16_TOL
    assert rel_err(layer.bias.loc.grad.cpu().numpy(), g["db_mu"]) < 1e-4
    assert rel_err(layer.bias.logscale.grad.cpu().numpy(), g["db_rho"]) < 1e-4


@pytest.mark.parametrize("mc,b,i,o,bias,expand", [
    (2, 128, 128, 128, True, False),      # one tile
    (1, 128, 64, 128, False, False),      # single k-block, no bias
    (3, 512, 256, 384, True, False),      # MT=4: full-batch tile, 512 TMEM columns
    (2, 256, 192, 256, True, False),      # MT=2
    (2, 200, 192, 200, True, False),      # ragged batch rows and ragged N tile
    (2, 300, 200, 136, True, False),      # K tail (200 % 64 = 8), batch tail in the dW reduction
    (1, 640, 128, 128, True, False),      # more than 512 rows: two row blocks
    (4, 512, 512, 256, True, True),       # MC-broadcast input (stride 0)
    (5, 64, 1024, 8, True, False),        # minimum width
])
def test_tc_philox_matches_fp64_oracle(plb, mc, b, i, o, bias, expand):
    """Production mode (eps synthesised on chip): dump the stream, evaluate the closed forms in fp64."""
    torch.manual_seed(mc * 1000 + b)
    plb.manual_seed(4242 + o)
    layer = plb.BayesLinear(i, o, bias=bias).cuda()
    layer.precision = "bf16"
    with torch.no_grad():
        layer.weight.logscale.uniform_(-5, -1)
        if bias:
            layer.bias.logscale.uniform_(-3, 0)
    if expand:
        x = torch.randn(b, i, device="cuda").unsqueeze(0).expand(mc, b, i)
    else:
        x = torch.randn(mc, b, i, device="cuda", requires_grad=True)
    gy = torch.randn(mc, b, o, device="cuda")
    out = layer(x)
    (out * gy).sum().backward()
    n = lambda t: t.detach().cpu().numpy()
    eps_w = n(layer.weight.eps)
    eps_b = n(layer.bias.eps) if bias else None
    want = cf.linear_fwd(n(x), n(layer.weight.loc), n(layer.weight.logscale), eps_w,
                         n(layer.bias.loc) if bias else None,
                         n(layer.bias.logscale) if bias else None, eps_b)
    assert rel_err(n(out), want) < BF16_TOL
    bw = cf.linear_bwd(n(gy), n(x), n(layer.weight.loc), n(layer.weight.logscale), eps_w,
                       n(layer.bias.logscale) if bias else None, eps_b)
    if not expand:
        assert rel_err(n(x.grad), bw["dx"]) < BF16_TOL
    assert rel_err(n(layer.weight.loc.grad), bw["dw_mu"]) < BF16_TOL
    assert rel_err(n(layer.weight.logscale.grad), bw["dw_rho"]) < BF16_TOL
    if bias:
        assert rel_err(n(layer.bias.loc.grad), bw["db_mu"]) < 1e-4
        assert rel_err(n(layer.bias.logscale.grad), bw["db_rho"]) < 1e-4


def test_tc_full_size_properties(plb):
    """C4 layer shape (MC=8 of 64, B=512, 4096x4096): size-independent properties instead of an
    oracle run -- linearity in x, MC-shard invariance and agreement with the fp32 path."""
    torch.manual_seed(9)
    plb.manual_seed(99)
    layer = plb.BayesLinear(4096, 4096).cuda()
    layer.precision = "bf16"
    x = torch.randn(8, 512, 4096, device="cuda")

    def fwd(inp, lo=0):
        layer._calls = 0
        plb.set_mc_shard(lo)
        try:
            return layer(inp)
        finally:
            plb.set_mc_shard(0)

    with torch.no_grad():
        y = fwd(x)
        # linearity: f(2x) - b = 2 (f(x) - b)
        y0 = fwd(torch.zeros_like(x))
        y2 = fwd(2 * x)
        assert rel_err((y2 - y0).cpu().numpy(), (2 * (y - y0)).cpu().numpy()) < 1e-3
        # MC shards see exactly the unsharded samples (same counter -> bit-identical)
        ya = fwd(x[:4], 0)
        yb = fwd(x[4:], 4)
        assert torch.equal(torch.cat([ya, yb]), y)
        # fp32 path on the same eps
        layer.precision = "fp32"
        yf = fwd(x[:2])
        assert rel_err(y[:2].cpu().numpy(), yf.cpu().numpy()) < BF16_TOL


def test_tc_rejects_unsupported_shapes_loudly(plb):
    """The C ABI refuses shapes the tensor-core kernels cannot tile (no silent wrong answer); the layer resolves a forced
    tensor-core precision to fp32 for such shapes up front instead of failing inside autograd."""
    from pytorch_probabilisticlayers_b200 import _cabi, functional as PF
    layer = plb.BayesLinear(30, 20).cuda()
    spec = PF.SampleSpec(seed=1, layer_id=0, precision=_cabi.PREC_BF16)
    with pytest.raises(RuntimeError, match="multiples of 8"):
        PF.bayes_linear(torch.zeros(2, 4, 30, device="cuda"), layer.weight.loc, layer.weight.logscale, layer.bias.loc,
                        layer.bias.logscale, None, None, spec)
    for prec in ("bf16", "tf32", "bf16_fast", "auto"):
        layer.precision = prec                   # forced and auto modes fall back to the fp32 kernels by shape
        assert layer(torch.zeros(2, 4, 30, device="cuda")).shape == (2, 4, 20)
        assert layer._last[3].precision == _cabi.PREC_FP32


# ---------------------------------------------------------------------------------------------
# BayesConv2d on the tensor-core kernels
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["conv_small", "conv_strided", "conv_k5"])
def test_tc_conv_injected_eps_matches_reference_golden(plb, name):
    g = load_golden(name)
    k, s, pad, d = (int(v) for v in g["hyper"])
    mc = g["x"].shape[0]
    cout, cin = g["w_mu"].shape[:2]
    layer = plb.BayesConv2d(cin, cout, k, stride=s, padding=pad, dilation=d, num_MC=mc,
                            prior_scale=float(g["prior"])).cuda()
    layer.precision = "bf16"
    with torch.no_grad():
        layer.weight.loc.copy_(dev(g["w_mu"]))
        layer.weight.logscale.copy_(dev(g["w_rho"]))
        layer.bias.loc.copy_(dev(g["b_mu"]))
        layer.bias.logscale.copy_(dev(g["b_rho"]))
    layer.inject_eps(dev(g["eps_w"]), dev(g["eps_b"]))
    x = dev(g["x"]).requires_grad_(True)
    out = layer(x)
    assert out.shape == g["out"].shape
    assert rel_err(out.detach().cpu().numpy(), g["out"]) < BF16_TOL
    (out * dev(g["gy"])).sum().backward()
    assert rel_err(x.grad.cpu().numpy(), g["dx"]) < BF16_TOL
    assert rel_err(layer.weight.loc.grad.cpu().numpy(), g["dw_mu"]) < BF16_TOL
    assert rel_err(layer.weight.logscale.grad.cpu().numpy(), g["dw_rho"]) < BF16_TOL
    assert rel_err(layer.bias.loc.grad.cpu().numpy(), g["db_mu"]) < 1e-4
    assert rel_err(layer.bias.logscale.grad.cpu().numpy(), g["db_rho"]) < 1e-4


@pytest.mark.parametrize("mc,b,cin,cout,hw,k,s,p,d,expand", [
    (2, 4, 8, 40, 12, 3, 1, 1, 1, False),
    (3, 2, 3, 96, 16, 5, 1, 2, 1, True),     # C3 conv1-like: K = 75 (unaligned rows), MC-broadcast input
    (2, 3, 16, 136, 9, 3, 2, 1, 1, False),   # stride 2, two Cout tiles
    (2, 2, 6, 33, 11, 3, 1, 2, 2, False),    # dilation 2
    (2, 32, 5, 32, 4, 3, 1, 1, 1, False),    # 4x4 planes: tiled col2im with several channels per CTA (cpb > Cin)
    (1, 4, 40, 64, 16, 5, 1, 2, 1, False),   # C3-like block: k5 p2, 16x16 planes, im2col with 4 output rows per CTA
    # the REAL C3 layer shapes (experiments/BNN.py:232-247; K = 2400 / 3200 / 6400), MC = 2, reference initialisation
    (2, 4, 96, 128, 16, 5, 1, 2, 1, "c3"),
    (2, 4, 128, 256, 8, 5, 1, 2, 1, "c3"),
    (2, 8, 256, 128, 4, 5, 1, 2, 1, "c3"),
])
def test_tc_conv_philox_matches_fp64_oracle(plb, mc, b, cin, cout, hw, k, s, p, d, expand):
    torch.manual_seed(cout)
    plb.manual_seed(777 + cin)
    layer = plb.BayesConv2d(cin, cout, k, stride=s, padding=p, dilation=d, num_MC=mc).cuda()
    layer.precision = "bf16"
    ref_init = expand == "c3"
    expand = expand is True
    with torch.no_grad():
        if not ref_init:
            layer.weight.logscale.uniform_(-4, -1)
            layer.bias.logscale.uniform_(-3, 0)
    if expand:
        x = torch.randn(b, cin, hw, hw, device="cuda").unsqueeze(0).expand(mc, b, cin, hw, hw)
    else:
        x = torch.randn(mc, b, cin, hw, hw, device="cuda", requires_grad=True)
    out = layer(x)
    gy = torch.randn_like(out)
    (out * gy).sum().backward()
    n = lambda t: t.detach().cpu().numpy()
    eps_w, eps_b = n(layer.weight.eps), n(layer.bias.eps)
    want = cf.conv2d_fwd(n(x), n(layer.weight.loc), n(layer.weight.logscale), eps_w,
                         n(layer.bias.loc), n(layer.bias.logscale), eps_b, s, p, d)
    assert rel_err(n(out), want) < BF16_TOL
    bw = cf.conv2d_bwd(n(gy), n(x), n(layer.weight.loc), n(layer.weight.logscale), eps_w,
                       n(layer.bias.logscale), eps_b, s, p, d)
    if not expand:
        assert rel_err(n(x.grad), bw["dx"]) < BF16_TOL
    assert rel_err(n(layer.weight.loc.grad), bw["dw_mu"]) < BF16_TOL
    assert rel_err(n(layer.weight.logscale.grad), bw["dw_rho"]) < BF16_TOL
    assert rel_err(n(layer.bias.loc.grad), bw["db_mu"]) < 1e-4
    assert rel_err(n(layer.bias.logscale.grad), bw["db_rho"]) < 1e-4


@pytest.mark.parametrize("mc,b,cin,cout,hw,k,s,p,d,expand", [
    (2, 3, 8, 16, 12, 3, 1, 1, 1, False),     # 16-byte staged loads, three output-row blocks
    (2, 3, 3, 16, 32, 5, 1, 2, 1, True),      # K = 75: tail chunk reads the zero region; 10 chunks share a warp pass
    (2, 3, 7, 16, 10, 3, 1, 1, 1, False),     # w % 4 != 0: scalar staging; ho = 10 leaves a short last block
    (2, 3, 16, 24, 9, 3, 2, 1, 1, False),     # stride 2
    (2, 2, 6, 16, 11, 3, 1, 2, 2, False),     # dilation 2
    (1, 2, 256, 16, 16, 5, 1, 2, 1, False),   # staging too large for several output rows per CTA
    (2, 4, 96, 128, 16, 5, 1, 2, 1, False),   # C3 layer 2
    (2, 8, 256, 128, 4, 5, 1, 2, 1, False),   # C3 layer 4 (4x4 planes)
])
def test_im2col_row_kernel_equals_elementwise_kernel(plb, monkeypatch, mc, b, cin, cout, hw, k, s, p, d, expand):
    """The staged im2col kernel and the element-wise one build the same bf16 matrix (PL_IM2COL_GENERIC=1, read per call
    by conv_tc.cu, selects the latter): the forward output is bit-identical, the weight gradients (pixel reduction split
    over CTAs with fp32 atomics, so order-dependent in the last bits) agree to 1e-5."""
    torch.manual_seed(cin + hw)
    layer = plb.BayesConv2d(cin, cout, k, stride=s, padding=p, dilation=d, num_MC=mc).cuda()
    layer.precision = "bf16"
    if expand:
        x0 = torch.randn(b, cin, hw, hw, device="cuda")
    else:
        x0 = torch.randn(mc, b, cin, hw, hw, device="cuda")

    def run(generic):
        if generic:
            monkeypatch.setenv("PL_IM2COL_GENERIC", "1")
        else:
            monkeypatch.delenv("PL_IM2COL_GENERIC", raising=False)
        plb.manual_seed(4242)
        layer._calls = 0          # same eps in both runs (the per-call seed counts forward calls)
        for q in layer.parameters():
            q.grad = None
        x = x0.unsqueeze(0).expand(mc, b, cin, hw, hw) if expand else x0.clone().requires_grad_(True)
        out = layer(x)
        gy = torch.sin(torch.arange(out.numel(), device="cuda", dtype=torch.float32)).view_as(out)
        (out * gy).sum().backward()
        torch.cuda.synchronize()
        return out.detach().clone(), layer.weight.loc.grad.clone(), layer.weight.logscale.grad.clone()

    a, g = run(False), run(True)
    assert torch.equal(a[0], g[0])
    assert torch.isfinite(a[0]).all() and a[0].abs().max() > 0
    for t_rows, t_generic in zip(a[1:], g[1:]):
        assert rel_err(t_rows.cpu().numpy(), t_generic.cpu().numpy()) < 1e-5


@pytest.mark.parametrize("mc,b,cin,cout,hw,k,p", [
    (2, 3, 8, 16, 12, 3, 1),       # 144-pixel planes: two images per CTA, the last image group is short
    (2, 5, 16, 24, 4, 3, 1),       # 4x4 planes: the whole batch in one CTA
    (1, 2, 20, 16, 16, 5, 2),      # channel groups 8 + 8 + 4
    (2, 4, 8, 16, 8, 1, 0),        # 1x1 kernel: one row per channel
    (2, 3, 8, 16, 8, 3, 2),        # padding = k - 1: the output planes are larger than the input planes
    (2, 5, 12, 16, 16, 3, 0),      # no padding: 14x14 output planes fall back to the element-wise kernel (196 % 8)
    (2, 5, 12, 16, 10, 3, 0),      # no padding, 8x8 output planes from 10x10 inputs
    (2, 3, 7, 16, 10, 3, 1),       # 100-pixel planes (not a multiple of 8): both runs take the element-wise kernel
    (2, 4, 96, 128, 16, 5, 2),     # C3 layer 2
    (2, 4, 128, 256, 8, 5, 2),     # C3 layer 3
    (2, 8, 256, 128, 4, 5, 2),     # C3 layer 4
])
def test_col2im_pipelined_kernel_equals_elementwise_kernel(plb, monkeypatch, mc, b, cin, cout, hw, k, p):
    """dX of the bf16 conv path: the bulk-copy pipelined col2im and the element-wise gather (PL_COL2IM_GENERIC=1, read
    per call by conv_tc.cu) sum the same bf16 dA entries in the same order -> bit-identical input gradients."""
    torch.manual_seed(cin * 3 + hw)
    layer = plb.BayesConv2d(cin, cout, k, stride=1, padding=p, num_MC=mc).cuda()
    layer.precision = "bf16"
    x0 = torch.randn(mc, b, cin, hw, hw, device="cuda")

    def run(generic):
        if generic:
            monkeypatch.setenv("PL_COL2IM_GENERIC", generic)
        else:
            monkeypatch.delenv("PL_COL2IM_GENERIC", raising=False)
        plb.manual_seed(99)
        layer._calls = 0
        x = x0.clone().requires_grad_(True)
        out = layer(x)
        gy = torch.cos(torch.arange(out.numel(), device="cuda", dtype=torch.float32)).view_as(out)
        (out * gy).sum().backward()
        torch.cuda.synchronize()
        return x.grad.clone()

    dx_default, dx_generic = run(None), run("1")
    assert torch.equal(dx_default, dx_generic)
    assert torch.isfinite(dx_default).all() and dx_default.abs().max() > 0


# ---------------------------------------------------------------------------------------------
# TF32 tensor-core mode (kind::tf32, fp32 operands straight from the caller's tensors)
# ---------------------------------------------------------------------------------------------
def test_tf32_injected_eps_matches_reference_golden(plb):
    g = load_golden("linear_tile")
    layer = plb.BayesLinear(128, 128, prior=float(g["prior"])).cuda()
    layer.precision = "tf32"
    with torch.no_grad():
        layer.weight.loc.copy_(dev(g["w_mu"]))
        layer.weight.logscale.copy_(dev(g["w_rho"]))
        layer.bias.loc.copy_(dev(g["b_mu"]))
        layer.bias.logscale.copy_(dev(g["b_rho"]))
    layer.inject_eps(dev(g["eps_w"]), dev(g["eps_b"]))
    x = dev(g["x"]).requires_grad_(True)
    out = layer(x)
    assert rel_err(out.detach().cpu().numpy(), g["out"]) < TF32_TOL
    (out * dev(g["gy"])).sum().backward()
    assert rel_err(x.grad.cpu().numpy(), g["dx"]) < TF32_TOL
    assert rel_err(layer.weight.loc.grad.cpu().numpy(), g["dw_mu"]) < TF32_TOL
    assert rel_err(layer.weight.logscale.grad.cpu().numpy(), g["dw_rho"]) < TF32_TOL
    assert rel_err(layer.bias.loc.grad.cpu().numpy(), g["db_mu"]) < 1e-4
    assert rel_err(layer.bias.logscale.grad.cpu().numpy(), g["db_rho"]) < 1e-4


@pytest.mark.parametrize("mc,b,i,o,bias,expand", [
    (2, 128, 128, 128, True, False),
    (1, 128, 32, 128, False, False),      # single k-block
    (3, 512, 256, 384, True, False),      # MT=4
    (2, 200, 192, 200, True, False),      # ragged rows / N tile
    (2, 300, 200, 136, True, False),      # K tail (200 % 32 = 8), batch tail (300 % 32 = 12)
    (4, 512, 512, 256, True, True),       # MC-broadcast input
])
def test_tf32_philox_matches_fp64_oracle(plb, mc, b, i, o, bias, expand):
    torch.manual_seed(mc * 100 + b)
    plb.manual_seed(31337 + o)
    layer = plb.BayesLinear(i, o, bias=bias).cuda()
    layer.precision = "tf32"
    with torch.no_grad():
        layer.weight.logscale.uniform_(-5, -1)
        if bias:
            layer.bias.logscale.uniform_(-3, 0)
    if expand:
        x = torch.randn(b, i, device="cuda").unsqueeze(0).expand(mc, b, i)
    else:
        x = torch.randn(mc, b, i, device="cuda", requires_grad=True)
    gy = torch.randn(mc, b, o, device="cuda")
    out = layer(x)
    (out * gy).sum().backward()
    n = lambda t: t.detach().cpu().numpy()
    eps_w = n(layer.weight.eps)
    eps_b = n(layer.bias.eps) if bias else None
    want = cf.linear_fwd(n(x), n(layer.weight.loc), n(layer.weight.logscale), eps_w,
                         n(layer.bias.loc) if bias else None,
                         n(layer.bias.logscale) if bias else None, eps_b)
    assert rel_err(n(out), want) < TF32_TOL
    bw = cf.linear_bwd(n(gy), n(x), n(layer.weight.loc), n(layer.weight.logscale), eps_w,
                       n(layer.bias.logscale) if bias else None, eps_b)
    if not expand:
        assert rel_err(n(x.grad), bw["dx"]) < TF32_TOL
    assert rel_err(n(layer.weight.loc.grad), bw["dw_mu"]) < TF32_TOL
    assert rel_err(n(layer.weight.logscale.grad), bw["dw_rho"]) < TF32_TOL


# ---------------------------------------------------------------------------------------------
# opt-in chain fusion: BayesLinear -> LeakyReLU -> BayesLinear ... with bf16 activation hand-off
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mc,b,dims,expand", [(3, 256, [256, 384, 128, 136], True),
                                              (2, 200, [192, 256, 200], False)])
def test_fused_mlp_equals_unfused_bf16_chain(plb, mc, b, dims, expand):
    from pytorch_probabilisticlayers_b200.fused import FusedBayesMLP, fuse_bayes_mlp
    torch.manual_seed(11)
    plb.manual_seed(2718)
    mods = []
    for li in range(len(dims) - 1):
        l = plb.BayesLinear(dims[li], dims[li + 1])
        l.precision = "bf16"
        mods.append(l)
        if li + 2 < len(dims):
            mods.append(torch.nn.LeakyReLU(0.01))
    net = torch.nn.Sequential(*mods).cuda()
    layers = [m for m in net if isinstance(m, plb.BayesLinear)]
    with torch.no_grad():
        for l in layers:
            l.weight.logscale.uniform_(-5, -2)
    fused = fuse_bayes_mlp(net, b)
    assert len(fused) == 1 and isinstance(fused[0], FusedBayesMLP)
    assert fused[0].layers[0] is layers[0]                    # same modules / parameters
    if expand:
        x = torch.randn(b, dims[0], device="cuda").unsqueeze(0).expand(mc, b, dims[0])
    else:
        x = torch.randn(mc, b, dims[0], device="cuda")
    gy = torch.randn(mc, b, dims[-1], device="cuda")

    def run(model):
        for l in layers:
            l._calls = 0
            l.zero_grad()
        out = model(x)
        (out * gy).sum().backward()
        return out.detach().clone(), [p.grad.clone() for l in layers for p in l.parameters()], \
            [l.kl_div.detach().clone() for l in layers]

    out_u, g_u, kl_u = run(net)
    out_f, g_f, kl_f = run(fused)
    assert torch.equal(out_f, out_u)                          # identical bf16 operands, same MMAs
    for k, (a, r) in enumerate(zip(g_f, g_u)):
        tol = 5e-3 if k % 4 >= 2 else 1e-5                    # bias grads see bf16-rounded dY
        assert rel_err(a.cpu().numpy(), r.cpu().numpy()) < tol, k
    for a, r in zip(kl_f, kl_u):
        assert torch.equal(a, r)


def test_fused_chain_passes_gradients_upstream_and_fuses_supported_prefix(plb):
    """A fused chain that is not the first module must hand dX to what precedes it, and a narrow head must not stop
    the wide layers in front of it from chaining (fuse_bayes_mlp fuses the supported part of a run)."""
    from pytorch_probabilisticlayers_b200.fused import fuse_bayes_mlp
    mc, b = 2, 128
    torch.manual_seed(5)
    plb.manual_seed(77)
    net = torch.nn.Sequential(plb.BayesLinear(24, 128), torch.nn.LeakyReLU(), plb.BayesLinear(128, 256),
                              torch.nn.LeakyReLU(), plb.BayesLinear(256, 128), torch.nn.LeakyReLU(),
                              plb.BayesLinear(128, 10)).cuda()
    layers = [m for m in net if isinstance(m, plb.BayesLinear)]
    for l in layers:
        l.precision = "auto"
    fused = fuse_bayes_mlp(net, b)
    kinds = [type(m).__name__ for m in fused]
    assert kinds == ["BayesLinear", "LeakyReLU", "FusedBayesMLP", "LeakyReLU", "BayesLinear"], kinds
    x0 = torch.randn(mc, b, 24, device="cuda")
    gy = torch.randn(mc, b, 10, device="cuda")

    def run(model):
        for l in layers:
            l._calls = 0                              # same per-forward seed in both runs
            l.zero_grad()
        x = x0.clone().requires_grad_(True)
        out = model(x)
        (out * gy).sum().backward()
        return [t.detach().clone() for t in (out, x.grad, layers[0].weight.loc.grad, layers[0].weight.logscale.grad,
                                             layers[1].weight.loc.grad)]
    ref, got = run(net), run(fused)
    for a, r in zip(got, ref):
        assert rel_err(a.cpu().numpy(), r.cpu().numpy()) < 5e-3


def test_tc_backward_mc_shards_sum_to_unsharded_gradients(plb):
    """Full-width layer (4096 x 1024), bf16 kernels: the gradients of two MC shards add up to the
    unsharded gradients (the property the one-all-reduce-per-step train step relies on)."""
    torch.manual_seed(21)
    plb.manual_seed(1212)
    layer = plb.BayesLinear(4096, 1024).cuda()
    layer.precision = "bf16"
    with torch.no_grad():
        layer.weight.logscale.uniform_(-6, -3)
    x = torch.randn(4, 512, 4096, device="cuda")
    gy = torch.randn(4, 512, 1024, device="cuda")

    def grads(lo, hi):
        layer._calls = 0
        layer.zero_grad()
        plb.set_mc_shard(lo)
        try:
            xs = x[lo:hi].clone().requires_grad_(True)
            (layer(xs) * gy[lo:hi]).sum().backward()
        finally:
            plb.set_mc_shard(0)
        return xs.grad, [p.grad.clone() for p in layer.parameters()]

    dx_full, g_full = grads(0, 4)
    dx_a, g_a = grads(0, 2)
    dx_b, g_b = grads(2, 4)
    assert torch.equal(torch.cat([dx_a, dx_b]), dx_full)            # per-sample work is identical
    for a, b, f in zip(g_a, g_b, g_full):
        assert rel_err((a + b).cpu().numpy(), f.cpu().numpy()) < 1e-5


def test_train_step_with_fused_chain_matches_module_by_module(plb):
    from pytorch_probabilisticlayers_b200.fused import fuse_bayes_mlp
    from pytorch_probabilisticlayers_b200.train import ELBOTrainStep
    dims, mc, b, ns = [256, 384, 256, 16], 4, 256, 1000     # 256 -> 16 head: fp32 kernels after a fused bf16 chain
    results = []
    for fuse in (False, True):
        torch.manual_seed(5)
        plb.manual_seed(99)
        plb.set_precision("auto")
        try:
            mods = [plb.MC_ExpansionLayer(mc, 2)]
            for li in range(3):
                mods.append(plb.BayesLinear(dims[li], dims[li + 1]))
                if li < 2:
                    mods.append(torch.nn.LeakyReLU())
            net = torch.nn.Sequential(*mods).cuda()
            layers = [m for m in net.modules() if isinstance(m, plb.BayesLinear)]
            for k, l in enumerate(layers):
                l.layer_id = 1000 + k                                # same eps stream in both runs
            if fuse:
                net = fuse_bayes_mlp(net, b)
            tr = ELBOTrainStep(net, plb.MC_CrossEntropyLoss(ns), ns, lr=1e-3)
            x = torch.randn(b, dims[0], device="cuda")
            y = torch.randint(0, dims[-1], (b,), device="cuda")
            stats = [tr.step(x, y).tolist() for _ in range(3)]
            results.append((stats, [p.detach().clone() for p in net.parameters()]))
        finally:
            plb.set_precision("fp32")
    (s0, p0), (s1, p1) = results
    for a, c in zip(s0, s1):
        assert abs(a[3] - c[3]) / abs(a[3]) < 1e-4                   # loss trajectory
    for a, c in zip(p0, p1):     # the chained dX / bias gradients see bf16-rounded dY: bf16 tolerance after 3 Adam steps
        assert rel_err(c.cpu().numpy(), a.cpu().numpy()) < BF16_TOL


def test_cuda_graph_step_equals_eager_steps(plb):
    """ELBOTrainStep.enable_cuda_graph: the whole step replayed from a CUDA graph, with the per-step part of the
    eps seed and Adam's bias-correction scalars advanced in device memory, follows the eager trajectory."""
    from pytorch_probabilisticlayers_b200.fused import fuse_bayes_mlp
    from pytorch_probabilisticlayers_b200.train import ELBOTrainStep
    dims, mc, b, ns = [256, 384, 256, 16], 4, 256, 1000     # fused bf16 chain + fp32 head + fused CE + fused Adam
    results = []
    for graph in (False, True):
        torch.manual_seed(5)
        plb.manual_seed(99)
        plb.set_precision("auto")
        try:
            mods = [plb.MC_ExpansionLayer(mc, 2)]
            for li in range(3):
                mods.append(plb.BayesLinear(dims[li], dims[li + 1]))
                if li < 2:
                    mods.append(torch.nn.LeakyReLU())
            net = torch.nn.Sequential(*mods).cuda()
            layers = [m for m in net.modules() if isinstance(m, plb.BayesLinear)]
            for k, l in enumerate(layers):
                l.layer_id = 2000 + k                                # same eps stream in both runs
            layers[-1].precision = "fp32"     # the 16-wide head in exact arithmetic (under 'auto' it would run as a padded
            net = fuse_bayes_mlp(net, b)      # tensor-core head whose atomic MC-split dW order changes from run to run)
            tr = ELBOTrainStep(net, plb.MC_CrossEntropyLoss(ns), ns, lr=1e-3)
            x = torch.randn(b, dims[0], device="cuda")
            y = torch.randint(0, dims[-1], (b,), device="cuda")
            if graph:
                tr.enable_cuda_graph(x, y, warmup=3)                 # 3 eager steps, then replays
                stats = [tr.step(x, y).tolist() for _ in range(5)]
            else:
                stats = [tr.step(x, y).tolist() for _ in range(8)][3:]
            torch.cuda.synchronize()
            results.append((stats, [p.detach().clone() for p in net.parameters()]))
        finally:
            plb.set_precision("fp32")
    (s0, p0), (s1, p1) = results
    assert len({tuple(s) for s in s1}) == 5                          # the replays really draw fresh eps each step
    # not bit-equal: torch.pow vs libm pow for Adam's bias corrections (1 ulp of fp32) and the atomics of the split
    # dW kernels reorder sums; both are amplified a little by the bf16 chain over 5 Adam steps
    for a, c in zip(s0, s1):
        for u, v in zip(a, c):
            assert abs(u - v) <= 2e-4 * max(abs(u), 1e-6), (a, c)
    for a, c in zip(p0, p1):     # Adam turns last-bit gradient differences of near-zero gradients into O(lr) steps
        assert rel_err(c.cpu().numpy(), a.cpu().numpy()) < BF16_TOL


@pytest.mark.parametrize("env", [{"PL_TC_2CTA": "1"}, {"PL_DW_2CTA": "1"}])
def test_opt_in_cta_pair_kernels_pass_the_oracle_tests(env):
    """The CTA-pair (cta_group::2) kernels are opt-in (they measured equal / slower, profiles/r02_experiments.md) but stay
    parity-tested: the library reads the switches once per process, so the oracle tests run in a child process."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", os.path.join(root, "tests", "test_gpu_tc.py"),
                        os.path.join(root, "tests", "test_gpu_noise.py"), "-k",
                        "philox_matches_fp64_oracle or baseline_shapes or fused_mlp_equals or noise_survives",
                        "-p", "no:cacheprovider"], capture_output=True, text=True, timeout=900, cwd=root, env=e)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("mc,b,i,o", [(4, 256, 1200, 10), (2, 128, 256, 3), (2, 160, 512, 20)])
def test_narrow_head_runs_on_tensor_cores_through_padded_views(plb, mc, b, i, o):
    """'auto' pads a narrow classifier head (C2's 1200 -> 10) to a multiple of 8 columns for the tcgen05 kernels; the first
    `o` columns draw the eps of the unpadded layer, so outputs, dX and parameter gradients match the fp32 path on the same
    eps within the bf16 tolerance, and the lazily dumped eps is the unpadded one."""
    from pytorch_probabilisticlayers_b200 import _cabi
    res = {}
    for prec in ("fp32", "auto"):
        torch.manual_seed(i + o)
        plb.manual_seed(2024)
        layer = plb.BayesLinear(i, o).cuda()
        layer.layer_id = 6000 + o
        layer.precision = prec
        with torch.no_grad():
            layer.weight.logscale.fill_(-4.0)
        x = torch.randn(mc, b, i, device="cuda", requires_grad=True)
        gy = torch.randn(mc, b, o, device="cuda")
        out = layer(x)
        assert out.shape == (mc, b, o)
        (out * gy).sum().backward()
        res[prec] = [t.detach().cpu().numpy() for t in (out, x.grad, layer.weight.loc.grad, layer.weight.logscale.grad,
                                                        layer.bias.loc.grad, layer.bias.logscale.grad, layer.weight.eps)]
        if prec == "auto":
            assert layer._last[3].precision == _cabi.PREC_BF16, "the narrow head did not take the tensor-core path"
    for a, b_ in zip(res["auto"][:6], res["fp32"][:6]):
        assert rel_err(a, b_) < BF16_TOL
    assert np.array_equal(res["auto"][6], res["fp32"][6])          # the same eps stream
