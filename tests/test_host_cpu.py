"""CPU-only checks: the C-ABI library loads and exports every symbol the header declares, argument
validation works without a GPU, and the host-side module mirror matches the reference's
construction semantics (state-dict names, shapes, seeded initialisation)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden


def header_symbols():
    text = open(os.path.join(ROOT, "include", "pl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pl_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from pytorch_probabilisticlayers_b200 import _cabi
    lib = _cabi.lib()
    syms = header_symbols()
    assert set(syms) == set(_cabi.EXPORTS), (syms, _cabi.EXPORTS)
    for s in syms:
        assert hasattr(lib, s), s
    assert lib.pl_version() == 100
    assert lib.pl_kl_workspace_bytes() >= 3 * 8 * 148


def test_argument_validation_without_gpu():
    from pytorch_probabilisticlayers_b200 import _cabi
    lib = _cabi.lib()
    assert lib.pl_kl_fwd(None, None, 10, 1.0, None, None, None, 0, None) == -1
    assert b"NULL" in lib.pl_last_error()
    assert lib.pl_linear_fwd(None) == -1
    a = _cabi.LinearArgs()
    a.mc, a.batch, a.in_features, a.out_features = 1, 1, -4, 1
    assert lib.pl_linear_fwd(C.byref(a)) == -1
    assert b"negative" in lib.pl_last_error()
    a.in_features = 4
    a.precision = 7
    assert lib.pl_linear_fwd(C.byref(a)) == -2          # unsupported precision
    c = _cabi.Conv2dArgs()
    c.mc, c.batch, c.cin, c.cout, c.h, c.w, c.k, c.stride, c.dilation = 1, 1, 1, 1, 4, 4, 3, 0, 1
    assert lib.pl_conv2d_fwd(C.byref(c)) == -1
    assert lib.pl_eps_dump(1, 0, 0, 1, 1, 1, None, 0, None) == -1


def test_struct_layout_matches_header():
    """ctypes mirrors must have the C struct's size (catches a field added on one side only)."""
    import subprocess, tempfile, textwrap
    from pytorch_probabilisticlayers_b200 import _cabi
    src = textwrap.dedent("""
        #include <stdio.h>
        #include "pl_b200.h"
        int main(void){ printf("%zu %zu %zu %zu\\n", sizeof(pl_linear_args), sizeof(pl_conv2d_args), sizeof(pl_ce_args), sizeof(pl_bnpool_args)); return 0; }
    """)
    with tempfile.TemporaryDirectory() as d:
        cfile = os.path.join(d, "s.c")
        open(cfile, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), cfile, "-o", exe])
        out = subprocess.check_output([exe]).split()
    assert int(out[0]) == C.sizeof(_cabi.LinearArgs)
    assert int(out[1]) == C.sizeof(_cabi.Conv2dArgs)
    assert int(out[2]) == C.sizeof(_cabi.CeArgs)
    assert int(out[3]) == C.sizeof(_cabi.BnPoolArgs)


def test_seeded_construction_matches_reference_state_dict():
    import pytorch_probabilisticlayers_b200 as p
    g = load_golden("seeded_init")
    for prefix, seed, make in (("lin__", 3, lambda: p.BayesLinear(7, 5)),
                               ("conv__", 4, lambda: p.BayesConv2d(2, 3, 3)),
                               ("sivi__", 5, lambda: p.SIVILayer(6, 1, 5))):
        torch.manual_seed(seed)
        sd = make().state_dict()
        want = {k[len(prefix):].replace("__", "."): v for k, v in g.items() if k.startswith(prefix)}
        assert set(sd) == set(want), (sorted(sd), sorted(want))
        for k, v in want.items():
            assert np.allclose(sd[k].numpy(), v, rtol=1e-6, atol=0), (prefix, k)


def test_cpu_tensors_fail_loudly_and_asserts_match_reference():
    import pytorch_probabilisticlayers_b200 as p
    layer = p.BayesLinear(4, 3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        layer(torch.zeros(2, 5, 4))
    with pytest.raises(AssertionError, match="Input tensor not of shape"):
        layer(torch.zeros(5, 4))
    with pytest.raises(AssertionError):
        p.BayesConv2d(2, 3, 3, num_MC=4)(torch.zeros(2, 1, 2, 5, 5))
    with pytest.raises(AssertionError):
        p.MC_CrossEntropyLoss(0)
    ex = p.MC_ExpansionLayer(num_MC=3, input_dim=2)
    assert ex(torch.zeros(5, 4)).shape == (3, 5, 4)
    assert ex(torch.zeros(3, 5, 4)).shape == (3, 5, 4)
    with pytest.raises(ValueError):
        ex(torch.zeros(4))


def test_dropin_import_paths():
    """The reference's experiments import pytorch_ProbabilisticLayers.src.* (experiments/BNN.py:42-47)."""
    from pytorch_ProbabilisticLayers.src.ProbabilisticLayers import (  # noqa: F401
        BayesAdaptiveInit_FlattenAndLinear, BayesConv2d, BayesLinear, BayesianNeuralNetwork,
        MC_Accuracy, MC_BatchNorm1D, MC_BatchNorm2D, MC_CrossEntropyLoss, MC_ExpansionLayer,
        MC_MaxPool2d, SIVILayer, VariationalNormal)
    from pytorch_ProbabilisticLayers.src.ProbabilisticLayers_Loss import MC_NLL  # noqa: F401
    from pytorch_ProbabilisticLayers.src.ProbabilisticLayers_Utils import (  # noqa: F401
        MC_GradientCorrection, RunningAverageMeter)
    from pytorch_ProbabilisticLayers.data.ProbabilisticLayers_SyntheticData import (  # noqa: F401
        generate_nonstationary_data)
    import pytorch_probabilisticlayers_b200 as p
    assert BayesLinear is p.BayesLinear


def test_fused_bn_act_pool_module_on_cpu_equals_the_three_modules():
    """bnpool.FusedBNActPool keeps the batch-norm object and, where the CUDA pass does not apply (CPU tensors, odd
    planes), runs the reference's three modules unchanged; fuse_bn_act_pool rewrites only matching triples."""
    import torch
    import pytorch_probabilisticlayers_b200 as plb
    from pytorch_probabilisticlayers_b200.bnpool import FusedBNActPool, fuse_bn_act_pool
    torch.manual_seed(3)
    net = torch.nn.Sequential(plb.MC_BatchNorm2D(4), torch.nn.LeakyReLU(0.02), plb.MC_MaxPool2d(2, 2),
                              plb.MC_BatchNorm2D(4), torch.nn.ReLU(), plb.MC_MaxPool2d(2, 2))
    fused = fuse_bn_act_pool(net)
    assert isinstance(fused[0], FusedBNActPool) and fused[0].bn is net[0] and fused[0].negative_slope == 0.02
    assert len(fused) == 4 and fused[1] is net[3]            # BN -> ReLU -> pool is left alone
    x = torch.randn(2, 3, 4, 6, 6)
    ref_bn = plb.MC_BatchNorm2D(4)
    want = plb.MC_MaxPool2d(2, 2)(torch.nn.functional.leaky_relu(ref_bn(x), 0.02))
    got = fused[0](x)
    assert torch.allclose(got, want, atol=1e-6)
    assert torch.allclose(net[0].running_var, ref_bn.running_var)


def test_integration_md_struct_mirror_matches_the_library():
    """The ctypes mirror of pl_linear_args printed in INTEGRATION.md (the stub a maintainer of the reference would paste)
    has the size the built library reports: the document cannot drift from include/pl_b200.h unnoticed."""
    import ctypes as C
    import os
    import re
    from pytorch_probabilisticlayers_b200 import _cabi
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "INTEGRATION.md")).read()
    m = re.search(r"class pl_linear_args\(C.Structure\):.*?\n    _fields_ = (\[.*?\])\n", src, re.S)
    assert m, "INTEGRATION.md no longer shows the pl_linear_args mirror"

    class Mirror(C.Structure):
        _fields_ = eval(m.group(1), {"C": C})
    assert C.sizeof(Mirror) == _cabi.lib().pl_struct_size(0) == C.sizeof(_cabi.LinearArgs)
