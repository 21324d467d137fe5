"""Out-of-bounds canaries around every buffer the tensor-core kernels write (compute-sanitizer is closed on this GPU pool,
so memory safety is checked with guard bands): pl_linear_fwd / pl_linear_bwd are called straight through the C ABI with
caller-owned buffers that carry a poisoned band on both sides; ragged shapes (row, N and K tails), every operand form,
MC-broadcast input (hoisted-mean launches) included."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

GUARD = 4096          # bytes on each side
POISON = 0x5A


@pytest.fixture(scope="module")
def plb():
    import pytorch_probabilisticlayers_b200 as p
    assert torch.cuda.is_available()
    return p


class Guarded:
    """A device buffer of ``nbytes`` with poisoned guard bands; ``.ptr`` is the 256-byte aligned payload."""

    def __init__(self, nbytes, fill=None):
        nbytes = int(nbytes)
        self.n = nbytes
        self.raw = torch.full((nbytes + 2 * GUARD + 512,), POISON, dtype=torch.uint8, device="cuda")
        base = self.raw.data_ptr()
        self.off = (-(base + GUARD)) % 256 + GUARD
        self.ptr = base + self.off
        if fill is not None:
            self.raw[self.off:self.off + nbytes] = fill

    def view(self, dtype, shape):
        return self.raw[self.off:self.off + self.n].view(dtype).view(shape)

    def intact(self):
        lo = self.raw[:self.off]
        hi = self.raw[self.off + self.n:]
        return bool((lo == POISON).all()) and bool((hi == POISON).all())


@pytest.mark.parametrize("precision", ["bf16", "tf32", "bf16_fast", "tf32_fast"])
@pytest.mark.parametrize("mc,b,i,o,broadcast", [(2, 200, 200, 136, False), (4, 300, 192, 200, True), (1, 130, 72, 8, False)])
def test_linear_tc_kernels_stay_inside_their_buffers(plb, precision, mc, b, i, o, broadcast):
    from pytorch_probabilisticlayers_b200 import _cabi
    lib = _cabi.lib()
    torch.manual_seed(1)
    f32 = torch.float32
    x = torch.randn((b, i) if broadcast else (mc, b, i), device="cuda")
    w_mu, w_rho = torch.randn(i, o, device="cuda") * 0.05, torch.full((i, o), -4.0, device="cuda")
    b_mu, b_rho = torch.zeros(o, device="cuda"), torch.full((o,), -3.0, device="cuda")
    dy = torch.randn(mc, b, o, device="cuda")
    a = _cabi.LinearArgs()
    a.mc, a.batch, a.in_features, a.out_features = mc, b, i, o
    a.precision, a.param_mode = _cabi.PRECISIONS[precision], _cabi.PARAM_SHARED_RHO
    a.x, a.x_mc_stride = x.data_ptr(), 0 if broadcast else b * i
    a.w_mu, a.w_rho, a.b_mu, a.b_rho = w_mu.data_ptr(), w_rho.data_ptr(), b_mu.data_ptr(), b_rho.data_ptr()
    a.seed, a.layer_id = 1234, 7
    y = Guarded(mc * b * o * 4)
    a.y = y.ptr
    state = Guarded(lib.pl_linear_state_bytes(C.byref(a)))
    a.fwd_state, a.fwd_state_bytes = state.ptr, state.n
    ws = Guarded(lib.pl_linear_workspace_bytes(C.byref(a)))
    a.workspace, a.workspace_bytes = ws.ptr, ws.n
    a.stream = torch.cuda.current_stream().cuda_stream
    _cabi.check(lib.pl_linear_fwd(C.byref(a)), "pl_linear_fwd")
    torch.cuda.synchronize()
    assert y.intact() and state.intact() and ws.intact(), "forward wrote outside a buffer"
    assert torch.isfinite(y.view(f32, (mc, b, o))).all()
    # backward into guarded gradient buffers
    a.dy = dy.data_ptr()
    dx = Guarded(mc * b * i * 4)
    dwm, dwr = Guarded(i * o * 4), Guarded(i * o * 4)
    dbm, dbr = Guarded(o * 4), Guarded(o * 4)
    a.dx = None if broadcast else dx.ptr
    a.dw_mu, a.dw_rho, a.db_mu, a.db_rho = dwm.ptr, dwr.ptr, dbm.ptr, dbr.ptr
    ws2 = Guarded(lib.pl_linear_workspace_bytes(C.byref(a)))
    a.workspace, a.workspace_bytes = ws2.ptr, ws2.n
    _cabi.check(lib.pl_linear_bwd(C.byref(a)), "pl_linear_bwd")
    torch.cuda.synchronize()
    for name, g in (("dx", dx), ("dw_mu", dwm), ("dw_rho", dwr), ("db_mu", dbm), ("db_rho", dbr), ("state", state),
                    ("workspace", ws2)):
        assert g.intact(), f"backward wrote outside {name}"
    assert torch.isfinite(dwm.view(f32, (i, o))).all() and torch.isfinite(dwr.view(f32, (i, o))).all()
    if not broadcast:
        assert torch.isfinite(dx.view(f32, (mc, b, i))).all()
