"""ctypes binding of libpl_b200.so (include/pl_b200.h).

There is no fallback: if the library is missing or a call fails the caller gets an exception.
PyTorch is only used for device memory and the current CUDA stream; every argument that crosses
the boundary is a raw pointer or a plain integer/float.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# PL_B200_LIB: developer override used by tools/build_probe.sh to time instrumented builds
LIB_PATH = os.environ.get("PL_B200_LIB") or os.path.join(_HERE, "lib", "libpl_b200.so")

PL_OK = 0
PREC_FP32, PREC_BF16, PREC_TF32, PREC_BF16_FAST, PREC_TF32_FAST = 0, 1, 2, 3, 4
PARAM_SHARED_RHO, PARAM_PER_MC_SIGMA = 0, 1
PRECISIONS = {"fp32": PREC_FP32, "bf16": PREC_BF16, "tf32": PREC_TF32, "bf16_fast": PREC_BF16_FAST,
              "tf32_fast": PREC_TF32_FAST}
BF16_CLASS = (PREC_BF16, PREC_BF16_FAST)       # bf16 MMA operands (activation chaining, bf16 x kept as backward state)

EXPORTS = (
    "pl_version", "pl_last_error", "pl_launch_count", "pl_device_supported", "pl_struct_size", "pl_eps_dump",
    "pl_eps_dump_dev",
    "pl_kl_workspace_bytes", "pl_kl_fwd", "pl_kl_bwd", "pl_adam_kl_step", "pl_adam_kl_step_dev",
    "pl_mc_cross_entropy_workspace_bytes", "pl_mc_cross_entropy",
    "pl_bn_stats", "pl_bn_act_pool_workspace_bytes", "pl_bn_act_pool_fwd", "pl_bn_act_pool_bwd",
    "pl_linear_workspace_bytes", "pl_linear_state_bytes", "pl_linear_fwd", "pl_linear_bwd",
    "pl_conv2d_workspace_bytes", "pl_conv2d_state_bytes", "pl_conv2d_fwd", "pl_conv2d_bwd",
    "pl_ipc_export", "pl_ipc_open", "pl_ipc_close",
    "pl_shard_push", "pl_shard_push_range", "pl_shard_barrier", "pl_shard_adam", "pl_shard_adam_range", "pl_shard_finish",
)


class LinearArgs(C.Structure):
    _fields_ = [
        ("mc", C.c_int32), ("batch", C.c_int32), ("in_features", C.c_int32),
        ("out_features", C.c_int32), ("precision", C.c_int32), ("param_mode", C.c_int32),
        ("x", C.c_void_p), ("x_mc_stride", C.c_int64),
        ("w_mu", C.c_void_p), ("w_rho", C.c_void_p), ("b_mu", C.c_void_p), ("b_rho", C.c_void_p),
        ("seed", C.c_uint64), ("layer_id", C.c_uint32), ("mc_global_offset", C.c_uint32),
        ("eps_w", C.c_void_p), ("eps_b", C.c_void_p),
        ("y", C.c_void_p), ("dy", C.c_void_p), ("dx", C.c_void_p),
        ("dw_mu", C.c_void_p), ("dw_rho", C.c_void_p), ("db_mu", C.c_void_p), ("db_rho", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("stream", C.c_void_p),
        ("fwd_state", C.c_void_p), ("fwd_state_bytes", C.c_size_t),
        ("x_lp", C.c_void_p), ("y_lp", C.c_void_p), ("y_act", C.c_int32), ("y_act_slope", C.c_float),
        ("dy_lp", C.c_void_p), ("dx_lp", C.c_void_p), ("x_act_slope", C.c_float),
        ("seed_dev", C.c_void_p),
        ("w_lp", C.c_void_p), ("w_lp_half", C.c_int64), ("dw_rho_raw", C.c_int32),
        ("dw_scatter_peers", C.c_void_p), ("dw_scatter_mu_off", C.c_int64), ("dw_scatter_rho_off", C.c_int64),
        ("dw_scatter_chunk", C.c_int64), ("dw_scatter_rank", C.c_int32), ("dw_scatter_world", C.c_int32),
        ("dw_scatter_bf16", C.c_int32),
    ]


class Conv2dArgs(C.Structure):
    _fields_ = [
        ("mc", C.c_int32), ("batch", C.c_int32), ("cin", C.c_int32), ("cout", C.c_int32),
        ("h", C.c_int32), ("w", C.c_int32), ("k", C.c_int32), ("stride", C.c_int32),
        ("padding", C.c_int32), ("dilation", C.c_int32), ("precision", C.c_int32),
        ("x", C.c_void_p), ("x_mc_stride", C.c_int64),
        ("w_mu", C.c_void_p), ("w_rho", C.c_void_p), ("b_mu", C.c_void_p), ("b_rho", C.c_void_p),
        ("seed", C.c_uint64), ("layer_id", C.c_uint32), ("mc_global_offset", C.c_uint32),
        ("eps_w", C.c_void_p), ("eps_b", C.c_void_p),
        ("y", C.c_void_p), ("dy", C.c_void_p), ("dx", C.c_void_p),
        ("dw_mu", C.c_void_p), ("dw_rho", C.c_void_p), ("db_mu", C.c_void_p), ("db_rho", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("stream", C.c_void_p),
        ("fwd_state", C.c_void_p), ("fwd_state_bytes", C.c_size_t),
        ("seed_dev", C.c_void_p),
    ]


class CeArgs(C.Structure):
    _fields_ = [
        ("mc", C.c_int32), ("batch", C.c_int32), ("classes", C.c_int32), ("grad_scale", C.c_float),
        ("pred", C.c_void_p), ("target", C.c_void_p), ("dpred", C.c_void_p), ("out", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("stream", C.c_void_p),
    ]


class BnPoolArgs(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("channels", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
        ("slope", C.c_float), ("batch_stats", C.c_int32),
        ("x", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p), ("mean", C.c_void_p),
        ("rstd", C.c_void_p), ("y", C.c_void_p), ("idx", C.c_void_p), ("dy", C.c_void_p),
        ("dx", C.c_void_p), ("dgamma", C.c_void_p), ("dbeta", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("stream", C.c_void_p),
    ]


_lib = None
_lock = threading.Lock()


def lib():
    """The loaded library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
                        f"g.build()'` (or `make -C pytorch_probabilisticlayers_b200/csrc`). "
                        "There is no CPU or PyTorch fallback for the variational layers.")
                l = C.CDLL(LIB_PATH)
                l.pl_version.restype = C.c_int
                l.pl_last_error.restype = C.c_char_p
                l.pl_device_supported.restype = C.c_int
                l.pl_launch_count.restype = C.c_uint64
                l.pl_struct_size.restype = C.c_size_t
                l.pl_struct_size.argtypes = [C.c_int32]
                l.pl_eps_dump.restype = C.c_int
                l.pl_eps_dump.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_int32, C.c_int32,
                                          C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
                l.pl_eps_dump_dev.restype = C.c_int
                l.pl_eps_dump_dev.argtypes = [C.c_uint64, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int32, C.c_int32,
                                              C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
                l.pl_kl_workspace_bytes.restype = C.c_size_t
                l.pl_kl_fwd.restype = C.c_int
                l.pl_kl_fwd.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
                l.pl_kl_bwd.restype = C.c_int
                l.pl_kl_bwd.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p,
                                        C.c_int32, C.c_void_p]
                l.pl_adam_kl_step.restype = C.c_int
                l.pl_adam_kl_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                              C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float,
                                              C.c_float, C.c_float, C.c_int32, C.c_void_p, C.c_void_p]
                l.pl_adam_kl_step_dev.restype = C.c_int
                l.pl_adam_kl_step_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                                  C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float,
                                                  C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
                l.pl_mc_cross_entropy_workspace_bytes.restype = C.c_size_t
                l.pl_mc_cross_entropy.restype = C.c_int
                l.pl_mc_cross_entropy.argtypes = [C.POINTER(CeArgs)]
                l.pl_bn_stats.restype = C.c_int
                l.pl_bn_stats.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
                l.pl_bn_act_pool_workspace_bytes.restype = C.c_size_t
                l.pl_bn_act_pool_workspace_bytes.argtypes = [C.c_int32]
                for name in ("pl_bn_act_pool_fwd", "pl_bn_act_pool_bwd"):
                    getattr(l, name).restype = C.c_int
                    getattr(l, name).argtypes = [C.POINTER(BnPoolArgs)]
                l.pl_linear_workspace_bytes.restype = C.c_size_t
                l.pl_linear_workspace_bytes.argtypes = [C.POINTER(LinearArgs)]
                l.pl_linear_state_bytes.restype = C.c_size_t
                l.pl_linear_state_bytes.argtypes = [C.POINTER(LinearArgs)]
                for name in ("pl_linear_fwd", "pl_linear_bwd"):
                    getattr(l, name).restype = C.c_int
                    getattr(l, name).argtypes = [C.POINTER(LinearArgs)]
                l.pl_conv2d_workspace_bytes.restype = C.c_size_t
                l.pl_conv2d_workspace_bytes.argtypes = [C.POINTER(Conv2dArgs)]
                l.pl_conv2d_state_bytes.restype = C.c_size_t
                l.pl_conv2d_state_bytes.argtypes = [C.POINTER(Conv2dArgs)]
                for name in ("pl_conv2d_fwd", "pl_conv2d_bwd"):
                    getattr(l, name).restype = C.c_int
                    getattr(l, name).argtypes = [C.POINTER(Conv2dArgs)]
                for name in ("pl_ipc_export", "pl_ipc_open", "pl_ipc_close", "pl_shard_push", "pl_shard_push_range", "pl_shard_barrier",
                             "pl_shard_adam", "pl_shard_adam_range", "pl_shard_finish"):
                    getattr(l, name).restype = C.c_int
                l.pl_ipc_export.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]
                l.pl_ipc_open.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p)]
                l.pl_ipc_close.argtypes = [C.c_void_p, C.c_int64]
                l.pl_shard_push.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
                l.pl_shard_push_range.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
                l.pl_shard_barrier.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
                l.pl_shard_adam_range.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                                  C.c_float, C.c_float, C.c_float, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
                l.pl_shard_adam.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float,
                                            C.c_float, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
                l.pl_shard_finish.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
                _lib = l
    return _lib


def check(rc, what):
    if rc != PL_OK:
        msg = lib().pl_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed with status {rc}: {msg}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def current_stream(device):
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda_f32(t, name):
    if not t.is_cuda:
        raise RuntimeError(
            f"{name} is on {t.device}: the B200 variational layers run only on CUDA "
            "(no CPU fallback; use the reference implementation for CPU runs)")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
