"""ctypes binding of include/dpi_b200.h (the C-ABI drop-in boundary).

There is no CPU fallback: if `libdpi_b200.so` is missing or cannot be loaded this module raises, and every
entry point returns an error (surfaced as RuntimeError) when no sm_100 device is present.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
# DPI_B200_LIB: developer override used by tools/variants.py to time experimental builds of the same library
LIB_PATH = os.environ.get("DPI_B200_LIB") or os.path.join(_PKG, "libdpi_b200.so")

DPI_MAX_LINEAR = 8
DPI_EQ_CHA, DPI_EQ_GBM, DPI_EQ_OU = 0, 1, 2
DPI_NET_ZERO, DPI_NET_MLP, DPI_NET_PISGRAD = 0, 1, 2
DPI_PREC_FP32, DPI_PREC_F16 = 0, 1
DPI_PREC_BF16 = 2  # distinct and rejected (DPI_ERR_UNSUPPORTED): bf16 operands are not built
DPI_PREC_TF32 = 3  # PISGradNet: layer-by-layer pipeline with TF32 tensor-core GEMMs
DPI_ERR_NONFINITE = -5
DPI_DP_HANDLE_BYTES = 64

EXPORTED_SYMBOLS = (
    "dpi_version", "dpi_last_error", "dpi_create", "dpi_destroy", "dpi_set_problem", "dpi_set_network",
    "dpi_sample_tx", "dpi_label_gen", "dpi_label_gen_host", "dpi_net_eval", "dpi_launch_count",
    "dpi_set_stage_timing", "dpi_stage_ms", "dpi_selftest_umma",
    "dpi_train_init", "dpi_train_param_count", "dpi_train_grad", "dpi_train_adam", "dpi_train_fit", "dpi_train_reset",
    "dpi_check_labels", "dpi_export_draws", "dpi_sample_tx_ex", "dpi_net_hessian_diag", "dpi_net_hessian",
    "dpi_train_dp_export", "dpi_train_dp_connect",
)


class dpi_problem(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("nx", C.c_int32), ("T", C.c_float), ("alpha", C.c_float), ("cha_k", C.c_float),
        ("gbm_nodes", C.c_int32), ("gbm_w", C.POINTER(C.c_double)), ("gbm_v", C.POINTER(C.c_double)),
        ("ou_K", C.c_int32), ("ou_mean", C.POINTER(C.c_double)), ("ou_diagvar", C.POINTER(C.c_double)),
        ("ou_pi", C.POINTER(C.c_double)), ("ou_theta", C.c_float), ("ou_mu", C.c_float),
        ("ou_alpha_init", C.c_float), ("sdgd_v", C.c_int32),
    ]


class dpi_linear(C.Structure):
    _fields_ = [("W", C.c_void_p), ("b", C.c_void_p), ("n_in", C.c_int32), ("n_out", C.c_int32)]


class dpi_network(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("n_main", C.c_int32), ("main", dpi_linear * DPI_MAX_LINEAR),
        ("n_t_encoder", C.c_int32), ("t_encoder", dpi_linear * DPI_MAX_LINEAR),
        ("n_smooth", C.c_int32), ("smooth_net", dpi_linear * DPI_MAX_LINEAR),
        ("timestep_phase", C.c_void_p), ("timestep_coeff", C.c_void_p),
    ]


class dpi_draws(C.Structure):
    _fields_ = [("zT", C.c_void_p), ("u_s", C.c_void_p), ("zS", C.c_void_p), ("idx", C.c_void_p)]


class dpi_label_config(C.Structure):
    _fields_ = [
        ("M_T", C.c_int32), ("M_I", C.c_int32), ("precision", C.c_int32), ("sample_bound", C.c_float),
        ("seed", C.c_uint64), ("point_index_base", C.c_int64), ("estimate_delta_t", C.c_float), ("reserved0", C.c_int32),
    ]


class dpi_train_config(C.Structure):
    _fields_ = [("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("loss_beta", C.c_float), ("fixed_weight", C.c_float)]


_lib = None


def load() -> C.CDLL:
    """Load libdpi_b200.so and declare the prototypes. Raises if the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -m deeppicarditeration_b200.build` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    H = C.c_void_p
    lib.dpi_version.restype = C.c_int
    lib.dpi_last_error.restype = C.c_char_p
    lib.dpi_create.argtypes = [C.c_int, C.POINTER(H)]
    lib.dpi_destroy.argtypes = [H]
    lib.dpi_set_problem.argtypes = [H, C.POINTER(dpi_problem)]
    lib.dpi_set_network.argtypes = [H, C.POINTER(dpi_network), C.c_void_p]
    lib.dpi_sample_tx.argtypes = [H, C.c_int64, C.c_int64, C.c_uint64, C.c_void_p, C.c_void_p]
    lib.dpi_label_gen.argtypes = [H, C.c_void_p, C.c_int64, C.POINTER(dpi_label_config), C.POINTER(dpi_draws),
                                  C.c_void_p, C.c_void_p]
    lib.dpi_label_gen_host.argtypes = [H, C.c_void_p, C.c_int64, C.POINTER(dpi_label_config), C.c_void_p, C.c_void_p,
                                       C.c_void_p]
    lib.dpi_sample_tx_ex.argtypes = [H, C.c_int64, C.c_int64, C.c_uint64, C.c_int32, C.c_float, C.c_void_p, C.c_void_p]
    lib.dpi_net_hessian_diag.argtypes = [H, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    lib.dpi_net_hessian.argtypes = [H, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    lib.dpi_train_dp_export.argtypes = [H, C.c_void_p]
    lib.dpi_train_dp_connect.argtypes = [H, C.c_int32, C.c_int32, C.c_void_p]
    lib.dpi_check_labels.argtypes = [H, C.c_void_p]
    lib.dpi_export_draws.argtypes = [H, C.c_int64, C.POINTER(dpi_label_config), C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]
    lib.dpi_net_eval.argtypes = [H, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.dpi_launch_count.argtypes = [H]
    lib.dpi_launch_count.restype = C.c_int64
    lib.dpi_set_stage_timing.argtypes = [H, C.c_int]
    lib.dpi_stage_ms.argtypes = [H, C.c_int, C.POINTER(C.c_float)]
    lib.dpi_selftest_umma.argtypes = [H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.dpi_train_init.argtypes = [H, C.POINTER(dpi_network), C.c_int64, C.c_void_p]
    lib.dpi_train_param_count.argtypes = [H]
    lib.dpi_train_param_count.restype = C.c_int64
    lib.dpi_train_grad.argtypes = [H, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(dpi_train_config), C.c_void_p,
                                   C.c_void_p, C.c_void_p]
    lib.dpi_train_adam.argtypes = [H, C.c_void_p, C.POINTER(dpi_train_config), C.c_int64, C.c_void_p]
    lib.dpi_train_reset.argtypes = [H, C.c_void_p]
    lib.dpi_train_fit.argtypes = [H, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.POINTER(dpi_train_config),
                                  C.c_int64, C.c_void_p, C.c_void_p]
    for name in EXPORTED_SYMBOLS:
        getattr(lib, name)  # AttributeError if a declared symbol is not exported
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().dpi_last_error().decode("utf-8", "replace")
        if rc == -3:
            raise NotImplementedError(f"{what}: {msg}")
        if rc == DPI_ERR_NONFINITE:
            raise FloatingPointError(f"{what}: {msg}")
        raise RuntimeError(f"{what} failed (status {rc}): {msg}")
