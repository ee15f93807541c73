"""ctypes binding of ``libertdiff_b200.so`` (C ABI: ``include/ertdiff_b200.h``).

There is no CPU fallback anywhere in this package: if the library is missing the import
raises, and every compute entry point fails loudly when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ERTDIFF_B200_LIB: development override (e.g. the -DUC_TIMING=1 build of scripts/gpu_umma_quick.sh)
LIB_PATH = os.environ.get("ERTDIFF_B200_LIB") or os.path.join(_HERE, "libertdiff_b200.so")

F32, F64 = 0, 1
LOOP_PERSISTENT, LOOP_GRAPH, LOOP_STREAM = 0, 1, 2
PREC_FP32, PREC_BF16, PREC_BF16X3 = 0, 1, 2
LOOP_MODES = {"persistent": LOOP_PERSISTENT, "graph": LOOP_GRAPH, "stream": LOOP_STREAM}
PRECISIONS = {"fp32": PREC_FP32, "bf16": PREC_BF16, "bf16x3": PREC_BF16X3}


class ErtdiffError(RuntimeError):
    pass


class ChainArgs(C.Structure):
    """``ertdiff_chain_args`` -- keep in sync with include/ertdiff_b200.h."""
    _fields_ = [
        ("B", C.c_int64), ("n_cond", C.c_int64), ("T", C.c_int32), ("num_steps", C.c_int32),
        ("temperature", C.c_double),
        ("d_betas", C.c_void_p), ("d_alphas", C.c_void_p), ("d_alpha_bar", C.c_void_p),
        ("d_cond_bias", C.c_void_p), ("d_x_T", C.c_void_p), ("d_noise", C.c_void_p),
        ("seed", C.c_uint64), ("offset", C.c_uint64), ("member_offset", C.c_int64),
        ("noise_member_stride_B", C.c_int64), ("loop_mode", C.c_int32), ("precision", C.c_int32),
        ("d_x_out", C.c_void_p), ("d_eps_trace", C.c_void_p),
    ]


# name -> (restype, argtypes); every symbol include/ertdiff_b200.h declares
SIGNATURES = {
    "ertdiff_abi_version": (C.c_int, []),
    "ertdiff_last_error": (C.c_char_p, []),
    "ertdiff_launch_count": (C.c_int64, []),
    "ertdiff_launch_count_reset": (None, []),
    "ertdiff_model_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int]),
    "ertdiff_model_destroy": (C.c_int, [C.c_void_p]),
    "ertdiff_model_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "ertdiff_model_last_chain_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "ertdiff_model_umma_status": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "ertdiff_debug_umma_timing": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]),
    "ertdiff_model_load": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_void_p]),
    "ertdiff_model_export": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_void_p]),
    "ertdiff_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                  C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "ertdiff_encode_condition": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                           C.c_void_p, C.c_void_p, C.c_void_p]),
    "ertdiff_encode_condition_prec": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                           C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "ertdiff_sample_chain": (C.c_int, [C.c_void_p, C.POINTER(ChainArgs), C.c_void_p]),
    "ertdiff_sample_model": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                       C.POINTER(ChainArgs), C.c_void_p]),
    "ertdiff_step_coefficients": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                            C.c_double, C.c_void_p, C.c_void_p]),
    "ertdiff_philox_normal": (C.c_int, [C.c_uint64, C.c_uint64, C.c_int64, C.c_int64, C.c_int32,
                                        C.c_int32, C.c_void_p, C.c_void_p]),
    "ertdiff_posterior_update": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                           C.c_float, C.c_int64, C.c_void_p, C.c_void_p]),
    "ertdiff_ensemble_moments": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p]),
    "ertdiff_ensemble_percentiles": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64,
                                               C.POINTER(C.c_double), C.c_int32, C.c_int,
                                               C.c_void_p, C.c_void_p]),
    "ertdiff_minmax": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]),
    "ertdiff_ensemble_kde_mode": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p,
                                            C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ertdiff_ensemble_kde_mode_auto": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int32, C.c_void_p,
                                                 C.c_void_p, C.c_void_p, C.c_void_p]),
    "ertdiff_interval_coverage": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int32,
                                            C.c_void_p, C.c_void_p]),
    "ertdiff_misfit_metrics": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int64,
                                         C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ertdiff_wasserstein_distance": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int64,
                                               C.c_void_p, C.c_void_p]),
    "ertdiff_ensemble_summary": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                           C.POINTER(C.c_double), C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64,
                                           C.c_void_p]),
    "ertdiff_pack_rows_f64": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.c_int32, C.c_int64, C.c_int64,
                                        C.c_void_p, C.c_void_p]),
    "ertdiff_check_bounds": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p]),
    "ertdiff_argsort_stable": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]),
    "ertdiff_peer_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_void_p]),
    "ertdiff_peer_connect": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ertdiff_peer_all_gather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.POINTER(C.c_void_p), C.c_void_p]),
    "ertdiff_peer_status": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "ertdiff_peer_destroy": (C.c_int, [C.c_void_p]),
    "ertdiff_debug_umma_gemm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "ertdiff_debug_chain_floor": (C.c_int, [C.c_void_p, C.c_int]),
    "ertdiff_debug_graph_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "ertdiff_untransform_bounds": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_float,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
}

_lib = None


def load():
    """Load the shared library (once) and attach the prototypes.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ErtdiffError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `make -C ert-conditional-diffusion-model_b200/csrc`. "
            "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.ertdiff_abi_version() != 1:
        raise ErtdiffError("libertdiff_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().ertdiff_last_error()
        raise ErtdiffError(f"{what or 'ertdiff call'} failed ({rc}): "
                           f"{msg.decode() if msg else 'unknown error'}")


def ptr(t):
    """Device (or host) address of a torch tensor / None."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
