"""ctypes binding of libd4s.so (include/d4s.h) -- the only route from Python to the CUDA kernels.

There is no CPU fallback: if the library cannot be loaded, or no CUDA device is visible, every compute
entry point raises `D4sError`.  torch is used only to own device memory and streams; the ABI sees raw
pointers (`tensor.data_ptr()`) and the current stream handle.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("D4S_LIB_PATH") or os.path.join(HERE, "libd4s.so")  # (override: A/B builds of the kernels)

ABI_VERSION = 15
MAX_LAYERS = 8
MAX_THETA = 10
MAX_THETA_WIDE = 32
MAX_WIDTH = 256
SCHED_STRIDE = 16
MAX_RANKS = 8
IPC_HANDLE_BYTES = 64

COV_GAUSS, COV_JAC = 0, 1
PRIOR_NONE, PRIOR_GAUSSIAN, PRIOR_UNIFORM, PRIOR_EXTERNAL = 0, 1, 2, 3
COMBINE_SUM, COMBINE_SHARED, COMBINE_PER_SAMPLE = 0, 1, 2

# schedule table columns (must match D4S_S_* in include/d4s.h)
S_T, S_ALPHA, S_SIGMA, S_INV_SIGMA, S_A_OVER_S2, S_C_THETA, S_C_SCORE, S_BRIDGE_STD, S_SQRT_ALPHA, S_COMBINE, \
    S_ONE_MINUS_N, S_S1_WEIGHT, S_TAME, S_CLIP = range(14)

EXPORTS = [
    "d4s_abi_version", "d4s_last_error", "d4s_device_count", "d4s_net_create", "d4s_net_free",
    "d4s_net_h1_padded", "d4s_workspace_floats", "d4s_partials_floats", "d4s_mats_floats", "d4s_score_partials", "d4s_finalize",
    "d4s_ddim_run", "d4s_graph_cache_clear", "d4s_ddim_run_sharded", "d4s_xchg_layout", "d4s_xchg_alloc", "d4s_xchg_free",
    "d4s_xchg_open", "d4s_xchg_close", "d4s_set_option", "d4s_get_profile", "d4s_prior_score", "d4s_ddim_update", "d4s_batched_inverse", "d4s_philox_normal", "d4s_launch_count", "d4s_reduce_pairs", "d4s_finalize_wide",
]


class D4sError(RuntimeError):
    pass


class NetDesc(C.Structure):
    _fields_ = [
        ("n_layers", C.c_int32),
        ("dims", C.c_int32 * (MAX_LAYERS + 1)),
        ("theta_cols", C.c_int32),
        ("weight", C.c_void_p * MAX_LAYERS),
        ("bias", C.c_void_p * MAX_LAYERS),
        ("emb_layers", C.c_int32),
        ("emb_dims", C.c_int32 * (MAX_LAYERS + 1)),
        ("emb_weight", C.c_void_p * MAX_LAYERS),
        ("emb_bias", C.c_void_p * MAX_LAYERS),
        ("out_act", C.c_int32),
        ("out_scale", C.c_float),
    ]


class Problem(C.Structure):
    _fields_ = [
        ("n_samples", C.c_int32),
        ("n_obs", C.c_int32),
        ("theta_dim", C.c_int32),
        ("cov_mode", C.c_int32),
        ("prior_kind", C.c_int32),
        ("paired", C.c_int32),
        ("obs_chunk", C.c_int32),
        ("n_obs_total", C.c_int32),
        ("obs_feat", C.c_void_p),
        ("obs_prec", C.c_void_p),
        ("prior_low", C.c_void_p),
        ("prior_high", C.c_void_p),
        ("clip_lo", C.c_float),
        ("clip_hi", C.c_float),
        ("seed", C.c_uint64),
        ("sample_offset", C.c_int64),
        ("workspace", C.c_void_p),
        ("workspace_floats", C.c_int64),
        ("obs_raw", C.c_void_p),
        ("x_dim", C.c_int32),
        ("ctx_samples", C.c_int32),
        ("obs_weight", C.c_void_p),
        ("ctx_qsum", C.c_void_p),
    ]


class Step(C.Structure):
    _fields_ = [
        ("sched", C.c_void_p),
        ("time_feat", C.c_void_p),
        ("mats", C.c_void_p),
        ("noise", C.c_void_p),
        ("step_index", C.c_int32),
        ("update", C.c_int32),
        ("affine", C.c_void_p),
    ]


class Exchange(C.Structure):
    _fields_ = [
        ("world", C.c_int32),
        ("rank", C.c_int32),
        ("seq_base", C.c_uint32),
        ("reserved", C.c_int32),
        ("data", C.c_void_p * MAX_RANKS),
        ("flags", C.c_void_p * MAX_RANKS),
    ]


_lib = None
_lock = threading.Lock()


def load():
    """Loads libd4s.so (built in-tree by d4s._build / __graft_entry__.build()).  Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise D4sError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  diffusions-for-sbi_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
        sig = {
            "d4s_abi_version": (C.c_int, []),
            "d4s_last_error": (C.c_char_p, []),
            "d4s_device_count": (C.c_int, []),
            "d4s_net_create": (C.c_int, [C.POINTER(vp), C.POINTER(NetDesc), vp]),
            "d4s_net_free": (None, [vp]),
            "d4s_net_h1_padded": (C.c_int, [vp]),
            "d4s_workspace_floats": (i64, [vp, C.POINTER(Problem)]),
            "d4s_partials_floats": (i64, [vp, C.POINTER(Problem)]),
            "d4s_mats_floats": (i32, [i32]),
            "d4s_score_partials": (C.c_int, [vp, C.POINTER(Problem), C.POINTER(Step), vp, vp, vp, vp]),
            "d4s_finalize": (C.c_int, [vp, C.POINTER(Problem), C.POINTER(Step), vp, vp, vp]),
            "d4s_ddim_run": (C.c_int, [vp, C.POINTER(Problem), i32, vp, vp, vp, vp, vp, C.POINTER(i32), vp, i32, vp]),
            "d4s_graph_cache_clear": (None, []),
            "d4s_ddim_run_sharded": (C.c_int, [vp, C.POINTER(Problem), i32, vp, vp, vp, vp, vp, C.POINTER(Exchange), vp]),
            "d4s_xchg_layout": (C.c_int, [vp, C.POINTER(Problem), i32, C.POINTER(i64), C.POINTER(i64)]),
            "d4s_xchg_alloc": (C.c_int, [i64, C.POINTER(vp), C.c_char_p]),
            "d4s_xchg_free": (C.c_int, [vp]),
            "d4s_xchg_open": (C.c_int, [C.c_char_p, C.POINTER(vp)]),
            "d4s_xchg_close": (C.c_int, [vp]),
            "d4s_set_option": (C.c_int, [C.c_char_p, i32]),
            "d4s_get_profile": (C.c_int, [C.POINTER(i64)]),
            "d4s_prior_score": (C.c_int, [i32, vp, i64, i32, vp, vp, vp, vp, vp, vp, vp]),
            "d4s_ddim_update": (C.c_int, [vp, vp, i64, i32, vp, vp, u64, i32, i64, C.c_float, C.c_float, vp]),
            "d4s_batched_inverse": (C.c_int, [vp, vp, i32, i32, vp]),
            "d4s_philox_normal": (C.c_int, [vp, i64, i32, u64, i32, i64, vp]),
            "d4s_launch_count": (i64, []),
            "d4s_reduce_pairs": (C.c_int, [vp, vp, vp, vp, i64, i32, i32, vp, vp]),
            "d4s_finalize_wide": (C.c_int, [vp, i64, i32, i32, vp, vp, i32, vp, vp, vp, vp, vp, u64, i32, i64, C.c_float,
                                            C.c_float, i32, vp, vp, vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.d4s_abi_version() != ABI_VERSION:
            raise D4sError(f"libd4s ABI {lib.d4s_abi_version()} != binding {ABI_VERSION}: rebuild the library")
        if os.environ.get("D4S_JAC_TC", "1") == "0":  # opt out of the tensor-core JAC kernel for the whole process (fp32 SIMT instead)
            lib.d4s_set_option(b"jac_tc", 0)
            _options["jac_tc"] = 0
        _lib = lib
    return _lib


_options: dict = {}


def set_option(key: str, value: int):
    """Library switches (see d4s_set_option in include/d4s.h): path, tc_steps_per_launch, profile, tc_pairing, jac_tc."""
    check(load().d4s_set_option(key.encode(), int(value)))
    _options[key] = int(value)


def options_key() -> tuple:
    """Current option state: part of every cache key whose value depends on the kernel path (workspace layouts do)."""
    return tuple(sorted(_options.items()))


PROFILE_PHASES = ["layer0_sample_part", "layer0", "prior_noise", "tensor_wait", "epilogues", "scores",
                  "prec_and_sums", "finalize", "mma_inner_layers", "mma_last_layer",  # 8, 9: MMA issue -> commit (MMA warp)
                  "tail", "fused_sums_tree"]


def get_profile():
    """Phase cycle counters of the last profiled tcgen05 launch (CTA 0): dict phase -> cycles."""
    buf = (C.c_int64 * 16)()
    check(load().d4s_get_profile(buf))
    return {name: int(buf[i]) for i, name in enumerate(PROFILE_PHASES)}


def check(rc: int):
    if rc != 0:
        raise D4sError(f"libd4s error {rc}: {load().d4s_last_error().decode()}")


def require_cuda():
    lib = load()
    if lib.d4s_device_count() < 1:
        raise D4sError("no CUDA device visible: diffusions-for-sbi_b200 runs on sm_100a only (no CPU fallback)")
    return lib
