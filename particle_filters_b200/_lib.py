"""ctypes binding of libsmc.so (the C-ABI declared in include/smc.h).

There is no CPU fallback: if the shared library is missing or does not match the
header this module raises, and every compute entry point needs a CUDA device.
"""

from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SMC_LIB_PATH", os.path.join(HERE, "_lib", "libsmc.so"))   # override: kernel tuning sweeps

SMC_ABI_VERSION = 2
SMC_F32, SMC_F64 = 0, 1
SMC_MODEL_STOCH_VOL, SMC_MODEL_LINEAR_GAUSSIAN, SMC_MODEL_LORENZ96 = 1, 2, 3
SMC_MODE_PROPAGATE, SMC_MODE_WEIGHT, SMC_MODE_STEP = 1, 2, 3
SMC_FLAG_WRITE_ANCESTORS, SMC_FLAG_NO_MOMENTS, SMC_FLAG_GENERIC_MOMENTS = 1, 2, 4
SMC_FLAG_POST_MOMENTS_GENERIC = 8
SMC_STATS_HEADER = 4


def stats_stride(nx: int) -> int:
    return SMC_STATS_HEADER + nx + nx * nx


class SirDesc(C.Structure):
    """Mirror of ``smc_sir_desc`` (include/smc.h)."""

    _fields_ = [
        ("n", C.c_int64), ("n_global", C.c_int64), ("offset", C.c_int64), ("ld", C.c_int64),
        ("nx", C.c_int32), ("nz", C.c_int32), ("dtype", C.c_int32), ("flags", C.c_int32),
        ("x0_dev", C.c_void_p), ("x1_dev", C.c_void_p), ("logw_dev", C.c_void_p),
        ("ancestors_dev", C.c_void_p), ("ws_dev", C.c_void_p), ("ws_bytes", C.c_uint64),
    ]


class LedhDesc(C.Structure):
    """Mirror of ``smc_ledh_desc`` (include/smc_ledh.h)."""

    _fields_ = [
        ("n", C.c_int64), ("ld", C.c_int64),
        ("nx", C.c_int32), ("nz", C.c_int32), ("dtype", C.c_int32), ("n_lambda_steps", C.c_int32),
        ("eta0_dev", C.c_void_p), ("eta_dev", C.c_void_p), ("theta_log_dev", C.c_void_p),
        ("s0_dev", C.c_void_p),
    ]


class ShardCtx(C.Structure):
    """Mirror of ``smc_shard_ctx`` (include/smc.h)."""

    _fields_ = [("mail_dev", C.c_void_p), ("peer_x_dev", C.c_void_p), ("rank", C.c_int32), ("n_shards", C.c_int32),
                ("stamp", C.c_uint64), ("peer_aux_dev", C.c_void_p), ("rsum_area_dev", C.c_void_p),
                ("mailbox_dev", C.c_void_p)]


_P = C.c_void_p
_D = C.POINTER(C.c_double)
_SIG = {
    # name: (restype, argtypes)
    "smc_abi_version": (C.c_int, []),
    "smc_launch_count": (C.c_uint64, []),
    "smc_last_error": (C.c_char_p, []),
    "smc_sizeof_sir_desc": (C.c_uint64, []),
    "smc_sir_workspace_bytes": (C.c_uint64, [C.c_int64, C.c_int32, C.c_int32]),
    "smc_sir_reset": (C.c_int, [C.POINTER(SirDesc), _D, _P]),
    "smc_sir_init_gaussian": (C.c_int, [C.POINTER(SirDesc), _D, _D, C.c_uint64, _P]),
    "smc_sir_propagate_weight": (C.c_int, [C.POINTER(SirDesc), C.c_int32, _D, C.c_int32, C.c_int32,
                                           _P, _P, C.c_uint64, C.c_int64, _P]),
    "smc_bridge_predict": (C.c_int, [C.POINTER(SirDesc), _P, _P, _D, C.c_double, C.c_int32, _P]),
    "smc_bridge_loglik": (C.c_int, [C.POINTER(SirDesc), _P, _D, _D, _P, _P]),
    "smc_sir_jitter": (C.c_int, [C.POINTER(SirDesc), _D, C.c_double, C.c_uint64, C.c_int64, _P]),
    "smc_sir_finalize": (C.c_int, [C.POINTER(SirDesc), C.c_double, C.c_double, C.c_uint64, C.c_int64,
                                   _P, _P, C.c_int32, _P]),
    "smc_sir_set_uniform": (C.c_int, [C.POINTER(SirDesc), C.c_double, _P]),
    "smc_partial_stride": (C.c_int64, [C.c_int32]),
    "smc_sir_reduce_partials": (C.c_int, [C.POINTER(SirDesc), _P, _P]),
    "smc_sir_resample": (C.c_int, [C.POINTER(SirDesc), C.c_uint64, _P, _P, C.c_int32, _P]),
    "smc_sir_shard_weight_total": (C.c_int, [C.POINTER(SirDesc), _P, _P]),
    "smc_sir_resample_sharded": (C.c_int, [C.POINTER(SirDesc), _P, C.c_int32, C.c_int32, _P, _P]),
    "smc_sir_reduce_moment_partials": (C.c_int, [C.POINTER(SirDesc), _P, _P]),
    "smc_dev_alloc": (C.c_int, [C.c_uint64, C.POINTER(C.c_void_p)]),
    "smc_dev_free": (C.c_int, [_P]),
    "smc_ipc_export": (C.c_int, [_P, C.c_char_p]),
    "smc_ipc_import": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "smc_ipc_close": (C.c_int, [_P]),
    "smc_sir_finish_step": (C.c_int, [C.POINTER(SirDesc), _P, _P, C.c_int32, _P]),
    "smc_sir_step": (C.c_int, [C.POINTER(SirDesc), C.c_int32, _D, C.c_int32, _P, _P, C.c_double,
                               C.c_double, C.c_uint64, C.c_int64, _P, _P]),
    "smc_sir_step_weight": (C.c_int, [C.POINTER(SirDesc), C.c_int32, _D, C.c_int32, _P, _P, C.c_double,
                                      C.c_double, C.c_uint64, C.c_int64, _P, _P]),
    "smc_sir_step_resample": (C.c_int, [C.POINTER(SirDesc), C.c_int32, _P, _P]),
    "smc_sir_step_fused": (C.c_int, [C.POINTER(SirDesc), _P, C.c_int32, _D, C.c_int32, _P, C.c_double, C.c_uint64,
                                     C.c_int64, _P, _P, C.c_int32, _P]),
    "smc_mailbox_bytes": (C.c_uint64, []),
    "smc_mailbox_error_offset": (C.c_uint64, []),
    "smc_sir_step_sharded": (C.c_int, [C.POINTER(SirDesc), C.POINTER(ShardCtx), C.c_int32, _D, C.c_int32, _P,
                                       C.c_double, C.c_uint64, C.c_int64, _P, _P]),
    "smc_ledh_info_counters": (None, [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "smc_sir_step_host": (C.c_int, [C.POINTER(SirDesc), C.c_int32, _D, C.c_int32, _P, _P, C.c_double, C.c_double,
                                    C.c_uint64, C.c_int64, _P, _P, C.c_int32, _P]),
    "smc_sizeof_shard_ctx": (C.c_uint64, []),
    "smc_shard_rsum_bytes": (C.c_uint64, []),
    "smc_sir_weight_sharded": (C.c_int, [C.POINTER(SirDesc), C.POINTER(ShardCtx), C.c_int32, _D, C.c_int32, _P,
                                         C.c_double, C.c_uint64, C.c_int64, _P]),
    "smc_sir_step_fused_sharded": (C.c_int, [C.POINTER(SirDesc), C.POINTER(ShardCtx), C.c_int32, C.c_int32, C.c_int32, _D,
                                             C.c_int32, _P, C.c_double, C.c_uint64, C.c_int64, _P, _P, _P]),
    "smc_sir_fused_sharded_flush": (C.c_int, [C.POINTER(SirDesc), C.POINTER(ShardCtx), C.c_double, C.c_uint64, C.c_int64,
                                              _P, _P, C.c_int32, _P]),
    "smc_systematic_ancestors": (C.c_int, [_P, C.c_int64, C.c_double, _P, _P, C.c_uint64, _P]),
    "smc_multinomial_workspace_bytes": (C.c_uint64, [C.c_int64]),
    "smc_multinomial_ancestors": (C.c_int, [_P, C.c_int64, C.c_int64, _P, C.c_uint64, C.c_int64, _P, _P,
                                            C.c_uint64, _P]),
    "smc_gather": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int64, C.c_int32, C.c_int32, _P]),
    "smc_sir_export_weights": (C.c_int, [C.POINTER(SirDesc), _P, _P]),
    "smc_sir_current_buffer": (C.c_int, [C.POINTER(SirDesc), C.POINTER(C.c_int32), _P]),
}

_LEDH_SIG = {
    "smc_sizeof_ledh_desc": (C.c_uint64, []),
    "smc_ledh_flow": (C.c_int, [C.POINTER(LedhDesc), _D, _D, _D, C.c_int32, _D, C.c_int32, _P]),
    "smc_ledh_flow_step_supplied": (C.c_int, [C.POINTER(LedhDesc), _D, _D, _D, _P, _P, C.c_double,
                                              C.c_double, C.c_int32, _P]),
    "smc_ledh_propagate_linear": (C.c_int, [C.POINTER(LedhDesc), _P, _D, _D, _P, C.c_uint64, C.c_int64, _P]),
    "smc_ledh_gauss_terms": (C.c_int, [C.POINTER(LedhDesc), _P, _D, _D, _D, _D, C.c_int32, _D, C.c_int32, _P, _P]),
    "smc_ledh_weights": (C.c_int, [C.POINTER(SirDesc), _P, _P, _P, _P]),
    "smc_weighted_moments": (C.c_int, [C.POINTER(SirDesc), _P, _P, _P]),
    "smc_affine_map": (C.c_int, [_P, _D, _D, C.c_int64, C.c_int64, C.c_int32, C.c_int32, _P]),
}
SMC_LEDH_H_LINEAR, SMC_LEDH_H_ACOUSTIC, SMC_LEDH_H_EXP = 1, 2, 3


class SmcError(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """Load libsmc.so; raise (never fall back) when it is absent or mismatched."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SmcError(
            f"{LIB_PATH} not found: build it with `python -m particle_filters_b200.build` "
            "(there is no CPU fallback for the SMC engine)")
    lib = C.CDLL(LIB_PATH)
    for table in (_SIG, _LEDH_SIG):
        for name, (res, args) in table.items():
            if not hasattr(lib, name):
                raise SmcError(f"libsmc.so does not export {name}")
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
    if lib.smc_abi_version() != SMC_ABI_VERSION:
        raise SmcError("libsmc.so ABI version mismatch; rebuild")
    if lib.smc_sizeof_sir_desc() != C.sizeof(SirDesc):
        raise SmcError("smc_sir_desc layout mismatch between include/smc.h and _lib.py")
    if lib.smc_sizeof_shard_ctx() != C.sizeof(ShardCtx):
        raise SmcError("smc_shard_ctx layout mismatch between include/smc.h and _lib.py")
    if lib.smc_sizeof_ledh_desc() != C.sizeof(LedhDesc):
        raise SmcError("smc_ledh_desc layout mismatch between include/smc_ledh.h and _lib.py")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise SmcError(load().smc_last_error().decode(errors="replace"))


def as_doubles(a):
    """Host float64 array -> (keepalive, double*) for params passed by pointer."""
    import numpy as np

    arr = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
    return arr, arr.ctypes.data_as(_D)
