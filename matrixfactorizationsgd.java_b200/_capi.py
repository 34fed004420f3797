"""ctypes binding of libmfsgd.so -- the same symbols, in the same way, that the Java host binds with
Panama FFM (java/MatrixFactorizationSGDGpu.java). Declarations mirror include/mfsgd.h one to one.

There is no fallback: if the CUDA library is missing, importing this module raises."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmfsgd.so")

OK, E_INVALID_ARG, E_CUDA, E_NCCL, E_OOM, E_STATE = 0, -1, -2, -3, -4, -5
MODE_DETERMINISTIC, MODE_HOGWILD, MODE_DSGD = 0, 1, 2
SCATTER_STORE, SCATTER_ATOMIC, SCATTER_ATOMIC_Q, SCATTER_ATOMIC_P = 0, 1, 2, 3
FLAG_TIME_KERNELS, FLAG_VIRTUAL_RING, FLAG_NO_SHUFFLE, FLAG_EXACT_ARITH, FLAG_SPLIT_SHARDS = 1, 2, 4, 8, 16
FLAG_MATERIALIZE_SHUFFLE = 32
MODEL_GLOBAL_MEAN, MODEL_BIASES = 1, 2
STORAGE_F32, STORAGE_F16 = 0, 1
ABI_VERSION = 3


class Config(C.Structure):
    _fields_ = [("n_users", C.c_int32), ("n_items", C.c_int32), ("k", C.c_int32), ("lr", C.c_float),
                ("lambda_", C.c_float), ("init_scale", C.c_float), ("seed", C.c_uint64), ("mode", C.c_int32),
                ("n_gpus", C.c_int32), ("stripes_per_gpu", C.c_int32), ("shards_per_gpu", C.c_int32),
                ("scatter", C.c_int32), ("flags", C.c_uint32), ("device", C.c_int32), ("world_size", C.c_int32),
                ("rank", C.c_int32), ("nccl_id", C.c_uint8 * 128), ("ctas_per_sm", C.c_int32),
                ("rounds", C.c_int32), ("hot_share", C.c_float), ("hot_chunk", C.c_int32), ("merge_boost", C.c_float),
                ("model", C.c_uint32), ("p_atomic_threshold", C.c_float), ("lr_decay", C.c_float),
                ("early_stop_patience", C.c_int32), ("early_stop_min_delta", C.c_float), ("p_storage", C.c_int32),
                ("reserved", C.c_int32 * 2)]


class EpochStats(C.Structure):
    _fields_ = [("updates", C.c_int64), ("epoch_ms", C.c_double), ("shuffle_ms", C.c_double),
                ("update_kernel_ms", C.c_double), ("update_launches", C.c_int32), ("total_launches", C.c_int32),
                ("heldout_rmse", C.c_double), ("cold_ms", C.c_double), ("hot_ms", C.c_double),
                ("exchange_ms", C.c_double)]


class SynthParams(C.Structure):
    _fields_ = [("n_total", C.c_int64), ("seed", C.c_uint64), ("log2_alpha_user", C.c_int32),
                ("log2_alpha_item", C.c_int32), ("c_user", C.c_double), ("c_item", C.c_double),
                ("planted_amplitude", C.c_float), ("noise_scale", C.c_float)]


class LayoutInfo(C.Structure):
    _fields_ = [("n_gpus", C.c_int32), ("stripes_per_gpu", C.c_int32), ("shards_per_gpu", C.c_int32),
                ("user_blocks", C.c_int32), ("item_blocks", C.c_int32), ("n_train_local", C.c_int64),
                ("n_heldout_local", C.c_int64), ("n_train_total", C.c_int64), ("rounds", C.c_int32),
                ("n_hot_items", C.c_int32), ("n_heavy_users", C.c_int32), ("run_length", C.c_int32)]


class Ratings(C.Structure):
    _fields_ = [("users", C.POINTER(C.c_int32)), ("items", C.POINTER(C.c_int32)), ("ratings", C.POINTER(C.c_float)),
                ("n", C.c_int64), ("n_users", C.c_int32), ("n_items", C.c_int32), ("user_ids", C.POINTER(C.c_int64)),
                ("item_ids", C.POINTER(C.c_int64)), ("format", C.c_int32), ("reserved", C.c_int32)]


class Ceilings(C.Structure):
    _fields_ = [("row_gather_scatter_gbs", C.c_double), ("row_gather_only_gbs", C.c_double), ("hbm_stream_copy_gbs", C.c_double),
                ("buffer_mb", C.c_double), ("l2_mb", C.c_double), ("sm_count", C.c_int32), ("reserved", C.c_int32)]


FORMAT_AUTO, FORMAT_TRIPLETS, FORMAT_NETFLIX_PRIZE = 0, 1, 2

_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float

# name -> (restype, argtypes): every symbol include/mfsgd.h declares
SIGNATURES = {
    "mfsgd_abi_version": (C.c_int, []),
    "mfsgd_last_error": (C.c_char_p, []),
    "mfsgd_device_count": (C.c_int, [C.POINTER(_i32)]),
    "mfsgd_config_default": (C.c_int, [C.POINTER(Config)]),
    "mfsgd_create": (C.c_int, [C.POINTER(Config), C.POINTER(_vp)]),
    "mfsgd_destroy": (None, [_vp]),
    "mfsgd_plan_layout": (C.c_int, [C.POINTER(Config), _i64, _i64, _i32, _i64, _i32, C.POINTER(_i32), C.POINTER(_i32),
                                    C.POINTER(_i32), C.POINTER(_i32)]),
    "mfsgd_plan_runs": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp, _i32, _i32, C.c_uint64, _i32, _f32, _vp, _vp, _vp, _vp,
                                  C.POINTER(_i64), _vp]),
    "mfsgd_read_ratings": (C.c_int, [C.c_char_p, _i32, C.POINTER(Ratings)]),
    "mfsgd_free_ratings": (None, [C.POINTER(Ratings)]),
    "mfsgd_load_ratings": (C.c_int, [_vp, _vp, _vp, _vp, _i64]),
    "mfsgd_load_ratings_sharded": (C.c_int, [_vp, _vp, _vp, _vp, _i64]),
    "mfsgd_load_heldout": (C.c_int, [_vp, _vp, _vp, _vp, _i64]),
    "mfsgd_generate_synthetic": (C.c_int, [_vp, C.POINTER(SynthParams), C.POINTER(_i64), C.POINTER(_i64)]),
    "mfsgd_init_factors": (C.c_int, [_vp]),
    "mfsgd_set_factors": (C.c_int, [_vp, _vp, _vp]),
    "mfsgd_get_factors": (C.c_int, [_vp, _vp, _vp]),
    "mfsgd_get_model": (C.c_int, [_vp, C.POINTER(C.c_float), _vp, _vp]),
    "mfsgd_set_biases": (C.c_int, [_vp, _vp, _vp]),
    "mfsgd_get_progress": (C.c_int, [_vp, C.POINTER(_i32), C.POINTER(C.c_float), C.POINTER(_i32)]),
    "mfsgd_get_partition": (C.c_int, [_vp] + [C.POINTER(_i32)] * 4),
    "mfsgd_train": (C.c_int, [_vp, _i32, C.POINTER(EpochStats)]),
    "mfsgd_train_traced": (C.c_int, [_vp, _i32, C.POINTER(EpochStats), _vp]),
    "mfsgd_set_eval_every_epoch": (C.c_int, [_vp, _i32]),
    "mfsgd_rmse": (C.c_int, [_vp, _vp, _vp, _vp, _i64, C.POINTER(C.c_double)]),
    "mfsgd_rmse_heldout": (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_i64)]),
    "mfsgd_rmse_train": (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_i64)]),
    "mfsgd_factorize": (C.c_int, [_vp, _vp, _vp, _i64, C.POINTER(Config), _i32, _vp, _vp]),
    "mfsgd_get_layout_info": (C.c_int, [_vp, C.POINTER(LayoutInfo)]),
    "mfsgd_get_bounds": (C.c_int, [_vp, _vp, _vp]),
    "mfsgd_get_records": (C.c_int, [_vp, _i32, _vp, _vp, C.POINTER(_i64)]),
    "mfsgd_shuffle_once": (C.c_int, [_vp, _i32]),
    "mfsgd_apply_updates_forced": (C.c_int, [_i32, _i32, _f32, _f32, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mfsgd_generate_to_host": (C.c_int, [_i32, C.POINTER(SynthParams), _i32, _i32, _i64, _i64, _vp, _vp, _vp, _vp]),
    "mfsgd_nccl_unique_id": (C.c_int, [_vp]),
    "mfsgd_measure_ceilings": (C.c_int, [_i32, C.c_double, C.POINTER(Ceilings)]),
    "mfsgd_release_cached_memory": (C.c_int, []),
    "mfsgd_host_alloc": (C.c_int, [C.POINTER(_vp), _i64]),
    "mfsgd_host_free": (C.c_int, [_vp]),
}


class MfsgdError(RuntimeError):
    """Non-zero return of an mfsgd_* call (the Java host raises IllegalStateException here)."""

    def __init__(self, code, message):
        super().__init__("mfsgd error %d: %s" % (code, message))
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError("libmfsgd.so is not built (%s missing): run `make` or __graft_entry__.build(); "
                          "there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.mfsgd_abi_version() != ABI_VERSION:
        raise ImportError("libmfsgd.so ABI %d != binding ABI %d" % (lib.mfsgd_abi_version(), ABI_VERSION))
    return lib


lib = _load()


def check(rc):
    if rc != OK:
        raise MfsgdError(rc, (lib.mfsgd_last_error() or b"").decode("utf-8", "replace"))


def ptr(a):
    return None if a is None else a.ctypes.data


def as_i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def as_f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)
