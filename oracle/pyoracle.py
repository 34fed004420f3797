"""ctypes binding of liboracle.so (CPU oracle -- TEST INFRASTRUCTURE, see oracle.cpp header).

Importers allowed: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline / --impl reference).
The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_DIR, "liboracle.so")

ORDER_SEQ, ORDER_WARP_TREE, ORDER_WARP_TREE_FMA, ORDER_WARP_TREE_FMA_PDELTA = 0, 1, 2, 3

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build():
    subprocess.check_call(["make", "-s", "-C", _DIR])


def _load():
    if not os.path.exists(_LIB):
        build()
    lib = C.CDLL(_LIB)
    lib.orc_hash64.restype = C.c_uint64
    lib.orc_hash64.argtypes = [C.c_uint64] * 3
    lib.orc_uniform.restype = C.c_float
    lib.orc_uniform.argtypes = [C.c_uint64] * 3
    lib.orc_default_init_scale.restype = C.c_float
    lib.orc_default_init_scale.argtypes = [C.c_int]
    lib.orc_init_factors.restype = None
    lib.orc_init_factors.argtypes = [_f32p, C.c_int64, C.c_int, C.c_uint64, C.c_uint64, C.c_float]
    lib.orc_shuffle.restype = None
    lib.orc_shuffle.argtypes = [C.c_uint64, C.c_int, C.c_int, _i32p]
    lib.orc_shuffle_mt.restype = None
    lib.orc_shuffle_mt.argtypes = [C.c_uint64, C.c_int, C.c_int, _i32p, C.c_int]
    lib.orc_sgd_update.restype = C.c_float
    lib.orc_sgd_update.argtypes = [_f32p, _f32p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int]
    lib.orc_train.restype = C.c_int
    lib.orc_train.argtypes = [_i32p, _i32p, _f32p, C.c_int64, _f32p, _f32p, C.c_int, C.c_int, C.c_int,
                              C.c_float, C.c_float, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_void_p]
    lib.orc_train_tape.restype = C.c_int
    lib.orc_train_tape.argtypes = [_i32p, _i32p, _f32p, C.c_int64, _f32p, _f32p, C.c_int, C.c_int, C.c_int,
                                   C.c_float, C.c_float, C.c_int, C.c_uint64, C.c_int, _i32p, _f32p, _f32p,
                                   _f32p, _f32p, _f32p]
    lib.orc_factorize.restype = C.c_int
    lib.orc_factorize.argtypes = [_i32p, _i32p, _f32p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_float,
                                  C.c_float, C.c_int, C.c_uint64, C.c_int, _f32p, _f32p]
    lib.orc_train_hogwild.restype = C.c_int
    lib.orc_train_hogwild.argtypes = [_i32p, _i32p, _f32p, C.c_int64, _f32p, _f32p, C.c_int, C.c_int, C.c_int,
                                      C.c_float, C.c_float, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int,
                                      C.POINTER(C.c_double)]
    lib.orc_rmse.restype = C.c_double
    lib.orc_rmse.argtypes = [_f32p, _f32p, C.c_int, _i32p, _i32p, _f32p, C.c_int64, C.c_int]
    lib.orc_skewed_rank.restype = C.c_int32
    lib.orc_skewed_rank.argtypes = [C.c_double, C.c_int32, C.c_int, C.c_double]
    lib.orc_scatter_id.restype = C.c_int32
    lib.orc_scatter_id.argtypes = [C.c_int32, C.c_int32]
    lib.orc_generate.restype = None
    lib.orc_generate.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_double,
                                 C.c_int, C.c_double, _i32p, _i32p, _f32p, _u8p, C.c_int]
    lib.orc_generate2.restype = None
    lib.orc_generate2.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_double,
                                  C.c_int, C.c_double, C.c_float, C.c_float, _i32p, _i32p, _f32p, _u8p, C.c_int]
    lib.orc_hardware_threads.restype = C.c_int
    _i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
    _u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
    lib.orc_bucket_perm_key.restype = C.c_uint64
    lib.orc_bucket_perm_key.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32]
    lib.orc_block_perm.restype = C.c_uint64
    lib.orc_block_perm.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
    lib.orc_block_perm_fill.restype = None
    lib.orc_block_perm_fill.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, _i64p]
    lib.orc_train_runs_launch.restype = C.c_int
    lib.orc_train_runs_launch.argtypes = [_i32p, _f32p, C.c_int64, _i64p, _i32p, _i32p, _f32p, _i64p, _i32p, _u32p, C.c_int64,
                                          C.c_int, C.c_uint64, C.c_uint32, _f32p, C.c_int32, _f32p, C.c_int32, C.c_int,
                                          C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.orc_global_mean.restype = C.c_float
    lib.orc_global_mean.argtypes = [_f32p, C.c_int64]
    lib.orc_train_model.restype = C.c_int
    lib.orc_train_model.argtypes = [_i32p, _i32p, _f32p, C.c_int64, _f32p, _f32p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                    C.c_float, C.c_float, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int]
    lib.orc_sgd_update_model.restype = C.c_float
    lib.orc_sgd_update_model.argtypes = [_f32p, _f32p, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int]
    lib.orc_rmse_model.restype = C.c_double
    lib.orc_rmse_model.argtypes = [_f32p, _f32p, C.c_void_p, C.c_void_p, C.c_int, _i32p, _i32p, _f32p, C.c_int64, C.c_int]
    lib.orc_learning_rate.restype = C.c_float
    lib.orc_learning_rate.argtypes = [C.c_float, C.c_float, C.c_int]
    lib.orc_train_early_stop.restype = C.c_int
    lib.orc_train_early_stop.argtypes = [_i32p, _i32p, _f32p, C.c_int64, _i32p, _i32p, _f32p, C.c_int64, _f32p, _f32p, C.c_void_p, C.c_void_p,
                                         C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_float, C.c_int, C.c_uint64,
                                         C.c_int, C.POINTER(C.c_double)]
    _u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
    lib.orc_f16_to_f32.restype = C.c_float
    lib.orc_f16_to_f32.argtypes = [C.c_uint16]
    lib.orc_f32_to_f16_rn.restype = C.c_uint16
    lib.orc_f32_to_f16_rn.argtypes = [C.c_float]
    lib.orc_sr_word.restype = C.c_uint32
    lib.orc_sr_word.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    lib.orc_store_f16_sr.restype = C.c_uint16
    lib.orc_store_f16_sr.argtypes = [C.c_float, C.c_uint32, C.c_int]
    lib.orc_f16_add_rn.restype = C.c_uint16
    lib.orc_f16_add_rn.argtypes = [C.c_uint16, C.c_uint16, C.c_int]
    lib.orc_init_factors_f16.restype = None
    lib.orc_init_factors_f16.argtypes = [_u16p, C.c_int64, C.c_int, C.c_uint64, C.c_uint64, C.c_float]
    lib.orc_sgd_update_mixed.restype = C.c_float
    lib.orc_sgd_update_mixed.argtypes = [_u16p, _f32p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_uint64, C.c_uint32,
                                         C.c_int32, C.c_int32, C.c_int]
    lib.orc_train_mixed.restype = C.c_int
    lib.orc_train_mixed.argtypes = [_i32p, _i32p, _f32p, C.c_int64, _u16p, _f32p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                    C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int]
    lib.orc_widen_f16.restype = None
    lib.orc_widen_f16.argtypes = [_u16p, C.c_int64, _f32p]
    lib.orc_set_tree_lanes.restype = None
    lib.orc_set_tree_lanes.argtypes = [C.c_int]
    return lib


lib = _load()


class tree_lanes:
    """with tree_lanes(8): ... -- ORDER_WARP_TREE* sums with that many lanes per rating (the GPU's run kernel:
    run_lanes(k)); outside, the default min(32, pow2ceil(k/4)) of the cold / deterministic / RMSE kernels."""

    def __init__(self, lanes):
        self.lanes = lanes

    def __enter__(self):
        lib.orc_set_tree_lanes(self.lanes)

    def __exit__(self, *exc):
        lib.orc_set_tree_lanes(0)


def run_lanes(k):
    """Lanes per rating of the GPU's run kernel at rank k (kernels_hot.cu run_geometry_for)."""
    return 8 if k <= 32 else (16 if k <= 64 else 32)


def hardware_threads():
    return int(lib.orc_hardware_threads())


def init_factors(n_rows, k, seed, stream, scale=None):
    out = np.empty((n_rows, k), dtype=np.float32)
    if scale is None:
        scale = lib.orc_default_init_scale(k)
    lib.orc_init_factors(out, n_rows, k, seed, stream, scale)
    return out


def shuffle(seed, epoch, n, threads=1):
    out = np.empty(n, dtype=np.int32)
    if threads > 1:
        lib.orc_shuffle_mt(seed, epoch, n, out, threads)
    else:
        lib.orc_shuffle(seed, epoch, n, out)
    return out


def generate(seed, start, count, n_users, n_items, l2au=2, cu=0.25, l2ai=3, ci=0.375, threads=None, amplitude=0.0,
             noise_scale=0.0):
    """amplitude / noise_scale 0 = the defaults of the stand-in (PLANTED_AMPLITUDE, 0.5)."""
    u = np.empty(count, dtype=np.int32)
    i = np.empty(count, dtype=np.int32)
    r = np.empty(count, dtype=np.float32)
    held = np.empty(count, dtype=np.uint8)
    lib.orc_generate2(seed, start, count, n_users, n_items, l2au, cu, l2ai, ci, amplitude, noise_scale, u, i, r, held,
                      threads or hardware_threads())
    return u, i, r, held.astype(bool)


def train(u, i, r, P, Q, lr, lam, epoch_begin, epoch_end, seed, order_mode=ORDER_SEQ, shuffled=True,
          trace=False):
    n_users, k = P.shape
    n_items = Q.shape[0]
    tr = np.empty((epoch_end - epoch_begin) * len(r), dtype=np.float32) if trace else None
    rc = lib.orc_train(u, i, r, len(r), P, Q, n_users, n_items, k, lr, lam, epoch_begin, epoch_end, seed,
                       order_mode, int(shuffled), tr.ctypes.data if trace else None)
    if rc:
        raise ValueError("oracle: bad triplets")
    return tr


def train_tape(u, i, r, P, Q, lr, lam, epoch, seed, order_mode=ORDER_SEQ):
    n_users, k = P.shape
    n = len(r)
    order = np.empty(n, dtype=np.int32)
    pre_p, pre_q, post_p, post_q = (np.empty((n, k), dtype=np.float32) for _ in range(4))
    err = np.empty(n, dtype=np.float32)
    rc = lib.orc_train_tape(u, i, r, n, P, Q, n_users, Q.shape[0], k, lr, lam, epoch, seed, order_mode,
                            order, pre_p, pre_q, post_p, post_q, err)
    if rc:
        raise ValueError("oracle: bad triplets")
    return order, pre_p, pre_q, post_p, post_q, err


def factorize(u, i, r, n_users, n_items, k, lr, lam, epochs, seed, order_mode=ORDER_SEQ):
    P = np.empty((n_users, k), dtype=np.float32)
    Q = np.empty((n_items, k), dtype=np.float32)
    rc = lib.orc_factorize(u, i, r, len(r), n_users, n_items, k, lr, lam, epochs, seed, order_mode, P, Q)
    if rc:
        raise ValueError("oracle: bad arguments")
    return P, Q


def train_hogwild(u, i, r, P, Q, lr, lam, epoch_begin, epoch_end, seed, threads, shuffled=True):
    secs = C.c_double(0.0)
    rc = lib.orc_train_hogwild(u, i, r, len(r), P, Q, P.shape[0], Q.shape[0], P.shape[1], lr, lam,
                               epoch_begin, epoch_end, seed, threads, int(shuffled), C.byref(secs))
    if rc:
        raise ValueError("oracle: bad arguments")
    return secs.value


def rmse(P, Q, u, i, r, order_mode=ORDER_SEQ):
    return float(lib.orc_rmse(P, Q, P.shape[1], u, i, r, len(r), order_mode))


def block_perm(n, seed, epoch, bucket_id):
    """The engine's per-epoch permutation of a bucket of n records (csrc/common.cuh block_perm, restated in oracle.cpp):
    out[j] = index inside the bucket of the record read at position j."""
    out = np.empty(n, dtype=np.int64)
    lib.orc_block_perm_fill(n, seed, epoch, bucket_id, out)
    return out


class RunPlan:
    """The run kernel's work for one ring member: units (mfsgd_plan_runs) plus the bucket every unit lies in."""

    def __init__(self, start, count, item, weight, block_off, member=0):
        self.start = np.ascontiguousarray(start, np.int64)
        self.count = np.ascontiguousarray(count, np.int32)
        self.item = np.ascontiguousarray(item, np.int32)
        self.weight = np.ascontiguousarray(weight, np.float32)
        off = np.asarray(block_off, np.int64)
        blk = np.searchsorted(off, self.start, side="right") - 1
        self.bstart = np.ascontiguousarray(off[blk], np.int64)
        self.bn = np.ascontiguousarray(off[blk + 1] - off[blk], np.int32)
        self.bid = np.ascontiguousarray(member * (len(off) - 1) + blk, np.uint32)


def global_mean(r):
    """Training mean of the model extension (stand-in globalMean: exact integer sum -> the same binary32 everywhere)."""
    return float(lib.orc_global_mean(np.ascontiguousarray(r, np.float32), len(r)))


def train_model(u, i, rc, P, Q, bu, bi, lr, lam, epoch_begin, epoch_end, seed, order_mode=ORDER_SEQ, shuffled=True):
    """factorizeModel's loop on CENTRED ratings rc = r - mu; bu / bi = bias arrays (float32, updated in place) or None."""
    rc_ = lib.orc_train_model(u, i, rc, len(rc), P, Q, None if bu is None else bu.ctypes.data, None if bi is None else bi.ctypes.data,
                              P.shape[0], Q.shape[0], P.shape[1], lr, lam, epoch_begin, epoch_end, seed, order_mode, int(shuffled))
    if rc_:
        raise ValueError("oracle: bad triplets")


def init_factors_f16(n_rows, k, seed, stream, scale=None):
    """initFactors narrowed to binary16 (round to nearest even): the mixed-precision storage of P."""
    rows = np.zeros((n_rows, k), np.uint16)
    lib.orc_init_factors_f16(rows, n_rows, k, seed, stream, lib.orc_default_init_scale(k) if scale is None else scale)
    return rows


def widen(P16):
    """binary16 bit patterns -> float32 (exact)."""
    out = np.zeros(P16.shape, np.float32)
    lib.orc_widen_f16(np.ascontiguousarray(P16), P16.size, out)
    return out


def train_mixed(u, i, r, P16, Q, lr, lam, epoch_begin, epoch_end, seed, order_mode=ORDER_SEQ, shuffled=True, sr=True):
    """The stand-in's loop with P kept in binary16 (oracle.cpp orc_train_mixed); sr=False rounds to nearest even instead;
    sr=2: rows moved by the binary16 difference (the engine's heavy-user red, two binary16 roundings)."""
    rc = lib.orc_train_mixed(u, i, r, len(r), P16, Q, P16.shape[0], Q.shape[0], Q.shape[1], lr, lam, epoch_begin, epoch_end, seed,
                             order_mode, int(shuffled), int(sr))
    if rc:
        raise ValueError("oracle: bad triplets")


def learning_rate(lr, decay, epoch):
    """Stand-in learningRate: lr * decay^epoch by repeated binary32 multiplication."""
    return float(lib.orc_learning_rate(lr, decay, epoch))


def train_early_stop(u, i, rc, vu, vi, vrc, P, Q, bu, bi, lr, lam, lr_decay, patience, min_delta, max_epochs, seed, order_mode=ORDER_SEQ):
    """Stand-in factorizeEarlyStop's loop on centred ratings; returns (epochs run, validation RMSE per epoch run)."""
    curve = (C.c_double * max(max_epochs, 1))()
    ran = lib.orc_train_early_stop(u, i, rc, len(rc), vu, vi, vrc, len(vrc), P, Q, None if bu is None else bu.ctypes.data,
                                   None if bi is None else bi.ctypes.data, P.shape[0], Q.shape[0], P.shape[1], lr, lam, lr_decay, patience,
                                   min_delta, max_epochs, seed, order_mode, curve)
    if ran < 0:
        raise ValueError("oracle: bad triplets")
    return ran, [curve[e] for e in range(ran)]


def rmse_model(P, Q, bu, bi, u, i, rc, order_mode=ORDER_SEQ):
    return float(lib.orc_rmse_model(P, Q, None if bu is None else bu.ctypes.data, None if bi is None else bi.ctypes.data, P.shape[1],
                                    u, i, rc, len(rc), order_mode))


def train_runs_launch(rec_u, rec_r, plan, lo, hi, P, Q, lr, lam, order_mode, resident, gpw=1, virt=False, seed=0, epoch=0,
                      u_base=0, i_base=0, always_add=False, bu=None, bi=None):
    """Twin of one sgd_update_runs_kernel launch over units [lo, hi) of `plan` (oracle.cpp orc_train_runs_launch)."""
    sl = slice(lo, hi)
    rc = lib.orc_train_runs_launch(rec_u, rec_r, len(rec_r), plan.start[sl].copy(), plan.count[sl].copy(), plan.item[sl].copy(),
                                   plan.weight[sl].copy(), plan.bstart[sl].copy(), plan.bn[sl].copy(), plan.bid[sl].copy(),
                                   hi - lo, int(virt), seed, epoch, P, u_base, Q, i_base, P.shape[1], lr, lam, order_mode,
                                   resident, gpw, int(always_add), None if bu is None else bu.ctypes.data,
                                   None if bi is None else bi.ctypes.data)
    if rc:
        raise ValueError("oracle: bad run plan (%d)" % rc)


def balanced_bounds(ids, n_rows, nblocks):
    """Rating-count-balanced row bounds, as the engine computes them (csrc/kernels_layout.cu balanced_bounds_kernel):
    bounds[b] = first row whose exclusive cumulative rating count reaches b * total / nblocks; bounds[nblocks] = n_rows."""
    cnt = np.bincount(ids, minlength=n_rows).astype(np.int64)
    cum = np.concatenate([[0], np.cumsum(cnt)])
    total = int(cum[-1])
    b = [int(np.searchsorted(cum[:-1], (total * j) // nblocks, side="left")) for j in range(nblocks)]
    b[0] = 0
    return np.array(b + [n_rows], dtype=np.int32)


def dsgd_order(u, i, user_bounds, item_bounds, seed, epoch):
    """Visiting order of the DSGD schedule for the sequential rule: sub-epoch s = 0..G-1, member g = 0..G-1 trains block
    (user stripe g, item group (g + s) % G); inside a block the stand-in's shuffled order of the epoch. Returns record indices."""
    G = len(user_bounds) - 1
    order = shuffle(seed, epoch, len(u)).astype(np.int64)
    g = (np.searchsorted(user_bounds, u[order], side="right") - 1).astype(np.int64)
    grp = (np.searchsorted(item_bounds, i[order], side="right") - 1).astype(np.int64)
    s = (grp - g) % G
    return order[np.argsort(s * G + g, kind="stable")]
