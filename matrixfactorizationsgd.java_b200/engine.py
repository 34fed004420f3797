"""Handle-level Python mirror of the C ABI (include/mfsgd.h): one method per mfsgd_* call.

This is host plumbing only -- all arithmetic happens in libmfsgd.so on the GPU."""
import ctypes as C

import numpy as np

from . import _capi as capi
from ._capi import Config, EpochStats, LayoutInfo, SynthParams, check, lib, ptr, as_f32, as_i32


def make_config(n_users, n_items, k, lr, lambda_, seed=20261018, mode=capi.MODE_HOGWILD, n_gpus=1,
                stripes_per_gpu=0, shards_per_gpu=0, scatter=capi.SCATTER_STORE, flags=0, device=0,
                world_size=1, rank=0, nccl_id=None, init_scale=0.0, ctas_per_sm=0, rounds=0, hot_share=0.0, hot_chunk=0,
                merge_boost=0.0, p_atomic_threshold=0.0, model=0, lr_decay=0.0, early_stop_patience=0,
                early_stop_min_delta=0.0, p_storage=0):
    cfg = Config()
    check(lib.mfsgd_config_default(C.byref(cfg)))
    cfg.n_users, cfg.n_items, cfg.k = int(n_users), int(n_items), int(k)
    cfg.lr, cfg.lambda_, cfg.init_scale = float(lr), float(lambda_), float(init_scale)
    cfg.seed, cfg.mode, cfg.n_gpus = int(seed), int(mode), int(n_gpus)
    cfg.stripes_per_gpu, cfg.shards_per_gpu = int(stripes_per_gpu), int(shards_per_gpu)
    cfg.scatter, cfg.flags, cfg.device = int(scatter), int(flags), int(device)
    cfg.world_size, cfg.rank, cfg.ctas_per_sm, cfg.rounds = int(world_size), int(rank), int(ctas_per_sm), int(rounds)
    cfg.hot_share, cfg.hot_chunk, cfg.merge_boost = float(hot_share), int(hot_chunk), float(merge_boost)
    cfg.p_atomic_threshold = float(p_atomic_threshold)
    cfg.model = int(model)
    cfg.lr_decay, cfg.early_stop_patience, cfg.early_stop_min_delta = float(lr_decay), int(early_stop_patience), float(early_stop_min_delta)
    cfg.p_storage = int(p_storage)
    if nccl_id is not None:
        C.memmove(cfg.nccl_id, bytes(nccl_id), 128)
    return cfg


def synth_params(n_total, seed=20261018, log2_alpha_user=2, c_user=0.25, log2_alpha_item=3, c_item=0.375, amplitude=0.0,
                 noise_scale=0.0):
    """amplitude / noise_scale 0 = the stand-in's defaults (noise-dominant); see workloads.SIGNAL for the signal-dominant variant."""
    return SynthParams(int(n_total), int(seed), int(log2_alpha_user), int(log2_alpha_item), float(c_user),
                       float(c_item), float(amplitude), float(noise_scale))


def synth_params_of(w, seed=20261018):
    """SynthParams of a workloads.Workload (shape, skew, and the signal-dominant variant's amplitude / noise scale)."""
    return synth_params(w.n_ratings, seed, w.log2_alpha_user, w.c_user, w.log2_alpha_item, w.c_item, w.amplitude, w.noise_scale)


def device_count():
    n = C.c_int32(0)
    rc = lib.mfsgd_device_count(C.byref(n))
    return n.value if rc == capi.OK else 0


def measure_ceilings(device=0, buffer_mb=61.0):
    """Measured ceilings of the update path's access pattern on `device` (mfsgd_measure_ceilings) as a dict."""
    out = capi.Ceilings()
    check(lib.mfsgd_measure_ceilings(int(device), float(buffer_mb), C.byref(out)))
    return {name: getattr(out, name) for name, _ in capi.Ceilings._fields_ if name != "reserved"}


def nccl_unique_id():
    buf = (C.c_uint8 * 128)()
    check(lib.mfsgd_nccl_unique_id(buf))
    return bytes(buf)


class Engine:
    """Owns one mfsgd_handle. Use as a context manager or call close()."""

    def __init__(self, cfg):
        self.cfg = cfg
        self._h = C.c_void_p()
        check(lib.mfsgd_create(C.byref(cfg), C.byref(self._h)))

    # -- lifecycle --
    def close(self):
        if self._h:
            lib.mfsgd_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- data --
    def load_ratings(self, users, items, ratings):
        u, i, r = as_i32(users), as_i32(items), as_f32(ratings)
        if not (len(u) == len(i) == len(r)):
            raise ValueError("triplet arrays differ in length")
        check(lib.mfsgd_load_ratings(self._h, ptr(u), ptr(i), ptr(r), len(r)))

    def load_ratings_sharded(self, users, items, ratings):
        """Multi-process ring: this rank's own slice of the triplets (collective call; see mfsgd_load_ratings_sharded)."""
        u, i, r = as_i32(users), as_i32(items), as_f32(ratings)
        if not (len(u) == len(i) == len(r)):
            raise ValueError("triplet arrays differ in length")
        check(lib.mfsgd_load_ratings_sharded(self._h, ptr(u), ptr(i), ptr(r), len(r)))

    def load_heldout(self, users, items, ratings):
        u, i, r = as_i32(users), as_i32(items), as_f32(ratings)
        if not (len(u) == len(i) == len(r)):
            raise ValueError("triplet arrays differ in length")
        check(lib.mfsgd_load_heldout(self._h, ptr(u), ptr(i), ptr(r), len(r)))

    def generate_synthetic(self, sp):
        nt, nh = C.c_int64(0), C.c_int64(0)
        check(lib.mfsgd_generate_synthetic(self._h, C.byref(sp), C.byref(nt), C.byref(nh)))
        return nt.value, nh.value

    # -- factors --
    def init_factors(self):
        check(lib.mfsgd_init_factors(self._h))

    def set_factors(self, P, Q):
        P, Q = as_f32(P), as_f32(Q)
        if P.shape != (self.cfg.n_users, self.cfg.k) or Q.shape != (self.cfg.n_items, self.cfg.k):
            raise ValueError("P/Q shapes do not match the configuration")
        check(lib.mfsgd_set_factors(self._h, ptr(P), ptr(Q)))

    def get_model(self):
        """(global mean, user biases, item biases) of the model extension; biases are None when MFSGD_MODEL_BIASES is off."""
        mu = C.c_float(0.0)
        if self.cfg.model & capi.MODEL_BIASES:
            bu = np.zeros(self.cfg.n_users, dtype=np.float32)
            bi = np.zeros(self.cfg.n_items, dtype=np.float32)
            check(lib.mfsgd_get_model(self._h, C.byref(mu), ptr(bu), ptr(bi)))
            return mu.value, bu, bi
        check(lib.mfsgd_get_model(self._h, C.byref(mu), None, None))
        return mu.value, None, None

    def progress(self):
        """(epochs trained since the load, learning rate of the next epoch, last train() ended on the early-stopping rule)."""
        ep, lr, stopped = C.c_int32(0), C.c_float(0.0), C.c_int32(0)
        check(lib.mfsgd_get_progress(self._h, C.byref(ep), C.byref(lr), C.byref(stopped)))
        return ep.value, lr.value, bool(stopped.value)

    def set_biases(self, user_bias, item_bias):
        bu, bi = as_f32(user_bias), as_f32(item_bias)
        if bu.shape != (self.cfg.n_users,) or bi.shape != (self.cfg.n_items,):
            raise ValueError("bias arrays must be [n_users] and [n_items]")
        check(lib.mfsgd_set_biases(self._h, ptr(bu), ptr(bi)))

    def get_factors(self):
        P = np.zeros((self.cfg.n_users, self.cfg.k), dtype=np.float32)
        Q = np.zeros((self.cfg.n_items, self.cfg.k), dtype=np.float32)
        check(lib.mfsgd_get_factors(self._h, ptr(P), ptr(Q)))
        return P, Q

    def partition(self):
        v = [C.c_int32(0) for _ in range(4)]
        check(lib.mfsgd_get_partition(self._h, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)

    # -- training / evaluation --
    def train(self, epochs, want_stats=True):
        stats = (EpochStats * max(epochs, 1))() if want_stats else None
        check(lib.mfsgd_train(self._h, epochs, stats))
        return list(stats)[:epochs] if want_stats else None

    def train_traced(self, epochs, n_train):
        stats = (EpochStats * max(epochs, 1))()
        trace = np.empty(epochs * n_train, dtype=np.float32)
        check(lib.mfsgd_train_traced(self._h, epochs, stats, ptr(trace)))
        return list(stats)[:epochs], trace

    def set_eval_every_epoch(self, on=True):
        check(lib.mfsgd_set_eval_every_epoch(self._h, int(bool(on))))

    def rmse(self, users, items, ratings):
        u, i, r = as_i32(users), as_i32(items), as_f32(ratings)
        if not (len(u) == len(i) == len(r)):
            raise ValueError("triplet arrays differ in length")
        out = C.c_double(0.0)
        check(lib.mfsgd_rmse(self._h, ptr(u), ptr(i), ptr(r), len(r), C.byref(out)))
        return out.value

    def _rmse_loaded(self, fn):
        rm, sse, n = C.c_double(0.0), C.c_double(0.0), C.c_int64(0)
        check(fn(self._h, C.byref(rm), C.byref(sse), C.byref(n)))
        return rm.value, sse.value, n.value

    def rmse_heldout(self):
        return self._rmse_loaded(lib.mfsgd_rmse_heldout)

    def rmse_train(self):
        return self._rmse_loaded(lib.mfsgd_rmse_train)

    # -- introspection --
    def layout_info(self):
        info = LayoutInfo()
        check(lib.mfsgd_get_layout_info(self._h, C.byref(info)))
        return info

    def bounds(self):
        info = self.layout_info()
        ub = np.zeros(info.user_blocks + 1, dtype=np.int32)
        ib = np.zeros(info.item_blocks + 1, dtype=np.int32)
        check(lib.mfsgd_get_bounds(self._h, ptr(ub), ptr(ib)))
        return ub, ib

    def records(self, member=0, with_marks=False):
        """(u, i, r, bucket offsets) of a ring member's current layout; with_marks keeps the heavy-user mark (bit 31 of u:
        negative values) that the run kernel and its oracle twin read, otherwise u is the plain id."""
        info = self.layout_info()
        n = C.c_int64(0)
        check(lib.mfsgd_get_records(self._h, member, None, None, C.byref(n)))
        recs = np.zeros((n.value, 3), dtype=np.int32)
        off = np.zeros(info.stripes_per_gpu * (info.item_blocks + info.n_hot_items) + 1, dtype=np.int64)
        check(lib.mfsgd_get_records(self._h, member, ptr(recs), ptr(off), C.byref(n)))
        u = recs[:, 0].copy()
        if not with_marks:
            u &= 0x7fffffff
        return u, recs[:, 1].copy(), recs[:, 2].copy().view(np.float32), off

    def shuffle_once(self, epoch):
        check(lib.mfsgd_shuffle_once(self._h, int(epoch)))


def apply_updates_forced(k, lr, lambda_, pre_p, pre_q, r, device=0):
    pre_p, pre_q, r = as_f32(pre_p), as_f32(pre_q), as_f32(r)
    n = len(r)
    post_p, post_q = np.empty_like(pre_p), np.empty_like(pre_q)
    err = np.empty(n, dtype=np.float32)
    check(lib.mfsgd_apply_updates_forced(device, k, lr, lambda_, n, ptr(pre_p), ptr(pre_q), ptr(r), ptr(post_p),
                                         ptr(post_q), ptr(err)))
    return post_p, post_q, err


def generate_to_host(sp, n_users, n_items, start, count, device=0):
    u = np.empty(count, dtype=np.int32)
    i = np.empty(count, dtype=np.int32)
    r = np.empty(count, dtype=np.float32)
    held = np.empty(count, dtype=np.uint8)
    check(lib.mfsgd_generate_to_host(device, C.byref(sp), n_users, n_items, start, count, ptr(u), ptr(i), ptr(r),
                                     ptr(held)))
    return u, i, r, held.astype(bool)
