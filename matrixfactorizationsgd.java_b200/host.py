"""Host-side mirror of the reference's entry point.

The reference's class is ``MatrixFactorizationSGD`` (/root/reference/README.md:1); its stand-in
(baseline/java/MatrixFactorizationSGD.java:109) exposes
``factorize(users, items, ratings, nUsers, nItems, k, lr, lambda, epochs, seed) -> Factors{P, Q}``.
The production host is Java over Panama FFM (java/MatrixFactorizationSGDGpu.java); no JDK exists in
this image, so this module is the same thin layer in Python over ctypes -- same names, argument
order, error behaviour (bad arguments raise before any GPU work, like the stand-in's
IllegalArgumentException) -- and is what the parity tests and bench.py call.
"""
import ctypes as C
from collections import namedtuple

import numpy as np

from . import _capi as capi
from .engine import Engine, make_config

Factors = namedtuple("Factors", ["P", "Q", "nUsers", "nItems", "k"])
#: the model extension's result (stand-in Model :257): userBias / itemBias are None when the biases are off
Model = namedtuple("Model", ["P", "Q", "userBias", "itemBias", "globalMean", "nUsers", "nItems", "k"])
#: stand-in EarlyStopResult (:339)
EarlyStopResult = namedtuple("EarlyStopResult", ["model", "epochsRun", "validationRmse"])
#: a parsed ratings file: dense triplets + the file ids of every row (RatingsFile.userIds[u] is the file's id of row u)
RatingsFile = namedtuple("RatingsFile", ["users", "items", "ratings", "nUsers", "nItems", "userIds", "itemIds", "format"])


def read_ratings(path, format=capi.FORMAT_AUTO):
    """mfsgd_read_ratings: MovieLens u.data / ratings.csv / ratings.dat or Netflix-Prize text -> RatingsFile
    (copies of the library's buffers; ids compacted in ascending file-id order). Host-only, needs no GPU."""
    out = capi.Ratings()
    capi.check(capi.lib.mfsgd_read_ratings(str(path).encode(), int(format), C.byref(out)))
    try:
        n, nu, ni = out.n, out.n_users, out.n_items
        grab = lambda p, m, dt: (np.ctypeslib.as_array(p, shape=(m,)).astype(dt, copy=True) if m > 0 else np.empty(0, dt))
        return RatingsFile(grab(out.users, n, np.int32), grab(out.items, n, np.int32), grab(out.ratings, n, np.float32), nu, ni,
                           grab(out.user_ids, nu, np.int64), grab(out.item_ids, ni, np.int64), out.format)
    finally:
        capi.lib.mfsgd_free_ratings(C.byref(out))


class MatrixFactorizationSGD:
    """Drop-in for the factorization path; every call runs on the GPU through libmfsgd.so."""

    #: execution modes of the GPU path (mfsgd_config.mode)
    DETERMINISTIC, HOGWILD, DSGD = capi.MODE_DETERMINISTIC, capi.MODE_HOGWILD, capi.MODE_DSGD

    @staticmethod
    def _check(users, items, ratings, nUsers, nItems, k, epochs):
        if not (len(users) == len(items) == len(ratings)):
            raise ValueError("triplet arrays differ in length")          # stand-in line 112-113
        if k <= 0 or nUsers <= 0 or nItems <= 0 or epochs < 0:
            raise ValueError("bad shape")                                # stand-in line 114-115

    @staticmethod
    def factorize(users, items, ratings, nUsers, nItems, k, lr, lambda_, epochs, seed,
                  mode=capi.MODE_HOGWILD, n_gpus=1, device=0, **cfg_kw):
        """One-shot form: maps 1:1 onto mfsgd_factorize (host arrays in, host P and Q out)."""
        MatrixFactorizationSGD._check(users, items, ratings, nUsers, nItems, k, epochs)
        u, i, r = capi.as_i32(users), capi.as_i32(items), capi.as_f32(ratings)
        cfg = make_config(nUsers, nItems, k, lr, lambda_, seed=seed, mode=mode, n_gpus=n_gpus, device=device, **cfg_kw)
        P = np.zeros((nUsers, k), dtype=np.float32)
        Q = np.zeros((nItems, k), dtype=np.float32)
        capi.check(capi.lib.mfsgd_factorize(capi.ptr(u), capi.ptr(i), capi.ptr(r), len(r), C.byref(cfg), int(epochs),
                                            capi.ptr(P), capi.ptr(Q)))
        return Factors(P, Q, nUsers, nItems, k)

    @staticmethod
    def factorizeMixed(users, items, ratings, nUsers, nItems, k, lr, lambda_, epochs, seed, mode=capi.MODE_HOGWILD, n_gpus=1, device=0,
                       **cfg_kw):
        """Stand-in factorizeMixed (:439): the rows of P kept as binary16 on the device (stochastic rounding from a counter hash),
        arithmetic in binary32; P comes back widened exactly."""
        if k % 4 != 0:
            raise ValueError("bad shape")                                # stand-in line 443
        return MatrixFactorizationSGD.factorize(users, items, ratings, nUsers, nItems, k, lr, lambda_, epochs, seed, mode=mode,
                                                n_gpus=n_gpus, device=device, p_storage=capi.STORAGE_F16, **cfg_kw)

    @staticmethod
    def factorizeModel(users, items, ratings, nUsers, nItems, k, lr, lambda_, epochs, seed, useGlobalMean, useBiases,
                       mode=capi.MODE_HOGWILD, n_gpus=1, device=0, **cfg_kw):
        """Stand-in factorizeModel (:305): r ~ mu + b_u + b_i + p_u . q_i. The mean is taken on the device while the ratings are
        counted, ratings are stored centred, the biases ride through the update kernels beside their rows."""
        MatrixFactorizationSGD._check(users, items, ratings, nUsers, nItems, k, epochs)
        bits = (capi.MODEL_GLOBAL_MEAN if useGlobalMean else 0) | (capi.MODEL_BIASES if useBiases else 0)
        cfg = make_config(nUsers, nItems, k, lr, lambda_, seed=seed, mode=mode, n_gpus=n_gpus, device=device, model=bits, **cfg_kw)
        with Engine(cfg) as eng:
            eng.load_ratings(users, items, ratings)
            eng.init_factors()
            if epochs > 0:
                eng.train(epochs, want_stats=False)
            P, Q = eng.get_factors()
            mu, bu, bi = eng.get_model()
        return Model(P, Q, bu, bi, mu, nUsers, nItems, k)

    @staticmethod
    def factorizeEarlyStop(users, items, ratings, vUsers, vItems, vRatings, nUsers, nItems, k, lr, lambda_, maxEpochs, seed,
                           useGlobalMean, useBiases, lrDecay, patience, minDelta, mode=capi.MODE_HOGWILD, n_gpus=1, device=0, **cfg_kw):
        """Stand-in factorizeEarlyStop (:350): the learning-rate schedule lr_(e+1) = lr_e * lrDecay and the early-stopping rule
        on the validation RMSE, both evaluated inside mfsgd_train (the validation set lives on the device)."""
        MatrixFactorizationSGD._check(users, items, ratings, nUsers, nItems, k, maxEpochs)
        if not (0.0 < lrDecay <= 1.0) or patience < 0 or not (0.0 <= minDelta < 1.0):
            raise ValueError("bad schedule")                                  # stand-in line 356-357
        if not (len(vUsers) == len(vItems) == len(vRatings)):
            raise ValueError("triplet arrays differ in length")               # before any GPU work, like the training triplets
        bits = (capi.MODEL_GLOBAL_MEAN if useGlobalMean else 0) | (capi.MODEL_BIASES if useBiases else 0)
        cfg = make_config(nUsers, nItems, k, lr, lambda_, seed=seed, mode=mode, n_gpus=n_gpus, device=device, model=bits,
                          lr_decay=lrDecay, early_stop_patience=patience, early_stop_min_delta=minDelta, **cfg_kw)
        with Engine(cfg) as eng:
            eng.load_ratings(users, items, ratings)
            eng.load_heldout(vUsers, vItems, vRatings)
            eng.init_factors()
            eng.set_eval_every_epoch(True)
            stats = eng.train(maxEpochs) if maxEpochs > 0 else []
            ran = eng.progress()[0]
            P, Q = eng.get_factors()
            mu, bu, bi = eng.get_model()
        return EarlyStopResult(Model(P, Q, bu, bi, mu, nUsers, nItems, k), ran, [s.heldout_rmse for s in stats[:ran]])

    @staticmethod
    def rmseModel(model, users, items, ratings, device=0):
        """Stand-in rmseModel (:389): e = (r - mu) - ((p_u . q_i + b_u) + b_i), evaluated by the RMSE kernel."""
        P, Q = capi.as_f32(model.P), capi.as_f32(model.Q)
        biased = model.userBias is not None
        cfg = make_config(P.shape[0], Q.shape[0], model.k, 1e-3, 0.0, mode=capi.MODE_HOGWILD, device=device, stripes_per_gpu=1,
                          model=capi.MODEL_BIASES if biased else 0)
        rc = (capi.as_f32(ratings) - np.float32(model.globalMean)).astype(np.float32)      # one binary32 subtraction, as :399
        with Engine(cfg) as eng:
            eng.load_ratings(np.empty(0, np.int32), np.empty(0, np.int32), np.empty(0, np.float32))
            eng.set_factors(P, Q)
            if biased:
                eng.set_biases(model.userBias, model.itemBias)
            return eng.rmse(users, items, rc)

    @staticmethod
    def rmse(P, Q, k, users, items, ratings, device=0):
        """Stand-in line 169: sqrt(mean (r - p_u.q_i)^2), evaluated by the RMSE kernel."""
        P, Q = capi.as_f32(P), capi.as_f32(Q)
        if P.ndim != 2 or Q.ndim != 2 or P.shape[1] != k or Q.shape[1] != k:
            raise ValueError("P and Q must be [rows, k]")
        cfg = make_config(P.shape[0], Q.shape[0], k, 1e-3, 0.0, mode=capi.MODE_HOGWILD, device=device,
                          stripes_per_gpu=1)
        with Engine(cfg) as eng:
            eng.load_ratings(np.empty(0, np.int32), np.empty(0, np.int32), np.empty(0, np.float32))
            eng.set_factors(P, Q)
            return eng.rmse(users, items, ratings)
