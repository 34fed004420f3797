"""GPU parity of the model extension: r ~ mu + b_u + b_i + p_u . q_i -- stand-in factorizeModel (:305),
sgdUpdateModel (:282), globalMean (:272), rmseModel (:389). Same bars as tests/test_gpu_parity.py: deterministic mode and every
conflict-free schedule bit for bit against the oracle's restatement, the averaged merge against its oracle twin to 1e-5,
Hogwild / DSGD held-out RMSE within 0.5 % (both sides) of the sequential oracle at equal epochs. Every call goes through the C ABI."""
import numpy as np
import pytest

import matrixfactorizationsgd.java_b200 as mf
from matrixfactorizationsgd.java_b200 import _capi as capi
import pyoracle as orc
from test_gpu_parity import (RMSE_TOL, SEED, MidSet, assert_curve_parity, assert_ring_rmse_parity, assert_rmse_parity, median_run,  # noqa: F401
                             plan_runs_of,
                             split)

pytestmark = pytest.mark.gpu

MEAN, BIASES = capi.MODEL_GLOBAL_MEAN, capi.MODEL_BIASES


def centred(r, bits):
    mu = orc.global_mean(r) if bits & MEAN else 0.0
    return np.float32(mu), (r - np.float32(mu)).astype(np.float32)


def zeros_or_none(n, bits):
    return np.zeros(n, np.float32) if bits & BIASES else None


@pytest.mark.parametrize("bits", [MEAN | BIASES, BIASES, MEAN])
@pytest.mark.parametrize("k", [8, 32, 100, 128])
def test_model_deterministic_mode_bit_exact(k, bits):
    nu, ni, n = 300, 200, 6000
    u, i, r, held = orc.generate(SEED + k, 0, n, nu, ni)
    (u, i, r), (hu, hi, hr) = split(u, i, r, held)
    got = mf.MatrixFactorizationSGD.factorizeModel(u, i, r, nu, ni, k, 0.02, 0.03, 3, SEED, bool(bits & MEAN), bool(bits & BIASES),
                                                   mode=capi.MODE_DETERMINISTIC)
    mu, rc = centred(r, bits)
    assert np.float32(got.globalMean) == mu                        # exact integer sum on both sides: the same binary32
    P, Q = orc.init_factors(nu, k, SEED, 0), orc.init_factors(ni, k, SEED, 1)
    bu, bi = zeros_or_none(nu, bits), zeros_or_none(ni, bits)
    orc.train_model(u, i, rc, P, Q, bu, bi, 0.02, 0.03, 0, 3, SEED, orc.ORDER_WARP_TREE)
    assert np.array_equal(got.P, P) and np.array_equal(got.Q, Q)
    if bits & BIASES:
        assert np.array_equal(got.userBias, bu) and np.array_equal(got.itemBias, bi) and np.abs(bu).max() > 0
    else:
        assert got.userBias is None and got.itemBias is None
    # the stand-in's sequential dot: last-ulp drift only
    Ps, Qs = orc.init_factors(nu, k, SEED, 0), orc.init_factors(ni, k, SEED, 1)
    bus, bis = zeros_or_none(nu, bits), zeros_or_none(ni, bits)
    orc.train_model(u, i, rc, Ps, Qs, bus, bis, 0.02, 0.03, 0, 3, SEED, orc.ORDER_SEQ)
    assert np.abs(got.P - Ps).max() < 1e-4 and np.abs(got.Q - Qs).max() < 1e-4
    # rmseModel through the RMSE kernel
    want = orc.rmse_model(Ps, Qs, bus, bis, hu, hi, (hr - mu).astype(np.float32))
    assert abs(mf.MatrixFactorizationSGD.rmseModel(got, hu, hi, hr) - want) / want < 1e-4


def test_model_off_is_the_reference_model():
    nu, ni, n, k = 300, 200, 6000, 32
    u, i, r, _ = orc.generate(SEED, 0, n, nu, ni)
    got = mf.MatrixFactorizationSGD.factorizeModel(u, i, r, nu, ni, k, 0.02, 0.03, 2, SEED, False, False, mode=capi.MODE_DETERMINISTIC)
    ref = mf.MatrixFactorizationSGD.factorize(u, i, r, nu, ni, k, 0.02, 0.03, 2, SEED, mode=capi.MODE_DETERMINISTIC)
    assert np.array_equal(got.P, ref.P) and np.array_equal(got.Q, ref.Q) and got.globalMean == 0.0 and got.userBias is None


@pytest.mark.parametrize("k", [8, 32, 64, 128, 256])
@pytest.mark.parametrize("arith", ["fast", "exact", "exact-atomic"])
def test_model_hogwild_kernel_bit_exact_on_conflict_free_data(k, arith):
    """Pairwise distinct users and items: every P, Q row and every bias entry is touched once per epoch, so the full-grid
    kernel equals the oracle in any order, biases included."""
    n = 5003
    rng = np.random.default_rng(k)
    u = rng.permutation(n).astype(np.int32)
    i = rng.permutation(n).astype(np.int32)
    r = (1 + 4 * rng.random(n)).astype(np.float32)
    bits = MEAN | BIASES
    mu, rc = centred(r, bits)
    order = orc.ORDER_WARP_TREE_FMA if arith == "fast" else orc.ORDER_WARP_TREE
    P, Q = orc.init_factors(n, k, SEED, 0), orc.init_factors(n, k, SEED, 1)
    bu, bi = np.zeros(n, np.float32), np.zeros(n, np.float32)
    orc.train_model(u, i, rc, P, Q, bu, bi, 0.02, 0.03, 0, 3, SEED, order)
    kw = dict(flags=0 if arith == "fast" else capi.FLAG_EXACT_ARITH,
              scatter=capi.SCATTER_ATOMIC if arith == "exact-atomic" else capi.SCATTER_STORE)
    for extra in (dict(stripes_per_gpu=1, shards_per_gpu=1), dict(stripes_per_gpu=3, shards_per_gpu=2),
                  dict(mode=capi.MODE_DSGD, n_gpus=4, stripes_per_gpu=2)):
        if "mode" in extra:
            if arith == "exact-atomic":
                continue
            kw2 = dict(kw, flags=kw["flags"] | capi.FLAG_VIRTUAL_RING)
        else:
            kw2 = kw
        got = mf.MatrixFactorizationSGD.factorizeModel(u, i, r, n, n, k, 0.02, 0.03, 3, SEED, True, True, **dict(kw2, **extra))
        assert np.float32(got.globalMean) == mu
        if arith == "exact-atomic":      # p + fl(delta) rounds once more than the store path; the bias add is the rule itself
            np.testing.assert_allclose(got.P, P, rtol=3e-7, atol=1e-9)
            np.testing.assert_allclose(got.Q, Q, rtol=3e-7, atol=1e-9)
            np.testing.assert_allclose(got.userBias, bu, rtol=3e-7, atol=1e-9)
            np.testing.assert_allclose(got.itemBias, bi, rtol=3e-7, atol=1e-9)
        else:
            assert np.array_equal(got.P, P) and np.array_equal(got.Q, Q)
            assert np.array_equal(got.userBias, bu) and np.array_equal(got.itemBias, bi)
    assert np.abs(bu).max() > 0 and np.abs(bi).max() > 0


@pytest.mark.parametrize("arith", ["fast", "exact", "fast-heavy"])
@pytest.mark.parametrize("k", [8, 32, 64, 100, 128, 256])
def test_model_run_kernel_exact_sequential_runs(k, arith):
    """Run path, one run per item: the run carries q_i AND b_i privately and stores both at its end; b_u is stored (or, for a
    heavy user, added in memory) beside p_u. Users pairwise distinct -> bit-exact against the sequential oracle."""
    n_hot, per_hot, n_cold = 5, 3000, 5003
    n = n_hot * per_hot + n_cold
    rng = np.random.default_rng(7)
    items = np.concatenate([np.repeat(np.arange(n_hot), per_hot), n_hot + np.arange(n_cold)]).astype(np.int32)
    i = items[rng.permutation(n)]
    u = rng.permutation(n).astype(np.int32)
    r = (1 + 4 * rng.random(n)).astype(np.float32)
    ni = n_hot + n_cold
    heavy, exact = arith.endswith("-heavy"), arith == "exact"
    cfg = mf.make_config(n, ni, k, 0.01, 0.03, seed=SEED, mode=capi.MODE_HOGWILD, stripes_per_gpu=1, rounds=1, hot_chunk=4096,
                         flags=capi.FLAG_NO_SHUFFLE | (capi.FLAG_EXACT_ARITH if exact else 0),
                         p_atomic_threshold=1e-9 if heavy else -1.0, model=MEAN | BIASES)
    with mf.Engine(cfg) as eng:
        eng.load_ratings(u, i, r)
        assert eng.layout_info().n_hot_items == n_hot
        ou, oi, orc_r, off = eng.records()              # stored ratings are centred
        eng.init_factors()
        eng.train(3)
        P, Q = eng.get_factors()
        mu, bu, bi = eng.get_model()
    mu_o, rc = centred(r, MEAN)
    assert np.float32(mu) == mu_o
    srt_g = np.lexsort((oi, ou))
    srt_o = np.lexsort((i, u))
    assert np.array_equal(orc_r[srt_g], rc[srt_o])      # the same centred binary32 values, record for record
    Po, Qo = orc.init_factors(n, k, SEED, 0), orc.init_factors(ni, k, SEED, 1)
    buo, bio = np.zeros(n, np.float32), np.zeros(ni, np.float32)
    order = orc.ORDER_WARP_TREE if exact else orc.ORDER_WARP_TREE_FMA
    run_order = orc.ORDER_WARP_TREE_FMA_PDELTA if heavy else order
    hot = oi < n_hot
    with orc.tree_lanes(orc.run_lanes(k)):
        orc.train_model(ou[hot].copy(), oi[hot].copy(), orc_r[hot].copy(), Po, Qo, buo, bio, 0.01, 0.03, 0, 3, SEED, run_order, shuffled=False)
    orc.train_model(ou[~hot].copy(), oi[~hot].copy(), orc_r[~hot].copy(), Po, Qo, buo, bio, 0.01, 0.03, 0, 3, SEED, run_order, shuffled=False)   # the cold kernel honours the heavy mark too
    assert np.array_equal(P, Po) and np.array_equal(Q, Qo)
    assert np.array_equal(bu, buo) and np.array_equal(bi, bio)


@pytest.mark.parametrize("shuffle", [False, True])
@pytest.mark.parametrize("k,rounds", [(128, 1), (32, 1), (128, 2)])
def test_model_run_kernel_averaged_merge_matches_its_oracle_twin(k, rounds, shuffle):
    """Several runs of one item per launch: b_i merges like q_i (weight * (b_run - b_start), added in memory)."""
    n_hot, pieces, chunk, n_cold, boost = 6, 8, 64, 1003, 1.25
    per_hot = pieces * chunk * rounds
    n = n_hot * per_hot + n_cold
    rng = np.random.default_rng(100 + k + rounds)
    items = np.concatenate([np.repeat(np.arange(n_hot), per_hot), n_hot + np.arange(n_cold)]).astype(np.int32)
    i = items[rng.permutation(n)]
    u = rng.permutation(n).astype(np.int32)
    r = (1 + 4 * rng.random(n)).astype(np.float32)
    nu, ni, lr, lam, epochs = n, n_hot + n_cold, 0.01, 0.03, 3
    cfg = mf.make_config(nu, ni, k, lr, lam, seed=SEED, mode=capi.MODE_HOGWILD, stripes_per_gpu=1, rounds=rounds, hot_chunk=chunk,
                         merge_boost=boost, flags=0 if shuffle else capi.FLAG_NO_SHUFFLE, p_atomic_threshold=-1.0, model=MEAN | BIASES)
    with mf.Engine(cfg) as eng:
        eng.load_ratings(u, i, r)
        ou, oi, orr, off = eng.records()
        eng.init_factors()
        eng.train(epochs)
        P, Q = eng.get_factors()
        _, bu, bi = eng.get_model()
    plan, visits = plan_runs_of(off, n_hot, np.arange(n_hot), rounds, chunk, boost)
    assert len(plan.start) == n_hot * pieces * rounds
    per_warp = 32 // orc.run_lanes(k)
    Po, Qo = orc.init_factors(nu, k, SEED, 0), orc.init_factors(ni, k, SEED, 1)
    buo, bio = np.zeros(nu, np.float32), np.zeros(ni, np.float32)
    cold = slice(int(off[0]), int(off[1]))
    for e in range(epochs):
        for rnd in range(rounds):
            lo, hi = int(visits[rnd]), int(visits[rnd + 1])
            grid = max(1, -(-(hi - lo) // (16 * per_warp)))
            with orc.tree_lanes(orc.run_lanes(k)):
                orc.train_runs_launch(ou, orr, plan, lo, hi, Po, Qo, lr, lam, orc.ORDER_WARP_TREE_FMA, grid * 8 * per_warp, per_warp,
                                      virt=shuffle, seed=SEED, epoch=e, bu=buo, bi=bio)
        orc.train_model(ou[cold].copy(), oi[cold].copy(), orr[cold].copy(), Po, Qo, buo, bio, lr, lam, e, e + 1, SEED,
                        orc.ORDER_WARP_TREE_FMA, shuffled=False)
    np.testing.assert_allclose(Q, Qo, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(P, Po, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(bu, buo, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(bi, bio, rtol=1e-5, atol=1e-7)
    assert np.abs(bi[:n_hot]).max() > 1e-3


class ModelMidSet(MidSet):
    """The mid-size sets under the extended model. Noise-dominant: the mean and the biases take the stiff common direction out of
    the factors and the sequential oracle ends below plain MF (which ends ~2 % ABOVE the constant predictor there).
    Signal-dominant: the sharp case -- the oracle ends 79 % below the constant predictor."""
    MODEL = "model"

    def __init__(self, signal):
        super().__init__(signal=signal)
        self.mu, self.rc = centred(self.train[2], MEAN)
        self.hc = (self.held[2] - self.mu).astype(np.float32)
        self.plain_rmse = self.fx["plain"]["shuffled"][-1]


@pytest.fixture(scope="module")
def model_midsize():
    m = ModelMidSet(signal=False)
    assert m.oracle_rmse < m.plain_rmse, (m.oracle_rmse, m.const_rmse, m.plain_rmse)
    return m


@pytest.fixture(scope="module")
def model_midsize_signal():
    m = ModelMidSet(signal=True)
    assert m.oracle_rmse < 0.3 * m.const_rmse
    return m


def model_hogwild_run(m, **cfg_kw):
    cfg = mf.make_config(m.nu, m.ni, m.k, m.lr, m.lam, seed=SEED, mode=capi.MODE_HOGWILD, model=MEAN | BIASES, **cfg_kw)
    with mf.Engine(cfg) as eng:
        eng.load_ratings(*m.train)
        eng.load_heldout(*m.held)
        eng.init_factors()
        eng.set_eval_every_epoch(True)
        stats = eng.train(m.epochs)
        got = eng.rmse(*m.held)
        mu, bu, bi = eng.get_model()
        P, Q = eng.get_factors()
    assert np.float32(mu) == m.mu
    assert abs(stats[-1].heldout_rmse - got) < 1e-9
    assert abs(got - orc.rmse_model(P, Q, bu, bi, m.held[0], m.held[1], m.hc)) / got < 1e-6
    return got, [s.heldout_rmse for s in stats]


@pytest.mark.parametrize("variant", ["default", "signal"])
def test_model_hogwild_rmse_parity(model_midsize, model_midsize_signal, variant):
    """Extended model, planned layout: 0.5 % both sides plus the half-epoch lag bound, median of three runs (test_gpu_parity.median_run)."""
    m = model_midsize_signal if variant == "signal" else model_midsize
    _, curve = median_run(lambda: model_hogwild_run(m))
    assert_curve_parity(curve, m.curve)


@pytest.mark.parametrize("variant", ["default", "signal"])
def test_model_hogwild_rmse_parity_forced_four_substripes_within_three_quarters_of_a_percent(model_midsize, model_midsize_signal, variant):
    """Forced into 4 sub-stripes of 3 450 users (see test_gpu_parity's test of the same name): 0.75 %, median of three."""
    m = model_midsize_signal if variant == "signal" else model_midsize
    got, _ = median_run(lambda: model_hogwild_run(m, stripes_per_gpu=4))
    assert abs(got / m.oracle_rmse - 1.0) <= 0.0075, (got, m.oracle_rmse)


@pytest.mark.parametrize("G,mu_,mi", [(2, 1, 1), (4, 2, 2), (8, 1, 1)])
@pytest.mark.parametrize("variant", ["default", "signal"])
def test_model_dsgd_virtual_ring_rmse_parity(model_midsize, model_midsize_signal, variant, G, mu_, mi):
    """The item biases travel round the ring with their Q shard group."""
    m = model_midsize_signal if variant == "signal" else model_midsize
    cfg = mf.make_config(m.nu, m.ni, m.k, m.lr, m.lam, seed=SEED, mode=capi.MODE_DSGD, n_gpus=G, stripes_per_gpu=mu_, shards_per_gpu=mi,
                         flags=capi.FLAG_VIRTUAL_RING, model=MEAN | BIASES)
    with mf.Engine(cfg) as eng:
        eng.load_ratings(*m.train)
        ub, ib = eng.bounds()
        eng.init_factors()
        eng.train(m.epochs)
        got = eng.rmse(*m.held)
        _, bu, bi = eng.get_model()
        P, Q = eng.get_factors()
    assert abs(got - orc.rmse_model(P, Q, bu, bi, m.held[0], m.held[1], m.hc)) / got < 1e-6      # factors and biases came home intact
    assert np.count_nonzero(bi) > 0.9 * np.count_nonzero(np.bincount(m.train[1], minlength=m.ni))      # every group's biases trained
    assert_ring_rmse_parity(got, m.oracle_rmse, m.dsgd_oracle_rmse(ub[::mu_], ib[::mi]))


# ------------------------------------------------------------------------------------------------
# learning-rate schedule and early stopping -- stand-in learningRate :331, factorizeEarlyStop :350
# ------------------------------------------------------------------------------------------------
def _small_model_set(k):
    nu, ni, n = 300, 200, 6000
    u, i, r, held = orc.generate(SEED + k, 0, n, nu, ni)
    return nu, ni, split(u, i, r, held)


@pytest.mark.parametrize("k", [8, 128])
def test_schedule_and_early_stop_deterministic_mode_bit_exact(k):
    """The schedule's rates are the stand-in's binary32 products, the stopping rule fires in the same epoch, the model that comes
    back is the oracle's bit for bit, and mfsgd_get_progress reports all of it."""
    nu, ni, ((u, i, r), (vu, vi, vr)) = _small_model_set(k)
    mu, rc = centred(r, MEAN)
    vrc = (vr - mu).astype(np.float32)
    for decay, patience, min_delta, max_epochs in ((0.8, 0, 0.0, 4), (1.0, 2, 0.5, 12), (0.9, 1, 0.5, 12)):
        got = mf.MatrixFactorizationSGD.factorizeEarlyStop(u, i, r, vu, vi, vr, nu, ni, k, 0.02, 0.03, max_epochs, SEED, True, True,
                                                           decay, patience, min_delta, mode=capi.MODE_DETERMINISTIC)
        P, Q = orc.init_factors(nu, k, SEED, 0), orc.init_factors(ni, k, SEED, 1)
        bu, bi = np.zeros(nu, np.float32), np.zeros(ni, np.float32)
        ran, curve = orc.train_early_stop(u, i, rc, vu, vi, vrc, P, Q, bu, bi, 0.02, 0.03, decay, patience, min_delta, max_epochs, SEED,
                                          orc.ORDER_WARP_TREE)
        assert got.epochsRun == ran == (max_epochs if patience == 0 else 1 + patience)
        assert np.array_equal(got.model.P, P) and np.array_equal(got.model.Q, Q)
        assert np.array_equal(got.model.userBias, bu) and np.array_equal(got.model.itemBias, bi)
        np.testing.assert_allclose(got.validationRmse, curve, rtol=1e-6)


def test_progress_reports_epochs_rate_and_the_stop():
    k = 32
    nu, ni, ((u, i, r), (vu, vi, vr)) = _small_model_set(k)
    cfg = mf.make_config(nu, ni, k, 0.02, 0.03, seed=SEED, mode=capi.MODE_HOGWILD, lr_decay=0.9, early_stop_patience=2,
                         early_stop_min_delta=0.5)
    with mf.Engine(cfg) as eng:
        eng.load_ratings(u, i, r)
        eng.init_factors()
        assert eng.progress() == (0, float(np.float32(0.02)), False)
        with pytest.raises(mf.MfsgdError) as ei:
            eng.train(3)
        assert ei.value.code == capi.E_STATE              # early stopping needs a held-out set
        eng.load_heldout(vu, vi, vr)
        stats = eng.train(10)
        ran, lr_next, stopped = eng.progress()
        assert (ran, stopped) == (3, True) and np.float32(lr_next) == np.float32(orc.learning_rate(0.02, 0.9, 3))
        assert [s.updates for s in stats] == [len(r)] * 3 + [0] * 7
        assert all(np.isfinite(s.heldout_rmse) for s in stats[:3]) and all(np.isnan(s.heldout_rmse) for s in stats[3:])
        eng.train(1)                                       # a new call starts with a clean strike count
        assert eng.progress()[0] == 4 and not eng.progress()[2]
    for bad in (dict(lr_decay=1.5), dict(lr_decay=-0.1), dict(early_stop_patience=-1), dict(early_stop_min_delta=1.0)):
        with pytest.raises(mf.MfsgdError) as ei:
            mf.Engine(mf.make_config(nu, ni, k, 0.02, 0.03, **bad))
        assert ei.value.code == capi.E_INVALID_ARG


@pytest.mark.parametrize("lr_scale,decay,tol", [(2.0, 0.85, 0.01), (1.0, 0.9, 0.02)])
def test_schedule_hogwild_rmse_lag_frozen_by_a_decaying_rate_within_one_and_two_percent(model_midsize_signal, lr_scale, decay, tol):
    """Hogwild under a decaying rate against the sequential oracle under the same schedule on the signal-dominant set, both sides.
    The exception to the 0.5 % bar, and why: the parallel execution trails the sequential one by a fraction of an epoch while the
    curve is steep (the hot items' concurrent runs are averaged); at a constant rate it catches up as the curve
    flattens (0.5 % at equal epochs, test_model_hogwild_rmse_parity), under a decaying rate that lag is frozen in at whatever the
    curve's slope was. Measured (tools/small_sweep.py, profiles/r02_experiments.md section 10, 5 runs each): twice the rate
    x 0.85 per epoch ends +0.72 +- 0.02 % above the oracle (its curve is flat by then), the workload's rate x 0.9 per epoch +1.6 %
    (the rate is down to 12 % while the oracle still falls 1 % per epoch). Sequential executions that differ only in the visiting
    order are within 0.1 % of each other here (tests/golden/order_spread.json), so this is the engine's lag, held to 1 % and 2 %."""
    m = model_midsize_signal
    P, Q = orc.init_factors(m.nu, m.k, SEED, 0), orc.init_factors(m.ni, m.k, SEED, 1)
    bu, bi = np.zeros(m.nu, np.float32), np.zeros(m.ni, np.float32)
    ran, curve = orc.train_early_stop(m.train[0], m.train[1], m.rc, m.held[0], m.held[1], m.hc, P, Q, bu, bi, lr_scale * m.lr, m.lam, decay, 0,
                                      0.0, m.epochs, SEED)
    got = mf.MatrixFactorizationSGD.factorizeEarlyStop(*m.train, *m.held, m.nu, m.ni, m.k, lr_scale * m.lr, m.lam, m.epochs, SEED, True, True,
                                                       decay, 0, 0.0)
    assert got.epochsRun == ran == m.epochs
    assert abs(got.validationRmse[-1] / curve[-1] - 1.0) <= tol, (got.validationRmse[-1], curve[-1])
    assert got.validationRmse[-1] >= curve[-1] * (1 - RMSE_TOL)         # a lag, never a lead beyond the bar


def test_model_state_and_argument_errors():
    with pytest.raises(mf.MfsgdError) as ei:
        mf.Engine(mf.make_config(10, 10, 8, 0.1, 0.1, model=4))
    assert ei.value.code == capi.E_INVALID_ARG
    with mf.Engine(mf.make_config(10, 10, 8, 0.1, 0.1, model=MEAN)) as eng:
        with pytest.raises(mf.MfsgdError) as ei:
            eng.get_model()
        assert ei.value.code == capi.E_STATE                # the mean is the loaded training set's
        eng.load_ratings(np.zeros(3, np.int32), np.zeros(3, np.int32), np.array([1, 2, 4.5], np.float32))
        mu, bu, bi = eng.get_model()
        assert np.float32(mu) == np.float32(2.5) and bu is None and bi is None
        with pytest.raises(mf.MfsgdError) as ei:
            eng.set_biases(np.zeros(10, np.float32), np.zeros(10, np.float32))
        assert ei.value.code == capi.E_STATE                # MFSGD_MODEL_BIASES is off
    with mf.Engine(mf.make_config(10, 10, 8, 0.1, 0.1, model=BIASES)) as eng:
        eng.load_ratings(np.zeros(3, np.int32), np.zeros(3, np.int32), np.ones(3, np.float32))
        with pytest.raises(ValueError):
            eng.set_biases(np.zeros(9, np.float32), np.zeros(10, np.float32))
        eng.init_factors()
        eng.set_biases(np.arange(10, dtype=np.float32), -np.arange(10, dtype=np.float32))
        mu, bu, bi = eng.get_model()
        assert mu == 0.0 and np.array_equal(bu, np.arange(10)) and np.array_equal(bi, -np.arange(10))
