"""CPU tests of the oracle (oracle/oracle.cpp) against the committed pins.

The reference has no tests or vectors (README.md:1-2 only); the pins are: published SplitMix64
outputs, the hand-computed KAT, and tests/golden/golden_small.json produced by the independent
NumPy restatement (tests/np_restatement.py, script tests/golden/make_golden.py).
"""
import numpy as np
import pytest

import np_restatement as npr
import pyoracle as orc

SEED = 20261018


def f32(bits):
    return np.array(bits, dtype=np.uint32).view(np.float32)


def test_splitmix64_published_vectors(golden):
    for n, want in enumerate(golden["splitmix64_seed0"]):
        assert orc.lib.orc_hash64(0, 0, n) == want
    for n, want in enumerate(golden["splitmix64_seed1234567"]):
        assert orc.lib.orc_hash64(1234567, 0, n) == want


def test_hash64_streams(golden):
    for stream, ctr, want in golden["hash64"]:
        assert orc.lib.orc_hash64(SEED, stream, ctr) == want


def test_uniform_range_and_exactness():
    for c in range(1000):
        x = orc.lib.orc_uniform(SEED, 0, c)
        assert 0.0 <= x < 1.0
        assert float(np.float32(x)) * 2**24 == int(float(x) * 2**24)


def test_kat_hand_computed(kat):
    for mode in (orc.ORDER_SEQ, orc.ORDER_WARP_TREE):
        # k=2 is padded to 4 with zeros for the tree order (k % 4 == 0 on the GPU path); zeros add exactly.
        p = np.array(kat["p"] + [0, 0], dtype=np.float32)
        q = np.array(kat["q"] + [0, 0], dtype=np.float32)
        e = orc.lib.orc_sgd_update(p, q, 4, kat["r"], kat["lr"], kat["lambda"], mode)
        assert abs(e - kat["e"]) < 1e-6
        np.testing.assert_allclose(p[:2], kat["p_new"], rtol=1e-6)
        np.testing.assert_allclose(q[:2], kat["q_new"], rtol=1e-6)
        assert p[2] == 0 and q[3] == 0


def test_kat_k2_sequential(kat):
    p = np.array(kat["p"], dtype=np.float32)
    q = np.array(kat["q"], dtype=np.float32)
    e = orc.lib.orc_sgd_update(p, q, 2, kat["r"], kat["lr"], kat["lambda"], orc.ORDER_SEQ)
    assert abs(e - kat["e"]) < 1e-6
    np.testing.assert_allclose(p, kat["p_new"], rtol=1e-6)
    np.testing.assert_allclose(q, kat["q_new"], rtol=1e-6)


def test_init_factors_golden(golden):
    np.testing.assert_array_equal(orc.init_factors(3, 8, SEED, 0).ravel(), f32(golden["init_p_k8_rows3"]))
    np.testing.assert_array_equal(orc.init_factors(2, 32, SEED, 1).ravel(), f32(golden["init_q_k32_rows2"]))


def test_shuffle_golden_and_permutation(golden):
    assert orc.shuffle(SEED, 0, 64).tolist() == golden["shuffle_e0_n64"]
    assert orc.shuffle(SEED, 3, 64).tolist() == golden["shuffle_e3_n64"]
    big = orc.shuffle(SEED, 5, 100_000)
    assert np.array_equal(np.sort(big), np.arange(100_000))
    assert np.array_equal(big, npr.shuffle(SEED, 5, 100_000))
    assert not np.array_equal(big, orc.shuffle(SEED, 6, 100_000))
    assert orc.shuffle(SEED, 0, 0).size == 0


@pytest.mark.parametrize("name,nu,ni,l2ai", [("ml100k", 943, 1682, 3), ("heavy", 10_000_000, 1_000_000, 4)])
def test_generator_golden(golden, name, nu, ni, l2ai):
    g = golden["gen_" + name]
    u, i, r, held = orc.generate(SEED, 0, 256, nu, ni, 2, 0.25, l2ai, 0.375)
    assert u.tolist() == g["u"] and i.tolist() == g["i"]
    np.testing.assert_array_equal(r, f32(g["r"]))
    assert held.astype(int).tolist() == g["held"]
    u, i, r, held = orc.generate(SEED, 10**9, 64, nu, ni, 2, 0.25, l2ai, 0.375, threads=3)
    assert u.tolist() == g["u_at_1e9"] and i.tolist() == g["i_at_1e9"]
    np.testing.assert_array_equal(r, f32(g["r_at_1e9"]))
    assert held.astype(int).tolist() == g["held_at_1e9"]


def test_generator_shape_properties():
    nu, ni, n = 480_000, 17_800, 2_000_000
    u, i, r, held = orc.generate(SEED, 0, n, nu, ni)
    assert u.min() >= 0 and u.max() < nu and i.min() >= 0 and i.max() < ni
    assert 1.0 <= r.min() and r.max() <= 5.0
    assert abs(held.mean() - 0.1) < 0.002
    counts = np.sort(np.bincount(i, minlength=ni))[::-1]
    top1 = counts[: ni // 100].sum() / n
    assert 0.27 < top1 < 0.33                      # top 1 % of items ~ 30 % of ratings
    assert counts[0] / n < 0.02                    # no single item dominates
    assert 3.3 < r.mean() < 3.7


def test_scatter_id_is_bijection():
    for count in (1, 2, 943, 1682, 17_800, 65_536):
        ids = np.array([orc.lib.orc_scatter_id(x, count) for x in range(count)])
        assert np.array_equal(np.sort(ids), np.arange(count))


def test_sgd_small_run_golden(golden):
    s = golden["sgd_small_shape"]
    u, i, r, _ = orc.generate(SEED, 0, s["n"], s["n_users"], s["n_items"])
    for mode, key in ((orc.ORDER_SEQ, "sgd_small_seq"), (orc.ORDER_WARP_TREE, "sgd_small_tree")):
        P, Q = orc.factorize(u, i, r, s["n_users"], s["n_items"], s["k"], s["lr"], s["lambda"], s["epochs"],
                             SEED, mode)
        np.testing.assert_array_equal(P.ravel(), f32(golden[key]["P"]))
        np.testing.assert_array_equal(Q.ravel(), f32(golden[key]["Q"]))
        assert abs(orc.rmse(P, Q, u, i, r) - golden[key]["rmse"]) < 1e-12


def test_update_k20_golden(golden):
    for mode, key in ((orc.ORDER_SEQ, "update_k20_seq"), (orc.ORDER_WARP_TREE, "update_k20_tree")):
        p = orc.init_factors(1, 20, SEED, 0, 0.5)[0].copy()
        q = orc.init_factors(1, 20, SEED, 1, 0.5)[0].copy()
        e = orc.lib.orc_sgd_update(p, q, 20, 3.25, 0.05, 0.02, mode)
        np.testing.assert_array_equal(p, f32(golden[key]["p"]))
        np.testing.assert_array_equal(q, f32(golden[key]["q"]))
        assert np.float32(e) == f32(golden[key]["e"])[0]


def test_oracle_vs_numpy_1k_ratings_k8():
    """SURVEY section 4 'oracle unit': C++ oracle vs NumPy float32 restatement, bit for bit."""
    nu, ni, k = 50, 80, 8
    u, i, r, _ = orc.generate(SEED + 1, 0, 1000, nu, ni)
    Pn, Qn = npr.factorize(u, i, r, nu, ni, k, 0.03, 0.02, 1, SEED + 1)
    Po, Qo = orc.factorize(u, i, r, nu, ni, k, 0.03, 0.02, 1, SEED + 1)
    np.testing.assert_array_equal(Pn, Po)
    np.testing.assert_array_equal(Qn, Qo)


def test_tree_and_seq_orders_agree_per_update_1e5():
    """Teacher-forced: from the same pre-update rows both summation orders land within 1e-5 relative."""
    nu, ni, k = 943, 1682, 32
    u, i, r, held = orc.generate(SEED, 0, 20_000, nu, ni)
    P = orc.init_factors(nu, k, SEED, 0)
    Q = orc.init_factors(ni, k, SEED, 1)
    order, pre_p, pre_q, post_p, post_q, err = orc.train_tape(u, i, r, P, Q, 0.01, 0.05, 0, SEED)
    for j in range(0, len(r), 7):
        p, q = pre_p[j].copy(), pre_q[j].copy()
        orc.lib.orc_sgd_update(p, q, k, r[order[j]], 0.01, 0.05, orc.ORDER_WARP_TREE)
        np.testing.assert_allclose(p, post_p[j], rtol=1e-5, atol=1e-8)
        np.testing.assert_allclose(q, post_q[j], rtol=1e-5, atol=1e-8)


def test_fma_arrangement_is_the_same_rule_to_a_few_ulp():
    """ORDER_WARP_TREE_FMA (the GPU full-grid arithmetic) vs the stand-in's rule, teacher-forced: << 1e-5."""
    nu, ni, k = 943, 1682, 32
    u, i, r, held = orc.generate(SEED, 0, 20_000, nu, ni)
    P = orc.init_factors(nu, k, SEED, 0)
    Q = orc.init_factors(ni, k, SEED, 1)
    order, pre_p, pre_q, post_p, post_q, err = orc.train_tape(u, i, r, P, Q, 0.01, 0.05, 0, SEED)
    worst = 0.0
    for j in range(0, len(r), 5):
        p, q = pre_p[j].copy(), pre_q[j].copy()
        e = orc.lib.orc_sgd_update(p, q, k, r[order[j]], 0.01, 0.05, orc.ORDER_WARP_TREE_FMA)
        worst = max(worst, np.max(np.abs(p - post_p[j]) / np.maximum(np.abs(post_p[j]), 1e-6)),
                    np.max(np.abs(q - post_q[j]) / np.maximum(np.abs(post_q[j]), 1e-6)))
        assert abs(e - err[j]) <= 1e-5 * max(1.0, abs(err[j]))
    assert worst < 2e-6


def test_trace_and_tape_consistent():
    nu, ni, k = 30, 40, 8
    u, i, r, _ = orc.generate(SEED, 0, 500, nu, ni)
    P1, Q1 = orc.init_factors(nu, k, SEED, 0), orc.init_factors(ni, k, SEED, 1)
    P2, Q2 = P1.copy(), Q1.copy()
    tr = orc.train(u, i, r, P1, Q1, 0.02, 0.05, 0, 1, SEED, trace=True)
    order, pre_p, pre_q, post_p, post_q, err = orc.train_tape(u, i, r, P2, Q2, 0.02, 0.05, 0, SEED)
    np.testing.assert_array_equal(tr, err)
    np.testing.assert_array_equal(P1, P2)
    assert np.array_equal(order, orc.shuffle(SEED, 0, 500))


def test_rejects_out_of_range_ids():
    u = np.array([0, 5], dtype=np.int32)
    i = np.array([0, 0], dtype=np.int32)
    r = np.ones(2, dtype=np.float32)
    with pytest.raises(ValueError):
        orc.factorize(u, i, r, 5, 3, 8, 0.1, 0.1, 1, 1)


def test_empty_input_is_identity():
    P, Q = orc.factorize(np.empty(0, np.int32), np.empty(0, np.int32), np.empty(0, np.float32), 3, 4, 8,
                         0.1, 0.1, 5, SEED)
    np.testing.assert_array_equal(P, orc.init_factors(3, 8, SEED, 0))
    np.testing.assert_array_equal(Q, orc.init_factors(4, 8, SEED, 1))
    assert orc.rmse(P, Q, np.empty(0, np.int32), np.empty(0, np.int32), np.empty(0, np.float32)) == 0.0


def test_ml100k_shaped_converges_and_hogwild_matches():
    """configs[0]: 943 x 1682, 100K ratings, k=32, 20 epochs, sequential and threaded."""
    nu, ni, k, lr, lam, epochs = 943, 1682, 32, 0.01, 0.05, 20
    u, i, r, held = orc.generate(SEED, 0, 100_000, nu, ni)
    tu, ti, tr = u[~held].copy(), i[~held].copy(), r[~held].copy()
    hu, hi, hr = u[held].copy(), i[held].copy(), r[held].copy()
    P, Q = orc.factorize(tu, ti, tr, nu, ni, k, lr, lam, epochs, SEED)
    seq = orc.rmse(P, Q, hu, hi, hr)
    assert seq < 0.45
    P2, Q2 = orc.init_factors(nu, k, SEED, 0), orc.init_factors(ni, k, SEED, 1)
    orc.train_hogwild(tu, ti, tr, P2, Q2, lr, lam, 0, epochs, SEED, threads=4)
    hog = orc.rmse(P2, Q2, hu, hi, hr)
    assert abs(hog - seq) / seq < 0.015            # thread interleaving varies with host load; the GPU bar (0.5 %) is tested on the GPU
    Pt, Qt = orc.factorize(tu, ti, tr, nu, ni, k, lr, lam, epochs, SEED, orc.ORDER_WARP_TREE)
    assert abs(orc.rmse(Pt, Qt, hu, hi, hr) - seq) / seq < 1e-4


@pytest.mark.parametrize("k,lanes", [(8, 8), (32, 8), (64, 16), (100, 32), (128, 32), (512, 32)])
def test_tree_lanes_override_matches_numpy(k, lanes):
    """orc.tree_lanes(L): the run kernel's summation geometry (never fewer than 8 lanes per rating) -- the oracle against
    the independent NumPy restatement, bit for bit, and back to the default geometry outside the with-block."""
    assert orc.run_lanes(k) == lanes
    rng = np.random.default_rng(k)
    p = rng.standard_normal(k).astype(np.float32)
    q = rng.standard_normal(k).astype(np.float32)
    want = np.float32(1.5) - npr.dot_warp_tree(p, q, lanes)
    with orc.tree_lanes(lanes):
        e = orc.lib.orc_sgd_update(p.copy(), q.copy(), k, 1.5, 0.01, 0.05, orc.ORDER_WARP_TREE)
    assert np.float32(e) == np.float32(want)
    e_default = orc.lib.orc_sgd_update(p.copy(), q.copy(), k, 1.5, 0.01, 0.05, orc.ORDER_WARP_TREE)
    assert np.float32(e_default) == np.float32(np.float32(1.5) - npr.dot_warp_tree(p, q))


# ------------------------------------------------------------------------------------------------
# round 2: signal-dominant generator variant, the engine's bucket permutation, the run kernel's twin
# ------------------------------------------------------------------------------------------------
SIGNAL = dict(amplitude=1.7320508, noise_scale=0.125)


def test_signal_variant_generator_golden_and_numpy(golden):
    g = golden["gen_ml100k_signal"]
    u, i, r, held = orc.generate(SEED, 0, 256, 943, 1682, **SIGNAL)
    assert u.tolist() == g["u"] and i.tolist() == g["i"] and r.view(np.uint32).tolist() == g["r"]
    assert held.astype(int).tolist() == g["held"]
    nu, ni, nr, nh = npr.generate(SEED, 5000, 300, 480_000, 17_800, 2, 0.25, 3, 0.375, 1.7320508, 0.125)
    ou, oi, orr, oh = orc.generate(SEED, 5000, 300, 480_000, 17_800, **SIGNAL)
    assert np.array_equal(nu, ou) and np.array_equal(ni, oi) and np.array_equal(nr.view(np.uint32), orr.view(np.uint32))
    # same users, items and split as the default variant: only the ratings differ
    du, di, dr, dh = orc.generate(SEED, 5000, 300, 480_000, 17_800)
    assert np.array_equal(du, ou) and np.array_equal(di, oi) and np.array_equal(dh, oh) and not np.array_equal(dr, orr)


def test_signal_variant_is_signal_dominant():
    """The variant exists so that RMSE parity bites: the planted signal (std ~1) dwarfs the noise (std 0.07), and a
    factorisation that only learns the mean scores ~0.94 while the sequential oracle gets far below it."""
    nu, ni, n, k = 2000, 800, 300_000, 16
    u, i, r, held = orc.generate(SEED, 0, n, nu, ni, **SIGNAL)
    tr = (u[~held].copy(), i[~held].copy(), r[~held].copy())
    ho = (u[held].copy(), i[held].copy(), r[held].copy())
    const = float(np.sqrt(np.mean((ho[2] - tr[2].mean()) ** 2)))
    assert 0.85 < const < 1.05
    P, Q = orc.factorize(*tr, nu, ni, k, 0.02, 0.02, 12, SEED)
    assert orc.rmse(P, Q, *ho) < 0.5 * const


@pytest.mark.parametrize("n", [1, 2, 5, 31, 32, 33, 63, 64, 65, 100, 1000, 4097, 70_001])
def test_block_perm_is_a_tile_coherent_bijection(n):
    p0 = orc.block_perm(n, SEED, 0, 17)
    assert sorted(p0.tolist()) == list(range(n))
    full = (n // 32) * 32
    # positions of one aligned group of 32 read one aligned group of 32 records (whole sectors); the tail stays the tail
    assert np.all((p0[:full].reshape(-1, 32) // 32) == (p0[:full:32] // 32)[:, None])
    assert np.all(p0[full:] >= full)
    if n >= 64:
        p1, pb = orc.block_perm(n, SEED, 1, 17), orc.block_perm(n, SEED, 0, 18)
        assert not np.array_equal(p0, p1) and not np.array_equal(p0, pb)      # keyed by epoch and by bucket
        assert np.mean(p0 != np.arange(n)) > 0.5


def _run_plan_one_bucket_per_item(counts, chunk, weights=None):
    """Units of a launch over consecutive per-item buckets of the given sizes, each cut into ceil(n / chunk) equal runs."""
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    start, count, item, weight = [], [], [], []
    for it, n in enumerate(counts):
        pieces = -(-n // chunk)
        for pc in range(pieces):
            lo, hi = off[it] + n * pc // pieces, off[it] + n * (pc + 1) // pieces
            start.append(lo); count.append(hi - lo); item.append(it)
            weight.append((weights or {}).get(pieces, 1.0 / pieces))
    return orc.RunPlan(start, count, item, weight, off), off


@pytest.mark.parametrize("k,order", [(8, orc.ORDER_SEQ), (32, orc.ORDER_WARP_TREE), (128, orc.ORDER_WARP_TREE_FMA)])
def test_run_twin_with_one_run_per_item_is_the_sequential_rule(k, order):
    rng = np.random.default_rng(k)
    counts = [40, 1, 300, 17, 64]
    n = sum(counts)
    u = rng.permutation(n).astype(np.int32)                    # pairwise distinct users: runs commute exactly
    r = (1 + 4 * rng.random(n)).astype(np.float32)
    plan, off = _run_plan_one_bucket_per_item(counts, 4096)
    items = np.repeat(np.arange(len(counts)), counts).astype(np.int32)
    P, Q = orc.init_factors(n, k, SEED, 0), orc.init_factors(len(counts), k, SEED, 1)
    Ps, Qs = P.copy(), Q.copy()
    orc.train_runs_launch(u, r, plan, 0, len(counts), P, Q, 0.02, 0.03, order, resident=2)
    orc.train(u, items, r, Ps, Qs, 0.02, 0.03, 0, 1, SEED, order, shuffled=False)
    assert np.array_equal(P, Ps) and np.array_equal(Q, Qs)


@pytest.mark.parametrize("resident,gpw", [(64, 1), (4, 1), (8, 4)])
def test_run_twin_weighted_merge_by_hand(resident, gpw):
    """Runs of an item that are in flight together all start from the launch-start q_i and add weight * (q_run - q_start);
    with fewer sub-warps than runs, later runs start from what the earlier ones merged."""
    k, lr, lam, chunk = 16, 0.02, 0.03, 32
    rng = np.random.default_rng(3)
    counts = [128, 96]
    n = sum(counts)
    u = rng.permutation(n).astype(np.int32)
    r = (1 + 4 * rng.random(n)).astype(np.float32)
    plan, off = _run_plan_one_bucket_per_item(counts, chunk, weights={4: 0.3125, 3: 0.4})
    P0, Q0 = orc.init_factors(n, k, SEED, 0), orc.init_factors(2, k, SEED, 1)
    P, Q = P0.copy(), Q0.copy()
    orc.train_runs_launch(u, r, plan, 0, len(plan.start), P, Q, lr, lam, orc.ORDER_SEQ, resident=resident, gpw=gpw)
    # by hand: waves of `resident` units in claim order (all runs equally long here except item 1's: 32 each too)
    Ph, Qh = P0.copy(), Q0.copy()
    units = list(range(len(plan.start)))
    while units:
        wave, units = units[:resident], units[resident:]
        start_q = Qh.copy()
        for j in wave:
            it = int(plan.item[j])
            q = start_q[it].copy()
            for t in range(int(plan.start[j]), int(plan.start[j] + plan.count[j])):
                orc.lib.orc_sgd_update(Ph[u[t]], q, k, float(r[t]), lr, lam, orc.ORDER_SEQ)
            Qh[it] = Qh[it] + (q - start_q[it]) * np.float32(plan.weight[j])
    assert np.array_equal(P, Ph)
    np.testing.assert_allclose(Q, Qh, rtol=1e-6, atol=1e-8)      # the adds of one wave come in slot order on both sides
    assert not np.allclose(Q, Q0)


def test_run_twin_reads_buckets_through_the_permutation():
    k = 8
    rng = np.random.default_rng(5)
    counts = [200, 70]
    n = sum(counts)
    u = rng.permutation(n).astype(np.int32)
    r = (1 + 4 * rng.random(n)).astype(np.float32)
    plan, off = _run_plan_one_bucket_per_item(counts, 4096)
    P, Q = orc.init_factors(n, k, SEED, 0), orc.init_factors(2, k, SEED, 1)
    orc.train_runs_launch(u, r, plan, 0, 2, P, Q, 0.02, 0.03, orc.ORDER_SEQ, resident=2, virt=True, seed=SEED, epoch=4)
    order = np.concatenate([off[b] + orc.block_perm(counts[b], SEED, 4, int(plan.bid[b])) for b in range(2)])
    items = np.repeat(np.arange(2), counts).astype(np.int32)
    Ps, Qs = orc.init_factors(n, k, SEED, 0), orc.init_factors(2, k, SEED, 1)
    orc.train(u[order].copy(), items, r[order].copy(), Ps, Qs, 0.02, 0.03, 0, 1, SEED, orc.ORDER_SEQ, shuffled=False)
    assert np.array_equal(P, Ps) and np.array_equal(Q, Qs)


# ------------------------------------------------------------------------------------------------
# model extension: global mean + biases -- stand-in factorizeModel :305, sgdUpdateModel :282
# ------------------------------------------------------------------------------------------------
def _np_model_epochs(u, i, rc, P, Q, bu, bi, lr, lam, epochs, seed):
    """Independent NumPy float32 restatement of factorizeModel's loop (sequential dot, one rounding per operation)."""
    f = np.float32
    lr, lam = f(lr), f(lam)
    for ep in range(epochs):
        for t in npr.shuffle(seed, ep, len(rc)):
            p, q = P[u[t]], Q[i[t]]
            dot = f(0)
            for a, b in zip(p, q):
                dot = f(dot + f(a * b))
            if bu is not None:
                pred = f(f(dot + bu[u[t]]) + bi[i[t]])
                e = f(rc[t] - pred)
                b0, b1 = bu[u[t]], bi[i[t]]
                bu[u[t]] = f(b0 + f(lr * f(e - f(lam * b0))))
                bi[i[t]] = f(b1 + f(lr * f(e - f(lam * b1))))
            else:
                e = f(rc[t] - dot)
            pn = (p + lr * (e * q - lam * p).astype(f)).astype(f)
            qn = (q + lr * (e * p - lam * q).astype(f)).astype(f)
            P[u[t]], Q[i[t]] = pn, qn


def test_global_mean_is_an_exact_integer_sum():
    rng = np.random.default_rng(1)
    r = (1 + 4 * rng.random(100_001)).astype(np.float32)
    mu = orc.global_mean(r)
    s = int(np.floor(r.astype(np.float64) * 1048576.0).astype(np.int64).sum())
    assert mu == np.float32(s / len(r) / 1048576.0)
    assert orc.global_mean(r[::-1].copy()) == mu and orc.global_mean(rng.permutation(r)) == mu     # any order
    assert abs(mu - r.astype(np.float64).mean()) < 1e-6
    assert orc.global_mean(r[:0]) == 0.0


@pytest.mark.parametrize("use_mean,use_bias", [(True, True), (False, True), (True, False)])
def test_model_extension_matches_numpy(use_mean, use_bias):
    nu, ni, n, k = 30, 40, 400, 8
    u, i, r, _ = orc.generate(SEED, 0, n, nu, ni)
    mu = orc.global_mean(r) if use_mean else 0.0
    rc = (r - np.float32(mu)).astype(np.float32)
    P, Q = orc.init_factors(nu, k, SEED, 0), orc.init_factors(ni, k, SEED, 1)
    bu = np.zeros(nu, np.float32) if use_bias else None
    bi = np.zeros(ni, np.float32) if use_bias else None
    Pn, Qn = P.copy(), Q.copy()
    bun, bin_ = (bu.copy(), bi.copy()) if use_bias else (None, None)
    orc.train_model(u, i, rc, P, Q, bu, bi, 0.02, 0.05, 0, 2, SEED)
    _np_model_epochs(u, i, rc, Pn, Qn, bun, bin_, 0.02, 0.05, 2, SEED)
    assert np.array_equal(P, Pn) and np.array_equal(Q, Qn)
    if use_bias:
        assert np.array_equal(bu, bun) and np.array_equal(bi, bin_) and np.abs(bu).max() > 0
    # without biases the extension is the reference rule on centred ratings
    if not use_bias:
        Pr, Qr = orc.init_factors(nu, k, SEED, 0), orc.init_factors(ni, k, SEED, 1)
        orc.train(u, i, rc, Pr, Qr, 0.02, 0.05, 0, 2, SEED)
        assert np.array_equal(P, Pr) and np.array_equal(Q, Qr)
    want = np.sqrt(np.mean([(float(rc[t]) - float(np.float32(np.float32(np.dot(P[u[t]].astype(np.float64), Q[i[t]].astype(np.float64))) +
                                                  (bu[u[t]] if use_bias else 0) + (bi[i[t]] if use_bias else 0)))) ** 2 for t in range(n)]))
    assert abs(orc.rmse_model(P, Q, bu, bi, u, i, rc) - want) / want < 1e-5


def test_model_extension_learns_past_the_mean_on_the_default_data():
    """Why the extension matters (round-1 review): on the noise-dominant sets plain MF ends above the constant predictor;
    with the global mean and biases the same rule, same epochs, ends below it."""
    nu, ni, n, k = 2000, 800, 300_000, 16
    u, i, r, held = orc.generate(SEED, 0, n, nu, ni)
    tr = (u[~held].copy(), i[~held].copy(), r[~held].copy())
    ho = (u[held].copy(), i[held].copy(), r[held].copy())
    mu = orc.global_mean(tr[2])
    const = float(np.sqrt(np.mean((ho[2] - np.float32(mu)) ** 2)))
    P, Q = orc.factorize(*tr, nu, ni, k, 0.005, 0.05, 10, SEED)
    plain = orc.rmse(P, Q, *ho)
    P, Q = orc.init_factors(nu, k, SEED, 0), orc.init_factors(ni, k, SEED, 1)
    bu, bi = np.zeros(nu, np.float32), np.zeros(ni, np.float32)
    orc.train_model(tr[0], tr[1], (tr[2] - np.float32(mu)).astype(np.float32), P, Q, bu, bi, 0.005, 0.05, 0, 10, SEED)
    ext = orc.rmse_model(P, Q, bu, bi, ho[0], ho[1], (ho[2] - np.float32(mu)).astype(np.float32))
    assert plain > const and ext < plain, (const, plain, ext)


def test_run_twin_carries_biases_like_q():
    k, lr, lam, chunk = 8, 0.02, 0.03, 32
    rng = np.random.default_rng(9)
    counts = [128, 64]
    n = sum(counts)
    u = rng.permutation(n).astype(np.int32)
    r = (rng.random(n) - 0.5).astype(np.float32)
    plan, off = _run_plan_one_bucket_per_item(counts, chunk, weights={4: 0.3125, 2: 0.625})
    P0, Q0 = orc.init_factors(n, k, SEED, 0), orc.init_factors(2, k, SEED, 1)
    P, Q, bu, bi = P0.copy(), Q0.copy(), np.zeros(n, np.float32), np.full(2, 0.25, np.float32)
    orc.train_runs_launch(u, r, plan, 0, len(plan.start), P, Q, lr, lam, orc.ORDER_SEQ, resident=64, bu=bu, bi=bi)
    # by hand: one wave, every run from the launch-start q_i and b_i
    Ph, Qh, buh, bih = P0.copy(), Q0.copy(), np.zeros(n, np.float32), np.full(2, 0.25, np.float32)
    bi_start = bih.copy()
    for j in range(len(plan.start)):
        it = int(plan.item[j])
        q, b = Q0[it].copy(), np.array([bi_start[it]], np.float32)
        for t in range(int(plan.start[j]), int(plan.start[j] + plan.count[j])):
            orc.lib.orc_sgd_update_model(Ph[u[t]], q, k, buh[u[t]:u[t] + 1].ctypes.data, b.ctypes.data, float(r[t]), lr, lam, orc.ORDER_SEQ)
        Qh[it] = Qh[it] + (q - Q0[it]) * np.float32(plan.weight[j])
        bih[it] = bih[it] + (b[0] - bi_start[it]) * np.float32(plan.weight[j])
    assert np.array_equal(P, Ph) and np.array_equal(bu, buh)
    np.testing.assert_allclose(Q, Qh, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(bi, bih, rtol=1e-6, atol=1e-8)
    assert np.abs(bu).max() > 0 and not np.allclose(bi, 0.25)


# ------------------------------------------------------------------------------------------------
# model extension: learning-rate schedule and early stopping -- stand-in learningRate :331, factorizeEarlyStop :350
# ------------------------------------------------------------------------------------------------
def test_learning_rate_schedule_is_repeated_binary32_multiplication():
    lr, d = np.float32(0.02), np.float32(0.9)
    want = lr
    for e in range(12):
        assert np.float32(orc.learning_rate(float(lr), float(d), e)) == want
        want = np.float32(want * d)
    assert orc.learning_rate(0.005, 1.0, 1000) == float(np.float32(0.005))


def _model_set(nu=60, ni=40, n=3000, k=8):
    u, i, r, held = orc.generate(SEED, 0, n, nu, ni)
    mu = np.float32(orc.global_mean(r[~held].copy()))
    tr = (u[~held].copy(), i[~held].copy(), (r[~held] - mu).astype(np.float32))
    va = (u[held].copy(), i[held].copy(), (r[held] - mu).astype(np.float32))
    return nu, ni, k, tr, va


def test_early_stop_loop_is_the_model_loop_with_a_schedule():
    nu, ni, k, tr, va = _model_set()
    for decay in (1.0, 0.8):
        P, Q = orc.init_factors(nu, k, SEED, 0), orc.init_factors(ni, k, SEED, 1)
        bu, bi = np.zeros(nu, np.float32), np.zeros(ni, np.float32)
        ran, curve = orc.train_early_stop(*tr, *va, P, Q, bu, bi, 0.02, 0.05, decay, 0, 0.0, 5, SEED)
        assert ran == 5 and len(curve) == 5
        Pm, Qm = orc.init_factors(nu, k, SEED, 0), orc.init_factors(ni, k, SEED, 1)
        bum, bim = np.zeros(nu, np.float32), np.zeros(ni, np.float32)
        for e in range(5):
            orc.train_model(*tr, Pm, Qm, bum, bim, orc.learning_rate(0.02, decay, e), 0.05, e, e + 1, SEED)
            assert curve[e] == orc.rmse_model(Pm, Qm, bum, bim, *va)
        assert np.array_equal(P, Pm) and np.array_equal(Q, Qm) and np.array_equal(bu, bum) and np.array_equal(bi, bim)


def test_early_stop_rule():
    nu, ni, k, tr, va = _model_set()
    fresh = lambda: (orc.init_factors(nu, k, SEED, 0), orc.init_factors(ni, k, SEED, 1), np.zeros(nu, np.float32), np.zeros(ni, np.float32))
    # an improvement of 50 % per epoch never happens after the first epoch: 1 + patience epochs run
    for patience in (1, 3):
        ran, curve = orc.train_early_stop(*tr, *va, *fresh(), 0.02, 0.05, 1.0, patience, 0.5, 20, SEED)
        assert ran == 1 + patience and len(curve) == ran
    # min_delta 0: stops `patience` epochs after the validation curve's first non-improvement, if it has one
    full_ran, full = orc.train_early_stop(*tr, *va, *fresh(), 0.1, 0.0, 1.0, 0, 0.0, 40, SEED)
    assert full_ran == 40
    best, strikes, want = float("inf"), 0, 40
    for e, v in enumerate(full):
        if v < best:
            best, strikes = v, 0
        else:
            strikes += 1
            if strikes >= 2:
                want = e + 1
                break
    ran, curve = orc.train_early_stop(*tr, *va, *fresh(), 0.1, 0.0, 1.0, 2, 0.0, 40, SEED)
    assert ran == want and curve == full[:ran]
    assert want < 40          # lr 0.1 without regularisation overfits this set within 40 epochs


def test_midsize_fixture_is_the_oracles():
    """tests/golden/oracle_rmse_midsize.json (the GPU parity tests' reference curves) re-derived here for its first epochs:
    shuffled order, plain and extended model, and the DSGD block order with 4 x 4 strata."""
    import json
    import os
    fx = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_rmse_midsize.json")))
    nu, ni, n, k = fx["n_users"], fx["n_items"], fx["n_ratings"], fx["k"]
    assert fx["seed"] == SEED
    for variant, model, G, epochs in (("default", False, 0, 2), ("signal", True, 0, 1), ("signal", False, 4, 1)):
        par = fx["params"][variant]
        u, i, r, held = orc.generate(SEED, 0, n, nu, ni, amplitude=par["amplitude"], noise_scale=par["noise_scale"])
        tu, ti, tr = u[~held].copy(), i[~held].copy(), r[~held].copy()
        hu, hi, hr = u[held].copy(), i[held].copy(), r[held].copy()
        assert (len(tr), len(hr)) == (fx[variant]["n_train"], fx[variant]["n_heldout"])
        mu = np.float32(orc.global_mean(tr)) if model else np.float32(0)
        rc, hc = (tr - mu).astype(np.float32), (hr - mu).astype(np.float32)
        P, Q = orc.init_factors(nu, k, SEED, 0), orc.init_factors(ni, k, SEED, 1)
        bu, bi = (np.zeros(nu, np.float32), np.zeros(ni, np.float32)) if model else (None, None)
        want = fx[variant]["model" if model else "plain"]["dsgd%d" % G if G else "shuffled"]
        assert len(want) == par["epochs"]
        for e in range(epochs):
            if G:
                o = orc.dsgd_order(tu, ti, orc.balanced_bounds(tu, nu, G), orc.balanced_bounds(ti, ni, G), SEED, e)
                orc.train_model(tu[o], ti[o], rc[o], P, Q, bu, bi, par["lr"], par["lam"], e, e + 1, SEED, shuffled=False)
            else:
                orc.train_model(tu, ti, rc, P, Q, bu, bi, par["lr"], par["lam"], e, e + 1, SEED)
            assert orc.rmse_model(P, Q, bu, bi, hu, hi, hc) == want[e]
        if not model and not G:        # orc_train_model without biases is orc_train: the plain curves are factorize()'s
            Pf, Qf = orc.factorize(tu, ti, tr, nu, ni, k, par["lr"], par["lam"], epochs, SEED)
            assert np.array_equal(P, Pf) and np.array_equal(Q, Qf)


# ------------------------------------------------------------------------------------------------
# mixed-precision factor storage -- stand-in srWord :413, storeF16Sr :420, sgdUpdateMixed :426, factorizeMixed :439
# ------------------------------------------------------------------------------------------------
def test_binary16_conversions_are_ieee():
    rng = np.random.default_rng(0)
    x = np.concatenate([(rng.standard_normal(50_000) * 0.2).astype(np.float32), (rng.standard_normal(5_000) * 1e-5).astype(np.float32),
                        np.array([0, -0.0, 65504, 65519.9, 65520, 1e-8, 2.98e-8, 2.9802322e-8, 6.1e-5, 6.0e-5, 1.0, -1.0], np.float32)])
    with np.errstate(over="ignore"):
        want = x.astype(np.float16).view(np.uint16)
    got = np.array([orc.lib.orc_f32_to_f16_rn(float(v)) for v in x], np.uint16)
    assert np.array_equal(got, want)                                         # round to nearest even, subnormals, overflow to inf
    h = np.arange(65536, dtype=np.uint16)
    w = orc.widen(h.reshape(1, -1)).ravel()
    ww = h.view(np.float16).astype(np.float32)
    ok = ~np.isnan(ww)
    assert np.array_equal(w.view(np.uint32)[ok], ww.view(np.uint32)[ok])     # widening is exact for every bit pattern


def test_binary16_add_twin_is_ieee():
    """orc_f16_add_rn (twin of the engine's HSUB2 + red.add.f16x2 on heavy users' binary16 rows): one rounding to nearest even of the
    exact sum / difference, signed zeros, subnormals, overflow -- against NumPy through float64 (exact for binary16 operands)."""
    rng = np.random.default_rng(3)
    a = rng.integers(0, 2 ** 16, 60_000).astype(np.uint16)
    b = rng.integers(0, 2 ** 16, 60_000).astype(np.uint16)
    near = rng.random(60_000) < 0.5                       # half the pairs are neighbours-ish, as in an SGD step
    b[near] = (a[near].astype(np.int32) + rng.integers(-40, 41, int(near.sum()))).clip(0, 65535).astype(np.uint16)
    fa, fb = a.view(np.float16), b.view(np.float16)
    ok = np.isfinite(fa) & np.isfinite(fb)
    a, b, fa, fb = a[ok], b[ok], fa[ok].astype(np.float64), fb[ok].astype(np.float64)
    for sub in (0, 1):
        with np.errstate(over="ignore"):
            want = (fa - fb if sub else fa + fb).astype(np.float16).view(np.uint16)
        got = np.array([orc.lib.orc_f16_add_rn(int(x), int(y), sub) for x, y in zip(a, b)], np.uint16)
        assert np.array_equal(got, want)
    # the heavy-row rule lands on the plain stochastic store whenever the binary16 difference is exact
    nu, ni, n, k = 300, 200, 6000, 16
    u, i, r, _ = orc.generate(SEED, 0, n, nu, ni)
    Pa, Qa = orc.init_factors_f16(nu, k, SEED, 0), orc.init_factors(ni, k, SEED, 1)
    Pb, Qb = Pa.copy(), Qa.copy()
    orc.train_mixed(u, i, r, Pa, Qa, 0.01, 0.03, 0, 1, SEED, sr=1)
    orc.train_mixed(u, i, r, Pb, Qb, 0.01, 0.03, 0, 1, SEED, sr=2)
    assert (Pa != Pb).mean() < 0.02 and np.abs(orc.widen(Pa) - orc.widen(Pb)).max() < 2e-3


def test_stochastic_rounding_is_unbiased_and_picks_a_neighbour():
    rng = np.random.default_rng(1)
    for v in (np.float32(0.1234567), np.float32(-0.0312345), np.float32(0.9999)):
        lo = np.float32(np.float16(v))                                       # a neighbour; the other one is one binary16 ulp away
        words = rng.integers(0, 2 ** 32, 40_000, dtype=np.uint64)
        vals = np.array([orc.lib.orc_f16_to_f32(orc.lib.orc_store_f16_sr(float(v), int(w), int(w) % 4)) for w in words])
        uniq = np.unique(vals)
        assert len(uniq) == 2 and uniq[0] < v < uniq[1] and lo in uniq
        ulp = float(uniq[1] - uniq[0])
        assert abs(vals.mean() - float(v)) < 0.01 * ulp                      # unbiased to a hundredth of an ulp (8 random bits + sampling)
    # values binary16 holds exactly are never moved
    for v in (0.5, -0.25, 0.0999755859375):
        for w in (0, 0xFFFFFFFF, 0x12345678):
            assert orc.lib.orc_f16_to_f32(orc.lib.orc_store_f16_sr(v, w, 1)) == v


def test_sr_word_known_values_and_independence_of_order():
    # the hash is a pure function of (seed, epoch, u, i, c): restated here in Python integers
    def word(seed, epoch, u, i, c):
        m = 0xFFFFFFFF
        x = ((seed ^ (seed >> 32)) + u * 0x9E3779B1 + i * 0x85EBCA77 + (epoch * 0x10001 + c) * 0xC2B2AE3D) & m
        x ^= x >> 16
        x = (x * 0x7FEB352D) & m
        x ^= x >> 15
        x = (x * 0x846CA68B) & m
        x ^= x >> 16
        return x
    for args in ((SEED, 0, 0, 0, 0), (SEED, 3, 479_999, 17_799, 31), (2 ** 63 + 5, 19, 9_999_999, 999_999, 127)):
        assert orc.lib.orc_sr_word(*args) == word(*args)
    # two conflict-free updates give the same rows whichever comes first
    k = 8
    P16 = orc.init_factors_f16(2, k, SEED, 0)
    Q = orc.init_factors(2, k, SEED, 1)
    a, b = (P16.copy(), Q.copy()), (P16.copy(), Q.copy())
    for (P_, Q_), order in ((a, (0, 1)), (b, (1, 0))):
        for t in order:
            orc.lib.orc_sgd_update_mixed(P_[t], Q_[t], k, 3.0 + t, 0.02, 0.05, orc.ORDER_SEQ, SEED, 4, t, t, 1)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_mixed_storage_tracks_binary32_training():
    """P kept in binary16 with stochastic rounding ends within 0.3 % of the binary32 oracle's held-out RMSE (noise-dominant and
    signal-dominant data)."""
    nu, ni, n, k = 2000, 800, 300_000, 16
    for amp, ns, lr, lam, ep in ((0.0, 0.0, 0.005, 0.05, 6), (1.7320508, 0.125, 0.02, 0.02, 10)):
        u, i, r, held = orc.generate(SEED, 0, n, nu, ni, amplitude=amp, noise_scale=ns)
        tr = (u[~held].copy(), i[~held].copy(), r[~held].copy())
        ho = (u[held].copy(), i[held].copy(), r[held].copy())
        P, Q = orc.factorize(*tr, nu, ni, k, lr, lam, ep, SEED)
        base = orc.rmse(P, Q, *ho)
        P16, Qm = orc.init_factors_f16(nu, k, SEED, 0), orc.init_factors(ni, k, SEED, 1)
        assert np.array_equal(P16, orc.init_factors(nu, k, SEED, 0).astype(np.float16).view(np.uint16))
        orc.train_mixed(*tr, P16, Qm, lr, lam, 0, ep, SEED)
        got = orc.rmse(orc.widen(P16), Qm, *ho)
        assert abs(got / base - 1.0) < 3e-3, (got, base)


def test_order_spread_of_the_full_size_sets_is_recorded_consistently():
    """tests/golden/order_spread_large.json (tools/order_spread_large.py): sequential executions of the rule on the Netflix-shaped
    sets that differ only in the visiting order. Order 0 is the committed oracle curve itself; the spread at the last epoch is the
    scale against which the 0.5 % parity bar has to be read on these sets."""
    import json, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = json.load(open(os.path.join(root, "tests", "golden", "order_spread_large.json")))["workloads"]
    for name in ("netflix", "netflix_signal"):
        w = d[name]
        fx = json.load(open(os.path.join(root, "tests", "golden", "oracle_rmse_%s.json" % name)))
        assert w["heldout_rmse_per_epoch_per_order"]["0"] == fx["heldout_rmse_per_epoch"][:w["epochs"]]
        assert len(w["heldout_rmse_per_epoch_per_order"]) >= 4 and w["epochs"] == 10
        finals = [c[-1] for c in w["heldout_rmse_per_epoch_per_order"].values()]
        assert abs((max(finals) - min(finals)) / np.mean(finals) - w["final_spread_rel"]) < 1e-12
    assert d["netflix"]["final_spread_rel"] < 0.002          # noise-dominant: the order hardly matters (< 0.2 %)
    assert 0.001 < d["netflix_signal"]["final_spread_rel"] < 0.006      # signal-dominant: a few tenths of a per cent
