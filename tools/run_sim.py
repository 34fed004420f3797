#!/usr/bin/env python
"""CPU simulation of the engine's Hogwild / DSGD run path (no GPU needed).

The real planner (libmfsgd.so: mfsgd_plan_layout / mfsgd_plan_runs, host-only hooks) lays out the runs, the oracle's twin of
the run kernel (oracle.cpp orc_train_runs_launch) executes every launch in the engine's visiting order with the engine's
merge rule. Used to study the convergence of merge rules and plans against the sequential oracle before spending GPU
time; tests/test_gpu_parity.py::test_run_kernel_averaged_merge_matches_its_oracle_twin checks the twin against the kernel.

  python tools/run_sim.py --shape mid --variant signal --gpus 1 --epochs 8 [--boost 1.5] [--hot-chunk 0]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle as orc  # noqa: E402
from matrixfactorizationsgd.java_b200 import _capi as capi  # noqa: E402  (host-only planner hooks; no GPU call is made)
from matrixfactorizationsgd.java_b200.engine import make_config  # noqa: E402

SEED = 20261018
MIN_RUN = 16
L2_BYTES = 126 << 20
SMS, CTAS_PER_SM = 148, 3
HOT_OFF = False

SHAPES = {  # n_users, n_items, n_ratings, k, epochs, log2_alpha_item
    "ml100k": (943, 1682, 100_000, 32, 20, 3),
    "mid": (13_800, 2_700, 2_000_000, 32, 8, 3),
    "mid128": (13_800, 2_700, 2_000_000, 128, 8, 3),
    "ml20m_10": (43_600, 8_500, 2_000_000, 128, 10, 3),
    "netflix_50": (9_600, 356, 2_000_000, 128, 10, 3),
    "netflix_10": (48_000, 1_780, 10_000_000, 128, 10, 3),
    "heavy": (100_000, 10_000, 4_000_000, 64, 3, 4),
}
VARIANTS = {"default": (0.0, 0.0, None, None), "signal": (1.7320508, 0.125, 0.02, 0.02)}   # amplitude, noise, lr, lambda


def balanced_bounds(counts, nblocks):
    cum = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    total = cum[-1]
    b = [int(np.searchsorted(cum[:-1], (total * j) // nblocks, side="left")) for j in range(nblocks)]
    b[0] = 0
    return np.array(b + [len(counts)], np.int32)


class Member:
    pass


def build_layout(u, i, r, nu, ni, k, G, mu_cfg, mi_cfg, rounds_cfg, chunk_cfg, mode, multi_process, seed, boost=0.0):
    cfg = make_config(nu, ni, k, 0.01, 0.01, seed=seed, mode=mode, n_gpus=G, stripes_per_gpu=mu_cfg, shards_per_gpu=mi_cfg,
                      rounds=rounds_cfg, hot_chunk=chunk_cfg, world_size=G if multi_process and G > 1 else 1)
    ucnt, icnt = np.bincount(u, minlength=nu), np.bincount(i, minlength=ni)
    # first pass of the planner for mu / mi (the other outputs need the member sizes)
    out = [C.c_int32() for _ in range(4)]
    capi.check(capi.lib.mfsgd_plan_layout(C.byref(cfg), L2_BYTES, len(r) // G, nu // G, len(r) // G, SMS * CTAS_PER_SM, *map(C.byref, out)))
    mu, mi = out[0].value, out[1].value
    UB, IB = G * mu, G * mi
    ub, ib = balanced_bounds(ucnt, UB), balanced_bounds(icnt, IB)
    owner_u = (np.searchsorted(ub, np.arange(nu), side="right") - 1).astype(np.int32)
    owner_i = (np.searchsorted(ib, np.arange(ni), side="right") - 1).astype(np.int32)
    thr = max(1e-6 * len(r), MIN_RUN * mu * G) if not HOT_OFF else 1e18
    hot_items = np.flatnonzero(icnt >= thr).astype(np.int32)
    H = len(hot_items)
    hot_index = np.full(ni, -1, np.int64)
    hot_index[hot_items] = np.arange(H)
    hot_block_lo = np.searchsorted(hot_items, ib).astype(np.int32)
    members = []
    for g in range(G):
        sel = (owner_u[u] >= g * mu) & (owner_u[u] < (g + 1) * mu)
        mu_, mi_, mr = u[sel], i[sel], r[sel]
        sa = owner_u[mu_] - g * mu
        hx = hot_index[mi_]
        bucket = np.where(hx >= 0, mu * IB + sa.astype(np.int64) * H + hx, sa.astype(np.int64) * IB + owner_i[mi_])
        order = np.argsort(bucket, kind="stable")
        m = Member()
        m.g = g
        m.u, m.i, m.r = mu_[order].copy(), mi_[order].copy(), mr[order].copy()
        nblk = mu * (IB + H)
        m.block_off = np.concatenate([[0], np.cumsum(np.bincount(bucket, minlength=nblk))]).astype(np.int64)
        members.append(m)
    m0 = members[0]
    run_records = int(m0.block_off[-1] - m0.block_off[mu * IB])
    per_warp = 32 // orc.run_lanes(k)
    capi.check(capi.lib.mfsgd_plan_layout(C.byref(cfg), L2_BYTES, len(m0.r), int(ub[mu] - ub[0]), run_records, SMS * CTAS_PER_SM,
                                          *map(C.byref, out)))
    rounds, chunk = out[2].value, out[3].value
    for m in members:
        cap = int(((m.block_off[mu * IB + 1:] - m.block_off[mu * IB:-1]) // chunk + rounds + 1).sum()) + 16
        st, ct, it = np.zeros(cap, np.int64), np.zeros(cap, np.int32), np.zeros(cap, np.int32)
        wt = np.zeros(cap, np.float32)
        n = C.c_int64(cap)
        m.visit_units = np.zeros(mu * rounds * IB + 1, np.int32)
        capi.check(capi.lib.mfsgd_plan_runs(capi.ptr(m.block_off), mu, H, IB, capi.ptr(hot_block_lo), capi.ptr(hot_items), rounds,
                                            chunk, seed, m.g, boost, capi.ptr(st), capi.ptr(ct), capi.ptr(it), capi.ptr(wt), C.byref(n),
                                            capi.ptr(m.visit_units)))
        nn = n.value
        m.plan = orc.RunPlan(st[:nn], ct[:nn], it[:nn], wt[:nn], m.block_off, member=m.g)
    info = dict(G=G, mu=mu, mi=mi, rounds=rounds, chunk=chunk, H=H, IB=IB, per_warp=per_warp)
    return members, info


def train_epoch(members, info, P, Q, lr, lam, epoch, seed, order_mode, always_add, virt=True):
    G, mu, mi, rounds, IB, per_warp = (info[x] for x in ("G", "mu", "mi", "rounds", "IB", "per_warp"))
    launches = 0
    for s in range(G):
        for m in members:
            grp = (m.g + s) % G
            for vis in range(mu * rounds):
                rnd, pos = divmod(vis, mu)
                hv = orc.lib.orc_hash64(seed, 10, (epoch << 24) ^ (s << 12) ^ rnd)
                sa = ((hv >> 1) + (pos if (hv & 1) else mu - 1 - pos)) % mu
                ib_lo, ib_hi = grp * mi, (grp + 1) * mi
                blo, bhi = int(m.block_off[sa * IB + ib_lo]), int(m.block_off[sa * IB + ib_hi])
                lo, hi = blo + (bhi - blo) * rnd // rounds, blo + (bhi - blo) * (rnd + 1) // rounds
                if hi > lo:   # cold records: the Hogwild kernel, modelled as sequential
                    orc.train(m.u[lo:hi].copy(), m.i[lo:hi].copy(), m.r[lo:hi].copy(), P, Q, lr, lam, epoch, epoch + 1, seed,
                              order_mode, shuffled=False)
                vkey = (sa * rounds + rnd) * IB
                ulo, uhi = int(m.visit_units[vkey + ib_lo]), int(m.visit_units[vkey + ib_hi])
                if uhi > ulo:
                    n_units = uhi - ulo
                    grid = max(1, min(SMS * CTAS_PER_SM, -(-n_units // (16 * per_warp))))
                    with orc.tree_lanes(orc.run_lanes(P.shape[1])):
                        orc.train_runs_launch(m.u, m.r, m.plan, ulo, uhi, P, Q, lr, lam, order_mode, grid * 8 * per_warp, per_warp,
                                              virt=virt, seed=seed, epoch=epoch, always_add=always_add)
                    launches += 1
    return launches


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="mid")
    ap.add_argument("--variant", default="signal")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--stripes", type=int, default=0)
    ap.add_argument("--shards", type=int, default=0)
    ap.add_argument("--rounds", type=int, default=0)
    ap.add_argument("--hot-chunk", type=int, default=0)
    ap.add_argument("--epochs", type=int, default=0)
    ap.add_argument("--boost", type=float, default=0.0, help="merge over-relaxation: w = min(1, boost / runs); 0 = the library default")
    ap.add_argument("--always-add", action="store_true")
    ap.add_argument("--lr", type=float, default=0.0)
    ap.add_argument("--lam", type=float, default=-1.0)
    ap.add_argument("--multi-process", action="store_true")
    ap.add_argument("--hot-off", action="store_true", help="no run path: every block is walked sequentially in arrival order")
    ap.add_argument("--sms", type=int, default=148, help="simulated SM count (scale the machine with the data)")
    a = ap.parse_args()
    global SMS, HOT_OFF
    SMS = a.sms
    HOT_OFF = a.hot_off
    nu, ni, n, k, epochs, l2ai = SHAPES[a.shape]
    amp, noise, vlr, vlam = VARIANTS[a.variant]
    lr = a.lr or vlr or (0.01 if a.shape == "ml100k" else 0.005)
    lam = a.lam if a.lam >= 0 else (vlam if vlam is not None else 0.05)
    epochs = a.epochs or epochs
    u, i, r, held = orc.generate(SEED, 0, n, nu, ni, 2, 0.25, l2ai, 0.375, amplitude=amp, noise_scale=noise)
    tu, ti, tr = u[~held].copy(), i[~held].copy(), r[~held].copy()
    hu, hi, hr = u[held].copy(), i[held].copy(), r[held].copy()
    const = float(np.sqrt(((hr - tr.astype(np.float64).mean()) ** 2).mean()))
    mode = capi.MODE_HOGWILD if a.gpus == 1 else capi.MODE_DSGD
    t0 = time.time()
    members, info = build_layout(tu, ti, tr, nu, ni, k, a.gpus, a.stripes, a.shards, a.rounds, a.hot_chunk, mode, a.multi_process, SEED, a.boost)
    Po, Qo = orc.init_factors(nu, k, SEED, 0), orc.init_factors(ni, k, SEED, 1)
    Ps, Qs = Po.copy(), Qo.copy()
    want, got = [], []
    for e in range(epochs):
        orc.train(tu, ti, tr, Po, Qo, lr, lam, e, e + 1, SEED)
        want.append(orc.rmse(Po, Qo, hu, hi, hr))
        nl = train_epoch(members, info, Ps, Qs, lr, lam, e, SEED, orc.ORDER_WARP_TREE_FMA, a.always_add)
        got.append(orc.rmse(Ps, Qs, hu, hi, hr))
    print(json.dumps({"shape": a.shape, "variant": a.variant, "lr": lr, "lambda": lam, "layout": info, "boost": a.boost,
                      "launches_per_epoch": nl, "units": len(members[0].plan.start),
                      "records_in_multi_run_slices": float((members[0].plan.count[members[0].plan.weight < 1.0]).sum() / max(1, members[0].plan.count.sum())),
                      "const_predictor_rmse": const, "oracle": want, "sim": got,
                      "rel_final": got[-1] / want[-1] - 1, "rel_per_epoch": [g / w - 1 for g, w in zip(got, want)],
                      "seconds": time.time() - t0}))


if __name__ == "__main__":
    main()
