"""CPU simulation of GPU-scale Hogwild staleness (design aid, not product, not a test).

Model: W updates are 'in flight' together: all read P,Q as of the window start, then write.
  plain : last writer wins per row (what st.global scatter does under a race)
  atomic: deltas of colliding rows add up (what red.global.add would do)
Compares held-out RMSE per epoch with the sequential oracle on an ML-20M-shaped (scaled) workload.
"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import pyoracle as orc

SEED = 20261018

def run(nu, ni, n, k, lr, lam, epochs, W, mode, u, i, r, hu, hi, hr):
    P = orc.init_factors(nu, k, SEED, 0); Q = orc.init_factors(ni, k, SEED, 1)
    out = []
    rng = np.random.default_rng(1)
    for ep in range(epochs):
        perm = rng.permutation(len(r))
        for s in range(0, len(r), W):
            b = perm[s:s+W]
            ub, ib, rb = u[b], i[b], r[b]
            p = P[ub]; q = Q[ib]
            e = (rb - np.einsum('ij,ij->i', p, q)).astype(np.float32)[:, None]
            dp = lr * (e * q - lam * p); dq = lr * (e * p - lam * q)
            if mode == "plain":
                P[ub] = p + dp; Q[ib] = q + dq
            else:
                for M, idx, d in ((P, ub, dp), (Q, ib, dq)):
                    o = np.argsort(idx, kind="stable"); si = idx[o]
                    starts = np.flatnonzero(np.r_[True, si[1:] != si[:-1]])
                    M[si[starts]] += np.add.reduceat(d[o], starts, axis=0)
        out.append(orc.rmse(P, Q, hu, hi, hr))
    return out

if __name__ == "__main__":
    nu, ni, n, k = 138_000, 27_000, 20_000_000, 32
    lr, lam, epochs = 0.005, 0.05, 6
    if len(sys.argv) > 1: n = int(sys.argv[1])
    u, i, r, held = orc.generate(SEED, 0, n, nu, ni)
    tu, ti, tr = u[~held].copy(), i[~held].copy(), r[~held].copy()
    hu, hi, hr = u[held].copy(), i[held].copy(), r[held].copy()
    P = orc.init_factors(nu, k, SEED, 0); Q = orc.init_factors(ni, k, SEED, 1)
    seq = []
    for ep in range(epochs):
        orc.train(tu, ti, tr, P, Q, lr, lam, ep, ep + 1, SEED)
        seq.append(orc.rmse(P, Q, hu, hi, hr))
    print("seq   ", ["%.5f" % x for x in seq], flush=True)
    for W in (8192, 32768):
        for mode in ("plain", "atomic"):
            t0 = time.time()
            res = run(nu, ni, n, k, lr, lam, epochs, W, mode, tu, ti, tr, hu, hi, hr)
            print(mode, W, ["%.5f" % x for x in res], "rel@end %.4f%%" % (100 * (res[-1] - seq[-1]) / seq[-1]),
                  "%.0fs" % (time.time() - t0), flush=True)
