"""Regenerates tests/golden/*.json from the independent NumPy restatement (tests/np_restatement.py).

The reference holds no golden vectors (README.md only), so these fixtures are project-made pins:
they freeze the written spec as evaluated by NumPy, and the C++ oracle and the
CUDA kernels must reproduce them bit for bit.   Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import np_restatement as npr  # noqa: E402

SEED = 20261018


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).ravel().tolist()


def main():
    g = {"seed": SEED}
    # SplitMix64 published outputs (seed 0 and seed 1234567): hash64(seed, 0, n) is its (n+1)-th output.
    g["splitmix64_seed0"] = [0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4, 0x06C45D188009454F]
    g["splitmix64_seed1234567"] = [6457827717110365317, 3203168211198807973, 9817491932198370423,
                                   4593380528125082431, 16408922859458223821]
    g["hash64"] = [[s, c, int(npr.hash64(SEED, s, c))] for s in range(9) for c in (0, 1, 12345, 2**32 + 7)]
    g["init_p_k8_rows3"] = bits(npr.init_factors(3, 8, SEED, 0, npr.default_init_scale(8)))
    g["init_q_k32_rows2"] = bits(npr.init_factors(2, 32, SEED, 1, npr.default_init_scale(32)))
    g["shuffle_e0_n64"] = npr.shuffle(SEED, 0, 64).tolist()
    g["shuffle_e3_n64"] = npr.shuffle(SEED, 3, 64).tolist()
    for name, (nu, ni, l2ai) in {"ml100k": (943, 1682, 3), "heavy": (10_000_000, 1_000_000, 4)}.items():
        u, i, r, held = npr.generate(SEED, 0, 256, nu, ni, 2, 0.25, l2ai, 0.375)
        u2, i2, r2, held2 = npr.generate(SEED, 10**9, 64, nu, ni, 2, 0.25, l2ai, 0.375)
        g["gen_" + name] = {"u": u.tolist(), "i": i.tolist(), "r": bits(r), "held": held.astype(int).tolist(),
                            "u_at_1e9": u2.tolist(), "i_at_1e9": i2.tolist(), "r_at_1e9": bits(r2),
                            "held_at_1e9": held2.astype(int).tolist()}
    # the signal-dominant variant: planted amplitude 1.7320508, noise scale 0.125
    u, i, r, held = npr.generate(SEED, 0, 256, 943, 1682, 2, 0.25, 3, 0.375, 1.7320508, 0.125)
    g["gen_ml100k_signal"] = {"u": u.tolist(), "i": i.tolist(), "r": bits(r), "held": held.astype(int).tolist()}
    # a small SGD run: 40 users x 60 items, 1000 ratings, k=8, 2 epochs, both summation orders
    nu, ni, k, lr, lam = 40, 60, 8, 0.02, 0.05
    u, i, r, held = npr.generate(SEED, 0, 1000, nu, ni, 2, 0.25, 3, 0.375)
    for tree in (False, True):
        P, Q = npr.factorize(u, i, r, nu, ni, k, lr, lam, 2, SEED, tree)
        g["sgd_small_tree" if tree else "sgd_small_seq"] = {
            "P": bits(P), "Q": bits(Q), "rmse": npr.rmse(P, Q, u, i, r)}
    g["sgd_small_shape"] = {"n_users": nu, "n_items": ni, "k": k, "lr": lr, "lambda": lam, "epochs": 2, "n": 1000}
    # k=20 (non power-of-two chunk count) single update, both orders
    rng_p = npr.init_factors(1, 20, SEED, 0, 0.5)[0]
    rng_q = npr.init_factors(1, 20, SEED, 1, 0.5)[0]
    for tree in (False, True):
        p, q = rng_p.copy(), rng_q.copy()
        e = npr.sgd_update(p, q, 3.25, 0.05, 0.02, tree)
        g["update_k20_tree" if tree else "update_k20_seq"] = {"p": bits(p), "q": bits(q), "e": bits([e])}
    with open(os.path.join(HERE, "golden_small.json"), "w") as f:
        json.dump(g, f)
    kat = {"k": 2, "p": [0.1, 0.2], "q": [0.3, 0.4], "r": 1.0, "lr": 0.1, "lambda": 0.01,
           "dot": 0.11, "e": 0.89, "p_new": [0.1266, 0.2354], "q_new": [0.3086, 0.4174],
           "note": "hand-computed in exact decimal arithmetic; q_new uses the OLD p (simultaneous update)"}
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1)
    print("wrote golden_small.json, kat.json")


if __name__ == "__main__":
    main()
