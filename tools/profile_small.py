"""ncu target: 64 user sub-stripes, rounds 1 (each launch = 1/64 of the ratings, the size of one block of an
8-member ring). usage: python tools/profile_small.py [hot_chunk] [epochs]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import matrixfactorizationsgd.java_b200 as mf
chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 0
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 2
w = mf.WORKLOADS["netflix"]
cfg = mf.make_config(w.n_users, w.n_items, w.k, w.lr, w.lambda_, seed=mf.SEED, mode=mf.capi.MODE_HOGWILD,
                     stripes_per_gpu=64, rounds=1, hot_chunk=chunk)
with mf.Engine(cfg) as eng:
    eng.generate_synthetic(mf.synth_params_of(w))
    eng.init_factors()
    st = eng.train(epochs)
    print("epoch_ms", [round(s.epoch_ms, 2) for s in st], "launches/epoch", st[-1].update_launches)
