"""Short single-GPU run for ncu: Netflix-shaped (or other) workload, a few epochs, chosen scatter mode.
usage: python tools/profile_target.py [workload] [scatter] [stripes] [epochs] [f32|f16]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import matrixfactorizationsgd.java_b200 as mf
wname = sys.argv[1] if len(sys.argv) > 1 else "netflix"
scatter = int(sys.argv[2]) if len(sys.argv) > 2 else 0
stripes = int(sys.argv[3]) if len(sys.argv) > 3 else 0
epochs = int(sys.argv[4]) if len(sys.argv) > 4 else 2
p_storage = mf.capi.STORAGE_F16 if len(sys.argv) > 5 and sys.argv[5] == "f16" else mf.capi.STORAGE_F32
w = mf.WORKLOADS[wname]
cfg = mf.make_config(w.n_users, w.n_items, w.k, w.lr, w.lambda_, seed=mf.SEED, mode=mf.capi.MODE_HOGWILD,
                     stripes_per_gpu=stripes, scatter=scatter, p_storage=p_storage)
with mf.Engine(cfg) as eng:
    eng.generate_synthetic(mf.synth_params_of(w))
    eng.init_factors()
    st = eng.train(epochs)
    print("epoch_ms", [round(s.epoch_ms, 2) for s in st], "launches/epoch", st[-1].update_launches)
