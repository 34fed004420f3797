"""Phase timings (MFSGD_TRACE=1) of the resident-caller sequence bench.py times as e2e for N > 1: a live handle is
re-loaded from host triplets, re-initialised, trained and read back. usage: python tools/reload_trace.py [epochs] [workload]"""
import ctypes as C, os, sys, time
import numpy as np
os.environ["MFSGD_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import matrixfactorizationsgd.java_b200 as mf
capi = mf.capi
sys.path.insert(0, ROOT)
import bench
epochs = int(sys.argv[1]) if len(sys.argv) > 1 else 30
w = mf.WORKLOADS[sys.argv[2] if len(sys.argv) > 2 else "netflix"]
eng = mf.Engine(mf.make_config(w.n_users, w.n_items, w.k, w.lr, w.lambda_, seed=mf.SEED, mode=capi.MODE_HOGWILD, flags=capi.FLAG_TIME_KERNELS))
eng.generate_synthetic(mf.synth_params_of(w))
eng.init_factors()
eng.train(3)
hu, hi, hr, pins = bench.host_training_set(mf, w, 0, pinned=True)
P = np.empty((w.n_users, w.k), np.float32); Q = np.empty((w.n_items, w.k), np.float32)
for rep in range(2):
    t0 = time.time()
    capi.check(capi.lib.mfsgd_load_ratings(eng._h, capi.ptr(hu), capi.ptr(hi), capi.ptr(hr), len(hr)))
    t1 = time.time()
    eng.init_factors()
    t2 = time.time()
    eng.train(epochs, want_stats=False)
    t3 = time.time()
    capi.check(capi.lib.mfsgd_get_factors(eng._h, capi.ptr(P), capi.ptr(Q)))
    t4 = time.time()
    print("rep %d: load %.3f init %.3f train %.3f get %.3f total %.3f s" % (rep, t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0), flush=True)
# (b) the same load on a fresh handle, while the first one is alive and after it is closed
def fresh(tag):
    e2 = mf.Engine(mf.make_config(w.n_users, w.n_items, w.k, w.lr, w.lambda_, seed=mf.SEED, mode=capi.MODE_HOGWILD))
    t0 = time.time()
    capi.check(capi.lib.mfsgd_load_ratings(e2._h, capi.ptr(hu), capi.ptr(hi), capi.ptr(hr), len(hr)))
    print("%s: fresh-handle load %.3f s" % (tag, time.time() - t0), flush=True)
    e2.close()
fresh("old handle alive")
eng.close()
fresh("old handle closed")
# (c) raw H2D speed of the same pinned buffers, through torch (no libmfsgd code involved)
import torch
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.time()
    x = torch.from_numpy(hu).cuda(); y = torch.from_numpy(hi).cuda(); z = torch.from_numpy(hr).cuda()
    torch.cuda.synchronize()
    print("torch H2D of the three arrays: %.3f s (%.1f GB/s)" % (time.time() - t0, 12e-9 * len(hr) / (time.time() - t0)), flush=True)

# (d) the one-shot entry point bench.py times at N = 1
cfg = mf.make_config(w.n_users, w.n_items, w.k, w.lr, w.lambda_, seed=mf.SEED, mode=capi.MODE_HOGWILD)
for rep in range(2):
    t0 = time.time()
    capi.check(capi.lib.mfsgd_factorize(capi.ptr(hu), capi.ptr(hi), capi.ptr(hr), len(hr), C.byref(cfg), epochs, capi.ptr(P), capi.ptr(Q)))
    print("mfsgd_factorize %d epochs: %.3f s" % (epochs, time.time() - t0), flush=True)
