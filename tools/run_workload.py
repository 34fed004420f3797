"""One BASELINE.json workload end to end on the GPU(s) of this process: synthetic ratings generated on the device,
w.epochs epochs with the held-out RMSE after every epoch, timings -> one JSON line (and gpurun_out/ when present).
usage: python tools/run_workload.py <ml100k|ml20m|netflix|yahoo|powerlaw> [virtual ring members G] [epochs]
The per-epoch curve is compared with tests/golden/oracle_rmse_<name>.json (the sequential CPU oracle) when that exists."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import matrixfactorizationsgd.java_b200 as mf
capi = mf.capi

name = sys.argv[1]
G = int(sys.argv[2]) if len(sys.argv) > 2 else 1
w = mf.WORKLOADS[name]
epochs = int(sys.argv[3]) if len(sys.argv) > 3 else w.epochs
kw = dict(seed=mf.SEED, flags=capi.FLAG_TIME_KERNELS)
for env, key, conv in (("RW_PAT", "p_atomic_threshold", float), ("RW_SCATTER", "scatter", int), ("RW_BOOST", "merge_boost", float), ("RW_ROUNDS", "rounds", int), ("RW_HOT_SHARE", "hot_share", float), ("RW_STRIPES", "stripes_per_gpu", int),
                       ("RW_HOT_CHUNK", "hot_chunk", int), ("RW_SHARDS", "shards_per_gpu", int)):
    if os.environ.get(env):
        kw[key] = conv(os.environ[env])
extra = int(os.environ.get("RW_FLAGS", "0"))      # e.g. 16 = MFSGD_FLAG_SPLIT_SHARDS: the item sub-shards on stream lanes, as a real ring runs them
kw["flags"] |= extra
if G > 1:
    kw.update(mode=capi.MODE_DSGD, n_gpus=G, flags=capi.FLAG_TIME_KERNELS | capi.FLAG_VIRTUAL_RING | extra)
else:
    kw.update(mode=capi.MODE_HOGWILD)
cfg = mf.make_config(w.n_users, w.n_items, w.k, w.lr, w.lambda_, **kw)
out = {"workload": w.name, "members": G, "k": w.k, "epochs": epochs}
with mf.Engine(cfg) as eng:
    t0 = time.time()
    nt, nh = eng.generate_synthetic(mf.synth_params_of(w))
    out["setup_s"] = time.time() - t0
    info = eng.layout_info()
    eng.init_factors()
    eng.set_eval_every_epoch(True)
    st = eng.train(epochs)
out.update(n_train=int(nt), n_heldout=int(nh), stripes_per_gpu=int(info.stripes_per_gpu), shards_per_gpu=int(info.shards_per_gpu),
           rounds=int(info.rounds), hot_items=int(info.n_hot_items), heavy_users=int(info.n_heavy_users), run_length=int(info.run_length),
           epoch_ms=[s.epoch_ms for s in st], heldout_rmse_per_epoch=[s.heldout_rmse for s in st],
           cold_ms=[s.cold_ms for s in st], hot_ms=[s.hot_ms for s in st], launches=[s.update_launches for s in st])
ms = float(np.median(out["epoch_ms"]))
out["gupdates_per_s"] = nt / ms / 1e6
out["roofline_frac_measured_peak"] = nt / (ms * 1e-3) * mf.bytes_per_update(w.k) / 6552.6e9
ref = os.path.join(ROOT, "tests", "golden", "oracle_rmse_%s.json" % name)
if os.path.exists(ref):
    fixture = json.load(open(ref))
    want = fixture.get("heldout_rmse_per_epoch", [])
    out["oracle_rmse_per_epoch"] = want
    out["constant_predictor_rmse"] = fixture.get("constant_predictor_rmse")
    if "dsgd%d" % G in fixture:      # the sequential rule in this ring's own block order (tools/oracle_reference_rmse.py --dsgd)
        out["oracle_dsgd_order_rmse_per_epoch"] = fixture["dsgd%d" % G]["heldout_rmse_per_epoch"]
    m = min(len(want), epochs)
    out["rel_diff_at_equal_epochs"] = [out["heldout_rmse_per_epoch"][e] / want[e] - 1 for e in range(m)]
line = json.dumps(out)
print(line, flush=True)
d = os.path.join(ROOT, "gpurun_out")
if os.path.isdir(d):
    open(os.path.join(d, "workload_%s_g%d%s.json" % (name, G, os.environ.get("RW_TAG", ""))), "w").write(line + "\n")
