#!/usr/bin/env python
"""bench.py -- SGD rating updates/sec on the Netflix-shaped workload (BASELINE.json configs[2]; the
config the metric and the 40 %-of-roofline target are quoted on), 1/2/4/8 B200.

  python bench.py --gpus N --steps K --warmup W            # N>1: launched by torch.distributed.run
  python bench.py --impl reference ...                      # the CPU path (oracle port), host cores

A step = one epoch = one pass of the hot path (reshuffle + update kernels [+ Q rotation]) over the
training records of the workload. Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "sgd_rating_updates_per_sec"
E2E_REPEATS = 2     # the end-to-end call is made twice; e2e.value is the FIRST call (what a caller's first call costs in a process
                    # whose CUDA context exists), the second is reported beside it
UNIT = "updates/s"


def load_workloads():
    """workloads.py by path: pure data, loads no native library (the reference arm must not map libmfsgd.so)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("mfsgd_workloads", os.path.join(ROOT, "matrixfactorizationsgd.java_b200", "workloads.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(workload, timeout_s=240):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel (sgd_update_runs_kernel), measured now,
    on this box, by a side run of the same workload under ncu (tools/profile_target.py, the 21st run-kernel launch: a
    steady-state visit of the second epoch). None when ncu is unavailable or fails -- never a stale constant."""
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None, "ncu not found"
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum", "--clock-control", "none", "-k",
           "regex:sgd_update_runs_kernel", "-s", "20", "-c", "1", "--csv", sys.executable, os.path.join(ROOT, "tools", "profile_target.py"),
           workload, "0", "0", "2"]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s)
    except Exception as e:      # noqa: BLE001
        return None, "ncu side run failed: %r" % (e,)
    vals = {}
    for line in out.stdout.splitlines():
        cells = [c.strip().strip('"') for c in line.split('","')]
        for name in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"):
            if name in cells:
                try:
                    unit, val = cells[cells.index(name) + 1], float(cells[cells.index(name) + 2].replace(",", ""))
                    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6}.get(unit, 1.0)
                    vals[name] = val * scale
                except Exception:      # noqa: BLE001
                    pass
    if "dram__bytes_read.sum" not in vals or "dram__bytes_write.sum" not in vals:
        return None, "ncu produced no dram counters (rc %d): %s" % (out.returncode, (out.stderr or out.stdout)[-200:].replace("\n", " "))
    return vals, "ncu side run in this bench process's box: launch 21 of sgd_update_runs_kernel (tools/profile_target.py)"


class stdout_to_stderr:
    """Route fd 1 to fd 2 while native libraries that print banners (NCCL's version line) initialise:
    this script's stdout carries exactly one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for row in self.rows:
            try:
                sm.append(float(row[1])); mx.append(float(row[2]))
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), row[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
def host_training_set(mf, w, device, pinned):
    """Host triplets of the workload's training records (generated on the GPU, held-out tenth removed)."""
    sp = mf.synth_params_of(w)
    us, is_, rs = [], [], []
    step = 25_000_000
    for start in range(0, w.n_ratings, step):
        u, i, r, held = mf.generate_to_host(sp, w.n_users, w.n_items, start, min(step, w.n_ratings - start), device)
        us.append(u[~held]); is_.append(i[~held]); rs.append(r[~held])
    n = sum(len(x) for x in rs)
    if not pinned:
        return np.concatenate(us), np.concatenate(is_), np.concatenate(rs), None
    ptrs, arrs = [], []
    for parts, dt in ((us, np.int32), (is_, np.int32), (rs, np.float32)):
        p = C.c_void_p()
        mf.capi.check(mf.capi.lib.mfsgd_host_alloc(C.byref(p), n * 4))
        a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int32 if dt == np.int32 else C.c_float)), shape=(n,))
        np.concatenate(parts, out=a)
        ptrs.append(p); arrs.append(a)
    return arrs[0], arrs[1], arrs[2], ptrs


def run_ours(args):
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"     # keep NCCL's version banner off stdout: this script prints ONE JSON line
    import matrixfactorizationsgd.java_b200 as mf
    capi = mf.capi
    w = mf.WORKLOADS[args.workload]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        # torch.distributed carries only bootstrap bytes and timing scalars (CPU tensors): gloo. The Q-shard
        # rotation runs over NCCL inside libmfsgd.so.
        dist.init_process_group("gloo", rank=rank, world_size=world)
    elif args.gpus > 1:
        raise SystemExit("N>1 runs one process per GPU: launch with python -m torch.distributed.run (see docstring)")

    def barrier():
        import torch
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(local_rank)

    def allmax(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def allsum(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t)
        return t.item()

    from matrixfactorizationsgd.java_b200 import ring
    flags = capi.FLAG_TIME_KERNELS
    common = dict(n_users=w.n_users, n_items=w.n_items, k=w.k, lr=w.lr, lambda_=w.lambda_, seed=mf.SEED,
                  stripes_per_gpu=args.stripes, shards_per_gpu=args.shards, scatter=args.scatter, flags=flags,
                  ctas_per_sm=args.ctas_per_sm, rounds=args.rounds, hot_share=args.hot_share, hot_chunk=args.hot_chunk,
                  p_storage=capi.STORAGE_F16 if args.p_storage == "f16" else capi.STORAGE_F32)
    if world > 1:
        with stdout_to_stderr():
            eng = ring.create_rank_engine(dist, rank, world, local_rank, **common)
    else:
        eng = mf.Engine(mf.make_config(mode=capi.MODE_HOGWILD, device=local_rank, **common))
    sp = mf.synth_params_of(w)
    t0 = time.time()
    n_train_local, n_held_local = eng.generate_synthetic(sp)
    setup_s = time.time() - t0
    info = eng.layout_info()
    eng.init_factors()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()           # nvidia-smi needs a moment to produce its first sample: start before the warm-up
    with stdout_to_stderr():
        if args.warmup > 0:
            eng.train(args.warmup)
    barrier()
    sampler.rows.clear()          # keep only samples taken during the timed region
    wall0 = time.time()
    stats = eng.train(args.steps)
    barrier()
    wall = time.time() - wall0
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = sum(s.epoch_ms for s in stats)                       # CUDA events on the engine's stream
    total_ms = allmax(dev_ms)
    updates = allsum(float(sum(s.updates for s in stats)))
    value = updates / (total_ms * 1e-3)
    kernel_ms = sum(s.update_kernel_ms for s in stats)
    n_launch = sum(s.update_launches for s in stats)
    launches_all = allsum(float(sum(s.total_launches for s in stats)))
    shuffle_ms = sum(s.shuffle_ms for s in stats) / max(1, len(stats))
    # held-out RMSE after warmup+steps epochs (evidence that the timed work is real training)
    _, sse, cnt = eng.rmse_heldout()
    heldout_rmse = ring.reduce_rmse(dist, sse, cnt) if dist is not None else float(np.sqrt(sse / max(cnt, 1)))
    if world == 1:
        eng.close()

    # Roofline of the update phase on rank 0: the run-kernel (sgd_update_runs_kernel) and cold (sgd_update_hogwild_kernel)
    # launches of an epoch run concurrently on two streams, so they are timed together:
    # one CUDA-event span per sub-epoch on the launching stream, fork to join. Algorithmic bytes = 12 + 16k per
    # update x the updates in the span.
    peak, peak_src = peaks()
    half_p = args.p_storage == "f16"
    bpu = mf.bytes_per_update(w.k) - (4 * w.k if half_p else 0)      # binary16 P rows: 2k B read + 2k B written instead of 4k + 4k
    rank_updates = float(sum(s.updates for s in stats))
    achieved = rank_updates * bpu / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src,
                "kernel": "sgd_update_runs_kernel (+ sgd_update_hogwild_kernel for rarely rated items; concurrent streams, timed as one span per sub-epoch)",
                "bytes_per_update": bpu, "updates_per_step": rank_updates / args.steps,
                "updates_per_launch": rank_updates / max(n_launch, 1),
                "algorithmic_bytes_per_launch": rank_updates * bpu / max(n_launch, 1),
                "avg_launch_ms": kernel_ms / max(n_launch, 1),
                "algorithmic_bytes_per_step": rank_updates * bpu / args.steps,
                "update_phase_ms_per_step": kernel_ms / args.steps, "update_launches_per_step": n_launch / args.steps,
                "kernel_share_of_step": kernel_ms / max(dev_ms, 1e-9), "frac_of_nominal_8TBs": achieved / 8000.0,
                "note": "frac > 1 is expected: the P sub-stripe and Q stay L2-resident (stratified blocks) and q_i rows "
                        "live in registers for a whole run, so DRAM traffic (see traffic) is a small fraction of the algorithmic "
                        "bytes; the binding resource is L2 sector throughput (see l2_bound)"}
    if world == 1 and rank == 0 and not args.no_traffic:
        vals, src = measured_traffic(args.workload)
        roofline["traffic_source"] = src
        if vals is not None:
            roofline["traffic"] = vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"]
            roofline["traffic_detail"] = {"unit": "bytes per launch", "dram_read": vals["dram__bytes_read.sum"],
                                          "dram_write": vals["dram__bytes_write.sum"],
                                          "launch_us_under_ncu": vals.get("gpu__time_duration.sum", 0.0) / 1e3,
                                          "per_step_estimate": (vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"]) * n_launch / args.steps}
    if rank == 0 and w.k == 128 and not args.no_ceilings:
        # The ceiling of the run kernel's access pattern, measured NOW in this process on this box (mfsgd_measure_ceilings):
        # random 512-B row gather + scatter inside an L2-resident buffer the size of one P sub-stripe, no arithmetic.
        ceil = mf.measure_ceilings(local_rank, 61.0)
        # what an update of the run kernel moves through L2: p_u read + written (2 x 4k B) and its record (12 B, whole sectors shared by a tile)
        l2_bytes = (4 if half_p else 8) * w.k + 12
        l2_achieved = rank_updates * l2_bytes / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0
        roofline["l2_bound"] = {"achieved": l2_achieved, "peak": ceil["row_gather_scatter_gbs"], "unit": "GB/s",
                                "frac": l2_achieved / ceil["row_gather_scatter_gbs"], "l2_bytes_per_update": l2_bytes,
                                "ceilings_measured_in_process": ceil,
                                "peak_source": "measured in this process after the timed region: mfsgd_measure_ceilings (csrc/kernels_diag.cu), "
                                               "random 512-B row gather + scatter in a 61 MB L2-resident buffer, no arithmetic"}

    # e2e: the reference-facing call with HOST buffers: H2D + bucketing + init + K epochs + D2H of P, Q
    e2e = None
    if not args.no_e2e:
        if args.trace_e2e:
            os.environ["MFSGD_TRACE"] = "1"
        hu, hi, hr, pins = host_training_set(mf, w, local_rank, pinned=True)
        n_host = len(hr)
        # pinned output buffers too (a Java caller would hand in off-heap segments from mfsgd_host_alloc)
        out_ptrs = []
        def pinned_f32(rows, cols):
            p = C.c_void_p()
            capi.check(capi.lib.mfsgd_host_alloc(C.byref(p), rows * cols * 4))
            out_ptrs.append(p)
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(rows, cols))
        P = pinned_f32(w.n_users, w.k)
        Q = pinned_f32(w.n_items, w.k)
        e2e_cfg = dict(common, flags=0)
        if world > 1:
            # Ring bootstrap (ncclCommInitRank and NCCL's lazy peer connections, 1-2 s for 8 ranks) is set-up, like process
            # start: the e2e call runs on the live ring handle used above. Timed: the handle-level calls a resident caller
            # makes -- load host triplets (H2D + bucketing), init, K epochs, read P and Q back.
            eng2 = eng
            e2e_all = []
            for _rep in range(E2E_REPEATS):
                barrier()
                t0 = time.time()
                lo_, hi_ = n_host * rank // world, n_host * (rank + 1) // world      # every rank uploads only its 1/N slice
                capi.check(capi.lib.mfsgd_load_ratings_sharded(eng2._h, capi.ptr(hu[lo_:hi_]), capi.ptr(hi[lo_:hi_]), capi.ptr(hr[lo_:hi_]),
                                                               hi_ - lo_))
                eng2.init_factors()
                eng2.train(args.steps, want_stats=False)
                capi.check(capi.lib.mfsgd_get_factors(eng2._h, capi.ptr(P), capi.ptr(Q)))
                barrier()
                e2e_all.append(allmax(time.time() - t0))
            e2e_s = e2e_all[0]
            eng2.close()
            e2e_call = ("mfsgd_load_ratings_sharded (each rank uploads 1/N of the triplets, records exchanged by stripe owner over NCCL) "
                        "+ init_factors + train + get_factors on a live ring handle (pinned host buffers)")
        else:
            cfg = mf.make_config(mode=capi.MODE_HOGWILD, device=local_rank, **e2e_cfg)
            e2e_all = []
            for _rep in range(E2E_REPEATS):
                barrier()
                t0 = time.time()
                capi.check(capi.lib.mfsgd_factorize(capi.ptr(hu), capi.ptr(hi), capi.ptr(hr), n_host, C.byref(cfg), args.steps,
                                                    capi.ptr(P), capi.ptr(Q)))
                barrier()
                e2e_all.append(allmax(time.time() - t0))
            e2e_s = e2e_all[0]
            e2e_call = "mfsgd_factorize(host triplets -> host P,Q), pinned host buffers"
        e2e_check = float(np.abs(P[:1000]).sum() + np.abs(Q[:1000]).sum())   # the result was really read back
        for p in pins + out_ptrs:
            capi.lib.mfsgd_host_free(p)
        e2e = {"value": float(n_host) * args.steps / e2e_s, "unit": UNIT,
               "h2d_bytes_per_step": 12.0 * n_host / args.steps,
               "d2h_bytes_per_step": 4.0 * w.k * (w.n_users + w.n_items) / args.steps,
               "seconds": e2e_s, "seconds_each_call": e2e_all,
               "value_second_call": float(n_host) * args.steps / e2e_all[-1],
               "call": e2e_call + " -- the FIRST of %d complete calls (the second is in seconds_each_call / value_second_call)" % E2E_REPEATS,
               "result_checksum": e2e_check,
               "epochs": args.steps}

    cpu = cpu_ml100k = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(w, sample=args.cpu_sample, threads=1)
        cpu_ml100k = cpu_baseline_ml100k()

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "f32" if not half_p else "f32 arithmetic, P rows stored as binary16 (stochastic rounding), Q binary32", "data": "synthetic",
               "config": {"workload": "%s: %d users x %d items, %d ratings (%d train), k=%d, lr=%g, lambda=%g" % (
                              w.name, w.n_users, w.n_items, w.n_ratings, int(info.n_train_total), w.k, w.lr, w.lambda_),
                          "parallelism": "hogwild-1gpu" if world == 1 else "dsgd-ring%d" % world,
                          "stripes_per_gpu": int(info.stripes_per_gpu), "shards_per_gpu": int(info.shards_per_gpu),
                          "rounds": int(info.rounds), "hot_items": int(info.n_hot_items), "run_length": int(info.run_length),
                          "rotation": ("ring window: copy-engine writes into the neighbour's memory over NVLink + sequence flags, pipelined over item sub-shards"
                                       if os.environ.get("MFSGD_RING_TRANSPORT", "window") != "nccl" else "ncclSend/ncclRecv, pipelined over item sub-shards") if world > 1 else "none",
                          "scatter": "store" if args.scatter == 0 else "atomic",
                          "arith": "fma (FFMA2 arrangement of the update rule, a few ulp per update from the stand-in's unfused rule; "
                                   "MFSGD_FLAG_EXACT_ARITH selects the unfused one)",
                          "merge": "runs of one item in one launch: weight min(1, 1.25 / runs)",
                          "l2": "inputs larger than L2: %.2f GB of records + %.0f MB of factors streamed per step" % (
                              12e-9 * info.n_train_total, 4e-6 * w.k * (w.n_users + w.n_items)),
                          "setup_seconds_excluded": setup_s},
               "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches_all), "clocks": clocks,
               "heldout_rmse": heldout_rmse, "epochs_trained": args.warmup + args.steps,
               "shuffle_ms_per_step": shuffle_ms, "wall_ms_per_step": wall * 1e3 / args.steps,
               "breakdown_ms_per_step_rank0": {"cold_kernels": sum(s.cold_ms for s in stats) / args.steps,
                                               "hot_kernels": sum(s.hot_ms for s in stats) / args.steps,
                                               "q_rotation": sum(s.exchange_ms for s in stats) / args.steps,
                                               "note": "cold and hot overlap (two streams); spans start at the sub-epoch fork"}}
        # the sequential oracle's held-out RMSE at the same epoch count, when the committed curve reaches that far
        try:
            fx = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_rmse_%s.json" % args.workload)))
            ep = args.warmup + args.steps
            if len(fx.get("heldout_rmse_per_epoch", [])) >= ep:
                want = fx["heldout_rmse_per_epoch"][ep - 1]
                out["rmse_vs_oracle"] = {"epochs": ep, "gpu": heldout_rmse, "oracle_sequential": want, "rel": heldout_rmse / want - 1.0,
                                         "constant_predictor_rmse": fx.get("constant_predictor_rmse")}
                key = "dsgd%d" % world
                if key in fx and len(fx[key]["heldout_rmse_per_epoch"]) >= ep:
                    out["rmse_vs_oracle"]["oracle_dsgd_order"] = fx[key]["heldout_rmse_per_epoch"][ep - 1]
                # the tests' bar (tests/test_gpu_parity.py assert_rmse_parity / assert_ring_rmse_parity): within 0.5 % of the sequential
                # oracle, two-sided; a ring may land anywhere between the shuffled and the DSGD-ordered sequential execution, +-0.5 %
                refs = [want] + ([out["rmse_vs_oracle"]["oracle_dsgd_order"]] if "oracle_dsgd_order" in out["rmse_vs_oracle"] else [])
                out["rmse_vs_oracle"]["within_half_percent"] = bool(min(refs) * 0.995 <= heldout_rmse <= max(refs) * 1.005)
        except Exception:      # noqa: BLE001  (no fixture for this workload)
            pass
        if cpu is not None:
            out["cpu_baseline"] = cpu
        if cpu_ml100k is not None:
            out["cpu_baseline_ml100k"] = cpu_ml100k
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------
def cpu_baseline(w, sample, threads):
    """The reference's CPU path (the C++ oracle port of the Java stand-in: no JDK in this image) on a
    bounded sample: the first `sample` records of the same synthetic set, full-size P and Q."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as orc
    seed = load_workloads().SEED
    n = min(sample, w.n_ratings)
    u, i, r, held = orc.generate(seed, 0, n, w.n_users, w.n_items, w.log2_alpha_user, w.c_user, w.log2_alpha_item, w.c_item,
                                 amplitude=w.amplitude, noise_scale=w.noise_scale)
    u, i, r = u[~held].copy(), i[~held].copy(), r[~held].copy()
    P = orc.init_factors(w.n_users, w.k, seed, 0)
    Q = orc.init_factors(w.n_items, w.k, seed, 1)
    if threads == 1:
        t0 = time.time()
        orc.train(u, i, r, P, Q, w.lr, w.lambda_, 0, 1, seed, shuffled=False)
        secs = time.time() - t0
    else:
        secs = orc.train_hogwild(u, i, r, P, Q, w.lr, w.lambda_, 0, 1, seed, threads, shuffled=False)
    return {"value": len(r) / secs, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "1 epoch over the first %d records (%d train) of %s, full-size P/Q, k=%d" % (n, len(r), w.name, w.k),
            "host_cores_available": orc.hardware_threads(), "seconds": secs}


def cpu_baseline_ml100k():
    """configs[0] (ML-100K-shaped, k=32, 20 epochs) through the reference's CPU path on THIS box's host cores, sequential and
    thread-parallel, in this run (north_star): oracle/oracle_cli is the C++ port of the stand-in's main()."""
    exe = os.path.join(ROOT, "oracle", "oracle_cli")
    if not os.path.exists(exe):
        return None
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as orc
    cores = orc.hardware_threads()
    rows = []
    for mode in ["seq"] + ["threads=%d" % t for t in sorted({2, 4, 8, cores}) if t <= cores]:
        try:
            line = subprocess.run([exe, "--mode", mode], capture_output=True, text=True, timeout=120).stdout.strip().splitlines()[-1]
            d = json.loads(line)
            rows.append({"mode": d["mode"], "threads": d["threads"], "updates_per_sec": d["updates_per_sec"],
                         "heldout_rmse": d["heldout_rmse"], "seconds": d["seconds"]})
        except Exception as e:      # noqa: BLE001
            rows.append({"mode": mode, "error": repr(e)})
    return {"workload": "ml100k-shaped: 943 users x 1682 items, 100000 ratings, k=32, 20 epochs (BASELINE.json configs[0])",
            "kind": "port", "host_cores": cores, "runs": rows,
            "note": "whole runs incl. the per-epoch shuffle sort; the set is so small that sort and thread start-up dominate the threaded runs"}


def run_reference(args):
    """--impl reference: the CPU implementation of the path (the stand-in's factorizeThreaded, C++ oracle port), all host
    threads, the SAME workload as the product arm: every training record, the per-epoch shuffle included in the step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as orc
    wl = load_workloads()          # pure data: this process never maps libmfsgd.so
    w, seed = wl.WORKLOADS[args.workload], wl.SEED
    threads = orc.hardware_threads()
    n = w.n_ratings if args.ref_sample <= 0 else min(args.ref_sample, w.n_ratings)
    us, is_, rs = [], [], []
    step = 25_000_000
    for start in range(0, n, step):
        u, i, r, held = orc.generate(seed, start, min(step, n - start), w.n_users, w.n_items, w.log2_alpha_user, w.c_user,
                                     w.log2_alpha_item, w.c_item, amplitude=w.amplitude, noise_scale=w.noise_scale)
        us.append(u[~held]); is_.append(i[~held]); rs.append(r[~held])
    u, i, r = np.concatenate(us), np.concatenate(is_), np.concatenate(rs)
    del us, is_, rs
    P = orc.init_factors(w.n_users, w.k, seed, 0)
    Q = orc.init_factors(w.n_items, w.k, seed, 1)
    # Bounded run: every step covers the whole workload unless this host is too slow for --warmup + --steps of them to end
    # within --ref-budget-s; then the remaining steps cover a prefix of the records (said in `sample`, and
    # same_workload_as_product_arm turns false). value = records of the timed steps / their wall time either way.
    n_train = len(r)
    wall = loops = spent = 0.0
    done = 0
    per_step = []
    for s in range(args.warmup + args.steps):
        t0 = time.time()
        lp = orc.train_hogwild(u, i, r, P, Q, w.lr, w.lambda_, s, s + 1, seed, threads, shuffled=True)
        dt = time.time() - t0
        spent += dt
        if s >= args.warmup:
            wall += dt
            loops += lp
            done += len(r)
            per_step.append(len(r))
        left = args.warmup + args.steps - (s + 1)
        room = max(args.ref_budget_s - spent, 1.0)
        if args.ref_budget_s > 0 and left > 0 and dt * left > room:
            keep = max(1_000_000, int(len(r) * room / (dt * left)))
            if keep < len(r):
                u, i, r = u[:keep].copy(), i[:keep].copy(), r[:keep].copy()
    value = done / wall
    full = n == w.n_ratings and all(c == n_train for c in per_step)
    if all(c == n_train for c in per_step):
        covered = "ALL %d records (%d train)" % (n, n_train) if n == w.n_ratings else "the first %d records (%d train)" % (n, n_train)
    else:
        covered = "a prefix of the %d training records (%d..%d per timed step: the %g s budget of the whole run required the cut)" % (
            n_train, min(per_step), max(per_step), args.ref_budget_s)
    sample = "each step = 1 Hogwild epoch (per-epoch shuffle + update loops), %d host threads, over %s of %s, full-size P/Q" % (
        threads, covered, w.name)
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall * 1e3 / args.steps, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "%s: %d users x %d items, %d ratings (%d train), k=%d, lr=%g, lambda=%g" % (
               w.name, w.n_users, w.n_items, n, n_train, w.k, w.lr, w.lambda_), "parallelism": "cpu-hogwild-%dthreads" % threads,
               "same_workload_as_product_arm": full, "shuffled_every_epoch": True},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                            "update_loops_only": done / loops},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "note": "reference = C++ oracle port of the Java stand-in's factorizeThreaded (no JDK in the image; /root/reference has no "
                   "source); the per-epoch order is the stand-in's (same permutation), produced with all host threads"}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="netflix")
    ap.add_argument("--stripes", type=int, default=0)
    ap.add_argument("--shards", type=int, default=0)
    ap.add_argument("--scatter", type=int, default=0)
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--rounds", type=int, default=0)
    ap.add_argument("--hot-share", type=float, default=0.0)
    ap.add_argument("--hot-chunk", type=int, default=0, help="longest run of the run kernel (0 = planned from the launch size)")
    ap.add_argument("--p-storage", default="f32", choices=["f32", "f16"],
                    help="f16: rows of P kept as binary16 with stochastic rounding (SURVEY.md 8f.3) -- a separate line, never the headline")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=12_000_000)
    ap.add_argument("--ref-sample", type=int, default=0, help="reference arm: records per step, 0 = the whole workload")
    ap.add_argument("--ref-budget-s", type=float, default=300.0,
                    help="reference arm: wall-clock budget of the whole run; later steps shrink to a prefix of the records if needed (0 = never)")
    ap.add_argument("--no-traffic", action="store_true", help="skip the ncu side run that measures roofline.traffic")
    ap.add_argument("--no-ceilings", action="store_true")
    ap.add_argument("--trace-e2e", action="store_true", help="MFSGD_TRACE=1 during the end-to-end calls (phase timings on stderr)")
    args = ap.parse_args()
    if args.steps < 1:
        raise SystemExit("--steps must be >= 1")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
