"""CPU-side tests of the boundary: libmfsgd.so loads, exports every symbol include/mfsgd.h declares,
validates arguments before touching the GPU, and fails loudly (no fallback) when no GPU exists.
No compute calls are made here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import matrixfactorizationsgd.java_b200 as mf
from matrixfactorizationsgd.java_b200 import _capi as capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "mfsgd.h")).read()
    return sorted(set(re.findall(r"MFSGD_API[^;(]*?\b(mfsgd_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree():
    syms = header_symbols()
    assert len(syms) >= 29
    assert syms == sorted(capi.SIGNATURES)


def test_library_exports_every_declared_symbol():
    out = subprocess.check_output(["nm", "-D", "--defined-only", capi.LIB_PATH], text=True)
    exported = set(re.findall(r" T (mfsgd_[a-z_0-9]+)", out))
    assert set(header_symbols()) <= exported
    raw = C.CDLL(capi.LIB_PATH)
    for name in header_symbols():
        assert getattr(raw, name) is not None


def test_struct_sizes_match_c_layout():
    # offsets the Java FFM layout (java/MatrixFactorizationSGDGpu.java) also hard-codes
    assert C.sizeof(capi.Config) == 248 and capi.Config.model.offset == 216 and capi.Config.lr_decay.offset == 224
    assert capi.Config.early_stop_min_delta.offset == 232
    assert capi.Config.seed.offset == 24 and capi.Config.nccl_id.offset == 68 and capi.Config.ctas_per_sm.offset == 196
    assert C.sizeof(capi.EpochStats) == 72
    assert C.sizeof(capi.SynthParams) == 48 and capi.SynthParams.planted_amplitude.offset == 40
    assert C.sizeof(capi.LayoutInfo) == 64
    assert C.sizeof(capi.Ratings) == 64 and capi.Ratings.n.offset == 24 and capi.Ratings.user_ids.offset == 40


def test_c_harness_dlopen_dlsym():
    exe = os.path.join(ROOT, "tests", "c", "abi_harness")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", ROOT, "harness"])
    out = subprocess.check_output([exe, capi.LIB_PATH], text=True)
    assert "abi_harness: OK" in out


@pytest.mark.parametrize("kw,needle", [
    (dict(k=6), "k=6"), (dict(k=0), "k=0"), (dict(k=516), "k=516"), (dict(n_users=0), "n_users"),
    (dict(lr=0.0), "lr"), (dict(lambda_=-1.0), "lambda"), (dict(mode=7), "mode"),
    (dict(n_gpus=2), "n_gpus == 1"), (dict(mode=capi.MODE_DSGD, n_gpus=0), "n_gpus"),
    (dict(scatter=9), "scatter"), (dict(mode=capi.MODE_DSGD, n_gpus=4, world_size=2), "world_size"),
    (dict(mode=capi.MODE_DETERMINISTIC, stripes_per_gpu=2), "DETERMINISTIC"), (dict(device=-1), "device"),
    (dict(rounds=-1), "rounds"), (dict(hot_chunk=-5), "hot_chunk"),
])
def test_invalid_config_rejected_before_gpu(kw, needle):
    base = dict(n_users=10, n_items=10, k=8, lr=0.1, lambda_=0.1)
    base.update(kw)
    with pytest.raises(mf.MfsgdError) as ei:
        mf.Engine(mf.make_config(**base))
    assert ei.value.code == capi.E_INVALID_ARG
    assert needle in str(ei.value)


def test_null_arguments():
    assert capi.lib.mfsgd_create(None, None) == capi.E_INVALID_ARG
    assert capi.lib.mfsgd_train(None, 1, None) == capi.E_INVALID_ARG
    assert capi.lib.mfsgd_init_factors(None) == capi.E_INVALID_ARG
    assert capi.lib.mfsgd_get_factors(None, None, None) == capi.E_INVALID_ARG
    assert capi.lib.mfsgd_rmse(None, None, None, None, 0, None) == capi.E_INVALID_ARG
    assert capi.lib.mfsgd_nccl_unique_id(None) == capi.E_INVALID_ARG
    capi.lib.mfsgd_destroy(None)   # no-op, must not crash
    assert b"null" in capi.lib.mfsgd_last_error()


def test_no_gpu_fails_loudly_no_fallback():
    if mf.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(mf.MfsgdError) as ei:
        mf.Engine(mf.make_config(10, 10, 8, 0.1, 0.1))
    assert ei.value.code == capi.E_CUDA and "no CPU fallback" in str(ei.value)
    with pytest.raises(mf.MfsgdError):
        mf.MatrixFactorizationSGD.factorize(np.zeros(1, np.int32), np.zeros(1, np.int32), np.ones(1, np.float32),
                                            1, 1, 8, 0.1, 0.1, 1, 1)


def test_host_mirror_argument_checks_match_stand_in():
    z = np.zeros(2, np.int32)
    with pytest.raises(ValueError):
        mf.MatrixFactorizationSGD.factorize(z, z[:1], np.ones(2, np.float32), 3, 3, 8, 0.1, 0.1, 1, 1)
    with pytest.raises(ValueError):
        mf.MatrixFactorizationSGD.factorize(z, z, np.ones(2, np.float32), 3, 3, 0, 0.1, 0.1, 1, 1)
    with pytest.raises(ValueError):
        mf.MatrixFactorizationSGD.factorize(z, z, np.ones(2, np.float32), 3, 3, 8, 0.1, 0.1, -1, 1)
    # the extension entry points (stand-in :305, :350, :439) reject what the stand-in rejects, before any GPU work
    M, r = mf.MatrixFactorizationSGD, np.ones(2, np.float32)
    with pytest.raises(ValueError, match="bad shape"):
        M.factorizeMixed(z, z, r, 3, 3, 6, 0.1, 0.1, 1, 1)                      # binary16 rows: k % 4 == 0 (:443)
    with pytest.raises(ValueError, match="differ in length"):
        M.factorizeModel(z, z, r[:1], 3, 3, 8, 0.1, 0.1, 1, 1, True, True)
    for decay, patience, delta in ((1.5, 1, 0.0), (0.0, 1, 0.0), (0.9, -1, 0.0), (0.9, 1, 1.0), (0.9, 1, -0.1)):
        with pytest.raises(ValueError, match="bad schedule"):                   # :356-357
            M.factorizeEarlyStop(z, z, r, z, z, r, 3, 3, 8, 0.1, 0.1, 1, 1, True, True, decay, patience, delta)
    with pytest.raises(ValueError, match="differ in length"):
        M.factorizeEarlyStop(z, z, r, z, z, r[:1], 3, 3, 8, 0.1, 0.1, 1, 1, True, True, 0.9, 1, 0.0)
    with pytest.raises(ValueError):
        M.rmse(np.zeros((3, 8), np.float32), np.zeros((3, 4), np.float32), 8, z, z, r)


def test_product_does_not_touch_the_oracle():
    """The product path must never import, link or load anything under oracle/."""
    pkg = os.path.join(ROOT, "matrixfactorizationsgd.java_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "pyoracle" not in text and "liboracle" not in text and "orc_" not in text, fn
    deps = subprocess.check_output(["ldd", capi.LIB_PATH], text=True)
    assert "oracle" not in deps


def test_workloads_match_baseline_json():
    import json
    cfgs = json.load(open(os.path.join(ROOT, "BASELINE.json")))["configs"]
    w = mf.WORKLOADS
    assert "943 users" in cfgs[0] and (w["ml100k"].n_users, w["ml100k"].n_items, w["ml100k"].k) == (943, 1682, 32)
    assert "20M ratings, k=128" in cfgs[1] and w["ml20m"].n_ratings == 20_000_000 and w["ml20m"].k == 128
    assert "480K users" in cfgs[2] and (w["netflix"].n_users, w["netflix"].n_items) == (480_000, 17_800)
    assert "700M ratings" in cfgs[3] and w["yahoo"].n_ratings == 700_000_000
    assert "k=64" in cfgs[4] and w["powerlaw"].k == 64 and w["powerlaw"].n_ratings == 2_000_000_000
    assert mf.bytes_per_update(128) == 2060 and mf.bytes_per_update(64) == 1036 and mf.bytes_per_update(32) == 524


def test_cpp_host_rejects_bad_arguments_before_any_gpu_work():
    """host/factorize_demo (C++ twin of the Java host): the stand-in's argument errors (bad shape :114, k % 4 :443, bad schedule :356)
    are raised by the host itself; without a GPU the first real call then fails loudly (exit 3) -- there is no CPU fallback."""
    exe = os.path.join(ROOT, "host", "factorize_demo")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", ROOT, "host"])
    out = subprocess.run([exe, capi.LIB_PATH], capture_output=True, text=True, timeout=300)
    assert "was not rejected" not in out.stderr
    if mf.device_count() > 0:
        assert out.returncode == 0, (out.stdout, out.stderr)
    else:
        assert out.returncode == 3 and "GPU path unavailable" in out.stderr and "no CPU fallback" in out.stderr
