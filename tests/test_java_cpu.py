"""The Java sources compile (JDK 22+ only; skipped in this image, which has no JDK -- see java/README.md)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def javac_release():
    exe = shutil.which("javac")
    if not exe:
        return 0
    out = subprocess.run([exe, "-version"], capture_output=True, text=True)
    m = re.search(r"javac (\d+)", out.stdout + out.stderr)
    return int(m.group(1)) if m else 0


def test_java_host_and_stand_in_compile(tmp_path):
    if javac_release() < 22:
        pytest.skip("no JDK 22+ on PATH")
    subprocess.check_call(["javac", "--release", "22", "-d", str(tmp_path), os.path.join(ROOT, "java", "MatrixFactorizationSGDGpu.java"),
                           os.path.join(ROOT, "baseline", "java", "MatrixFactorizationSGD.java")])
    assert (tmp_path / "MatrixFactorizationSGDGpu.class").exists() and (tmp_path / "MatrixFactorizationSGD.class").exists()


# ---- without a JDK: the FFM binding's hand-written layout and descriptors against the ctypes binding, which the C harness and
# ---- tests/test_abi_cpu.py pin to include/mfsgd.h. A drifted offset in the Java file would corrupt the config silently.
JAVA = os.path.join(ROOT, "java", "MatrixFactorizationSGDGpu.java")
_JSIZE = {"JAVA_INT": 4, "JAVA_FLOAT": 4, "JAVA_LONG": 8, "JAVA_BYTE": 1, "JAVA_DOUBLE": 8, "ADDRESS": 8}


def _java_config_layout():
    src = open(JAVA).read()
    body = src[src.index("StructLayout CONFIG = MemoryLayout.structLayout("):]
    body = body[:body.index(");  /*") if ");  /*" in body else body.index(");")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for m in re.finditer(r"MemoryLayout\.sequenceLayout\((\d+),\s*(JAVA_\w+)\)\.withName\(\"(\w+)\"\)|(JAVA_\w+)\.withName\(\"(\w+)\"\)", body):
        if m.group(3):
            fields.append((m.group(3), _JSIZE[m.group(2)] * int(m.group(1)), _JSIZE[m.group(2)]))
        else:
            fields.append((m.group(5), _JSIZE[m.group(4)], _JSIZE[m.group(4)]))
    return fields


def test_java_config_layout_matches_the_c_struct():
    import matrixfactorizationsgd.java_b200 as mf
    Config = mf.capi.Config
    fields = _java_config_layout()
    assert [n for n, _, _ in fields] == [n.rstrip("_") for n, _ in Config._fields_]      # same members, same order
    off = 0
    for name, size, align in fields:
        assert off % align == 0, "the Java layout would need padding before %s; FFM structLayout does not insert any" % name
        c = getattr(Config, name if hasattr(Config, name) else name + "_")
        assert (c.offset, c.size) == (off, size), (name, c.offset, c.size, off, size)
        off += size
    import ctypes
    assert off == ctypes.sizeof(Config) == 248


def test_java_hard_coded_offsets_and_descriptors_match_the_binding():
    import ctypes as C
    import matrixfactorizationsgd.java_b200 as mf
    capi = mf.capi
    src = open(JAVA).read()
    # cfg.set(JAVA_X, <offset>, <javaName>) lines of the config writer
    want = {"nUsers": "n_users", "nItems": "n_items", "k": "k", "lr": "lr", "lambda": "lambda_", "seed": "seed", "mode": "mode", "nGpus": "n_gpus"}
    seen = 0
    for m in re.finditer(r"cfg\.set\((JAVA_\w+),\s*(\d+),\s*(\w+)\)", src):
        if m.group(3) in want:
            f = getattr(capi.Config, want[m.group(3)])
            assert (f.offset, f.size) == (int(m.group(2)), _JSIZE[m.group(1)]), m.group(0)
            seen += 1
    assert seen >= 6
    # the extension rows' config members and the epoch-stats entry factorizeEarlyStop reads its validation curve from
    consts = dict((m.group(1), int(m.group(2))) for m in re.finditer(r"\b(OFF_\w+|EPOCH_STATS_BYTES)\s*=\s*(\d+)", src))
    for jname, cname in (("OFF_MODEL", "model"), ("OFF_LR_DECAY", "lr_decay"), ("OFF_ES_PATIENCE", "early_stop_patience"),
                         ("OFF_ES_MIN_DELTA", "early_stop_min_delta"), ("OFF_P_STORAGE", "p_storage")):
        assert consts[jname] == getattr(capi.Config, cname).offset, jname
    assert consts["EPOCH_STATS_BYTES"] == C.sizeof(capi.EpochStats)
    assert consts["OFF_STATS_HELDOUT_RMSE"] == capi.EpochStats.heldout_rmse.offset
    assert (capi.MODEL_GLOBAL_MEAN, capi.MODEL_BIASES, capi.STORAGE_F16) == (1, 2, 1) and "MODEL_GLOBAL_MEAN = 1, MODEL_BIASES = 2" in src
    # the host mirrors the stand-in's public entry points for every built row of SURVEY.md section 8
    for method in ("factorize", "factorizeMixed", "factorizeModel", "factorizeEarlyStop", "rmse", "rmseModel", "readRatings"):
        assert re.search(r"public static [\w.]+ %s\(" % method, src), method
    # struct mfsgd_ratings as readRatings reads it
    for name, offset in (("users", 0), ("items", 8), ("ratings", 16), ("n", 24), ("n_users", 32), ("n_items", 36), ("user_ids", 40), ("item_ids", 48)):
        assert getattr(capi.Ratings, name).offset == offset
    assert C.sizeof(capi.Ratings) == 64
    # every downcall: the symbol exists in the header's binding, with the same number and width class of arguments
    def klass(t):
        if t is None:
            return "void"
        if t in (C.c_int32, C.c_int, C.c_uint32):
            return "JAVA_INT"
        if t in (C.c_int64, C.c_uint64):
            return "JAVA_LONG"
        if t is C.c_float:
            return "JAVA_FLOAT"
        if t is C.c_double:
            return "JAVA_DOUBLE"
        return "ADDRESS"          # pointers, c_void_p, c_char_p
    n = 0
    for m in re.finditer(r"down\(\"(mfsgd_\w+)\",\s*FunctionDescriptor\.(of|ofVoid)\(([^;]*?)\)\);", src, flags=re.S):
        name, kind, args = m.group(1), m.group(2), [a.strip() for a in m.group(3).split(",") if a.strip()]
        assert name in capi.SIGNATURES, name
        res, argtypes = capi.SIGNATURES[name]
        if kind == "of":
            assert klass(res) == args[0], (name, "return")
            args = args[1:]
        else:
            assert res is None, name
        assert [klass(t) for t in argtypes] == args, (name, [klass(t) for t in argtypes], args)
        n += 1
    assert n >= 18
