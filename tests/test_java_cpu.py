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
