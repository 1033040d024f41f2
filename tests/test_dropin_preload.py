"""Drop-in proof: a C program written against the reference API, linked to the unmodified reference
library, runs with libmosfhet_b200.so LD_PRELOADed; the interposed hot-path calls execute on the
GPU and match a private (dlmopen) copy of the reference on the same keys and inputs."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_ld_preload_dropin(tmp_path):
    from mosfhet_b200 import _lib
    from oracle import ref as reflib
    variant = reflib.best_variant()
    if variant is None:
        pytest.skip("oracle/_ref not built")
    ref_so = os.path.join(reflib.REF_DIR, f"libmosfhet_{variant}.so")
    exe = str(tmp_path / "dropin_main")
    subprocess.check_call(["gcc", "-O1", "-w", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "dropin", "dropin_main.c"), "-o", exe,
                           "-L", reflib.REF_DIR, f"-l:libmosfhet_{variant}.so", f"-Wl,-rpath,{reflib.REF_DIR}", "-ldl", "-lm"])
    env = dict(os.environ, LD_PRELOAD=_lib.LIB_PATH)
    out = subprocess.run([exe, ref_so], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "DROPIN OK" in out.stdout, out.stdout + out.stderr
