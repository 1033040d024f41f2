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


def _build_dropin(tmp_path, name):
    from oracle import ref as reflib
    variant = reflib.best_variant()
    if variant is None:
        pytest.skip("oracle/_ref not built")
    from mosfhet_b200 import _lib
    exe = str(tmp_path / name)
    # the reference first (key generation, encryption), then the CUDA library for its mb200_* entry points; at run time
    # LD_PRELOAD puts the CUDA library in front so that the shared names resolve to it
    subprocess.check_call(["gcc", "-O1", "-w", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "dropin", name + ".c"), "-o", exe,
                           "-L", reflib.REF_DIR, f"-l:libmosfhet_{variant}.so", _lib.LIB_PATH,
                           f"-Wl,-rpath,{reflib.REF_DIR}", f"-Wl,-rpath,{os.path.dirname(_lib.LIB_PATH)}", "-ldl", "-lm"])
    return exe


@pytest.mark.gpu
def test_multi_gpu_mode_in_the_library(tmp_path):
    """mb200_init_multi: the batched drop-in call sharded over the visible devices equals the one-device result word for
    word.  With ONE visible device the mode degenerates to a single worker (still exercised: worker thread, shard loop)."""
    from mosfhet_b200 import _lib, api
    ndev = max(1, api.device_count())
    exe = _build_dropin(tmp_path, "dropin_multi")
    env = dict(os.environ, LD_PRELOAD=_lib.LIB_PATH)
    # small key, batch spread over min(ndev, 8) devices
    out = subprocess.run([exe, str(min(ndev, 8)), "640", "64", "1024", "3", "6", "7", "2"], env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "DROPIN MULTI OK" in out.stdout, out.stdout + out.stderr
