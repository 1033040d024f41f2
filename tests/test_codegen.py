"""Code-generation guard (CPU, no GPU needed): the two benchmark kernels sit on the 255-register cliff and ptxas'
allocation has flipped into spilling for reasons as remote as the size of the kernel parameter struct (141 -> 200 ms
per 4096 Level-2 bootstraps).  This pins the resource usage of the default instantiations in the built objects."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "mosfhet_b200", "build")


def _usage(obj):
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe) and not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    if not os.path.exists(obj):
        pytest.skip(f"{obj} not built in this checkout (python mosfhet_b200/build.py)")
    out = subprocess.run([exe, "-res-usage", obj], capture_output=True, text=True).stdout
    res = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+)", out):
        res[m.group(1)] = (int(m.group(2)), int(m.group(3)))
    return res


def test_default_bootstrap_kernels_do_not_spill():
    res = _usage(os.path.join(BUILD, "blind_rotate_k1.o"))
    # <LOGM, L, LB, MINB, PKALL, PF, DIRECT>: Level 1 (N=1024, l=3, batches of 2+1 levels) and Level 2 (N=2048, l=4)
    level1 = [v for k, v in res.items() if "Li9ELi3ELi2ELi1ELb1ELi0ELb0E" in k]
    level2 = [v for k, v in res.items() if "Li10ELi4ELi2ELi1ELb0ELi0ELb0E" in k]
    assert level1 and level2, "default instantiations missing from blind_rotate_k1.o"
    assert level1[0][1] == 0, f"Level-1 kernel spills: REG/STACK = {level1[0]}"
    assert level2[0][1] <= 16, f"Level-2 kernel spills more than its 16-byte baseline: REG/STACK = {level2[0]}"


def test_keyswitch_kernel_register_budget():
    res = _usage(os.path.join(BUILD, "keyswitch.o"))
    ks = [v for k, v in res.items() if "keyswitch_warp_kernelILi10ELb0E" in k]
    assert ks, "TLWE key-switch instantiation (640-word rows) missing"
    assert ks[0][0] <= 72, f"28 warps per SM need <= 72 registers, got {ks[0]}"


def test_keyswitch_kernel_keeps_three_row_loads_in_flight():
    """At the 72-register cap the allocator decides how many 128-bit row loads rotate: with two the Level-1 key switch
    takes 4.2 ms per 4096 ciphertexts instead of 3.5 (measured, profiles/r2m_ks_sweep.log)."""
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    obj = os.path.join(BUILD, "keyswitch.o")
    if not os.path.exists(exe) or not os.path.exists(obj):
        pytest.skip("cuobjdump or keyswitch.o not available")
    sass = subprocess.run([exe, "-sass", "-fun", "_ZN2mb21keyswitch_warp_kernelILi10ELb0EEEvNS_11TableKsArgsE", obj],
                          capture_output=True, text=True).stdout
    dests = set(re.findall(r"LDG\.E\.128\.CONSTANT (R\d+),", sass))
    assert len(dests) >= 3, f"row loads rotate over {sorted(dests)} only"


def test_k1q_kernels_fit_four_warps_per_scheduler():
    """The T = M/4 kernels exist to run 16 warps per SM: 128 registers at most, and only a few spilled words."""
    res = _usage(os.path.join(BUILD, "blind_rotate_k1q.o"))
    # <LOGM, L, LB, PKALL, VAR>
    level1 = [v for k, v in res.items() if "k1q_kernelILi9ELi3ELi2ELb1ELi7E" in k]
    level2 = [v for k, v in res.items() if "k1q_kernelILi10ELi4ELi2ELb0ELi7E" in k]
    assert level1 and level2, "benchmark instantiations missing from blind_rotate_k1q.o"
    for name, (reg, stack) in (("level 1", level1[0]), ("level 2", level2[0])):
        assert reg <= 128, f"{name}: {reg} registers"
    assert level1[0][1] <= 64, f"Level-1 k1q kernel spills: REG/STACK = {level1[0]}"
