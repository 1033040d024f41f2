"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU, exports every
symbol include/mosfhet_b200.h declares, and the handle layouts agree between the header, the ctypes
mirror and (when mounted) the reference's own mosfhet.h.  No compute entry point is called."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mosfhet_b200.h")


@pytest.fixture(scope="module")
def lib():
    from mosfhet_b200 import build, _lib
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    return _lib.load()


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?[A-Za-z_][A-Za-z0-9_ \*]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", src, flags=re.M)
    return sorted(set(n for n in names if n not in ("defined",)))


def test_header_functions_exported(lib):
    from mosfhet_b200 import _lib
    names = declared_functions()
    assert len(names) >= 55
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    # and the ctypes prototype table covers the header exactly
    assert sorted(_lib.PROTOTYPES) == names


def test_dropin_names_present(lib):
    for n in ("functional_bootstrap", "functional_bootstrap_wo_extract", "programmable_bootstrap", "blind_rotate",
              "trgsw_mul_trlwe_DFT", "trlwe_from_DFT", "trlwe_extract_tlwe", "tlwe_keyswitch"):
        assert hasattr(lib, n)


def test_no_gpu_probe_does_not_abort(lib):
    assert lib.mb200_device_count() >= 0
    assert lib.mb200_version().startswith(b"mosfhet_b200")


LAYOUT_PROBE = r"""
#include <stdio.h>
#include <stddef.h>
%s
#define P(T, f) printf(#T "." #f " %%zu\n", offsetof(struct T, f))
#define S(T) printf(#T " %%zu\n", sizeof(struct T))
int main(void) {
  S(_TorusPolynomial); P(_TorusPolynomial, coeffs); P(_TorusPolynomial, N);
  S(_DFT_Polynomial); P(_DFT_Polynomial, coeffs); P(_DFT_Polynomial, N);
  S(_TLWE); P(_TLWE, a); P(_TLWE, b); P(_TLWE, n);
  S(_TLWE_KS_Key); P(_TLWE_KS_Key, s); P(_TLWE_KS_Key, base_bit); P(_TLWE_KS_Key, t); P(_TLWE_KS_Key, n);
  S(_TRLWE); P(_TRLWE, a); P(_TRLWE, b); P(_TRLWE, k);
  S(_TRLWE_DFT); P(_TRLWE_DFT, a); P(_TRLWE_DFT, b); P(_TRLWE_DFT, k);
  S(_TRGSW_DFT); P(_TRGSW_DFT, samples); P(_TRGSW_DFT, l); P(_TRGSW_DFT, Bg_bit);
  S(_Bootstrap_Key); P(_Bootstrap_Key, s); P(_Bootstrap_Key, su); P(_Bootstrap_Key, n); P(_Bootstrap_Key, k);
  P(_Bootstrap_Key, N); P(_Bootstrap_Key, Bg_bit); P(_Bootstrap_Key, l); P(_Bootstrap_Key, unfolding);
  S(_TRGSW); P(_TRGSW, samples); P(_TRGSW, l); P(_TRGSW, Bg_bit);
  S(_Generic_KS_Key); P(_Generic_KS_Key, s); P(_Generic_KS_Key, base_bit); P(_Generic_KS_Key, t); P(_Generic_KS_Key, n);
  P(_Generic_KS_Key, include_b);
  S(_TRLWE_KS_Key); P(_TRLWE_KS_Key, s); P(_TRLWE_KS_Key, base_bit); P(_TRLWE_KS_Key, t); P(_TRLWE_KS_Key, k);
  return 0;
}
"""


def probe_layout(include_line, extra_flags=()):
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "p.c")
        open(src, "w").write(LAYOUT_PROBE % include_line)
        exe = os.path.join(td, "p")
        subprocess.check_call(["gcc", "-w", "-o", exe, src, *extra_flags])
        return subprocess.check_output([exe], text=True)


def ctypes_layout():
    from mosfhet_b200 import abi
    pairs = [("_TorusPolynomial", abi.TorusPolynomialS), ("_DFT_Polynomial", abi.DFTPolynomialS), ("_TLWE", abi.TLWES),
             ("_TLWE_KS_Key", abi.TLWEKSKeyS), ("_TRLWE", abi.TRLWES), ("_TRLWE_DFT", abi.TRLWEDFTS),
             ("_TRGSW_DFT", abi.TRGSWDFTS), ("_Bootstrap_Key", abi.BootstrapKeyS), ("_TRGSW", abi.TRGSWS),
             ("_Generic_KS_Key", abi.GenericKSKeyS), ("_TRLWE_KS_Key", abi.TRLWEKSKeyS)]
    lines = []
    for name, cls in pairs:
        lines.append(f"{name} {C.sizeof(cls)}")
        for f, _ in cls._fields_:
            lines.append(f"{name}.{f} {getattr(cls, f).offset}")
    return "\n".join(lines) + "\n"


def test_handle_layouts_match_header():
    mine = probe_layout(f'#include "{HEADER}"')
    assert mine == ctypes_layout()


@pytest.mark.skipif(not os.path.exists("/root/reference/include/mosfhet.h"), reason="reference tree not mounted")
def test_handle_layouts_match_reference_header():
    ref = probe_layout('#include "/root/reference/include/mosfhet.h"', ["-DPORTABLE_BUILD"])
    mine = probe_layout(f'#include "{HEADER}"')
    assert ref == mine


def test_slot_exponent_tables(lib):
    """mb200_host_slot_exponents agrees with the oracle's tables (both pinned to the live reference
    by tests/test_oracle_golden.py::test_slot_order_*)."""
    from mosfhet_b200 import api
    from oracle import oracle as O
    for layout in (1, 2, 3):
        for N in (16, 256, 2048):
            assert np.array_equal(api.host_slot_exponents(layout, N), O.slot_exponents(layout, N))


def test_host_struct_builders_roundtrip():
    from mosfhet_b200 import abi
    rng = np.random.default_rng(0)
    t = rng.integers(0, 2**64, size=9, dtype=np.uint64)
    assert np.array_equal(abi.tlwe_to_flat(abi.HostTLWE(t).handle), t)
    p = rng.integers(0, 2**64, size=(3, 32), dtype=np.uint64)
    assert np.array_equal(abi.trlwe_to_flat(abi.HostTRLWE(p).handle), p)
    bsk = rng.standard_normal((3, 4, 2, 32))
    assert np.array_equal(abi.bootstrap_key_to_flat(abi.HostBootstrapKey(bsk, 1, 2, 8).handle), bsk)
    ksk = rng.integers(0, 2**64, size=(8, 2, 3, 5), dtype=np.uint64)
    assert np.array_equal(abi.ks_key_to_flat(abi.HostKSKey(ksk, 2).handle), ksk)
    assert abi.HostTRLWE(p).polys.ctypes.data % 64 == 0
    # the handle types of the circuit bootstrap / unfolded keys
    g = rng.integers(0, 2**64, size=(5, 2, 3, 2, 16), dtype=np.uint64)
    hg = abi.HostGenericKSKey(g, 2, 1)
    assert np.array_equal(abi.generic_ks_key_to_flat(hg.handle), g) and hg.struct.n == 4 and hg.struct.include_b == 1
    r = rng.standard_normal((1, 3, 2, 16))
    assert np.array_equal(abi.trlwe_ks_key_to_flat(abi.HostTRLWEKSKey(r, 2).handle), r)
    pair = abi.HostTRLWEKSKeyPair(rng.standard_normal((2, 3, 2, 16)), 2)
    assert pair.handle[0].contents.t == 3 and pair.handle[1].contents.k == 1
    tg = rng.integers(0, 2**64, size=(4, 2, 16), dtype=np.uint64)
    assert np.array_equal(abi.trgsw_to_flat(abi.HostTRGSW(tg, 2, 8).handle, 1), tg)
    su = rng.integers(0, 2**64, size=(8, 4, 2, 16), dtype=np.uint64)       # n = 4, unfolding = 2 -> 2 groups x 4
    hu = abi.HostUnfoldedBootstrapKey(su, 4, 2, 1, 2, 8)
    assert hu.struct.unfolding == 2 and np.array_equal(abi.trgsw_to_flat(hu.struct.su[5], 1), su[5])
