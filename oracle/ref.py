"""TEST INFRASTRUCTURE ONLY -- ctypes loader for the UNMODIFIED reference CPU library built by
``oracle/build_ref.sh`` into ``oracle/_ref/`` (never imported by the mosfhet_b200 package).

Picks the ISA variant the current host can run (``/proc/cpuinfo``) and declares the prototypes
of the reference functions the tests drive (``/root/reference/include/mosfhet.h``).
"""
from __future__ import annotations

import ctypes as C
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from mosfhet_b200 import abi  # noqa: E402  (struct mirrors only; no compute)

REF_DIR = os.path.join(HERE, "_ref")

# host FFT slot-order id (include/mosfhet_b200.h) of each variant
VARIANT_LAYOUT = {"avx512": 1, "fma": 1, "portable": 2}


def cpu_flags() -> set:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def runnable_variants() -> list:
    fl = cpu_flags()
    out = []
    if {"avx512f", "avx512dq", "avx512vl", "avx512bw", "avx512cd", "vaes", "aes", "rdrand", "pclmulqdq"} <= fl:
        out.append("avx512")
    if {"avx2", "fma", "aes", "rdrand", "pclmulqdq", "bmi2"} <= fl:
        out.append("fma")
    out.append("portable")
    return [v for v in out if os.path.exists(os.path.join(REF_DIR, f"libmosfhet_{v}.so"))]


def best_variant() -> str | None:
    env = os.environ.get("MOSFHET_REF_VARIANT")
    vs = runnable_variants()
    if env:
        return env if env in vs else None
    return vs[0] if vs else None


_SIGS = {
    # name: (restype, argtypes)
    "tlwe_new_binary_key": (abi.TLWE_Key, [C.c_int, C.c_double]),
    "trlwe_new_binary_key": (abi.TRLWE_Key, [C.c_int, C.c_int, C.c_double]),
    "trlwe_extract_tlwe_key": (None, [abi.TLWE_Key, abi.TRLWE_Key]),
    "trgsw_new_key": (abi.TRGSW_Key, [abi.TRLWE_Key, C.c_int, C.c_int]),
    "new_bootstrap_key": (abi.Bootstrap_Key, [abi.TRGSW_Key, abi.TLWE_Key, C.c_int]),
    "tlwe_new_KS_key": (abi.TLWE_KS_Key, [abi.TLWE_Key, abi.TLWE_Key, C.c_int, C.c_int]),
    "tlwe_new_sample": (abi.TLWE, [C.c_uint64, abi.TLWE_Key]),
    "tlwe_alloc_sample": (abi.TLWE, [C.c_int]),
    "tlwe_phase": (C.c_uint64, [abi.TLWE, abi.TLWE_Key]),
    "trlwe_alloc_new_sample": (abi.TRLWE, [C.c_int, C.c_int]),
    "trlwe_alloc_new_DFT_sample": (abi.TRLWE_DFT, [C.c_int, C.c_int]),
    "trlwe_new_sample": (abi.TRLWE, [abi.TorusPolynomial, abi.TRLWE_Key]),
    "trlwe_new_noiseless_trivial_sample": (abi.TRLWE, [abi.TorusPolynomial, C.c_int, C.c_int]),
    "trlwe_phase": (None, [abi.TorusPolynomial, abi.TRLWE, abi.TRLWE_Key]),
    "trlwe_torus_packing": (None, [abi.TRLWE, C.POINTER(C.c_uint64), C.c_int]),
    "trlwe_torus_packing_many_LUT": (None, [abi.TRLWE, C.POINTER(C.c_uint64), C.c_int, C.c_int]),
    "trgsw_new_sample": (abi.TRGSW, [C.c_uint64, abi.TRGSW_Key]),
    "trgsw_new_exp_sample": (abi.TRGSW, [C.c_int, abi.TRGSW_Key]),
    "trgsw_new_monomial_sample": (abi.TRGSW, [C.c_int64, C.c_int, abi.TRGSW_Key]),
    "trgsw_alloc_new_DFT_sample": (abi.TRGSW_DFT, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "trgsw_to_DFT": (None, [abi.TRGSW_DFT, abi.TRGSW]),
    "polynomial_new_torus_polynomial": (abi.TorusPolynomial, [C.c_int]),
    "polynomial_new_DFT_polynomial": (abi.DFTPolynomial, [C.c_int]),
    "polynomial_decompose_i": (None, [abi.TorusPolynomial, abi.TorusPolynomial, C.c_int, C.c_int, C.c_int]),
    "polynomial_torus_to_DFT": (None, [abi.DFTPolynomial, abi.TorusPolynomial]),
    "polynomial_DFT_to_torus": (None, [abi.TorusPolynomial, abi.DFTPolynomial]),
    "torus_polynomial_mul_by_xai": (None, [abi.TorusPolynomial, abi.TorusPolynomial, C.c_int]),
    "torus_polynomial_mul_by_xai_minus_1": (None, [abi.TorusPolynomial, abi.TorusPolynomial, C.c_int]),
    "init_fft": (None, [C.c_int]),
    "generate_random_bytes": (None, [C.c_uint64, C.c_void_p]),
    # the hot path itself
    "functional_bootstrap": (None, [abi.TLWE, abi.TRLWE, abi.TLWE, abi.Bootstrap_Key, C.c_int]),
    "functional_bootstrap_wo_extract": (None, [abi.TRLWE, abi.TRLWE, abi.TLWE, abi.Bootstrap_Key, C.c_int]),
    "programmable_bootstrap": (None, [abi.TLWE, abi.TRLWE, abi.TLWE, abi.Bootstrap_Key, C.c_int, C.c_int, C.c_int]),
    "multivalue_bootstrap_CLOT21": (None, [C.POINTER(abi.TLWE), abi.TRLWE, abi.TLWE, abi.Bootstrap_Key, C.c_int, C.c_int]),
    "blind_rotate": (None, [abi.TRLWE, C.POINTER(C.c_uint64), C.POINTER(abi.TRGSW_DFT), C.c_int]),
    "trgsw_mul_trlwe_DFT": (None, [abi.TRLWE_DFT, abi.TRLWE, abi.TRGSW_DFT]),
    "trlwe_from_DFT": (None, [abi.TRLWE, abi.TRLWE_DFT]),
    "trlwe_extract_tlwe": (None, [abi.TLWE, abi.TRLWE, C.c_int]),
    "tlwe_keyswitch": (None, [abi.TLWE, abi.TLWE, abi.TLWE_KS_Key]),
    "free_tlwe": (None, [abi.TLWE]),
    "free_trlwe": (None, [C.c_void_p]),
    "free_trgsw": (None, [C.c_void_p]),
    "free_bootstrap_key": (None, [abi.Bootstrap_Key]),
    "free_tlwe_ks_key": (None, [abi.TLWE_KS_Key]),
}


class RefLib:
    """The reference CPU library (one ISA variant), loaded RTLD_LOCAL."""

    def __init__(self, variant: str | None = None):
        variant = variant or best_variant()
        if variant is None:
            raise RuntimeError("no runnable reference build under oracle/_ref (run oracle/build_ref.sh)")
        self.variant = variant
        self.layout = VARIANT_LAYOUT[variant]
        self.path = os.path.join(REF_DIR, f"libmosfhet_{variant}.so")
        self.lib = C.CDLL(self.path, mode=os.RTLD_LOCAL)
        for name, (res, args) in _SIGS.items():
            fn = getattr(self.lib, name)
            fn.restype = res
            fn.argtypes = args

    def __getattr__(self, name):
        return getattr(self.lib, name)


_cached = {}


def load(variant: str | None = None) -> RefLib:
    key = variant or best_variant()
    if key not in _cached:
        _cached[key] = RefLib(key)
    return _cached[key]


def available() -> bool:
    return best_variant() is not None
