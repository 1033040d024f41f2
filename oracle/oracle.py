"""TEST INFRASTRUCTURE ONLY -- numpy front-end to ``oracle/liboracle.so`` (oracle.c), the plain-C
CPU restatement of the reference hot path.  Never imported by the mosfhet_b200 package; only
tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference`` legs use it.

All arrays are numpy, in the flat layouts of ``include/mosfhet_b200.h`` section 3; Fourier-domain
data is in the oracle's NATURAL slot order (slot s <-> root exponent 1+4s) unless a function
says otherwise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(HERE, "liboracle.so")

_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_int = C.c_int


def build() -> None:
    src = os.path.join(HERE, "oracle.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", _SO, src, "-lm"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO, mode=os.RTLD_LOCAL)
        sig = {
            "oracle_torus2int": (C.c_uint64, [C.c_uint64, _int]),
            "oracle_double2torus": (C.c_uint64, [C.c_double]),
            "oracle_decompose_i": (None, [_i64p, _u64p, _int, _int, _int, _int]),
            "oracle_mul_by_xai": (None, [_u64p, _u64p, _int, _int]),
            "oracle_mul_by_xai_minus_1": (None, [_u64p, _u64p, _int, _int]),
            "oracle_int_to_dft": (None, [_f64p, _i64p, _int]),
            "oracle_torus_to_dft": (None, [_f64p, _u64p, _int]),
            "oracle_f64_to_torus": (C.c_uint64, [C.c_double, _int]),
            "oracle_dft_to_torus": (None, [_u64p, _f64p, _int, _int]),
            "oracle_slot_exponents": (None, [_int, _int, _i32p]),
            "oracle_permute_from_host": (None, [_f64p, _f64p, _int, _i32p]),
            "oracle_permute_to_host": (None, [_f64p, _f64p, _int, _i32p]),
            "oracle_trgsw_mul_trlwe_dft": (None, [_f64p, _u64p, _f64p, _int, _int, _int, _int]),
            "oracle_trlwe_from_dft": (None, [_u64p, _f64p, _int, _int, _int]),
            "oracle_blind_rotate": (None, [_u64p, _u64p, _f64p, _int, _int, _int, _int, _int, _int]),
            "oracle_functional_bootstrap_wo_extract": (None, [_u64p, _u64p, _u64p, _f64p] + [_int] * 7),
            "oracle_extract_tlwe": (None, [_u64p, _u64p, _int, _int, _int]),
            "oracle_extract_tlwe_addto": (None, [_u64p, _u64p, _int, _int, _int]),
            "oracle_extract_tlwe_subto": (None, [_u64p, _u64p, _int, _int, _int]),
            "oracle_mv_extract_tlwe": (None, [_u64p, _u64p, _int, _int, _int]),
            "oracle_mv_extract_tlwe_scaling": (None, [_u64p, _u64p, _int, _int, _int]),
            "oracle_mv_extract_tlwe_scaling_acc": (None, [_u64p, _u64p, _int, _int, _int, _int]),
            "oracle_functional_bootstrap": (None, [_u64p, _u64p, _u64p, _f64p] + [_int] * 7),
            "oracle_programmable_preprocess": (None, [_u64p, _u64p, _int, _int, _int, _int]),
            "oracle_programmable_bootstrap": (None, [_u64p, _u64p, _u64p, _f64p] + [_int] * 9),
            "oracle_multivalue_bootstrap_CLOT21": (None, [_u64p, _u64p, _u64p, _f64p] + [_int] * 8),
            "oracle_multivalue_phase1": (None, [_u64p, _u64p, _f64p] + [_int] * 7),
            "oracle_multivalue_phase2": (None, [_u64p, _i32p, _u64p, _int, _int, _int, _int]),
            "oracle_table_keyswitch_trlwe": (None, [_u64p, _u64p, _u64p] + [_int] * 6),
            "oracle_circuit_bootstrap_2": (None, [_u64p, _u64p, _f64p, _u64p, _u64p] + [_int] * 9),
            "oracle_functional_bootstrap_trgsw_phase1": (None, [_u64p, _f64p, _u64p, _f64p] + [_int] * 9),
            "oracle_functional_bootstrap_trgsw_phase2": (None, [_u64p, _f64p, _u64p] + [_int] * 5),
            "oracle_unfold_group": (None, [_u64p, _u64p, _u64p] + [_int] * 5),
            "oracle_blind_rotate_unfolded": (None, [_u64p, _u64p, _u64p] + [_int] * 7),
            "oracle_functional_bootstrap_unfolded_wo_extract": (None, [_u64p, _u64p, _u64p, _u64p] + [_int] * 8),
            "oracle_trlwe_keyswitch": (None, [_u64p, _u64p, _f64p] + [_int] * 6),
            "oracle_trlwe_priv_keyswitch_2": (None, [_u64p, _u64p, _f64p] + [_int] * 4),
            "oracle_circuit_bootstrap": (None, [_u64p, _u64p, _f64p, _u64p, _u64p] + [_int] * 10),
            "oracle_circuit_bootstrap_3": (None, [_u64p, _u64p, _f64p, _f64p, _u64p] + [_int] * 10),
            "oracle_tlwe_keyswitch": (None, [_u64p, _u64p, _u64p, _int, _int, _int, _int]),
            "oracle_tlwe_phase": (C.c_uint64, [_u64p, _u64p, _int]),
            "oracle_trlwe_phase": (None, [_u64p, _u64p, _u64p, _int, _int]),
            "oracle_trgsw_mul_trlwe_exact": (None, [_u64p, _u64p, _u64p, _int, _int, _int, _int]),
            "oracle_blind_rotate_exact": (None, [_u64p, _u64p, _u64p, _int, _int, _int, _int, _int]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


# ---- thin numpy wrappers ----------------------------------------------------------------------
def torus2int(x: int, log_scale: int) -> int:
    return int(lib().oracle_torus2int(int(x) & (2**64 - 1), log_scale))


def double2torus(x: float) -> int:
    return int(lib().oracle_double2torus(float(x)))


def decompose_i(poly, Bg_bit, l, i):
    poly = _c(poly, np.uint64)
    out = np.empty(poly.shape[0], np.int64)
    lib().oracle_decompose_i(out, poly, poly.shape[0], Bg_bit, l, i)
    return out


def mul_by_xai(poly, a):
    poly = _c(poly, np.uint64)
    out = np.empty_like(poly)
    lib().oracle_mul_by_xai(out, poly, poly.shape[0], int(a))
    return out


def mul_by_xai_minus_1(poly, a):
    poly = _c(poly, np.uint64)
    out = np.empty_like(poly)
    lib().oracle_mul_by_xai_minus_1(out, poly, poly.shape[0], int(a))
    return out


def torus_to_dft(poly):
    poly = _c(poly, np.uint64)
    out = np.empty(poly.shape[0], np.float64)
    lib().oracle_torus_to_dft(out, poly, poly.shape[0])
    return out


def dft_to_torus(dft, mode=0):
    dft = _c(dft, np.float64)
    out = np.empty(dft.shape[0], np.uint64)
    lib().oracle_dft_to_torus(out, dft, dft.shape[0], mode)
    return out


def slot_exponents(layout, N):
    e = np.empty(N // 2, np.int32)
    lib().oracle_slot_exponents(layout, N, e)
    return e


def permute_from_host(host_dft, layout_or_exponents):
    """host-order DFT polynomial(s) [..., N] -> natural order."""
    host_dft = _c(host_dft, np.float64)
    N = host_dft.shape[-1]
    e = slot_exponents(layout_or_exponents, N) if np.isscalar(layout_or_exponents) else _c(layout_or_exponents, np.int32)
    flat = host_dft.reshape(-1, N)
    out = np.empty_like(flat)
    for r in range(flat.shape[0]):
        lib().oracle_permute_from_host(out[r], flat[r], N, e)
    return out.reshape(host_dft.shape)


def permute_to_host(nat_dft, layout_or_exponents):
    nat_dft = _c(nat_dft, np.float64)
    N = nat_dft.shape[-1]
    e = slot_exponents(layout_or_exponents, N) if np.isscalar(layout_or_exponents) else _c(layout_or_exponents, np.int32)
    flat = nat_dft.reshape(-1, N)
    out = np.empty_like(flat)
    for r in range(flat.shape[0]):
        lib().oracle_permute_to_host(out[r], flat[r], N, e)
    return out.reshape(nat_dft.shape)


def trgsw_mul_trlwe_dft(trlwe, trgsw_dft, l, Bg_bit):
    trlwe = _c(trlwe, np.uint64)
    trgsw_dft = _c(trgsw_dft, np.float64)
    k, N = trlwe.shape[0] - 1, trlwe.shape[1]
    out = np.empty((k + 1, N), np.float64)
    lib().oracle_trgsw_mul_trlwe_dft(out, trlwe, trgsw_dft, N, k, l, Bg_bit)
    return out


def trlwe_from_dft(dft, mode=0):
    dft = _c(dft, np.float64)
    k, N = dft.shape[0] - 1, dft.shape[1]
    out = np.empty((k + 1, N), np.uint64)
    lib().oracle_trlwe_from_dft(out, dft, N, k, mode)
    return out


def blind_rotate(acc, a, bsk, l, Bg_bit, mode=0):
    acc = _c(acc, np.uint64).copy()
    a = _c(a, np.uint64)
    bsk = _c(bsk, np.float64)
    k, N = acc.shape[0] - 1, acc.shape[1]
    lib().oracle_blind_rotate(acc, a, bsk, a.shape[0], N, k, l, Bg_bit, mode)
    return acc


def functional_bootstrap_wo_extract(tv, tlwe_in, bsk, l, Bg_bit, torus_base, mode=0):
    tv = _c(tv, np.uint64)
    tlwe_in = _c(tlwe_in, np.uint64)
    bsk = _c(bsk, np.float64)
    k, N = tv.shape[0] - 1, tv.shape[1]
    n = tlwe_in.shape[0] - 1
    out = np.empty((k + 1, N), np.uint64)
    lib().oracle_functional_bootstrap_wo_extract(out, tv, tlwe_in, bsk, n, N, k, l, Bg_bit, torus_base, mode)
    return out


def extract_tlwe(trlwe, idx=0):
    trlwe = _c(trlwe, np.uint64)
    k, N = trlwe.shape[0] - 1, trlwe.shape[1]
    out = np.empty(k * N + 1, np.uint64)
    lib().oracle_extract_tlwe(out, trlwe, N, k, idx)
    return out


def extract_tlwe_acc(out, trlwe, idx, sign):
    """trlwe_extract_tlwe_addto (sign = +1) / _subto (sign = -1), trlwe.c:554-578; returns the updated copy of ``out``."""
    trlwe = _c(trlwe, np.uint64)
    k, N = trlwe.shape[0] - 1, trlwe.shape[1]
    out = _c(out, np.uint64).copy()
    (lib().oracle_extract_tlwe_addto if sign > 0 else lib().oracle_extract_tlwe_subto)(out, trlwe, N, k, idx)
    return out


def mv_extract_tlwe(trlwe, amount):
    """trlwe_mv_extract_tlwe (trlwe.c:580-589) -> [amount, k*N+1]."""
    trlwe = _c(trlwe, np.uint64)
    k, N = trlwe.shape[0] - 1, trlwe.shape[1]
    out = np.empty((amount, k * N + 1), np.uint64)
    lib().oracle_mv_extract_tlwe(out, trlwe, N, k, amount)
    return out


def mv_extract_tlwe_scaling(trlwe, scale, out=None, sign=0):
    """trlwe_mv_extract_tlwe_scaling (sign = 0, trlwe.c:591-600), _scaling_addto (+1, :602-610), _scaling_subto (-1, :612-620)."""
    trlwe = _c(trlwe, np.uint64)
    k, N = trlwe.shape[0] - 1, trlwe.shape[1]
    if sign == 0:
        res = np.empty(k * N + 1, np.uint64)
        lib().oracle_mv_extract_tlwe_scaling(res, trlwe, N, k, scale)
        return res
    res = _c(out, np.uint64).copy()
    lib().oracle_mv_extract_tlwe_scaling_acc(res, trlwe, N, k, scale, sign)
    return res


def functional_bootstrap(tv, tlwe_in, bsk, l, Bg_bit, torus_base, mode=0):
    tv = _c(tv, np.uint64)
    tlwe_in = _c(tlwe_in, np.uint64)
    bsk = _c(bsk, np.float64)
    k, N = tv.shape[0] - 1, tv.shape[1]
    n = tlwe_in.shape[0] - 1
    out = np.empty(k * N + 1, np.uint64)
    lib().oracle_functional_bootstrap(out, tv, tlwe_in, bsk, n, N, k, l, Bg_bit, torus_base, mode)
    return out


def programmable_preprocess(tlwe_in, N, kappa, theta):
    tlwe_in = _c(tlwe_in, np.uint64)
    out = np.empty_like(tlwe_in)
    lib().oracle_programmable_preprocess(out, tlwe_in, tlwe_in.shape[0] - 1, N, kappa, theta)
    return out


def programmable_bootstrap(tv, tlwe_in, bsk, l, Bg_bit, precision, kappa, theta, mode=0):
    tv = _c(tv, np.uint64)
    tlwe_in = _c(tlwe_in, np.uint64)
    bsk = _c(bsk, np.float64)
    k, N = tv.shape[0] - 1, tv.shape[1]
    n = tlwe_in.shape[0] - 1
    out = np.empty(k * N + 1, np.uint64)
    lib().oracle_programmable_bootstrap(out, tv, tlwe_in, bsk, n, N, k, l, Bg_bit, precision, kappa, theta, mode)
    return out


def multivalue_bootstrap_CLOT21(tv, tlwe_in, bsk, l, Bg_bit, torus_base, n_luts, mode=0):
    tv = _c(tv, np.uint64)
    tlwe_in = _c(tlwe_in, np.uint64)
    bsk = _c(bsk, np.float64)
    k, N = tv.shape[0] - 1, tv.shape[1]
    n = tlwe_in.shape[0] - 1
    out = np.empty((n_luts, k * N + 1), np.uint64)
    lib().oracle_multivalue_bootstrap_CLOT21(out, tv, tlwe_in, bsk, n, N, k, l, Bg_bit, torus_base, n_luts, mode)
    return out


def multivalue_phase1(tlwe_in, bsk, N, k, l, Bg_bit, torus_base, mode=0):
    tlwe_in = _c(tlwe_in, np.uint64)
    bsk = _c(bsk, np.float64)
    out = np.empty((torus_base + 1, k + 1, N), np.uint64)
    lib().oracle_multivalue_phase1(out, tlwe_in, bsk, tlwe_in.shape[0] - 1, N, k, l, Bg_bit, torus_base, mode)
    return out


def multivalue_phase2(lut_ints, rot, torus_base, log_torus_base):
    rot = _c(rot, np.uint64)
    lut = _c(lut_ints, np.int32)
    k, N = rot.shape[1] - 1, rot.shape[2]
    out = np.empty(k * N + 1, np.uint64)
    lib().oracle_multivalue_phase2(out, lut, rot, N, k, torus_base, log_torus_base)
    return out


def table_keyswitch_trlwe(tlwe_in, table, include_b, base_bit):
    """table: [n_in + include_b, t, 2^base_bit-1, k+1, N] (trlwe_packing1_keyswitch / trlwe_priv_keyswitch)."""
    tlwe_in = _c(tlwe_in, np.uint64)
    table = _c(table, np.uint64)
    ne, t, _, kp1, N = table.shape
    out = np.empty((kp1, N), np.uint64)
    lib().oracle_table_keyswitch_trlwe(out, tlwe_in, table, ne - include_b, include_b, N, kp1 - 1, t, base_bit)
    return out


def circuit_bootstrap_2(tlwe_in, bsk, kska, kskb, l, Bg_bit, Bg_out, base_bit, mode=0):
    tlwe_in = _c(tlwe_in, np.uint64)
    bsk = _c(bsk, np.float64)
    kska, kskb = _c(kska, np.uint64), _c(kskb, np.uint64)
    n = tlwe_in.shape[0] - 1
    t, kp1, N = kskb.shape[1], kskb.shape[3], kskb.shape[4]
    out = np.empty((2 * l, kp1, N), np.uint64)
    lib().oracle_circuit_bootstrap_2(out, tlwe_in, bsk, kska, kskb, n, N, kp1 - 1, l, Bg_bit, Bg_out, t, base_bit, mode)
    return out


def trlwe_keyswitch(trlwe_in, ksk, base_bit, mode=0):
    """ksk: natural-order doubles [k_in, t, k_out+1, N] (keyswitch.c:162-193)."""
    trlwe_in = _c(trlwe_in, np.uint64)
    ksk = _c(ksk, np.float64)
    k_in, t, kp1, N = ksk.shape
    assert trlwe_in.shape == (k_in + 1, N)
    out = np.empty((kp1, N), np.uint64)
    lib().oracle_trlwe_keyswitch(out, trlwe_in, ksk, N, k_in, kp1 - 1, t, base_bit, mode)
    return out


def trlwe_priv_keyswitch_2(trlwe_in, ksk2, base_bit, mode=0):
    """ksk2: natural-order doubles [2, t, 2, N], the pair of trlwe_new_priv_KS_key (keyswitch.c:52-63)."""
    trlwe_in = _c(trlwe_in, np.uint64)
    ksk2 = _c(ksk2, np.float64)
    _, t, _, N = ksk2.shape
    out = np.empty((2, N), np.uint64)
    lib().oracle_trlwe_priv_keyswitch_2(out, trlwe_in, ksk2, N, t, base_bit, mode)
    return out


def circuit_bootstrap(tlwe_in, bsk, kska, kskb, l, Bg_bit, l_out, Bg_out, base_bit, mode=0):
    tlwe_in = _c(tlwe_in, np.uint64)
    bsk = _c(bsk, np.float64)
    kska, kskb = _c(kska, np.uint64), _c(kskb, np.uint64)
    n = tlwe_in.shape[0] - 1
    t, kp1, N = kskb.shape[1], kskb.shape[3], kskb.shape[4]
    out = np.empty((2 * l_out, kp1, N), np.uint64)
    lib().oracle_circuit_bootstrap(out, tlwe_in, bsk, kska, kskb, n, N, kp1 - 1, l, Bg_bit, l_out, Bg_out, t, base_bit, mode)
    return out


def circuit_bootstrap_3(tlwe_in, bsk, kska2, kskb, l, Bg_bit, Bg_out, base_bit_a, base_bit_b, mode=0):
    tlwe_in = _c(tlwe_in, np.uint64)
    bsk = _c(bsk, np.float64)
    kska2, kskb = _c(kska2, np.float64), _c(kskb, np.uint64)
    n = tlwe_in.shape[0] - 1
    t_b, N = kskb.shape[1], kskb.shape[4]
    out = np.empty((2 * l, 2, N), np.uint64)
    lib().oracle_circuit_bootstrap_3(out, tlwe_in, bsk, kska2, kskb, n, N, l, Bg_bit, Bg_out, kska2.shape[1], base_bit_a,
                                     t_b, base_bit_b, mode)
    return out


def functional_bootstrap_trgsw_phase1(tlwe_in, bsk, l, Bg_bit, l_out, Bg_out, torus_base, mode=0):
    """-> (torus rows [(k+1)l_out, k+1, N], natural-order DFT rows of the same shape)  (bootstrap.c:286-296)."""
    tlwe_in = _c(tlwe_in, np.uint64)
    bsk = _c(bsk, np.float64)
    n, kp1, N = bsk.shape[0], bsk.shape[2], bsk.shape[3]
    rows = kp1 * l_out
    out_t, out_d = np.empty((rows, kp1, N), np.uint64), np.empty((rows, kp1, N), np.float64)
    lib().oracle_functional_bootstrap_trgsw_phase1(out_t, out_d, tlwe_in, bsk, n, N, kp1 - 1, l, Bg_bit, l_out, Bg_out,
                                                   torus_base, mode)
    return out_t, out_d


def functional_bootstrap_trgsw_phase2(trgsw_dft, tv, l, Bg_bit, mode=0):
    trgsw_dft = _c(trgsw_dft, np.float64)
    tv = _c(tv, np.uint64)
    kp1, N = tv.shape
    out = np.empty((kp1 - 1) * N + 1, np.uint64)
    lib().oracle_functional_bootstrap_trgsw_phase2(out, trgsw_dft, tv, N, kp1 - 1, l, Bg_bit, mode)
    return out


def unfold_group(a, su, group, unfolding, l):
    """su: torus [groups * 2^u, (k+1)l, k+1, N] -> the group's combined TRGSW [(k+1)l, k+1, N] (bootstrap.c:132-140)."""
    a = _c(a, np.uint64)
    su = _c(su, np.uint64)
    kp1, N = su.shape[2], su.shape[3]
    out = np.empty(su.shape[1:], np.uint64)
    lib().oracle_unfold_group(out, a, su, group, unfolding, N, kp1 - 1, l)
    return out


def blind_rotate_unfolded(acc, a, su, size, unfolding, l, Bg_bit, mode=0):
    acc = _c(acc, np.uint64).copy()
    a = _c(a, np.uint64)
    su = _c(su, np.uint64)
    kp1, N = acc.shape
    lib().oracle_blind_rotate_unfolded(acc, a, su, size, unfolding, N, kp1 - 1, l, Bg_bit, mode)
    return acc


def functional_bootstrap_unfolded_wo_extract(tv, tlwe_in, su, unfolding, l, Bg_bit, torus_base, mode=0):
    tv = _c(tv, np.uint64)
    tlwe_in = _c(tlwe_in, np.uint64)
    su = _c(su, np.uint64)
    kp1, N = tv.shape
    out = np.empty((kp1, N), np.uint64)
    lib().oracle_functional_bootstrap_unfolded_wo_extract(out, tv, tlwe_in, su, tlwe_in.shape[0] - 1, unfolding, N,
                                                          kp1 - 1, l, Bg_bit, torus_base, mode)
    return out


def tlwe_keyswitch(tlwe_in, ksk, base_bit):
    tlwe_in = _c(tlwe_in, np.uint64)
    ksk = _c(ksk, np.uint64)
    n_in, t, _, w = ksk.shape
    out = np.empty(w, np.uint64)
    lib().oracle_tlwe_keyswitch(out, tlwe_in, ksk, n_in, w - 1, t, base_bit)
    return out


def tlwe_phase(c, s):
    c = _c(c, np.uint64)
    s = _c(s, np.uint64)
    return int(lib().oracle_tlwe_phase(c, s, s.shape[0]))


def trlwe_phase(c, s):
    c = _c(c, np.uint64)
    s = _c(s, np.uint64).reshape(-1)
    k, N = c.shape[0] - 1, c.shape[1]
    out = np.empty(N, np.uint64)
    lib().oracle_trlwe_phase(out, c, s, N, k)
    return out


def trgsw_mul_trlwe_exact(trlwe, trgsw_torus, l, Bg_bit):
    trlwe = _c(trlwe, np.uint64)
    trgsw_torus = _c(trgsw_torus, np.uint64)
    k, N = trlwe.shape[0] - 1, trlwe.shape[1]
    out = np.empty((k + 1, N), np.uint64)
    lib().oracle_trgsw_mul_trlwe_exact(out, trlwe, trgsw_torus, N, k, l, Bg_bit)
    return out


def blind_rotate_exact(acc, a, bsk_torus, l, Bg_bit):
    acc = _c(acc, np.uint64).copy()
    a = _c(a, np.uint64)
    bsk_torus = _c(bsk_torus, np.uint64)
    k, N = acc.shape[0] - 1, acc.shape[1]
    lib().oracle_blind_rotate_exact(acc, a, bsk_torus, a.shape[0], N, k, l, Bg_bit)
    return acc


def signed_diff(a, b):
    """(a - b) mod 2^64 interpreted as signed int64 (torus distance)."""
    return (np.asarray(a, np.uint64) - np.asarray(b, np.uint64)).view(np.int64)
