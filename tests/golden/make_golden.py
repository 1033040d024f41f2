"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference library.

Run in the build container (needs oracle/_ref built by oracle/build_ref.sh from /root/reference):

    python tests/golden/make_golden.py

The reference has no golden vectors and its RNG cannot be seeded, so the
fixtures record, for small parameter sets, the keys and inputs the reference drew and every output
of the hot-path functions on them: polynomial_decompose_i, polynomial_torus_to_DFT,
polynomial_DFT_to_torus, the rotations, trgsw_mul_trlwe_DFT, trlwe_from_DFT, blind_rotate,
functional_bootstrap[_wo_extract], programmable_bootstrap, multivalue_bootstrap_CLOT21,
trlwe_extract_tlwe and tlwe_keyswitch.  Fourier-domain arrays are stored in the slot order of the
build that produced them (``layout``: 1 = SPQLIOS, 2 = FFNT).
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from mosfhet_b200 import abi  # noqa: E402
from oracle import ref as reflib  # noqa: E402

SETS = {
    # name: n, N, k, l, Bg_bit, t, base_bit, lwe_sigma, rlwe_sigma, batch
    "tiny_k1": dict(n=12, N=256, k=1, l=2, Bg_bit=8, t=3, base_bit=2, lwe_sigma=2.0**-30, rlwe_sigma=2.0**-45, batch=6),
    "tiny_k2": dict(n=8, N=128, k=2, l=3, Bg_bit=6, t=4, base_bit=3, lwe_sigma=2.0**-30, rlwe_sigma=2.0**-45, batch=4),
    "small_l4": dict(n=10, N=512, k=1, l=4, Bg_bit=9, t=3, base_bit=4, lwe_sigma=2.0**-30, rlwe_sigma=2.0**-44, batch=4),
}


def rand_u64(R, count):
    buf = abi.aligned_empty(max(count, 64), np.uint64)[:count]   # aes_prng wants aligned, >= 256 B
    R.generate_random_bytes(max(count, 64) * 8, buf.ctypes.data_as(C.c_void_p))
    return buf


def key_words(ptr, n):
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(np.uint64).copy()


def gen(name, P, variant, full=True):
    R = reflib.load(variant)
    n, N, k, l, Bg_bit, t, base_bit = (P[x] for x in ("n", "N", "k", "l", "Bg_bit", "t", "base_bit"))
    B = P["batch"]
    R.init_fft(N)
    out = dict(params=np.array([n, N, k, l, Bg_bit, t, base_bit], np.int32), layout=np.int32(R.layout))

    key_lwe = R.tlwe_new_binary_key(n, P["lwe_sigma"])
    key_rlwe = R.trlwe_new_binary_key(N, k, P["rlwe_sigma"])
    key_ext = R.tlwe_new_binary_key(k * N, P["rlwe_sigma"])
    R.trlwe_extract_tlwe_key(key_ext, key_rlwe)
    trgsw_key = R.trgsw_new_key(key_rlwe, l, Bg_bit)
    out["lwe_key"] = key_words(key_lwe.contents.s, n)
    out["rlwe_key"] = np.stack([key_words(key_rlwe.contents.s[i].contents.coeffs, N) for i in range(k)])
    out["ext_key"] = key_words(key_ext.contents.s, k * N)

    # --- bootstrapping key, built row by row exactly as new_bootstrap_key_wo_unfolding
    #     (bootstrap.c:3-21) so that the torus-domain TRGSW can be recorded as well
    rows = (k + 1) * l
    bsk_host = np.empty((n, rows, k + 1, N), np.float64)
    bsk_torus = np.empty((n, rows, k + 1, N), np.uint64)
    dft = R.trgsw_alloc_new_DFT_sample(l, Bg_bit, k, N)
    for i in range(n):
        tmp = R.trgsw_new_monomial_sample(int(out["lwe_key"][i]), 0, trgsw_key)
        R.trgsw_to_DFT(dft, tmp)
        bsk_torus[i] = abi.trgsw_to_flat(tmp, k)
        bsk_host[i] = abi.trgsw_dft_to_flat(dft, k)
        R.free_trgsw(tmp)
    out["bsk_host"] = bsk_host
    out["bsk_torus"] = bsk_torus
    hbsk = abi.HostBootstrapKey(bsk_host, k, l, Bg_bit)

    # --- single-polynomial primitives
    poly = rand_u64(R, N)
    hp = abi.HostTRLWE(np.stack([poly] * (k + 1)))         # reuse its polynomial structs
    p_in = hp.struct.b
    p_out = R.polynomial_new_torus_polynomial(N)
    d_out = R.polynomial_new_DFT_polynomial(N)
    out["poly"] = poly
    dec = []
    for j in range(l):
        R.polynomial_decompose_i(p_out, p_in, Bg_bit, l, j)
        dec.append(key_words(p_out.contents.coeffs, N))
    out["poly_decomp"] = np.stack(dec).view(np.int64)
    # forward transform of a digit polynomial (the only kind of input the path transforms)
    hd = abi.HostTRLWE(np.stack([dec[0]] * (k + 1)))
    R.polynomial_torus_to_DFT(d_out, hd.struct.b)
    out["digit_dft_host"] = np.ctypeslib.as_array(d_out.contents.coeffs, shape=(N,)).copy()
    # forward transform of the monomial X (pins the slot order)
    mono = np.zeros(N, np.uint64); mono[1] = 1
    hm = abi.HostTRLWE(np.stack([mono] * (k + 1)))
    R.polynomial_torus_to_DFT(d_out, hm.struct.b)
    out["monomial_dft_host"] = np.ctypeslib.as_array(d_out.contents.coeffs, shape=(N,)).copy()
    rots = []
    rot_amounts = np.array([0, 1, N - 1, N, N + 1, 2 * N - 1, N // 3, N + N // 5], np.int32)
    for a in rot_amounts:
        R.torus_polynomial_mul_by_xai(p_out, p_in, int(a))
        r1 = key_words(p_out.contents.coeffs, N)
        R.torus_polynomial_mul_by_xai_minus_1(p_out, p_in, int(a))
        rots.append(np.stack([r1, key_words(p_out.contents.coeffs, N)]))
    out["rot_amounts"] = rot_amounts
    out["rot_out"] = np.stack(rots)

    # --- one external product + inverse transform (trgsw.c:385, trlwe.c:629)
    msg = rand_u64(R, N)
    hmsg = abi.HostTRLWE(np.stack([msg] * (k + 1)))
    c_in = R.trlwe_new_sample(hmsg.struct.b, key_rlwe)
    ep_dft = R.trlwe_alloc_new_DFT_sample(k, N)
    ep_out = R.trlwe_alloc_new_sample(k, N)
    R.trgsw_mul_trlwe_DFT(ep_dft, c_in, hbsk.trgsw[0].handle)
    R.trlwe_from_DFT(ep_out, ep_dft)
    out["ep_in"] = abi.trlwe_to_flat(c_in)
    out["ep_dft_host"] = abi.trlwe_dft_to_flat(ep_dft)
    out["ep_out"] = abi.trlwe_to_flat(ep_out)

    if not full:
        return out

    # --- key-switching key and key switch (tlwe.c:193-212, 289-303)
    ksk = R.tlwe_new_KS_key(key_lwe, key_ext, t, base_bit)
    out["ksk"] = abi.ks_key_to_flat(ksk)

    # --- bootstraps
    torus_base = 4
    lut_vals = rand_u64(R, torus_base)
    tv = R.trlwe_alloc_new_sample(k, N)
    R.trlwe_torus_packing(tv, lut_vals.ctypes.data_as(C.POINTER(C.c_uint64)), torus_base)
    out["lut_vals"] = lut_vals
    out["tv"] = abi.trlwe_to_flat(tv)
    msgs = np.arange(B) % torus_base
    ins, br_out, fbwo_out, fb_out, ks_out, pb_out, ext_out = [], [], [], [], [], [], []
    for b in range(B):
        m = int(msgs[b]) << (64 - 3)                       # int2torus(m, log2(2*torus_base))
        c = R.tlwe_new_sample(m, key_lwe)
        ins.append(abi.tlwe_to_flat(c))
        # blind_rotate in place on a copy of tv (bootstrap.c:107)
        acc = abi.HostTRLWE(out["tv"])
        R.blind_rotate(acc.handle, c.contents.a, hbsk.struct.s, n)
        br_out.append(acc.flat().copy())
        wo = abi.HostTRLWE.zeros(k, N)
        R.functional_bootstrap_wo_extract(wo.handle, tv, c, hbsk.handle, torus_base)
        fbwo_out.append(wo.flat().copy())
        o = R.tlwe_alloc_sample(k * N)
        R.functional_bootstrap(o, tv, c, hbsk.handle, torus_base)
        fb_out.append(abi.tlwe_to_flat(o))
        ko = R.tlwe_alloc_sample(n)
        R.tlwe_keyswitch(ko, o, ksk)
        ks_out.append(abi.tlwe_to_flat(ko))
        po = R.tlwe_alloc_sample(k * N)
        R.programmable_bootstrap(po, tv, c, hbsk.handle, 3, 1, 1)
        pb_out.append(abi.tlwe_to_flat(po))
        ex = []
        for idx in (0, 1, N // 2, N - 1):
            eo = R.tlwe_alloc_sample(k * N)
            R.trlwe_extract_tlwe(eo, wo.handle, idx)
            ex.append(abi.tlwe_to_flat(eo))
        ext_out.append(np.stack(ex))
    out["msgs"] = msgs.astype(np.int32)
    out["tlwe_in"] = np.stack(ins)
    out["blind_rotate_out"] = np.stack(br_out)
    out["fb_wo_extract_out"] = np.stack(fbwo_out)
    out["fb_out"] = np.stack(fb_out)
    out["ks_out"] = np.stack(ks_out)
    out["pb_out"] = np.stack(pb_out)
    out["pb_args"] = np.array([3, 1, 1], np.int32)
    out["extract_idx"] = np.array([0, 1, N // 2, N - 1], np.int32)
    out["extract_out"] = np.stack(ext_out)

    # --- multivalue_bootstrap_CLOT21 (bootstrap.c:222-230) with n_luts LUTs of torus_base=2
    n_luts, tb2 = 4, 2
    lut2 = rand_u64(R, n_luts * tb2)
    tv2 = R.trlwe_alloc_new_sample(k, N)
    R.trlwe_torus_packing_many_LUT(tv2, lut2.ctypes.data_as(C.POINTER(C.c_uint64)), tb2, n_luts)
    outs = (abi.TLWE * n_luts)(*[R.tlwe_alloc_sample(k * N) for _ in range(n_luts)])
    c = R.tlwe_new_sample(1 << (64 - 2), key_lwe)            # message 1 of torus_base 2
    R.multivalue_bootstrap_CLOT21(outs, tv2, c, hbsk.handle, tb2, n_luts)
    out["mv_lut"] = lut2
    out["mv_tv"] = abi.trlwe_to_flat(tv2)
    out["mv_in"] = abi.tlwe_to_flat(c)
    out["mv_out"] = np.stack([abi.tlwe_to_flat(outs[i]) for i in range(n_luts)])
    out["mv_args"] = np.array([tb2, n_luts], np.int32)
    return out


def gen_mv(P, variant):
    """multivalue_bootstrap_phase1 / phase2 (bootstrap.c:232-265), as test_functional_mv_bootstrap (tests.c:1793-1827)."""
    R = reflib.load(variant)
    R.multivalue_bootstrap_phase1.restype = None
    R.multivalue_bootstrap_phase1.argtypes = [C.POINTER(abi.TRLWE), abi.TLWE, abi.Bootstrap_Key, C.c_int]
    R.multivalue_bootstrap_phase2.restype = None
    R.multivalue_bootstrap_phase2.argtypes = [abi.TLWE, C.POINTER(C.c_int), C.POINTER(abi.TRLWE), C.c_int, C.c_int]
    n, N, k, l, Bg_bit = (P[x] for x in ("n", "N", "k", "l", "Bg_bit"))
    R.init_fft(N)
    out = dict(params=np.array([n, N, k, l, Bg_bit, P["t"], P["base_bit"]], np.int32), layout=np.int32(R.layout))
    key_lwe = R.tlwe_new_binary_key(n, P["lwe_sigma"])
    key_rlwe = R.trlwe_new_binary_key(N, k, P["rlwe_sigma"])
    key_ext = R.tlwe_new_binary_key(k * N, P["rlwe_sigma"])
    R.trlwe_extract_tlwe_key(key_ext, key_rlwe)
    trgsw_key = R.trgsw_new_key(key_rlwe, l, Bg_bit)
    bk = R.new_bootstrap_key(trgsw_key, key_lwe, 1)
    out["lwe_key"] = key_words(key_lwe.contents.s, n)
    out["ext_key"] = key_words(key_ext.contents.s, k * N)
    out["bsk_host"] = abi.bootstrap_key_to_flat(bk)
    tb, log_tb = 4, 2
    luts = np.array([[1, 2, 3, 0], [3, 3, 0, 1], [0, 1, 2, 3]], np.int32)
    ins, p1, p2 = [], [], []
    for m in range(tb):
        c = R.tlwe_new_sample(m << 61, key_lwe)
        ins.append(abi.tlwe_to_flat(c))
        rots = [abi.HostTRLWE.zeros(k, N) for _ in range(tb + 1)]
        arr = abi.handle_array(rots, abi.TRLWE)
        R.multivalue_bootstrap_phase1(arr, c, bk, tb)
        p1.append(np.stack([r.polys.copy() for r in rots]))
        row = []
        for lut in luts:
            o = R.tlwe_alloc_sample(k * N)
            R.multivalue_bootstrap_phase2(o, (C.c_int * tb)(*[int(x) for x in lut]), arr, tb, log_tb)
            row.append(abi.tlwe_to_flat(o))
        p2.append(np.stack(row))
    out["mv_luts"] = luts
    out["mv_in"] = np.stack(ins)
    out["mv_phase1"] = np.stack(p1)
    out["mv_phase2"] = np.stack(p2)
    return out


def gen_cb(variant):
    """circuit_bootstrap_2 (bootstrap.c:324-345) and its two table key switches, from a build with
    UNCOMPRESSED key-switching rows (A_PRNG=none: the fma / portable variants)."""
    R = reflib.load(variant)
    sig = {
        "trlwe_new_priv_SK_KS_key_N2": (abi.Generic_KS_Key, [abi.TRLWE_Key, abi.TLWE_Key, C.c_int, C.c_int]),
        "trlwe_new_packing1_KS_key": (abi.Generic_KS_Key, [abi.TRLWE_Key, abi.TLWE_Key, C.c_int, C.c_int]),
        "trlwe_priv_keyswitch": (None, [abi.TRLWE, abi.TLWE, abi.Generic_KS_Key]),
        "trlwe_packing1_keyswitch": (None, [abi.TRLWE, abi.TLWE, abi.Generic_KS_Key]),
        "circuit_bootstrap_2": (None, [abi.TRGSW, abi.TLWE, abi.Bootstrap_Key, abi.Generic_KS_Key, abi.Generic_KS_Key]),
        "trgsw_alloc_new_sample": (abi.TRGSW, [C.c_int, C.c_int, C.c_int, C.c_int]),
        "trlwe_new_priv_KS_key": (C.POINTER(abi.TRLWE_KS_Key), [abi.TRLWE_Key, abi.TRLWE_Key, C.c_int, C.c_int]),
        "trlwe_keyswitch": (None, [abi.TRLWE, abi.TRLWE, abi.TRLWE_KS_Key]),
        "trlwe_priv_keyswitch_2": (None, [abi.TRLWE, abi.TRLWE, C.POINTER(abi.TRLWE_KS_Key)]),
        "circuit_bootstrap": (None, [abi.TRGSW, abi.TLWE, abi.Bootstrap_Key, abi.Generic_KS_Key, abi.Generic_KS_Key]),
        "circuit_bootstrap_3": (None, [abi.TRGSW, abi.TLWE, abi.Bootstrap_Key, C.POINTER(abi.TRLWE_KS_Key), abi.Generic_KS_Key]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(R.lib, name); fn.restype, fn.argtypes = res, args
    n, N, k, l, Bg_bit, t, base_bit = 10, 64, 1, 2, 8, 6, 2
    R.init_fft(N)
    out = dict(params=np.array([n, N, k, l, Bg_bit, t, base_bit], np.int32), layout=np.int32(R.layout))
    key_lwe = R.tlwe_new_binary_key(n, 2.0 ** -30)
    key_rlwe = R.trlwe_new_binary_key(N, k, 2.0 ** -45)
    key_ext = R.tlwe_new_binary_key(k * N, 2.0 ** -45)
    R.trlwe_extract_tlwe_key(key_ext, key_rlwe)
    trgsw_key = R.trgsw_new_key(key_rlwe, l, Bg_bit)
    bk = R.new_bootstrap_key(trgsw_key, key_lwe, 1)
    kska = R.lib.trlwe_new_priv_SK_KS_key_N2(key_rlwe, key_ext, t, base_bit)
    kskb = R.lib.trlwe_new_packing1_KS_key(key_rlwe, key_ext, t, base_bit)
    out["lwe_key"] = key_words(key_lwe.contents.s, n)
    out["rlwe_key"] = np.stack([key_words(key_rlwe.contents.s[i].contents.coeffs, N) for i in range(k)])
    out["bsk_host"] = abi.bootstrap_key_to_flat(bk)
    out["kska"] = abi.generic_ks_key_to_flat(kska)
    out["kskb"] = abi.generic_ks_key_to_flat(kskb)
    # FFT-based private key switch (keyswitch.c:40-63) with its own decomposition (t2 digits of base_bit2 bits)
    t2, base_bit2 = 10, 2
    kska2 = R.lib.trlwe_new_priv_KS_key(key_rlwe, key_rlwe, t2, base_bit2)
    out["kska2"] = np.stack([abi.trlwe_ks_key_to_flat(kska2[i])[0] for i in range(2)])      # [2, t2, 2, N]
    out["params2"] = np.array([t2, base_bit2], np.int32)
    ins, cbs, pk, pv, tls = [], [], [], [], []
    cb1, cb3, rk_in, rk_out, pv2_out = [], [], [], [], []
    for m in (0, 1, 0, 1):
        c = R.tlwe_new_sample(m << 62, key_lwe)            # LWE(m/4), as tests.c:980, 998
        ins.append(abi.tlwe_to_flat(c))
        o = R.lib.trgsw_alloc_new_sample(l, Bg_bit, k, N)
        R.lib.circuit_bootstrap_2(o, c, bk, kska, kskb)
        cbs.append(abi.trgsw_to_flat(o, k))
        tl = rand_u64(R, k * N + 1)                          # the key switches alone on a random TLWE
        tls.append(tl.copy())
        ht = abi.HostTLWE(tl)
        a = abi.HostTRLWE.zeros(k, N); R.lib.trlwe_priv_keyswitch(a.handle, ht.handle, kska); pv.append(a.polys.copy())
        b = abi.HostTRLWE.zeros(k, N); R.lib.trlwe_packing1_keyswitch(b.handle, ht.handle, kskb); pk.append(b.polys.copy())
        o1 = R.lib.trgsw_alloc_new_sample(l, Bg_bit, k, N)
        R.lib.circuit_bootstrap(o1, c, bk, kska, kskb)
        cb1.append(abi.trgsw_to_flat(o1, k))
        o3 = R.lib.trgsw_alloc_new_sample(l, Bg_bit, k, N)
        R.lib.circuit_bootstrap_3(o3, c, bk, kska2, kskb)
        cb3.append(abi.trgsw_to_flat(o3, k))
        rin = abi.HostTRLWE(rand_u64(R, (k + 1) * N).reshape(k + 1, N))     # the FFT key switches alone
        rk_in.append(rin.polys.copy())
        ro = abi.HostTRLWE.zeros(k, N); R.lib.trlwe_keyswitch(ro.handle, rin.handle, kska2[0]); rk_out.append(ro.polys.copy())
        ro = abi.HostTRLWE.zeros(k, N); R.lib.trlwe_priv_keyswitch_2(ro.handle, rin.handle, kska2); pv2_out.append(ro.polys.copy())
    out["cb_in"] = np.stack(ins); out["cb_out"] = np.stack(cbs)
    out["ks_in"] = np.stack(tls); out["priv_out"] = np.stack(pv); out["pack_out"] = np.stack(pk)
    out["cb_msgs"] = np.array([0, 1, 0, 1], np.int32)
    out["cb1_out"] = np.stack(cb1); out["cb3_out"] = np.stack(cb3)
    out["rks_in"] = np.stack(rk_in); out["rks_out"] = np.stack(rk_out); out["priv2_out"] = np.stack(pv2_out)
    return out


def gen_r4(variant):
    """TRGSW-accumulator bootstrap (bootstrap.c:267-306) and the unfolded blind rotation (bootstrap.c:23-48,
    124-190) with an unfolding = 2 key, recorded from the unmodified reference."""
    R = reflib.load(variant)
    sig = {
        "functional_bootstrap_trgsw_phase1": (None, [abi.TRGSW_DFT, abi.TLWE, abi.Bootstrap_Key, C.c_int]),
        "functional_bootstrap_trgsw_phase2": (None, [abi.TLWE, abi.TRGSW_DFT, abi.TRLWE]),
        "blind_rotate_unfolded": (None, [abi.TRLWE, C.POINTER(C.c_uint64), C.POINTER(abi.TRGSW), C.c_int, C.c_int]),
        "multivalue_bootstrap_UBR_phase1": (None, [C.POINTER(abi.TRGSW_DFT), abi.TLWE, abi.Bootstrap_Key]),
        "multivalue_bootstrap_UBR_phase2": (None, [abi.TLWE, abi.TRLWE, abi.TLWE, C.POINTER(abi.TRGSW_DFT), abi.Bootstrap_Key, C.c_int]),
        "trgsw_alloc_new_DFT_sample": (abi.TRGSW_DFT, [C.c_int, C.c_int, C.c_int, C.c_int]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(R.lib, name); fn.restype, fn.argtypes = res, args
    n, N, k, l, Bg_bit, t, base_bit = 12, 64, 1, 2, 10, 4, 2
    unfolding = 2
    R.init_fft(N)
    out = dict(params=np.array([n, N, k, l, Bg_bit, t, base_bit], np.int32), layout=np.int32(R.layout),
               unfolding=np.int32(unfolding))
    key_lwe = R.tlwe_new_binary_key(n, 2.0 ** -30)
    key_rlwe = R.trlwe_new_binary_key(N, k, 2.0 ** -45)
    key_ext = R.tlwe_new_binary_key(k * N, 2.0 ** -45)
    R.trlwe_extract_tlwe_key(key_ext, key_rlwe)
    trgsw_key = R.trgsw_new_key(key_rlwe, l, Bg_bit)
    bk = R.new_bootstrap_key(trgsw_key, key_lwe, 1)
    bku = R.new_bootstrap_key(trgsw_key, key_lwe, unfolding)
    out["lwe_key"] = key_words(key_lwe.contents.s, n)
    out["rlwe_key"] = np.stack([key_words(key_rlwe.contents.s[i].contents.coeffs, N) for i in range(k)])
    out["ext_key"] = key_words(key_ext.contents.s, k * N)
    out["bsk_host"] = abi.bootstrap_key_to_flat(bk)
    n_su = (n // unfolding) * (1 << unfolding)
    out["su"] = np.stack([abi.trgsw_to_flat(bku.contents.su[i], k) for i in range(n_su)])
    torus_base = 4
    lut = rand_u64(R, torus_base)
    tvh = abi.HostTRLWE(np.stack([np.zeros(N, np.uint64), np.repeat(lut, N // torus_base)]))
    out["lut"] = lut; out["tv"] = tvh.polys.copy()
    ins, p1, p2, fbu, bru, ubr1, ubr2, bru_in = [], [], [], [], [], [], [], []
    for m in (0, 1, 2, 3):
        c = R.tlwe_new_sample(m << 61, key_lwe)                                   # m/8 as tests.c:1753 (m = 1)
        ins.append(abi.tlwe_to_flat(c))
        g = R.lib.trgsw_alloc_new_DFT_sample(l, Bg_bit, k, N)
        R.lib.functional_bootstrap_trgsw_phase1(g, c, bk, torus_base)
        p1.append(abi.trgsw_dft_to_flat(g, k))
        o = abi.HostTLWE.zeros(k * N)
        R.lib.functional_bootstrap_trgsw_phase2(o.handle, g, tvh.handle)
        p2.append(o.flat())
        # functional_bootstrap_wo_extract through the unfolding = 2 key (bootstrap.c:197)
        w = abi.HostTRLWE.zeros(k, N)
        R.functional_bootstrap_wo_extract(w.handle, tvh.handle, c, bku, torus_base)
        fbu.append(w.polys.copy())
        acc = abi.HostTRLWE(rand_u64(R, (k + 1) * N).reshape(k + 1, N))            # blind_rotate_unfolded alone
        bru_in.append(acc.polys.copy())
        a_in = abi.tlwe_to_flat(c)[:n].copy()
        R.lib.blind_rotate_unfolded(acc.handle, a_in.ctypes.data_as(C.POINTER(C.c_uint64)), bku.contents.su, n, unfolding)
        bru.append(acc.polys.copy())
        sa = [R.lib.trgsw_alloc_new_DFT_sample(l, Bg_bit, k, N) for _ in range(n // unfolding)]
        sa_arr = (abi.TRGSW_DFT * len(sa))(*sa)
        R.lib.multivalue_bootstrap_UBR_phase1(sa_arr, c, bku)
        ubr1.append(np.stack([abi.trgsw_dft_to_flat(x, k) for x in sa]))
        o = abi.HostTLWE.zeros(k * N)
        R.lib.multivalue_bootstrap_UBR_phase2(o.handle, tvh.handle, c, sa_arr, bku, torus_base)
        ubr2.append(o.flat())
    out["bru_in"] = np.stack(bru_in)
    out["r4_in"] = np.stack(ins); out["trgsw_p1"] = np.stack(p1); out["trgsw_p2"] = np.stack(p2)
    out["fbu_out"] = np.stack(fbu); out["bru_out"] = np.stack(bru)
    out["ubr_p1"] = np.stack(ubr1); out["ubr_p2"] = np.stack(ubr2)
    out["msgs"] = np.arange(4, dtype=np.int32)
    return out


def gen_mvx(variant):
    """Integer extraction family of the multi-ciphertext caller (trlwe.c:554-620): recorded from the reference on random
    TRLWE inputs for (k, N) = (1, 256) and (2, 128); every output is bit-exact by construction."""
    R = reflib.load(variant)
    for name, res, args in (("trlwe_extract_tlwe_addto", None, [abi.TLWE, abi.TRLWE, C.c_int]),
                            ("trlwe_extract_tlwe_subto", None, [abi.TLWE, abi.TRLWE, C.c_int]),
                            ("trlwe_mv_extract_tlwe", None, [C.POINTER(abi.TLWE), abi.TRLWE, C.c_int]),
                            ("trlwe_mv_extract_tlwe_scaling", None, [abi.TLWE, abi.TRLWE, C.c_int]),
                            ("trlwe_mv_extract_tlwe_scaling_addto", None, [abi.TLWE, abi.TRLWE, C.c_int]),
                            ("trlwe_mv_extract_tlwe_scaling_subto", None, [abi.TLWE, abi.TRLWE, C.c_int])):
        fn = getattr(R.lib, name)
        fn.restype, fn.argtypes = res, args
    out = dict(params=np.array([0, 0, 0, 0, 0, 0, 0], np.int32), layout=np.int32(R.layout))
    shapes = [(1, 256), (2, 128)]
    out["shapes"] = np.array(shapes, np.int32)
    out["idx_list"] = np.array([0, 1, 17, 127], np.int32)
    out["scales"] = np.array([1, 2, 4, 8, 16], np.int32)
    out["amounts"] = np.array([2, 4, 8], np.int32)
    for si, (k, N) in enumerate(shapes):
        W = k * N + 1
        tr = abi.HostTRLWE(rand_u64(R, (k + 1) * N).reshape(k + 1, N))
        start = rand_u64(R, W)
        out[f"s{si}_trlwe"] = tr.polys.copy()
        out[f"s{si}_start"] = start.copy()
        for idx in out["idx_list"]:
            for nm, sign in (("addto", +1), ("subto", -1)):
                o = abi.HostTLWE(start)
                getattr(R.lib, f"trlwe_extract_tlwe_{nm}")(o.handle, tr.handle, int(idx))
                out[f"s{si}_extract_{nm}_{int(idx)}"] = o.flat()
        for amount in out["amounts"]:
            outs = [abi.HostTLWE.zeros(k * N) for _ in range(int(amount))]
            R.lib.trlwe_mv_extract_tlwe(abi.handle_array(outs, abi.TLWE), tr.handle, int(amount))
            out[f"s{si}_mv_{int(amount)}"] = np.stack([o.flat() for o in outs])
        for scale in out["scales"]:
            o = abi.HostTLWE.zeros(k * N)
            R.lib.trlwe_mv_extract_tlwe_scaling(o.handle, tr.handle, int(scale))
            out[f"s{si}_scaling_{int(scale)}"] = o.flat()
            for nm in ("addto", "subto"):
                o = abi.HostTLWE(start)
                getattr(R.lib, f"trlwe_mv_extract_tlwe_scaling_{nm}")(o.handle, tr.handle, int(scale))
                out[f"s{si}_scaling_{nm}_{int(scale)}"] = o.flat()
    return out


def main():
    if "--mvx-only" in sys.argv:
        data = gen_mvx("avx512")
        path = os.path.join(HERE, "tiny_mvx.npz")
        np.savez_compressed(path, **data)
        print(path, os.path.getsize(path) // 1024, "KiB")
        return
    if "--r4-only" in sys.argv:
        data = gen_r4("fma")
        path = os.path.join(HERE, "tiny_r4_spqlios.npz")
        np.savez_compressed(path, **data)
        print(path, os.path.getsize(path) // 1024, "KiB")
        return
    if "--cb-only" in sys.argv:
        data = gen_cb("fma")
        path = os.path.join(HERE, "tiny_cb_spqlios.npz")
        np.savez_compressed(path, **data)
        print(path, os.path.getsize(path) // 1024, "KiB")
        return
    if "--mv-only" in sys.argv:
        data = gen_mv(SETS["tiny_k1"], "avx512")
        path = os.path.join(HERE, "tiny_k1_mv_spqlios.npz")
        np.savez_compressed(path, **data)
        print(path, os.path.getsize(path) // 1024, "KiB")
        return
    for name, P in SETS.items():
        data = gen(name, P, "avx512")
        path = os.path.join(HERE, f"{name}_spqlios.npz")
        np.savez_compressed(path, **data)
        print(path, os.path.getsize(path) // 1024, "KiB")
    # FFNT slot order: primitives + one external product only
    data = gen("tiny_k1", SETS["tiny_k1"], "portable", full=False)
    path = os.path.join(HERE, "tiny_k1_ffnt.npz")
    np.savez_compressed(path, **data)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
