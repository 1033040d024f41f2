"""TEST INFRASTRUCTURE ONLY -- full-size parity of the CUDA path against the UNMODIFIED reference on the SAME keys.

Keys, test vector and inputs are made by the reference library itself exactly as ``/root/reference/test/benchmark.c:97-114``
does (``tlwe_new_binary_key``, ``trlwe_new_binary_key``, ``trlwe_extract_tlwe_key``, ``trgsw_new_key``,
``tlwe_new_KS_key``, ``new_bootstrap_key(.., 1)``, ``tlwe_new_sample``); the same in-memory handle trees are then given
to the reference's ``functional_bootstrap`` + ``tlwe_keyswitch`` and to ``libmosfhet_b200.so`` through the drop-in batch
entry points.  Used by ``tests/test_fullsize_parity.py`` and by ``bench.py``'s ``parity`` field (checker side only: the
product never imports this module).

What is compared:
* decrypted messages of PBS+KS outputs: identical;
* phase of the bootstrap output under the extracted key, GPU vs reference, per kernel policy -- next to the same figure
  between two builds of the reference itself (AVX-512 vs FMA SPQLIOS) on the same keys, which is the yardstick when the
  gadget is coarse (two correct FFTs flip different decomposition roundings);
* ``tlwe_keyswitch_batch`` on the reference's own bootstrap outputs (padded with fresh encryptions to ``ks_count``
  ciphertexts so that the full-batch kernel instantiation runs): bit-exact.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from mosfhet_b200 import abi  # noqa: E402  (struct mirrors only)
from oracle import ref as reflib  # noqa: E402

U64 = np.uint64


def _signed(a, b):
    return np.abs((np.asarray(a, U64) - np.asarray(b, U64)).view(np.int64)).astype(np.float64)


def _log2(x):
    return float(np.log2(max(float(x), 1.0)))


class ReferenceSetup:
    """Keys and inputs drawn by the reference (benchmark.c:97-114) for one parameter set."""

    def __init__(self, P, n_inputs: int, torus_base: int = 4, variant: str | None = None, seed: int = 1):
        self.P, self.torus_base, self.n_inputs = P, torus_base, n_inputs
        R = self.R = reflib.load(variant)
        R.init_fft(P.N)
        self.key_tlwe = R.tlwe_new_binary_key(P.n, P.lwe_sigma)
        self.key_tlwe_out = R.tlwe_new_binary_key(P.k * P.N, P.rlwe_sigma)
        self.key_trlwe = R.trlwe_new_binary_key(P.N, P.k, P.rlwe_sigma)
        R.trlwe_extract_tlwe_key(self.key_tlwe_out, self.key_trlwe)
        self.trgsw_key = R.trgsw_new_key(self.key_trlwe, P.l, P.Bg_bit)
        self.ksk = R.tlwe_new_KS_key(self.key_tlwe, self.key_tlwe_out, P.t, P.base_bit)
        self.bk = R.new_bootstrap_key(self.trgsw_key, self.key_tlwe, 1)
        log = int(np.log2(2 * torus_base))
        rng = np.random.default_rng(seed)
        self.msgs = rng.integers(0, torus_base, n_inputs)
        self.inputs = [R.tlwe_new_sample(int(m) << (64 - log), self.key_tlwe) for m in self.msgs]
        # test vector: LUT m -> 3m+1 mod torus_base, packed as trlwe_torus_packing does (trlwe.c:662-667)
        self.lut = ((3 * np.arange(torus_base) + 1) % torus_base).astype(U64) << U64(64 - log)
        lut_c = (C.c_uint64 * torus_base)(*[int(x) for x in self.lut])
        self.tv = R.trlwe_new_noiseless_trivial_sample(None, P.k, P.N)
        R.trlwe_torus_packing(self.tv, lut_c, torus_base)

    # ---- the reference's own path, a few host threads (ctypes releases the GIL; the hot path has thread-local scratch)
    def reference_pbs(self, R=None, threads: int = 8):
        R = R or self.R
        outs = [R.tlwe_alloc_sample(self.P.k * self.P.N) for _ in self.inputs]

        def one(i):
            R.functional_bootstrap(outs[i], self.tv, self.inputs[i], self.bk, self.torus_base)
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(one, range(len(self.inputs))))
        return outs

    def reference_ks(self, ins, threads: int = 8):
        R = self.R
        outs = [R.tlwe_alloc_sample(self.P.n) for _ in ins]

        def one(i):
            R.tlwe_keyswitch(outs[i], ins[i], self.ksk)
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(one, range(len(ins))))
        return outs

    def phases(self, cts, key):
        return np.array([self.R.tlwe_phase(c, key) for c in cts], dtype=U64)

    def decode(self, phases):
        log = int(np.log2(2 * self.torus_base))
        return ((phases + (U64(1) << U64(63 - log))) >> U64(64 - log)).astype(np.int64)


def reference_parity(P, api, n_inputs: int = 64, ks_count: int = 640, policies=((0, "auto"),), second_variant: bool = True):
    """Runs the comparison described in the module docstring; returns a JSON-able dict (no assertions here)."""
    t0 = time.perf_counter()
    S = ReferenceSetup(P, n_inputs)
    R = S.R
    res = {"params": f"n={P.n} N={P.N} l={P.l} Bg_bit={P.Bg_bit} t={P.t} base_bit={P.base_bit}", "inputs": n_inputs,
           "reference_variant": R.variant, "key_setup_s": round(time.perf_counter() - t0, 2)}
    t0 = time.perf_counter()
    mid_ref = S.reference_pbs()
    out_ref = S.reference_ks(mid_ref)
    res["reference_s"] = round(time.perf_counter() - t0, 2)
    ph_mid_ref = S.phases(mid_ref, S.key_tlwe_out)
    dec_ref = S.decode(S.phases(out_ref, S.key_tlwe))
    want = (3 * S.msgs + 1) % S.torus_base
    res["reference_wrong"] = int((dec_ref % (2 * S.torus_base) != want).sum())

    # yardstick: a second build of the reference (other SIMD path of the same FFT), same keys and inputs
    if second_variant:
        others = [v for v in reflib.runnable_variants() if v != R.variant and reflib.VARIANT_LAYOUT[v] == R.layout]
        if others:
            R2 = reflib.load(others[0])
            R2.init_fft(P.N)
            d = _signed(S.phases(S.reference_pbs(R2), S.key_tlwe_out), ph_mid_ref)
            res["ref_vs_ref"] = {"variant": others[0], "pbs_phase_max_log2": _log2(d.max()), "pbs_phase_rms_log2": _log2(np.sqrt((d ** 2).mean()))}

    # ---- the CUDA path on the same handle trees -----------------------------------------------------------------------------
    api.set_host_fft_layout(R.layout)
    api.register_bootstrap_key(S.bk)
    api.register_ks_key(S.ksk)
    res["gpu"] = {}
    for pol, name in policies:
        api.set_kernel_policy(pol)
        mid = [R.tlwe_alloc_sample(P.k * P.N) for _ in S.inputs]
        out = [R.tlwe_alloc_sample(P.n) for _ in S.inputs]
        api.functional_bootstrap_batch(mid, S.tv, S.inputs, S.bk, S.torus_base)
        kern = api.last_blind_rotate_kernel()
        api.functional_bootstrap_keyswitch_batch(out, S.tv, S.inputs, S.bk, S.ksk, S.torus_base)
        d = _signed(S.phases(mid, S.key_tlwe_out), ph_mid_ref)
        dec = S.decode(S.phases(out, S.key_tlwe))
        res["gpu"][name] = {"kernel": kern, "messages_identical": bool(np.array_equal(dec, dec_ref)),
                            "pbs_phase_max_log2": _log2(d.max()), "pbs_phase_rms_log2": _log2(np.sqrt((d ** 2).mean()))}
    api.set_kernel_policy(0)

    # ---- key switch alone on the reference's bootstrap outputs: integer, bit-exact ------------------------------------------------
    extra = max(0, ks_count - n_inputs)
    rnd = np.random.default_rng(7).integers(0, 2 ** 63, extra, dtype=np.int64).astype(U64)
    ks_in = list(mid_ref) + [R.tlwe_new_sample(int(x), S.key_tlwe_out) for x in rnd]
    ks_ref = S.reference_ks(ks_in)
    ks_gpu = [R.tlwe_alloc_sample(P.n) for _ in ks_in]
    api.tlwe_keyswitch_batch(ks_gpu, ks_in, S.ksk)
    a = np.stack([abi.tlwe_to_flat(c) for c in ks_gpu])
    b = np.stack([abi.tlwe_to_flat(c) for c in ks_ref])
    res["keyswitch"] = {"count": len(ks_in), "bit_exact": bool(np.array_equal(a, b)), "mismatching_words": int((a != b).sum())}
    api.release_bootstrap_key(S.bk)
    api.release_ks_key(S.ksk)
    # the reference's objects are left to the process (free_* of 250 k separately allocated rows costs seconds)
    res["ok"] = bool(res["keyswitch"]["bit_exact"] and res["reference_wrong"] == 0 and
                     all(g["messages_identical"] for g in res["gpu"].values()))
    return res
