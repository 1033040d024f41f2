"""Host-side mirror of the reference operators for the accelerated path.

Three levels, all thin bindings over ``libmosfhet_b200.so`` (no computation happens in Python):

1. **Reference names over reference handles** -- ``functional_bootstrap(out, tv, in_, key, torus_base)``
   etc. take ``abi.Host*`` objects (or raw ctypes handles produced by any C library with the
   ``mosfhet.h`` layouts) and have the argument meaning, ownership and error behaviour of
   ``/root/reference/include/mosfhet.h:227,263,277,344,409-412,425``.
2. **Batched variants** over lists of handles.
3. **Flat calls** on numpy arrays (host) and raw device pointers / torch CUDA tensors (device).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib, abi
from .params import Params

FFT_AUTO, FFT_SPQLIOS, FFT_FFNT, FFT_NATURAL = 0, 1, 2, 3


def lib():
    return _lib.load()


def _h(x):
    return getattr(x, "handle", x)


# ---------------------------------------------------------------------------------------------
# runtime
# ---------------------------------------------------------------------------------------------
def device_count() -> int:
    return int(lib().mb200_device_count())


def init(device: int = 0) -> None:
    """Select the CUDA device; aborts the process if there is none (no CPU fallback)."""
    lib().mb200_init(device)


def init_multi(ndev: int = 0) -> int:
    """Multi-GPU mode inside the library: batched drop-in calls shard over ``ndev`` devices (0 = all); returns the count."""
    return int(lib().mb200_init_multi(ndev))


def require_gpu() -> None:
    if device_count() < 1:
        raise RuntimeError("mosfhet_b200: no CUDA device visible and there is no CPU fallback")


def shutdown() -> None:
    lib().mb200_shutdown()


def synchronize() -> None:
    lib().mb200_device_synchronize()


def set_host_fft_layout(layout: int) -> None:
    lib().mb200_set_host_fft_layout(layout)


def host_slot_exponents(layout: int, N: int) -> np.ndarray:
    e = np.empty(N // 2, np.int32)
    lib().mb200_host_slot_exponents(layout, N, e.ctypes.data_as(C.POINTER(C.c_int32)))
    return e


def launch_count() -> int:
    return int(lib().mb200_launch_count())


def reset_launch_count() -> None:
    lib().mb200_reset_launch_count()


def last_blind_rotate_kernel() -> str:
    return lib().mb200_last_blind_rotate_kernel().decode()


def measure_fp64_tflops(iters: int = 20000) -> float:
    return float(lib().mb200_measure_fp64_tflops(iters))


def set_kernel_policy(policy: int) -> None:
    lib().mb200_set_kernel_policy(policy)


# ---------------------------------------------------------------------------------------------
# (1) reference names over reference handles
# ---------------------------------------------------------------------------------------------
def functional_bootstrap(out, tv, in_, key, torus_base: int) -> None:
    lib().functional_bootstrap(_h(out), _h(tv), _h(in_), _h(key), torus_base)


def functional_bootstrap_wo_extract(out, tv, in_, key, torus_base: int) -> None:
    lib().functional_bootstrap_wo_extract(_h(out), _h(tv), _h(in_), _h(key), torus_base)


def programmable_bootstrap(out, tv, in_, key, precision: int, kappa: int, theta: int) -> None:
    lib().programmable_bootstrap(_h(out), _h(tv), _h(in_), _h(key), precision, kappa, theta)


def blind_rotate(tv, a, s, size: int) -> None:
    """In place on ``tv`` (bootstrap.c:118).  ``a``: uint64 array / pointer; ``s``: TRGSW_DFT array."""
    if isinstance(a, np.ndarray):
        a = a.ctypes.data_as(C.POINTER(C.c_uint64))
    lib().blind_rotate(_h(tv), a, s, size)


def blind_rotate_unfolded(tv, a, s, size: int, unfolding: int) -> None:
    """In place on ``tv`` (bootstrap.c:124-148).  ``s``: array of torus-domain TRGSW (``Bootstrap_Key->su``)."""
    if isinstance(a, np.ndarray):
        a = a.ctypes.data_as(C.POINTER(C.c_uint64))
    lib().blind_rotate_unfolded(_h(tv), a, s, size, unfolding)


def functional_bootstrap_trgsw_phase1(out, in_, key, torus_base: int) -> None:
    """TLWE -> TRGSW_DFT(X^-phase) (bootstrap.c:286-296)."""
    lib().functional_bootstrap_trgsw_phase1(_h(out), _h(in_), _h(key), torus_base)


def functional_bootstrap_trgsw_phase2(out, in_, tv) -> None:
    lib().functional_bootstrap_trgsw_phase2(_h(out), _h(in_), _h(tv))


def functional_bootstrap_trgsw_phase1_batch(outs, ins, key, torus_base: int) -> None:
    lib().functional_bootstrap_trgsw_phase1_batch(abi.handle_array(outs, abi.TRGSW_DFT), abi.handle_array(ins, abi.TLWE),
                                                  _h(key), torus_base, len(ins))


def functional_bootstrap_trgsw_phase2_batch(outs, ins, tvs) -> None:
    lib().functional_bootstrap_trgsw_phase2_batch(abi.handle_array(outs, abi.TLWE), abi.handle_array(ins, abi.TRGSW_DFT),
                                                  abi.handle_array(tvs, abi.TRLWE), len(tvs), len(ins))


def multivalue_bootstrap_UBR_phase1(outs, in_, key) -> None:
    """outs: n/unfolding TRGSW_DFT handles (bootstrap.c:151-172)."""
    lib().multivalue_bootstrap_UBR_phase1(abi.handle_array(outs, abi.TRGSW_DFT), _h(in_), _h(key))


def multivalue_bootstrap_UBR_phase2(out, tv, in_, sa, key, torus_base: int) -> None:
    lib().multivalue_bootstrap_UBR_phase2(_h(out), _h(tv), _h(in_), abi.handle_array(sa, abi.TRGSW_DFT), _h(key), torus_base)


def bootstrap_trgsw_phase1_dev(bsk, d_out_trgsw, d_in, l_out, Bg_bit_out, torus_base, count, stream=None):
    lib().mb200_bootstrap_trgsw_phase1_dev(bsk.handle, _ptr(d_out_trgsw), _ptr(d_in), l_out, Bg_bit_out, torus_base, count,
                                           _ptr(stream))


def trgsw_mul_trlwe_DFT(out, in1, in2) -> None:
    lib().trgsw_mul_trlwe_DFT(_h(out), _h(in1), _h(in2))


def trlwe_from_DFT(out, in_) -> None:
    lib().trlwe_from_DFT(_h(out), _h(in_))


def trlwe_extract_tlwe(out, in_, idx: int) -> None:
    lib().trlwe_extract_tlwe(_h(out), _h(in_), idx)


def tlwe_keyswitch(out, in_, ks_key) -> None:
    lib().tlwe_keyswitch(_h(out), _h(in_), _h(ks_key))


def multivalue_bootstrap_CLOT21(out_list, tv, in_, key, torus_base: int, n_luts: int) -> None:
    lib().multivalue_bootstrap_CLOT21(abi.handle_array(out_list, abi.TLWE), _h(tv), _h(in_), _h(key), torus_base, n_luts)


def trlwe_packing1_keyswitch(out, in_, ks_key) -> None:
    lib().trlwe_packing1_keyswitch(_h(out), _h(in_), _h(ks_key))


def trlwe_priv_keyswitch(out, in_, ks_key) -> None:
    lib().trlwe_priv_keyswitch(_h(out), _h(in_), _h(ks_key))


def trlwe_packing1_keyswitch_batch(outs, ins, ks_key) -> None:
    lib().trlwe_packing1_keyswitch_batch(abi.handle_array(outs, abi.TRLWE), abi.handle_array(ins, abi.TLWE), _h(ks_key), len(ins))


def trlwe_priv_keyswitch_batch(outs, ins, ks_key) -> None:
    lib().trlwe_priv_keyswitch_batch(abi.handle_array(outs, abi.TRLWE), abi.handle_array(ins, abi.TLWE), _h(ks_key), len(ins))


def circuit_bootstrap_2(out, in_, key, kska, kskb) -> None:
    """TLWE -> torus-domain TRGSW (bootstrap.c:324-345)."""
    lib().circuit_bootstrap_2(_h(out), _h(in_), _h(key), _h(kska), _h(kskb))


def circuit_bootstrap_2_batch(outs, ins, key, kska, kskb) -> None:
    lib().circuit_bootstrap_2_batch(abi.handle_array(outs, abi.TRGSW), abi.handle_array(ins, abi.TLWE), _h(key),
                                    _h(kska), _h(kskb), len(ins))


def circuit_bootstrap(out, in_, key, kska, kskb) -> None:
    """One functional bootstrap per gadget level (bootstrap.c:309-322)."""
    lib().circuit_bootstrap(_h(out), _h(in_), _h(key), _h(kska), _h(kskb))


def circuit_bootstrap_3(out, in_, key, kska_pair, kskb) -> None:
    """circuit_bootstrap_2 with the FFT-based private key switch (bootstrap.c:347-366); kska_pair: TRLWE_KS_Key[2]."""
    lib().circuit_bootstrap_3(_h(out), _h(in_), _h(key), _h(kska_pair), _h(kskb))


def circuit_bootstrap_batch(outs, ins, key, kska, kskb) -> None:
    lib().circuit_bootstrap_batch(abi.handle_array(outs, abi.TRGSW), abi.handle_array(ins, abi.TLWE), _h(key),
                                  _h(kska), _h(kskb), len(ins))


def circuit_bootstrap_3_batch(outs, ins, key, kska_pair, kskb) -> None:
    lib().circuit_bootstrap_3_batch(abi.handle_array(outs, abi.TRGSW), abi.handle_array(ins, abi.TLWE), _h(key),
                                    _h(kska_pair), _h(kskb), len(ins))


def trlwe_keyswitch(out, in_, ks_key) -> None:
    """FFT-based TRLWE key switch (keyswitch.c:162-193); out may be in_."""
    lib().trlwe_keyswitch(_h(out), _h(in_), _h(ks_key))


def trlwe_priv_keyswitch_2(out, in_, ks_key_pair) -> None:
    lib().trlwe_priv_keyswitch_2(_h(out), _h(in_), _h(ks_key_pair))


def trlwe_keyswitch_batch(outs, ins, ks_key) -> None:
    lib().trlwe_keyswitch_batch(abi.handle_array(outs, abi.TRLWE), abi.handle_array(ins, abi.TRLWE), _h(ks_key), len(ins))


def trlwe_priv_keyswitch_2_batch(outs, ins, ks_key_pair) -> None:
    lib().trlwe_priv_keyswitch_2_batch(abi.handle_array(outs, abi.TRLWE), abi.handle_array(ins, abi.TRLWE),
                                       _h(ks_key_pair), len(ins))


def release_trlwe_ks_key(key) -> None:
    lib().mb200_release_trlwe_ks_key(_h(key))


def release_trlwe_priv_ks_key(pair) -> None:
    lib().mb200_release_trlwe_priv_ks_key(_h(pair))


def trlwe_fft_ks_dev(row_set, mode: int, d_out, d_in, count, stream=None):
    lib().mb200_trlwe_fft_ks_dev(row_set.handle, mode, _ptr(d_out), _ptr(d_in), count, _ptr(stream))


def circuit_bootstrap_variant_dev(variant, bsk, kska, kska_fft, kskb, d_out_trgsw, d_in, l_out, Bg_bit_out, count, stream=None):
    lib().mb200_circuit_bootstrap_variant_dev(variant, bsk.handle, kska.handle if kska else None,
                                              kska_fft.handle if kska_fft else None, kskb.handle, _ptr(d_out_trgsw),
                                              _ptr(d_in), l_out, Bg_bit_out, count, _ptr(stream))


def release_generic_ks_key(key) -> None:
    lib().mb200_release_generic_ks_key(_h(key))


def multivalue_bootstrap_phase1(out_list, in_, key, torus_base: int) -> None:
    """out_list: torus_base + 1 TRLWEs (bootstrap.c:232-243)."""
    lib().multivalue_bootstrap_phase1(abi.handle_array(out_list, abi.TRLWE), _h(in_), _h(key), torus_base)


def multivalue_bootstrap_phase2(out, lut_ints, rotated_tv, torus_base: int, log_torus_base: int) -> None:
    lut = (C.c_int * len(lut_ints))(*[int(x) for x in lut_ints])
    lib().multivalue_bootstrap_phase2(_h(out), lut, abi.handle_array(rotated_tv, abi.TRLWE), torus_base, log_torus_base)


def register_bootstrap_key(key) -> None:
    lib().mb200_register_bootstrap_key(_h(key))


def release_bootstrap_key(key) -> None:
    lib().mb200_release_bootstrap_key(_h(key))


def register_ks_key(key) -> None:
    lib().mb200_register_ks_key(_h(key))


def release_ks_key(key) -> None:
    lib().mb200_release_ks_key(_h(key))


# ---------------------------------------------------------------------------------------------
# (2) batched variants over lists of handles
# ---------------------------------------------------------------------------------------------
def _tvs(tv):
    tvs = tv if isinstance(tv, (list, tuple)) else [tv]
    return abi.handle_array(tvs, abi.TRLWE), len(tvs)


def functional_bootstrap_batch(out, tv, in_, key, torus_base: int) -> None:
    tva, tvc = _tvs(tv)
    lib().functional_bootstrap_batch(abi.handle_array(out, abi.TLWE), tva, tvc, abi.handle_array(in_, abi.TLWE),
                                     _h(key), torus_base, len(in_))


def functional_bootstrap_wo_extract_batch(out, tv, in_, key, torus_base: int) -> None:
    tva, tvc = _tvs(tv)
    lib().functional_bootstrap_wo_extract_batch(abi.handle_array(out, abi.TRLWE), tva, tvc,
                                                abi.handle_array(in_, abi.TLWE), _h(key), torus_base, len(in_))


def programmable_bootstrap_batch(out, tv, in_, key, precision: int, kappa: int, theta: int) -> None:
    tva, tvc = _tvs(tv)
    lib().programmable_bootstrap_batch(abi.handle_array(out, abi.TLWE), tva, tvc, abi.handle_array(in_, abi.TLWE),
                                       _h(key), precision, kappa, theta, len(in_))


def tlwe_keyswitch_batch(out, in_, ks_key) -> None:
    lib().tlwe_keyswitch_batch(abi.handle_array(out, abi.TLWE), abi.handle_array(in_, abi.TLWE), _h(ks_key), len(in_))


def functional_bootstrap_keyswitch_batch(out, tv, in_, key, ks_key, torus_base: int) -> None:
    tva, tvc = _tvs(tv)
    lib().functional_bootstrap_keyswitch_batch(abi.handle_array(out, abi.TLWE), tva, tvc,
                                               abi.handle_array(in_, abi.TLWE), _h(key), _h(ks_key), torus_base, len(in_))


def blind_rotate_batch(tv_list, a_list, s, size: int) -> None:
    ptrs = (C.POINTER(C.c_uint64) * len(a_list))(*[a.ctypes.data_as(C.POINTER(C.c_uint64)) for a in a_list])
    lib().blind_rotate_batch(abi.handle_array(tv_list, abi.TRLWE), ptrs, s, size, len(tv_list))


def trgsw_mul_trlwe_DFT_batch(out, in1, in2) -> None:
    in2s = in2 if isinstance(in2, (list, tuple)) else [in2]
    lib().trgsw_mul_trlwe_DFT_batch(abi.handle_array(out, abi.TRLWE_DFT), abi.handle_array(in1, abi.TRLWE),
                                    abi.handle_array(in2s, abi.TRGSW_DFT), len(in2s), len(in1))


def trgsw_cmux_batch(out, in1, in2, selector) -> None:
    """out[i] = in1[i] + selector (.) (in2[i] - in1[i])  (vertical_packing.c:24-33)."""
    lib().trgsw_cmux_batch(abi.handle_array(out, abi.TRLWE), abi.handle_array(in1, abi.TRLWE),
                           abi.handle_array(in2, abi.TRLWE), _h(selector), len(in1))


def trlwe_from_DFT_batch(out, in_) -> None:
    lib().trlwe_from_DFT_batch(abi.handle_array(out, abi.TRLWE), abi.handle_array(in_, abi.TRLWE_DFT), len(in_))


def trlwe_extract_tlwe_batch(out, in_, idx) -> None:
    idx = np.ascontiguousarray(idx, np.int32)
    lib().trlwe_extract_tlwe_batch(abi.handle_array(out, abi.TLWE), abi.handle_array(in_, abi.TRLWE),
                                   idx.ctypes.data_as(C.POINTER(C.c_int)), len(idx), len(in_))


# extraction family of the multi-ciphertext caller (trlwe.c:554-620, integer.c:94-100)
def trlwe_extract_tlwe_addto(out, in_, idx: int) -> None:
    lib().trlwe_extract_tlwe_addto(_h(out), _h(in_), idx)


def trlwe_extract_tlwe_subto(out, in_, idx: int) -> None:
    lib().trlwe_extract_tlwe_subto(_h(out), _h(in_), idx)


def trlwe_mv_extract_tlwe(outs, in_, amount: int) -> None:
    lib().trlwe_mv_extract_tlwe(abi.handle_array(outs, abi.TLWE), _h(in_), amount)


def trlwe_mv_extract_tlwe_scaling(out, in_, scale: int, mode: int = 0) -> None:
    """mode 0: trlwe_mv_extract_tlwe_scaling, +1: _scaling_addto, -1: _scaling_subto."""
    fn = {0: "trlwe_mv_extract_tlwe_scaling", 1: "trlwe_mv_extract_tlwe_scaling_addto", -1: "trlwe_mv_extract_tlwe_scaling_subto"}[mode]
    getattr(lib(), fn)(_h(out), _h(in_), scale)


def trlwe_extract_tlwe_acc_batch(outs, ins, idx, mode: int) -> None:
    idx = np.ascontiguousarray(np.atleast_1d(idx), np.int32)
    lib().trlwe_extract_tlwe_acc_batch(abi.handle_array(outs, abi.TLWE), abi.handle_array(ins, abi.TRLWE),
                                       idx.ctypes.data_as(C.POINTER(C.c_int)), len(idx), mode, len(ins))


def trlwe_mv_extract_tlwe_scaling_batch(outs, ins, scale: int, mode: int = 0) -> None:
    lib().trlwe_mv_extract_tlwe_scaling_batch(abi.handle_array(outs, abi.TLWE), abi.handle_array(ins, abi.TRLWE), scale, mode, len(ins))


def trlwe_mv_extract_tlwe_batch(out_lists, ins, amount: int) -> None:
    arrs = [abi.handle_array(o, abi.TLWE) for o in out_lists]
    ptrs = (C.POINTER(abi.TLWE) * len(arrs))(*[C.cast(a, C.POINTER(abi.TLWE)) for a in arrs])
    lib().trlwe_mv_extract_tlwe_batch(ptrs, abi.handle_array(ins, abi.TRLWE), amount, len(ins))


def tlwe_keyswitch_bootstrap_mv_extract_batch(digits, carries, tv, ks_key, key, torus_base: int, scale_digit: int,
                                              scale_carry: int = 1) -> None:
    """One digit step of integer.c:94-100 for a batch: digits[i] -= mv_scaling(PBS(KS(digits[i])), scale_digit) and
    (carries given) carries[i] += mv_scaling(.., scale_carry)."""
    tva, tvc = _tvs(tv)
    lib().tlwe_keyswitch_bootstrap_mv_extract_batch(abi.handle_array(digits, abi.TLWE),
                                                    abi.handle_array(carries, abi.TLWE) if carries is not None else None,
                                                    tva, tvc, _h(ks_key), _h(key), torus_base, scale_digit, scale_carry, len(digits))


# ---------------------------------------------------------------------------------------------
# (3) flat calls: resident keys, host-buffer and device-resident batches
# ---------------------------------------------------------------------------------------------
def _ptr(x):
    """Raw address of a numpy array, a torch tensor, or an int."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    return int(x)


class BootstrapKey:
    """Fourier-domain bootstrapping key resident in HBM (``mb200_bsk_t``)."""

    def __init__(self, params: Params, handle: int, keepalive=None):
        self.params, self.handle, self._keep = params, handle, keepalive

    @classmethod
    def from_host(cls, params: Params, bsk_host: np.ndarray, layout: int) -> "BootstrapKey":
        """``bsk_host``: float64 [n, (k+1)l, k+1, N] in the slot order ``layout`` of the CPU build."""
        bsk_host = np.ascontiguousarray(bsk_host, np.float64)
        assert bsk_host.shape == (params.n, (params.k + 1) * params.l, params.k + 1, params.N), bsk_host.shape
        p = params.c()
        h = lib().mb200_bsk_from_host(C.byref(p), bsk_host.ctypes.data_as(C.POINTER(C.c_double)), layout)
        return cls(params, h)

    @classmethod
    def synthesize(cls, params: Params, lwe_key: np.ndarray, rlwe_key: np.ndarray, seed: int = 1) -> "BootstrapKey":
        lwe_key = np.ascontiguousarray(lwe_key, np.uint64)
        rlwe_key = np.ascontiguousarray(rlwe_key, np.uint64).reshape(-1)
        assert lwe_key.shape[0] == params.n and rlwe_key.shape[0] == params.k * params.N
        p = params.c()
        u = C.POINTER(C.c_uint64)
        h = lib().mb200_bsk_synthesize(C.byref(p), lwe_key.ctypes.data_as(u), rlwe_key.ctypes.data_as(u),
                                       params.rlwe_sigma, seed)
        return cls(params, h)

    @classmethod
    def from_torus_dev(cls, params: Params, d_trgsw, stream=None) -> "BootstrapKey":
        """trgsw_to_DFT (trgsw.c:345) of ``params.n`` torus-domain TRGSW samples held in device memory."""
        p = params.c()
        return cls(params, lib().mb200_bsk_from_torus_dev(C.byref(p), _ptr(d_trgsw), _ptr(stream)))

    @classmethod
    def adopt(cls, params: Params, device_buffer) -> "BootstrapKey":
        """Wrap caller-owned device memory already in the resident layout (e.g. an NCCL receive buffer)."""
        p = params.c()
        h = lib().mb200_bsk_adopt_device(C.byref(p), _ptr(device_buffer))
        return cls(params, h, keepalive=device_buffer)

    @property
    def device_ptr(self) -> int:
        return int(lib().mb200_bsk_device_ptr(self.handle))

    @property
    def nbytes(self) -> int:
        p = self.params.c()
        return int(lib().mb200_bsk_device_bytes(C.byref(p)))

    def free(self) -> None:
        if self.handle:
            lib().mb200_bsk_free(self.handle)
            self.handle = None


class KeySwitchKey:
    """TLWE key-switching table resident in HBM (``mb200_ksk_t``)."""

    def __init__(self, params: Params, handle: int, keepalive=None):
        self.params, self.handle, self._keep = params, handle, keepalive

    @classmethod
    def from_host(cls, params: Params, ksk: np.ndarray) -> "KeySwitchKey":
        ksk = np.ascontiguousarray(ksk, np.uint64)
        assert ksk.shape == (params.k * params.N, params.t, (1 << params.base_bit) - 1, params.n + 1), ksk.shape
        p = params.c()
        h = lib().mb200_ksk_from_host(C.byref(p), ksk.ctypes.data_as(C.POINTER(C.c_uint64)))
        return cls(params, h)

    @classmethod
    def synthesize(cls, params: Params, rlwe_key: np.ndarray, lwe_key: np.ndarray, seed: int = 2) -> "KeySwitchKey":
        rlwe_key = np.ascontiguousarray(rlwe_key, np.uint64).reshape(-1)
        lwe_key = np.ascontiguousarray(lwe_key, np.uint64)
        p = params.c()
        u = C.POINTER(C.c_uint64)
        h = lib().mb200_ksk_synthesize(C.byref(p), rlwe_key.ctypes.data_as(u), lwe_key.ctypes.data_as(u),
                                       params.lwe_sigma, seed)
        return cls(params, h)

    @classmethod
    def adopt(cls, params: Params, device_buffer) -> "KeySwitchKey":
        p = params.c()
        h = lib().mb200_ksk_adopt_device(C.byref(p), _ptr(device_buffer))
        return cls(params, h, keepalive=device_buffer)

    @property
    def device_ptr(self) -> int:
        return int(lib().mb200_ksk_device_ptr(self.handle))

    @property
    def nbytes(self) -> int:
        p = self.params.c()
        return int(lib().mb200_ksk_device_bytes(C.byref(p)))

    def free(self) -> None:
        if self.handle:
            lib().mb200_ksk_free(self.handle)
            self.handle = None


class GenericKSKey:
    """TRLWE-row key-switching key resident in HBM (``mb200_gksk_t``; k = 1)."""

    def __init__(self, handle, n_in, include_b, N, t, base_bit):
        self.handle, self.n_in, self.include_b, self.N, self.t, self.base_bit = handle, n_in, include_b, N, t, base_bit

    @classmethod
    def from_host(cls, rows: np.ndarray, include_b: int, base_bit: int) -> "GenericKSKey":
        rows = np.ascontiguousarray(rows, np.uint64)
        ne, t, _, kp1, N = rows.shape
        assert kp1 == 2
        h = lib().mb200_gksk_from_host(rows.ctypes.data_as(C.POINTER(C.c_uint64)), ne - include_b, include_b, N, t, base_bit)
        return cls(h, ne - include_b, include_b, N, t, base_bit)

    @classmethod
    def synthesize(cls, in_key, out_rlwe_key, include_b, t, base_bit, sigma, seed=5) -> "GenericKSKey":
        in_key = np.ascontiguousarray(in_key, np.uint64)
        out_rlwe_key = np.ascontiguousarray(out_rlwe_key, np.uint64).reshape(-1)
        u = C.POINTER(C.c_uint64)
        h = lib().mb200_gksk_synthesize(in_key.ctypes.data_as(u), out_rlwe_key.ctypes.data_as(u), in_key.shape[0],
                                        include_b, out_rlwe_key.shape[0], t, base_bit, sigma, seed)
        return cls(h, in_key.shape[0], include_b, out_rlwe_key.shape[0], t, base_bit)

    @property
    def device_ptr(self) -> int:
        return int(lib().mb200_gksk_device_ptr(self.handle))

    @property
    def shape(self):
        return (self.n_in + self.include_b, self.t, (1 << self.base_bit) - 1, 2, self.N)

    def free(self) -> None:
        if self.handle:
            lib().mb200_gksk_free(self.handle)
            self.handle = None


def trlwe_ks_dev(ksk: GenericKSKey, d_out, d_in, count, stream=None):
    lib().mb200_trlwe_ks_dev(ksk.handle, _ptr(d_out), _ptr(d_in), count, _ptr(stream))


def circuit_bootstrap_dev(bsk, kska, kskb, d_out_trgsw, d_in, Bg_bit_out, count, stream=None):
    lib().mb200_circuit_bootstrap_dev(bsk.handle, kska.handle, kskb.handle, _ptr(d_out_trgsw), _ptr(d_in), Bg_bit_out,
                                      count, _ptr(stream))


# ---- host-buffer batches (numpy or pinned torch tensors in, numpy out) ------------------------
def pbs_ks_host(bsk: BootstrapKey, ksk: KeySwitchKey, tv, tlwe_in, torus_base: int, out=None):
    """functional_bootstrap then tlwe_keyswitch for every row of ``tlwe_in`` [count, n+1]."""
    P = bsk.params
    count = tlwe_in.shape[0]
    tv_count = 1 if tv.ndim == 2 else tv.shape[0]
    if out is None:
        out = np.empty((count, P.n + 1), np.uint64)
    lib().mb200_pbs_ks_host(bsk.handle, ksk.handle, _ptr(out), _ptr(tv), tv_count, _ptr(tlwe_in), torus_base, count)
    return out


def pbs_host(bsk: BootstrapKey, tv, tlwe_in, torus_base: int, out=None):
    P = bsk.params
    count = tlwe_in.shape[0]
    tv_count = 1 if tv.ndim == 2 else tv.shape[0]
    if out is None:
        out = np.empty((count, P.k * P.N + 1), np.uint64)
    lib().mb200_pbs_host(bsk.handle, _ptr(out), _ptr(tv), tv_count, _ptr(tlwe_in), torus_base, count)
    return out


def ks_host(ksk: KeySwitchKey, tlwe_in, out=None):
    P = ksk.params
    count = tlwe_in.shape[0]
    if out is None:
        out = np.empty((count, P.n + 1), np.uint64)
    lib().mb200_ks_host(ksk.handle, _ptr(out), _ptr(tlwe_in), count)
    return out


# ---- device-resident batches (raw device pointers / torch CUDA tensors; async on `stream`) ----
def pbs_dev(bsk, d_out, d_tv, tv_count, d_in, torus_base, count, stream=None):
    lib().mb200_pbs_dev(bsk.handle, _ptr(d_out), _ptr(d_tv), tv_count, _ptr(d_in), torus_base, count, _ptr(stream))


def pbs_wo_extract_dev(bsk, d_out, d_tv, tv_count, d_in, torus_base, count, stream=None):
    lib().mb200_pbs_wo_extract_dev(bsk.handle, _ptr(d_out), _ptr(d_tv), tv_count, _ptr(d_in), torus_base, count,
                                   _ptr(stream))


def blind_rotate_dev(bsk, d_acc, d_a, a_stride, size, count, stream=None):
    lib().mb200_blind_rotate_dev(bsk.handle, _ptr(d_acc), _ptr(d_a), a_stride, size, count, _ptr(stream))


def extract_dev(d_out, d_trlwe, idx, N, k, count, stream=None):
    idx = np.ascontiguousarray(idx, np.int32)
    lib().mb200_extract_dev(_ptr(d_out), _ptr(d_trlwe), idx.ctypes.data_as(C.POINTER(C.c_int)), len(idx), N, k, count,
                            _ptr(stream))


def ks_dev(ksk, d_out, d_in, count, stream=None):
    lib().mb200_ks_dev(ksk.handle, _ptr(d_out), _ptr(d_in), count, _ptr(stream))


def pbs_ks_dev(bsk, ksk, d_out, d_tv, tv_count, d_in, d_scratch, torus_base, count, stream=None):
    lib().mb200_pbs_ks_dev(bsk.handle, ksk.handle, _ptr(d_out), _ptr(d_tv), tv_count, _ptr(d_in), _ptr(d_scratch),
                           torus_base, count, _ptr(stream))


def extprod_dev(trgsw_set, sel, d_out, d_in, count, stream=None):
    sel = np.ascontiguousarray(sel, np.int32)
    lib().mb200_extprod_dev(trgsw_set.handle, sel.ctypes.data_as(C.POINTER(C.c_int)), _ptr(d_out), _ptr(d_in), count,
                            _ptr(stream))


def cmux_dev(trgsw_set, sel: int, d_out, d_in1, d_in2, count, stream=None):
    lib().mb200_cmux_dev(trgsw_set.handle, sel, _ptr(d_out), _ptr(d_in1), _ptr(d_in2), count, _ptr(stream))


def vertical_packing_batch_dev(bits, d_luts, d_out_tlwe, size: int, E: int, stream=None):
    """E evaluations in lockstep; bits: E*size TRGSWs (evaluation-major), d_luts LUT-major [n_luts][E][2N] (consumed)."""
    lib().mb200_vertical_packing_batch_dev(bits.handle, _ptr(d_luts), _ptr(d_out_tlwe), size, E, _ptr(stream))


def vertical_packing_dev(bits, d_luts, d_out_tlwe, size: int, stream=None):
    """CGGI vertical packing over resident TRGSW bit encryptions; consumes d_luts (vertical_packing.c:36-52)."""
    lib().mb200_vertical_packing_dev(bits.handle, _ptr(d_luts), _ptr(d_out_tlwe), size, _ptr(stream))


def torus_to_dft_dev(d_out, d_in, N, count, stream=None):
    lib().mb200_torus_to_dft_dev(_ptr(d_out), _ptr(d_in), N, count, _ptr(stream))


def dft_to_torus_dev(d_out, d_in, N, count, stream=None):
    lib().mb200_dft_to_torus_dev(_ptr(d_out), _ptr(d_in), N, count, _ptr(stream))
