"""Loads ``libmosfhet_b200.so`` (the CUDA C-ABI library) and declares its prototypes.

There is deliberately no fallback: if the shared object is missing, or a compute entry point is
called without a CUDA device, the call fails loudly (ImportError here, abort() in the library).
"""
from __future__ import annotations

import ctypes as C
import os

from . import abi

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, os.environ.get("MB200_LIB_NAME", "libmosfhet_b200.so"))

_P = C.POINTER
_params_p = _P(abi.ParamsS)
_u64p = _P(C.c_uint64)
_vp = C.c_void_p

# every symbol include/mosfhet_b200.h declares: name -> (restype, argtypes)
PROTOTYPES = {
    # (1) drop-in names
    "functional_bootstrap": (None, [abi.TLWE, abi.TRLWE, abi.TLWE, abi.Bootstrap_Key, C.c_int]),
    "functional_bootstrap_wo_extract": (None, [abi.TRLWE, abi.TRLWE, abi.TLWE, abi.Bootstrap_Key, C.c_int]),
    "programmable_bootstrap": (None, [abi.TLWE, abi.TRLWE, abi.TLWE, abi.Bootstrap_Key, C.c_int, C.c_int, C.c_int]),
    "blind_rotate": (None, [abi.TRLWE, _u64p, _P(abi.TRGSW_DFT), C.c_int]),
    "trgsw_mul_trlwe_DFT": (None, [abi.TRLWE_DFT, abi.TRLWE, abi.TRGSW_DFT]),
    "trlwe_from_DFT": (None, [abi.TRLWE, abi.TRLWE_DFT]),
    "trlwe_extract_tlwe": (None, [abi.TLWE, abi.TRLWE, C.c_int]),
    "tlwe_keyswitch": (None, [abi.TLWE, abi.TLWE, abi.TLWE_KS_Key]),
    "multivalue_bootstrap_CLOT21": (None, [_P(abi.TLWE), abi.TRLWE, abi.TLWE, abi.Bootstrap_Key, C.c_int, C.c_int]),
    "trlwe_packing1_keyswitch": (None, [abi.TRLWE, abi.TLWE, abi.Generic_KS_Key]),
    "trlwe_priv_keyswitch": (None, [abi.TRLWE, abi.TLWE, abi.Generic_KS_Key]),
    "circuit_bootstrap_2": (None, [abi.TRGSW, abi.TLWE, abi.Bootstrap_Key, abi.Generic_KS_Key, abi.Generic_KS_Key]),
    "functional_bootstrap_trgsw_phase1": (None, [abi.TRGSW_DFT, abi.TLWE, abi.Bootstrap_Key, C.c_int]),
    "functional_bootstrap_trgsw_phase2": (None, [abi.TLWE, abi.TRGSW_DFT, abi.TRLWE]),
    "blind_rotate_unfolded": (None, [abi.TRLWE, _u64p, _P(abi.TRGSW), C.c_int, C.c_int]),
    "multivalue_bootstrap_UBR_phase1": (None, [_P(abi.TRGSW_DFT), abi.TLWE, abi.Bootstrap_Key]),
    "multivalue_bootstrap_UBR_phase2": (None, [abi.TLWE, abi.TRLWE, abi.TLWE, _P(abi.TRGSW_DFT), abi.Bootstrap_Key, C.c_int]),
    "functional_bootstrap_trgsw_phase1_batch": (None, [_P(abi.TRGSW_DFT), _P(abi.TLWE), abi.Bootstrap_Key, C.c_int, C.c_int]),
    "functional_bootstrap_trgsw_phase2_batch": (None, [_P(abi.TLWE), _P(abi.TRGSW_DFT), _P(abi.TRLWE), C.c_int, C.c_int]),
    "blind_rotate_unfolded_batch": (None, [_P(abi.TRLWE), _P(_u64p), _P(abi.TRGSW), C.c_int, C.c_int, C.c_int]),
    "mb200_bootstrap_trgsw_phase1_dev": (None, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "circuit_bootstrap": (None, [abi.TRGSW, abi.TLWE, abi.Bootstrap_Key, abi.Generic_KS_Key, abi.Generic_KS_Key]),
    "circuit_bootstrap_3": (None, [abi.TRGSW, abi.TLWE, abi.Bootstrap_Key, _P(abi.TRLWE_KS_Key), abi.Generic_KS_Key]),
    "trlwe_keyswitch": (None, [abi.TRLWE, abi.TRLWE, abi.TRLWE_KS_Key]),
    "trlwe_priv_keyswitch_2": (None, [abi.TRLWE, abi.TRLWE, _P(abi.TRLWE_KS_Key)]),
    "circuit_bootstrap_batch": (None, [_P(abi.TRGSW), _P(abi.TLWE), abi.Bootstrap_Key, abi.Generic_KS_Key, abi.Generic_KS_Key, C.c_int]),
    "circuit_bootstrap_3_batch": (None, [_P(abi.TRGSW), _P(abi.TLWE), abi.Bootstrap_Key, _P(abi.TRLWE_KS_Key), abi.Generic_KS_Key, C.c_int]),
    "trlwe_keyswitch_batch": (None, [_P(abi.TRLWE), _P(abi.TRLWE), abi.TRLWE_KS_Key, C.c_int]),
    "trlwe_priv_keyswitch_2_batch": (None, [_P(abi.TRLWE), _P(abi.TRLWE), _P(abi.TRLWE_KS_Key), C.c_int]),
    "mb200_register_trlwe_ks_key": (None, [abi.TRLWE_KS_Key]),
    "mb200_release_trlwe_ks_key": (None, [abi.TRLWE_KS_Key]),
    "mb200_register_trlwe_priv_ks_key": (None, [_P(abi.TRLWE_KS_Key)]),
    "mb200_release_trlwe_priv_ks_key": (None, [_P(abi.TRLWE_KS_Key)]),
    "mb200_trlwe_fft_ks_dev": (None, [_vp, C.c_int, _vp, _vp, C.c_int, _vp]),
    "mb200_circuit_bootstrap_variant_dev": (None, [C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp]),
    "trlwe_packing1_keyswitch_batch": (None, [_P(abi.TRLWE), _P(abi.TLWE), abi.Generic_KS_Key, C.c_int]),
    "trlwe_priv_keyswitch_batch": (None, [_P(abi.TRLWE), _P(abi.TLWE), abi.Generic_KS_Key, C.c_int]),
    "circuit_bootstrap_2_batch": (None, [_P(abi.TRGSW), _P(abi.TLWE), abi.Bootstrap_Key, abi.Generic_KS_Key, abi.Generic_KS_Key, C.c_int]),
    "mb200_register_generic_ks_key": (None, [abi.Generic_KS_Key]),
    "mb200_release_generic_ks_key": (None, [abi.Generic_KS_Key]),
    "multivalue_bootstrap_phase1": (None, [_P(abi.TRLWE), abi.TLWE, abi.Bootstrap_Key, C.c_int]),
    "multivalue_bootstrap_phase2": (None, [abi.TLWE, _P(C.c_int), _P(abi.TRLWE), C.c_int, C.c_int]),
    # (2) batched
    "multivalue_bootstrap_phase1_batch": (None, [_P(_P(abi.TRLWE)), _P(abi.TLWE), abi.Bootstrap_Key, C.c_int, C.c_int]),
    "multivalue_bootstrap_phase2_batch": (None, [_P(abi.TLWE), _P(_P(C.c_int)), C.c_int, _P(_P(abi.TRLWE)), C.c_int, C.c_int, C.c_int]),
    "functional_bootstrap_batch": (None, [_P(abi.TLWE), _P(abi.TRLWE), C.c_int, _P(abi.TLWE), abi.Bootstrap_Key, C.c_int, C.c_int]),
    "functional_bootstrap_wo_extract_batch": (None, [_P(abi.TRLWE), _P(abi.TRLWE), C.c_int, _P(abi.TLWE), abi.Bootstrap_Key, C.c_int, C.c_int]),
    "programmable_bootstrap_batch": (None, [_P(abi.TLWE), _P(abi.TRLWE), C.c_int, _P(abi.TLWE), abi.Bootstrap_Key, C.c_int, C.c_int, C.c_int, C.c_int]),
    "blind_rotate_batch": (None, [_P(abi.TRLWE), _P(_u64p), _P(abi.TRGSW_DFT), C.c_int, C.c_int]),
    "trgsw_mul_trlwe_DFT_batch": (None, [_P(abi.TRLWE_DFT), _P(abi.TRLWE), _P(abi.TRGSW_DFT), C.c_int, C.c_int]),
    "trgsw_cmux_batch": (None, [_P(abi.TRLWE), _P(abi.TRLWE), _P(abi.TRLWE), abi.TRGSW_DFT, C.c_int]),
    "trlwe_from_DFT_batch": (None, [_P(abi.TRLWE), _P(abi.TRLWE_DFT), C.c_int]),
    "trlwe_extract_tlwe_batch": (None, [_P(abi.TLWE), _P(abi.TRLWE), _P(C.c_int), C.c_int, C.c_int]),
    "tlwe_keyswitch_batch": (None, [_P(abi.TLWE), _P(abi.TLWE), abi.TLWE_KS_Key, C.c_int]),
    "trlwe_extract_tlwe_addto": (None, [abi.TLWE, abi.TRLWE, C.c_int]),
    "trlwe_extract_tlwe_subto": (None, [abi.TLWE, abi.TRLWE, C.c_int]),
    "trlwe_mv_extract_tlwe": (None, [_P(abi.TLWE), abi.TRLWE, C.c_int]),
    "trlwe_mv_extract_tlwe_scaling": (None, [abi.TLWE, abi.TRLWE, C.c_int]),
    "trlwe_mv_extract_tlwe_scaling_addto": (None, [abi.TLWE, abi.TRLWE, C.c_int]),
    "trlwe_mv_extract_tlwe_scaling_subto": (None, [abi.TLWE, abi.TRLWE, C.c_int]),
    "trlwe_extract_tlwe_acc_batch": (None, [_P(abi.TLWE), _P(abi.TRLWE), _P(C.c_int), C.c_int, C.c_int, C.c_int]),
    "trlwe_mv_extract_tlwe_batch": (None, [_P(_P(abi.TLWE)), _P(abi.TRLWE), C.c_int, C.c_int]),
    "trlwe_mv_extract_tlwe_scaling_batch": (None, [_P(abi.TLWE), _P(abi.TRLWE), C.c_int, C.c_int, C.c_int]),
    "tlwe_keyswitch_bootstrap_mv_extract_batch": (None, [_P(abi.TLWE), _P(abi.TLWE), _P(abi.TRLWE), C.c_int, abi.TLWE_KS_Key,
                                                         abi.Bootstrap_Key, C.c_int, C.c_int, C.c_int, C.c_int]),
    "mb200_init_multi": (C.c_int, [C.c_int]),
    "mb200_multi_device_count": (C.c_int, []),
    "free_bootstrap_key": (None, [abi.Bootstrap_Key]),
    "free_tlwe_ks_key": (None, [abi.TLWE_KS_Key]),
    "free_trlwe_generic_ks_key": (None, [C.c_void_p]),
    "free_trlwe_ks_key": (None, [C.c_void_p]),
    "functional_bootstrap_keyswitch_batch": (None, [_P(abi.TLWE), _P(abi.TRLWE), C.c_int, _P(abi.TLWE), abi.Bootstrap_Key, abi.TLWE_KS_Key, C.c_int, C.c_int]),
    "multivalue_bootstrap_CLOT21_batch": (None, [_P(_P(abi.TLWE)), _P(abi.TRLWE), C.c_int, _P(abi.TLWE), abi.Bootstrap_Key, C.c_int, C.c_int, C.c_int]),
    # runtime / keys
    "mb200_init": (C.c_int, [C.c_int]),
    "mb200_shutdown": (None, []),
    "mb200_device_count": (C.c_int, []),
    "mb200_version": (C.c_char_p, []),
    "mb200_device_synchronize": (None, []),
    "mb200_set_host_fft_layout": (None, [C.c_int]),
    "mb200_get_host_fft_layout": (C.c_int, []),
    "mb200_host_slot_exponents": (None, [C.c_int, C.c_int, _P(C.c_int32)]),
    "mb200_register_bootstrap_key": (None, [abi.Bootstrap_Key]),
    "mb200_release_bootstrap_key": (None, [abi.Bootstrap_Key]),
    "mb200_register_ks_key": (None, [abi.TLWE_KS_Key]),
    "mb200_release_ks_key": (None, [abi.TLWE_KS_Key]),
    # (3) flat
    "mb200_bsk_device_bytes": (C.c_size_t, [_params_p]),
    "mb200_ksk_device_bytes": (C.c_size_t, [_params_p]),
    "mb200_bsk_from_host": (_vp, [_params_p, _P(C.c_double), C.c_int]),
    "mb200_ksk_from_host": (_vp, [_params_p, _u64p]),
    "mb200_bsk_adopt_device": (_vp, [_params_p, _vp]),
    "mb200_ksk_adopt_device": (_vp, [_params_p, _vp]),
    "mb200_bsk_device_ptr": (_vp, [_vp]),
    "mb200_ksk_device_ptr": (_vp, [_vp]),
    "mb200_bsk_free": (None, [_vp]),
    "mb200_ksk_free": (None, [_vp]),
    "mb200_bsk_synthesize": (_vp, [_params_p, _u64p, _u64p, C.c_double, C.c_uint64]),
    "mb200_ksk_synthesize": (_vp, [_params_p, _u64p, _u64p, C.c_double, C.c_uint64]),
    "mb200_bsk_from_torus_dev": (_vp, [_params_p, _vp, _vp]),
    "mb200_gksk_from_host": (_vp, [_u64p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "mb200_gksk_synthesize": (_vp, [_u64p, _u64p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_uint64]),
    "mb200_gksk_device_ptr": (_vp, [_vp]),
    "mb200_gksk_free": (None, [_vp]),
    "mb200_trlwe_ks_dev": (None, [_vp, _vp, _vp, C.c_int, _vp]),
    "mb200_circuit_bootstrap_dev": (None, [_vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp]),
    "mb200_pbs_dev": (None, [_vp, _vp, _vp, C.c_int, _vp, C.c_int, C.c_int, _vp]),
    "mb200_pbs_wo_extract_dev": (None, [_vp, _vp, _vp, C.c_int, _vp, C.c_int, C.c_int, _vp]),
    "mb200_blind_rotate_dev": (None, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp]),
    "mb200_extract_dev": (None, [_vp, _vp, _P(C.c_int), C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "mb200_ks_dev": (None, [_vp, _vp, _vp, C.c_int, _vp]),
    "mb200_pbs_ks_dev": (None, [_vp, _vp, _vp, _vp, C.c_int, _vp, _vp, C.c_int, C.c_int, _vp]),
    "mb200_extprod_dev": (None, [_vp, _P(C.c_int), _vp, _vp, C.c_int, _vp]),
    "mb200_cmux_dev": (None, [_vp, C.c_int, _vp, _vp, _vp, C.c_int, _vp]),
    "mb200_vertical_packing_dev": (None, [_vp, _vp, _vp, C.c_int, _vp]),
    "mb200_vertical_packing_batch_dev": (None, [_vp, _vp, _vp, C.c_int, C.c_int, _vp]),
    "mb200_torus_to_dft_dev": (None, [_vp, _vp, C.c_int, C.c_int, _vp]),
    "mb200_dft_to_torus_dev": (None, [_vp, _vp, C.c_int, C.c_int, _vp]),
    "mb200_pbs_ks_host": (None, [_vp, _vp, _vp, _vp, C.c_int, _vp, C.c_int, C.c_int]),
    "mb200_pbs_host": (None, [_vp, _vp, _vp, C.c_int, _vp, C.c_int, C.c_int]),
    "mb200_ks_host": (None, [_vp, _vp, _vp, C.c_int]),
    "mb200_launch_count": (C.c_uint64, []),
    "mb200_reset_launch_count": (None, []),
    "mb200_last_blind_rotate_kernel": (C.c_char_p, []),
    "mb200_set_kernel_policy": (None, [C.c_int]),
    "mb200_measure_fp64_tflops": (C.c_double, [C.c_int]),
}

_lib = None


def load():
    """The loaded library (RTLD_LOCAL, so it can coexist with the reference CPU library in tests)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m mosfhet_b200.build` "
                "(or __graft_entry__.build()); there is no CPU fallback")
        lib = C.CDLL(LIB_PATH, mode=os.RTLD_LOCAL)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)          # AttributeError if the .so does not export it
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib
