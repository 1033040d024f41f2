"""Extra measurements that ride in bench.py's JSON line (`extra`), each bounded to a few seconds on one B200:

* ``level2_batch1``  BASELINE configs[0]: ONE functional_bootstrap + tlwe_keyswitch at the default (TFHEpp Level 2)
  parameters through the drop-in handle entry points -- host TLWE in, host TLWE out, copies inside the timed region --
  on keys made by the unmodified reference (oracle/_ref), the output checked against the reference's own result.
* ``next``           BASELINE configs[3] / [4]: multi-value bootstrap and circuit bootstrap at batch 8192, and the leveled
  LUT (vertical packing) over 2048 evaluations, device resident with device-made keys, every LUT output decrypted.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def level2_batch1(api, reps: int = 12):
    """configs[0] through the reference-facing calls.  Needs oracle/_ref (checker side); returns a dict."""
    from mosfhet_b200.params import LEVEL2 as P
    from oracle import parity, ref as reflib
    if not reflib.available():
        return {"unavailable": "oracle/_ref not built"}
    S = parity.ReferenceSetup(P, 8)
    R = S.R
    api.set_host_fft_layout(R.layout)
    api.register_bootstrap_key(S.bk)
    api.register_ks_key(S.ksk)
    mid = R.tlwe_alloc_sample(P.k * P.N)
    out = R.tlwe_alloc_sample(P.n)
    ts_ref, ts = [], []
    want = []
    for i in range(len(S.inputs)):                              # the reference itself, one host thread (benchmark.c:262-265)
        t0 = time.perf_counter()
        R.functional_bootstrap(mid, S.tv, S.inputs[i], S.bk, S.torus_base)
        R.tlwe_keyswitch(out, mid, S.ksk)
        ts_ref.append(time.perf_counter() - t0)
        want.append(int(S.decode(S.phases([out], S.key_tlwe))[0]))
    got = []
    for r in range(reps):
        i = r % len(S.inputs)
        t0 = time.perf_counter()
        api.functional_bootstrap(mid, S.tv, S.inputs[i], S.bk, S.torus_base)
        api.tlwe_keyswitch(out, mid, S.ksk)
        ts.append(time.perf_counter() - t0)
        got.append((i, int(S.decode(S.phases([out], S.key_tlwe))[0])))
    kern = api.last_blind_rotate_kernel()
    api.release_bootstrap_key(S.bk)
    api.release_ks_key(S.ksk)
    ok = all(want[i] == g for i, g in got)
    return {"config": "configs[0]: functional_bootstrap + tlwe_keyswitch, Level 2 (n=632, N=2048, l=4, Bg_bit=9, t=8, base_bit=4), batch 1",
            "path": "drop-in handle entry points, host TLWE in / out (H2D + kernels + D2H + host scatter inside the timed region)",
            "ms_per_bootstrap_keyswitch": float(np.median(ts[2:]) * 1e3), "ms_min": float(min(ts[2:]) * 1e3),
            "reference_cpu_ms": float(np.median(ts_ref[1:]) * 1e3), "reference_variant": R.variant, "reference_threads": 1,
            "kernel": kern, "outputs_match_reference": bool(ok), "checked": len(got)}


def next_configs(api, batch: int = 8192):
    """configs[3] / [4], device resident (keys synthesised on the device, CUDA events on the launching stream)."""
    import torch
    from mosfhet_b200 import synthetic as syn
    from mosfhet_b200.params import LEVEL2, Params

    st = torch.cuda.Stream()
    res = {}

    def dev_time(fn, reps=2):
        ts = []
        for _ in range(reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); fn(); e1.record(st); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        return min(ts[1:])

    def i64(a):
        return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()

    # config 3a: multivalue_bootstrap_CLOT21 (8 LUTs per input, bootstrap.c:222-230) at the default parameters
    P = LEVEL2
    lwe_k, rlwe_k = syn.binary_key(P.n, 1), syn.binary_key(P.N, 2)
    bsk = api.BootstrapKey.synthesize(P, lwe_k, rlwe_k, seed=3)
    msgs = np.arange(batch) % 2
    d_in = i64(syn.tlwe_encrypt(syn.encode(msgs, 2), lwe_k, P.lwe_sigma, seed=4))
    lut16 = np.arange(16, dtype=np.uint64) << np.uint64(59)
    d_tv = i64(syn.test_vector(lut16, P.N, 1))
    d_acc = torch.empty((batch, 2, P.N), dtype=torch.int64, device="cuda")
    d_mv = torch.empty((batch, 8, P.N + 1), dtype=torch.int64, device="cuda")
    idx = np.arange(8, dtype=np.int32) * (P.N // 16)

    def mv():
        api.pbs_wo_extract_dev(bsk, d_acc, d_tv, 1, d_in, 16, batch, st.cuda_stream)
        api.extract_dev(d_mv, d_acc, idx, P.N, 1, batch, st.cuda_stream)
    ms = dev_time(mv, 1)
    got = d_mv[:64].cpu().numpy().view(np.uint64)
    # LUT j of input m sits at slot m*8 + j of the 16-slot test vector
    ok = True
    for j in range(8):
        ph = syn.tlwe_phase(got[:, j], rlwe_k)
        ok &= bool((syn.torus_distance(ph, lut16[(msgs[:64] * 8 + j) % 16]) <= (1 << 58)).all())
    res["multivalue_bootstrap_CLOT21"] = {"value": batch / ms * 1e3, "unit": "multi-value bootstraps/s", "batch": batch,
                                          "luts_per_input": 8, "ms_per_batch": ms, "params": "level2", "outputs_ok": ok,
                                          "kernel": api.last_blind_rotate_kernel()}
    bsk.free()
    del d_acc, d_mv

    # config 3b: circuit_bootstrap_2 (bootstrap.c:325-347) on the N = 1024 ring with the Level-2 gadget
    Pc = Params(632, 1024, 1, 4, 9, 6, 4, 2.0 ** -30, 2.0 ** -55)
    lwe_k, rlwe_k = syn.binary_key(Pc.n, 5), syn.binary_key(Pc.N, 6)
    bskc = api.BootstrapKey.synthesize(Pc, lwe_k, rlwe_k, seed=7)
    ka = api.GenericKSKey.synthesize(rlwe_k, rlwe_k, 1, 6, 4, 2.0 ** -55, seed=8)
    kb = api.GenericKSKey.synthesize(rlwe_k, rlwe_k, 0, 6, 4, 2.0 ** -55, seed=9)
    d_in = i64(syn.tlwe_encrypt((np.arange(batch, dtype=np.uint64) % 2) << np.uint64(62), lwe_k, Pc.lwe_sigma, seed=10))
    d_g = torch.empty((batch, 8, 2, Pc.N), dtype=torch.int64, device="cuda")
    ms = dev_time(lambda: api.circuit_bootstrap_dev(bskc, ka, kb, d_g, d_in, 9, batch, st.cuda_stream), 1)
    res["circuit_bootstrap_2"] = {"value": batch / ms * 1e3, "unit": "circuit bootstraps/s", "batch": batch, "ms_per_batch": ms,
                                  "params": "n=632 N=1024 l=4 Bg_bit=9 t=6 base_bit=4", "kernel": api.last_blind_rotate_kernel()}
    for h in (bskc, ka, kb):
        h.free()
    del d_g

    # config 4: leveled LUT = vertical packing (applications/leveled_lut/vertical_packing.c:36-52), 13 input bits
    E, Nl, ll, Bgl, size = 2048, 1024, 2, 10, 13
    bits = np.random.default_rng(1).integers(0, 2, size=E * size).astype(np.uint64)
    Pl = Params(E * size, Nl, 1, ll, Bgl, 3, 2, 2.0 ** -30, 2.0 ** -55)
    rk = syn.binary_key(Nl, 11)
    tb = api.BootstrapKey.synthesize(Pl, bits, rk, seed=12)
    lut = np.random.default_rng(2).integers(0, 1 << 10, size=(8, Nl), dtype=np.uint64)
    luts = np.zeros((8, 2, Nl), np.uint64)
    luts[:, 1, :] = lut << np.uint64(54)
    d_l0 = i64(np.repeat(luts[:, None], E, axis=1))
    d_l = torch.empty_like(d_l0)
    d_r = torch.empty((E, Nl + 1), dtype=torch.int64, device="cuda")

    def vp():
        d_l.copy_(d_l0)                                         # the tree consumes its LUTs
        api.vertical_packing_batch_dev(tb, d_l, d_r, size, E, st.cuda_stream)
    with torch.cuda.stream(st):
        ms = dev_time(vp, 1)
    out = d_r.cpu().numpy().view(np.uint64)
    vals = (bits.reshape(E, size) << np.arange(size, dtype=np.uint64)).sum(axis=1)
    ph = syn.tlwe_phase(out, rk)
    ok = bool(np.array_equal(((ph + (np.uint64(1) << np.uint64(53))) >> np.uint64(54)) & np.uint64(1023), lut.reshape(-1)[vals]))
    res["leveled_lut"] = {"value": E / ms * 1e3, "unit": "LUT evaluations/s", "batch": E, "input_bits": size, "ms_per_batch": ms,
                          "params": f"N={Nl} l={ll} Bg_bit={Bgl}", "all_outputs_decrypt_to_lut": ok}
    tb.free()
    return res
