"""Batch-1 latency breakdown: blind rotation, key switch (device timed) and the whole host call."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mosfhet_b200 import api, synthetic as syn
from mosfhet_b200.params import NAMED
api.init(0)
for wl in ("level1", "level2"):
    P = NAMED[wl]
    lwe_key, rlwe_key = syn.binary_key(P.n, 1), syn.binary_key(P.N, 2)
    bsk = api.BootstrapKey.synthesize(P, lwe_key, rlwe_key, seed=3)
    ksk = api.KeySwitchKey.synthesize(P, rlwe_key, lwe_key, seed=4)
    st = torch.cuda.Stream()
    for B in (1, 8):
        cts = syn.tlwe_encrypt(syn.encode(np.arange(B) % 4, 4), lwe_key, P.lwe_sigma, seed=4)
        lut = syn.encode((3 * np.arange(4) + 1) % 4, 4)
        tv = syn.test_vector(lut, P.N, 1)
        d_in = torch.from_numpy(cts.view(np.int64)).cuda()
        d_tv = torch.from_numpy(tv.view(np.int64)).cuda()
        d_mid = torch.empty((B, P.N + 1), dtype=torch.int64, device="cuda")
        d_out = torch.empty((B, P.n + 1), dtype=torch.int64, device="cuda")
        def ev(fn):
            ts = []
            for _ in range(6):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st); fn(); e1.record(st); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
            return min(ts[1:])
        t_pbs = ev(lambda: api.pbs_dev(bsk, d_mid, d_tv, 1, d_in, 4, B, st.cuda_stream))
        t_ks = ev(lambda: api.ks_dev(ksk, d_out, d_mid, B, st.cuda_stream))
        hs = []
        for _ in range(6):
            t0 = time.perf_counter(); api.pbs_ks_host(bsk, ksk, tv, cts, 4); hs.append((time.perf_counter() - t0) * 1e3)
        print(f"{wl} B={B}: blind rotation {t_pbs:.3f} ms ({api.last_blind_rotate_kernel()}), key switch {t_ks:.3f} ms, "
              f"host call (H2D + both + D2H) {min(hs[1:]):.3f} ms", flush=True)
    bsk.free(); ksk.free()
