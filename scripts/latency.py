"""Small-batch latency of the blind-rotation kernels (k1, k1q, k1h, and the 2-CTA cluster kernel k1c)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mosfhet_b200 import api, synthetic as syn
from mosfhet_b200.params import NAMED
api.init(0)
for wl in ("level1", "level2"):
    P = NAMED[wl]
    lwe_key, rlwe_key = syn.binary_key(P.n, 1), syn.binary_key(P.N, 2)
    bsk = api.BootstrapKey.synthesize(P, lwe_key, rlwe_key, seed=3)
    st = torch.cuda.Stream()
    for B in [int(v) for v in os.environ.get("BATCHES", "1,16,74,148,296,444,592,888,1184").split(",")]:
        msgs = np.arange(B) % 4
        cts = syn.tlwe_encrypt(syn.encode(msgs, 4), lwe_key, P.lwe_sigma, seed=4)
        lut = syn.encode((3 * np.arange(4) + 1) % 4, 4)
        d_in = torch.from_numpy(cts.view(np.int64)).cuda()
        d_tv = torch.from_numpy(syn.test_vector(lut, P.N, 1).view(np.int64)).cuda()
        d_out = torch.empty((B, P.N + 1), dtype=torch.int64, device="cuda")
        for tag, pol in (("k1", 2), ("k1q", 5), ("k1h", 3), ("k1c", 4), ("auto", 0)):
            api.set_kernel_policy(pol)
            ts = []
            for it in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st); api.pbs_dev(bsk, d_out, d_tv, 1, d_in, 4, B, st.cuda_stream); e1.record(st)
                torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
            out = d_out.cpu().numpy().view(np.uint64)
            ok = syn.torus_distance(syn.tlwe_phase(out, rlwe_key), lut[msgs]).max() <= (1 << 58)
            print(f"{wl} B={B:4d} {tag:6s} {min(ts[1:]):8.3f} ms  {api.last_blind_rotate_kernel():40s} ok={ok}", flush=True)
        api.set_kernel_policy(0)
    bsk.free()
