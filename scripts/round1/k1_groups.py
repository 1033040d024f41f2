"""Times the k1 kernel with G ciphertexts per CTA (MB200_K1_G) on the GPU box."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mosfhet_b200 import api, synthetic as syn
from mosfhet_b200.params import NAMED

api.init(0)
for B in (4096, 4097):
    for wl, gs in (("level1", [1, 2, 3]), ("level2", [1, 2])):
        P = NAMED[wl]
        lwe_key, rlwe_key = syn.binary_key(P.n, 1), syn.binary_key(P.N, 2)
        bsk = api.BootstrapKey.synthesize(P, lwe_key, rlwe_key, seed=3)
        msgs = np.arange(B) % 4
        cts = syn.tlwe_encrypt(syn.encode(msgs, 4), lwe_key, P.lwe_sigma, seed=4)
        cts[5, 3] = 0          # a step with a_i = 0 inside a lockstep group
        lut = syn.encode((3 * np.arange(4) + 1) % 4, 4)
        d_in = torch.from_numpy(cts.view(np.int64)).cuda()
        d_tv = torch.from_numpy(syn.test_vector(lut, P.N, 1).view(np.int64)).cuda()
        d_out = torch.empty((B, P.N + 1), dtype=torch.int64, device="cuda")
        st = torch.cuda.Stream()
        ref = None
        for g in gs:
            os.environ["MB200_K1_G"] = str(g)
            ts = []
            for it in range(4 if B == 4096 else 1):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                api.pbs_dev(bsk, d_out, d_tv, 1, d_in, 4, B, st.cuda_stream)
                e1.record(st)
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            out = d_out.cpu().numpy().view(np.uint64)
            ok = syn.torus_distance(syn.tlwe_phase(out, rlwe_key), lut[msgs]).max() <= (1 << 58)
            if ref is None:
                ref = out.copy()
            same = bool(np.array_equal(ref, out))     # same arithmetic per ciphertext -> identical words
            ms = min(ts[1:]) if len(ts) > 1 else ts[0]
            print(f"B={B} {wl} {api.last_blind_rotate_kernel():52s} {ms:8.2f} ms  {B/ms*1e3:9.0f} PBS/s  fp64 {P.flops_per_pbs()*B/ms*1e-9:6.2f} TF  ok={ok} same_as_g1={same}", flush=True)
        bsk.free()
