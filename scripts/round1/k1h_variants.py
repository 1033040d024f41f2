"""Times k1 (T=M/8) vs k1h (T=M/4) blind-rotation kernels on the GPU box."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mosfhet_b200 import api, synthetic as syn
from mosfhet_b200.params import NAMED, Params

api.init(0)
B = int(os.environ.get("BATCH", "4096"))
cases = [("level1", NAMED["level1"]), ("set1", NAMED["set1"]), ("n512l3", Params(500, 512, 1, 3, 6, 5, 2, 2.0**-15, 2.0**-25))]
for wl, P in cases:
    lwe_key, rlwe_key = syn.binary_key(P.n, 1), syn.binary_key(P.N, 2)
    bsk = api.BootstrapKey.synthesize(P, lwe_key, rlwe_key, seed=3)
    msgs = np.arange(B) % 4
    cts = syn.tlwe_encrypt(syn.encode(msgs, 4), lwe_key, 2.0**-20, seed=4)
    lut = syn.encode((3 * np.arange(4) + 1) % 4, 4)
    d_in = torch.from_numpy(cts.view(np.int64)).cuda()
    d_tv = torch.from_numpy(syn.test_vector(lut, P.N, 1).view(np.int64)).cuda()
    d_out = torch.empty((B, P.N + 1), dtype=torch.int64, device="cuda")
    st = torch.cuda.Stream()
    for tag, env in (("k1", {"MB200_K1H": "0"}), ("k1h-3", {"MB200_K1H": "1", "MB200_K1H_MINB": "3"}),
                     ("k1h-2", {"MB200_K1H": "1", "MB200_K1H_MINB": "2"}), ("k1h-4", {"MB200_K1H": "1", "MB200_K1H_MINB": "4"})):
        if tag in ("k1h-2", "k1h-4") and wl != "level1":
            continue
        os.environ.update(env)
        ts = []
        for it in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            api.pbs_dev(bsk, d_out, d_tv, 1, d_in, 4, B, st.cuda_stream)
            e1.record(st)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        out = d_out.cpu().numpy().view(np.uint64)
        ok = syn.torus_distance(syn.tlwe_phase(out, rlwe_key), lut[msgs]).max() <= (1 << 58)
        ms = min(ts[1:])
        print(f"{wl:8s} {tag:6s} {api.last_blind_rotate_kernel():46s} {ms:8.2f} ms  {B/ms*1e3:9.0f} PBS/s  fp64 {P.flops_per_pbs()*B/ms*1e-9:6.2f} TF  ok={ok}", flush=True)
    bsk.free()
