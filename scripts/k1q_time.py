"""Times the blind-rotation kernels against each other on one B200 (device-resident inputs, CUDA events on the
launching stream, best of 3 after a warm-up): policy 2 = k1 (T = M/8), 5 = k1q (T = M/4)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mosfhet_b200 import api, synthetic as syn
from mosfhet_b200.params import NAMED

api.init(0)
B = int(os.environ.get("BATCH", "4096"))
peak = api.measure_fp64_tflops()
print(f"fp64 peak (measured) {peak:.2f} TFLOP/s", flush=True)
for wl in os.environ.get("WLS", "level1,level2").split(","):
    P = NAMED[wl]
    lwe_key, rlwe_key = syn.binary_key(P.n, 1), syn.binary_key(P.N, 2)
    bsk = api.BootstrapKey.synthesize(P, lwe_key, rlwe_key, seed=3)
    msgs = np.arange(B) % 4
    cts = syn.tlwe_encrypt(syn.encode(msgs, 4), lwe_key, P.lwe_sigma, seed=4)
    lut = syn.encode((3 * np.arange(4) + 1) % 4, 4)
    d_in = torch.from_numpy(cts.view(np.int64)).cuda()
    d_tv = torch.from_numpy(syn.test_vector(lut, P.N, 1).view(np.int64)).cuda()
    d_out = torch.empty((B, P.N + 1), dtype=torch.int64, device="cuda")
    st = torch.cuda.Stream()
    for pol in [int(x) for x in os.environ.get("POLICIES", "2,5").split(",")]:
        api.set_kernel_policy(pol)
        ts = []
        for it in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            api.pbs_dev(bsk, d_out, d_tv, 1, d_in, 4, B, st.cuda_stream)
            e1.record(st)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        out = d_out.cpu().numpy().view(np.uint64)
        dist = syn.torus_distance(syn.tlwe_phase(out, rlwe_key), lut[msgs])
        ok = dist.max() <= (1 << 58)
        ms = min(ts[1:])
        tf = P.flops_per_pbs() * B / ms * 1e-9
        print(f"{wl:7s} {api.last_blind_rotate_kernel():44s} {ms:8.2f} ms  {B/ms*1e3:9.0f} PBS/s  {tf:6.2f} TF  frac {tf/peak:.3f}  "
              f"ok={ok} maxdist=2^{np.log2(float(dist.max()) + 1):.1f}", flush=True)
    api.set_kernel_policy(0)
    bsk.free()
