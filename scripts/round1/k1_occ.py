"""Level-1 occupancy experiment: 2+1 level batches (4 CTAs/SM) vs the default 3-level batch (3 CTAs/SM)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mosfhet_b200 import api, synthetic as syn
from mosfhet_b200.params import NAMED

api.init(0)
B = int(os.environ.get("BATCH", "4096"))
P = NAMED[os.environ.get("WL", "level1")]
lwe_key, rlwe_key = syn.binary_key(P.n, 1), syn.binary_key(P.N, 2)
bsk = api.BootstrapKey.synthesize(P, lwe_key, rlwe_key, seed=3)
msgs = np.arange(B) % 4
cts = syn.tlwe_encrypt(syn.encode(msgs, 4), lwe_key, P.lwe_sigma, seed=4)
lut = syn.encode((3 * np.arange(4) + 1) % 4, 4)
d_in = torch.from_numpy(cts.view(np.int64)).cuda()
d_tv = torch.from_numpy(syn.test_vector(lut, P.N, 1).view(np.int64)).cuda()
d_out = torch.empty((B, P.N + 1), dtype=torch.int64, device="cuda")
st = torch.cuda.Stream()
variants = [tuple(int(x) for x in v.split(",")) for v in os.environ.get("VARIANTS", "3,1,1;2,1,0;2,4,0;2,4,1").split(";")]
for v in variants:
    os.environ["MB200_K1_LB"], os.environ["MB200_K1_MINB"], os.environ["MB200_K1_PF"] = (str(x) for x in v)
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
    print(f"{api.last_blind_rotate_kernel():46s} {ms:8.2f} ms  {B/ms*1e3:9.0f} PBS/s  ok={ok}", flush=True)
