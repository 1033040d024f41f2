"""Key-switch launch time against the batch size (device resident, synthetic key).  The warp-per-ciphertext sweep is one
serial chain, so below a full wave (28 warps x SMs) the launcher slices every ciphertext's sweep over several warps; the
sliced results must equal the unsliced ones bit for bit (integer sums, any order)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mosfhet_b200 import api, synthetic as syn
from mosfhet_b200.params import NAMED

api.init(0)
WL = os.environ.get("WL", "level1")
P = NAMED[WL]
ksk = api.KeySwitchKey.synthesize(P, syn.binary_key(P.k * P.N, 1002), syn.binary_key(P.n, 1001), seed=12)
dev = torch.device("cuda:0")
rng = np.random.default_rng(3)
BIG = 8192
x = torch.from_numpy(rng.integers(-2**63, 2**63 - 1, size=(BIG, P.k * P.N + 1), dtype=np.int64)).to(dev)
ref = torch.empty((BIG, P.n + 1), dtype=torch.int64, device=dev)
st = torch.cuda.Stream(dev)
with torch.cuda.stream(st):
    api.ks_dev(ksk, ref, x, BIG, st.cuda_stream)            # two full waves: one warp per ciphertext, no slicing
torch.cuda.synchronize()
for B in [int(v) for v in os.environ.get("BATCHES", "1,16,64,256,592,1024,1500,2048,2072,3000,4096,8192").split(",")]:
    d_out = torch.empty((B, P.n + 1), dtype=torch.int64, device=dev)
    with torch.cuda.stream(st):
        for _ in range(3):
            api.ks_dev(ksk, d_out, x, B, st.cuda_stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(5):
            api.ks_dev(ksk, d_out, x, B, st.cuda_stream)
        e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    same = bool(torch.equal(d_out, ref[:B]))
    print(f"{WL} batch {B:6d}: {ms:7.3f} ms  {ms / B * 1e3:8.3f} us/ct  identical_to_unsliced={same}", flush=True)
