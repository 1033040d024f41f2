"""Batch-1 PBS time right after a full-batch launch (what bench.py measures) vs in a quiet loop."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mosfhet_b200 import api, synthetic as syn
from mosfhet_b200.params import LEVEL1 as P
api.init(0)
lwe_key, rlwe_key = syn.binary_key(P.n, 1), syn.binary_key(P.N, 2)
bsk = api.BootstrapKey.synthesize(P, lwe_key, rlwe_key, seed=3)
B = 4096
cts = syn.tlwe_encrypt(syn.encode(np.arange(B) % 4, 4), lwe_key, P.lwe_sigma, seed=4)
lut = syn.encode((3 * np.arange(4) + 1) % 4, 4)
d_in = torch.from_numpy(cts.view(np.int64)).cuda()
d_tv = torch.from_numpy(syn.test_vector(lut, P.N, 1).view(np.int64)).cuda()
d_out = torch.empty((B, P.N + 1), dtype=torch.int64, device="cuda")
st = torch.cuda.Stream()
def one(count):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st); api.pbs_dev(bsk, d_out, d_tv, 1, d_in, 4, count, st.cuda_stream); e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)
import subprocess
def smi(tag):
    q = "clocks.sm,clocks.mem,clocks.gr,pstate,power.draw,temperature.gpu,clocks_throttle_reasons.active"
    print(tag, subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip())
print("quiet  :", " ".join(f"{one(1):.2f}" for _ in range(8)))
smi("quiet smi :")
print("quiet B=148:", " ".join(f"{one(148):.2f}" for _ in range(4)))
for _ in range(3): one(B)
print("loaded :", " ".join(f"{one(1):.2f}" for _ in range(12)))
smi("loaded smi:")
print("loaded B=148:", " ".join(f"{one(148):.2f}" for _ in range(4)))
print("loaded B=1 again:", " ".join(f"{one(1):.2f}" for _ in range(4)))
import subprocess
print(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip())
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
flush.fill_(1); torch.cuda.synchronize()
print("flushed:", " ".join(f"{one(1):.2f}" for _ in range(8)))
for _ in range(2): one(B)
api.set_kernel_policy(2)
print("loaded, k1 kernel at batch 1:", " ".join(f"{one(1):.2f}" for _ in range(6)))
api.set_kernel_policy(4)
print("loaded, k1c kernel at batch 1:", " ".join(f"{one(1):.2f}" for _ in range(6)))
api.set_kernel_policy(0)
import time; time.sleep(2.0)
print("after 2 s idle:", " ".join(f"{one(1):.2f}" for _ in range(6)))
