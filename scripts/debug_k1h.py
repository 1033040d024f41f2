import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mosfhet_b200 import api, synthetic as syn
from mosfhet_b200.params import Params
api.init(0)
P = Params(96, 1024, 1, 3, 6, 7, 2, 2.0**-20, 2.0**-30)
lwe_key, rlwe_key = syn.binary_key(P.n, 1), syn.binary_key(P.N, 2)
bsk = api.BootstrapKey.synthesize(P, lwe_key, rlwe_key, seed=3)
B = 64
msgs = np.arange(B) % 4
cts = syn.tlwe_encrypt(syn.encode(msgs, 4), lwe_key, P.lwe_sigma, seed=4)
lut = syn.encode((3 * np.arange(4) + 1) % 4, 4)
tv = syn.test_vector(lut, P.N, 1)
outs = {}
for name, pol in (("generic", 1), ("k1", 2), ("k1h", 3)):
    api.set_kernel_policy(pol)
    outs[name] = api.pbs_host(bsk, tv, cts, 4).copy()
    print(name, api.last_blind_rotate_kernel())
ph = {k: syn.tlwe_phase(v, rlwe_key) for k, v in outs.items()}
for a, b in (("k1", "generic"), ("k1h", "generic"), ("k1h", "k1")):
    d = syn.torus_distance(ph[a], ph[b])
    print(a, "vs", b, "max 2^%.1f" % np.log2(d.max() + 1), "bad cts:", np.nonzero(d > (1 << 44))[0][:20])
bad = np.nonzero(syn.torus_distance(ph["k1h"], ph["generic"]) > (1 << 44))[0]
rot = ((cts[:, :P.n] + (np.uint64(1) << np.uint64(52))) >> np.uint64(53)).astype(np.int64)
for c in bad[:6]:
    print("ct", c, "zeros at", np.nonzero(rot[c] == 0)[0], "min/max rot", rot[c].min(), rot[c].max(), "b rot", int((cts[c, P.n] >> np.uint64(53))))
good = [c for c in range(B) if c not in set(bad)][:4]
for c in good:
    print("good ct", c, "zeros at", np.nonzero(rot[c] == 0)[0], "min/max rot", rot[c].min(), rot[c].max())
