"""Small k1q launches for compute-sanitizer racecheck: benchmark shapes, short rotation (n = 12), 3 ciphertexts."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mosfhet_b200 import api, synthetic as syn
from mosfhet_b200.params import Params
api.init(0)
for N, l, Bg in ((1024, 3, 6), (2048, 4, 9), (1024, 2, 8), (2048, 3, 6)):
    P = Params(12, N, 1, l, Bg, 3, 2, 2.0 ** -30, 2.0 ** -50)
    lwe_key, rlwe_key = syn.binary_key(P.n, 5), syn.binary_key(P.N, 6)
    bsk = api.BootstrapKey.synthesize(P, lwe_key, rlwe_key, seed=N + l)
    msgs = np.arange(3) % 4
    cts = syn.tlwe_encrypt(syn.encode(msgs, 4), lwe_key, 2.0 ** -30, seed=7)
    lut = syn.encode((3 * np.arange(4) + 1) % 4, 4)
    tv = syn.test_vector(lut, P.N, 1)
    outs = {}
    for pol in (5, 1):
        api.set_kernel_policy(pol)
        outs[pol] = api.pbs_host(bsk, tv, cts, 4).copy()
        print(N, l, api.last_blind_rotate_kernel(), flush=True)
    d = syn.torus_distance(syn.tlwe_phase(outs[5], rlwe_key), syn.tlwe_phase(outs[1], rlwe_key)).max()
    print("  k1q vs generic phase distance 2^%.1f" % np.log2(float(d) + 1), flush=True)
    bsk.free()
api.set_kernel_policy(0)
