import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import load_golden
from mosfhet_b200 import abi, api
from mosfhet_b200.params import Params
from oracle import oracle as O
api.init(0)
for name in ("tiny_k1_spqlios", "tiny_k2_spqlios", "small_l4_spqlios"):
    g = load_golden(name); P = g["P"]
    Pm = Params(P["n"], P["N"], P["k"], P["l"], P["Bg_bit"], P["t"], P["base_bit"])
    ksk = api.KeySwitchKey.from_host(Pm, g["ksk"])
    got = api.ks_host(ksk, g["fb_out"])
    for b in range(got.shape[0]):
        bad = np.nonzero(got[b] != g["ks_out"][b])[0]
        print(name, "flat", b, "mismatch cols", bad[:10], len(bad))
    hksk = abi.HostKSKey(g["ksk"], P["base_bit"])
    for b in range(2):
        out = abi.HostTLWE.zeros(P["n"])
        api.tlwe_keyswitch(out, abi.HostTLWE(g["fb_out"][b]), hksk)
        bad = np.nonzero(out.flat() != g["ks_out"][b])[0]
        print(name, "struct", b, "mismatch cols", bad[:10], len(bad))
    api.release_ks_key(hksk); ksk.free()
