"""Host-side timeline of functional_bootstrap_keyswitch_batch over reference-made handles (MB200_TRACE=1)."""
import os, sys, time
os.environ["MB200_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mosfhet_b200 import abi, api
from mosfhet_b200.params import NAMED
from oracle import parity
api.init(0)
P = NAMED[os.environ.get("WL", "level1")]
B = int(os.environ.get("BATCH", "4096"))
S = parity.ReferenceSetup(P, B)
R = S.R
api.set_host_fft_layout(R.layout)
api.register_bootstrap_key(S.bk); api.register_ks_key(S.ksk)
outs = [R.tlwe_alloc_sample(P.n) for _ in range(B)]
a_out, a_in, a_tv = abi.handle_array(outs, abi.TLWE), abi.handle_array(S.inputs, abi.TLWE), abi.handle_array([S.tv], abi.TRLWE)
fn = api.lib().functional_bootstrap_keyswitch_batch
for it in range(3):
    t0 = time.perf_counter()
    fn(a_out, a_tv, 1, a_in, S.bk, S.ksk, 4, B)
    print(f"call {it}: {1e3 * (time.perf_counter() - t0):.2f} ms", file=sys.stderr, flush=True)
