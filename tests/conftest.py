import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_SETS = ["tiny_k1_spqlios", "tiny_k2_spqlios", "small_l4_spqlios"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    n, N, k, l, Bg_bit, t, base_bit = (int(x) for x in g["params"])
    g["P"] = dict(n=n, N=N, k=k, l=l, Bg_bit=Bg_bit, t=t, base_bit=base_bit)
    g["layout"] = int(g["layout"])
    return g


@pytest.fixture(params=GOLDEN_SETS)
def golden(request):
    return load_golden(request.param)


@pytest.fixture
def golden_mv():
    return load_golden("tiny_k1_mv_spqlios")


@pytest.fixture
def golden_cb():
    return load_golden("tiny_cb_spqlios")


@pytest.fixture
def golden_r4():
    g = load_golden("tiny_r4_spqlios")
    g["unfolding"] = int(g["unfolding"])
    return g


@pytest.fixture
def golden_ffnt():
    return load_golden("tiny_k1_ffnt")
