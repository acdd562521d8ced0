import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def prototypes():
    """The reference's Nyquist(M) h/g fixtures (tests/golden/prototypes.npz, see make_golden.py)."""
    P = np.load(os.path.join(GOLDEN, "prototypes.npz"))
    return {k: P[k] for k in P.files}


def proto(prototypes, M, m, r):
    import btk_b200

    key = f"h_{M}_{m}_{r}"
    if key in prototypes:
        return prototypes[key], prototypes[f"g_{M}_{m}_{r}"]
    return btk_b200.workloads.kaiser_prototype(M, m, r)


def golden_cases():
    return sorted(f[len("golden_"):-4] for f in os.listdir(GOLDEN) if f.startswith("golden_") and f.endswith(".npz"))


def load_golden(name):
    Z = np.load(os.path.join(GOLDEN, f"golden_{name}.npz"))
    return {k: Z[k] for k in Z.files}
