import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu through gpurun)")


def load_golden_cases():
    z = np.load(os.path.join(GOLDEN, "chain_small.npz"))
    cases = []
    for line in z["meta"]:
        idx, name, space, clip, grid, k = str(line).split("|")
        cases.append(dict(idx=int(idx), name=name, space=space, clip=float(clip), grid=int(grid), k=int(k),
                          inp=z[f"in_{name}"], out=z[f"out_{idx}"]))
    gate = [tuple(str(g).split("|")) for g in z["gate"]]
    return cases, gate, z


def load_sha_pins():
    pins = []
    with open(os.path.join(GOLDEN, "chain_sha1.txt")) as fh:
        for line in fh:
            if line.startswith("#") or not line.strip():
                continue
            h, w, space, grid, k, seed, sha = line.strip().split("|")
            pins.append(dict(h=int(h), w=int(w), space=space, grid=int(grid), k=int(k), seed=int(seed), sha=sha))
    return pins


@pytest.fixture(scope="session")
def ctx():
    """GPU context; only -m gpu tests may request it."""
    import rvb200
    return rvb200.default_context()
