import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with -m gpu on the GPU box")


def load_golden(case):
    """Golden fixture -> dict of numpy arrays (column-major bytes restored to their R shape)."""
    with open(os.path.join(GOLDEN, case + ".json")) as f:
        meta = json.load(f)
    out = dict(meta=meta)
    for name, dim in meta["dims"].items():
        raw = np.fromfile(os.path.join(GOLDEN, "%s.%s.bin" % (case, name)), dtype="<f8")
        out[name] = raw.reshape(dim[::-1]).T          # column-major -> (rows, cols) view
    n = meta["dims"]["assay"][0]
    ctl = np.zeros(n, dtype=bool)
    ctl[meta["ctl_index_0based"]] = True
    out["ctl"] = ctl
    return out


@pytest.fixture(scope="session")
def golden_tiny():
    return load_golden("tiny")


@pytest.fixture(scope="session")
def golden_mid():
    return load_golden("mid")


@pytest.fixture(scope="session")
def engine():
    """One engine handle for the whole GPU session (the R shim also keeps one per session)."""
    from ruviii_b200 import Engine
    eng = Engine(devices=[0])
    yield eng
    eng.close()
