import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def load_golden():
    z = np.load(os.path.join(ROOT, "tests", "golden", "exact_gp_golden.npz"))
    cases = []
    for i in range(int(z["n_cases"])):
        pre = "c%02d_" % i
        cases.append({k[len(pre):]: z[k] for k in z.files if k.startswith(pre)})
    for c in cases:
        c["kernel"] = str(c["kernel"])
        c["variance"] = float(c["variance"])
        c["noise"] = float(c["noise"])
        c["ARD"] = bool(c["ARD"])
        c["lml"] = float(c["lml"])
    return cases


def normwise(a, b):
    """The parity gate of SURVEY.md section 8(d): ||a - b||_inf / ||b||_inf."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.fixture(scope="session")
def golden_cases():
    return load_golden()


@pytest.fixture(scope="session")
def gpu_handle():
    from thesis_code_b200 import _gpx
    h = _gpx.Handle(0)
    yield h
    h.close()


def load_reference_pinned():
    """tests/golden/reference_pinned.npz: outputs of REFERENCE code executed in the build container
    (tests/golden/make_golden_reference.py): {tag: {field: array}} for the GPVanillaModel cases v0..v2 and the
    NormalizerModel<GPVanillaModel> cases n0..n2."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "reference_pinned.npz"))
    cases = {}
    for k in z.files:
        tag, field = k.split("_", 1)
        cases.setdefault(tag, {})[field] = z[k]
    return cases


@pytest.fixture(scope="session")
def reference_pinned():
    return load_reference_pinned()


@pytest.fixture(scope="session")
def golden_large():
    path = os.path.join(ROOT, "tests", "golden", "exact_gp_golden_large.npz")
    z = np.load(path)
    out = {}
    for k in z.files:
        tag, field = k.split("_", 1)
        out.setdefault(tag, {})[field] = z[k]
    return out
