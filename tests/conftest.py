import importlib
import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def pkg():
    """The product package (selfsupervised-nvsf_b200/), with the CUDA library built."""
    p = importlib.import_module("selfsupervised-nvsf_b200")
    from importlib import import_module
    import_module("selfsupervised-nvsf_b200.build").build()
    return p


@pytest.fixture(scope="session")
def synth():
    return importlib.import_module("selfsupervised-nvsf_b200.synth")


@pytest.fixture(scope="session")
def oracle():
    """numpy front-end of the C oracle (test infrastructure)."""
    from oracle import raymarching_oracle
    raymarching_oracle.lib()
    return raymarching_oracle


@pytest.fixture(scope="session")
def ref_ext():
    """The reference's own extension rebuilt for sm_100a, or None when it is not in the tree."""
    path = os.path.join(ROOT, "oracle", "_ref", "_raymarching_ref.so")
    if not os.path.exists(path) or not _has_cuda():
        return None
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location("_raymarching_ref", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def bits(a):
    """View a float32 array as uint32 for bit-exact comparisons."""
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_bits_equal(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if a.dtype.kind == "f":
        same = bits(a) == bits(b)
    else:
        same = a == b
    if not same.all():
        bad = np.argwhere(~same)
        i = tuple(bad[0])
        raise AssertionError(f"{what}: {len(bad)} of {a.size} elements differ; first at {i}: "
                             f"{a[i]!r} vs {b[i]!r}")
