import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def sfe():
    """The product package (ctypes binding of libslamfe.so)."""
    return importlib.import_module("slam-robot_b200")


@pytest.fixture(scope="session")
def synth():
    return importlib.import_module("slam-robot_b200.synth")


@pytest.fixture(scope="session")
def po():
    from oracle import pyoracle
    return pyoracle


@pytest.fixture(scope="session")
def fe(sfe):
    """One FrontEnd context on cuda:0; GPU tests fail loudly if it cannot be created."""
    f = sfe.FrontEnd(0)
    yield f
    f.close()


@pytest.fixture(scope="session")
def pair640(synth):
    A, B = synth.make_pairs(1, 1, 480, 640)
    return A[0].numpy(), B[0].numpy()


def assert_bits_equal(a, b, what=""):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    ai, bi = a.view(np.uint32 if a.dtype == np.float32 else a.dtype), b.view(np.uint32 if b.dtype == np.float32 else b.dtype)
    bad = np.flatnonzero(ai.ravel() != bi.ravel())
    assert bad.size == 0, "%s: %d of %d entries differ; first at %d: %r vs %r" % (
        what, bad.size, ai.size, bad[0], a.ravel()[bad[0]], b.ravel()[bad[0]])
