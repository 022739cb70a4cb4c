import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))


@pytest.fixture(scope="session")
def oracle():
    """The CPU restatement (oracle/visfd_oracle.cpp) -- the checker, never the product."""
    from oracle.pyoracle import Oracle
    return Oracle("port")


@pytest.fixture(scope="session")
def ref_oracle():
    """The unmodified reference (oracle/_ref), when it has been built and shipped."""
    from oracle.pyoracle import Oracle, have
    if not have("reference"):
        pytest.skip("oracle/_ref/libvisfd_ref.so not present")
    return Oracle("reference")


@pytest.fixture(scope="session")
def ctx():
    """A CUDA context on cuda:0; fails loudly (no fallback) if the library or GPU is missing."""
    import visfd_b200
    c = visfd_b200.Context(0)
    yield c
    c.close()
