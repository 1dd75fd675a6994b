import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "refdata: needs the reference tree at /root/reference (skipped elsewhere)")


@pytest.fixture(scope="session")
def rt():
    """The product package (its directory name is not an identifier)."""
    import importlib
    return importlib.import_module("2015-raytracing_b200")


@pytest.fixture(scope="session")
def oracle_lib():
    """Kernel oracle: the reference's own code.cl text when it was compiled (oracle/_ref),
    else our C restatement."""
    from oracle import refcl
    return refcl.load_best()


@pytest.fixture(scope="session")
def ref_lib(oracle_lib):
    """The reference's own kernel text (oracle/_ref) -- needed for the assignments the plain-C restatement does not
    cover (A01-A09); skipped where only the restatement is available."""
    if oracle_lib.kind != "reference":
        pytest.skip("oracle/_ref (the reference's own kernels) is not built here")
    return oracle_lib


@pytest.fixture(scope="session")
def gpu_ctx(rt):
    ctx = rt.lib.Context(0)
    yield ctx
    ctx.close()
