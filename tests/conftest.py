import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs oracle/_ref/libbsref.so (the compiled reference)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.bindings import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    from oracle.bindings import Reference, reference_available
    if not reference_available():
        pytest.skip("oracle/_ref/libbsref.so not built (reference tree absent)")
    return Reference(calc_threads=2)


def pytest_sessionfinish(session, exitstatus):
    """BSGPU_REDZONE=1 (tests/test_gpu_redzones.py runs the parity modules like that in a child process): after the last test
    every red zone behind a device buffer of the library is read back; a zone a kernel wrote into fails the run."""
    if not os.environ.get("BSGPU_REDZONE"):
        return
    import ctypes as C
    so = os.path.join(ROOT, "bs_call_b200", "libbsgpu.so")
    lib = C.CDLL(os.environ.get("BSGPU_LIB_PATH", so))
    a, b = C.c_ulonglong(0), C.c_ulonglong(0)
    rc = lib.bsgpu_debug_redzones(C.byref(a), C.byref(b))
    print("\nredzones: checked %d corrupt %d rc %d" % (a.value, b.value, rc))
    if b.value or rc != 1:
        session.exitstatus = 1
