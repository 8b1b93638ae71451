import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ctx():
    """One dgpu context on cuda:0 through the C ABI. Fails loudly (no CPU fallback)."""
    import delly_b200
    c = delly_b200.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    return pyoracle.oracle()


@pytest.fixture(scope="session")
def ref():
    """The reference compiled verbatim (oracle/_ref). Built here; travels prebuilt to the GPU box."""
    from oracle import pyoracle
    r = pyoracle.ref()
    if r is None:
        pytest.skip("oracle/_ref/libdelly_ref.so not available")
    return r


@pytest.fixture(scope="session")
def ref3():
    """The reference's bolog.h / coverage.h (_computeGLs, _generateProbes) compiled verbatim (oracle/_ref)."""
    from oracle import pyoracle
    r = pyoracle.ref3()
    if r is None:
        pytest.skip("oracle/_ref/libdelly_ref3.so not available")
    return r


@pytest.fixture(scope="session")
def ref4():
    """The reference's genotype.h (genotypeLR) compiled verbatim, htslib served from memory (oracle/_ref)."""
    from oracle import pyoracle
    r = pyoracle.ref4()
    if r is None:
        pytest.skip("oracle/_ref/libdelly_ref4.so not available")
    return r


@pytest.fixture(scope="session")
def ref5():
    """The reference's shortpe.h (assembleSplitReads) compiled verbatim, htslib served from memory (oracle/_ref)."""
    from oracle import pyoracle
    r = pyoracle.ref5()
    if r is None:
        pytest.skip("oracle/_ref/libdelly_ref5.so not available")
    return r


@pytest.fixture(scope="session")
def ref6():
    """The reference's util.h compiled itself (getLibraryParams), htslib served from memory (oracle/_ref)."""
    from oracle import pyoracle
    r = pyoracle.ref6()
    if r is None:
        pytest.skip("oracle/_ref/libdelly_ref6.so not available")
    return r


@pytest.fixture(scope="session")
def ref7():
    """The reference's modvcf.h (vcfOutput) compiled verbatim over a recording VCF/BCF stand-in (oracle/_ref)."""
    from oracle import pyoracle
    r = pyoracle.ref7()
    if r is None:
        pytest.skip("oracle/_ref/libdelly_ref7.so not available")
    return r


@pytest.fixture(scope="session")
def ref8():
    """The reference's svanno.h (annotateSV) compiled verbatim with its own edlib (oracle/_ref)."""
    from oracle import pyoracle
    r = pyoracle.ref8()
    if r is None:
        pytest.skip("oracle/_ref/libdelly_ref8.so not available")
    return r
