import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


# GPU cases written after the round's last GPU run have only been exercised in the CPU suite (through tests/standin): they are collected
# last, so that with `-x` a surprise in one of them cannot hide the results of the cases already verified on a B200.
_NOT_YET_RUN_ON_GPU = ("test_svanno.py", "test_methyl.py", "test_lr_full_chain.py", "test_genotype_mode.py", "test_multi_sample.py", "test_seq_identity.py", "test_svtset.py", "test_edlib_compat.py", "test_example_data.py")


def pytest_collection_modifyitems(config, items):
    items.sort(key=lambda it: os.path.basename(str(it.fspath)) in _NOT_YET_RUN_ON_GPU)   # stable: the order inside both groups is kept


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


@pytest.fixture(scope="session")
def standin(ref):
    """tests/standin/host_standin.cpp built next to the tests (never into the product libraries)."""
    import ctypes as C
    import subprocess
    import delly_b200
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    refdir = os.path.join(root, "oracle", "_ref")
    pkg = os.path.join(root, "delly_b200")
    out = os.path.join(here, "standin", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libhost_standin.so")
    src = os.path.join(here, "standin", "host_standin.cpp")
    deps = [src] + [os.path.join(pkg, "host", f) for f in os.listdir(os.path.join(pkg, "host"))]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in deps):
        delly_b200.lib()   # the CUDA library must exist: the stand-in links against it for every other entry point
        r = subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-w", "-Wl,-Bsymbolic", "-o", so, src, "-L" + refdir, "-l:libdelly_ref.so",
                            "-L" + pkg, "-l:libdelly_b200.so", "-Wl,-rpath," + refdir, "-Wl,-rpath," + pkg], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    lib = C.CDLL(so)
    lib.standin_ctx.restype = C.c_void_p
    return lib


@pytest.fixture(params=[pytest.param("gpu", marks=pytest.mark.gpu), "edlib-standin"])
def hostdev(request):
    """(host library, context handle) for the batched mirrors whose only device calls are edit distances / edit paths:
    "gpu" = the real libraries on a B200 (`-m gpu`); "edlib-standin" = the CPU suite's stand-in (tests/standin/host_standin.cpp), where
    those calls are forwarded to the reference's edlib, so the host logic around them is checked without a GPU."""
    import ctypes as C
    import delly_b200
    if request.param == "gpu":
        return delly_b200.hostlib(), request.getfixturevalue("ctx").h
    lib = request.getfixturevalue("standin")
    return lib, C.c_void_p(lib.standin_ctx())


@pytest.fixture(scope="session")
def ref9():
    """The complete long-read chain of the reference (annotation and methylation included) over in-memory alignments (oracle/_ref)."""
    from oracle import pyoracle
    r = pyoracle.ref9()
    if r is None:
        pytest.skip("oracle/_ref/libdelly_ref9.so not available")
    return r


@pytest.fixture(scope="session")
def ref10():
    """`delly merge`'s sequence-identity helpers (src/merge.h:187-243, build-time extract) with the reference's edlib (oracle/_ref)."""
    from oracle import pyoracle
    r = pyoracle.ref10()
    if r is None:
        pytest.skip("oracle/_ref/libdelly_ref10.so not available")
    return r
