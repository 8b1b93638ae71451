"""include/dgpu_edlib.h: the edlib C API (src/edlib.h) served by the B200 library for call sites that are not batched.
CPU: it compiles as C in place of edlib.h, exports every entry point, and the CIGAR conversion equals the reference's; single alignments
through the API (host marshalling, bounds, equalities, tasks, empty sequences) equal edlibAlign — on the GPU through the real library,
in the CPU suite with the device entry points forwarded to the reference's edlib (tests/standin)."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

import delly_b200
from delly_b200 import synth
from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ALPHA = np.frombuffer(b"ACGT", np.uint8)


class Pair(C.Structure):
    _fields_ = [("first", C.c_char), ("second", C.c_char)]


class Config(C.Structure):
    _fields_ = [("k", C.c_int), ("mode", C.c_int), ("task", C.c_int), ("eq", C.POINTER(Pair)), ("neq", C.c_int)]


class Result(C.Structure):
    _fields_ = [("status", C.c_int), ("editDistance", C.c_int), ("endLocations", C.POINTER(C.c_int)), ("startLocations", C.POINTER(C.c_int)),
                ("numLocations", C.c_int), ("alignment", C.POINTER(C.c_ubyte)), ("alignmentLength", C.c_int), ("alphabetLength", C.c_int)]


def _bind(lib):
    lib.dgpu_edlibNewAlignConfig.restype = Config
    lib.dgpu_edlibNewAlignConfig.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(Pair), C.c_int]
    lib.dgpu_edlibDefaultAlignConfig.restype = Config
    lib.dgpu_edlibAlign.restype = Result
    lib.dgpu_edlibAlign.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, Config]
    lib.dgpu_edlibFreeAlignResult.argtypes = [Result]
    lib.dgpu_edlibFreeAlignResult.restype = None
    lib.dgpu_edlibAlignmentToCigar.restype = C.c_void_p
    lib.dgpu_edlibAlignmentToCigar.argtypes = [C.c_char_p, C.c_int, C.c_int]
    return lib


def test_header_compiles_as_c_in_place_of_edlib_h():
    """A C translation unit written against edlib.h (the reference's call pattern, src/coverage.h:107-115 and src/split.h:485-490) compiles
    unchanged against include/dgpu_edlib.h and links with the host library."""
    src = r'''
#include <stddef.h>
#include "dgpu_edlib.h"
int probe(const char* q, int ql, const char* t, int tl) {
  EdlibAlignResult r = edlibAlign(q, ql, t, tl, edlibNewAlignConfig(2 * ql, EDLIB_MODE_HW, EDLIB_TASK_DISTANCE, NULL, 0));
  int d = (r.status == EDLIB_STATUS_OK) ? r.editDistance : -2;
  edlibFreeAlignResult(r);
  EdlibEqualityPair eq[1] = {{'N', 'A'}};
  EdlibAlignResult p = edlibAlign(q, ql, t, tl, edlibNewAlignConfig(-1, EDLIB_MODE_NW, EDLIB_TASK_PATH, eq, 1));
  if (p.status == EDLIB_STATUS_OK && p.alignment) { char* c = edlibAlignmentToCigar(p.alignment, p.alignmentLength, EDLIB_CIGAR_EXTENDED); if (c) d += c[0] == '='; }
  if (p.numLocations > 0) d += p.endLocations[0] - p.startLocations[0];
  edlibFreeAlignResult(p);
  EdlibAlignConfig def = edlibDefaultAlignConfig();
  return d + def.k + (EDLIB_EDOP_MATCH + EDLIB_EDOP_INSERT + EDLIB_EDOP_DELETE + EDLIB_EDOP_MISMATCH);
}
'''
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "probe.c"), "w").write(src)
        r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"), "-o", os.path.join(td, "libprobe.so"),
                            os.path.join(td, "probe.c"), "-L" + os.path.join(ROOT, "delly_b200"), "-l:libdelly_b200_host.so"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_cigar_conversion_matches_reference(ref):
    H = _bind(delly_b200.hostlib())
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    rng = np.random.default_rng(2)
    for it in range(400):
        n = int(rng.integers(0, 200))
        ops = rng.choice([0, 1, 2, 3], size=n, p=[0.7, 0.1, 0.1, 0.1]).astype(np.uint8)
        if it % 7 == 0 and n:
            ops[:] = ops[0]                                   # one run
        if it % 11 == 0 and n:
            ops[int(rng.integers(0, n))] = 4                  # an invalid operation code
        for fmt in (0, 1, 2):
            out = C.create_string_buffer(2 * n + 8)
            L = ref.ref_edlib_cigar(ops.tobytes(), n, fmt, out, len(out))
            p = H.dgpu_edlibAlignmentToCigar(ops.tobytes(), n, fmt)
            if L < 0:
                # the reference only notices an invalid code where a run starts (an invalid code inside a run indexes its table out of bounds);
                # this library rejects it wherever it is
                assert (p is None) or (4 in ops.tolist())
            else:
                assert p is not None and C.string_at(p) == out.value, (it, fmt, ops.tolist())
            if p:
                libc.free(p)


def _cases(seed, n):
    rng = np.random.default_rng(seed)
    for i in range(n):
        tl = int(rng.choice([0, 1, 7, 60, 150, 700]))
        t = ALPHA[rng.integers(0, 4, size=tl)]
        kind = i % 4
        if kind == 0:
            q = ALPHA[rng.integers(0, 4, size=int(rng.choice([0, 1, 30, 200])))]
        elif tl > 8:
            a = int(rng.integers(0, tl // 2)); q = synth.mutate(rng, t[a:a + int(rng.integers(4, tl - a + 1))], sub=0.05, ins=0.02, dele=0.02)
        else:
            q = t.copy()
        if kind == 3 and len(q):
            q = q.copy(); q[rng.integers(0, len(q), size=max(1, len(q) // 8))] = ord("N")
        mode = int(rng.integers(0, 3)); task = int(rng.integers(0, 3))
        k = int(rng.choice([-1, -1, 0, 3, 20, 10 ** 6]))
        eq = b"NANCNGNT" if (kind == 3 and rng.random() < 0.7) else b""
        yield bytes(q), bytes(t), k, mode, task, eq


def test_single_alignments_match_edlib(hostdev, ref):
    H, ctxh = hostdev
    _bind(H)
    if hasattr(H, "standin_edlib_compat_init"):
        H.standin_edlib_compat_init()
    n_path = n_minus = n_eq = 0
    for q, t, k, mode, task, eq in _cases(17, 500):
        d, e, s, ops = po.edit_distance(ref, q, t, k=k, mode=mode, task=task, eq=eq)
        pairs = (Pair * max(1, len(eq) // 2))()
        for j in range(len(eq) // 2):
            pairs[j].first = eq[2 * j:2 * j + 1]; pairs[j].second = eq[2 * j + 1:2 * j + 2]
        cfg = H.dgpu_edlibNewAlignConfig(k, mode, task, pairs if eq else None, len(eq) // 2)
        r = H.dgpu_edlibAlign(q, len(q), t, len(t), cfg)
        try:
            assert r.status == 0
            assert r.editDistance == d, (q, t, k, mode, task, eq, r.editDistance, d)
            assert r.alphabetLength == len(set(q) | set(t))
            if d < 0:
                assert r.numLocations == 0 and not r.endLocations
                n_minus += 1
                continue
            assert r.numLocations == 1 and r.endLocations[0] == e, (q, t, k, mode, task, r.endLocations[0], e)
            if task >= 1 and len(q) and len(t):
                assert r.startLocations[0] == s
            if task == 2 and len(q) and len(t):
                assert bytes(r.alignment[:r.alignmentLength]) == ops
                n_path += 1
            n_eq += bool(eq)
        finally:
            H.dgpu_edlibFreeAlignResult(r)
    assert n_path > 60 and n_minus > 20 and n_eq > 40
    dflt = H.dgpu_edlibDefaultAlignConfig()
    assert (dflt.k, dflt.mode, dflt.task, dflt.neq) == (-1, 0, 0, 0)
