"""`delly merge`'s sequence identity between two SV calls (_seqIdentity / _bestSeqIdentity / _minRotation, src/merge.h:187-243) as batched
NW edit-distance rounds (delly_b200/host/seqidentity.hpp) against the reference's own functions (build-time extract, oracle/ref_wrap10.cpp)."""
import ctypes as C

import numpy as np
import pytest

import delly_b200
from delly_b200 import synth

ALPHA = np.frombuffer(b"ACGT", np.uint8)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_min_rotation_matches_reference(ref10):
    H = delly_b200.hostlib()
    rng = np.random.default_rng(1)
    for it in range(600):
        n = int(rng.integers(0, 60))
        kind = it % 3
        if kind == 0:
            s = ALPHA[rng.integers(0, 4, size=n)]
        elif kind == 1:
            s = ALPHA[rng.integers(0, 2, size=n)]                                 # many ties
        else:
            u = ALPHA[rng.integers(0, 4, size=int(rng.integers(1, 7)))]
            s = np.tile(u, n // len(u) + 1)[:n]                                   # periodic: several minimal rotations, one string
        s = np.concatenate([s, np.zeros(1, np.uint8)]).astype(np.uint8)
        a = C.create_string_buffer(n + 1); b = C.create_string_buffer(n + 1)
        la = ref10.ref_min_rotation(_p(s), n, a); lb = H.dh_min_rotation(_p(s), n, b)
        assert la == lb == n and a.raw[:n] == b.raw[:n], (it, s[:n].tobytes())


def _pairs(seed, n):
    rng = np.random.default_rng(seed)
    seqs, pos = [], []
    for i in range(n):
        L = int(rng.choice([0, 1, 5, 40, 300, 1500]))
        a = ALPHA[rng.integers(0, 4, size=L)]
        kind = i % 5
        posoff = int(rng.integers(0, 50))
        if kind == 0 or L == 0:
            b = ALPHA[rng.integers(0, 4, size=int(rng.choice([0, 3, 40, 350])))]        # unrelated (or empty)
        elif kind == 1:
            b = synth.mutate(rng, a, sub=0.02, ins=0.01, dele=0.01)                     # the same insertion, a few errors
        elif kind == 2:
            f = int(rng.integers(0, max(1, L)))
            b = np.concatenate([a[L - f:], a[:L - f]]); posoff = f                      # rotated by the breakpoint offset
        elif kind == 3:
            f = int(rng.integers(0, max(1, L)))
            b = synth.mutate(rng, np.concatenate([a[f:], a[:f]]), sub=0.03, ins=0.0, dele=0.0)   # rotated the other way: only the canonical rotation helps
        else:
            b = synth.mutate(rng, a, sub=0.2, ins=0.05, dele=0.05)                      # diverged beyond the bound
        pos.append(posoff)
        seqs += [a, b]
    arena, off, ln = synth.pack(seqs)
    return arena, off[0::2].copy(), ln[0::2].copy(), off[1::2].copy(), ln[1::2].copy(), np.array(pos, np.int32)


@pytest.mark.parametrize("min_id,cutoff", [(0.9, 1000), (0.7, 2000), (0.0, 2000), (1.0, 100)])
def test_seq_identity_batches_match_reference(hostdev, ref10, min_id, cutoff):
    """Both helpers for 300 pairs: empty sequences, unrelated / near-identical / rotated / diverged pairs around the bound
    k = (1 - minId) * length, with the rotation attempts gated by the sequence cut-off. Exact double equality."""
    H, ctxh = hostdev
    ref10.ref_seq_identity.restype = C.c_double; ref10.ref_best_seq_identity.restype = C.c_double
    arena, ao, al, bo, bl, pos = _pairs(5, 300)
    n = len(ao)
    for best in (0, 1):
        got = np.zeros(n, np.float64)
        rc = H.dh_seq_identity_batch(ctxh, _p(arena), _p(ao), _p(al), _p(bo), _p(bl), _p(pos), n, C.c_double(min_id), cutoff, best, _p(got))
        assert rc == 0, rc
        exp = np.zeros(n, np.float64)
        for i in range(n):
            a = arena[ao[i]:ao[i] + al[i]].tobytes(); b = arena[bo[i]:bo[i] + bl[i]].tobytes()
            exp[i] = (ref10.ref_best_seq_identity(a, len(a), b, len(b), int(pos[i]), C.c_double(min_id), cutoff) if best
                      else ref10.ref_seq_identity(a, len(a), b, len(b), C.c_double(min_id)))
        bad = np.nonzero(exp.view(np.uint64) != got.view(np.uint64))[0]
        assert len(bad) == 0, (best, bad[:5], exp[bad[:5]], got[bad[:5]], al[bad[:5]], bl[bad[:5]])
        if best and 0 < min_id < 1:
            plain = np.array([ref10.ref_seq_identity(arena[ao[i]:ao[i] + al[i]].tobytes(), int(al[i]), arena[bo[i]:bo[i] + bl[i]].tobytes(), int(bl[i]), C.c_double(min_id))
                              for i in range(n)])
            assert (exp > plain).sum() >= 20                                            # the rotation rounds rescue pairs
        assert (exp == -1.0).any() and ((exp > 0.9) & (exp < 1.0)).any()
