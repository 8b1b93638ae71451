"""CpG methylation around SV breakpoints from MM / ML tags (src/methyl.h as used by genotypeLR, src/genotype.h:306-324,383-388):
the host mirror (delly_b200/host/methyl.hpp inside genotypeLRBatch) against genotypeLR compiled VERBATIM over in-memory records that
carry the tags (oracle/ref_wrap4.cpp::ref_genotype_lr_methyl).
CPU: the host logic with the two device entry points it uses (edit distance, edit path) forwarded to the reference's edlib at link time
(tests/standin/host_standin.cpp — test infrastructure, see its header). GPU: the same comparison through the real libraries."""
import ctypes as C

import numpy as np
import pytest

import delly_b200
from test_host_genotype import _lr_geno_case

COMP = {65: 84, 67: 71, 71: 67, 84: 65}


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _with_tags(d, seed):
    """MM / ML tags for the records of an _lr_geno_case: mostly 'C+m?' / 'C+m.' / 'C+m' lists over a random subset of the read's cytosines
    (original read orientation), plus the shapes the parser has to get right: two interleaved codes (C+mh), the reverse-strand notation (G-m),
    an unrelated modification first (A+a: it consumes ML entries), a ChEBI number instead of a letter code, an ML array shorter than the
    hit list, no ML tag, an MM tag without positions, an empty MM tag, no MM tag."""
    rng = np.random.default_rng(seed)
    rec, reads = d["rec"], d["reads"]
    flags = np.zeros(len(rec), np.uint8)
    mms, mls = [], []
    for i, r in enumerate(rec):
        seq = reads[r[7]:r[7] + r[4]]
        if r[2] & 16:
            seq = np.array([COMP.get(int(x), int(x)) for x in seq[::-1]], np.uint8)
        kind = int(rng.choice(10, p=[0.42, 0.12, 0.1, 0.08, 0.06, 0.05, 0.04, 0.04, 0.04, 0.05]))
        mm, ml = b"", []

        def listed(base, frac):
            occ = np.nonzero(seq == ord(base))[0]
            k = int(len(occ) * frac)
            if k == 0:
                return []
            ranks = np.sort(rng.choice(len(occ), size=k, replace=False))
            return np.diff(np.concatenate([[-1], ranks])) - 1

        def probs(n):
            return list(rng.choice([0, 10, 127, 128, 200, 255], size=n, p=[0.15, 0.2, 0.1, 0.1, 0.25, 0.2]))

        if kind == 0 or kind == 1 or kind == 2:
            deltas = listed("C", 0.7 if kind != 1 else 0.3)
            head = [b"C+m?", b"C+m.", b"C+m"][kind]
            mm = head + b"".join(b",%d" % x for x in deltas) + b";"
            ml = probs(len(deltas))
        elif kind == 3:   # two codes on the same positions: the hit list interleaves them
            deltas = listed("C", 0.5)
            mm = b"C+mh?" + b"".join(b",%d" % x for x in deltas) + b";"
            ml = probs(2 * len(deltas))
        elif kind == 4:   # reverse-strand notation: G-m counts the read's cytosines through the complement
            deltas = listed("C", 0.6)
            mm = b"G-m?" + b"".join(b",%d" % x for x in deltas) + b";"
            ml = probs(len(deltas))
        elif kind == 5:   # another modification first, then 5mC
            da = listed("A", 0.1); dc = listed("C", 0.6)
            mm = b"A+a?" + b"".join(b",%d" % x for x in da) + b";C+m?" + b"".join(b",%d" % x for x in dc) + b";"
            ml = probs(len(da) + len(dc))
        elif kind == 6:   # ChEBI code (no letter): no hits from this token; ML shorter than the hit list of the next
            dc = listed("C", 0.6)
            mm = b"C+76792?,1,2;C+m," + b",".join(b"%d" % x for x in dc) + b";"
            ml = probs(max(0, len(dc) - 5))
        elif kind == 7:   # no ML tag: every listed position counts as methylated
            deltas = listed("C", 0.4)
            mm = b"C+m?" + b"".join(b",%d" % x for x in deltas) + b";"
            ml = None
        elif kind == 8:   # MM without positions / empty MM
            mm = b"C+m?;" if rng.random() < 0.5 else b""
            ml = []
        else:
            mm = None
        if mm is not None:
            flags[i] |= 1
        if mm is not None and ml is not None:
            flags[i] |= 2
        mms.append(mm or b""); mls.append(np.array(ml if ml is not None else [], np.uint8))
    mm_off = np.zeros(len(rec) + 1, np.uint32); mm_off[1:] = np.cumsum([len(x) for x in mms])
    ml_off = np.zeros(len(rec) + 1, np.uint32); ml_off[1:] = np.cumsum([len(x) for x in mls])
    d = dict(d)
    d.update(tagflags=flags, mm=np.frombuffer(b"".join(mms) + b"\0", np.uint8).copy(), mm_off=mm_off,
             ml=np.concatenate(mls + [np.zeros(1, np.uint8)]), ml_off=ml_off)
    return d


def _call(fn, lead, d, cap, geno_cap, window, prob, depth, tags=True):
    nsv, nrec = len(d["sv"]), len(d["rec"])
    q = np.zeros(400000, np.uint8); ro = np.zeros(nsv + 1, np.uint32); ao = np.zeros(nsv + 1, np.uint32)
    hp = np.zeros((nsv, 5), np.int32); rc = np.zeros((nsv, 3), np.int32); me = np.full((nsv, 16), -7, np.int32)
    args = [_p(d["contig"]), _p(d["coff"]), _p(d["clen"]), 2, _p(d["rec"]), nrec, _p(d["cig"]), _p(d["reads"]), _p(d["sv"]), nsv, _p(d["cons"]), _p(d["co"]),
            _p(d["cl"]), 1, 100, 300, cap, C.c_float(0.9), geno_cap, _p(q), len(q), _p(ro), _p(ao), _p(hp), _p(rc)]
    if tags:
        args += [_p(d["tagflags"]), _p(d["mm"]), _p(d["mm_off"]), _p(d["ml"]), _p(d["ml_off"]), window, prob, depth, _p(me)]
    r = fn(*lead, *args)
    assert r >= 0, r
    return r, q[:r].copy(), ro, ao, hp, rc, me


def _check(e, g, d, window):
    assert e[0] == g[0]
    for k in range(1, 6):
        assert np.array_equal(e[k], g[k]), k
    bad = np.nonzero((e[6] != g[6]).any(axis=1))[0]
    assert len(bad) == 0, (bad[:4], d["sv"][bad[:4]], e[6][bad[:4]], g[6][bad[:4]])
    me, sv = e[6], d["sv"]
    assert (me != -7).all()
    called = (me[:, 0:8] >= 0).sum()
    assert called > 4 * len(sv) // 2                                        # percentages on both alleles of most SVs
    ins = (sv[:, 4] == 4) & (sv[:, 5] >= 40)
    assert ((me[ins, 1] >= 0) | (me[ins, 2] >= 0)).sum() >= max(1, ins.sum() // 2)   # calls inside insertions (edit-path placement)
    assert len(np.unique(me[:, 0:8][me[:, 0:8] >= 0])) > 10               # a spread of percentages
    assert (me[:, 8:12] > 0).any() and (me[:, 12:16] > 0).any()


# ------------------------------------------------------------------------------------ CPU
def test_genotype_lr_host_logic_with_reference_edlib(standin, ref4):
    """genotypeLRBatch without tags through the stand-in: the CPU-side counterpart of test_genotype_lr_matches_reference."""
    d = _lr_geno_case(4242)
    e = _call(ref4.ref_genotype_lr, (), d, 250, 60, 0, 0, 0, tags=False)
    g = _call(standin.dh_genotype_lr, (C.c_void_p(standin.standin_ctx()),), d, 250, 60, 0, 0, 0, tags=False)
    assert e[0] == g[0] and e[0] > len(d["sv"]) * 4
    for k in range(1, 6):
        assert np.array_equal(e[k], g[k]), k


@pytest.mark.parametrize("seed,window,prob,depth,cap", [(4242, 300, 128, 1, 250), (77, 1000, 200, 3, 250), (78, 60, 256, 2, 12), (79, 150, 1, 1, 250)])
def test_genotype_lr_methyl_host_logic_with_reference_edlib(standin, ref4, seed, window, prob, depth, cap):
    """Methylation windows of every SV type, tag shapes incl. the odd ones, insertion placement through the edit path — identical
    MethylInfo (and unchanged genotyping outputs) for several window sizes / probability thresholds (256 wraps to 0 like the
    reference's uint8 cast) / CpG depth gates / read caps."""
    d = _with_tags(_lr_geno_case(seed), seed + 1)
    e = _call(ref4.ref_genotype_lr_methyl, (), d, cap, 60, window, prob, depth)
    g = _call(standin.dh_genotype_lr_methyl, (C.c_void_p(standin.standin_ctx()),), d, cap, 60, window, prob, depth)
    _check(e, g, d, window)


def test_methyl_no_tags_gives_no_calls(standin, ref4):
    d = _with_tags(_lr_geno_case(5), 6)
    d["tagflags"][:] = 0
    e = _call(ref4.ref_genotype_lr_methyl, (), d, 250, 60, 300, 128, 1)
    g = _call(standin.dh_genotype_lr_methyl, (C.c_void_p(standin.standin_ctx()),), d, 250, 60, 300, 128, 1)
    assert np.array_equal(e[6], g[6])
    assert (e[6][:, 0:8] == -1).all() and (e[6][:, 8:12] == 0).all() and (e[6][:, 12:16] == -1).all()


# ------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("seed,window,prob,depth", [(4242, 300, 128, 1), (77, 1000, 200, 3)])
def test_genotype_lr_methyl_matches_reference(ctx, ref4, seed, window, prob, depth):
    """The same comparison through the real libraries: NW distances and HW edit paths on the device."""
    H = delly_b200.hostlib()
    d = _with_tags(_lr_geno_case(seed), seed + 1)
    e = _call(ref4.ref_genotype_lr_methyl, (), d, 250, 60, window, prob, depth)
    g = _call(H.dh_genotype_lr_methyl, (ctx.h,), d, 250, 60, window, prob, depth)
    _check(e, g, d, window)
