"""Reference-based SV annotation (src/svanno.h: breakpoint homology, mobile-element / NUMT / LTR classes, tandem repeats):
the host mirror (delly_b200/host/svanno.hpp, annotateSVBatch) against annotateSV compiled VERBATIM with the reference's own
edlib (oracle/ref_wrap8.cpp). The mobile-element templates are data of the reference and are fetched from that library.
CPU tests cover everything that needs no edit distance (deletions, duplications, inversions, symbolic / short / purely
periodic insertions); the GPU tests add the insertions whose class comes from HW edit distances on the device."""
import ctypes as C

import numpy as np
import pytest

import delly_b200
from delly_b200 import synth

ALPHA = np.frombuffer(b"ACGT", np.uint8)
COMP = {65: 84, 67: 71, 71: 67, 84: 65, 78: 78}


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _templates(ref8):
    seqs = []
    for w in range(1, 8):
        n = ref8.ref_mei_template(w, None, 0)
        b = C.create_string_buffer(n)
        assert ref8.ref_mei_template(w, b, n) == n
        seqs.append(np.frombuffer(b.raw[:n], np.uint8).copy())
    off = np.zeros(8, np.uint32)
    off[1:] = np.cumsum([len(s) for s in seqs])
    return seqs, np.concatenate(seqs), off


def _rc(a):
    return np.array([COMP.get(int(x), int(x)) for x in a[::-1]], np.uint8)


def _noisy(rng, a, rate):
    return synth.mutate(rng, a, sub=rate / 3, ins=rate / 3, dele=rate / 3)


def _chromosome(seed, n=60000):
    """Random chromosome with planted features: tandem-repeat tracts, a duplicated block (breakpoint homology), a soft-masked stretch."""
    rng = np.random.default_rng(seed)
    g = ALPHA[rng.integers(0, 4, size=n)].copy()
    feats = {}
    # tandem repeats: (start, unit length, copies, noise)
    trs = [(5000, 5, 80, 0.0), (9000, 17, 30, 0.03), (14000, 2, 150, 0.0), (20000, 31, 20, 0.05), (26000, 64, 12, 0.02)]
    for (s, ul, cp, noise) in trs:
        unit = ALPHA[rng.integers(0, 4, size=ul)]
        tract = np.tile(unit, cp)
        if noise:
            flip = rng.random(len(tract)) < noise
            tract = np.where(flip, ALPHA[rng.integers(0, 4, size=len(tract))], tract)
        g[s:s + len(tract)] = tract
    feats["trs"] = [(s, ul, ul * cp) for (s, ul, cp, _) in trs]
    # homologous blocks: X ends at a, its (noisy) copy ends at b -> DEL/DUP/INV (a, b) has homology leftwards; same rightwards
    homs = []
    for (a, b, L, noise) in [(31000, 33500, 400, 0.0), (35000, 35900, 300, 0.04), (38000, 44000, 2500, 0.02), (46000, 46400, 150, 0.08)]:
        blk = g[a - L:a].copy()
        flip = rng.random(L) < noise
        g[b - L:b] = np.where(flip, ALPHA[rng.integers(0, 4, size=L)], blk)
        blk2 = g[a + 1:a + 1 + L // 2].copy()
        g[b + 1:b + 1 + L // 2] = blk2
        homs.append((a, b))
    feats["homs"] = homs
    g[31000 - 200:31000 - 100] |= 0x20   # soft-masked (lower case) bases inside a homologous block
    g[5100:5200] |= 0x20
    return g, feats


def _cases(seed, g, feats, tpl=None):
    """[svt, svStart, svEnd], alleles. tpl=None: host-only cases; otherwise the mobile-element / flank-repeat insertions too."""
    rng = np.random.default_rng(seed)
    n = len(g)
    svs, alleles = [], []

    def add(svt, s, e, al=b""):
        svs.append([svt, int(s), int(e)])
        alleles.append(bytes(al))

    # deletions / duplications / inversions: random, planted homology, inside tandem repeats, chromosome ends, > 50 kb
    for _ in range(30):
        s = int(rng.integers(100, n - 3000)); L = int(rng.integers(1, 2500))
        add(int(rng.choice([0, 1, 2, 3])), s, s + L, b"" if rng.random() < 0.5 else b"A,<DEL>")
    for (a, b) in feats["homs"]:
        for svt in (0, 1, 2, 3):
            add(svt, a, b)
            add(svt, a + int(rng.integers(-3, 4)), b + int(rng.integers(-3, 4)))
    for (s, ul, span) in feats["trs"]:
        add(2, s + ul, s + ul + 5 * ul)                  # whole copies deleted
        add(2, s + 3, s + 3 + max(10, 3 * ul + 1))
        add(2, s - 7, s + span // 2)
        add(2, s + 1, s + 1 + 9)                          # 9 bp: below the length gate
        add(3, s + ul, s + ul + 4 * ul)                   # duplication: homology only
    add(2, 1, 40); add(2, 3, 30000); add(2, n - 500, n - 2); add(2, n - 40, n - 1); add(2, 2000, 2000 + 55000); add(3, 0, 500)
    add(2, 700, 700); add(2, 900, 850)                    # svEnd <= svStart: nothing to do
    # insertions without a device distance: symbolic / empty / one base / short / purely periodic ALT; target-site duplication
    add(4, 1200, 1201, b"A,<INS>"); add(4, 1300, 1301, b""); add(4, 1400, 1401, b"AC"); add(4, 1500, 1501, b"T,TACGTAC"); add(4, 1600, 1601, b"noalt")
    for ul, cp in [(1, 30), (3, 10), (7, 9), (25, 3), (49, 2), (2, 49)]:
        unit = ALPHA[rng.integers(0, 4, size=ul)]
        if ul > 1 and len(set(unit.tolist())) == 1:
            unit[0] = ord("A") if unit[0] != ord("A") else ord("C")
        s = int(rng.integers(2000, 4000))
        add(4, s, s + 1, bytes(g[s - 1:s]) + b"," + bytes(g[s - 1:s]) + bytes(np.tile(unit, cp)))
    for tsd in (5, 17, 60, 100, 130):
        s = int(rng.integers(50000, 56000))
        dup = g[s - 1:s - 1 + tsd].copy()
        if tsd >= 17:
            dup[tsd // 2] = ord("A") if dup[tsd // 2] != ord("A") else ord("G")     # one edit inside the duplication
        rep = np.tile(np.frombuffer(b"CA", np.uint8), 20)                           # periodic tail keeps the case host-only (< 100 bp gates aside)
        alt = np.concatenate([dup, rep])[:99 + 1] if tsd < 100 else None
        if alt is not None:
            add(4, s, s + 1, b"n," + bytes(alt).lower())
    if tpl is None:
        return np.array(svs, np.int32), alleles

    # ---- insertions whose class needs HW edit distances
    alu, l1, sva, numt, ltr, hervk, polya = tpl

    def ins(body, s=None):
        s = int(rng.integers(47000, 58000)) if s is None else s
        add(4, s, s + 1, bytes(g[s - 1:s]) + b"," + bytes(g[s - 1:s]) + bytes(body))

    full = [np.concatenate([alu, polya[:30]]), np.concatenate([l1[-1800:], polya[:40]]), np.concatenate([sva[-1500:], polya[:25]]), numt[3000:4200], ltr,
            hervk[2000:3500]]
    for k, body in enumerate(full):
        for rate in (0.0, 0.04, 0.12, 0.3):
            b2 = _noisy(rng, body, rate)
            ins(b2)
            ins(_rc(b2))
    ins(np.concatenate([alu[:150]]))                                             # half an Alu: fraction below / near the threshold
    ins(np.concatenate([ALPHA[rng.integers(0, 4, size=60)], alu, polya[:20], ALPHA[rng.integers(0, 4, size=60)]]))   # insertion longer than the template
    ins(np.concatenate([ALPHA[rng.integers(0, 4, size=200)], _rc(alu), ALPHA[rng.integers(0, 4, size=300)]]))       # coverage-limited
    ins(np.concatenate([l1, polya]))                                              # full-length LINE1
    ins(np.concatenate([numt[:15500]]))                                           # capped at 15000 bases
    ins(_noisy(rng, hervk, 0.05))
    for L in (12, 99, 100, 149, 150, 299, 300, 999, 1000, 2500):                  # unrelated sequence around the per-class length gates
        ins(ALPHA[rng.integers(0, 4, size=L)])
    ins(np.where(rng.random(400) < 0.1, ord("N"), ALPHA[rng.integers(0, 4, size=400)]).astype(np.uint8))
    # flank repeats: the insertion is a noisy run of the repeat unit next to it (autocorrelation of the insertion alone fails)
    for (s, ul, span) in feats["trs"]:
        unit = g[s:s + ul] & 0xDF
        for where in (s + span, s, s + span // 2):                                # right end (left flank repeats), left end, inside
            for L, rate in ((40, 0.25), (90, 0.3), (180, 0.35), (700, 0.2), (60, 0.6)):
                body = _noisy(rng, np.tile(unit, L // ul + 2)[:L], rate)
                if len(body) >= 1:
                    ins(body, where)
    for where in (30, 210, len(g) - 30, len(g) - 210):                            # flank windows clipped by the chromosome ends
        ins(ALPHA[rng.integers(0, 4, size=50)], where)
    return np.array(svs, np.int32), alleles


def _pack_alleles(alleles):
    off = np.zeros(len(alleles) + 1, np.uint32)
    off[1:] = np.cumsum([len(a) for a in alleles])
    blob = np.frombuffer(b"".join(alleles) + b"\0", np.uint8).copy()
    return blob, off


def _run_ref(ref8, g, svs, alleles, mei=0.8, tr=0.85):
    blob, off = _pack_alleles(alleles)
    out = np.zeros((len(svs), 5), np.int32)
    gz = np.concatenate([g, np.zeros(1, np.uint8)])
    ref8.ref_annotate_sv(_p(gz), len(g), _p(svs), len(svs), _p(blob), _p(off), C.c_float(mei), C.c_float(tr), _p(out))
    return out


def _run_host(ctxh, tpl_arena, tpl_off, g, svs, alleles, mei=0.8, tr=0.85):
    H = delly_b200.hostlib()
    blob, off = _pack_alleles(alleles)
    out = np.zeros((len(svs), 5), np.int32)
    gz = np.concatenate([g, np.zeros(1, np.uint8)])
    rc = H.dh_annotate_sv(ctxh, _p(tpl_arena) if tpl_arena is not None else None, _p(tpl_off) if tpl_off is not None else None, _p(gz), len(g), _p(svs), len(svs),
                          _p(blob), _p(off), C.c_float(mei), C.c_float(tr), _p(out))
    return rc, out


# ------------------------------------------------------------------------------------ CPU
def test_detect_tandem_repeat_matches_reference(ref8):
    H = delly_b200.hostlib()
    rng = np.random.default_rng(77)
    H.dh_detect_tandem_repeat.restype = C.c_int
    ref8.ref_detect_tandem_repeat.restype = C.c_int
    hits = 0
    for it in range(400):
        kind = it % 4
        L = int(rng.integers(0, 700))
        if kind == 0:
            s = ALPHA[rng.integers(0, 4, size=L)]
        else:
            ul = int(rng.integers(1, 130))
            s = np.tile(ALPHA[rng.integers(0, 4, size=ul)], L // ul + 1)[:L].copy()
            if kind >= 2 and L:
                flip = rng.random(L) < (0.1 if kind == 2 else 0.2)
                s[flip] = ALPHA[rng.integers(0, 4, size=int(flip.sum()))]
        s = np.concatenate([s, np.zeros(1, np.uint8)]).astype(np.uint8)
        for frac, maxp in ((0.85, 100), (0.7, 100), (0.95, 40)):
            c1, c2 = C.c_float(), C.c_float()
            p1 = ref8.ref_detect_tandem_repeat(_p(s), L, maxp, C.c_float(frac), C.byref(c1))
            p2 = H.dh_detect_tandem_repeat(_p(s), L, maxp, C.c_float(frac), C.byref(c2))
            assert (p1, c1.value) == (p2, c2.value), (it, L, frac, maxp)
            hits += p1 > 0
    assert hits > 200


@pytest.mark.parametrize("seed", [3, 4])
def test_annotate_sv_host_only_cases_match_reference(ref8, seed):
    """Deletions, duplications, inversions and the insertions that need no edit distance: homology lengths, DEL:TR / INS:TR periods and
    copy numbers identical to annotateSV. No device context is passed: these cases make no device call."""
    g, feats = _chromosome(seed)
    svs, alleles = _cases(seed + 10, g, feats, None)
    e = _run_ref(ref8, g, svs, alleles)
    rc, got = _run_host(None, None, None, g, svs, alleles)
    assert rc == 0
    bad = np.nonzero((e != got).any(axis=1))[0]
    assert len(bad) == 0, (bad[:5], svs[bad[:5]], e[bad[:5]], got[bad[:5]])
    assert (e[:, 2] > 100).sum() >= 8 and (e[:, 2] > 0).sum() > 30          # planted homologies are found
    assert ((e[:, 1] == 7) & (svs[:, 0] == 2)).sum() >= 8                  # DEL:TR
    assert ((e[:, 1] == 7) & (svs[:, 0] == 4)).sum() >= 5                  # INS:TR by autocorrelation
    assert len(np.unique(e[e[:, 1] == 7][:, 3])) >= 6                       # several distinct periods
    # other thresholds
    e2 = _run_ref(ref8, g, svs, alleles, 0.5, 0.6)
    rc, g2 = _run_host(None, None, None, g, svs, alleles, 0.5, 0.6)
    assert rc == 0 and np.array_equal(e2, g2) and not np.array_equal(e, e2)


def test_annotate_sv_without_device_fails_loudly(ref8):
    """An insertion that needs template distances and no device context: an error, never a host computation."""
    g, feats = _chromosome(5)
    seqs, arena, off = _templates(ref8)
    svs = np.array([[4, 1000, 1001]], np.int32)
    rc, _ = _run_host(None, arena, off, g, svs, [b"A,A" + bytes(seqs[0])])
    assert rc == -3  # DGPU_ERR_NODEVICE


@pytest.mark.parametrize("seed", [11, 12])
def test_annotate_sv_host_logic_with_reference_distances(standin, ref8, seed):
    """The complete case set of the GPU test below (mobile elements on both strands, length gates, flank repeats ...) with the device call
    replaced by the reference's edlib at link time: job construction, query / target swap, class selection and the flank-repeat round
    of annotateSVBatch give annotateSV's result for every SV."""
    g, feats = _chromosome(seed)
    seqs, arena, off = _templates(ref8)
    svs, alleles = _cases(seed + 10, g, feats, seqs)
    blob, aoff = _pack_alleles(alleles)
    gz = np.concatenate([g, np.zeros(1, np.uint8)])
    for mei, tr in ((0.8, 0.85), (0.55, 0.85), (0.3, 0.6)):
        e = _run_ref(ref8, g, svs, alleles, mei, tr)
        got = np.zeros((len(svs), 5), np.int32)
        rc = standin.dh_annotate_sv(C.c_void_p(standin.standin_ctx()), _p(arena), _p(off), _p(gz), len(g), _p(svs), len(svs), _p(blob), _p(aoff), C.c_float(mei),
                                    C.c_float(tr), _p(got))
        assert rc == 0
        bad = np.nonzero((e != got).any(axis=1))[0]
        assert len(bad) == 0, (mei, tr, bad[:5], svs[bad[:5]], e[bad[:5]], got[bad[:5]])
        if mei == 0.8:
            for t in range(1, 7):
                assert (e[:, 1] == t).sum() >= 2, t
            assert ((e[:, 1] == 7) & (svs[:, 0] == 4)).sum() >= 12 and ((e[:, 1] == 0) & (svs[:, 0] == 4)).sum() >= 10


# ------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("seed", [11, 12])
def test_annotate_sv_matches_reference(ctx, ref8, seed):
    """All SV kinds of one chromosome in one batch, incl. mobile-element / NUMT / LTR / HERVK insertions on both strands at several
    divergences, insertions longer than the template, unrelated sequence around the per-class length gates, and noisy repeat
    expansions next to reference repeats (flank-repeat round)."""
    g, feats = _chromosome(seed)
    seqs, arena, off = _templates(ref8)
    svs, alleles = _cases(seed + 10, g, feats, seqs)
    e = _run_ref(ref8, g, svs, alleles)
    rc, got = _run_host(ctx.h, arena, off, g, svs, alleles)
    assert rc == 0, rc
    bad = np.nonzero((e != got).any(axis=1))[0]
    assert len(bad) == 0, (bad[:5], svs[bad[:5]], e[bad[:5]], got[bad[:5]], [len(alleles[i]) for i in bad[:5]])
    for t in range(1, 7):
        assert (e[:, 1] == t).sum() >= 2, t                                 # every class is called
    assert ((e[:, 1] >= 1) & (e[:, 1] <= 6) & (e[:, 0] == 1)).sum() >= 6    # reverse strand
    assert ((e[:, 1] >= 1) & (e[:, 1] <= 6) & (e[:, 0] == 0)).sum() >= 6
    assert ((e[:, 1] == 7) & (svs[:, 0] == 4)).sum() >= 12                  # INS:TR incl. the flank-repeat round
    assert ((e[:, 1] == 0) & (svs[:, 0] == 4)).sum() >= 10                  # unclassified insertions
    # a lower mobile-element threshold reclassifies the diverged copies
    e2 = _run_ref(ref8, g, svs, alleles, 0.55, 0.85)
    rc, g2 = _run_host(ctx.h, arena, off, g, svs, alleles, 0.55, 0.85)
    assert rc == 0 and np.array_equal(e2, g2) and not np.array_equal(e, e2)
