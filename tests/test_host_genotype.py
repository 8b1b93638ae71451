"""Genotyping pass, host side (delly_b200/host/gl.hpp, genotype.hpp) against the reference compiled verbatim
(oracle/_ref/libdelly_ref3.so: src/bolog.h _computeGLs, src/coverage.h _generateProbes).
CPU tests: GT / GL / GQ bit-for-bit, PL / phasing / RCN / FT against a restatement of src/modvcf.h:671-715.
GPU test: generateProbesBatch (one device alignment batch for all precise SVs) against _generateProbes."""
import ctypes as C

import numpy as np
import pytest

import delly_b200
from delly_b200 import synth
from test_host_split import _genome, _p, _sv_cases


def _support_cases(seed, n):
    """(ref qualities, alt qualities) per sample: empty, one-sided, balanced, deep, saturated qualities."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        kind = i % 8
        nr, na = int(rng.integers(0, 40)), int(rng.integers(0, 40))
        if kind == 0: nr, na = 0, 0
        elif kind == 1: na = 0
        elif kind == 2: nr = 0
        elif kind == 3: nr, na = int(rng.integers(200, 400)), int(rng.integers(200, 400))
        hi = 256 if kind in (4, 5) else 61
        r = rng.integers(0, hi, size=nr).astype(np.uint8)
        a = rng.integers(0, hi, size=na).astype(np.uint8)
        if kind == 6: r[:] = 0
        if kind == 7 and na: a[:] = 255
        out.append((r, a))
    return out


def test_compute_gls_matches_reference(ref3):
    H = delly_b200.hostlib()
    for r, a in _support_cases(1, 600):
        eg = np.zeros(3, np.float32); eq = np.zeros(1, np.int32); et = np.zeros(2, np.int32)
        gg = np.zeros(3, np.float32); gq = np.zeros(1, np.int32); gt = np.zeros(2, np.int32)
        rr = np.ascontiguousarray(np.concatenate([r, [0]]).astype(np.uint8)); aa = np.ascontiguousarray(np.concatenate([a, [0]]).astype(np.uint8))
        ref3.ref_compute_gls(_p(rr), len(r), _p(aa), len(a), _p(eg), _p(eq), _p(et))
        H.dh_compute_gls(_p(rr), len(r), _p(aa), len(a), _p(gg), _p(gq), _p(gt))
        assert np.array_equal(eg.view(np.uint32), gg.view(np.uint32)), (r, a, eg, gg)   # bit pattern, not tolerance
        assert eq[0] == gq[0] and np.array_equal(et, gt)


def test_sample_format_matches_restatement(ref3):
    """GT/GL/GQ from the compiled reference; PL, het phasing, RCN and FT per src/modvcf.h:671-715 restated here."""
    H = delly_b200.hostlib()
    cases = _support_cases(2, 400)
    rng = np.random.default_rng(3)
    n = len(cases)
    extra = np.zeros((n, 6), np.int32)
    extra[:, 0] = np.where(rng.random(n) < 0.5, -1, rng.integers(1, 10 ** 6, size=n))
    extra[:, 1] = rng.integers(0, 6, size=n); extra[:, 2] = rng.integers(0, 6, size=n)
    extra[:, 3] = rng.integers(0, 500, size=n) * (rng.random(n) < 0.9); extra[:, 4] = rng.integers(0, 10 ** 6, size=n)
    extra[:, 5] = rng.integers(0, 500, size=n) * (rng.random(n) < 0.9)
    extra[0, 3:] = (1, 10 ** 9, 0)  # copy number clamped at 100000
    refq = np.concatenate([c[0] for c in cases] + [np.zeros(1, np.uint8)]); altq = np.concatenate([c[1] for c in cases] + [np.zeros(1, np.uint8)])
    ro = np.concatenate([[0], np.cumsum([len(c[0]) for c in cases])]).astype(np.uint32)
    ao = np.concatenate([[0], np.cumsum([len(c[1]) for c in cases])]).astype(np.uint32)
    out = np.zeros((n, 9), np.int32); gls = np.zeros((n, 3), np.float32)
    H.dh_sample_format(n, _p(refq), _p(ro), _p(altq), _p(ao), _p(extra), _p(out), _p(gls))
    UN0, UN1, PH0, PH1 = 2, 4, 3, 5   # htslib: bcf_gt_unphased(a) = (a+1)<<1, phased |1
    for i, (r, a) in enumerate(cases):
        eg = np.zeros(3, np.float32); eq = np.zeros(1, np.int32); et = np.zeros(2, np.int32)
        rr = np.concatenate([r, [0]]).astype(np.uint8); aa = np.concatenate([a, [0]]).astype(np.uint8)
        ref3.ref_compute_gls(_p(rr), len(r), _p(aa), len(a), _p(eg), _p(eq), _p(et))
        missing = et[0] == 0
        pl = [np.iinfo(np.int32).min] * 3 if missing else [int(max(np.float32(0), np.round(np.float32(-10) * eg[k]))) for k in range(3)]
        gt = et.tolist()
        ps, h1, h2, rcl, rc, rcr = extra[i].tolist()
        if ps != -1 and gt == [UN0, UN1] and (h1 + h2) > 0 and h1 != h2:
            gt = [PH1, PH0] if h1 > h2 else [PH0, PH1]
        rcn = -1
        if rcl + rcr > 0:
            cn = min(max(2.0 * rc / (rcl + rcr), 0.0), 100000.0)
            rcn = int(np.floor(cn + 0.5))
        exp = gt + [int(eq[0])] + pl + [rcn, 0 if eq[0] < 15 else 1, 1 if missing else 0]
        assert out[i].tolist() == exp, (i, out[i].tolist(), exp)
        assert np.array_equal(gls[i].view(np.uint32), eg.view(np.uint32))


@pytest.mark.parametrize("mode", ["sr", "lr"])
def test_generate_probes_matches_reference(hostdev, ref, ref3, mode):
    """_generateProbes (src/coverage.h:164-263): REF/ALT probes of both breakpoints, breakpoint regions (per contig, in the
    reference's order), symbolic alleles — for SVs refined by the reference's own alignConsensus, precise and imprecise."""
    H, ctxh = hostdev
    g1, g2 = _genome(21), _genome(22)
    fq, mfs, indel, mcw = (0.95, 13, 1000, 100) if mode == "sr" else (0.9, 30, 10000, 300)
    svs, cons = _sv_cases(31, g1, g2, n=220, cons_range=(80, 260) if mode == "sr" else (200, 500), with_ins=True)
    n = len(svs)
    rec = np.zeros((n, 8), np.int32)
    cons_final = []
    for i in range(n):
        sv = svs[i]
        seq, snd = (g1, g1) if sv[4] < 5 else (g2, g1)
        out = np.zeros(10, np.int32); srq = C.c_float(); al = C.create_string_buffer(8192); all_ = C.c_int()
        co = C.create_string_buffer(4096); col = C.c_int()
        cb = cons[i].tobytes()
        ok = ref.ref_align_consensus(seq, len(g1), snd, len(g2), _p(sv.copy()), cb, len(cb), 0, C.c_float(fq), mfs, indel, mcw, _p(out),
                                     C.byref(srq), al, 8192, C.byref(all_), co, C.byref(col))
        # the discovery pass leaves refined coordinates on precise SVs and the cluster's coordinates on the others
        rec[i] = [sv[0], out[0] if ok else sv[1], sv[2], out[1] if ok else sv[3], sv[4], out[2] if ok else sv[5], 1 if ok else 0, i]
        cons_final.append(co.raw[:col.value])
    assert rec[:, 6].sum() > n // 3
    arena, off, ln = synth.pack([np.frombuffer(c, np.uint8) for c in cons_final])
    # contig 0 = g1 ("chrA"), contig 1 = g2 ("chrB"): translocations have chr = 1, chr2 = 0
    def run(fn, *lead):
        pa = np.zeros(4 << 20, np.uint8); po = np.zeros((n, 4), np.uint64); pl = np.zeros((n, 4), np.uint32)
        reg = np.zeros((4 * n, 9), np.int32); al = np.zeros((n, 256), np.uint8); all_ = np.zeros(n, np.int32); on = np.zeros(2, np.uint8)
        r = fn(*lead, g1, len(g1), g2, len(g2), n, _p(rec), _p(arena), _p(off), _p(ln), C.c_float(fq), mfs, indel, mcw, _p(pa), C.c_uint64(len(pa)),
               _p(po), _p(pl), _p(reg), 4 * n, _p(al), _p(all_), _p(on))
        assert r >= 0, r
        probes = [[pa[int(po[i, k]):int(po[i, k]) + int(pl[i, k])].tobytes() for k in range(4)] for i in range(n)]
        return r, probes, reg[:r].tolist(), [al[i, :all_[i]].tobytes() for i in range(n)], on.tolist()
    er, eprobes, ereg, eal, eon = run(ref3.ref_generate_probes)
    gr, gprobes, greg, gal, gon = run(H.dh_generate_probes, ctxh)
    assert er == gr and er > n // 2
    assert eprobes == gprobes
    assert ereg == greg
    assert eal == gal and eon == gon
    assert sum(1 for p in eprobes if p[0]) > n // 4


# ---- long-read genotyping pass (genotypeLR, src/genotype.h:93-397) ----------------------------------------------------

def _lr_geno_case(seed, nsv=40, reads_per_bp=6):
    """Two contigs, SVs of every type with a junction-spanning consensus, and reads: REF-like over either breakpoint, ALT-like
    (aligned with the event in the CIGAR or soft-clipped at it), unrelated ones; some secondary / supplementary / duplicate /
    mapq-0 records, HP/PS tags on a third of the reads. Returns everything in the flat layout of ref_genotype_lr."""
    rng = np.random.default_rng(seed)
    G = [synth.random_genome(rng, 90000), synth.random_genome(rng, 70000)]
    G[0][30000:30030] = ord("N")
    L = 700
    svs, cons, recs, cigs, reads = [], [], [], [], []

    def add_read(tid, pos, ops, seq, flag=None):
        if flag is None:
            flag = int(rng.choice([0, 16, 0, 16, 2048, 256, 1024], p=[0.4, 0.4, 0.05, 0.05, 0.04, 0.03, 0.03]))
        hp = int(rng.choice([0, 0, 1, 2])); ps = int(rng.integers(1, 1000)) if hp and rng.random() < 0.8 else -1
        recs.append([tid, pos, flag, int(rng.choice([0, 20, 60], p=[0.05, 0.15, 0.8])), len(seq), len(cigs), len(ops), sum(len(r) for r in reads), hp, ps])
        cigs.extend((ln << 4) | op for op, ln in ops)
        reads.append(seq)

    def noisy(a):
        return synth.sub_noise(rng, a.copy(), 0.04)

    for i in range(nsv):
        svt = int(rng.choice([0, 1, 2, 3, 4, 5, 6, 7, 8]))
        s = int(rng.integers(5000, 50000)); size = int(rng.integers(1, 9)) if rng.random() < 0.3 else int(rng.integers(60, 3000)); e = s + size
        if svt >= 5:
            c1, c2 = 1, 0
            A, B = G[1], G[0]
            e = int(rng.integers(5000, 50000))
        else:
            c1 = c2 = 0
            A = B = G[0]
        ins = 0
        if svt == 2: c = np.concatenate([A[s - L:s], B[e:e + L]])
        elif svt == 3: c = np.concatenate([B[e - L:e], A[s:s + L]])
        elif svt == 4:
            ins = int(rng.integers(1, 9)) if size < 10 else int(rng.integers(40, 300)); e = s + 1
            c = np.concatenate([A[s - L:s], synth._ACGT[rng.integers(0, 4, size=ins)], A[s:s + L]])
        elif svt in (0, 5): c = np.concatenate([A[s - L:s], synth.revcomp(B[e - L:e])])
        elif svt in (1, 6): c = np.concatenate([synth.revcomp(A[s:s + L]), B[e:e + L]])
        elif svt == 7: c = np.concatenate([A[s - L:s], B[e:e + L]])
        else: c = np.concatenate([B[e - L:e], A[s:s + L]])
        svs.append([c1, s, c2, e, svt, ins, L, i]); cons.append(c.astype(np.uint8))
        for _ in range(reads_per_bp):
            # REF-like reads over each breakpoint
            for tid, p in ((c1, s), (c2, e)):
                a = int(rng.integers(250, 900)); b = int(rng.integers(250, 900))
                add_read(tid, p - a, [(0, a + b)], noisy(G[tid][p - a:p + b]))
            # ALT-like read: a consensus window around the junction
            a = int(rng.integers(250, 650)); b = int(rng.integers(250, 650))
            seq = noisy(c[L - a:L + ins + b])
            if svt == 2: ops = [(0, a), (2, size), (0, b)]
            elif svt == 4: ops = [(0, a), (1, ins), (0, b)]
            else: ops = [(0, a), (4, len(seq) - a)]
            add_read(c1, s - a, ops, seq)
            if rng.random() < 0.3:  # a leading soft clip before the breakpoint
                add_read(c1, s - a + 40, [(4, 40), (0, a - 40), (4, len(seq) - a)], seq)
    for _ in range(300):   # unrelated reads
        tid = int(rng.integers(0, 2)); p = int(rng.integers(1000, 60000)); ln = int(rng.integers(150, 2500))
        add_read(tid, p, [(0, ln)], noisy(G[tid][p:p + ln]))
    rec = np.array(recs, np.int64)
    order = np.lexsort((np.arange(len(rec)), rec[:, 1], rec[:, 0]))
    rec = np.ascontiguousarray(rec[order].astype(np.int32))
    contig = np.concatenate(G); coff = np.array([0, len(G[0])], np.uint32); clen = np.array([len(G[0]), len(G[1])], np.uint32)
    carena, co, cl = synth.pack(cons)
    return dict(contig=contig, coff=coff, clen=clen, rec=rec, cig=np.array(cigs, np.uint32), reads=np.concatenate(reads).astype(np.uint8),
                sv=np.array(svs, np.int32), cons=carena, co=co.astype(np.uint32), cl=cl.astype(np.uint32))


@pytest.mark.parametrize("cap", [250, 5])
def test_genotype_lr_matches_reference(hostdev, ref4, cap):
    geno_cap = 60 if cap == 250 else 25   # 60: the per-read quality formula is visible below the cap (tiny events give small deltas)
    """The whole long-read genotyping pass against genotypeLR run verbatim over the same in-memory alignments: per-SV REF/ALT
    quality lists (order included), haplotype counts and phase set, read-depth of the SV body and flanks."""
    H, ctxh = hostdev
    d = _lr_geno_case(4242)
    nsv, nrec = len(d["sv"]), len(d["rec"])
    outs = []
    for fn, lead in ((ref4.ref_genotype_lr, ()), (H.dh_genotype_lr, (ctxh,))):
        q = np.zeros(200000, np.uint8); ro = np.zeros(nsv + 1, np.uint32); ao = np.zeros(nsv + 1, np.uint32)
        hp = np.zeros((nsv, 5), np.int32); rc = np.zeros((nsv, 3), np.int32)
        r = fn(*lead, _p(d["contig"]), _p(d["coff"]), _p(d["clen"]), 2, _p(d["rec"]), nrec, _p(d["cig"]), _p(d["reads"]), _p(d["sv"]), nsv, _p(d["cons"]),
               _p(d["co"]), _p(d["cl"]), 1, 100, 300, cap, C.c_float(0.9), geno_cap, _p(q), len(q), _p(ro), _p(ao), _p(hp), _p(rc))
        assert r >= 0, r
        outs.append((r, q[:r].copy(), ro.copy(), ao.copy(), hp.copy(), rc.copy()))
    e, g = outs
    assert e[0] == g[0] and e[0] > (nsv * 4 if cap > 5 else nsv)
    for k in range(1, 6):
        assert np.array_equal(e[k], g[k]), k
    assert (e[4][:, :4].sum() > 0) and (e[5].sum() > 0)
    nref = int(e[2][nsv] - e[2][0]); nalt = int(e[3][nsv] - e[3][0])
    assert nref > nsv and (nalt > nsv // 2 or cap == 5)    # both alleles are genotyped
    if cap == 250:
        assert len(np.unique(e[1])) >= 3                        # qualities below the cap occur


# ---- short-read genotyping pass, junction reads (annotateCoverage, src/coverage.h:265-548) -------------------------------

def _sr_geno_case(ref, seed, nsv, reads_per_sv):
    """SVs refined by the reference's own alignConsensus, then 150 bp reads over their breakpoints: REF-like (plain 150M), ALT-like
    (consensus window, soft-clipped at the junction on either side), unrelated; duplicates / secondary / mate-unmapped / low mapq mixed in."""
    rng = np.random.default_rng(seed)
    g1, g2 = _genome(seed + 1), _genome(seed + 2)
    G = [np.frombuffer(g1, np.uint8), np.frombuffer(g2, np.uint8)]
    svs, cons = _sv_cases(seed + 3, g1, g2, n=nsv, cons_range=(120, 260), with_ins=True)
    rows, cons_final = [], []
    for i in range(len(svs)):
        sv = svs[i]
        seq, snd = (g1, g1) if sv[4] < 5 else (g2, g1)
        out = np.zeros(10, np.int32); srq = C.c_float(); al = C.create_string_buffer(8192); all_ = C.c_int()
        co = C.create_string_buffer(4096); col = C.c_int()
        cb = cons[i].tobytes()
        ok = ref.ref_align_consensus(seq, len(g1), snd, len(g2), _p(sv.copy()), cb, len(cb), 0, C.c_float(0.95), 13, 1000, 100, _p(out), C.byref(srq), al, 8192,
                                     C.byref(all_), co, C.byref(col))
        rows.append([sv[0], out[0] if ok else sv[1], sv[2], out[1] if ok else sv[3], sv[4], out[2] if ok else sv[5], 1 if ok else 0, i, int(rng.integers(0, 3)),
                     int(out[3])])
        cons_final.append(np.frombuffer(co.raw[:col.value], np.uint8))
    rows = np.array(rows, np.int32)
    recs, cigs, reads = [], [], []
    RL = 150

    def add(tid, pos, ops, seq):
        flag = int(rng.choice([0, 16, 1, 17, 1024, 256, 2048, 9], p=[0.3, 0.3, 0.15, 0.15, 0.03, 0.03, 0.02, 0.02]))
        recs.append([tid, pos, flag, int(rng.choice([0, 3, 20, 60], p=[0.03, 0.04, 0.2, 0.73])), len(seq), len(cigs), len(ops), RL * len(reads), tid,
                     max(0, pos + int(rng.integers(-400, 400))), 0, len(recs)])
        cigs.extend((ln << 4) | op for op, ln in ops)
        reads.append(seq)

    for r in rows:
        chr_, s, chr2, e, svt, ins, precise, i, _, cbp = [int(x) for x in r]
        c = cons_final[i]
        for _ in range(reads_per_sv):
            k = rng.random()
            if k < 0.45:      # REF-like over one of the breakpoints
                tid, p = (chr_, s) if rng.random() < 0.5 else (chr2, e)
                a = int(rng.integers(20, RL - 20))
                if p - a < 0 or p - a + RL > len(G[tid]): continue
                add(tid, p - a, [(0, RL)], synth.sub_noise(rng, np.char.upper(G[tid][p - a:p - a + RL].view("S1")).view(np.uint8).copy(), 0.01))
            elif k < 0.9 and precise and len(c) >= RL + 10:   # ALT-like: a consensus window over the junction, clipped at it
                a = int(rng.integers(20, RL - 20))
                st = min(max(cbp - a, 0), len(c) - RL)
                seq = synth.sub_noise(rng, c[st:st + RL].copy(), 0.01)
                a = cbp - st
                if rng.random() < 0.5: add(chr_, max(s - a, 0), [(0, max(a, 1)), (4, RL - max(a, 1))], seq)
                else: add(chr2, e, [(4, max(a, 1)), (0, RL - max(a, 1))], seq)
            else:
                tid = int(rng.integers(0, 2)); p = int(rng.integers(100, 50000))
                add(tid, p, [(0, RL)], np.char.upper(G[tid][p:p + RL].view("S1")).view(np.uint8).copy())
    rec = np.array(recs, np.int64)
    order = np.lexsort((np.arange(len(rec)), rec[:, 1], rec[:, 0]))
    rec = np.ascontiguousarray(rec[order].astype(np.int32))
    carena, co_, cl_ = synth.pack(cons_final)
    return dict(g1=g1, g2=g2, rec=rec, cig=np.array(cigs, np.uint32), reads=np.concatenate(reads).astype(np.uint8), sv=np.ascontiguousarray(rows[:, :9]),
                cons=carena, co=co_.astype(np.uint32), cl=cl_.astype(np.uint32))


@pytest.mark.parametrize("shape", [(120, 40, 250), (120, 40, 6), (60, 3600, 250)])
def test_annotate_junction_reads_matches_reference(hostdev, ref, ref3, shape):
    """The junction-read half of the short-read genotyping pass against annotateCoverage run verbatim over the same in-memory
    alignments: per-SV REF/ALT quality lists, order included. (60, 3600): > 131072 jobs on one contig, so the reference flushes a
    batch mid-contig and consults the counts it merged — the batch boundary is part of the result."""
    nsv, rps, cap = shape
    H, ctxh = hostdev
    d = _sr_geno_case(ref, 900 + rps, nsv, rps)
    n, nrec = len(d["sv"]), len(d["rec"])
    outs = []
    lib = np.array([300, 100, 500, 600], np.int32)
    for fn, lead in ((ref3.ref_annotate_junction_reads, ()), (H.dh_annotate_junction_reads, (ctxh,))):
        q = np.zeros(2_000_000, np.uint8); ro = np.zeros(n + 1, np.uint32); ao = np.zeros(n + 1, np.uint32)
        extra = ()
        if not lead:   # the reference hook also reports the spanning / read-depth half (checked in test_annotate_spanning_...)
            sq = np.zeros(2_000_000, np.uint8); so1 = np.zeros(n + 1, np.uint32); so2 = np.zeros(n + 1, np.uint32); src = np.zeros((n, 3), np.int32)
            extra = (_p(lib), _p(sq), len(sq), _p(so1), _p(so2), _p(src))
        r = fn(*lead, d["g1"], len(d["g1"]), d["g2"], len(d["g2"]), _p(d["rec"]), nrec, _p(d["cig"]), _p(d["reads"]), _p(d["sv"]), n, _p(d["cons"]), _p(d["co"]),
               _p(d["cl"]), C.c_float(0.95), 13, 1000, 100, 5, cap, 1, _p(q), len(q), _p(ro), _p(ao), *extra)
        assert r >= 0, r
        outs.append((r, q[:r].copy(), ro.copy(), ao.copy()))
    e, g = outs
    assert e[0] == g[0] and e[0] > n
    assert np.array_equal(e[1], g[1]) and np.array_equal(e[2], g[2]) and np.array_equal(e[3], g[3])
    assert int(e[3][n] - e[3][0]) > n // 4     # ALT support is found


def _sr_pair_case(seed, nsv=60, pairs_per_sv=30):
    """Paired-end records around imprecise SVs (peSupport > 0): normal FR pairs spanning a breakpoint, abnormal pairs of the SV's own
    type with the mate near the other breakpoint (and some of a different type / far away), soft-clipped mates, low mapq, duplicates,
    unpaired reads, mates at the same position. 12 columns, sorted by (tid, pos)."""
    rng = np.random.default_rng(seed)
    L = [120000, 90000]
    RL = 100
    svs, recs, cigs = [], [], []
    nid = [0]

    def rec(tid, pos, flag, mapq, ops, mtid, mpos, isize, name):
        recs.append([tid, pos, flag, mapq, sum(l for o, l in ops if o in (0, 1, 4, 7, 8)), len(cigs), len(ops), 0, mtid, mpos, isize, name])
        cigs.extend((ln << 4) | op for op, ln in ops)

    def pair(t1, p1, rev1, t2, p2, rev2, clip=False):
        name = nid[0]; nid[0] += 1
        mq1, mq2 = [int(x) for x in rng.choice([2, 20, 60], p=[0.05, 0.25, 0.7], size=2)]
        isz = (p2 + RL - p1) if t1 == t2 else 0
        f1 = 0x1 | 0x40 | (0x10 if rev1 else 0) | (0x20 if rev2 else 0)
        f2 = 0x1 | 0x80 | (0x10 if rev2 else 0) | (0x20 if rev1 else 0)
        if rng.random() < 0.03: f1 |= 0x400
        ops1 = [(0, RL - 20), (4, 20)] if clip else ([(0, 40), (2, 3), (0, RL - 40)] if rng.random() < 0.1 else [(0, RL)])
        rec(t1, p1, f1, mq1, ops1, t2, p2, isz, name)
        rec(t2, p2, f2, mq2, [(7, RL // 2), (8, 1), (0, RL - RL // 2 - 1)] if rng.random() < 0.1 else [(0, RL)], t1, p1, -isz, name)

    for i in range(nsv):
        svt = int(rng.choice([0, 1, 2, 3, 4, 5, 6, 7, 8]))
        s = int(rng.integers(3000, 80000)); size = int(rng.choice([int(rng.integers(300, 900)), int(rng.integers(1500, 6000))]))
        if svt >= 5: c1, c2, e = 1, 0, int(rng.integers(3000, 80000))
        elif svt == 4: c1, c2, e = 0, 0, s + 1
        else: c1, c2, e = 0, 0, s + size
        svs.append([c1, s, c2, e, svt, 0, 0, i, int(rng.integers(0, 4))])
        for _ in range(pairs_per_sv):
            k = rng.random()
            if k < 0.4:   # normal FR pair over one of the breakpoints
                tid, p = (c1, s) if rng.random() < 0.5 else (c2, e)
                ins = int(rng.integers(80, 650)); a = int(rng.integers(0, max(ins - 10, 1)))
                pair(tid, max(p - a, 0), False, tid, max(p - a, 0) + ins - RL if ins > RL else max(p - a, 0), True, clip=rng.random() < 0.1)
            elif k < 0.85:  # abnormal pair of some SV type near the breakpoints
                t = svt if rng.random() < 0.8 else int(rng.choice([0, 1, 2, 3]))
                d1, d2 = int(rng.integers(-450, 100)), int(rng.integers(-100, 450))
                if t >= 5:
                    ct = t - 5
                    pair(c2, max(e + d1, 0), ct in (1, 3) if ct < 2 else ct == 3, c1, max(s + d2, 0), ct in (1,) if ct < 2 else ct == 2)
                elif t == 0: pair(c1, max(s + d1, 0), False, c2, max(e + d1, 0), False)
                elif t == 1: pair(c1, max(s + d2, 0), True, c2, max(e + d2, 0), True)
                elif t == 2: pair(c1, max(s + d1, 0), False, c2, max(e + d2, 0), True)
                else: pair(c1, max(s + d2, 0), True, c2, max(e + d1, 0), False)
            elif k < 0.9:  # both mates at the same position
                p = max(s + int(rng.integers(-200, 200)), 0)
                pair(c1, p, False, c1, p, True)
            else:          # unpaired read
                rec(c1, max(s + int(rng.integers(-200, 200)), 0), int(rng.choice([0, 16])), 60, [(0, RL)], -1, -1, 0, nid[0]); nid[0] += 1
    r = np.array(recs, np.int64)
    order = np.lexsort((np.arange(len(r)), r[:, 1], r[:, 0]))
    return dict(L=L, rec=np.ascontiguousarray(r[order].astype(np.int32)), cig=np.array(cigs, np.uint32), sv=np.array(svs, np.int32))


@pytest.mark.parametrize("lib", [(300, 100, 500, 600), (0, 0, 0, 0)])
def test_annotate_spanning_and_depth_matches_reference(ref3, lib):
    """The spanning-pair / read-depth half of annotateCoverage (pure host logic) against the reference function run verbatim over
    in-memory alignments: REF and ALT spanning qualities per SV (order included) and the left / body / right read-depth.
    lib median 0 = single-end library: no spanning counts, depth only."""
    H = delly_b200.hostlib()
    d = _sr_pair_case(77)
    n, nrec = len(d["sv"]), len(d["rec"])
    g = [b"A" * d["L"][0], b"A" * d["L"][1]]
    reads = np.full(200, ord("A"), np.uint8)   # every record points at the same dummy bases (the SVs are imprecise: no probes, no jobs)
    libv = np.array(lib, np.int32)
    cons = np.zeros(1, np.uint8); zo = np.zeros(n, np.uint32)
    q = np.zeros(1000, np.uint8); ro = np.zeros(n + 1, np.uint32); ao = np.zeros(n + 1, np.uint32)
    es = np.zeros(200000, np.uint8); esr = np.zeros(n + 1, np.uint32); esa = np.zeros(n + 1, np.uint32); erc = np.zeros((n, 3), np.int32)
    r = ref3.ref_annotate_junction_reads(g[0], len(g[0]), g[1], len(g[1]), _p(d["rec"]), nrec, _p(d["cig"]), _p(reads), _p(d["sv"]), n, _p(cons), _p(zo), _p(zo),
                                         C.c_float(0.95), 13, 1000, 100, 5, 250, 1, _p(q), len(q), _p(ro), _p(ao), _p(libv), _p(es), len(es), _p(esr), _p(esa), _p(erc))
    assert r == 0, r
    gs = np.zeros(200000, np.uint8); gsr = np.zeros(n + 1, np.uint32); gsa = np.zeros(n + 1, np.uint32); grc = np.zeros((n, 3), np.int32)
    sp = H.dh_annotate_spanning(len(g[0]), len(g[1]), _p(d["rec"]), nrec, _p(d["cig"]), _p(d["sv"]), n, 1000, 5, _p(libv), _p(gs), len(gs), _p(gsr), _p(gsa), _p(grc))
    assert sp >= 0 and sp == int(esa[n])
    assert np.array_equal(esr, gsr) and np.array_equal(esa, gsa) and np.array_equal(es[:sp], gs[:sp])
    assert np.array_equal(erc, grc) and erc.sum() > 0
    if lib[0]:
        assert int(esr[n] - esr[0]) > n and int(esa[n] - esa[0]) > n     # both kinds of spanning support occur
    else:
        assert sp == 0


# ---- discovery front end (scanPEandSR, src/shortpe.h:285-533) -----------------------------------------------------------

def _hash_string(s):
    h = 37
    for ch in s.encode():
        h = ((h * 54059) ^ (ch * 76963)) & 0xFFFFFFFF
    return h


def _scan_case(seed):
    """The paired-end records of _sr_pair_case plus split reads: a primary alignment clipped at one breakpoint and a supplementary one
    clipped at the other (same read), in the strand / clip-side combination of the SV type; reads with a long deletion / insertion in
    the CIGAR; everything re-sorted by (tid, pos)."""
    d = _sr_pair_case(seed, nsv=80, pairs_per_sv=24)
    rng = np.random.default_rng(seed + 1)
    recs = d["rec"].tolist(); cigs = d["cig"].tolist()
    name = int(d["rec"][:, 11].max()) + 1
    RL = 100

    def rec(tid, pos, flag, ops, nm):
        recs.append([tid, max(pos, 0), flag, int(rng.choice([3, 20, 60], p=[0.05, 0.25, 0.7])), sum(l for o, l in ops if o in (0, 1, 4, 7, 8)), len(cigs), len(ops), 0, tid,
                     max(pos, 0) + 250, 0, nm])
        cigs.extend((ln << 4) | op for op, ln in ops)

    for sv in d["sv"]:
        c1, s, c2, e, svt = [int(x) for x in sv[:5]]
        for _ in range(int(rng.integers(0, 9))):
            a = int(rng.integers(30, 70)); j1, j2 = int(rng.integers(-3, 4)), int(rng.integers(-3, 4))
            rd = int(rng.choice([0x40, 0x80]))
            ct = svt if svt < 5 else svt - 5
            if svt == 4:
                rec(c1, s + j1 - a, rd, [(0, a), (1, int(rng.integers(30, 60))), (0, RL - a)], name)
            elif ct == 2:   # deletion-type: right clip at the start, left clip at the end, same strand
                if rng.random() < 0.3 and svt == 2: rec(c1, s + j1 - a, rd, [(0, a), (2, e - s), (0, RL - a)], name)
                else:
                    rec(c1, s + j1 - a, rd, [(0, a), (4, RL - a)], name)
                    rec(c2, e + j2, rd | 0x800, [(5, a), (0, RL - a)], name)
            elif ct == 3:   # duplication-type: left clip at the start, right clip at the end
                rec(c1, s + j1, rd, [(4, a), (0, RL - a)], name)
                rec(c2, e + j2 - (RL - a), rd | 0x800, [(0, RL - a), (5, a)], name)
            elif ct == 0:   # 3to3: right clips on opposite strands
                rec(c1, s + j1 - a, rd, [(0, a), (4, RL - a)], name)
                rec(c2, e + j2 - (RL - a), rd | 0x800 | 0x10, [(0, RL - a), (5, a)], name)
            else:           # 5to5: left clips on opposite strands
                rec(c1, s + j1, rd, [(4, a), (0, RL - a)], name)
                rec(c2, e + j2, rd | 0x800 | 0x10, [(5, a), (0, RL - a)], name)
            name += 1
    r = np.array(recs, np.int64)
    order = np.lexsort((np.arange(len(r)), r[:, 1], r[:, 0]))
    return dict(L=d["L"], rec=np.ascontiguousarray(r[order].astype(np.int32)), cig=np.array(cigs, np.uint32))


@pytest.mark.gpu
def test_scan_pe_sr_device_pair_scan_matches_reference(ctx, ref5):
    """Same comparison with the pair scans of both cluster() flavours on the device (clusterGpu)."""
    _scan_compare(ref5, delly_b200.hostlib(), ctx.h, 300)


@pytest.mark.parametrize("median", [300, 0])
def test_scan_pe_sr_matches_reference(ref5, median):
    """The discovery front end against scanPEandSR run verbatim over in-memory alignments: paired-end SVs, split-read SVs, the split-read
    store and the abnormal-pair count. median 0 = single-end library (no paired-end evidence)."""
    _scan_compare(ref5, delly_b200.hostlib(), None, median)


def _scan_compare(ref5, H, ctx_h, median):
    d = _scan_case(31)
    nrec = len(d["rec"])
    tl = np.array(d["L"], np.uint32)
    lib = np.array([100, median, 20, 500, 600], np.int32)
    ref5.ref_hash_sr_name.restype = C.c_uint64
    names = [f"q{int(r[11])}".encode() for r in d["rec"]]
    seeds = np.array([ref5.ref_hash_sr_name(nm, 1 if (int(r[2]) & 0x80) else 0) for nm, r in zip(names, d["rec"])], np.uint64)
    nh = np.array([_hash_string(nm.decode()) for nm in names], np.uint32)
    outs = []
    for which in ("ref", "ours"):
        pe = np.zeros((4096, 12), np.int32); sr = np.zeros((4096, 14), np.int32); st = np.zeros((40000, 3), np.int32); ss = np.zeros(40000, np.uint64)
        no = np.zeros(3, np.int32); ab = C.c_uint32()
        if which == "ref":
            rc = ref5.ref_scan_pe_sr(_p(tl), 2, _p(d["rec"]), nrec, _p(d["cig"]), _p(lib), 1, 20, 25, 25, 40, 2, 1000, _p(pe), _p(sr), 4096, _p(st), _p(ss), 40000, _p(no),
                                     C.byref(ab))
        else:
            rc = H.dh_scan_pe_sr(ctx_h, _p(tl), 2, _p(d["rec"]), _p(seeds), _p(nh), nrec, _p(d["cig"]), _p(lib), 1, 20, 25, 25, 40, 2, 1000, _p(pe), _p(sr), 4096, _p(st),
                                 _p(ss), 40000, _p(no), C.byref(ab))
        assert rc == 0, rc
        outs.append((no.copy(), pe[:no[0]].copy(), sr[:no[1]].copy(), st[:no[2]].copy(), ss[:no[2]].copy(), ab.value))
    e, g = outs
    assert e[0].tolist() == g[0].tolist(), (e[0], g[0])
    assert np.array_equal(e[1], g[1]) and np.array_equal(e[2], g[2])
    assert np.array_equal(e[3], g[3]) and np.array_equal(e[4], g[4]) and e[5] == g[5]
    assert e[0][1] > 20 and e[0][2] > 40
    if median: assert e[0][0] > 10 and e[5] > 100
    else: assert e[0][0] == 0


def test_merge_sort_matches_reference(ref5):
    """mergeSort (src/shortpe.h:536-621): paired-end SVs refined by matching split-read SVs, split-read-only SVs appended unless a better
    precise duplicate sits within 10 bp; the output order (sorted after every append) included."""
    H = delly_b200.hostlib()
    rng = np.random.default_rng(8)
    pe, sr = [], []
    for i in range(300):
        svt = int(rng.integers(0, 9)); chr_ = int(rng.integers(0, 3)); chr2 = chr_ if svt < 5 else int(rng.integers(0, 3))
        s = int(rng.integers(1000, 60000)); e = s + int(rng.integers(50, 5000))
        ci = int(rng.choice([50, 120, 300]))
        if rng.random() < 0.7:
            pe.append([chr_, s + int(rng.integers(-80, 80)), chr2, e + int(rng.integers(-80, 80)), -ci, ci, -ci, ci, int(rng.integers(2, 30)), 0, int(rng.integers(10, 60)), 0,
                       int(rng.integers(20, 500)), 0, 0, svt, 0, 0, 100000 + i, 0])
        for _ in range(int(rng.choice([0, 1, 1, 2, 3]))):   # split-read calls of the same event, a few bp apart (precise duplicates)
            h = int(rng.integers(0, 6)); sup = int(rng.choice([0, 2, 3, 3, 8])); q = int(rng.choice([0, 950, 1000]))
            sr.append([chr_, s + int(rng.integers(-4, 5)), chr2, e + int(rng.integers(-4, 5)), -h, h, -h, h, 0, sup, 0, int(rng.integers(10, 60)), int(rng.integers(20, 300)),
                       int(rng.integers(0, 40)), h, svt if rng.random() < 0.9 else int(rng.integers(0, 9)), 1, int(rng.integers(20, 200)), len(sr), q])
    pe = np.array(pe, np.int32); sr = np.array(sr, np.int32)
    outs = []
    for lib in (ref5.ref_merge_sort, H.dh_merge_sort):
        out = np.zeros((len(pe) + len(sr), 20), np.int32)
        n = lib(_p(pe), len(pe), _p(sr), len(sr), _p(out), len(out))
        assert n > len(pe)
        outs.append(out[:n].copy())
    assert np.array_equal(outs[0], outs[1])
    assert (outs[0][:, 16] == 1).sum() > 50 and ((outs[0][:, 8] > 0) & (outs[0][:, 16] == 1)).sum() > 20   # refined paired-end SVs exist


# ---- the whole short-read call path (dellyRun's stage sequence, src/delly.h:127-178) ------------------------------------

def _simulate_sr_sample(seed, n_del=14, cov=30):
    """A diploid sample with heterozygous and homozygous deletions on contig 0: 100 bp read pairs (insert ~ N(300, 15)) sampled from both
    haplotypes and 'aligned' by construction — a read over a deletion junction becomes a soft-clipped primary alignment of its longer
    part plus a supplementary (hard-clipped) alignment of the rest. Returns the flat record layout of the oracle hooks."""
    rng = np.random.default_rng(seed)
    RL = 100
    ref0 = synth._ACGT[rng.integers(0, 4, size=70000, dtype=np.uint8)]; ref1 = synth._ACGT[rng.integers(0, 4, size=20000, dtype=np.uint8)]
    starts = np.sort(rng.choice(np.arange(4000, 64000, 4000), size=n_del, replace=False)) + rng.integers(0, 500, size=n_del)
    sizes = rng.integers(300, 2000, size=n_del)
    zyg = rng.choice([1, 2], size=n_del, p=[0.7, 0.3])   # 1 = heterozygous (haplotype 1 only), 2 = homozygous
    haps = []
    for h in (0, 1):
        dels = [(int(s), int(s + z)) for s, z, g in zip(starts, sizes, zyg) if g == 2 or h == 1]
        pieces, prev, junc, shift = [], 0, [], 0
        for s, e in dels:
            pieces.append(ref0[prev:s]); prev = e
            junc.append((s - shift, s, e)); shift += e - s      # (donor position of the junction, ref start, ref end)
        pieces.append(ref0[prev:])
        haps.append((np.concatenate(pieces), junc))
    recs, cigs, reads = [], [], []
    npairs = int(len(ref0) * cov / (2 * RL))

    def place(donor, junc, a):
        """alignments of donor[a:a+RL]: list of (ref pos, ops) with the primary first"""
        b = a + RL
        for j, s, e in junc:
            if a < j < b:
                L, R = j - a, b - j
                base = s - L      # ref position of the left part
                left = (base, [(0, L), (4, R)]); right = (e, [(4, L), (0, R)])
                if min(L, R) < 20: return [left if L >= R else right]
                lsup = (base, [(0, L), (5, R)]); rsup = (e, [(5, L), (0, R)])
                return [left, rsup] if L >= R else [right, lsup]
        sh = sum(e - s for j, s, e in junc if j <= a)
        return [(a + sh, [(0, RL)])]

    for pid in range(npairs):
        tid = 0 if rng.random() < 0.85 else 1
        if tid == 1:
            donor, junc = ref1, []
        else:
            donor, junc = haps[int(rng.integers(0, 2))]
        ins = int(np.clip(rng.normal(300, 15), 220, 380))
        a1 = int(rng.integers(0, len(donor) - ins)); a2 = a1 + ins - RL
        al = [place(donor, junc, a1), place(donor, junc, a2)]
        p = [al[0][0][0], al[1][0][0]]
        span = p[1] + RL - p[0]
        for k in (0, 1):
            seq = synth.sub_noise(rng, donor[(a1, a2)[k]:(a1, a2)[k] + RL].copy(), 0.003)
            base = 0x1 | (0x40 if k == 0 else 0x80) | (0x10 if k == 1 else 0x20)
            for ai, (pos, ops) in enumerate(al[k]):
                recs.append([tid, pos, base | (0x800 if ai else 0), 60, RL, len(cigs), len(ops), RL * len(reads), tid, p[1 - k], span if k == 0 else -span, pid])
                cigs.extend((ln << 4) | op for op, ln in ops)
                reads.append(seq)
    r = np.array(recs, np.int64)
    order = np.lexsort((np.arange(len(r)), r[:, 1], r[:, 0]))
    contig = np.concatenate([ref0, ref1])
    truth = [(int(s), int(s + z), int(g)) for s, z, g in zip(starts, sizes, zyg)]
    return dict(contig=contig, coff=np.array([0, len(ref0)], np.uint32), clen=np.array([len(ref0), len(ref1)], np.uint32), rec=np.ascontiguousarray(r[order].astype(np.int32)),
                cig=np.array(cigs, np.uint32), reads=np.concatenate(reads).astype(np.uint8), truth=truth)


def test_delly_sr_call_matches_reference_end_to_end(hostdev, ref5):
    """Alignments in, genotyped SV records out: the stage sequence of dellyRun with every stage the batched mirror of this repository
    (device kernels for the realignments, the MSA and the split alignments) against the same sequence of the reference's own functions
    compiled verbatim. Compared per SV: coordinates, confidence intervals, supports, qualities, consensus, GT / GQ / PL / GL bits,
    DR / DV / RR / RV, read-depth. The simulated deletions must be found and genotyped."""
    H, ctxh = hostdev
    d = _simulate_sr_sample(2024)
    nrec = len(d["rec"])
    lib = np.array([100, 300, 15, 200, 400, 480], np.int32)
    ref5.ref_hash_sr_name.restype = C.c_uint64
    names = [f"q{int(r[11])}".encode() for r in d["rec"]]
    seeds = np.array([ref5.ref_hash_sr_name(nm, 1 if (int(r[2]) & 0x80) else 0) for nm, r in zip(names, d["rec"])], np.uint64)
    nh = np.array([_hash_string(nm.decode()) for nm in names], np.uint32)
    outs = []
    for which in ("ref", "ours"):
        sv = np.zeros((512, 20), np.int32); fmt = np.zeros((512, 14), np.int32); gl = np.zeros((512, 3), np.float32)
        co = np.zeros((512, 1024), np.uint8); cl = np.zeros(512, np.int32)
        common = (_p(d["cig"]), _p(d["reads"]), _p(lib), _p(sv), 512, _p(fmt), _p(gl), _p(co), 1024, _p(cl))
        if which == "ref":
            n = ref5.ref_delly_sr_call(_p(d["contig"]), _p(d["coff"]), _p(d["clen"]), 2, _p(d["rec"]), nrec, *common)
        else:
            n = H.dh_delly_sr_call(ctxh, _p(d["contig"]), _p(d["coff"]), _p(d["clen"]), 2, _p(d["rec"]), _p(seeds), _p(nh), nrec, *common)
        assert n > 0, n
        outs.append((n, sv[:n].copy(), fmt[:n].copy(), gl[:n].copy(), [co[i, :cl[i]].tobytes() for i in range(n)]))
    e, g = outs
    assert e[0] == g[0]
    assert np.array_equal(e[1], g[1]), np.argwhere(e[1] != g[1])[:5]
    assert np.array_equal(e[2], g[2]), np.argwhere(e[2] != g[2])[:5]
    assert np.array_equal(e[3].view(np.uint32), g[3].view(np.uint32))
    assert e[4] == g[4]
    # and the calls are the planted deletions, precise, with the right genotype
    found = 0
    for s, en, zyg in d["truth"]:
        for i in range(e[0]):
            v = e[1][i]
            if v[15] == 2 and v[16] == 1 and abs(int(v[1]) - s) <= 3 and abs(int(v[3]) - en) <= 3:
                gt = (int(e[2][i][0]) >> 1) - 1, (int(e[2][i][1]) >> 1) - 1
                found += (gt == ((0, 1) if zyg == 1 else (1, 1)))
                break
    assert found >= len(d["truth"]) - 2, (found, len(d["truth"]))


# ---- the whole long-read call path (runTegua's stage sequence, src/tegua.h:104-193) -------------------------------------

def _simulate_lr_sample(seed, n_sv=12, cov=24, insert_pool=None):
    """A diploid sample with heterozygous / homozygous deletions and insertions on contig 0, sequenced with 2-5 kb reads (5 % substitutions
    plus short indels, all reflected in the CIGAR): a read over an event carries it in its CIGAR (aM <size>D bM / aM <len>I bM).
    insert_pool: sequences to insert instead of random ones (mobile elements, repeat expansions); deletions then alternate with
    deletions of a planted tandem repeat."""
    rng = np.random.default_rng(seed)
    ref0 = synth._ACGT[rng.integers(0, 4, size=90000, dtype=np.uint8)]; ref1 = synth._ACGT[rng.integers(0, 4, size=20000, dtype=np.uint8)]
    starts = np.sort(rng.choice(np.arange(6000, 84000, 6000), size=n_sv, replace=False)) + rng.integers(0, 800, size=n_sv)
    kinds = rng.choice([2, 4], size=n_sv)
    sizes = rng.integers(150, 1200, size=n_sv)
    zyg = rng.choice([1, 2], size=n_sv, p=[0.65, 0.35])
    inserts = [synth._ACGT[rng.integers(0, 4, size=int(z), dtype=np.uint8)] for z in sizes]
    if insert_pool is not None:
        rng2 = np.random.default_rng(seed + 1000)
        n_ins = 0
        for i in range(n_sv):
            if kinds[i] == 4:
                inserts[i] = np.asarray(insert_pool[n_ins % len(insert_pool)], np.uint8)
                sizes[i] = len(inserts[i]); n_ins += 1
            elif i % 2 == 0:   # the deleted segment is a run of a short unit (DEL:TR)
                unit = synth._ACGT[rng2.integers(0, 4, size=int(rng2.integers(3, 40)))]
                s0 = int(starts[i])
                ref0[s0 - 50:s0 + int(sizes[i]) + 50] = np.resize(unit, int(sizes[i]) + 100)
    recs, cigs, reads = [], [], []
    total = int(len(ref0) * cov / 3500)
    for rid in range(total):
        tid = 0 if rng.random() < 0.9 else 1
        L = int(rng.integers(2000, 5001))
        hap = int(rng.integers(0, 2))
        G = ref0 if tid == 0 else ref1
        p = int(rng.integers(0, len(G) - L - 2000))
        # walk the reference from p, applying this haplotype's events, emitting (op, len, bases)
        ops, seq, rp, left = [], [], p, L
        ev = [(int(s), int(k), int(z), i) for i, (s, k, z, g) in enumerate(zip(starts, kinds, sizes, zyg)) if tid == 0 and (g == 2 or hap == 1) and s > p + 300]
        for s, k, z, i in ev:
            if s - rp >= left - 300: break
            ops.append((0, s - rp)); seq.append(G[rp:s]); left -= s - rp; rp = s
            if k == 2: ops.append((2, z)); rp += z
            else: ops.append((1, z)); seq.append(inserts[i]); left -= z
        if left > 0: ops.append((0, left)); seq.append(G[rp:rp + left])
        # noise inside the aligned blocks: substitutions, and short indels that split an M block
        nops, nseq = [], []
        for (op, ln), sq in zip([o for o in ops if o[0] != 2] if False else ops, iter(seq + [None] * len(ops))):
            pass
        si = 0
        for op, ln in ops:
            if op == 2: nops.append((2, ln)); continue
            sq = seq[si].copy(); si += 1
            if op == 1: nops.append((1, ln)); nseq.append(sq); continue
            sq = synth.sub_noise(rng, sq, 0.05)
            cut = 0
            while len(sq) - cut > 400 and rng.random() < 0.6:
                step = int(rng.integers(150, 400)); kind = rng.random()
                nops.append((0, step)); nseq.append(sq[cut:cut + step]); cut += step
                if kind < 0.5: nops.append((1, 2)); nseq.append(synth._ACGT[rng.integers(0, 4, size=2, dtype=np.uint8)])
                else: nops.append((2, 2)); cut += 0   # a 2-bp deletion: the read simply lacks them (sequence taken from the shifted reference is fine for a test)
            nops.append((0, len(sq) - cut)); nseq.append(sq[cut:])
        s = np.concatenate(nseq)
        flag = int(rng.choice([0, 16]))
        if rng.random() < 0.03: flag |= 0x400
        recs.append([tid, p, flag, int(rng.choice([0, 20, 60], p=[0.03, 0.12, 0.85])), len(s), len(cigs), len(nops), sum(len(r) for r in reads), tid, 0, 0, rid])
        cigs.extend((ln << 4) | op for op, ln in nops)
        reads.append(s)
    r = np.array(recs, np.int64)
    order = np.lexsort((np.arange(len(r)), r[:, 1], r[:, 0]))
    truth = [(int(s), int(k), int(z), int(g)) for s, k, z, g in zip(starts, kinds, sizes, zyg)]
    return dict(contig=np.concatenate([ref0, ref1]), coff=np.array([0, len(ref0)], np.uint32), clen=np.array([len(ref0), len(ref1)], np.uint32),
                rec=np.ascontiguousarray(r[order].astype(np.int32)), cig=np.array(cigs, np.uint32), reads=np.concatenate(reads).astype(np.uint8), truth=truth)


LR_CFG = np.array([1, 25, 30, 75, 2, 1000, 15, 100, 10000, 400, 250, 25], np.int32)   # lr settings (src/tegua.h:230-266), a smaller consensus window


def test_delly_lr_call_matches_reference_end_to_end(hostdev, ref5):
    """Long-read alignments in, genotyped SV records out: runTegua's stage sequence with the batched mirrors of this repository against the
    same sequence of the reference's own functions compiled verbatim (junction scan, clustering, assembly with msaEdlib / msaWfa,
    neighbour de-duplication, genotypeLR, GLs). Every record field must agree; the planted events must be called."""
    H, ctxh = hostdev
    d = _simulate_lr_sample(777)
    nrec = len(d["rec"])
    ref5.ref_hash_lr_name5.restype = C.c_uint64
    seeds = np.array([ref5.ref_hash_lr_name5(f"q{int(r[11])}".encode()) for r in d["rec"]], np.uint64)
    outs = []
    for which in ("ref", "ours"):
        sv = np.zeros((512, 20), np.int32); fmt = np.zeros((512, 14), np.int32); gl = np.zeros((512, 3), np.float32)
        co = np.zeros((512, 8192), np.uint8); cl = np.zeros(512, np.int32)
        tail = (_p(d["cig"]), _p(d["reads"]), _p(LR_CFG), C.c_float(0.9), C.c_float(0.5), _p(sv), 512, _p(fmt), _p(gl), _p(co), 8192, _p(cl))
        if which == "ref":
            n = ref5.ref_delly_lr_call(_p(d["contig"]), _p(d["coff"]), _p(d["clen"]), 2, _p(d["rec"]), nrec, *tail)
        else:
            n = H.dh_delly_lr_call(ctxh, _p(d["contig"]), _p(d["coff"]), _p(d["clen"]), 2, _p(d["rec"]), _p(seeds), nrec, *tail)
        assert n > 0, n
        outs.append((n, sv[:n].copy(), fmt[:n].copy(), gl[:n].copy(), [co[i, :cl[i]].tobytes() for i in range(n)]))
    e, g = outs
    assert e[0] == g[0]
    assert np.array_equal(e[1], g[1]), np.argwhere(e[1] != g[1])[:5]
    assert np.array_equal(e[2], g[2]), np.argwhere(e[2] != g[2])[:5]
    assert np.array_equal(e[3].view(np.uint32), g[3].view(np.uint32))
    assert e[4] == g[4]
    found = 0
    for s, k, z, zyg in d["truth"]:
        for i in range(e[0]):
            v = e[1][i]
            if v[15] == k and v[16] == 1 and abs(int(v[1]) - s) <= 10:
                found += 1
                break
    assert found >= len(d["truth"]) - 3, (found, len(d["truth"]))


def test_host_stages_accept_empty_inputs(ref5):
    """No records / no SVs: every host stage returns cleanly with empty results (as the reference's functions do)."""
    H = delly_b200.hostlib()
    z32 = np.zeros(16, np.int32); zu32 = np.zeros(16, np.uint32); zu64 = np.zeros(16, np.uint64); zu8 = np.zeros(16, np.uint8)
    tl = np.array([1000, 1000], np.uint32); lib = np.array([100, 300, 20, 500, 600], np.int32)
    pe = np.zeros((4, 12), np.int32); sr = np.zeros((4, 14), np.int32); st = np.zeros((4, 3), np.int32); ss = np.zeros(4, np.uint64); no = np.zeros(3, np.int32)
    ab = C.c_uint32(7)
    for which in ("ref", "ours"):
        if which == "ref":
            rc = ref5.ref_scan_pe_sr(_p(tl), 2, _p(z32), 0, _p(zu32), _p(lib), 1, 20, 25, 25, 40, 2, 1000, _p(pe), _p(sr), 4, _p(st), _p(ss), 4, _p(no), C.byref(ab))
        else:
            rc = H.dh_scan_pe_sr(None, _p(tl), 2, _p(z32), _p(zu64), _p(zu32), 0, _p(zu32), _p(lib), 1, 20, 25, 25, 40, 2, 1000, _p(pe), _p(sr), 4, _p(st), _p(ss), 4, _p(no),
                                 C.byref(ab))
        assert rc == 0 and no.tolist() == [0, 0, 0] and ab.value == 0
    out = np.zeros((4, 20), np.int32)
    assert ref5.ref_merge_sort(_p(z32), 0, _p(z32), 0, _p(out), 4) == 0 and H.dh_merge_sort(_p(z32), 0, _p(z32), 0, _p(out), 4) == 0
    sro = np.zeros(1, np.uint32); sao = np.zeros(1, np.uint32); rc3 = np.zeros((1, 3), np.int32)
    assert H.dh_annotate_spanning(1000, 1000, _p(z32), 0, _p(zu32), _p(z32), 0, 1000, 5, _p(np.array([300, 100, 500, 600], np.int32)), _p(zu8), 16, _p(sro), _p(sao), _p(rc3)) == 0
    rs = np.zeros(4, np.uint64); ro = np.zeros(5, np.uint32); jn = np.zeros((4, 7), np.int32); nr = C.c_int(9)
    assert H.dh_find_junctions(_p(z32), _p(zu64), 0, _p(zu32), 1, 25, 25, C.c_float(0.5), 1, _p(rs), _p(ro), 4, _p(jn), 4, C.byref(nr)) == 0 and nr.value == 0


def test_cluster_sr_reads_lr_matches_reference(ref5):
    """The long-read discovery front end (findJunctions -> fetchSVs -> sort -> cluster -> read store) against _clusterSRReads run verbatim
    over the simulated long-read sample: clustered SVs and every (read, SV, junction offset) entry of the store."""
    H = delly_b200.hostlib()
    d = _simulate_lr_sample(778, n_sv=13, cov=20)
    nrec = len(d["rec"])
    ref5.ref_hash_lr_name5.restype = C.c_uint64
    seeds = np.array([ref5.ref_hash_lr_name5(f"q{int(r[11])}".encode()) for r in d["rec"]], np.uint64)
    tl = d["clen"]
    outs = []
    for which in ("ref", "ours"):
        sv = np.zeros((1024, 14), np.int32); st = np.zeros((20000, 6), np.int32); ss = np.zeros(20000, np.uint64); no = np.zeros(2, np.int32)
        if which == "ref":
            rc = ref5.ref_cluster_sr_reads(_p(tl), 2, _p(d["rec"]), nrec, _p(d["cig"]), _p(LR_CFG), C.c_float(0.5), _p(sv), 1024, _p(st), _p(ss), 20000, _p(no))
        else:
            rc = H.dh_cluster_sr_reads(None, _p(tl), 2, _p(d["rec"]), _p(seeds), nrec, _p(d["cig"]), _p(LR_CFG), C.c_float(0.5), _p(sv), 1024, _p(st), _p(ss), 20000, _p(no))
        assert rc == 0, rc
        outs.append((no.copy(), sv[:no[0]].copy(), st[:no[1]].copy(), ss[:no[1]].copy()))
    e, g = outs
    assert e[0].tolist() == g[0].tolist() and e[0][0] >= 10 and e[0][1] > 60
    assert np.array_equal(e[1], g[1]) and np.array_equal(e[2], g[2]) and np.array_equal(e[3], g[3])


@pytest.mark.parametrize("case", ["sim", "rf", "few", "single"])
def test_get_library_params_matches_reference(ref6, case):
    """getLibraryParams (src/util.h:771-891, util.h compiled itself) over in-memory alignments: read length, insert-size median / MAD,
    normal-pair window, deletion cut-off. Cases: the simulated sample, a mostly reverse-forward library (kept as single-end), too few
    pairs, unpaired reads only."""
    H = delly_b200.hostlib()
    d = _simulate_sr_sample(99, n_del=6, cov=12)
    rec = d["rec"].copy()
    if case == "rf": rec[:, 2] ^= 0x30          # swap the strands of every pair: the FR share drops below one half
    if case == "few": rec = np.ascontiguousarray(rec[:600])
    if case == "single": rec[:, 2] &= ~0x1
    outs = []
    for fn in (ref6.ref_get_library_params, H.dh_get_library_params):
        o = np.zeros(7, np.int32)
        assert fn(_p(d["clen"]), 2, _p(rec), len(rec), _p(d["cig"]), 9, 5, _p(o)) == 0
        outs.append(o.tolist())
    assert outs[0] == outs[1], outs
    if case == "sim": assert outs[0][0] == 100 and 290 <= outs[0][1] <= 310 and outs[0][6] >= 500
    if case in ("rf", "single"): assert outs[0][1] == 0 and outs[0][0] == 100
    if case == "few": assert outs[0][0] == 0 and outs[0][1] == 0


# ---- BCF record construction (vcfOutput, src/modvcf.h:344-791) -----------------------------------------------------------

def _vcf_case(seed=12, n=240):
    """Random SV records of every type with count maps, in the flat layout of ref_vcf_output."""
    rng = np.random.default_rng(seed)
    sv = np.zeros((n, 25), np.int32); alle = np.zeros((n, 512), np.uint8); al = np.zeros(n, np.int32); cons = np.zeros((n, 512), np.uint8); cl = np.zeros(n, np.int32)
    quals, jr, ja, sr_, sa = [], [0], [0], [0], [0]
    lists = [[], [], [], []]
    hp = np.zeros((n, 5), np.int32); rc = np.zeros((n, 3), np.int32)
    tl = np.array([100000, 80000, 50000], np.uint32)
    bases = np.frombuffer(b"ACGT", np.uint8)
    for i in range(n):
        svt = int(rng.integers(0, 9)); precise = int(rng.random() < 0.6)
        chr_ = int(rng.integers(0, 3)); chr2 = chr_ if svt < 5 else int((chr_ + 1 + rng.integers(0, 2)) % 3)
        s = int(rng.choice([0, 1, int(rng.integers(2, 49000))])); e = s + int(rng.integers(1, 3000)) if svt != 4 else s + 1
        if rng.random() < 0.05: e = int(tl[chr2]) + 50      # beyond the contig end: clamped
        ci = int(rng.choice([1, 40, 200]))
        sv[i, :19] = [chr_, s, chr2, e, -ci, ci, -ci, ci, int(rng.integers(0, 12)), int(rng.integers(0, 12)) if precise else 0, int(rng.integers(0, 61)), int(rng.integers(0, 61)),
                      int(rng.choice([-5, 0, 300, 20000])), int(rng.integers(0, 300)) if svt == 4 else 0, int(rng.integers(0, 20)), svt, precise, int(rng.integers(0, 200)), i]
        sv[i, 19] = np.float32(rng.choice([0.0, 0.91, 1.0])).view(np.int32)
        sv[i, 20] = int(rng.choice([-1, -1, 7])); sv[i, 21] = int(rng.integers(1, 4))
        sv[i, 22] = int(rng.choice([0, 0, 9])); sv[i, 23] = int(rng.choice([0, 0, 0, 2, 4, 6])); sv[i, 24] = int(rng.integers(0, 2))
        ref_base = "ACGTN"[int(rng.integers(0, 5))]
        k = rng.random()
        if svt >= 5:
            a = ref_base + "," + ["%s]chr%d:%d]" % (ref_base, chr2, e), "[chr%d:%d[%s" % (chr2, e, ref_base), "%s[chr%d:%d[" % (ref_base, chr2, e), "]chr%d:%d]%s" % (chr2, e, ref_base)][svt - 5]
        elif precise and svt in (2, 4) and k < 0.5:   # sequence-resolved alleles, IUPAC codes in the ALT
            r_ = "".join("ACGT"[int(x)] for x in rng.integers(0, 4, size=int(rng.integers(1, 40))))
            a_ = "".join("ACGTRYSWKMBDHVNacgtu"[int(x)] for x in rng.integers(0, 20, size=int(rng.integers(1, 40))))
            a = r_ + "," + a_
        else:
            a = ref_base + ",<" + ["INV", "INV", "DEL", "DUP", "INS"][svt] + ">"
        ab = a.encode(); alle[i, :len(ab)] = np.frombuffer(ab, np.uint8); al[i] = len(ab)
        if precise and rng.random() < 0.85:
            c = bases[rng.integers(0, 4, size=int(rng.integers(30, 400)))]; cons[i, :len(c)] = c; cl[i] = len(c)
        for li in range(4):
            m = int(rng.choice([0, 0, 1, 3, 12])); lists[li].append(rng.integers(0, 61, size=m).astype(np.uint8))
        hp[i] = [int(rng.integers(0, 4)), int(rng.integers(0, 4)), int(rng.integers(0, 4)), int(rng.integers(0, 4)), int(rng.choice([-1, -1, 17]))]
        rc[i] = [int(rng.integers(0, 300)) * int(rng.random() < 0.9), int(rng.integers(0, 600)), int(rng.integers(0, 300)) * int(rng.random() < 0.9)]
    q = np.concatenate([np.concatenate(l) if len(l) else np.zeros(0, np.uint8) for l in lists] + [np.zeros(1, np.uint8)])
    offs = []
    base = 0
    for li in range(4):
        o = np.concatenate([[0], np.cumsum([len(x) for x in lists[li]])]).astype(np.uint32) + base
        offs.append(np.ascontiguousarray(o.astype(np.uint32))); base = int(o[-1])
    return dict(n=n, tl=tl, sv=sv, alle=alle, al=al, cons=cons, cl=cl, q=q, offs=offs, hp=hp, rc=rc)


@pytest.mark.parametrize("geno_mode", [0, 1])
def test_vcf_records_match_reference(ref7, geno_mode):
    """Everything vcfOutput hands to htslib — header lines, and per record CHROM/POS/QUAL/ID/alleles/FILTER, every INFO and FORMAT key
    with its values, in call order — against the reference function run verbatim over a recording VCF/BCF stand-in. All SV types, precise
    and imprecise, sequence-resolved and symbolic / breakend alleles (incl. IUPAC codes), annotation subtypes, missing genotypes."""
    H = delly_b200.hostlib()
    d = _vcf_case()
    n, tl, sv, alle, al, cons, cl, q, offs, hp, rc = (d[k] for k in ("n", "tl", "sv", "alle", "al", "cons", "cl", "q", "offs", "hp", "rc"))
    outs = []
    for fn in (ref7.ref_vcf_output, H.dh_vcf_output):
        out = np.zeros(1 << 20, np.uint8)
        L = fn(_p(tl), 3, _p(sv), n, _p(alle), 512, _p(al), _p(cons), 512, _p(cl), _p(q), _p(offs[0]), _p(offs[1]), _p(offs[2]), _p(offs[3]), _p(hp), _p(rc), geno_mode, _p(out), len(out))
        assert 0 < L < len(out)
        outs.append(out[:L].tobytes().decode().split("\n"))
    e, g = outs
    assert len(e) == len(g), (len(e), len(g))
    for a, b in zip(e, g):
        assert a == b, (a, b)
    nrec = sum(1 for l in e if l.startswith("R "))
    assert nrec > n // 3 and (geno_mode == 0 or nrec > n // 2)


@pytest.mark.parametrize("depth", [1, 4])
def test_vcf_records_with_annotation_and_methylation_match_reference(ref7, depth):
    """The long-read extras of the record: tandem-repeat annotation (SUBTYPE INS:TR / DEL:TR, TRPERIOD, TRCOPIES), every mobile-element
    subtype with its strand, and the methylation FORMAT fields (MR / MA / MNC / MDV with the depth gate and the per-type missing
    rules) from a per-SV MethylInfo."""
    H = delly_b200.hostlib()
    d = _vcf_case(seed=31, n=200)
    n, tl, sv, alle, al, cons, cl, q, offs, hp, rc = (d[k] for k in ("n", "tl", "sv", "alle", "al", "cons", "cl", "q", "offs", "hp", "rc"))
    rng = np.random.default_rng(5)
    sv[:, 23] = rng.choice([0, 1, 2, 3, 4, 5, 6, 7, 7], size=n)
    anno_tr = np.zeros((n, 2), np.int32)
    anno_tr[:, 0] = rng.integers(1, 100, size=n)
    anno_tr[:, 1] = (rng.integers(10, 4000, size=n) / anno_tr[:, 0]).astype(np.float32).view(np.int32)
    me = np.zeros((n, 16), np.int32)
    me[:, 0:8] = rng.choice([-1, 0, 1, 37, 50, 99, 100], size=(n, 8))
    me[:, 8:12] = rng.choice([-1, 0, 1, 2, 17], size=(n, 4))
    me[:, 12:16] = rng.choice([-1, 0, 1, 3, 4, 5, 30], size=(n, 4))
    outs = []
    for fn in (ref7.ref_vcf_output_ex, H.dh_vcf_output_ex):
        out = np.zeros(1 << 20, np.uint8)
        L = fn(_p(tl), 3, _p(sv), n, _p(alle), 512, _p(al), _p(cons), 512, _p(cl), _p(q), _p(offs[0]), _p(offs[1]), _p(offs[2]), _p(offs[3]), _p(hp), _p(rc), 1, _p(out), len(out),
               _p(anno_tr), _p(me), depth)
        assert 0 < L < len(out)
        outs.append(out[:L].tobytes().decode().split("\n"))
    e, g = outs
    assert len(e) == len(g), (len(e), len(g))
    for a, b in zip(e, g):
        assert a == b, (a, b)
    text = "\n".join(e)
    for key in ("INS:TR", "DEL:TR", "TRPERIOD", "TRCOPIES", "INS:ME:ALU", "INS:ME:LINE1", "INS:ME:SVA", "INS:NUMT", "INS:LTR", "INS:HERVK", "INSSTRAND=-", "INSSTRAND=+"):
        assert key in text, key
    mr = [l for l in e if l.startswith("R ") and "F:MR=" in l]
    assert len(mr) > n // 3 and any("F:MA=37" in l or ",37" in l for l in mr)
