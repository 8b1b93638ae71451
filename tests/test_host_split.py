"""Host C++ mirror of the consensus->breakpoint logic (delly_b200/host/split.hpp, genotype.hpp, msa.hpp).
CPU tests: _getSVRef / _findSplit / longestHomology against the compiled reference.
GPU tests: alignConsensusBatch / processBatch / msaBatch (device alignment + host logic) against the reference chain."""
import ctypes as C

import numpy as np
import pytest

import delly_b200
from delly_b200 import synth
from oracle import pyoracle as po


def _p(a):
    return C.c_void_p(a.ctypes.data)


def _genome(seed, n=60000):
    rng = np.random.default_rng(seed)
    g = synth.random_genome(rng, n)
    g[5000:5040] = ord("N")
    # a little lower case and IUPAC, as FASTA references have
    g[7000:7100] = np.char.lower(g[7000:7100].view("S1")).view(np.uint8)
    g[9000] = ord("R")
    return g.tobytes()


def _sv_cases(seed, g1, g2, n=120, cons_range=(80, 260), realistic=True, with_ins=False):
    """SV records + consensus sequences spanning the planted junction (some offset / noisy / unrelated)."""
    rng = np.random.default_rng(seed)
    G1, G2 = np.frombuffer(g1, np.uint8), np.frombuffer(g2, np.uint8)
    svs, cons = [], []
    for _ in range(n):
        svt = int(rng.choice([0, 1, 2, 3, 4, 4, 5, 6, 7, 8])) if with_ins else int(rng.choice([0, 1, 2, 3, 5, 6, 7, 8]))
        L = int(rng.integers(*cons_range)); off = int(rng.integers(L // 4, 3 * L // 4))
        p1 = int(rng.integers(10000, 30000))
        size = int(rng.choice([int(rng.integers(30, 90)), int(rng.integers(120, 900)), int(rng.integers(1100, 4000))]))
        p2 = p1 + size if svt < 5 else int(rng.integers(10000, 40000))
        A, B = G1, (G1 if svt < 5 else G2)
        ct = svt if svt < 5 else svt - 5
        up = lambda x: np.char.upper(x.view("S1")).view(np.uint8)  # noqa: E731
        if svt == 4:
            ins = int(rng.integers(20, 120))
            p2 = p1 + 1
            c = np.concatenate([A[p1 - off:p1], synth._ACGT[rng.integers(0, 4, size=ins)], A[p1:p1 + (L - off)]])
        elif svt == 2 or ct == 2 and svt >= 5:
            c = np.concatenate([A[p1 - off:p1], B[p2:p2 + (L - off)]])
        elif svt == 3 or ct == 3 and svt >= 5:
            c = np.concatenate([B[p2 - off:p2], A[p1:p1 + (L - off)]])
        elif ct == 0:
            c = np.concatenate([A[p1 - off:p1], synth.revcomp(up(B[p2 - (L - off):p2]))])
        else:
            c = np.concatenate([synth.revcomp(up(A[p1:p1 + off])), B[p2:p2 + (L - off)]])
        c = up(c)
        r = rng.random()
        if r < 0.1:
            c = synth._ACGT[rng.integers(0, 4, size=L)]
        elif r < 0.5:
            c = synth.mutate(rng, c, sub=0.01, ins=0.003, dele=0.003)
        # the caller's coordinates are approximate (cluster means): jitter them
        j1, j2 = int(rng.integers(-8, 9)), int(rng.integers(-8, 9))
        chr_, chr2 = (0, 0) if svt < 5 else (1, 0)   # translocations: chr is the higher contig index (src/junction.h:256)
        if svt >= 5:
            svs.append([chr_, p2 + j2, chr2, p1 + j1, svt, 0])
        elif svt == 4:
            svs.append([chr_, p1 + j1, chr2, p1 + j1 + 1, svt, ins + int(rng.integers(-5, 6))])
        else:
            svs.append([chr_, p1 + j1, chr2, p2 + j2, svt, 0])
        cons.append(c)
    return np.array(svs, np.int32), cons


def test_get_sv_ref_matches_reference(ref):
    H = delly_b200.hostlib()
    g1, g2 = _genome(1), _genome(2)
    svs, cons = _sv_cases(3, g1, g2, n=300)
    n_nonempty = 0
    for i in range(len(svs)):
        sv = svs[i].copy()
        # for translocations contig 1 is "chr" and contig 0 is "chr2": seq = contig of chr, sndSeq = contig of chr2
        seq, snd = (g1, g1) if sv[4] < 5 else (g2, g1)
        lens = (len(g1), len(g1)) if sv[4] < 5 else None
        outs = []
        for lib, fn in ((H, "dh_get_sv_ref"), (ref, "ref_get_sv_ref")):
            buf = C.create_string_buffer(20000); L = C.c_int()
            # both wrappers index target_len by contig id: pass contig 0 = g1 and contig 1 = g2 lengths consistently
            s_in = sv.copy()
            rc = getattr(lib, fn)(seq, len(g1) if sv[4] < 5 else len(g1), snd, len(g2), _p(s_in), len(cons[i]), 13, 1000, 100, buf, 20000, C.byref(L))
            assert rc == 0
            outs.append(buf.raw[:L.value])
        assert outs[0] == outs[1], (i, sv)
        n_nonempty += len(outs[0]) > 0
    assert n_nonempty > 250


def test_find_split_and_homology_match_reference(ref):
    H = delly_b200.hostlib()
    O = po.oracle()
    b = synth.k3_consref_batch(150, seed=17, cons_range=(60, 200), ref_cap=700)
    n_ok = 0
    rng = np.random.default_rng(0)
    for i in range(len(b["c_off"])):
        c = b["seqs"][b["c_off"][i]:b["c_off"][i] + b["c_len"][i]].tobytes()
        r = b["seqs"][b["r_off"][i]:b["r_off"][i] + b["r_len"][i]].tobytes()
        ok, r0, r1 = po.long_needle(O, c, r)
        if not ok:
            continue
        for fq, mfs in ((0.95, 13), (0.9, 25)):
            res = []
            for lib, fn in ((H, "dh_find_split"), (ref, "ref_find_split")):
                ad = np.zeros(6, np.int32); pid = C.c_float()
                okk = getattr(lib, fn)(c, len(c), r, len(r), r0 + r1, len(r0), int(b["svt"][i]), C.c_float(fq), mfs, _p(ad), C.byref(pid))
                res.append((okk, ad.tolist(), pid.value))
            assert res[0] == res[1], (i, res)
            n_ok += res[0][0]
        a = c[: int(rng.integers(1, 40))]; bb = r[: int(rng.integers(1, 40))]
        assert H.dh_longest_homology(a, len(a), bb, len(bb), -1) == ref.ref_longest_homology(a, len(a), bb, len(bb), -1)
        assert H.dh_longest_homology(a, len(a), a[1:] + b"A", len(a), -1) == ref.ref_longest_homology(a, len(a), a[1:] + b"A", len(a), -1)
    assert n_ok > 40


# ----------------------------------------------------------------------------------------------- GPU
@pytest.mark.parametrize("mode", ["sr", "lr_realign"])
def test_align_consensus_batch_matches_reference(hostdev, ref, mode):
    H, ctxh = hostdev
    g1, g2 = _genome(11), _genome(12)
    realign = 1 if mode == "lr_realign" else 0
    fq, mfs, indel, mcw = (0.95, 13, 1000, 100) if mode == "sr" else (0.9, 30, 10000, 300)
    svs, cons = _sv_cases(13 + realign, g1, g2, n=200, cons_range=(80, 260) if mode == "sr" else (200, 500), with_ins=True)
    if realign:  # half of the consensus sequences arrive on the other strand
        cons = [synth.revcomp(c) if i % 2 else c for i, c in enumerate(cons)]
    n = len(svs)
    # reference: one alignConsensus call per SV
    exp = []
    for i in range(n):
        sv = svs[i]
        seq, snd = (g1, g1) if sv[4] < 5 else (g2, g1)
        out = np.zeros(10, np.int32); srq = C.c_float(); al = C.create_string_buffer(8192); all_ = C.c_int()
        co = C.create_string_buffer(4096); col = C.c_int()
        cb = cons[i].tobytes()
        ok = ref.ref_align_consensus(seq, len(g1), snd, len(g2), _p(sv.copy()), cb, len(cb), realign, C.c_float(fq), mfs, indel, mcw, _p(out),
                                     C.byref(srq), al, 8192, C.byref(all_), co, C.byref(col))
        exp.append((ok, out.tolist(), srq.value, al.raw[:all_.value], co.raw[:col.value]))
    # ours: ONE batch over all SVs; the hook takes contig 0 and contig 1
    for sel, (seq, snd) in ((svs[:, 4] >= 0, (g1, g2)),):
        idx = np.nonzero(sel)[0]
        arena, off, ln = synth.pack([cons[i] for i in idx])
        m = len(idx)
        sv_in = np.ascontiguousarray(svs[idx])
        out = np.zeros((m, 10), np.int32); srq = np.zeros(m, np.float32); al = np.zeros((m, 8192), np.uint8); all_ = np.zeros(m, np.int32)
        co = np.zeros((m, 4096), np.uint8); col = np.zeros(m, np.int32); okk = np.zeros(m, np.uint8)
        rc = H.dh_align_consensus_batch(ctxh, seq, len(g1), snd, len(g2), m, _p(sv_in), _p(arena), _p(off), _p(ln), realign, C.c_float(fq), mfs,
                                        indel, mcw, _p(out), _p(srq), _p(al), 8192, _p(all_), _p(co), 4096, _p(col), _p(okk))
        assert rc == 0
        for k, i in enumerate(idx):
            got = (int(okk[k]), out[k].tolist(), float(srq[k]), al[k, :all_[k]].tobytes(), co[k, :col[k]].tobytes())
            assert got == exp[i], (i, svs[i].tolist(), got[:3], exp[i][:3])
    assert sum(e[0] for e in exp) > n // 3


def test_split_align_matches_reference(hostdev, ref):
    """_consRefAlignment for insertions = splitAlign (six edlib PATH calls, src/split.h:480-537) + row swap."""
    H, ctxh = hostdev
    rng = np.random.default_rng(44)
    g = np.frombuffer(_genome(21), np.uint8)
    n_ok = 0
    for it in range(120):
        L = int(rng.integers(60, 140)); ins = int(rng.integers(20, 150)); p = int(rng.integers(10000, 30000))
        cons = np.concatenate([np.char.upper(g[p - L:p].view("S1")).view(np.uint8), synth._ACGT[rng.integers(0, 4, size=ins)],
                               np.char.upper(g[p:p + L].view("S1")).view(np.uint8)])
        if it % 3 == 0:
            cons = synth.mutate(rng, cons, sub=0.01, ins=0.004, dele=0.004)
        if it % 10 == 9:
            cons = synth._ACGT[rng.integers(0, 4, size=len(cons))]
        w = max((len(cons) - ins) // 3, 13)
        ref_s = np.char.upper(g[p - w:p + 1 + w].view("S1")).view(np.uint8).tobytes()
        cb = cons.tobytes()
        res = []
        for lib, fn, extra in ((H, "dh_cons_ref_alignment", (ctxh,)), (ref, "ref_cons_ref_alignment", ())):
            rows = C.create_string_buffer(4 * (len(cb) + len(ref_s)) + 64); al = C.c_int()
            okk = getattr(lib, fn)(*extra, cb, len(cb), ref_s, len(ref_s), 4, rows, C.c_long(len(rows)), C.byref(al))
            assert okk >= 0, okk
            res.append((okk, rows.raw[:2 * al.value] if okk else b""))
        assert res[0] == res[1], (it, len(cb), len(ref_s), res[0][0], res[1][0])
        n_ok += res[0][0]
    assert n_ok > 60


def test_process_batch_matches_reference(hostdev, ref):
    """process_batch (src/coverage.h:412-441): type/qual from two HW distances, scored in double like the reference."""
    H, ctxh = hostdev
    b = synth.k1_genotype_batch(6000, seed=77)
    n = len(b["q_off"]) // 2
    arena = b["seqs"]
    cons_off, cons_len = b["q_off"][0::2].copy(), b["q_len"][0::2].copy()
    ref_off, ref_len = b["q_off"][1::2].copy(), b["q_len"][1::2].copy()
    seq_off, seq_len = b["t_off"][0::2].copy(), b["t_len"][0::2].copy()
    qual = np.random.default_rng(1).integers(0, 61, size=n).astype(np.uint8)
    typ = np.zeros(n, np.uint8); qo = np.zeros(n, np.uint8)
    rc = H.dh_process_batch(ctxh, n, _p(arena), _p(cons_off), _p(cons_len), _p(ref_off), _p(ref_len), _p(seq_off), _p(seq_len), _p(qual),
                            C.c_float(0.95), _p(typ), _p(qo))
    assert rc == 0
    d, _ = po.edit_distance_batch(ref, arena, b["q_off"], b["q_len"], b["t_off"], b["t_len"], b["k"], 2, threads=8)
    fq = np.float64(np.float32(0.95))
    exp_t, exp_q = [], []
    for i in range(n):
        sa = ((1.0 - fq) * float(cons_len[i])) / float(d[2 * i] + 1) if d[2 * i] != -1 else 0.0
        sr = ((1.0 - fq) * float(ref_len[i])) / float(d[2 * i + 1] + 1) if d[2 * i + 1] != -1 else 0.0
        if sr > 0.7 or sa > 0.7:
            if sr > sa:
                exp_t.append(ord("R")); exp_q.append(min(255, min(int(sr * 35), int(qual[i]))))
            else:
                exp_t.append(ord("A")); exp_q.append(min(255, min(int(sa * 35), int(qual[i]))))
        else:
            exp_t.append(ord("N")); exp_q.append(0)
    assert np.array_equal(typ, np.array(exp_t, np.uint8)) and np.array_equal(qo, np.array(exp_q, np.uint8))
    assert (typ == ord("R")).sum() > 500 and (typ == ord("A")).sum() > 500 and (typ == ord("N")).sum() > 50


def test_msa_batch_matches_reference(hostdev, ref):
    H, ctxh = hostdev
    b = synth.k2_msa_batch(40, seed=5, read_len=100, max_off=60, err=0.01)
    ncl = len(b["cluster_off"]) - 1
    cons = np.zeros((ncl, 4096), np.uint8); clen = np.zeros(ncl, np.int32); rows = np.zeros(ncl, np.int32)
    rc = H.dh_msa_batch(ctxh, _p(b["seqs"]), _p(b["read_off"]), _p(b["read_len"]), _p(b["cluster_off"]), ncl, 2, _p(cons), 4096, _p(clen), _p(rows))
    assert rc == 0
    for i in range(ncl):
        reads = [b["seqs"][b["read_off"][r]:b["read_off"][r] + b["read_len"][r]].tobytes() for r in range(b["cluster_off"][i], b["cluster_off"][i + 1])]
        er, ec, _ = po.msa(ref, reads, 2)
        assert cons[i, :clen[i]].tobytes() == ec and rows[i] == er


@pytest.mark.parametrize("shape", [(300, 700, 10), (1800, 3200, 4)])
def test_msa_edlib_batch_matches_reference(hostdev, shape):
    """Long-read consensus: msaEdlib (src/assemble.h:385-473) batched (all-pairs NW distance + progressive IUPAC-aware NW paths)."""
    R2 = po.ref2()
    if R2 is None:
        pytest.fail("oracle/_ref/libdelly_ref2.so not available: run __graft_entry__.build() where /root/reference exists")
    H, ctxh = hostdev
    lo, hi, ncl = shape
    rng = np.random.default_rng(lo)
    reads, coff = [], [0]
    for _ in range(ncl):
        L = int(rng.integers(lo, hi)); base = synth._ACGT[rng.integers(0, 4, size=L + 200)]
        for _ in range(int(rng.integers(3, 13))):
            a = int(rng.integers(0, 100)); r = base[a:a + L + int(rng.integers(-40, 40))]
            reads.append(synth.mutate(rng, r, sub=0.03, ins=0.02, dele=0.02))
        coff.append(len(reads))
    arena, off, ln = synth.pack(reads)
    coff = np.array(coff, np.uint32)
    cons = np.zeros((ncl, 8192), np.uint8); clen = np.zeros(ncl, np.int32); rows = np.zeros(ncl, np.int32)
    rc = H.dh_msa_edlib_batch(ctxh, _p(arena), _p(off), _p(ln), _p(coff), ncl, 2, _p(cons), 8192, _p(clen), _p(rows))
    assert rc == 0, rc
    for i in range(ncl):
        a, b = int(coff[i]), int(coff[i + 1])
        buf = C.create_string_buffer(8192); cl = C.c_int()
        o2 = (off[a:b] - off[a]).astype(np.uint32); l2 = np.ascontiguousarray(ln[a:b])
        sub = np.ascontiguousarray(arena[int(off[a]):int(off[b - 1]) + int(ln[b - 1])])
        er = R2.ref_msa_edlib(_p(sub), _p(o2), _p(l2), b - a, 2, buf, 8192, C.byref(cl))
        assert er >= 0
        assert rows[i] == er and cons[i, :clen[i]].tobytes() == buf.raw[:cl.value], (i, b - a)


@pytest.mark.parametrize("shape", [(300, 700, 10, True), (1200, 2600, 4, True), (400, 900, 6, False)])
def test_msa_wfa_batch_matches_reference(hostdev, shape):
    """Long-read insertion consensus: msaWfa (src/assemble.h:549-725) batched — k-mer diagonal overlaps (NW distance),
    superstring rounds (NW paths), progressive rounds (HW paths with IUPAC equalities), _trimConsensus (HW distance + paths)."""
    R2 = po.ref2()
    if R2 is None:
        pytest.fail("oracle/_ref/libdelly_ref2.so not available: run __graft_entry__.build() where /root/reference exists")
    H, ctxh = hostdev
    lo, hi, ncl, flanks = shape
    rng = np.random.default_rng(lo + 7)
    reads, coff = [], [0]
    FS = 128
    pre = np.zeros((ncl, FS), np.uint8); suf = np.zeros((ncl, FS), np.uint8)
    plen = np.zeros(ncl, np.int32); slen = np.zeros(ncl, np.int32)
    for ci in range(ncl):
        L = int(rng.integers(lo, hi)); base = synth._ACGT[rng.integers(0, 4, size=L + 400)]
        ins_at = 200 + L // 3
        for j in range(int(rng.integers(3, 13))):
            a = int(rng.integers(0, 150)); r = base[a:a + L + int(rng.integers(100, 250))]
            r = synth.mutate(rng, r, sub=0.03, ins=0.02, dele=0.02)
            if ci % 3 == 1 and j % 2 == 1:
                r = np.ascontiguousarray(synth.revcomp(r))
            reads.append(r)
        coff.append(len(reads))
        if flanks:
            # reference flanks left / right of the insertion point; every third cluster sees them reverse-complemented
            p = base[ins_at - 100:ins_at]; s = base[ins_at + 60:ins_at + 160]
            if ci % 3 == 2:
                p, s = synth.revcomp(s), synth.revcomp(p)
            pre[ci, :len(p)] = p; plen[ci] = len(p); suf[ci, :len(s)] = s; slen[ci] = len(s)
    arena, off, ln = synth.pack(reads)
    coff = np.array(coff, np.uint32)
    CAP = 16384
    cons = np.zeros((ncl, CAP), np.uint8); clen = np.zeros(ncl, np.int32); rows = np.zeros(ncl, np.int32)
    rc = H.dh_msa_wfa_batch(ctxh, _p(arena), _p(off), _p(ln), _p(coff), ncl, 2, _p(pre), _p(suf), FS, _p(plen), _p(slen), _p(cons), CAP, _p(clen), _p(rows))
    assert rc == 0, rc
    for i in range(ncl):
        a, b = int(coff[i]), int(coff[i + 1])
        buf = C.create_string_buffer(CAP); cl = C.c_int()
        o2 = (off[a:b] - off[a]).astype(np.uint32); l2 = np.ascontiguousarray(ln[a:b])
        sub = np.ascontiguousarray(arena[int(off[a]):int(off[b - 1]) + int(ln[b - 1])])
        er = R2.ref_msa_wfa(_p(sub), _p(o2), _p(l2), b - a, 2, pre[i].tobytes(), int(plen[i]), suf[i].tobytes(), int(slen[i]), buf, CAP, C.byref(cl))
        assert er >= 0
        assert rows[i] == er and cons[i, :clen[i]].tobytes() == buf.raw[:cl.value], (i, b - a, clen[i], cl.value)


# ---- split-read assembly stage (assembleSplitReads, src/shortpe.h:48-282) -------------------------------------------------

def _assembly_case(seed, nsv=160):
    """Clustered SVs with approximate coordinates and their split reads as BAM-like records: windows of the junction sequence (noise,
    duplicates, more reads than maxReadPerSV, SVs with 0/1 reads), reads at the second breakpoint of inversions / translocations in the
    other orientation (the stage flips them back), filtered records (duplicate / secondary / low mapq), reads not in the store."""
    rng = np.random.default_rng(seed)
    g1, g2 = _genome(seed + 1), _genome(seed + 2)
    svs, cons = _sv_cases(seed + 3, g1, g2, n=nsv, cons_range=(220, 330), with_ins=True)
    recs, cigs, reads, store = [], [], [], []
    name = [0]

    def add(tid, pos, seq, svid, flag=0, mapq=None, stored=True):
        f = flag | (0x80 if rng.random() < 0.5 else 0x40) | 0x1
        recs.append([tid, pos, f, int(rng.choice([0, 20, 60], p=[0.04, 0.2, 0.76])) if mapq is None else mapq, len(seq), len(cigs), 1, sum(len(r) for r in reads), tid,
                     pos + 200, 0, name[0]])
        cigs.append((len(seq) << 4) | 0)
        reads.append(seq)
        if stored: store.append([tid, pos, name[0], 1 if (f & 0x80) else 0, svid])
        name[0] += 1

    for i in range(len(svs)):
        chr_, s, chr2, e, svt, ins = [int(x) for x in svs[i]]
        c = cons[i] if rng.random() < 0.5 else synth.revcomp(cons[i])   # either strand of the junction sequence
        k = int(rng.choice([0, 1, 2, 5, 12, 20, 30]))
        for _ in range(k):
            rl = int(rng.integers(100, 151)); st = int(rng.integers(0, len(c) - rl + 1))
            w = synth.sub_noise(rng, c[st:st + rl].copy(), float(rng.choice([0, 0.01, 0.03])))
            second = rng.random() < 0.3
            # where the read sits decides whether the stage reverse-complements it (src/shortpe.h:115-137, src/split.h:55-68):
            # reads on the flipped side are stored reverse-complemented so that every read of an SV ends up on one strand
            if svt == 0 and second: tid, pos, seq = chr_, s + int(rng.integers(0, 40)), synth.revcomp(w)
            elif svt == 1 and second: tid, pos, seq = chr_, e + int(rng.integers(0, 40)), synth.revcomp(w)
            elif svt == 5: tid, pos, seq = (chr2, e - 60, synth.revcomp(w)) if second else (chr_, s - 60, w)
            elif svt == 6: tid, pos, seq = (chr2, e - 60, w) if second else (chr_, s - 60, synth.revcomp(w))
            elif svt >= 7: tid, pos, seq = (chr2, e - 60, w) if second else (chr_, s - 60, w)
            else: tid, pos, seq = chr_, max(s - 100, 0), w
            r = rng.random()
            if r < 0.05: add(tid, pos, seq, i, flag=0x400)
            elif r < 0.08: add(tid, pos, seq, i, flag=0x100)
            elif r < 0.12: add(tid, pos, seq, i, stored=False)
            else: add(tid, pos, seq, i)
            if rng.random() < 0.1: add(tid, pos, seq.copy(), i)   # identical read: de-duplicated by the store
    rec = np.array(recs, np.int64)
    order = np.lexsort((np.arange(len(rec)), rec[:, 1], rec[:, 0]))
    rec = np.ascontiguousarray(rec[order].astype(np.int32))
    # contig order of the test genomes: translocations have chr = 1 (g2), chr2 = 0 (g1)
    contig = np.frombuffer(g1 + g2, np.uint8).copy()
    sv7 = np.array([[svs[i][0], svs[i][1], svs[i][2], svs[i][3], svs[i][4], svs[i][5], i] for i in range(len(svs))], np.int32)
    return dict(contig=contig, coff=np.array([0, len(g1)], np.uint32), clen=np.array([len(g1), len(g2)], np.uint32), rec=rec, cig=np.array(cigs, np.uint32),
                reads=np.concatenate(reads).astype(np.uint8), store=np.array(store, np.int32), sv=sv7)


def test_assemble_split_reads_matches_reference(hostdev, ref5):
    """The split-read assembly stage against assembleSplitReads run verbatim over the same in-memory alignments: per SV the consensus,
    refined coordinates, support, qualities, homology / insertion lengths, confidence intervals and alleles."""
    H, ctxh = hostdev
    d = _assembly_case(5150)
    n, nrec = len(d["sv"]), len(d["rec"])
    outs = []
    for fn, lead in ((ref5.ref_assemble_split_reads, ()), (H.dh_assemble_split_reads, (ctxh,))):
        so = np.zeros((n, 13), np.int32); srq = np.zeros(n, np.float32); co = np.zeros((n, 2048), np.uint8); cl = np.zeros(n, np.int32)
        al = np.zeros((n, 4096), np.uint8); all_ = np.zeros(n, np.int32)
        rc = fn(*lead, _p(d["contig"]), _p(d["coff"]), _p(d["clen"]), 2, _p(d["rec"]), nrec, _p(d["cig"]), _p(d["reads"]), _p(d["store"]), len(d["store"]), _p(d["sv"]), n,
                20, 1, 2, C.c_float(0.95), 13, 1000, 100, _p(so), _p(srq), _p(co), 2048, _p(cl), _p(al), 4096, _p(all_))
        assert rc == 0, rc
        outs.append((so.copy(), srq.copy(), [co[i, :cl[i]].tobytes() for i in range(n)], [al[i, :all_[i]].tobytes() for i in range(n)]))
    e, g = outs
    for i in range(n):
        assert g[2][i] == e[2][i], (i, d["sv"][i].tolist(), len(g[2][i]), len(e[2][i]))
        assert g[0][i].tolist() == e[0][i].tolist(), (i, d["sv"][i].tolist(), g[0][i].tolist(), e[0][i].tolist())
        assert g[3][i] == e[3][i]
    assert np.array_equal(e[1].view(np.uint32), g[1].view(np.uint32))
    precise = int(e[0][:, 8].sum())
    assert precise >= n // 10 and (e[0][:, 2] > 1).sum() >= precise


# ---- long-read assembly stage (assemble, src/assemble.h:736-964) ----------------------------------------------------------

def _lr_assembly_case(seed, nsv=36):
    """Clustered long-read SVs (all types) with their supporting reads as BAM-like records: 1.2-2.6 kb noisy windows of the junction
    sequence with the junction offset recorded per read (SeqSlice.sstart), forward and reverse flags, reads at the second breakpoint of
    inversions / translocations on the other strand, srSupport equal to / above / far above the number of stored reads (in-scan trigger,
    left-over path, candidate cap), single-read SVs, filtered records."""
    rng = np.random.default_rng(seed)
    G = [synth.random_genome(rng, 70000), synth.random_genome(rng, 60000)]
    F = 1300
    svs, recs, cigs, reads, store = [], [], [], [], []
    name = [0]
    for i in range(nsv):
        svt = int(rng.choice([0, 1, 2, 3, 4, 4, 5, 6, 7, 8]))
        s = int(rng.integers(5000, 40000)); size = int(rng.choice([int(rng.integers(200, 900)), int(rng.integers(1500, 6000))]))
        if svt >= 5: c1, c2, e = 1, 0, int(rng.integers(5000, 40000))
        elif svt == 4: c1, c2, e = 0, 0, s + 1
        else: c1, c2, e = 0, 0, s + size
        A, B = G[c1], G[c2]
        ct = svt if svt < 5 else svt - 5
        ins = np.zeros(0, np.uint8)
        if svt == 4:
            ins = synth._ACGT[rng.integers(0, 4, size=int(rng.integers(60, 500)))]
            J = np.concatenate([A[s - F:s], ins, A[s:s + F]])
        elif ct == 2: J = np.concatenate([A[s - F:s], B[e:e + F]])
        elif ct == 3: J = np.concatenate([B[e - F:e], A[s:s + F]])
        elif ct == 0: J = np.concatenate([A[s - F:s], synth.revcomp(B[e - F:e])])
        else: J = np.concatenate([synth.revcomp(A[s:s + F]), B[e:e + F]])
        jpos = F
        k = int(rng.choice([1, 3, 6, 10, 16, 70]))
        ci = int(rng.choice([0, 20, 60]))
        extra = int(rng.choice([0, 0, 3, 100]))
        svs.append([c1, s + int(rng.integers(-5, 6)), c2, e + int(rng.integers(-5, 6)), svt, len(ins), i, k + extra, -ci, ci, -ci, ci])
        for _ in range(k):
            a = int(rng.integers(500, 1250)); b = int(rng.integers(500, 1250))
            w = synth.mutate(rng, J[jpos - a:jpos + len(ins) + b].copy(), sub=0.02, ins=0.01, dele=0.01)
            off = min(a, len(w) - 1)     # junction offset in the read (indel noise moves it a little: the stage only needs it roughly)
            rev = rng.random() < 0.5
            second = rng.random() < 0.35
            tid, pos = c1, s - a
            flip = False
            if svt >= 5:
                if second: tid, pos = c2, e - 300
                flip = (ct == 0 and tid == c2) or (ct == 1 and tid == c1)
            elif svt == 0:
                if second: pos = e + 100
                flip = pos > (svs[-1][1] + svs[-1][3]) // 2
            elif svt == 1:
                flip = rev
            seq = synth.revcomp(w) if flip else w
            center = (len(w) - off) if flip else off
            sstart = (len(seq) - center) if rev else center
            flag = (16 if rev else 0)
            r = rng.random()
            if r < 0.04: flag |= 0x400
            elif r < 0.08: flag |= 0x800
            recs.append([tid, max(pos, 0), flag, 60, len(seq), len(cigs), 1, sum(len(x) for x in reads), tid, 0, 0, name[0]])
            cigs.append((len(seq) << 4) | 0)
            reads.append(seq)
            if rng.random() > 0.05: store.append([tid, max(pos, 0), name[0], i, sstart, len(ins)])
            name[0] += 1
    rec = np.array(recs, np.int64)
    order = np.lexsort((np.arange(len(rec)), rec[:, 1], rec[:, 0]))
    return dict(contig=np.concatenate(G), coff=np.array([0, len(G[0])], np.uint32), clen=np.array([len(G[0]), len(G[1])], np.uint32),
                rec=np.ascontiguousarray(rec[order].astype(np.int32)), cig=np.array(cigs, np.uint32), reads=np.concatenate(reads).astype(np.uint8),
                store=np.array(store, np.int32), sv=np.array(svs, np.int32))


def test_assemble_lr_matches_reference(hostdev, ref5):
    """The long-read assembly stage against assemble() run verbatim over the same in-memory alignments: per SV the consensus, refined
    coordinates, homology / insertion length, consensus breakpoint, confidence intervals, alleles and the alignment quality bits."""
    H, ctxh = hostdev
    d = _lr_assembly_case(6262)
    n, nrec = len(d["sv"]), len(d["rec"])
    ref5.ref_hash_lr_name5.restype = C.c_uint64
    # the reference derives the read id from the query name: translate the store's name ids for its hook, keep plain ids for ours
    outs = []
    for fn, lead in ((ref5.ref_assemble_lr, ()), (H.dh_assemble_lr, (ctxh,))):
        so = np.zeros((n, 13), np.int32); srq = np.zeros(n, np.float32); co = np.zeros((n, 8192), np.uint8); cl = np.zeros(n, np.int32)
        al = np.zeros((n, 16384), np.uint8); all_ = np.zeros(n, np.int32)
        rc = fn(*lead, _p(d["contig"]), _p(d["coff"]), _p(d["clen"]), 2, _p(d["rec"]), nrec, _p(d["cig"]), _p(d["reads"]), _p(d["store"]), len(d["store"]), _p(d["sv"]), n,
                15, 2, C.c_float(0.9), 100, 10000, 300, _p(so), _p(srq), _p(co), 8192, _p(cl), _p(al), 16384, _p(all_))
        assert rc == 0, rc
        outs.append((so.copy(), srq.copy(), [co[i, :cl[i]].tobytes() for i in range(n)], [al[i, :all_[i]].tobytes() for i in range(n)]))
    e, g = outs
    for i in range(n):
        assert g[2][i] == e[2][i], (i, d["sv"][i].tolist(), len(g[2][i]), len(e[2][i]))
        assert g[0][i].tolist() == e[0][i].tolist(), (i, d["sv"][i].tolist(), g[0][i].tolist(), e[0][i].tolist())
        assert g[3][i] == e[3][i]
    assert np.array_equal(e[1].view(np.uint32), g[1].view(np.uint32))
    assert int(e[0][:, 8].sum()) >= n // 4
