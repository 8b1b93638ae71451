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


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["sr", "lr"])
def test_generate_probes_matches_reference(ctx, ref, ref3, mode):
    """_generateProbes (src/coverage.h:164-263): REF/ALT probes of both breakpoints, breakpoint regions (per contig, in the
    reference's order), symbolic alleles — for SVs refined by the reference's own alignConsensus, precise and imprecise."""
    H = delly_b200.hostlib()
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
    gr, gprobes, greg, gal, gon = run(H.dh_generate_probes, ctx.h)
    assert er == gr and er > n // 2
    assert eprobes == gprobes
    assert ereg == greg
    assert eal == gal and eon == gon
    assert sum(1 for p in eprobes if p[0]) > n // 4
