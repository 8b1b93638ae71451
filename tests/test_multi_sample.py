"""Several samples in one call set (`delly call a.bam b.bam`): discovery pools the files (per-file junction / pair scans concatenated before
sorting and clustering, split reads collected contig by contig and file by file), genotyping is per file over the joint SV list. The chain
of batched mirrors (dellySrCallMulti) against the same stage sequence of the reference's own functions compiled verbatim and run over
TWO in-memory alignment files (oracle/ref_wrap5.cpp::ref_delly_sr_call_multi)."""
import ctypes as C

import numpy as np

from test_host_genotype import LR_CFG, _hash_string, _simulate_lr_sample, _simulate_sr_sample


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_delly_sr_call_two_samples_matches_reference(hostdev, ref5):
    H, ctxh = hostdev
    d = _simulate_sr_sample(515, n_del=12, cov=26)
    # two sequencing runs of the same genome: the read pairs are dealt to two files (arenas shared, each file coordinate-sorted)
    rec = d["rec"]
    which = (rec[:, 11] % 5 < 2).astype(np.int64)         # 60 % of the pairs to file 0, 40 % to file 1
    rec2 = np.ascontiguousarray(np.concatenate([rec[which == 0], rec[which == 1]]))
    file_off = np.array([0, int((which == 0).sum()), len(rec)], np.uint32)
    lib = np.array([[100, 300, 15, 200, 400, 480], [100, 305, 18, 190, 420, 500]], np.int32)
    ref5.ref_hash_sr_name.restype = C.c_uint64
    names = [f"q{int(r[11])}".encode() for r in rec2]
    seeds = np.array([ref5.ref_hash_sr_name(nm, 1 if (int(r[2]) & 0x80) else 0) for nm, r in zip(names, rec2)], np.uint64)
    nh = np.array([_hash_string(nm.decode()) for nm in names], np.uint32)
    CAP = 256
    outs = []
    for w in ("ref", "ours"):
        sv = np.zeros((CAP, 20), np.int32); fmt = np.full((2, CAP, 14), -5, np.int32); gl = np.zeros((2, CAP, 3), np.float32)
        co = np.zeros((CAP, 1024), np.uint8); cl = np.zeros(CAP, np.int32)
        tail = (_p(file_off), 2, _p(d["cig"]), _p(d["reads"]), _p(lib), _p(sv), CAP, _p(fmt), _p(gl), _p(co), 1024, _p(cl))
        if w == "ref":
            n = ref5.ref_delly_sr_call_multi(_p(d["contig"]), _p(d["coff"]), _p(d["clen"]), 2, _p(rec2), *tail)
        else:
            n = H.dh_delly_sr_call_multi(ctxh, _p(d["contig"]), _p(d["coff"]), _p(d["clen"]), 2, _p(rec2), _p(seeds), _p(nh), *tail)
        assert n > 0, n
        outs.append((n, sv[:n].copy(), fmt[:, :n].copy(), gl[:, :n].copy(), [co[i, :cl[i]].tobytes() for i in range(n)]))
    e, g = outs
    assert e[0] == g[0]
    assert np.array_equal(e[1], g[1]), np.argwhere(e[1] != g[1])[:5]
    assert np.array_equal(e[2], g[2]), np.argwhere(e[2] != g[2])[:5]
    assert np.array_equal(e[3].view(np.uint32), g[3].view(np.uint32))
    assert e[4] == g[4]
    # the planted deletions are called from the pooled evidence, and the two files are genotyped separately
    found = sum(1 for s, en, zyg in d["truth"] if any(v[15] == 2 and v[16] == 1 and abs(int(v[1]) - s) <= 3 and abs(int(v[3]) - en) <= 3 for v in e[1]))
    assert found >= len(d["truth"]) - 2, (found, len(d["truth"]))
    assert (e[2] != -5).all() and not np.array_equal(e[2][0], e[2][1])
    assert (e[2][0][:, 10] + e[2][0][:, 11]).sum() > (e[2][1][:, 10] + e[2][1][:, 11]).sum()    # fewer junction reads in the smaller file


def test_delly_lr_call_two_samples_matches_reference(hostdev, ref5):
    """Two long-read files: every discovery stage sees the records contig by contig and file by file; genotyping is per file."""
    H, ctxh = hostdev
    d = _simulate_lr_sample(909, n_sv=12, cov=30)
    rec = d["rec"]
    which = (rec[:, 11] % 3 == 0).astype(np.int64)         # two thirds of the reads to file 0
    rec2 = np.ascontiguousarray(np.concatenate([rec[which == 0], rec[which == 1]]))
    file_off = np.array([0, int((which == 0).sum()), len(rec)], np.uint32)
    ref5.ref_hash_lr_name5.restype = C.c_uint64
    seeds = np.array([ref5.ref_hash_lr_name5(f"q{int(r[11])}".encode()) for r in rec2], np.uint64)
    CAP = 256
    outs = []
    for w in ("ref", "ours"):
        sv = np.zeros((CAP, 20), np.int32); fmt = np.full((2, CAP, 14), -5, np.int32); gl = np.zeros((2, CAP, 3), np.float32)
        co = np.zeros((CAP, 8192), np.uint8); cl = np.zeros(CAP, np.int32)
        tail = (_p(file_off), 2, _p(d["cig"]), _p(d["reads"]), _p(LR_CFG), C.c_float(0.9), C.c_float(0.5), _p(sv), CAP, _p(fmt), _p(gl), _p(co), 8192, _p(cl))
        if w == "ref":
            n = ref5.ref_delly_lr_call_multi(_p(d["contig"]), _p(d["coff"]), _p(d["clen"]), 2, _p(rec2), *tail)
        else:
            n = H.dh_delly_lr_call_multi(ctxh, _p(d["contig"]), _p(d["coff"]), _p(d["clen"]), 2, _p(rec2), _p(seeds), *tail)
        assert n > 0, n
        outs.append((n, sv[:n].copy(), fmt[:, :n].copy(), gl[:, :n].copy(), [co[i, :cl[i]].tobytes() for i in range(n)]))
    e, g = outs
    assert e[0] == g[0]
    assert np.array_equal(e[1], g[1]), np.argwhere(e[1] != g[1])[:5]
    assert np.array_equal(e[2], g[2]), np.argwhere(e[2] != g[2])[:5]
    assert np.array_equal(e[3].view(np.uint32), g[3].view(np.uint32))
    assert e[4] == g[4]
    found = sum(1 for s, k, z, zyg in d["truth"] if any(v[15] == k and v[16] == 1 and abs(int(v[1]) - s) <= 10 for v in e[1]))
    assert found >= len(d["truth"]) - 3, (found, len(d["truth"]))
    assert (e[2] != -5).all() and not np.array_equal(e[2][0], e[2][1])


import pytest

import delly_b200
from test_host_genotype import _vcf_case


@pytest.mark.parametrize("geno_mode", [0, 1])
def test_vcf_records_three_samples_match_reference(ref7, geno_mode):
    """The multi-sample record: one FORMAT column per sample (GT / GQ / PL / FT / read-depth / DR / DV / RR / RV / HP / PS / methylation), AC / AN
    over all genotypes, and the discovery filter on the ALT support summed over the samples — against vcfOutput run verbatim with three files."""
    H = delly_b200.hostlib()
    F = 3
    cases = [_vcf_case(seed=40 + f, n=150) for f in range(F)]
    d = cases[0]
    n, tl, sv, alle, al, cons, cl = (d[k] for k in ("n", "tl", "sv", "alle", "al", "cons", "cl"))
    # per-sample count maps: the quality lists of the three cases behind each other in one arena, offsets file-major
    q = np.concatenate([c["q"] for c in cases])
    base, offs = 0, [[], [], [], []]
    for c in cases:
        for k in range(4):
            offs[k].append(c["offs"][k].astype(np.uint32) + np.uint32(base))
        base += len(c["q"])
    offs = [np.ascontiguousarray(np.concatenate(o)) for o in offs]
    hp = np.ascontiguousarray(np.concatenate([c["hp"] for c in cases])); rc = np.ascontiguousarray(np.concatenate([c["rc"] for c in cases]))
    rng = np.random.default_rng(8)
    me = np.zeros((F, n, 16), np.int32)
    me[:, :, 0:8] = rng.choice([-1, 0, 37, 100], size=(F, n, 8)); me[:, :, 8:12] = rng.choice([-1, 0, 2, 9], size=(F, n, 4)); me[:, :, 12:16] = rng.choice([-1, 0, 3, 12], size=(F, n, 4))
    outs = []
    for fn in (ref7.ref_vcf_output_multi, H.dh_vcf_output_multi):
        out = np.zeros(1 << 21, np.uint8)
        L = fn(_p(tl), 3, _p(sv), n, _p(alle), 512, _p(al), _p(cons), 512, _p(cl), _p(q), _p(offs[0]), _p(offs[1]), _p(offs[2]), _p(offs[3]), _p(hp), _p(rc), geno_mode, _p(out),
               len(out), None, _p(me), 2, F)
        assert 0 < L < len(out)
        outs.append(out[:L].tobytes().decode().split("\n"))
    e, g = outs
    assert len(e) == len(g), (len(e), len(g))
    for a, b in zip(e, g):
        assert a == b, (a, b)
    assert sum(1 for l in e if l.startswith("S ")) == F
    recs = [l for l in e if l.startswith("R ")]
    assert len(recs) > n // 3 and all(l.split("F:GQ=")[1].split(";")[0].count(",") == F - 1 for l in recs)


def test_delly_sr_genotype_mode_two_samples_round_trip(hostdev, ref5):
    """Genotyping mode over two files: the two-sample discovery output as a site list reproduces the records and both samples' genotype fields."""
    from test_genotype_mode import BITS, _pack_sites
    H, ctxh = hostdev
    d = _simulate_sr_sample(515, n_del=12, cov=26)
    rec = d["rec"]
    which = (rec[:, 11] % 5 < 2).astype(np.int64)
    rec2 = np.ascontiguousarray(np.concatenate([rec[which == 0], rec[which == 1]]))
    file_off = np.array([0, int((which == 0).sum()), len(rec)], np.uint32)
    lib = np.array([[100, 300, 15, 200, 400, 480], [100, 305, 18, 190, 420, 500]], np.int32)
    ref5.ref_hash_sr_name.restype = C.c_uint64
    names = [f"q{int(r[11])}".encode() for r in rec2]
    seeds = np.array([ref5.ref_hash_sr_name(nm, 1 if (int(r[2]) & 0x80) else 0) for nm, r in zip(names, rec2)], np.uint64)
    nh = np.array([_hash_string(nm.decode()) for nm in names], np.uint32)
    CAP = 256

    def run(fn, extra):
        sv = np.zeros((CAP, 20), np.int32); fmt = np.full((2, CAP, 14), -5, np.int32); gl = np.zeros((2, CAP, 3), np.float32)
        co = np.zeros((CAP, 1024), np.uint8); cl = np.zeros(CAP, np.int32)
        n = fn(ctxh, _p(d["contig"]), _p(d["coff"]), _p(d["clen"]), 2, _p(rec2), _p(seeds), _p(nh), _p(file_off), 2, _p(d["cig"]), _p(d["reads"]), _p(lib), _p(sv), CAP,
               _p(fmt), _p(gl), _p(co), 1024, _p(cl), *extra)
        assert n > 0, n
        return n, sv[:n].copy(), fmt[:, :n].copy(), gl[:, :n].copy(), [co[i, :cl[i]].tobytes() for i in range(n)]

    e = run(H.dh_delly_sr_call_multi, ())
    rows, strings = [], []
    for i in range(e[0]):
        v = e[1][i]; svt = int(v[15])
        svtype = [b"INV", b"INV", b"DEL", b"DUP", b"INS"][svt] if svt < 5 else b"BND"
        ct = [b"3to3", b"5to5", b"3to5", b"5to3", b"NtoN"][svt if svt < 5 else svt - 5]
        mask = sum(1 << BITS[k] for k in ("SVMETHOD", "SVTYPE", "CT", "PE", "HOMLEN", "SR", "CIPOS", "CIEND", "MAPQ", "SRMAPQ", "SRQ"))
        mask |= (1 << BITS["SVLEN"]) if svt == 4 else (1 << BITS["INSLEN"])
        mask |= (1 << BITS["END"]) if svt < 5 else ((1 << BITS["CHR2"]) | (1 << BITS["POS2"]))
        if v[16]:
            mask |= (1 << BITS["CONSENSUS"]) | (1 << BITS["CONSBP"])
        rows.append([int(v[0]), int(v[1]) - 1, int(np.float32(v[12]).view(np.int32)), mask, int(v[8]), int(v[13]), int(v[13]), int(v[14]), int(v[9]), int(v[3]), int(v[3]),
                     int(v[17]), int(v[4]), int(v[5]), int(v[6]), int(v[7]), int(v[10]), int(v[11]), int(v[19]), 0, 1, int(v[16])])
        strings.append([b"N", b"<" + svtype + b">", b"EMBL.DELLYv1.3.3", svtype, ct, b"chr%d" % int(v[2]), e[4][i]])
    site, strs, off = _pack_sites(rows, strings)
    g = run(H.dh_delly_sr_genotype_multi, (_p(site), len(site), _p(strs), _p(off)))
    assert e[0] == g[0]
    assert np.array_equal(e[1], g[1]) and np.array_equal(e[2], g[2]) and np.array_equal(e[3].view(np.uint32), g[3].view(np.uint32)) and e[4] == g[4]
