"""Genotyping mode (`delly call -v sites.bcf`, BASELINE config 4): the site list of a Delly BCF turned back into SV records
(vcfParse, src/modvcf.h:156-339, compiled VERBATIM over an in-memory BCF reader stand-in: oracle/ref_wrap7.cpp::ref_vcf_parse) and the
chain sites -> annotateCoverage -> genotype fields (dellySrGenotype), checked by a round trip through the discovery path."""
import ctypes as C

import numpy as np
import pytest

import delly_b200
from test_host_genotype import LR_CFG, _hash_string, _simulate_lr_sample, _simulate_sr_sample

BITS = dict(SVMETHOD=0, SVTYPE=1, CT=2, PE=3, INSLEN=4, SVLEN=5, HOMLEN=6, SR=7, END=8, CHR2=9, POS2=10, CONSENSUS=11, CONSBP=12, CIPOS=13, CIEND=14, MAPQ=15, SRMAPQ=16,
            SRQ=17, ALLELEID=18, NALLELE=19)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _pack_sites(rows, strings):
    site = np.array(rows, np.int32).reshape(-1, 22)
    flat = [s for seven in strings for s in seven]
    off = np.zeros(len(flat) + 1, np.uint32)
    off[1:] = np.cumsum([len(s) for s in flat])
    return np.ascontiguousarray(site), np.frombuffer(b"".join(flat) + b"\0", np.uint8).copy(), off


def _random_sites(seed, n, method_first=b"EMBL.DELLYv1.3.3"):
    rng = np.random.default_rng(seed)
    rows, strings = [], []
    for i in range(n):
        svtype = [b"DEL", b"DUP", b"INV", b"INS", b"BND", b"CNV", b"FOO"][int(rng.choice(7, p=[0.25, 0.12, 0.15, 0.15, 0.2, 0.05, 0.08]))]
        ct = [b"3to3", b"5to5", b"3to5", b"5to3", b"NtoN", b"bad"][int(rng.integers(0, 6))]
        mask = 0
        for name, p in (("SVMETHOD", 0.97), ("SVTYPE", 0.95), ("CT", 0.95), ("PE", 0.6), ("INSLEN", 0.6), ("SVLEN", 0.8), ("HOMLEN", 0.6), ("SR", 0.6), ("END", 0.85),
                        ("CHR2", 0.8), ("POS2", 0.8), ("CONSENSUS", 0.7), ("CONSBP", 0.8), ("CIPOS", 0.7), ("CIEND", 0.7), ("MAPQ", 0.7), ("SRMAPQ", 0.7), ("SRQ", 0.7),
                        ("ALLELEID", 0.3), ("NALLELE", 0.6)):
            if rng.random() < p:
                mask |= 1 << BITS[name]
        if i == 0:
            mask |= 1 << BITS["SVMETHOD"]
        pos0 = int(rng.integers(0, 90000))
        ref = [b"A", b"N", b"ACGTACGTAC", b".", b"", b"ACG"][int(rng.integers(0, 6))]
        alt = [b"<DEL>", b"<INS>", b"A", b"ACGTTTTGCA", b".", b"", b"A]chr1:1234]", b"[chr0:77[C", b"<DUP>"][int(rng.integers(0, 9))]
        chr2 = [b"chr0", b"chr1", b"chr2", b"chrUn"][int(rng.integers(0, 4))]
        cons = bytes(np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=int(rng.integers(0, 300)))])
        method = method_first if (i == 0 or rng.random() < 0.9) else b"OTHER"
        qual = np.float32(rng.choice([0.0, 12.7, 60.0, 300.0, 10000.0])).view(np.int32)
        srq = np.float32(rng.choice([0.0, 0.87, 1.0])).view(np.int32)
        rows.append([int(rng.integers(0, 3)), pos0, int(qual), mask, int(rng.integers(0, 30)), int(rng.integers(0, 200)), int(rng.integers(1, 500)), int(rng.integers(0, 40)),
                     int(rng.integers(0, 30)), pos0 + int(rng.integers(1, 5000)), int(rng.integers(1, 90000)), int(rng.integers(0, 300)), -int(rng.integers(0, 60)),
                     int(rng.integers(0, 60)), -int(rng.integers(0, 60)), int(rng.integers(0, 60)), int(rng.choice([0, 37, 60, 255, 256, 300])), int(rng.choice([0, 20, 60, 260])),
                     int(srq), int(rng.integers(0, 9)), int(rng.integers(1, 4)), int(rng.random() < 0.6)])
        strings.append([ref, alt, method, svtype, ct, chr2, cons])
    return _pack_sites(rows, strings)


def _parse(fn, site, strs, off, has_consbp=1, ncontig=3):
    sv = np.full((len(site) + 1, 22), -77, np.int32)
    al = np.zeros((len(site) + 1, 64), np.uint8); alen = np.zeros(len(site) + 1, np.int32)
    co = np.zeros((len(site) + 1, 512), np.uint8); clen = np.zeros(len(site) + 1, np.int32)
    n = fn(ncontig, _p(site), len(site), _p(strs), _p(off), has_consbp, _p(sv), len(sv), _p(al), 64, _p(alen), _p(co), 512, _p(clen))
    assert n >= 0
    return n, sv[:n].copy(), [al[i, :alen[i]].tobytes() for i in range(n)], [co[i, :clen[i]].tobytes() for i in range(n)]


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_vcf_parse_matches_reference(ref7, seed):
    """Every field vcfParse fills, for random site lists: all SV types and orientations (incl. undecodable ones), every INFO key present or
    absent, sequence-resolved / symbolic / breakend / missing alleles, unknown mate contigs, MAPQ beyond uint8, records it skips."""
    H = delly_b200.hostlib()
    site, strs, off = _random_sites(seed, 400)
    e = _parse(ref7.ref_vcf_parse, site, strs, off)
    g = _parse(H.dh_vcf_parse, site, strs, off)
    assert e[0] == g[0] and 250 < e[0] < 400
    assert np.array_equal(e[1], g[1]), np.argwhere(e[1] != g[1])[:5]
    assert e[2] == g[2] and e[3] == g[3]
    assert len(set(e[1][:, 15].tolist())) >= 9 and (e[1][:, 16] == 0).any() and (e[1][:, 2] == -1).any()


@pytest.mark.parametrize("case", ["not-delly", "no-consbp-key", "late-foreign-record"])
def test_vcf_parse_stops_like_reference(ref7, case):
    H = delly_b200.hostlib()
    site, strs, off = _random_sites(9, 50, method_first=b"GATK" if case == "not-delly" else b"EMBL.DELLYv1.3.3")
    if case == "late-foreign-record":
        site[0, 3] &= ~1   # the first record carries no SVMETHOD: the file never qualifies before a record has to be parsed
    e = _parse(ref7.ref_vcf_parse, site, strs, off, has_consbp=0 if case == "no-consbp-key" else 1)
    g = _parse(H.dh_vcf_parse, site, strs, off, has_consbp=0 if case == "no-consbp-key" else 1)
    assert e[0] == g[0] == 0


def test_delly_sr_genotype_mode_round_trip(hostdev, ref5):
    """Discovery output written as a site list and genotyped again in genotyping mode reproduces every record and every genotype field of
    the discovery run (same SVs, same alignments): vcfParseSites + sort / renumber + annotateCoverage + genotype fields as one chain."""
    H, ctxh = hostdev
    d = _simulate_sr_sample(2024)
    nrec = len(d["rec"])
    lib = np.array([100, 300, 15, 200, 400, 480], np.int32)
    ref5.ref_hash_sr_name.restype = C.c_uint64
    names = [f"q{int(r[11])}".encode() for r in d["rec"]]
    seeds = np.array([ref5.ref_hash_sr_name(nm, 1 if (int(r[2]) & 0x80) else 0) for nm, r in zip(names, d["rec"])], np.uint64)
    nh = np.array([_hash_string(nm.decode()) for nm in names], np.uint32)

    def run(fn, extra):
        sv = np.zeros((512, 20), np.int32); fmt = np.zeros((512, 14), np.int32); gl = np.zeros((512, 3), np.float32)
        co = np.zeros((512, 1024), np.uint8); cl = np.zeros(512, np.int32)
        n = fn(ctxh, _p(d["contig"]), _p(d["coff"]), _p(d["clen"]), 2, _p(d["rec"]), _p(seeds), _p(nh), nrec, _p(d["cig"]), _p(d["reads"]), _p(lib), _p(sv), 512, _p(fmt),
               _p(gl), _p(co), 1024, _p(cl), *extra)
        assert n > 0, n
        return n, sv[:n].copy(), fmt[:n].copy(), gl[:n].copy(), [co[i, :cl[i]].tobytes() for i in range(n)]

    e = run(H.dh_delly_sr_call, ())
    rows, strings = [], []
    order = np.random.default_rng(0).permutation(e[0])      # file order need not be sorted: the chain sorts and renumbers
    for i in order:
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
    g = run(H.dh_delly_sr_genotype, (_p(site), len(site), _p(strs), _p(off)))
    assert e[0] == g[0] and e[0] >= 10
    assert np.array_equal(e[1], g[1]), np.argwhere(e[1] != g[1])[:5]
    assert np.array_equal(e[2], g[2]), np.argwhere(e[2] != g[2])[:5]
    assert np.array_equal(e[3].view(np.uint32), g[3].view(np.uint32))
    assert e[4] == g[4]


def test_delly_lr_genotype_mode_round_trip(hostdev, ref5):
    """Long-read genotyping mode: the discovery output as a site list (file order = the sorted call order, ids are not renumbered in this mode)
    reproduces the records and the genotype fields of the discovery run."""
    H, ctxh = hostdev
    d = _simulate_lr_sample(777)
    nrec = len(d["rec"])
    ref5.ref_hash_lr_name5.restype = C.c_uint64
    seeds = np.array([ref5.ref_hash_lr_name5(f"q{int(r[11])}".encode()) for r in d["rec"]], np.uint64)
    sv = np.zeros((512, 20), np.int32); fmt = np.zeros((512, 14), np.int32); gl = np.zeros((512, 3), np.float32)
    co = np.zeros((512, 8192), np.uint8); cl = np.zeros(512, np.int32)
    n = H.dh_delly_lr_call(ctxh, _p(d["contig"]), _p(d["coff"]), _p(d["clen"]), 2, _p(d["rec"]), _p(seeds), nrec, _p(d["cig"]), _p(d["reads"]), _p(LR_CFG), C.c_float(0.9),
                           C.c_float(0.5), _p(sv), 512, _p(fmt), _p(gl), _p(co), 8192, _p(cl))
    assert n >= 8
    rows, strings = [], []
    for i in range(n):
        v = sv[i]; svt = int(v[15])
        svtype = [b"INV", b"INV", b"DEL", b"DUP", b"INS"][svt] if svt < 5 else b"BND"
        ct = [b"3to3", b"5to5", b"3to5", b"5to3", b"NtoN"][svt if svt < 5 else svt - 5]
        mask = sum(1 << BITS[k] for k in ("SVMETHOD", "SVTYPE", "CT", "PE", "HOMLEN", "SR", "CIPOS", "CIEND", "MAPQ", "SRMAPQ", "SRQ", "END"))
        mask |= (1 << BITS["SVLEN"]) if svt == 4 else (1 << BITS["INSLEN"])
        if v[16]:
            mask |= (1 << BITS["CONSENSUS"]) | (1 << BITS["CONSBP"])
        rows.append([int(v[0]), int(v[1]) - 1, int(np.float32(v[12]).view(np.int32)), mask, int(v[8]), int(v[13]), int(v[13]), int(v[14]), int(v[9]), int(v[3]), int(v[3]),
                     int(v[17]), int(v[4]), int(v[5]), int(v[6]), int(v[7]), int(v[10]), int(v[11]), int(v[19]), 0, 1, int(v[16])])
        strings.append([b"N", b"<" + svtype + b">", b"EMBL.DELLYv1.3.3", svtype, ct, b"chr%d" % int(v[2]), co[i, :cl[i]].tobytes()])
    site, strs, off = _pack_sites(rows, strings)
    sv2 = np.zeros((512, 20), np.int32); fmt2 = np.zeros((512, 14), np.int32); gl2 = np.zeros((512, 3), np.float32)
    n2 = H.dh_delly_lr_genotype(ctxh, _p(d["contig"]), _p(d["coff"]), _p(d["clen"]), 2, _p(d["rec"]), nrec, _p(d["cig"]), _p(d["reads"]), _p(LR_CFG), C.c_float(0.9), _p(site),
                                len(site), _p(strs), _p(off), _p(sv2), 512, _p(fmt2), _p(gl2))
    assert n2 == n
    assert np.array_equal(sv[:n], sv2[:n]), np.argwhere(sv[:n] != sv2[:n])[:5]
    assert np.array_equal(fmt[:n], fmt2[:n]), np.argwhere(fmt[:n] != fmt2[:n])[:5]
    assert np.array_equal(gl[:n].view(np.uint32), gl2[:n].view(np.uint32))
