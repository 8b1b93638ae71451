"""msa (src/msa.h:185-239): LCS/UPGMA/progressive gotoh/consensus. Oracle vs compiled reference (CPU),
CUDA vs oracle / golden (GPU)."""
import os

import numpy as np
import pytest

from delly_b200 import synth
from oracle import pyoracle as po

GOLD = os.path.join(os.path.dirname(__file__), "golden", "msa.npz")


def _clusters(b):
    for i in range(len(b["cluster_off"]) - 1):
        yield [b["seqs"][b["read_off"][r]: b["read_off"][r] + b["read_len"][r]].tobytes()
               for r in range(b["cluster_off"][i], b["cluster_off"][i + 1])]


def _special_batch():
    """Identical reads, N bases, single read, two reads, unrelated reads, very different lengths, homopolymers."""
    rng = np.random.default_rng(21)
    A = lambda n: synth._ACGT[rng.integers(0, 4, size=n)].tobytes()  # noqa: E731
    base = A(120)
    cl = [
        [base, base, base],
        [base],
        [base, base[10:] + A(10)],
        [A(80), A(90), A(100), A(70)],
        [base[:60], base, base[30:], base[50:110]],
        [b"A" * 50, b"A" * 70, b"A" * 30 + b"C" * 30],
        [base[:50] + b"N" * 5 + base[55:], base, base[5:] + b"NNNNN"],
        [b"ACGT" * 20, b"CGTA" * 20, b"GTAC" * 20, b"TACG" * 20, b"ACGT" * 19],
        [A(33), A(32), A(31), A(64), A(65), A(63)],
        [base] * 20,
        [A(300), A(280)],
        [A(10), A(200), A(12)],
    ]
    big = A(900)  # alignments wider than one 512-column DP strip, reads longer than the bit-parallel LCS width
    cl.append([big[:700], big[100:800], big[200:900], big[150:650]])
    cl.append([A(600), A(520)])
    cl.append([big[:520], big[400:900], big[300:560]])
    reads, coff = [], [0]
    for c in cl:
        reads += c
        coff.append(len(reads))
    arena, off, ln = synth.pack(reads)
    return dict(seqs=arena, read_off=off, read_len=ln, cluster_off=np.array(coff, np.uint32))


def test_oracle_matches_reference(ref):
    O = po.oracle()
    for b in (synth.k2_msa_batch(25, seed=4, read_len=80, max_off=50, err=0.02), _special_batch()):
        for reads in _clusters(b):
            for mc in (2, 3):
                a = po.msa(O, reads, mc, want_alignment=True)
                e = po.msa(ref, reads, mc, want_alignment=True)
                assert a == e, ([len(r) for r in reads], mc)


def test_oracle_gotoh_matches_reference(ref):
    """White-box: profile-profile gotoh on hand-made sub-alignments (gaps, different row counts)."""
    O = po.oracle()
    rng = np.random.default_rng(8)
    for _ in range(60):
        def rand_aln(rows, L):
            out = []
            for _ in range(rows):
                a = synth._ACGT[rng.integers(0, 4, size=L)].copy()
                lead, trail = int(rng.integers(0, L // 3)), int(rng.integers(0, L // 3))
                a[:lead] = ord("-"); a[L - trail:] = ord("-")
                a[rng.integers(lead, L - trail, size=2)] = ord("-")
                out.append(a.tobytes())
            return out
        a1 = rand_aln(int(rng.integers(1, 5)), int(rng.integers(12, 60)))
        a2 = rand_aln(int(rng.integers(1, 5)), int(rng.integers(12, 60)))
        # every column needs at least one covering row for a defined profile: patch with a full row
        a1.append(synth._ACGT[rng.integers(0, 4, size=len(a1[0]))].tobytes())
        a2.append(synth._ACGT[rng.integers(0, 4, size=len(a2[0]))].tobytes())
        assert po.gotoh(O, a1, a2) == po.gotoh(ref, a1, a2)
        assert po.consensus(O, a1 + [a1[0]], 2) == po.consensus(ref, a1 + [a1[0]], 2)


def test_golden_msa(oracle):
    g = np.load(GOLD, allow_pickle=False)
    b = {k: g[k] for k in ("seqs", "read_off", "read_len", "cluster_off")}
    for i, reads in enumerate(_clusters(b)):
        rows, cons, aln = po.msa(oracle, reads, 2, want_alignment=True)
        o, L = int(g["cons_off"][i]), int(g["cons_len"][i])
        assert cons == g["cons"][o:o + L].tobytes()
        assert rows == g["n_rows"][i]
        ao, aL = int(g["aln_off"][i]), int(g["aln_cols"][i])
        assert b"".join(aln) == g["aln"][ao:ao + rows * aL].tobytes()


def _check(ctx, b, mc=2):
    cons, nrows, status, alns = ctx.msa(b["seqs"], b["read_off"], b["read_len"], b["cluster_off"], min_clique=mc,
                                        want_alignment=True)
    O = po.oracle()
    for i, reads in enumerate(_clusters(b)):
        rows, ecs, ealn = po.msa(O, reads, mc, want_alignment=True)
        assert alns[i] == ealn, (i, [len(r) for r in reads])
        assert cons[i] == ecs, (i, [len(r) for r in reads])
        assert nrows[i] == rows


@pytest.mark.gpu
def test_cuda_special(ctx):
    _check(ctx, _special_batch())
    _check(ctx, _special_batch(), mc=3)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(80, 50, 0.02, 60), (150, 120, 0.005, 60), (150, 120, 0.03, 30), (250, 200, 0.01, 12)])
def test_cuda_matches_oracle(ctx, shape):
    rl, mo, err, n = shape
    _check(ctx, synth.k2_msa_batch(n, seed=31 + rl, read_len=rl, max_off=mo, err=err))


@pytest.mark.gpu
def test_cuda_golden(ctx):
    g = np.load(GOLD, allow_pickle=False)
    cons, nrows, status, alns = ctx.msa(g["seqs"], g["read_off"], g["read_len"], g["cluster_off"], want_alignment=True)
    for i in range(len(cons)):
        o, L = int(g["cons_off"][i]), int(g["cons_len"][i])
        assert cons[i] == g["cons"][o:o + L].tobytes()
        ao, aL = int(g["aln_off"][i]), int(g["aln_cols"][i])
        assert b"".join(alns[i]) == g["aln"][ao:ao + int(nrows[i]) * aL].tobytes()


@pytest.mark.gpu
def test_cuda_rejects_loudly(ctx):
    """Clusters the device path does not support must be reported, never silently mis-computed."""
    import delly_b200
    reads = [b"ACGTRACGT" * 5, b"ACGTACGT" * 5]
    arena, off, ln = synth.pack(reads)
    with pytest.raises(delly_b200.DgpuError):
        ctx.msa(arena, off, ln, np.array([0, 2], np.uint32))


@pytest.mark.gpu
def test_cuda_reads_longer_than_256_bp(ctx):
    """2 x 300 bp libraries: every read pair takes the plain-DP LCS path (both longer than the 256-column bit vectors); results as the reference's"""
    _check(ctx, synth.k2_msa_batch(6, seed=77, read_len=300, max_off=180, err=0.01))
    _check(ctx, synth.k2_msa_batch(4, seed=78, read_len=290, max_off=100, err=0.02), mc=3)


@pytest.mark.gpu
def test_cuda_oversize_cluster_fails_alone(ctx):
    """A cluster beyond the kernel's 32 reads (translocation clusters are uncapped in the reference, src/shortpe.h:140-146) is flagged with its
    own status; the other clusters of the batch are computed as always (the batched host mirrors then treat that one SV's consensus as failed)."""
    b = synth.k2_msa_batch(8, seed=79, read_len=100, max_off=60, err=0.01)
    rng = np.random.default_rng(80)
    big = [synth._ACGT[rng.integers(0, 4, size=100)].tobytes() for _ in range(60)]
    reads = [b["seqs"][b["read_off"][r]:b["read_off"][r] + b["read_len"][r]].tobytes() for r in range(len(b["read_off"]))]
    coff = list(b["cluster_off"])
    k = 3                                               # splice the big cluster in as cluster 3
    allreads = reads[:coff[k]] + big + reads[coff[k]:]
    newoff = coff[:k + 1] + [coff[k] + 60] + [c + 60 for c in coff[k + 1:]]
    arena, off, ln = synth.pack(allreads)
    cons, nrows, status = ctx.msa(arena, off, ln, np.array(newoff, np.uint32), check=False)
    assert status[k] == 1 and cons[k] == b"" and nrows[k] == 0
    O = po.oracle()
    for i in range(len(newoff) - 1):
        if i == k:
            continue
        rs = allreads[newoff[i]:newoff[i + 1]]
        rows, ecs, _ = po.msa(O, rs, 2)
        assert status[i] == 0 and cons[i] == ecs and nrows[i] == rows
