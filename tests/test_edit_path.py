"""edlib PATH task (distance, start/end location, edit operations) in the traceback regime: CUDA vs the compiled reference."""
import numpy as np
import pytest

from delly_b200 import synth
from oracle import pyoracle as po

ALPHA = np.frombuffer(b"ACGT", np.uint8)


def _jobs(seed, n, qmax, tmax, mode):
    rng = np.random.default_rng(seed)
    seqs = []
    for _ in range(n):
        tl = int(rng.integers(1, tmax + 1))
        t = ALPHA[rng.integers(0, 4, size=tl)]
        r = rng.random()
        if r < 0.7 and tl > 6:
            if mode == 2:
                a = int(rng.integers(0, tl - 3)); b = int(rng.integers(a + 2, min(tl, a + qmax) + 1)); q = t[a:b]
            elif mode == 1:
                q = t[:int(rng.integers(2, min(tl, qmax) + 1))]
            else:
                q = t[:qmax]
            rate = float(rng.choice([0, 0.03, 0.1, 0.25]))
            q = synth.mutate(rng, q, sub=rate / 3, ins=rate / 3, dele=rate / 3)[:qmax]
            if len(q) == 0:
                q = ALPHA[rng.integers(0, 4, size=3)]
        else:
            q = ALPHA[rng.integers(0, 4, size=int(rng.integers(1, qmax + 1)))]
        if rng.random() < 0.15:  # repeats: many co-optimal paths, the tie rules decide
            q = np.resize(np.frombuffer(b"AC", np.uint8), len(q)); t = np.resize(np.frombuffer(b"AC", np.uint8), len(t))
        seqs += [q, t]
    arena, off, ln = synth.pack(seqs)
    return dict(seqs=arena, q_off=off[0::2].copy(), q_len=ln[0::2].copy(), t_off=off[1::2].copy(), t_len=ln[1::2].copy())


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("shape", [(60, 120, 500), (150, 300, 300), (300, 700, 120), (64, 64, 300)])
def test_cuda_path_matches_reference(ctx, ref, mode, shape):
    qmax, tmax, n = shape
    b = _jobs(300 + mode + qmax, n, qmax, tmax, mode)
    d, st, en, ops, status = ctx.edit_path(b["seqs"], b["q_off"], b["q_len"], b["t_off"], b["t_len"], mode)
    assert not status.any()
    for i in range(n):
        q = b["seqs"][b["q_off"][i]: b["q_off"][i] + b["q_len"][i]].tobytes()
        t = b["seqs"][b["t_off"][i]: b["t_off"][i] + b["t_len"][i]].tobytes()
        rd, re, rs, rops = po.edit_distance(ref, q, t, -1, mode, task=2)
        assert (d[i], en[i], st[i]) == (rd, re, rs), (i, mode, len(q), len(t))
        assert ops[i] == rops, (i, mode, len(q), len(t))


def _long_jobs(seed, n, lo, hi, mode, rate):
    rng = np.random.default_rng(seed)
    seqs = []
    for _ in range(n):
        tl = int(rng.integers(lo, hi)); t = ALPHA[rng.integers(0, 4, size=tl)]
        if mode == 2:
            a = int(rng.integers(0, tl // 3)); q = t[a:a + int(rng.integers(tl // 3, 2 * tl // 3))]
        elif mode == 1:
            q = t[:int(rng.integers(tl // 2, tl))]
        else:
            q = t
        q = synth.sub_noise(rng, q, rate)
        # a few indels
        for _ in range(int(rng.integers(0, 12))):
            p = int(rng.integers(1, len(q) - 1))
            q = np.concatenate([q[:p], ALPHA[rng.integers(0, 4, size=int(rng.integers(1, 30)))], q[p:]]) if rng.random() < 0.5 else np.concatenate([q[:p], q[p + int(rng.integers(1, 30)):]])
        if rng.random() < 0.2:
            q = np.resize(np.frombuffer(b"ACG", np.uint8), len(q)); t = np.resize(np.frombuffer(b"ACG", np.uint8), len(t))  # repeats: ties
        seqs += [q, t]
    arena, off, ln = synth.pack(seqs)
    return dict(seqs=arena, q_off=off[0::2].copy(), q_len=ln[0::2].copy(), t_off=off[1::2].copy(), t_len=ln[1::2].copy())


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_cuda_path_hirschberg_regime_matches_reference(ctx, ref, mode):
    """Problems above edlib's 1 MiB switch: the device splits them with the same Hirschberg recursion (1-3 levels here)."""
    b = _long_jobs(900 + mode, 14, 1800, 5200, mode, 0.06)
    d, st, en, ops, status = ctx.edit_path(b["seqs"], b["q_off"], b["q_len"], b["t_off"], b["t_len"], mode)
    assert not status.any()
    nsplit = 0
    for i in range(len(d)):
        q = b["seqs"][b["q_off"][i]: b["q_off"][i] + b["q_len"][i]].tobytes()
        t = b["seqs"][b["t_off"][i]: b["t_off"][i] + b["t_len"][i]].tobytes()
        rd, re, rs, rops = po.edit_distance(ref, q, t, -1, mode, task=2)
        assert (d[i], en[i], st[i]) == (rd, re, rs), (i, mode, len(q), len(t))
        assert ops[i] == rops, (i, mode, len(q), len(t))
        nsplit += 20 * ((len(q) + 63) // 64) * (re - rs + 1) + 8 * (re - rs + 1) >= 1 << 20
    assert nsplit >= 4


IUPAC_EQ = b"MAMCRARGWAWTBAB-SCSGYCYTDCD-KGKTEGE-FTF-"  # the 20 pairs of src/assemble.h:425


@pytest.mark.gpu
@pytest.mark.parametrize("size", [(80, 200), (1500, 4000)])
def test_cuda_path_with_iupac_equalities_matches_reference(ctx, ref, size):
    """msaEdlib's call (src/assemble.h:447): read vs IUPAC/gapped consensus string, NW PATH with 20 additional equalities."""
    rng = np.random.default_rng(size[0])
    seqs = []
    amb = np.frombuffer(b"MRWBSYDKEF-", np.uint8)
    for _ in range(60 if size[0] < 1000 else 8):
        tl = int(rng.integers(*size)); t = ALPHA[rng.integers(0, 4, size=tl)].copy()
        q = synth.mutate(rng, t, sub=0.03, ins=0.01, dele=0.01)
        k = max(1, tl // 15)
        t[rng.integers(0, tl, size=k)] = amb[rng.integers(0, len(amb), size=k)]
        seqs += [q, t]
    arena, off, ln = synth.pack(seqs)
    b = dict(seqs=arena, q_off=off[0::2].copy(), q_len=ln[0::2].copy(), t_off=off[1::2].copy(), t_len=ln[1::2].copy())
    d, st, en, ops, status = ctx.edit_path(b["seqs"], b["q_off"], b["q_len"], b["t_off"], b["t_len"], 0, eq=IUPAC_EQ)
    assert not status.any()
    for i in range(len(d)):
        q = b["seqs"][b["q_off"][i]: b["q_off"][i] + b["q_len"][i]].tobytes()
        t = b["seqs"][b["t_off"][i]: b["t_off"][i] + b["t_len"][i]].tobytes()
        rd, re, rs, rops = po.edit_distance(ref, q, t, -1, 0, task=2, eq=IUPAC_EQ)
        assert (d[i], en[i], st[i]) == (rd, re, rs), (i, len(q), len(t))
        assert ops[i] == rops, (i, len(q), len(t))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("size", [(40, 300), (800, 2600)])
def test_cuda_infix_path_with_iupac_equalities_matches_reference(ctx, ref, mode, size):
    """msaWfa's call (src/assemble.h:693): read placed inside an IUPAC/gapped consensus, HW PATH with the 20 additional
    equalities (SHW covered too). Some queries carry IUPAC / gap bytes themselves so the exact compare path runs."""
    rng = np.random.default_rng(size[0] + mode)
    seqs = []
    amb = np.frombuffer(b"MRWBSYDKEF-", np.uint8)
    for j in range(50 if size[0] < 500 else 8):
        tl = int(rng.integers(*size)); t = ALPHA[rng.integers(0, 4, size=tl)].copy()
        a = int(rng.integers(0, tl // 3)); e = int(rng.integers(2 * tl // 3, tl))
        q = synth.mutate(rng, t[a:e] if mode == 2 else t[:e], sub=0.03, ins=0.01, dele=0.01)
        k = max(1, tl // 15)
        t[rng.integers(0, tl, size=k)] = amb[rng.integers(0, len(amb), size=k)]
        if j % 5 == 0 and len(q) > 4:
            q = q.copy(); q[rng.integers(0, len(q), size=3)] = amb[rng.integers(0, len(amb), size=3)]
        seqs += [q, t]
    arena, off, ln = synth.pack(seqs)
    b = dict(seqs=arena, q_off=off[0::2].copy(), q_len=ln[0::2].copy(), t_off=off[1::2].copy(), t_len=ln[1::2].copy())
    d, st, en, ops, status = ctx.edit_path(b["seqs"], b["q_off"], b["q_len"], b["t_off"], b["t_len"], mode, eq=IUPAC_EQ)
    assert not status.any()
    for i in range(len(d)):
        q = b["seqs"][b["q_off"][i]: b["q_off"][i] + b["q_len"][i]].tobytes()
        t = b["seqs"][b["t_off"][i]: b["t_off"][i] + b["t_len"][i]].tobytes()
        rd, re, rs, rops = po.edit_distance(ref, q, t, -1, mode, task=2, eq=IUPAC_EQ)
        assert (d[i], en[i], st[i]) == (rd, re, rs), (i, mode, len(q), len(t))
        assert ops[i] == rops, (i, mode, len(q), len(t))


@pytest.mark.gpu
@pytest.mark.parametrize("eq", [b"", IUPAC_EQ])
def test_cuda_path_band_and_wavefront_segments_mixed(ctx, ref, eq):
    """One batch whose segments take every route of edit_path.cu: similar pairs (banded bit-parallel half columns and leaves, classes G=2..16),
    pairs with a long gap or very different lengths, unrelated pairs (no band class certifies them: wavefront kernels), short pairs
    (no split), lengths on block boundaries. Paths must be the reference's, op for op."""
    rng = np.random.default_rng(4343 if eq else 4242)
    amb = np.frombuffer(b"MRWBSYDKEF-N", np.uint8)
    seqs = []
    for i in range(40):
        tl = int(rng.choice([64, 65, 127, 128, 129, 700, 1024, 1500, 2048, 2500, 3300, 4100]))
        t = ALPHA[rng.integers(0, 4, size=tl)].copy()
        r = i % 5
        if r == 0: q = synth.sub_noise(rng, t, float(rng.choice([0.0, 0.01, 0.05])))
        elif r == 1: q = synth.mutate(rng, t, sub=0.03, ins=0.02, dele=0.02)
        elif r == 2:   # one long gap
            g = int(rng.integers(10, max(11, tl // 3))); a = int(rng.integers(0, tl - g))
            q = synth.sub_noise(rng, np.concatenate([t[:a], t[a + g:]]), 0.02)
        elif r == 3: q = ALPHA[rng.integers(0, 4, size=int(rng.integers(max(2, tl // 2), tl + 200)))]
        else: q = synth.mutate(rng, t, sub=0.08, ins=0.04, dele=0.04)
        if eq:
            k = max(1, tl // 20)
            t[rng.integers(0, tl, size=k)] = amb[rng.integers(0, len(amb), size=k)]
        if len(q) == 0: q = t[:1]
        seqs += [q, t]
    arena, off, ln = synth.pack(seqs)
    b = dict(seqs=arena, q_off=off[0::2].copy(), q_len=ln[0::2].copy(), t_off=off[1::2].copy(), t_len=ln[1::2].copy())
    d, st, en, ops, status = ctx.edit_path(b["seqs"], b["q_off"], b["q_len"], b["t_off"], b["t_len"], 0, eq=eq)
    assert not status.any()
    for i in range(len(d)):
        q = b["seqs"][b["q_off"][i]: b["q_off"][i] + b["q_len"][i]].tobytes()
        t = b["seqs"][b["t_off"][i]: b["t_off"][i] + b["t_len"][i]].tobytes()
        rd, re, rs, rops = po.edit_distance(ref, q, t, -1, 0, task=2, eq=eq)
        assert (d[i], en[i], st[i]) == (rd, re, rs), (i, len(q), len(t))
        assert ops[i] == rops, (i, len(q), len(t), d[i])
