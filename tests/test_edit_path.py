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


@pytest.mark.gpu
def test_cuda_path_flags_hirschberg_regime(ctx):
    """Sizes for which the reference switches to Hirschberg are reported, not silently traced back another way."""
    rng = np.random.default_rng(1)
    q = ALPHA[rng.integers(0, 4, size=3000)]; t = ALPHA[rng.integers(0, 4, size=3000)]
    arena, off, ln = synth.pack([q, t])
    d, st, en, ops, status = ctx.edit_path(arena, off[0:1].copy(), ln[0:1].copy(), off[1:2].copy(), ln[1:2].copy(), 0)
    assert status[0] == 2 and ops[0] == b""
