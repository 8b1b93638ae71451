"""Generates the committed golden vectors from the COMPILED REFERENCE (oracle/_ref/libdelly_ref.so).

Run in the dev container (where /root/reference exists):  python tests/golden/make_golden.py
The reference ships no golden vectors of its own; these pin the oracle and the CUDA path to the
behaviour of the reference's own code on fixed seeded inputs.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import pyoracle as po  # noqa: E402


def golden_edit_distance(R):
    from test_edit_distance import _random_jobs
    out = {}
    for mode in (0, 1, 2):
        b = _random_jobs(4242 + mode, 600, 180, 240, mode, weird=True)
        n = len(b["q_off"])
        dist = np.empty(n, np.int32); end = np.empty(n, np.int32)
        for i in range(n):
            q = b["seqs"][b["q_off"][i]: b["q_off"][i] + b["q_len"][i]].tobytes()
            t = b["seqs"][b["t_off"][i]: b["t_off"][i] + b["t_len"][i]].tobytes()
            d, e, _, _ = po.edit_distance(R, q, t, int(b["k"][i]), mode)
            dist[i] = d; end[i] = e
        for k, v in b.items():
            out[f"m{mode}_{k}"] = v
        out[f"m{mode}_dist"] = dist
        out[f"m{mode}_end"] = end
    np.savez_compressed(os.path.join(HERE, "edit_distance.npz"), **out)


def golden_long_needle(R):
    from delly_b200 import synth
    from test_long_needle import _edge_batch, _jobs
    b1 = synth.k3_consref_batch(60, seed=77, cons_range=(60, 200), ref_cap=700)
    b2 = _edge_batch(seed=9)
    seqs = []
    for b in (b1, b2):
        for c, r in _jobs(b):
            if len(c) and len(r):
                seqs += [c, r]
    arena, off, ln = synth.pack(seqs)
    out = dict(seqs=arena, c_off=off[0::2].copy(), c_len=ln[0::2].copy(), r_off=off[1::2].copy(), r_len=ln[1::2].copy())
    oks, alns, aoff, alen = [], [], [], []
    pos = 0
    for c, r in _jobs(out):
        ok, r0, r1 = po.long_needle(R, c, r)
        oks.append(int(ok)); aoff.append(pos); alen.append(len(r0))
        alns.append(np.frombuffer(r0 + r1, np.uint8)); pos += 2 * len(r0)
    out.update(ok=np.array(oks, np.uint8), aln=np.concatenate(alns), aln_off=np.array(aoff, np.uint64), aln_len=np.array(alen, np.uint32))
    np.savez_compressed(os.path.join(HERE, "long_needle.npz"), **out)


def golden_msa(R):
    from delly_b200 import synth
    from test_msa import _special_batch, _clusters
    b1 = synth.k2_msa_batch(30, seed=99, read_len=100, max_off=70, err=0.01)
    b2 = _special_batch()
    reads, coff = [], [0]
    for b in (b1, b2):
        for cl in _clusters(b):
            reads += cl; coff.append(len(reads))
    arena, off, ln = synth.pack(reads)
    out = dict(seqs=arena, read_off=off, read_len=ln, cluster_off=np.array(coff, np.uint32))
    cons, coffs, clen, nrows, alns, aoff, acols = [], [], [], [], [], [], []
    cp = ap = 0
    for cl in _clusters(out):
        rows, cs, aln = po.msa(R, cl, 2, want_alignment=True)
        cons.append(np.frombuffer(cs, np.uint8)); coffs.append(cp); clen.append(len(cs)); cp += len(cs)
        nrows.append(rows); flat = b"".join(aln)
        alns.append(np.frombuffer(flat, np.uint8)); aoff.append(ap); acols.append(len(aln[0])); ap += len(flat)
    out.update(cons=np.concatenate(cons), cons_off=np.array(coffs, np.uint64), cons_len=np.array(clen, np.uint32),
               n_rows=np.array(nrows, np.uint32), aln=np.concatenate(alns), aln_off=np.array(aoff, np.uint64),
               aln_cols=np.array(acols, np.uint32))
    np.savez_compressed(os.path.join(HERE, "msa.npz"), **out)


if __name__ == "__main__":
    R = po.ref()
    assert R is not None, "needs oracle/_ref (build in the dev container)"
    which = sys.argv[1:] or ["edit_distance", "long_needle", "msa"]
    for w in which:
        globals()["golden_" + w](R)
        print("wrote", w)
