#!/usr/bin/env python
"""Quick kernel-time probe of the K3 / K2 / K5 families and of the annotateSV (k9) / merge-identity (k10) job shapes (development aid; bench.py is
the reported number). k9 / k10 were added after the round-1 GPU budget was spent: not yet run."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import delly_b200
from delly_b200 import synth
ctx = delly_b200.Context(0); ctx.set_profiling(True)
which = sys.argv[1:] or ["k3", "k2"]
if "k3" in which:
    b = synth.k3_consref_batch(4096, seed=1003, fast=True)
    rep = 4
    co, cl, ro, rl = [np.tile(b[k], rep) for k in ("c_off", "c_len", "r_off", "r_len")]
    cells = int(((cl.astype(np.int64) + 1) * (rl.astype(np.int64) + 1)).sum())
    for i in range(3):
        t = time.time(); ctx.long_needle(b["seqs"], co, cl, ro, rl); dt = time.time() - t
        k = ctx.last_kernel_ms()
    print(f"K3 longNeedle: {len(co)} jobs kernel {k:.2f} ms -> {len(co)/k*1e3:.0f} aln/s, {3*cells/k/1e6:.1f} GCUPS (host call {dt*1e3:.0f} ms)")
if "k3w" in which:  # the widest short-read windows (n 1300..2000)
    b = synth.k3_consref_batch(4096, seed=1004, cons_range=(340, 500), fast=True)
    co, cl, ro, rl = [np.tile(b[k], 2) for k in ("c_off", "c_len", "r_off", "r_len")]
    cells = int(((cl.astype(np.int64) + 1) * (rl.astype(np.int64) + 1)).sum())
    for i in range(3):
        ctx.long_needle(b["seqs"], co, cl, ro, rl); k = ctx.last_kernel_ms()
    print(f"K3w longNeedle wide: {len(co)} jobs (m~{cl.mean():.0f}, n~{rl.mean():.0f}, n max {rl.max()}) kernel {k:.2f} ms -> {len(co)/k*1e3:.0f} aln/s, {3*cells/k/1e6:.1f} GCUPS")
if "k5" in which:
    b = synth.k3_consref_batch(int(os.environ.get("K5N", "64")), seed=2001, cons_range=(2000, 4000), err=0.05, fast=True, genome_len=4_000_000)
    co, cl, ro, rl = [b[k] for k in ("c_off", "c_len", "r_off", "r_len")]
    cells = int(((cl.astype(np.int64) + 1) * (rl.astype(np.int64) + 1)).sum())
    for i in range(2):
        ctx.long_needle(b["seqs"], co, cl, ro, rl); k = ctx.last_kernel_ms()
    print(f"K5 longNeedle lr: {len(co)} jobs (m~{cl.mean():.0f}, n~{rl.mean():.0f}) kernel {k:.2f} ms -> {len(co)/k*1e3:.1f} aln/s, {3*cells/k/1e6:.1f} GCUPS")
if "k6" in which:  # long-read NW distance pairs (banded passes)
    rng = np.random.default_rng(2002)
    seqs = []
    for _ in range(20000):
        L = int(rng.integers(200, 4001)); t = synth._ACGT[rng.integers(0, 4, size=L, dtype=np.uint8)]
        seqs += [synth.sub_noise(rng, t, 0.08), t]
    arena, off, ln = synth.pack(seqs)
    rep = int(os.environ.get("K6REP", "10"))
    qo, ql, to, tl_ = [np.tile(x, rep) for x in (off[0::2].copy(), ln[0::2].copy(), off[1::2].copy(), ln[1::2].copy())]
    for i in range(3):
        ctx.edit_distance(arena, qo, ql, to, tl_, None, 0); k = ctx.last_kernel_ms()
    print(f"K6 NW distance: {len(qo)} jobs kernels {k:.2f} ms -> {len(qo)/k*1e3:.0f} alignments/s")
if "k2" in which:
    b = synth.k2_msa_batch(int(os.environ.get("K2N", "2048")), seed=1002, fast=True)
    for i in range(3):
        ctx.msa(b["seqs"], b["read_off"], b["read_len"], b["cluster_off"]); k = ctx.last_kernel_ms()
    n = len(b["cluster_off"]) - 1
    print(f"K2 msa: {n} clusters kernel {k:.2f} ms -> {n/k*1e3:.0f} clusters/s")
if "k9" in which:  # annotateSV shape (src/svanno.h:124-170): every insertion against six templates, both strands, HW distance, unbounded
    rng = np.random.default_rng(3001)
    ACGT = np.frombuffer(b"ACGT", np.uint8)
    tlen = [281 + 111, 6018 + 111, 2901 + 111, 16569, 968, 9472]           # template lengths of the reference (Alu / LINE1 / SVA + polyA, NUMT, LTR, HERVK)
    tpl = [ACGT[rng.integers(0, 4, size=n)] for n in tlen for _ in (0, 1)]
    nins = int(os.environ.get("K9N", "2000"))
    ins = [ACGT[rng.integers(0, 4, size=int(rng.integers(300, 6000)))] for _ in range(nins)]
    arena, off, ln = synth.pack(tpl + ins)
    qo, ql, to, tl_ = [], [], [], []
    for i in range(nins):
        io, il = off[12 + i], ln[12 + i]
        for t in range(12):
            if il > ln[t]: qo.append(off[t]); ql.append(ln[t]); to.append(io); tl_.append(il)
            else: qo.append(io); ql.append(il); to.append(off[t]); tl_.append(ln[t])
    qo, ql, to, tl_ = [np.array(x, np.uint32) for x in (qo, ql, to, tl_)]
    cells = int((ql.astype(np.int64) * tl_.astype(np.int64)).sum())
    for i in range(3):
        ctx.edit_distance(arena, qo, ql, to, tl_, None, 2); k = ctx.last_kernel_ms()
    print(f"K9 annotateSV HW: {len(qo)} jobs ({nins} insertions x 12) kernel {k:.2f} ms -> {nins/k*1e3:.0f} insertions/s, {cells/k/1e6:.1f} GCUPS-equivalent")
if "k10" in which:  # delly merge identity shape (src/merge.h:210-223): NW distance under k = (1 - 0.9) * length
    rng = np.random.default_rng(3002)
    ACGT = np.frombuffer(b"ACGT", np.uint8)
    n = int(os.environ.get("K10N", "200000"))
    seqs = []
    for i in range(n):
        a = ACGT[rng.integers(0, 4, size=int(rng.integers(50, 1500)))]
        b = synth.mutate(rng, a, sub=0.02, ins=0.01, dele=0.01) if i % 2 else ACGT[rng.integers(0, 4, size=int(rng.integers(50, 1500)))]
        seqs += [a, b]
    arena, off, ln = synth.pack(seqs)
    qo, ql, to, tl_ = off[0::2].copy(), ln[0::2].copy(), off[1::2].copy(), ln[1::2].copy()
    kk = (0.1 * np.maximum(ql, tl_)).astype(np.int32)
    for i in range(3):
        ctx.edit_distance(arena, qo, ql, to, tl_, kk, 0); k = ctx.last_kernel_ms()
    print(f"K10 merge identity NW (bounded): {n} pairs kernel {k:.2f} ms -> {n/k*1e3:.0f} pairs/s")
