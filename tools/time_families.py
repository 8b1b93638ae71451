#!/usr/bin/env python
"""Quick kernel-time probe of the K3 / K2 / K5 families (development aid; bench.py is the reported number)."""
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
if "k2" in which:
    b = synth.k2_msa_batch(int(os.environ.get("K2N", "2048")), seed=1002, fast=True)
    for i in range(3):
        ctx.msa(b["seqs"], b["read_off"], b["read_len"], b["cluster_off"]); k = ctx.last_kernel_ms()
    n = len(b["cluster_off"]) - 1
    print(f"K2 msa: {n} clusters kernel {k:.2f} ms -> {n/k*1e3:.0f} clusters/s")
