#!/usr/bin/env python
"""Short driver for ncu captures: a few launches of every kernel family (never used for bench numbers)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import delly_b200  # noqa: E402
from delly_b200 import synth  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
ctx = delly_b200.Context(0)
if which in ("all", "k1"):
    import bench
    b = bench.make_batch(2_000_000, 1001)
    for _ in range(3):
        ctx.edit_distance(b["seqs"], b["q_off"], b["q_len"], b["t_off"], b["t_len"], b["k"], 2)
if which in ("all", "k3"):
    b = synth.k3_consref_batch(4096, seed=1003, fast=True)
    for _ in range(2):
        ctx.long_needle(b["seqs"], b["c_off"], b["c_len"], b["r_off"], b["r_len"])
if which in ("all", "k2"):
    b = synth.k2_msa_batch(int(os.environ.get("K2N", "1024")), seed=1002, fast=True)
    for _ in range(2):
        ctx.msa(b["seqs"], b["read_off"], b["read_len"], b["cluster_off"])
if which == "k5":
    b = synth.k3_consref_batch(int(os.environ.get("K5N", "148")), seed=2001, cons_range=(2000, 4000), err=0.05, fast=True, genome_len=4_000_000)
    ctx.long_needle(b["seqs"], b["c_off"], b["c_len"], b["r_off"], b["r_len"])
ctx.close()
print("done")
