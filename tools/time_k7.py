"""K7 (msaEdlibBatch) timing aid: stage clock of the host mirror (DGPU_TIMING=1) for a given number of clusters.
usage: python tools/time_k7.py [clusters] [repeats]"""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import delly_b200
from delly_b200 import synth

ncl = int(sys.argv[1]) if len(sys.argv) > 1 else 48
rep = int(sys.argv[2]) if len(sys.argv) > 2 else 2
os.environ["DGPU_TIMING"] = "1"
ctx = delly_b200.Context(0)
H = delly_b200.hostlib()
rng = np.random.default_rng(2003)
reads, coff = [], [0]
for _ in range(ncl):
    L = int(rng.integers(2000, 4000)); base = synth._ACGT[rng.integers(0, 4, size=L + 200)]
    for _ in range(15):
        a = int(rng.integers(0, 100)); r = base[a:a + L + int(rng.integers(-40, 40))]
        reads.append(synth.mutate_fast(rng, r, sub=0.03, ins=0.02, dele=0.02))
    coff.append(len(reads))
arena, off, ln = synth.pack(reads)
coff = np.array(coff, np.uint32)
cons = np.zeros((ncl, 8192), np.uint8); clen = np.zeros(ncl, np.int32); rows = np.zeros(ncl, np.int32)
P = lambda a: C.c_void_p(a.ctypes.data)
for i in range(rep):
    t0 = time.perf_counter()
    rc = H.dh_msa_edlib_batch(ctx.h, P(arena), P(off), P(ln), P(coff), ncl, 2, P(cons), 8192, P(clen), P(rows))
    dt = time.perf_counter() - t0
    print(f"call {i}: rc {rc}, {dt * 1e3:.1f} ms, {ncl / dt:.1f} clusters/s", flush=True)
