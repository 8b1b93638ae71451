import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, ctypes as C
import delly_b200
from delly_b200 import synth
from oracle import pyoracle as po
from test_host_split import _genome, _sv_cases, _p
ctx = delly_b200.Context(0)
H = delly_b200.hostlib(); R = po.ref(); O = po.oracle()
g1, g2 = _genome(11), _genome(12)
svs, cons = _sv_cases(14, g1, g2, n=160, cons_range=(200, 500))
cons = [synth.revcomp(c) if i % 2 else c for i, c in enumerate(cons)]
seqs = []
for i in range(len(svs)):
    sv = svs[i]
    seq, snd = (g1, g1) if sv[4] < 5 else (g2, g1)
    buf = C.create_string_buffer(20000); L = C.c_int()
    R.ref_get_sv_ref(seq, len(g1), snd, len(g2), _p(sv.copy()), len(cons[i]), 30, 10000, 300, buf, 20000, C.byref(L))
    ref = np.frombuffer(buf.raw[:L.value], np.uint8)
    seqs += [ref, cons[i], ref, synth.revcomp(cons[i])]
arena, off, ln = synth.pack(seqs)
b = dict(seqs=arena, q_off=off[0::2].copy(), q_len=ln[0::2].copy(), t_off=off[1::2].copy(), t_len=ln[1::2].copy())
k = np.full(len(b["q_off"]), -1, np.int32)
gd = ctx.edit_distance(b["seqs"], b["q_off"], b["q_len"], b["t_off"], b["t_len"], k, 0)
od, _ = po.edit_distance_batch(O, b["seqs"], b["q_off"], b["q_len"], b["t_off"], b["t_len"], k, 0, threads=8)
rd, _ = po.edit_distance_batch(R, b["seqs"], b["q_off"], b["q_len"], b["t_off"], b["t_len"], k, 0, threads=8)
bad = np.nonzero(gd != od)[0]
print("gpu!=oracle", len(bad), "oracle!=ref", int((od != rd).sum()))
for i in bad[:10]:
    print(i, b["q_len"][i], b["t_len"][i], gd[i], od[i], rd[i])
