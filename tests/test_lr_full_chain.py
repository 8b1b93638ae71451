"""The complete long-read chain — junction scan, clustering, assembly, neighbour de-duplication, genotyping WITH methylation (MM / ML tags) and
the reference-based annotation (mobile elements, tandem repeats, breakpoint homology) — against the same sequence of the reference's own
functions compiled verbatim (oracle/ref_wrap5.cpp built with ORACLE_FULL_LR: libdelly_ref9.so). Runs on the GPU through the real libraries
and, in the CPU suite, with the alignment entry points forwarded to the reference's own functions (tests/standin/host_standin.cpp)."""
import ctypes as C

import numpy as np

from test_host_genotype import LR_CFG, _simulate_lr_sample
from test_methyl import _with_tags
from test_svanno import _noisy, _rc, _templates


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_delly_lr_call_full_chain_matches_reference(hostdev, ref9, ref8):
    H, ctxh = hostdev
    seqs, tpl_arena, tpl_off = _templates(ref8)
    alu, l1, sva, numt, ltr, hervk, polya = seqs
    rng = np.random.default_rng(3)
    pool = [np.concatenate([alu, polya[:25]]), _rc(np.concatenate([alu, polya[:30]])), _noisy(rng, np.concatenate([l1[-900:], polya[:30]]), 0.03),
            np.tile(np.frombuffer(b"CAGGT", np.uint8), 60), ltr[:700], _noisy(rng, numt[5000:5600], 0.02), rng.choice(np.frombuffer(b"ACGT", np.uint8), size=500)]
    d = _with_tags(_simulate_lr_sample(4321, n_sv=13, cov=26, insert_pool=pool), 99)
    nrec = len(d["rec"])
    ref9.ref_hash_lr_name5.restype = C.c_uint64
    seeds = np.array([ref9.ref_hash_lr_name5(f"q{int(r[11])}".encode()) for r in d["rec"]], np.uint64)
    outs = []
    for which in ("ref", "ours"):
        sv = np.zeros((512, 20), np.int32); fmt = np.zeros((512, 14), np.int32); gl = np.zeros((512, 3), np.float32)
        co = np.zeros((512, 8192), np.uint8); cl = np.zeros(512, np.int32)
        an = np.full((512, 5), -9, np.int32); me = np.full((512, 16), -9, np.int32); al = np.zeros((512, 4096), np.uint8); all_ = np.zeros(512, np.int32)
        head = (_p(d["cig"]), _p(d["reads"]), _p(LR_CFG), C.c_float(0.9), C.c_float(0.5), _p(sv), 512, _p(fmt), _p(gl), _p(co), 8192, _p(cl), _p(d["tagflags"]), _p(d["mm"]),
                _p(d["mm_off"]), _p(d["ml"]), _p(d["ml_off"]), 400, 128, 1)
        tail = (C.c_float(0.8), C.c_float(0.85), _p(an), _p(me), _p(al), 4096, _p(all_))
        if which == "ref":
            n = ref9.ref_delly_lr_call_ex(_p(d["contig"]), _p(d["coff"]), _p(d["clen"]), 2, _p(d["rec"]), nrec, *head, *tail)
        else:
            n = H.dh_delly_lr_call_ex(ctxh, _p(d["contig"]), _p(d["coff"]), _p(d["clen"]), 2, _p(d["rec"]), _p(seeds), nrec, *head, _p(tpl_arena), _p(tpl_off), *tail)
        assert n > 0, n
        outs.append((n, sv[:n].copy(), fmt[:n].copy(), gl[:n].copy(), [co[i, :cl[i]].tobytes() for i in range(n)], an[:n].copy(), me[:n].copy(),
                     [al[i, :all_[i]].tobytes() for i in range(n)]))
    e, g = outs
    assert e[0] == g[0]
    assert np.array_equal(e[1], g[1]), np.argwhere(e[1] != g[1])[:5]
    assert np.array_equal(e[2], g[2]), np.argwhere(e[2] != g[2])[:5]
    assert np.array_equal(e[3].view(np.uint32), g[3].view(np.uint32))
    assert e[4] == g[4] and e[7] == g[7]
    assert np.array_equal(e[5], g[5]), (e[5][(e[5] != g[5]).any(axis=1)][:4], g[5][(e[5] != g[5]).any(axis=1)][:4])
    assert np.array_equal(e[6], g[6]), (e[6][(e[6] != g[6]).any(axis=1)][:4], g[6][(e[6] != g[6]).any(axis=1)][:4])
    # the planted events are called, the inserted mobile elements / repeats are recognised, and methylation is reported
    found = sum(1 for s, k, z, zyg in d["truth"] if any(v[15] == k and v[16] == 1 and abs(int(v[1]) - s) <= 60 for v in e[1]))
    assert found >= len(d["truth"]) - 3, (found, len(d["truth"]))
    types = set(e[5][:, 1].tolist())
    assert 1 in types and 7 in types and len(types - {0}) >= 3, types
    assert (e[5][:, 0] == 1).any() and (e[5][:, 2] > 0).any()
    assert (e[6] != -9).all() and (e[6][:, 0:8] >= 0).sum() > 2 * e[0]
