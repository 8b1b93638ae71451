"""The reference's own example data (example/sr.bam, example/lr.bam, example/ref.fa — the only fixtures it ships, SURVEY section 8c / 8d
config 1) through both call chains: library estimation, discovery and genotyping of `delly sr` on sr.bam, discovery and genotyping of
`delly lr` on lr.bam (tool defaults), the batched mirrors against the reference's stage functions compiled verbatim — every record field.
The records come from tests/golden/example_{sr,lr}.npz (the BAMs parsed by tests/golden/make_example_fixture.py)."""
import ctypes as C
import os

import numpy as np
import pytest

import delly_b200
from test_host_genotype import _hash_string

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _load(which):
    d = dict(np.load(os.path.join(GOLD, "example_%s.npz" % which)))
    # the hooks take two contig slots; the example genome has one contig ("18")
    d["contig"] = np.concatenate([d["contig"], np.frombuffer(b"ACGT" * 64, np.uint8)])
    d["coff"] = np.array([0, int(d["clen"][0])], np.uint32)
    d["clen"] = np.array([int(d["clen"][0]), 256], np.uint32)
    keep = d["rec"][:, 0] >= 0                       # unplaced records are behind every iterator of the reference
    d["rec"] = np.ascontiguousarray(d["rec"][keep])
    return d


def test_example_sr_library_and_call_chain(hostdev, ref5, ref6):
    H, ctxh = hostdev
    Hp = delly_b200.hostlib()
    d = _load("sr")
    rec, nrec = d["rec"], len(d["rec"])
    assert nrec > 25000
    libs = []
    for fn in (ref6.ref_get_library_params, Hp.dh_get_library_params):
        o = np.zeros(7, np.int32)
        assert fn(_p(d["clen"]), 2, _p(rec), nrec, _p(d["cig"]), 9, 5, _p(o)) == 0
        libs.append(o.tolist())
    assert libs[0] == libs[1] and libs[0][0] > 50 and libs[0][1] > 100, libs      # read size and insert-size median of the example library
    l = libs[0]
    lib = np.array([l[0], l[1], l[2], l[3], l[5], l[6]], np.int32)
    ref5.ref_hash_sr_name.restype = C.c_uint64
    names = [f"q{int(r[11])}".encode() for r in rec]
    seeds = np.array([ref5.ref_hash_sr_name(nm, 1 if (int(r[2]) & 0x80) else 0) for nm, r in zip(names, rec)], np.uint64)
    nh = np.array([_hash_string(nm.decode()) for nm in names], np.uint32)
    outs = []
    for which in ("ref", "ours"):
        sv = np.zeros((512, 20), np.int32); fmt = np.zeros((512, 14), np.int32); gl = np.zeros((512, 3), np.float32)
        co = np.zeros((512, 2048), np.uint8); cl = np.zeros(512, np.int32)
        common = (_p(d["cig"]), _p(d["reads"]), _p(lib), _p(sv), 512, _p(fmt), _p(gl), _p(co), 2048, _p(cl))
        if which == "ref":
            n = ref5.ref_delly_sr_call(_p(d["contig"]), _p(d["coff"]), _p(d["clen"]), 2, _p(rec), nrec, *common)
        else:
            n = H.dh_delly_sr_call(ctxh, _p(d["contig"]), _p(d["coff"]), _p(d["clen"]), 2, _p(rec), _p(seeds), _p(nh), nrec, *common)
        assert n > 0, n
        outs.append((n, sv[:n].copy(), fmt[:n].copy(), gl[:n].copy(), [co[i, :cl[i]].tobytes() for i in range(n)]))
    e, g = outs
    assert e[0] == g[0]
    assert np.array_equal(e[1], g[1]), np.argwhere(e[1] != g[1])[:5]
    assert np.array_equal(e[2], g[2]), np.argwhere(e[2] != g[2])[:5]
    assert np.array_equal(e[3].view(np.uint32), g[3].view(np.uint32))
    assert e[4] == g[4]
    assert (e[1][:, 16] == 1).any()          # precise (split-read) calls exist in the example


def test_example_lr_call_chain(hostdev, ref5):
    H, ctxh = hostdev
    d = _load("lr")
    rec, nrec = d["rec"], len(d["rec"])
    assert nrec > 800
    # `delly lr` defaults (src/tegua.h:227-266): map-qual 1, minclip 25, minrefsep 30, maxreadsep 500, clique 3, pruning 1000, max-reads 15, flank 100,
    # indel-size 10000, cons-window 1000, max-geno-count 250, read-cap 25; flank quality 0.9, extension 0.5
    cfg = np.array([1, 25, 30, 500, 3, 1000, 15, 100, 10000, 1000, 250, 25], np.int32)
    ref5.ref_hash_lr_name5.restype = C.c_uint64
    seeds = np.array([ref5.ref_hash_lr_name5(f"q{int(r[11])}".encode()) for r in rec], np.uint64)
    outs = []
    for which in ("ref", "ours"):
        sv = np.zeros((512, 20), np.int32); fmt = np.zeros((512, 14), np.int32); gl = np.zeros((512, 3), np.float32)
        co = np.zeros((512, 16384), np.uint8); cl = np.zeros(512, np.int32)
        tail = (_p(d["cig"]), _p(d["reads"]), _p(cfg), C.c_float(0.9), C.c_float(0.5), _p(sv), 512, _p(fmt), _p(gl), _p(co), 16384, _p(cl))
        if which == "ref":
            n = ref5.ref_delly_lr_call(_p(d["contig"]), _p(d["coff"]), _p(d["clen"]), 2, _p(rec), nrec, *tail)
        else:
            n = H.dh_delly_lr_call(ctxh, _p(d["contig"]), _p(d["coff"]), _p(d["clen"]), 2, _p(rec), _p(seeds), nrec, *tail)
        assert n > 0, n
        outs.append((n, sv[:n].copy(), fmt[:n].copy(), gl[:n].copy(), [co[i, :cl[i]].tobytes() for i in range(n)]))
    e, g = outs
    assert e[0] == g[0]
    assert np.array_equal(e[1], g[1]), np.argwhere(e[1] != g[1])[:5]
    assert np.array_equal(e[2], g[2]), np.argwhere(e[2] != g[2])[:5]
    assert np.array_equal(e[3].view(np.uint32), g[3].view(np.uint32))
    assert e[4] == g[4]


def test_example_lr_full_chain_with_annotation(hostdev, ref9, ref8):
    """lr.bam through the complete chain (annotation and methylation switched on; the example reads carry no MM / ML tags): records, alleles,
    SVAnno and MethylInfo identical to the reference's chain."""
    from test_svanno import _templates
    H, ctxh = hostdev
    d = _load("lr")
    rec, nrec = d["rec"], len(d["rec"])
    seqs, tpl_arena, tpl_off = _templates(ref8)
    cfg = np.array([1, 25, 30, 500, 3, 1000, 15, 100, 10000, 1000, 250, 25], np.int32)
    ref9.ref_hash_lr_name5.restype = C.c_uint64
    seeds = np.array([ref9.ref_hash_lr_name5(f"q{int(r[11])}".encode()) for r in rec], np.uint64)
    flags = np.zeros(nrec, np.uint8)
    zoff = np.zeros(nrec + 1, np.uint32); z8 = np.zeros(4, np.uint8)
    outs = []
    for which in ("ref", "ours"):
        sv = np.zeros((64, 20), np.int32); fmt = np.zeros((64, 14), np.int32); gl = np.zeros((64, 3), np.float32)
        co = np.zeros((64, 16384), np.uint8); cl = np.zeros(64, np.int32)
        an = np.full((64, 5), -9, np.int32); me = np.full((64, 16), -9, np.int32); al = np.zeros((64, 16384), np.uint8); all_ = np.zeros(64, np.int32)
        head = (_p(d["cig"]), _p(d["reads"]), _p(cfg), C.c_float(0.9), C.c_float(0.5), _p(sv), 64, _p(fmt), _p(gl), _p(co), 16384, _p(cl), _p(flags), _p(z8), _p(zoff), _p(z8),
                _p(zoff), 1000, 128, 1)
        tail = (C.c_float(0.8), C.c_float(0.85), _p(an), _p(me), _p(al), 16384, _p(all_))
        if which == "ref":
            n = ref9.ref_delly_lr_call_ex(_p(d["contig"]), _p(d["coff"]), _p(d["clen"]), 2, _p(rec), nrec, *head, *tail)
        else:
            n = H.dh_delly_lr_call_ex(ctxh, _p(d["contig"]), _p(d["coff"]), _p(d["clen"]), 2, _p(rec), _p(seeds), nrec, *head, _p(tpl_arena), _p(tpl_off), *tail)
        assert n > 0, n
        outs.append((n, sv[:n].copy(), fmt[:n].copy(), gl[:n].copy().view(np.uint32), [co[i, :cl[i]].tobytes() for i in range(n)], an[:n].copy(), me[:n].copy(),
                     [al[i, :all_[i]].tobytes() for i in range(n)]))
    e, g = outs
    assert e[0] == g[0]
    for k in (1, 2, 3, 5, 6):
        assert np.array_equal(e[k], g[k]), k
    assert e[4] == g[4] and e[7] == g[7]
    assert (e[5] != -9).all() and (e[6][:, 8:12] == 0).all()      # annotated; no methylation calls without tags
