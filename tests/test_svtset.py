"""The `-t` option (c.svtset: compute only the listed SV types; src/shortpe.h:405,457-461,486,503; src/junction.h:463-475): both call chains
with a restricted type set against the reference's chains with the same restriction."""
import ctypes as C

import numpy as np
import pytest

from test_host_genotype import LR_CFG, _hash_string, _simulate_lr_sample, _simulate_sr_sample


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("types", [(2,), (0, 1, 3, 4)])
def test_delly_sr_call_with_type_restriction(hostdev, ref5, types):
    H, ctxh = hostdev
    mask = sum(1 << t for t in types)
    d = _simulate_sr_sample(2024)
    nrec = len(d["rec"])
    lib = np.array([100, 300, 15, 200, 400, 480], np.int32)
    ref5.ref_hash_sr_name.restype = C.c_uint64
    names = [f"q{int(r[11])}".encode() for r in d["rec"]]
    seeds = np.array([ref5.ref_hash_sr_name(nm, 1 if (int(r[2]) & 0x80) else 0) for nm, r in zip(names, d["rec"])], np.uint64)
    nh = np.array([_hash_string(nm.decode()) for nm in names], np.uint32)
    outs = []
    try:
        ref5.ref_set_svtset(mask); H.dh_set_svtset(mask)
        for which in ("ref", "ours"):
            sv = np.zeros((512, 20), np.int32); fmt = np.zeros((512, 14), np.int32); gl = np.zeros((512, 3), np.float32)
            co = np.zeros((512, 1024), np.uint8); cl = np.zeros(512, np.int32)
            common = (_p(d["cig"]), _p(d["reads"]), _p(lib), _p(sv), 512, _p(fmt), _p(gl), _p(co), 1024, _p(cl))
            if which == "ref":
                n = ref5.ref_delly_sr_call(_p(d["contig"]), _p(d["coff"]), _p(d["clen"]), 2, _p(d["rec"]), nrec, *common)
            else:
                n = H.dh_delly_sr_call(ctxh, _p(d["contig"]), _p(d["coff"]), _p(d["clen"]), 2, _p(d["rec"]), _p(seeds), _p(nh), nrec, *common)
            assert n >= 0, n
            outs.append((n, sv[:n].copy(), fmt[:n].copy(), gl[:n].copy()))
    finally:
        ref5.ref_set_svtset(0); H.dh_set_svtset(0)
    e, g = outs
    assert e[0] == g[0] and np.array_equal(e[1], g[1]) and np.array_equal(e[2], g[2]) and np.array_equal(e[3].view(np.uint32), g[3].view(np.uint32))
    assert set(e[1][:, 15].tolist()) <= set(types)
    if types == (2,):
        assert e[0] >= len(d["truth"]) - 2          # the deletions are all there
    else:
        assert (e[1][:, 15] != 2).all()              # and none of them when deletions are not asked for


@pytest.mark.parametrize("types", [(4,), (2,)])
def test_delly_lr_call_with_type_restriction(hostdev, ref5, types):
    H, ctxh = hostdev
    mask = sum(1 << t for t in types)
    d = _simulate_lr_sample(777)
    nrec = len(d["rec"])
    ref5.ref_hash_lr_name5.restype = C.c_uint64
    seeds = np.array([ref5.ref_hash_lr_name5(f"q{int(r[11])}".encode()) for r in d["rec"]], np.uint64)
    outs = []
    try:
        ref5.ref_set_svtset(mask); H.dh_set_svtset(mask)
        for which in ("ref", "ours"):
            sv = np.zeros((512, 20), np.int32); fmt = np.zeros((512, 14), np.int32); gl = np.zeros((512, 3), np.float32)
            co = np.zeros((512, 8192), np.uint8); cl = np.zeros(512, np.int32)
            tail = (_p(d["cig"]), _p(d["reads"]), _p(LR_CFG), C.c_float(0.9), C.c_float(0.5), _p(sv), 512, _p(fmt), _p(gl), _p(co), 8192, _p(cl))
            if which == "ref":
                n = ref5.ref_delly_lr_call(_p(d["contig"]), _p(d["coff"]), _p(d["clen"]), 2, _p(d["rec"]), nrec, *tail)
            else:
                n = H.dh_delly_lr_call(ctxh, _p(d["contig"]), _p(d["coff"]), _p(d["clen"]), 2, _p(d["rec"]), _p(seeds), nrec, *tail)
            assert n > 0, n
            outs.append((n, sv[:n].copy(), fmt[:n].copy(), gl[:n].copy()))
    finally:
        ref5.ref_set_svtset(0); H.dh_set_svtset(0)
    e, g = outs
    assert e[0] == g[0] and np.array_equal(e[1], g[1]) and np.array_equal(e[2], g[2]) and np.array_equal(e[3].view(np.uint32), g[3].view(np.uint32))
    assert set(e[1][:, 15].tolist()) == set(types)
    want = sum(1 for s, k, z, zyg in d["truth"] if k == types[0])
    assert e[0] >= want - 2
