"""The sharded (one process per GPU) call chains of delly_b200/host/gather.hpp (SURVEY section 8e / VERDICT r1 item 2): the SV list cut into
contiguous per-rank ranges by cost, every rank finishing its own range, ONE all-gatherv of the full records, ids = local index + exclusive
prefix sum, order restored by concatenation. The N-rank result must equal the 1-rank result record for record — and through it the
reference's (the 1-rank chain is pinned against the reference's stage sequence in tests/test_host_genotype.py).

  * in-process simulation (dh_delly_sr_call_sharded_sim): N ranks as N host threads, each with its own context ([gpu]: N contexts on one B200;
    [edlib-standin]: the CPU stand-in), exchanging through a barrier — discovery and genotyping mode, N = 2, 3, 5;
  * world-size-2 gloo: two real processes, the exchange through a callback that all-gathers with torch.distributed (CPU stand-in);
  * the wire format round trip.
The NCCL exchange itself (dgpu_gather_records) runs in bench.py --gpus N (needs N GPUs) and is checked there against the 1-rank result."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import delly_b200
from test_host_genotype import _hash_string, _simulate_sr_sample

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _inputs(ref5, seed=2024):
    d = _simulate_sr_sample(seed)
    ref5.ref_hash_sr_name.restype = C.c_uint64
    names = [f"q{int(r[11])}".encode() for r in d["rec"]]
    d["seeds"] = np.array([ref5.ref_hash_sr_name(nm, 1 if (int(r[2]) & 0x80) else 0) for nm, r in zip(names, d["rec"])], np.uint64)
    d["nh"] = np.array([_hash_string(nm.decode()) for nm in names], np.uint32)
    d["lib"] = np.array([100, 300, 15, 200, 400, 480], np.int32)
    return d


def _bufs(cap=512, stride=1024):
    return dict(sv=np.zeros((cap, 20), np.int32), fmt=np.zeros((cap, 14), np.int32), gl=np.zeros((cap, 3), np.float32), co=np.zeros((cap, stride), np.uint8),
                cl=np.zeros(cap, np.int32), cap=cap, stride=stride)


def _result(n, b):
    assert n > 0, n
    return (n, b["sv"][:n].copy(), b["fmt"][:n].copy(), b["gl"][:n].copy().view(np.uint32), [b["co"][i, :b["cl"][i]].tobytes() for i in range(n)])


def _same(a, b):
    assert a[0] == b[0]
    for k in (1, 2, 3):
        assert np.array_equal(a[k], b[k]), (k, np.argwhere(a[k] != b[k])[:5])
    assert a[4] == b[4]


def _contexts(hostdev_param, H, ctxh, n):
    """n context handles: the stand-in's token n times, or n real contexts on cuda:0"""
    if hostdev_param == "edlib-standin":
        return [ctxh] * n, []
    made = [delly_b200.Context(0) for _ in range(n)]
    return [m.h for m in made], made


def _sites_from(e):
    """the discovery result as the site rows dh_delly_sr_genotype takes (as tests/test_genotype_mode.py builds them), in shuffled file order"""
    from test_genotype_mode import BITS, _pack_sites
    rows, strings = [], []
    for i in np.random.default_rng(0).permutation(e[0]):
        v = e[1][i]; svt = int(v[15])
        svtype = [b"INV", b"INV", b"DEL", b"DUP", b"INS"][svt] if svt < 5 else b"BND"
        ct = [b"3to3", b"5to5", b"3to5", b"5to3", b"NtoN"][svt if svt < 5 else svt - 5]
        mask = sum(1 << BITS[k] for k in ("SVMETHOD", "SVTYPE", "CT", "PE", "HOMLEN", "SR", "CIPOS", "CIEND", "MAPQ", "SRMAPQ", "SRQ"))
        mask |= (1 << BITS["SVLEN"]) if svt == 4 else (1 << BITS["INSLEN"])
        mask |= (1 << BITS["END"]) if svt < 5 else ((1 << BITS["CHR2"]) | (1 << BITS["POS2"]))
        if v[16]:
            mask |= (1 << BITS["CONSENSUS"]) | (1 << BITS["CONSBP"])
        rows.append([int(v[0]), int(v[1]) - 1, int(np.float32(v[12]).view(np.int32)), mask, int(v[8]), int(v[13]), int(v[13]), int(v[14]), int(v[9]), int(v[3]), int(v[3]),
                     int(v[17]), int(v[4]), int(v[5]), int(v[6]), int(v[7]), int(v[10]), int(v[11]), int(v[19]), 0, 1, int(v[16])])
        strings.append([b"N", b"<" + svtype + b">", b"EMBL.DELLYv1.3.3", svtype, ct, b"chr%d" % int(v[2]), e[4][i]])
    return _pack_sites(rows, strings)


@pytest.mark.parametrize("nranks", [2, 4])
def test_sharded_genotyping_mode_equals_single_rank(hostdev, ref5, request, nranks):
    """`delly sr -v sites.bcf` sharded by site ranges (BASELINE configs[3] shape): the discovery calls as a shuffled site list"""
    H, ctxh = hostdev
    param = request.node.callspec.params["hostdev"]
    d = _inputs(ref5)
    nrec = len(d["rec"])
    head = (_p(d["contig"]), _p(d["coff"]), _p(d["clen"]), 2, _p(d["rec"]), _p(d["seeds"]), _p(d["nh"]), nrec, _p(d["cig"]), _p(d["reads"]), _p(d["lib"]))
    disc = _bufs()
    nd = H.dh_delly_sr_call(ctxh, *head, _p(disc["sv"]), disc["cap"], _p(disc["fmt"]), _p(disc["gl"]), _p(disc["co"]), disc["stride"], _p(disc["cl"]))
    e = _result(nd, disc)
    e = (e[0], e[1], e[2], disc["gl"][:nd].copy(), e[4])
    site, strs, off = _sites_from(e)
    one = _bufs()
    n1 = H.dh_delly_sr_genotype(ctxh, *head, _p(one["sv"]), one["cap"], _p(one["fmt"]), _p(one["gl"]), _p(one["co"]), one["stride"], _p(one["cl"]), _p(site), len(site), _p(strs),
                                _p(off))
    r1 = _result(n1, one)
    handles, keep = _contexts(param, H, ctxh, nranks)
    arr = (C.c_void_p * nranks)(*[h if isinstance(h, int) else h.value for h in handles])
    sh = _bufs()
    ns = H.dh_delly_sr_call_sharded_sim(arr, nranks, *head, _p(site), len(site), _p(strs), _p(off), _p(sh["sv"]), sh["cap"], _p(sh["fmt"]), _p(sh["gl"]), _p(sh["co"]),
                                        sh["stride"], _p(sh["cl"]))
    assert ns != -100, "ranks disagree with each other"
    _same(r1, _result(ns, sh))
    for m in keep:
        m.close()


@pytest.mark.parametrize("nranks", [2, 3, 5])
def test_sharded_discovery_equals_single_rank(hostdev, ref5, request, nranks):
    H, ctxh = hostdev
    param = request.node.callspec.params["hostdev"]
    d = _inputs(ref5)
    nrec = len(d["rec"])
    one = _bufs()
    n1 = H.dh_delly_sr_call(ctxh, _p(d["contig"]), _p(d["coff"]), _p(d["clen"]), 2, _p(d["rec"]), _p(d["seeds"]), _p(d["nh"]), nrec, _p(d["cig"]), _p(d["reads"]), _p(d["lib"]),
                            _p(one["sv"]), one["cap"], _p(one["fmt"]), _p(one["gl"]), _p(one["co"]), one["stride"], _p(one["cl"]))
    r1 = _result(n1, one)
    handles, keep = _contexts(param, H, ctxh, nranks)
    arr = (C.c_void_p * nranks)(*[h if isinstance(h, int) else h.value for h in handles])
    sh = _bufs()
    ns = H.dh_delly_sr_call_sharded_sim(arr, nranks, _p(d["contig"]), _p(d["coff"]), _p(d["clen"]), 2, _p(d["rec"]), _p(d["seeds"]), _p(d["nh"]), nrec, _p(d["cig"]),
                                        _p(d["reads"]), _p(d["lib"]), None, 0, None, None, _p(sh["sv"]), sh["cap"], _p(sh["fmt"]), _p(sh["gl"]), _p(sh["co"]), sh["stride"],
                                        _p(sh["cl"]))
    assert ns != -100, "ranks disagree with each other"
    _same(r1, _result(ns, sh))
    assert (r1[1][:, 16] == 1).sum() >= 8          # precise calls went through the sharded assembly
    for m in keep:
        m.close()


def test_payload_round_trip():
    H = delly_b200.hostlib()
    for nsv, ns, seed in ((0, 1, 1), (1, 1, 2), (37, 2, 3), (400, 3, 4)):
        assert H.dh_shard_payload_roundtrip(nsv, ns, seed) > 0


def test_partition_is_contiguous_and_balanced():
    H = delly_b200.hostlib()
    rng = np.random.default_rng(5)
    for n, r in ((0, 4), (3, 8), (1000, 8), (17, 2)):
        cost = rng.integers(1, 10_000, size=max(n, 1)).astype(np.uint64)[:n]
        b = np.zeros(r + 1, np.uint64)
        H.dh_partition_by_cost(_p(cost) if n else None, n, r, _p(b))
        assert b[0] == 0 and b[-1] == n and (np.diff(b.astype(np.int64)) >= 0).all()
        if n >= 100:
            tot = cost.sum(); parts = [cost[int(b[i]):int(b[i + 1])].sum() for i in range(r)]
            assert max(parts) <= tot / r + cost.max()


_WORKER = r'''
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, sys.argv[3]); sys.path.insert(0, os.path.join(sys.argv[3], "tests"))
import torch, torch.distributed as dist
rank, world = int(sys.argv[1]), int(sys.argv[2])
os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = sys.argv[4]
dist.init_process_group("gloo", rank=rank, world_size=world)
from delly_b200 import gather as dg
from oracle import pyoracle
import test_sharded as T
ref5 = pyoracle.ref5()
lib = C.CDLL(sys.argv[5])
lib.standin_ctx.restype = C.c_void_p
libc = C.CDLL(None); libc.malloc.restype = C.c_void_p; libc.malloc.argtypes = [C.c_size_t]
CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int))
def gather_cb(user, local, nbytes, all_p, counts_p, n_p):
    parts = dg.all_gather_bytes(C.string_at(local, nbytes) if nbytes else b"")
    blob = b"".join(parts)
    a = libc.malloc(max(len(blob), 1)); C.memmove(a, blob, len(blob))
    cnt = libc.malloc(8 * len(parts)); C.memmove(cnt, np.array([len(p) for p in parts], np.uint64).tobytes(), 8 * len(parts))
    all_p[0] = a; counts_p[0] = cnt; n_p[0] = len(parts)
    return 0
cb = CB(gather_cb)
d = T._inputs(ref5)
b = T._bufs()
p = T._p
n = lib.dh_delly_sr_call_sharded(C.c_void_p(lib.standin_ctx()), p(d["contig"]), p(d["coff"]), p(d["clen"]), 2, p(d["rec"]), p(d["seeds"]), p(d["nh"]), len(d["rec"]), p(d["cig"]),
                                 p(d["reads"]), p(d["lib"]), None, 0, None, None, rank, world, cb, None, p(b["sv"]), b["cap"], p(b["fmt"]), p(b["gl"]), p(b["co"]), b["stride"],
                                 p(b["cl"]), None)
assert n > 0, n
np.savez(sys.argv[6] + ".%d.npz" % rank, n=n, sv=b["sv"][:n], fmt=b["fmt"][:n], gl=b["gl"][:n], cl=b["cl"][:n], co=b["co"][:n])
dist.barrier(); dist.destroy_process_group()
'''


def test_world2_gloo_processes_equal_single_rank(standin, ref5, tmp_path):
    """two real processes; the exchange is torch.distributed (gloo) behind the gather callback"""
    d = _inputs(ref5)
    one = _bufs()
    ctxh = C.c_void_p(standin.standin_ctx())
    n1 = standin.dh_delly_sr_call(ctxh, _p(d["contig"]), _p(d["coff"]), _p(d["clen"]), 2, _p(d["rec"]), _p(d["seeds"]), _p(d["nh"]), len(d["rec"]), _p(d["cig"]), _p(d["reads"]),
                                  _p(d["lib"]), _p(one["sv"]), one["cap"], _p(one["fmt"]), _p(one["gl"]), _p(one["co"]), one["stride"], _p(one["cl"]))
    r1 = _result(n1, one)
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    so = os.path.join(HERE, "standin", "_build", "libhost_standin.so")
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), str(r), "2", ROOT, port, so, str(tmp_path / "out")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    for r in range(2):
        z = np.load(str(tmp_path / "out") + ".%d.npz" % r)
        n = int(z["n"])
        got = (n, z["sv"], z["fmt"], z["gl"].view(np.uint32), [z["co"][i, :z["cl"][i]].tobytes() for i in range(n)])
        _same(r1, got)


@pytest.mark.parametrize("nranks", [2, 3])
def test_sharded_lr_chain_equals_single_rank(hostdev, ref9, ref8, request, nranks):
    """`delly lr` as N ranks (dellyLrCallSharded): the per-SV consensus queue and the genotyping / annotation cut by cost, records exchanged —
    SVs, genotype fields, likelihoods, consensus, alleles, annotation and methylation must equal the single-rank chain (itself pinned against the
    reference's in tests/test_lr_full_chain.py on the same sample)."""
    from test_host_genotype import LR_CFG, _simulate_lr_sample
    from test_methyl import _with_tags
    from test_svanno import _noisy, _rc, _templates
    H, ctxh = hostdev
    param = request.node.callspec.params["hostdev"]
    seqs, tpl_arena, tpl_off = _templates(ref8)
    alu, l1, sva, numt, ltr, hervk, polya = seqs
    rng = np.random.default_rng(3)
    pool = [np.concatenate([alu, polya[:25]]), _rc(np.concatenate([alu, polya[:30]])), _noisy(rng, np.concatenate([l1[-900:], polya[:30]]), 0.03),
            np.tile(np.frombuffer(b"CAGGT", np.uint8), 60), ltr[:700], _noisy(rng, numt[5000:5600], 0.02), rng.choice(np.frombuffer(b"ACGT", np.uint8), size=500)]
    d = _with_tags(_simulate_lr_sample(4321, n_sv=13, cov=26, insert_pool=pool), 99)
    nrec = len(d["rec"])
    ref9.ref_hash_lr_name5.restype = C.c_uint64
    seeds = np.array([ref9.ref_hash_lr_name5(f"q{int(r[11])}".encode()) for r in d["rec"]], np.uint64)
    handles, keep = _contexts(param, H, ctxh, nranks)
    arr = (C.c_void_p * nranks)(*[h if isinstance(h, int) else h.value for h in handles])
    outs = []
    for which in ("one", "sharded"):
        sv = np.zeros((512, 20), np.int32); fmt = np.zeros((512, 14), np.int32); gl = np.zeros((512, 3), np.float32)
        co = np.zeros((512, 8192), np.uint8); cl = np.zeros(512, np.int32)
        an = np.full((512, 5), -9, np.int32); me = np.full((512, 16), -9, np.int32); al = np.zeros((512, 4096), np.uint8); all_ = np.zeros(512, np.int32)
        head = (_p(d["cig"]), _p(d["reads"]), _p(LR_CFG), C.c_float(0.9), C.c_float(0.5), _p(sv), 512, _p(fmt), _p(gl), _p(co), 8192, _p(cl), _p(d["tagflags"]), _p(d["mm"]),
                _p(d["mm_off"]), _p(d["ml"]), _p(d["ml_off"]), 400, 128, 1)
        tail = (C.c_float(0.8), C.c_float(0.85), _p(an), _p(me), _p(al), 4096, _p(all_))
        if which == "one":
            n = H.dh_delly_lr_call_ex(ctxh, _p(d["contig"]), _p(d["coff"]), _p(d["clen"]), 2, _p(d["rec"]), _p(seeds), nrec, *head, _p(tpl_arena), _p(tpl_off), *tail)
        else:
            n = H.dh_delly_lr_call_ex_sharded_sim(arr, nranks, _p(d["contig"]), _p(d["coff"]), _p(d["clen"]), 2, _p(d["rec"]), _p(seeds), nrec, *head, _p(tpl_arena),
                                                  _p(tpl_off), *tail)
            assert n != -100, "ranks disagree with each other"
        assert n > 0, n
        outs.append((n, sv[:n].copy(), fmt[:n].copy(), gl[:n].copy().view(np.uint32), [co[i, :cl[i]].tobytes() for i in range(n)], an[:n].copy(), me[:n].copy(),
                     [al[i, :all_[i]].tobytes() for i in range(n)]))
    e, g = outs
    assert e[0] == g[0] and e[0] >= 8
    for k in (1, 2, 3, 5, 6):
        assert np.array_equal(e[k], g[k]), (k, np.argwhere(e[k] != g[k])[:5])
    assert e[4] == g[4] and e[7] == g[7]
    for m in keep:
        m.close()
