"""Host C++ mirror of clustering / junction selection (delly_b200/host/cluster.hpp, junction.hpp) against the
reference's cluster.h / junction.h compiled verbatim (oracle/_ref/libdelly_ref2.so). Pure host logic: no GPU."""
import ctypes as C

import numpy as np
import pytest

import delly_b200
from oracle import pyoracle as po


def _p(a):
    return C.c_void_p(a.ctypes.data)


@pytest.fixture(scope="module")
def libs():
    r2 = po.ref2()
    if r2 is None:
        pytest.fail("oracle/_ref/libdelly_ref2.so not available: run __graft_entry__.build() where /root/reference exists")
    return delly_b200.hostlib(), r2


def _sr_records(seed, svt, n_clusters=40, noise=200, nchr=3):
    rng = np.random.default_rng(seed)
    rows, ids = [], []
    rid = 10
    for _ in range(n_clusters):
        chr_ = int(rng.integers(0, nchr))
        chr2 = chr_ if svt < 5 else int(rng.integers(0, nchr))
        p = int(rng.integers(1000, 200000))
        size = int(rng.integers(20, 5000)) if svt != 4 else 1
        ins = int(rng.integers(20, 400)) if svt == 4 else 0
        k = int(rng.integers(1, 12))
        jit = int(rng.choice([2, 10, 60]))
        for _ in range(k):
            rows.append([chr_, p + int(rng.integers(-jit, jit + 1)), chr2, p + size + int(rng.integers(-jit, jit + 1)),
                         int(rng.integers(0, 1000)), int(rng.integers(0, 150)), int(rng.integers(0, 61)),
                         ins + int(rng.integers(-5, 6)) if svt == 4 else int(rng.integers(0, 30))])
            rid += int(rng.choice([1, 2, 3]))  # consecutive ids = mates of one pair
            ids.append(rid)
    for _ in range(noise):
        chr_ = int(rng.integers(0, nchr)); p = int(rng.integers(1000, 200000))
        rows.append([chr_, p, chr_ if svt < 5 else int(rng.integers(0, nchr)), p + int(rng.integers(1, 8000)), 0, 0, int(rng.integers(0, 61)),
                     int(rng.integers(0, 300))])
        rid += 5; ids.append(rid)
    a = np.array(rows, np.int32); ids = np.array(ids, np.uint64)
    order = np.lexsort((a[:, 3], a[:, 2], a[:, 1], a[:, 0]))  # SRBamRecord::operator<
    return np.ascontiguousarray(a[order]), np.ascontiguousarray(ids[order]), nchr


@pytest.mark.parametrize("svt", [0, 1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("pruning", [1000, 7])
def test_cluster_sr_matches_reference(libs, svt, pruning):
    H, R = libs
    br, ids, nchr = _sr_records(100 + svt, svt)
    n = len(br)
    outs = []
    for lib, fn in ((H, "dh_cluster_sr"), (R, "ref_cluster_sr")):
        svid = np.zeros(n, np.int32); sv = np.zeros((4096, 14), np.int32)
        cnt = getattr(lib, fn)(_p(br), _p(ids), n, svt, 2, 40, pruning, nchr, _p(svid), _p(sv), 4096)
        assert cnt >= 0
        outs.append((cnt, svid.copy(), sv[:cnt].copy()))
    assert outs[0][0] == outs[1][0] and outs[0][0] > 5
    assert np.array_equal(outs[0][1], outs[1][1])
    assert np.array_equal(outs[0][2], outs[1][2])


def _pe_records(seed, svt, n_clusters=40, noise=150):
    rng = np.random.default_rng(seed)
    rows = []
    for _ in range(n_clusters):
        tid = int(rng.integers(0, 3)); mtid = tid if svt < 5 else (tid + 1 + int(rng.integers(0, 2))) % 4
        p = int(rng.integers(5000, 300000)); size = int(rng.integers(400, 6000))
        k = int(rng.integers(1, 10)); med = 300; isz = int(rng.choice([450, 600]))
        for _ in range(k):
            a = p + int(rng.integers(-150, 150)); b = (p + size if svt < 5 else int(rng.integers(5000, 300000)) // 1000 * 0 + p + size) + int(rng.integers(-150, 150))
            pos, mpos = (b, a) if svt < 5 else (a, b)   # the reference records the later mate for intra-chromosomal pairs
            rows.append([tid, pos, mtid, mpos, 100, 100, med, 20, isz, 0, int(rng.integers(1, 61))])
    for _ in range(noise):
        tid = int(rng.integers(0, 3)); p = int(rng.integers(5000, 300000))
        rows.append([tid, p + int(rng.integers(300, 9000)), tid if svt < 5 else (tid + 1) % 4, p, 100, 100, 300, 20, 450, 0, int(rng.integers(1, 61))])
    a = np.array(rows, np.int32)
    if svt < 5:
        order = np.lexsort((a[:, 8], np.maximum(a[:, 1], a[:, 3]), np.minimum(a[:, 1], a[:, 3])))
    else:
        order = np.lexsort((a[:, 8], a[:, 3], a[:, 1]))
    return np.ascontiguousarray(a[order])


@pytest.mark.parametrize("svt", [0, 1, 2, 3, 5, 6, 7, 8])
def test_cluster_pe_matches_reference(libs, svt):
    H, R = libs
    rec = _pe_records(200 + svt, svt)
    n = len(rec)
    outs = []
    for lib, fn in ((H, "dh_cluster_pe"), (R, "ref_cluster_pe")):
        sv = np.zeros((4096, 12), np.int32)
        cnt = getattr(lib, fn)(_p(rec), n, svt, 2, 1000, C.c_uint32(600), _p(sv), 4096)
        assert cnt >= 0
        outs.append((cnt, sv[:cnt].copy()))
    assert outs[0][0] == outs[1][0]
    assert np.array_equal(outs[0][1], outs[1][1])
    if svt in (2, 3):
        assert outs[0][0] > 3


def _junction_reads(seed, nreads=3000):
    rng = np.random.default_rng(seed)
    junc, off, ids = [], [0], []
    for r in range(nreads):
        k = int(rng.choice([1, 2, 2, 3, 4]))
        base = int(rng.integers(1000, 100000)); chr_ = int(rng.integers(0, 3)); fw = int(rng.integers(0, 2))
        js = []
        for _ in range(k):
            kind = rng.random()
            refidx = chr_ if kind < 0.8 else int(rng.integers(0, 3))
            forward = fw if kind < 0.7 else 1 - fw
            refpos = base + int(rng.integers(-30, 3000)) if rng.random() < 0.7 else base + int(rng.integers(-30, 40))
            seqpos = int(rng.integers(5, 400))
            js.append([forward, int(rng.integers(0, 2)), refidx, base if rng.random() < 0.7 else -1, refpos, seqpos, int(rng.integers(0, 61))])
        js.sort(key=lambda j: (j[5], j[2], j[4], j[1]))  # Junction::operator<
        junc += js; off.append(len(junc)); ids.append(1000 + 3 * r)
    return np.array(junc, np.int32), np.array(off, np.uint32), np.array(ids, np.uint64)


def test_select_junctions_matches_reference(libs):
    H, R = libs
    for seed, mrs in ((5, 40), (6, 500)):
        junc, off, ids = _junction_reads(seed)
        outs = []
        for lib, fn in ((H, "dh_select_junctions"), (R, "ref_select_junctions")):
            out = np.zeros((200000, 9), np.int32); oid = np.zeros(200000, np.uint64); cnt = np.zeros(9, np.int32)
            tot = getattr(lib, fn)(_p(junc), _p(off), _p(ids), len(ids), mrs, 25, _p(out), _p(oid), 200000, _p(cnt))
            assert tot >= 0
            outs.append((tot, out[:tot].copy(), oid[:tot].copy(), cnt.copy()))
        assert outs[0][0] == outs[1][0]
        assert np.array_equal(outs[0][3], outs[1][3]) and (outs[0][3] > 0).sum() >= 8
        assert np.array_equal(outs[0][1], outs[1][1])
        assert np.array_equal(outs[0][2], outs[1][2])


# ---- the same clustering with the pair scan on the device (dgpu_cluster_edges_* + clusterGpu) ---------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("svt", [0, 1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("pruning", [1000, 7])
def test_cluster_sr_gpu_matches_reference(ctx, libs, svt, pruning):
    H, R = libs
    br, ids, nchr = _sr_records(300 + svt, svt, n_clusters=400, noise=3000)
    n = len(br)
    outs = []
    for fn, lead in ((H.dh_cluster_sr_gpu, (ctx.h,)), (R.ref_cluster_sr, ())):
        svid = np.zeros(n, np.int32); sv = np.zeros((8192, 14), np.int32)
        cnt = fn(*lead, _p(br), _p(ids), n, svt, 2, 40, pruning, nchr, _p(svid), _p(sv), 8192)
        assert cnt >= 0, cnt
        outs.append((cnt, svid.copy(), sv[:cnt].copy()))
    assert outs[0][0] == outs[1][0] and outs[0][0] > 50
    assert np.array_equal(outs[0][1], outs[1][1])
    assert np.array_equal(outs[0][2], outs[1][2])


@pytest.mark.gpu
@pytest.mark.parametrize("svt", [0, 1, 2, 3, 5, 6, 7, 8])
def test_cluster_pe_gpu_matches_reference(ctx, libs, svt):
    H, R = libs
    rec = _pe_records(400 + svt, svt, n_clusters=400, noise=2000)
    n = len(rec)
    outs = []
    for fn, lead in ((H.dh_cluster_pe_gpu, (ctx.h,)), (R.ref_cluster_pe, ())):
        sv = np.zeros((8192, 12), np.int32)
        cnt = fn(*lead, _p(rec), n, svt, 2, 1000, C.c_uint32(600), _p(sv), 8192)
        assert cnt >= 0, cnt
        outs.append((cnt, sv[:cnt].copy()))
    assert outs[0][0] == outs[1][0]
    assert np.array_equal(outs[0][1], outs[1][1])
    if svt in (2, 3):
        assert outs[0][0] > 30


@pytest.mark.gpu
def test_cluster_edges_capacity_and_empty(ctx):
    """Counting mode (edge_cap = 0), a too-small capacity (DGPU_ERR_CAPACITY with the needed count) and n = 0."""
    import delly_b200
    L = delly_b200.lib()
    br, ids, nchr = _sr_records(77, 2, n_clusters=50, noise=100)
    n = len(br)
    cols = [np.ascontiguousarray(br[:, k]) for k in (0, 1, 2, 3, 7)]
    off = np.zeros(n + 1, np.uint32); total = C.c_uint64()
    rc = L.dgpu_cluster_edges_sr(ctx.h, *[_p(c) for c in cols], C.c_uint64(n), 2, 40, _p(off), None, None, C.c_uint64(0), C.byref(total))
    assert rc == -4 and total.value > 0 and off[n] == total.value          # DGPU_ERR_CAPACITY, offsets still valid
    ej = np.zeros(total.value, np.uint32); ew = np.zeros(total.value, np.uint32)
    rc = L.dgpu_cluster_edges_sr(ctx.h, *[_p(c) for c in cols], C.c_uint64(n), 2, 40, _p(off), _p(ej), _p(ew), C.c_uint64(total.value), C.byref(total))
    assert rc == 0
    # every edge satisfies the predicates of src/cluster.h:373-376 and the weight of :425; targets increase per source
    for i in range(n):
        js = ej[off[i]:off[i + 1]]
        assert np.all(js > i) and np.all(np.diff(js.astype(np.int64)) > 0)
        var = min(1000, max(40, int(abs(0.1 * (int(br[i, 3]) - int(br[i, 1]))))))
        for j, w in zip(js, ew[off[i]:off[i + 1]]):
            assert br[j, 0] == br[i, 0] and br[j, 1] - br[i, 1] <= var and abs(int(br[j, 3]) - int(br[i, 3])) < var
            assert w == abs(int(br[j, 3]) - int(br[i, 3])) + abs(int(br[j, 1]) - int(br[i, 1]))
    rc = L.dgpu_cluster_edges_sr(ctx.h, None, None, None, None, None, C.c_uint64(0), 2, 40, _p(off), None, None, C.c_uint64(0), C.byref(total))
    assert rc == 0 and total.value == 0 and off[0] == 0


# ---- CIGAR -> junction scans (src/junction.h:352-437 long-read with indel look-ahead; src/shortpe.h:355-389 short-read) -----

def _cigar_records(seed, n, indel_then_match=False):
    """Records [tid, pos, flag, mapq, read id, cigar_off, n_cigar] + BAM-encoded CIGARs: clips, aligned blocks (M/=/X), short and
    long indels (also adjacent ones unless indel_then_match), N skips; several alignments per read (supplementary), both strands."""
    rng = np.random.default_rng(seed)
    rec, cig = [], []
    for i in range(n):
        rid = int(rng.integers(0, max(2, n // 2)))
        flag = int(rng.choice([0, 16, 2048, 2064, 256, 1024, 4, 512], p=[0.3, 0.3, 0.12, 0.12, 0.06, 0.04, 0.03, 0.03]))
        ops = []
        if rng.random() < 0.6:
            ops.append((int(rng.choice([4, 5])), int(rng.choice([3, 20, 26, 300]))))
        for _ in range(int(rng.integers(1, 9))):
            ops.append((int(rng.choice([0, 7, 8], p=[0.8, 0.1, 0.1])), int(rng.choice([1, 5, 30, 200, 2000]))))
            k = rng.random()
            if k < 0.75:
                nind = 1 if indel_then_match else int(rng.integers(1, 3))
                for _ in range(nind):
                    ops.append((int(rng.choice([1, 2, 3], p=[0.45, 0.45, 0.1])), int(rng.choice([1, 10, 25, 26, 60, 500]))))
        if ops[-1][0] in (1, 2, 3):
            ops.append((0, int(rng.integers(1, 100))))
        if rng.random() < 0.6:
            ops.append((int(rng.choice([4, 5])), int(rng.choice([3, 20, 26, 300]))))
        rec.append([int(rng.integers(-1, 3)) if rng.random() < 0.05 else int(rng.integers(0, 3)), int(rng.integers(0, 1_000_000)), flag,
                    int(rng.integers(0, 61)), rid, len(cig), len(ops)])
        cig += [(ln << 4) | op for op, ln in ops]
    return np.array(rec, np.int32), np.array(cig, np.uint32)


@pytest.mark.parametrize("case", [("lr", 0.5), ("lr", 0.7), ("lr", 0.05), ("sr", -1.0)])
def test_find_junctions_matches_reference(libs, case):
    """findJunctions of the reference (compiled verbatim, htslib served from memory) against cigarJunctionsLR; the short-read scan of
    scanPEandSR equals it without the look-ahead (extension threshold below zero, every long indel followed by an aligned block)."""
    H, R = libs
    mode, ext = case
    rec, cig = _cigar_records(500 + int(ext * 100), 4000, indel_then_match=(mode == "sr"))
    n = len(rec)
    R.ref_hash_lr_name.restype = C.c_uint64
    seeds = np.array([R.ref_hash_lr_name(f"r{int(r[4])}".encode()) for r in rec], np.uint64)
    outs = []
    for which in ("ref", "ours"):
        rs = np.zeros(n, np.uint64); ro = np.zeros(n + 1, np.uint32); jn = np.zeros((40 * n, 7), np.int32); nr = C.c_int()
        if which == "ref":
            cnt = R.ref_find_junctions(_p(rec), n, _p(cig), 3, 1, 25, 25, C.c_float(ext), _p(rs), _p(ro), n, _p(jn), 40 * n, C.byref(nr))
        else:
            cnt = H.dh_find_junctions(_p(rec), _p(seeds), n, _p(cig), 1, 25, 25, C.c_float(ext), 1 if mode == "lr" else 0, _p(rs), _p(ro), n, _p(jn), 40 * n,
                                      C.byref(nr))
        assert cnt > 0, cnt
        outs.append((cnt, nr.value, rs[:nr.value].copy(), ro[:nr.value + 1].copy(), jn[:cnt].copy()))
    assert outs[0][0] == outs[1][0] and outs[0][1] == outs[1][1] and outs[0][0] > n
    assert np.array_equal(outs[0][2], outs[1][2]) and np.array_equal(outs[0][3], outs[1][3])
    assert np.array_equal(outs[0][4], outs[1][4])
