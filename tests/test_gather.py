"""N>1 host logic on CPU: world_size-2 gloo run of the record all-gatherv and of the shard partitioning."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _records(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        out.append(dict(chr=int(rng.integers(0, 3)), svStart=int(rng.integers(0, 10000)), chr2=int(rng.integers(0, 3)), svEnd=int(rng.integers(0, 20000)),
                        srSupport=int(rng.integers(0, 20)), peSupport=int(rng.integers(0, 20)), svt=int(rng.integers(0, 9)), id=i,
                        srAlignQuality=float(np.float32(rng.random())), precise=bool(rng.integers(0, 2)),
                        alleles=b"A,<DEL>" if i % 2 else b"", consensus=bytes(rng.integers(65, 70, size=int(rng.integers(0, 300)), dtype=np.uint8))))
    return out


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from delly_b200 import gather
    allr = _records(57, 1)
    lo, hi = gather.shard_range(len(allr), rank, world)
    got = gather.gather_sv_records(allr[lo:hi])
    parts = gather.all_gather_bytes(b"" if rank == 0 else b"xyz" * rank)  # empty shard on one rank
    q.put((rank, got, parts, (lo, hi)))
    dist.destroy_process_group()


def test_gather_world2_gloo():
    sys.path.insert(0, ROOT)
    from delly_b200 import gather
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    allr = [dict({k: 0 for k in gather.SV_INT_FIELDS}, **r) for r in _records(57, 1)]
    exp = sorted(allr, key=lambda r: (r["chr"], r["svStart"], r["chr2"], r["svEnd"], -r["peSupport"], -r["srSupport"]))
    for i, r in enumerate(exp):
        r["id"] = i
    shards = sorted(x[3] for x in res)
    assert shards[0][0] == 0 and shards[0][1] == shards[1][0] and shards[1][1] == 57
    for rank, got, parts, _ in res:
        assert got == exp  # every rank reconstructs the same globally ordered, renumbered list
        assert parts == [b"", b"xyz"]


def test_pack_roundtrip_and_shards():
    sys.path.insert(0, ROOT)
    from delly_b200 import gather
    recs = _records(23, 9)
    assert gather.unpack_sv_records(gather.pack_sv_records(recs)) == [dict({k: 0 for k in gather.SV_INT_FIELDS}, **r) for r in recs]
    for n in (0, 1, 7, 64):
        for w in (1, 2, 3, 8):
            spans = [gather.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n and all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
