"""The one exchange step of the sharded path: an all-gatherv of the per-rank finished call
records before BCF emission. SVs are sharded by contiguous id range / chromosome across ranks (one process per
GPU); each rank realigns and genotypes its shard with no data-path collective; then every rank (or rank 0)
needs all records to restore the reference's global order (sort(svs), renumber ids — src/delly.h:156-158).

Implementation: all_gather of the byte counts, then all_gather of the payload padded to the longest shard —
two NCCL collectives over NVLink/NVSwitch on the GPU box (backend "nccl", tensors on the device), or gloo on
CPU in the tests. Payloads are tiny (100-500 B per SV), so this is latency- not bandwidth-bound.

Record wire format (little endian), mirroring StructuralVariantRecord (src/tags.h:93-118):
  int32 x 20: chr svStart chr2 svEnd ciposlow ciposhigh ciendlow ciendhigh srSupport srMapQuality mapq insLen svt id
              homLen peSupport peMapQuality consBp alleleid nallele
  float32 srAlignQuality, uint8 precise, uint32 len(alleles), uint32 len(consensus), bytes alleles, bytes consensus
"""
import struct

import numpy as np
import torch
import torch.distributed as dist

SV_INT_FIELDS = ("chr", "svStart", "chr2", "svEnd", "ciposlow", "ciposhigh", "ciendlow", "ciendhigh", "srSupport", "srMapQuality",
                 "mapq", "insLen", "svt", "id", "homLen", "peSupport", "peMapQuality", "consBp", "alleleid", "nallele")
_HDR = struct.Struct("<20ifBII")


def pack_sv_records(records):
    """records: iterable of dicts with SV_INT_FIELDS + srAlignQuality, precise, alleles (bytes), consensus (bytes)."""
    out = bytearray()
    for r in records:
        al, co = bytes(r.get("alleles", b"")), bytes(r.get("consensus", b""))
        out += _HDR.pack(*[int(r.get(k, 0)) for k in SV_INT_FIELDS], float(r.get("srAlignQuality", 0.0)), 1 if r.get("precise") else 0,
                         len(al), len(co))
        out += al + co
    return bytes(out)


def unpack_sv_records(buf):
    recs, p = [], 0
    while p < len(buf):
        vals = _HDR.unpack_from(buf, p)
        p += _HDR.size
        r = dict(zip(SV_INT_FIELDS, vals[:20]))
        r["srAlignQuality"] = vals[20]; r["precise"] = bool(vals[21])
        la, lc = vals[22], vals[23]
        r["alleles"] = bytes(buf[p:p + la]); p += la
        r["consensus"] = bytes(buf[p:p + lc]); p += lc
        recs.append(r)
    return recs


def all_gather_bytes(local, device=None):
    """All-gatherv of one byte string per rank. Returns the list of every rank's bytes (rank order)."""
    world = dist.get_world_size()
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu"))
    n = torch.tensor([len(local)], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    buf = torch.zeros(cap, dtype=torch.uint8, device=dev)
    if len(local):
        buf[:len(local)] = torch.frombuffer(bytearray(local), dtype=torch.uint8).to(dev)
    parts = [torch.empty(cap, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(parts, buf)
    return [bytes(p[:s].cpu().numpy().tobytes()) for p, s in zip(parts, sizes)]


def shard_range(n_items, rank, world):
    """Contiguous, balanced [lo, hi) shard of n_items for `rank` (SV ids / cluster ids / reads)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_sv_records(local_records, id_offset_fix=True):
    """Gather every rank's SV records, restore the reference's global order and renumber ids
    (concatenate in rank order -> sort with StructuralVariantRecord::operator< -> id = index; src/delly.h:156-158)."""
    parts = all_gather_bytes(pack_sv_records(local_records))
    allrecs = []
    for p in parts:
        allrecs += unpack_sv_records(p)
    allrecs.sort(key=lambda r: (r["chr"], r["svStart"], r["chr2"], r["svEnd"], -r["peSupport"], -r["srSupport"]))
    if id_offset_fix:
        for i, r in enumerate(allrecs):
            r["id"] = i
    return allrecs
