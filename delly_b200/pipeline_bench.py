"""Pipeline-level measurement: the whole `delly sr` chain — BAM + FASTA in, BCF out — through the product binding
(delly_b200/bin/delly_b200: htslib IO + batched stage mirrors + device kernels) next to the reference's own drivers (oracle/_ref/delly_ref:
src/delly.h compiled verbatim against the same htslib, run with all host threads), on a synthetic sample written by delly_b200/bin/simbam
(uniform random genome, 150 bp FR pairs at 30x, planted het / hom deletions, tandem duplications, inversions). The two BCFs must be
identical byte for byte (BGZF inflated, ##fileDate aside) — "bit-exact BCF vs CPU ref" of BASELINE.json's metric, at a size beyond the example
files. Used by bench.py (never by the product path)."""
import gzip
import json
import os
import re
import subprocess
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "delly_b200", "bin")
DELLY_REF = os.path.join(ROOT, "oracle", "_ref", "delly_ref")


def simulate(prefix, genome_len, contigs, n_sv, cov=30, seed=1, types="DEL,DUP,INV", threads=8):
    t0 = time.perf_counter()
    r = subprocess.run([os.path.join(BIN, "simbam"), "--out", prefix, "--genome-len", str(genome_len), "--contigs", str(contigs), "--n-sv", str(n_sv), "--cov", str(cov),
                        "--seed", str(seed), "--types", types, "--threads", str(threads)], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("simbam failed: " + r.stderr[-500:])
    m = re.search(r"simbam: (\d+) records", r.stderr)
    return {"records": int(m.group(1)) if m else None, "seconds": time.perf_counter() - t0, "bam_bytes": os.path.getsize(prefix + ".bam")}


def inflate(path):
    return re.sub(rb"##fileDate=\d+\n", b"##fileDate=X\n", gzip.open(path, "rb").read(), count=1)


def count_records(path):
    """number of BCF records = lines of the VCF text view; cheap proxy: occurrences of the SVMETHOD value in the inflated stream"""
    return inflate(path).count(b"EMBL.DELLYv")


def run_reference(prefix, out, threads, sites=None):
    cmd = [DELLY_REF, "sr", "-g", prefix + ".fa", "-o", out, "-h", str(threads)] + (["-v", sites] if sites else []) + [prefix + ".bam"]
    t0 = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, text=True)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError("delly_ref failed: " + r.stderr[-500:])
    return dt


def ours_cmd(prefix, out, device=0, sites=None, rank=0, nranks=1, comm_file=None, timing=None, threads=4):
    cmd = [os.path.join(BIN, "delly_b200"), "sr", "-g", prefix + ".fa", "-o", out, "-h", str(threads), "--device", str(device)]
    if sites:
        cmd += ["-v", sites]
    if nranks > 1:
        cmd += ["--rank", str(rank), "--nranks", str(nranks), "--comm-file", comm_file]
    if timing:
        cmd += ["--timing", timing]
    return cmd + [prefix + ".bam"]


def run_ours(prefix, out, device=0, sites=None, timing=None, threads=4):
    t0 = time.perf_counter()
    r = subprocess.run(ours_cmd(prefix, out, device=device, sites=sites, timing=timing, threads=threads), capture_output=True, text=True)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError("delly_b200 failed: " + r.stderr[-800:])
    stages = json.load(open(timing)) if timing and os.path.exists(timing) else None
    return dt, stages
