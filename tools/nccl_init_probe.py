#!/usr/bin/env python
"""Development aid: wall time of an N-rank run of the binding on the reference's tiny example (dominated by CUDA context + NCCL communicator start-up)
under different NCCL environment settings. usage: tools/nccl_init_probe.py N"""
import os, subprocess, sys, time
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ex = os.path.join(root, "oracle", "_ref", "example")
exe = os.path.join(root, "delly_b200", "bin", "delly_b200")
def run(extra):
    for f in os.listdir("/tmp"):
        if f.startswith("probe."):
            os.remove(os.path.join("/tmp", f))
    env = dict(os.environ, **extra)
    t0 = time.perf_counter()
    ps = [subprocess.Popen([exe, "sr", "--rank", str(r), "--nranks", str(N), "--comm-file", "/tmp/probe.id", "--device", str(r), "-g", os.path.join(ex, "ref.fa"),
                            "-o", f"/tmp/probe.{r}.bcf", os.path.join(ex, "sr.bam")], env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for r in range(N)]
    rcs = [p.wait() for p in ps]
    print(f"{time.perf_counter() - t0:6.2f} s  rc={rcs}  {extra}", flush=True)
sets = [{}, {}, {"NCCL_NVLS_ENABLE": "0"}, {"NCCL_NVLS_ENABLE": "0", "NCCL_MAX_NCHANNELS": "2", "NCCL_MIN_NCHANNELS": "1"},
        {"NCCL_NVLS_ENABLE": "0", "NCCL_MAX_NCHANNELS": "2", "NCCL_MIN_NCHANNELS": "1", "NCCL_IB_DISABLE": "1", "NCCL_NET_DISABLE": "1"},
        {"NCCL_NVLS_ENABLE": "0", "NCCL_MAX_NCHANNELS": "2", "NCCL_MIN_NCHANNELS": "1", "NCCL_IB_DISABLE": "1", "NCCL_CUMEM_ENABLE": "0"},
        {"NCCL_MAX_NCHANNELS": "2", "NCCL_MIN_NCHANNELS": "1"}]
for s in sets:
    run(s)
t0 = time.perf_counter()
subprocess.run([exe, "sr", "-g", os.path.join(ex, "ref.fa"), "-o", "/tmp/probe.one.bcf", os.path.join(ex, "sr.bam")], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
print(f"{time.perf_counter() - t0:6.2f} s  single rank")
