#!/usr/bin/env python
"""Development aid: the binding on a simulated sample with DGPU_TRACE=1 (one line per device batch call) and the stage clock, discovery and
genotyping mode. usage: tools/pipeline_trace.py GENOME_LEN N_SV"""
import json, os, subprocess, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from delly_b200 import pipeline_bench as pb
glen, nsv = int(sys.argv[1]), int(sys.argv[2])
d = tempfile.mkdtemp(prefix="dtrace"); pre = os.path.join(d, "s")
pb.simulate(pre, glen, 4, nsv, seed=11, threads=16)
env = dict(os.environ, DGPU_TRACE="1")
pb.run_ours(pre, pre + ".w.bcf")
for name, sites in (("discovery", None), ("genotyping", pre + ".w.bcf")):
    r = subprocess.run(pb.ours_cmd(pre, pre + ".o.bcf", sites=sites, timing=pre + ".t.json"), capture_output=True, text=True, env=env)
    print("==", name, r.returncode)
    print("\n".join(l for l in r.stderr.splitlines() if l.startswith("[dgpu]") or l.startswith("[msa")))
    print(json.load(open(pre + ".t.json")))
