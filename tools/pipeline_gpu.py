#!/usr/bin/env python
"""Run the pipeline-level comparison on the GPU box: simbam -> delly_ref sr (all host threads) vs delly_b200 sr (one B200), discovery and
genotyping mode, BCFs compared byte for byte. usage: tools/pipeline_gpu.py GENOME_LEN N_SV [CONTIGS] ; prints one JSON line."""
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from delly_b200 import pipeline_bench as pb

glen, nsv = int(sys.argv[1]), int(sys.argv[2])
contigs = int(sys.argv[3]) if len(sys.argv) > 3 else 4
cores = len(os.sched_getaffinity(0))
d = tempfile.mkdtemp(prefix="dpipe")
pre = os.path.join(d, "s")
out = {"genome_len": glen, "n_sv": nsv, "contigs": contigs, "host_threads": cores}
out["simulate"] = pb.simulate(pre, glen, contigs, nsv, seed=11, threads=min(cores, 16))
ref_bcf, our_bcf = pre + ".ref.bcf", pre + ".ours.bcf"
pb.run_ours(pre, our_bcf, timing=pre + ".warm.json")           # warm-up (CUDA context, page cache)
t_ours, stages = pb.run_ours(pre, our_bcf, timing=pre + ".t.json")
t_ref = pb.run_reference(pre, ref_bcf, threads=cores)
out["discovery"] = {"reference_s": t_ref, "ours_s": t_ours, "speedup": t_ref / t_ours, "bcf_identical": pb.inflate(ref_bcf) == pb.inflate(our_bcf),
                    "records": pb.count_records(ref_bcf), "stages": stages}
g_ref, g_our = pre + ".gref.bcf", pre + ".gours.bcf"
t_ours, stages = pb.run_ours(pre, g_our, sites=ref_bcf, timing=pre + ".tg.json")
t_ref = pb.run_reference(pre, g_ref, threads=cores, sites=ref_bcf)
out["genotyping_mode"] = {"reference_s": t_ref, "ours_s": t_ours, "speedup": t_ref / t_ours, "bcf_identical": pb.inflate(g_ref) == pb.inflate(g_our),
                          "records": pb.count_records(g_ref), "stages": stages}
print(json.dumps(out))
