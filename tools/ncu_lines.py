#!/usr/bin/env python
"""Attribute ncu warp-stall samples to source lines: joins the SASS page of an .ncu-rep with nvdisasm line info
of the matching cubin (development aid).  usage: ncu_lines.py report.ncu-rep cubin-name source.cu [top]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, cubin, src = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "delly_b200", "libdelly_b200.so")], cwd=tmp, capture_output=True)
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin + ".sm_100a.cubin")], capture_output=True, text=True).stdout
kernel = None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kname = rows[0][1]
hdr, data = rows[1], rows[2:]
# walk the disassembly of the profiled function only
short = os.environ.get("NCU_FN") or re.search(r"(\w+)[<(]", kname).group(1)  # NCU_FN: mangled-name substring for template instances
cur, seq, infn = None, [], False
for l in sass.split("\n"):
    if l.startswith("\t.section") or l.startswith(".section"):
        infn = (".text." in l) and (short in l)
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if infn and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        seq.append(cur)
si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
print(f"{kname}: {len(data)} SASS rows, {len(seq)} disassembled")
agg, ins = collections.Counter(), collections.Counter()
for k, r in enumerate(data):
    if k < len(seq) and seq[k]:
        agg[seq[k]] += int(r[si] or 0)
        ins[seq[k]] += int(r[ii] or 0)
tot = sum(agg.values())
text = open(src).read().split("\n")
base = os.path.basename(src)
for (f, ln), v in agg.most_common(top):
    t = text[ln - 1].strip()[:100] if f == base else ""
    print(f"{v / tot * 100:5.1f}% {ins[(f, ln)]:>11} {f}:{ln} {t}")
