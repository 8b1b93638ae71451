#!/usr/bin/env python
"""Turn the reference's own example data (example/sr.bam, example/lr.bam, example/ref.fa — its only fixtures) into the flat
record layout the test hooks take, as compressed fixtures next to this script. BAM is parsed here directly (BGZF = concatenated gzip members;
record layout: SAM/BAM specification section 4.2), so no htslib build is needed.

    python tests/golden/make_example_fixture.py /root/reference/example

Layout of <name>.npz: contig (uint8 arena), coff / clen, names of the contigs; rec (n x 12 int32: tid, pos, flag, mapq, l_qseq, cigar_off, n_cigar,
seq_off, mtid, mpos, isize, name id — mates share the id), cig (uint32, BAM-encoded), reads (uint8 ASCII arena), hp / ps (int32, 0 / -1 = no tag),
tagflags + mm / mm_off + ml / ml_off (MM:Z and ML:B:C tags), in file order (coordinate-sorted)."""
import gzip
import os
import struct
import sys

import numpy as np

SEQ = b"=ACMGRSVTWYHKDBN"


def read_fasta(path):
    names, seqs, cur = [], [], []
    for line in open(path, "rb"):
        line = line.strip()
        if line.startswith(b">"):
            if names:
                seqs.append(b"".join(cur))
            names.append(line[1:].split()[0].decode()); cur = []
        else:
            cur.append(line)
    seqs.append(b"".join(cur))
    return names, seqs


def parse_aux(buf):
    """-> dict tag -> (type, value)"""
    out, p = {}, 0
    sizes = {"A": 1, "c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4}
    fmts = {"c": "<b", "C": "<B", "s": "<h", "S": "<H", "i": "<i", "I": "<I", "f": "<f"}
    while p + 3 <= len(buf):
        tag = buf[p:p + 2].decode(); ty = chr(buf[p + 2]); p += 3
        if ty == "Z" or ty == "H":
            e = buf.index(b"\0", p); out[tag] = (ty, buf[p:e]); p = e + 1
        elif ty == "B":
            sub = chr(buf[p]); n = struct.unpack_from("<I", buf, p + 1)[0]; p += 5
            out[tag] = ("B" + sub, buf[p:p + n * sizes[sub]]); p += n * sizes[sub]
        elif ty == "A":
            out[tag] = (ty, buf[p:p + 1]); p += 1
        else:
            out[tag] = (ty, struct.unpack_from(fmts[ty], buf, p)[0]); p += sizes[ty]
    return out


def read_bam(path):
    data = gzip.open(path, "rb").read()
    assert data[:4] == b"BAM\1"
    l_text = struct.unpack_from("<i", data, 4)[0]
    p = 8 + l_text
    n_ref = struct.unpack_from("<i", data, p)[0]; p += 4
    refs = []
    for _ in range(n_ref):
        l_name = struct.unpack_from("<i", data, p)[0]; p += 4
        name = data[p:p + l_name - 1].decode(); p += l_name
        l_ref = struct.unpack_from("<i", data, p)[0]; p += 4
        refs.append((name, l_ref))
    recs = []
    while p < len(data):
        bs = struct.unpack_from("<i", data, p)[0]; p += 4
        tid, pos, l_read_name, mapq, _bin, n_cigar, flag, l_seq, mtid, mpos, tlen = struct.unpack_from("<iiBBHHHiiii", data, p)
        q = p + 32
        name = data[q:q + l_read_name - 1]; q += l_read_name
        cigar = np.frombuffer(data, "<u4", n_cigar, q).copy(); q += 4 * n_cigar
        packed = np.frombuffer(data, np.uint8, (l_seq + 1) // 2, q); q += (l_seq + 1) // 2
        codes = np.empty(2 * len(packed), np.uint8); codes[0::2] = packed >> 4; codes[1::2] = packed & 15
        seq = np.frombuffer(SEQ, np.uint8)[codes[:l_seq]]
        q += l_seq   # qualities
        aux = parse_aux(data[q:p + bs])
        recs.append((tid, pos, flag, mapq, l_seq, n_cigar, mtid, mpos, tlen, name, cigar, seq, aux))
        p += bs
    return refs, recs


def main(src):
    here = os.path.dirname(os.path.abspath(__file__))
    fa_names, fa_seqs = read_fasta(os.path.join(src, "ref.fa"))
    for which in ("sr", "lr"):
        refs, recs = read_bam(os.path.join(src, which + ".bam"))
        assert [r[0] for r in refs] == fa_names and [r[1] for r in refs] == [len(s) for s in fa_seqs]
        ids, rows, cigs, reads, hp, ps, flags, mms, mls = {}, [], [], [], [], [], [], [], []
        coff = 0
        roff = 0
        for (tid, pos, flag, mapq, l_seq, n_cigar, mtid, mpos, tlen, name, cigar, seq, aux) in recs:
            nid = ids.setdefault(name, len(ids))
            rows.append([tid, pos, flag, mapq, l_seq, coff, n_cigar, roff, mtid, mpos, tlen, nid])
            cigs.append(cigar); coff += n_cigar
            reads.append(seq); roff += l_seq
            hp.append(int(aux["HP"][1]) if "HP" in aux and aux["HP"][0] in "cCsSiI" else 0)
            ps.append(int(aux["PS"][1]) if "PS" in aux and aux["PS"][0] in "cCsSiI" else -1)
            f = 0
            mm, ml = b"", b""
            if "MM" in aux and aux["MM"][0] == "Z": f |= 1; mm = aux["MM"][1]
            if "ML" in aux and aux["ML"][0] == "BC": f |= 2; ml = aux["ML"][1]
            flags.append(f); mms.append(mm); mls.append(ml)
        mm_off = np.zeros(len(rows) + 1, np.uint32); mm_off[1:] = np.cumsum([len(x) for x in mms])
        ml_off = np.zeros(len(rows) + 1, np.uint32); ml_off[1:] = np.cumsum([len(x) for x in mls])
        out = os.path.join(here, "example_%s.npz" % which)
        np.savez_compressed(out, contig=np.frombuffer(b"".join(fa_seqs), np.uint8), coff=np.cumsum([0] + [len(s) for s in fa_seqs[:-1]]).astype(np.uint32),
                            clen=np.array([len(s) for s in fa_seqs], np.uint32), names=np.array(fa_names), rec=np.array(rows, np.int32),
                            cig=np.concatenate(cigs).astype(np.uint32) if cigs else np.zeros(0, np.uint32), reads=np.concatenate(reads).astype(np.uint8),
                            hp=np.array(hp, np.int32), ps=np.array(ps, np.int32), tagflags=np.array(flags, np.uint8),
                            mm=np.frombuffer(b"".join(mms) + b"\\0", np.uint8), mm_off=mm_off, ml=np.frombuffer(b"".join(mls) + b"\\0", np.uint8), ml_off=ml_off)
        print(which, len(rows), "records,", len(ids), "names,", os.path.getsize(out) // 1024, "KiB,", "tags:", sum(1 for f in flags if f), "HP:", sum(1 for h in hp if h))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/example")
