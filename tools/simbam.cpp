// simbam — synthetic short-read sample generator for the pipeline-level benchmarks and whole-file parity tests ("reference genome
// = uniform random ACGT ...; reads = reference substrings with substitution error; SVs planted at uniform positions"). Writes a FASTA (+ .fai),
// a coordinate-sorted BAM (+ .bai) and a truth table, through htslib — the inputs `delly sr` / `delly_b200 sr` read.
//
// A diploid sample: deletions, tandem duplications and inversions planted on every contig (heterozygous on haplotype 1 only, or homozygous), paired
// reads (FR, insert ~ N(isize, isd)) sampled uniformly from both haplotypes and "aligned by construction": a read over an SV junction becomes a
// soft-clipped primary alignment of its longer part plus — when the shorter part is >= 20 bp — a supplementary, hard-clipped alignment of the rest
// (what bwa mem reports); pairs over a junction get the insert size / orientation the reference genome implies (the discordant-pair signal).
// Deterministic for a given seed. Not part of the product path; bench.py and tests/ run it.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include <htslib/faidx.h>
#include <htslib/sam.h>

namespace {

struct Sv { int type; int64_t s, e; int zyg; };   // type 2 = DEL, 3 = DUP (tandem), 0 = INV (both junctions); [s, e) on the reference
struct Seg { int64_t dstart, dend, rstart; bool rev; };   // donor [dstart, dend) = reference rstart.. (forward) or the reverse complement of reference [rstart, rstart+len)
struct Aln { int64_t pos; bool rev; std::vector<uint32_t> cigar; bool supp; };
struct Rec { int32_t tid; int64_t pos; uint16_t flag; uint8_t mapq; std::vector<uint32_t> cigar; int32_t mtid; int64_t mpos; int64_t isize; uint64_t pair; uint32_t seqIdx; bool seqRev; };

char comp(char c) { switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return 'N'; } }

struct Hap {
  std::string seq;
  std::vector<Seg> segs;
};

Hap buildHap(std::string const& ref, std::vector<Sv> const& svs, int h) {
  Hap hp;
  int64_t prev = 0;
  auto addFwd = [&](int64_t a, int64_t b) { if (b <= a) return; hp.segs.push_back(Seg{(int64_t) hp.seq.size(), (int64_t) hp.seq.size() + (b - a), a, false}); hp.seq.append(ref, (size_t) a, (size_t) (b - a)); };
  auto addRev = [&](int64_t a, int64_t b) {
    hp.segs.push_back(Seg{(int64_t) hp.seq.size(), (int64_t) hp.seq.size() + (b - a), a, true});
    for (int64_t i = b - 1; i >= a; --i) hp.seq.push_back(comp(ref[(size_t) i]));
  };
  for (Sv const& v : svs) {
    const bool carries = (v.zyg == 2) || (h == 1);
    if (!carries) continue;
    addFwd(prev, v.s);
    if (v.type == 2) prev = v.e;                                  // deletion: skip [s, e)
    else if (v.type == 3) { addFwd(v.s, v.e); prev = v.s; }         // tandem duplication: [s, e) twice
    else { addRev(v.s, v.e); prev = v.e; }                          // inversion
  }
  addFwd(prev, (int64_t) ref.size());
  return hp;
}

// alignments of donor [a, a + L): the read is cut at segment borders; the longest piece is the primary (others soft-clipped), the second longest
// (>= 20 bp) a supplementary with hard clips. Pieces on reversed segments align on the reverse strand.
std::vector<Aln> place(Hap const& hp, int64_t a, int L) {
  std::vector<Aln> out;
  const int64_t b = a + L;
  auto it = std::upper_bound(hp.segs.begin(), hp.segs.end(), a, [](int64_t x, Seg const& s) { return x < s.dend; });
  struct Piece { int off, len; int64_t rpos; bool rev; };
  std::vector<Piece> pieces;
  for (; it != hp.segs.end() && it->dstart < b; ++it) {
    const int64_t lo = std::max(a, it->dstart), hi = std::min(b, it->dend);
    if (hi <= lo) continue;
    Piece p; p.off = (int) (lo - a); p.len = (int) (hi - lo); p.rev = it->rev;
    p.rpos = it->rev ? it->rstart + (it->dend - hi) : it->rstart + (lo - it->dstart);
    pieces.push_back(p);
  }
  // merge pieces that are contiguous on the reference in the same orientation (a haplotype without the SV has several segments that join up)
  std::vector<Piece> m;
  for (Piece const& p : pieces) {
    if (!m.empty() && !m.back().rev && !p.rev && m.back().rpos + m.back().len == p.rpos) m.back().len += p.len;
    else m.push_back(p);
  }
  std::vector<int> order(m.size());
  for (size_t i = 0; i < m.size(); ++i) order[i] = (int) i;
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return m[x].len > m[y].len; });
  for (size_t k = 0; k < order.size() && k < 2; ++k) {
    Piece const& p = m[order[k]];
    if (k == 1 && p.len < 20) break;
    Aln al; al.pos = p.rpos; al.rev = p.rev; al.supp = (k == 1);
    int left = p.off, right = L - p.off - p.len;
    if (p.rev) std::swap(left, right);   // the record shows the read reverse-complemented
    const uint32_t clip = al.supp ? BAM_CHARD_CLIP : BAM_CSOFT_CLIP;
    if (left) al.cigar.push_back(((uint32_t) left << 4) | clip);
    al.cigar.push_back(((uint32_t) p.len << 4) | BAM_CMATCH);
    if (right) al.cigar.push_back(((uint32_t) right << 4) | clip);
    out.push_back(al);
  }
  return out;
}

const char* arg(int argc, char** argv, const char* name, const char* def) {
  for (int i = 1; i + 1 < argc; ++i) if (!std::strcmp(argv[i], name)) return argv[i + 1];
  return def;
}

}  // namespace

int main(int argc, char** argv) {
  const std::string out = arg(argc, argv, "--out", "sim");
  const int64_t glen = std::atoll(arg(argc, argv, "--genome-len", "1000000"));
  const int ncontig = std::atoi(arg(argc, argv, "--contigs", "2"));
  const double cov = std::atof(arg(argc, argv, "--cov", "30"));
  const int RL = std::atoi(arg(argc, argv, "--read-len", "150"));
  const double imean = std::atof(arg(argc, argv, "--isize", "350")), isd = std::atof(arg(argc, argv, "--isd", "20"));
  const int nsv = std::atoi(arg(argc, argv, "--n-sv", "100"));
  const uint64_t seed = (uint64_t) std::atoll(arg(argc, argv, "--seed", "1"));
  const int threads = std::atoi(arg(argc, argv, "--threads", "4"));
  const double err = std::atof(arg(argc, argv, "--err", "0.003"));
  const std::string types = arg(argc, argv, "--types", "DEL");   // comma list of DEL,DUP,INV
  const std::string sample = arg(argc, argv, "--sample", "sim");
  std::mt19937_64 rng(seed);
  auto uni = [&](uint64_t n) { return (uint64_t) (rng() % n); };
  std::vector<int> typeCodes;
  if (types.find("DEL") != std::string::npos) typeCodes.push_back(2);
  if (types.find("DUP") != std::string::npos) typeCodes.push_back(3);
  if (types.find("INV") != std::string::npos) typeCodes.push_back(0);
  if (typeCodes.empty()) typeCodes.push_back(2);
  // genome
  std::vector<std::string> contigs((size_t) ncontig);
  std::vector<int64_t> clen((size_t) ncontig);
  for (int c = 0; c < ncontig; ++c) { clen[c] = glen / ncontig; contigs[c].resize((size_t) clen[c]); for (auto& ch : contigs[c]) ch = "ACGT"[rng() & 3]; }
  {
    FILE* f = std::fopen((out + ".fa").c_str(), "w");
    if (!f) { std::perror("fa"); return 1; }
    for (int c = 0; c < ncontig; ++c) {
      std::fprintf(f, ">chr%d\n", c + 1);
      for (int64_t i = 0; i < clen[c]; i += 60) { std::fwrite(contigs[c].data() + i, 1, (size_t) std::min<int64_t>(60, clen[c] - i), f); std::fputc('\n', f); }
    }
    std::fclose(f);
    if (fai_build((out + ".fa").c_str()) != 0) { std::fprintf(stderr, "fai_build failed\n"); return 1; }
  }
  // SVs: evenly spaced slots, jittered start, size 300..2000
  std::vector<std::vector<Sv> > svs((size_t) ncontig);
  FILE* truth = std::fopen((out + ".truth.tsv").c_str(), "w");
  {
    const int per = std::max(1, nsv / ncontig);
    for (int c = 0; c < ncontig; ++c) {
      const int64_t slot = (clen[c] - 8000) / per;
      if (slot < 4000) { std::fprintf(stderr, "too many SVs for this genome length (slot %lld bp)\n", (long long) slot); return 1; }
      for (int k = 0; k < per; ++k) {
        Sv v; v.type = typeCodes[uni(typeCodes.size())];
        const int64_t size = 300 + (int64_t) uni(1700);
        v.s = 4000 + k * slot + (int64_t) uni((uint64_t) std::max<int64_t>(1, slot - size - 1500));
        v.e = v.s + size; v.zyg = (uni(10) < 7) ? 1 : 2;
        svs[c].push_back(v);
        if (truth) std::fprintf(truth, "chr%d\t%lld\t%lld\t%s\t%s\n", c + 1, (long long) v.s, (long long) v.e, v.type == 2 ? "DEL" : v.type == 3 ? "DUP" : "INV", v.zyg == 2 ? "hom" : "het");
      }
    }
  }
  if (truth) std::fclose(truth);
  // reads
  std::vector<Rec> recs;
  std::vector<std::string> seqs;   // read sequences in donor orientation
  std::normal_distribution<double> insd(imean, isd);
  uint64_t pairId = 0;
  for (int c = 0; c < ncontig; ++c) {
    Hap haps[2] = {buildHap(contigs[c], svs[c], 0), buildHap(contigs[c], svs[c], 1)};
    const uint64_t npairs = (uint64_t) ((double) clen[c] * cov / (2.0 * RL));
    for (uint64_t p = 0; p < npairs; ++p, ++pairId) {
      Hap const& hp = haps[rng() & 1];
      int ins = (int) insd(rng);
      ins = std::max(RL + 10, std::min(ins, (int) (imean + 6 * isd)));
      if ((int64_t) hp.seq.size() < ins + 10) continue;
      const int64_t a1 = (int64_t) uni((uint64_t) ((int64_t) hp.seq.size() - ins)), a2 = a1 + ins - RL;
      std::vector<Aln> al[2] = {place(hp, a1, RL), place(hp, a2, RL)};
      if (al[0].empty() || al[1].empty()) continue;
      // donor strand: read 1 forward, read 2 reverse; a piece on an inverted segment flips
      const bool rv[2] = {al[0][0].rev, !al[1][0].rev};
      for (int k = 0; k < 2; ++k) {
        std::string s = hp.seq.substr((size_t) (k == 0 ? a1 : a2), (size_t) RL);
        for (auto& ch : s) if ((double) (rng() >> 11) * (1.0 / 9007199254740992.0) < err) ch = "ACGT"[rng() & 3];
        seqs.push_back(s);
        const uint32_t sidx = (uint32_t) seqs.size() - 1;
        for (Aln const& x : al[k]) {
          Rec r; r.tid = c; r.pos = x.pos; r.mapq = 60; r.cigar = x.cigar; r.pair = pairId; r.seqIdx = sidx;
          const bool myRev = (k == 0) ? x.rev : !x.rev;
          r.seqRev = x.rev;   // stored sequence = donor piece, reverse-complemented when the segment is inverted
          r.flag = (uint16_t) (BAM_FPAIRED | (k == 0 ? BAM_FREAD1 : BAM_FREAD2) | (myRev ? BAM_FREVERSE : 0) | (rv[1 - k] ? BAM_FMREVERSE : 0) | (x.supp ? BAM_FSUPPLEMENTARY : 0));
          r.mtid = c; r.mpos = al[1 - k][0].pos;
          const int64_t p0 = al[0][0].pos, p1 = al[1][0].pos;
          const int64_t span = (p1 >= p0) ? (p1 + RL - p0) : -(p0 + RL - p1);
          r.isize = (k == 0) ? span : -span;
          recs.push_back(r);
        }
      }
    }
  }
  std::vector<uint32_t> order(recs.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = (uint32_t) i;
  std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return recs[x].tid != recs[y].tid ? recs[x].tid < recs[y].tid : recs[x].pos < recs[y].pos; });
  // BAM
  samFile* fp = sam_open((out + ".bam").c_str(), "wb1");
  if (!fp) { std::fprintf(stderr, "cannot write %s.bam\n", out.c_str()); return 1; }
  if (threads > 1) hts_set_threads(fp, threads);
  sam_hdr_t* hdr = sam_hdr_init();
  sam_hdr_add_line(hdr, "HD", "VN", "1.6", "SO", "coordinate", NULL);
  for (int c = 0; c < ncontig; ++c) { const std::string nm = "chr" + std::to_string(c + 1), ln = std::to_string(clen[c]); sam_hdr_add_line(hdr, "SQ", "SN", nm.c_str(), "LN", ln.c_str(), NULL); }
  sam_hdr_add_line(hdr, "RG", "ID", "rg1", "SM", sample.c_str(), NULL);
  if (sam_hdr_write(fp, hdr) != 0) { std::fprintf(stderr, "header write failed\n"); return 1; }
  bam1_t* b = bam_init1();
  std::string qual((size_t) RL, (char) 30), seqbuf;
  for (uint32_t oi : order) {
    Rec const& r = recs[oi];
    char qname[40];
    std::snprintf(qname, sizeof(qname), "p%llu", (unsigned long long) r.pair);
    std::string const& s = seqs[r.seqIdx];
    int hardL = 0, hardR = 0;
    if ((r.cigar.front() & 0xf) == BAM_CHARD_CLIP) hardL = (int) (r.cigar.front() >> 4);
    if (r.cigar.size() > 1 && (r.cigar.back() & 0xf) == BAM_CHARD_CLIP) hardR = (int) (r.cigar.back() >> 4);
    if (r.seqRev) { seqbuf.clear(); for (int i = RL - 1; i >= 0; --i) seqbuf.push_back(comp(s[(size_t) i])); }
    else seqbuf = s;
    const std::string shown = seqbuf.substr((size_t) hardL, (size_t) (RL - hardL - hardR));
    if (bam_set1(b, std::strlen(qname), qname, r.flag, r.tid, r.pos, r.mapq, r.cigar.size(), r.cigar.data(), r.mtid, r.mpos, r.isize, shown.size(), shown.data(),
                 qual.data(), 0) < 0 || sam_write1(fp, hdr, b) < 0) { std::fprintf(stderr, "record write failed\n"); return 1; }
  }
  bam_destroy1(b);
  sam_hdr_destroy(hdr);
  sam_close(fp);
  if (sam_index_build((out + ".bam").c_str(), 0) != 0) { std::fprintf(stderr, "index build failed\n"); return 1; }
  std::fprintf(stderr, "simbam: %zu records, %d contigs x %lld bp, %d SVs -> %s.bam\n", recs.size(), ncontig, (long long) (glen / ncontig), nsv, out.c_str());
  return 0;
}
