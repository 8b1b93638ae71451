// oracle/_ref wrapper, part 2 (TEST INFRASTRUCTURE ONLY): the reference's clustering and junction selection
// compiled VERBATIM from /root/reference/src/cluster.h and junction.h. util.h (real Boost filesystem / iostreams /
// icl) and pangenome.h are masked by pre-defining their include guards; the util.h symbols these headers name are
// restated below with their reference lines. Nothing from the reference is copied into this repository.
#define UTIL_H
#define PANGENOME_H
#include <cmath>
#include <cstring>
#include <limits>
#include <numeric>
#include <map>
#include <set>
#include <unordered_map>
#include <unordered_set>
#include "shim/prelude.h"

namespace boost {
template <typename K, typename V> using unordered_map = std::unordered_map<K, V>;
namespace posix_time {
struct ptime {};
struct second_clock { static ptime local_time() { return ptime(); } };
inline std::string to_simple_string(ptime const&) { return "now"; }
}  // namespace posix_time
namespace filesystem {
struct path {
  std::string s;
  path() {}
  path(std::string const& x) : s(x) {}
  std::string const& string() const { return s; }
};
inline std::ostream& operator<<(std::ostream& o, path const& p) { return o << p.s; }
}  // namespace filesystem
}  // namespace boost

#include <htslib/faidx.h>
#include "tags.h"

namespace torali {
// util.h:430-438 — query length from the CIGAR
inline uint32_t readLength(bam1_t const* rec) {
  uint32_t const* cigar = bam_get_cigar(rec);
  uint32_t slen = 0;
  for (uint32_t i = 0; i < rec->core.n_cigar; ++i)
    if ((bam_cigar_op(cigar[i]) == BAM_CMATCH) || (bam_cigar_op(cigar[i]) == BAM_CEQUAL) || (bam_cigar_op(cigar[i]) == BAM_CDIFF) ||
        (bam_cigar_op(cigar[i]) == BAM_CINS) || (bam_cigar_op(cigar[i]) == BAM_CSOFT_CLIP) || (bam_cigar_op(cigar[i]) == BAM_CHARD_CLIP))
      slen += bam_cigar_oplen(cigar[i]);
  return slen;
}
// opaque read ids (util.h:519-535 use boost::hash; any injective-enough id works for the functions tested here)
inline std::size_t hash_lr(bam1_t* rec) { return std::hash<std::string>()(bam_get_qname(rec)); }
inline std::size_t hash_sr(bam1_t* rec) { return std::hash<std::string>()(bam_get_qname(rec)) * 2 + ((rec->core.flag & BAM_FREAD2) ? 1 : 0); }
inline bool isBamCram(std::string const&) { return true; }
inline std::string _addID(int32_t const) { return "SV"; }
inline std::string _addOrientation(int32_t const) { return "NtoN"; }
template <typename TConfig, typename A, typename B> inline void _alternateAlignments(TConfig const&, A&, B&) {}
struct Graph { std::map<std::string, std::size_t> smap; };
template <typename TConfig> inline bool parseGfa(TConfig const&, Graph&) { return false; }
}  // namespace torali

#include "cluster.h"

namespace {
struct RefConfig2 {
  uint16_t minCliqueSize;
  uint32_t minRefSep, maxReadSep, graphPruning;
  int32_t nchr;
  // findJunctions (src/junction.h:319-460)
  uint16_t minMapQual = 0;
  uint32_t minClip = 0;
  float indelExtension = 0.5f;
  std::vector<boost::filesystem::path> files;
  boost::filesystem::path genome;
};

// ---- in-memory stand-in for the handful of htslib calls findJunctions makes (no htslib library is linked) ----
struct MemRecord { int32_t tid, pos; uint16_t flag; uint8_t mapq; std::vector<uint8_t> data; uint16_t l_qname; uint32_t n_cigar; };
std::vector<MemRecord> g_records;   // the "BAM file", in file order
int32_t g_ntargets = 0;
struct MemInterval { uint32_t lo, hi; uint32_t lower() const { return lo; } uint32_t upper() const { return hi; } };
}  // namespace

extern "C" {

// SR clustering (src/cluster.h:324-442). br: n x [chr,pos,chr2,pos2,rstart,sstart,qual,inslen] + ids; must be sorted.
// sv_out: up to cap x 14 ints [chr,svStart,chr2,svEnd,ciposlow,ciposhigh,ciendlow,ciendhigh,srSupport,srMapQuality,mapq,insLen,svt,id]
int ref_cluster_sr(const int32_t* br8, const uint64_t* ids, int n, int svt, int minClique, int maxReadSep, int graphPruning, int nchr,
                   int32_t* svid_out, int32_t* sv_out, int cap) {
  RefConfig2 c; c.minCliqueSize = (uint16_t) minClique; c.maxReadSep = maxReadSep; c.graphPruning = graphPruning; c.nchr = nchr; c.minRefSep = 0;
  std::vector<torali::SRBamRecord> br;
  for (int i = 0; i < n; ++i) {
    const int32_t* r = br8 + 8 * i;
    br.push_back(torali::SRBamRecord(r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], (std::size_t) ids[i]));
  }
  std::vector<torali::StructuralVariantRecord> sv;
  torali::cluster(c, br, sv, svt);
  for (int i = 0; i < n; ++i) svid_out[i] = br[i].svid;
  if ((int) sv.size() > cap) return -1;
  for (std::size_t i = 0; i < sv.size(); ++i) {
    int32_t* o = sv_out + 14 * i;
    o[0] = sv[i].chr; o[1] = sv[i].svStart; o[2] = sv[i].chr2; o[3] = sv[i].svEnd; o[4] = sv[i].ciposlow; o[5] = sv[i].ciposhigh;
    o[6] = sv[i].ciendlow; o[7] = sv[i].ciendhigh; o[8] = sv[i].srSupport; o[9] = sv[i].srMapQuality; o[10] = sv[i].mapq; o[11] = sv[i].insLen;
    o[12] = sv[i].svt; o[13] = sv[i].id;
  }
  return (int) sv.size();
}

// PE clustering (src/cluster.h:528-629). rec: n x [tid,pos,mtid,mpos,alen,malen,Median,Mad,maxNormalISize,flag,MapQuality]; must be sorted.
// sv_out: cap x 12 [chr,svStart,chr2,svEnd,ciposlow,ciposhigh,ciendlow,ciendhigh,peSupport,peMapQuality,mapq,svt]
int ref_cluster_pe(const int32_t* rec11, int n, int svt, int minClique, int graphPruning, uint32_t varisize, int32_t* sv_out, int cap) {
  RefConfig2 c; c.minCliqueSize = (uint16_t) minClique; c.graphPruning = graphPruning; c.maxReadSep = 0; c.nchr = 0; c.minRefSep = 0;
  std::vector<torali::BamAlignRecord> v;
  bam1_t bb; memset(&bb, 0, sizeof(bb)); bam1_t* b = &bb;  // no htslib library is linked: a zeroed record is enough for the constructor
  for (int i = 0; i < n; ++i) {
    const int32_t* r = rec11 + 11 * i;
    b->core.tid = r[0]; b->core.pos = r[1]; b->core.mtid = r[2]; b->core.mpos = r[3]; b->core.flag = (uint16_t) r[9];
    v.push_back(torali::BamAlignRecord(b, (uint8_t) r[10], (uint16_t) r[4], (uint16_t) r[5], r[6], r[7], r[8]));
  }
  std::vector<torali::StructuralVariantRecord> sv;
  torali::cluster(c, v, sv, varisize, svt);
  if ((int) sv.size() > cap) return -1;
  for (std::size_t i = 0; i < sv.size(); ++i) {
    int32_t* o = sv_out + 12 * i;
    o[0] = sv[i].chr; o[1] = sv[i].svStart; o[2] = sv[i].chr2; o[3] = sv[i].svEnd; o[4] = sv[i].ciposlow; o[5] = sv[i].ciposhigh;
    o[6] = sv[i].ciendlow; o[7] = sv[i].ciendhigh; o[8] = sv[i].peSupport; o[9] = sv[i].peMapQuality; o[10] = sv[i].mapq; o[11] = sv[i].svt;
  }
  return (int) sv.size();
}

// Junction selection (src/junction.h:60-316: select* + bridgeInsertions, driven as fetchSVs :463-475 does with an empty svtset).
// Reads are given in the caller's order; junc: [forward,scleft,refidx,rstart,refpos,seqpos,qual] x total, read_off[nreads+1].
// out: per svt (0..8) records [chr,pos,chr2,pos2,rstart,sstart,qual,inslen,primaryChr] + id, concatenated; out_cnt[9].
int ref_select_junctions(const int32_t* junc7, const uint32_t* read_off, const uint64_t* read_id, int nreads, int maxReadSep, int minRefSep,
                         int32_t* out9, uint64_t* out_id, int cap, int32_t* out_cnt) {
  RefConfig2 c; c.maxReadSep = maxReadSep; c.minRefSep = minRefSep; c.minCliqueSize = 2; c.graphPruning = 1000; c.nchr = 0;
  // an ordered container: iteration = ascending read id (the reference iterates a boost::unordered_map, whose order is unspecified;
  // record ORDER within each svt list is irrelevant downstream because the lists are sorted before clustering, src/shortpe.h:490)
  typedef std::map<std::size_t, std::vector<torali::Junction> > TReadBp;
  TReadBp readBp;
  for (int r = 0; r < nreads; ++r) {
    std::vector<torali::Junction> v;
    for (uint32_t k = read_off[r]; k < read_off[r + 1]; ++k) {
      const int32_t* j = junc7 + 7 * k;
      v.push_back(torali::Junction(j[0] != 0, j[1] != 0, j[2], j[3], j[4], j[5], (uint16_t) j[6]));
    }
    readBp.insert(std::make_pair((std::size_t) read_id[r], v));
  }
  std::vector<std::vector<torali::SRBamRecord> > br(2 * DELLY_SVT_TRANS, std::vector<torali::SRBamRecord>());
  torali::selectDeletions(c, readBp, br);
  torali::selectDuplications(c, readBp, br);
  torali::selectInversions(c, readBp, br);
  torali::selectInsertions(c, readBp, br);
  torali::bridgeInsertions(readBp, br);
  torali::selectTranslocations(c, readBp, br);
  int pos = 0;
  for (int svt = 0; svt < 9; ++svt) {
    out_cnt[svt] = (int32_t) br[svt].size();
    for (auto const& x : br[svt]) {
      if (pos >= cap) return -1;
      int32_t* o = out9 + 9 * pos;
      o[0] = x.chr; o[1] = x.pos; o[2] = x.chr2; o[3] = x.pos2; o[4] = x.rstart; o[5] = x.sstart; o[6] = x.qual; o[7] = x.inslen; o[8] = x.primaryChr;
      out_id[pos] = x.id;
      ++pos;
    }
  }
  return pos;
}

// ---- findJunctions (src/junction.h:319-460) verbatim over an in-memory record list ------------------------------
hts_itr_t* sam_itr_queryi(const hts_idx_t*, int tid, hts_pos_t beg, hts_pos_t end) {
  hts_itr_t* it = (hts_itr_t*) calloc(1, sizeof(hts_itr_t));
  it->tid = tid; it->beg = beg; it->end = end; it->i = 0;
  return it;
}
int hts_itr_next(BGZF*, hts_itr_t* it, void* r, void*) {
  bam1_t* b = (bam1_t*) r;
  while (it->i < (int) g_records.size()) {
    MemRecord& m = g_records[it->i++];
    if (m.tid != it->tid || m.pos < it->beg || m.pos >= it->end) continue;
    memset(&b->core, 0, sizeof(b->core));
    b->core.tid = m.tid; b->core.pos = m.pos; b->core.flag = m.flag; b->core.qual = m.mapq; b->core.l_qname = m.l_qname; b->core.n_cigar = m.n_cigar;
    b->data = m.data.data(); b->l_data = (int) m.data.size(); b->m_data = (uint32_t) m.data.size();
    return 0;
  }
  return -1;
}
int hts_itr_multi_next(htsFile*, hts_itr_t*, void*) { return -1; }
void hts_itr_destroy(hts_itr_t* it) { free(it); }
htsFile* hts_open(const char*, const char*) { htsFile* f = (htsFile*) calloc(1, sizeof(htsFile)); f->is_bgzf = 1; return f; }
int hts_close(htsFile* f) { free(f); return 0; }
int hts_set_fai_filename(htsFile*, const char*) { return 0; }
hts_idx_t* sam_index_load(htsFile*, const char*) { return (hts_idx_t*) &g_ntargets; }
void hts_idx_destroy(hts_idx_t*) {}
sam_hdr_t* sam_hdr_read(samFile*) { sam_hdr_t* h = (sam_hdr_t*) calloc(1, sizeof(sam_hdr_t)); h->n_targets = g_ntargets; return h; }
void sam_hdr_destroy(sam_hdr_t* h) { free(h); }
bam1_t* bam_init1(void) { return (bam1_t*) calloc(1, sizeof(bam1_t)); }
void bam_destroy1(bam1_t* b) { free(b); }   // the record data belongs to g_records
void hts_log(enum htsLogLevel, const char*, const char*, ...) {}

// rec: n x [tid, pos, flag, mapq, read id, cigar_off, n_cigar]; cigar: BAM-encoded uint32 (len << 4 | op).
// out: reads in ascending seed order: read_seed[k], read_off[k+1]; junctions 7 ints [forward, scleft, refidx, rstart, refpos, seqpos, qual].
int ref_find_junctions(const int32_t* rec7, int n, const uint32_t* cigar, int ntargets, int minMapQual, int minClip, int minRefSep, float indelExtension,
                       uint64_t* read_seed, uint32_t* read_off, int read_cap, int32_t* junc7, int junc_cap, int* n_reads) {
  RefConfig2 c; c.minCliqueSize = 2; c.maxReadSep = 0; c.graphPruning = 0; c.nchr = ntargets; c.minRefSep = minRefSep;
  c.minMapQual = (uint16_t) minMapQual; c.minClip = minClip; c.indelExtension = indelExtension;
  c.files.push_back(boost::filesystem::path("in-memory.bam")); c.genome = boost::filesystem::path("in-memory.fa");
  g_ntargets = ntargets;
  g_records.clear();
  for (int i = 0; i < n; ++i) {
    const int32_t* r = rec7 + 7 * i;
    MemRecord m; m.tid = r[0]; m.pos = r[1]; m.flag = (uint16_t) r[2]; m.mapq = (uint8_t) r[3];
    std::string qn = "r" + std::to_string(r[4]);
    m.l_qname = (uint16_t) ((qn.size() + 1 + 3) & ~3u);   // NUL-terminated, padded to 4 like htslib does
    m.n_cigar = (uint32_t) r[6];
    m.data.assign(m.l_qname + 4 * (size_t) m.n_cigar, 0);
    memcpy(m.data.data(), qn.data(), qn.size());
    memcpy(m.data.data() + m.l_qname, cigar + r[5], 4 * (size_t) m.n_cigar);
    g_records.push_back(m);
  }
  std::vector<std::vector<MemInterval> > validRegions(ntargets);
  for (int t = 0; t < ntargets; ++t) validRegions[t].push_back(MemInterval{0u, 0x7fffffffu});
  typedef std::map<std::size_t, std::vector<torali::Junction> > TReadBp;   // ordered: ascending seed (see ref_select_junctions)
  TReadBp readBp;
  std::set<std::size_t> validSR;
  std::streambuf* old = std::cerr.rdbuf(nullptr);
  torali::findJunctions(c, validRegions, readBp, validSR);
  std::cerr.rdbuf(old);
  int k = 0, pos = 0;
  read_off[0] = 0;
  for (auto const& kv : readBp) {
    if (k >= read_cap) return -1;
    read_seed[k] = kv.first;
    for (auto const& j : kv.second) {
      if (pos >= junc_cap) return -1;
      int32_t* o = junc7 + 7 * pos++;
      o[0] = j.forward; o[1] = j.scleft; o[2] = j.refidx; o[3] = j.rstart; o[4] = j.refpos; o[5] = j.seqpos; o[6] = j.qual;
    }
    read_off[++k] = (uint32_t) pos;
  }
  *n_reads = k;
  return pos;
}

// the read id findJunctions derives from the query name (hash_lr as restated above), so that callers can map seeds back
uint64_t ref_hash_lr_name(const char* qname) { return (uint64_t) std::hash<std::string>()(qname); }

// msaEdlib (src/assemble.h:385-473) on one cluster of reads
int ref_msa_edlib(const char* arena, const uint32_t* off, const uint32_t* len, int nreads, int minClique, char* cons, int cons_cap, int* cons_len) {
  RefConfig2 c; c.minCliqueSize = (uint16_t) minClique; c.maxReadSep = 0; c.minRefSep = 0; c.graphPruning = 0; c.nchr = 0;
  std::vector<std::string> sps;
  for (int i = 0; i < nreads; ++i) sps.push_back(std::string(arena + off[i], len[i]));
  std::string cs;
  int rows = torali::msaEdlib(c, sps, cs);
  *cons_len = (int) cs.size();
  if ((int) cs.size() > cons_cap) return -1;
  memcpy(cons, cs.data(), cs.size());
  return rows;
}

// msaWfa (src/assemble.h:549-725, incl. _trimConsensus when both flanks are given) on one cluster of reads
int ref_msa_wfa(const char* arena, const uint32_t* off, const uint32_t* len, int nreads, int minClique, const char* prefix, int prefix_len, const char* suffix,
                int suffix_len, char* cons, int cons_cap, int* cons_len) {
  RefConfig2 c; c.minCliqueSize = (uint16_t) minClique; c.maxReadSep = 0; c.minRefSep = 0; c.graphPruning = 0; c.nchr = 0;
  std::vector<std::string> sps;
  for (int i = 0; i < nreads; ++i) sps.push_back(std::string(arena + off[i], len[i]));
  std::string cs;
  int rows = torali::msaWfa(c, sps, cs, std::string(prefix, prefix_len), std::string(suffix, suffix_len));
  *cons_len = (int) cs.size();
  if ((int) cs.size() > cons_cap) return -1;
  memcpy(cons, cs.data(), cs.size());
  return rows;
}

}  // extern "C"
