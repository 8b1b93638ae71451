// delly_b200 — the drop-in binding: the `delly sr` / `delly lr` command-line surface (BAM + FASTA in, BCF out) over the batched
// B200 path. It is what a maintainer gets by replacing the per-item alignment calls of the reference's drivers with the batched
// C ABI (include/dgpu.h): htslib reads the alignments into record lists (bindings/hts_io.hpp), the stage mirrors of delly_b200/host
// run the reference's stage sequence with every alignment batch on the device, htslib writes the BCF.
//   src/delly.h:199-400  `delly sr` options and checks -> dellyRun (:86-196)
//   src/tegua.h:209-440  `delly lr` options and checks -> runTegua (:78-207)
// Parity: tests/test_bcf_parity.py compares the BCF written here byte for byte (uncompressed stream, ##fileDate aside) with the BCF the
// reference's own drivers write (oracle/_ref/delly_ref = src/delly.h + src/tegua.h compiled verbatim against the same htslib).
// No CPU fallback: without an sm_100 device dgpu_ctx_create fails and so does this program.
#include <condition_variable>
#include <cstdlib>
#include <mutex>
#include <cstring>
#include <iostream>
#include <string>
#include <sys/stat.h>
#include <unistd.h>
#include <thread>
#include <vector>

#include "hts_io.hpp"
#include "../delly_b200/host/gather.hpp"
#include "../include/dgpu.h"

using namespace dellyb200;

namespace {

struct Options {
  bool lr = false;
  std::string svtype = "ALL", mode = "ont";
  std::string genome, exclude, outfile = "-", vcffile, dumpfile, meifile;
  std::vector<std::string> files;
  bool hasExclude = false, hasVcf = false, hasOutfile = false, help = false;
  int device = 0, ioThreads = 16;
  int rank = 0, nranks = 1;          // one process per GPU: --rank r --nranks N --comm-file path (rank 0 publishes the NCCL id there)
  std::string commFile, timingFile;  // --timing file: stage wall-clock times as one JSON object
  float indelExtension = 0.5f;
  MethylConfig methyl;
  AnnoConfig anno;
  Config c;
};

bool fileOk(std::string const& p) { struct stat st; return ::stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0; }

// src/util.h:370-395 (_svTypesToCompute): DEL -> 2, INS -> 4, DUP -> 3, INV -> 0,1, BND -> 5..8, ALL -> no restriction
bool svTypesToCompute(Config& c, std::string const& svtype) {
  c.svtset.clear();
  std::size_t b = 0;
  bool all = false;
  while (b <= svtype.size()) {
    std::size_t e = svtype.find(',', b);
    if (e == std::string::npos) e = svtype.size();
    const std::string t = svtype.substr(b, e - b);
    if (t == "ALL") all = true;
    else if (t == "DEL") c.svtset.insert(2);
    else if (t == "INS") c.svtset.insert(4);
    else if (t == "DUP") c.svtset.insert(3);
    else if (t == "INV") { c.svtset.insert(0); c.svtset.insert(1); }
    else if (t == "BND") { c.svtset.insert(5); c.svtset.insert(6); c.svtset.insert(7); c.svtset.insert(8); }
    else return false;
    b = e + 1;
  }
  if (all) c.svtset.clear();
  return true;
}

void usage(Options const& o) {
  std::cerr << "\nUsage: delly_b200 " << (o.lr ? "lr" : "sr") << " [OPTIONS] -g <ref.fa> <sample1.sort.bam> <sample2.sort.bam> ...\n"
            << "  options follow `delly " << (o.lr ? "lr" : "sr") << "` (same letters and defaults); extra: --device N, --io-threads N"
            << (o.lr ? ", --mei <templates.fa>" : "") << "\n";
}

// `delly sr` (src/delly.h:207-246) and `delly lr` (src/tegua.h:215-275) share most letters; the long names are the reference's.
bool parse(int argc, char** argv, Options& o) {
  Config& c = o.c;
  if (o.lr) { c = Config::longRead(); c.minCliqueSize = 3; c.minRefSep = 30; c.maxReadPerSV = 15; }
  else c = Config::shortRead();
  c.maxThreads = 4;
  o.methyl.methylProb = 128; o.methyl.methylWindow = 1000; o.methyl.minCpgDepth = 5;
  auto need = [&](int& i) -> const char* { if (i + 1 >= argc) { std::cerr << "missing value for " << argv[i] << std::endl; std::exit(2); } return argv[++i]; };
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    std::string v;
    bool inl = false;
    if (a.size() > 2 && a[0] == '-' && a[1] == '-') { const std::size_t e = a.find('='); if (e != std::string::npos) { v = a.substr(e + 1); a = a.substr(0, e); inl = true; } }
    else if (a.size() > 2 && a[0] == '-' && a[1] != '-') { v = a.substr(2); a = a.substr(0, 2); inl = true; }
    auto val = [&]() -> std::string { return inl ? v : std::string(need(i)); };
    auto is = [&](const char* s, const char* l) { return a == s || a == l; };
    if (a.empty() || a[0] != '-' || a == "-") { o.files.push_back(a); continue; }
    if (is("-?", "--help")) o.help = true;
    else if (is("-t", "--svtype")) o.svtype = val();
    else if (is("-g", "--genome")) o.genome = val();
    else if (is("-x", "--exclude")) { o.exclude = val(); o.hasExclude = true; }
    else if (is("-o", "--outfile")) { o.outfile = val(); o.hasOutfile = true; }
    else if (is("-q", "--map-qual")) c.minMapQual = (uint16_t) std::stoi(val());
    else if (is("-c", "--minclip")) c.minClip = (uint32_t) std::stoul(val());
    else if (is("-z", "--min-clique-size")) c.minCliqueSize = (uint16_t) std::stoi(val());
    else if (is("-m", "--minrefsep")) c.minRefSep = (uint32_t) std::stoul(val());
    else if (is("-n", "--maxreadsep")) c.maxReadSep = (uint32_t) std::stoul(val());
    else if (is("-p", "--max-reads")) c.maxReadPerSV = (uint32_t) std::stoul(val());
    else if (is("-v", "--vcffile")) { o.vcffile = val(); o.hasVcf = true; }
    else if (is("-u", "--geno-qual")) c.minGenoQual = (uint16_t) std::stoi(val());
    else if (is("-d", "--dump")) o.dumpfile = val();
    else if (is("-w", "--cons-window")) c.minConsWindow = std::stoi(val());
    else if (a == "--device") o.device = std::stoi(val());
    else if (a == "--io-threads") o.ioThreads = std::stoi(val());
    else if (a == "--rank") o.rank = std::stoi(val());
    else if (a == "--nranks") o.nranks = std::stoi(val());
    else if (a == "--comm-file") o.commFile = val();
    else if (a == "--timing") o.timingFile = val();
    else if (!o.lr && is("-h", "--threads")) c.maxThreads = (uint32_t) std::stoul(val());
    else if (!o.lr && is("-r", "--qual-tra")) c.minTraQual = (uint16_t) std::stoi(val());
    else if (!o.lr && is("-s", "--mad-cutoff")) c.madCutoff = (uint16_t) std::stoi(val());
    else if (!o.lr && is("-a", "--max-geno-count")) c.maxGenoReadCount = (uint32_t) std::stoul(val());
    else if (!o.lr && is("-j", "--pruning")) c.graphPruning = (uint32_t) std::stoul(val());
    else if (o.lr && is("-y", "--technology")) o.mode = val();
    else if (o.lr && is("-f", "--flank-size")) c.minimumFlankSize = std::stoi(val());
    else if (o.lr && is("-a", "--flank-quality")) c.flankQuality = std::stof(val());
    else if (o.lr && is("-i", "--indel-size")) c.indelsize = std::stoi(val());
    else if (o.lr && is("-k", "--mei-minfrac")) o.anno.meiMinFrac = std::stof(val());
    else if (o.lr && is("-r", "--tr-minfrac")) o.anno.trMinFrac = std::stof(val());
    else if (o.lr && is("-j", "--methyl-window")) o.methyl.methylWindow = std::stoi(val());
    else if (o.lr && is("-e", "--methyl-prob")) o.methyl.methylProb = (uint16_t) std::stoi(val());
    else if (o.lr && a == "--min-cpg-depth") o.methyl.minCpgDepth = (uint32_t) std::stoul(val());
    else if (o.lr && is("-b", "--max-geno-count")) c.maxGenoReadCount = (uint32_t) std::stoul(val());
    else if (o.lr && a == "--pruning") c.graphPruning = (uint32_t) std::stoul(val());
    else if (o.lr && a == "--extension") o.indelExtension = std::stof(val());
    else if (o.lr && a == "--read-cap") c.genoCap = std::stoi(val());
    else if (o.lr && a == "--mei") o.meifile = val();
    else if (o.lr && is("-l", "--alt-align")) { std::cerr << "alternate alignments (-l) are outside the accelerated path (SURVEY section 2)" << std::endl; return false; }
    else { std::cerr << "unrecognised option '" << argv[i] << "'" << std::endl; return false; }
  }
  return true;
}

// the template sequences of the reference's annotation step (class MEI, src/svanno.h:30-39) are data of the reference; the build
// extracts them into a FASTA next to the binary (tools/extract_mei.py), `--mei` names another file. Records: >alu >line1 >sva >numt >soloLTR >hervk >polyA
bool loadMei(std::string const& path, MeiTemplates& mei) {
  std::ifstream f(path.c_str());
  if (!f.is_open()) return false;
  std::string line, name;
  int found = 0;
  while (std::getline(f, line)) {
    if (line.empty()) continue;
    if (line[0] == '>') { name = line.substr(1); continue; }
    int k = (name == "alu") ? 1 : (name == "line1") ? 2 : (name == "sva") ? 3 : (name == "numt") ? 4 : (name == "soloLTR") ? 5 : (name == "hervk") ? 6 : (name == "polyA") ? 0 : -1;
    if (k < 0) continue;
    if (k == 0) mei.polyA += line; else mei.seq[k] += line;
    ++found;
  }
  return found >= 7;
}

std::string exeDir(const char* argv0) {
  char buf[4096];
  const ssize_t n = ::readlink("/proc/self/exe", buf, sizeof(buf) - 1);
  std::string p = (n > 0) ? std::string(buf, (std::size_t) n) : std::string(argv0);
  const std::size_t k = p.rfind('/');
  return (k == std::string::npos) ? "." : p.substr(0, k);
}

// The device side of the process starts while the host reads its inputs: creating the CUDA context (and, with several ranks, the NCCL
// communicator plus one tiny warm-up exchange that sets up the NVLink connections) costs 0.5 - 2 s; a background thread pays it during the BAM scan.
class DeviceSession {
 public:
  DeviceSession(Options const& o) : o_(o) { th_ = std::thread([this] { start(); }); }
  ~DeviceSession() {
    if (th_.joinable()) th_.join();
    if (comm_) dgpu_comm_destroy(commCtx_, comm_);
    if (commCtx_) dgpu_ctx_destroy(commCtx_);
    if (ctx_) dgpu_ctx_destroy(ctx_);
  }
  // the compute context (needed by the first device stage); the communicator keeps starting in the background until the first exchange needs it
  bool wait() { waitFor(1); return ctx_ != nullptr; }
  bool waitComm() { waitFor(2); return commOk_; }
  dgpu_ctx* ctx() { return ctx_; }
  Shard shard() {
    Shard sh;
    if (o_.nranks > 1) {
      sh.rank = o_.rank; sh.nranks = o_.nranks;
      sh.gather = [this](std::string const& local, std::vector<std::string>& parts) -> int {
        if (!waitComm()) return DGPU_ERR_NCCL;
        return gatherStrings(commCtx_, comm_, local, parts);
      };
    }
    return sh;
  }
  static int gatherStrings(dgpu_ctx* ctx, void* comm, std::string const& local, std::vector<std::string>& parts) {
    void* all = nullptr; uint64_t* counts = nullptr; int n = 0;
    const int grc = dgpu_gather_records(ctx, comm, local.data(), (uint64_t) local.size(), &all, &counts, &n);
    if (grc) return grc;
    parts.clear();
    const char* p = (const char*) all;
    for (int r = 0; r < n; ++r) { parts.emplace_back(p, (std::size_t) counts[r]); p += counts[r]; }
    dgpu_free_host(all); dgpu_free_host(counts);
    return DGPU_OK;
  }
 private:
  void start();
  void reach(int stage) { { std::lock_guard<std::mutex> g(m_); stage_ = stage; } cv_.notify_all(); }
  void waitFor(int stage) { std::unique_lock<std::mutex> g(m_); cv_.wait(g, [&] { return stage_ >= stage; }); }
  Options const& o_;
  std::thread th_;
  std::mutex m_;
  std::condition_variable cv_;
  int stage_ = 0;              // 1: compute context ready (or failed), 2: communicator ready (or failed / not needed)
  dgpu_ctx* ctx_ = nullptr;
  dgpu_ctx* commCtx_ = nullptr;   // the exchange has its own context (stream, staging buffers): it starts while the compute context already works
  void* comm_ = nullptr;
  bool commOk_ = false;
};

struct Inputs {
  std::vector<io::AlignmentFile> files;
  std::vector<std::string> sampleName;
  std::vector<uint32_t> target_len;
  std::vector<std::string> target_name;
  io::TRegionsGenome validRegions, wholeContigs;
  io::Genome genome;
};

bool openInputs(Options& o, Inputs& in) {
  if (!fileOk(o.genome)) { std::cerr << "Reference file is missing: " << o.genome << std::endl; return false; }
  {
    faidx_t* fai = fai_load(o.genome.c_str());
    if (!fai) { if (fai_build(o.genome.c_str()) == -1) { std::cerr << "Fail to open genome fai index for " << o.genome << std::endl; return false; } }
    else fai_destroy(fai);
  }
  in.files.resize(o.files.size());
  in.sampleName.resize(o.files.size());
  o.c.nchr = 0;
  for (std::size_t f = 0; f < o.files.size(); ++f) {
    if (!fileOk(o.files[f])) { std::cerr << "Alignment file is missing: " << o.files[f] << std::endl; return false; }
    if (!in.files[f].open(o.files[f], o.genome, o.ioThreads)) return false;
    if (!o.c.nchr) o.c.nchr = in.files[f].hdr->n_targets;
    else if (o.c.nchr != in.files[f].hdr->n_targets) { std::cerr << "BAM files have different number of chromosomes!" << std::endl; return false; }
    std::string sampleName = "unknown";
    io::getSMTag(std::string(in.files[f].hdr->text ? in.files[f].hdr->text : sam_hdr_str(in.files[f].hdr)), io::fileStem(o.files[f]), sampleName);
    in.sampleName[f] = sampleName;
  }
  io::checkSampleNames(in.sampleName);
  if (o.hasExclude && !fileOk(o.exclude)) { std::cerr << "Exclude file is missing: " << o.exclude << std::endl; return false; }
  if (o.hasVcf && !fileOk(o.vcffile)) { std::cerr << "Input VCF/BCF file is missing: " << o.vcffile << std::endl; return false; }
  bam_hdr_t* hdr = in.files[0].hdr;
  for (int32_t i = 0; i < hdr->n_targets; ++i) { in.target_len.push_back(hdr->target_len[i]); in.target_name.push_back(hdr->target_name[i]); }
  if (!io::parseExcludeIntervals(o.hasExclude, o.exclude, hdr, in.validRegions)) { std::cerr << "Delly couldn't parse exclude intervals!" << std::endl; return false; }
  in.wholeContigs.assign(hdr->n_targets, std::vector<io::Interval>());
  for (int32_t i = 0; i < hdr->n_targets; ++i) in.wholeContigs[i].push_back(io::Interval{0u, hdr->target_len[i]});
  return in.genome.load(o.genome, hdr, o.lr);
}


void writeTiming(Options const& o, StageClock const& clock, std::size_t nsv) {
  if (o.timingFile.empty()) return;
  std::ofstream f((o.timingFile + (o.nranks > 1 ? "." + std::to_string(o.rank) : std::string())).c_str());
  double total = 0;
  f << "{\"rank\": " << o.rank << ", \"nranks\": " << o.nranks << ", \"svs\": " << nsv << ", \"stages_ms\": {";
  for (std::size_t i = 0; i < clock.ms.size(); ++i) { f << (i ? ", " : "") << "\"" << clock.ms[i].first << "\": " << clock.ms[i].second; total += clock.ms[i].second; }
  f << "}, \"total_ms\": " << total << "}" << std::endl;
}

// One process per GPU: rank 0 creates the NCCL id and publishes it in --comm-file (written under a temporary name, then renamed); the other
// ranks wait for the file. The communicator lives as long as the process.
bool initComm(Options const& o, dgpu_ctx* ctx, void** comm) {
  uint8_t id[128];
  if (o.commFile.empty()) { std::cerr << "--nranks > 1 needs --comm-file" << std::endl; return false; }
  if (o.rank == 0) {
    if (dgpu_comm_unique_id(id) != DGPU_OK) { std::cerr << "NCCL is not available (libnccl.so.2)" << std::endl; return false; }
    const std::string tmp = o.commFile + ".tmp";
    { std::ofstream f(tmp.c_str(), std::ios::binary); f.write((const char*) id, 128); }
    if (std::rename(tmp.c_str(), o.commFile.c_str()) != 0) { std::cerr << "cannot publish " << o.commFile << std::endl; return false; }
  } else {
    for (int tries = 0; tries < 6000; ++tries) {
      std::ifstream f(o.commFile.c_str(), std::ios::binary);
      if (f.good()) { f.read((char*) id, 128); if (f.gcount() == 128) break; }
      if (tries == 5999) { std::cerr << "timed out waiting for " << o.commFile << std::endl; return false; }
      usleep(10000);
    }
  }
  const int rc = dgpu_comm_init(ctx, o.nranks, o.rank, id, comm);
  if (rc != DGPU_OK) { std::cerr << "dgpu_comm_init failed (" << rc << "): " << dgpu_last_error(ctx) << std::endl; return false; }
  return true;
}

void DeviceSession::start() {
  const int rc = dgpu_ctx_create(o_.device, &ctx_);
  if (rc != DGPU_OK) { std::cerr << "delly_b200: no usable sm_100 device (dgpu_ctx_create = " << rc << "); there is no CPU fallback on this path" << std::endl; ctx_ = nullptr; reach(2); return; }
  reach(1);
  if (o_.nranks > 1) {
    std::vector<std::string> parts;
    if (dgpu_ctx_create(o_.device, &commCtx_) != DGPU_OK) { commCtx_ = nullptr; std::cerr << "delly_b200: cannot create the exchange context" << std::endl; }
    else if (!initComm(o_, commCtx_, &comm_)) { }
    else if (gatherStrings(commCtx_, comm_, std::string("warm-up"), parts) != DGPU_OK || (int) parts.size() != o_.nranks) std::cerr << "NCCL warm-up exchange failed: " << dgpu_last_error(commCtx_) << std::endl;
    else commOk_ = true;
  } else commOk_ = true;
  reach(2);
}

int runSr(Options& o, DeviceSession& dev) {
  StageClock clock;
  Inputs in;
  if (!openInputs(o, in)) return 1;
  clock.lap("open_inputs");
  Config& c = o.c;
  const std::size_t F = in.files.size();
  // records of every file, as the reference's iterators return them
  std::vector<std::vector<SrRecord> > recs(F);
  std::vector<std::vector<SrRecord> const*> samples;
  for (std::size_t f = 0; f < F; ++f) {
    if (!in.files[f].readRegionsParallel(o.genome, in.validRegions, recs[f], io::toSrRecord, o.ioThreads)) { std::cerr << "cannot read " << o.files[f] << std::endl; return 1; }
    samples.push_back(&recs[f]);
  }
  // exclude file: scan, clustering, assembly and the library estimate see the valid regions (src/shortpe.h:349, src/util.h:810), the genotyping pass
  // the whole contigs (src/coverage.h) — a second record list
  std::vector<std::vector<SrRecord> > recsWhole(o.hasExclude ? F : 0);
  std::vector<std::vector<SrRecord> const*> genoSamples;
  if (o.hasExclude) {
    for (std::size_t f = 0; f < F; ++f) {
      if (!in.files[f].readRegionsParallel(o.genome, in.wholeContigs, recsWhole[f], io::toSrRecord, o.ioThreads)) { std::cerr << "cannot read " << o.files[f] << std::endl; return 1; }
      genoSamples.push_back(&recsWhole[f]);
    }
    c.contigExcluded.assign(in.validRegions.size(), 0);
    for (std::size_t r = 0; r < in.validRegions.size(); ++r) c.contigExcluded[r] = in.validRegions[r].empty();
  }
  clock.lap("read_bam");
  std::vector<LibraryInfo> libs(F);
  for (std::size_t f = 0; f < F; ++f) {
    getLibraryParams(c, in.target_len, recs[f], libs[f]);
    if (libs[f].rs == 0) { std::cerr << "Sample has not enough data to estimate library parameters! File: " << o.files[f] << std::endl; return 1; }
  }
  clock.lap("library");
  if (!dev.wait()) return 3;
  clock.lap("wait_device");
  dgpu_ctx* ctx = dev.ctx();
  const Shard shard = dev.shard();
  SrMultiCallSet cs;
  int rc;
  if (!o.hasVcf) rc = dellySrCallSharded(ctx, c, libs, in.target_len, in.target_name, in.genome.cseq, samples, shard, cs, &clock, o.hasExclude ? &genoSamples : nullptr);
  else {
    std::vector<VcfSite> sites;
    bool hasConsBp = false;
    if (!io::readSites(o.vcffile, sites, hasConsBp)) return 1;
    rc = dellySrGenotypeSharded(ctx, c, libs, in.target_len, in.target_name, in.genome.cseq, sites, hasConsBp, o.hasExclude ? genoSamples : samples, shard, cs, &clock);
    if (rc == DGPU_ERR_ARG && !cs.sample.empty()) { std::cerr << "Error: Delly genotyping requires a Delly BCF file from v1.1.7 or later!" << std::endl; rc = DGPU_OK; }
  }
  if (rc) { std::cerr << "delly_b200: device path failed (" << rc << "): " << dgpu_last_error(ctx) << std::endl; return 1; }
  if (o.rank != 0) { writeTiming(o, clock, cs.svs.size()); return 0; }   // every rank holds the complete call set; rank 0 writes it
  // src/delly.h:161-178: the count maps stay empty (and the header gets no sample column) when there is nothing to genotype
  std::vector<VcfSample> vs;
  if (!cs.svs.empty())
    for (std::size_t f = 0; f < F; ++f) {
      VcfSample s; s.name = in.sampleName[f]; s.jctMap = &cs.sample[f].jctMap; s.rcMap = &cs.sample[f].rcMap; s.spanMap = &cs.sample[f].spanMap;
      vs.push_back(s);
    }
  io::HtsVcfWriter w(o.outfile, in.target_name);
  if (!w.ok()) { std::cerr << "cannot open " << o.outfile << std::endl; return 1; }
  vcfRecordsTo(w, cs.svs, vs, in.target_name, in.target_len, o.genome, io::todayIso(), o.hasVcf, 0);
  w.close();
  clock.lap("write_bcf");
  writeTiming(o, clock, cs.svs.size());
  for (std::size_t f = 0; f < F; ++f)
    std::cerr << "Sample:" << in.sampleName[f] << ",ReadSize=" << libs[f].rs << ",Median=" << libs[f].median << ",MAD=" << libs[f].mad << ",UniqueDiscordantPairs=" << libs[f].abnormal_pairs << std::endl;
  return 0;
}

int runLr(Options& o, DeviceSession& dev, const char* argv0) {
  Inputs in;
  if (!openInputs(o, in)) return 1;
  Config& c = o.c;
  const std::size_t F = in.files.size();
  if (o.mode == "pb") o.indelExtension = 0.7f;
  else if (o.mode == "ont") o.indelExtension = 0.5f;
  MeiTemplates mei;
  const std::string meiPath = o.meifile.empty() ? exeDir(argv0) + "/../data/_mei_templates.fa" : o.meifile;
  if (!loadMei(meiPath, mei)) { std::cerr << "cannot read the annotation templates " << meiPath << " (see --mei)" << std::endl; return 1; }
  std::vector<std::vector<LrRecord> > recs(F);
  std::vector<std::vector<std::size_t> > ids(F);
  std::vector<LrSample> samples;
  for (std::size_t f = 0; f < F; ++f) {
    if (!in.files[f].readRegionsParallel(o.genome, in.validRegions, recs[f], io::toLrRecord, o.ioThreads, &ids[f], io::hashLr)) { std::cerr << "cannot read " << o.files[f] << std::endl; return 1; }
    samples.push_back(LrSample{&recs[f], &ids[f]});
  }
  std::vector<std::vector<LrRecord> > recsWhole(o.hasExclude ? F : 0);   // exclude file: the genotyping pass still reads whole contigs
  std::vector<std::vector<std::size_t> > idsWhole(o.hasExclude ? F : 0);
  std::vector<LrSample> genoSamples;
  if (o.hasExclude)
    for (std::size_t f = 0; f < F; ++f) {
      if (!in.files[f].readRegionsParallel(o.genome, in.wholeContigs, recsWhole[f], io::toLrRecord, o.ioThreads, &idsWhole[f], io::hashLr)) { std::cerr << "cannot read " << o.files[f] << std::endl; return 1; }
      genoSamples.push_back(LrSample{&recsWhole[f], &idsWhole[f]});
    }
  if (!dev.wait()) return 3;
  dgpu_ctx* ctx = dev.ctx();
  LrMultiCallSet cs;
  int rc;
  const Shard shard = dev.shard();
  if (!o.hasVcf) rc = dellyLrCallSharded(ctx, c, o.indelExtension, in.target_len, in.target_name, in.genome.cseq, samples, shard, cs, &mei, o.anno, &o.methyl, o.hasExclude ? &genoSamples : nullptr);
  else {
    std::vector<VcfSite> sites;
    bool hasConsBp = false;
    if (!io::readSites(o.vcffile, sites, hasConsBp)) return 1;
    rc = dellyLrGenotypeSharded(ctx, c, in.target_len, in.target_name, in.genome.cseq, sites, hasConsBp, o.hasExclude ? genoSamples : samples, shard, cs, &mei, o.anno, &o.methyl);
    if (rc == DGPU_ERR_ARG && !cs.sample.empty()) { std::cerr << "Error: Delly genotyping requires a Delly BCF file from v1.1.7 or later!" << std::endl; rc = DGPU_OK; }
  }
  if (rc) { std::cerr << "delly_b200: device path failed (" << rc << "): " << dgpu_last_error(ctx) << std::endl; return 1; }
  if (o.rank != 0) return 0;   // every rank holds the complete call set; rank 0 writes it
  // long reads have no spanning pairs: empty lists per SV (src/tegua.h:173-190)
  std::vector<SpanningCount> noSpan(cs.svs.size());
  std::vector<VcfSample> vs;
  for (std::size_t f = 0; f < F; ++f) {
    VcfSample s; s.name = in.sampleName[f]; s.jctMap = &cs.sample[f].jctMap; s.rcMap = &cs.sample[f].rcMap; s.spanMap = &noSpan; s.methylMap = &cs.sample[f].methyl;
    vs.push_back(s);
  }
  io::HtsVcfWriter w(o.outfile, in.target_name);
  if (!w.ok()) { std::cerr << "cannot open " << o.outfile << std::endl; return 1; }
  vcfRecordsTo(w, cs.svs, vs, in.target_name, in.target_len, o.genome, io::todayIso(), o.hasVcf, o.methyl.minCpgDepth);
  w.close();
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 2 || (std::strcmp(argv[1], "sr") && std::strcmp(argv[1], "lr"))) { std::cerr << "usage: delly_b200 sr|lr [OPTIONS] -g <ref.fa> <in.bam> ..." << std::endl; return 1; }
  Options o;
  o.lr = !std::strcmp(argv[1], "lr");
  if (!parse(argc - 1, argv + 1, o)) return 1;
  if (o.help || o.files.empty() || o.genome.empty()) { usage(o); return 0; }
  if (!svTypesToCompute(o.c, o.svtype)) { std::cerr << "Please specify a valid SV type, i.e., -t INV or -t DEL,INV without spaces." << std::endl; return 1; }
  const uint32_t hw = std::thread::hardware_concurrency() ? std::thread::hardware_concurrency() : 1;
  o.c.maxThreads = std::min<uint32_t>(std::max<uint32_t>(1, o.c.maxThreads), hw);   // src/delly.h:268
  if (o.c.minCliqueSize < 2) o.c.minCliqueSize = 2;
  if (o.c.minMapQual > o.c.minTraQual) o.c.minTraQual = o.c.minMapQual;
  if (o.c.minGenoQual < 5 && !o.lr) o.c.minGenoQual = 5;   // src/delly.h:386
  if (!o.dumpfile.empty()) { std::cerr << "the SV-read dump file (-d) is outside the accelerated path" << std::endl; return 1; }
  DeviceSession dev(o);
  const int r = o.lr ? runLr(o, dev, argv[0]) : runSr(o, dev);
  DeviceLimitLog const& lim = deviceLimitLog();
  if (lim.total())
    std::cerr << "Warning: " << lim.msaClusters.load() << " read cluster(s), " << lim.pathJobs.load() << " edit-path job(s) and " << lim.needleJobs.load()
              << " consensus alignment(s) exceeded a device limit (include/dgpu.h) and were treated as failed alignments of their SV / read; the reference"
                 " has no such limits, so these records may differ from its output." << std::endl;
  // Every output is written and closed at this point. Tearing the CUDA context, the NCCL communicator and a few GB of device scratch down
  // costs a few hundred ms and produces nothing: leave that to the operating system (DGPU_CLEAN_EXIT=1 keeps the orderly teardown, e.g. under
  // compute-sanitizer).
  if (!getenv("DGPU_CLEAN_EXIT")) { std::cout.flush(); std::cerr.flush(); fflush(nullptr); _exit(r); }
  return r;
}
