// oracle/_ref/delly_ref (TEST INFRASTRUCTURE ONLY): the reference's OWN command-line drivers — torali::delly (`delly sr`,
// src/delly.h:199-400 -> dellyRun :86-196) and torali::tegua (`delly lr`, src/tegua.h:209-440 -> runTegua :78-207) — compiled VERBATIM
// from /root/reference/src together with every header they include (util.h, shortpe.h, coverage.h, genotype.h, junction.h, cluster.h,
// assemble.h, modvcf.h, svanno.h, methyl.h, ...) and linked against the REAL htslib (tools/build_htslib.sh builds the reference's vendored
// htslib 1.21). Only Boost is served by stand-ins (oracle/shim11, oracle/shim: program_options, filesystem, icl interval_set, date_time,
// iostreams gzip sink, multi_array, dynamic_bitset, hash helpers; boost::unordered_map -> std::unordered_map) because Boost is not installed
// here, and pangenome.h (`-l` alternate alignments, out of scope) is masked by its include guard.
// This binary reads the same BAM + FASTA and writes the BCF the reference writes: it pins whole-file parity of the product binding
// (bindings/delly_b200_main.cpp) on example/sr.bam and example/lr.bam. Nothing from the reference is copied into this repository.
#define PANGENOME_H
#define ORACLE_REAL_UTIL_H
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <iostream>
#include <fstream>
#include <limits>
#include <numeric>
#include <map>
#include <set>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <boost/filesystem.hpp>
#include <boost/unordered_map.hpp>
#include <boost/program_options.hpp>
#include "shim/prelude.h"
#include <htslib/faidx.h>
#include <htslib/vcf.h>
#include <htslib/sam.h>
#include "util.h"

namespace torali {
// pangenome.h is masked: the names junction.h mentions in its alternate-alignment branch (never taken: no `-l`)
struct Graph { std::map<std::string, std::size_t> smap; };
template <typename TConfig> inline bool parseGfa(TConfig const&, Graph&) { return false; }
template <typename TConfig, typename TRegions, typename TGraph, typename TSR> inline void _findGraphSRBreakpoints(TConfig const&, TRegions const&, TGraph const&, TSR&) {}
template <typename TConfig> inline bool _checkAlternateAlignments(TConfig const&) { return false; }
}  // namespace torali

#define MAX_CN 10
#include "delly.h"
#include "tegua.h"

int main(int argc, char** argv) {
  if (argc < 2) { std::cerr << "usage: delly_ref sr|lr [options]" << std::endl; return 1; }
  const std::string cmd(argv[1]);
  try {
    if (cmd == "sr") return torali::delly(argc - 1, argv + 1);       // src/delly.cpp:66-68
    else if (cmd == "lr") return torali::tegua(argc - 1, argv + 1);  // src/delly.cpp:69-71
  } catch (std::exception const& e) { std::cerr << "error: " << e.what() << std::endl; return 2; }
  std::cerr << "Unrecognized command " << cmd << std::endl;
  return 1;
}
