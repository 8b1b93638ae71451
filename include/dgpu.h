/*
 * dgpu.h — C ABI of libdelly_b200: batched, B200-native (sm_100a) replacements for the
 * per-item alignment calls on Delly's split-read / consensus realignment path.
 *
 * The reference (dellytools/delly @ 3a22fe2) has no FFI; it is one translation unit of
 * inline templates. Each entry point below names the reference call site(s) it replaces
 * (paths relative to the reference root). A maintainer binds them by replacing the
 * per-item loop at that call site with one batched call.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes only. Return 0 (DGPU_OK) or a negative
 *    DGPU_ERR_* code; no exceptions cross the boundary.
 *  - Sequences are 1 byte per base in one packed arena (`seqs`); equality is byte
 *    equality exactly like edlib / the reference's char compares (so IUPAC codes,
 *    'N' and lower case keep the reference's semantics).
 *  - Every batch call exists in two forms:
 *      dgpu_<op>        host pointers in/out; the call stages H2D, launches, copies the
 *                       results D2H and returns when they are in the caller's buffers.
 *                       Pinned caller buffers are DMA'd directly.
 *      dgpu_<op>_dev    device pointers in/out, enqueued on `stream`
 *                       (a cudaStream_t passed as void*; NULL = the context's stream).
 *  - Caller buffers. Host forms check every job against the sizes given (sequences inside `seqs_bytes`, output slots inside `ops_bytes` /
 *    `aln_bytes` / `cons_bytes`) and return DGPU_ERR_ARG / DGPU_ERR_CAPACITY before anything is launched. Device forms trust their arguments
 *    (checking would need a device round trip); their arena must be readable in whole 16-byte words: 16-byte aligned, with >= 16 readable bytes
 *    after `seqs_bytes` (the kernels fetch the arena with aligned 128-bit loads).
 *  - A context is bound to one CUDA device and must not be shared between host threads
 *    (same contract as one edlib call per pool worker, src/coverage.h:420-426).
 *  - There is NO CPU fallback: if no CUDA device is usable, dgpu_ctx_create fails.
 */
#ifndef DGPU_H
#define DGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DGPU_OK 0
#define DGPU_ERR_CUDA (-1)        /* a CUDA runtime call failed; see dgpu_last_error() */
#define DGPU_ERR_ARG (-2)         /* invalid argument */
#define DGPU_ERR_NODEVICE (-3)    /* no usable sm_100 device */
#define DGPU_ERR_CAPACITY (-4)    /* an output/workspace capacity was exceeded */
#define DGPU_ERR_UNSUPPORTED (-5) /* shape outside what the device path supports */
#define DGPU_ERR_NCCL (-6)

/* Alignment modes and tasks: numerically identical to EdlibAlignMode / EdlibAlignTask
 * (src/edlib.h:36-71) so call sites can pass their existing enum values. */
#define DGPU_MODE_NW 0
#define DGPU_MODE_SHW 1
#define DGPU_MODE_HW 2

typedef struct dgpu_ctx dgpu_ctx;

/* ---- context ------------------------------------------------------------------ */
int dgpu_ctx_create(int device, dgpu_ctx** out);
void dgpu_ctx_destroy(dgpu_ctx* ctx);
/* Block until everything queued on the context's stream has finished. */
int dgpu_ctx_sync(dgpu_ctx* ctx);
/* Text of the last CUDA/NCCL error seen by this context (never NULL). */
const char* dgpu_last_error(dgpu_ctx* ctx);
const char* dgpu_strerror(int code);
/* ABI version (major*1000+minor). */
int dgpu_version(void);
/* Number of kernel launches issued through this context so far (bench accounting). */
uint64_t dgpu_launch_count(dgpu_ctx* ctx);
/* jobs refused PER ITEM so far because they exceed a device limit (the call still succeeds; the item comes back flagged: ok = 0 / status != 0) */
uint64_t dgpu_unsupported_count(dgpu_ctx* ctx);
/* Stream-asynchronous device forms. By default dgpu_edit_distance_dev reads its job-class counts back once per call (one small D2H + stream
 * synchronisation) so that it launches exactly the kernels the batch needs. With max_seq_len > 0 the caller guarantees that no query or target of
 * later calls is longer; the call then enqueues everything on `stream` and returns without any host synchronisation (scratch buffers are sized from
 * the bound and grown only between calls; it can be followed by further work on the stream, overlapped with copies on other streams, or captured
 * into a CUDA graph once the scratch exists). 0 switches back. dgpu_long_needle_dev / dgpu_msa_dev / dgpu_edit_path_dev always synchronise once:
 * their workspace depends on the shapes in the batch. */
int dgpu_set_async_bound(dgpu_ctx* ctx, uint32_t max_seq_len);
/* Measurement hooks (bench.py): when profiling is on, every batch call brackets its dominant
 * kernels (not the staging copies or the binning pre-pass) with CUDA events on the launching
 * stream; dgpu_last_kernel_ms waits for and returns that span for the most recent call. */
int dgpu_set_profiling(dgpu_ctx* ctx, int on);
float dgpu_last_kernel_ms(dgpu_ctx* ctx);
/* Measured integer-ALU peak of this device: a LOP3/IADD3 micro-kernel, result in 1e12 int32 op/s.
 * MEASURED_PEAKS.json carries no integer peak, and this path is integer-pipe bound. */
int dgpu_int_peak(dgpu_ctx* ctx, double* tera_ops_per_s);

/* ---- Myers bit-vector edit distance (replaces edlibAlign with EDLIB_TASK_DISTANCE) --
 * Call sites: _editDistanceHW, src/coverage.h:107-115 (two per AlignJob in process_batch,
 * :412-441, and the dump path :520-521); _editDistanceNW, src/genotype.h:22-29
 * (:276,:284); orientation check src/split.h:567-568; msaEdlib pairs src/assemble.h:390.
 *
 * Job i aligns query seqs[q_off[i] .. +q_len[i]) against target seqs[t_off[i] .. +t_len[i])
 * in `mode` with bound k[i] (k < 0: unbounded, like edlib's auto-doubling k,
 * src/edlib.cpp:192-210). dist[i] = edit distance, or -1 if it exceeds k — bit-identical to
 * EdlibAlignResult.editDistance, including the HW clamp k=min(k,|q|) (src/edlib.cpp:563-565),
 * the NW rules (src/edlib.cpp:740-746) and the empty-sequence rules (src/edlib.cpp:158-177).
 * end_loc (may be NULL): endLocations[0] (leftmost optimal end, src/edlib.cpp:656-691), -1 if none.
 * `k` may be NULL (all unbounded). Offsets are 32-bit: one arena is at most 4 GiB.
 */
int dgpu_edit_distance(dgpu_ctx* ctx, const uint8_t* seqs, uint64_t seqs_bytes,
                       const uint32_t* q_off, const uint32_t* q_len,
                       const uint32_t* t_off, const uint32_t* t_len,
                       const int32_t* k, int mode, uint64_t n,
                       int32_t* dist, int32_t* end_loc);
int dgpu_edit_distance_dev(dgpu_ctx* ctx, const uint8_t* seqs, uint64_t seqs_bytes,
                           const uint32_t* q_off, const uint32_t* q_len,
                           const uint32_t* t_off, const uint32_t* t_len,
                           const int32_t* k, int mode, uint64_t n,
                           int32_t* dist, int32_t* end_loc, void* stream);

/* ---- edit distance + locations + alignment path (replaces edlibAlign with EDLIB_TASK_PATH) --
 * Call sites (all with k = -1): splitAlign, src/split.h:485,489,497,507,524,527 (HW and SHW);
 * _trimConsensus src/assemble.h:351,356; msaEdlib/msaWfa src/assemble.h:447,656,693 (NW/HW).
 * Per job: dist (editDistance), start_loc / end_loc (startLocations[0] / endLocations[0], src/edlib.cpp:213-258),
 * ops = the alignment array (0 match, 1 insert, 2 delete, 3 mismatch; src/edlib.h:84-87) at ops + ops_off[i]
 * (reserve q_len + t_len bytes), ops_len (alignmentLength). Problems above edlib's 1 MiB switch are split with
 * Hirschberg's recursion exactly like the reference (src/edlib.cpp:1189-1212, :1232-1397), so paths are identical
 * at every size. status: 0 ok; 2 = aligned target slice longer than 16384 (not supported, path NOT produced);
 * 3 = internal consistency failure.
 * The _ex forms take edlib's additionalEqualities (src/edlib.h:100-106) as 2*n_eq bytes (first,second pairs,
 * n_eq <= 32) in every mode — the reference's call sites are src/assemble.h:425-447 (NW, msaEdlib) and
 * src/assemble.h:663-693 (HW, msaWfa).
 * The *_dev forms stage job geometry through the host between device rounds (they synchronise the stream).
 */
int dgpu_edit_path(dgpu_ctx* ctx, const uint8_t* seqs, uint64_t seqs_bytes,
                   const uint32_t* q_off, const uint32_t* q_len,
                   const uint32_t* t_off, const uint32_t* t_len, int mode, uint64_t n,
                   int32_t* dist, int32_t* start_loc, int32_t* end_loc,
                   uint8_t* ops, const uint64_t* ops_off, uint64_t ops_bytes,
                   uint32_t* ops_len, uint32_t* status);
int dgpu_edit_path_dev(dgpu_ctx* ctx, const uint8_t* seqs, uint64_t seqs_bytes,
                       const uint32_t* q_off, const uint32_t* q_len,
                       const uint32_t* t_off, const uint32_t* t_len, int mode, uint64_t n,
                       int32_t* dist, int32_t* start_loc, int32_t* end_loc,
                       uint8_t* ops, const uint64_t* ops_off,
                       uint32_t* ops_len, uint32_t* status, void* stream);

int dgpu_edit_path_ex(dgpu_ctx* ctx, const uint8_t* seqs, uint64_t seqs_bytes,
                      const uint32_t* q_off, const uint32_t* q_len,
                      const uint32_t* t_off, const uint32_t* t_len, int mode,
                      const uint8_t* eq_pairs, uint32_t n_eq, uint64_t n,
                      int32_t* dist, int32_t* start_loc, int32_t* end_loc,
                      uint8_t* ops, const uint64_t* ops_off, uint64_t ops_bytes,
                      uint32_t* ops_len, uint32_t* status);
int dgpu_edit_path_ex_dev(dgpu_ctx* ctx, const uint8_t* seqs, uint64_t seqs_bytes,
                          const uint32_t* q_off, const uint32_t* q_len,
                          const uint32_t* t_off, const uint32_t* t_len, int mode,
                          const uint8_t* eq_pairs, uint32_t n_eq, uint64_t n,
                          int32_t* dist, int32_t* start_loc, int32_t* end_loc,
                          uint8_t* ops, const uint64_t* ops_off,
                          uint32_t* ops_len, uint32_t* status, void* stream);

/* ---- consensus vs SV-reference split alignment (replaces longNeedle) --------------------
 * Call sites: _consRefAlignment for svt != 4, src/split.h:555 (reached from alignConsensus,
 * src/shortpe.h:186,253, src/assemble.h:849,859,916,926) and _generateProbes, src/coverage.h:214.
 * Implements longNeedle(cons, svRef, align, AlignConfig<true,false>, DnaScore(1,-1,-1,-1)),
 * src/needle.h:45-222, bit-for-bit (DP, forward/reverse score check, best-join tie rules,
 * traceback priority, stitched alignment).
 *
 * Job i aligns consensus seqs[c_off[i]..+c_len[i]) (rows) against the SV reference window
 * seqs[r_off[i]..+r_len[i]) (columns). Outputs per job:
 *   ok[i]       1 = longNeedle returned true; 0 = false (scores disagree / no split found)
 *   aln_len[i]  alignment columns L (0 when ok == 0)
 *   aln         row 0 (consensus) at aln + aln_off[i], row 1 (reference) at aln + aln_off[i] +
 *               (c_len[i] + r_len[i]); the caller reserves 2*(c_len+r_len) bytes per job.
 *   info        optional [consLeft, refLeft, refRight, bestScore] per job (white-box tests)
 * Limits: c_len + r_len <= 32000 (scores are stored as int16); otherwise DGPU_ERR_UNSUPPORTED.
 */
int dgpu_long_needle(dgpu_ctx* ctx, const uint8_t* seqs, uint64_t seqs_bytes,
                     const uint32_t* c_off, const uint32_t* c_len,
                     const uint32_t* r_off, const uint32_t* r_len, uint64_t n,
                     uint8_t* aln, const uint64_t* aln_off, uint64_t aln_bytes,
                     uint32_t* aln_len, uint8_t* ok, int32_t* info);
int dgpu_long_needle_dev(dgpu_ctx* ctx, const uint8_t* seqs, uint64_t seqs_bytes,
                         const uint32_t* c_off, const uint32_t* c_len,
                         const uint32_t* r_off, const uint32_t* r_len, uint64_t n,
                         uint8_t* aln, const uint64_t* aln_off,
                         uint32_t* aln_len, uint8_t* ok, int32_t* info, void* stream);

/* ---- per-cluster MSA + consensus (replaces msa) -----------------------------------------
 * Call sites: msa(c, seqStore[svid], consensus), src/shortpe.h:185 and :252 (one call per SV
 * inside the thread-pool workers of assembleSplitReads). Implements src/msa.h:185-239 bit-for-bit:
 * LCS similarity matrix, UPGMA guide tree, progressive end-gap-free affine gotoh() on column
 * profiles (src/gotoh.h:71-174, src/align.h:96-171) and the coverage-thresholded consensus vote.
 *
 * Cluster i consists of reads cluster_off[i] .. cluster_off[i+1]-1 (indices into read_off/read_len),
 * IN THE ORDER THE REFERENCE ITERATES ITS std::unordered_set (the caller preserves that order).
 * Outputs per cluster: consensus bytes at cons + cons_off[i] (reserve sum of the cluster's read
 * lengths), cons_len[i], n_rows[i] (= msa()'s return value), status[i]:
 *   0 ok; 1 more than 32 reads; 2 a read or an intermediate alignment exceeds 1023 columns / the
 *   per-cluster workspace; 3 a read contains a byte outside ACGTN (the reference's float profile is
 *   NaN there) — such clusters are NOT computed and the caller must treat them as errors.
 * Optional white-box output: the root alignment (n_rows x aln_cols, row-major) at aln + aln_off[i].
 * Scoring: (match, mismatch, go, ge) = DnaScore, e.g. (5,-4,-10,-1); requires go <= 0 and ge <= 0.
 */
int dgpu_msa(dgpu_ctx* ctx, const uint8_t* seqs, uint64_t seqs_bytes,
             const uint32_t* read_off, const uint32_t* read_len, uint32_t nreads,
             const uint32_t* cluster_off, uint32_t nclusters,
             int match, int mismatch, int go, int ge, int min_clique,
             uint8_t* cons, const uint64_t* cons_off, uint64_t cons_bytes,
             uint32_t* cons_len, uint32_t* n_rows, uint32_t* status,
             uint8_t* aln, const uint64_t* aln_off, uint64_t aln_bytes, uint32_t* aln_cols);
int dgpu_msa_dev(dgpu_ctx* ctx, const uint8_t* seqs, uint64_t seqs_bytes,
                 const uint32_t* read_off, const uint32_t* read_len,
                 const uint32_t* cluster_off, uint32_t nclusters,
                 int match, int mismatch, int go, int ge, int min_clique,
                 uint8_t* cons, const uint64_t* cons_off,
                 uint32_t* cons_len, uint32_t* n_rows, uint32_t* status,
                 uint8_t* aln, const uint64_t* aln_off, uint32_t* aln_cols, void* stream);

/* ---- clustering graphs: candidate edges ------------------------------------------
 * Replaces the windowed pair scans inside cluster() — src/cluster.h:371-431 (SRBamRecord) and :551-623
 * (BamAlignRecord): which pairs (i, j > i) of the SORTED records connect, and the edge weight. Records are passed
 * column-wise (int32). edge_off[n+1] receives the CSR offsets (edges of record i at edge_off[i] .. edge_off[i+1],
 * in increasing j — the order the reference's double loop meets them), edge_j / edge_w the targets and weights.
 * *n_edges = total number of edges; if it exceeds edge_cap the call fills nothing beyond edge_cap and returns
 * DGPU_ERR_CAPACITY (call again with larger arrays; edge_cap = 0 just counts). The component bookkeeping,
 * graphPruning and clique growth that consume the edges are sequential and stay with the caller
 * (delly_b200/host/cluster.hpp: clusterGpu).
 *   SR: chr, pos, chr2, pos2, inslen of SRBamRecord (src/tags.h:62-80), svt 0..8, max_read_sep = c.maxReadSep.
 *   PE: pos, mpos, mtid, alen, Median, maxNormalISize of BamAlignRecord (src/cluster.h:24-50), svt 0..3 / 5..8,
 *       varisize as passed to cluster() (src/shortpe.h:512-515). */
int dgpu_cluster_edges_sr(dgpu_ctx* ctx, const int32_t* chr, const int32_t* pos, const int32_t* chr2, const int32_t* pos2,
                          const int32_t* inslen, uint64_t n, int svt, uint32_t max_read_sep,
                          uint32_t* edge_off, uint32_t* edge_j, uint32_t* edge_w, uint64_t edge_cap, uint64_t* n_edges);
int dgpu_cluster_edges_pe(dgpu_ctx* ctx, const int32_t* pos, const int32_t* mpos, const int32_t* mtid, const int32_t* alen,
                          const int32_t* median, const int32_t* max_normal_isize, uint64_t n, int svt, uint32_t varisize,
                          uint32_t* edge_off, uint32_t* edge_j, uint32_t* edge_w, uint64_t edge_cap, uint64_t* n_edges);

/* ---- multi-GPU: the one exchange step ----------------------------------------------
 * The reference is one process: it builds the SV list, sorts and renumbers it (src/delly.h:155-158, src/tegua.h:149-156) and writes
 * the BCF. Sharded over GPUs (one process per GPU; contiguous ranges of the sorted SV list per rank, delly_b200/host/gather.hpp) every
 * rank finishes the records of its own range; this is the exchange that brings them together before emission: an all-gatherv of one
 * serialised byte string per rank over NCCL (ncclAllGather of the byte counts, then a grouped ncclBroadcast of exactly each payload).
 *   dgpu_comm_unique_id / dgpu_comm_init / dgpu_comm_destroy  thin wrappers of ncclGetUniqueId / ncclCommInitRank / ncclCommDestroy
 *       for callers that have no communicator yet (rank 0 creates the 128-byte id and hands it to the other ranks out of band). NCCL is
 *       bound at run time (dlopen libnccl.so.2): without it these calls return DGPU_ERR_NCCL and the rest of the library is unaffected.
 *   dgpu_gather_records  local / local_bytes = this rank's payload (host memory); *all = the payloads of ranks 0..n-1 back to back and
 *       *counts = their n sizes, both malloc'd by the callee (release with dgpu_free_host); *nranks = n. comm == NULL: single process
 *       (the output is a copy of the input). Collective: every rank of `comm` must call it, with its own context. */
int dgpu_comm_unique_id(uint8_t* id128);
int dgpu_comm_init(dgpu_ctx* ctx, int nranks, int rank, const uint8_t* id128, void** comm);
int dgpu_comm_destroy(dgpu_ctx* ctx, void* comm);
int dgpu_nccl_version(void);   /* ncclGetVersion of the bound library, 0 if NCCL is not available */
int dgpu_gather_records(dgpu_ctx* ctx, void* comm /* ncclComm_t */, const void* local, uint64_t local_bytes, void** all, uint64_t** counts, int* nranks);
void dgpu_free_host(void* p);

#ifdef __cplusplus
}
#endif
#endif /* DGPU_H */
