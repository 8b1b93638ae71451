/* dgpu_edlib.h — the edlib C API (src/edlib.h:30-271 of the reference) served by the B200 library, so that call sites that
 * have not been batched yet keep compiling unchanged: include this header instead of edlib.h and
 * link libdelly_b200_host.so. Same type names, field order and constant values as edlib.h; the functions are the dgpu_edlib*
 * entry points below, reached through the edlib names by the macros at the end (macros rather than exported symbols called
 * edlibAlign, so that a process that also links the CPU edlib — the test oracle does — never sees two definitions).
 *
 * One call = one device round trip (a one-job dgpu_edit_distance / dgpu_edit_path_ex on a per-thread context created on first
 * use, device $DGPU_DEVICE or 0): correct but latency-bound. It exists for source compatibility; throughput comes from the
 * batched calls in dgpu.h. There is no CPU fallback: without a usable device every result has status EDLIB_STATUS_ERROR.
 *
 * Differences from edlib, by design: of the optimal end / start locations only the FIRST is reported (numLocations is 1 when
 * an alignment exists) — every call site of the reference reads index 0 only.
 */
#ifndef DGPU_EDLIB_H
#define DGPU_EDLIB_H

#ifdef __cplusplus
extern "C" {
#endif

#define EDLIB_STATUS_OK 0
#define EDLIB_STATUS_ERROR 1

typedef enum { EDLIB_MODE_NW, EDLIB_MODE_SHW, EDLIB_MODE_HW } EdlibAlignMode;          /* global, prefix, infix */
typedef enum { EDLIB_TASK_DISTANCE, EDLIB_TASK_LOC, EDLIB_TASK_PATH } EdlibAlignTask;
typedef enum { EDLIB_CIGAR_STANDARD, EDLIB_CIGAR_EXTENDED } EdlibCigarFormat;           /* M/I/D or =/X/I/D */

#define EDLIB_EDOP_MATCH 0
#define EDLIB_EDOP_INSERT 1
#define EDLIB_EDOP_DELETE 2
#define EDLIB_EDOP_MISMATCH 3

typedef struct {
  char first;
  char second;
} EdlibEqualityPair;

typedef struct {
  int k;                                           /* upper bound on the distance, negative = none */
  EdlibAlignMode mode;
  EdlibAlignTask task;
  const EdlibEqualityPair* additionalEqualities;   /* may be NULL */
  int additionalEqualitiesLength;
} EdlibAlignConfig;

typedef struct {
  int status;
  int editDistance;             /* -1 if larger than k */
  int* endLocations;            /* malloc'd by the callee, released by edlibFreeAlignResult; NULL if no alignment */
  int* startLocations;          /* LOC and PATH tasks only */
  int numLocations;
  unsigned char* alignment;     /* PATH task only: EDLIB_EDOP_* codes */
  int alignmentLength;
  int alphabetLength;           /* distinct bytes over both sequences */
} EdlibAlignResult;

EdlibAlignConfig dgpu_edlibNewAlignConfig(int k, EdlibAlignMode mode, EdlibAlignTask task, const EdlibEqualityPair* additionalEqualities,
                                          int additionalEqualitiesLength);
EdlibAlignConfig dgpu_edlibDefaultAlignConfig(void);
EdlibAlignResult dgpu_edlibAlign(const char* query, int queryLength, const char* target, int targetLength, const EdlibAlignConfig config);
void dgpu_edlibFreeAlignResult(EdlibAlignResult result);
/* malloc'd, NUL-terminated; NULL for an unknown format or an operation code above 3 (src/edlib.cpp:296-345) */
char* dgpu_edlibAlignmentToCigar(const unsigned char* alignment, int alignmentLength, EdlibCigarFormat cigarFormat);

#ifndef DGPU_EDLIB_NO_ALIASES
#define edlibNewAlignConfig dgpu_edlibNewAlignConfig
#define edlibDefaultAlignConfig dgpu_edlibDefaultAlignConfig
#define edlibAlign dgpu_edlibAlign
#define edlibFreeAlignResult dgpu_edlibFreeAlignResult
#define edlibAlignmentToCigar dgpu_edlibAlignmentToCigar
#endif

#ifdef __cplusplus
}
#endif
#endif /* DGPU_EDLIB_H */
