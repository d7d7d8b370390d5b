// Alignment-phase types shared by align.cu and report.cu (internal).
#pragma once
#include "common.cuh"
#include "mapper.cuh"

namespace damgpu {

constexpr int ALIGN_WARPS = 4;            // warps (jobs in flight) per CTA
constexpr int ALIGN_W     = 128;          // diagonal window in shared memory
constexpr int ALIGN_W_BIG = 8192;         // diagonal window of the overflow kernel (global memory)
#define ALIGN_STATE_BYTES(W) ((size_t) (W) * (2 * 8 + 10 * 4))

// What the waves need of _Align_Spec (align.c:183-191); tables built on the host (align.c:207-269)
struct AlignSpecD
{ int spacing, ave_path;
  const int16_t *score, *table;           // device, 32768 entries each
};

// One candidate chain to extend (a warp's unit of work)
struct AlignJob { int read, cand, first, count, status, pad; };

// One kept local alignment: A path (read vs contig) and B path (contig vs read)
struct AlnRec
{ int       next;                         // next alignment of the same job
  int       comp, bread, pad;
  int       a[6];                         // abpos, bbpos, aepos, bepos, diffs, tlen
  int       b[6];
  long long atrace, btrace;               // offsets into the uint16 trace pool
};

struct AlignArgs
{ AlignJob        *jobs;
  const int       *job_list;              // optional indirection (overflow re-runs)
  int              njobs;
  int             *job_counter;
  const Candidate *cand;
  const uint32_t  *jumps;
  const uint8_t   *bases_a, *bases_ac, *bases_b;    // reads, reverse-complemented reads, reference
  const int64_t   *boff_a, *boff_b;
  const int32_t   *rlen_a, *rlen_b;
  AlignSpecD       spec;
  int              kmer, do_b;
  // per-warp scratch
  unsigned char   *big_state;
  void            *cells;
  int              cells_small, cells_big;
  uint16_t        *tscratch;
  int              tcap, tcap_big;
  // outputs
  AlnRec          *alns;   int *aln_top;   int aln_cap;
  uint16_t        *traces; unsigned long long *trace_top; long long trace_cap;
  int             *nfailed;
  unsigned long long *stats;              // nalign, nwaves, ncells, empty-band events
};

void launch_align(const AlignArgs &A, bool big, int nblocks, cudaStream_t stream);

}  // namespace damgpu
