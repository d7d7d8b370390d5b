// Alignment-phase types shared by align.cu and report.cu (internal).
#pragma once
#include "common.cuh"
#include "mapper.cuh"

namespace damgpu {

constexpr int ALIGN_WARPS = 4;            // warps (jobs in flight) per CTA
#ifndef ALIGN_MINB
#define ALIGN_MINB 5                      // resident CTAs per SM the register allocation aims at
#endif
#ifndef ALIGN_W_V
#define ALIGN_W_V 128
#endif
constexpr int ALIGN_W     = ALIGN_W_V;    // diagonal window in shared memory
constexpr int ALIGN_W_BIG = 8192;         // diagonal window of the overflow kernel (global memory)
#define ALIGN_STATE_BYTES(W) ((size_t) (W) * (2 * 8 + 10 * 4))
constexpr int DUO_WARPS   = 4;            // duo kernel: warps per CTA (two jobs each)
constexpr int DUO_W       = 64;           // duo kernel: diagonal window of a half in wide mode (global memory)
#define LANE_ARENA(tps) (16 * (tps) + 512)   // duo kernel: Pebble cells per job, tps = read length / spacing

// What the waves need of _Align_Spec (align.c:183-191); tables built on the host (align.c:207-269)
struct AlignSpecD
{ int spacing, ave_path;
  const int16_t *score, *table;           // device, 32768 entries each
};

// One candidate chain to extend (a warp's unit of work).  nalign / nwaves / ncells: what the duo kernel added
// to the statistics for this job, so that k_unwind can take it back when it hands the job to the warp kernel.
struct AlignJob { int read, cand, first, count, status; unsigned nalign, nwaves, ncells; };

// One kept local alignment: A path (read vs contig) and B path (contig vs read)
struct AlnRec
{ int       next;                         // next alignment of the same job
  int       comp, bread, pad;
  int       a[6];                         // abpos, bbpos, aepos, bepos, diffs, tlen
  int       b[6];
  long long atrace, btrace;               // offsets into the uint16 trace pool
};

// Duo kernel (align_duo.cu): what k_unwind needs of one wave call, and the calls whose traces
// make up one kept alignment (forward then reverse; a DUB_TRIM re-run stands alone).
struct LaneCall
{ long long cells;                        // first Pebble of the call in the arena
  int       dir, mida, aoff;              // direction, start anti-diagonal, A trace phase
  int       ha, hb;                       // heads of the A and B Pebble chains
  int       x, y, d;                      // end point (a, b) and differences
  int       pad;
};
struct LaneUnwind
{ int      ncalls;                        // -1: record was produced by the warp kernels
  int      acomp, job, pad;
  LaneCall call[2];
};

struct AlignArgs
{ AlignJob        *jobs;
  const int       *job_list;              // optional indirection (overflow re-runs)
  int              njobs;
  int             *job_counter;
  const Candidate *cand;
  const uint32_t  *jumps;
  const uint8_t   *bases_a, *bases_ac, *bases_b;    // reads, reverse-complemented reads, reference
  const uint32_t  *pk_a, *pk_ac, *pk_b;             // the same images 2 bits per base (warp kernel)
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
  // duo kernel
  int             *duo_win;               // wide-band windows, one per half-warp slot (duo_window_bytes)
  void            *lane_cells;            // Pebble arena: per job 8*(rlen/spacing)+256 cells
  const long long *lane_cell_base;        // per read: arena index of its first job
  const int64_t   *lane_job_off;          // per read: index of its first job
  LaneUnwind      *unwind;                // per alignment record (aln_cap)
};

void launch_align(const AlignArgs &A, bool big, int nblocks, cudaStream_t stream);
void launch_align_duo(const AlignArgs &A, int njobs, cudaStream_t stream);
size_t duo_window_bytes(int nblocks);
int  duo_max_blocks();
void launch_unwind(const AlignArgs &A, int max_alns, cudaStream_t stream);

}  // namespace damgpu
