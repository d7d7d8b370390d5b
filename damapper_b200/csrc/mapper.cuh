// Library-wide parameters and the per-reads-block mapper state (internal).
#pragma once
#include <string>
#include <vector>
#include "common.cuh"
#include "index.cuh"
#include "seeds.cuh"

namespace damgpu {

// Set_Filter_Params state (map.c:114-150) + the globals of map.h:16-23
struct Params
{ int kmer = 0, suppress = 0, nthreads = 1, nshift = 0;
  int verbose = 0, profile = 0, spacing = 100;
  double best_tie = 1.0;
  std::string sort_path = "/tmp";
  uint64_t mem_limit = 0, mem_physical = 0;
};
extern Params g_par;

// Candidate chain (Candidate, map.c:1386-1397); `chain` indexes the jump pool directly and the
// chain's `length` (da,db) pairs are contiguous there (the reference packs them 5 per Jump).
struct Candidate
{ int       next;            // next candidate of the same read (newest first), -1 end, -2 freed
  int       score, length;
  int       bread, comp;
  int       afirst, alast, bfirst, blast;
  int       pad;
  long long chain;
};

// Per-reads-block state that persists from Match_Filter calls to Reporter
// (static Report_Arg *parmr, map.c:1441-1461,2885).
struct Mapper
{ const DeviceBlock *reads = nullptr;
  int       *head = nullptr;                 // per read: DAZZ_READ.coff (map.c:1578,1875)
  Candidate *cand = nullptr;  int *cand_top = nullptr;  int cand_cap = 0;
  uint32_t  *jumps = nullptr; unsigned long long *jump_top = nullptr; uint64_t jump_cap = 0;
  int16_t   *cover = nullptr;                // -p difference array (map.c:1580-1587)
  int64_t   *coff = nullptr;                 // per-read offset into cover
  std::vector<int64_t> h_coff;
  int       *overflow = nullptr;
  int        ctop_bound = 0;                 // host-side upper bounds of *cand_top / *jump_top
  unsigned long long jtop_bound = 0;
  int        spacing = 0;
  int64_t    last_nhits = 0;
  int        last_limit = 0;
};

Mapper *mapper_new(const DeviceBlock *reads);
void    mapper_reset(Mapper *m);
void    mapper_free(Mapper *m);
// chain_thread over all reads for the sorted seeds of one Match_Filter call
// With async the kernel runs on the chain stream, takes ownership of ss->hits and the call returns at
// once; chain_sync() waits for it (every reader of the candidate pools calls it first).
void    chain_seeds(Mapper *m, SeedSet *ss, int bstart, int comp, cudaStream_t stream, bool async);
void    chain_sync();

}  // namespace damgpu
