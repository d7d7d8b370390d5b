// Sorted seed hits of one Match_Filter call (internal).
#pragma once
#include <vector>
#include "common.cuh"
#include "index.cuh"

namespace damgpu {

struct SeedSet
{ SeedPair *hits = nullptr;      // nhits records sorted by (aread,bread,apos,bpos) + sentinel
  int64_t   nhits = 0;
  int       limit = 0;           // cap on run products that was applied (map.c:3015)
  std::vector<unsigned long long> histo;   // hitgram[MAXGRAM] (+ total of all products)
};

// A = reads block/index, B = reference block/index (map.c naming)
SeedSet *merge_join(const KmerIndex *aidx, const DeviceBlock *ablock, const KmerIndex *bidx,
                    const DeviceBlock *bblock, int K, uint64_t mem_limit, cudaStream_t stream);
void     free_seeds(SeedSet *ss);
void     join_times(float out[4]);   // ms of the prefix table build and the match kernel of the last call

}  // namespace damgpu
