// Reporter output (internal): canonical record streams as they go into the .las files.
#pragma once
#include <vector>
#include "common.cuh"
#include "mapper.cuh"

namespace damgpu {

// Host buffers the record streams are copied into: page-locked memory from a small pool (capi.cu), so
// the device-to-host copy of a few MB runs at PCIe speed instead of through the driver's bounce buffers.
void *pinned_get(size_t bytes);
void  pinned_put(void *p);
template <class T> struct PinnedAlloc
{ typedef T value_type;
  PinnedAlloc() {}
  template <class U> PinnedAlloc(const PinnedAlloc<U> &) {}
  T   *allocate(size_t n) { return static_cast<T *>(pinned_get(n * sizeof(T))); }
  void deallocate(T *p, size_t) { pinned_put(p); }
  template <class U> bool operator==(const PinnedAlloc<U> &) const { return true; }
  template <class U> bool operator!=(const PinnedAlloc<U> &) const { return false; }
};
typedef std::vector<uint8_t, PinnedAlloc<uint8_t>> ByteVec;

struct ReportOut
{ ByteVec a, b;                            // 40-byte records (padding zeroed) + trace bytes
  std::vector<int64_t> read_off_a, read_off_b;   // per read byte offsets into a / b (nreads+1)
  std::vector<int>     read_nrec_a, read_nrec_b; // per read record counts
  int64_t nrec_a = 0, nrec_b = 0;
  ByteVec prof;                            // -p track bytes, (rlen-1)/S+2 per read
  int64_t nalign = 0, nwaves = 0, ncells = 0, empty_band = 0, h2_events = 0;
  int64_t trace_fails = 0;                 // records failing Check_Trace_Points (align.c:3194), checked on the device
  int     overflow_jobs = 0;               // alignment jobs re-run by the overflow kernel
  float   ms_align = 0.f;
};

void build_align_spec(double ave_corr, const float freq[4], int *ave_path, int16_t *score,
                      int16_t *table);
ReportOut *reporter(Mapper *m, const DeviceBlock *ref, double ave_corr, const float freq[4],
                    int do_a, int do_b, cudaStream_t stream);

}  // namespace damgpu
