// Shared declarations of the damgpu CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

namespace damgpu {

// KmerPos / SeedPair of the reference (map.c:78-89), little-endian layout, 16 bytes each.
struct __align__(16) KmerPos  { uint64_t code; int32_t rpos; int32_t read; };
struct __align__(16) SeedPair { int32_t diag; int32_t apos; int32_t bread; int32_t aread; };

// Fatal-error hook: the reference's core calls Clean_Exit(1) (map.h:39); the C-ABI lets the
// host driver install that callback, the default prints and exits.
void fatal(const char *fmt, ...);

#define CUDA_CHECK(x)                                                                   \
  do { cudaError_t e_ = (x);                                                            \
       if (e_ != cudaSuccess)                                                           \
         damgpu::fatal("CUDA error %s at %s:%d (%s)", cudaGetErrorString(e_), __FILE__, \
                       __LINE__, #x);                                                   \
  } while (0)

#define KERNEL_CHECK() CUDA_CHECK(cudaGetLastError())

// Counts kernel launches issued by the library (bench.py reports it as gpu_launches).
extern unsigned long long g_launches;
#define LAUNCH(kernel, grid, block, smem, stream, ...)                \
  do { kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);    \
       damgpu::g_launches += 1;                                       \
       KERNEL_CHECK();                                                \
  } while (0)

// Device memory comes from the stream-ordered pool of the default stream (cudaMallocAsync) with
// an unlimited release threshold: after the first step every buffer is recycled inside the
// process instead of being mapped/unmapped by the driver (GB-sized cudaMalloc/cudaFree calls
// cost more host time than the kernels that use them).
template <typename T> static inline T *dalloc(size_t n)
{ T *p = nullptr;
  if (n == 0) n = 1;
  CUDA_CHECK(cudaMallocAsync((void **) &p, n * sizeof(T), 0));
  return p;
}

static inline void dfree(void *p) { if (p) CUDA_CHECK(cudaFreeAsync(p, 0)); }

int sm_count();

// ---- radix_sort.cu -------------------------------------------------------------------
// Stable LSD radix sort of n 16-byte records on the key bytes listed in `bytes` (least
// significant first), 8-bit digits, one read + one write of the array per pass
// (lex_sort/lex_thread, map.c:181-444).  `hist` is a device array [npass][256] holding the
// digit histograms of every pass (filled by the producer kernel or by radix_histogram).
// Sorts ping-pong between a and b and returns the buffer holding the result.
void *radix_sort16(void *a, void *b, uint32_t n, const int *bytes, int npass, uint32_t *hist,
                   cudaStream_t stream);
// histogram of all pass bytes in one read of the array
void radix_histogram(const void *recs, uint32_t n, const int *bytes, int npass, uint32_t *hist,
                     cudaStream_t stream);
// event-timed duration of the radix passes of the last radix_sort16 call when timing is on
extern bool   g_time_kernels;

}  // namespace damgpu
