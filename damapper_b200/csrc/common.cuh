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
extern bool g_debug_sync;          // DAMGPU_DEBUG_SYNC=1: synchronise and check after every launch
#define LAUNCH(kernel, grid, block, smem, stream, ...)                \
  do { kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);    \
       damgpu::g_launches += 1;                                       \
       KERNEL_CHECK();                                                \
       if (damgpu::g_debug_sync)                                      \
         { cudaError_t e_ = cudaDeviceSynchronize();                  \
           if (e_ != cudaSuccess)                                     \
             damgpu::fatal("kernel %s failed: %s (%s:%d)", #kernel, cudaGetErrorString(e_), __FILE__, __LINE__); \
         }                                                            \
  } while (0)

// Device memory comes from a caching allocator owned by the library (capi.cu): freed blocks go
// to size-keyed free lists and are handed out again for requests of (almost) the same size, so
// after the first step no GB-sized cudaMalloc/cudaFree reaches the driver (those calls cost more
// host time than the kernels using the buffers).  All work runs on one stream, so reuse is
// stream-ordered.
void *cache_alloc(size_t bytes);
void  cache_free(void *p);

template <typename T> static inline T *dalloc(size_t n)
{ if (n == 0) n = 1;
  return static_cast<T *>(cache_alloc(n * sizeof(T)));
}

static inline void dfree(void *p) { if (p) cache_free(p); }

int sm_count();

// DAMGPU_TRACE=1: wall-clock per named phase (device synchronised at every mark), to stderr
void trace_mark(const char *name);
#define TRACE(name) do { if (damgpu::g_trace) damgpu::trace_mark(name); } while (0)
extern bool g_trace;
// first tier of the alignment phase: 1 = two jobs per warp (align_duo.cu, default), 0 = warp per
// job (align.cu, also the tier that re-runs what outgrows the first).  DAMGPU_ALIGN=warp|duo.
extern int g_align_tier;
extern bool g_chain_async;         // chain kernel on its own stream (DAMGPU_SYNC_CHAIN=1 turns it off)

// ---- radix_sort.cu -------------------------------------------------------------------
// Stable LSD radix sort of n 16-byte records on the key bytes listed in `bytes` (least
// significant first), 8-bit digits, one read + one write of the array per pass
// (lex_sort/lex_thread, map.c:181-444).  `hist` is a device array [npass][256] holding the
// digit histograms of every pass (filled by the producer kernel or by radix_histogram).
// Sorts ping-pong between a and b and returns the buffer holding the result.
void *radix_sort16(void *a, void *b, uint32_t n, const int *bytes, int npass, uint32_t *hist,
                   cudaStream_t stream);
// bytes / ms / pass launches / calls of the sorts since the last reset (damgpu_time_kernels on)
void radix_totals(double out[4], int reset);
// histogram of all pass bytes in one read of the array
void radix_histogram(const void *recs, uint32_t n, const int *bytes, int npass, uint32_t *hist,
                     cudaStream_t stream);
// event-timed duration of the radix passes of the last radix_sort16 call when timing is on
extern bool   g_time_kernels;
extern int    g_radix_pf;
extern bool   g_radix_reload;      // radix pass variant (radix_sort.cu)

}  // namespace damgpu
