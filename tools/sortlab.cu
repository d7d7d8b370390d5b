// Dev harness (not product): variants of the onesweep radix pass over 16-byte records, timed with
// CUDA events on random data of the C2 reads size and checked on the device for digit order +
// stability + multiset equality.   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr uint32_t FLAG_AGG = 1u << 30, FLAG_INC = 2u << 30, VAL_MASK = (1u << 30) - 1;

__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t *p)
{ uint32_t v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_relaxed(uint32_t *p, uint32_t v)
{ asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }

// digit = bits [shift, shift+BITS) of 64-bit word `word` (0: x,y  1: z,w)
template <int BITS>
__device__ __forceinline__ uint32_t rec_digit(const uint4 &r, int word, int shift)
{ const uint32_t lo = word ? r.z : r.x, hi = word ? r.w : r.y;
  const uint32_t v = (shift < 32) ? __funnelshift_r(lo, hi, shift) : (hi >> (shift - 32));
  return v & ((1u << BITS) - 1);
}

// ---------------- data + check kernels ----------------
__device__ __forceinline__ uint64_t mix(uint64_t x)
{ x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; }

__global__ void k_fill(uint4 *a, uint32_t n, int keybits, uint64_t seed)
{ for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    { uint64_t c = mix(i + seed) & ((keybits == 64) ? ~0ull : ((1ull << keybits) - 1));
      a[i] = make_uint4((uint32_t) c, (uint32_t) (c >> 32), i, 0x5a5a0000u ^ (i >> 7));
    }
}

__global__ void k_sum(const uint4 *a, uint32_t n, unsigned long long *sum)
{ unsigned long long s = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    { uint4 r = a[i]; s += mix(((uint64_t) r.x | ((uint64_t) r.y << 32)) ^ mix(((uint64_t) r.z << 32) | r.w)); }
  atomicAdd(sum, s);
}

// after LSD passes over key bits [0, donebits): order must be (key & mask, original index) strictly increasing
__global__ void k_check(const uint4 *a, uint32_t n, int donebits, unsigned long long *bad)
{ const uint64_t mask = (donebits >= 64) ? ~0ull : ((1ull << donebits) - 1);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i + 1 < n; i += gridDim.x * blockDim.x)
    { uint4 p = a[i], q = a[i + 1];
      uint64_t kp = ((uint64_t) p.x | ((uint64_t) p.y << 32)) & mask, kq = ((uint64_t) q.x | ((uint64_t) q.y << 32)) & mask;
      if (kp > kq || (kp == kq && p.z >= q.z)) atomicAdd(bad, 1ull);
    }
}

__global__ void k_copy(const uint4 *__restrict__ in, uint4 *__restrict__ out, uint32_t n)
{ for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    __stcs(out + i, __ldcs(in + i));
}

template <int BITS>
__global__ void k_hist(const uint4 *__restrict__ a, uint32_t n, int word, int shift, uint32_t *hist)
{ __shared__ uint32_t sh[1 << BITS];
  for (int i = threadIdx.x; i < (1 << BITS); i += blockDim.x) sh[i] = 0;
  __syncthreads();
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    atomicAdd(&sh[rec_digit<BITS>(a[i], word, shift)], 1u);
  __syncthreads();
  for (int i = threadIdx.x; i < (1 << BITS); i += blockDim.x) if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

__global__ void k_exscan(uint32_t *h, int nb)      // single thread, tiny
{ uint32_t run = 0; for (int i = 0; i < nb; i++) { uint32_t c = h[i]; h[i] = run; run += c; } }

// ---------------- V0: the committed kernel (256x8, 8-bit), with ablation switches ----------------
// ABL bit0: skip look-back wait   bit1: skip ranking (identity positions)
template <int ABL>
__global__ void __launch_bounds__(256)
k_pass_v0(const uint4 *__restrict__ in, uint4 *__restrict__ out, uint32_t n, int word, int shift,
          const uint32_t *__restrict__ gbase, uint32_t *tile_state, uint32_t *tile_counter)
{ constexpr int RS_THREADS = 256, RS_ITEMS = 8, RS_TILE = 2048, RS_WARPS = 8;
  __shared__ uint4    stage[RS_TILE];
  __shared__ uint32_t s_dbase[256];
  __shared__ uint32_t s_delta[256];
  __shared__ uint32_t s_wsum[RS_WARPS];
  __shared__ uint32_t s_tile;
  uint32_t (*whist)[256] = reinterpret_cast<uint32_t (*)[256]>(stage);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
  for (int i = lane; i < 256; i += 32) whist[warp][i] = 0;
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint32_t base = tile * (uint32_t) RS_TILE + warp * (32 * RS_ITEMS) + lane;
  const uint32_t nvalid = (n - tile * (uint32_t) RS_TILE < (uint32_t) RS_TILE) ? n - tile * (uint32_t) RS_TILE : (uint32_t) RS_TILE;
  uint4 rec[RS_ITEMS]; uint32_t dig[RS_ITEMS], rank[RS_ITEMS];
#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++)
    { uint32_t idx = base + i * 32;
      if (idx < n) { rec[i] = __ldcs(in + idx); dig[i] = rec_digit<8>(rec[i], word, shift); }
      else { rec[i] = make_uint4(0, 0, 0, 0); dig[i] = 255; }
    }
  const uint32_t lt = (1u << lane) - 1;
  if (ABL & 2)
    {
#pragma unroll
      for (int i = 0; i < RS_ITEMS; i++) { rank[i] = 0; if (lane == 0) whist[warp][dig[i]] += 1; }
    }
  else
    {
#pragma unroll
      for (int i = 0; i < RS_ITEMS; i++)
        { uint32_t peers = __match_any_sync(0xffffffffu, dig[i]);
          uint32_t prev  = whist[warp][dig[i]];
          __syncwarp();
          if ((peers & lt) == 0) whist[warp][dig[i]] = prev + __popc(peers);
          __syncwarp();
          rank[i] = prev + __popc(peers & lt);
        }
    }
  __syncthreads();
  { const int d = tid;
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; w++) { uint32_t c = whist[w][d]; whist[w][d] = run; run += c; }
    if (d == 255) run -= (uint32_t) RS_TILE - nvalid;
    uint32_t *st = tile_state + (size_t) tile * 256 + d;
    st_relaxed(st, (tile == 0 ? FLAG_INC : FLAG_AGG) | run);
    uint32_t x = run;
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_wsum[warp] = x;
    __syncthreads();
    uint32_t add = 0;
    for (int w = 0; w < warp; w++) add += s_wsum[w];
    const uint32_t dbase = x + add - run;
    uint32_t excl = 0;
    if (tile > 0 && !(ABL & 1))
      { const uint32_t *p = st - 256;
        while (true)
          { uint32_t v = ld_relaxed(p);
            if (v & FLAG_INC) { excl += v & VAL_MASK; break; }
            if (v & FLAG_AGG) { excl += v & VAL_MASK; p -= 256; continue; }
            __nanosleep(20);
          }
        st_relaxed(st, FLAG_INC | (excl + run));
      }
    if (ABL & 1) excl = tile * 8;
    s_dbase[d] = dbase;
    s_delta[d] = gbase[d] + excl - dbase;
  }
  __syncthreads();
  uint32_t pos[RS_ITEMS];
#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++)
    pos[i] = (ABL & 2) ? (uint32_t) (warp * 256 + i * 32 + lane) : s_dbase[dig[i]] + whist[warp][dig[i]] + rank[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++) stage[pos[i]] = rec[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++)
    { uint32_t j = tid + i * RS_THREADS;
      if (j < nvalid)
        { uint4 r = stage[j];
          uint32_t o = (ABL & 2) ? tile * RS_TILE + j : j + s_delta[rec_digit<8>(r, word, shift)];
          if (ABL & 1) o = (o < n) ? o : (o % n);
          __stcs(out + o, r);
        }
    }
}

// ---------------- V1: early counts ----------------
// 1. per-warp digit histogram with plain shared atomics  2. sync; thread-per-digit scan over warps,
// publish the tile aggregate at once  3. match-based stable ranking with the per-warp counters now
// holding exclusive offsets (leader atomicAdd + shuffle) giving final staged positions
// 4. stage  5. look-back (predecessors have had the whole ranking phase to publish)  6. write.
template <int THREADS, int ITEMS, int BITS, bool ALIAS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_pass_v1(const uint4 *__restrict__ in, uint4 *__restrict__ out, uint32_t n, int word, int shift,
          const uint32_t *__restrict__ gbase, uint32_t *tile_state, uint32_t *tile_counter)
{ constexpr int NB = 1 << BITS, WARPS = THREADS / 32, TILE = THREADS * ITEMS;
  constexpr int DPT = (NB + THREADS - 1) / THREADS;        // digits owned per thread (blocked)
  extern __shared__ uint4 dyn[];
  uint4    *stage = dyn;
  uint32_t *whist = ALIAS ? reinterpret_cast<uint32_t *>(dyn) : reinterpret_cast<uint32_t *>(dyn + TILE);
  __shared__ uint32_t s_dbase[NB];
  __shared__ uint32_t s_delta[NB];
  __shared__ uint32_t s_wsum[WARPS];
  __shared__ uint32_t s_tile;
  static_assert(!ALIAS || (size_t) WARPS * NB * 4 <= (size_t) TILE * 16, "alias");

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lt = (1u << lane) - 1;
  if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
  uint32_t *wh = whist + warp * NB;
#pragma unroll
  for (int i = lane; i < NB; i += 32) wh[i] = 0;
  __syncthreads();
  const uint32_t tile  = s_tile;
  const uint32_t tbase = tile * (uint32_t) TILE;
  const uint32_t nvalid = (n - tbase < (uint32_t) TILE) ? n - tbase : (uint32_t) TILE;
  const uint32_t base = tbase + warp * (32 * ITEMS) + lane;

  uint4 rec[ITEMS]; uint32_t dig[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
    { const uint32_t idx = base + i * 32;
      rec[i] = (idx < n) ? __ldcs(in + idx) : make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    }
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
    { dig[i] = rec_digit<BITS>(rec[i], word, shift);       // padding: all ones -> digit NB-1, ranks last
      atomicAdd(&wh[dig[i]], 1u);
    }
  __syncthreads();                                         // (A)

  uint32_t cnt[DPT], dbase[DPT];
  { uint32_t tsum = 0;
#pragma unroll
    for (int j = 0; j < DPT; j++)
      { const int d = tid * DPT + j;
        uint32_t run = 0;
        if (d < NB)
          {
#pragma unroll
            for (int w = 0; w < WARPS; w++) { uint32_t c = whist[w * NB + d]; whist[w * NB + d] = run; run += c; }
            if (d == NB - 1) run -= (uint32_t) TILE - nvalid;
            st_relaxed(tile_state + (size_t) tile * NB + d, (tile == 0 ? FLAG_INC : FLAG_AGG) | run);
          }
        cnt[j] = run; tsum += run;
      }
    uint32_t x = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_wsum[warp] = x;
    __syncthreads();                                       // (B1)
    uint32_t add = 0;
#pragma unroll
    for (int w = 0; w < WARPS; w++) if (w < warp) add += s_wsum[w];
    uint32_t e = x + add - tsum;
#pragma unroll
    for (int j = 0; j < DPT; j++)
      { const int d = tid * DPT + j;
        dbase[j] = e;
        if (d < NB) s_dbase[d] = e;
        e += cnt[j];
      }
  }
  __syncthreads();                                         // (B2)

  uint32_t pos[ITEMS];
  { uint32_t peers[ITEMS], old[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; i++) peers[i] = __match_any_sync(0xffffffffu, dig[i]);
#pragma unroll
    for (int i = 0; i < ITEMS; i++)
      { old[i] = 0;
        if ((peers[i] & lt) == 0) old[i] = atomicAdd(&wh[dig[i]], (uint32_t) __popc(peers[i]));
      }
#pragma unroll
    for (int i = 0; i < ITEMS; i++)
      pos[i] = s_dbase[dig[i]] + __shfl_sync(0xffffffffu, old[i], __ffs(peers[i]) - 1) + __popc(peers[i] & lt);
  }
  if (ALIAS) __syncthreads();                              // (C) counters die, staging live
#pragma unroll
  for (int i = 0; i < ITEMS; i++) stage[pos[i]] = rec[i];

#pragma unroll
  for (int j = 0; j < DPT; j++)
    { const int d = tid * DPT + j;
      if (d < NB)
        { uint32_t excl = 0;
          if (tile > 0)
            { const uint32_t *p = tile_state + (size_t) (tile - 1) * NB + d;
              while (true)
                { uint32_t v = ld_relaxed(p);
                  if (v & FLAG_INC) { excl += v & VAL_MASK; break; }
                  if (v & FLAG_AGG) { excl += v & VAL_MASK; p -= NB; continue; }
                }
              st_relaxed(tile_state + (size_t) tile * NB + d, FLAG_INC | (excl + cnt[j]));
            }
          s_delta[d] = gbase[d] + excl - dbase[j];
        }
    }
  __syncthreads();                                         // (D)
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
    { const uint32_t j = tid + i * THREADS;
      if (j < nvalid)
        { const uint4 r = stage[j];
          __stcs(out + (j + s_delta[rec_digit<BITS>(r, word, shift)]), r);
        }
    }
}


// ---------------- V2: rank first (match or ballots), optional staging ----------------
template <int BITS> __device__ __forceinline__ uint32_t match_ballot(uint32_t dig)
{ uint32_t m = 0xffffffffu;
#pragma unroll
  for (int b = 0; b < BITS; b++)
    { const bool p = (dig >> b) & 1;
      const uint32_t v = __ballot_sync(0xffffffffu, p);
      m &= p ? v : ~v;
    }
  return m;
}

template <int THREADS, int ITEMS, int BITS, int MATCHMODE, bool STAGED, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_pass_v2(const uint4 *__restrict__ in, uint4 *__restrict__ out, uint32_t n, int word, int shift,
          const uint32_t *__restrict__ gbase, uint32_t *tile_state, uint32_t *tile_counter)
{ constexpr int NB = 1 << BITS, WARPS = THREADS / 32, TILE = THREADS * ITEMS;
  constexpr int DPT = (NB + THREADS - 1) / THREADS;
  extern __shared__ uint4 dyn[];
  uint4    *stage = dyn;                                   // STAGED only; aliases the counters
  uint32_t *whist = reinterpret_cast<uint32_t *>(dyn);     // [WARPS][NB]
  __shared__ uint32_t s_dbase[STAGED ? NB : 1];
  __shared__ uint32_t s_delta[NB];
  __shared__ uint32_t s_wsum[WARPS];
  __shared__ uint32_t s_tile;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lt = (1u << lane) - 1;
  if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
  uint32_t *wh = whist + warp * NB;
#pragma unroll
  for (int i = lane; i < NB; i += 32) wh[i] = 0;
  __syncthreads();
  const uint32_t tile  = s_tile;
  const uint32_t tbase = tile * (uint32_t) TILE;
  const uint32_t nvalid = (n - tbase < (uint32_t) TILE) ? n - tbase : (uint32_t) TILE;
  const uint32_t base = tbase + warp * (32 * ITEMS) + lane;

  uint4 rec[ITEMS]; uint32_t rank[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
    { const uint32_t idx = base + i * 32;
      rec[i] = (idx < n) ? __ldcs(in + idx) : make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    }
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
    { const uint32_t dig = rec_digit<BITS>(rec[i], word, shift);
      const uint32_t peers = MATCHMODE ? match_ballot<BITS>(dig) : __match_any_sync(0xffffffffu, dig);
      uint32_t old = 0;
      if ((peers & lt) == 0) old = atomicAdd(&wh[dig], (uint32_t) __popc(peers));
      rank[i] = __shfl_sync(0xffffffffu, old, __ffs(peers) - 1) + __popc(peers & lt);
    }
  __syncthreads();                                         // (A)

  uint32_t cnt[DPT], dbase[DPT];
  { uint32_t tsum = 0;
#pragma unroll
    for (int j = 0; j < DPT; j++)
      { const int d = tid * DPT + j;
        uint32_t run = 0;
        if (d < NB)
          {
#pragma unroll
            for (int w = 0; w < WARPS; w++) { uint32_t c = whist[w * NB + d]; whist[w * NB + d] = run; run += c; }
            if (d == NB - 1) run -= (uint32_t) TILE - nvalid;
            st_relaxed(tile_state + (size_t) tile * NB + d, (tile == 0 ? FLAG_INC : FLAG_AGG) | run);
          }
        cnt[j] = run; tsum += run;
      }
    if (STAGED)
      { uint32_t x = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) s_wsum[warp] = x;
        __syncthreads();                                   // (B1)
        uint32_t add = 0;
#pragma unroll
        for (int w = 0; w < WARPS; w++) if (w < warp) add += s_wsum[w];
        uint32_t e = x + add - tsum;
#pragma unroll
        for (int j = 0; j < DPT; j++)
          { const int d = tid * DPT + j;
            dbase[j] = e;
            if (d < NB) s_dbase[d] = e;
            e += cnt[j];
          }
        __syncthreads();                                   // (B2)
#pragma unroll
        for (int i = 0; i < ITEMS; i++)
          { const uint32_t dig = rec_digit<BITS>(rec[i], word, shift);
            rank[i] += s_dbase[dig] + wh[dig];
          }
        __syncthreads();                                   // (C) counters die, staging live
#pragma unroll
        for (int i = 0; i < ITEMS; i++) stage[rank[i]] = rec[i];
      }
    else
      {
#pragma unroll
        for (int j = 0; j < DPT; j++) dbase[j] = 0;
      }
  }
#pragma unroll
  for (int j = 0; j < DPT; j++)
    { const int d = tid * DPT + j;
      if (d < NB)
        { uint32_t excl = 0;
          if (tile > 0)
            { const uint32_t *p = tile_state + (size_t) (tile - 1) * NB + d;
              while (true)
                { uint32_t v = ld_relaxed(p);
                  if (v & FLAG_INC) { excl += v & VAL_MASK; break; }
                  if (v & FLAG_AGG) { excl += v & VAL_MASK; p -= NB; continue; }
                }
              st_relaxed(tile_state + (size_t) tile * NB + d, FLAG_INC | (excl + cnt[j]));
            }
          s_delta[d] = gbase[d] + excl - dbase[j];
        }
    }
  __syncthreads();                                         // (D)
  if (STAGED)
    {
#pragma unroll
      for (int i = 0; i < ITEMS; i++)
        { const uint32_t j = tid + i * THREADS;
          if (j < nvalid)
            { const uint4 r = stage[j];
              __stcs(out + (j + s_delta[rec_digit<BITS>(r, word, shift)]), r);
            }
        }
    }
  else
    {
#pragma unroll
      for (int i = 0; i < ITEMS; i++)
        { const uint32_t idx = base + i * 32;
          const uint32_t dig = rec_digit<BITS>(rec[i], word, shift);
          if (idx < n)
            __stcs(out + (s_delta[dig] + wh[dig] + rank[i]), rec[i]);
        }
    }
}

template <int THREADS, int ITEMS, int BITS, int MATCHMODE, bool STAGED, int MINB, bool TMAST>
__global__ void __launch_bounds__(THREADS, MINB)
k_pass_v3(const uint4 *__restrict__ in, uint4 *__restrict__ out, uint32_t n, int word, int shift,
          const uint32_t *__restrict__ gbase, uint32_t *tile_state, uint32_t *tile_counter)
{ constexpr int NB = 1 << BITS, WARPS = THREADS / 32, TILE = THREADS * ITEMS;
  constexpr int DPT = (NB + THREADS - 1) / THREADS;
  extern __shared__ uint4 dyn[];
  uint4    *stage = dyn;                                   // STAGED only; aliases the counters
  uint32_t *whist = reinterpret_cast<uint32_t *>(dyn);     // [WARPS][NB]
  __shared__ uint32_t s_dbase[STAGED ? NB : 1];
  __shared__ uint32_t s_delta[NB];
  __shared__ uint32_t s_wsum[WARPS];
  __shared__ uint32_t s_tile;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lt = (1u << lane) - 1;
  if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
  uint32_t *wh = whist + warp * NB;
#pragma unroll
  for (int i = lane; i < NB; i += 32) wh[i] = 0;
  __syncthreads();
  const uint32_t tile  = s_tile;
  const uint32_t tbase = tile * (uint32_t) TILE;
  const uint32_t nvalid = (n - tbase < (uint32_t) TILE) ? n - tbase : (uint32_t) TILE;
  const uint32_t base = tbase + warp * (32 * ITEMS) + lane;

  uint4 rec[ITEMS]; uint32_t rank[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
    { const uint32_t idx = base + i * 32;
      rec[i] = (idx < n) ? __ldcs(in + idx) : make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    }
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
    { const uint32_t dig = rec_digit<BITS>(rec[i], word, shift);
      const uint32_t peers = MATCHMODE ? match_ballot<BITS>(dig) : __match_any_sync(0xffffffffu, dig);
      uint32_t old = 0;
      if ((peers & lt) == 0) old = atomicAdd(&wh[dig], (uint32_t) __popc(peers));
      rank[i] = __shfl_sync(0xffffffffu, old, __ffs(peers) - 1) + __popc(peers & lt);
    }
  __syncthreads();                                         // (A)

  uint32_t cnt[DPT], dbase[DPT], gdst[DPT];
  { uint32_t tsum = 0;
#pragma unroll
    for (int j = 0; j < DPT; j++)
      { const int d = tid * DPT + j;
        uint32_t run = 0;
        if (d < NB)
          {
#pragma unroll
            for (int w = 0; w < WARPS; w++) { uint32_t c = whist[w * NB + d]; whist[w * NB + d] = run; run += c; }
            if (d == NB - 1) run -= (uint32_t) TILE - nvalid;
            st_relaxed(tile_state + (size_t) tile * NB + d, (tile == 0 ? FLAG_INC : FLAG_AGG) | run);
          }
        cnt[j] = run; tsum += run;
      }
    if (STAGED)
      { uint32_t x = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) s_wsum[warp] = x;
        __syncthreads();                                   // (B1)
        uint32_t add = 0;
#pragma unroll
        for (int w = 0; w < WARPS; w++) if (w < warp) add += s_wsum[w];
        uint32_t e = x + add - tsum;
#pragma unroll
        for (int j = 0; j < DPT; j++)
          { const int d = tid * DPT + j;
            dbase[j] = e;
            if (d < NB) s_dbase[d] = e;
            e += cnt[j];
          }
        __syncthreads();                                   // (B2)
#pragma unroll
        for (int i = 0; i < ITEMS; i++)
          { const uint32_t dig = rec_digit<BITS>(rec[i], word, shift);
            rank[i] += s_dbase[dig] + wh[dig];
          }
        __syncthreads();                                   // (C) counters die, staging live
#pragma unroll
        for (int i = 0; i < ITEMS; i++) stage[rank[i]] = rec[i];
        if (TMAST) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
    else
      {
#pragma unroll
        for (int j = 0; j < DPT; j++) dbase[j] = 0;
      }
  }
#pragma unroll
  for (int j = 0; j < DPT; j++)
    { const int d = tid * DPT + j;
      if (d < NB)
        { uint32_t excl = 0;
          if (tile > 0)
            { const uint32_t *p = tile_state + (size_t) (tile - 1) * NB + d;
              while (true)
                { uint32_t v = ld_relaxed(p);
                  if (v & FLAG_INC) { excl += v & VAL_MASK; break; }
                  if (v & FLAG_AGG) { excl += v & VAL_MASK; p -= NB; continue; }
                }
              st_relaxed(tile_state + (size_t) tile * NB + d, FLAG_INC | (excl + cnt[j]));
            }
          s_delta[d] = gbase[d] + excl - dbase[j];
          gdst[j] = gbase[d] + excl;
        }
    }
  __syncthreads();                                         // (D)
  if (STAGED && TMAST)
    {
#pragma unroll
      for (int j = 0; j < DPT; j++)
        { const int d = tid * DPT + j;
          if (d < NB && cnt[j] > 0)
            { const uint32_t sa = (uint32_t) __cvta_generic_to_shared(stage + dbase[j]);
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                           :: "l"(out + gdst[j]), "r"(sa), "r"(cnt[j] * 16u) : "memory");
            }
        }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  else if (STAGED)
    {
#pragma unroll
      for (int i = 0; i < ITEMS; i++)
        { const uint32_t j = tid + i * THREADS;
          if (j < nvalid)
            { const uint4 r = stage[j];
              __stcs(out + (j + s_delta[rec_digit<BITS>(r, word, shift)]), r);
            }
        }
    }
  else
    {
#pragma unroll
      for (int i = 0; i < ITEMS; i++)
        { const uint32_t idx = base + i * 32;
          const uint32_t dig = rec_digit<BITS>(rec[i], word, shift);
          if (idx < n)
            __stcs(out + (s_delta[dig] + wh[dig] + rank[i]), rec[i]);
        }
    }
}

template <int THREADS, int ITEMS, int BITS, int MATCHMODE, bool STAGED, int MINB, bool TMAST, int PROBE>
__global__ void __launch_bounds__(THREADS, MINB)
k_pass_v4(const uint4 *__restrict__ in, uint4 *__restrict__ out, uint32_t n, int word, int shift,
          const uint32_t *__restrict__ gbase, uint32_t *tile_state, uint32_t *tile_counter)
{ constexpr int NB = 1 << BITS, WARPS = THREADS / 32, TILE = THREADS * ITEMS;
  constexpr int DPT = (NB + THREADS - 1) / THREADS;
  extern __shared__ uint4 dyn[];
  uint4    *stage = dyn;                                   // STAGED only; aliases the counters
  uint32_t *whist = reinterpret_cast<uint32_t *>(dyn);     // [WARPS][NB]
  __shared__ uint32_t s_dbase[STAGED ? NB : 1];
  __shared__ uint32_t s_delta[NB];
  __shared__ uint32_t s_wsum[WARPS];
  __shared__ uint32_t s_tile;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lt = (1u << lane) - 1;
  if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
  uint32_t *wh = whist + warp * NB;
#pragma unroll
  for (int i = lane; i < NB; i += 32) wh[i] = 0;
  __syncthreads();
  const uint32_t tile  = s_tile;
  const uint32_t tbase = tile * (uint32_t) TILE;
  const uint32_t nvalid = (n - tbase < (uint32_t) TILE) ? n - tbase : (uint32_t) TILE;
  const uint32_t base = tbase + warp * (32 * ITEMS) + lane;

  uint4 rec[ITEMS]; uint32_t rank[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
    { const uint32_t idx = base + i * 32;
      rec[i] = (idx < n) ? __ldcs(in + idx) : make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    }
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
    { const uint32_t dig = rec_digit<BITS>(rec[i], word, shift);
      const uint32_t peers = MATCHMODE ? match_ballot<BITS>(dig) : __match_any_sync(0xffffffffu, dig);
      uint32_t old = 0;
      if ((peers & lt) == 0) old = atomicAdd(&wh[dig], (uint32_t) __popc(peers));
      rank[i] = __shfl_sync(0xffffffffu, old, __ffs(peers) - 1) + __popc(peers & lt);
    }
  __syncthreads();                                         // (A)

  uint32_t cnt[DPT], dbase[DPT], gdst[DPT];
  { uint32_t tsum = 0;
#pragma unroll
    for (int j = 0; j < DPT; j++)
      { const int d = tid * DPT + j;
        uint32_t run = 0;
        if (d < NB)
          {
#pragma unroll
            for (int w = 0; w < WARPS; w++) { uint32_t c = whist[w * NB + d]; whist[w * NB + d] = run; run += c; }
            if (d == NB - 1) run -= (uint32_t) TILE - nvalid;
            st_relaxed(tile_state + (size_t) tile * NB + d, (tile == 0 ? FLAG_INC : FLAG_AGG) | run);
          }
        cnt[j] = run; tsum += run;
      }
    if (STAGED)
      { uint32_t x = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) s_wsum[warp] = x;
        __syncthreads();                                   // (B1)
        uint32_t add = 0;
#pragma unroll
        for (int w = 0; w < WARPS; w++) if (w < warp) add += s_wsum[w];
        uint32_t e = x + add - tsum;
#pragma unroll
        for (int j = 0; j < DPT; j++)
          { const int d = tid * DPT + j;
            dbase[j] = e;
            if (d < NB) s_dbase[d] = e;
            e += cnt[j];
          }
        __syncthreads();                                   // (B2)
#pragma unroll
        for (int i = 0; i < ITEMS; i++)
          { const uint32_t dig = rec_digit<BITS>(rec[i], word, shift);
            rank[i] += s_dbase[dig] + wh[dig];
          }
        __syncthreads();                                   // (C) counters die, staging live
#pragma unroll
        for (int i = 0; i < ITEMS; i++) stage[rank[i]] = rec[i];
        if (TMAST) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
    else
      {
#pragma unroll
        for (int j = 0; j < DPT; j++) dbase[j] = 0;
      }
  }
#pragma unroll
  for (int j = 0; j < DPT; j++)
    { const int d = tid * DPT + j;
      if (d < NB)
        { uint32_t excl = 0;
          if (tile > 0)
            { const uint32_t *col = tile_state + d;
              int64_t t = (int64_t) tile - 1;
              bool done = false;
              while (!done)
                { uint32_t v[PROBE];
#pragma unroll
                  for (int k = 0; k < PROBE; k++)
                    v[k] = (t - k >= 0) ? ld_relaxed(col + (size_t) (t - k) * NB) : FLAG_INC;
                  int used = PROBE;
#pragma unroll
                  for (int k = 0; k < PROBE; k++)
                    if (!done && used == PROBE)
                      { if (v[k] & FLAG_INC)      { excl += v[k] & VAL_MASK; done = true; }
                        else if (v[k] & FLAG_AGG) excl += v[k] & VAL_MASK;
                        else                      used = k;
                      }
                  t -= used;
                }
              st_relaxed(tile_state + (size_t) tile * NB + d, FLAG_INC | (excl + cnt[j]));
            }
          s_delta[d] = gbase[d] + excl - dbase[j];
          gdst[j] = gbase[d] + excl;
        }
    }
  __syncthreads();                                         // (D)
  if (STAGED && TMAST)
    {
#pragma unroll
      for (int j = 0; j < DPT; j++)
        { const int d = tid * DPT + j;
          if (d < NB && cnt[j] > 0)
            { const uint32_t sa = (uint32_t) __cvta_generic_to_shared(stage + dbase[j]);
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                           :: "l"(out + gdst[j]), "r"(sa), "r"(cnt[j] * 16u) : "memory");
            }
        }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  else if (STAGED)
    {
#pragma unroll
      for (int i = 0; i < ITEMS; i++)
        { const uint32_t j = tid + i * THREADS;
          if (j < nvalid)
            { const uint4 r = stage[j];
              __stcs(out + (j + s_delta[rec_digit<BITS>(r, word, shift)]), r);
            }
        }
    }
  else
    {
#pragma unroll
      for (int i = 0; i < ITEMS; i++)
        { const uint32_t idx = base + i * 32;
          const uint32_t dig = rec_digit<BITS>(rec[i], word, shift);
          if (idx < n)
            __stcs(out + (s_delta[dig] + wh[dig] + rank[i]), rec[i]);
        }
    }
}

template <int THREADS, int ITEMS, int BITS, int MATCHMODE, bool STAGED, int MINB, bool TMAST, int PROBE>
__global__ void __launch_bounds__(THREADS, MINB)
k_pass_v7(const uint4 *__restrict__ in, uint4 *__restrict__ out, uint32_t n, int word, int shift,
          const uint32_t *__restrict__ gbase, uint32_t *tile_state, uint32_t *tile_counter)
{ constexpr int NB = 1 << BITS, WARPS = THREADS / 32, TILE = THREADS * ITEMS;
  constexpr int DPT = (NB + THREADS - 1) / THREADS;
  extern __shared__ uint4 dyn[];
  uint4    *stage = dyn;                                   // STAGED only; aliases the counters
  uint32_t *whist = reinterpret_cast<uint32_t *>(dyn);     // [WARPS][NB]
  __shared__ uint32_t s_dbase[STAGED ? NB : 1];
  __shared__ uint32_t s_delta[NB];
  __shared__ uint32_t s_wsum[WARPS];
  __shared__ uint32_t s_tile;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lt = (1u << lane) - 1;
  if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
  uint32_t *wh = whist + warp * NB;
#pragma unroll
  for (int i = lane; i < NB; i += 32) wh[i] = 0;
  __syncthreads();
  const uint32_t tile  = s_tile;
  const uint32_t tbase = tile * (uint32_t) TILE;
  const uint32_t nvalid = (n - tbase < (uint32_t) TILE) ? n - tbase : (uint32_t) TILE;
  const uint32_t base = tbase + warp * (32 * ITEMS) + lane;

  uint4 rec[ITEMS]; uint32_t rank[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
    { const uint32_t idx = base + i * 32;
      rec[i] = (idx < n) ? __ldcs(in + idx) : make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    }
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
    { const uint32_t dig = rec_digit<BITS>(rec[i], word, shift);
      const uint32_t peers = MATCHMODE ? match_ballot<BITS>(dig) : __match_any_sync(0xffffffffu, dig);
      const uint32_t prev = wh[dig];
      __syncwarp();
      if ((peers & lt) == 0) wh[dig] = prev + (uint32_t) __popc(peers);
      __syncwarp();
      rank[i] = prev + __popc(peers & lt);
    }
  __syncthreads();                                         // (A)

  uint32_t cnt[DPT], dbase[DPT], gdst[DPT];
  { uint32_t tsum = 0;
#pragma unroll
    for (int j = 0; j < DPT; j++)
      { const int d = tid * DPT + j;
        uint32_t run = 0;
        if (d < NB)
          {
#pragma unroll
            for (int w = 0; w < WARPS; w++) { uint32_t c = whist[w * NB + d]; whist[w * NB + d] = run; run += c; }
            if (d == NB - 1) run -= (uint32_t) TILE - nvalid;
            st_relaxed(tile_state + (size_t) tile * NB + d, (tile == 0 ? FLAG_INC : FLAG_AGG) | run);
          }
        cnt[j] = run; tsum += run;
      }
    if (STAGED)
      { uint32_t x = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) s_wsum[warp] = x;
        __syncthreads();                                   // (B1)
        uint32_t add = 0;
#pragma unroll
        for (int w = 0; w < WARPS; w++) if (w < warp) add += s_wsum[w];
        uint32_t e = x + add - tsum;
#pragma unroll
        for (int j = 0; j < DPT; j++)
          { const int d = tid * DPT + j;
            dbase[j] = e;
            if (d < NB) s_dbase[d] = e;
            e += cnt[j];
          }
        __syncthreads();                                   // (B2)
#pragma unroll
        for (int i = 0; i < ITEMS; i++)
          { const uint32_t dig = rec_digit<BITS>(rec[i], word, shift);
            rank[i] += s_dbase[dig] + wh[dig];
          }
        __syncthreads();                                   // (C) counters die, staging live
#pragma unroll
        for (int i = 0; i < ITEMS; i++) stage[rank[i]] = rec[i];
        if (TMAST) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
    else
      {
#pragma unroll
        for (int j = 0; j < DPT; j++) dbase[j] = 0;
      }
  }
#pragma unroll
  for (int j = 0; j < DPT; j++)
    { const int d = tid * DPT + j;
      if (d < NB)
        { uint32_t excl = 0;
          if (tile > 0)
            { const uint32_t *col = tile_state + d;
              int64_t t = (int64_t) tile - 1;
              bool done = false;
              while (!done)
                { uint32_t v[PROBE];
#pragma unroll
                  for (int k = 0; k < PROBE; k++)
                    v[k] = (t - k >= 0) ? ld_relaxed(col + (size_t) (t - k) * NB) : FLAG_INC;
                  int used = PROBE;
#pragma unroll
                  for (int k = 0; k < PROBE; k++)
                    if (!done && used == PROBE)
                      { if (v[k] & FLAG_INC)      { excl += v[k] & VAL_MASK; done = true; }
                        else if (v[k] & FLAG_AGG) excl += v[k] & VAL_MASK;
                        else                      used = k;
                      }
                  t -= used;
                }
              st_relaxed(tile_state + (size_t) tile * NB + d, FLAG_INC | (excl + cnt[j]));
            }
          s_delta[d] = gbase[d] + excl - dbase[j];
          gdst[j] = gbase[d] + excl;
        }
    }
  __syncthreads();                                         // (D)
  if (STAGED && TMAST)
    {
#pragma unroll
      for (int j = 0; j < DPT; j++)
        { const int d = tid * DPT + j;
          if (d < NB && cnt[j] > 0)
            { const uint32_t sa = (uint32_t) __cvta_generic_to_shared(stage + dbase[j]);
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                           :: "l"(out + gdst[j]), "r"(sa), "r"(cnt[j] * 16u) : "memory");
            }
        }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  else if (STAGED)
    {
#pragma unroll
      for (int i = 0; i < ITEMS; i++)
        { const uint32_t j = tid + i * THREADS;
          if (j < nvalid)
            { const uint4 r = stage[j];
              __stcs(out + (j + s_delta[rec_digit<BITS>(r, word, shift)]), r);
            }
        }
    }
  else
    {
#pragma unroll
      for (int i = 0; i < ITEMS; i++)
        { const uint32_t idx = base + i * 32;
          const uint32_t dig = rec_digit<BITS>(rec[i], word, shift);
          if (idx < n)
            __stcs(out + (s_delta[dig] + wh[dig] + rank[i]), rec[i]);
        }
    }
}

// tuned ballot matcher: per bit one LOP3 (bit -> predicate), one VOTE, one SEL, one LOP3 (accumulate lanes that differ)
template <int BITS> __device__ __forceinline__ uint32_t match_ballot2(uint32_t dig)
{ uint32_t diff = 0;
#pragma unroll
  for (int b = 0; b < BITS; b++)
    { uint32_t v, sx;
      asm("{ .reg .pred p; .reg .b32 t;\n\t"
          "and.b32 t, %2, %3;\n\t"
          "setp.ne.u32 p, t, 0;\n\t"
          "vote.sync.ballot.b32 %0, p, 0xffffffff;\n\t"
          "selp.b32 %1, 0xffffffff, 0, p;\n\t}"
          : "=r"(v), "=r"(sx) : "r"(dig), "r"(1u << b));
      diff |= v ^ sx;                                    // lanes whose bit b differs from mine
    }
  return ~diff;
}

// collision-table matcher: lanes publish their id under their digit; only digits held by several
// lanes are resolved with ballots (about two per 32 random 8-bit digits)
__device__ __forceinline__ uint32_t match_table(uint32_t dig, uint32_t *tag, int lane)
{ tag[dig] = lane;
  __syncwarp();
  const uint32_t w = tag[dig];
  __syncwarp();
  uint32_t rem = __ballot_sync(0xffffffffu, w != (uint32_t) lane);
  uint32_t peers = 1u << lane;
  while (rem)
    { const int l = __ffs(rem) - 1;
      const uint32_t d0 = __shfl_sync(0xffffffffu, dig, l);
      const uint32_t grp = __ballot_sync(0xffffffffu, dig == d0);
      if (dig == d0) peers = grp;
      rem &= ~grp;
    }
  return peers;
}

// ---------------- V9: v8 + PRMT digits (word as template), unchecked loads on full tiles, lean probe-8 look-back ----------------
template <int THREADS, int ITEMS, int MINB, int PROBE, int W32>
__global__ void __launch_bounds__(THREADS, MINB)
k_pass_v9(const uint4 *__restrict__ in, uint4 *__restrict__ out, uint32_t n, uint32_t psel,
          const uint32_t *__restrict__ gbase, uint32_t *tile_state, uint32_t *tile_counter)
{ constexpr int NB = 256, WARPS = THREADS / 32, TILE = THREADS * ITEMS;
  extern __shared__ uint4 dyn[];
  uint4    *stage = dyn;
  uint32_t *whist = reinterpret_cast<uint32_t *>(dyn);
  __shared__ uint32_t s_wsum[WARPS];
  __shared__ uint32_t s_tile;
#define DIG9(r) __byte_perm((W32 == 0) ? (r).x : (W32 == 1) ? (r).y : (W32 == 2) ? (r).z : (r).w, 0, psel)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lt = (1u << lane) - 1;
  if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
  uint32_t *wh = whist + warp * NB;
#pragma unroll
  for (int i = lane; i < NB; i += 32) wh[i] = 0;
  __syncthreads();
  const uint32_t tile  = s_tile;
  const uint32_t tbase = tile * (uint32_t) TILE;
  const uint32_t nvalid = (n - tbase < (uint32_t) TILE) ? n - tbase : (uint32_t) TILE;
  const uint4 *src = in + tbase + warp * (32 * ITEMS) + lane;

  uint4 rec[ITEMS];
  if (nvalid == (uint32_t) TILE)
    {
#pragma unroll
      for (int i = 0; i < ITEMS; i++) rec[i] = __ldcs(src + i * 32);
    }
  else
    { const uint32_t base = tbase + warp * (32 * ITEMS) + lane;
#pragma unroll
      for (int i = 0; i < ITEMS; i++)
        rec[i] = (base + i * 32 < n) ? __ldcs(src + i * 32) : make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    }
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
    atomicAdd(&wh[DIG9(rec[i])], 1u);
  __syncthreads();

  uint32_t cnt = 0, dbase;
  uint32_t *st = tile_state + (size_t) tile * 256 + (tid & 255);
  if (tid < 256)
    {
#pragma unroll
      for (int w = 0; w < WARPS; w++) cnt += whist[w * 256 + tid];
      if (tid == 255) cnt -= (uint32_t) TILE - nvalid;
      st_relaxed(st, (tile == 0 ? FLAG_INC : FLAG_AGG) | cnt);
    }
  { uint32_t x = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_wsum[warp] = x;
    __syncthreads();
    uint32_t add = 0;
#pragma unroll
    for (int w = 0; w < WARPS; w++) if (w < warp) add += s_wsum[w];
    dbase = x + add - cnt;
  }
  if (tid < 256)
    { uint32_t run = dbase;
#pragma unroll
      for (int w = 0; w < WARPS; w++) { const uint32_t c = whist[w * 256 + tid]; whist[w * 256 + tid] = run; run += c; }
    }
  __syncthreads();

  uint32_t pos[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
    { const uint32_t dig   = DIG9(rec[i]);
      const uint32_t peers = match_ballot2<8>(dig);
      uint32_t old = 0;
      if ((peers & lt) == 0) old = atomicAdd(&wh[dig], (uint32_t) __popc(peers));
      pos[i] = __shfl_sync(0xffffffffu, old, __ffs(peers) - 1) + __popc(peers & lt);
    }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < ITEMS; i++) stage[pos[i]] = rec[i];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

  uint32_t gdst = 0;
  if (tid < 256)
    { uint32_t excl = 0;
      if (tile > 0)
        { const uint32_t *p = tile_state + (size_t) (tile - 1) * 256 + tid;
          uint32_t left = tile;
          while (true)
            { uint32_t v[PROBE];
#pragma unroll
              for (int k = 0; k < PROBE; k++)
                v[k] = ((uint32_t) k < left) ? ld_relaxed(p - (size_t) k * 256) : FLAG_INC;
              uint32_t adv = 0; bool stop = false, inc = false;
#pragma unroll
              for (int k = 0; k < PROBE; k++)
                if (!stop)
                  { if (v[k] == 0) stop = true;                       // not published yet
                    else
                      { excl += v[k] & VAL_MASK; adv += 1;
                        if (v[k] & FLAG_INC) { stop = true; inc = true; }
                      }
                  }
              if (inc) break;
              p -= (size_t) adv * 256; left -= adv;
            }
          st_relaxed(st, FLAG_INC | (excl + cnt));
        }
      gdst = gbase[tid] + excl;
    }
  __syncthreads();
  if (tid < 256 && cnt > 0)
    { const uint32_t sa = (uint32_t) __cvta_generic_to_shared(stage + dbase);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                   :: "l"(out + gdst), "r"(sa), "r"(cnt * 16u) : "memory");
    }
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#undef DIG9
}

// ---------------- V10 (v9 + mixed matcher / highest-lane leader): v8 + PRMT digits (word as template), unchecked loads on full tiles, lean probe-8 look-back ----------------
template <int THREADS, int ITEMS, int MINB, int PROBE, int W32, int MIX, int LH>
__global__ void __launch_bounds__(THREADS, MINB)
k_pass_v10(const uint4 *__restrict__ in, uint4 *__restrict__ out, uint32_t n, uint32_t psel,
          const uint32_t *__restrict__ gbase, uint32_t *tile_state, uint32_t *tile_counter)
{ constexpr int NB = 256, WARPS = THREADS / 32, TILE = THREADS * ITEMS;
  extern __shared__ uint4 dyn[];
  uint4    *stage = dyn;
  uint32_t *whist = reinterpret_cast<uint32_t *>(dyn);
  __shared__ uint32_t s_wsum[WARPS];
  __shared__ uint32_t s_tile;
#define DIG9(r) __byte_perm((W32 == 0) ? (r).x : (W32 == 1) ? (r).y : (W32 == 2) ? (r).z : (r).w, 0, psel)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lt = (1u << lane) - 1;
  if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
  uint32_t *wh = whist + warp * NB;
#pragma unroll
  for (int i = lane; i < NB; i += 32) wh[i] = 0;
  __syncthreads();
  const uint32_t tile  = s_tile;
  const uint32_t tbase = tile * (uint32_t) TILE;
  const uint32_t nvalid = (n - tbase < (uint32_t) TILE) ? n - tbase : (uint32_t) TILE;
  const uint4 *src = in + tbase + warp * (32 * ITEMS) + lane;

  uint4 rec[ITEMS];
  if (nvalid == (uint32_t) TILE)
    {
#pragma unroll
      for (int i = 0; i < ITEMS; i++) rec[i] = __ldcs(src + i * 32);
    }
  else
    { const uint32_t base = tbase + warp * (32 * ITEMS) + lane;
#pragma unroll
      for (int i = 0; i < ITEMS; i++)
        rec[i] = (base + i * 32 < n) ? __ldcs(src + i * 32) : make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    }
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
    atomicAdd(&wh[DIG9(rec[i])], 1u);
  __syncthreads();

  uint32_t cnt = 0, dbase;
  uint32_t *st = tile_state + (size_t) tile * 256 + (tid & 255);
  if (tid < 256)
    {
#pragma unroll
      for (int w = 0; w < WARPS; w++) cnt += whist[w * 256 + tid];
      if (tid == 255) cnt -= (uint32_t) TILE - nvalid;
      st_relaxed(st, (tile == 0 ? FLAG_INC : FLAG_AGG) | cnt);
    }
  { uint32_t x = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_wsum[warp] = x;
    __syncthreads();
    uint32_t add = 0;
#pragma unroll
    for (int w = 0; w < WARPS; w++) if (w < warp) add += s_wsum[w];
    dbase = x + add - cnt;
  }
  if (tid < 256)
    { uint32_t run = dbase;
#pragma unroll
      for (int w = 0; w < WARPS; w++) { const uint32_t c = whist[w * 256 + tid]; whist[w * 256 + tid] = run; run += c; }
    }
  __syncthreads();

  uint32_t pos[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
    { const uint32_t dig   = DIG9(rec[i]);
      const uint32_t peers = (MIX > 0 && (i % MIX) == MIX - 1) ? __match_any_sync(0xffffffffu, dig) : match_ballot2<8>(dig);
      uint32_t old = 0;
      if (LH)
        { const int leader = 31 - __clz(peers);
          if (lane == leader) old = atomicAdd(&wh[dig], (uint32_t) __popc(peers));
          pos[i] = __shfl_sync(0xffffffffu, old, leader) + __popc(peers & lt);
        }
      else
        { if ((peers & lt) == 0) old = atomicAdd(&wh[dig], (uint32_t) __popc(peers));
          pos[i] = __shfl_sync(0xffffffffu, old, __ffs(peers) - 1) + __popc(peers & lt);
        }
    }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < ITEMS; i++) stage[pos[i]] = rec[i];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

  uint32_t gdst = 0;
  if (tid < 256)
    { uint32_t excl = 0;
      if (tile > 0)
        { const uint32_t *p = tile_state + (size_t) (tile - 1) * 256 + tid;
          uint32_t left = tile;
          while (true)
            { uint32_t v[PROBE];
#pragma unroll
              for (int k = 0; k < PROBE; k++)
                v[k] = ((uint32_t) k < left) ? ld_relaxed(p - (size_t) k * 256) : FLAG_INC;
              uint32_t adv = 0; bool stop = false, inc = false;
#pragma unroll
              for (int k = 0; k < PROBE; k++)
                if (!stop)
                  { if (v[k] == 0) stop = true;                       // not published yet
                    else
                      { excl += v[k] & VAL_MASK; adv += 1;
                        if (v[k] & FLAG_INC) { stop = true; inc = true; }
                      }
                  }
              if (inc) break;
              p -= (size_t) adv * 256; left -= adv;
            }
          st_relaxed(st, FLAG_INC | (excl + cnt));
        }
      gdst = gbase[tid] + excl;
    }
  __syncthreads();
  if (tid < 256 && cnt > 0)
    { const uint32_t sa = (uint32_t) __cvta_generic_to_shared(stage + dbase);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                   :: "l"(out + gdst), "r"(sa), "r"(cnt * 16u) : "memory");
    }
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#undef DIG9
}

// ---------------- V5: early counts + ballot ranking straight to staged positions + probes + TMA store ----------------
template <int THREADS, int ITEMS, int BITS, int MINB, int PROBE>
__global__ void __launch_bounds__(THREADS, MINB)
k_pass_v5(const uint4 *__restrict__ in, uint4 *__restrict__ out, uint32_t n, int word, int shift,
          const uint32_t *__restrict__ gbase, uint32_t *tile_state, uint32_t *tile_counter)
{ constexpr int NB = 1 << BITS, WARPS = THREADS / 32, TILE = THREADS * ITEMS;
  constexpr int DPT = (NB + THREADS - 1) / THREADS;
  extern __shared__ uint4 dyn[];
  uint4    *stage = dyn;                                   // aliases the counters
  uint32_t *whist = reinterpret_cast<uint32_t *>(dyn);     // [WARPS][NB]
  __shared__ uint32_t s_wsum[WARPS];
  __shared__ uint32_t s_tile;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lt = (1u << lane) - 1;
  if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
  uint32_t *wh = whist + warp * NB;
#pragma unroll
  for (int i = lane; i < NB; i += 32) wh[i] = 0;
  __syncthreads();
  const uint32_t tile  = s_tile;
  const uint32_t tbase = tile * (uint32_t) TILE;
  const uint32_t nvalid = (n - tbase < (uint32_t) TILE) ? n - tbase : (uint32_t) TILE;
  const uint32_t base = tbase + warp * (32 * ITEMS) + lane;

  uint4 rec[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
    { const uint32_t idx = base + i * 32;
      rec[i] = (idx < n) ? __ldcs(in + idx) : make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    }
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
    atomicAdd(&wh[rec_digit<BITS>(rec[i], word, shift)], 1u);
  __syncthreads();                                         // (A)

  uint32_t cnt[DPT], dbase[DPT];
  { uint32_t tsum = 0;
#pragma unroll
    for (int j = 0; j < DPT; j++)
      { const int d = tid * DPT + j;
        uint32_t run = 0;
        if (d < NB)
          {
#pragma unroll
            for (int w = 0; w < WARPS; w++) run += whist[w * NB + d];
            if (d == NB - 1) run -= (uint32_t) TILE - nvalid;
            st_relaxed(tile_state + (size_t) tile * NB + d, (tile == 0 ? FLAG_INC : FLAG_AGG) | run);
          }
        cnt[j] = run; tsum += run;
      }
    uint32_t x = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_wsum[warp] = x;
    __syncthreads();                                       // (B1)
    uint32_t add = 0;
#pragma unroll
    for (int w = 0; w < WARPS; w++) if (w < warp) add += s_wsum[w];
    uint32_t e = x + add - tsum;
#pragma unroll
    for (int j = 0; j < DPT; j++)
      { const int d = tid * DPT + j;
        dbase[j] = e;
        if (d < NB)
          { uint32_t run = e;                              // per-warp counters become staged start positions
#pragma unroll
            for (int w = 0; w < WARPS; w++) { uint32_t c = whist[w * NB + d]; whist[w * NB + d] = run; run += c; }
          }
        e += cnt[j];
      }
  }
  __syncthreads();                                         // (B2)

  uint32_t pos[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
    { const uint32_t dig = rec_digit<BITS>(rec[i], word, shift);
      const uint32_t peers = match_ballot<BITS>(dig);
      uint32_t old = 0;
      if ((peers & lt) == 0) old = atomicAdd(&wh[dig], (uint32_t) __popc(peers));
      pos[i] = __shfl_sync(0xffffffffu, old, __ffs(peers) - 1) + __popc(peers & lt);
    }
  __syncthreads();                                         // (C) counters die, staging live
#pragma unroll
  for (int i = 0; i < ITEMS; i++) stage[pos[i]] = rec[i];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

  uint32_t gdst[DPT];
#pragma unroll
  for (int j = 0; j < DPT; j++)
    { const int d = tid * DPT + j;
      if (d < NB)
        { uint32_t excl = 0;
          if (tile > 0)
            { const uint32_t *col = tile_state + d;
              int64_t t = (int64_t) tile - 1;
              bool done = false;
              while (!done)
                { uint32_t v[PROBE];
#pragma unroll
                  for (int k = 0; k < PROBE; k++)
                    v[k] = (t - k >= 0) ? ld_relaxed(col + (size_t) (t - k) * NB) : FLAG_INC;
                  int used = PROBE;
#pragma unroll
                  for (int k = 0; k < PROBE; k++)
                    if (!done && used == PROBE)
                      { if (v[k] & FLAG_INC)      { excl += v[k] & VAL_MASK; done = true; }
                        else if (v[k] & FLAG_AGG) excl += v[k] & VAL_MASK;
                        else                      used = k;
                      }
                  t -= used;
                }
              st_relaxed(tile_state + (size_t) tile * NB + d, FLAG_INC | (excl + cnt[j]));
            }
          gdst[j] = gbase[d] + excl;
        }
    }
  __syncthreads();                                         // (D)
#pragma unroll
  for (int j = 0; j < DPT; j++)
    { const int d = tid * DPT + j;
      if (d < NB && cnt[j] > 0)
        { const uint32_t sa = (uint32_t) __cvta_generic_to_shared(stage + dbase[j]);
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                       :: "l"(out + gdst[j]), "r"(sa), "r"(cnt[j] * 16u) : "memory");
        }
    }
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ---------------- V8 (v5 + matcher choice): early counts + ballot ranking straight to staged positions + probes + TMA store ----------------
template <int THREADS, int ITEMS, int BITS, int MINB, int PROBE, int MM>
__global__ void __launch_bounds__(THREADS, MINB)
k_pass_v8(const uint4 *__restrict__ in, uint4 *__restrict__ out, uint32_t n, int word, int shift,
          const uint32_t *__restrict__ gbase, uint32_t *tile_state, uint32_t *tile_counter)
{ constexpr int NB = 1 << BITS, WARPS = THREADS / 32, TILE = THREADS * ITEMS;
  constexpr int DPT = (NB + THREADS - 1) / THREADS;
  extern __shared__ uint4 dyn[];
  uint4    *stage = dyn;                                   // aliases the counters
  uint32_t *whist = reinterpret_cast<uint32_t *>(dyn);     // [WARPS][NB]
  __shared__ uint32_t s_wsum[WARPS];
  __shared__ uint32_t s_tile;
  __shared__ uint32_t s_tag[MM == 2 ? WARPS * 256 : 1];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lt = (1u << lane) - 1;
  if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
  uint32_t *wh = whist + warp * NB;
#pragma unroll
  for (int i = lane; i < NB; i += 32) wh[i] = 0;
  __syncthreads();
  const uint32_t tile  = s_tile;
  const uint32_t tbase = tile * (uint32_t) TILE;
  const uint32_t nvalid = (n - tbase < (uint32_t) TILE) ? n - tbase : (uint32_t) TILE;
  const uint32_t base = tbase + warp * (32 * ITEMS) + lane;

  uint4 rec[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
    { const uint32_t idx = base + i * 32;
      rec[i] = (idx < n) ? __ldcs(in + idx) : make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    }
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
    atomicAdd(&wh[rec_digit<BITS>(rec[i], word, shift)], 1u);
  __syncthreads();                                         // (A)

  uint32_t cnt[DPT], dbase[DPT];
  { uint32_t tsum = 0;
#pragma unroll
    for (int j = 0; j < DPT; j++)
      { const int d = tid * DPT + j;
        uint32_t run = 0;
        if (d < NB)
          {
#pragma unroll
            for (int w = 0; w < WARPS; w++) run += whist[w * NB + d];
            if (d == NB - 1) run -= (uint32_t) TILE - nvalid;
            st_relaxed(tile_state + (size_t) tile * NB + d, (tile == 0 ? FLAG_INC : FLAG_AGG) | run);
          }
        cnt[j] = run; tsum += run;
      }
    uint32_t x = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_wsum[warp] = x;
    __syncthreads();                                       // (B1)
    uint32_t add = 0;
#pragma unroll
    for (int w = 0; w < WARPS; w++) if (w < warp) add += s_wsum[w];
    uint32_t e = x + add - tsum;
#pragma unroll
    for (int j = 0; j < DPT; j++)
      { const int d = tid * DPT + j;
        dbase[j] = e;
        if (d < NB)
          { uint32_t run = e;                              // per-warp counters become staged start positions
#pragma unroll
            for (int w = 0; w < WARPS; w++) { uint32_t c = whist[w * NB + d]; whist[w * NB + d] = run; run += c; }
          }
        e += cnt[j];
      }
  }
  __syncthreads();                                         // (B2)

  uint32_t pos[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
    { const uint32_t dig = rec_digit<BITS>(rec[i], word, shift);
      const uint32_t peers = (MM == 2) ? match_table(dig, s_tag + warp * 256, lane) : (MM == 1) ? match_ballot2<BITS>(dig) : match_ballot<BITS>(dig);
      uint32_t old = 0;
      if ((peers & lt) == 0) old = atomicAdd(&wh[dig], (uint32_t) __popc(peers));
      pos[i] = __shfl_sync(0xffffffffu, old, __ffs(peers) - 1) + __popc(peers & lt);
    }
  __syncthreads();                                         // (C) counters die, staging live
#pragma unroll
  for (int i = 0; i < ITEMS; i++) stage[pos[i]] = rec[i];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

  uint32_t gdst[DPT];
#pragma unroll
  for (int j = 0; j < DPT; j++)
    { const int d = tid * DPT + j;
      if (d < NB)
        { uint32_t excl = 0;
          if (tile > 0)
            { const uint32_t *col = tile_state + d;
              int64_t t = (int64_t) tile - 1;
              bool done = false;
              while (!done)
                { uint32_t v[PROBE];
#pragma unroll
                  for (int k = 0; k < PROBE; k++)
                    v[k] = (t - k >= 0) ? ld_relaxed(col + (size_t) (t - k) * NB) : FLAG_INC;
                  int used = PROBE;
#pragma unroll
                  for (int k = 0; k < PROBE; k++)
                    if (!done && used == PROBE)
                      { if (v[k] & FLAG_INC)      { excl += v[k] & VAL_MASK; done = true; }
                        else if (v[k] & FLAG_AGG) excl += v[k] & VAL_MASK;
                        else                      used = k;
                      }
                  t -= used;
                }
              st_relaxed(tile_state + (size_t) tile * NB + d, FLAG_INC | (excl + cnt[j]));
            }
          gdst[j] = gbase[d] + excl;
        }
    }
  __syncthreads();                                         // (D)
#pragma unroll
  for (int j = 0; j < DPT; j++)
    { const int d = tid * DPT + j;
      if (d < NB && cnt[j] > 0)
        { const uint32_t sa = (uint32_t) __cvta_generic_to_shared(stage + dbase[j]);
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                       :: "l"(out + gdst[j]), "r"(sa), "r"(cnt[j] * 16u) : "memory");
        }
    }
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ---------------- V6 (blockIdx tile ids, ABL): early counts + ballot ranking straight to staged positions + probes + TMA store ----------------
template <int THREADS, int ITEMS, int BITS, int MINB, int PROBE, int ABL>
__global__ void __launch_bounds__(THREADS, MINB)
k_pass_v6(const uint4 *__restrict__ in, uint4 *__restrict__ out, uint32_t n, int word, int shift,
          const uint32_t *__restrict__ gbase, uint32_t *tile_state, uint32_t *tile_counter, unsigned long long *tm)
{ constexpr int NB = 1 << BITS, WARPS = THREADS / 32, TILE = THREADS * ITEMS;
  constexpr int DPT = (NB + THREADS - 1) / THREADS;
  extern __shared__ uint4 dyn[];
  uint4    *stage = dyn;                                   // aliases the counters
  uint32_t *whist = reinterpret_cast<uint32_t *>(dyn);     // [WARPS][NB]
  __shared__ uint32_t s_wsum[WARPS];
  __shared__ uint32_t s_tile;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  long long tk[9]; int nk = 0;
#define MARK() do { if ((ABL & 4) && tid == 0) tk[nk++] = clock64(); } while (0)
  MARK();
  const uint32_t lt = (1u << lane) - 1;
  uint32_t *wh = whist + warp * NB;
#pragma unroll
  for (int i = lane; i < NB; i += 32) wh[i] = 0;
  __syncwarp();
  const uint32_t tile  = blockIdx.x;
  const uint32_t tbase = tile * (uint32_t) TILE;
  const uint32_t nvalid = (n - tbase < (uint32_t) TILE) ? n - tbase : (uint32_t) TILE;
  const uint32_t base = tbase + warp * (32 * ITEMS) + lane;

  uint4 rec[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
    { const uint32_t idx = base + i * 32;
      rec[i] = (idx < n) ? __ldcs(in + idx) : make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    }
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
    atomicAdd(&wh[rec_digit<BITS>(rec[i], word, shift)], 1u);
  __syncthreads();                                         // (A)
  MARK();

  uint32_t cnt[DPT], dbase[DPT];
  { uint32_t tsum = 0;
#pragma unroll
    for (int j = 0; j < DPT; j++)
      { const int d = tid * DPT + j;
        uint32_t run = 0;
        if (d < NB)
          {
#pragma unroll
            for (int w = 0; w < WARPS; w++) run += whist[w * NB + d];
            if (d == NB - 1) run -= (uint32_t) TILE - nvalid;
            st_relaxed(tile_state + (size_t) tile * NB + d, (tile == 0 ? FLAG_INC : FLAG_AGG) | run);
          }
        cnt[j] = run; tsum += run;
      }
    uint32_t x = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_wsum[warp] = x;
    __syncthreads();                                       // (B1)
    uint32_t add = 0;
#pragma unroll
    for (int w = 0; w < WARPS; w++) if (w < warp) add += s_wsum[w];
    uint32_t e = x + add - tsum;
#pragma unroll
    for (int j = 0; j < DPT; j++)
      { const int d = tid * DPT + j;
        dbase[j] = e;
        if (d < NB)
          { uint32_t run = e;                              // per-warp counters become staged start positions
#pragma unroll
            for (int w = 0; w < WARPS; w++) { uint32_t c = whist[w * NB + d]; whist[w * NB + d] = run; run += c; }
          }
        e += cnt[j];
      }
  }
  __syncthreads();                                         // (B2)
  MARK();

  uint32_t pos[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
    { const uint32_t dig = rec_digit<BITS>(rec[i], word, shift);
      const uint32_t peers = (ABL & 2) ? (1u << lane) : match_ballot<BITS>(dig);
      uint32_t old = 0;
      if ((peers & lt) == 0) old = atomicAdd(&wh[dig], (uint32_t) __popc(peers));
      pos[i] = __shfl_sync(0xffffffffu, old, __ffs(peers) - 1) + __popc(peers & lt);
    }
  __syncthreads();                                         // (C) counters die, staging live
  MARK();
#pragma unroll
  for (int i = 0; i < ITEMS; i++) stage[pos[i]] = rec[i];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

  MARK();
  uint32_t gdst[DPT];
#pragma unroll
  for (int j = 0; j < DPT; j++)
    { const int d = tid * DPT + j;
      if (d < NB)
        { uint32_t excl = 0;
          if (tile > 0 && !(ABL & 1))
            { const uint32_t *col = tile_state + d;
              int64_t t = (int64_t) tile - 1;
              bool done = false;
              while (!done)
                { uint32_t v[PROBE];
#pragma unroll
                  for (int k = 0; k < PROBE; k++)
                    v[k] = (t - k >= 0) ? ld_relaxed(col + (size_t) (t - k) * NB) : FLAG_INC;
                  int used = PROBE;
#pragma unroll
                  for (int k = 0; k < PROBE; k++)
                    if (!done && used == PROBE)
                      { if (v[k] & FLAG_INC)      { excl += v[k] & VAL_MASK; done = true; }
                        else if (v[k] & FLAG_AGG) excl += v[k] & VAL_MASK;
                        else                      used = k;
                      }
                  t -= used;
                }
              st_relaxed(tile_state + (size_t) tile * NB + d, FLAG_INC | (excl + cnt[j]));
            }
          gdst[j] = gbase[d] + excl;
        }
    }
  MARK();
  __syncthreads();                                         // (D)
  MARK();
#pragma unroll
  for (int j = 0; j < DPT; j++)
    { const int d = tid * DPT + j;
      if (d < NB && cnt[j] > 0)
        { const uint32_t sa = (uint32_t) __cvta_generic_to_shared(stage + dbase[j]);
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                       :: "l"(out + gdst[j]), "r"(sa), "r"(cnt[j] * 16u) : "memory");
        }
    }
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  MARK();
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  MARK();
  if ((ABL & 4) && tid == 0)
    for (int k = 1; k < nk; k++) atomicAdd(tm + k, (unsigned long long) (tk[k] - tk[k - 1]));
#undef MARK
}

// cost of the two matching primitives in isolation: cycles per warp-level match at full occupancy
template <int MODE>
__global__ void k_matchbench(uint32_t *outp, int iters)
{ uint32_t x = mix(blockIdx.x * 1024 + threadIdx.x), acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++)
    { uint32_t dig = x & 255;
      uint32_t m = MODE ? match_ballot<8>(dig) : __match_any_sync(0xffffffffu, dig);
      acc += m;
      x = x * 1664525u + 1013904223u + (m & 1);
      x ^= x >> 13;
    }
  long long t1 = clock64();
  if (acc == 0x12345) outp[0] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) outp[1 + MODE] = (uint32_t) ((t1 - t0) / iters);
}

// ---------------- skeleton study: staged copy, no ranking, no look-back ----------------
// MODE 0: STG, identity   1: TMA runs, identity   2: TMA runs, 256-way scatter   3: STG, 256-way scatter
// MODE 4: direct copy through registers (no smem), tile-structured
template <int THREADS, int ITEMS, int MODE>
__global__ void __launch_bounds__(THREADS)
k_stage_copy(const uint4 *__restrict__ in, uint4 *__restrict__ out, uint32_t n, uint32_t ntiles)
{ constexpr int TILE = THREADS * ITEMS, RUN = TILE / 256;
  extern __shared__ uint4 stage[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t tile = blockIdx.x, tbase = tile * (uint32_t) TILE;
  const uint32_t base = tbase + warp * (32 * ITEMS) + lane;
  uint4 rec[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
    { const uint32_t idx = base + i * 32;
      rec[i] = (idx < n) ? __ldcs(in + idx) : make_uint4(0, 0, 0, 0);
    }
  if (MODE == 4)
    {
#pragma unroll
      for (int i = 0; i < ITEMS; i++)
        { const uint32_t idx = base + i * 32;
          if (idx < n) __stcs(out + idx, rec[i]);
        }
      return;
    }
#pragma unroll
  for (int i = 0; i < ITEMS; i++)
    stage[warp * (32 * ITEMS) + i * 32 + lane] = rec[i];
  if (MODE == 1 || MODE == 2) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const uint32_t per = (ntiles * (uint32_t) RUN);          // records per digit stream
  if (MODE == 0 || MODE == 3)
    {
#pragma unroll
      for (int i = 0; i < ITEMS; i++)
        { const uint32_t j = tid + i * THREADS;
          const uint32_t d = j / RUN, o = j % RUN;
          const uint32_t dst = (MODE == 0) ? tbase + j : d * per + tile * RUN + o;
          if (dst < n) __stcs(out + dst, stage[j]);
        }
    }
  else
    { if (tid < 256)
        { const uint32_t d = tid;
          const uint32_t dst = (MODE == 1) ? tbase + d * RUN : d * per + tile * RUN;
          if (dst + RUN <= n)
            { const uint32_t sa = (uint32_t) __cvta_generic_to_shared(stage + d * RUN);
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                           :: "l"(out + dst), "r"(sa), "r"((uint32_t) RUN * 16u) : "memory");
            }
        }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

// ---------------- driver ----------------
struct Ctx { uint4 *src, *a, *b; uint32_t n; uint32_t *hist, *state; unsigned long long *scal; int keybits; unsigned long long refsum; };

template <int THREADS, int ITEMS, int MODE>
static void run_stage(const char *name, Ctx &c)
{ constexpr int TILE = THREADS * ITEMS;
  const uint32_t ntiles = c.n / TILE;                      // whole tiles only
  const uint32_t n = ntiles * TILE;
  const size_t smem = (size_t) TILE * 16;
  auto kern = k_stage_copy<THREADS, ITEMS, MODE>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  float best = 1e30f;
  for (int r = 0; r < 5; r++)
    { cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0); kern<<<ntiles, THREADS, smem>>>(c.src, c.a, n, ntiles); cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
      float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms);
    }
  int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem);
  printf("%-44s %.3f ms  %.0f GB/s (%.1f%%)  [%d CTAs/SM]\n", name, best, 32.0 * n / 1e6 / best, 100 * 32.0 * n / 1e6 / best / 6544.7, occ);
  fflush(stdout);
}


static float run_timed(const char *name, Ctx &c, int bits, int npass, size_t tile, bool verify,
                       void (*launch)(Ctx &, const uint4 *, uint4 *, int word, int shift, uint32_t ntiles, int p))
{ const uint32_t ntiles = (uint32_t) ((c.n + tile - 1) / tile);
  const int nb = 1 << bits;
  // histograms of every pass from the source
  CK(cudaMemset(c.hist, 0, sizeof(uint32_t) * nb * npass));
  for (int p = 0; p < npass; p++)
    { if (bits == 8) k_hist<8><<<148 * 8, 256>>>(c.src, c.n, 0, p * bits, c.hist + p * nb);
      else if (bits == 10) k_hist<10><<<148 * 8, 256>>>(c.src, c.n, 0, p * bits, c.hist + p * nb);
      else k_hist<11><<<148 * 8, 256>>>(c.src, c.n, 0, p * bits, c.hist + p * nb);
      k_exscan<<<1, 1>>>(c.hist + p * nb, nb);
    }
  float best = 1e30f, tot = 0;
  const int reps = 4;
  for (int rep = 0; rep < reps; rep++)
    { CK(cudaMemcpy(c.a, c.src, (size_t) c.n * 16, cudaMemcpyDeviceToDevice));
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      const uint4 *s = c.a; uint4 *d = c.b;
      cudaEventRecord(e0);
      for (int p = 0; p < npass; p++)
        { CK(cudaMemsetAsync(c.state, 0, sizeof(uint32_t) * ((size_t) ntiles * nb + 1)));
          launch(c, s, d, 0, p * bits, ntiles, p);
          const uint4 *t = s; s = d; d = (uint4 *) t;
        }
      cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1));
      CK(cudaGetLastError());
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0) { best = std::min(best, ms); tot += ms; }
      if (rep == reps - 1 && verify)
        { CK(cudaMemset(c.scal, 0, 16));
          k_check<<<148 * 8, 256>>>(s, c.n, std::min(npass * bits, 64), c.scal);
          k_sum<<<148 * 8, 256>>>(s, c.n, c.scal + 1);
          unsigned long long h[2]; CK(cudaMemcpy(h, c.scal, 16, cudaMemcpyDeviceToHost));
          printf("  [%s] check: bad pairs %llu, multiset %s\n", name, h[0], h[1] == c.refsum ? "ok" : "MISMATCH");
        }
      cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
  const double gb = 32.0 * c.n * npass / 1e6;
  printf("%-34s %d pass(es) x %2d bits: best %.3f ms  mean %.3f ms  = %.3f ms/pass  %.0f GB/s/pass (%.1f%% of 6544.7)\n",
         name, npass, bits, best, tot / (reps - 1), best / npass, gb / best, 100 * gb / best / 6544.7);
  fflush(stdout);
  return best;
}

template <int ABL> static void L_v0(Ctx &c, const uint4 *s, uint4 *d, int word, int shift, uint32_t ntiles, int p)
{ k_pass_v0<ABL><<<ntiles, 256>>>(s, d, c.n, word, shift, c.hist + p * 256, c.state, c.state + (size_t) ntiles * 256); }

template <int THREADS, int ITEMS, int BITS, bool ALIAS, int MINB>
static void L_v1(Ctx &c, const uint4 *s, uint4 *d, int word, int shift, uint32_t ntiles, int p)
{ constexpr int NB = 1 << BITS;
  const size_t smem = (size_t) THREADS * ITEMS * 16 + (ALIAS ? 0 : (size_t) (THREADS / 32) * NB * 4);
  static bool set = false;
  if (!set) { CK(cudaFuncSetAttribute(k_pass_v1<THREADS, ITEMS, BITS, ALIAS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)); set = true;
              int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_pass_v1<THREADS, ITEMS, BITS, ALIAS, MINB>, THREADS, smem);
              cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_pass_v1<THREADS, ITEMS, BITS, ALIAS, MINB>);
              printf("  v1<%d,%d,%d,%d,%d>: %d regs, %zu B dyn smem, %d CTAs/SM\n", THREADS, ITEMS, BITS, (int) ALIAS, MINB, fa.numRegs, smem, occ); }
  k_pass_v1<THREADS, ITEMS, BITS, ALIAS, MINB><<<ntiles, THREADS, smem>>>(s, d, c.n, word, shift, c.hist + p * NB, c.state, c.state + (size_t) ntiles * NB);
}

template <int THREADS, int ITEMS, int BITS, int MATCHMODE, bool STAGED, int MINB>
static void L_v2(Ctx &c, const uint4 *s, uint4 *d, int word, int shift, uint32_t ntiles, int p)
{ constexpr int NB = 1 << BITS;
  const size_t smem = std::max((size_t) (STAGED ? THREADS * ITEMS * 16 : 0), (size_t) (THREADS / 32) * NB * 4);
  static bool set = false;
  auto kern = k_pass_v2<THREADS, ITEMS, BITS, MATCHMODE, STAGED, MINB>;
  if (!set) { CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)); set = true;
              int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem);
              cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
              printf("  v2<%d,%d,%d,%d,%d,%d>: %d regs, %zu B dyn smem, %d CTAs/SM\n", THREADS, ITEMS, BITS, MATCHMODE, (int) STAGED, MINB, fa.numRegs, smem, occ); }
  kern<<<ntiles, THREADS, smem>>>(s, d, c.n, word, shift, c.hist + p * NB, c.state, c.state + (size_t) ntiles * NB);
}

template <int THREADS, int ITEMS, int BITS, int MATCHMODE, int MINB, bool TMAST>
static void L_v3(Ctx &c, const uint4 *s, uint4 *d, int word, int shift, uint32_t ntiles, int p)
{ constexpr int NB = 1 << BITS;
  const size_t smem = std::max((size_t) (THREADS * ITEMS * 16), (size_t) (THREADS / 32) * NB * 4);
  static bool set = false;
  auto kern = k_pass_v3<THREADS, ITEMS, BITS, MATCHMODE, true, MINB, TMAST>;
  if (!set) { CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)); set = true;
              int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem);
              cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
              printf("  v3<%d,%d,%d,%d,%d,tma=%d>: %d regs, %zu B dyn smem, %d CTAs/SM\n", THREADS, ITEMS, BITS, MATCHMODE, MINB, (int) TMAST, fa.numRegs, smem, occ); }
  kern<<<ntiles, THREADS, smem>>>(s, d, c.n, word, shift, c.hist + p * NB, c.state, c.state + (size_t) ntiles * NB);
}

template <int THREADS, int ITEMS, int BITS, int MINB, bool TMAST, int PROBE>
static void L_v4(Ctx &c, const uint4 *s, uint4 *d, int word, int shift, uint32_t ntiles, int p)
{ constexpr int NB = 1 << BITS;
  const size_t smem = std::max((size_t) (THREADS * ITEMS * 16), (size_t) (THREADS / 32) * NB * 4);
  static bool set = false;
  auto kern = k_pass_v4<THREADS, ITEMS, BITS, 1, true, MINB, TMAST, PROBE>;
  if (!set) { CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)); set = true;
              int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem);
              cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
              printf("  v4<%d,%d,%d,%d,tma=%d,probe=%d>: %d regs, %zu B dyn smem, %d CTAs/SM\n", THREADS, ITEMS, BITS, MINB, (int) TMAST, PROBE, fa.numRegs, smem, occ); }
  kern<<<ntiles, THREADS, smem>>>(s, d, c.n, word, shift, c.hist + p * NB, c.state, c.state + (size_t) ntiles * NB);
}

template <int THREADS, int ITEMS, int BITS, int MINB, int PROBE>
static void L_v5(Ctx &c, const uint4 *s, uint4 *d, int word, int shift, uint32_t ntiles, int p)
{ constexpr int NB = 1 << BITS;
  const size_t smem = std::max((size_t) (THREADS * ITEMS * 16), (size_t) (THREADS / 32) * NB * 4);
  static bool set = false;
  auto kern = k_pass_v5<THREADS, ITEMS, BITS, MINB, PROBE>;
  if (!set) { CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)); set = true;
              int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem);
              cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
              printf("  v5<%d,%d,%d,%d,probe=%d>: %d regs, %zu B dyn smem, %d CTAs/SM\n", THREADS, ITEMS, BITS, MINB, PROBE, fa.numRegs, smem, occ); }
  kern<<<ntiles, THREADS, smem>>>(s, d, c.n, word, shift, c.hist + p * NB, c.state, c.state + (size_t) ntiles * NB);
}

template <int THREADS, int ITEMS, int BITS, int MINB, int PROBE, int ABL>
static void L_v6(Ctx &c, const uint4 *s, uint4 *d, int word, int shift, uint32_t ntiles, int p)
{ constexpr int NB = 1 << BITS;
  const size_t smem = std::max((size_t) (THREADS * ITEMS * 16), (size_t) (THREADS / 32) * NB * 4);
  static bool set = false;
  auto kern = k_pass_v6<THREADS, ITEMS, BITS, MINB, PROBE, ABL>;
  if (!set) { CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)); set = true; }
  kern<<<ntiles, THREADS, smem>>>(s, d, c.n, word, shift, c.hist + p * NB, c.state, c.state + (size_t) ntiles * NB, c.scal + 8);
}

template <int THREADS, int ITEMS, int BITS, int MINB, bool TMAST, int PROBE>
static void L_v7(Ctx &c, const uint4 *s, uint4 *d, int word, int shift, uint32_t ntiles, int p)
{ constexpr int NB = 1 << BITS;
  const size_t smem = std::max((size_t) (THREADS * ITEMS * 16), (size_t) (THREADS / 32) * NB * 4);
  static bool set = false;
  auto kern = k_pass_v7<THREADS, ITEMS, BITS, 1, true, MINB, TMAST, PROBE>;
  if (!set) { CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)); set = true;
              int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem);
              cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
              printf("  v7<%d,%d,%d,%d,tma=%d,probe=%d>: %d regs, %zu B dyn smem, %d CTAs/SM\n", THREADS, ITEMS, BITS, MINB, (int) TMAST, PROBE, fa.numRegs, smem, occ); }
  kern<<<ntiles, THREADS, smem>>>(s, d, c.n, word, shift, c.hist + p * NB, c.state, c.state + (size_t) ntiles * NB);
}

template <int THREADS, int ITEMS, int BITS, int MINB, int PROBE, int MM>
static void L_v8(Ctx &c, const uint4 *s, uint4 *d, int word, int shift, uint32_t ntiles, int p)
{ constexpr int NB = 1 << BITS;
  const size_t smem = std::max((size_t) (THREADS * ITEMS * 16), (size_t) (THREADS / 32) * NB * 4);
  static bool set = false;
  auto kern = k_pass_v8<THREADS, ITEMS, BITS, MINB, PROBE, MM>;
  if (!set) { CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)); set = true;
              int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem);
              cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
              printf("  v8<%d,%d,%d,%d,probe=%d,mm=%d>: %d regs, %zu B dyn smem, %d CTAs/SM\n", THREADS, ITEMS, BITS, MINB, PROBE, MM, fa.numRegs, smem, occ); }
  kern<<<ntiles, THREADS, smem>>>(s, d, c.n, word, shift, c.hist + p * NB, c.state, c.state + (size_t) ntiles * NB);
}

template <int THREADS, int ITEMS, int MINB, int PROBE>
static void L_v9(Ctx &c, const uint4 *s, uint4 *d, int word, int shift, uint32_t ntiles, int p)
{ const size_t smem = (size_t) THREADS * ITEMS * 16;
  const int byte = word * 8 + shift / 8, w32 = byte >> 2;
  const uint32_t psel = 0x4440u | (uint32_t) (byte & 3);
  static bool set = false;
  if (!set)
    { CK(cudaFuncSetAttribute(k_pass_v9<THREADS, ITEMS, MINB, PROBE, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
      CK(cudaFuncSetAttribute(k_pass_v9<THREADS, ITEMS, MINB, PROBE, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
      set = true;
      int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_pass_v9<THREADS, ITEMS, MINB, PROBE, 0>, THREADS, smem);
      cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_pass_v9<THREADS, ITEMS, MINB, PROBE, 0>);
      printf("  v9<%d,%d,%d,probe=%d>: %d regs, %zu B dyn smem, %d CTAs/SM\n", THREADS, ITEMS, MINB, PROBE, fa.numRegs, smem, occ);
    }
  if (w32 == 0) k_pass_v9<THREADS, ITEMS, MINB, PROBE, 0><<<ntiles, THREADS, smem>>>(s, d, c.n, psel, c.hist + p * 256, c.state, c.state + (size_t) ntiles * 256);
  else          k_pass_v9<THREADS, ITEMS, MINB, PROBE, 1><<<ntiles, THREADS, smem>>>(s, d, c.n, psel, c.hist + p * 256, c.state, c.state + (size_t) ntiles * 256);
}

template <int THREADS, int ITEMS, int MINB, int PROBE, int MIX, int LH>
static void L_v10(Ctx &c, const uint4 *s, uint4 *d, int word, int shift, uint32_t ntiles, int p)
{ const size_t smem = (size_t) THREADS * ITEMS * 16;
  const int byte = word * 8 + shift / 8, w32 = byte >> 2;
  const uint32_t psel = 0x4440u | (uint32_t) (byte & 3);
  static bool set = false;
  if (!set)
    { CK(cudaFuncSetAttribute(k_pass_v10<THREADS, ITEMS, MINB, PROBE, 0, MIX, LH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
      CK(cudaFuncSetAttribute(k_pass_v10<THREADS, ITEMS, MINB, PROBE, 1, MIX, LH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
      set = true;
    }
  if (w32 == 0) k_pass_v10<THREADS, ITEMS, MINB, PROBE, 0, MIX, LH><<<ntiles, THREADS, smem>>>(s, d, c.n, psel, c.hist + p * 256, c.state, c.state + (size_t) ntiles * 256);
  else          k_pass_v10<THREADS, ITEMS, MINB, PROBE, 1, MIX, LH><<<ntiles, THREADS, smem>>>(s, d, c.n, psel, c.hist + p * 256, c.state, c.state + (size_t) ntiles * 256);
}

int main(int argc, char **argv)
{ uint32_t n = argc > 1 ? (uint32_t) atoll(argv[1]) : 139813248u;
  const char *which = argc > 2 ? argv[2] : "all";
  Ctx c; c.n = n; c.keybits = 40;
  CK(cudaMalloc(&c.src, (size_t) (n + 2) * 16)); CK(cudaMalloc(&c.a, (size_t) (n + 2) * 16)); CK(cudaMalloc(&c.b, (size_t) (n + 2) * 16));
  CK(cudaMalloc(&c.hist, sizeof(uint32_t) * 2048 * 8));
  CK(cudaMalloc(&c.state, sizeof(uint32_t) * ((size_t) (n / 1024 + 2) * 256 + 64) * 4));
  CK(cudaMalloc(&c.scal, 256)); CK(cudaMemset(c.scal, 0, 256));
  k_fill<<<148 * 8, 256>>>(c.src, n, c.keybits, 12345);
  CK(cudaMemset(c.scal, 0, 16));
  k_sum<<<148 * 8, 256>>>(c.src, n, c.scal + 1);
  unsigned long long h[2]; CK(cudaMemcpy(h, c.scal, 16, cudaMemcpyDeviceToHost)); c.refsum = h[1];
  // plain copy for reference
  { float best = 1e30f;
    for (int r = 0; r < 5; r++)
      { cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0); k_copy<<<148 * 16, 512>>>(c.src, c.a, n); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms);
      }
    printf("plain copy kernel: %.3f ms = %.0f GB/s\n", best, 32.0 * n / 1e6 / best);
  }
  auto want = [&](const char *k)
    { if (!strcmp(which, "all")) return true;
      const char *p = which;
      while (*p)
        { const char *q = strchr(p, ','); size_t len = q ? (size_t) (q - p) : strlen(p);
          if (len > 0 && !strncmp(k, p, len)) return true;
          p += len; if (*p == ',') p++;
        }
      return false;
    };
  if (want("v0"))  run_timed("v0 committed 256x8", c, 8, 5, 2048, true, L_v0<0>);
  if (want("abl")) { run_timed("v0 no look-back (wrong)", c, 8, 1, 2048, false, L_v0<1>);
                     run_timed("v0 no ranking (wrong)", c, 8, 1, 2048, false, L_v0<2>);
                     run_timed("v0 neither (wrong)", c, 8, 1, 2048, false, L_v0<3>); }
  if (want("v1a")) run_timed("v1 256x8 alias", c, 8, 5, 2048, true, L_v1<256, 8, 8, true, 1>);
  if (want("v1j")) run_timed("v1 256x8 alias minb3", c, 8, 5, 2048, true, L_v1<256, 8, 8, true, 3>);
  if (want("v1k")) run_timed("v1 256x8 alias minb4", c, 8, 5, 2048, true, L_v1<256, 8, 8, true, 4>);
  if (want("v1l")) run_timed("v1 512x8 alias minb2", c, 8, 5, 4096, true, L_v1<512, 8, 8, true, 2>);
  if (want("v1m")) run_timed("v1 128x8 alias minb8", c, 8, 5, 1024, true, L_v1<128, 8, 8, true, 8>);
  if (want("v1b")) run_timed("v1 256x8 noalias", c, 8, 5, 2048, true, L_v1<256, 8, 8, false, 1>);
  if (want("v1c")) run_timed("v1 512x8 alias", c, 8, 5, 4096, true, L_v1<512, 8, 8, true, 1>);
  if (want("v1d")) run_timed("v1 256x12 alias", c, 8, 5, 3072, true, L_v1<256, 12, 8, true, 1>);
  if (want("v1e")) run_timed("v1 256x16 alias", c, 8, 5, 4096, true, L_v1<256, 16, 8, true, 1>);
  if (want("v1f")) run_timed("v1 384x8 alias", c, 8, 5, 3072, true, L_v1<384, 8, 8, true, 1>);
  if (want("v1g")) run_timed("v1 512x8 10bit alias", c, 10, 4, 4096, true, L_v1<512, 8, 10, true, 1>);
  if (want("v1h")) run_timed("v1 512x16 10bit alias", c, 10, 4, 8192, true, L_v1<512, 16, 10, true, 1>);
  if (want("v1i")) run_timed("v1 1024x8 10bit alias", c, 10, 4, 8192, true, L_v1<1024, 8, 10, true, 1>);
  if (want("mb"))
    { uint32_t *o; CK(cudaMalloc(&o, 64)); CK(cudaMemset(o, 0, 64));
      k_matchbench<0><<<148 * 2, 1024>>>(o, 2000); k_matchbench<1><<<148 * 2, 1024>>>(o, 2000);
      uint32_t h3[3]; CK(cudaMemcpy(h3, o, 12, cudaMemcpyDeviceToHost));
      printf("match bench (64 warps/SM): match_any %u cycles/iter/warp, ballot-match %u cycles/iter/warp\n", h3[1], h3[2]);
    }
  if (want("v2a")) run_timed("v2 512x8 match staged", c, 8, 5, 4096, true, L_v2<512, 8, 8, 0, true, 2>);
  if (want("v2b")) run_timed("v2 512x8 ballot staged", c, 8, 5, 4096, true, L_v2<512, 8, 8, 1, true, 2>);
  if (want("v2c")) run_timed("v2 512x8 match direct", c, 8, 5, 4096, true, L_v2<512, 8, 8, 0, false, 2>);
  if (want("v2d")) run_timed("v2 512x8 ballot direct", c, 8, 5, 4096, true, L_v2<512, 8, 8, 1, false, 2>);
  if (want("v2e")) run_timed("v2 256x8 ballot direct", c, 8, 5, 2048, true, L_v2<256, 8, 8, 1, false, 4>);
  if (want("v2f")) run_timed("v2 256x8 ballot staged", c, 8, 5, 2048, true, L_v2<256, 8, 8, 1, true, 4>);
  if (want("v2g")) run_timed("v2 1024x8 ballot direct", c, 8, 5, 8192, true, L_v2<1024, 8, 8, 1, false, 1>);
  if (want("v2h")) run_timed("v2 512x4 ballot direct", c, 8, 5, 2048, true, L_v2<512, 4, 8, 1, false, 3>);
  if (want("v2i")) run_timed("v2 1024x8 10b ballot direct", c, 10, 4, 8192, true, L_v2<1024, 8, 10, 1, false, 1>);
  if (want("v2j")) run_timed("v2 1024x8 10b ballot staged", c, 10, 4, 8192, true, L_v2<1024, 8, 10, 1, true, 1>);
  if (want("v2k")) run_timed("v2 512x8 10b ballot direct", c, 10, 4, 4096, true, L_v2<512, 8, 10, 1, false, 2>);
  if (want("v3a")) run_timed("v3 512x8 ballot staged tma-store", c, 8, 5, 4096, true, L_v3<512, 8, 8, 1, 2, true>);
  if (want("v3b")) run_timed("v3 1024x8 ballot staged", c, 8, 5, 8192, true, L_v3<1024, 8, 8, 1, 1, false>);
  if (want("v3c")) run_timed("v3 1024x8 ballot staged tma-store", c, 8, 5, 8192, true, L_v3<1024, 8, 8, 1, 1, true>);
  if (want("v3d")) run_timed("v3 256x16 ballot staged", c, 8, 5, 4096, true, L_v3<256, 16, 8, 1, 2, false>);
  if (want("v3e")) run_timed("v3 256x16 ballot staged tma-store", c, 8, 5, 4096, true, L_v3<256, 16, 8, 1, 2, true>);
  if (want("v3f")) run_timed("v3 512x16 ballot staged", c, 8, 5, 8192, true, L_v3<512, 16, 8, 1, 1, false>);
  if (want("v3g")) run_timed("v3 512x16 ballot staged tma-store", c, 8, 5, 8192, true, L_v3<512, 16, 8, 1, 1, true>);
  if (want("v3h")) run_timed("v3 1024x8 10b ballot staged tma-store", c, 10, 4, 8192, true, L_v3<1024, 8, 10, 1, 1, true>);
  if (want("v3i")) run_timed("v3 512x16 10b ballot staged tma-store", c, 10, 4, 8192, true, L_v3<512, 16, 10, 1, 1, true>);
  if (want("v3j")) run_timed("v3 384x8 ballot staged", c, 8, 5, 3072, true, L_v3<384, 8, 8, 1, 3, false>);
  if (want("v3k")) run_timed("v3 256x8 ballot staged tma-store", c, 8, 5, 2048, true, L_v3<256, 8, 8, 1, 4, true>);
  if (want("v3l")) run_timed("v3 512x12 ballot staged tma-store", c, 8, 5, 6144, true, L_v3<512, 12, 8, 1, 1, true>);
  if (want("v4a")) run_timed("v4 512x8 tma probe4", c, 8, 5, 4096, true, L_v4<512, 8, 8, 2, true, 4>);
  if (want("v4b")) run_timed("v4 512x8 tma probe8", c, 8, 5, 4096, true, L_v4<512, 8, 8, 2, true, 8>);
  if (want("v4c")) run_timed("v4 512x8 tma probe16", c, 8, 5, 4096, true, L_v4<512, 8, 8, 2, true, 16>);
  if (want("v4d")) run_timed("v4 384x8 tma probe8", c, 8, 5, 3072, true, L_v4<384, 8, 8, 3, true, 8>);
  if (want("v4e")) run_timed("v4 256x8 tma probe8", c, 8, 5, 2048, true, L_v4<256, 8, 8, 4, true, 8>);
  if (want("v4f")) run_timed("v4 256x16 tma probe8", c, 8, 5, 4096, true, L_v4<256, 16, 8, 2, true, 8>);
  if (want("v4g")) run_timed("v4 1024x8 tma probe8", c, 8, 5, 8192, true, L_v4<1024, 8, 8, 1, true, 8>);
  if (want("v4h")) run_timed("v4 512x8 stg probe8", c, 8, 5, 4096, true, L_v4<512, 8, 8, 2, false, 8>);
  if (want("v4i")) run_timed("v4 1024x8 10b tma probe8", c, 10, 4, 8192, true, L_v4<1024, 8, 10, 1, true, 8>);
  if (want("v4j")) run_timed("v4 512x8 10b tma probe8", c, 10, 4, 4096, true, L_v4<512, 8, 10, 2, true, 8>);
  if (want("v5a")) run_timed("v5 512x8 probe4", c, 8, 5, 4096, true, L_v5<512, 8, 8, 2, 4>);
  if (want("v5b")) run_timed("v5 512x8 probe8", c, 8, 5, 4096, true, L_v5<512, 8, 8, 2, 8>);
  if (want("v5c")) run_timed("v5 512x8 probe1", c, 8, 5, 4096, true, L_v5<512, 8, 8, 2, 1>);
  if (want("v5d")) run_timed("v5 384x8 probe4", c, 8, 5, 3072, true, L_v5<384, 8, 8, 3, 4>);
  if (want("v5e")) run_timed("v5 256x8 probe4", c, 8, 5, 2048, true, L_v5<256, 8, 8, 4, 4>);
  if (want("v5f")) run_timed("v5 256x16 probe4", c, 8, 5, 4096, true, L_v5<256, 16, 8, 2, 4>);
  if (want("v5g")) run_timed("v5 1024x8 probe4", c, 8, 5, 8192, true, L_v5<1024, 8, 8, 1, 4>);
  if (want("v5h")) run_timed("v5 1024x8 10b probe4", c, 10, 4, 8192, true, L_v5<1024, 8, 10, 1, 4>);
  if (want("v5i")) run_timed("v5 512x8 10b probe4", c, 10, 4, 4096, true, L_v5<512, 8, 10, 2, 4>);
  if (want("v5j")) run_timed("v5 512x6 probe4", c, 8, 5, 3072, true, L_v5<512, 6, 8, 2, 4>);
  if (want("v5k")) run_timed("v5 640x8 probe4", c, 8, 5, 5120, true, L_v5<640, 8, 8, 1, 4>);
  if (want("v6a")) run_timed("v6 512x8 probe4", c, 8, 5, 4096, true, L_v6<512, 8, 8, 2, 4, 0>);
  if (want("v6b")) run_timed("v6 384x8 probe4", c, 8, 5, 3072, true, L_v6<384, 8, 8, 3, 4, 0>);
  if (want("v6c")) run_timed("v6 256x16 probe4", c, 8, 5, 4096, true, L_v6<256, 16, 8, 2, 4, 0>);
  if (want("v6d")) run_timed("v6 256x8 probe4", c, 8, 5, 2048, true, L_v6<256, 8, 8, 4, 4, 0>);
  if (want("v6e")) run_timed("v6 384x8 no-lookback (wrong)", c, 8, 1, 3072, false, L_v6<384, 8, 8, 3, 4, 1>);
  if (want("v6f")) run_timed("v6 384x8 no-ballot (wrong)", c, 8, 1, 3072, false, L_v6<384, 8, 8, 3, 4, 2>);
  if (want("v6g")) run_timed("v6 384x8 neither (wrong)", c, 8, 1, 3072, false, L_v6<384, 8, 8, 3, 4, 3>);
  if (want("v6h")) run_timed("v6 384x12 probe4", c, 8, 5, 4608, true, L_v6<384, 12, 8, 2, 4, 0>);
  if (want("v6i")) run_timed("v6 320x8 probe4", c, 8, 5, 2560, true, L_v6<320, 8, 8, 3, 4, 0>);
  if (want("v6j")) run_timed("v6 256x12 probe4", c, 8, 5, 3072, true, L_v6<256, 12, 8, 3, 4, 0>);
  if (want("v6t"))
    { CK(cudaMemset(c.scal, 0, 256));
      run_timed("v6 384x8 timing", c, 8, 1, 3072, false, L_v6<384, 8, 8, 3, 4, 4>);
      unsigned long long t[9]; CK(cudaMemcpy(t, c.scal + 8, 72, cudaMemcpyDeviceToHost));
      const double nt = 4.0 * ((c.n + 3071) / 3072);
      const char *nm[9] = { "", "start->A (zero,load,hist,sync)", "A->B2 (scan,publish)", "B2->C (ballot rank, sync)", "C->staged (STS,fence)", "look-back", "sync D", "TMA issue", "TMA read wait" };
      for (int k = 1; k < 9; k++) printf("    %-34s %8.0f cycles/CTA\n", nm[k], t[k] / nt);
    }
  if (want("sk"))
    { run_stage<384, 12, 4>("sk 384x12 direct regs copy", c);
      run_stage<384, 12, 0>("sk 384x12 staged STG identity", c);
      run_stage<384, 12, 1>("sk 384x12 staged TMA identity", c);
      run_stage<384, 12, 2>("sk 384x12 staged TMA scatter256", c);
      run_stage<384, 12, 3>("sk 384x12 staged STG scatter256", c);
      run_stage<256, 8, 4>("sk 256x8 direct regs copy", c);
      run_stage<256, 8, 0>("sk 256x8 staged STG identity", c);
      run_stage<256, 8, 2>("sk 256x8 staged TMA scatter256", c);
      run_stage<512, 8, 2>("sk 512x8 staged TMA scatter256", c);
      run_stage<512, 16, 2>("sk 512x16 staged TMA scatter256", c);
      run_stage<256, 16, 2>("sk 256x16 staged TMA scatter256", c);
      run_stage<1024, 8, 2>("sk 1024x8 staged TMA scatter256", c);
    }
  if (want("v7a")) run_timed("v7 512x8 tma probe4 rmw", c, 8, 5, 4096, true, L_v7<512, 8, 8, 2, true, 4>);
  if (want("v7b")) run_timed("v7 384x12 tma probe4 rmw", c, 8, 5, 4608, true, L_v7<384, 12, 8, 2, true, 4>);
  if (want("v7c")) run_timed("v7 384x8 tma probe4 rmw", c, 8, 5, 3072, true, L_v7<384, 8, 8, 3, true, 4>);
  if (want("v7d")) run_timed("v7 256x16 tma probe4 rmw", c, 8, 5, 4096, true, L_v7<256, 16, 8, 2, true, 4>);
  if (want("v7e")) run_timed("v7 256x12 tma probe4 rmw", c, 8, 5, 3072, true, L_v7<256, 12, 8, 3, true, 4>);
  if (want("v7f")) run_timed("v4 384x12 tma probe4 atomics", c, 8, 5, 4608, true, L_v4<384, 12, 8, 2, true, 4>);
  if (want("v8a")) run_timed("v8 384x12 ballot (ref)", c, 8, 5, 4608, true, L_v8<384, 12, 8, 2, 4, 0>);
  if (want("v8b")) run_timed("v8 384x12 ballot2 tuned", c, 8, 5, 4608, true, L_v8<384, 12, 8, 2, 4, 1>);
  if (want("v8c")) run_timed("v8 384x12 table matcher", c, 8, 5, 4608, true, L_v8<384, 12, 8, 2, 4, 2>);
  if (want("v8d")) run_timed("v8 384x8 ballot2 tuned", c, 8, 5, 3072, true, L_v8<384, 8, 8, 3, 4, 1>);
  if (want("v8e")) run_timed("v8 384x8 table matcher", c, 8, 5, 3072, true, L_v8<384, 8, 8, 3, 4, 2>);
  if (want("v8f")) run_timed("v8 512x8 table matcher", c, 8, 5, 4096, true, L_v8<512, 8, 8, 2, 4, 2>);
  if (want("v8g")) run_timed("v8 256x16 table matcher", c, 8, 5, 4096, true, L_v8<256, 16, 8, 2, 4, 2>);
  if (want("v8h")) run_timed("v8 256x12 table matcher", c, 8, 5, 3072, true, L_v8<256, 12, 8, 3, 4, 2>);
  if (want("v9a")) run_timed("v9 384x12 probe8", c, 8, 5, 4608, true, L_v9<384, 12, 2, 8>);
  if (want("v9b")) run_timed("v9 384x12 probe4", c, 8, 5, 4608, true, L_v9<384, 12, 2, 4>);
  if (want("v9c")) run_timed("v9 384x8 probe8", c, 8, 5, 3072, true, L_v9<384, 8, 3, 8>);
  if (want("v9d")) run_timed("v9 512x8 probe8", c, 8, 5, 4096, true, L_v9<512, 8, 2, 8>);
  if (want("v9e")) run_timed("v9 256x12 probe8", c, 8, 5, 3072, true, L_v9<256, 12, 3, 8>);
  if (want("v9f")) run_timed("v9 256x16 probe8", c, 8, 5, 4096, true, L_v9<256, 16, 2, 8>);
  if (want("v9g")) run_timed("v9 384x12 probe16", c, 8, 5, 4608, true, L_v9<384, 12, 2, 16>);
  if (want("v10a")) run_timed("v10 384x12 ballots, ffs leader (=v9)", c, 8, 5, 4608, true, L_v10<384, 12, 2, 8, 0, 0>);
  if (want("v10b")) run_timed("v10 384x12 ballots, highest-lane leader", c, 8, 5, 4608, true, L_v10<384, 12, 2, 8, 0, 1>);
  if (want("v10c")) run_timed("v10 384x12 every 4th MATCH.ANY", c, 8, 5, 4608, true, L_v10<384, 12, 2, 8, 4, 1>);
  if (want("v10d")) run_timed("v10 384x12 every 3rd MATCH.ANY", c, 8, 5, 4608, true, L_v10<384, 12, 2, 8, 3, 1>);
  if (want("v10e")) run_timed("v10 384x12 every 6th MATCH.ANY", c, 8, 5, 4608, true, L_v10<384, 12, 2, 8, 6, 1>);
  if (want("v10f")) run_timed("v10 384x12 every 2nd MATCH.ANY", c, 8, 5, 4608, true, L_v10<384, 12, 2, 8, 2, 1>);
  return 0;
}
