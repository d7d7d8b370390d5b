// Stable LSD radix sort of 16-byte records, 8-bit digits, single pass per digit
// ("onesweep": per-tile digit counts are chained through a decoupled look-back so every
// pass reads the array once and writes it once = 32 B per record per pass).
//
// Replaces lex_sort/lex_thread of the reference (map.c:181-444).  The reference sorts the
// key bytes flagged in bytes[16], least significant first, with a stable scatter; any
// stable LSD sort over the same bytes yields the same array (SURVEY.md section 4 items 1,2).
#include "common.cuh"

namespace damgpu {

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS   = 8;
constexpr int RS_TILE    = RS_THREADS * RS_ITEMS;      // records per tile
constexpr int RS_WARPS   = RS_THREADS / 32;

// look-back word: [31:30] flag, [29:0] value
constexpr uint32_t FLAG_AGG = 1u << 30;
constexpr uint32_t FLAG_INC = 2u << 30;
constexpr uint32_t VAL_MASK = (1u << 30) - 1;

__device__ __forceinline__ uint32_t rec_byte(const uint4 &r, int byte)
{ uint32_t w = (byte < 8) ? ((byte < 4) ? r.x : r.y) : ((byte < 12) ? r.z : r.w);
  return (w >> ((byte & 3) * 8)) & 0xffu;
}

__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t *p)
{ uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void st_relaxed(uint32_t *p, uint32_t v)
{ asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }

// ---- histogram of every pass byte in one read ---------------------------------------

struct PassBytes { int npass; int byte[16]; };

__global__ void __launch_bounds__(256)
k_radix_histogram(const uint4 *__restrict__ recs, uint32_t n, PassBytes pb, uint32_t *hist)
{ extern __shared__ uint32_t sh[];                    // [npass][256]
  for (int i = threadIdx.x; i < pb.npass * 256; i += blockDim.x)
    sh[i] = 0;
  __syncthreads();
  for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t) gridDim.x * blockDim.x)
    { uint4 r = recs[i];
      for (int p = 0; p < pb.npass; p++)
        atomicAdd(&sh[p * 256 + rec_byte(r, pb.byte[p])], 1u);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < pb.npass * 256; i += blockDim.x)
    if (sh[i])
      atomicAdd(&hist[i], sh[i]);
}

// hist[p][d] -> exclusive prefix over d (in place), one block per pass
__global__ void __launch_bounds__(256) k_radix_prefix(uint32_t *hist)
{ __shared__ uint32_t ws[8];
  uint32_t *h = hist + blockIdx.x * 256;
  uint32_t v = h[threadIdx.x], x = v;
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int o = 1; o < 32; o <<= 1)
    { uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
  if (lane == 31) ws[w] = x;
  __syncthreads();
  uint32_t add = 0;
  for (int i = 0; i < w; i++) add += ws[i];
  h[threadIdx.x] = x + add - v;
}

// ---- one scatter pass -----------------------------------------------------------------
//
// Tile = 2048 records, warp-striped: warp w owns records [w*256, w*256+256) of the tile, item
// i of lane l is record i*32+l of that chunk.  Ranking is a warp-level multisplit
// (__match_any_sync) on per-warp shared-memory counters, which keeps the order of equal
// digits (stability).  Thread d (0..255) owns digit d for the cross-warp scan, the look-back
// and the publication of the tile's counts.  Records are then staged in shared memory in
// sorted order so the global writes are coalesced runs per digit.

__global__ void __launch_bounds__(RS_THREADS)
k_radix_pass(const uint4 *__restrict__ in, uint4 *__restrict__ out, uint32_t n, int byte,
             const uint32_t *__restrict__ gbase, uint32_t *tile_state, uint32_t *tile_counter)
{ __shared__ uint4    stage[RS_TILE];                 // 32 KB; first 8 KB alias the counters
  __shared__ uint32_t s_dbase[256];                   // first local rank of digit d in the tile
  __shared__ uint32_t s_delta[256];                   // global index - local rank for digit d
  __shared__ uint32_t s_wsum[RS_WARPS];
  __shared__ uint32_t s_tile;
  uint32_t (*whist)[256] = reinterpret_cast<uint32_t (*)[256]>(stage);   // [RS_WARPS][256]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0)
    s_tile = atomicAdd(tile_counter, 1u);
  for (int i = lane; i < 256; i += 32)
    whist[warp][i] = 0;
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint32_t base = tile * (uint32_t) RS_TILE + warp * (32 * RS_ITEMS) + lane;
  const uint32_t nvalid = (n - tile * (uint32_t) RS_TILE < (uint32_t) RS_TILE)
                              ? n - tile * (uint32_t) RS_TILE : (uint32_t) RS_TILE;

  uint4    rec[RS_ITEMS];
  uint32_t dig[RS_ITEMS], rank[RS_ITEMS];
#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++)
    { uint32_t idx = base + i * 32;
      if (idx < n)
        { rec[i] = __ldcs(in + idx);
          dig[i] = rec_byte(rec[i], byte);
        }
      else
        { rec[i] = make_uint4(0, 0, 0, 0);
          dig[i] = 255;                               // padding ranks after every real 255
        }
    }

  const uint32_t lt = (1u << lane) - 1;
#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++)
    { uint32_t peers = __match_any_sync(0xffffffffu, dig[i]);
      uint32_t prev  = whist[warp][dig[i]];
      __syncwarp();
      if ((peers & lt) == 0)
        whist[warp][dig[i]] = prev + __popc(peers);
      __syncwarp();
      rank[i] = prev + __popc(peers & lt);
    }
  __syncthreads();

  // digit d: exclusive scan over warps, tile count, look-back
  { const int d = tid;
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; w++)
      { uint32_t c = whist[w][d];
        whist[w][d] = run;
        run += c;
      }
    if (d == 255)
      run -= (uint32_t) RS_TILE - nvalid;             // padding is not counted
    uint32_t *st = tile_state + (size_t) tile * 256 + d;
    st_relaxed(st, (tile == 0 ? FLAG_INC : FLAG_AGG) | run);

    // exclusive scan of the tile counts over digits -> s_dbase
    uint32_t x = run;
    for (int o = 1; o < 32; o <<= 1)
      { uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
    if (lane == 31) s_wsum[warp] = x;
    __syncthreads();
    uint32_t add = 0;
    for (int w = 0; w < warp; w++) add += s_wsum[w];
    const uint32_t dbase = x + add - run;

    uint32_t excl = 0;
    if (tile > 0)
      { const uint32_t *p = st - 256;
        while (true)
          { uint32_t v = ld_relaxed(p);
            if (v & FLAG_INC) { excl += v & VAL_MASK; break; }
            if (v & FLAG_AGG) { excl += v & VAL_MASK; p -= 256; continue; }
            __nanosleep(20);
          }
        st_relaxed(st, FLAG_INC | (excl + run));
      }
    s_dbase[d] = dbase;
    s_delta[d] = gbase[d] + excl - dbase;
  }
  __syncthreads();

  uint32_t pos[RS_ITEMS];
#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++)
    pos[i] = s_dbase[dig[i]] + whist[warp][dig[i]] + rank[i];
  __syncthreads();                                    // counters die, staging area is live
#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++)
    stage[pos[i]] = rec[i];
  __syncthreads();

#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++)
    { uint32_t j = tid + i * RS_THREADS;
      if (j < nvalid)
        { uint4 r = stage[j];
          __stcs(out + (j + s_delta[rec_byte(r, byte)]), r);
        }
    }
}

void radix_histogram(const void *recs, uint32_t n, const int *bytes, int npass, uint32_t *hist,
                     cudaStream_t stream)
{ PassBytes pb;
  pb.npass = npass;
  for (int i = 0; i < npass; i++) pb.byte[i] = bytes[i];
  CUDA_CHECK(cudaMemsetAsync(hist, 0, sizeof(uint32_t) * 256 * npass, stream));
  if (n == 0) return;
  int grid = sm_count() * 8;
  LAUNCH(k_radix_histogram, grid, 256, sizeof(uint32_t) * 256 * npass, stream,
         (const uint4 *) recs, n, pb, hist);
}

void *radix_sort16(void *a, void *b, uint32_t n, const int *bytes, int npass, uint32_t *hist,
                   cudaStream_t stream)
{ if (n == 0 || npass == 0)
    return a;
  if (n >= (1u << 30))
    fatal("radix_sort16: %u records exceed the 2^30 limit of the 32-bit look-back words", n);
  const uint32_t ntiles = (n + RS_TILE - 1) / RS_TILE;
  uint32_t *state = dalloc<uint32_t>((size_t) ntiles * 256 + 1);
  uint32_t *counter = state + (size_t) ntiles * 256;

  LAUNCH(k_radix_prefix, npass, 256, 0, stream, hist);
  uint4 *src = (uint4 *) a, *dst = (uint4 *) b;
  for (int p = 0; p < npass; p++)
    { CUDA_CHECK(cudaMemsetAsync(state, 0, sizeof(uint32_t) * ((size_t) ntiles * 256 + 1), stream));
      LAUNCH(k_radix_pass, ntiles, RS_THREADS, 0, stream, src, dst, n, bytes[p],
             hist + p * 256, state, counter);
      uint4 *t = src; src = dst; dst = t;
    }
  CUDA_CHECK(cudaStreamSynchronize(stream));
  dfree(state);
  return src;
}

}  // namespace damgpu
