// Stable LSD radix sort of 16-byte records, 8-bit digits, single pass per digit
// ("onesweep": per-tile digit counts are chained through a decoupled look-back so every
// pass reads the array once and writes it once = 32 B per record per pass; the sorted tile
// leaves shared memory through TMA bulk stores, one per digit run).
//
// Replaces lex_sort/lex_thread of the reference (map.c:181-444).  The reference sorts the
// key bytes flagged in bytes[16], least significant first, with a stable scatter; any
// stable LSD sort over the same bytes yields the same array (SURVEY.md section 4 items 1,2).
#include <vector>
#include "common.cuh"

namespace damgpu {

#ifndef RS_THREADS_V                                   // build-time variants (tools/sort_bench.py)
#define RS_THREADS_V 384
#define RS_ITEMS_V   12
#define RS_MINB_V    2
#endif
constexpr int RS_THREADS = RS_THREADS_V;
constexpr int RS_ITEMS   = RS_ITEMS_V;
constexpr int RS_TILE    = RS_THREADS * RS_ITEMS;      // records per tile (4608 = 72 KB staged)
constexpr int RS_WARPS   = RS_THREADS / 32;
constexpr int RS_PROBE   = 8;                          // predecessors fetched per look-back round
constexpr int RS_SMEM    = RS_TILE * 16;               // dynamic shared memory per CTA

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

// Lanes of the warp holding the same 8-bit digit: eight ballots (MATCH.ANY costs ~2.6x as much on
// sm_100 when most of the 32 digits are distinct; tools/sortlab.cu).  The pass is bound by the ALU
// pipe, so the loop is spelled out: per bit one AND+SETP (bit -> predicate), the VOTE, one SELP
// and one LOP3 accumulating the lanes that differ from this one.
__device__ __forceinline__ uint32_t match_digit(uint32_t dig)
{ uint32_t diff = 0;
#pragma unroll
  for (int b = 0; b < 8; b++)
    { uint32_t v, sx;
      asm("{ .reg .pred p; .reg .b32 t;\n\t"
          "and.b32 t, %2, %3;\n\t"
          "setp.ne.u32 p, t, 0;\n\t"
          "vote.sync.ballot.b32 %0, p, 0xffffffff;\n\t"
          "selp.b32 %1, 0xffffffff, 0, p;\n\t}"
          : "=r"(v), "=r"(sx) : "r"(dig), "r"(1u << b));
      diff |= v ^ sx;                                   // lanes whose bit b differs from mine
    }
  return ~diff;
}

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
// Tile = 4608 records (384 threads x 12), warp-striped: warp w owns records [w*384, w*384+384)
// of the tile, item i of lane l is record i*32+l of that chunk, so tile order = (warp, item,
// lane).  Tile ids come from an atomic ticket, so a tile's predecessors are always held by CTAs
// that started earlier and the look-back cannot deadlock.
//   1. early counts: per-warp digit histograms with plain shared atomics; thread d sums digit d
//      over the warps and publishes the tile aggregate at once, long before it is needed by
//      successors;
//   2. the per-warp counters are rewritten as staged start positions (exclusive over digits,
//      then over warps);
//   3. stable ranking: the lanes holding the same digit are found with ballots, the lowest of
//      them bumps the warp's counter with one shared atomic and broadcasts the old value (the
//      twelve items are issued in order and shared atomics of a warp execute in program order,
//      which keeps equal digits in input order);
//   4. records are staged in shared memory in sorted order (the counters die here);
//   5. decoupled look-back for digit d, RS_PROBE predecessors per round trip;
//   6. thread d hands the run of digit d to the TMA engine (cp.async.bulk shared -> global:
//      16-byte records make every run a legal bulk copy), no per-record store instructions.

// W32 = 32-bit word of the record holding the key byte (compile time), psel = PRMT selector that
// moves that byte to bits 0..7: the digit costs one instruction wherever it is needed.
// RELOAD: only the key word of every record is loaded before the ranking (the 64-byte sectors come
// from DRAM either way and stay in L2); the whole records are read a second time, from L2, when they
// are staged.  No record lives in registers across the ranking, so three CTAs fit an SM instead of two.
template <int W32, bool RELOAD>
__global__ void __launch_bounds__(RS_THREADS, RELOAD ? 3 : RS_MINB_V)
k_radix_pass(const uint4 *__restrict__ in, uint4 *__restrict__ out, uint32_t n, uint32_t psel,
             const uint32_t *__restrict__ gbase, uint32_t *tile_state, uint32_t *tile_counter,
             uint32_t pf_dist)
{
#define REC_DIGIT(r) __byte_perm((W32 == 0) ? (r).x : (W32 == 1) ? (r).y : (W32 == 2) ? (r).z : (r).w, 0, psel)
  extern __shared__ uint4 stage[];                    // 72 KB; the first 12 KB alias the counters
  __shared__ uint32_t s_wsum[RS_WARPS];
  __shared__ uint32_t s_tile;
  uint32_t *whist = reinterpret_cast<uint32_t *>(stage);                 // [RS_WARPS][256]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lt = (1u << lane) - 1;

  if (tid == 0)
    s_tile = atomicAdd(tile_counter, 1u);
  uint32_t *wh = whist + warp * 256;
#pragma unroll
  for (int i = lane; i < 256; i += 32)
    wh[i] = 0;
  __syncthreads();
  const uint32_t tile   = s_tile;
  // the tile that a CTA will pick up pf_dist tickets from now is pulled into L2 by the copy engine,
  // so DRAM keeps streaming while the resident CTAs are in their ranking phases
  if (tid == 0 && pf_dist != 0 && (uint64_t) (tile + pf_dist + 1) * RS_TILE <= n)
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;"
                 :: "l"(in + (size_t) (tile + pf_dist) * RS_TILE), "r"((uint32_t) RS_SMEM) : "memory");
  const uint32_t tbase  = tile * (uint32_t) RS_TILE;
  const uint32_t nvalid = (n - tbase < (uint32_t) RS_TILE) ? n - tbase : (uint32_t) RS_TILE;
  const uint4 *src = in + tbase + warp * (32 * RS_ITEMS) + lane;

  uint4 rec[RELOAD ? 1 : RS_ITEMS];
  uint32_t dg[RS_ITEMS];
  const uint32_t rbase = tbase + warp * (32 * RS_ITEMS) + lane;         // index of this thread's item 0
  if (RELOAD)
    { const uint32_t *kw = reinterpret_cast<const uint32_t *>(src) + W32;
      if (nvalid == (uint32_t) RS_TILE)
        {
#pragma unroll
          for (int i = 0; i < RS_ITEMS; i++)
            dg[i] = __byte_perm(__ldg(kw + i * 128), 0, psel);
        }
      else
        {
#pragma unroll
          for (int i = 0; i < RS_ITEMS; i++)
            dg[i] = (rbase + i * 32 < n) ? __byte_perm(__ldg(kw + i * 128), 0, psel) : 255u;
        }
    }
  else
    { if (nvalid == (uint32_t) RS_TILE)               // every tile but the last: no bounds tests
        {
#pragma unroll
          for (int i = 0; i < RS_ITEMS; i++)
            rec[RELOAD ? 0 : i] = __ldcs(src + i * 32);
        }
      else                                            // padding = all ones: digit 255, ranks after
        {                                             // every real 255 of the tile
#pragma unroll
          for (int i = 0; i < RS_ITEMS; i++)
            rec[RELOAD ? 0 : i] = (rbase + i * 32 < n) ? __ldcs(src + i * 32)
                                         : make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
        }
#pragma unroll
      for (int i = 0; i < RS_ITEMS; i++)
        dg[i] = REC_DIGIT(rec[RELOAD ? 0 : i]);
    }
#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++)
    atomicAdd(&wh[dg[i]], 1u);
  __syncthreads();

  // digit d = tid (threads 256.. only take part in the barriers)
  uint32_t cnt = 0, dbase;
  uint32_t *st = tile_state + (size_t) tile * 256 + (tid & 255);
  if (tid < 256)
    {
#pragma unroll
      for (int w = 0; w < RS_WARPS; w++)
        cnt += whist[w * 256 + tid];
      if (tid == 255)
        cnt -= (uint32_t) RS_TILE - nvalid;           // padding is not counted
      st_relaxed(st, (tile == 0 ? FLAG_INC : FLAG_AGG) | cnt);
    }
  { uint32_t x = cnt;                                 // exclusive scan over digits
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
      { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
    if (lane == 31) s_wsum[warp] = x;
    __syncthreads();
    uint32_t add = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; w++)
      if (w < warp) add += s_wsum[w];
    dbase = x + add - cnt;
  }
  if (tid < 256)
    { uint32_t run = dbase;
#pragma unroll
      for (int w = 0; w < RS_WARPS; w++)
        { const uint32_t c = whist[w * 256 + tid];
          whist[w * 256 + tid] = run;
          run += c;
        }
    }
  __syncthreads();

  uint32_t pos[RS_ITEMS];
#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++)
    { const uint32_t dig   = dg[i];
      const uint32_t peers = match_digit(dig);
      uint32_t old = 0;
      if ((peers & lt) == 0)
        old = atomicAdd(&wh[dig], (uint32_t) __popc(peers));
      pos[i] = __shfl_sync(0xffffffffu, old, __ffs(peers) - 1) + __popc(peers & lt);
    }
  __syncthreads();                                    // counters die, staging area is live
  if (RELOAD)
    { if (nvalid == (uint32_t) RS_TILE)
        { uint4 r[RS_ITEMS / 2];                      // second read of the tile, from L2; six in flight
#pragma unroll
          for (int h = 0; h < 2; h++)
            {
#pragma unroll
              for (int i = 0; i < RS_ITEMS / 2; i++)
                r[i] = __ldcs(src + (h * (RS_ITEMS / 2) + i) * 32);
#pragma unroll
              for (int i = 0; i < RS_ITEMS / 2; i++)
                stage[pos[h * (RS_ITEMS / 2) + i]] = r[i];
            }
        }
      else
        {
#pragma unroll
          for (int i = 0; i < RS_ITEMS; i++)
            if (rbase + i * 32 < n)                   // padding ranks behind every real record: not staged
              stage[pos[i]] = __ldcs(src + i * 32);
        }
    }
  else
    {
#pragma unroll
      for (int i = 0; i < RS_ITEMS; i++)
        stage[pos[i]] = rec[RELOAD ? 0 : i];
    }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staged data -> async proxy

  uint32_t gdst = 0;
  if (tid < 256)
    { uint32_t excl = 0;
      if (tile > 0)
        { const uint32_t *p = st - 256;                 // predecessor tile, same digit
          uint32_t left = tile;
          while (true)
            { uint32_t v[RS_PROBE];
#pragma unroll
              for (int k = 0; k < RS_PROBE; k++)
                v[k] = ((uint32_t) k < left) ? ld_relaxed(p - (size_t) k * 256) : FLAG_INC;
              uint32_t adv = 0;
              bool stop = false, inc = false;
#pragma unroll
              for (int k = 0; k < RS_PROBE; k++)
                if (!stop)
                  { if (v[k] == 0) stop = true;         // not published yet: ask again from here
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
  __syncthreads();                                    // staging complete

  if (tid < 256 && cnt > 0)
    { const uint32_t src = (uint32_t) __cvta_generic_to_shared(stage + dbase);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                   :: "l"(out + gdst), "r"(src), "r"(cnt * 16u) : "memory");
    }
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory may be released
#undef REC_DIGIT
}

bool g_radix_reload = false;                           // DAMGPU_RADIX=reload|regs
int  g_radix_pf = 0;                                   // DAMGPU_RADIX_PF: L2 prefetch distance in tiles
typedef void (*radix_pass_fn)(const uint4 *, uint4 *, uint32_t, uint32_t, const uint32_t *, uint32_t *, uint32_t *, uint32_t);
static radix_pass_fn radix_pass_for(int byte)
{ switch (byte >> 2)
    { case 0:  return g_radix_reload ? k_radix_pass<0, true> : k_radix_pass<0, false>;
      case 1:  return g_radix_reload ? k_radix_pass<1, true> : k_radix_pass<1, false>;
      case 2:  return g_radix_reload ? k_radix_pass<2, true> : k_radix_pass<2, false>;
      default: return g_radix_reload ? k_radix_pass<3, true> : k_radix_pass<3, false>;
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

// With damgpu_time_kernels on, every sort leaves a pair of events around its passes; radix_totals()
// resolves them: bytes (32 per record and pass), ms, pass launches and calls since the last reset.
struct PendingSort { cudaEvent_t e0, e1; double bytes; int npass; };
static std::vector<PendingSort> g_pending;
static double g_tot[4] = { 0., 0., 0., 0. };

void radix_totals(double out[4], int reset)
{ for (PendingSort &p : g_pending)
    { float ms = 0.f;
      if (cudaEventSynchronize(p.e1) == cudaSuccess && cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess)
        { g_tot[0] += p.bytes; g_tot[1] += ms; g_tot[2] += p.npass; g_tot[3] += 1; }
      cudaEventDestroy(p.e0); cudaEventDestroy(p.e1);
    }
  g_pending.clear();
  for (int i = 0; i < 4; i++) out[i] = g_tot[i];
  if (reset) g_tot[0] = g_tot[1] = g_tot[2] = g_tot[3] = 0.;
}

void *radix_sort16(void *a, void *b, uint32_t n, const int *bytes, int npass, uint32_t *hist,
                   cudaStream_t stream)
{ if (n == 0 || npass == 0)
    return a;
  if (n >= (1u << 30))
    fatal("radix_sort16: %u records exceed the 2^30 limit of the 32-bit look-back words", n);
  const uint32_t ntiles = (n + RS_TILE - 1) / RS_TILE;
  // tile states + ticket counter of EVERY pass, cleared by one memset (a memset per pass is a launch
  // per pass: on the 5-10 M record lists of a step the passes are 25-45 us each)
  const size_t per_pass = (size_t) ntiles * 256 + 1;
  uint32_t *state = dalloc<uint32_t>(per_pass * (size_t) npass);

  static int attr_set = -1;
  if (attr_set != (int) g_radix_reload)
    { for (int w = 0; w < 4; w++)
        { CUDA_CHECK(cudaFuncSetAttribute(radix_pass_for(4 * w), cudaFuncAttributeMaxDynamicSharedMemorySize, RS_SMEM));
          if (g_radix_reload)
            CUDA_CHECK(cudaFuncSetAttribute(radix_pass_for(4 * w), cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        }
      attr_set = (int) g_radix_reload;
    }
  LAUNCH(k_radix_prefix, npass, 256, 0, stream, hist);
  CUDA_CHECK(cudaMemsetAsync(state, 0, sizeof(uint32_t) * per_pass * (size_t) npass, stream));
  PendingSort ps;
  if (g_time_kernels)
    { cudaEventCreate(&ps.e0); cudaEventCreate(&ps.e1);
      ps.bytes = 32. * (double) n * npass; ps.npass = npass;
      cudaEventRecord(ps.e0, stream);
    }
  uint4 *src = (uint4 *) a, *dst = (uint4 *) b;
  for (int p = 0; p < npass; p++)
    { uint32_t *st = state + per_pass * (size_t) p;
      LAUNCH(radix_pass_for(bytes[p]), ntiles, RS_THREADS, RS_SMEM, stream, src, dst, n,
             0x4440u | (uint32_t) (bytes[p] & 3), hist + p * 256, st, st + (size_t) ntiles * 256, (uint32_t) g_radix_pf);
      uint4 *t = src; src = dst; dst = t;
    }
  if (g_time_kernels)
    { cudaEventRecord(ps.e1, stream);
      g_pending.push_back(ps);
    }
  dfree(state);                                        // stream-ordered reuse (cache_alloc)
  return src;
}

}  // namespace damgpu
