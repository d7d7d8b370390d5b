// Deferred reads index: Sort_Kmers of a reads block without sorting what can never match.
//
// Match_Filter only ever uses the reads index (map.c:655-822) through the merge-join against one
// reference block's index (map.c:881-1002): a reads record whose code is absent from the reference
// list contributes nothing -- no run pair, no histogram entry, no seed.  With noisy long reads that
// is ~95 % of the list (an error-free 20-mer has probability 0.85^20 = 4 %).  So the reads side is
// kept DEFERRED: the call records the block, and the first Match_Filter against a reference block
//   1. sets two bits (in one 64-bit word) per reference code AND per reverse-complemented code in a bitmap
//      (k_ref_bitmap) -- both orientations at once, so the complement call of the same block
//      (damapper.c:847-861) finds the filtered list ready;
//   2. re-runs the extraction with a bitmap test per k-mer and an ordered compaction
//      (k_extract_filtered: ballot ranks inside a tile, chained scan across tiles), which keeps the
//      (read, rpos) extraction order the stable sort relies on;
//   3. sorts the survivors with the same LSD radix passes.
// The result is a sub-list of the reference's sorted list that contains every record whose code
// occurs in the reference block (plus a few percent false positives of the hash), in the same
// relative order, hence the same run pairs, the same `gram` histogram, the same seeds.  `alen` of the
// hit-cap formula (map.c:3002-3007) stays the full count (limit_len).
//
// A filtered list is keyed by a signature of the reference list that is invariant under
// complementing the block (sum of a mix of the canonical codes, and the length).  Anything that
// needs the whole list (download/export, -t, masks, a third reference block) materialises it with
// sort_kmers().
#include "common.cuh"
#include "index.cuh"

namespace damgpu {

int g_filter_mode = 1;                                 // 0 = off, 1 = auto, 2 = always
int g_filter_log2 = 0;                                 // DAMGPU_FILTER_BITS: log2 of the bitmap size (0 = auto)
static const int64_t FILTER_MIN_KMERS = 4000000;       // auto: lists shorter than this are simply sorted
static const int     FILTER_MAX_BUILDS = 2;            // auto: distinct reference blocks before the full sort

constexpr int FX_THREADS = 256;
constexpr int FX_ITEMS   = 16;
constexpr int FX_TILE    = FX_THREADS * FX_ITEMS;
constexpr int FX_MAXSPAN = 512;
constexpr int FX_WARPS   = FX_THREADS / 32;
constexpr int FX_BATCH   = 4;                          // bitmap lookups in flight per thread
constexpr uint64_t FX_AGG = 1ull << 62, FX_INC = 2ull << 62, FX_VAL = (1ull << 62) - 1;

__device__ __forceinline__ uint64_t mix64(uint64_t x)
{ x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}

// Blocked Bloom filter with two bits per code: the top bits of the hash pick a 64-bit word of the
// bitmap, two 6-bit fields of the same hash pick the bits inside it, so inserting is one 64-bit
// atomicOr and testing one 8-byte load (the same 32-byte sector a single-bit test would fetch).  With
// 32 bits per reference k-mer and both orientations inserted ~7 % of the bits are set and ~0.6 % of the
// absent codes pass, against 3.4 % with one bit per code.
// (the hash is kept as its top 32 bits: word index from the top, bit fields from the bottom)
__device__ __forceinline__ uint32_t hash_of(uint64_t code) { return (uint32_t) ((code * 0x9E3779B97F4A7C15ull) >> 32); }
__device__ __forceinline__ uint32_t word_of(uint32_t h, int wshift) { return h >> wshift; }
__device__ __forceinline__ uint64_t bits_of(uint32_t h)      // second field from a remix: the low bits of h
{ return (1ull << (h & 63)) | (1ull << ((h * 0x9E3779B1u) >> 26)); }   // above bit 5 also feed the word index

// code of the reverse complement of a K-mer (2 bits per base, first base most significant)
__device__ __forceinline__ uint64_t rc_code(uint64_t c, int K)
{ uint64_t x = __brevll(c);                            // reverses the bases and the bits inside each base
  x = ((x & 0x5555555555555555ull) << 1) | ((x >> 1) & 0x5555555555555555ull);
  x >>= (64 - 2 * K);
  const uint64_t kmask = (K == 32) ? ~0ull : ((1ull << (2 * K)) - 1);
  return (~x) & kmask;
}

// sig[0], sig[1] += two mixes of the canonical code over the list (invariant under complementing the
// block); with bitmap != null also sets the bits of every code and of its reverse complement
__global__ void __launch_bounds__(256)
k_ref_bitmap(const KmerPos *__restrict__ B, int blen, int K, int hshift, unsigned long long *bitmap,
             unsigned long long *sig)
{ unsigned long long s = 0, s2 = 0;
  for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < blen;
       i += (int64_t) gridDim.x * blockDim.x)
    { const uint64_t c = __ldg(&B[i].code), r = rc_code(c, K);
      s += mix64(c < r ? c : r);
      s2 += mix64((c < r ? c : r) ^ 0x9e3779b97f4a7c15ull);
      if (bitmap != nullptr)
        { const uint32_t hc = hash_of(c), hr = hash_of(r);
          atomicOr(&bitmap[word_of(hc, hshift)], (unsigned long long) bits_of(hc));
          atomicOr(&bitmap[word_of(hr, hshift)], (unsigned long long) bits_of(hr));
        }
    }
  for (int o = 16; o > 0; o >>= 1)
    { s += __shfl_down_sync(0xffffffffu, s, o); s2 += __shfl_down_sync(0xffffffffu, s2, o); }
  if ((threadIdx.x & 31) == 0)
    { if (s) atomicAdd(sig, s);
      if (s2) atomicAdd(sig + 1, s2);
    }
}

__device__ __forceinline__ uint64_t fx_ld(const uint64_t *p)
{ uint64_t v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fx_st(uint64_t *p, uint64_t v)
{ asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }

// per tile of FX_TILE base positions: r0 = read holding the tile's first position (largest r with
// boff[r] <= t0) and l2 = last read starting before the tile's end; built once per block
__global__ void __launch_bounds__(256)
k_tile_reads(const int64_t *__restrict__ boff, int nreads, int64_t ntiles, int32_t *__restrict__ tab)
{ const int64_t tile = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (tile >= ntiles) return;
  const int64_t t0 = tile * FX_TILE;
  int lo = 0, hi = nreads - 1;
  while (lo < hi)
    { int mid = (lo + hi + 1) >> 1;
      if (boff[mid] <= t0) lo = mid; else hi = mid - 1;
    }
  int l2 = lo, h2 = nreads - 1;                         // the last READ starting before the tile's end
  while (l2 < h2)
    { int mid = (l2 + h2 + 1) >> 1;
      if (boff[mid] < t0 + FX_TILE) l2 = mid; else h2 = mid - 1;
    }
  tab[2 * tile] = lo;
  tab[2 * tile + 1] = l2;
}

void ensure_tile_tab(const DeviceBlock *blk, cudaStream_t stream)
{ if (blk->tile_tab != nullptr || blk->nreads == 0)
    return;
  const int64_t ntiles = (blk->total + FX_TILE - 1) / FX_TILE;
  blk->tile_tab = dalloc<int32_t>((size_t) 2 * ntiles);
  LAUNCH(k_tile_reads, (unsigned) ((ntiles + 255) / 256), 256, 0, stream, blk->boff, blk->nreads,
         ntiles, blk->tile_tab);
}

// code of the k-mer whose last base sits at offset `off` of the tile (window index off+32)
__device__ __forceinline__ uint64_t window_code(const uint64_t *s_pack, int off, uint64_t kmask)
{ const int e = off + 32 + 1;                         // one past the last base, in bases
  const int wj = (e - 1) >> 5;                        // word holding the last base
  const int sh = 2 * (32 - (e - (wj << 5)));          // free low bits in that word
  const uint64_t hiw = (wj > 0) ? s_pack[wj - 1] : 0ull, low = s_pack[wj];
  const uint64_t c = (sh == 0) ? low : ((low >> sh) | (hiw << (64 - sh)));
  return c & kmask;
}

// tuple_thread (map.c:466-579) with a membership test: position q of the block image is the last
// base of a k-mer of read r iff rpos >= K-1 and q is not the terminator; the k-mer survives iff its
// hash bit is set.  Survivors are written in extraction order: tiles are handed out by a ticket,
// ranks inside a tile come from ballots (item-major, then warp, then lane = ascending q), the tile's
// base from a chained scan over the tile totals (look-back by a whole warp, 32 predecessors per
// round trip).  counters: [0] ticket, [1] survivors, [2] overflow.
#ifndef FX_MINB
#define FX_MINB 4
#endif
__global__ void __launch_bounds__(FX_THREADS, FX_MINB)
k_extract_filtered(const uint8_t *__restrict__ bases, const int64_t *__restrict__ boff, int nreads,
                   int64_t total, int K, const int32_t *__restrict__ tile_tab,
                   const unsigned long long *__restrict__ bitmap, int hshift, KmerPos *__restrict__ out,
                   uint32_t cap, uint64_t *tile_state, uint32_t *counters)
{ __shared__ uint64_t s_pack[FX_TILE / 32 + 2];
  __shared__ int64_t  s_boff[FX_MAXSPAN + 2];
  __shared__ uint32_t s_cnt[FX_ITEMS * FX_WARPS];
  __shared__ uint32_t s_wtot[4];
  __shared__ uint32_t s_tile;
  __shared__ uint64_t s_excl;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t kmask = (K == 32) ? ~0ull : ((1ull << (2 * K)) - 1);
  const uint32_t lt = (1u << lane) - 1;

  const int64_t ntiles = (total + FX_TILE - 1) / FX_TILE;
  while (true)
    { __syncthreads();
      if (tid == 0)
        s_tile = atomicAdd(&counters[0], 1u);
      __syncthreads();
      const int64_t tile = s_tile;
      if (tile >= ntiles)
        break;
      const int64_t t0 = tile * FX_TILE;
      const int r0 = __ldg(&tile_tab[2 * tile]);
      const int nspan = __ldg(&tile_tab[2 * tile + 1]) - r0 + 1;    // entries r0 .. l2, plus boff[l2+1]
      const bool cached = (nspan <= FX_MAXSPAN);

      for (int c = tid; c < (FX_TILE + 32) / 16; c += FX_THREADS)
        { int64_t q = t0 - 32 + (int64_t) c * 16;
          uint32_t w = 0;
          // bases is 16-byte aligned and q a multiple of 16.  The alignment test is spelled out because a
          // round-1 build without it faulted on B200 ("misaligned address" at this load).  Round 2 fixed an
          // out-of-bounds read of boff in k_tile_reads / the s_boff staging (l2 could reach nreads); since
          // then the build without the test (-DFX_NO_ALIGN_GUARD) runs the full-size C2 steps clean, so
          // the test is belt and braces (compute-sanitizer is closed on this pool: no memcheck run)
#ifdef FX_NO_ALIGN_GUARD                                    /* dev: reproduce the fault under compute-sanitizer */
          if (q >= 0 && q + 16 <= total)
#else
          if (q >= 0 && q + 16 <= total && (((uintptr_t) (bases + q)) & 15) == 0)
#endif
            { uint4 v = __ldcs(reinterpret_cast<const uint4 *>(bases + q));
              uint32_t x[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
              for (int j = 0; j < 4; j++)
                { uint32_t u = x[j];
                  w = (w << 8) | ((u & 3) << 6) | (((u >> 8) & 3) << 4) |
                      (((u >> 16) & 3) << 2) | ((u >> 24) & 3);
                }
            }
          else
            { for (int j = 0; j < 16; j++)
                { int64_t p = q + j;
                  uint32_t b = (p >= 0 && p < total) ? bases[p] : 0;
                  w = (w << 2) | (b & 3);
                }
            }
          reinterpret_cast<uint32_t *>(s_pack)[c ^ 1] = w;
        }
      if (cached)
        for (int i = tid; i <= nspan; i += FX_THREADS)
          s_boff[i] = boff[r0 + i];
      __syncthreads();

      // membership of the thread's 16 positions, four lookups in flight at a time
      uint32_t keep = 0;
#pragma unroll
      for (int half = 0; half < FX_ITEMS / FX_BATCH; half++)
        { uint32_t hsh[FX_BATCH];
          uint64_t word[FX_BATCH];
          uint32_t valid = 0;
#pragma unroll
          for (int j = 0; j < FX_BATCH; j++)
            { const int off = tid + (half * FX_BATCH + j) * FX_THREADS;
              const int64_t q = t0 + off;
              hsh[j] = 0;
              if (q >= total)
                continue;
              int64_t b0, b1;
              if (cached)
                { int lo = 0, hi = nspan - 1;
                  while (lo < hi)
                    { int mid = (lo + hi + 1) >> 1;
                      if (s_boff[mid] <= q) lo = mid; else hi = mid - 1;
                    }
                  b0 = s_boff[lo]; b1 = s_boff[lo + 1];
                }
              else
                { int lo = r0, hi = nreads - 1;
                  while (lo < hi)
                    { int mid = (lo + hi + 1) >> 1;
                      if (boff[mid] <= q) lo = mid; else hi = mid - 1;
                    }
                  b0 = boff[lo]; b1 = boff[lo + 1];
                }
              if (q - b0 < K - 1 || q >= b1 - 1)
                continue;
              hsh[j] = hash_of(window_code(s_pack, off, kmask));
              valid |= 1u << j;
            }
#pragma unroll
          for (int j = 0; j < FX_BATCH; j++)
            word[j] = ((valid >> j) & 1) ? __ldg(&bitmap[word_of(hsh[j], hshift)]) : 0ull;
#pragma unroll
          for (int j = 0; j < FX_BATCH; j++)
            { const int it = half * FX_BATCH + j;
              const uint64_t need = bits_of(hsh[j]);
              const uint32_t k = ((word[j] & need) == need) ? 1u : 0u;    // word is 0 where no k-mer ends
              keep |= k << it;
              const uint32_t b = __ballot_sync(0xffffffffu, k);
              if (lane == 0)
                s_cnt[it * FX_WARPS + warp] = __popc(b);
            }
        }
      __syncthreads();
      // exclusive scan of the 128 (item, warp) counts
      if (tid < FX_ITEMS * FX_WARPS)
        { const uint32_t v = s_cnt[tid];
          uint32_t x = v;
          for (int o = 1; o < 32; o <<= 1)
            { uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
              if (lane >= o) x += y;
            }
          if (lane == 31) s_wtot[warp] = x;
          s_cnt[tid] = x - v;
        }
      __syncthreads();
      uint32_t tot = 0;
      { uint32_t add = 0;
        for (int w = 0; w < 4; w++)
          { if (w < warp) add += s_wtot[w];
            tot += s_wtot[w];
          }
        if (tid < FX_ITEMS * FX_WARPS)
          s_cnt[tid] += add;
      }
      if (warp == 0)                                    // chained scan over the tiles
        { uint64_t *st = tile_state + tile;
          if (lane == 0)
            fx_st(st, (tile == 0 ? FX_INC : FX_AGG) | tot);
          uint64_t excl = 0;
          if (tile > 0)
            { int64_t j = tile - 1;                     // nearest predecessor not yet summed
              while (true)
                { const int64_t idx = j - lane;
                  const uint64_t v = (idx >= 0) ? fx_ld(tile_state + idx) : FX_INC;
                  const uint32_t notready = __ballot_sync(0xffffffffu, (v >> 62) == 0);
                  const uint32_t inc = __ballot_sync(0xffffffffu, (v & FX_INC) != 0);
                  const int first = inc ? (__ffs(inc) - 1) : 32;
                  const uint32_t upto = (first >= 31) ? 0xffffffffu : ((2u << first) - 1);
                  if (notready & upto)
                    { __nanosleep(40);
                      continue;
                    }
                  const uint32_t mine = ((upto >> lane) & 1u) ? (uint32_t) (v & FX_VAL) : 0u;
                  excl += __reduce_add_sync(0xffffffffu, mine);
                  if (first < 32)
                    break;
                  j -= 32;
                }
              if (lane == 0)
                fx_st(st, FX_INC | (excl + tot));
            }
          if (lane == 0)
            { s_excl = excl;
              if (tile == ntiles - 1)
                counters[1] = (uint32_t) (excl + tot);
              if (excl + tot > cap)
                counters[2] = 1;
            }
        }
      __syncthreads();
      const uint64_t base = s_excl;
      if (tot == 0)
        continue;
#pragma unroll
      for (int it = 0; it < FX_ITEMS; it++)
        { const uint32_t k = (keep >> it) & 1u;
          const uint32_t b = __ballot_sync(0xffffffffu, k);
          if (!k)
            continue;
          const uint64_t pos = base + s_cnt[it * FX_WARPS + warp] + __popc(b & lt);
          if (pos >= cap)
            continue;
          const int off = tid + it * FX_THREADS;
          const int64_t q = t0 + off;
          int r; int64_t b0;
          if (cached)
            { int lo = 0, hi = nspan - 1;
              while (lo < hi)
                { int mid = (lo + hi + 1) >> 1;
                  if (s_boff[mid] <= q) lo = mid; else hi = mid - 1;
                }
              r = r0 + lo; b0 = s_boff[lo];
            }
          else
            { int lo = r0, hi = nreads - 1;
              while (lo < hi)
                { int mid = (lo + hi + 1) >> 1;
                  if (boff[mid] <= q) lo = mid; else hi = mid - 1;
                }
              r = lo; b0 = boff[lo];
            }
          KmerPos kp;
          kp.code = window_code(s_pack, off, kmask); kp.rpos = (int) (q - b0); kp.read = r;
          __stcs(reinterpret_cast<uint4 *>(out + pos), *reinterpret_cast<uint4 *>(&kp));
        }
    }
}

__global__ void k_filter_sentinels(KmerPos *list, int64_t n)   // map.c:772-773
{ list[n].code = 0xffffffffffffffffull; list[n].rpos = 0; list[n].read = 0;
  list[n + 1].code = 0;                 list[n + 1].rpos = 0; list[n + 1].read = 0;
}

// Sort_Kmers with the list left unbuilt (reads side).  -t and masks need the whole list: built now.
KmerIndex *sort_kmers_deferred(const DeviceBlock *blk, int K, int suppress, cudaStream_t stream)
{ const int64_t kmers64 = blk->total - (int64_t) K * blk->nreads;
  if (g_filter_mode == 0 || suppress > 0 || blk->mask_off != nullptr || kmers64 <= 0 ||
      (g_filter_mode == 1 && kmers64 < FILTER_MIN_KMERS))
    return sort_kmers(blk, K, suppress, stream);
  if (kmers64 > 0x7fffffffll)                          // `int kmers`, map.c:663,676
    fatal("Sort_Kmers: block holds %lld k-mers, more than 2^31-1", (long long) kmers64);
  if (kmers64 >= (1ll << 30))                          // look-back words of the radix pass are 32 bits wide
    fatal("Sort_Kmers: block holds %lld k-mers; this build sorts at most 2^30-1 per block: split the "
          "database into smaller blocks (DBsplit -s)", (long long) kmers64);
  for (int i = 0; i < blk->nreads; i++)
    if (blk->h_rlen[i] < K)                            // damapper.c:403-410
      fatal("Sort_Kmers: block contains reads < %dbp long", K);
  KmerIndex *idx = new KmerIndex();
  idx->deferred = true;
  idx->src = blk;
  idx->K = K;
  idx->len = (int) kmers64;
  return idx;
}

// build the whole list of a deferred index (download, export, third reference block)
void materialize_index(KmerIndex *idx, cudaStream_t stream)
{ if (!idx->deferred)
    return;
  KmerIndex *full = sort_kmers(idx->src, idx->K, 0, stream);
  idx->list = full->list; full->list = nullptr;
  idx->len = full->len;
  idx->ms_extract = full->ms_extract; idx->ms_sort = full->ms_sort; idx->npass = full->npass;
  idx->deferred = false;
  free_index(full);
  if (idx->filt != nullptr)
    { free_index(idx->filt);
      idx->filt = nullptr;
    }
}

static KmerIndex *build_filtered(const KmerIndex *a, const KmerIndex *b, unsigned long long *bitmap,
                                 int hshift, cudaStream_t stream)
{ const DeviceBlock *blk = a->src;
  const int K = a->K;
  const uint32_t n = (uint32_t) a->len;
  int bytes[16], npass = 0;
  for (int i = 0; i < 2 * K; i += 8)                   // mersort, map.c:670-673
    bytes[npass++] = i >> 3;

  const int64_t ntiles = (blk->total + FX_TILE - 1) / FX_TILE;
  ensure_tile_tab(blk, stream);                          // reads of every tile, once per block
  uint32_t *hist = dalloc<uint32_t>(256 * 16);
  uint64_t *state = dalloc<uint64_t>((size_t) ntiles + 2);
  uint32_t *counters = reinterpret_cast<uint32_t *>(state + ntiles);      // 3 words used
  int grid = sm_count() * FX_MINB;                     // resident CTAs, tiles by ticket
  if (grid > ntiles) grid = (int) ntiles;

  cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
  if (g_time_kernels)
    { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
      cudaEventRecord(e0, stream);
    }
  // survivors are a few percent of the list; if a repeat-rich block overflows the guess, run again
  // with room for everything
  uint32_t cap = n / 6 + 4096;
  if (cap > n) cap = n;
  KmerPos *lst = nullptr;
  uint32_t res[3] = { 0, 0, 0 };
  while (true)
    { lst = dalloc<KmerPos>((size_t) cap + 2);
      CUDA_CHECK(cudaMemsetAsync(state, 0, sizeof(uint64_t) * ((size_t) ntiles + 2), stream));
      LAUNCH(k_extract_filtered, grid, FX_THREADS, 0, stream, blk->bases, blk->boff, blk->nreads,
             blk->total, K, blk->tile_tab, bitmap, hshift, lst, cap, state, counters);
      CUDA_CHECK(cudaMemcpyAsync(res, counters, sizeof(res), cudaMemcpyDeviceToHost, stream));
      CUDA_CHECK(cudaStreamSynchronize(stream));
      if (res[2] == 0)
        break;
      dfree(lst);
      cap = n;
    }
  const uint32_t kept = res[1];
  if (g_time_kernels) cudaEventRecord(e1, stream);

  KmerIndex *f = new KmerIndex();
  f->limit_len = a->len;
  f->npass = npass;
  if (kept > 0)
    { KmerPos *tmp = dalloc<KmerPos>((size_t) kept + 2);
      radix_histogram(lst, kept, bytes, npass, hist, stream);     // one read of the survivors
      KmerPos *rez = (KmerPos *) radix_sort16(lst, tmp, kept, bytes, npass, hist, stream);
      LAUNCH(k_filter_sentinels, 1, 1, 0, stream, rez, (int64_t) kept);
      dfree(rez == lst ? tmp : lst);
      f->list = rez;
      f->len = (int) kept;
    }
  else
    { // no reads k-mer occurs in the reference block: an empty list with its sentinels
      LAUNCH(k_filter_sentinels, 1, 1, 0, stream, lst, (int64_t) 0);
      f->list = lst;
      f->len = 0;
    }
  if (g_time_kernels)
    { cudaEventRecord(e2, stream);
      cudaEventSynchronize(e2);
      cudaEventElapsedTime(&f->ms_extract, e0, e1);
      cudaEventElapsedTime(&f->ms_sort, e1, e2);
      cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
    }
  dfree(hist); dfree(state);
  (void) b;
  return f;
}

float g_filter_times[4] = { 0, 0, 0, 0 };              // bitmap ms, filtered extraction ms, sort ms, survivors

// The list Match_Filter joins against reference index b: a itself, or the filtered view of a
const KmerIndex *reads_view(const KmerIndex *ca, const KmerIndex *b, cudaStream_t stream)
{ if (!ca->deferred)
    return ca;
  KmerIndex *a = const_cast<KmerIndex *>(ca);
  if (b == nullptr || b->len == 0)
    { materialize_index(a, stream);
      return a;
    }
  const int K = a->K;
  int grid = (b->len + 255) / 256;
  if (grid > sm_count() * 8) grid = sm_count() * 8;
  if (a->filt != nullptr)
    { bool same = (b->src_uid != 0 && a->filt_uid == b->src_uid);      // the same block, either strand
      if (!same && a->filt_blen == b->len)                             // another object: compare the hashes
        { unsigned long long *sig = dalloc<unsigned long long>(2);
          unsigned long long h[2] = { 0, 0 };
          CUDA_CHECK(cudaMemsetAsync(sig, 0, sizeof(unsigned long long) * 2, stream));
          LAUNCH(k_ref_bitmap, grid, 256, 0, stream, b->list, b->len, K, 0, (unsigned long long *) nullptr, sig);
          CUDA_CHECK(cudaMemcpyAsync(h, sig, sizeof(h), cudaMemcpyDeviceToHost, stream));
          if (!a->filt_sig_known)
            CUDA_CHECK(cudaMemcpyAsync(a->filt_sig, a->filt_dsig, sizeof(a->filt_sig), cudaMemcpyDeviceToHost, stream));
          CUDA_CHECK(cudaStreamSynchronize(stream));
          a->filt_sig_known = true;
          dfree(sig);
          same = (h[0] == a->filt_sig[0] && h[1] == a->filt_sig[1]);
          if (same && b->src_uid != 0)
            a->filt_uid = b->src_uid;                                  // next time without the pass
        }
      if (same)
        { g_filter_times[0] = g_filter_times[1] = g_filter_times[2] = 0.f;
          return a->filt;
        }
    }
  // bitmap: ~32 bits per reference k-mer (four bits set per k-mer: two per orientation, 12 % of the
  // bits at most).  In automatic mode it has to stay resident in L2 (2^29 bits = 64 MB): a lookup that
  // goes to DRAM costs about what sorting the record costs, so a reference list that would fill more
  // than 35 % of such a bitmap (12 % false positives), or that is longer than 1.5x the reads list, gets
  // the plain full sort.
  int lg = g_filter_log2;
  if (lg == 0)
    { lg = 20;
      while (lg < 32 && (1ll << lg) < 32ll * b->len) lg++;
      if (g_filter_mode == 1 && lg > 29) lg = 29;
    }
  if (lg < 10) lg = 10;
  if (lg > 32) lg = 32;
  if (g_filter_mode == 1 &&
      (a->nfilt >= FILTER_MAX_BUILDS || 4.0 * b->len > 0.35 * (double) (1ll << lg) ||
       (double) b->len > 1.5 * (double) a->len))
    { materialize_index(a, stream);
      return a;
    }
  if (a->filt != nullptr)
    { free_index(a->filt);
      a->filt = nullptr;
    }
  const size_t words = (size_t) 1 << (lg - 6);          // 64-bit words
  const int    wshift = 32 - (lg - 6);                 // the hash is 32 bits wide
  unsigned long long *bitmap = dalloc<unsigned long long>(words);
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (g_time_kernels)
    { cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0, stream);
    }
  CUDA_CHECK(cudaMemsetAsync(bitmap, 0, words * sizeof(unsigned long long), stream));
  if (a->filt_dsig == nullptr)
    a->filt_dsig = dalloc<unsigned long long>(2);
  CUDA_CHECK(cudaMemsetAsync(a->filt_dsig, 0, sizeof(unsigned long long) * 2, stream));
  LAUNCH(k_ref_bitmap, grid, 256, 0, stream, b->list, b->len, K, wshift, bitmap, a->filt_dsig);
  if (g_time_kernels) cudaEventRecord(e1, stream);
  KmerIndex *f = build_filtered(a, b, bitmap, wshift, stream);
  if (g_time_kernels)
    { cudaEventSynchronize(e1);
      cudaEventElapsedTime(&g_filter_times[0], e0, e1);
      cudaEventDestroy(e0); cudaEventDestroy(e1);
      g_filter_times[1] = f->ms_extract; g_filter_times[2] = f->ms_sort;
    }
  g_filter_times[3] = (float) f->len;
  dfree(bitmap);
  a->filt = f;
  a->filt_uid = b->src_uid;
  a->filt_sig_known = false;
  a->filt_blen = b->len;
  a->nfilt += 1;
  return f;
}

}  // namespace damgpu
