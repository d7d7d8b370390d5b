// Merge-join of the two code-sorted k-mer lists and the seed sort
// (count_thread/merge_thread map.c:881-1002, limit :2992-3055, seed sort :2917-2936,3110-3126).
//
// A = reads index (alen records), B = reference index (blen records), both followed by the
// sentinels ~0, 0.  A run = maximal stretch of equal codes.  The join is partitioned over A:
// every A run head looks its code up in B through a prefix table (LUT over the top P bits of
// the code) + a short binary search, and run pairs present in both lists are compacted, IN A
// ORDER, into a run list with a chained (decoupled look-back) scan.  After the host has
// derived `limit` from the run-product histogram, seeds are emitted at scanned offsets in
// exactly the reference's emission order (code, a, b), so that the stable sort on the
// reference's key bytes (apos, bread, aread) reproduces its array bit for bit.
#include "common.cuh"
#include "index.cuh"
#include "seeds.cuh"

namespace damgpu {

constexpr int MAXGRAM = 10000;                        // map.c:32

struct __align__(16) Run { int32_t ia, na, jb, nb; };

__device__ __forceinline__ uint64_t ld_relaxed64(const uint64_t *p)
{ uint64_t v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed64(uint64_t *p, uint64_t v)
{ asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }

// ---- prefix table over B ---------------------------------------------------------------
// lut[p] = first index i with (B[i].code >> shift) >= p, p in [0, 2^P]; lut[2^P] = blen
// Four independent loads in flight per thread; the predecessor's prefix comes from the neighbouring
// lane (lane 0 reads it).
constexpr int LUT_UNROLL = 4;
__global__ void __launch_bounds__(256)
k_build_lut(const KmerPos *__restrict__ B, int blen, int shift, uint32_t np, uint32_t *__restrict__ lut)
{ const int64_t base = ((int64_t) blockIdx.x * LUT_UNROLL) * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  int64_t cur[LUT_UNROLL], prev[LUT_UNROLL];
#pragma unroll
  for (int u = 0; u < LUT_UNROLL; u++)
    { const int64_t i = base + (int64_t) u * blockDim.x;
      cur[u] = (i < blen) ? (int64_t) (__ldg(&B[i].code) >> shift) : (int64_t) np;
    }
#pragma unroll
  for (int u = 0; u < LUT_UNROLL; u++)
    { const int64_t i = base + (int64_t) u * blockDim.x;
      prev[u] = __shfl_up_sync(0xffffffffu, cur[u], 1);
      if (lane == 0)
        prev[u] = (i == 0) ? -1 : ((i <= blen) ? (int64_t) (__ldg(&B[i - 1].code) >> shift) : (int64_t) np);
    }
#pragma unroll
  for (int u = 0; u < LUT_UNROLL; u++)
    { const int64_t i = base + (int64_t) u * blockDim.x;
      if (i <= blen)
        for (int64_t p = prev[u] + 1; p <= cur[u]; p++)
          lut[p] = (uint32_t) i;
    }
}

// end of the run of code c that starts at `s` (galloping, runs can be long on repeats)
__device__ __forceinline__ int run_end(const KmerPos *__restrict__ L, int s, int len, uint64_t c)
{ int step = 1, lo = s, hi;
  while (true)
    { hi = (lo + step < len) ? lo + step : len;       // invariant: L[lo] == c
      if (hi >= len || L[hi].code != c) break;
      lo = hi; step <<= 1;
    }
  // first index in (lo, hi] that is not c
  while (hi - lo > 1)
    { int mid = (int) (((int64_t) lo + hi) >> 1);
      if (L[mid].code == c) lo = mid; else hi = mid;
    }
  return hi;
}

// ---- match driver run heads against the target list, ordered compaction of matching run pairs --
// The kernel's "A" is the driver (the SHORTER of the two lists: one lookup per distinct driver
// code), its "B" the target searched through the prefix table; with swap != 0 the driver is the
// reference list and the emitted Run has the roles put back.  Both lists are code-sorted, so the
// run list comes out in code order either way.
constexpr int JM_THREADS = 256;
constexpr int JM_ITEMS   = 4;
constexpr int JM_TILE    = JM_THREADS * JM_ITEMS;
constexpr uint64_t J_AGG = 1ull << 62, J_INC = 2ull << 62, J_VAL = (1ull << 62) - 1;

__global__ void __launch_bounds__(JM_THREADS)
k_join_match(const KmerPos *__restrict__ A, int alen, const KmerPos *__restrict__ B, int blen,
             const uint32_t *__restrict__ lut, int shift, int swap, Run *__restrict__ runs,
             uint64_t *tile_state, uint32_t *tile_counter, uint32_t *nruns_out)
{ __shared__ uint32_t s_tile, s_wsum[JM_THREADS / 32];
  __shared__ uint64_t s_excl;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0)
    s_tile = atomicAdd(tile_counter, 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  const int64_t  t0 = (int64_t) tile * JM_TILE;
  const bool last_tile = (t0 + JM_TILE >= alen);

  // blocked arrangement: thread owns elements t0 + tid*JM_ITEMS + j, so run order = thread order
  Run mine[JM_ITEMS];
  int nm = 0;
  const int64_t i0 = t0 + (int64_t) tid * JM_ITEMS;
  uint64_t prev = (i0 > 0 && i0 <= alen) ? A[i0 - 1].code : 0;
#pragma unroll
  for (int j = 0; j < JM_ITEMS; j++)
    { const int64_t i = i0 + j;
      if (i >= alen) break;
      const uint64_t c = A[i].code;
      const bool head = (i == 0) || (c != prev);
      prev = c;
      if (!head) continue;
      // B lookup
      const uint32_t p = (uint32_t) (c >> shift);
      int lo = (int) lut[p], hi = (int) lut[p + 1];
      while (lo < hi)                                  // lower_bound
        { int mid = (lo + hi) >> 1;
          if (B[mid].code < c) lo = mid + 1; else hi = mid;
        }
      if (lo >= blen || B[lo].code != c) continue;
      const int e  = run_end(B, lo, blen, c);
      const int ae = run_end(A, (int) i, alen, c);
      if (swap)                                        // driver = reference list, target = reads list
        { mine[nm].ia = lo; mine[nm].na = e - lo; mine[nm].jb = (int) i; mine[nm].nb = ae - (int) i; }
      else
        { mine[nm].ia = (int) i; mine[nm].na = ae - (int) i; mine[nm].jb = lo; mine[nm].nb = e - lo; }
      nm++;
    }

  // block exclusive scan of nm
  uint32_t x = nm;
  for (int o = 1; o < 32; o <<= 1)
    { uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
  if (lane == 31) s_wsum[warp] = x;
  __syncthreads();
  uint32_t add = 0, total = 0;
  for (int w = 0; w < JM_THREADS / 32; w++)
    { if (w < warp) add += s_wsum[w];
      total += s_wsum[w];
    }
  const uint32_t local = add + x - nm;

  if (tid == 0)
    { uint64_t *st = tile_state + tile;
      uint64_t excl = 0;
      st_relaxed64(st, (tile == 0 ? J_INC : J_AGG) | total);
      if (tile > 0)
        { const uint64_t *p = st - 1;
          while (true)
            { uint64_t v = ld_relaxed64(p);
              if (v & J_INC) { excl += v & J_VAL; break; }
              if (v & J_AGG) { excl += v & J_VAL; p -= 1; continue; }
              __nanosleep(20);
            }
          st_relaxed64(st, J_INC | (excl + total));
        }
      s_excl = excl;
      if (last_tile)
        *nruns_out = (uint32_t) (excl + total);
    }
  __syncthreads();
  const uint64_t o = s_excl + local;
  for (int j = 0; j < nm; j++)
    *reinterpret_cast<int4 *>(runs + o + j) = *reinterpret_cast<int4 *>(&mine[j]);
}

// ---- histogram of run products (count_thread, map.c:919-927) ------------------------------
__global__ void __launch_bounds__(256)
k_run_histogram(const Run *__restrict__ runs, const uint32_t *__restrict__ nruns_p,
                unsigned long long *gram, unsigned long long *total)
{ __shared__ uint32_t sh[2048];
  const uint32_t nruns = *nruns_p;
  for (int i = threadIdx.x; i < 2048; i += 256) sh[i] = 0;
  __syncthreads();
  unsigned long long sum = 0;
  for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < nruns; i += gridDim.x * 256)
    { Run r = runs[i];
      int64_t ct = (int64_t) r.na * r.nb;
      sum += ct;
      if (ct < 2048) atomicAdd(&sh[ct], 1u);
      else if (ct < MAXGRAM) atomicAdd(&gram[ct], 1ull);
    }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0 && sum) atomicAdd(total, sum);
  __syncthreads();
  for (int i = threadIdx.x; i < 2048; i += 256)
    if (sh[i]) atomicAdd(&gram[i], (unsigned long long) sh[i]);
}

// ---- seed offsets: exclusive scan of (product < limit ? product : 0) over the run list ------
__device__ __forceinline__ uint64_t run_hits(const Run &r, int64_t limit)
{ int64_t ct = (int64_t) r.na * r.nb;
  return (ct < limit) ? (uint64_t) ct : 0ull;
}

__global__ void __launch_bounds__(256)
k_hits_blocksum(const Run *__restrict__ runs, uint32_t nruns, int64_t limit, uint64_t *bsum)
{ __shared__ uint64_t ws[8];
  uint32_t i = blockIdx.x * 1024 + threadIdx.x * 4;
  uint64_t s = 0;
  for (int j = 0; j < 4; j++)
    if (i + j < nruns) s += run_hits(runs[i + j], limit);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0)
    { uint64_t t = 0;
      for (int w = 0; w < 8; w++) t += ws[w];
      bsum[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024) k_scan_u64(uint64_t *bsum, int nb, uint64_t *total)
{ __shared__ uint64_t part[1024];
  const int t = threadIdx.x;
  const int per = (nb + 1023) / 1024;
  const int lo = t * per, hi = (lo + per < nb) ? lo + per : nb;
  uint64_t s = 0;
  for (int i = lo; i < hi; i++) s += bsum[i];
  part[t] = s;
  __syncthreads();
  if (t == 0)
    { uint64_t run = 0;
      for (int i = 0; i < 1024; i++)
        { uint64_t c = part[i]; part[i] = run; run += c; }
      *total = run;
    }
  __syncthreads();
  uint64_t run = part[t];
  for (int i = lo; i < hi; i++)
    { uint64_t c = bsum[i]; bsum[i] = run; run += c; }
}

__global__ void __launch_bounds__(256)
k_hits_offsets(const Run *__restrict__ runs, uint32_t nruns, int64_t limit,
               const uint64_t *__restrict__ bsum, uint64_t *__restrict__ off)
{ __shared__ uint64_t ws[8];
  uint32_t i = blockIdx.x * 1024 + threadIdx.x * 4;
  uint64_t v[4], s = 0;
  for (int j = 0; j < 4; j++)
    { v[j] = (i + j < nruns) ? run_hits(runs[i + j], limit) : 0; s += v[j]; }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint64_t x = s;
  for (int o = 1; o < 32; o <<= 1)
    { uint64_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
  if (lane == 31) ws[w] = x;
  __syncthreads();
  uint64_t add = bsum[blockIdx.x];
  for (int k = 0; k < w; k++) add += ws[k];
  uint64_t o = add + x - s;
  for (int j = 0; j < 4; j++)
    if (i + j < nruns)
      { off[i + j] = o; o += v[j]; }
}

// ---- emission (merge_thread, map.c:984-994) + histograms for the seed sort -------------------
struct SeedBytes { int npass; int byte[16]; };

__global__ void __launch_bounds__(256)
k_join_emit(const KmerPos *__restrict__ A, const KmerPos *__restrict__ B,
            const Run *__restrict__ runs, uint32_t nruns, const uint64_t *__restrict__ off,
            uint64_t nhits, SeedPair *__restrict__ hits, SeedBytes sb, uint32_t *hist)
{ extern __shared__ uint32_t sh[];
  for (int i = threadIdx.x; i < sb.npass * 256; i += 256) sh[i] = 0;
  __syncthreads();
  for (uint64_t s = (uint64_t) blockIdx.x * 256 + threadIdx.x; s < nhits;
       s += (uint64_t) gridDim.x * 256)
    { uint32_t lo = 0, hi = nruns;                     // largest u with off[u] <= s
      while (hi - lo > 1)
        { uint32_t mid = (lo + hi) >> 1;
          if (off[mid] <= s) lo = mid; else hi = mid;
        }
      const Run r = runs[lo];
      const uint32_t local = (uint32_t) (s - off[lo]);
      const uint32_t a = local / (uint32_t) r.nb, b = local - a * (uint32_t) r.nb;
      const KmerPos ka = A[r.ia + a], kb = B[r.jb + b];
      SeedPair h;
      h.diag = ka.rpos - kb.rpos; h.apos = ka.rpos; h.bread = kb.read; h.aread = ka.read;
      uint4 v = *reinterpret_cast<uint4 *>(&h);
      *reinterpret_cast<uint4 *>(hits + s) = v;
      const uint32_t w[4] = { v.x, v.y, v.z, v.w };
      for (int p = 0; p < sb.npass; p++)
        atomicAdd(&sh[p * 256 + ((w[sb.byte[p] >> 2] >> ((sb.byte[p] & 3) * 8)) & 0xff)], 1u);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < sb.npass * 256; i += 256)
    if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

__global__ void k_seed_sentinel(SeedPair *hits, uint64_t n)   // map.c:3123-3126
{ hits[n].aread = 0x7fffffff; hits[n].bread = 0x7fffffff; hits[n].diag = 0x7fffffff;
  hits[n].apos = 0;
}

// `limit` from the histogram, map.c:2992-3015 (host arithmetic, identical expressions)
static int compute_limit(const unsigned long long *histo, uint64_t mem_limit, int64_t asize,
                         int64_t bsize, int alen, int blen)
{ if (mem_limit == 0)
    return 0x7fffffff;
  int64_t avail = (int64_t) ((uint64_t) (int64_t) (mem_limit - (uint64_t) (asize + bsize)) / 16u);
  if (avail > alen + 2 * (int64_t) blen)
    avail = (avail - alen) / 2;
  else
    avail = avail - (alen + (int64_t) blen);
  avail = (int64_t) (avail * .98);
  int64_t tom = 0;
  int j;
  for (j = 0; j < MAXGRAM; j++)
    { tom += j * (int64_t) histo[j];
      if (tom > avail)
        break;
    }
  return j;
}

float g_join_times[4] = { 0, 0, 0, 0 };                 // lut ms, match ms, alen, blen of the last call
static cudaEvent_t g_join_ev[3] = { nullptr, nullptr, nullptr };

void join_times(float out[4])
{ if (g_join_ev[2] != nullptr && cudaEventSynchronize(g_join_ev[2]) == cudaSuccess)
    { cudaEventElapsedTime(&g_join_times[0], g_join_ev[0], g_join_ev[1]);
      cudaEventElapsedTime(&g_join_times[1], g_join_ev[1], g_join_ev[2]);
    }
  for (int i = 0; i < 4; i++) out[i] = g_join_times[i];
}

SeedSet *merge_join(const KmerIndex *aidx, const DeviceBlock *ablock, const KmerIndex *bidx,
                    const DeviceBlock *bblock, int K, uint64_t mem_limit, cudaStream_t stream)
{ SeedSet *ss = new SeedSet();
  if (aidx->len == 0 || bidx->len == 0)
    return ss;
  if (bidx->deferred)
    materialize_index(const_cast<KmerIndex *>(bidx), stream);
  const int full_alen = aidx->len;                     // `alen` of the hit cap: the whole reads list
  aidx = reads_view(aidx, bidx, stream);               // deferred reads index: its filtered view
  const int alen = aidx->len, blen = bidx->len;
  if (alen == 0)                                       // no reads k-mer occurs in the reference block
    { ss->histo.assign(MAXGRAM + 1, 0);
      ss->limit = compute_limit(ss->histo.data(), mem_limit, ablock->sizeof_db, bblock->sizeof_db,
                                full_alen, blen);
      if (mem_limit > 0 && ss->limit <= 1)               // map.c:3017-3028
        fatal("Insufficient memory allocation (%.1fGb), reduce block size or increase allocation",
              (1. * mem_limit) / 0x40000000ll);
      ss->hits = dalloc<SeedPair>(1);
      LAUNCH(k_seed_sentinel, 1, 1, 0, stream, ss->hits, (uint64_t) 0);
      return ss;
    }
  const KmerPos *A = aidx->list, *B = bidx->list;

  TRACE(nullptr);
  // driver = shorter list; target = longer list, searched through a prefix table over the top P
  // bits of the code (about 4 target records per bucket), kept with the index it was built over.
  const bool swap = (blen < alen);
  const KmerPos *D = swap ? B : A, *T = swap ? A : B;
  const int dlen = swap ? blen : alen, tlen = swap ? alen : blen;
  int P = 1;
  while ((1ll << P) * 4 < tlen && P < 26) P++;
  if (P > 2 * K) P = 2 * K;
  const int shift = 2 * K - P;
  const uint32_t np = 1u << P;
  uint32_t *lut;
  cudaEvent_t *ev = g_join_ev;                           // lut | match (bench: damgpu_last_join_times)
  if (g_time_kernels)
    { if (ev[0] == nullptr)
        for (int i = 0; i < 3; i++) cudaEventCreate(&ev[i]);
      cudaEventRecord(ev[0], stream);
    }
  const KmerIndex *tidx = swap ? aidx : bidx;          // the table stays with the list it indexes: a reads
  if (tidx->lut != nullptr)                             // index serves both strands and every reference block,
    lut = tidx->lut;                                    // a resident reference index every reads block
  else
    { lut = dalloc<uint32_t>((size_t) np + 2);
      LAUNCH(k_build_lut, (tlen + 256 * LUT_UNROLL) / (256 * LUT_UNROLL), 256, 0, stream, T, tlen, shift, np, lut);
      tidx->lut = lut;
    }

  if (g_time_kernels) cudaEventRecord(ev[1], stream);
  const uint32_t ntiles = (uint32_t) (((int64_t) dlen + JM_TILE - 1) / JM_TILE);
  uint64_t *state = dalloc<uint64_t>((size_t) ntiles + 2);
  CUDA_CHECK(cudaMemsetAsync(state, 0, sizeof(uint64_t) * ((size_t) ntiles + 2), stream));
  uint32_t *counter = reinterpret_cast<uint32_t *>(state + ntiles);      // [0]=tile counter, [1]=nruns
  // worst case every driver record heads a matching run (normally a few percent do)
  Run *runs = dalloc<Run>((size_t) dlen + 1);
  LAUNCH(k_join_match, ntiles, JM_THREADS, 0, stream, D, dlen, T, tlen, lut, shift, swap ? 1 : 0,
         runs, state, counter, counter + 1);
  if (g_time_kernels)
    { cudaEventRecord(ev[2], stream);                    // read by join_times(), no sync here
      g_join_times[2] = (float) alen; g_join_times[3] = (float) blen;
    }
  TRACE("join: lut+match launch");
  // histogram of the run products; the run count stays on the device until both are read back
  unsigned long long *gram = dalloc<unsigned long long>(MAXGRAM + 1);
  CUDA_CHECK(cudaMemsetAsync(gram, 0, sizeof(unsigned long long) * (MAXGRAM + 1), stream));
  { int grid = (int) (((int64_t) dlen + 255) / 256);
    if (grid > sm_count() * 8) grid = sm_count() * 8;
    LAUNCH(k_run_histogram, grid, 256, 0, stream, runs, counter + 1, gram, gram + MAXGRAM);
  }
  uint32_t nruns = 0;
  ss->histo.resize(MAXGRAM + 1);
  CUDA_CHECK(cudaMemcpyAsync(&nruns, counter + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
  CUDA_CHECK(cudaMemcpyAsync(ss->histo.data(), gram, sizeof(unsigned long long) * (MAXGRAM + 1),
                             cudaMemcpyDeviceToHost, stream));
  CUDA_CHECK(cudaStreamSynchronize(stream));
  dfree(state);
  dfree(gram);

  TRACE("join: match sync+histogram");
  const int limit = compute_limit(ss->histo.data(), mem_limit, ablock->sizeof_db,
                                  bblock->sizeof_db, full_alen, blen);
  ss->limit = limit;
  if (mem_limit > 0 && limit <= 1)                     // map.c:3017-3028
    fatal("Insufficient memory allocation (%.1fGb), reduce block size or increase allocation",
          (1. * mem_limit) / 0x40000000ll);

  // seed offsets
  uint64_t nhits = 0;
  uint64_t *off = dalloc<uint64_t>((size_t) nruns + 2);
  if (nruns > 0)
    { int nb = (int) ((nruns + 1023) / 1024);
      uint64_t *bsum = dalloc<uint64_t>(nb + 1);
      LAUNCH(k_hits_blocksum, nb, 256, 0, stream, runs, nruns, (int64_t) limit, bsum);
      LAUNCH(k_scan_u64, 1, 1024, 0, stream, bsum, nb, bsum + nb);
      LAUNCH(k_hits_offsets, nb, 256, 0, stream, runs, nruns, (int64_t) limit, bsum, off);
      CUDA_CHECK(cudaMemcpyAsync(&nhits, bsum + nb, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
      CUDA_CHECK(cudaStreamSynchronize(stream));
      dfree(bsum);
    }
  ss->nhits = (int64_t) nhits;

  TRACE("join: offsets");
  // pairsort key bytes, map.c:2917-2936
  int bytes[16], npass = 0;
  { int64_t powr; int nbyte;
    powr = 1; for (nbyte = 0; powr < ablock->maxlen; nbyte++) powr <<= 8;
    for (int i = 4; i < 4 + nbyte; i++) bytes[npass++] = i;
    powr = 1; for (nbyte = 0; powr < bblock->nreads; nbyte++) powr <<= 8;
    for (int i = 8; i < 8 + nbyte; i++) bytes[npass++] = i;
    powr = 1; for (nbyte = 0; powr < ablock->nreads; nbyte++) powr <<= 8;
    for (int i = 12; i < 12 + nbyte; i++) bytes[npass++] = i;
  }

  SeedPair *h1 = dalloc<SeedPair>((size_t) nhits + 1);
  SeedPair *h2 = dalloc<SeedPair>((size_t) nhits + 1);
  uint32_t *hist = dalloc<uint32_t>(256 * 16);
  CUDA_CHECK(cudaMemsetAsync(hist, 0, sizeof(uint32_t) * 256 * 16, stream));
  if (nhits > 0)
    { SeedBytes sb; sb.npass = npass;
      for (int i = 0; i < npass; i++) sb.byte[i] = bytes[i];
      uint64_t g64 = (nhits + 255) / 256;
      int grid = (g64 > (uint64_t) sm_count() * 16) ? sm_count() * 16 : (int) g64;
      LAUNCH(k_join_emit, grid, 256, sizeof(uint32_t) * 256 * npass, stream, A, B, runs, nruns, off,
             nhits, h1, sb, hist);
    }
  if (nhits >= (1ull << 30))
    fatal("Match_Filter: %llu seed hits exceed the 2^30 sort limit; lower -M or use -t",
          (unsigned long long) nhits);
  TRACE("join: emit");
  SeedPair *rez = (SeedPair *) radix_sort16(h1, h2, (uint32_t) nhits, bytes, npass, hist, stream);
  LAUNCH(k_seed_sentinel, 1, 1, 0, stream, rez, nhits);
  dfree(rez == h1 ? h2 : h1);
  dfree(hist); dfree(off); dfree(runs);
  TRACE("join: seed sort");
  ss->hits = rez;
  return ss;
}

void free_seeds(SeedSet *ss)
{ if (ss == nullptr) return;
  dfree(ss->hits);
  delete ss;
}

}  // namespace damgpu
