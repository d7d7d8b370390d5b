// K-mer tuple extraction and index build (Sort_Kmers, map.c:655-822; tuple_thread :466-579).
//
// Extraction walks the block image of Load_All_Reads (one byte per base, a 4 after every read)
// by absolute base position q: position q of read r (rpos = q - boff[r]) is the last base of
// a k-mer iff rpos >= k-1 and q is not the read's terminator.  Its slot in the list is
// g = q + 1 - (r+1)*k, i.e. extraction order (read, rpos) -- the order the reference's
// per-thread loops produce -- so the later stable sort on the code bytes gives the
// reference's (code, read, rpos) order.
#include "common.cuh"
#include "index.cuh"

namespace damgpu {

constexpr int EX_THREADS = 256;
constexpr int EX_ITEMS   = 16;
constexpr int EX_TILE    = EX_THREADS * EX_ITEMS;     // base positions per tile (= the tile of blk->tile_tab)
static_assert(EX_TILE == 4096, "tile_tab is built for 4096-position tiles");
constexpr int EX_MAXSPAN = 512;                       // read starts cached per tile

// Tile-local 2-bit packing: word j (64 bit) holds bases 32j..32j+31 of the window, first base
// in the most significant bits, so the code of a k-mer is a funnel shift of two words.
__global__ void __launch_bounds__(EX_THREADS)
k_extract(const uint8_t *__restrict__ bases, const int64_t *__restrict__ boff, int nreads,
          int64_t total, int K, int npass, KmerPos *__restrict__ list, uint32_t *hist,
          const int64_t *__restrict__ mask_off, const int32_t *__restrict__ mask_pts,
          const int32_t *__restrict__ tile_tab)
{ // window = [t0-32, t0+EX_TILE): 32 bases of left context (K <= 32)
  __shared__ uint64_t s_pack[EX_TILE / 32 + 2];
  __shared__ int64_t  s_boff[EX_MAXSPAN + 2];
  __shared__ uint32_t s_hist[8 * 256];

  const int tid = threadIdx.x;
  const uint64_t kmask = (K == 32) ? ~0ull : ((1ull << (2 * K)) - 1);

  for (int i = tid; i < npass * 256; i += EX_THREADS)
    s_hist[i] = 0;

  const int64_t ntiles = (total + EX_TILE - 1) / EX_TILE;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
    { const int64_t t0 = tile * EX_TILE;
      __syncthreads();

      // pack 16 bases per thread; threads 0..(EX_TILE+32)/16-1 cover the window
      for (int c = tid; c < (EX_TILE + 32) / 16; c += EX_THREADS)
        { int64_t q = t0 - 32 + (int64_t) c * 16;
          uint32_t w = 0;
          if (q >= 0 && q + 16 <= total)
            { uint4 v = *reinterpret_cast<const uint4 *>(bases + q);   // bases is 16B aligned
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
          reinterpret_cast<uint32_t *>(s_pack)[c ^ 1] = w;   // big-endian pairs inside a u64
        }

      // reads overlapping the tile, from the per-tile table (k_tile_reads): r0 = read containing t0
      // (largest r with boff[r] <= t0), l2 = last read starting before the tile end
      const int r0 = __ldg(&tile_tab[2 * tile]);
      const int nspan = __ldg(&tile_tab[2 * tile + 1]) - r0 + 1;      // entries r0 .. l2, plus boff[l2+1]
      __syncthreads();
      const bool cached = (nspan <= EX_MAXSPAN);
      if (cached)
        for (int i = tid; i <= nspan; i += EX_THREADS)
          s_boff[i] = boff[r0 + i];
      __syncthreads();

#pragma unroll 4
      for (int it = 0; it < EX_ITEMS; it++)
        { const int off = tid + it * EX_THREADS;
          const int64_t q = t0 + off;
          if (q >= total)
            break;
          // read of position q
          int r;
          int64_t b0, b1;
          if (cached)
            { int lo = 0, hi = nspan - 1;
              while (lo < hi)
                { int mid = (lo + hi + 1) >> 1;
                  if (s_boff[mid] <= q) lo = mid; else hi = mid - 1;
                }
              r = r0 + lo; b0 = s_boff[lo]; b1 = s_boff[lo + 1];
            }
          else
            { int lo = r0, hi = nreads - 1;
              while (lo < hi)
                { int mid = (lo + hi + 1) >> 1;
                  if (boff[mid] <= q) lo = mid; else hi = mid - 1;
                }
              r = lo; b0 = boff[lo]; b1 = boff[lo + 1];
            }
          const int rpos = (int) (q - b0);
          if (rpos < K - 1 || q >= b1 - 1)
            continue;
          // code = bases q-K+1 .. q, window index of q is off+32
          const int e = off + 32 + 1;                 // one past the last base, in bases
          const int wj = (e - 1) >> 5;                // word holding the last base
          const int sh = 2 * (32 - (e - (wj << 5)));  // free low bits in that word
          uint64_t hiw = (wj > 0) ? s_pack[wj - 1] : 0ull, low = s_pack[wj];
          uint64_t code = (sh == 0) ? low : ((low >> sh) | (hiw << (64 - sh)));
          code &= kmask;
          const int64_t g = q + 1 - (int64_t) (r + 1) * K;
          KmerPos kp;
          kp.code = code; kp.rpos = rpos; kp.read = r;
          if (mask_off != nullptr)                    // -m, map.c:481-543: the k-mer must lie inside
            { const int64_t mb = mask_off[r], mf = mask_off[r + 1];     // an unmasked segment
              int64_t lo = 0, hi = (mf - mb) >> 1;    // first interval whose begin is > rpos
              while (lo < hi)
                { const int64_t mid = (lo + hi) >> 1;
                  if (mask_pts[mb + 2 * mid] > rpos) hi = mid; else lo = mid + 1;
                }
              const int p = (lo == 0) ? 0 : mask_pts[mb + 2 * lo - 1];  // end of the interval before
              if (rpos - (K - 1) < p)
                { kp.code = ~0ull; kp.rpos = -1; kp.read = -1;          // dropped by the compaction
                  *reinterpret_cast<uint4 *>(list + g) = *reinterpret_cast<uint4 *>(&kp);
                  continue;
                }
            }
          *reinterpret_cast<uint4 *>(list + g) = *reinterpret_cast<uint4 *>(&kp);
          for (int p = 0; p < npass; p++)
            atomicAdd(&s_hist[p * 256 + ((code >> (8 * p)) & 0xff)], 1u);
        }
    }
  __syncthreads();
  for (int i = tid; i < npass * 256; i += EX_THREADS)
    if (s_hist[i])
      atomicAdd(&hist[i], s_hist[i]);
}

__global__ void k_set_sentinels(KmerPos *list, int64_t n)   // map.c:772-773
{ list[n].code = 0xffffffffffffffffull; list[n].rpos = 0; list[n].read = 0;
  list[n + 1].code = 0;                 list[n + 1].rpos = 0; list[n + 1].read = 0;
}

// -t: keep[i] = 1 iff the run of equal codes holding i is shorter than t (map.c:590-636).
// On a sorted list the run holding i has length >= t iff some window of t consecutive
// records covering i has equal end codes.
__global__ void k_suppress_flags(const KmerPos *__restrict__ list, int64_t n, int t,
                                 uint32_t *__restrict__ keep)
{ int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t c = list[i].code;
  // leftmost equal within t-1 to the left, rightmost equal within t-1 to the right
  int64_t lo = i - (t - 1) < 0 ? 0 : i - (t - 1), hi = i;
  while (lo < hi)
    { int64_t mid = (lo + hi) >> 1;
      if (list[mid].code == c) hi = mid; else lo = mid + 1;
    }
  int64_t left = lo;
  lo = i; hi = i + (t - 1) >= n ? n - 1 : i + (t - 1);
  while (lo < hi)
    { int64_t mid = (lo + hi + 1) >> 1;
      if (list[mid].code == c) lo = mid; else hi = mid - 1;
    }
  keep[i] = (lo - left + 1 < t) ? 1u : 0u;
}

// -m: keep[i] = 1 iff slot i holds a real k-mer (masked slots carry read == -1)
__global__ void k_mask_flags(const KmerPos *__restrict__ list, int64_t n, uint32_t *__restrict__ keep)
{ int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keep[i] = (list[i].read != -1) ? 1u : 0u;
}

// block-wise exclusive scan of keep flags (three-kernel scan; -t is off the default path)
__global__ void __launch_bounds__(256) k_scan_blocks(const uint32_t *keep, int64_t n,
                                                     uint32_t *bsum)
{ __shared__ uint32_t ws[8];
  int64_t base = (int64_t) blockIdx.x * 2048;
  uint32_t s = 0;
  for (int j = 0; j < 8; j++)
    { int64_t i = base + threadIdx.x * 8 + j;
      if (i < n) s += keep[i];
    }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0)
    { uint32_t t = 0;
      for (int i = 0; i < 8; i++) t += ws[i];
      bsum[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024) k_scan_bsum(uint32_t *bsum, int nb, uint32_t *total)
{ __shared__ uint32_t part[1024];
  const int t = threadIdx.x;
  const int per = (nb + 1023) / 1024;
  const int lo = t * per, hi = (lo + per < nb) ? lo + per : nb;
  uint32_t s = 0;
  for (int i = lo; i < hi; i++) s += bsum[i];
  part[t] = s;
  __syncthreads();
  if (t == 0)
    { uint32_t run = 0;
      for (int i = 0; i < 1024; i++)
        { uint32_t c = part[i]; part[i] = run; run += c; }
      *total = run;
    }
  __syncthreads();
  uint32_t run = part[t];
  for (int i = lo; i < hi; i++)
    { uint32_t c = bsum[i]; bsum[i] = run; run += c; }
}

__global__ void __launch_bounds__(256) k_compact(const KmerPos *__restrict__ src,
                                                 const uint32_t *__restrict__ keep, int64_t n,
                                                 const uint32_t *__restrict__ bsum,
                                                 KmerPos *__restrict__ dst)
{ __shared__ uint32_t ws[8];
  int64_t base = (int64_t) blockIdx.x * 2048 + threadIdx.x * 8;
  uint32_t f[8], s = 0;
  for (int j = 0; j < 8; j++)
    { f[j] = (base + j < n) ? keep[base + j] : 0; s += f[j]; }
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t x = s;
  for (int o = 1; o < 32; o <<= 1)
    { uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
  if (lane == 31) ws[w] = x;
  __syncthreads();
  uint32_t add = bsum[blockIdx.x];
  for (int i = 0; i < w; i++) add += ws[i];
  uint32_t o = add + x - s;
  for (int j = 0; j < 8; j++)
    if (f[j])
      dst[o++] = src[base + j];
}

DeviceBlock *upload_block(const uint8_t *bases, const int64_t *boff, const int32_t *rlen,
                          int nreads, int tfirst, int maxlen, int64_t totlen, int64_t sizeof_db,
                          cudaStream_t stream)
{ static uint64_t next_uid = 0;
  DeviceBlock *blk = new DeviceBlock();
  blk->uid = ++next_uid;
  blk->nreads = nreads; blk->tfirst = tfirst; blk->maxlen = maxlen; blk->totlen = totlen;
  blk->sizeof_db = sizeof_db;
  blk->total = boff[nreads];
  // image = leading 4 + bases; bases[0] is 16-byte aligned; BLOCK_SLACK bytes on both sides may be
  // read (never interpreted) by the window prefetch of the alignment kernels
  blk->raw = dalloc<uint8_t>((size_t) blk->total + 2 * BLOCK_SLACK);
  blk->bases = blk->raw + BLOCK_SLACK;
  CUDA_CHECK(cudaMemcpyAsync(blk->bases - 1, bases - 1, (size_t) blk->total + 1,
                             cudaMemcpyHostToDevice, stream));
  blk->boff = dalloc<int64_t>(nreads + 1);
  CUDA_CHECK(cudaMemcpyAsync(blk->boff, boff, sizeof(int64_t) * (nreads + 1),
                             cudaMemcpyHostToDevice, stream));
  blk->rlen = dalloc<int32_t>(nreads + 1);
  CUDA_CHECK(cudaMemcpyAsync(blk->rlen, rlen, sizeof(int32_t) * nreads,
                             cudaMemcpyHostToDevice, stream));
  blk->h_boff.assign(boff, boff + nreads + 1);
  blk->h_rlen.assign(rlen, rlen + nreads);
  CUDA_CHECK(cudaStreamSynchronize(stream));
  return blk;
}

// .bps -> Load_All_Reads image (Uncompress_Read DB.c:342-363 + the 4 separators of DB.c:1402-1433).
// Parallel over the PACKED BYTES, not over the reads: a reference block is a handful of contigs of tens
// of Mbp each (the first version, one warp per read, took ~50 ms for a 250 Mbp block of 8 contigs against
// 0.1 ms for the same bases in 25 000 reads).  A CTA takes 4096 consecutive packed bytes; its first thread
// brackets the reads they belong to (two binary searches over poff), every thread then finds the read of
// its byte inside that bracket (usually one read: no step) and writes the four bases.  Bytes between reads
// (a trimmed DB leaves gaps) belong to nobody.  The separators are written by a grid-stride loop over reads.
constexpr int UNP_TILE = 4096;
__global__ void __launch_bounds__(256)
k_unpack_bps(const uint8_t *__restrict__ packed, const int64_t *__restrict__ poff,
             const int64_t *__restrict__ boff, int nreads, int64_t packed_bytes, uint8_t *__restrict__ bases)
{ __shared__ int s_lo, s_hi;
  const int64_t t0 = (int64_t) blockIdx.x * UNP_TILE;
  if (threadIdx.x == 0)
    { const int64_t t1 = ((t0 + UNP_TILE < packed_bytes) ? t0 + UNP_TILE : packed_bytes) - 1;
      int lo = 0, hi = nreads - 1;
      while (lo < hi)                                    // last read with poff <= t0
        { const int mid = (lo + hi + 1) >> 1;
          if (poff[mid] <= t0) lo = mid; else hi = mid - 1;
        }
      s_lo = lo;
      hi = nreads - 1;
      while (lo < hi)                                    // last read with poff <= t1
        { const int mid = (lo + hi + 1) >> 1;
          if (poff[mid] <= t1) lo = mid; else hi = mid - 1;
        }
      s_hi = lo;
    }
  __syncthreads();
  for (int u = 0; u < UNP_TILE / 256; u++)
    { const int64_t t = t0 + (int64_t) u * 256 + threadIdx.x;
      if (t >= packed_bytes) break;
      int lo = s_lo, hi = s_hi;
      while (lo < hi)
        { const int mid = (lo + hi + 1) >> 1;
          if (poff[mid] <= t) lo = mid; else hi = mid - 1;
        }
      const int64_t j = t - poff[lo], b0 = boff[lo];
      const int len = (int) (boff[lo + 1] - b0 - 1);
      if (j < 0 || j >= ((len + 3) >> 2)) continue;      // a byte no read of the block owns
      const uint32_t c = packed[t];
      const int i = (int) (4 * j);
      uint8_t *d = bases + b0;
      d[i] = (uint8_t) (c >> 6);
      if (i + 1 < len) d[i + 1] = (uint8_t) ((c >> 4) & 3);
      if (i + 2 < len) d[i + 2] = (uint8_t) ((c >> 2) & 3);
      if (i + 3 < len) d[i + 3] = (uint8_t) (c & 3);
    }
  for (int64_t r = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; r < nreads; r += (int64_t) gridDim.x * blockDim.x)
    { bases[boff[r + 1] - 1] = 4;
      if (r == 0) bases[-1] = 4;
    }
}

DeviceBlock *upload_block_packed(const uint8_t *packed, const int64_t *poff, int64_t packed_bytes,
                                 const int64_t *boff, const int32_t *rlen, int nreads, int tfirst,
                                 int maxlen, int64_t totlen, int64_t sizeof_db, cudaStream_t stream)
{ static uint64_t next_uid = 1ull << 40;
  DeviceBlock *blk = new DeviceBlock();
  blk->uid = ++next_uid;
  blk->nreads = nreads; blk->tfirst = tfirst; blk->maxlen = maxlen; blk->totlen = totlen;
  blk->sizeof_db = sizeof_db;
  blk->total = boff[nreads];
  TRACE(nullptr);
  blk->raw = dalloc<uint8_t>((size_t) blk->total + 2 * BLOCK_SLACK);
  blk->bases = blk->raw + BLOCK_SLACK;
  blk->boff = dalloc<int64_t>(nreads + 1);
  blk->rlen = dalloc<int32_t>(nreads + 1);
  uint8_t *d_packed = dalloc<uint8_t>((size_t) packed_bytes + 16);
  int64_t *d_poff = dalloc<int64_t>(nreads + 1);
  TRACE("upload_packed: allocations");
  CUDA_CHECK(cudaMemcpyAsync(d_packed, packed, (size_t) packed_bytes, cudaMemcpyHostToDevice, stream));
  CUDA_CHECK(cudaMemcpyAsync(d_poff, poff, sizeof(int64_t) * nreads, cudaMemcpyHostToDevice, stream));
  CUDA_CHECK(cudaMemcpyAsync(blk->boff, boff, sizeof(int64_t) * (nreads + 1), cudaMemcpyHostToDevice, stream));
  CUDA_CHECK(cudaMemcpyAsync(blk->rlen, rlen, sizeof(int32_t) * nreads, cudaMemcpyHostToDevice, stream));
  if (nreads > 0)
    { int64_t grid = (packed_bytes + UNP_TILE - 1) / UNP_TILE;
      if (grid < 1) grid = 1;
      LAUNCH(k_unpack_bps, (int) grid, 256, 0, stream, d_packed, d_poff, blk->boff, nreads, (int64_t) packed_bytes,
             blk->bases);
    }
  else
    CUDA_CHECK(cudaMemsetAsync(blk->bases - 1, 4, 1, stream));
  blk->h_boff.assign(boff, boff + nreads + 1);
  blk->h_rlen.assign(rlen, rlen + nreads);
  CUDA_CHECK(cudaStreamSynchronize(stream));
  TRACE("upload_packed: copies + unpack");
  dfree(d_packed); dfree(d_poff);
  return blk;
}

void set_block_mask(DeviceBlock *blk, const int64_t *mask_off, const int32_t *mask_pts,
                    cudaStream_t stream)
{ if (mask_off == nullptr)
    return;
  const int n = blk->nreads;
  blk->nmask = mask_off[n];
  blk->mask_off = dalloc<int64_t>((size_t) n + 1);
  blk->mask_pts = dalloc<int32_t>((size_t) blk->nmask + 1);
  CUDA_CHECK(cudaMemcpyAsync(blk->mask_off, mask_off, sizeof(int64_t) * ((size_t) n + 1),
                             cudaMemcpyHostToDevice, stream));
  if (blk->nmask > 0)
    CUDA_CHECK(cudaMemcpyAsync(blk->mask_pts, mask_pts, sizeof(int32_t) * (size_t) blk->nmask,
                               cudaMemcpyHostToDevice, stream));
  CUDA_CHECK(cudaStreamSynchronize(stream));
}

// Mask intervals of a complemented block (complement_DB, damapper.c:471-522): the point list of
// every read is reversed and each point x becomes rlen - x.  One thread per read.
__global__ void k_mirror_mask(const int64_t *__restrict__ mask_off, int32_t *mask_pts,
                              const int32_t *__restrict__ rlen, int nreads)
{ const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nreads) return;
  const int len = rlen[r];
  int64_t k = mask_off[r], j = mask_off[r + 1] - 1;
  while (k < j)
    { const int x = mask_pts[j];
      mask_pts[j--] = len - mask_pts[k];
      mask_pts[k++] = len - x;
    }
  if (k == j)
    mask_pts[k] = len - mask_pts[k];
}

// In-place reverse complement of every read (complement, damapper.c:417-431).  One thread per
// base position of the first half of its read: position q of read r swaps with its mirror.
__global__ void __launch_bounds__(256)
k_complement(uint8_t *bases, const int64_t *__restrict__ boff, int nreads, int64_t total)
{ for (int64_t q = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; q < total;
       q += (int64_t) gridDim.x * blockDim.x)
    { int lo = 0, hi = nreads - 1;                    // read holding position q
      while (lo < hi)
        { int mid = (lo + hi + 1) >> 1;
          if (boff[mid] <= q) lo = mid; else hi = mid - 1;
        }
      const int64_t b0 = boff[lo];
      const int len = (int) (boff[lo + 1] - b0 - 1);
      const int i = (int) (q - b0);
      if (i >= (len + 1) / 2)
        continue;                                     // second half or the terminator
      const int j = len - 1 - i;
      const uint8_t a = bases[b0 + i], b = bases[b0 + j];
      bases[b0 + i] = (uint8_t) (3 - b);
      bases[b0 + j] = (uint8_t) (3 - a);
    }
}

// Out-of-place reverse complement of every read, one warp per read (grid-stride): position i of
// the copy is 3 - base[len-1-i]; the terminators (and the leading 4) are copied as they are.
__global__ void __launch_bounds__(256)
k_revcomp_copy(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
               const int64_t *__restrict__ boff, int nreads)
{ const int lane = threadIdx.x & 31;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < nreads; r += nw)
    { const int64_t b0 = boff[r];
      const int len = (int) (boff[r + 1] - b0 - 1);
      const uint8_t *s = src + b0;
      uint8_t *d = dst + b0;
      if (lane == 0)
        { d[len] = s[len];
          if (r == 0) d[-1] = s[-1];
        }
      // head: bytes up to the first 4-byte aligned destination address
      const int headn = (int) ((4 - ((uintptr_t) d & 3)) & 3);
      const int h = headn < len ? headn : len;
      if (lane < h)
        d[lane] = (uint8_t) (3 - s[len - 1 - lane]);
      const int nwords = (len - h) >> 2;
      for (int w = lane; w < nwords; w += 32)
        { const int i = h + 4 * w;                  // destination bytes i..i+3 <- source len-1-i .. len-4-i
          const uint8_t *q = s + (len - 4 - i);
          const uint32_t v = (uint32_t) q[0] | ((uint32_t) q[1] << 8) | ((uint32_t) q[2] << 16) | ((uint32_t) q[3] << 24);
          *reinterpret_cast<uint32_t *>(d + i) = 0x03030303u - __byte_perm(v, 0, 0x0123);
        }
      const int done = h + 4 * nwords;
      if (lane < len - done)
        d[done + lane] = (uint8_t) (3 - s[len - 1 - (done + lane)]);
    }
}

void revcomp_copy_block(const DeviceBlock *blk, uint8_t *dst_bases, cudaStream_t stream)
{ if (blk->nreads == 0) return;
  int grid = (blk->nreads + 7) / 8;
  if (grid > sm_count() * 16) grid = sm_count() * 16;
  LAUNCH(k_revcomp_copy, grid, 256, 0, stream, blk->bases, dst_bases, blk->boff, blk->nreads);
}

void complement_block(DeviceBlock *blk, cudaStream_t stream)
{ if (blk->nreads == 0) return;
  int64_t nb = (blk->total + 255) / 256;
  int grid = nb < (int64_t) sm_count() * 16 ? (int) nb : sm_count() * 16;
  LAUNCH(k_complement, grid, 256, 0, stream, blk->bases, blk->boff, blk->nreads, blk->total);
  if (blk->mask_off != nullptr)
    LAUNCH(k_mirror_mask, (blk->nreads + 255) / 256, 256, 0, stream, blk->mask_off, blk->mask_pts,
           blk->rlen, blk->nreads);
}

void free_block(DeviceBlock *blk)
{ if (blk == nullptr) return;
  dfree(blk->raw); dfree(blk->boff); dfree(blk->rlen);
  dfree(blk->mask_off); dfree(blk->mask_pts); dfree(blk->tile_tab);
  delete blk;
}

// Sort_Kmers: returns a device list of *len records followed by the two sentinels
KmerIndex *sort_kmers(const DeviceBlock *blk, int K, int suppress, cudaStream_t stream)
{ const int     nreads = blk->nreads;
  const int64_t kmers64 = blk->total - (int64_t) K * nreads;
  KmerIndex    *idx = new KmerIndex();
  idx->src_uid = blk->uid;

  if (kmers64 <= 0)
    return idx;
  if (kmers64 > 0x7fffffffll)                          // `int kmers`, map.c:663,676
    fatal("Sort_Kmers: block holds %lld k-mers, more than 2^31-1", (long long) kmers64);
  if (kmers64 >= (1ll << 30))                          // look-back words of the radix pass are 32 bits wide
    fatal("Sort_Kmers: block holds %lld k-mers; this build sorts at most 2^30-1 per block: split the "
          "database into smaller blocks (DBsplit -s)", (long long) kmers64);
  for (int i = 0; i < nreads; i++)
    if (blk->h_rlen[i] < K)                            // damapper.c:403-410
      fatal("Sort_Kmers: block contains reads < %dbp long", K);
  const uint32_t n = (uint32_t) kmers64;

  int bytes[16], npass = 0;
  for (int i = 0; i < 2 * K; i += 8)                   // mersort, map.c:670-673
    bytes[npass++] = i >> 3;

  TRACE(nullptr);
  KmerPos  *a = dalloc<KmerPos>((size_t) n + 2);
  KmerPos  *b = dalloc<KmerPos>((size_t) n + 2);
  uint32_t *hist = dalloc<uint32_t>(256 * 16);
  CUDA_CHECK(cudaMemsetAsync(hist, 0, sizeof(uint32_t) * 256 * 16, stream));

  int grid = sm_count() * 4;
  int64_t ntiles = (blk->total + EX_TILE - 1) / EX_TILE;
  if (grid > ntiles) grid = (int) ntiles;
  cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
  if (g_time_kernels)
    { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
      cudaEventRecord(e0, stream);
    }
  ensure_tile_tab(blk, stream);
  LAUNCH(k_extract, grid, EX_THREADS, 0, stream, blk->bases, blk->boff, nreads, blk->total, K,
         npass, a, hist, blk->mask_off, blk->mask_pts, blk->tile_tab);
  uint32_t nlist = n;                                  // records that go into the sort
  if (blk->mask_off != nullptr)                        // -m: squeeze out the masked slots, in order
    { uint32_t *keep = dalloc<uint32_t>(n);
      int nb = (int) ((n + 2047) / 2048);
      uint32_t *bsum = dalloc<uint32_t>(nb + 1);
      LAUNCH(k_mask_flags, (n + 255) / 256, 256, 0, stream, a, (int64_t) n, keep);
      LAUNCH(k_scan_blocks, nb, 256, 0, stream, keep, (int64_t) n, bsum);
      LAUNCH(k_scan_bsum, 1, 1024, 0, stream, bsum, nb, bsum + nb);
      LAUNCH(k_compact, nb, 256, 0, stream, a, keep, (int64_t) n, bsum, b);
      CUDA_CHECK(cudaMemcpyAsync(&nlist, bsum + nb, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
      CUDA_CHECK(cudaStreamSynchronize(stream));
      dfree(keep); dfree(bsum);
      KmerPos *t = a; a = b; b = t;
    }
  TRACE("sort_kmers: alloc+extract");
  if (g_time_kernels) cudaEventRecord(e1, stream);
  KmerPos *rez = (KmerPos *) radix_sort16(a, b, nlist, bytes, npass, hist, stream);
  if (g_time_kernels)
    { cudaEventRecord(e2, stream);
      cudaEventSynchronize(e2);
      float t1, t2;
      cudaEventElapsedTime(&t1, e0, e1);
      cudaEventElapsedTime(&t2, e1, e2);
      idx->ms_extract = t1; idx->ms_sort = t2; idx->npass = npass;
      cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
    }
  TRACE("sort_kmers: radix");
  KmerPos *other = (rez == a) ? b : a;
  dfree(hist);

  uint32_t kept = nlist;
  if (suppress > 0 && nlist > 0)                       // map.c:726-770
    { const uint32_t n = nlist;
      uint32_t *keep = dalloc<uint32_t>(n);
      int nb = (int) ((n + 2047) / 2048);
      uint32_t *bsum = dalloc<uint32_t>(nb + 1);
      LAUNCH(k_suppress_flags, (n + 255) / 256, 256, 0, stream, rez, (int64_t) n, suppress, keep);
      LAUNCH(k_scan_blocks, nb, 256, 0, stream, keep, (int64_t) n, bsum);
      LAUNCH(k_scan_bsum, 1, 1024, 0, stream, bsum, nb, bsum + nb);
      LAUNCH(k_compact, nb, 256, 0, stream, rez, keep, (int64_t) n, bsum, other);
      CUDA_CHECK(cudaMemcpyAsync(&kept, bsum + nb, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
      CUDA_CHECK(cudaStreamSynchronize(stream));
      dfree(keep); dfree(bsum);
      KmerPos *t = rez; rez = other; other = t;
    }
  dfree(other);
  if (kept == 0)
    { dfree(rez);
      return idx;
    }
  LAUNCH(k_set_sentinels, 1, 1, 0, stream, rez, (int64_t) kept);
  TRACE("sort_kmers: tail");
  idx->list = rez;
  idx->len  = (int) kept;
  return idx;
}

void free_index(KmerIndex *idx)
{ if (idx == nullptr) return;
  dfree(idx->list);
  dfree(idx->lut);
  dfree(idx->filt_dsig);
  free_index(idx->filt);
  free_block(idx->block);
  delete idx;
}

}  // namespace damgpu
