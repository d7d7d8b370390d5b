// Sparse k-mer chaining and the candidate dominance filter (chain_thread, map.c:1463-1922).
//
// One thread per read: a read's (read, contig) seed groups are chained one after the other
// because the per-read candidate list they feed is updated sequentially (map.c:1668-1766).
// The reference's splay tree is an ordered set keyed (diag desc, apos desc); the result does
// not depend on its shape, so each thread keeps the active set as a
// sorted slice of a global scratch array (slot i of every scratch array belongs to seed i,
// hence no per-thread sizing).  Candidates and their Jump lists persist in device pools
// across Match_Filter calls; per-read list heads replace DAZZ_READ.coff (map.c:1875).
#include <algorithm>
#include <vector>
#include "common.cuh"
#include "mapper.cuh"

namespace damgpu {

constexpr int HITMIN = 3, MAX_GAP = 1000, MIN_PIECE = 300;   // map.c:34-37

struct ChainScratch                 // one slot per seed
{ int *from, *orig, *cost, *dead, *E;
  int4 *S;                          // spill area of the active set (node, diag, apos, -)
};

constexpr int CH_WARPS = 4;         // reads in flight per CTA
constexpr int CH_SCAP  = 64;        // active-set entries kept in shared memory per warp (general path)
constexpr int CH_NCAP  = 1024;      // chain nodes kept in shared memory per warp (fast path)
constexpr int CH_CTAS  = 6;         // CTAs per SM the shared memory allows (37 KB each)

// A chain node of the fast path: one 8-byte shared-memory word (the general path keeps
// from/orig/cost/dead in global scratch).  A group on the fast path has at most CH_NCAP nodes and a
// node adds at most K <= 32 to the cost of its predecessor, so 16 bits hold every field.
struct __align__(8) NodeSm
{ unsigned short cost;
  short from;
  unsigned short od;                // orig | dead << 15
  unsigned short pad;
};
static_assert(32 * CH_NCAP <= 65535 && CH_NCAP <= 0x7fff, "NodeSm fields are 16 bits");

// node accessors of the two paths (same interface)
struct AccG
{ const SeedPair *hits; int64_t g0; int *from_; const int *orig_, *cost_;
  __device__ __forceinline__ int  apos(int n) const { return hits[g0 + n].apos + 1; }
  __device__ __forceinline__ int  bpos(int n) const { return hits[g0 + n].apos + 1 - hits[g0 + n].diag; }
  __device__ __forceinline__ int  from(int n) const { return from_[n]; }
  __device__ __forceinline__ void set_from(int n, int v) const { from_[n] = v; }
  __device__ __forceinline__ int  orig(int n) const { return orig_[n]; }
  __device__ __forceinline__ int  cost(int n) const { return cost_[n]; }
};
struct AccS
{ NodeSm *N; const SeedPair *hits; int64_t g0;
  __device__ __forceinline__ int  apos(int n) const { return __ldg(&hits[g0 + n].apos) + 1; }
  __device__ __forceinline__ int  bpos(int n) const
  { const int2 v = __ldg(reinterpret_cast<const int2 *>(hits + g0 + n));      // (diag, apos)
    return v.y + 1 - v.x;
  }
  __device__ __forceinline__ int  from(int n) const { return N[n].from; }
  __device__ __forceinline__ void set_from(int n, int v) const { N[n].from = (short) v; }
  __device__ __forceinline__ int  orig(int n) const { return N[n].od & 0x7fff; }
  __device__ __forceinline__ int  cost(int n) const { return N[n].cost; }
};

// remove the entry with key (d,a) from the sorted active set; all lanes, uniform
__device__ __forceinline__ void set_remove(int4 *S, int &ns, int d, int a, int lane)
{ int idx = -1;
  for (int base = 0; base < ns && idx < 0; base += 32)
    { const int j = base + lane;
      int4 e = (j < ns) ? S[j] : make_int4(-1, 0, 0, 0);
      const unsigned b = __ballot_sync(0xffffffffu, j < ns && e.y == d && e.z == a);
      if (b) idx = base + __ffs(b) - 1;
    }
  if (idx < 0) return;
  for (int base = idx; base < ns - 1; base += 32)
    { const int j = base + lane;
      int4 v = make_int4(0, 0, 0, 0);
      if (j + 1 < ns) v = S[j + 1];
      __syncwarp();
      if (j < ns - 1) S[j] = v;
      __syncwarp();
    }
  ns -= 1;
}

// Candidate test + dominance filter + Jump list for chain end h (map.c:1642-1767); one lane
template <class Acc>
__device__ void consider(const Acc nd, int h, int ar, int br,
                         int comp, int K, int profile, int spacing,
                         Candidate *cand, int *cand_top, int cand_cap,
                         uint32_t *jumps, unsigned long long *jump_top, unsigned long long jump_cap,
                         int &chead, int16_t *cover, const int64_t *coff, int *overflow)
{ const int oh = nd.orig(h);
  const int ab = nd.apos(oh) - K, bb = nd.bpos(oh) - K;
  const int ae = nd.apos(h), be = nd.bpos(h);
  const int hc = nd.cost(h);

  if (profile)                                  // map.c:1654-1666
    { int16_t *cnt = cover + coff[ar];
      const int tb = ab / spacing, te = (ae - 1) / spacing + 1;
      const int cb = cnt[tb], ce = cnt[te];
      if (cb < 0x7fff && ce > -0xffff)
        { cnt[tb] = (int16_t) (cb + 1);
          cnt[te] = (int16_t) (ce - 1);
        }
    }

  int c = -1, d, e;                             // dominance filter, map.c:1675-1713
  for (d = chead; d >= 0; d = e)
    { Candidate *D = cand + d;
      const bool A = (D->afirst < ab + MIN_PIECE && D->alast > ae - MIN_PIECE);
      const bool B = (ab < D->afirst + MIN_PIECE && ae > D->alast - MIN_PIECE);
      e = D->next;
      if (A && .9 * (double) D->score >= (double) hc)
        break;
      if (B && (double) D->score <= .9 * (double) hc)
        { if (c < 0) chead = e; else cand[c].next = e;
          D->next = -2;
        }
      else
        c = d;
    }
  if (d >= 0)
    return;

  d = atomicAdd(cand_top, 1);
  if (d >= cand_cap) { *overflow = 1; return; }
  Candidate *D = cand + d;
  D->next = chead; chead = d;
  D->bread = br; D->comp = comp; D->score = hc;
  D->afirst = ab; D->alast = ae; D->bfirst = bb; D->blast = be;

  int len = 0;                                  // chain_length, map.c:1243-1260 (splices persist)
  { int x = h, y = nd.from(h);
    int ax = nd.apos(x), bx = nd.bpos(x);
    while (y >= 0)
      { const int ay = nd.apos(y), by = nd.bpos(y);
        const int da = ax - ay;
        if (da == bx - by && da < 100)
          { y = nd.from(y);
            nd.set_from(x, y);
          }
        else
          { len += 1; x = y; ax = ay; bx = by; y = nd.from(x); }
      }
  }
  D->length = len;
  unsigned long long jo = 0;
  if (len > 0)
    { jo = atomicAdd(jump_top, (unsigned long long) len);
      if (jo + len > jump_cap) { *overflow = 1; return; }
      int g = h, k = 0;
      int ag = nd.apos(g), bg = nd.bpos(g);
      for (int f = nd.from(h); f >= 0; f = nd.from(f))    // map.c:1746-1759
        { const int af = nd.apos(f), bf = nd.bpos(f);
          const uint32_t da = (uint16_t) (ag - af);
          const uint32_t db = (uint16_t) (bg - bf);
          jumps[jo + k++] = da | (db << 16);
          g = f; ag = af; bg = bf;
        }
    }
  D->chain = (long long) jo;
}

// One warp per read, reads handed out by a counter (a read with its true location in this
// orientation carries ~10x the seeds of one without).  A (read, contig) group starts on the FAST
// path: the active set is one entry per lane in registers (insert / remove / predOf / leftmost /
// succOf are a ballot and a shuffle each) and the chain nodes sit in shared memory.  When the set
// would pass 32 entries or the group 1024 nodes, the state is written out and the GENERAL path takes
// over for the rest of the group: sorted slice in shared memory (spilling to global scratch), nodes
// in global scratch, the lanes scan the slice 32 entries at a time.  Both give the same result.
// Hand-out order of the reads: longest seed lists first (a read with its true location in this
// orientation carries ~10x the seeds of one without, and the kernel ends with its longest reads).
// first[r] = index of read r's first seed, bucket = bit length of its seed count; reads are ordered
// by descending bucket (order inside a bucket is irrelevant: reads are independent of each other).
__global__ void __launch_bounds__(256)
k_read_seed_spans(const SeedPair *__restrict__ hits, int64_t nhits, int nreads, int64_t *__restrict__ first,
                  int *__restrict__ bucket, int *bucket_count)
{ const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nreads) return;
  int64_t lo = 0, hi = nhits;
  while (lo < hi)
    { int64_t mid = (lo + hi) >> 1;
      if (hits[mid].aread < r) lo = mid + 1; else hi = mid;
    }
  const int64_t b = lo;
  hi = nhits;
  while (lo < hi)
    { int64_t mid = (lo + hi) >> 1;
      if (hits[mid].aread <= r) lo = mid + 1; else hi = mid;
    }
  const int64_t cnt = lo - b;
  first[r] = (cnt > 0) ? b : -1;
  const int k = (cnt > 0) ? 64 - __clzll((unsigned long long) cnt) : 0;      // 0..40
  bucket[r] = k;
  atomicAdd(&bucket_count[k], 1);
}

__global__ void k_read_order(const int *__restrict__ bucket, int nreads, int *bucket_count, int *__restrict__ order)
{ __shared__ int start[64];
  if (threadIdx.x == 0)
    { int run = 0;
      for (int k = 63; k >= 0; k--)                      // descending bit length
        { start[k] = run; run += bucket_count[k]; }
    }
  __syncthreads();
  for (int r = threadIdx.x; r < nreads; r += blockDim.x)
    order[atomicAdd(&start[bucket[r]], 1)] = r;
}

__global__ void __launch_bounds__(CH_WARPS * 32)
k_chain(const SeedPair *__restrict__ hits, int64_t nhits, int nreads, int K, int bstart, int comp,
        int profile, int spacing, ChainScratch sc, Candidate *cand, int *cand_top, int cand_cap,
        uint32_t *jumps, unsigned long long *jump_top, unsigned long long jump_cap,
        int *head, int16_t *cover, const int64_t *__restrict__ coff, int *overflow, int *read_counter,
        const int *__restrict__ order, const int64_t *__restrict__ first)
{ __shared__ int4   s_S[CH_WARPS][CH_SCAP];
  __shared__ NodeSm s_N[CH_WARPS][CH_NCAP];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int hithr = HITMIN * K;
  const unsigned FULL = 0xffffffffu;

  while (true)
  { int ar = 0;
    if (lane == 0) ar = atomicAdd(read_counter, 1);
    ar = __shfl_sync(FULL, ar, 0);
    if (ar >= nreads) break;
    ar = order[ar];                                    // ticket -> read, longest seed lists first
    int64_t nidx = first[ar];                          // seed range of read ar
    if (nidx < 0) break;                               // only reads without seeds are left

    int chead = head[ar];

    while (nidx < nhits && hits[nidx].aread == ar)
      { const int     br = hits[nidx].bread;
        const int64_t g0 = nidx;                         // group base: node n <-> seed g0+n
        int *from = sc.from + g0, *orig = sc.orig + g0, *cost = sc.cost + g0, *dead = sc.dead + g0;
        int *E = sc.E + g0;
        int4 *S = s_S[warp];
        NodeSm *N = s_N[warp];
        bool spilled = false, fast = true;
        int nn = 0, ns = 0, nexp = 0, qhead = 0;

        // ---- fast path ------------------------------------------------------------------------
        { int e_n = 0, e_d = 0, e_a = 0;                 // lane j < ns: j-th entry in key order
          int qd = 0, qa = 0;                            // (diag, apos) of node qhead when qhead < nn
          int64_t cbase = nidx;
          int4 hv = make_int4(0, 0, -1, -1);             // (diag, apos, bread, aread) of seed cbase+lane
          if (cbase + lane < nhits) hv = *reinterpret_cast<const int4 *>(hits + cbase + lane);
          unsigned same = __ballot_sync(FULL, hv.w == ar && hv.z == br);
          while (true)
            { int off = (int) (nidx - cbase);
              if (off == 32)
                { cbase = nidx; off = 0;
                  hv = make_int4(0, 0, -1, -1);
                  if (cbase + lane < nhits) hv = *reinterpret_cast<const int4 *>(hits + cbase + lane);
                  same = __ballot_sync(FULL, hv.w == ar && hv.z == br);
                }
              if (((same >> off) & 1u) == 0)             // end of the group
                break;
              if (ns == 32 || nn == CH_NCAP)             // hand over to the general path
                { for (int i = lane; i < nn; i += 32)
                    { const NodeSm x = N[i];
                      from[i] = x.from; orig[i] = x.od & 0x7fff; cost[i] = x.cost; dead[i] = x.od >> 15;
                    }
                  if (lane < ns) S[lane] = make_int4(e_n, e_d, e_a, 0);
                  __syncwarp();
                  fast = false;
                  break;
                }
              const int diag = __shfl_sync(FULL, hv.x, off);
              const int apos = __shfl_sync(FULL, hv.y, off) + 1;
              const int bpos = apos - diag;

              while (qhead < nn && qa < apos - MAX_GAP)  // map.c:1787-1796
                { const int qi = qhead++;
                  const unsigned short qod = N[qi].od;
                  const int xd = qd, xa = qa;
                  if (qhead < nn)                        // key of the next node to expire
                    { const int2 v = __ldg(reinterpret_cast<const int2 *>(hits + g0 + qhead));
                      qd = v.x; qa = v.y + 1;
                    }
                  if ((qod & 0x8000) == 0)
                    { const unsigned mt = __ballot_sync(FULL, lane < ns && e_d == xd && e_a == xa);
                      if (mt)
                        { const int idx = __ffs(mt) - 1;
                          const int t_n = __shfl_down_sync(FULL, e_n, 1), t_d = __shfl_down_sync(FULL, e_d, 1);
                          const int t_a = __shfl_down_sync(FULL, e_a, 1);
                          if (lane >= idx) { e_n = t_n; e_d = t_d; e_a = t_a; }
                          ns -= 1;
                        }
                      if ((N[qod & 0x7fff].od & 0x7fff) == qi)
                        { if (lane == 0) E[nexp] = qi;
                          nexp += 1;
                        }
                    }
                }
              if (qhead == nn)                           // the queue is empty: this seed is its head
                { qd = diag; qa = apos; }

              const int n = nn++;
              int pos;                                   // insert, key order diag desc, apos desc
              { const unsigned gt = __ballot_sync(FULL, lane < ns && (diag > e_d || (diag == e_d && apos > e_a)));
                pos = gt ? __ffs(gt) - 1 : ns;
                const int t_n = __shfl_up_sync(FULL, e_n, 1), t_d = __shfl_up_sync(FULL, e_d, 1);
                const int t_a = __shfl_up_sync(FULL, e_a, 1);
                if (lane > pos) { e_n = t_n; e_d = t_d; e_a = t_a; }
                if (lane == pos) { e_n = n; e_d = diag; e_a = apos; }
                ns += 1;
              }
              // predOf + leftmost (map.c:1806-1808), succOf (map.c:1809)
              int l = -1, ld = 0, la = 0, r = -1, rd = 0, ra = 0;
              { const unsigned lb = __ballot_sync(FULL, lane < pos && (e_a - e_d) >= bpos - MAX_GAP);
                if (lb)
                  { const int src = 31 - __clz(lb);
                    ld = __shfl_sync(FULL, e_d, src);
                    const unsigned eq = __ballot_sync(FULL, lane <= src && e_d == ld);
                    const int s2 = __ffs(eq) - 1;        // equal diagonals are contiguous: first of the run
                    l  = __shfl_sync(FULL, e_n, s2);
                    la = __shfl_sync(FULL, e_a, s2);
                  }
                const unsigned rb = __ballot_sync(FULL, lane > pos && lane < ns && (e_a - e_d) <= bpos);
                if (rb)
                  { const int src = __ffs(rb) - 1;
                    r  = __shfl_sync(FULL, e_n, src);
                    rd = __shfl_sync(FULL, e_d, src);
                    ra = __shfl_sync(FULL, e_a, src);
                  }
              }
              int lcost = 0, rcost = 0;                  // map.c:1810-1826
              if (l >= 0)
                lcost = N[l].cost + ((apos >= la + K) ? K : apos - la);
              if (r >= 0)
                { const int rb_ = ra - rd;
                  rcost = N[r].cost + ((bpos >= rb_ + K) ? K : bpos - rb_);
                }
              if (lcost > rcost)
                rcost = 0;
              else
                lcost = 0;

              if (lcost > 0 || rcost > 0)                // map.c:1828-1857
                { const int p = (lcost > 0) ? l : r;
                  const int c = (lcost > 0) ? lcost : rcost;
                  const int pd = (lcost > 0) ? ld : rd, pa = (lcost > 0) ? la : ra;
                  const NodeSm P = N[p];
                  const int o = (P.from < 0) ? p : (P.od & 0x7fff);
                  const unsigned short ood = N[o].od;
                  const bool best = (c >= N[ood & 0x7fff].cost);
                  bool drop = false;
                  if (best)
                    { int dd = pd - diag;
                      if (dd < 0) dd = -dd;
                      drop = ((double) dd <= .2 * (double) (apos - pa));
                    }
                  __syncwarp();
                  if (lane == 0)
                    { NodeSm x;
                      x.cost = (unsigned short) c; x.from = (short) p; x.od = (unsigned short) o; x.pad = 0;
                      N[n] = x;
                      if (best) N[o].od = (unsigned short) ((ood & 0x8000) | n);
                      if (drop) N[p].od = (unsigned short) (N[p].od | 0x8000);
                    }
                  if (drop)
                    { const unsigned mt = __ballot_sync(FULL, lane < ns && e_d == pd && e_a == pa);
                      if (mt)
                        { const int idx = __ffs(mt) - 1;
                          const int t_n = __shfl_down_sync(FULL, e_n, 1), t_d = __shfl_down_sync(FULL, e_d, 1);
                          const int t_a = __shfl_down_sync(FULL, e_a, 1);
                          if (lane >= idx) { e_n = t_n; e_d = t_d; e_a = t_a; }
                          ns -= 1;
                        }
                    }
                  __syncwarp();
                }
              else
                { __syncwarp();
                  if (lane == 0)
                    { NodeSm x;
                      x.cost = (unsigned short) K; x.from = -1; x.od = (unsigned short) n; x.pad = 0;
                      N[n] = x;
                    }
                  __syncwarp();
                }
              nidx += 1;
            }
          if (fast)                                      // live set to shared memory for the walk below
            { if (lane < ns) S[lane] = make_int4(e_n, e_d, e_a, 0);
              __syncwarp();
            }
        }

        // ---- general path (the rest of the group after a hand-over) ---------------------------
        if (!fast)
        for ( ; nidx < nhits; nidx++)
          { const SeedPair hp = hits[nidx];
            if (hp.aread != ar || hp.bread != br) break;
            const int apos = hp.apos + 1;
            const int diag = hp.diag;
            const int bpos = apos - diag;

            while (qhead < nn && hits[g0 + qhead].apos + 1 < apos - MAX_GAP)   // map.c:1787-1796
              { const int q = qhead++;
                if (!dead[q])
                  { set_remove(S, ns, hits[g0 + q].diag, hits[g0 + q].apos + 1, lane);
                    if (orig[orig[q]] == q)
                      { if (lane == 0) E[nexp] = q;
                        nexp += 1;
                      }
                  }
              }

            const int n = nn++;
            if (!spilled && ns + 1 > CH_SCAP)            // active set outgrew shared memory
              { int4 *G = sc.S + g0;
                for (int j = lane; j < ns; j += 32) G[j] = S[j];
                __syncwarp();
                S = G; spilled = true;
              }

            // insert position: key order diag desc, apos desc (add, map.c:1101)
            int pos = ns;
            for (int base = 0; base < ns; base += 32)
              { const int j = base + lane;
                int4 e = (j < ns) ? S[j] : make_int4(0, 0, 0, 0);
                const unsigned b = __ballot_sync(0xffffffffu,
                                                 j < ns && (diag > e.y || (diag == e.y && apos > e.z)));
                if (b) { pos = base + __ffs(b) - 1; break; }
              }
            for (int top = ns; top > pos; top -= 32)     // shift right, top chunk first
              { const int j = top - 1 - lane;
                int4 v = make_int4(0, 0, 0, 0);
                if (j >= pos) v = S[j];
                __syncwarp();
                if (j >= pos) S[j + 1] = v;
                __syncwarp();
              }
            if (lane == 0) S[pos] = make_int4(n, diag, apos, 0);
            __syncwarp();
            ns += 1;

            // predOf + leftmost (map.c:1806-1808): nearest predecessor with bpos' >= bpos-MAX_GAP,
            // replaced by the max-apos node on its diagonal
            int l = -1, lj = -1, ld = 0, la = 0;
            for (int base = pos - 1; base >= 0 && l < 0; base -= 32)
              { const int j = base - lane;
                int4 e = (j >= 0) ? S[j] : make_int4(0, 0, 0, 0);
                const unsigned b = __ballot_sync(0xffffffffu, j >= 0 && (e.z - e.y) >= bpos - MAX_GAP);
                if (b)
                  { const int src = __ffs(b) - 1;
                    lj = base - src;
                    l  = __shfl_sync(0xffffffffu, e.x, src);
                    ld = __shfl_sync(0xffffffffu, e.y, src);
                    la = __shfl_sync(0xffffffffu, e.z, src);
                  }
              }
            if (l >= 0)
              for (int base = lj - 1; base >= 0; base -= 32)
                { const int j = base - lane;
                  int4 e = (j >= 0) ? S[j] : make_int4(0, 0, 0, 0);
                  const unsigned b = __ballot_sync(0xffffffffu, j >= 0 && e.y == ld);
                  const int run = (b == 0xffffffffu) ? 32 : __ffs(~b) - 1;
                  if (run > 0)
                    { l  = __shfl_sync(0xffffffffu, e.x, run - 1);
                      la = __shfl_sync(0xffffffffu, e.z, run - 1);
                    }
                  if (run < 32) break;
                }
            // succOf (map.c:1809): nearest successor with bpos' <= bpos
            int r = -1, rd = 0, ra = 0;
            for (int base = pos + 1; base < ns && r < 0; base += 32)
              { const int j = base + lane;
                int4 e = (j < ns) ? S[j] : make_int4(0, 0, 0, 0);
                const unsigned b = __ballot_sync(0xffffffffu, j < ns && (e.z - e.y) <= bpos);
                if (b)
                  { const int src = __ffs(b) - 1;
                    r  = __shfl_sync(0xffffffffu, e.x, src);
                    rd = __shfl_sync(0xffffffffu, e.y, src);
                    ra = __shfl_sync(0xffffffffu, e.z, src);
                  }
              }

            int lcost = 0, rcost = 0;                   // map.c:1810-1826
            if (l >= 0)
              lcost = cost[l] + ((apos >= la + K) ? K : apos - la);
            if (r >= 0)
              { const int rb = ra - rd;
                rcost = cost[r] + ((bpos >= rb + K) ? K : bpos - rb);
              }
            if (lcost > rcost)
              rcost = 0;
            else
              lcost = 0;

            if (lcost > 0 || rcost > 0)                 // map.c:1828-1857
              { const int p = (lcost > 0) ? l : r;
                const int c = (lcost > 0) ? lcost : rcost;
                const int pd = (lcost > 0) ? ld : rd, pa = (lcost > 0) ? la : ra;
                const int o = (from[p] < 0) ? p : orig[p];
                const bool best = (c >= cost[orig[o]]);
                __syncwarp();
                if (lane == 0)
                  { from[n] = p; cost[n] = c; orig[n] = o; dead[n] = 0;
                    if (best) orig[o] = n;
                  }
                __syncwarp();
                if (best)
                  { int dd = pd - diag;
                    if (dd < 0) dd = -dd;
                    if ((double) dd <= .2 * (double) (apos - pa))
                      { set_remove(S, ns, pd, pa, lane);
                        if (lane == 0) dead[p] = 1;
                        __syncwarp();
                      }
                  }
              }
            else
              { if (lane == 0)
                  { from[n] = -1; cost[n] = K; orig[n] = n; dead[n] = 0; }
                __syncwarp();
              }
          }

        // candidates of the group: live set in key order, then expired (newest first), map.c:1634-1767
        __syncwarp();
        if (lane == 0)
          { const AccS as = { N, hits, g0 };
            const AccG ag = { hits, g0, from, orig, cost };
            for (int pass = 0; pass < 2; pass++)
              for (int jj = 0; jj < (pass == 0 ? ns : nexp); jj++)
                { const int h = (pass == 0) ? S[jj].x : E[nexp - 1 - jj];
                  if (fast)
                    { if (as.cost(h) >= hithr && as.orig(as.orig(h)) == h)
                        consider(as, h, ar, br + bstart, comp, K, profile, spacing,
                                 cand, cand_top, cand_cap, jumps, jump_top, jump_cap, chead, cover, coff,
                                 overflow);
                    }
                  else
                    { if (cost[h] >= hithr && orig[orig[h]] == h)
                        consider(ag, h, ar, br + bstart, comp, K, profile, spacing,
                                 cand, cand_top, cand_cap, jumps, jump_top, jump_cap, chead, cover, coff,
                                 overflow);
                    }
                }
          }
        chead = __shfl_sync(0xffffffffu, chead, 0);
        __syncwarp();
      }
    if (lane == 0) head[ar] = chead;
  }
}

// ---- the chain kernel runs on its own stream -------------------------------------------------
// k_chain is bound by the latency of its longest read (about 1 ms on the bench workload whatever
// the number of reads) and leaves most of the machine idle; nothing that follows on the main
// stream until the next chain call or the Reporter touches what it reads or writes (seeds,
// scratch, candidate/jump pools, list heads, -p counters).  So it is launched on a second,
// non-blocking stream and the complement + index + merge-join of the next orientation overlap it.
// Buffers it uses are released only after the main stream has been made to wait for it.
struct ChainAsync
{ cudaStream_t stream = nullptr;
  cudaEvent_t  ready = nullptr, done = nullptr;
  bool         busy = false;
  std::vector<void *> pending;
};
static ChainAsync g_ca;

static void chain_release(cudaStream_t main_stream, bool host_wait)
{ if (!g_ca.busy)
    return;
  if (host_wait)
    CUDA_CHECK(cudaStreamSynchronize(g_ca.stream));
  else
    CUDA_CHECK(cudaStreamWaitEvent(main_stream, g_ca.done, 0));
  for (void *p : g_ca.pending)
    dfree(p);
  g_ca.pending.clear();
  g_ca.busy = false;
}

void chain_sync() { chain_release(0, true); }

static void ensure_pools(Mapper *m, int64_t nhits)
{ // a candidate is the best end of a distinct chain origin with >= 3 seeds and the from-paths
  // of distinct origins are disjoint: at most nhits/3 new candidates and nhits new jumps
  // Upper bounds kept on the host (no read-back): every call adds at most nhits/3 candidates and
  // nhits jumps.  The true tops are fetched only when a pool has to grow.
  int h_ctop = m->ctop_bound; unsigned long long h_jtop = m->jtop_bound;
  if ((int64_t) h_ctop + nhits / 3 + 16 > m->cand_cap || h_jtop + (uint64_t) nhits + 16 > m->jump_cap)
    { chain_sync();                                      // the pools are about to be read and moved
      CUDA_CHECK(cudaMemcpy(&h_ctop, m->cand_top, sizeof(int), cudaMemcpyDeviceToHost));
      CUDA_CHECK(cudaMemcpy(&h_jtop, m->jump_top, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    }
  int64_t need_c = (int64_t) h_ctop + nhits / 3 + 16;
  uint64_t need_j = h_jtop + (uint64_t) nhits + 16;
  if (need_c > 0x7ffffff0ll)
    fatal("Match_Filter: candidate pool exceeds 2^31 entries");
  if (need_c > m->cand_cap)
    { int64_t cap = need_c + need_c / 2;
      if (cap > 0x7ffffff0ll) cap = 0x7ffffff0ll;
      Candidate *nc = dalloc<Candidate>((size_t) cap);
      if (h_ctop > 0)
        CUDA_CHECK(cudaMemcpy(nc, m->cand, sizeof(Candidate) * (size_t) h_ctop, cudaMemcpyDeviceToDevice));
      dfree(m->cand);
      m->cand = nc; m->cand_cap = (int) cap;
    }
  if (need_j > m->jump_cap)
    { uint64_t cap = need_j + need_j / 2;
      uint32_t *nj = dalloc<uint32_t>((size_t) cap);
      if (h_jtop > 0)
        CUDA_CHECK(cudaMemcpy(nj, m->jumps, sizeof(uint32_t) * (size_t) h_jtop, cudaMemcpyDeviceToDevice));
      dfree(m->jumps);
      m->jumps = nj; m->jump_cap = cap;
    }
  m->ctop_bound = (int) std::min<int64_t>(need_c, 0x7ffffff0ll);
  m->jtop_bound = need_j;
}

Mapper *mapper_new(const DeviceBlock *reads)
{ Mapper *m = new Mapper();
  m->reads = reads;
  const int n = reads->nreads;
  m->head = dalloc<int>((size_t) n + 1);
  m->cand_top = dalloc<int>(4);
  m->jump_top = dalloc<unsigned long long>(2);
  m->overflow = dalloc<int>(1);
  std::vector<int64_t> coff(n + 1);
  int64_t tot = 0;
  for (int i = 0; i < n; i++)
    { coff[i] = tot;
      tot += (reads->h_rlen[i] - 1) / g_par.spacing + 2;
    }
  coff[n] = tot;
  m->h_coff = coff;
  m->coff = dalloc<int64_t>((size_t) n + 1);
  CUDA_CHECK(cudaMemcpy(m->coff, coff.data(), sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice));
  m->cover = dalloc<int16_t>((size_t) tot + 2);
  m->spacing = g_par.spacing;
  mapper_reset(m);
  return m;
}

void mapper_reset(Mapper *m)                           // start != 0, map.c:1574-1588,1491-1506
{ const int n = m->reads->nreads;
  chain_sync();
  CUDA_CHECK(cudaMemset(m->head, 0xff, sizeof(int) * ((size_t) n + 1)));
  CUDA_CHECK(cudaMemset(m->cand_top, 0, sizeof(int) * 4));
  CUDA_CHECK(cudaMemset(m->jump_top, 0, sizeof(unsigned long long) * 2));
  CUDA_CHECK(cudaMemset(m->overflow, 0, sizeof(int)));
  m->ctop_bound = 0; m->jtop_bound = 0;
  CUDA_CHECK(cudaMemset(m->cover, 0, sizeof(int16_t) * ((size_t) m->h_coff[n] + 2)));
}

void mapper_free(Mapper *m)
{ if (m == nullptr) return;
  chain_sync();
  dfree(m->head); dfree(m->cand_top); dfree(m->jump_top); dfree(m->overflow);
  dfree(m->coff); dfree(m->cover); dfree(m->cand); dfree(m->jumps);
  delete m;
}

void chain_seeds(Mapper *m, SeedSet *ss, int bstart, int comp, cudaStream_t stream, bool async)
{ const int64_t nhits = ss->nhits;
  if (nhits == 0) return;
  if (m->spacing != g_par.spacing)
    fatal("SPACING changed after the mapper was created");
  if (g_ca.stream == nullptr)
    { CUDA_CHECK(cudaStreamCreateWithFlags(&g_ca.stream, cudaStreamNonBlocking));
      CUDA_CHECK(cudaEventCreateWithFlags(&g_ca.ready, cudaEventDisableTiming));
      CUDA_CHECK(cudaEventCreateWithFlags(&g_ca.done, cudaEventDisableTiming));
    }
  TRACE(nullptr);
  chain_release(stream, false);                        // the previous call's buffers go back now
  ensure_pools(m, nhits);
  TRACE("chain: ensure_pools");
  ChainScratch sc;
  int *scratch = dalloc<int>((size_t) nhits * 5);
  sc.from = scratch; sc.orig = scratch + nhits; sc.cost = scratch + 2 * nhits;
  sc.dead = scratch + 3 * nhits; sc.E = scratch + 4 * nhits;
  sc.S = dalloc<int4>((size_t) nhits);
  const int n = m->reads->nreads;
  CUDA_CHECK(cudaEventRecord(g_ca.ready, stream));     // seeds, pools and scratch are ready
  CUDA_CHECK(cudaStreamWaitEvent(g_ca.stream, g_ca.ready, 0));
  int *read_counter = m->cand_top + 2;                  // reads are handed out by a counter
  CUDA_CHECK(cudaMemsetAsync(read_counter, 0, sizeof(int), g_ca.stream));
  int64_t *first = dalloc<int64_t>((size_t) n + 1);      // hand-out order: longest seed lists first
  int *order = dalloc<int>((size_t) 2 * n + 64);
  int *bucket = order + n, *bucket_count = order + 2 * n;
  CUDA_CHECK(cudaMemsetAsync(bucket_count, 0, sizeof(int) * 64, g_ca.stream));
  LAUNCH(k_read_seed_spans, (n + 255) / 256, 256, 0, g_ca.stream, ss->hits, nhits, n, first, bucket, bucket_count);
  LAUNCH(k_read_order, 1, 1024, 0, g_ca.stream, bucket, n, bucket_count, order);
  int grid = (n + CH_WARPS - 1) / CH_WARPS;
  if (grid > sm_count() * CH_CTAS) grid = sm_count() * CH_CTAS;
  LAUNCH(k_chain, grid, CH_WARPS * 32, 0, g_ca.stream, ss->hits, nhits, n, g_par.kmer,
         bstart, comp, g_par.profile, g_par.spacing, sc, m->cand, m->cand_top, m->cand_cap, m->jumps,
         m->jump_top, (unsigned long long) m->jump_cap, m->head, m->cover, m->coff, m->overflow, read_counter,
         order, first);
  CUDA_CHECK(cudaEventRecord(g_ca.done, g_ca.stream));
  g_ca.busy = true;
  g_ca.pending.push_back(scratch);
  g_ca.pending.push_back(sc.S);
  g_ca.pending.push_back(first);
  g_ca.pending.push_back(order);
  if (async)                                           // the seeds now belong to the pending call
    { g_ca.pending.push_back(ss->hits);
      ss->hits = nullptr;
    }
  else
    chain_sync();
  TRACE("chain: kernel");
  // m->overflow (internal sizing error, cannot happen with the bounds above) is checked by the
  // Reporter, which has to synchronise anyway
}

}  // namespace damgpu
