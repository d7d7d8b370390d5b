// Sparse k-mer chaining and the candidate dominance filter (chain_thread, map.c:1463-1922).
//
// One thread per read: a read's (read, contig) seed groups are chained one after the other
// because the per-read candidate list they feed is updated sequentially (map.c:1668-1766).
// The reference's splay tree is an ordered set keyed (diag desc, apos desc); the result does
// not depend on its shape, so each thread keeps the active set as a
// sorted slice of a global scratch array (slot i of every scratch array belongs to seed i,
// hence no per-thread sizing).  Candidates and their Jump lists persist in device pools
// across Match_Filter calls; per-read list heads replace DAZZ_READ.coff (map.c:1875).
#include "common.cuh"
#include "mapper.cuh"

namespace damgpu {

constexpr int HITMIN = 3, MAX_GAP = 1000, MIN_PIECE = 300;   // map.c:34-37

struct ChainScratch                 // one slot per seed
{ int *from, *orig, *cost, *dead, *S, *E; };

__device__ __forceinline__ int seed_apos(const SeedPair &h) { return h.apos + 1; }   // map.c:1784

__global__ void __launch_bounds__(64)
k_chain(const SeedPair *__restrict__ hits, int64_t nhits, int nreads, int K, int bstart, int comp,
        int profile, int spacing, ChainScratch sc, Candidate *cand, int *cand_top, int cand_cap,
        uint32_t *jumps, unsigned long long *jump_top, unsigned long long jump_cap,
        int *head, int16_t *cover, const int64_t *__restrict__ coff, int *overflow)
{ const int ar = blockIdx.x * blockDim.x + threadIdx.x;
  if (ar >= nreads) return;

  // seed range of read ar (hits are sorted by aread first)
  int64_t lo = 0, hi = nhits;
  while (lo < hi)
    { int64_t mid = (lo + hi) >> 1;
      if (hits[mid].aread < ar) lo = mid + 1; else hi = mid;
    }
  int64_t nidx = lo;
  if (nidx >= nhits || hits[nidx].aread != ar) return;

  const int hithr = HITMIN * K;
  int chead = head[ar];

  while (nidx < nhits && hits[nidx].aread == ar)
    { const int     br = hits[nidx].bread;
      const int64_t g0 = nidx;                         // group base: node n <-> seed g0+n
      int *from = sc.from + g0, *orig = sc.orig + g0, *cost = sc.cost + g0, *dead = sc.dead + g0;
      int *S = sc.S + g0, *E = sc.E + g0;
      int nn = 0, ns = 0, nexp = 0, qhead = 0;

#define APOS(n) (hits[g0 + (n)].apos + 1)
#define DIAG(n) (hits[g0 + (n)].diag)
#define BPOS(n) (APOS(n) - DIAG(n))

      for ( ; nidx < nhits && hits[nidx].aread == ar && hits[nidx].bread == br; nidx++)
        { const int apos = hits[nidx].apos + 1;
          const int diag = hits[nidx].diag;
          const int bpos = apos - diag;
          int pos, l, r, lcost, rcost, j;

          while (qhead < nn && APOS(qhead) < apos - MAX_GAP)       // map.c:1787-1796
            { const int q = qhead++;
              if (!dead[q])
                { for (j = 0; j < ns; j++)
                    if (S[j] == q) break;
                  for ( ; j < ns - 1; j++) S[j] = S[j + 1];
                  ns -= 1;
                  if (orig[orig[q]] == q)
                    E[nexp++] = q;
                }
            }

          const int n = nn++;
          dead[n] = 0;
          for (pos = 0; pos < ns; pos++)              // key: diag desc, apos desc (map.c:1101)
            { const int x = S[pos];
              const int xd = DIAG(x);
              if (diag > xd || (diag == xd && apos > APOS(x)))
                break;
            }
          for (j = ns; j > pos; j--) S[j] = S[j - 1];
          S[pos] = n;
          ns += 1;

          l = -1;                                     // predOf + leftmost, map.c:1806-1808
          for (j = pos - 1; j >= 0; j--)
            if (BPOS(S[j]) >= bpos - MAX_GAP)
              { l = S[j];
                while (j > 0 && DIAG(S[j - 1]) == DIAG(l))
                  l = S[--j];
                break;
              }
          r = -1;                                     // succOf, map.c:1809
          for (j = pos + 1; j < ns; j++)
            if (BPOS(S[j]) <= bpos)
              { r = S[j];
                break;
              }

          lcost = rcost = 0;                          // map.c:1810-1826
          if (l >= 0)
            lcost = cost[l] + ((apos >= APOS(l) + K) ? K : apos - APOS(l));
          if (r >= 0)
            rcost = cost[r] + ((bpos >= BPOS(r) + K) ? K : bpos - BPOS(r));
          if (lcost > rcost)
            rcost = 0;
          else
            lcost = 0;

          if (lcost > 0 || rcost > 0)                 // map.c:1828-1857
            { const int p = (lcost > 0) ? l : r;
              const int c = (lcost > 0) ? lcost : rcost;
              from[n] = p;
              cost[n] = c;
              const int o = (from[p] < 0) ? p : orig[p];
              orig[n] = o;
              if (c >= cost[orig[o]])
                { int dd = DIAG(p) - diag;
                  orig[o] = n;
                  if (dd < 0) dd = -dd;
                  if ((double) dd <= .2 * (double) (apos - APOS(p)))
                    { for (j = 0; j < ns; j++)
                        if (S[j] == p) break;
                      for ( ; j < ns - 1; j++) S[j] = S[j + 1];
                      ns -= 1;
                      dead[p] = 1;
                    }
                }
            }
          else
            { from[n] = -1;
              cost[n] = K;
              orig[n] = n;
            }
        }

      // candidates of the group: live set in key order, then expired (newest first), map.c:1634-1767
      for (int pass = 0; pass < 2; pass++)
        for (int jj = 0; jj < (pass == 0 ? ns : nexp); jj++)
          { const int h = (pass == 0) ? S[jj] : E[nexp - 1 - jj];
            if (!(cost[h] >= hithr && orig[orig[h]] == h))
              continue;
            const int ab = APOS(orig[h]) - K, bb = BPOS(orig[h]) - K;
            const int ae = APOS(h), be = BPOS(h);
            const int hc = cost[h];

            if (profile)                              // map.c:1654-1666
              { int16_t *cnt = cover + coff[ar];
                const int tb = ab / spacing, te = (ae - 1) / spacing + 1;
                const int cb = cnt[tb], ce = cnt[te];
                if (cb < 0x7fff && ce > -0xffff)
                  { cnt[tb] = (int16_t) (cb + 1);
                    cnt[te] = (int16_t) (ce - 1);
                  }
              }

            int c = -1, d, e;                         // dominance filter, map.c:1675-1713
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
              continue;

            d = atomicAdd(cand_top, 1);
            if (d >= cand_cap) { *overflow = 1; return; }
            Candidate *D = cand + d;
            D->next = chead; chead = d;
            D->bread = br + bstart; D->comp = comp; D->score = hc;
            D->afirst = ab; D->alast = ae; D->bfirst = bb; D->blast = be;

            // chain_length, map.c:1243-1260 (the splices persist)
            int len = 0;
            { int x = h, y = from[h];
              while (y >= 0)
                { const int da = APOS(x) - APOS(y);
                  if (da == BPOS(x) - BPOS(y) && da < 100)
                    y = from[x] = from[y];
                  else
                    { len += 1; x = y; y = from[x]; }
                }
            }
            D->length = len;
            unsigned long long jo = 0;
            if (len > 0)
              { jo = atomicAdd(jump_top, (unsigned long long) len);
                if (jo + len > jump_cap) { *overflow = 1; return; }
                int g = h, k = 0;
                for (int f = from[h]; f >= 0; f = from[f])          // map.c:1746-1759
                  { const uint32_t da = (uint16_t) (APOS(g) - APOS(f));
                    const uint32_t db = (uint16_t) (BPOS(g) - BPOS(f));
                    jumps[jo + k++] = da | (db << 16);
                    g = f;
                  }
              }
            D->chain = (long long) jo;
          }
#undef APOS
#undef DIAG
#undef BPOS
    }
  head[ar] = chead;
}

static void ensure_pools(Mapper *m, int64_t nhits)
{ // a candidate is the best end of a distinct chain origin with >= 3 seeds and the from-paths
  // of distinct origins are disjoint: at most nhits/3 new candidates and nhits new jumps
  int h_ctop = 0; unsigned long long h_jtop = 0;
  CUDA_CHECK(cudaMemcpy(&h_ctop, m->cand_top, sizeof(int), cudaMemcpyDeviceToHost));
  CUDA_CHECK(cudaMemcpy(&h_jtop, m->jump_top, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
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
  CUDA_CHECK(cudaMemset(m->head, 0xff, sizeof(int) * ((size_t) n + 1)));
  CUDA_CHECK(cudaMemset(m->cand_top, 0, sizeof(int) * 4));
  CUDA_CHECK(cudaMemset(m->jump_top, 0, sizeof(unsigned long long) * 2));
  CUDA_CHECK(cudaMemset(m->overflow, 0, sizeof(int)));
  CUDA_CHECK(cudaMemset(m->cover, 0, sizeof(int16_t) * ((size_t) m->h_coff[n] + 2)));
}

void mapper_free(Mapper *m)
{ if (m == nullptr) return;
  dfree(m->head); dfree(m->cand_top); dfree(m->jump_top); dfree(m->overflow);
  dfree(m->coff); dfree(m->cover); dfree(m->cand); dfree(m->jumps);
  delete m;
}

void chain_seeds(Mapper *m, const SeedSet *ss, int bstart, int comp, cudaStream_t stream)
{ const int64_t nhits = ss->nhits;
  if (nhits == 0) return;
  if (m->spacing != g_par.spacing)
    fatal("SPACING changed after the mapper was created");
  ensure_pools(m, nhits);
  ChainScratch sc;
  int *scratch = dalloc<int>((size_t) nhits * 6);
  sc.from = scratch; sc.orig = scratch + nhits; sc.cost = scratch + 2 * nhits;
  sc.dead = scratch + 3 * nhits; sc.S = scratch + 4 * nhits; sc.E = scratch + 5 * nhits;
  const int n = m->reads->nreads;
  LAUNCH(k_chain, (n + 63) / 64, 64, 0, stream, ss->hits, nhits, n, g_par.kmer, bstart, comp,
         g_par.profile, g_par.spacing, sc, m->cand, m->cand_top, m->cand_cap, m->jumps,
         m->jump_top, (unsigned long long) m->jump_cap, m->head, m->cover, m->coff, m->overflow);
  int ovf = 0;
  CUDA_CHECK(cudaMemcpyAsync(&ovf, m->overflow, sizeof(int), cudaMemcpyDeviceToHost, stream));
  CUDA_CHECK(cudaStreamSynchronize(stream));
  dfree(scratch);
  if (ovf)
    fatal("Match_Filter: candidate/jump pool overflow (internal sizing error)");
}

}  // namespace damgpu
