// Wave (O(ND)) local alignment with on-the-fly trace points, one warp per candidate chain
// (forward_wave align.c:353-1011, reverse_wave :1015-1720, Local_Alignment :1727-1946,
// the seeding loop of report_thread map.c:2487-2579).
//
// Parallel formulation: within a wave every diagonal depends only on the
// previous wave, so lane j of the warp owns the j-th diagonal in the reference's scan order
// (chunks of 32 when the band is wider) and the three order-dependent reductions -- the chain of
// strict record breakers (besta/lasta/trim*), aclip (last writer) and bclip (extreme k) -- are
// resolved with ballots in that order.  Per-diagonal state (V, M, T, HA, HB double-buffered;
// NA, NB) lives in a circular window indexed by k & (W-1): shared memory (W=128) in the normal
// kernel, global memory (W=8192) in the overflow kernel that re-runs the rare job whose band or
// Pebble pool outgrew the fast configuration.  Pebble cells are bump-allocated per warp with a
// warp prefix sum; cell indices differ from the reference's, the link structure (hence the
// traces) does not.
#include "common.cuh"
#include "mapper.cuh"
#include "align.cuh"

namespace damgpu {

constexpr int      TRIM_LEN = 15, DUB_TRIM = 45, PATH_LEN = 60;     // align.c:162-176
constexpr uint64_t PATH_TOP = 0x1000000000000000ull, PATH_INT = 0x0fffffffffffffffull;
constexpr int      TRIM_MASK = 0x7fff, TRIM_MLAG = 250, WAVE_LAG = 30;
constexpr int      IMAX = 0x7fffffff;

enum { ERR_NONE = 0, ERR_BAND = 1, ERR_CELLS = 2, ERR_TRACE = 3, ERR_MULTI = 4 };

struct __align__(16) Pebble { int ptr, diag, diff, mark; };           // align.c:344-349

struct WaveMem
{ int      *V[2], *M[2], *HA[2], *HB[2], *NA, *NB;
  uint64_t *T[2];
  int       wmask, wsize;
  Pebble   *cells;
  int       cmax;
  uint16_t *tbuf;                       // 4*tcap entries: A trace window, then B trace window
  int       tcap;
};

struct PathD { int abpos, bbpos, aepos, bepos, diffs, tlen; uint16_t *trace; };

struct WaveStats { unsigned long long nwaves, ncells, nalign, empty; };

template <int DIR> __device__ __forceinline__ bool LT(int a, int b) { return DIR > 0 ? a < b : a > b; }
template <int DIR> __device__ __forceinline__ bool GE(int a, int b) { return DIR > 0 ? a >= b : a <= b; }

// A sequence of the 2-bit packed block image (k_pack2bit, report.cu): base x of the sequence is bit
// pair s + x of the word stream w.  The separators of the byte image carry no information here;
// the ends of the sequence are its length.
struct PSeq { const uint32_t *__restrict__ w; int s, len; };

// sixteen consecutive bases starting at bit pair p (any sign: the images have slack on both sides)
__device__ __forceinline__ uint32_t win16(const uint32_t *__restrict__ w, int p)
{ const int i = p >> 4;
  return __funnelshift_r(__ldg(w + i), __ldg(w + i + 1), (unsigned) (p & 15) * 2);
}

// Slide along diagonal k from b-coordinate y while bases match (align.c:748-768 / 1403-1423).
// Returns the new y; hit: 1 = ran into the end of B, 2 = into the end of A.  Sixteen bases per
// step: the first differing bit pair of the xor says where the run of matches stops, the distance
// to the nearer sequence end caps it.  The byte loop of the reference tests B's terminator first,
// then the mismatch, then A's terminator: a run that reaches both ends together reports B's.
template <int DIR>
__device__ __forceinline__ int slide(const PSeq &A, const PSeq &B, int k, int y, int &hit)
{ int lim, run = 0;
  int pa = A.s + y + k, pb = B.s + y;
  if (DIR > 0)
    lim = min(B.len - y, A.len - (y + k));
  else
    { lim = min(y, y + k);
      pa -= 16; pb -= 16;
    }
  while (run < lim)
    { const uint32_t x = win16(A.w, pa) ^ win16(B.w, pb);
      if (x != 0)
        { run += (DIR > 0) ? ((__ffs(x) - 1) >> 1) : (__clz(x) >> 1);
          break;
        }
      run += 16;
      pa += 16 * DIR; pb += 16 * DIR;
    }
  if (run > lim) run = lim;
  y += DIR * run;
  if (DIR > 0)
    hit = (y == B.len) ? 1 : ((y + k == A.len) ? 2 : 0);
  else
    hit = (y == 0) ? 1 : ((y + k == 0) ? 2 : 0);
  return y;
}

// One forward (DIR=+1) or reverse (DIR=-1) extension from anti-diagonal mida on diagonal k0.
// All lanes return the same value; path fields are warp-uniform.
template <int DIR, bool DOB>
__device__ int wave(const WaveMem &wm, const AlignSpecD &sp, const PSeq &aseq,
                    const PSeq &bseq, PathD &apath, PathD &bpath, int k0, int mida,
                    int aoff, int boff, int *start_diag, WaveStats &st)
{ const int lane = threadIdx.x & 31;
  const int TS = sp.spacing, PATH_AVE = sp.ave_path;
  const int SENT = (DIR > 0) ? -1 : IMAX;
  const int WM = wm.wmask;
  Pebble *cells = wm.cells;
#define IX(k) ((k) & WM)

  int hgh = k0, low = k0, dif = 0, avail = 0, cur = 0;
  // Register path: while the band (with its two new diagonals) spans at most 32 diagonals, lane
  // (k & 31) keeps diagonal k's state in registers and neighbours are reached with shuffles.
  int rV = SENT, rM = 0, rHA = 0, rHB = 0, rNA = 0, rNB = 0;
  uint64_t rT = 0;
  bool generic = false;
  int more = 1, aclip, bclip;
  int besta, besty, trima, trimy, trimd, trimha, trimhb;
  int morea, morey, mored, moreha, morehb, morem, lasta;

  besta = trima = morea = lasta = mida;
  besty = trimy = morey = (mida - hgh) >> 1;
  trimd = mored = 0;
  trimha = moreha = 0;
  trimhb = morehb = 1;
  morem = -1;
  aclip = (DIR > 0) ? IMAX : -IMAX;
  bclip = (DIR > 0) ? -IMAX : IMAX;

  // ---- wave 0 on the single start diagonal (align.c:433-556 / 1093-1214), computed redundantly
  { const int k = k0;
    int y = (mida - k) >> 1, c, ha, hb, na, nb, hit;
    if (DIR > 0)
      { na = (((y + k) + (TS - aoff)) / TS - 1) * TS + aoff;
        nb = ((y + (TS - boff)) / TS - 1) * TS + boff;
        if (lane == 0)
          { cells[0] = Pebble{ -1, k, 0, na };
            cells[1] = Pebble{ -1, k, 0, nb };
          }
        na += TS; nb += TS;
      }
    else
      { na = (((y + k) + (TS - aoff) - 1) / TS - 1) * TS + aoff;
        nb = ((y + (TS - boff) - 1) / TS - 1) * TS + boff;
        if (lane == 0)
          { cells[0] = Pebble{ -1, k, 0, y + k };
            cells[1] = Pebble{ -1, k, 0, y };
          }
      }
    ha = 0; hb = 1; avail = 2;
    y = slide<DIR>(aseq, bseq, k, y, hit);
    if (hit)
      { more = 0;
        if (hit == 1) bclip = k; else aclip = k;
      }
    c = (y << 1) + k;
    while (GE<DIR>(y + k, na))
      { if (avail >= wm.cmax) return ERR_CELLS;
        if (lane == 0) cells[avail] = Pebble{ ha, k, 0, na };
        ha = avail++;
        na += DIR * TS;
      }
    while (GE<DIR>(y, nb))
      { if (avail >= wm.cmax) return ERR_CELLS;
        if (lane == 0) cells[avail] = Pebble{ hb, k, 0, nb };
        hb = avail++;
        nb += DIR * TS;
      }
    if (LT<DIR>(besta, c))
      { besta = trima = lasta = c;
        besty = trimy = y;
        trimha = ha;
        trimhb = hb;
      }
    rV = c; rT = PATH_INT; rM = PATH_LEN; rHA = ha; rHB = hb; rNA = na; rNB = nb;
  }

  // boundary handling after a wave (align.c:558-583,848-875 / 1216-1241,1502-1529)
#define GETM(k)  (generic ? wm.M[cur][IX(k)]  : __shfl_sync(0xffffffffu, rM, (k) & 31))
#define GETV(k)  (generic ? wm.V[cur][IX(k)]  : __shfl_sync(0xffffffffu, rV, (k) & 31))
#define GETHA(k) (generic ? wm.HA[cur][IX(k)] : __shfl_sync(0xffffffffu, rHA, (k) & 31))
#define GETHB(k) (generic ? wm.HB[cur][IX(k)] : __shfl_sync(0xffffffffu, rHB, (k) & 31))
#define CLIP_AFTER_WAVE(SETD)                                                                   \
  if (more == 0)                                                                                \
    { const int o_ = (DIR > 0) ? 0 : -1;                                                        \
      const int yb_ = besty + o_, xa_ = besta - besty + o_;    /* neither is a terminator */     \
      if (yb_ >= 0 && yb_ < bseq.len && xa_ >= 0 && xa_ < aseq.len)                             \
        more = 1;                                                                               \
      const bool aclipped = (DIR > 0) ? (hgh >= aclip) : (low <= aclip);                        \
      if (aclipped)                                                                             \
        { if (DIR > 0) hgh = aclip - 1; else low = aclip + 1;                                   \
          const int m_ = GETM(aclip);                                                           \
          if (morem <= m_)                                                                      \
            { morem = m_; morea = GETV(aclip);                                                  \
              morey = (morea - aclip) / 2; SETD                                                 \
              moreha = GETHA(aclip); morehb = GETHB(aclip);                                     \
            }                                                                                   \
        }                                                                                       \
      const bool bclipped = (DIR > 0) ? (low <= bclip) : (hgh >= bclip);                        \
      if (bclipped)                                                                             \
        { if (DIR > 0) low = bclip + 1; else hgh = bclip - 1;                                   \
          const int m_ = GETM(bclip);                                                           \
          if (morem <= m_)                                                                      \
            { morem = m_; morea = GETV(bclip);                                                  \
              morey = (morea - bclip) / 2; SETD                                                 \
              moreha = GETHA(bclip); morehb = GETHB(bclip);                                     \
            }                                                                                   \
        }                                                                                       \
      aclip = (DIR > 0) ? IMAX : -IMAX;                                                         \
      bclip = (DIR > 0) ? -IMAX : IMAX;                                                         \
    }

  CLIP_AFTER_WAVE(;)

  // ---- successive waves, register path (align.c:592-898 / 1248-1552)
  while (more && GE<DIR>(lasta, besta - DIR * TRIM_MLAG))
    { if (hgh < low)                    // empty band: the reference would read stale cells; stop
        { st.empty += 1;
          break;
        }
      if (hgh - low + 3 > 32)           // band outgrows the warp: continue in the window arrays
        { const int k = low + ((lane - low) & 31);
          if (k <= hgh)
            { wm.V[0][IX(k)] = rV; wm.T[0][IX(k)] = rT; wm.M[0][IX(k)] = rM;
              wm.HA[0][IX(k)] = rHA; wm.HB[0][IX(k)] = rHB; wm.NA[IX(k)] = rNA; wm.NB[IX(k)] = rNB;
            }
          __syncwarp();
          cur = 0;
          generic = true;
          break;
        }
      low -= 1;
      hgh += 1;
      dif += 1;
      const int  base = low;                           // diagonal of rotated-ballot bit 0
      const int  width = hgh - low + 1;
      const int  k = low + ((lane - low) & 31);
      const bool act = (k <= hgh);
      const int  lup = (lane + 1) & 31, ldn = (lane + 31) & 31;

      // new outer diagonals inherit NA/NB from their inner neighbour (align.c:678-690)
      { const int inner = (k == low) ? lup : ldn;
        const int nai = __shfl_sync(0xffffffffu, rNA, inner);
        if (act && (k == low || k == hgh)) rNA = nai;
        if (DOB)
          { const int nbi = __shfl_sync(0xffffffffu, rNB, inner);
            if (act && (k == low || k == hgh)) rNB = nbi;
          }
      }
      const int vold = (act && k > low && k < hgh) ? rV : SENT;
      const int lp = (DIR > 0) ? lup : ldn, ln = (DIR > 0) ? ldn : lup;   // lanes of k+DIR, k-DIR
      int vp = __shfl_sync(0xffffffffu, vold, lp);
      int vn = __shfl_sync(0xffffffffu, vold, ln);
      { const int kp = k + DIR, kn = k - DIR;
        if (kp < low || kp > hgh) vp = SENT;
        if (kn < low || kn > hgh) vn = SENT;
      }
      int c, srcl;
      if (LT<DIR>(vold, vn))                            // align.c:712-741 / 1367-1396
        { if (LT<DIR>(vn, vp)) { c = vp + DIR; srcl = lp; }
          else                 { c = vn + DIR; srcl = ln; }
        }
      else
        { if (LT<DIR>(vold, vp)) { c = vp + DIR; srcl = lp; }
          else                   { c = vold + 2 * DIR; srcl = lane; }
        }
      int m  = __shfl_sync(0xffffffffu, rM, srcl);
      int ha = __shfl_sync(0xffffffffu, rHA, srcl);
      int hb = 0;
      if (DOB) hb = __shfl_sync(0xffffffffu, rHB, srcl);
      uint64_t b;
      { const unsigned tl = __shfl_sync(0xffffffffu, (unsigned) rT, srcl);
        const unsigned th = __shfl_sync(0xffffffffu, (unsigned) (rT >> 32), srcl);
        b = ((uint64_t) th << 32) | tl;
      }
      int y = 0, hit = 0, cntA = 0, cntB = 0, skipA = 0, skipB = 0;
      if (act)
        { if ((b & PATH_TOP) != 0) m -= 1;
          b <<= 1;
          const int y0 = (c - k) >> 1;
          y = slide<DIR>(aseq, bseq, k, y0, hit);
          const int r = (DIR > 0) ? y - y0 : y0 - y;
          if (r > 0)                                    // closed form of align.c:764-767
            { const int rr = (r < 61) ? r : 61;
              const uint64_t mask = ((1ull << rr) - 1) << (61 - rr);
              m += rr - __popcll(b & mask);
              b = (r >= 64) ? ~0ull : ((b << r) | ((1ull << r) - 1));
            }
          c = (y << 1) + k;
          if (GE<DIR>(y + k, rNA))                      // align.c:771-793 / 1426-1448
            { const int over = DIR * (y + k - rNA);
              cntA = (over < TS) ? 1 : over / TS + 1;
              const int d0 = DIR * (cells[ha].mark - rNA);
              if (d0 >= 0)
                { skipA = (d0 < TS) ? 1 : d0 / TS + 1;
                  if (skipA > cntA) skipA = cntA;
                }
            }
          if (DOB && GE<DIR>(y, rNB))                   // align.c:795-817 / 1449-1471
            { const int over = DIR * (y - rNB);
              cntB = (over < TS) ? 1 : over / TS + 1;
              const int d0 = DIR * (cells[hb].mark - rNB);
              if (d0 >= 0)
                { skipB = (d0 < TS) ? 1 : d0 / TS + 1;
                  if (skipB > cntB) skipB = cntB;
                }
            }
        }
      else
        c = SENT;

      const int need = (cntA - skipA) + (cntB - skipB);
      if (__ballot_sync(0xffffffffu, need > 0))         // Pebble allocation (rare)
        { int pre = need;
          for (int o = 1; o < 32; o <<= 1)
            { const int t = __shfl_up_sync(0xffffffffu, pre, o);
              if (lane >= o) pre += t;
            }
          const int total = __shfl_sync(0xffffffffu, pre, 31);
          if (avail + total > wm.cmax) return ERR_CELLS;
          int idx = avail + pre - need;
          for (int i = skipA; i < cntA; i++)
            { cells[idx] = Pebble{ ha, k, dif, rNA + DIR * TS * i };
              ha = idx++;
            }
          for (int i = skipB; i < cntB; i++)
            { cells[idx] = Pebble{ hb, k, dif, rNB + DIR * TS * i };
              hb = idx++;
            }
          avail += total;
          __syncwarp();
        }
      if (act)
        { rNA += DIR * TS * cntA;
          if (DOB) { rNB += DIR * TS * cntB; rHB = hb; }
          rV = c; rT = b; rM = m; rHA = ha;
        }

#define ROT(x)      __funnelshift_r((x), (x), base & 31)           /* bit i <-> diagonal base+i */
#define LAST_SCAN(x)  ((DIR > 0) ? __ffs(x) - 1 : 31 - __clz(x))    /* last in scan order */
#define FIRST_SCAN(x) ((DIR > 0) ? 31 - __clz(x) : __ffs(x) - 1)
      // record breakers in scan order (align.c:819-833 / 1473-1487): prefix maximum along the
      // scan direction over the circular lane layout
      { const int cv = act ? ((DIR > 0) ? c : -c) : -IMAX;
        const int bv = (DIR > 0) ? besta : -besta;
        // only points beyond besta can break the record: walk them in scan order with a running
        // maximum (there are seldom more than two or three); bit i of bm <-> diagonal base+i
        unsigned bm = ROT(__ballot_sync(0xffffffffu, cv > bv));
        if (bm & (bm - 1))
          { unsigned rest = bm;
            int rm = bv;
            bm = 0;
            while (rest)
              { const int i = FIRST_SCAN(rest);
                rest &= ~(1u << i);
                const int t = __shfl_sync(0xffffffffu, cv, (base + i) & 31);
                if (t > rm) { rm = t; bm |= 1u << i; }
              }
          }
        const bool brk = act && ((bm >> (k - base)) & 1u);
        if (bm)
          { const bool good = brk && (m >= PATH_AVE);
            bool trim = false;
            if (good)
              { const int lo15 = (int) (b & TRIM_MASK), hi15 = (int) ((b >> TRIM_LEN) & TRIM_MASK);
                if (__ldg(sp.table + lo15) >= 0)
                  if (__ldg(sp.table + hi15) + __ldg(sp.score + lo15) >= 0)
                    trim = true;
              }
            const unsigned gm = __ballot_sync(0xffffffffu, good);
            const unsigned tm = __ballot_sync(0xffffffffu, trim);
            const int lb = (base + LAST_SCAN(bm)) & 31;
            besta = __shfl_sync(0xffffffffu, c, lb);
            besty = __shfl_sync(0xffffffffu, y, lb);
            if (gm)
              lasta = __shfl_sync(0xffffffffu, c, (base + LAST_SCAN(ROT(gm))) & 31);
            if (tm)
              { const int lt = (base + LAST_SCAN(ROT(tm))) & 31;
                trima  = __shfl_sync(0xffffffffu, c, lt);
                trimy  = __shfl_sync(0xffffffffu, y, lt);
                trimha = __shfl_sync(0xffffffffu, ha, lt);
                if (DOB) trimhb = __shfl_sync(0xffffffffu, hb, lt);
                trimd  = dif;
              }
          }
      }
      { const unsigned am = __ballot_sync(0xffffffffu, act && hit == 2);
        const unsigned bb = __ballot_sync(0xffffffffu, act && hit == 1);
        if (am | bb)
          { more = 0;
            if (am) aclip = base + LAST_SCAN(ROT(am));        // last writer in scan order
            if (bb) bclip = base + FIRST_SCAN(ROT(bb));       // extreme k towards the scan start
          }
      }

      CLIP_AFTER_WAVE(mored = dif;)

      // trim the band to within WAVE_LAG of the best point (align.c:877-885 / 1531-1539)
      { const int n = besta - DIR * WAVE_LAG;
        const unsigned g = __ballot_sync(0xffffffffu, act && k >= low && k <= hgh && !LT<DIR>(rV, n));
        if (g)
          { const unsigned rg = ROT(g);
            low = base + __ffs(rg) - 1;
            hgh = base + 31 - __clz(rg);
          }
        else
          hgh = low - 1;
      }
      st.nwaves += 1;
      st.ncells += (hgh - low) + 1;
    }
#undef ROT
#undef LAST_SCAN
#undef FIRST_SCAN

  // ---- successive waves, window path (band wider than the warp)
  while (generic && more && GE<DIR>(lasta, besta - DIR * TRIM_MLAG))
    { if (hgh < low)                    // empty band: the reference would read stale cells; stop
        { st.empty += 1;
          break;
        }
      low -= 1;
      hgh += 1;
      const int width = hgh - low + 1;
      if (width + 3 > wm.wsize) return ERR_BAND;
      const int nxt = cur ^ 1;
      if (lane == 0)
        { wm.NA[IX(low)] = wm.NA[IX(low + 1)]; wm.NB[IX(low)] = wm.NB[IX(low + 1)];
          wm.NA[IX(hgh)] = wm.NA[IX(hgh - 1)]; wm.NB[IX(hgh)] = wm.NB[IX(hgh - 1)];
          wm.V[cur][IX(low)] = SENT; wm.V[cur][IX(hgh)] = SENT;
          wm.V[cur][IX(low - 1)] = SENT; wm.V[cur][IX(hgh + 1)] = SENT;
        }
      __syncwarp();
      dif += 1;

      for (int s0 = 0; s0 < width; s0 += 32)
        { const int  s = s0 + lane;
          const bool act = (s < width);
          const int  k = (DIR > 0) ? hgh - s : low + s;
          int c = SENT, y = 0, m = 0, ha = 0, hb = 0, hit = 0;
          int na = 0, nb = 0, cntA = 0, cntB = 0, skipA = 0, skipB = 0;
          uint64_t b = 0;

          if (act)
            { const int kp = k + DIR, kn = k - DIR;
              const int vp = wm.V[cur][IX(kp)], vc = wm.V[cur][IX(k)], vn = wm.V[cur][IX(kn)];
              int src;
              if (LT<DIR>(vc, vn))                      // align.c:712-741 / 1367-1396
                { if (LT<DIR>(vn, vp)) { c = vp + DIR; src = kp; }
                  else                 { c = vn + DIR; src = kn; }
                }
              else
                { if (LT<DIR>(vc, vp)) { c = vp + DIR; src = kp; }
                  else                 { c = vc + 2 * DIR; src = k; }
                }
              m = wm.M[cur][IX(src)]; b = wm.T[cur][IX(src)];
              ha = wm.HA[cur][IX(src)]; hb = wm.HB[cur][IX(src)];

              if ((b & PATH_TOP) != 0) m -= 1;
              b <<= 1;

              const int y0 = (c - k) >> 1;
              y = slide<DIR>(aseq, bseq, k, y0, hit);
              const int r = (DIR > 0) ? y - y0 : y0 - y;
              if (r > 0)                                // closed form of align.c:764-767
                { const int rr = (r < 61) ? r : 61;
                  const uint64_t mask = ((1ull << rr) - 1) << (61 - rr);
                  m += rr - __popcll(b & mask);
                  b = (r >= 64) ? ~0ull : ((b << r) | ((1ull << r) - 1));
                }
              c = (y << 1) + k;

              na = wm.NA[IX(k)]; nb = wm.NB[IX(k)];
              if (GE<DIR>(y + k, na))                   // align.c:771-793 / 1426-1448
                { cntA = (DIR * (y + k - na)) / TS + 1;
                  const int d0 = DIR * (cells[ha].mark - na);  // coordinates the inherited path
                  if (d0 >= 0)                                 // has already crossed are skipped
                    skipA = (d0 / TS + 1 < cntA) ? d0 / TS + 1 : cntA;
                }
              if (GE<DIR>(y, nb))                       // align.c:795-817 / 1449-1471
                { cntB = (DIR * (y - nb)) / TS + 1;
                  const int d0 = DIR * (cells[hb].mark - nb);
                  if (d0 >= 0)
                    skipB = (d0 / TS + 1 < cntB) ? d0 / TS + 1 : cntB;
                }
            }

          // Pebble allocation: warp exclusive prefix sum of the cells each lane needs
          int need = (cntA - skipA) + (cntB - skipB), pre = need;
          for (int o = 1; o < 32; o <<= 1)
            { int t = __shfl_up_sync(0xffffffffu, pre, o);
              if (lane >= o) pre += t;
            }
          const int total = __shfl_sync(0xffffffffu, pre, 31);
          if (avail + total > wm.cmax) return ERR_CELLS;
          if (act)
            { int idx = avail + pre - need;
              for (int i = skipA; i < cntA; i++)
                { cells[idx] = Pebble{ ha, k, dif, na + DIR * TS * i };
                  ha = idx++;
                }
              for (int i = skipB; i < cntB; i++)
                { cells[idx] = Pebble{ hb, k, dif, nb + DIR * TS * i };
                  hb = idx++;
                }
              wm.NA[IX(k)] = na + DIR * TS * cntA;
              wm.NB[IX(k)] = nb + DIR * TS * cntB;
              wm.V[nxt][IX(k)] = c; wm.T[nxt][IX(k)] = b; wm.M[nxt][IX(k)] = m;
              wm.HA[nxt][IX(k)] = ha; wm.HB[nxt][IX(k)] = hb;
            }
          avail += total;

          // record breakers in scan order (align.c:819-833 / 1473-1487)
          const int cv = act ? ((DIR > 0) ? c : -c) : -IMAX;       // "larger is better"
          int pm = cv;
          for (int o = 1; o < 32; o <<= 1)
            { int t = __shfl_up_sync(0xffffffffu, pm, o);
              if (lane >= o && t > pm) pm = t;
            }
          int before = __shfl_up_sync(0xffffffffu, pm, 1);         // max over earlier lanes
          const int bv = (DIR > 0) ? besta : -besta;
          if (lane == 0 || before < bv) before = bv;
          const bool brk = act && (cv > before);
          const unsigned bm = __ballot_sync(0xffffffffu, brk);
          if (bm)
            { const bool good = brk && (m >= PATH_AVE);
              bool trim = false;
              if (good)
                { const int lo15 = (int) (b & TRIM_MASK), hi15 = (int) ((b >> TRIM_LEN) & TRIM_MASK);
                  if (__ldg(sp.table + lo15) >= 0)
                    if (__ldg(sp.table + hi15) + __ldg(sp.score + lo15) >= 0)
                      trim = true;
                }
              const unsigned gm = __ballot_sync(0xffffffffu, good);
              const unsigned tm = __ballot_sync(0xffffffffu, trim);
              const int lb = 31 - __clz(bm);
              besta = __shfl_sync(0xffffffffu, c, lb);
              besty = __shfl_sync(0xffffffffu, y, lb);
              if (gm)
                lasta = __shfl_sync(0xffffffffu, c, 31 - __clz(gm));
              if (tm)
                { const int lt = 31 - __clz(tm);
                  trima  = __shfl_sync(0xffffffffu, c, lt);
                  trimy  = __shfl_sync(0xffffffffu, y, lt);
                  trimha = __shfl_sync(0xffffffffu, ha, lt);
                  trimhb = __shfl_sync(0xffffffffu, hb, lt);
                  trimd  = dif;
                }
            }
          const unsigned am = __ballot_sync(0xffffffffu, act && hit == 2);
          const unsigned bb = __ballot_sync(0xffffffffu, act && hit == 1);
          if (am | bb) more = 0;
          if (am)                                       // last writer in scan order
            aclip = __shfl_sync(0xffffffffu, k, 31 - __clz(am));
          if (bb)                                       // k closest to the scan start
            { const int kb = __shfl_sync(0xffffffffu, k, __ffs(bb) - 1);
              if (DIR > 0) { if (bclip < kb) bclip = kb; }
              else         { if (bclip > kb) bclip = kb; }
            }
        }
      __syncwarp();
      cur = nxt;

      CLIP_AFTER_WAVE(mored = dif;)

      // trim the band to within WAVE_LAG of the best point (align.c:877-885 / 1531-1539)
      { const int n = besta - DIR * WAVE_LAG;
        int nh = low - 1, nl = hgh + 1;
        for (int s0 = 0; s0 <= hgh - low; s0 += 32)
          { const int  k = low + s0 + lane;
            const bool good = (k <= hgh) && !LT<DIR>(wm.V[cur][IX(k)], n);
            const unsigned g = __ballot_sync(0xffffffffu, good);
            if (g)
              { const int first = low + s0 + __ffs(g) - 1, last = low + s0 + 31 - __clz(g);
                if (first < nl) nl = first;
                if (last > nh) nh = last;
              }
          }
        if (nh >= nl) { hgh = nh; low = nl; }
        else          hgh = low - 1;
      }
      st.nwaves += 1;
      st.ncells += (hgh - low) + 1;
    }
#undef CLIP_AFTER_WAVE
#undef GETM
#undef GETV
#undef GETHA
#undef GETHB

  // ---- unwind the Pebble chains into trace pairs (align.c:900-1007 / 1554-1717), lane 0
  __syncwarp();
  int err = ERR_NONE;
  int r_x = 0, r_y = 0, r_d = 0, r_at = 0, r_bt = 0, r_start = k0;
  if (lane == 0)
    { uint16_t *atrace = apath.trace, *btrace = bpath.trace;
      uint16_t *const alo = wm.tbuf, *const ahi = wm.tbuf + 2 * wm.tcap;
      uint16_t *const blo = ahi, *const bhi = wm.tbuf + 4 * wm.tcap;
      int atlen = 0, btlen = 0, trimx, a, bq, k, h, d, e;

      if (morem >= 0)                                   // REACH = 1 (damapper.c:796)
        { trimx = morea - morey; trimy = morey; trimd = mored; trimha = moreha; trimhb = morehb; }
      else
        trimx = trima - trimy;

      // A chain
      a = -1;
      for (h = trimha; h >= 0; h = bq)
        { bq = cells[h].ptr; cells[h].ptr = a; a = h; }
      h = a;
      k = cells[h].diag;
      if (DIR > 0)
        { bq = (mida - k) / 2;
          e = 0;
          for (h = cells[h].ptr; h >= 0; h = cells[h].ptr)
            { k = cells[h].diag; a = cells[h].mark - k; d = cells[h].diff;
              if (atrace + atlen + 2 > ahi) { err = ERR_TRACE; break; }
              atrace[atlen++] = (uint16_t) (d - e);
              atrace[atlen++] = (uint16_t) (a - bq);
              bq = a; e = d;
            }
          if (!err)
            { if (bq + k != trimx)
                { atrace[atlen++] = (uint16_t) (trimd - e);
                  atrace[atlen++] = (uint16_t) (trimy - bq);
                }
              else if (bq != trimy)
                { atrace[atlen - 1] = (uint16_t) (atrace[atlen - 1] + (trimy - bq));
                  atrace[atlen - 2] = (uint16_t) (atrace[atlen - 2] + (trimd - e));
                }
            }
        }
      else
        { bq = cells[h].mark - k;
          e = 0; a = 0; d = 0;
          if ((bq + k) % TS != aoff)
            { h = cells[h].ptr;
              if (h < 0) { a = trimy; d = trimd; }
              else       { k = cells[h].diag; a = cells[h].mark - k; d = cells[h].diff; }
              if (apath.tlen == 0)
                { atrace[--atlen] = (uint16_t) (bq - a);
                  atrace[--atlen] = (uint16_t) (d - e);
                }
              else
                { atrace[1] = (uint16_t) (atrace[1] + (bq - a));
                  atrace[0] = (uint16_t) (atrace[0] + (d - e));
                }
              bq = a; e = d;
            }
          if (h >= 0)
            { for (h = cells[h].ptr; h >= 0; h = cells[h].ptr)
                { k = cells[h].diag; a = cells[h].mark - k;
                  if (atrace + atlen - 4 < alo) { err = ERR_TRACE; break; }
                  atrace[--atlen] = (uint16_t) (bq - a);
                  d = cells[h].diff;
                  atrace[--atlen] = (uint16_t) (d - e);
                  bq = a; e = d;
                }
              if (!err)
                { if (bq + k != trimx)
                    { atrace[--atlen] = (uint16_t) (bq - trimy);
                      atrace[--atlen] = (uint16_t) (trimd - e);
                    }
                  else if (bq != trimy)
                    { atrace[atlen + 1] = (uint16_t) (atrace[atlen + 1] + (bq - trimy));
                      atrace[atlen]     = (uint16_t) (atrace[atlen] + (trimd - e));
                    }
                }
            }
        }

      if (DOB)
        {
      // B chain
      a = -1;
      for (h = trimhb; h >= 0; h = bq)
        { bq = cells[h].ptr; cells[h].ptr = a; a = h; }
      h = a;
      k = cells[h].diag;
      if (DIR > 0)
        { bq = (mida + k) / 2;
          e = 0;
          r_start = k;                                  // low = k, align.c:971 -> *mind
          for (h = cells[h].ptr; h >= 0 && !err; h = cells[h].ptr)
            { k = cells[h].diag; a = cells[h].mark + k; d = cells[h].diff;
              if (btrace + btlen + 2 > bhi) { err = ERR_TRACE; break; }
              btrace[btlen++] = (uint16_t) (d - e);
              btrace[btlen++] = (uint16_t) (a - bq);
              bq = a; e = d;
            }
          if (!err)
            { if (bq - k != trimy)
                { btrace[btlen++] = (uint16_t) (trimd - e);
                  btrace[btlen++] = (uint16_t) (trimx - bq);
                }
              else if (bq != trimx)
                { btrace[btlen - 1] = (uint16_t) (btrace[btlen - 1] + (trimx - bq));
                  btrace[btlen - 2] = (uint16_t) (btrace[btlen - 2] + (trimd - e));
                }
            }
        }
      else
        { bq = cells[h].mark + k;
          e = 0;
          if ((bq - k) % TS != boff)
            { h = cells[h].ptr;
              if (h < 0) { a = trimx; d = trimd; }
              else       { k = cells[h].diag; a = cells[h].mark + k; d = cells[h].diff; }
              if (bpath.tlen == 0)
                { btrace[--btlen] = (uint16_t) (bq - a);
                  btrace[--btlen] = (uint16_t) (bq - a);         // sic, align.c:1670-1671 (H3)
                }
              else
                { btrace[1] = (uint16_t) (btrace[1] + (bq - a));
                  btrace[0] = (uint16_t) (btrace[0] + (d - e));
                }
              bq = a; e = d;
            }
          if (h >= 0)
            { for (h = cells[h].ptr; h >= 0 && !err; h = cells[h].ptr)
                { k = cells[h].diag; a = cells[h].mark + k;
                  if (btrace + btlen - 4 < blo) { err = ERR_TRACE; break; }
                  btrace[--btlen] = (uint16_t) (bq - a);
                  d = cells[h].diff;
                  btrace[--btlen] = (uint16_t) (d - e);
                  bq = a; e = d;
                }
              if (!err)
                { if (bq - k != trimy)
                    { btrace[--btlen] = (uint16_t) (bq - trimx);
                      btrace[--btlen] = (uint16_t) (trimd - e);
                    }
                  else if (bq != trimx)
                    { btrace[btlen + 1] = (uint16_t) (btrace[btlen + 1] + (bq - trimx));
                      btrace[btlen]     = (uint16_t) (btrace[btlen] + (trimd - e));
                    }
                }
            }
        }
        }
      r_x = trimx; r_y = trimy; r_d = trimd; r_at = atlen; r_bt = btlen;
    }
  err   = __shfl_sync(0xffffffffu, err, 0);
  r_x   = __shfl_sync(0xffffffffu, r_x, 0);   r_y  = __shfl_sync(0xffffffffu, r_y, 0);
  r_d   = __shfl_sync(0xffffffffu, r_d, 0);   r_at = __shfl_sync(0xffffffffu, r_at, 0);
  r_bt  = __shfl_sync(0xffffffffu, r_bt, 0);  r_start = __shfl_sync(0xffffffffu, r_start, 0);
  if (err) return err;

  if (DIR > 0)
    { apath.aepos = r_x; apath.bepos = r_y; apath.diffs = r_d;
      apath.tlen = r_at; bpath.tlen = r_bt;
      *start_diag = r_start;
    }
  else
    { apath.abpos = r_x; apath.bbpos = r_y; apath.diffs = apath.diffs + r_d;
      apath.tlen = apath.tlen - r_at; apath.trace = apath.trace + r_at;
      bpath.tlen = bpath.tlen - r_bt; bpath.trace = bpath.trace + r_bt;
    }
  __syncwarp();
  return ERR_NONE;
#undef IX
}

// Local_Alignment as damapper calls it: (dg,dg,ad,-1,-1), reach = 1 (align.c:1727-1946).
// DOB = false skips the B-side Pebble chain: the B path is only consumed with -C (map.c:2556-2572).
template <bool DOB>
__device__ int local_alignment(const WaveMem &wm, const AlignSpecD &sp, const PSeq &aseq, int alen,
                               const PSeq &bseq, int blen, int acomp, int dg, int anti,
                               PathD &apath, PathD &bpath, WaveStats &st)
{ const int lane = threadIdx.x & 31;
  int aoff = 0, boff = 0, low = dg, err;
  apath.trace = wm.tbuf + wm.tcap;                       // room on both sides
  bpath.trace = wm.tbuf + 3 * wm.tcap;
  apath.tlen = bpath.tlen = 0;
  apath.diffs = 0;
  apath.abpos = apath.bbpos = apath.aepos = apath.bepos = 0;
  bpath.abpos = bpath.bbpos = bpath.aepos = bpath.bepos = bpath.diffs = 0;
  if (((anti - dg) >> 1) < 0)
    return ERR_MULTI;                                    // never for a seed inside both sequences
  if (acomp)
    aoff = alen % sp.spacing;                            // align.c:1794-1797
  st.nalign += 1;

  if ((err = wave<1, DOB>(wm, sp, aseq, bseq, apath, bpath, dg, anti, aoff, boff, &low, st)) != 0)
    return err;
  const bool fshort = ((apath.aepos + apath.bepos) - anti < DUB_TRIM);
  if ((err = wave<-1, DOB>(wm, sp, aseq, bseq, apath, bpath, low, anti, aoff, boff, &low, st)) != 0)
    return err;
  const bool rshort = (anti - (apath.abpos + apath.bbpos) < DUB_TRIM);

  if (fshort)
    { if (rshort)
        { apath.aepos = apath.abpos = (apath.abpos + apath.aepos) / 2;
          apath.bepos = apath.bbpos = (apath.bbpos + apath.bepos) / 2;
          apath.tlen = 0;
          bpath.tlen = 0;
        }
      else
        { low = apath.abpos - apath.bbpos;
          anti = apath.abpos + apath.bbpos;
          apath.tlen = bpath.tlen = 0;
          if ((err = wave<1, DOB>(wm, sp, aseq, bseq, apath, bpath, low, anti, aoff, boff, &low, st)) != 0)
            return err;
        }
    }
  else if (rshort)
    { low = apath.aepos - apath.bepos;
      anti = apath.aepos + apath.bepos;
      apath.tlen = bpath.tlen = 0;
      apath.diffs = 0;
      if ((err = wave<-1, DOB>(wm, sp, aseq, bseq, apath, bpath, low, anti, aoff, boff, &low, st)) != 0)
        return err;
    }

  bpath.diffs = apath.diffs;
  bpath.aepos = apath.bepos; bpath.bepos = apath.aepos;  // align.c:1863-1866 / 1908-1911
  bpath.abpos = apath.bbpos; bpath.bbpos = apath.abpos;
  if (acomp)                                             // align.c:1858-1884
    { apath.abpos = alen - bpath.bepos;
      apath.bbpos = blen - bpath.aepos;
      apath.aepos = alen - bpath.bbpos;
      apath.bepos = blen - bpath.abpos;
      if (lane == 0)
        { uint16_t *trace = apath.trace, p;
          int i = apath.tlen - 2, j = 0;
          while (j < i)
            { p = trace[i]; trace[i] = trace[j]; trace[j] = p;
              p = trace[i + 1]; trace[i + 1] = trace[j + 1]; trace[j + 1] = p;
              i -= 2; j += 2;
            }
        }
      __syncwarp();
    }
  return ERR_NONE;
}

// ---- job kernel: one warp per candidate (map.c:2460-2579) ----------------------------------

template <bool BIG, bool DOB>
__global__ void __launch_bounds__(ALIGN_WARPS * 32, ALIGN_MINB)
k_align(AlignArgs A)
{ extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int gw = blockIdx.x * ALIGN_WARPS + wib;          // global warp id -> scratch slot

  WaveMem wm;
  { const int W = BIG ? ALIGN_W_BIG : ALIGN_W;
    unsigned char *base = BIG ? (A.big_state + (size_t) gw * ALIGN_STATE_BYTES(ALIGN_W_BIG))
                              : (smem + (size_t) wib * ALIGN_STATE_BYTES(ALIGN_W));
    uint64_t *t = reinterpret_cast<uint64_t *>(base);
    wm.T[0] = t; wm.T[1] = t + W;
    int *p = reinterpret_cast<int *>(t + 2 * W);
    wm.V[0] = p; wm.V[1] = p + W; wm.M[0] = p + 2 * W; wm.M[1] = p + 3 * W;
    wm.HA[0] = p + 4 * W; wm.HA[1] = p + 5 * W; wm.HB[0] = p + 6 * W; wm.HB[1] = p + 7 * W;
    wm.NA = p + 8 * W; wm.NB = p + 9 * W;
    wm.wsize = W; wm.wmask = W - 1;
    wm.cmax = BIG ? A.cells_big : A.cells_small;
    wm.cells = reinterpret_cast<Pebble *>(A.cells) + (size_t) gw * wm.cmax;
    wm.tcap = BIG ? A.tcap_big : A.tcap;
    wm.tbuf = A.tscratch + (size_t) gw * 4 * wm.tcap;
  }

  const int hithr = 3 * A.kmer;                          // HITMIN*Kmer, map.c:2419

  while (true)
    { WaveStats st = { 0, 0, 0, 0 };                     // of this job: counted only when the job succeeds
      int j = 0;
      if (lane == 0) j = atomicAdd(A.job_counter, 1);
      j = __shfl_sync(0xffffffffu, j, 0);
      if (j >= A.njobs) break;
      const int jid = A.job_list ? A.job_list[j] : j;
      AlignJob &job = A.jobs[jid];
      const Candidate cd = A.cand[job.cand];
      const int ar = job.read, br = cd.bread, cm = cd.comp;
      const int alen = A.rlen_a[ar], blen = A.rlen_b[br];
      PSeq aseq, bseq;
      { const int64_t ob = A.boff_b[br], oa = A.boff_a[ar];
        bseq.w = A.pk_b + (ob >> 4); bseq.s = (int) (ob & 15); bseq.len = blen;
        aseq.w = (cm ? A.pk_ac : A.pk_a) + (oa >> 4); aseq.s = (int) (oa & 15); aseq.len = alen;
      }

      int apos = cd.alast, bpos = cd.blast, alast = alen + 1;
      int first = -1, last = -1, count = 0, status = 0;
      for (int n = 0; n < cd.length; n++)
        { const uint32_t jp = A.jumps[cd.chain + n];
          apos -= (int) (jp & 0xffff);
          bpos -= (int) (jp >> 16);
          if (apos >= alast)
            continue;
          int dg, ad;
          if (cm) { const int ac = alen - apos, bc = blen - bpos; dg = ac - bc; ad = ac + bc; }
          else    { dg = apos - bpos; ad = apos + bpos; }
          PathD ap, bp;
          const int err = local_alignment<DOB>(wm, A.spec, aseq, alen, bseq, blen, cm, dg, ad, ap, bp, st);
          if (err)
            { status = err;
              break;
            }
          if (ap.aepos - ap.abpos < hithr)
            continue;
          alast = ap.abpos;
          // keep it: copy both paths and traces out of the warp scratch
          int rec = 0; long long to = 0;
          const int tl = ap.tlen + (A.do_b ? bp.tlen : 0);
          if (lane == 0)
            { rec = atomicAdd(A.aln_top, 1);
              to  = (long long) atomicAdd(A.trace_top, (unsigned long long) tl);
            }
          rec = __shfl_sync(0xffffffffu, rec, 0);
          to  = __shfl_sync(0xffffffffu, to, 0);
          if (rec >= A.aln_cap || to + tl > A.trace_cap)
            { status = ERR_TRACE + 10;                    // output pools too small: host grows them
              break;
            }
          for (int i = lane; i < ap.tlen; i += 32)
            A.traces[to + i] = ap.trace[i];
          if (A.do_b)
            for (int i = lane; i < bp.tlen; i += 32)
              A.traces[to + ap.tlen + i] = bp.trace[i];
          if (lane == 0)
            { AlnRec r;
              r.next = -1; r.comp = cm; r.bread = br; r.pad = 0;
              r.a[0] = ap.abpos; r.a[1] = ap.bbpos; r.a[2] = ap.aepos; r.a[3] = ap.bepos;
              r.a[4] = ap.diffs; r.a[5] = ap.tlen;
              r.b[0] = bp.abpos; r.b[1] = bp.bbpos; r.b[2] = bp.aepos; r.b[3] = bp.bepos;
              r.b[4] = bp.diffs; r.b[5] = bp.tlen;
              r.atrace = to; r.btrace = to + ap.tlen;
              A.alns[rec] = r;
              if (last >= 0) A.alns[last].next = rec;
            }
          if (first < 0) first = rec;
          last = rec;
          count += 1;
          __syncwarp();
        }
      if (lane == 0)
        { job.first = (status == 0) ? first : -1;
          job.count = (status == 0) ? count : 0;
          job.status = status;
          if (status != 0)
            atomicAdd(A.nfailed, 1);                     // re-run whole by the next round, which counts it
          else
            { atomicAdd(&A.stats[0], st.nalign); atomicAdd(&A.stats[1], st.nwaves);
              atomicAdd(&A.stats[2], st.ncells); atomicAdd(&A.stats[3], st.empty);
            }
        }
    }
}

void launch_align(const AlignArgs &A, bool big, int nblocks, cudaStream_t stream)
{ const bool dob = (A.do_b != 0);
  if (big)
    { if (dob) LAUNCH((k_align<true, true>), nblocks, ALIGN_WARPS * 32, 0, stream, A);
      else     LAUNCH((k_align<true, false>), nblocks, ALIGN_WARPS * 32, 0, stream, A);
    }
  else
    { const size_t smem = (size_t) ALIGN_WARPS * ALIGN_STATE_BYTES(ALIGN_W);
      if (dob) LAUNCH((k_align<false, true>), nblocks, ALIGN_WARPS * 32, smem, stream, A);
      else     LAUNCH((k_align<false, false>), nblocks, ALIGN_WARPS * 32, smem, stream, A);
    }
}

}  // namespace damgpu
