// Wave (O(ND)) local alignment, ONE THREAD PER CANDIDATE CHAIN ("lane kernel") -- the first tier of
// the alignment phase (forward_wave align.c:353-1011, reverse_wave :1015-1720, Local_Alignment
// :1727-1946, the seeding loop of report_thread map.c:2487-2579).
//
// Why lanes: after WAVE_LAG trimming a wave spans ~8 diagonals (max ~30 at 15% error), so a warp
// per alignment leaves most lanes idle and spends ~390 warp instructions per wave on cross-lane
// bookkeeping (align.cu).  Here every lane owns a whole alignment and walks its band serially,
// exactly like the scalar recurrence, which costs ~45 warp instructions per alignment-wave.
//
//  * Direction is data, not code: a reverse wave is a forward wave on negated coordinates
//    (k' = -k, c' = -c, y' = -y) over the sequences read backwards, so lanes extending in
//    opposite directions execute the same loop body.
//  * Per-diagonal state (V, T, M|HB|HA, NA[, NB]) lives in shared memory, word-interleaved by lane
//    (bank == lane: conflict-free whatever diagonal each lane touches), window of 64 diagonals.
//  * Pebble cells go to a per-job arena in global memory and are NOT unwound here: a lane only
//    records the chain heads of each wave call; k_unwind (one thread per kept alignment) turns
//    them into trace pairs afterwards, so pointer chasing never serialises a warp.
//  * A job that outgrows the window or its arena is marked failed and re-run by the warp kernels
//    of align.cu (the host loop in report.cu).
#include "common.cuh"
#include "mapper.cuh"
#include "align.cuh"

namespace damgpu {

namespace {

constexpr int      TRIM_LEN = 15, DUB_TRIM = 45, PATH_LEN = 60;     // align.c:162-176
constexpr uint64_t PATH_TOP = 0x1000000000000000ull, PATH_INT = 0x0fffffffffffffffull;
constexpr int      TRIM_MASK = 0x7fff, TRIM_MLAG = 250, WAVE_LAG = 30;
constexpr int      IMAX = 0x7fffffff;
constexpr int      LSENT = -0x3fffffff;                 // "no point on this diagonal"
constexpr int      LCELLS = 4095;                       // cells per wave call (12-bit handles)

enum { LERR_NONE = 0, LERR_BAND = 1, LERR_CELLS = 2, LERR_TRACE = 3, LERR_MULTI = 4, LERR_POOL = 13 };
enum { PH_IDLE = 0, PH_SEED, PH_START, PH_WAVE, PH_ENDCALL, PH_FINISH, PH_JOBEND, PH_DONE };

struct __align__(16) LPebble { int ptr, diag, diff, mark; };         // align.c:344-349

// eight bases seq[p], seq[p+dir], .., seq[p+7*dir], first one in the low byte
__device__ __forceinline__ uint64_t bases8(const uint8_t *seq, int p, int dir)
{ const uint8_t *s = seq + (dir > 0 ? p : p - 7);
  const uintptr_t a = reinterpret_cast<uintptr_t>(s);
  const uint64_t *w = reinterpret_cast<const uint64_t *>(a & ~(uintptr_t) 7);
  const unsigned sh = (unsigned) (a & 7) * 8;
  const uint64_t lo = w[0], hi = w[1];
  uint64_t v = sh ? ((lo >> sh) | (hi << (64 - sh))) : lo;
  if (dir < 0)
    { const uint32_t l = __byte_perm((uint32_t) (v >> 32), 0, 0x0123);
      const uint32_t h = __byte_perm((uint32_t) v, 0, 0x0123);
      v = ((uint64_t) h << 32) | l;
    }
  return v;
}

}  // namespace

// ---- the lane kernel ------------------------------------------------------------------------
template <bool DOB>
__global__ void __launch_bounds__(LANE_WARPS * 32, 1)
k_align_lane(AlignArgs A)
{ extern __shared__ uint32_t lsm[];
  constexpr int W = LANE_W, NF = DOB ? 6 : 5;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int *sm = reinterpret_cast<int *>(lsm) + (size_t) wib * LANE_WARP_WORDS(DOB);
  // Sequence windows: the next LANE_WIN*4 bases of A and of B in VIRTUAL order (the order the
  // wave reads them: ascending addresses forward, descending in a reverse wave), word-interleaved
  // by lane like the state.  Word j of a window holds virtual bytes u = 4j..4j+3 (u = v - phase,
  // phase chosen so that 16-byte chunks are aligned in memory).  One chunk per sequence is
  // prefetched per wave: the load is issued before the band loop and stored after it.
  uint32_t *const winA = reinterpret_cast<uint32_t *>(sm) + NF * W * 32;
  uint32_t *const winB = winA + LANE_WIN * 32;
  // good record breakers of the current wave (c, y, hb<<12|ha, T low, T high), up to 4 per lane: the
  // trim test needs three dependent table loads; deferring it to the end of the wave exposes their
  // latency once per wave instead of once per diagonal (some lane of the warp breaks the record
  // on almost every diagonal)
  int *const brk = reinterpret_cast<int *>(winB + LANE_WIN * 32);
#define BRK(e, f) brk[(((e) & 3) * 5 + (f)) * 32 + lane]
#define TRIM_TEST(tl_, th_, ok_)                                                            \
  { const uint64_t b_ = ((uint64_t) (uint32_t) (th_) << 32) | (uint32_t) (tl_);             \
    const int lo15 = (int) (b_ & TRIM_MASK), hi15 = (int) ((b_ >> TRIM_LEN) & TRIM_MASK);   \
    ok_ = false;                                                                            \
    if (__ldg(A.spec.table + lo15) >= 0)                                                    \
      if (__ldg(A.spec.table + hi15) + __ldg(A.spec.score + lo15) >= 0)                     \
        ok_ = true;                                                                         \
  }
#define WINW(win, j) (win)[(((j) & (LANE_WIN - 1)) << 5) + lane]
#define FLD(f, k) sm[((((f) * W) + ((k) & (W - 1))) << 5) + lane]
#define SV(k)  FLD(0, k)
#define STL(k) FLD(1, k)
#define STH(k) FLD(2, k)
#define SMH(k) FLD(3, k)               /* M << 24 | HB << 12 | HA */
#define SNA(k) FLD(4, k)
#define SNB(k) FLD(5, k)

  const int TS = A.spec.spacing, PATH_AVE = A.spec.ave_path;
  const int hithr = 3 * A.kmer;                           // HITMIN*Kmer, map.c:2419
  LPebble *const arena = reinterpret_cast<LPebble *>(A.lane_cells);

  // job
  int phase = PH_IDLE, jid = -1, ar = 0, br = 0, cm = 0, alen = 0, blen = 0;
  const uint8_t *aseq = nullptr, *bseq = nullptr;
  long long chain = 0, abase = 0;
  int clen = 0, sn = 0, apos = 0, bpos = 0, alast = 0, first = -1, last = -1, count = 0, status = 0;
  int acap = 0, atop = 0, astart = 0;
  // alignment
  int anti = 0, dg = 0, aoff = 0, call = 0, ncall = 0, fshort = 0;
  int p_ab = 0, p_bb = 0, p_ae = 0, p_be = 0, p_df = 0;
  LaneCall c0, c1;
  c0.cells = c1.cells = 0; c0.dir = c1.dir = 0; c0.mida = c1.mida = 0; c0.aoff = c1.aoff = 0;
  c0.ha = c1.ha = 0; c0.hb = c1.hb = 0; c0.x = c1.x = 0; c0.y = c1.y = 0; c0.d = c1.d = 0;
  // wave (primed coordinates: everything multiplied by dir)
  int dir = 1, off = 0, k0 = 0, mida = 0, low = 0, hgh = 0, dif = 0, avail = 0, more = 1;
  int aclip = IMAX, bclip = -IMAX, cbase = 0;
  int besta = 0, besty = 0, trima = 0, trimy = 0, trimd = 0, trimha = 0, trimhb = 1;
  int morea = 0, morey = 0, mored = 0, moreha = 0, morehb = 1, morem = -1, lasta = 0;
  unsigned long long nwaves = 0, ncells = 0, nalign = 0, nempty = 0;
  int pha = 0, phb = 0, aulo = 0, auhi = 0, bulo = 0, buhi = 0;      // window phases and extents (u units)

  // 16 virtual bases starting at u == uu (multiple of 16) of sequence `seq`, as four window words
#define CHUNK_LOAD(seq, ph, uu, q)                                                          \
  { const int v_ = (uu) + (ph);                                                             \
    const uint8_t *g_ = (dir > 0) ? (seq) + v_ : (seq) - 16 - v_;                           \
    q = *reinterpret_cast<const uint4 *>(g_);                                               \
    if (dir < 0)                                                                            \
      { const uint32_t x_ = __byte_perm(q.w, 0, 0x0123), y_ = __byte_perm(q.z, 0, 0x0123),  \
                       z_ = __byte_perm(q.y, 0, 0x0123), w_ = __byte_perm(q.x, 0, 0x0123);  \
        q = make_uint4(x_, y_, z_, w_);                                                     \
      }                                                                                     \
  }
#define CHUNK_STORE(win, uu, q)                                                             \
  { const int j_ = (uu) >> 2;                                                               \
    WINW(win, j_) = q.x; WINW(win, j_ + 1) = q.y; WINW(win, j_ + 2) = q.z; WINW(win, j_ + 3) = q.w; }
  // eight virtual bases starting at virtual position v of a sequence: window, else global memory
#define BASES8(seq, win, ph, ulo, uhi, v, out)                                              \
  { const int u_ = (v) - (ph), j_ = u_ >> 2;                                                \
    if (u_ >= (ulo) && (j_ + 3) * 4 <= (uhi))                                               \
      { const uint32_t w0_ = WINW(win, j_), w1_ = WINW(win, j_ + 1), w2_ = WINW(win, j_ + 2); \
        const unsigned s_ = (unsigned) (u_ & 3) * 8;                                        \
        out = ((uint64_t) __funnelshift_r(w1_, w2_, s_) << 32) | __funnelshift_r(w0_, w1_, s_); \
      }                                                                                     \
    else                                                                                    \
      out = bases8(seq, dir * (v) + off, dir);                                              \
  }

#define NEWCELL(dst, pp, kk, dd, mm)                                                        \
  { if (avail >= LCELLS || atop + avail >= acap) { status = LERR_CELLS; }                   \
    else { arena[abase + cbase + avail] = LPebble{ (pp), dir * (kk), (dd), dir * (mm) };    \
           (dst) = avail++; } }
#define BP(yp) bseq[dir * (yp) + off]
#define AP(xp) aseq[dir * (xp) + off]

  // slide along primed diagonal kp from primed y while bases match; hit: 1 = end of B, 2 = end of A
#define SLIDE(kp, y, hit)                                                                   \
  { hit = 0;                                                                                \
    while (true)                                                                            \
      { uint64_t wa, wb;                                                                    \
        BASES8(aseq, winA, pha, aulo, auhi, (y) + (kp), wa);                                \
        BASES8(bseq, winB, phb, bulo, buhi, (y), wb);                                       \
        const uint64_t x_ = wa ^ wb, e_ = wb & 0x0404040404040404ull;                       \
        if ((x_ | e_) == 0) { (y) += 8; continue; }                                         \
        const int ix = x_ ? (__ffsll((long long) x_) - 1) >> 3 : 8;                         \
        const int ie = e_ ? (__ffsll((long long) e_) - 1) >> 3 : 8;                         \
        if (ie <= ix) { hit = 1; (y) += ie; }                                               \
        else { if (((wa >> (8 * ix)) & 0xff) == 4) hit = 2; (y) += ix; }                    \
        break;                                                                              \
      }                                                                                     \
  }

  // boundary handling after a wave (align.c:558-583,848-875 / 1216-1241,1502-1529)
#define CLIP_AFTER_WAVE                                                                     \
  if (more == 0)                                                                            \
    { if (BP(besty) != 4 && AP(besta - besty) != 4)                                         \
        more = 1;                                                                           \
      if (hgh >= aclip)                                                                     \
        { hgh = aclip - 1;                                                                  \
          const int mh_ = SMH(aclip), m_ = (int) ((unsigned) mh_ >> 24);                    \
          if (morem <= m_)                                                                  \
            { morem = m_; morea = SV(aclip); morey = (morea - aclip) / 2; mored = dif;      \
              moreha = mh_ & 0xfff; morehb = (mh_ >> 12) & 0xfff;                           \
            }                                                                               \
        }                                                                                   \
      if (low <= bclip)                                                                     \
        { low = bclip + 1;                                                                  \
          const int mh_ = SMH(bclip), m_ = (int) ((unsigned) mh_ >> 24);                    \
          if (morem <= m_)                                                                  \
            { morem = m_; morea = SV(bclip); morey = (morea - bclip) / 2; mored = dif;      \
              moreha = mh_ & 0xfff; morehb = (mh_ >> 12) & 0xfff;                           \
            }                                                                               \
        }                                                                                   \
      aclip = IMAX; bclip = -IMAX;                                                          \
    }

  while (true)
    { // ---- scalar state machine: until this lane sits in a wave or has run out of jobs
      while (phase != PH_WAVE && phase != PH_DONE)
        { if (phase == PH_IDLE)
            { const int j = atomicAdd(A.job_counter, 1);
              if (j >= A.njobs) { phase = PH_DONE; break; }
              jid = A.job_list ? A.job_list[j] : j;
              const AlignJob job = A.jobs[jid];
              const Candidate cd = A.cand[job.cand];
              ar = job.read; br = cd.bread; cm = cd.comp;
              alen = A.rlen_a[ar]; blen = A.rlen_b[br];
              bseq = A.bases_b + A.boff_b[br];
              aseq = (cm ? A.bases_ac : A.bases_a) + A.boff_a[ar];
              chain = cd.chain; clen = cd.length; sn = 0;
              apos = cd.alast; bpos = cd.blast; alast = alen + 1;
              first = last = -1; count = 0; status = 0;
              acap = LANE_ARENA(alen / TS);
              abase = A.lane_cell_base[ar] + (long long) (jid - (int) A.lane_job_off[ar]) * acap;
              atop = 0;
              phase = PH_SEED;
            }
          else if (phase == PH_SEED)                       // map.c:2487-2513
            { if (status != 0 || sn >= clen) { phase = PH_JOBEND; continue; }
              const uint32_t jp = A.jumps[chain + sn];
              sn += 1;
              apos -= (int) (jp & 0xffff);
              bpos -= (int) (jp >> 16);
              if (apos >= alast)
                continue;
              if (cm) { const int ac = alen - apos, bc = blen - bpos; dg = ac - bc; anti = ac + bc; }
              else    { dg = apos - bpos; anti = apos + bpos; }
              if (((anti - dg) >> 1) < 0) { status = LERR_MULTI; phase = PH_JOBEND; continue; }
              aoff = cm ? alen % TS : 0;                   // align.c:1794-1797
              nalign += 1;
              p_ab = p_bb = p_ae = p_be = p_df = 0;
              astart = atop;
              call = 0; ncall = 0; fshort = 0;
              dir = 1; k0 = dg; mida = anti;               // forward wave from the seed
              phase = PH_START;
            }
          else if (phase == PH_START)                      // wave 0 (align.c:433-556 / 1093-1214)
            { // here k0 and mida are ACTUAL coordinates; they are primed below
              off = (dir < 0) ? -1 : 0;
              cbase = atop; avail = 0; dif = 0; more = 1;
              aclip = IMAX; bclip = -IMAX;
              const int k = k0;
              int y = (mida - k) >> 1, na, nb, ha = 0, hb = 1;
              if (dir > 0)
                { na = (((y + k) + (TS - aoff)) / TS - 1) * TS + aoff;
                  nb = ((y + TS) / TS - 1) * TS;
                  NEWCELL(ha, -1, k, 0, na);               // dir == 1: stored as they are
                  NEWCELL(hb, -1, k, 0, nb);
                  na += TS; nb += TS;
                }
              else
                { na = (((y + k) + (TS - aoff) - 1) / TS - 1) * TS + aoff;
                  nb = ((y + TS - 1) / TS - 1) * TS;
                  if (atop + 2 >= acap) status = LERR_CELLS;
                  else
                    { arena[abase + cbase]     = LPebble{ -1, k, 0, y + k };
                      arena[abase + cbase + 1] = LPebble{ -1, k, 0, y };
                      avail = 2;
                    }
                }
              if (status != 0) { phase = PH_JOBEND; continue; }
              // primed from here on
              const int kp = dir * k;
              int yp = dir * y, nap = dir * na, nbp = dir * nb, hit;
              k0 = kp; mida = dir * mida;
              low = hgh = kp;
              besta = trima = morea = lasta = mida;
              besty = trimy = morey = yp;
              trimd = mored = 0; trimha = moreha = 0; trimhb = morehb = 1; morem = -1;
              // windows: chunk phases, then the first LANE_WIN*4 bases from the start position on
              pha = (int) ((dir > 0 ? (0 - reinterpret_cast<uintptr_t>(aseq)) : reinterpret_cast<uintptr_t>(aseq)) & 15);
              phb = (int) ((dir > 0 ? (0 - reinterpret_cast<uintptr_t>(bseq)) : reinterpret_cast<uintptr_t>(bseq)) & 15);
              aulo = auhi = ((yp + kp) - pha - 16) & ~15;
              bulo = buhi = (yp - phb - 16) & ~15;
              for (int i_ = 0; i_ < LANE_WIN / 4; i_++)
                { uint4 qa, qb;
                  CHUNK_LOAD(aseq, pha, auhi, qa);
                  CHUNK_LOAD(bseq, phb, buhi, qb);
                  CHUNK_STORE(winA, auhi, qa);
                  CHUNK_STORE(winB, buhi, qb);
                  auhi += 16; buhi += 16;
                }
              SLIDE(kp, yp, hit);
              if (hit)
                { more = 0;
                  if (hit == 1) bclip = kp; else aclip = kp;
                }
              const int c = (yp << 1) + kp;
              while (yp + kp >= nap && status == 0)
                { NEWCELL(ha, ha, kp, 0, nap); nap += TS; }
              if (DOB)
                while (yp >= nbp && status == 0)
                  { NEWCELL(hb, hb, kp, 0, nbp); nbp += TS; }
              if (status != 0) { phase = PH_JOBEND; continue; }
              if (besta < c)
                { besta = trima = lasta = c;
                  besty = trimy = yp;
                  trimha = ha; trimhb = hb;
                }
              SV(kp) = c; STL(kp) = (int) (uint32_t) PATH_INT; STH(kp) = (int) (uint32_t) (PATH_INT >> 32);
              SMH(kp) = (PATH_LEN << 24) | (hb << 12) | ha;
              SNA(kp) = nap;
              if (DOB) SNB(kp) = nbp;
              CLIP_AFTER_WAVE
              phase = PH_WAVE;
            }
          else if (phase == PH_ENDCALL)
            { int tx, ty, td, tha, thb;
              if (morem >= 0) { tx = morea - morey; ty = morey; td = mored; tha = moreha; thb = morehb; }
              else            { tx = trima - trimy; ty = trimy; td = trimd; tha = trimha; thb = trimhb; }
              LaneCall cc;
              cc.cells = abase + cbase; cc.dir = dir; cc.mida = dir * mida; cc.aoff = aoff;
              cc.ha = tha; cc.hb = thb; cc.x = dir * tx; cc.y = dir * ty; cc.d = td;
              atop += avail;
              if (dir > 0) { p_ae = cc.x; p_be = cc.y; p_df = td; }
              else         { p_ab = cc.x; p_bb = cc.y; p_df += td; }
              if (call == 0)                               // forward done: reverse from the seed
                { c0 = cc; ncall = 1;
                  fshort = ((p_ae + p_be) - anti < DUB_TRIM);
                  call = 1; dir = -1; k0 = dg; mida = anti;
                  phase = PH_START;
                }
              else if (call == 1)                          // align.c:1810-1854
                { c1 = cc; ncall = 2;
                  const int rshort = (anti - (p_ab + p_bb) < DUB_TRIM);
                  if (fshort && rshort)
                    { p_ae = p_ab = (p_ab + p_ae) / 2;
                      p_be = p_bb = (p_bb + p_be) / 2;
                      ncall = 0;
                      phase = PH_FINISH;
                    }
                  else if (fshort)
                    { call = 2; dir = 1; k0 = p_ab - p_bb; mida = p_ab + p_bb;
                      phase = PH_START;
                    }
                  else if (rshort)
                    { call = 2; dir = -1; k0 = p_ae - p_be; mida = p_ae + p_be; p_df = 0;
                      phase = PH_START;
                    }
                  else
                    phase = PH_FINISH;
                }
              else                                         // the re-run replaces both traces
                { c0 = cc; ncall = 1;
                  phase = PH_FINISH;
                }
            }
          else if (phase == PH_FINISH)                     // align.c:1857-1912, map.c:2514-2579
            { int a_ab = p_ab, a_bb = p_bb, a_ae = p_ae, a_be = p_be;
              const int b_ab = p_bb, b_bb = p_ab, b_ae = p_be, b_be = p_ae;
              if (cm)
                { a_ab = alen - b_be; a_bb = blen - b_ae; a_ae = alen - b_bb; a_be = blen - b_ab; }
              if (a_ae - a_ab < hithr)
                { atop = astart;                           // dropped: its cells are released
                  phase = PH_SEED;
                  continue;
                }
              alast = a_ab;
              const int rec = atomicAdd(A.aln_top, 1);
              if (rec >= A.aln_cap) { status = LERR_POOL; phase = PH_JOBEND; continue; }
              AlnRec r;
              r.next = -1; r.comp = cm; r.bread = br; r.pad = 0;
              r.a[0] = a_ab; r.a[1] = a_bb; r.a[2] = a_ae; r.a[3] = a_be; r.a[4] = p_df; r.a[5] = 0;
              r.b[0] = b_ab; r.b[1] = b_bb; r.b[2] = b_ae; r.b[3] = b_be; r.b[4] = p_df; r.b[5] = 0;
              r.atrace = 0; r.btrace = 0;
              A.alns[rec] = r;
              LaneUnwind u;
              u.ncalls = ncall; u.acomp = cm; u.job = jid; u.pad = 0;
              u.call[0] = c0; u.call[1] = c1;
              A.unwind[rec] = u;
              if (last >= 0) A.alns[last].next = rec;
              if (first < 0) first = rec;
              last = rec;
              count += 1;
              phase = PH_SEED;
            }
          else if (phase == PH_JOBEND)
            { AlignJob &job = A.jobs[jid];
              job.first = (status == 0) ? first : -1;
              job.count = (status == 0) ? count : 0;
              job.status = status;
              if (status != 0)
                { atomicAdd(A.nfailed, 1);
                  atomicAdd(&A.stats[4], 1ull << (16 * (status > 3 ? 3 : status - 1)));   // why (trace aid)
                }
              phase = PH_IDLE;
            }
        }
      if (__all_sync(0xffffffffu, phase == PH_DONE))
        break;

      // ---- one wave for every lane that is inside a call (align.c:592-898 / 1248-1552)
      if (phase == PH_WAVE)
        { bool go = more && (lasta >= besta - TRIM_MLAG);
          if (go && hgh < low)                            // empty band: the reference would read
            { nempty += 1; go = false; }                  // stale cells; stop (as align.cu does)
          if (go && hgh - low + 6 > W)
            { status = LERR_BAND; go = false; }
          if (!go)
            phase = (status != 0) ? PH_JOBEND : PH_ENDCALL;
          else
            { low -= 1; hgh += 1; dif += 1;
              SNA(low) = SNA(low + 1); SNA(hgh) = SNA(hgh - 1);
              if (DOB) { SNB(low) = SNB(low + 1); SNB(hgh) = SNB(hgh - 1); }

              // window prefetch: one chunk per sequence when the front gets close to the end
              const bool fa = (((besta + hgh) >> 1) + 40 - pha > auhi);
              const bool fb = (((besta - low) >> 1) + 40 - phb > buhi);
              uint4 qa = make_uint4(0, 0, 0, 0), qb = make_uint4(0, 0, 0, 0);
              if (fa) CHUNK_LOAD(aseq, pha, auhi, qa);
              if (fb) CHUNK_LOAD(bseq, phb, buhi, qb);

              // rolling old states: p = diagonal k+1, c = diagonal k, m = diagonal k-1
              int nbk = 0, bk0 = 0;                       // good breakers of this wave: [bk0, nbk)
              int ap = LSENT, ac = LSENT, am;
              int pmh = 0, cmh = 0, mmh = 0;
              uint32_t ptl = 0, pth = 0, ctl = 0, cth = 0, mtl = 0, mth = 0;
              for (int k = hgh; k >= low; k--)
                { if (k - 1 > low)
                    { am = SV(k - 1); mtl = (uint32_t) STL(k - 1); mth = (uint32_t) STH(k - 1); mmh = SMH(k - 1); }
                  else
                    am = LSENT;
                  int c, mh; uint32_t tl, th;
                  if (ac < am)                            // align.c:712-741 / 1367-1396
                    { if (am < ap) { c = ap + 1; mh = pmh; tl = ptl; th = pth; }
                      else         { c = am + 1; mh = mmh; tl = mtl; th = mth; }
                    }
                  else
                    { if (ac < ap) { c = ap + 1; mh = pmh; tl = ptl; th = pth; }
                      else         { c = ac + 2; mh = cmh; tl = ctl; th = cth; }
                    }
                  uint64_t b = ((uint64_t) th << 32) | tl;
                  int m = (int) ((unsigned) mh >> 24), ha = mh & 0xfff, hb = (mh >> 12) & 0xfff;
                  if ((b & PATH_TOP) != 0) m -= 1;
                  b <<= 1;

                  const int y0 = (c - k) >> 1;
                  int y = y0, hit;
                  SLIDE(k, y, hit);
                  const int r = y - y0;
                  if (r > 0)                              // closed form of align.c:764-767
                    { const int rr = (r < 61) ? r : 61;
                      const uint64_t mask = ((1ull << rr) - 1) << (61 - rr);
                      m += rr - __popcll(b & mask);
                      b = (r >= 64) ? ~0ull : ((b << r) | ((1ull << r) - 1));
                    }
                  c = (y << 1) + k;
                  if (hit)
                    { more = 0;
                      if (hit == 1) { if (bclip < k) bclip = k; }
                      else          aclip = k;
                    }

                  int na = SNA(k);                        // align.c:771-793 / 1426-1448
                  if (y + k >= na)
                    { do
                        { if (dir * arena[abase + cbase + ha].mark < na)
                            NEWCELL(ha, ha, k, dif, na);
                          na += TS;
                        }
                      while (y + k >= na);
                      SNA(k) = na;
                    }
                  if (DOB)                                // align.c:795-817 / 1449-1471
                    { int nb = SNB(k);
                      if (y >= nb)
                        { do
                            { if (dir * arena[abase + cbase + hb].mark < nb)
                                NEWCELL(hb, hb, k, dif, nb);
                              nb += TS;
                            }
                          while (y >= nb);
                          SNB(k) = nb;
                        }
                    }

                  if (c > besta)                          // align.c:819-833 / 1473-1487
                    { besta = c; besty = y;
                      if (m >= PATH_AVE)
                        { lasta = c;
                          if (nbk - bk0 == 4)             // buffer full: settle the oldest now
                            { bool ok;
                              TRIM_TEST(BRK(bk0, 3), BRK(bk0, 4), ok);
                              if (ok)
                                { trima = BRK(bk0, 0); trimy = BRK(bk0, 1); trimd = dif;
                                  trimha = BRK(bk0, 2) & 0xfff; trimhb = (BRK(bk0, 2) >> 12) & 0xfff;
                                }
                              bk0 += 1;
                            }
                          BRK(nbk, 0) = c; BRK(nbk, 1) = y; BRK(nbk, 2) = (hb << 12) | ha;
                          BRK(nbk, 3) = (int) (uint32_t) b; BRK(nbk, 4) = (int) (uint32_t) (b >> 32);
                          nbk += 1;
                        }
                    }

                  // rotate the old states, then overwrite diagonal k
                  ap = ac; pmh = cmh; ptl = ctl; pth = cth;
                  ac = am; cmh = mmh; ctl = mtl; cth = mth;
                  SV(k) = c; STL(k) = (int) (uint32_t) b; STH(k) = (int) (uint32_t) (b >> 32);
                  SMH(k) = (m << 24) | (hb << 12) | ha;
                }

              if (fa)
                { CHUNK_STORE(winA, auhi, qa);
                  auhi += 16;
                  if (auhi - aulo > LANE_WIN * 4) aulo = auhi - LANE_WIN * 4;
                }
              if (fb)
                { CHUNK_STORE(winB, buhi, qb);
                  buhi += 16;
                  if (buhi - bulo > LANE_WIN * 4) bulo = buhi - LANE_WIN * 4;
                }
              // the trim point is the LAST good breaker that passes the table test (align.c:824-832)
              for (int e = nbk - 1; e >= bk0; e--)
                { bool ok;
                  TRIM_TEST(BRK(e, 3), BRK(e, 4), ok);
                  if (ok)
                    { trima = BRK(e, 0); trimy = BRK(e, 1); trimd = dif;
                      trimha = BRK(e, 2) & 0xfff; trimhb = (BRK(e, 2) >> 12) & 0xfff;
                      break;
                    }
                }
              if (status != 0)
                phase = PH_JOBEND;
              else
                { CLIP_AFTER_WAVE
                  // trim the band to within WAVE_LAG of the best point (align.c:877-885 / 1531-1539)
                  const int n = besta - WAVE_LAG;
                  while (hgh >= low)
                    if (SV(hgh) < n)
                      hgh -= 1;
                    else
                      { while (SV(low) < n)
                          low += 1;
                        break;
                      }
                  nwaves += 1;
                  ncells += (hgh - low) + 1;
                }
            }
        }
    }
  atomicAdd(&A.stats[0], nalign); atomicAdd(&A.stats[1], nwaves);
  atomicAdd(&A.stats[2], ncells); atomicAdd(&A.stats[3], nempty);
#undef FLD
#undef SV
#undef STL
#undef STH
#undef SMH
#undef SNA
#undef SNB
#undef NEWCELL
#undef BP
#undef AP
#undef SLIDE
#undef CLIP_AFTER_WAVE
#undef WINW
#undef BRK
#undef TRIM_TEST
#undef CHUNK_LOAD
#undef CHUNK_STORE
#undef BASES8
}

// ---- unwinding: Pebble chains -> trace pairs (align.c:900-1007 / 1554-1717), one thread per
// alignment kept by the lane kernel ------------------------------------------------------------
struct UPath { int tlen; uint16_t *trace; };

// one wave call; the A chain always, the B chain with dob.  Returns 0 or LERR_TRACE.
__device__ int unwind_call(const LaneCall &cc, LPebble *cells, int TS, int dob, UPath &apath, UPath &bpath,
                           uint16_t *alo, uint16_t *ahi, uint16_t *blo, uint16_t *bhi)
{ uint16_t *atrace = apath.trace, *btrace = bpath.trace;
  const int DIR = cc.dir, mida = cc.mida, aoff = cc.aoff, boff = 0;
  const int trimx = cc.x, trimy = cc.y, trimd = cc.d;
  int atlen = 0, btlen = 0, a, bq, k, h, d, e, err = 0;

  a = -1;                                               // A chain
  for (h = cc.ha; h >= 0; h = bq)
    { bq = cells[h].ptr; cells[h].ptr = a; a = h; }
  h = a;
  k = cells[h].diag;
  if (DIR > 0)
    { bq = (mida - k) / 2;
      e = 0;
      for (h = cells[h].ptr; h >= 0; h = cells[h].ptr)
        { k = cells[h].diag; a = cells[h].mark - k; d = cells[h].diff;
          if (atrace + atlen + 2 > ahi) { err = LERR_TRACE; break; }
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
              if (atrace + atlen - 4 < alo) { err = LERR_TRACE; break; }
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

  if (dob && !err)                                      // B chain
    { a = -1;
      for (h = cc.hb; h >= 0; h = bq)
        { bq = cells[h].ptr; cells[h].ptr = a; a = h; }
      h = a;
      k = cells[h].diag;
      if (DIR > 0)
        { bq = (mida + k) / 2;
          e = 0;
          for (h = cells[h].ptr; h >= 0; h = cells[h].ptr)
            { k = cells[h].diag; a = cells[h].mark + k; d = cells[h].diff;
              if (btrace + btlen + 2 > bhi) { err = LERR_TRACE; break; }
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
            { for (h = cells[h].ptr; h >= 0; h = cells[h].ptr)
                { k = cells[h].diag; a = cells[h].mark + k;
                  if (btrace + btlen - 4 < blo) { err = LERR_TRACE; break; }
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
  if (err) return err;
  if (DIR > 0)
    { apath.tlen = atlen; bpath.tlen = btlen; }
  else
    { apath.tlen = apath.tlen - atlen; apath.trace = apath.trace + atlen;
      bpath.tlen = bpath.tlen - btlen; bpath.trace = bpath.trace + btlen;
    }
  return 0;
}

__global__ void __launch_bounds__(128)
k_unwind(AlignArgs A)
{ const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int naln = *A.aln_top;
  if (naln > A.aln_cap) naln = A.aln_cap;
  if (i >= naln) return;
  const LaneUnwind u = A.unwind[i];
  if (u.ncalls < 0) return;                             // not a lane-kernel record
  if (A.jobs[u.job].status != 0) return;                // the job failed later: it is re-run whole
  uint16_t *tb = A.lane_tscratch + (size_t) i * 4 * A.tcap;
  uint16_t *const alo = tb, *const ahi = tb + 2 * A.tcap, *const blo = ahi, *const bhi = tb + 4 * A.tcap;
  UPath ap, bp;
  ap.trace = tb + A.tcap; bp.trace = tb + 3 * A.tcap; ap.tlen = bp.tlen = 0;
  int err = 0;
  LPebble *arena = reinterpret_cast<LPebble *>(A.lane_cells);
  for (int c = 0; c < u.ncalls && !err; c++)
    err = unwind_call(u.call[c], arena + u.call[c].cells, A.spec.spacing, A.do_b, ap, bp, alo, ahi, blo, bhi);
  if (err)
    { if (atomicExch(&A.jobs[u.job].status, err) == 0)
        atomicAdd(A.nfailed, 1);
      return;
    }
  if (u.acomp)                                          // align.c:1858-1884
    { uint16_t *trace = ap.trace, p;
      int ii = ap.tlen - 2, j = 0;
      while (j < ii)
        { p = trace[ii]; trace[ii] = trace[j]; trace[j] = p;
          p = trace[ii + 1]; trace[ii + 1] = trace[j + 1]; trace[j + 1] = p;
          ii -= 2; j += 2;
        }
    }
  const int tl = ap.tlen + (A.do_b ? bp.tlen : 0);
  const long long to = (long long) atomicAdd(A.trace_top, (unsigned long long) tl);
  if (to + tl > A.trace_cap)
    { if (atomicExch(&A.jobs[u.job].status, LERR_POOL) == 0)
        atomicAdd(A.nfailed, 1);
      return;
    }
  for (int t = 0; t < ap.tlen; t++)
    A.traces[to + t] = ap.trace[t];
  if (A.do_b)
    for (int t = 0; t < bp.tlen; t++)
      A.traces[to + ap.tlen + t] = bp.trace[t];
  AlnRec &r = A.alns[i];
  r.a[5] = ap.tlen; r.b[5] = bp.tlen;
  r.atrace = to; r.btrace = to + ap.tlen;
}

int lane_warps(bool dob) { return dob ? LANE_WARPS - 1 : LANE_WARPS; }   // leave the L1 some room

size_t lane_smem_bytes(bool dob)
{ return (size_t) lane_warps(dob) * (dob ? LANE_WARP_WORDS(true) : LANE_WARP_WORDS(false)) * sizeof(uint32_t); }

void launch_align_lane(const AlignArgs &A, int nblocks, cudaStream_t stream)
{ const bool dob = (A.do_b != 0);
  const size_t smem = lane_smem_bytes(dob);
  static bool attr_set = false;
  if (!attr_set)
    { CUDA_CHECK(cudaFuncSetAttribute(k_align_lane<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int) lane_smem_bytes(true)));
      CUDA_CHECK(cudaFuncSetAttribute(k_align_lane<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int) lane_smem_bytes(false)));
      attr_set = true;
    }
  if (dob) LAUNCH(k_align_lane<true>, nblocks, lane_warps(true) * 32, smem, stream, A);
  else     LAUNCH(k_align_lane<false>, nblocks, lane_warps(false) * 32, smem, stream, A);
}

void launch_unwind(const AlignArgs &A, int max_alns, cudaStream_t stream)
{ if (max_alns <= 0) return;
  LAUNCH(k_unwind, (max_alns + 127) / 128, 128, 0, stream, A);
}

}  // namespace damgpu
