// Wave (O(ND)) local alignment, G CANDIDATE CHAINS PER WARP, each owned by a fixed group of 32/G lanes
// ("group kernel") -- forward_wave align.c:353-1011, reverse_wave :1015-1720, Local_Alignment
// :1727-1946, the seeding loop of report_thread map.c:2487-2579.
//
// After WAVE_LAG trimming a wave spans ~8 diagonals, so a warp per alignment (align.cu) keeps a
// quarter of its lanes busy.  Here a group of 8 (or 16) lanes owns a job: lane i of the group takes
// the i-th diagonal of the band in scan order, 8 at a time when the band is wider, and the four
// groups of a warp advance their waves in lock step.
//  * Group-uniform state (best point, trim point, clips, job bookkeeping) is held redundantly by
//    every lane of the group, like the warp kernel does for the warp: no owner lane, nothing to
//    publish.  Side effects (job fetch, Pebble cells of wave 0, records) are done by lane 0 of the group.
//  * Per-diagonal state (V, T, M|HB|HA double-buffered; NA, NB) lives in shared memory, a window of
//    PACK_W diagonals per group indexed by k & (PACK_W-1).
//  * The order-dependent reductions (record breakers, aclip = last writer, bclip = extreme k) use
//    ballots restricted to the group's lanes and a prefix maximum over width-limited shuffles.
//  * Direction is data: a reverse wave is a forward wave on negated coordinates over the sequences
//    read backwards, so groups extending in opposite directions share every instruction.
//  * Pebble cells go to a per-job arena in global memory; chains are unwound afterwards by k_unwind
//    (align_lane.cu), one thread per kept alignment.
// A job that outgrows the window or its arena is marked failed and re-run by the warp kernels of
// align.cu (host loop in report.cu).
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
enum { PH_IDLE = 0, PH_SEED, PH_SEEDGO, PH_START, PH_WAVE, PH_ENDCALL, PH_FINISH, PH_JOBEND, PH_DONE };
enum { H_LOW = 0, H_HGH, H_DIR, H_OFF, H_BESTA, H_DIF, H_CUR, H_AVAIL, H_STATUS, H_CAP,
       H_ASEQ, H_ASEQ2, H_BSEQ, H_BSEQ2, H_CELLS, H_CELLS2, H_WORDS };

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

// slide along primed diagonal kp from primed y while bases match; hit: 1 = end of B, 2 = end of A
__device__ __forceinline__ int slide8(const uint8_t *aseq, const uint8_t *bseq, int dir, int off,
                                      int kp, int y, int &hit)
{ hit = 0;
  while (true)
    { const uint64_t wa = bases8(aseq, dir * (y + kp) + off, dir);
      const uint64_t wb = bases8(bseq, dir * y + off, dir);
      const uint64_t x = wa ^ wb, e = wb & 0x0404040404040404ull;
      if ((x | e) == 0) { y += 8; continue; }
      const int ix = x ? (__ffsll((long long) x) - 1) >> 3 : 8;
      const int ie = e ? (__ffsll((long long) e) - 1) >> 3 : 8;
      if (ie <= ix) { hit = 1; return y + ie; }
      if (((wa >> (8 * ix)) & 0xff) == 4) hit = 2;
      return y + ix;
    }
}

}  // namespace

#ifndef GROUP_MINB
#define GROUP_MINB 1
#endif
template <int G, bool DOB>
__global__ void __launch_bounds__(PACK_WARPS * 32, GROUP_MINB)
k_align_group(AlignArgs A)
{ extern __shared__ int psm[];
  constexpr int W = PACK_W, SLOTW = PACK_SLOT_WORDS(DOB);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int *const hdr = psm + (size_t) wib * PACK_WARP_WORDS(G, DOB);
  int *const stt = hdr + G * H_WORDS;
  constexpr int SEGL = 32 / G;                          // lanes per group
  const int g = lane / SEGL, li = lane % SEGL, gl0 = g * SEGL;
  const unsigned gmask = (SEGL == 32) ? 0xffffffffu : (((1u << SEGL) - 1) << gl0);
  const bool lead = (li == 0);
#define HD(g, f)        hdr[(g) * H_WORDS + (f)]
#define ST(g, b, f, k)  stt[(g) * SLOTW + (((b) * 4 + (f)) * W) + ((k) & (W - 1))]
#define NAW(g, k)       stt[(g) * SLOTW + 8 * W + ((k) & (W - 1))]
#define NBW(g, k)       stt[(g) * SLOTW + 9 * W + ((k) & (W - 1))]
  // the group's own slot, current buffer
#define SV(k)  ST(g, cur, 0, k)
#define STL(k) ST(g, cur, 1, k)
#define STH(k) ST(g, cur, 2, k)
#define SMH(k) ST(g, cur, 3, k)             /* M << 24 | HB << 12 | HA */
#define SNA(k) NAW(g, k)
#define SNB(k) NBW(g, k)

  const int TS = A.spec.spacing, PATH_AVE = A.spec.ave_path;
  const int hithr = 3 * A.kmer;                           // HITMIN*Kmer, map.c:2419
  LPebble *const arena = reinterpret_cast<LPebble *>(A.lane_cells);

  // ---- group-uniform state (every lane of the group holds the same values) -------------------
  int phase = PH_IDLE, jid = -1, ar = 0, br = 0, cm = 0, alen = 0, blen = 0;
  const uint8_t *aseq = nullptr, *bseq = nullptr;
  long long chain = 0, abase = 0;
  int clen = 0, sn = 0, apos = 0, bpos = 0, alast = 0, first = -1, last = -1, count = 0, status = 0;
  int acap = 0, atop = 0, astart = 0, cellcap = 0;
  int anti = 0, dg = 0, aoff = 0, call = 0, ncall = 0, fshort = 0;
  int p_ab = 0, p_bb = 0, p_ae = 0, p_be = 0, p_df = 0;
  LaneCall c0, c1;
  c0.cells = c1.cells = 0; c0.dir = c1.dir = 0; c0.mida = c1.mida = 0; c0.aoff = c1.aoff = 0;
  c0.ha = c1.ha = 0; c0.hb = c1.hb = 0; c0.x = c1.x = 0; c0.y = c1.y = 0; c0.d = c1.d = 0;
  c0.pad = c1.pad = 0;
  // wave (primed coordinates: everything multiplied by dir)
  int dir = 1, off = 0, k0 = 0, mida = 0, low = 0, hgh = 0, dif = 0, more = 1, cur = 0;
  int aclip = IMAX, bclip = -IMAX, cbase = 0;
  int besta = 0, besty = 0, trima = 0, trimy = 0, trimd = 0, trimha = 0, trimhb = 1;
  int morea = 0, morey = 0, mored = 0, moreha = 0, morehb = 1, morem = -1, lasta = 0;
  unsigned long long nwaves = 0, ncells = 0, nalign = 0, nempty = 0, jobw0 = 0, joba0 = 0;

  // cell allocation of wave 0: group-uniform counter, lane 0 of the group writes
  int avail0 = 0;
#define ONEWCELL(dst, pp, kk, dd, mm)                                                       \
  { if (avail0 >= LCELLS || atop + avail0 >= acap) { status = LERR_CELLS; }                 \
    else { if (lead) arena[abase + cbase + avail0] = LPebble{ (pp), dir * (kk), (dd), dir * (mm) }; \
           (dst) = avail0; avail0 += 1; } }
#define BP(yp) bseq[dir * (yp) + off]
#define AP(xp) aseq[dir * (xp) + off]
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
    { // ---- owner lanes: scalar state machine until the slot sits in a wave or has no more jobs
      while (phase != PH_WAVE && phase != PH_DONE && phase != PH_SEED)
        { if (phase == PH_IDLE)
            { int j = 0;
              if (lead) j = atomicAdd(A.job_counter, 1);
              j = __shfl_sync(gmask, j, gl0);
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
              jobw0 = nwaves; joba0 = nalign;
              acap = LANE_ARENA(alen / TS);
              abase = A.lane_cell_base[ar] + (long long) (jid - (int) A.lane_job_off[ar]) * acap;
              atop = 0;
              phase = PH_SEED;
            }
          else if (phase == PH_SEEDGO)                     // map.c:2487-2513, seed chosen below
            {               if (cm) { const int ac = alen - apos, bc = blen - bpos; dg = ac - bc; anti = ac + bc; }
              else    { dg = apos - bpos; anti = apos + bpos; }
              if (((anti - dg) >> 1) < 0) { status = LERR_MULTI; phase = PH_JOBEND; continue; }
              aoff = cm ? alen % TS : 0;                   // align.c:1794-1797
              if (lead) nalign += 1;
              p_ab = p_bb = p_ae = p_be = p_df = 0;
              astart = atop;
              call = 0; ncall = 0; fshort = 0;
              dir = 1; k0 = dg; mida = anti;               // forward wave from the seed
              phase = PH_START;
            }
          else if (phase == PH_START)                      // wave 0 (align.c:433-556 / 1093-1214)
            { // here k0 and mida are ACTUAL coordinates; they are primed below
              off = (dir < 0) ? -1 : 0;
              cbase = atop; dif = 0; more = 1; cur = 0;
              avail0 = 0;
              aclip = IMAX; bclip = -IMAX;
              const int k = k0;
              int y = (mida - k) >> 1, na, nb, ha = 0, hb = 1;
              if (dir > 0)
                { na = (((y + k) + (TS - aoff)) / TS - 1) * TS + aoff;
                  nb = ((y + TS) / TS - 1) * TS;
                  ONEWCELL(ha, -1, k, 0, na);              // dir == 1: stored as they are
                  ONEWCELL(hb, -1, k, 0, nb);
                  na += TS; nb += TS;
                }
              else
                { na = (((y + k) + (TS - aoff) - 1) / TS - 1) * TS + aoff;
                  nb = ((y + TS - 1) / TS - 1) * TS;
                  if (atop + 2 >= acap) status = LERR_CELLS;
                  else
                    { if (lead)
                        { arena[abase + cbase]     = LPebble{ -1, k, 0, y + k };
                          arena[abase + cbase + 1] = LPebble{ -1, k, 0, y };
                        }
                      avail0 = 2;
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
              yp = slide8(aseq, bseq, dir, off, kp, yp, hit);
              if (hit)
                { more = 0;
                  if (hit == 1) bclip = kp; else aclip = kp;
                }
              const int c = (yp << 1) + kp;
              while (yp + kp >= nap && status == 0)
                { ONEWCELL(ha, ha, kp, 0, nap); nap += TS; }
              if (DOB)
                while (yp >= nbp && status == 0)
                  { ONEWCELL(hb, hb, kp, 0, nbp); nbp += TS; }
              if (status != 0) { phase = PH_JOBEND; continue; }
              if (besta < c)
                { besta = trima = lasta = c;
                  besty = trimy = yp;
                  trimha = ha; trimhb = hb;
                }
              __syncwarp(gmask);
              if (lead)
                { SV(kp) = c; STL(kp) = (int) (uint32_t) PATH_INT; STH(kp) = (int) (uint32_t) (PATH_INT >> 32);
                  SMH(kp) = (PATH_LEN << 24) | (hb << 12) | ha;
                  SNA(kp) = nap;
                  if (DOB) SNB(kp) = nbp;
                  HD(g, H_AVAIL) = avail0;
                  HD(g, H_STATUS) = 0;
                }
              __syncwarp(gmask);
              CLIP_AFTER_WAVE
              cellcap = acap - atop; if (cellcap > LCELLS) cellcap = LCELLS;
              phase = PH_WAVE;
            }
          else if (phase == PH_ENDCALL)
            { int tx, ty, td, tha, thb;
              if (morem >= 0) { tx = morea - morey; ty = morey; td = mored; tha = moreha; thb = morehb; }
              else            { tx = trima - trimy; ty = trimy; td = trimd; tha = trimha; thb = trimhb; }
              LaneCall cc;
              cc.cells = abase + cbase; cc.dir = dir; cc.mida = dir * mida; cc.aoff = aoff;
              cc.ha = tha; cc.hb = thb; cc.x = dir * tx; cc.y = dir * ty; cc.d = td; cc.pad = 0;
              atop += HD(g, H_AVAIL);
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
              int rec = 0;
              if (lead) rec = atomicAdd(A.aln_top, 1);
              rec = __shfl_sync(gmask, rec, gl0);
              if (rec >= A.aln_cap) { status = LERR_POOL; phase = PH_JOBEND; continue; }
              AlnRec r;
              r.next = -1; r.comp = cm; r.bread = br; r.pad = 0;
              r.a[0] = a_ab; r.a[1] = a_bb; r.a[2] = a_ae; r.a[3] = a_be; r.a[4] = p_df; r.a[5] = 0;
              r.b[0] = b_ab; r.b[1] = b_bb; r.b[2] = b_ae; r.b[3] = b_be; r.b[4] = p_df; r.b[5] = 0;
              r.atrace = 0; r.btrace = 0;
              if (lead) A.alns[rec] = r;
              LaneUnwind u;
              u.ncalls = ncall; u.acomp = cm; u.job = jid; u.pad = 0;
              u.call[0] = c0; u.call[1] = c1;
              if (lead)
                { A.unwind[rec] = u;
                  if (last >= 0) A.alns[last].next = rec;
                }
              if (first < 0) first = rec;
              last = rec;
              count += 1;
              phase = PH_SEED;
            }
          else if (phase == PH_JOBEND)
            { if (lead)
                { AlignJob &job = A.jobs[jid];
                  job.first = (status == 0) ? first : -1;
                  job.count = (status == 0) ? count : 0;
                  job.status = status;
                  atomicMax(&A.stats[6], ((nwaves - jobw0) << 20) | (nalign - joba0));   // longest job (trace aid)
                  if (status != 0)
                    { atomicAdd(A.nfailed, 1);
                      atomicAdd(&A.stats[4], 1ull << (16 * (status > 3 ? 3 : status - 1)));   // why (trace aid)
                    }
                }
              phase = PH_IDLE;
            }
        }
      // ---- next seed of every slot that needs one (map.c:2487-2498: walk the chain's jumps until
      // a seed falls before the last kept alignment).  Cooperative: 32 jumps per step, prefix sums
      // of the (da, db) pairs, first lane whose a-position is below `alast`.
      { unsigned seek = __ballot_sync(0xffffffffu, phase == PH_SEED);
        const bool any = (seek != 0);
        while (seek)
          { const int sg = (__ffs(seek) - 1) / SEGL, sl = sg * SEGL;           // group and its first lane
            seek &= ~((SEGL == 32) ? 0xffffffffu : (((1u << SEGL) - 1) << sl));
            const long long ch = __shfl_sync(0xffffffffu, chain, sl);
            int s_sn = __shfl_sync(0xffffffffu, sn, sl), s_ap = __shfl_sync(0xffffffffu, apos, sl);
            int s_bp = __shfl_sync(0xffffffffu, bpos, sl);
            const int s_cl = __shfl_sync(0xffffffffu, clen, sl), s_al = __shfl_sync(0xffffffffu, alast, sl);
            const int s_st = __shfl_sync(0xffffffffu, status, sl);
            int found = -1, fa = 0, fb = 0;
            while (s_st == 0 && s_sn < s_cl)
              { const int i = s_sn + lane;
                const uint32_t jp = (i < s_cl) ? A.jumps[ch + i] : 0u;
                int da = (int) (jp & 0xffff), db = (int) (jp >> 16);
#pragma unroll
                for (int o = 1; o < 32; o <<= 1)
                  { const int ta = __shfl_up_sync(0xffffffffu, da, o), tb = __shfl_up_sync(0xffffffffu, db, o);
                    if (lane >= o) { da += ta; db += tb; }
                  }
                const int ap_i = s_ap - da, bp_i = s_bp - db;
                const unsigned hm = __ballot_sync(0xffffffffu, i < s_cl && ap_i < s_al);
                if (hm)
                  { const int l = __ffs(hm) - 1;
                    fa = __shfl_sync(0xffffffffu, ap_i, l); fb = __shfl_sync(0xffffffffu, bp_i, l);
                    found = s_sn + l + 1;
                    break;
                  }
                s_ap -= __shfl_sync(0xffffffffu, da, 31); s_bp -= __shfl_sync(0xffffffffu, db, 31);
                s_sn += 32;
              }
            if (g == sg)
              { if (found >= 0) { sn = found; apos = fa; bpos = fb; phase = PH_SEEDGO; }
                else            phase = PH_JOBEND;
              }
          }
        if (any)
          continue;                                      // let those groups start their calls first
      }
      if (__all_sync(0xffffffffu, phase == PH_DONE))
        break;

      // ---- top of a wave (align.c:592-690 / 1248-1345), every lane for its group
      int width = 0;
      if (phase == PH_WAVE)
        { bool go = more && (lasta >= besta - TRIM_MLAG);
          if (go && hgh < low)                            // empty band: the reference would read
            { if (lead) nempty += 1;                      // stale cells; stop (as align.cu does)
              go = false;
            }
          if (go && hgh - low + 6 > W)
            { status = LERR_BAND; go = false; }
          if (!go)
            phase = (status != 0) ? PH_JOBEND : PH_ENDCALL;
          else
            { low -= 1; hgh += 1; dif += 1;
              if (lead)
                { SNA(low) = SNA(low + 1); SNA(hgh) = SNA(hgh - 1);
                  if (DOB) { SNB(low) = SNB(low + 1); SNB(hgh) = SNB(hgh - 1); }
                }
              width = hgh - low + 1;
            }
        }
      __syncwarp();
      const int maxw = __reduce_max_sync(0xffffffffu, width);

      // ---- lane li of a group takes diagonal hgh - (base + li) of its band, SEGL per chunk
      for (int base = 0; base < maxw; base += SEGL)
        { const int  j = base + li;
          const bool act = (j < width);
          const int  ob = cur, nb_ = cur ^ 1;
          LPebble *const cells = arena + abase + cbase;

          int c = -IMAX, y = 0, m = 0, ha = 0, hb = 0, hit = 0, k = hgh - j;
          uint64_t b = 0;
          if (act)
            { // old band = (low, hgh) exclusive: the two outer diagonals are new this wave
              const int ap = (k + 1 < hgh) ? ST(g, ob, 0, k + 1) : LSENT;
              const int ac = (k > low && k < hgh) ? ST(g, ob, 0, k) : LSENT;
              const int am = (k - 1 > low) ? ST(g, ob, 0, k - 1) : LSENT;
              int src;
              if (ac < am)                                // align.c:712-741 / 1367-1396
                { if (am < ap) { c = ap + 1; src = k + 1; }
                  else         { c = am + 1; src = k - 1; }
                }
              else
                { if (ac < ap) { c = ap + 1; src = k + 1; }
                  else         { c = ac + 2; src = k; }
                }
              const int mh = ST(g, ob, 3, src);
              b = ((uint64_t) (uint32_t) ST(g, ob, 2, src) << 32) | (uint32_t) ST(g, ob, 1, src);
              m = (int) ((unsigned) mh >> 24); ha = mh & 0xfff; hb = (mh >> 12) & 0xfff;
              if ((b & PATH_TOP) != 0) m -= 1;
              b <<= 1;

              const int y0 = (c - k) >> 1;
              y = slide8(aseq, bseq, dir, off, k, y0, hit);
              const int r = y - y0;
              if (r > 0)                                  // closed form of align.c:764-767
                { const int rr = (r < 61) ? r : 61;
                  const uint64_t mask = ((1ull << rr) - 1) << (61 - rr);
                  m += rr - __popcll(b & mask);
                  b = (r >= 64) ? ~0ull : ((b << r) | ((1ull << r) - 1));
                }
              c = (y << 1) + k;

              int na = NAW(g, k);                         // align.c:771-793 / 1426-1448
              if (y + k >= na)
                { do
                    { if (dir * cells[ha].mark < na)
                        { const int av = atomicAdd(&HD(g, H_AVAIL), 1);
                          if (av >= cellcap) HD(g, H_STATUS) = LERR_CELLS;
                          else { cells[av] = LPebble{ ha, dir * k, dif, dir * na }; ha = av; }
                        }
                      na += TS;
                    }
                  while (y + k >= na);
                  NAW(g, k) = na;
                }
              if (DOB)                                    // align.c:795-817 / 1449-1471
                { int nb = NBW(g, k);
                  if (y >= nb)
                    { do
                        { if (dir * cells[hb].mark < nb)
                            { const int av = atomicAdd(&HD(g, H_AVAIL), 1);
                              if (av >= cellcap) HD(g, H_STATUS) = LERR_CELLS;
                              else { cells[av] = LPebble{ hb, dir * k, dif, dir * nb }; hb = av; }
                            }
                          nb += TS;
                        }
                      while (y >= nb);
                      NBW(g, k) = nb;
                    }
                }
              ST(g, nb_, 0, k) = c; ST(g, nb_, 1, k) = (int) (uint32_t) b; ST(g, nb_, 2, k) = (int) (uint32_t) (b >> 32);
              ST(g, nb_, 3, k) = (m << 24) | (hb << 12) | ha;
            }

          // record breakers in scan order within the group (align.c:819-833 / 1473-1487)
          int pm = c;
#pragma unroll
          for (int o = 1; o < SEGL; o <<= 1)
            { const int t = __shfl_up_sync(0xffffffffu, pm, o, SEGL);
              if (li >= o && t > pm) pm = t;
            }
          int before = __shfl_up_sync(0xffffffffu, pm, 1, SEGL);
          if (li == 0) before = -IMAX;
          const bool brk = act && c > besta && c > before;
          const bool good = brk && (m >= PATH_AVE);
          bool trim = false;
          if (good)
            { const int lo15 = (int) (b & TRIM_MASK), hi15 = (int) ((b >> TRIM_LEN) & TRIM_MASK);
              if (__ldg(A.spec.table + lo15) >= 0)
                if (__ldg(A.spec.table + hi15) + __ldg(A.spec.score + lo15) >= 0)
                  trim = true;
            }
          const unsigned xb = __ballot_sync(0xffffffffu, brk) & gmask;
          const unsigned xg = __ballot_sync(0xffffffffu, good) & gmask;
          const unsigned xt = __ballot_sync(0xffffffffu, trim) & gmask;
          const unsigned x2 = __ballot_sync(0xffffffffu, act && hit == 2) & gmask;
          const unsigned x1 = __ballot_sync(0xffffffffu, act && hit == 1) & gmask;
          const int lb = xb ? 31 - __clz(xb) : lane, lg = xg ? 31 - __clz(xg) : lane;
          const int lt = xt ? 31 - __clz(xt) : lane;
          const int l2 = x2 ? 31 - __clz(x2) : lane, l1 = x1 ? __ffs(x1) - 1 : lane;
          const int cb_ = __shfl_sync(0xffffffffu, c, lb), yb_ = __shfl_sync(0xffffffffu, y, lb);
          const int cg_ = __shfl_sync(0xffffffffu, c, lg);
          const int ct_ = __shfl_sync(0xffffffffu, c, lt), yt_ = __shfl_sync(0xffffffffu, y, lt);
          const int at_ = __shfl_sync(0xffffffffu, ha, lt), bt_ = __shfl_sync(0xffffffffu, hb, lt);
          const int k2_ = __shfl_sync(0xffffffffu, k, l2), k1_ = __shfl_sync(0xffffffffu, k, l1);
          if (xb) { besta = cb_; besty = yb_; }
          if (xg) lasta = cg_;
          if (xt) { trima = ct_; trimy = yt_; trimd = dif; trimha = at_; trimhb = bt_; }
          if (x2 | x1) more = 0;
          if (x2) aclip = k2_;                            // last writer in scan order
          if (x1) { if (bclip < k1_) bclip = k1_; }       // extreme k towards the scan start
        }
      __syncwarp();

      // ---- end of the wave, every lane for its group
      if (width > 0)
        { cur ^= 1;
          if (HD(g, H_STATUS) != 0) status = HD(g, H_STATUS);
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
              if (lead)
                { nwaves += 1;
                  ncells += (hgh - low) + 1;
                }
            }
        }
      __syncwarp();
    }
  if (lead)
    { atomicAdd(&A.stats[0], nalign); atomicAdd(&A.stats[1], nwaves);
      atomicAdd(&A.stats[2], ncells); atomicAdd(&A.stats[3], nempty);
    }
}

template <int G>
static void launch_group_g(const AlignArgs &A, int nblocks, cudaStream_t stream)
{ const bool dob = (A.do_b != 0);
  const size_t smem = (size_t) PACK_WARPS * (dob ? PACK_WARP_WORDS(G, true) : PACK_WARP_WORDS(G, false)) * sizeof(int);
  static bool attr_set = false;
  if (!attr_set)
    { CUDA_CHECK(cudaFuncSetAttribute(k_align_group<G, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int) (PACK_WARPS * PACK_WARP_WORDS(G, true) * sizeof(int))));
      CUDA_CHECK(cudaFuncSetAttribute(k_align_group<G, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int) (PACK_WARPS * PACK_WARP_WORDS(G, false) * sizeof(int))));
      attr_set = true;
    }
  if (dob) LAUNCH((k_align_group<G, true>), nblocks, PACK_WARPS * 32, smem, stream, A);
  else     LAUNCH((k_align_group<G, false>), nblocks, PACK_WARPS * 32, smem, stream, A);
}

// blocks such that every SM holds as many CTAs as fit; slots are persistent (jobs from a counter)
void launch_align_group(const AlignArgs &A, int njobs, cudaStream_t stream)
{ const int G = g_align_slots;
  const int per_block = PACK_WARPS * G;
  int nblocks = (njobs + per_block - 1) / per_block;
  const int cap = sm_count() * 8;
  if (nblocks > cap) nblocks = cap;
  if (G == 8)      launch_group_g<8>(A, nblocks, stream);
  else if (G == 2) launch_group_g<2>(A, nblocks, stream);
  else             launch_group_g<4>(A, nblocks, stream);
}

}  // namespace damgpu
