// Wave (O(ND)) local alignment, TWO candidate chains per warp ("duo" kernel): forward_wave
// align.c:353-1011, reverse_wave :1015-1720, Local_Alignment :1727-1946, the seeding loop of
// report_thread map.c:2487-2579.
//
// After WAVE_LAG trimming a wave spans ~6 diagonals (+2 new ones per wave), so a whole warp per
// alignment (align.cu) keeps a fifth of its lanes busy.  Here each 16-lane half of a warp owns one
// job and both halves run the SAME instruction stream:
//  * every shuffle and ballot is issued with the full mask from warp-uniform control flow; a half
//    only looks at its own 16 result bits / its own 16-lane shuffle segment, so nothing diverges
//    between the halves inside a wave.
//  * direction is data: a reverse wave is a forward wave in primed coordinates (everything times
//    dir) over sequences read backwards from the 2-bit images, so a half extending forward and a
//    half extending backward share every instruction.
//  * diagonal k' of a half lives in lane (-k') & 15: the scan order of the reference (hgh down to
//    low) is increasing lane number from lane (-hgh) & 15, the state of a diagonal never moves
//    between lanes, neighbours are one shuffle away.  Per diagonal: V, the 61-column match history
//    (kept inverted: 1 = difference, so a difference shifts in a one and a slide is one shift; M of
//    the reference is 61 - popcount and is never stored), the heads of the A and B Pebble chains
//    with the marks of those head cells (so the "already crossed" test of align.c:772 needs no
//    load), and the next trace coordinates NA, NB.
//  * the wave-uniform scalars the reference updates per diagonal (trima/trimy/trimha/trimhb,
//    besty) are captured in the registers of the lane that produced them; only the lane number is
//    kept per half, and the values are fetched once when the wave call ends.
//  * 1 % of the waves are wider than 16 diagonals (59 % of the calls meet one): the half then
//    spills its diagonals to a 64-slot window in global memory, runs those waves 16 diagonals at a
//    time from there and returns to registers when the band has shrunk.  Bands over 64 diagonals
//    (chimeric junctions, 30 %-error stretches: the band grows by two diagonals per wave until the wave
//    dies; a 256-slot window kept them here and was slower than the hand-off: 4.3 against 3.7 ms on C5)
//    fail the job, which is re-run by the warp kernel of align.cu (host loop in report.cu).
//  * everything that happens once per wave call or per job (job fetch, seed walk, wave 0,
//    Local_Alignment's F/R/re-run logic, record output) is a small scalar state machine run by
//    lane 0 of the half on a control record in shared memory.
//  * Pebble cells go to a per-job arena in global memory in any order (cell numbers differ from
//    the reference's, the links do not); k_unwind (align_unwind.cu) turns the chains of the kept
//    alignments into trace pairs afterwards, one thread per alignment.
#include "common.cuh"
#include "mapper.cuh"
#include "align.cuh"

namespace damgpu {

namespace {

constexpr int      TRIM_LEN = 15, DUB_TRIM = 45, PATH_LEN = 60;     // align.c:162-176
constexpr int      TRIM_MASK = 0x7fff, TRIM_MLAG = 250, WAVE_LAG = 30;
constexpr int      IMAX = 0x7fffffff;
constexpr int      NEG = -0x3fffffff;                   // "no point on this diagonal"
constexpr uint64_t HIST61 = 0x1fffffffffffffffull;      // the columns M counts: bits 0..PATH_LEN
constexpr uint64_t HIST0  = 0xf000000000000000ull;      // ~PATH_INT: the history before wave 0
constexpr unsigned FULL = 0xffffffffu;

enum { DERR_NONE = 0, DERR_BAND = 1, DERR_CELLS = 2, DERR_TRACE = 3, DERR_MULTI = 4, DERR_POOL = 13 };
enum { PH_IDLE = 0, PH_SEED, PH_SEEDGO, PH_START, PH_WAVE, PH_ENDCALL, PH_FINISH, PH_JOBEND, PH_DONE };

struct __align__(16) LPebble { int ptr, diag, diff, mark; };         // align.c:344-349

// window fields (wide mode, global memory): inherited state is double-buffered
enum { F_V = 0, F_TL, F_TH, F_HA, F_HB, F_MA, F_MB, F_INH };
constexpr int DUO_WIN_WORDS = (2 * F_INH + 2) * DUO_W;

// Control record of one half (shared memory).  Job and alignment bookkeeping, the parameters of
// the wave call that is about to start, the results of the one that has just ended.
struct __align__(16) DuoCtl
{ // job
  int phase, jid, ar, br, cm, alen, blen, clen, sn, apos, bpos, alast, first, last, count, status;
  long long chain, abase;
  const uint32_t *aw, *bw;              // 2-bit images of the two sequences
  int as, bs;                           // bit-pair offset of base 0 in them
  int acap, atop, astart, cbase;
  // alignment (Local_Alignment)
  int anti, dg, aoff, call, ncall, fshort, p_ab, p_bb, p_ae, p_be, p_df;
  LaneCall c0, c1;
  // wave call: constants
  int dir, mida, asd, bsd, aend, bend, alo, blo, cellcap;
  // wave call: start state / end state
  int low, hgh, more, besta, lasta, besty, avail, dif;
  int trima, trimy, trimd, trimha, trimhb;
  int k0, v0, ha0, hb0, na0, nb0, ma0, mb0;
  int aclip, bclip, morea, morey, mored, moreha, morehb, morem;
  // statistics
  unsigned long long nwaves, ncells, nalign, nempty, jobw0, joba0, jobc0;
};

__device__ __forceinline__ uint32_t win16(const uint32_t *__restrict__ w, int p)
{ const int i = p >> 4;
  return __funnelshift_r(__ldg(w + i), __ldg(w + i + 1), (unsigned) (p & 15) * 2);
}

// Slide along primed diagonal kp from primed b-coordinate yp while bases match (align.c:748-768 /
// 1403-1423).  Sixteen bases per step.  asd/bsd = bit-pair position of base 0 (forward) or of base
// -16 (reverse); aend/bend = primed end coordinates.  hit: 1 = ran into the end of B, 2 = of A (the
// reference tests B's terminator first).
__device__ __forceinline__ int slide2(const uint32_t *__restrict__ aw, const uint32_t *__restrict__ bw,
                                      int asd, int bsd, int dir, int aend, int bend,
                                      int kp, int yp, int &hit)
{ const int lim = min(bend - yp, aend - (yp + kp));
  int pa = asd + dir * (yp + kp), pb = bsd + dir * yp, run = 0;
  while (run < lim)
    { uint32_t x = win16(aw, pa) ^ win16(bw, pb);
      if (dir > 0) x = __brev(x);
      const int t = __clz(x) >> 1;                      // x == 0: 16
      run += t;
      if (t < 16) break;
      pa += 16 * dir; pb += 16 * dir;
    }
  if (run > lim) run = lim;
  yp += run;
  hit = (yp == bend) ? 1 : ((yp + kp == aend) ? 2 : 0);
  return yp;
}

// M of the reference for an inverted history word
__device__ __forceinline__ int hist_m(uint64_t t) { return PATH_LEN + 1 - __popcll(t & HIST61); }

// the trim test of align.c:824-826 on an inverted history word; sc0 = score[0]: the score table is
// 1000 popc(p) + score[0] (set_table, align.c:199-213: mscore + dscore = FRACTION)
__device__ __forceinline__ bool trim_ok(uint64_t t, const AlignSpecD &sp, int sc0)
{ const uint32_t b = ~(uint32_t) t;
  const int lo15 = (int) (b & TRIM_MASK), hi15 = (int) ((b >> TRIM_LEN) & TRIM_MASK);
  const int t_lo = __ldg(sp.table + lo15), t_hi = __ldg(sp.table + hi15);   // both in flight at once
  return (t_lo >= 0) && (t_hi + 1000 * __popc(lo15) + sc0 >= 0);
}

// how many trace coordinates n, n+TS, .. the point p has reached, and how many of those the
// inherited chain (head mark mk) has crossed already (align.c:771-793, generalised to several)
__device__ __forceinline__ void crossings(int p, int n, int mk, int TS, int &cnt, int &skip)
{ cnt = 0; skip = 0;
  if (p >= n)
    { const int over = p - n;
      cnt = (over < TS) ? 1 : over / TS + 1;
      const int d0 = mk - n;
      if (d0 >= 0)
        { skip = (d0 < TS) ? 1 : d0 / TS + 1;
          if (skip > cnt) skip = cnt;
        }
    }
}

// ---- the scalar side: lane 0 of a half advances its control record until a wave call is ready
// to run (PH_WAVE) or there is no more work (PH_DONE) --------------------------------------------
template <bool DOB>
__device__ __forceinline__ void duo_control(DuoCtl &C, const AlignArgs &A)
{ const int TS = A.spec.spacing;
  const int hithr = 3 * A.kmer;                           // HITMIN*Kmer, map.c:2419
  LPebble *const arena = reinterpret_cast<LPebble *>(A.lane_cells);
  while (true)
    { if (C.phase == PH_IDLE)
        { const int j = atomicAdd(A.job_counter, 1);
          if (j >= A.njobs) { C.phase = PH_DONE; return; }
          const int jid = A.job_list ? A.job_list[j] : j;
          const AlignJob job = A.jobs[jid];
          const Candidate cd = A.cand[job.cand];
          C.jid = jid; C.ar = job.read; C.br = cd.bread; C.cm = cd.comp;
          C.alen = A.rlen_a[C.ar]; C.blen = A.rlen_b[C.br];
          { const int64_t ob = A.boff_b[C.br], oa = A.boff_a[C.ar];
            C.bw = A.pk_b + (ob >> 4); C.bs = (int) (ob & 15);
            C.aw = (C.cm ? A.pk_ac : A.pk_a) + (oa >> 4); C.as = (int) (oa & 15);
          }
          C.chain = cd.chain; C.clen = cd.length; C.sn = 0;
          C.apos = cd.alast; C.bpos = cd.blast; C.alast = C.alen + 1;
          C.first = C.last = -1; C.count = 0; C.status = 0;
          C.jobw0 = C.nwaves; C.joba0 = C.nalign; C.jobc0 = C.ncells;
          C.acap = LANE_ARENA(C.alen / TS);
          C.abase = A.lane_cell_base[C.ar] + (long long) (jid - (int) A.lane_job_off[C.ar]) * C.acap;
          C.atop = 0;
          C.phase = PH_SEED;
        }
      else if (C.phase == PH_SEED)                         // map.c:2487-2498
        { bool found = false;
          int sn = C.sn, apos = C.apos, bpos = C.bpos;
          while (sn < C.clen)
            { const uint32_t jp = A.jumps[C.chain + sn];
              sn += 1;
              apos -= (int) (jp & 0xffff);
              bpos -= (int) (jp >> 16);
              if (apos < C.alast) { found = true; break; }
            }
          C.sn = sn; C.apos = apos; C.bpos = bpos;
          C.phase = found ? PH_SEEDGO : PH_JOBEND;
        }
      else if (C.phase == PH_SEEDGO)                       // map.c:2499-2513, align.c:1727-1805
        { if (C.cm) { const int ac = C.alen - C.apos, bc = C.blen - C.bpos; C.dg = ac - bc; C.anti = ac + bc; }
          else      { C.dg = C.apos - C.bpos; C.anti = C.apos + C.bpos; }
          if (((C.anti - C.dg) >> 1) < 0) { C.status = DERR_MULTI; C.phase = PH_JOBEND; continue; }
          C.aoff = C.cm ? C.alen % TS : 0;                 // align.c:1794-1797
          C.nalign += 1;
          C.p_ab = C.p_bb = C.p_ae = C.p_be = C.p_df = 0;
          C.astart = C.atop;
          C.call = 0; C.ncall = 0; C.fshort = 0;
          C.dir = 1; C.low = C.dg; C.mida = C.anti;        // forward wave from the seed
          C.phase = PH_START;
        }
      else if (C.phase == PH_START)                        // wave 0 (align.c:433-583 / 1093-1241)
        { // C.low (start diagonal) and C.mida are ACTUAL coordinates here; primed below
          const int dir = C.dir, k = C.low;
          LPebble *const cells = arena + C.abase + C.atop;
          const int cap = C.acap - C.atop;
          int y = (C.mida - k) >> 1, na, nb, ha = 0, hb = 1, ma, mb, avail = 2;
          if (cap < 2) { C.status = DERR_CELLS; C.phase = PH_JOBEND; continue; }
          if (dir > 0)
            { na = (((y + k) + (TS - C.aoff)) / TS - 1) * TS + C.aoff;
              nb = ((y + TS) / TS - 1) * TS;
              cells[0] = LPebble{ -1, k, 0, na };
              cells[1] = LPebble{ -1, k, 0, nb };
              ma = na; mb = nb;
              na += TS; nb += TS;
            }
          else
            { na = (((y + k) + (TS - C.aoff) - 1) / TS - 1) * TS + C.aoff;
              nb = ((y + TS - 1) / TS - 1) * TS;
              cells[0] = LPebble{ -1, k, 0, y + k };
              cells[1] = LPebble{ -1, k, 0, y };
              ma = y + k; mb = y;
            }
          // primed from here on
          const int kp = dir * k;
          int yp = dir * y, nap = dir * na, nbp = dir * nb, hit;
          ma *= dir; mb *= dir;
          C.cbase = C.atop;
          C.mida = dir * C.mida;
          C.asd = C.as + (dir < 0 ? -16 : 0); C.bsd = C.bs + (dir < 0 ? -16 : 0);
          C.aend = (dir > 0) ? C.alen : 0; C.bend = (dir > 0) ? C.blen : 0;
          C.alo = (dir > 0) ? 0 : -C.alen; C.blo = (dir > 0) ? 0 : -C.blen;
          C.cellcap = cap;
          int more = 1, aclip = IMAX, bclip = -IMAX, low = kp, hgh = kp;
          int besta = C.mida, lasta = C.mida, besty = yp;
          C.trima = C.morea = C.mida; C.trimy = C.morey = yp;
          C.trimd = C.mored = 0; C.trimha = C.moreha = 0; C.trimhb = C.morehb = 1; C.morem = -1;
          yp = slide2(C.aw, C.bw, C.asd, C.bsd, dir, C.aend, C.bend, kp, yp, hit);
          if (hit)
            { more = 0;
              if (hit == 1) bclip = kp; else aclip = kp;
            }
          const int c = (yp << 1) + kp;
          int status = 0;
          while (yp + kp >= nap)
            { if (avail >= cap) { status = DERR_CELLS; break; }
              cells[avail] = LPebble{ ha, k, 0, dir * nap };
              ha = avail++; ma = nap;
              nap += TS;
            }
          if (DOB)
            while (yp >= nbp && status == 0)
              { if (avail >= cap) { status = DERR_CELLS; break; }
                cells[avail] = LPebble{ hb, k, 0, dir * nbp };
                hb = avail++; mb = nbp;
                nbp += TS;
              }
          if (status != 0) { C.status = status; C.phase = PH_JOBEND; continue; }
          if (besta < c)
            { besta = C.trima = lasta = c;
              besty = C.trimy = yp;
              C.trimha = ha; C.trimhb = hb;
            }
          if (more == 0)                                   // align.c:558-583 on the one diagonal
            { const int xb = besta - besty;
              if (besty >= C.blo && besty < C.bend && xb >= C.alo && xb < C.aend)
                more = 1;
              if (hgh >= aclip)
                { hgh = aclip - 1;
                  if (C.morem <= PATH_LEN)
                    { C.morem = PATH_LEN; C.morea = c; C.morey = (c - aclip) / 2;
                      C.moreha = ha; C.morehb = hb;
                    }
                }
              if (low <= bclip)
                { low = bclip + 1;
                  if (C.morem <= PATH_LEN)
                    { C.morem = PATH_LEN; C.morea = c; C.morey = (c - bclip) / 2;
                      C.moreha = ha; C.morehb = hb;
                    }
                }
              aclip = IMAX; bclip = -IMAX;
            }
          C.aclip = aclip; C.bclip = bclip;
          C.low = low; C.hgh = hgh; C.more = more; C.besta = besta; C.lasta = lasta; C.besty = besty;
          C.avail = avail; C.dif = 0;
          C.v0 = c; C.ha0 = ha; C.hb0 = hb; C.na0 = nap; C.nb0 = nbp; C.ma0 = ma; C.mb0 = mb;
          C.k0 = kp;
          C.phase = PH_WAVE;
          return;
        }
      else if (C.phase == PH_ENDCALL)                      // align.c:895-898, 1810-1854
        { int tx, ty, td, tha, thb;
          if (C.morem >= 0) { tx = C.morea - C.morey; ty = C.morey; td = C.mored; tha = C.moreha; thb = C.morehb; }
          else              { tx = C.trima - C.trimy; ty = C.trimy; td = C.trimd; tha = C.trimha; thb = C.trimhb; }
          const int dir = C.dir;
          LaneCall cc;
          cc.cells = C.abase + C.cbase; cc.dir = dir; cc.mida = dir * C.mida; cc.aoff = C.aoff;
          cc.ha = tha; cc.hb = thb; cc.x = dir * tx; cc.y = dir * ty; cc.d = td; cc.pad = 0;
          C.atop += C.avail;
          if (dir > 0) { C.p_ae = cc.x; C.p_be = cc.y; C.p_df = td; }
          else         { C.p_ab = cc.x; C.p_bb = cc.y; C.p_df += td; }
          if (C.call == 0)                                 // forward done: reverse from the seed
            { C.c0 = cc; C.ncall = 1;
              C.fshort = ((C.p_ae + C.p_be) - C.anti < DUB_TRIM);
              C.call = 1; C.dir = -1; C.low = C.dg; C.mida = C.anti;
              C.phase = PH_START;
            }
          else if (C.call == 1)
            { C.c1 = cc; C.ncall = 2;
              const int rshort = (C.anti - (C.p_ab + C.p_bb) < DUB_TRIM);
              if (C.fshort && rshort)
                { C.p_ae = C.p_ab = (C.p_ab + C.p_ae) / 2;
                  C.p_be = C.p_bb = (C.p_bb + C.p_be) / 2;
                  C.ncall = 0;
                  C.phase = PH_FINISH;
                }
              else if (C.fshort)
                { C.call = 2; C.dir = 1; C.low = C.p_ab - C.p_bb; C.mida = C.p_ab + C.p_bb;
                  C.phase = PH_START;
                }
              else if (rshort)
                { C.call = 2; C.dir = -1; C.low = C.p_ae - C.p_be; C.mida = C.p_ae + C.p_be; C.p_df = 0;
                  C.phase = PH_START;
                }
              else
                C.phase = PH_FINISH;
            }
          else                                             // the re-run replaces both traces
            { C.c0 = cc; C.ncall = 1;
              C.phase = PH_FINISH;
            }
        }
      else if (C.phase == PH_FINISH)                       // align.c:1857-1912, map.c:2514-2579
        { int a_ab = C.p_ab, a_bb = C.p_bb, a_ae = C.p_ae, a_be = C.p_be;
          const int b_ab = C.p_bb, b_bb = C.p_ab, b_ae = C.p_be, b_be = C.p_ae;
          if (C.cm)
            { a_ab = C.alen - b_be; a_bb = C.blen - b_ae; a_ae = C.alen - b_bb; a_be = C.blen - b_ab; }
          if (a_ae - a_ab < hithr)
            { C.atop = C.astart;                           // dropped: its cells are released
              C.phase = PH_SEED;
              continue;
            }
          C.alast = a_ab;
          const int rec = atomicAdd(A.aln_top, 1);
          if (rec >= A.aln_cap) { C.status = DERR_POOL; C.phase = PH_JOBEND; continue; }
          AlnRec r;
          r.next = -1; r.comp = C.cm; r.bread = C.br; r.pad = 0;
          r.a[0] = a_ab; r.a[1] = a_bb; r.a[2] = a_ae; r.a[3] = a_be; r.a[4] = C.p_df; r.a[5] = 0;
          r.b[0] = b_ab; r.b[1] = b_bb; r.b[2] = b_ae; r.b[3] = b_be; r.b[4] = C.p_df; r.b[5] = 0;
          r.atrace = 0; r.btrace = 0;
          A.alns[rec] = r;
          LaneUnwind u;
          u.ncalls = C.ncall; u.acomp = C.cm; u.job = C.jid; u.pad = 0;
          u.call[0] = C.c0; u.call[1] = C.c1;
          A.unwind[rec] = u;
          if (C.last >= 0) A.alns[C.last].next = rec;
          if (C.first < 0) C.first = rec;
          C.last = rec;
          C.count += 1;
          C.phase = PH_SEED;
        }
      else if (C.phase == PH_JOBEND)
        { AlignJob &job = A.jobs[C.jid];
          job.first = (C.status == 0) ? C.first : -1;
          job.count = (C.status == 0) ? C.count : 0;
          job.status = C.status;
          atomicMax(&A.stats[6], ((C.nwaves - C.jobw0) << 20) | (C.nalign - C.joba0));   // longest job (trace aid)
          if (C.status != 0)
            { atomicAdd(A.nfailed, 1);
              atomicAdd(&A.stats[4], 1ull << (16 * (C.status > 3 ? 3 : C.status - 1)));   // why (trace aid)
              C.nwaves = C.jobw0; C.nalign = C.joba0; C.ncells = C.jobc0;   // the re-run counts them
            }
          else                                             // k_unwind may still fail the job: it then takes these back
            { job.nalign = (unsigned) (C.nalign - C.joba0); job.nwaves = (unsigned) (C.nwaves - C.jobw0);
              job.ncells = (unsigned) (C.ncells - C.jobc0);
            }
          C.phase = PH_IDLE;
        }
      else
        return;                                            // PH_WAVE, PH_DONE
    }
}

}  // namespace

#ifndef DUO_MINB
#define DUO_MINB 5
#endif
#ifdef DUO_DEBUG_DIV
__device__ unsigned long long g_duo_dbg[64];
#define DBG_EV(b) (dbg_ev |= (b))
#define DBG_CP(i) do { if (__activemask() != FULL && lead) atomicAdd(&g_duo_dbg[56 + (i)], 1ull); } while (0)
#else
#define DBG_EV(b)
#define DBG_CP(i)
#endif

template <bool DOB>
__global__ void __launch_bounds__(DUO_WARPS * 32, DUO_MINB)
k_align_duo(const __grid_constant__ AlignArgs A)
{ extern __shared__ __align__(16) unsigned char dsm[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int half = lane >> 4, hl = lane & 15;
  const unsigned dupsel = half ? 0x3232u : 0x1010u;       // PRMT selector: own ballot half, twice
  const bool lead = (hl == 0);
  DuoCtl &C = reinterpret_cast<DuoCtl *>(dsm)[wib * 2 + half];
  // the wide-band window of this slot: global memory (L1/L2-resident while in use; 1 % of the waves on plain
  // reads, more on chimeric / low-quality ones where it saves re-running the job in the warp kernel)
  int *const win = A.duo_win + ((size_t) (blockIdx.x * DUO_WARPS + wib) * 2 + half) * DUO_WIN_WORDS;
#define WF(f, b, k) win[((f) * 2 + (b)) * DUO_W + ((k) & (DUO_W - 1))]
#define WNA(k)      win[(2 * F_INH) * DUO_W + ((k) & (DUO_W - 1))]
#define WNB(k)      win[(2 * F_INH + 1) * DUO_W + ((k) & (DUO_W - 1))]
#define SHF(v, l)   __shfl_sync(FULL, (v), (l), 16)
  // a half's 16 ballot bits, duplicated into both halves of a word
#define DUP16(b)    __byte_perm((b), 0u, dupsel)
#define MY16(b)     (DUP16(b) & 0xffffu)
  // ... rotated into scan order: bits i and i+16 <-> diagonal kh - i (lane (s + i) & 15).  The lowest
  // scan index set is __ffs(x) - 1, the highest 15 - __clz(x).
#define ROT16(b)    __funnelshift_r(DUP16(b), DUP16(b), s)

  const int TS = A.spec.spacing;
  const int mgood = PATH_LEN + 1 - A.spec.ave_path;       // m >= PATH_AVE <=> popcount <= mgood
  const int sc0 = __ldg(A.spec.score);                    // score[p] = 1000 popc(p) + score[0] (align.c:199-213)
  LPebble *const arena = reinterpret_cast<LPebble *>(A.lane_cells);

  // per-diagonal state of the lane's diagonal (narrow mode)
  int rV = NEG, rHA = 0, rHB = 0, rNA = 0, rNB = 0, rMA = 0, rMB = 0;
  uint64_t rT = 0;
  // captured by the lane that produced them
  int tC = 0, tY = 0, tHA = 0, tHB = 0, bY = 0;
  // uniform per half
  int low = 0, hgh = 0, dif = 0, avail = 0, besta = 0, lasta = 0, trimd = 0, more = 0, status = 0;
  int tlane = 0, blane = 0, cur = 0, cellcap = 0;
  int dir = 1, asd = 0, bsd = 0, aend = 0, bend = 0;
  const uint32_t *aw = nullptr, *bw = nullptr;
  LPebble *cells = nullptr;
  unsigned ncells = 0;
  bool waving = false, wide = false, done = false;
#ifdef DUO_DEBUG_DIV
  unsigned dbg_ev = 0; bool dbg_split = false;
#endif

  if (lead)
    { C.phase = PH_IDLE;
      C.nwaves = C.ncells = C.nalign = C.nempty = 0;
    }
  __syncwarp();

  while (true)
    { bool nar;
      // The two halves must run every wave as ONE warp.  Divergence the compiler does not expect (a half
      // waiting for the other's lane 0 in the control section) leaves two thread groups that leapfrog
      // through the same code, each sync instruction taking its slow collective path and every
      // instruction issuing twice (measured: 17 active threads per instruction, 22 ms instead of 6).
      // Groups merge when both are runnable at the same instruction; a YIELD at the loop head makes the
      // group that arrives first let the other catch up.  There is no intrinsic for YIELD: the
      // compiler emits one at the head of a loop that contains a volatile load (a possible spin wait).
      { const int yield_ = *reinterpret_cast<volatile int *>(&C.phase); (void) yield_; }
#ifdef DUO_DEBUG_DIV
      { const bool split = (__activemask() != FULL);
        if (lead) atomicAdd(&g_duo_dbg[split ? 1 : 0], 1ull);
        if (split && !dbg_split && lead) atomicAdd(&g_duo_dbg[8 + (dbg_ev & 31)], 1ull);
        if (!split && dbg_split && lead) atomicAdd(&g_duo_dbg[40 + (dbg_ev & 15)], 1ull);
        dbg_split = split; dbg_ev = 0;
      }
#endif
      { const bool go = waving && (more != 0) && (lasta >= besta - TRIM_MLAG) && (status == 0) && (hgh >= low);
        nar = go && !wide && (hgh - low + 3 <= 16);
      }
      if (!__all_sync(FULL, nar))
        { // =========== everything that is not "both halves run a narrow wave" ===========
          // ---- a wave call ends (align.c:592 / 1248)
          { bool go = waving && (more != 0) && (lasta >= besta - TRIM_MLAG) && (status == 0);
            if (go && hgh < low)                          // empty band: the reference would read
              { if (lead) C.nempty += 1;                  // stale cells; stop (as align.cu does)
                go = false;
              }
            if (go && hgh - low + 3 > DUO_W)
              { status = DERR_BAND; go = false; }
            const bool ending = waving && !go;
            if (__any_sync(FULL, ending))
              { DBG_EV(1);
                const int xC = SHF(tC, tlane), xY = SHF(tY, tlane), xHA = SHF(tHA, tlane), xHB = SHF(tHB, tlane);
                if (ending)
                  { if (lead)
                      { C.trima = xC; C.trimy = xY; C.trimd = trimd; C.trimha = xHA; C.trimhb = xHB;
                        C.avail = avail;
                        C.nwaves += (unsigned) dif; C.ncells += ncells;
                        if (status != 0) { C.status = status; C.phase = PH_JOBEND; }
                        else             C.phase = PH_ENDCALL;
                      }
                    waving = false;
                  }
                __syncwarp();
              }
          }
          // ---- control: a half that is not inside a wave call lets its lane 0 run the state machine
          // (all 32 lanes walk through here together: every warp-level sync stays full-mask)
          { const bool ctl = !waving && !done;
            if (__any_sync(FULL, ctl))
              { DBG_EV(2);
                if (ctl && lead) duo_control<DOB>(C, A);
                __syncwarp();
                if (ctl)
                  { if (C.phase == PH_DONE)
                      done = true;
                    else
                      { dir = C.dir; asd = C.asd; bsd = C.bsd; aend = C.aend; bend = C.bend;
                        aw = C.aw; bw = C.bw;
                        cells = arena + C.abase + C.cbase; cellcap = C.cellcap;
                        low = C.low; hgh = C.hgh; more = C.more; besta = C.besta; lasta = C.lasta;
                        avail = C.avail; dif = 0; trimd = 0; status = 0;
                        tC = C.trima; tY = C.trimy; tHA = C.trimha; tHB = C.trimhb; bY = C.besty;
                        tlane = 0; blane = 0;             // every lane holds the wave-0 values
                        if (((-C.k0) & 15) == hl)         // the one diagonal of wave 0
                          { rV = C.v0; rT = HIST0; rHA = C.ha0; rHB = C.hb0; rNA = C.na0; rNB = C.nb0;
                            rMA = C.ma0; rMB = C.mb0;
                          }
                        ncells = 0;
                        waving = true; wide = false; cur = 0;
                      }
                  }
                __syncwarp();
              }
          }
          if (__all_sync(FULL, done))
            break;
          const bool go = waving && (more != 0) && (lasta >= besta - TRIM_MLAG) && (status == 0) && (hgh >= low)
                          && (hgh - low + 3 <= DUO_W);

          // ---- registers <-> window when the next wave needs more / no more than 16 lanes
          { const int wnext = hgh - low + 3;
            const bool spill = go && !wide && wnext > 16, fill = go && wide && wnext <= 16;
            if (__any_sync(FULL, spill || fill))
              { DBG_EV(4);
                const int s0 = (-hgh) & 15, r0 = (hl - s0) & 15, k = hgh - r0;
                const bool have = (r0 < hgh - low + 1);
                if (spill)
                  { if (have)
                      { WF(F_V, 0, k) = rV; WF(F_TL, 0, k) = (int) (uint32_t) rT; WF(F_TH, 0, k) = (int) (uint32_t) (rT >> 32);
                        WF(F_HA, 0, k) = rHA; WF(F_HB, 0, k) = rHB; WF(F_MA, 0, k) = rMA; WF(F_MB, 0, k) = rMB;
                        WNA(k) = rNA; WNB(k) = rNB;
                      }
                    cur = 0; wide = true;
                  }
                __syncwarp();
                if (fill)
                  { if (have)
                      { rV = WF(F_V, cur, k);
                        rT = ((uint64_t) (uint32_t) WF(F_TH, cur, k) << 32) | (uint32_t) WF(F_TL, cur, k);
                        rHA = WF(F_HA, cur, k); rHB = WF(F_HB, cur, k); rMA = WF(F_MA, cur, k); rMB = WF(F_MB, cur, k);
                        rNA = WNA(k); rNB = WNB(k);
                      }
                    wide = false;
                  }
                __syncwarp();
              }
          }

          // ---- wide wave: 17..64 diagonals, state in the window, 16 at a time
          const bool wid = go && wide;
          if (__any_sync(FULL, wid))
            { DBG_EV(8);
              if (wid) { low -= 1; hgh += 1; dif += 1; }
              const int kh = hgh, kl = low;
              const int width = wid ? hgh - low + 1 : 0;
              const int nxt = cur ^ 1;
              if (wid && lead)
                { WNA(kl) = WNA(kl + 1); WNA(kh) = WNA(kh - 1);
                  if (DOB) { WNB(kl) = WNB(kl + 1); WNB(kh) = WNB(kh - 1); }
                }
              __syncwarp();
              int aclip = IMAX, bclip = -IMAX, rm = besta;
              const int nch = (max(width, __shfl_xor_sync(FULL, width, 16)) + 15) >> 4;   // warp-uniform
              for (int ch = 0; ch < nch; ch++)
                { const int r = ch * 16 + hl, k = kh - r;  // ballot bit i of a half <-> r = ch*16 + i
                  const bool act = wid && (r < width);
                  int c = NEG, y = 0, hit = 0, ha = 0, hb = 0, ma = 0, mb = 0, na = 0, nb = 0;
                  int cntA = 0, skipA = 0, cntB = 0, skipB = 0;
                  uint64_t t = 0;
                  if (act)
                    { // old band = (kl, kh) exclusive: the two outer diagonals are new this wave
                      const int vp = (k + 1 < kh) ? WF(F_V, cur, k + 1) : NEG;
                      const int vc = (k > kl && k < kh) ? WF(F_V, cur, k) : NEG;
                      const int vn = (k - 1 > kl) ? WF(F_V, cur, k - 1) : NEG;
                      int src;
                      if (vc < vn)                        // align.c:712-741 / 1367-1396
                        { if (vn < vp) { c = vp + 1; src = k + 1; }
                          else         { c = vn + 1; src = k - 1; }
                        }
                      else
                        { if (vc < vp) { c = vp + 1; src = k + 1; }
                          else         { c = vc + 2; src = k; }
                        }
                      t = ((uint64_t) (uint32_t) WF(F_TH, cur, src) << 32) | (uint32_t) WF(F_TL, cur, src);
                      ha = WF(F_HA, cur, src); ma = WF(F_MA, cur, src);
                      if (DOB) { hb = WF(F_HB, cur, src); mb = WF(F_MB, cur, src); }
                      na = WNA(k); nb = WNB(k);
                      t = (t << 1) | 1ull;
                      const int y0 = (c - k) >> 1;
                      y = slide2(aw, bw, asd, bsd, dir, aend, bend, k, y0, hit);
                      const int run = y - y0;
                      t = (run >= 64) ? 0ull : (t << run);
                      c = (y << 1) + k;
                      crossings(y + k, na, ma, TS, cntA, skipA);
                      if (DOB) crossings(y, nb, mb, TS, cntB, skipB);
                    }
                  { int remA = cntA - skipA, remB = cntB - skipB;
                    int nxA = na + TS * skipA, nxB = nb + TS * skipB;
                    unsigned pa = __ballot_sync(FULL, remA > 0), pb = __ballot_sync(FULL, remB > 0);
                    while (pa | pb)
                      { const unsigned a16 = MY16(pa), b16 = MY16(pb), lt = (1u << hl) - 1;
                        const int na_ = __popc(a16), tot = na_ + __popc(b16);
                        if (avail + tot > cellcap)
                          { if (tot) status = DERR_CELLS;
                            remA = remB = 0;
                          }
                        else
                          { if (remA > 0)
                              { const int ix = avail + __popc(a16 & lt);
                                cells[ix] = LPebble{ ha, dir * k, dif, dir * nxA };
                                ha = ix; ma = nxA; nxA += TS; remA -= 1;
                              }
                            if (remB > 0)
                              { const int ix = avail + na_ + __popc(b16 & lt);
                                cells[ix] = LPebble{ hb, dir * k, dif, dir * nxB };
                                hb = ix; mb = nxB; nxB += TS; remB -= 1;
                              }
                            avail += tot;
                          }
                        pa = __ballot_sync(FULL, remA > 0); pb = __ballot_sync(FULL, remB > 0);
                      }
                  }
                  if (act)
                    { WNA(k) = na + TS * cntA; WF(F_V, nxt, k) = c;
                      WF(F_TL, nxt, k) = (int) (uint32_t) t; WF(F_TH, nxt, k) = (int) (uint32_t) (t >> 32);
                      WF(F_HA, nxt, k) = ha; WF(F_MA, nxt, k) = ma;
                      if (DOB) { WNB(k) = nb + TS * cntB; WF(F_HB, nxt, k) = hb; WF(F_MB, nxt, k) = mb; }
                    }
                  // record breakers, the running maximum carries over from the first chunk
                  { const unsigned cb = __ballot_sync(FULL, c > besta);
                    unsigned rest = MY16(cb), brk = 0;
                    while (__any_sync(FULL, rest != 0))
                      { const int i = __ffs(rest) - 1;
                        const int v = SHF(c, i);
                        if (rest != 0 && v > rm) { rm = v; brk |= 1u << i; }
                        rest &= rest - 1;
                      }
                    const bool isb = act && ((brk >> hl) & 1u);
                    const bool good = isb && (__popcll(t & HIST61) <= mgood);
                    const bool trim = good && trim_ok(t, A.spec, sc0);
                    const unsigned gs = MY16(__ballot_sync(FULL, good)), ts = MY16(__ballot_sync(FULL, trim));
                    const int cg = SHF(c, 31 - __clz(gs));
                    if (brk)
                      { blane = 31 - __clz(brk);
                        if (hl == blane) bY = y;
                        if (gs) lasta = cg;
                        if (ts)
                          { tlane = 31 - __clz(ts);
                            trimd = dif;
                            if (hl == tlane) { tC = c; tY = y; tHA = ha; tHB = hb; }
                          }
                      }
                  }
                  { const unsigned as_ = MY16(__ballot_sync(FULL, act && hit == 2));
                    const unsigned bs_ = MY16(__ballot_sync(FULL, act && hit == 1));
                    if (as_ | bs_) more = 0;
                    if (as_) aclip = kh - (ch * 16 + 31 - __clz(as_));        // last writer in scan order
                    if (bs_ && bclip == -IMAX) bclip = kh - (ch * 16 + __ffs(bs_) - 1);   // first in scan order
                  }
                }
              if (wid) besta = rm;
              __syncwarp();
              if (wid) cur = nxt;
              const int ybw = SHF(bY, blane), moremw = C.morem;
              __syncwarp();
              if (wid && more == 0)                       // align.c:848-875 / 1502-1529
                { const int yb = ybw, xb = besta - yb;
                  if (yb >= C.blo && yb < bend && xb >= C.alo && xb < aend)
                    more = 1;
                  int morem = moremw;
                  if (hgh >= aclip)
                    { hgh = aclip - 1;
                      const uint64_t ta = ((uint64_t) (uint32_t) WF(F_TH, cur, aclip) << 32) | (uint32_t) WF(F_TL, cur, aclip);
                      const int am = hist_m(ta);
                      if (morem <= am)
                        { morem = am;
                          const int av = WF(F_V, cur, aclip);
                          if (lead)
                            { C.morem = am; C.morea = av; C.morey = (av - aclip) / 2; C.mored = dif;
                              C.moreha = WF(F_HA, cur, aclip); C.morehb = WF(F_HB, cur, aclip);
                            }
                        }
                    }
                  if (low <= bclip)
                    { low = bclip + 1;
                      const uint64_t tb2 = ((uint64_t) (uint32_t) WF(F_TH, cur, bclip) << 32) | (uint32_t) WF(F_TL, cur, bclip);
                      const int bm = hist_m(tb2);
                      if (morem <= bm)
                        { const int bv = WF(F_V, cur, bclip);
                          if (lead)
                            { C.morem = bm; C.morea = bv; C.morey = (bv - bclip) / 2; C.mored = dif;
                              C.moreha = WF(F_HA, cur, bclip); C.morehb = WF(F_HB, cur, bclip);
                            }
                        }
                    }
                }
              __syncwarp();
              // trim the band (align.c:877-885 / 1531-1539)
              { const int n = besta - WAVE_LAG;
                int first = -1, last = -1;                // scan indices of the first / last diagonal kept
                for (int ch = 0; ch < nch; ch++)
                  { const int r = ch * 16 + hl, k = kh - r;
                    const unsigned g = MY16(__ballot_sync(FULL, wid && r < width && k >= low && k <= hgh
                                                                && WF(F_V, cur, k) >= n));
                    if (g)
                      { if (first < 0) first = ch * 16 + __ffs(g) - 1;
                        last = ch * 16 + 31 - __clz(g);
                      }
                  }
                if (wid)
                  { if (first >= 0)
                      { hgh = kh - first;
                        low = kh - last;
                      }
                    else
                      hgh = low - 1;
                    ncells += (unsigned) (hgh - low + 1);
                  }
              }
            }

          nar = go && !wide;
          if (!__any_sync(FULL, nar))
            continue;
          __syncwarp();
        }

      // =================================== narrow wave: one diagonal per lane, state in registers
      DBG_CP(0);
      if (nar) { low -= 1; hgh += 1; dif += 1; }
      const int kh = hgh;                                 // lane mapping of this wave
      const int s = (-kh) & 15, r = (hl - s) & 15, k = kh - r;
      const int width = hgh - low + 1;
      const bool act = nar && (r < width);
      const bool old = act && r > 0 && r < width - 1;     // the diagonal existed before this wave
      const int vold = old ? rV : NEG;
      const int vp = SHF(vold, hl - 1), vn = SHF(vold, hl + 1);     // diagonals k+1, k-1
      int c, srcl;
      if (vold < vn)                                      // align.c:712-741 / 1367-1396
        { if (vn < vp) { c = vp + 1; srcl = hl - 1; }
          else         { c = vn + 1; srcl = hl + 1; }
        }
      else
        { if (vold < vp) { c = vp + 1; srcl = hl - 1; }
          else           { c = vold + 2; srcl = hl; }
        }
      uint64_t t;
      { const unsigned tl = SHF((unsigned) rT, srcl), th = SHF((unsigned) (rT >> 32), srcl);
        t = ((uint64_t) th << 32) | tl;
      }
      int ha = SHF(rHA, srcl), ma = SHF(rMA, srcl), hb = 0, mb = 0;
      // a new outer diagonal takes NA/NB of its inner neighbour (align.c:678-690), which is its source
      { const int nas = SHF(rNA, srcl);
        if (act && !old) rNA = nas;
      }
      if (DOB)
        { hb = SHF(rHB, srcl); mb = SHF(rMB, srcl);
          const int nbs = SHF(rNB, srcl);
          if (act && !old) rNB = nbs;
        }
      int y = 0, hit = 0, remA = 0, remB = 0, nxA = 0, nxB = 0;
      DBG_CP(1);
      if (act)
        { t = (t << 1) | 1ull;                             // the difference
          const int y0 = (c - k) >> 1;
          y = slide2(aw, bw, asd, bsd, dir, aend, bend, k, y0, hit);
          const int run = y - y0;
          t = (run >= 64) ? 0ull : (t << run);             // matches
          c = (y << 1) + k;
          int cnt, skip;
          crossings(y + k, rNA, ma, TS, cnt, skip);        // align.c:771-793 / 1426-1448
          remA = cnt - skip; nxA = rNA + TS * skip; rNA += TS * cnt;
          if (DOB)
            { crossings(y, rNB, mb, TS, cnt, skip);        // align.c:795-817 / 1449-1471
              remB = cnt - skip; nxB = rNB + TS * skip; rNB += TS * cnt;
            }
        }
      else
        c = NEG;

      // Pebble cells (numbered by ballot rank, one A and one B cell per lane and round); sequence ends
      DBG_CP(2);
      bool anyhit = false;
      if (__any_sync(FULL, (remA > 0) || (remB > 0) || (hit != 0)))
        { DBG_EV(16);
          unsigned pa = __ballot_sync(FULL, remA > 0), pb = __ballot_sync(FULL, remB > 0);
          while (pa | pb)
            { const unsigned a16 = MY16(pa), b16 = MY16(pb), lt = (1u << hl) - 1;
              const int na_ = __popc(a16), tot = na_ + __popc(b16);
              if (avail + tot > cellcap)
                { if (tot) status = DERR_CELLS;
                  remA = remB = 0;
                }
              else
                { if (remA > 0)
                    { const int ix = avail + __popc(a16 & lt);
                      cells[ix] = LPebble{ ha, dir * k, dif, dir * nxA };
                      ha = ix; ma = nxA; nxA += TS; remA -= 1;
                    }
                  if (remB > 0)
                    { const int ix = avail + na_ + __popc(b16 & lt);
                      cells[ix] = LPebble{ hb, dir * k, dif, dir * nxB };
                      hb = ix; mb = nxB; nxB += TS; remB -= 1;
                    }
                  avail += tot;
                }
              pa = __ballot_sync(FULL, remA > 0); pb = __ballot_sync(FULL, remB > 0);
            }
          anyhit = __any_sync(FULL, hit != 0);
        }
      DBG_CP(3);
      if (act)
        { rV = c; rT = t; rHA = ha; rMA = ma;
          if (DOB) { rHB = hb; rMB = mb; }
        }

      // record breakers in scan order (align.c:819-833 / 1473-1487): a point breaks the record when it
      // lies beyond besta and beyond every point before it in the scan -- an exclusive prefix maximum
      // over the circular lane order
      { int pm = c;                                       // NEG on lanes outside the band
#pragma unroll
        for (int o = 1; o < 16; o <<= 1)
          { const int q = SHF(pm, hl - o);
            if (r >= o) pm = max(pm, q);
          }
        int before = SHF(pm, hl - 1);
        before = (r == 0) ? besta : max(before, besta);
        const bool isb = act && (c > before);
        const bool good = isb && (__popcll(t & HIST61) <= mgood);
        const bool trim = good && trim_ok(t, A.spec, sc0);
        const unsigned rb = ROT16(__ballot_sync(FULL, isb)), rg = ROT16(__ballot_sync(FULL, good));
        const unsigned rt = ROT16(__ballot_sync(FULL, trim));
        const int bl = (s + 15 - __clz(rb)) & 15, gl = s + 15 - __clz(rg);
        const int cb = SHF(c, bl), cg = SHF(c, gl);
        if (rb)
          { besta = cb; blane = bl;
            if (hl == bl) bY = y;
          }
        if (rg) lasta = cg;
        if (rt)
          { tlane = (s + 15 - __clz(rt)) & 15;
            trimd = dif;
            if (hl == tlane) { tC = c; tY = y; tHA = ha; tHB = hb; }
          }
      }

      // sequence ends (align.c:752-763,848-875 / 1407-1418,1502-1529)
      DBG_CP(4);
      if (anyhit)
        { const unsigned as_ = ROT16(__ballot_sync(FULL, act && hit == 2)), bs_ = ROT16(__ballot_sync(FULL, act && hit == 1));
          int aclip = IMAX, bclip = -IMAX;
          if (as_ | bs_) more = 0;
          if (as_) aclip = kh - (15 - __clz(as_));        // last writer in scan order
          if (bs_) bclip = kh - (__ffs(bs_) - 1);         // extreme k towards the scan start
          const bool clip = nar && (more == 0);
          const int yb = SHF(bY, blane);
          const int la = (-aclip) & 15, lb = (-bclip) & 15;
          const int mloc = hist_m(rT), morem0 = C.morem, alo = C.alo, blo = C.blo;
          __syncwarp();
          const int am = SHF(mloc, la), av = SHF(rV, la), aha = SHF(rHA, la), ahb = SHF(rHB, la);
          const int bm = SHF(mloc, lb), bv = SHF(rV, lb), bha = SHF(rHA, lb), bhb = SHF(rHB, lb);
          if (clip)
            { const int xb = besta - yb;
              if (yb >= blo && yb < bend && xb >= alo && xb < aend)
                more = 1;
              int morem = morem0;
              if (hgh >= aclip)
                { hgh = aclip - 1;
                  if (morem <= am)
                    { morem = am;
                      if (lead)
                        { C.morem = am; C.morea = av; C.morey = (av - aclip) / 2; C.mored = dif;
                          C.moreha = aha; C.morehb = ahb;
                        }
                    }
                }
              if (low <= bclip)
                { low = bclip + 1;
                  if (morem <= bm)
                    { if (lead)
                        { C.morem = bm; C.morea = bv; C.morey = (bv - bclip) / 2; C.mored = dif;
                          C.moreha = bha; C.morehb = bhb;
                        }
                    }
                }
            }
          __syncwarp();
        }

      // trim the band to within WAVE_LAG of the best point (align.c:877-885 / 1531-1539)
      DBG_CP(5);
      { const unsigned g = ROT16(__ballot_sync(FULL, act && k >= low && k <= hgh && rV >= besta - WAVE_LAG));
        if (nar)
          { if (g)
              { hgh = kh - (__ffs(g) - 1);
                low = kh - (15 - __clz(g));
              }
            else
              hgh = low - 1;
            ncells += (unsigned) (hgh - low + 1);
          }
      }
    }

  if (lead)
    { atomicAdd(&A.stats[0], C.nalign); atomicAdd(&A.stats[1], C.nwaves);
      atomicAdd(&A.stats[2], C.ncells); atomicAdd(&A.stats[3], C.nempty);
    }
#undef WF
#undef WNA
#undef WNB
#undef SHF
#undef DUP16
#undef MY16
#undef ROT16
}

size_t duo_smem_bytes()
{ return (size_t) DUO_WARPS * 2 * sizeof(DuoCtl); }

// bytes of wide-band windows for a grid of up to `nblocks` CTAs (AlignArgs.duo_win)
size_t duo_window_bytes(int nblocks)
{ return (size_t) nblocks * DUO_WARPS * 2 * DUO_WIN_WORDS * sizeof(int); }

int duo_max_blocks()
{ return sm_count() * 8; }

// persistent half-warp slots: as many CTAs as fit on every SM, jobs from a counter
void launch_align_duo(const AlignArgs &A, int njobs, cudaStream_t stream)
{ const bool dob = (A.do_b != 0);
  const size_t smem = duo_smem_bytes();
  static int per_sm = 0;
  if (per_sm == 0)
    { CUDA_CHECK(cudaFuncSetAttribute(k_align_duo<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
      CUDA_CHECK(cudaFuncSetAttribute(k_align_duo<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
      int a = 0, b = 0;
      CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_align_duo<true>, DUO_WARPS * 32, smem));
      CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_align_duo<false>, DUO_WARPS * 32, smem));
      per_sm = a < b ? a : b;
      if (per_sm < 1) per_sm = 1;
    }
  // persistent slots (halves) pull jobs from a counter; the largest resident grid is the fastest (a grid
  // shrunk to whole job rounds -- every slot busy to the end -- was 12 % slower: fewer warps hide less latency)
  const int per_block = DUO_WARPS * 2;
  const int cap = sm_count() * per_sm;
  int nblocks = (njobs + per_block - 1) / per_block;
  if (nblocks > cap) nblocks = cap;
  if (nblocks > duo_max_blocks()) nblocks = duo_max_blocks();
  if (dob) LAUNCH(k_align_duo<true>, nblocks, DUO_WARPS * 32, smem, stream, A);
  else     LAUNCH(k_align_duo<false>, nblocks, DUO_WARPS * 32, smem, stream, A);
#ifdef DUO_DEBUG_DIV
  { unsigned long long h[64];
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(h, g_duo_dbg, sizeof(h));
    fprintf(stderr, "[duo dbg] rounds converged %llu split %llu\n  became split after events:", h[0], h[1]);
    for (int i = 0; i < 32; i++) if (h[8 + i]) fprintf(stderr, " %d:%llu", i, h[8 + i]);
    fprintf(stderr, "\n  became converged after events:");
    for (int i = 0; i < 16; i++) if (h[40 + i]) fprintf(stderr, " %d:%llu", i, h[40 + i]);
    fprintf(stderr, "\n  split at checkpoints:");
    for (int i = 0; i < 8; i++) fprintf(stderr, " %d:%llu", i, h[56 + i]);
    fprintf(stderr, "\n");
    memset(h, 0, sizeof(h)); cudaMemcpyToSymbol(g_duo_dbg, h, sizeof(h));
  }
#endif
}

}  // namespace damgpu
