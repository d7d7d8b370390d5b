// Unwinding of Pebble chains into trace pairs (align.c:900-1007 / 1554-1717), one thread per
// (kept alignment, chain).  The duo kernel (align_duo.cu) leaves the cells of every wave call in
// the job's arena and records per alignment which calls make up its traces (LaneUnwind).
//
// The reference reverses the linked chain in place, walks it from the start and appends pairs; that
// is two dependent passes over memory written long ago.  Here the chain is walked ONCE, from its head
// backwards, and every pair goes straight to its final place in the trace pool:
//  * consecutive cells of a chain are exactly one trace spacing apart (a path meets every trace
//    coordinate once, and the "already crossed" rule of align.c:772 keeps one cell per coordinate),
//    so the number of cells n follows from the marks of the head and of the first cell; the walk
//    checks that it arrives at the first cell after n steps, anything else fails the job (it is
//    re-run by the warp kernel, which unwinds the reference's way);
//  * with n, the end-of-chain decision (append a last pair / stretch the last one, align.c:946-955,
//    1617-1626) and the first-pair decision of a reverse call (align.c:1585-1600) known up front,
//    the trace length is known before the walk: the pool is claimed first, pair j of the forward
//    call and push q of the reverse call have fixed positions (mirrored when the A read is
//    complemented, align.c:1858-1884);
//  * the additions the reference makes to pairs written earlier (reverse call stretching the forward
//    call's first pair) are applied afterwards by the same thread; uint16 additions commute.
#include "common.cuh"
#include "mapper.cuh"
#include "align.cuh"

namespace damgpu {

namespace {

enum { UERR_TRACE = 3, UERR_POOL = 13 };
enum { T_NONE = 0, T_PUSH = 1, T_ADJ = 2 };

struct Cell { int ptr, diag, diff, mark; };                           // Pebble, align.c:344-349

__device__ __forceinline__ Cell load_cell(const int4 *cells, int i)
{ const int4 v = __ldg(cells + i);
  Cell c; c.ptr = v.x; c.diag = v.y; c.diff = v.z; c.mark = v.w;
  return c;
}

// What one wave call contributes to one chain's trace, decided from the two ends of the chain.
struct Plan
{ const int4 *cells;
  int dir, sk, fi, head, n;             // sk: +1 A chain (a = mark - diag), -1 B chain (a = mark + diag)
  int a0;                               // a of the first cell (d of it is 0)
  int aL, eL;                           // a and diff of the last element
  int T1, Q, trimd;                     // tested trim coordinate, the other one, differences
  int tail, special, virt, base;        // T_*; reverse: first-pair handling, virtual first element, push base
  int pairs;                            // pairs (forward) / pushes (reverse) this call adds
};

__device__ void make_plan(Plan &P, const LaneCall &cc, const int4 *arena, int which, int TS, int prevtlen)
{ P.cells = arena + cc.cells;
  P.dir = cc.dir; P.sk = which ? -1 : 1; P.fi = which; P.head = which ? cc.hb : cc.ha;
  P.T1 = which ? cc.y : cc.x; P.Q = which ? cc.x : cc.y; P.trimd = cc.d;
  const Cell F = load_cell(P.cells, P.fi), H = load_cell(P.cells, P.head);
  const int off = which ? 0 : cc.aoff;
  int kL;
  P.special = T_NONE; P.virt = 0; P.base = 0;
  if (P.dir > 0)
    { P.n = (H.mark - F.mark) / TS;
      P.a0 = (cc.mida - P.sk * F.diag) / 2;
      if (P.n > 0) { P.aL = H.mark - P.sk * H.diag; P.eL = H.diff; kL = H.diag; }
      else         { P.aL = P.a0; P.eL = 0; kL = F.diag; }
      P.tail = (P.aL + P.sk * kL != P.T1) ? T_PUSH : ((P.aL != P.Q) ? T_ADJ : T_NONE);
      P.pairs = P.n + (P.tail == T_PUSH);
    }
  else
    { P.n = (F.mark - H.mark + TS - 1) / TS;
      P.a0 = F.mark - P.sk * F.diag;
      const bool offgrid = (F.mark % TS != off);
      if (P.n > 0) { P.aL = H.mark - P.sk * H.diag; P.eL = H.diff; kL = H.diag; }
      else         { P.aL = P.a0; P.eL = 0; kL = F.diag; }
      if (offgrid)
        { P.special = (prevtlen == 0) ? T_PUSH : T_ADJ;
          P.base = (P.special == T_PUSH) ? 0 : -1;
          P.virt = (P.n == 0);
        }
      if (P.virt)
        { P.tail = T_NONE;
          P.pairs = (P.special == T_PUSH) ? 1 : 0;
        }
      else
        { P.tail = (P.aL + P.sk * kL != P.T1) ? T_PUSH : ((P.aL != P.Q) ? T_ADJ : T_NONE);
          P.pairs = P.base + P.n + (P.tail == T_PUSH);
        }
    }
}

__device__ __forceinline__ uint32_t mkpair(int lo, int hi)
{ return (uint32_t) (uint16_t) lo | ((uint32_t) (uint16_t) hi << 16); }

}  // namespace

__global__ void __launch_bounds__(128)
k_unwind(const __grid_constant__ AlignArgs A)
{ const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = A.do_b ? (t >> 1) : t, which = A.do_b ? (t & 1) : 0;
  int naln = *A.aln_top;
  if (naln > A.aln_cap) naln = A.aln_cap;
  if (i >= naln) return;
  const LaneUnwind &u = A.unwind[i];
  const int ncalls = u.ncalls;
  if (ncalls < 0) return;                               // a record of the warp kernel
  const int job = u.job, acomp = u.acomp;
  if (A.jobs[job].status != 0) return;                  // the job failed later: it is re-run whole
  const int TS = A.spec.spacing;
  const int4 *arena = reinterpret_cast<const int4 *>(A.lane_cells);

  // calls: forward then reverse (ncalls 2), or one stand-alone call of either direction, or none
  Plan F, R;
  bool hasF = false, hasR = false;
  int nF = 0, nR = 0;
  if (ncalls >= 1)
    { const LaneCall c0 = u.call[0];
      if (c0.dir > 0) { make_plan(F, c0, arena, which, TS, 0); hasF = true; nF = F.pairs; }
      else            { make_plan(R, c0, arena, which, TS, 0); hasR = true; nR = R.pairs; }
    }
  if (ncalls >= 2)
    { const LaneCall c1 = u.call[1];
      make_plan(R, c1, arena, which, TS, 2 * nF); hasR = true; nR = R.pairs;
    }
  const int T = nF + nR, tl = 2 * T;
  AlnRec &r = A.alns[i];
  const long long to = (long long) atomicAdd(A.trace_top, (unsigned long long) tl);
  int err = 0;
  if (to + tl > A.trace_cap) err = UERR_POOL;
  uint32_t *const out = reinterpret_cast<uint32_t *>(A.traces + to);  // `to` is even: every length is
  const bool flip = (which == 0) && acomp;              // align.c:1858-1884 mirrors the A trace only
#define POS(p) (flip ? T - 1 - (p) : (p))

  if (!err && hasF)                                     // pairs nR .. nR+nF-1, pair j-1 from cells j-1, j
    { if (F.tail == T_PUSH)
        out[POS(nR + F.n)] = mkpair(F.trimd - F.eL, F.Q - F.aL);
      int idx = F.head;
      Cell cur = load_cell(F.cells, idx);
      for (int j = F.n; j >= 1; j--)
        { const int pi = cur.ptr;
          if (pi < 0) { err = UERR_TRACE; break; }
          const Cell prev = load_cell(F.cells, pi);
          const int aj = cur.mark - F.sk * cur.diag;
          const int ap = (j == 1) ? F.a0 : prev.mark - F.sk * prev.diag;
          const int dp = (j == 1) ? 0 : prev.diff;
          int lo = cur.diff - dp, hi = aj - ap;
          if (j == F.n && F.tail == T_ADJ) { lo += F.trimd - F.eL; hi += F.Q - F.aL; }
          out[POS(nR + j - 1)] = mkpair(lo, hi);
          cur = prev; idx = pi;
        }
      if (!err && idx != F.fi) err = UERR_TRACE;
    }
  if (!err && hasR)                                     // push q sits at pair nR - q
    { int add_lo = 0, add_hi = 0;                       // what the call adds to the pair after its own
      if (R.virt)
        { const int lo = R.trimd, hi = R.a0 - R.Q;
          if (R.special == T_PUSH) out[POS(nR - 1)] = mkpair(which ? hi : lo, hi);   // sic, align.c:1670-1671 (H3)
          else                     { add_lo += lo; add_hi += hi; }
        }
      else
        { if (R.tail == T_PUSH)
            out[POS(nR - (R.base + R.n + 1))] = mkpair(R.trimd - R.eL, R.aL - R.Q);
          else if (R.tail == T_ADJ && R.base + R.n == 0)
            { add_lo += R.trimd - R.eL; add_hi += R.aL - R.Q; }
          int idx = R.head;
          Cell cur = load_cell(R.cells, idx);
          for (int j = R.n; j >= 1; j--)
            { const int pi = cur.ptr;
              if (pi < 0) { err = UERR_TRACE; break; }
              const Cell prev = load_cell(R.cells, pi);
              const int aj = cur.mark - R.sk * cur.diag, ap = prev.mark - R.sk * prev.diag;
              int lo = cur.diff - prev.diff, hi = ap - aj;
              const int q = R.base + j;
              if (q == 0)
                { add_lo += lo; add_hi += hi; }
              else
                { if (j == 1 && R.special == T_PUSH && which) lo = hi;               // sic (H3)
                  if (j == R.n && R.tail == T_ADJ) { lo += R.trimd - R.eL; hi += R.aL - R.Q; }
                  out[POS(nR - q)] = mkpair(lo, hi);
                }
              cur = prev; idx = pi;
            }
          if (!err && idx != R.fi) err = UERR_TRACE;
        }
      if (!err && (add_lo | add_hi) != 0 && nF > 0)     // stretches the forward call's first pair
        { const uint32_t v = out[POS(nR)];
          out[POS(nR)] = mkpair((int) (v & 0xffff) + add_lo, (int) (v >> 16) + add_hi);
        }
    }
#undef POS
  if (err)
    { if (atomicExch(&A.jobs[job].status, err) == 0)
        { atomicAdd(A.nfailed, 1);
          // the warp kernel re-runs the job and counts its alignments, waves and cells again: what the duo
          // kernel counted for it leaves the statistics (one thread per job gets here)
          const AlignJob &jb = A.jobs[job];
          atomicAdd(&A.stats[0], 0ull - (unsigned long long) jb.nalign);
          atomicAdd(&A.stats[1], 0ull - (unsigned long long) jb.nwaves);
          atomicAdd(&A.stats[2], 0ull - (unsigned long long) jb.ncells);
        }
      return;
    }
  if (which == 0) { r.a[5] = tl; r.atrace = to; }
  else            { r.b[5] = tl; r.btrace = to; }
}

void launch_unwind(const AlignArgs &A, int max_alns, cudaStream_t stream)
{ if (max_alns <= 0) return;
  const long long nthreads = (long long) max_alns * (A.do_b ? 2 : 1);
  LAUNCH(k_unwind, (int) ((nthreads + 127) / 128), 128, 0, stream, A);
}

}  // namespace damgpu
