// Unwinding of Pebble chains into trace pairs (align.c:900-1007 / 1554-1717), one thread per kept
// alignment.  The duo kernel (align_duo.cu) leaves the cells of every wave call in the job's arena
// and records per alignment which calls make up its traces (LaneUnwind); pointer chasing through
// a few hundred cells is latency, so it runs here with one thread per alignment instead of
// stalling a warp of the wave kernel.
#include "common.cuh"
#include "mapper.cuh"
#include "align.cuh"

namespace damgpu {

namespace {
enum { UERR_TRACE = 3, UERR_POOL = 13 };
struct __align__(16) UPebble { int ptr, diag, diff, mark; };         // align.c:344-349
}  // namespace

struct UPath { int tlen; uint16_t *trace; };

// one wave call; the A chain always, the B chain with dob.  Returns 0 or UERR_TRACE.
__device__ int unwind_call(const LaneCall &cc, UPebble *cells, int TS, int dob, UPath &apath, UPath &bpath,
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
          if (atrace + atlen + 2 > ahi) { err = UERR_TRACE; break; }
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
              if (atrace + atlen - 4 < alo) { err = UERR_TRACE; break; }
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
              if (btrace + btlen + 2 > bhi) { err = UERR_TRACE; break; }
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
                  if (btrace + btlen - 4 < blo) { err = UERR_TRACE; break; }
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
  UPebble *arena = reinterpret_cast<UPebble *>(A.lane_cells);
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
    { if (atomicExch(&A.jobs[u.job].status, UERR_POOL) == 0)
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

void launch_unwind(const AlignArgs &A, int max_alns, cudaStream_t stream)
{ if (max_alns <= 0) return;
  LAUNCH(k_unwind, (max_alns + 127) / 128, 128, 0, stream, A);
}

}  // namespace damgpu
