// Reporter phase (map.c:3227-3319): candidate chains -> wave alignments (align.cu) -> per read
// redundancy removal / fusion (Entwine, Fusion, Handle_Redundancies map.c:1953-2268), group
// sorts (:2304-2341), alignment-link DP (:2630-2710), chain selection and flags (:2712-2816),
// repeat-profile finalisation (:2835-2845).  One thread per read for the second half: it is a
// few hundred integer operations per read; the records and traces it emits are copied to the
// host, which only writes the .las files (align.c:3115-3122).
#include <math.h>
#include <string.h>
#include <algorithm>
#include "common.cuh"
#include "mapper.cuh"
#include "align.cuh"
#include "report.cuh"

namespace damgpu {

constexpr double CHAIN_OFF = 500., CHAIN_OVL = 400., CHAIN_PLAY = 1.4, DIFF_SCORE = 2.3;   // map.c:42-47
constexpr int    TIE_SCORE = 50, TIE_GAP = 500;                                            // map.c:48-49
constexpr uint32_t COMP_FLAG = 0x1, START_FLAG = 0x4, NEXT_FLAG = 0x8, BEST_FLAG = 0x10;  // align.h:127-135
constexpr long long FUS_BASE = 1ll << 40;       // trace offsets >= FUS_BASE live in the fusion buffer

struct OPath { long long trace; int tlen, diffs, abpos, bbpos, aepos, bepos; };
struct Ovl   { OPath path; uint32_t flags; int aread, bread, pad; };
struct Links { int score, link, mark; };
struct Zones { int beg, end, top; };

struct ReportArgs
{ int              nreads, tfirst, spacing, do_a, do_b, small, profile;
  double           best_tie;
  const int       *head;                  // per read candidate list
  const Candidate *cand;
  const int64_t   *job_off;               // per read: first job (jobs are in read, list order)
  const AlignJob  *jobs;
  const AlnRec    *alns;
  uint16_t        *traces;                // alignment traces (align.cu)
  uint16_t        *ftraces;               // fusion buffer
  const int64_t   *fus_off;               // per read offset/capacity into ftraces (nreads+1)
  const int64_t   *ovl_off;               // per read offset into the Ovl/Links/Zones scratch
  const int32_t   *rlen;
  Ovl             *amatch, *bmatch, *tmp;
  Links           *linker;
  int             *perm;
  Zones           *part;
  uint8_t         *out_a, *out_b;         // per-read output regions
  const int64_t   *outa_off, *outb_off;   // nreads+1
  int64_t         *used_a, *used_b;       // bytes written per read
  int             *nrec_a, *nrec_b;
  const int16_t   *cover;
  const int64_t   *coff;
  uint8_t         *prof;
  double           spow[41];
  int             *error;                 // 1 = fusion overflow, 2 = trace value > 255, 3 = output overflow
  unsigned long long *h2_events;
};

__device__ __forceinline__ uint16_t *tptr(const ReportArgs &R, long long off)
{ return (off >= FUS_BASE) ? R.ftraces + (off - FUS_BASE) : R.traces + off; }

__device__ __forceinline__ int iabs(int x) { return x < 0 ? -x : x; }

// Entwine, map.c:1953-2058
__device__ int entwine(const ReportArgs &R, const OPath *jpath, const OPath *kpath, int *where)
{ const int S = R.spacing;
  const uint16_t *ktrace = tptr(R, kpath->trace), *jtrace = tptr(R, jpath->trace);
  int ac, b2, y2, ae, i, j, k, den = 0, min = 10000;

  y2 = jpath->bbpos; j = jpath->abpos / S;
  b2 = kpath->bbpos; k = kpath->abpos / S;
  if (jpath->abpos == kpath->abpos)
    { min = iabs(y2 - b2);
      if (min == 0) *where = kpath->abpos;
    }
  if (j < k)
    { ac = k * S; j = 1 + 2 * (k - j); k = 1;
      for (i = 1; i < j; i += 2) y2 += jtrace[i];
    }
  else
    { ac = j * S; k = 1 + 2 * (j - k); j = 1;
      for (i = 1; i < k; i += 2) b2 += ktrace[i];
    }
  ae = jpath->aepos;
  if (ae > kpath->aepos) ae = kpath->aepos;
  while (1)
    { ac += S;
      if (ac >= ae) break;
      y2 += jtrace[j]; b2 += ktrace[k];
      j += 2; k += 2;
      i = iabs(y2 - b2);
      if (i <= min)
        { min = i;
          if (i == 0) *where = ac;
        }
      den += 1;
    }
  if (jpath->aepos == kpath->aepos)
    { i = iabs(jpath->bepos - kpath->bepos);
      if (i <= min)
        { min = i;
          if (i == 0) *where = kpath->aepos;
        }
    }
  return (den == 0) ? -1 : min;
}

// Fusion, map.c:2065-2109; appends to the read's slice of the fusion buffer
__device__ void fusion(const ReportArgs &R, OPath *path1, int ap, const OPath *path2,
                       long long *ftop, long long fend)
{ const int S = R.spacing;
  const int k1 = 2 * ((ap / S) - (path1->abpos / S));
  const int k2 = 2 * ((ap / S) - (path2->abpos / S));
  int len = k1 + (path2->tlen - k2), diff = 0, k;
  if (*ftop + len > fend)
    { *R.error = 1;
      return;
    }
  uint16_t *trace = R.ftraces + *ftop;
  const long long at = *ftop;
  *ftop += len;
  len = 0;
  if (k1 > 0)
    { const uint16_t *t = tptr(R, path1->trace);
      for (k = 0; k < k1; k += 2)
        { trace[len++] = t[k]; trace[len++] = t[k + 1]; diff += t[k]; }
    }
  if (k2 < path2->tlen)
    { const uint16_t *t = tptr(R, path2->trace);
      for (k = k2; k < path2->tlen; k += 2)
        { trace[len++] = t[k]; trace[len++] = t[k + 1]; diff += t[k]; }
    }
  path1->aepos = path2->aepos;
  path1->bepos = path2->bepos;
  path1->diffs = diff;
  path1->trace = FUS_BASE + at;
  path1->tlen  = len;
}

// Handle_Redundancies, map.c:2116-2268
__device__ int handle_redundancies(const ReportArgs &R, Ovl *amatch, int novls, Ovl *bmatch, int cm,
                                   long long *ftop, long long fend)
{ const bool hasB = (bmatch != nullptr);
  int awhen = 0, bwhen = 0, dist;
  for (int j = 1; j < novls; j++)
    { OPath *jpath = &amatch[j].path, *jmath = hasB ? &bmatch[j].path : nullptr;
      for (int k = j - 1; k >= 0; k--)
        { OPath *kpath = &amatch[k].path, *kmath = hasB ? &bmatch[k].path : nullptr;
          if (kpath->abpos < 0)
            continue;
          if (jpath->abpos < kpath->abpos)
            { if (kpath->abpos <= jpath->aepos && kpath->bbpos <= jpath->bepos)
                { dist = entwine(R, jpath, kpath, &awhen);
                  if (dist == 0)
                    { if (kpath->aepos > jpath->aepos)
                        { if (hasB)
                            { if (cm)
                                { dist = entwine(R, kmath, jmath, &bwhen);
                                  if (dist != 0) continue;
                                  fusion(R, jpath, awhen, kpath, ftop, fend);
                                  fusion(R, kmath, bwhen, jmath, ftop, fend);
                                  *jmath = *kmath;
                                }
                              else
                                { dist = entwine(R, jmath, kmath, &bwhen);
                                  if (dist != 0) continue;
                                  fusion(R, jpath, awhen, kpath, ftop, fend);
                                  fusion(R, jmath, bwhen, kmath, ftop, fend);
                                }
                            }
                          else
                            fusion(R, jpath, awhen, kpath, ftop, fend);
                        }
                      kpath->abpos = -1;
                      break;
                    }
                }
            }
          else
            { if (jpath->abpos <= kpath->aepos && jpath->bbpos <= kpath->bepos)
                { dist = entwine(R, kpath, jpath, &awhen);
                  if (dist == 0)
                    { if (kpath->abpos == jpath->abpos)
                        { if (kpath->aepos > jpath->aepos)
                            { *jpath = *kpath;
                              if (hasB) *jmath = *kmath;
                            }
                        }
                      else if (jpath->aepos > kpath->aepos)
                        { if (hasB)
                            { if (cm)
                                { dist = entwine(R, jmath, kmath, &bwhen);
                                  if (dist != 0) continue;
                                  fusion(R, kpath, awhen, jpath, ftop, fend);
                                  *jpath = *kpath;
                                  fusion(R, jmath, bwhen, kmath, ftop, fend);
                                }
                              else
                                { dist = entwine(R, kmath, jmath, &bwhen);
                                  if (dist != 0) continue;
                                  fusion(R, kpath, awhen, jpath, ftop, fend);
                                  *jpath = *kpath;
                                  fusion(R, kmath, bwhen, jmath, ftop, fend);
                                  *jmath = *kmath;
                                }
                            }
                          else
                            { fusion(R, kpath, awhen, jpath, ftop, fend);
                              *jpath = *kpath;
                            }
                        }
                      else
                        { *jpath = *kpath;
                          if (hasB) *jmath = *kmath;
                        }
                      kpath->abpos = -1;
                      break;
                    }
                }
            }
        }
    }
  int no = 0;
  for (int j = 0; j < novls; j++)
    if (amatch[j].path.abpos >= 0)
      { if (hasB) bmatch[no] = bmatch[j];
        amatch[no++] = amatch[j];
      }
  return no;
}

// The group sorts AMATCH/BN_MATCH/BC_MATCH_SORT (map.c:2304-2341) under glibc's indirect merge
// sort: key order, equal keys in descending original index.
__device__ void group_sort(Ovl *v, int n, int which, Ovl *t)
{ for (int i = 0; i < n; i++)
    t[i] = v[n - 1 - i];
  for (int i = 1; i < n; i++)
    { const Ovl x = t[i];
      int j;
      for (j = i - 1; j >= 0; j--)
        { bool before;
          if (which == 0)      before = (x.path.abpos > t[j].path.abpos);
          else if (which == 1) before = (x.path.bbpos > t[j].path.bbpos);
          else                 before = (x.path.bepos < t[j].path.bepos);
          if (!before) break;
          t[j + 1] = t[j];
        }
      t[j + 1] = x;
    }
  for (int i = 0; i < n; i++)
    v[i] = t[i];
}

// special_log, map.c:2270-2302 (the pow table comes from the host's libm)
__device__ int special_log(const ReportArgs &R, int cover)
{ if (cover <= 1) return cover;
  if (cover >= 10000) return 40;
  const double x = cover;
  int l = 0, r = 41;
  while (l < r)
    { const int m = (l + r) >> 1;
      if (R.spow[m] <= x) l = m + 1; else r = m;
    }
  return l - 1;
}

// Write_Overlap (align.c:3115-3122) + Compress_TraceTo8 (:3124-3142) into the read's output slice
__device__ void emit(const ReportArgs &R, uint8_t *out, int64_t *used, int64_t cap, int *nrec,
                     const Ovl *o)
{ const int tb = R.small ? 1 : 2;
  const int64_t need = 40 + (int64_t) o->path.tlen * tb;
  if (*used + need > cap)
    { *R.error = 3;
      return;
    }
  uint8_t *p = out + *used;
  const int32_t w[10] = { o->path.tlen, o->path.diffs, o->path.abpos, o->path.bbpos, o->path.aepos,
                          o->path.bepos, (int32_t) o->flags, o->aread, o->bread, 0 };
  for (int i = 0; i < 10; i++)
    { const uint32_t v = (uint32_t) w[i];
      p[4 * i] = (uint8_t) v; p[4 * i + 1] = (uint8_t) (v >> 8);
      p[4 * i + 2] = (uint8_t) (v >> 16); p[4 * i + 3] = (uint8_t) (v >> 24);
    }
  p += 40;
  const uint16_t *t = tptr(R, o->path.trace);
  if (R.small)
    for (int j = 0; j < o->path.tlen; j++)
      { if (t[j] > 255) *R.error = 2;                   // H9: the reference exits here
        p[j] = (uint8_t) t[j];
      }
  else
    for (int j = 0; j < o->path.tlen; j++)
      { p[2 * j] = (uint8_t) t[j]; p[2 * j + 1] = (uint8_t) (t[j] >> 8); }
  *used += need;
  *nrec += 1;
}

__global__ void __launch_bounds__(64) k_report(ReportArgs R)
{ const int ar = blockIdx.x * blockDim.x + threadIdx.x;
  if (ar >= R.nreads) return;

  const int64_t o0 = R.ovl_off[ar];
  Ovl   *amatch = R.amatch + o0, *bmatch = R.do_b ? R.bmatch + o0 : nullptr, *tmp = R.tmp + o0;
  Links *linker = R.linker + o0;
  int   *perm = R.perm + o0;
  Zones *part = R.part + o0;
  long long ftop = R.fus_off[ar];
  const long long fend = R.fus_off[ar + 1];
  uint8_t *outa = R.out_a + R.outa_off[ar], *outb = R.do_b ? R.out_b + R.outb_off[ar] : nullptr;
  const int64_t capa = R.outa_off[ar + 1] - R.outa_off[ar];
  const int64_t capb = R.do_b ? R.outb_off[ar + 1] - R.outb_off[ar] : 0;
  int64_t useda = 0, usedb = 0;
  int nreca = 0, nrecb = 0;

  const int alen = R.rlen[ar];
  const int atck = (alen - 1) / R.spacing + 1;
  int novl = 0, lovl = 0;

  // gather the alignments of the read's candidates in list order (map.c:2460-2610)
  { int64_t j = R.job_off[ar];
    for (int c = R.head[ar]; c >= 0; j++)
      { const AlignJob &job = R.jobs[j];
        const int br = R.cand[c].bread, cm = R.cand[c].comp;
        for (int a = job.first; a >= 0; a = R.alns[a].next)
          { const AlnRec &r = R.alns[a];
            Ovl &A = amatch[novl];
            A.aread = ar + R.tfirst; A.bread = br; A.flags = cm ? COMP_FLAG : 0; A.pad = 0;
            A.path.abpos = r.a[0]; A.path.bbpos = r.a[1]; A.path.aepos = r.a[2]; A.path.bepos = r.a[3];
            A.path.diffs = r.a[4]; A.path.tlen = r.a[5]; A.path.trace = r.atrace;
            if (R.do_b)
              { Ovl &B = bmatch[novl];
                B.aread = br; B.bread = ar + R.tfirst; B.flags = cm ? COMP_FLAG : 0; B.pad = 0;
                B.path.abpos = r.b[0]; B.path.bbpos = r.b[1]; B.path.aepos = r.b[2]; B.path.bepos = r.b[3];
                B.path.diffs = r.b[4]; B.path.tlen = r.b[5]; B.path.trace = r.btrace;
              }
            novl += 1;
          }
        const int d = R.cand[c].next;
        if (d < 0 || R.cand[d].bread != br || R.cand[d].comp != cm)
          { if (novl - lovl > 1)
              novl = lovl + handle_redundancies(R, amatch + lovl, novl - lovl,
                                                R.do_b ? bmatch + lovl : nullptr, cm, &ftop, fend);
            if (novl - lovl > 1)
              { group_sort(amatch + lovl, novl - lovl, 0, tmp);
                if (R.do_b)
                  group_sort(bmatch + lovl, novl - lovl, cm ? 2 : 1, tmp);
              }
            lovl = novl;
          }
        c = d;
      }
  }

  if (novl > 0)
    { // link DP, map.c:2630-2710
      int br;
      lovl = 0;
      linker[0].link = -1;
      linker[0].score = (int) __dsub_rn((double) (amatch[0].path.aepos - amatch[0].path.abpos),
                                        __dmul_rn(DIFF_SCORE, (double) amatch[0].path.diffs));
      linker[0].mark = 1;
      perm[0] = 0;
      br = amatch[0].bread;
      for (int c = 1; c < novl; c++)
        { const OPath *cpath = &amatch[c].path;
          linker[c].link = -1;
          linker[c].score = (int) __dsub_rn((double) (cpath->aepos - cpath->abpos),
                                            __dmul_rn(DIFF_SCORE, (double) cpath->diffs));
          linker[c].mark = 1;
          perm[c] = c;
          if (amatch[c].bread != br)
            { br = amatch[c].bread;
              lovl = c;
              continue;
            }
          const uint32_t cor = amatch[c].flags & COMP_FLAG;
          for (int d = c - 1; d >= lovl; d--)
            { const uint32_t dor = amatch[d].flags & COMP_FLAG;
              if (dor != cor) continue;
              const OPath *dpath = &amatch[d].path;
              if (dor) { if (dpath->bepos < cpath->bepos) continue; }
              else     { if (dpath->bbpos < cpath->bbpos) continue; }
              if ((double) dpath->abpos <= __dsub_rn((double) cpath->aepos, CHAIN_OVL) ||
                  (double) dpath->bbpos <= __dsub_rn((double) cpath->bepos, CHAIN_OVL))
                continue;
              const double rat = __ddiv_rn(__dadd_rn((double) (dpath->abpos - cpath->aepos), CHAIN_OFF),
                                           __dadd_rn((double) (dpath->bbpos - cpath->bepos), CHAIN_OFF));
              if (1. > __dmul_rn(rat, CHAIN_PLAY) || rat > CHAIN_PLAY)
                continue;
              const int scr = (int) __dsub_rn((double) (linker[d].score + (cpath->aepos - cpath->abpos)),
                                              __dmul_rn(DIFF_SCORE, (double) cpath->diffs));
              const int scr2 = linker[c].score;
              if (scr < scr2 - TIE_SCORE)
                continue;
              if (scr <= scr2 + TIE_SCORE)
                { const int gap = dpath->abpos - cpath->aepos;
                  int gap2;
                  if (linker[d].link >= 0)
                    { if (linker[c].link < 0)             // H2: the reference reads amatch[-1] here
                        { atomicAdd(R.h2_events, 1ull);
                          gap2 = 0;
                        }
                      else
                        gap2 = amatch[linker[c].link].path.aepos - dpath->abpos;
                    }
                  else
                    gap2 = 0;
                  if (gap > gap2 + TIE_GAP)
                    continue;
                  if (gap >= gap2 - TIE_GAP)
                    { if (scr < scr2) continue;
                      if (scr == scr2 && gap >= gap2) continue;
                    }
                }
              linker[c].link = d;
              linker[c].score = scr;
              linker[d].mark = 0;
            }
        }

      // LINK_SORT (map.c:2355-2360,2712): score descending, stable
      for (int i = 1; i < novl; i++)
        { const int x = perm[i];
          int j;
          for (j = i - 1; j >= 0 && linker[perm[j]].score < linker[x].score; j--)
            perm[j + 1] = perm[j];
          perm[j + 1] = x;
        }

      // selection, map.c:2714-2816
      int nparts = 0;
      for (int c = 0; c < novl && linker[perm[c]].score >= 0; c++)
        if (linker[perm[c]].mark == 1)
          { int p, b, e, q, n, best;
            b = e = perm[c];
            for (p = linker[b].link; p >= 0 && linker[p].mark >= 0; p = linker[p].link)
              e = p;
            for (p = 0; p < nparts; p++)
              if (amatch[b].path.abpos < part[p].end - 100 && amatch[e].path.aepos > part[p].beg + 100)
                break;
            if (p >= nparts)
              { part[p].beg = amatch[b].path.abpos;
                part[p].end = amatch[e].path.aepos;
                part[p].top = linker[b].score;
                best = 1;
                nparts += 1;
              }
            else
              { if ((double) linker[b].score < __dmul_rn(R.best_tie, (double) part[p].top))
                  continue;
                best = (linker[b].score == part[p].top);
              }

            q = -1;
            for (p = b; 1; p = n)
              { linker[p].mark = -1;
                if (R.do_a)
                  { if (p == b)
                      { amatch[p].flags |= START_FLAG;
                        if (best) amatch[p].flags |= BEST_FLAG;
                      }
                    else
                      amatch[p].flags |= NEXT_FLAG;
                    emit(R, outa, &useda, capa, &nreca, amatch + p);
                  }
                n = linker[p].link;
                if (R.do_b)
                  { if (bmatch[p].flags & COMP_FLAG)
                      { linker[p].link = q;
                        q = p;
                      }
                    else
                      { if (p == b)
                          { bmatch[p].flags |= START_FLAG;
                            if (best) bmatch[p].flags |= BEST_FLAG;
                          }
                        else
                          bmatch[p].flags |= NEXT_FLAG;
                        emit(R, outb, &usedb, capb, &nrecb, bmatch + p);
                      }
                  }
                if (p == e)
                  break;
              }
            if (R.do_b && (bmatch[b].flags & COMP_FLAG))
              { e = b;
                b = q;
                for (p = b; 1; p = linker[p].link)
                  { if (p == b)
                      { bmatch[p].flags |= START_FLAG;
                        if (best) bmatch[p].flags |= BEST_FLAG;
                      }
                    else
                      bmatch[p].flags |= NEXT_FLAG;
                    emit(R, outb, &usedb, capb, &nrecb, bmatch + p);
                    if (p == e)
                      break;
                  }
              }
          }
    }

  R.used_a[ar] = useda; R.nrec_a[ar] = nreca;
  if (R.do_b) { R.used_b[ar] = usedb; R.nrec_b[ar] = nrecb; }

  if (R.profile)                                        // map.c:2835-2845
    { const int16_t *cnt = R.cover + R.coff[ar];
      uint8_t *log = R.prof + R.coff[ar];
      int c = 0;
      for (int i = 0; i <= atck; i++)
        { c += cnt[i];
          log[i] = (uint8_t) special_log(R, c);
        }
    }
}

// ---- 2-bit block images for the warp alignment kernel -----------------------------------------
// Word j of the packed image holds bytes 16j .. 16j+15 of the byte image, first base in the low
// bits; `out` and `bases` are both the origin of the image, j runs over [first, first + nwords).
__device__ __forceinline__ uint32_t pack4(uint32_t w)
{ uint32_t x = w & 0x03030303u;
  x = (x | (x >> 6)) & 0x000f000fu;
  return (x | (x >> 12)) & 0xffu;
}

__global__ void __launch_bounds__(256)
k_pack2bit(const uint8_t *__restrict__ bases, long long first, long long nwords, uint32_t *__restrict__ out)
{ const long long stride = (long long) gridDim.x * blockDim.x;
  for (long long j = first + (long long) blockIdx.x * blockDim.x + threadIdx.x; j < first + nwords; j += stride)
    { const uint4 v = *reinterpret_cast<const uint4 *>(bases + 16 * j);
      out[j] = pack4(v.x) | (pack4(v.y) << 8) | (pack4(v.z) << 16) | (pack4(v.w) << 24);
    }
}

constexpr long long PK_LEAD = 8;                        // words readable before the image origin

static uint32_t *pack_image(const uint8_t *bases, int64_t total, cudaStream_t stream)
{ const long long nwords = total / 16 + 8 + PK_LEAD;    // BLOCK_SLACK covers the bytes this reads
  uint32_t *raw = dalloc<uint32_t>((size_t) nwords + 8);
  long long blocks = (nwords + 255) / 256;
  if (blocks > (long long) sm_count() * 16) blocks = (long long) sm_count() * 16;
  LAUNCH(k_pack2bit, (int) blocks, 256, 0, stream, bases, -PK_LEAD, nwords, raw + PK_LEAD);
  return raw;
}

// ---- job enumeration ----------------------------------------------------------------------
__global__ void k_count_cands(const int *head, const Candidate *cand, int nreads, int *cnt)
{ const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nreads) return;
  int n = 0;
  for (int c = head[r]; c >= 0; c = cand[c].next) n++;
  cnt[r] = n;
}

__global__ void k_fill_jobs(const int *head, const Candidate *cand, int nreads,
                            const int64_t *job_off, AlignJob *jobs)
{ const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nreads) return;
  int64_t j = job_off[r];
  for (int c = head[r]; c >= 0; c = cand[c].next, j++)
    { AlignJob jb; jb.read = r; jb.cand = c; jb.first = -1; jb.count = 0; jb.status = 0;
      jb.nalign = jb.nwaves = jb.ncells = 0;
      jobs[j] = jb;
    }
}

// per read: number of alignments and total A/B trace lengths after the alignment phase
__global__ void k_read_totals(int nreads, const int64_t *job_off, const AlignJob *jobs,
                              const AlnRec *alns, int *novl, int64_t *alen_sum, int64_t *blen_sum)
{ const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nreads) return;
  int n = 0; int64_t sa = 0, sb = 0;
  for (int64_t j = job_off[r]; j < job_off[r + 1]; j++)
    for (int a = jobs[j].first; a >= 0; a = alns[a].next)
      { n++; sa += alns[a].a[5]; sb += alns[a].b[5]; }
  novl[r] = n; alen_sum[r] = sa; blen_sum[r] = sb;
}

__global__ void k_collect_failed(const AlignJob *jobs, int njobs, int *list, int *n)
{ const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= njobs) return;
  if (jobs[j].status != 0)
    list[atomicAdd(n, 1)] = j;
}

// New_Align_Spec tables, align.c:207-269 (host; identical double expressions)
static const double Bias_Factor[10] = { .690, .690, .690, .690, .780, .850, .900, .933, .966, 1.000 };

static void set_table(int bit, int prefix, int score, int max, int mscore, int dscore,
                      int16_t *table, int16_t *tscore)
{ if (bit >= 15)
    { table[prefix]  = (int16_t) (score - max);
      tscore[prefix] = (int16_t) score;
    }
  else
    { if (score > max) max = score;
      set_table(bit + 1, (prefix << 1), score - dscore, max, mscore, dscore, table, tscore);
      set_table(bit + 1, (prefix << 1) | 1, score + mscore, max, mscore, dscore, table, tscore);
    }
}

void build_align_spec(double ave_corr, const float freq[4], int *ave_path, int16_t *score,
                      int16_t *table)
{ double match = freq[0] + freq[3];
  if ((match <= 0.) == (match > 0.)) match = .5;
  if (match > .5) match = 1. - match;
  int bias = (int) ((match + .025) * 20. - 1.);
  if (match < .2)
    { fprintf(stderr, "Warning: Base bias worse than 80/20%% ! (New_Align_Spec)\n");
      fprintf(stderr, "         Capping bias at this ratio.\n");
      bias = 3;
    }
  *ave_path = (int) (60 * (1. - Bias_Factor[bias] * (1. - ave_corr)));
  const int mscore = (int) (1000 * Bias_Factor[bias] * (1. - ave_corr));
  const int dscore = 1000 - mscore;
  set_table(0, 0, 0, 0, mscore, dscore, table, score);
}

// sizes of read r's slices in the second half of the Reporter: Ovl records, fusion trace scratch
// (map.c:2129-2268 can fuse novl-1 pairs, each up to 5*(rlen/S)+16 values), output bytes of the
// M and R families (40-byte records + trace bytes, fused traces included)
__global__ void __launch_bounds__(256)
k_read_sizes(int n, int S, int tb, const int32_t *__restrict__ rlen, const int *__restrict__ novl,
             const int64_t *__restrict__ asum, const int64_t *__restrict__ bsum, int64_t *__restrict__ sz)
{ const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int64_t no = novl[r];
  const int64_t fus = (no > 1) ? (no - 1) * (5 * (rlen[r] / S) + 16) : 0;
  sz[r] = no;
  sz[(size_t) (n + 1) + r] = fus;
  sz[2 * (size_t) (n + 1) + r] = 40ll * no + tb * (asum[r] + fus);
  sz[3 * (size_t) (n + 1) + r] = 40ll * no + tb * (bsum[r] + fus);
}

__global__ void k_gather_totals(int n, const int64_t *a, const int64_t *b, const int64_t *c, const int64_t *d,
                                int64_t *tot)
{ if (threadIdx.x == 0)
    { tot[0] = a[n]; tot[1] = b[n]; tot[2] = c[n]; tot[3] = d[n]; }
}

// exclusive scan of used[0..n) into off[0..n] (single CTA; n = reads of the block)
__global__ void __launch_bounds__(1024) k_scan_used(const int64_t *__restrict__ used, int n, int64_t *off)
{ __shared__ int64_t part[1024];
  const int t = threadIdx.x;
  const int per = (n + 1023) / 1024;
  const int lo = t * per, hi = (lo + per < n) ? lo + per : n;
  int64_t s = 0;
  for (int i = lo; i < hi; i++) s += used[i];
  part[t] = s;
  __syncthreads();
  if (t == 0)
    { int64_t run = 0;
      for (int i = 0; i < 1024; i++)
        { const int64_t c = part[i]; part[i] = run; run += c; }
      off[n] = run;
    }
  __syncthreads();
  int64_t run = part[t];
  for (int i = lo; i < hi; i++)
    { off[i] = run; run += used[i]; }
}

// per-read output slots (worst-case sized) -> dense stream in read order, one warp per read
__global__ void __launch_bounds__(256)
k_compact_out(const uint8_t *__restrict__ src, const int64_t *__restrict__ src_off,
              const int64_t *__restrict__ used, const int64_t *__restrict__ dst_off, int n,
              uint8_t *__restrict__ dst)
{ const int lane = threadIdx.x & 31;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += nw)
    { const uint8_t *s = src + src_off[r];
      uint8_t *d = dst + dst_off[r];
      const int64_t len = used[r];
      for (int64_t i = lane; i < len; i += 32)
        d[i] = s[i];
    }
}

// Check_Trace_Points (align.c:3194-3236) on every record of the dense stream, one thread per read:
// the number of trace points must match the A interval and the B advances must add up to the B
// interval.  Counts the records that fail (the LAcheck pass of
// HPC.damapper, at full speed on the device).
__global__ void __launch_bounds__(256)
k_check_trace(const uint8_t *__restrict__ dense, const int64_t *__restrict__ off,
              const int *__restrict__ nrec, int n, int tspace, int tbytes, unsigned long long *fails)
{ const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const uint8_t *p = dense + off[r];
  int bad = 0;
  for (int k = nrec[r]; k > 0; k--)
    { int h[6];                                         // tlen, diffs, abpos, bbpos, aepos, bepos
      for (int j = 0; j < 6; j++)
        h[j] = (int) ((uint32_t) p[4 * j] | ((uint32_t) p[4 * j + 1] << 8) |
                      ((uint32_t) p[4 * j + 2] << 16) | ((uint32_t) p[4 * j + 3] << 24));
      const int tlen = h[0];
      const uint8_t *t = p + 40;
      if (((h[4] - 1) / tspace - h[2] / tspace) * 2 != tlen - 2)
        bad++;
      else
        { int b = h[3];
          for (int i = 1; i < tlen; i += 2)
            b += (tbytes == 1) ? (int) t[i] : (int) ((uint32_t) t[2 * i] | ((uint32_t) t[2 * i + 1] << 8));
          if (b != h[5])
            bad++;
        }
      p = t + (int64_t) tlen * tbytes;
    }
  if (bad)
    atomicAdd(fails, (unsigned long long) bad);
}

__global__ void k_flip_byte(uint8_t *p) { *p = (uint8_t) (*p + 1); }   // DAMGPU_TEST_CORRUPT_TRACE (tests only)

template <typename T> static std::vector<T> d2h(const T *d, size_t n)
{ std::vector<T> v(n);
  if (n) CUDA_CHECK(cudaMemcpy(v.data(), d, sizeof(T) * n, cudaMemcpyDeviceToHost));
  return v;
}

template <typename T> static T *h2d(const std::vector<T> &v)
{ T *d = dalloc<T>(v.size() + 1);
  if (!v.empty()) CUDA_CHECK(cudaMemcpy(d, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
  return d;
}

ReportOut *reporter(Mapper *m, const DeviceBlock *ref, double ave_corr, const float freq[4],
                    int do_a, int do_b, cudaStream_t stream)
{ ReportOut *out = new ReportOut();
  const DeviceBlock *rd = m->reads;
  const int n = rd->nreads, S = g_par.spacing;
  if (m->spacing != S)
    fatal("SPACING changed after the mapper was created");

  TRACE(nullptr);
  // alignment spec tables
  std::vector<int16_t> tables(65536);
  int ave_path = 0;
  build_align_spec(ave_corr, freq, &ave_path, tables.data(), tables.data() + 32768);
  int16_t *d_tables = h2d(tables);

  // reverse-complemented copy of the reads (complement, map.c:1940-1948, once per block)
  DeviceBlock rc = *rd;
  rc.raw = dalloc<uint8_t>((size_t) rd->total + 2 * BLOCK_SLACK);
  rc.bases = rc.raw + BLOCK_SLACK;
  revcomp_copy_block(rd, rc.bases, stream);
  uint32_t *pk_a = pack_image(rd->bases, rd->total, stream);
  uint32_t *pk_ac = pack_image(rc.bases, rd->total, stream);
  uint32_t *pk_b = pack_image(ref->bases, ref->total, stream);
  chain_sync();                                          // the last chain call may still be running

  TRACE("report: spec+rc copy");
  // jobs = live candidates in (read, list order)
  int *d_cnt = dalloc<int>(n + 1);
  LAUNCH(k_count_cands, (n + 255) / 256, 256, 0, stream, m->head, m->cand, n, d_cnt);
  std::vector<int> cnt = d2h(d_cnt, n);
  { int ovf = 0;                                         // deferred check of the chain kernels
    CUDA_CHECK(cudaMemcpy(&ovf, m->overflow, sizeof(int), cudaMemcpyDeviceToHost));
    if (ovf)
      fatal("Match_Filter: candidate/jump pool overflow (internal sizing error)");
  }
  std::vector<int64_t> job_off(n + 1);
  int64_t njobs64 = 0;
  for (int i = 0; i < n; i++) { job_off[i] = njobs64; njobs64 += cnt[i]; }
  job_off[n] = njobs64;
  if (njobs64 > 0x7ffffff0ll) fatal("Reporter: too many candidates");
  const int njobs = (int) njobs64;
  int64_t  *d_job_off = h2d(job_off);
  AlignJob *d_jobs = dalloc<AlignJob>((size_t) njobs + 1);
  LAUNCH(k_fill_jobs, (n + 255) / 256, 256, 0, stream, m->head, m->cand, n, d_job_off, d_jobs);

  TRACE("report: jobs");
  // scratch + output pools of the alignment phase
  const int nblocks = std::min((njobs + ALIGN_WARPS - 1) / ALIGN_WARPS, sm_count() * 8);
  const int nwarps = std::max(1, nblocks) * ALIGN_WARPS;
  const int nbig = 64;                                 // warps of the overflow kernel
  const int cells_small = 4096, cells_big = 1 << 20;
  const int tcap = 2 * ((2 * rd->maxlen) / S + 8), tcap_big = 8 * tcap;
  const size_t cell_bytes = std::max((size_t) nwarps * cells_small, (size_t) nbig * cells_big) * 16;
  const size_t tscr = std::max((size_t) nwarps * 4 * tcap, (size_t) nbig * 4 * tcap_big);
  void          *d_cells = dalloc<unsigned char>(cell_bytes);
  uint16_t      *d_tscr = dalloc<uint16_t>(tscr);
  unsigned char *d_big = nullptr;
  int           *d_ctr = dalloc<int>(8);               // job counter, aln_top, nfailed, nlist
  unsigned long long *d_ull = dalloc<unsigned long long>(10);  // trace_top, stats[7] (1..7), trace-check failures (8)
  CUDA_CHECK(cudaMemsetAsync(d_ctr, 0, sizeof(int) * 8, stream));
  CUDA_CHECK(cudaMemsetAsync(d_ull, 0, sizeof(unsigned long long) * 10, stream));

  int       aln_cap = njobs * 2 + 1024;
  long long trace_cap = (long long) aln_cap * (4 * (rd->maxlen / S + 4));
  AlnRec   *d_alns = dalloc<AlnRec>((size_t) aln_cap);
  uint16_t *d_traces = dalloc<uint16_t>((size_t) trace_cap);

  // duo kernel (first tier): per-job Pebble arenas, unwind records
  std::vector<long long> cell_base(n + 1);
  long long ncell = 0;
  for (int i = 0; i < n; i++)
    { cell_base[i] = ncell;
      ncell += (long long) cnt[i] * LANE_ARENA(rd->h_rlen[i] / S);
    }
  cell_base[n] = ncell;
  long long  *d_cell_base = h2d(cell_base);
  void       *d_lane_cells = dalloc<unsigned char>((size_t) ncell * 16 + 16);
  LaneUnwind *d_unwind = dalloc<LaneUnwind>((size_t) aln_cap);
  int        *d_duo_win = dalloc<int>(duo_window_bytes(std::min((njobs + 7) / 8 + 1, duo_max_blocks())) / sizeof(int) + 1);
  CUDA_CHECK(cudaMemsetAsync(d_unwind, 0xff, sizeof(LaneUnwind) * (size_t) aln_cap, stream));

  AlignArgs A;
  memset(&A, 0, sizeof(A));
  A.jobs = d_jobs; A.job_list = nullptr; A.njobs = njobs; A.job_counter = d_ctr;
  A.cand = m->cand; A.jumps = m->jumps;
  A.bases_a = rd->bases; A.bases_ac = rc.bases; A.bases_b = ref->bases;
  A.pk_a = pk_a + PK_LEAD; A.pk_ac = pk_ac + PK_LEAD; A.pk_b = pk_b + PK_LEAD;
  A.boff_a = rd->boff; A.boff_b = ref->boff; A.rlen_a = rd->rlen; A.rlen_b = ref->rlen;
  A.spec.spacing = S; A.spec.ave_path = ave_path; A.spec.score = d_tables; A.spec.table = d_tables + 32768;
  A.kmer = g_par.kmer; A.do_b = do_b;
  A.big_state = nullptr; A.cells = d_cells; A.cells_small = cells_small; A.cells_big = cells_big;
  A.tscratch = d_tscr; A.tcap = tcap; A.tcap_big = tcap_big;
  A.alns = d_alns; A.aln_top = d_ctr + 1; A.aln_cap = aln_cap;
  A.traces = d_traces; A.trace_top = d_ull; A.trace_cap = trace_cap;
  A.nfailed = d_ctr + 2; A.stats = d_ull + 1;
  A.lane_cells = d_lane_cells; A.lane_cell_base = d_cell_base; A.lane_job_off = d_job_off;
  A.unwind = d_unwind; A.duo_win = d_duo_win;

  TRACE("report: alloc");
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (g_time_kernels)
    { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, stream); }
  if (njobs > 0)
    { if (g_align_tier != 0)
        { launch_align_duo(A, njobs, stream);
          launch_unwind(A, aln_cap, stream);
        }
      else
        launch_align(A, false, nblocks, stream);
    }
  if (g_time_kernels)
    { cudaEventRecord(e1, stream); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&out->ms_align, e0, e1);
      cudaEventDestroy(e0); cudaEventDestroy(e1);
    }

  TRACE("report: align kernel");
  // re-run the jobs that outgrew the fast configuration / the output pools
  int *d_list = nullptr;
  for (int round = 0; ; round++)
    { int ctr[4];
      CUDA_CHECK(cudaMemcpyAsync(ctr, d_ctr, sizeof(int) * 4, cudaMemcpyDeviceToHost, stream));
      CUDA_CHECK(cudaStreamSynchronize(stream));
      const int nfailed = ctr[2];
      if (nfailed == 0)
        break;
      if (round >= 8)
        fatal("Reporter: %d alignment jobs still fail after %d overflow rounds", nfailed, round);
      out->overflow_jobs += nfailed;
      // grow the output pools (keeps what is there)
      { const int used = std::min(ctr[1], aln_cap);
        unsigned long long tused = 0;
        CUDA_CHECK(cudaMemcpy(&tused, d_ull, sizeof(tused), cudaMemcpyDeviceToHost));
        if ((long long) tused > trace_cap) tused = (unsigned long long) trace_cap;
        const int ncap = aln_cap * 2;
        const long long tcap2 = trace_cap * 2;
        AlnRec *na = dalloc<AlnRec>((size_t) ncap);
        uint16_t *nt = dalloc<uint16_t>((size_t) tcap2);
        CUDA_CHECK(cudaMemcpy(na, d_alns, sizeof(AlnRec) * (size_t) used, cudaMemcpyDeviceToDevice));
        CUDA_CHECK(cudaMemcpy(nt, d_traces, sizeof(uint16_t) * (size_t) tused, cudaMemcpyDeviceToDevice));
        dfree(d_alns); dfree(d_traces);
        d_alns = na; d_traces = nt; aln_cap = ncap; trace_cap = tcap2;
        CUDA_CHECK(cudaMemcpy(d_ctr + 1, &used, sizeof(int), cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(d_ull, &tused, sizeof(tused), cudaMemcpyHostToDevice));
        A.alns = d_alns; A.aln_cap = aln_cap; A.traces = d_traces; A.trace_cap = trace_cap;
      }
      if (d_list == nullptr) d_list = dalloc<int>((size_t) njobs + 1);
      if (d_big == nullptr) d_big = dalloc<unsigned char>((size_t) nbig * ALIGN_STATE_BYTES(ALIGN_W_BIG));
      int zero[4] = { 0, 0, 0, 0 };
      CUDA_CHECK(cudaMemcpy(d_ctr + 3, zero, sizeof(int), cudaMemcpyHostToDevice));
      LAUNCH(k_collect_failed, (njobs + 255) / 256, 256, 0, stream, d_jobs, njobs, d_list, d_ctr + 3);
      CUDA_CHECK(cudaMemcpy(d_ctr, zero, sizeof(int), cudaMemcpyHostToDevice));       // job counter
      CUDA_CHECK(cudaMemcpy(d_ctr + 2, zero, sizeof(int), cudaMemcpyHostToDevice));   // nfailed
      A.job_list = d_list; A.njobs = nfailed; A.big_state = d_big;
      if (g_align_tier != 0 && round == 0)                   // second tier: warp per job, window 128
        launch_align(A, false, std::min((nfailed + ALIGN_WARPS - 1) / ALIGN_WARPS, nblocks), stream);
      else
        launch_align(A, true, nbig / ALIGN_WARPS, stream);
    }

  std::vector<unsigned long long> stats = d2h(d_ull, 8);
  out->nalign = (int64_t) stats[1]; out->nwaves = (int64_t) stats[2]; out->ncells = (int64_t) stats[3];
  out->empty_band = (int64_t) stats[4];
  if (g_trace && stats[7] != 0)
    fprintf(stderr, "[trace] longest job: %llu waves in %llu alignments; mean %.0f waves per job\n",
            stats[7] >> 20, stats[7] & 0xfffff, njobs ? (double) stats[2] / njobs : 0.);
  if (g_trace && stats[5] != 0)
    fprintf(stderr, "[trace] duo kernel hand-offs: band %llu, cells %llu, trace %llu, other %llu\n",
            stats[5] & 0xffff, (stats[5] >> 16) & 0xffff, (stats[5] >> 32) & 0xffff, stats[5] >> 48);

  TRACE("report: overflow loop");
  // per-read sizing of the second half
  int     *d_novl = dalloc<int>(n + 1);
  int64_t *d_asum = dalloc<int64_t>(n + 1), *d_bsum = dalloc<int64_t>(n + 1);
  LAUNCH(k_read_totals, (n + 255) / 256, 256, 0, stream, n, d_job_off, d_jobs, d_alns, d_novl, d_asum, d_bsum);
  const int tb = (S <= 125) ? 1 : 2;                   // TRACE_XOVR, align.h:21
  // per-read capacities (records, fusion traces, output bytes of both families) and their exclusive
  // scans stay on the device; only the four totals come back
  int64_t *d_sz = dalloc<int64_t>((size_t) 4 * (n + 1));
  int64_t *d_ovl_off = dalloc<int64_t>((size_t) n + 1), *d_fus_off = dalloc<int64_t>((size_t) n + 1);
  int64_t *d_outa_off = dalloc<int64_t>((size_t) n + 1), *d_outb_off = dalloc<int64_t>((size_t) n + 1);
  int64_t *d_tot = dalloc<int64_t>(4);
  LAUNCH(k_read_sizes, (n + 255) / 256, 256, 0, stream, n, S, tb, rd->rlen, d_novl, d_asum, d_bsum, d_sz);
  LAUNCH(k_scan_used, 1, 1024, 0, stream, d_sz, n, d_ovl_off);
  LAUNCH(k_scan_used, 1, 1024, 0, stream, d_sz + (n + 1), n, d_fus_off);
  LAUNCH(k_scan_used, 1, 1024, 0, stream, d_sz + 2 * (size_t) (n + 1), n, d_outa_off);
  LAUNCH(k_scan_used, 1, 1024, 0, stream, d_sz + 3 * (size_t) (n + 1), n, d_outb_off);
  LAUNCH(k_gather_totals, 1, 32, 0, stream, n, d_ovl_off, d_fus_off, d_outa_off, d_outb_off, d_tot);
  int64_t h_tot[4] = { 0, 0, 0, 0 };
  CUDA_CHECK(cudaMemcpyAsync(h_tot, d_tot, sizeof(h_tot), cudaMemcpyDeviceToHost, stream));
  CUDA_CHECK(cudaStreamSynchronize(stream));
  const int64_t to = h_tot[0], tf = h_tot[1], ta = h_tot[2], tbb = h_tot[3];

  ReportArgs R;
  memset(&R, 0, sizeof(R));
  R.nreads = n; R.tfirst = rd->tfirst; R.spacing = S; R.do_a = do_a; R.do_b = do_b;
  R.small = (tb == 1); R.profile = g_par.profile; R.best_tie = g_par.best_tie;
  R.head = m->head; R.cand = m->cand; R.jobs = d_jobs; R.alns = d_alns;
  R.traces = d_traces; R.rlen = rd->rlen; R.cover = m->cover; R.coff = m->coff;
  R.job_off = d_job_off; R.ovl_off = d_ovl_off; R.fus_off = d_fus_off;
  R.outa_off = d_outa_off; R.outb_off = d_outb_off;
  R.ftraces = dalloc<uint16_t>((size_t) tf + 1);
  R.amatch = dalloc<Ovl>((size_t) to + 1); R.tmp = dalloc<Ovl>((size_t) to + 1);
  R.bmatch = do_b ? dalloc<Ovl>((size_t) to + 1) : nullptr;
  R.linker = dalloc<Links>((size_t) to + 1); R.perm = dalloc<int>((size_t) to + 1);
  R.part = dalloc<Zones>((size_t) to + 1);
  R.out_a = dalloc<uint8_t>((size_t) ta + 1);
  R.out_b = do_b ? dalloc<uint8_t>((size_t) tbb + 1) : nullptr;
  R.used_a = dalloc<int64_t>(n + 1); R.used_b = dalloc<int64_t>(n + 1);
  R.nrec_a = dalloc<int>(n + 1); R.nrec_b = dalloc<int>(n + 1);
  R.prof = dalloc<uint8_t>((size_t) m->h_coff[n] + 1);
  for (int i = 0; i <= 40; i++) R.spow[i] = pow(10., i / 10.);       // map.c:2279-2280
  R.error = d_ctr + 4; R.h2_events = d_ull + 6;
  TRACE("report: sizing+alloc");
  LAUNCH(k_report, (n + 63) / 64, 64, 0, stream, R);
  int rerr = 0;
  CUDA_CHECK(cudaMemcpyAsync(&rerr, d_ctr + 4, sizeof(int), cudaMemcpyDeviceToHost, stream));
  CUDA_CHECK(cudaStreamSynchronize(stream));
  if (rerr == 2)
    fatal("Compression of trace to bytes fails, value too big");       // align.c:3133 (H9)
  if (rerr != 0)
    fatal("Reporter: internal buffer overflow (code %d)", rerr);

  TRACE("report: k_report");
  // copy out and compact per read (record order = read order, as the per-thread files concatenate)
  { int64_t *d_off = dalloc<int64_t>((size_t) n + 1);
    const int cgrid = std::min((n + 7) / 8, sm_count() * 16);
    for (int fam = 0; fam < (do_b ? 2 : 1); fam++)
      { const uint8_t *src = fam ? R.out_b : R.out_a;
        const int64_t *soff = fam ? d_outb_off : d_outa_off;
        const int64_t *used = fam ? R.used_b : R.used_a;
        ByteVec &dstv = fam ? out->b : out->a;
        std::vector<int64_t> &roff = fam ? out->read_off_b : out->read_off_a;
        std::vector<int>     &nrec = fam ? out->read_nrec_b : out->read_nrec_a;
        LAUNCH(k_scan_used, 1, 1024, 0, stream, used, n, d_off);
        roff = d2h(d_off, (size_t) n + 1);
        const int64_t tot = roff[n];
        uint8_t *dense = dalloc<uint8_t>((size_t) tot + 1);
        if (n > 0 && tot > 0)
          LAUNCH(k_compact_out, cgrid, 256, 0, stream, src, soff, used, d_off, n, dense);
        dstv.resize((size_t) tot);
        if (tot > 0)
          CUDA_CHECK(cudaMemcpyAsync(dstv.data(), dense, (size_t) tot, cudaMemcpyDeviceToHost, stream));
        if (fam == 0 && tot > 48 && getenv("DAMGPU_TEST_CORRUPT_TRACE") != nullptr)
          LAUNCH(k_flip_byte, 1, 1, 0, stream, dense + 41);    // test hook: a record the checker must reject
        if (n > 0 && tot > 0)
          LAUNCH(k_check_trace, (n + 255) / 256, 256, 0, stream, dense, d_off, fam ? R.nrec_b : R.nrec_a, n,
                 S, (S <= 125) ? 1 : 2, d_ull + 8);       // TRACE_XOVR, align.h:45
        nrec = d2h(fam ? R.nrec_b : R.nrec_a, (size_t) n);
        int64_t total = 0;
        for (int i = 0; i < n; i++) total += nrec[i];
        if (fam) out->nrec_b = total; else out->nrec_a = total;
        CUDA_CHECK(cudaStreamSynchronize(stream));
        dfree(dense);
      }
    dfree(d_off);
    if (g_par.profile)
      { out->prof.resize((size_t) m->h_coff[n]);
        if (!out->prof.empty())
          CUDA_CHECK(cudaMemcpy(out->prof.data(), R.prof, out->prof.size(), cudaMemcpyDeviceToHost));
      }
    out->h2_events = (int64_t) d2h(d_ull + 6, 1)[0];
    out->trace_fails = (int64_t) d2h(d_ull + 8, 1)[0];
  }

  TRACE("report: copy out");
  dfree(R.ftraces); dfree(R.amatch); dfree(R.tmp); dfree(R.bmatch); dfree(R.linker); dfree(R.perm);
  dfree(R.part); dfree(R.out_a); dfree(R.out_b); dfree(R.used_a); dfree(R.used_b);
  dfree(R.nrec_a); dfree(R.nrec_b); dfree(R.prof);
  dfree(d_ovl_off); dfree(d_fus_off); dfree(d_outa_off); dfree(d_outb_off);
  dfree(d_novl); dfree(d_asum); dfree(d_bsum); dfree(d_sz); dfree(d_tot);
  dfree(d_alns); dfree(d_traces); dfree(d_list); dfree(d_big);
  dfree(d_cells); dfree(d_tscr); dfree(d_ctr); dfree(d_ull);
  dfree(d_cell_base); dfree(d_lane_cells); dfree(d_unwind); dfree(d_duo_win);
  dfree(d_jobs); dfree(d_job_off); dfree(d_cnt); dfree(rc.raw); dfree(d_tables);
  dfree(pk_a); dfree(pk_ac); dfree(pk_b);
  TRACE("report: frees");
  return out;
}

}  // namespace damgpu
