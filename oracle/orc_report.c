/* TEST INFRASTRUCTURE ONLY -- oracle restatement of the Reporter phase
 * (report_thread map.c:2362-2871, Entwine/Fusion/Handle_Redundancies :1953-2268,
 * special_log :2270-2302, Reporter :3227-3319).  Output = canonical record streams. */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#include "orc_internal.h"
#include "orc_align.h"

#define HITMIN       3
#define CHAIN_OFF    500.
#define CHAIN_OVL    400.
#define CHAIN_PLAY   1.4
#define DIFF_SCORE   2.3
#define TIE_SCORE    50
#define TIE_GAP      500
#define COMP_FLAG    0x1
#define START_FLAG   0x4
#define NEXT_FLAG    0x8
#define BEST_FLAG    0x10
#define TRACE_XOVR   125

typedef struct                     /* Path + Overlap, align.h:89-95,336-341 */
  { int64_t trace;                 /* offset into the per-read trace buffer */
    int tlen, diffs, abpos, bbpos, aepos, bepos;
  } OPath;

typedef struct
  { OPath    path;
    uint32_t flags;
    int      aread, bread;
  } Ovl;

typedef struct { uint16_t *trace; int64_t top, max; } TBuf;
typedef struct { int score, link, mark; } Links;
typedef struct { int beg, end, top; } Zones;

static int SPACING;

static void tbuf_room(TBuf *tb, int64_t len)
{ if (tb->top + len >= tb->max)
    { tb->max = (int64_t) (1.2*(tb->top+len)) + 20000;
      tb->trace = (uint16_t *) realloc(tb->trace,sizeof(uint16_t)*tb->max);
    }
}

static int iabs(int x) { return (x < 0 ? -x : x); }

/* Entwine, map.c:1953-2058 */
static int Entwine(OPath *jpath, OPath *kpath, TBuf *tbuf, int *where)
{ int ac, b2, y2, ae, i, j, k, den, min;
  uint16_t *ktrace = tbuf->trace + kpath->trace;
  uint16_t *jtrace = tbuf->trace + jpath->trace;

  min = 10000;
  den = 0;

  y2 = jpath->bbpos;
  j  = jpath->abpos/SPACING;
  b2 = kpath->bbpos;
  k  = kpath->abpos/SPACING;

  if (jpath->abpos == kpath->abpos)
    { min = iabs(y2-b2);
      if (min == 0)
        *where = kpath->abpos;
    }

  if (j < k)
    { ac = k*SPACING;
      j = 1 + 2*(k-j);
      k = 1;
      for (i = 1; i < j; i += 2)
        y2 += jtrace[i];
    }
  else
    { ac = j*SPACING;
      k = 1 + 2*(j-k);
      j = 1;
      for (i = 1; i < k; i += 2)
        b2 += ktrace[i];
    }

  ae = jpath->aepos;
  if (ae > kpath->aepos)
    ae = kpath->aepos;

  while (1)
    { ac += SPACING;
      if (ac >= ae)
        break;
      y2 += jtrace[j];
      b2 += ktrace[k];
      j += 2;
      k += 2;
      i = iabs(y2-b2);
      if (i <= min)
        { min = i;
          if (i == 0)
            *where = ac;
        }
      den += 1;
    }

  if (jpath->aepos == kpath->aepos)
    { i = iabs(jpath->bepos-kpath->bepos);
      if (i <= min)
        { min = i;
          if (i == 0)
            *where = kpath->aepos;
        }
    }

  if (den == 0)
    return (-1);
  return (min);
}

/* Fusion, map.c:2065-2109 */
static void Fusion(OPath *path1, int ap, OPath *path2, TBuf *tbuf)
{ int k, k1, k2, len, diff;
  uint16_t *trace;

  k1 = 2 * ((ap/SPACING) - (path1->abpos/SPACING));
  k2 = 2 * ((ap/SPACING) - (path2->abpos/SPACING));
  len = k1+(path2->tlen-k2);
  tbuf_room(tbuf,len);
  trace = tbuf->trace + tbuf->top;
  tbuf->top += len;

  diff = 0;
  len  = 0;
  if (k1 > 0)
    { uint16_t *t = tbuf->trace + path1->trace;
      for (k = 0; k < k1; k += 2)
        { trace[len++] = t[k];
          trace[len++] = t[k+1];
          diff += t[k];
        }
    }
  if (k2 < path2->tlen)
    { uint16_t *t = tbuf->trace + path2->trace;
      for (k = k2; k < path2->tlen; k += 2)
        { trace[len++] = t[k];
          trace[len++] = t[k+1];
          diff += t[k];
        }
    }

  path1->aepos = path2->aepos;
  path1->bepos = path2->bepos;
  path1->diffs = diff;
  path1->trace = trace - tbuf->trace;
  path1->tlen  = len;
}

/* Handle_Redundancies, map.c:2116-2268 */
static int Handle_Redundancies(Ovl *amatch, int novls, Ovl *bmatch, TBuf *tbuf, int cm)
{ OPath *jpath, *kpath, *jmath = NULL, *kmath = NULL;
  int j, k, no, dist, awhen = 0, bwhen = 0, hasB;

  hasB = (bmatch != NULL);

  for (j = 1; j < novls; j++)
    { jpath = &(amatch[j].path);
      if (hasB) jmath = &(bmatch[j].path);
      for (k = j-1; k >= 0; k--)
        { kpath = &(amatch[k].path);
          if (hasB) kmath = &(bmatch[k].path);

          if (kpath->abpos < 0)
            continue;

          if (jpath->abpos < kpath->abpos)
            { if (kpath->abpos <= jpath->aepos && kpath->bbpos <= jpath->bepos)
                { dist = Entwine(jpath,kpath,tbuf,&awhen);
                  if (dist == 0)
                    { if (kpath->aepos > jpath->aepos)
                        { if (hasB)
                            { if (cm)
                                { dist = Entwine(kmath,jmath,tbuf,&bwhen);
                                  if (dist != 0)
                                    continue;
                                  Fusion(jpath,awhen,kpath,tbuf);
                                  Fusion(kmath,bwhen,jmath,tbuf);
                                  *jmath = *kmath;
                                }
                              else
                                { dist = Entwine(jmath,kmath,tbuf,&bwhen);
                                  if (dist != 0)
                                    continue;
                                  Fusion(jpath,awhen,kpath,tbuf);
                                  Fusion(jmath,bwhen,kmath,tbuf);
                                }
                            }
                          else
                            Fusion(jpath,awhen,kpath,tbuf);
                        }
                      kpath->abpos = -1;
                      break;
                    }
                }
            }
          else
            { if (jpath->abpos <= kpath->aepos && jpath->bbpos <= kpath->bepos)
                { dist = Entwine(kpath,jpath,tbuf,&awhen);
                  if (dist == 0)
                    { if (kpath->abpos == jpath->abpos)
                        { if (kpath->aepos > jpath->aepos)
                            { *jpath = *kpath;
                              if (hasB)
                                *jmath = *kmath;
                            }
                        }
                      else if (jpath->aepos > kpath->aepos)
                        { if (hasB)
                            { if (cm)
                                { dist = Entwine(jmath,kmath,tbuf,&bwhen);
                                  if (dist != 0)
                                    continue;
                                  Fusion(kpath,awhen,jpath,tbuf);
                                  *jpath = *kpath;
                                  Fusion(jmath,bwhen,kmath,tbuf);
                                }
                              else
                                { dist = Entwine(kmath,jmath,tbuf,&bwhen);
                                  if (dist != 0)
                                    continue;
                                  Fusion(kpath,awhen,jpath,tbuf);
                                  *jpath = *kpath;
                                  Fusion(kmath,bwhen,jmath,tbuf);
                                  *jmath = *kmath;
                                }
                            }
                          else
                            { Fusion(kpath,awhen,jpath,tbuf);
                              *jpath = *kpath;
                            }
                        }
                      else
                        { *jpath = *kpath;
                          if (hasB)
                            *jmath = *kmath;
                        }
                      kpath->abpos = -1;
                      break;
                    }
                }
            }
        }
    }

  no = 0;
  for (j = 0; j < novls; j++)
    if (amatch[j].path.abpos >= 0)
      { if (hasB)
          bmatch[no] = bmatch[j];
        amatch[no++] = amatch[j];
      }
  return (no);
}

/* special_log, map.c:2270-2302 */
static int special_log(int cover)
{ static double Spow[51];
  static int    first = 1;
  int    l, r, m;
  double x;

  if (first)
    { first = 0;
      for (m = 0; m <= 40; m++)
        Spow[m] = pow(10.,m/10.);
    }
  if (cover <= 1)
    return (cover);
  else if (cover >= 10000)
    return (40);
  x = cover;
  l = 0;
  r = 41;
  while (l < r)
    { m = ((l+r) >> 1);
      if (Spow[m] <= x)
        l = m+1;
      else
        r = m;
    }
  return (l-1);
}

/* The three group sorts (map.c:2304-2341) under glibc's stable, indirect merge sort:
 * key order, equal keys in DESCENDING original index.  which: 0 = AMATCH
 * (abpos desc), 1 = BN_MATCH (bbpos desc), 2 = BC_MATCH (bepos asc). */
static void group_sort(Ovl *v, int n, int which)
{ int i, j;
  for (i = 1; i < n; i++)                   /* insertion sort from a reversed copy = stable
                                               on descending original index */
    ;
  { Ovl *t = (Ovl *) malloc(sizeof(Ovl)*n);
    for (i = 0; i < n; i++)
      t[i] = v[n-1-i];
    for (i = 1; i < n; i++)
      { Ovl x = t[i];
        for (j = i-1; j >= 0; j--)
          { int before;               /* does x sort strictly before t[j] ? */
            if (which == 0)
              before = (x.path.abpos > t[j].path.abpos);
            else if (which == 1)
              before = (x.path.bbpos > t[j].path.bbpos);
            else
              before = (x.path.bepos < t[j].path.bepos);
            if ( ! before)
              break;
            t[j+1] = t[j];
          }
        t[j+1] = x;
      }
    memcpy(v,t,sizeof(Ovl)*n);
    free(t);
  }
}

static void emit(uint8_t **buf, int64_t *len, int64_t *max, Ovl *o, TBuf *tbuf, int small)
{ int64_t need = 40 + (int64_t) o->path.tlen*2;
  uint8_t *p;
  uint16_t *t = tbuf->trace + o->path.trace;
  int32_t  w[10];
  int      j;

  if (*len + need > *max)
    { *max = (int64_t) (1.5*(*len+need)) + (1<<20);
      *buf = (uint8_t *) realloc(*buf,*max);
    }
  p = *buf + *len;
  w[0] = o->path.tlen; w[1] = o->path.diffs; w[2] = o->path.abpos; w[3] = o->path.bbpos;
  w[4] = o->path.aepos; w[5] = o->path.bepos; w[6] = (int32_t) o->flags;
  w[7] = o->aread; w[8] = o->bread; w[9] = 0;                 /* padding zeroed (H1) */
  memcpy(p,w,40);
  p += 40;
  if (small)
    { for (j = 0; j < o->path.tlen; j++)                      /* Compress_TraceTo8, H9 */
        { if (t[j] > 255)
            { fprintf(stderr,"oracle: Compression of trace to bytes fails, value too big\n");
              exit (1);
            }
          p[j] = (uint8_t) t[j];
        }
      *len += 40 + o->path.tlen;
    }
  else
    { memcpy(p,t,sizeof(uint16_t)*o->path.tlen);
      *len += 40 + 2*(int64_t) o->path.tlen;
    }
}

static void push_trace(TBuf *tbuf, OPath *dst, const orc_path *src)
{ tbuf_room(tbuf,src->tlen+1);
  dst->tlen = src->tlen; dst->diffs = src->diffs;
  dst->abpos = src->abpos; dst->bbpos = src->bbpos; dst->aepos = src->aepos; dst->bepos = src->bepos;
  dst->trace = tbuf->top;
  memmove(tbuf->trace+tbuf->top,src->trace,sizeof(uint16_t)*src->tlen);
  tbuf->top += src->tlen;
}

void orc_report(orc_mapper *m, const orc_block *ref,
                const uint8_t **abuf, int64_t *alen_out, int64_t *anrec,
                const uint8_t **bbuf, int64_t *blen_out, int64_t *bnrec,
                const uint8_t **prof, int64_t *proflen)
{ const orc_block *rd = &m->reads;
  int       K = m->par.kmer, doA = m->par.do_a, doB = m->par.do_b;
  int       hithr = HITMIN*K, small;
  orc_work *work = orc_work_new();
  orc_aspec spec;
  int16_t  *tables = (int16_t *) malloc(sizeof(int16_t)*65536);
  uint8_t  *acomp = (uint8_t *) malloc(rd->maxlen+4);
  TBuf      tb;
  Ovl      *amatch, *bmatch;
  Links    *linker, **perm;
  Zones    *part;
  int       Omax = 128, ar;

  SPACING = m->par.spacing;
  small   = (SPACING <= TRACE_XOVR);
  spec.spacing = SPACING;
  spec.score   = tables;
  spec.table   = tables + 32768;
  orc_align_spec(m->par.ave_corr,m->par.freq,&spec.ave_path,tables,tables+32768);

  tb.max = 40000; tb.top = 0;
  tb.trace = (uint16_t *) malloc(sizeof(uint16_t)*tb.max);
  amatch = (Ovl *) malloc(sizeof(Ovl)*Omax);
  bmatch = (Ovl *) malloc(sizeof(Ovl)*Omax);
  linker = (Links *) malloc(sizeof(Links)*Omax);
  perm   = (Links **) malloc(sizeof(Links *)*Omax);
  part   = (Zones *) malloc(sizeof(Zones)*Omax);

  m->alen = m->blen_out = m->anrec = m->bnrec = 0;
  m->h2_events = 0;
  m->proflen = m->coff[rd->nreads];
  free(m->prof);
  m->prof = (uint8_t *) calloc(m->proflen+1,1);

  for (ar = 0; ar < rd->nreads; ar++)
    { int alen = rd->rlen[ar];
      int atck = (alen-1)/SPACING + 1;
      int novl = 0, lovl = 0, hascomp = 0;
      int c, d;
      const uint8_t *aseq0 = rd->bases + rd->boff[ar];

      tb.top = 0;
      for (c = m->head[ar]; c >= 0; c = d)                   /* map.c:2460-2610 */
        { orc_cand *C = m->cand+c;
          int br = C->c.bread, cm = C->c.comp;
          int blen = ref->rlen[br];
          const uint8_t *bseq = ref->bases + ref->boff[br];
          const uint8_t *aseq;
          int apos, bpos, alast, n;

          if (cm)
            { if ( ! hascomp)
                { int i;                                      /* complement, map.c:1940-1948 */
                  acomp[0] = 4;
                  for (i = 0; i < alen; i++)
                    acomp[alen-i] = (uint8_t) (3-aseq0[i]);
                  acomp[alen+1] = 4;
                  hascomp = 1;
                }
              aseq = acomp+1;
            }
          else
            aseq = aseq0;

          apos  = C->c.alast;
          bpos  = C->c.blast;
          alast = alen + 1;
          for (n = 0; n < C->c.length; n++)
            { apos -= C->jumps[2*n];
              bpos -= C->jumps[2*n+1];
              if (apos < alast)
                { int dg, ad;
                  orc_path ap, bp;

                  if (cm)
                    { int ac = alen - apos, bc = blen - bpos;
                      dg = ac-bc;
                      ad = ac+bc;
                    }
                  else
                    { dg = apos-bpos;
                      ad = apos+bpos;
                    }
                  memset(&ap,0,sizeof(ap)); memset(&bp,0,sizeof(bp));
                  orc_local_align(work,&spec,aseq,alen,bseq,blen,cm,dg,dg,ad,&ap,&bp);
                  if (ap.aepos - ap.abpos >= hithr)
                    { alast = ap.abpos;
                      if (novl >= Omax-1)
                        { Omax = (int) (1.2*novl) + 128;
                          amatch = (Ovl *) realloc(amatch,sizeof(Ovl)*Omax);
                          bmatch = (Ovl *) realloc(bmatch,sizeof(Ovl)*Omax);
                          linker = (Links *) realloc(linker,sizeof(Links)*Omax);
                          perm   = (Links **) realloc(perm,sizeof(Links *)*Omax);
                          part   = (Zones *) realloc(part,sizeof(Zones)*Omax);
                        }
                      amatch[novl].aread = ar + rd->tfirst;
                      amatch[novl].bread = br;
                      amatch[novl].flags = (cm ? COMP_FLAG : 0);
                      push_trace(&tb,&amatch[novl].path,&ap);
                      if (doB)
                        { bmatch[novl].aread = br;
                          bmatch[novl].bread = ar + rd->tfirst;
                          bmatch[novl].flags = (cm ? COMP_FLAG : 0);
                          push_trace(&tb,&bmatch[novl].path,&bp);
                        }
                      novl += 1;
                    }
                }
            }

          d = C->next;
          if (d < 0 || m->cand[d].c.bread != br || m->cand[d].c.comp != cm)
            { if (novl-lovl > 1)
                novl = lovl + Handle_Redundancies(amatch+lovl,novl-lovl,
                                                  doB ? bmatch+lovl : NULL,&tb,cm);
              if (novl-lovl > 1)
                { group_sort(amatch+lovl,novl-lovl,0);
                  if (doB)
                    group_sort(bmatch+lovl,novl-lovl,cm ? 2 : 1);
                }
              lovl = novl;
            }
        }

      /* link DP, map.c:2630-2710 */
      if (novl > 0)
        { int br;
          lovl = 0;
          linker[0].link  = -1;
          linker[0].score = (int) ((amatch[0].path.aepos - amatch[0].path.abpos)
                                   - DIFF_SCORE * amatch[0].path.diffs);
          linker[0].mark  = 1;
          perm[0] = linker;
          br = amatch[0].bread;
          for (c = 1; c < novl; c++)
            { OPath *cpath = &(amatch[c].path);
              int cor, dor;

              linker[c].link  = -1;
              linker[c].score = (int) ((cpath->aepos - cpath->abpos) - DIFF_SCORE * cpath->diffs);
              linker[c].mark  = 1;
              perm[c] = linker+c;

              if (amatch[c].bread != br)
                { br = amatch[c].bread;
                  lovl = c;
                  continue;
                }

              cor = (amatch[c].flags & COMP_FLAG);
              for (d = c-1; d >= lovl; d--)
                if ((dor = (amatch[d].flags & COMP_FLAG)) == cor)
                  { OPath *dpath = &(amatch[d].path);
                    int    scr, scr2, gap, gap2;
                    double rat;

                    if (dor)
                      { if (dpath->bepos < cpath->bepos)
                          continue;
                      }
                    else
                      { if (dpath->bbpos < cpath->bbpos)
                          continue;
                      }

                    if (dpath->abpos <= cpath->aepos - CHAIN_OVL ||
                        dpath->bbpos <= cpath->bepos - CHAIN_OVL)
                      continue;

                    rat = ( (dpath->abpos - cpath->aepos + CHAIN_OFF) /
                            (dpath->bbpos - cpath->bepos + CHAIN_OFF) );

                    if (1. > rat*CHAIN_PLAY || rat > CHAIN_PLAY)
                      continue;

                    scr  = (int) (linker[d].score + (cpath->aepos - cpath->abpos)
                                                  - DIFF_SCORE * cpath->diffs);
                    scr2 = linker[c].score;
                    if (scr < scr2 - TIE_SCORE)
                      continue;

                    if (scr <= scr2 + TIE_SCORE)
                      { gap = dpath->abpos - cpath->aepos;
                        if (linker[d].link >= 0)
                          { if (linker[c].link < 0)             /* H2: reads amatch[-1] */
                              { m->h2_events += 1;
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
                          { if (scr < scr2)
                              continue;
                            if (scr == scr2)
                              { if (gap >= gap2)
                                  continue;
                              }
                          }
                      }

                    linker[c].link  = d;
                    linker[c].score = scr;
                    linker[d].mark  = 0;
                  }
            }

          /* LINK_SORT: score descending, stable (map.c:2355-2360,2712) */
          { int i, j;
            for (i = 1; i < novl; i++)
              { Links *x = perm[i];
                for (j = i-1; j >= 0 && perm[j]->score < x->score; j--)
                  perm[j+1] = perm[j];
                perm[j+1] = x;
              }
          }

          /* selection, map.c:2714-2816 */
          { int nparts = 0;
            for (c = 0; c < novl && perm[c]->score >= 0; c++)
              if (perm[c]->mark == 1)
                { int p, b, e, q, n, best;

                  b = e = (int) (perm[c] - linker);
                  for (p = linker[b].link; p >= 0 && linker[p].mark >= 0; p = linker[p].link)
                    e = p;

                  for (p = 0; p < nparts; p++)
                    if (amatch[b].path.abpos < part[p].end-100 &&
                        amatch[e].path.aepos > part[p].beg+100)
                      break;
                  if (p >= nparts)
                    { part[p].beg = amatch[b].path.abpos;
                      part[p].end = amatch[e].path.aepos;
                      part[p].top = linker[b].score;
                      best = 1;
                      nparts += 1;
                    }
                  else
                    { if (linker[b].score < m->par.best_tie * part[p].top)
                        continue;
                      best = (linker[b].score == part[p].top);
                    }

                  q = -1;
                  for (p = b; 1; p = n)
                    { linker[p].mark = -1;
                      if (doA)
                        { if (p == b)
                            { amatch[p].flags |= START_FLAG;
                              if (best)
                                amatch[p].flags |= BEST_FLAG;
                            }
                          else
                            amatch[p].flags |= NEXT_FLAG;
                          emit(&m->abuf,&m->alen,&m->amax,amatch+p,&tb,small);
                          m->anrec += 1;
                        }
                      n = linker[p].link;
                      if (doB)
                        { if (bmatch[p].flags & COMP_FLAG)
                            { linker[p].link = q;
                              q = p;
                            }
                          else
                            { if (p == b)
                                { bmatch[p].flags |= START_FLAG;
                                  if (best)
                                    bmatch[p].flags |= BEST_FLAG;
                                }
                              else
                                bmatch[p].flags |= NEXT_FLAG;
                              emit(&m->bbuf,&m->blen_out,&m->bmax,bmatch+p,&tb,small);
                              m->bnrec += 1;
                            }
                        }
                      if (p == e)
                        break;
                    }
                  if (doB && (bmatch[b].flags & COMP_FLAG))
                    { e = b;
                      b = q;
                      for (p = b; 1; p = linker[p].link)
                        { if (p == b)
                            { bmatch[p].flags |= START_FLAG;
                              if (best)
                                bmatch[p].flags |= BEST_FLAG;
                            }
                          else
                            bmatch[p].flags |= NEXT_FLAG;
                          emit(&m->bbuf,&m->blen_out,&m->bmax,bmatch+p,&tb,small);
                          m->bnrec += 1;
                          if (p == e)
                            break;
                        }
                    }
                }
          }
        }

      if (m->par.profile)                                    /* map.c:2835-2845 */
        { int16_t *cnt = m->cover + m->coff[ar];
          uint8_t *log = m->prof + m->coff[ar];
          int i, cc = 0;
          for (i = 0; i <= atck; i++)
            { cc += cnt[i];
              log[i] = (uint8_t) special_log(cc);
            }
        }
    }

  m->nalign = work->nalign; m->nwaves = work->nwaves; m->ncells = work->ncells;
  if (work->empty_band)
    fprintf(stderr,"oracle: note, %lld waves met an empty band\n",(long long) work->empty_band);

  free(part); free(perm); free(linker); free(bmatch); free(amatch);
  free(tb.trace); free(acomp); free(tables);
  { extern void orc_bandstats_print(void); orc_bandstats_print(); }
  orc_work_free(work);

  *abuf = m->abuf; *alen_out = m->alen; *anrec = m->anrec;
  *bbuf = m->bbuf; *blen_out = m->blen_out; *bnrec = m->bnrec;
  *prof = m->prof; *proflen = m->par.profile ? m->proflen : 0;
}

void orc_report_stats(const orc_mapper *m, int64_t *nalign, int64_t *nwaves, int64_t *ncells,
                      int64_t *h2_events)
{ *nalign = m->nalign; *nwaves = m->nwaves; *ncells = m->ncells; *h2_events = m->h2_events; }
