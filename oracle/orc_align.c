/* TEST INFRASTRUCTURE ONLY -- oracle restatement of the wave local alignment
 * (forward_wave align.c:353-1011, reverse_wave :1015-1720, Local_Alignment :1727-1946,
 * New_Align_Spec :207-269) as damapper uses it: lbord = hbord = -1, reach = 1, a != b.
 *
 * Per-diagonal state lives in circular arrays indexed by (k & WMASK) instead of the
 * reference's re-centred vector; the band is far narrower than the window. */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "orc_align.h"

#define TRIM_LEN   15
#define DUB_TRIM   45
#define PATH_LEN   60
#define PATH_TOP   0x1000000000000000ull
#define PATH_INT   0x0fffffffffffffffull
#define TRIM_MASK  0x7fff
#define TRIM_MLAG  250
#define WAVE_LAG   30
#define FRACTION   1000
#define IMAX       0x7fffffff

#define WBITS 13
#define WSIZE (1<<WBITS)
#define WMASK (WSIZE-1)
#define IX(k) ((k) & WMASK)

static const double Bias_Factor[10] = { .690, .690, .690, .690, .780,
                                        .850, .900, .933, .966, 1.000 };

typedef struct { int mscore, dscore; int16_t *table, *score; } TBits;

static void set_table(int bit, int prefix, int score, int max, TBits *p)   /* align.c:207-218 */
{ if (bit >= TRIM_LEN)
    { p->table[prefix] = (int16_t) (score-max);
      p->score[prefix] = (int16_t) score;
    }
  else
    { if (score > max)
        max = score;
      set_table(bit+1,(prefix<<1),score - p->dscore,max,p);
      set_table(bit+1,(prefix<<1) | 1,score + p->mscore,max,p);
    }
}

void orc_align_spec(double ave_corr, const float freq[4], int *ave_path,
                    int16_t *score, int16_t *table)                         /* align.c:222-269 */
{ TBits  parms;
  double match;
  int    bias;

  match = freq[0] + freq[3];
  if ((match <= 0.) == (match > 0.))
    match = .5;
  if (match > .5)
    match = 1.-match;
  bias = (int) ((match+.025)*20.-1.);
  if (match < .2)
    bias = 3;
  *ave_path    = (int) (PATH_LEN * (1. - Bias_Factor[bias] * (1. - ave_corr)));
  parms.mscore = (int) (FRACTION * Bias_Factor[bias] * (1. - ave_corr));
  parms.dscore = FRACTION - parms.mscore;
  parms.score  = score;
  parms.table  = table;
  set_table(0,0,0,0,&parms);
}

orc_work *orc_work_new(void)
{ orc_work *w = (orc_work *) calloc(1,sizeof(orc_work));
  w->V  = (int *) calloc(WSIZE,sizeof(int));
  w->M  = (int *) calloc(WSIZE,sizeof(int));
  w->HA = (int *) calloc(WSIZE,sizeof(int));
  w->HB = (int *) calloc(WSIZE,sizeof(int));
  w->NA = (int *) calloc(WSIZE,sizeof(int));
  w->NB = (int *) calloc(WSIZE,sizeof(int));
  w->T  = (uint64_t *) calloc(WSIZE,sizeof(uint64_t));
  w->cmax  = 20000;
  w->cells = (orc_pebble *) malloc(sizeof(orc_pebble)*w->cmax);
  return (w);
}

void orc_work_free(orc_work *w)
{ if (w == NULL) return;
  free(w->V); free(w->M); free(w->HA); free(w->HB); free(w->NA); free(w->NB); free(w->T);
  free(w->cells); free(w->tbuf);
  free(w);
}

/* ORC_BANDSTATS=1: histogram of band widths per wave and of the per-call maxima (dev aid) */
static long long g_bandhist[64], g_maxhist[64], g_cellhist[64];
static int g_bandstats = -1, g_callmax = 0;
static void band_note(int width)
{ if (g_bandstats < 0) g_bandstats = (getenv("ORC_BANDSTATS") != NULL);
  if (!g_bandstats) return;
  g_bandhist[width < 63 ? width : 63] += 1;
  if (width > g_callmax) g_callmax = width;
}
static void call_note(int avail)
{ if (g_bandstats <= 0) return;
  g_maxhist[g_callmax < 63 ? g_callmax : 63] += 1; g_callmax = 0;
  { int b = avail / 64; g_cellhist[b < 63 ? b : 63] += 1; }
}
void orc_bandstats_print(void)
{ int i;
  if (g_bandstats <= 0) return;
  fprintf(stderr,"band width per wave:"); for (i = 0; i < 64; i++) if (g_bandhist[i]) fprintf(stderr," %d:%lld",i,g_bandhist[i]);
  fprintf(stderr,"\nmax band per call:"); for (i = 0; i < 64; i++) if (g_maxhist[i]) fprintf(stderr," %d:%lld",i,g_maxhist[i]);
  fprintf(stderr,"\ncells per call (x64):"); for (i = 0; i < 64; i++) if (g_cellhist[i]) fprintf(stderr," %d:%lld",i,g_cellhist[i]);
  fprintf(stderr,"\n");
}

static inline int new_cell(orc_work *w, int *avail, int ptr, int diag, int diff, int mark)
{ orc_pebble *pb;
  if (*avail >= w->cmax)
    { w->cmax  = (int) (w->cmax*1.2) + 10000;
      w->cells = (orc_pebble *) realloc(w->cells,sizeof(orc_pebble)*w->cmax);
    }
  pb = w->cells + *avail;
  pb->ptr = ptr; pb->diag = diag; pb->diff = diff; pb->mark = mark;
  return ((*avail)++);
}

static void band_check(int low, int hgh)
{ band_note(hgh-low+1);
  if (hgh-low+4 >= WSIZE)
    { fprintf(stderr,"oracle: wave band wider than %d diagonals\n",WSIZE); exit (1); }
}

/* forward_wave, align.c:353-1011 */
static void forward_wave(orc_work *work, const orc_aspec *spec, const uint8_t *aseq,
                         const uint8_t *bseq, orc_path *apath, orc_path *bpath,
                         int *mind, int maxd, int mida, int aoff, int boff)
{ int *V = work->V, *M = work->M, *HA = work->HA, *HB = work->HB, *NA = work->NA, *NB = work->NB;
  uint64_t *T = work->T;
  orc_pebble *cells;
  int avail = 0;
  int TS = spec->spacing, PATH_AVE = spec->ave_path;
  const int16_t *SCORE = spec->score, *TABLE = spec->table;

  int hgh, low, dif;
  int besta, besty, trima, trimy, trimd, trimha, trimhb;
  int morea, morey, mored, moreha, morehb, more, morem, lasta, aclip, bclip;

  hgh = maxd;
  low = *mind;
  dif = 0;

  more  = 1;
  aclip =  IMAX;
  bclip = -IMAX;

  besta  = trima  = morea = lasta = mida;
  besty  = trimy  = morey = (mida-hgh) >> 1;
  trimd  = mored  = 0;
  trimha = moreha = 0;
  trimhb = morehb = 1;
  morem  = -1;

  { int k;                                                   /* 0-wave, align.c:433-556 */
    band_check(low,hgh);
    for (k = hgh; k >= low; k--)
      { int y, c, d, ha, hb, na, nb;
        const uint8_t *a = aseq + k;

        y  = (mida-k) >> 1;
        na = (((y+k)+(TS-aoff))/TS-1)*TS+aoff;
        ha = new_cell(work,&avail,-1,k,0,na);
        na += TS;
        nb = ((y+(TS-boff))/TS-1)*TS+boff;
        hb = new_cell(work,&avail,-1,k,0,nb);
        nb += TS;

        while (1)
          { c = bseq[y];
            if (c == 4)
              { more = 0;
                if (bclip < k)
                  bclip = k;
                break;
              }
            d = a[y];
            if (c != d)
              { if (d == 4)
                  { more  = 0;
                    aclip = k;
                  }
                break;
              }
            y += 1;
          }
        c = (y << 1) + k;

        while (y+k >= na)
          { ha = new_cell(work,&avail,ha,k,0,na);
            na += TS;
          }
        while (y >= nb)
          { hb = new_cell(work,&avail,hb,k,0,nb);
            nb += TS;
          }

        if (c > besta)
          { besta  = trima = lasta = c;
            besty  = trimy = y;
            trimha = ha;
            trimhb = hb;
          }

        V[IX(k)]  = c;
        T[IX(k)]  = PATH_INT;
        M[IX(k)]  = PATH_LEN;
        HA[IX(k)] = ha;
        HB[IX(k)] = hb;
        NA[IX(k)] = na;
        NB[IX(k)] = nb;
      }
  }

  if (more == 0)                                             /* align.c:558-583 */
    { if (bseq[besty] != 4 && aseq[besta - besty] != 4)
        more = 1;
      if (hgh >= aclip)
        { hgh = aclip-1;
          if (morem <= M[IX(aclip)])
            { morem  = M[IX(aclip)];
              morea  = V[IX(aclip)];
              morey  = (morea - aclip)/2;
              moreha = HA[IX(aclip)];
              morehb = HB[IX(aclip)];
            }
        }
      if (low <= bclip)
        { low = bclip+1;
          if (morem <= M[IX(bclip)])
            { morem  = M[IX(bclip)];
              morea  = V[IX(bclip)];
              morey  = (morea - bclip)/2;
              moreha = HA[IX(bclip)];
              morehb = HB[IX(bclip)];
            }
        }
      aclip =  IMAX;
      bclip = -IMAX;
    }

  while (more && lasta >= besta - TRIM_MLAG)                 /* align.c:592-898 */
    { int      k, n, ua, ub, am, ac, ap;
      uint64_t t;

      if (hgh < low)           /* empty band: the reference reads stale cells here (never
                                  observed); both oracle and product stop instead */
        { work->empty_band += 1;
          break;
        }

      low -= 1;
      hgh += 1;
      band_check(low,hgh);

      NA[IX(low)] = NA[IX(low+1)];                           /* minp = -INT32_MAX */
      NB[IX(low)] = NB[IX(low+1)];
      V[IX(low)]  = -1;

      NA[IX(hgh)] = NA[IX(hgh-1)];                           /* maxp = INT32_MAX */
      NB[IX(hgh)] = NB[IX(hgh-1)];
      V[IX(hgh)]  = am = -1;

      dif += 1;

      ac = V[IX(hgh+1)] = V[IX(low-1)] = -1;
      t  = PATH_INT;
      n  = PATH_LEN;
      ua = ub = -1;
      for (k = hgh; k >= low; k--)
        { int y, m, ha, hb, c, d;
          uint64_t b;
          const uint8_t *a = aseq + k;

          ap = ac;
          ac = am;
          am = V[IX(d = k-1)];

          if (ac < am)
            if (am < ap)
              { c = ap+1; m = n; b = t; ha = ua; hb = ub; }
            else
              { c = am+1; m = M[IX(d)]; b = T[IX(d)]; ha = HA[IX(d)]; hb = HB[IX(d)]; }
          else
            if (ac < ap)
              { c = ap+1; m = n; b = t; ha = ua; hb = ub; }
            else
              { c = ac+2; m = M[IX(k)]; b = T[IX(k)]; ha = HA[IX(k)]; hb = HB[IX(k)]; }

          if ((b & PATH_TOP) != 0)
            m -= 1;
          b <<= 1;

          y = (c-k) >> 1;
          while (1)
            { c = bseq[y];
              if (c == 4)
                { more = 0;
                  if (bclip < k)
                    bclip = k;
                  break;
                }
              d = a[y];
              if (c != d)
                { if (d == 4)
                    { more  = 0;
                      aclip = k;
                    }
                  break;
                }
              y += 1;
              if ((b & PATH_TOP) == 0)
                m += 1;
              b = (b << 1) | 1;
            }
          c = (y << 1) + k;

          while (y+k >= NA[IX(k)])
            { if (work->cells[ha].mark < NA[IX(k)])
                ha = new_cell(work,&avail,ha,k,dif,NA[IX(k)]);
              NA[IX(k)] += TS;
            }
          while (y >= NB[IX(k)])
            { if (work->cells[hb].mark < NB[IX(k)])
                hb = new_cell(work,&avail,hb,k,dif,NB[IX(k)]);
              NB[IX(k)] += TS;
            }

          if (c > besta)
            { besta = c;
              besty = y;
              if (m >= PATH_AVE)
                { lasta = c;
                  if (TABLE[b & TRIM_MASK] >= 0)
                    if (TABLE[(b >> TRIM_LEN) & TRIM_MASK] + SCORE[b & TRIM_MASK] >= 0)
                      { trima  = c;
                        trimy  = y;
                        trimd  = dif;
                        trimha = ha;
                        trimhb = hb;
                      }
                }
            }

          t  = T[IX(k)];
          n  = M[IX(k)];
          ua = HA[IX(k)];
          ub = HB[IX(k)];
          V[IX(k)]  = c;
          T[IX(k)]  = b;
          M[IX(k)]  = m;
          HA[IX(k)] = ha;
          HB[IX(k)] = hb;
        }

      if (more == 0)
        { if (bseq[besty] != 4 && aseq[besta-besty] != 4)
            more = 1;
          if (hgh >= aclip)
            { hgh = aclip-1;
              if (morem <= M[IX(aclip)])
                { morem  = M[IX(aclip)];
                  morea  = V[IX(aclip)];
                  morey  = (morea - aclip)/2;
                  mored  = dif;
                  moreha = HA[IX(aclip)];
                  morehb = HB[IX(aclip)];
                }
            }
          if (low <= bclip)
            { low = bclip+1;
              if (morem <= M[IX(bclip)])
                { morem  = M[IX(bclip)];
                  morea  = V[IX(bclip)];
                  morey  = (morea - bclip)/2;
                  mored  = dif;
                  moreha = HA[IX(bclip)];
                  morehb = HB[IX(bclip)];
                }
            }
          aclip =  IMAX;
          bclip = -IMAX;
        }

      n = besta - WAVE_LAG;
      while (hgh >= low)
        if (V[IX(hgh)] < n)
          hgh -= 1;
        else
          { while (V[IX(low)] < n)
              low += 1;
            break;
          }

      work->nwaves += 1;                                     /* WAVE_STATS, align.c:887-893 */
      work->ncells += (hgh-low)+1;
    }

  call_note(avail);
  cells = work->cells;
  { uint16_t *atrace = apath->trace;                         /* align.c:900-1007 */
    uint16_t *btrace = bpath->trace;
    int atlen, btlen, trimx, a, b, k, h, d, e;

    if (morem >= 0)                                          /* REACH = 1 */
      { trimx  = morea-morey;
        trimy  = morey;
        trimd  = mored;
        trimha = moreha;
        trimhb = morehb;
      }
    else
      trimx = trima-trimy;

    atlen = btlen = 0;

    a = -1;
    for (h = trimha; h >= 0; h = b)
      { b = cells[h].ptr;
        cells[h].ptr = a;
        a = h;
      }
    h = a;

    k = cells[h].diag;
    b = (mida-k)/2;
    e = 0;
    for (h = cells[h].ptr; h >= 0; h = cells[h].ptr)
      { k = cells[h].diag;
        a = cells[h].mark - k;
        d = cells[h].diff;
        atrace[atlen++] = (uint16_t) (d-e);
        atrace[atlen++] = (uint16_t) (a-b);
        b = a;
        e = d;
      }
    if (b+k != trimx)
      { atrace[atlen++] = (uint16_t) (trimd-e);
        atrace[atlen++] = (uint16_t) (trimy-b);
      }
    else if (b != trimy)
      { atrace[atlen-1] = (uint16_t) (atrace[atlen-1] + (trimy-b));
        atrace[atlen-2] = (uint16_t) (atrace[atlen-2] + (trimd-e));
      }

    a = -1;
    for (h = trimhb; h >= 0; h = b)
      { b = cells[h].ptr;
        cells[h].ptr = a;
        a = h;
      }
    h = a;

    k = cells[h].diag;
    b = (mida+k)/2;
    e = 0;
    low = k;
    for (h = cells[h].ptr; h >= 0; h = cells[h].ptr)
      { k = cells[h].diag;
        a = cells[h].mark + k;
        d = cells[h].diff;
        btrace[btlen++] = (uint16_t) (d-e);
        btrace[btlen++] = (uint16_t) (a-b);
        b = a;
        e = d;
      }
    if (b-k != trimy)
      { btrace[btlen++] = (uint16_t) (trimd-e);
        btrace[btlen++] = (uint16_t) (trimx-b);
      }
    else if (b != trimx)
      { btrace[btlen-1] = (uint16_t) (btrace[btlen-1] + (trimx-b));
        btrace[btlen-2] = (uint16_t) (btrace[btlen-2] + (trimd-e));
      }

    apath->aepos = trimx;
    apath->bepos = trimy;
    apath->diffs = trimd;
    apath->tlen  = atlen;
    bpath->tlen  = btlen;
  }

  *mind = low;
}

/* reverse_wave, align.c:1015-1720 */
static void reverse_wave(orc_work *work, const orc_aspec *spec, const uint8_t *aseq0,
                         const uint8_t *bseq0, orc_path *apath, orc_path *bpath,
                         int mind, int maxd, int mida, int aoff, int boff)
{ const uint8_t *aseq = aseq0 - 1;
  const uint8_t *bseq = bseq0 - 1;
  int *V = work->V, *M = work->M, *HA = work->HA, *HB = work->HB, *NA = work->NA, *NB = work->NB;
  uint64_t *T = work->T;
  orc_pebble *cells;
  int avail = 0;
  int TS = spec->spacing, PATH_AVE = spec->ave_path;
  const int16_t *SCORE = spec->score, *TABLE = spec->table;

  int hgh, low, dif;
  int besta, besty, trima, trimy, trimd, trimha, trimhb;
  int morea, morey, mored, moreha, morehb, more, morem, lasta, aclip, bclip;

  hgh = maxd;
  low = mind;
  dif = 0;

  more  = 1;
  aclip = -IMAX;
  bclip =  IMAX;

  besta  = trima  = morea = lasta = mida;
  besty  = trimy  = morey = (mida-hgh) >> 1;
  trimd  = mored  = 0;
  trimha = moreha = 0;
  trimhb = morehb = 1;
  morem  = -1;

  { int k;                                                   /* align.c:1093-1214 */
    band_check(low,hgh);
    for (k = low; k <= hgh; k++)
      { int y, c, d, ha, hb, na, nb;
        const uint8_t *a = aseq + k;

        y  = (mida-k) >> 1;
        na = (((y+k)+(TS-aoff)-1)/TS-1)*TS+aoff;
        ha = new_cell(work,&avail,-1,k,0,y+k);
        nb = ((y+(TS-boff)-1)/TS-1)*TS+boff;
        hb = new_cell(work,&avail,-1,k,0,y);

        while (1)
          { c = bseq[y];
            if (c == 4)
              { more = 0;
                if (bclip > k)
                  bclip = k;
                break;
              }
            d = a[y];
            if (c != d)
              { if (d == 4)
                  { more  = 0;
                    aclip = k;
                  }
                break;
              }
            y -= 1;
          }
        c = (y << 1) + k;

        while (y+k <= na)
          { ha = new_cell(work,&avail,ha,k,0,na);
            na -= TS;
          }
        while (y <= nb)
          { hb = new_cell(work,&avail,hb,k,0,nb);
            nb -= TS;
          }

        if (c < besta)
          { besta  = trima = lasta = c;
            besty  = trimy = y;
            trimha = ha;
            trimhb = hb;
          }

        V[IX(k)]  = c;
        T[IX(k)]  = PATH_INT;
        M[IX(k)]  = PATH_LEN;
        HA[IX(k)] = ha;
        HB[IX(k)] = hb;
        NA[IX(k)] = na;
        NB[IX(k)] = nb;
      }
  }

  if (more == 0)                                             /* align.c:1216-1241 */
    { if (bseq[besty] != 4 && aseq[besta - besty] != 4)
        more = 1;
      if (low <= aclip)
        { low = aclip+1;
          if (morem <= M[IX(aclip)])
            { morem  = M[IX(aclip)];
              morea  = V[IX(aclip)];
              morey  = (morea - aclip)/2;
              moreha = HA[IX(aclip)];
              morehb = HB[IX(aclip)];
            }
        }
      if (hgh >= bclip)
        { hgh = bclip-1;
          if (morem <= M[IX(bclip)])
            { morem  = M[IX(bclip)];
              morea  = V[IX(bclip)];
              morey  = (morea - bclip)/2;
              moreha = HA[IX(bclip)];
              morehb = HB[IX(bclip)];
            }
        }
      aclip = -IMAX;
      bclip =  IMAX;
    }

  while (more && lasta <= besta + TRIM_MLAG)                 /* align.c:1248-1552 */
    { int      k, n, ua, ub, am, ac, ap;
      uint64_t t;

      if (hgh < low)
        { work->empty_band += 1;
          break;
        }

      low -= 1;
      hgh += 1;
      band_check(low,hgh);

      NA[IX(low)] = NA[IX(low+1)];
      NB[IX(low)] = NB[IX(low+1)];
      V[IX(low)]  = ap = IMAX;

      NA[IX(hgh)] = NA[IX(hgh-1)];
      NB[IX(hgh)] = NB[IX(hgh-1)];
      V[IX(hgh)]  = IMAX;

      dif += 1;

      ac = V[IX(hgh+1)] = V[IX(low-1)] = IMAX;
      t  = PATH_INT;
      n  = PATH_LEN;
      ua = ub = -1;
      for (k = low; k <= hgh; k++)
        { int y, m, ha, hb, c, d;
          uint64_t b;
          const uint8_t *a = aseq + k;

          am = ac;
          ac = ap;
          ap = V[IX(d = k+1)];

          if (ac > ap)
            if (ap > am)
              { c = am-1; m = n; b = t; ha = ua; hb = ub; }
            else
              { c = ap-1; m = M[IX(d)]; b = T[IX(d)]; ha = HA[IX(d)]; hb = HB[IX(d)]; }
          else
            if (ac > am)
              { c = am-1; m = n; b = t; ha = ua; hb = ub; }
            else
              { c = ac-2; m = M[IX(k)]; b = T[IX(k)]; ha = HA[IX(k)]; hb = HB[IX(k)]; }

          if ((b & PATH_TOP) != 0)
            m -= 1;
          b <<= 1;

          y = (c-k) >> 1;
          while (1)
            { c = bseq[y];
              if (c == 4)
                { more = 0;
                  if (bclip > k)
                    bclip = k;
                  break;
                }
              d = a[y];
              if (c != d)
                { if (d == 4)
                    { more  = 0;
                      aclip = k;
                    }
                  break;
                }
              y -= 1;
              if ((b & PATH_TOP) == 0)
                m += 1;
              b = (b << 1) | 1;
            }
          c = (y << 1) + k;

          while (y+k <= NA[IX(k)])
            { if (work->cells[ha].mark > NA[IX(k)])
                ha = new_cell(work,&avail,ha,k,dif,NA[IX(k)]);
              NA[IX(k)] -= TS;
            }
          while (y <= NB[IX(k)])
            { if (work->cells[hb].mark > NB[IX(k)])
                hb = new_cell(work,&avail,hb,k,dif,NB[IX(k)]);
              NB[IX(k)] -= TS;
            }

          if (c < besta)
            { besta = c;
              besty = y;
              if (m >= PATH_AVE)
                { lasta = c;
                  if (TABLE[b & TRIM_MASK] >= 0)
                    if (TABLE[(b >> TRIM_LEN) & TRIM_MASK] + SCORE[b & TRIM_MASK] >= 0)
                      { trima  = c;
                        trimy  = y;
                        trimd  = dif;
                        trimha = ha;
                        trimhb = hb;
                      }
                }
            }

          t  = T[IX(k)];
          n  = M[IX(k)];
          ua = HA[IX(k)];
          ub = HB[IX(k)];
          V[IX(k)]  = c;
          T[IX(k)]  = b;
          M[IX(k)]  = m;
          HA[IX(k)] = ha;
          HB[IX(k)] = hb;
        }

      if (more == 0)
        { if (bseq[besty] != 4 && aseq[besta - besty] != 4)
            more = 1;
          if (low <= aclip)
            { low = aclip+1;
              if (morem <= M[IX(aclip)])
                { morem  = M[IX(aclip)];
                  morea  = V[IX(aclip)];
                  morey  = (morea - aclip)/2;
                  mored  = dif;
                  moreha = HA[IX(aclip)];
                  morehb = HB[IX(aclip)];
                }
            }
          if (hgh >= bclip)
            { hgh = bclip-1;
              if (morem <= M[IX(bclip)])
                { morem  = M[IX(bclip)];
                  morea  = V[IX(bclip)];
                  morey  = (morea - bclip)/2;
                  mored  = dif;
                  moreha = HA[IX(bclip)];
                  morehb = HB[IX(bclip)];
                }
            }
          aclip = -IMAX;
          bclip =  IMAX;
        }

      n = besta + WAVE_LAG;
      while (hgh >= low)
        if (V[IX(hgh)] > n)
          hgh -= 1;
        else
          { while (V[IX(low)] > n)
              low += 1;
            break;
          }

      work->nwaves += 1;
      work->ncells += (hgh-low)+1;
    }

  call_note(avail);
  cells = work->cells;
  { uint16_t *atrace = apath->trace;                         /* align.c:1554-1717 */
    uint16_t *btrace = bpath->trace;
    int atlen, btlen, trimx, a, b, k, h, d, e;

    if (morem >= 0)
      { trimx  = morea-morey;
        trimy  = morey;
        trimd  = mored;
        trimha = moreha;
        trimhb = morehb;
      }
    else
      trimx = trima-trimy;

    atlen = btlen = 0;

    a = -1;
    for (h = trimha; h >= 0; h = b)
      { b = cells[h].ptr;
        cells[h].ptr = a;
        a = h;
      }
    h = a;

    k = cells[h].diag;
    b = cells[h].mark - k;
    e = 0;
    a = 0; d = 0;
    if ((b+k)%TS != aoff)
      { h = cells[h].ptr;
        if (h < 0)
          { a = trimy;
            d = trimd;
          }
        else
          { k = cells[h].diag;
            a = cells[h].mark - k;
            d = cells[h].diff;
          }
        if (apath->tlen == 0)
          { atrace[--atlen] = (uint16_t) (b-a);
            atrace[--atlen] = (uint16_t) (d-e);
          }
        else
          { atrace[1] = (uint16_t) (atrace[1] + (b-a));
            atrace[0] = (uint16_t) (atrace[0] + (d-e));
          }
        b = a;
        e = d;
      }
    if (h >= 0)
      { for (h = cells[h].ptr; h >= 0; h = cells[h].ptr)
          { k = cells[h].diag;
            a = cells[h].mark - k;
            atrace[--atlen] = (uint16_t) (b-a);
            d = cells[h].diff;
            atrace[--atlen] = (uint16_t) (d-e);
            b = a;
            e = d;
          }
        if (b+k != trimx)
          { atrace[--atlen] = (uint16_t) (b-trimy);
            atrace[--atlen] = (uint16_t) (trimd-e);
          }
        else if (b != trimy)
          { atrace[atlen+1] = (uint16_t) (atrace[atlen+1] + (b-trimy));
            atrace[atlen]   = (uint16_t) (atrace[atlen]   + (trimd-e));
          }
      }

    a = -1;
    for (h = trimhb; h >= 0; h = b)
      { b = cells[h].ptr;
        cells[h].ptr = a;
        a = h;
      }
    h = a;

    k = cells[h].diag;
    b = cells[h].mark + k;
    e = 0;
    if ((b-k)%TS != boff)
      { h = cells[h].ptr;
        if (h < 0)
          { a = trimx;
            d = trimd;
          }
        else
          { k = cells[h].diag;
            a = cells[h].mark + k;
            d = cells[h].diff;
          }
        if (bpath->tlen == 0)
          { btrace[--btlen] = (uint16_t) (b-a);
            btrace[--btlen] = (uint16_t) (b-a);          /* sic, align.c:1670-1671 (H3) */
          }
        else
          { btrace[1] = (uint16_t) (btrace[1] + (b-a));
            btrace[0] = (uint16_t) (btrace[0] + (d-e));
          }
        b = a;
        e = d;
      }

    if (h >= 0)
      { for (h = cells[h].ptr; h >= 0; h = cells[h].ptr)
          { k = cells[h].diag;
            a = cells[h].mark + k;
            btrace[--btlen] = (uint16_t) (b-a);
            d = cells[h].diff;
            btrace[--btlen] = (uint16_t) (d-e);
            b = a;
            e = d;
          }
        if (b-k != trimy)
          { btrace[--btlen] = (uint16_t) (b-trimx);
            btrace[--btlen] = (uint16_t) (trimd-e);
          }
        else if (b != trimx)
          { btrace[btlen+1] = (uint16_t) (btrace[btlen+1] + (b-trimx));
            btrace[btlen]   = (uint16_t) (btrace[btlen]   + (trimd-e));
          }
      }

    apath->abpos = trimx;
    apath->bbpos = trimy;
    apath->diffs = apath->diffs + trimd;
    apath->tlen  = apath->tlen  - atlen;
    apath->trace = atrace + atlen;
    bpath->tlen  = bpath->tlen  - btlen;
    bpath->trace = btrace + btlen;
  }
}

/* Local_Alignment, align.c:1727-1946, called as (dg,dg,ad,-1,-1) (map.c:2513).
 * On return apath/bpath->trace point into work->tbuf. */
void orc_local_align(orc_work *work, const orc_aspec *spec, const uint8_t *aseq, int alen,
                     const uint8_t *bseq, int blen, int acomp, int low, int hgh, int anti,
                     orc_path *apath, orc_path *bpath)
{ int aoff, boff, fshort, rshort;
  int maxtp;

  if (alen < blen)                                           /* align.c:1752-1765 */
    maxtp = 2*(blen/spec->spacing+2);
  else
    maxtp = 2*(alen/spec->spacing+2);
  if (4*maxtp > work->tmax)
    { work->tmax = (int) (4*maxtp*1.2) + 10000;
      work->tbuf = (uint16_t *) realloc(work->tbuf,sizeof(uint16_t)*work->tmax);
    }
  apath->trace = work->tbuf + maxtp;
  bpath->trace = apath->trace + 2*maxtp;

  while (((anti-hgh) >> 1) < 0)
    hgh -= 1;

  if (acomp)                                                 /* align.c:1794-1805 */
    { aoff = alen % spec->spacing;
      boff = 0;
    }
  else
    { aoff = 0;
      boff = 0;
    }

  work->nalign += 1;
  forward_wave(work,spec,aseq,bseq,apath,bpath,&low,hgh,anti,aoff,boff);
  fshort = ((apath->aepos + apath->bepos) - anti < DUB_TRIM);

  reverse_wave(work,spec,aseq,bseq,apath,bpath,low,low,anti,aoff,boff);
  rshort = (anti - (apath->abpos + apath->bbpos) < DUB_TRIM);

  if (fshort)
    { if (rshort)
        { apath->aepos = apath->abpos = (apath->abpos+apath->aepos)/2;
          apath->bepos = apath->bbpos = (apath->bbpos+apath->bepos)/2;
          bpath->aepos = bpath->abpos = (bpath->abpos+bpath->aepos)/2;   /* uninitialised in the
                                                 reference; overwritten below (align.c:1857-1912) */
          bpath->bepos = bpath->bbpos = (bpath->bbpos+bpath->bepos)/2;
          apath->tlen  = 0;
          bpath->tlen  = 0;
        }
      else
        { low  = apath->abpos - apath->bbpos;
          anti = apath->abpos + apath->bbpos;
          apath->tlen = bpath->tlen = 0;
          forward_wave(work,spec,aseq,bseq,apath,bpath,&low,low,anti,aoff,boff);
        }
    }
  else
    { if (rshort)
        { low  = apath->aepos - apath->bepos;
          anti = apath->aepos + apath->bepos;
          apath->tlen = bpath->tlen = 0;
          apath->diffs = 0;
          reverse_wave(work,spec,aseq,bseq,apath,bpath,low,low,anti,aoff,boff);
        }
    }

  bpath->diffs = apath->diffs;
  if (acomp)                                                 /* align.c:1858-1884 */
    { uint16_t *trace = apath->trace;
      uint16_t  p;
      int       i, j;

      bpath->aepos = apath->bepos;
      bpath->bepos = apath->aepos;
      bpath->abpos = apath->bbpos;
      bpath->bbpos = apath->abpos;

      apath->abpos = alen - bpath->bepos;
      apath->bbpos = blen - bpath->aepos;
      apath->aepos = alen - bpath->bbpos;
      apath->bepos = blen - bpath->abpos;
      i = apath->tlen-2;
      j = 0;
      while (j < i)
        { p = trace[i];
          trace[i] = trace[j];
          trace[j] = p;
          p = trace[i+1];
          trace[i+1] = trace[j+1];
          trace[j+1] = p;
          i -= 2;
          j += 2;
        }
    }
  else                                                       /* align.c:1907-1912 */
    { bpath->aepos = apath->bepos;
      bpath->bepos = apath->aepos;
      bpath->abpos = apath->bbpos;
      bpath->bbpos = apath->abpos;
    }
}

int orc_local_alignment(const uint8_t *aseq, int alen, const uint8_t *bseq, int blen, int acomp,
                        int dg, int ad, int spacing, int ave_path,
                        const int16_t *score, const int16_t *table,
                        int apath[6], uint16_t *atrace, int bpath[6], uint16_t *btrace, int tcap)
{ orc_work *w = orc_work_new();
  orc_aspec spec;
  orc_path  ap, bp;
  int       rc = 0;

  spec.spacing = spacing; spec.ave_path = ave_path; spec.score = score; spec.table = table;
  memset(&ap,0,sizeof(ap)); memset(&bp,0,sizeof(bp));
  orc_local_align(w,&spec,aseq,alen,bseq,blen,acomp,dg,dg,ad,&ap,&bp);
  apath[0] = ap.abpos; apath[1] = ap.bbpos; apath[2] = ap.aepos; apath[3] = ap.bepos;
  apath[4] = ap.diffs; apath[5] = ap.tlen;
  bpath[0] = bp.abpos; bpath[1] = bp.bbpos; bpath[2] = bp.aepos; bpath[3] = bp.bepos;
  bpath[4] = bp.diffs; bpath[5] = bp.tlen;
  if (ap.tlen > tcap || bp.tlen > tcap)
    rc = -1;
  else
    { memcpy(atrace,ap.trace,sizeof(uint16_t)*ap.tlen);
      memcpy(btrace,bp.trace,sizeof(uint16_t)*bp.tlen);
    }
  orc_work_free(w);
  return (rc);
}
