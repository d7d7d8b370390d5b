/* TEST INFRASTRUCTURE ONLY -- oracle restatement of sparse k-mer chaining and the candidate
 * filter (chain_thread, map.c:1463-1922).  The reference's splay tree is replaced by a sorted
 * array with the same key (diag descending, apos descending): the result does not depend on
 * tree shape (verified against TEST_CHAIN dumps). */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "orc_internal.h"

#define HITMIN    3      /* map.c:34 */
#define MAX_GAP   1000   /* map.c:36 */
#define MIN_PIECE 300    /* map.c:37 */

typedef struct
  { int apos, bpos, diag;
    int from;          /* index of predecessor node or -1 (Splay.from) */
    int orig;          /* Splay.orig: origin of the chain; for an origin, its best end */
    int cost;
    int dead;          /* Splay.bmin == -1 */
  } Node;

orc_mapper *orc_mapper_new(const orc_params *p, const orc_block *reads)
{ orc_mapper *m = (orc_mapper *) calloc(1,sizeof(orc_mapper));
  int64_t tot = 0;
  int i;

  m->par   = *p;
  m->reads = *reads;
  m->head  = (int *) malloc(sizeof(int)*(reads->nreads+1));
  m->coff  = (int64_t *) malloc(sizeof(int64_t)*(reads->nreads+1));
  for (i = 0; i < reads->nreads; i++)
    { m->head[i] = -1;
      m->coff[i] = tot;
      tot += (reads->rlen[i]-1)/p->spacing + 2;
    }
  m->coff[reads->nreads] = tot;
  m->cover = (int16_t *) calloc(tot+1,sizeof(int16_t));
  m->bidx  = orc_sort_kmers(reads,p->kmer,p->suppress,&m->blen);
  return (m);
}

void orc_mapper_free(orc_mapper *m)
{ int i;
  if (m == NULL) return;
  for (i = 0; i < m->ncand; i++)
    free(m->cand[i].jumps);
  free(m->cand); free(m->head); free(m->coff); free(m->cover); free(m->bidx);
  free(m->abuf); free(m->bbuf); free(m->prof);
  free(m);
}

static int new_cand(orc_mapper *m)
{ if (m->ncand >= m->cmax)
    { m->cmax = (int) (1.5*m->cmax) + 1024;
      m->cand = (orc_cand *) realloc(m->cand,sizeof(orc_cand)*m->cmax);
    }
  memset(m->cand+m->ncand,0,sizeof(orc_cand));
  return (m->ncand++);
}

/* chain_length, map.c:1243-1260: splices same-diagonal predecessors closer than 100 out of
 * the from-chain (the mutation is kept, later candidates of the group see it). */
static int chain_length(Node *nd, int x)
{ int y, n = 0, da;
  y = nd[x].from;
  while (y >= 0)
    { da = nd[x].apos - nd[y].apos;
      if (da == nd[x].bpos - nd[y].bpos && da < 100)
        y = nd[x].from = nd[y].from;
      else
        { n += 1;
          x = y;
          y = nd[x].from;
        }
    }
  return (n);
}

/* Candidate test + dominance filter + Jump list for chain end h, map.c:1642-1767 */
static void consider(orc_mapper *m, Node *nd, int h, int ar, int br, int comp)
{ int K = m->par.kmer, S = m->par.spacing;
  int ab, bb, ae, be, c, d, e;

  ab = nd[nd[h].orig].apos - K;
  bb = nd[nd[h].orig].bpos - K;
  ae = nd[h].apos;
  be = nd[h].bpos;

  if (m->par.profile)
    { int16_t *cnt = m->cover + m->coff[ar];
      int tb = ab/S, te = (ae-1)/S+1;
      int cb = cnt[tb], ce = cnt[te];
      if (cb < 0x7fff && ce > -0xffff)       /* sic, map.c:1662 */
        { cnt[tb] = (int16_t) (cb+1);
          cnt[te] = (int16_t) (ce-1);
        }
    }

  c = -1;
  for (d = m->head[ar]; d >= 0; d = e)
    { orc_cand *D = m->cand+d;
      int A = (D->c.afirst < ab+MIN_PIECE && D->c.alast > ae-MIN_PIECE);
      int B = (ab < D->c.afirst+MIN_PIECE && ae > D->c.alast-MIN_PIECE);
      e = D->next;
      if (A && .9 * D->c.score >= nd[h].cost)
        break;
      if (B && D->c.score <= .9 * nd[h].cost)
        { if (c < 0)                         /* DEL_CELL, map.c:1538-1545 */
            m->head[ar] = e;
          else
            m->cand[c].next = e;
          D->next = -2;                      /* freed */
          free(D->jumps); D->jumps = NULL;
          m->nlive -= 1;
        }
      else
        c = d;
    }
  if (d >= 0)
    return;

  d = new_cand(m);
  { orc_cand *D = m->cand+d;
    int len, f, g, n;
    D->next = m->head[ar];
    m->head[ar] = d;
    D->c.read   = ar;
    D->c.bread  = br;
    D->c.comp   = comp;
    D->c.score  = nd[h].cost;
    D->c.afirst = ab;
    D->c.alast  = ae;
    D->c.bfirst = bb;
    D->c.blast  = be;
    D->c.length = len = chain_length(nd,h);
    D->jumps = (int32_t *) malloc(sizeof(int32_t)*2*(len+1));
    g = h; n = 0;
    for (f = nd[h].from; f >= 0; f = nd[f].from)
      { D->jumps[2*n]   = (uint16_t) (nd[g].apos - nd[f].apos);   /* uint16 fields, map.c:1382 */
        D->jumps[2*n+1] = (uint16_t) (nd[g].bpos - nd[f].bpos);
        n += 1;
        g = f;
      }
    if (n != len)
      { fprintf(stderr,"oracle: chain length mismatch\n"); exit (1); }
    m->nlive += 1;
  }
}

/* chain_thread, map.c:1463-1922, for all reads (one "thread"). */
void orc_chain_seeds(orc_mapper *m, const orc_seed *hits, int64_t nhits, int bstart,
                     int comp, int start)
{ int     K = m->par.kmer;
  int     hithr = HITMIN*K;
  int64_t nidx, f;
  int     maxk = 0, i;
  Node   *nd;
  int    *S, *expired;

  if (start)                                   /* map.c:1574-1588, INIT_FREE_SPACE */
    { for (i = 0; i < m->ncand; i++)
        { free(m->cand[i].jumps); m->cand[i].jumps = NULL; }
      m->ncand = 0;
      m->nlive = 0;
      for (i = 0; i < m->reads.nreads; i++)
        m->head[i] = -1;
      memset(m->cover,0,sizeof(int16_t)*m->coff[m->reads.nreads]);
    }
  if (nhits == 0)
    return;

  for (nidx = 0; nidx < nhits; )               /* largest (aread,bread) group, map.c:1554-1566 */
    { f = nidx++;
      while (nidx < nhits && hits[nidx].aread == hits[f].aread && hits[nidx].bread == hits[f].bread)
        nidx += 1;
      if (nidx-f > maxk)
        maxk = (int) (nidx-f);
    }
  nd = (Node *) malloc(sizeof(Node)*(maxk+1));
  S  = (int *) malloc(sizeof(int)*(maxk+1));
  expired = (int *) malloc(sizeof(int)*(maxk+1));

  nidx = 0;
  while (nidx < nhits)
    { int ar = hits[nidx].aread, br = hits[nidx].bread;
      int nn = 0, ns = 0, nexp = 0, qhead = 0;

      for ( ; nidx < nhits && hits[nidx].aread == ar && hits[nidx].bread == br; nidx++)
        { int apos = hits[nidx].apos + 1;                       /* map.c:1784-1785 */
          int bpos = apos - hits[nidx].diag;
          int diag = apos - bpos;
          int n, pos, l, r, lcost, rcost, j;

          while (qhead < nn && nd[qhead].apos < apos-MAX_GAP)   /* map.c:1787-1796 */
            { int q = qhead++;
              if ( ! nd[q].dead)
                { for (j = 0; j < ns; j++)
                    if (S[j] == q) break;
                  memmove(S+j,S+j+1,sizeof(int)*(ns-j-1));
                  ns -= 1;
                  if (nd[nd[q].orig].orig == q)
                    expired[nexp++] = q;      /* pushed at the FRONT in the reference */
                }
            }

          n = nn++;
          nd[n].apos = apos; nd[n].bpos = bpos; nd[n].diag = diag; nd[n].dead = 0;

          for (pos = 0; pos < ns; pos++)     /* key order: diag desc, apos desc (add, map.c:1101) */
            { Node *x = nd+S[pos];
              if (diag > x->diag || (diag == x->diag && apos > x->apos))
                break;
            }
          memmove(S+pos+1,S+pos,sizeof(int)*(ns-pos));
          S[pos] = n;
          ns += 1;

          l = -1;                             /* predOf + leftmost, map.c:1806-1808 */
          for (j = pos-1; j >= 0; j--)
            if (nd[S[j]].bpos >= bpos-MAX_GAP)
              { l = S[j];
                while (j > 0 && nd[S[j-1]].diag == nd[l].diag)
                  l = S[--j];
                break;
              }
          r = -1;                             /* succOf, map.c:1809 */
          for (j = pos+1; j < ns; j++)
            if (nd[S[j]].bpos <= bpos)
              { r = S[j];
                break;
              }

          lcost = rcost = 0;                  /* map.c:1810-1826 */
          if (l >= 0)
            lcost = nd[l].cost + ((apos >= nd[l].apos+K) ? K : apos - nd[l].apos);
          if (r >= 0)
            rcost = nd[r].cost + ((bpos >= nd[r].bpos+K) ? K : bpos - nd[r].bpos);
          if (lcost > rcost)
            rcost = 0;
          else
            lcost = 0;

          if (lcost > 0 || rcost > 0)         /* map.c:1828-1857 */
            { int p = (lcost > 0) ? l : r;
              int c = (lcost > 0) ? lcost : rcost;
              int o;
              nd[n].from = p;
              nd[n].cost = c;
              o = nd[n].orig = (nd[p].from < 0) ? p : nd[p].orig;
              if (c >= nd[nd[o].orig].cost)
                { int dd = nd[p].diag - nd[n].diag;
                  nd[o].orig = n;
                  if (dd < 0) dd = -dd;
                  if (dd <= .2*(nd[n].apos - nd[p].apos))
                    { for (j = 0; j < ns; j++)
                        if (S[j] == p) break;
                      memmove(S+j,S+j+1,sizeof(int)*(ns-j-1));
                      ns -= 1;
                      nd[p].dead = 1;
                    }
                }
            }
          else
            { nd[n].from = -1;
              nd[n].cost = K;
              nd[n].orig = n;
            }
        }

      { int j;                                  /* map.c:1634-1767 */
        for (j = 0; j < ns; j++)
          { int h = S[j];
            if (nd[h].cost >= hithr && nd[nd[h].orig].orig == h)
              consider(m,nd,h,ar,br+bstart,comp);
          }
        for (j = nexp-1; j >= 0; j--)
          { int h = expired[j];
            if (nd[h].cost >= hithr && nd[nd[h].orig].orig == h)
              consider(m,nd,h,ar,br+bstart,comp);
          }
      }
    }

  free(expired);
  free(S);
  free(nd);
}

/* Match_Filter, map.c:2889-3209 */
void orc_match_filter(orc_mapper *m, const orc_block *ref, int comp, int start)
{ int       alen;
  int64_t   nhits;
  int       limit;
  orc_kmer *aidx;
  orc_seed *seeds;

  aidx = orc_sort_kmers(ref,m->par.kmer,m->par.suppress,&alen);
  if (alen == 0 || m->blen == 0)                /* map.c:2955-2956 (no reset happens either) */
    { free(aidx);
      return;
    }
  seeds = orc_merge_join(m->bidx,m->blen,aidx,alen,m->par.mem_limit,
                         m->reads.sizeof_db,ref->sizeof_db,
                         m->reads.maxlen,m->reads.nreads,ref->nreads,&nhits,&limit,NULL);
  free(aidx);
  m->last_nhits = nhits;
  m->last_limit = limit;
  orc_chain_seeds(m,seeds,nhits,ref->tfirst,comp,start);
  free(seeds);
}

int64_t orc_num_candidates(const orc_mapper *m)
{ return (m->nlive); }

int64_t orc_get_candidates(const orc_mapper *m, orc_candidate *out, int32_t *jcnt,
                           int32_t *jumps, int64_t jmax)
{ int64_t n = 0, nj = 0;
  int i, c, k;
  for (i = 0; i < m->reads.nreads; i++)
    for (c = m->head[i]; c >= 0; c = m->cand[c].next)
      { out[n] = m->cand[c].c;
        jcnt[n] = m->cand[c].c.length;
        for (k = 0; k < m->cand[c].c.length; k++)
          { if (nj < jmax)
              { jumps[2*nj]   = m->cand[c].jumps[2*k];
                jumps[2*nj+1] = m->cand[c].jumps[2*k+1];
              }
            nj += 1;
          }
        n += 1;
      }
  return (nj);
}

int64_t orc_get_cover(const orc_mapper *m, int16_t *out, int64_t max)
{ int64_t tot = m->coff[m->reads.nreads];
  if (out != NULL)
    memcpy(out,m->cover,sizeof(int16_t)*(tot < max ? tot : max));
  return (tot);
}
