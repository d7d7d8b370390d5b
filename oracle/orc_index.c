/* TEST INFRASTRUCTURE ONLY -- oracle restatement of the index build and the merge-join.
 * See damapper_oracle.h.  Single-threaded, plain C; citations are into /root/reference. */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "damapper_oracle.h"

#define MAXGRAM 10000           /* map.c:32 */

/* Stable LSD radix sort, 8-bit digits, over the key bytes flagged in bytes[16]
 * (lex_sort/lex_thread, map.c:181-444, restated for one thread).  Sorts n 16-byte records
 * ping-ponging between src and trg; returns the buffer holding the result. */
static void *lsd_sort(const int bytes[16], void *vsrc, void *vtrg, int64_t n)
{ typedef struct { uint64_t p[2]; } rec;
  rec *src = (rec *) vsrc, *trg = (rec *) vtrg, *x;
  int64_t cnt[256], i, s, t;
  int b, j;

  for (b = 0; b < 16; b++)
    { if (!bytes[b]) continue;
      memset(cnt,0,sizeof(cnt));
      for (i = 0; i < n; i++)
        cnt[((const uint8_t *) (src+i))[b]] += 1;
      s = 0;
      for (j = 0; j < 256; j++)
        { t = cnt[j]; cnt[j] = s; s += t; }
      for (i = 0; i < n; i++)
        trg[cnt[((const uint8_t *) (src+i))[b]]++] = src[i];
      x = src; src = trg; trg = x;
    }
  return (src);
}

/* Sort_Kmers, map.c:655-822; tuple_thread (unmasked branch) map.c:545-576;
 * -t suppression compsize/compress_thread map.c:590-636,726-770; sentinels :772-773. */
orc_kmer *orc_sort_kmers(const orc_block *blk, int kmer, int suppress, int *len)
{ int      nreads = blk->nreads;
  int64_t  kmers64 = blk->boff[nreads] - (int64_t) kmer * nreads;
  uint64_t kmask = (kmer == 32) ? ~0ull : ((1ull << (2*kmer)) - 1);
  orc_kmer *src, *trg, *rez;
  int      mersort[16];
  int64_t  n;
  int      i, kmers;

  *len = 0;
  if (kmers64 <= 0 || kmers64 > 0x7fffffff)
    return (NULL);
  kmers = (int) kmers64;

  src = (orc_kmer *) malloc(sizeof(orc_kmer)*((size_t) kmers+2));
  trg = (orc_kmer *) malloc(sizeof(orc_kmer)*((size_t) kmers+2));
  if (src == NULL || trg == NULL)
    { fprintf(stderr,"oracle: out of memory\n"); exit (1); }

  n = 0;
  for (i = 0; i < nreads; i++)
    { const uint8_t *s = blk->bases + blk->boff[i];
      int64_t  b = 0, f = 0, a;
      if (blk->mask_off != NULL)                     /* masked branch, map.c:481-543 */
        { b = blk->mask_off[i]; f = blk->mask_off[i+1]; }
      for (a = b; a <= f; a += 2)                    /* unmasked segment [p,q) before interval a */
        { int      p = (a == b) ? 0 : blk->mask_pts[a-1];
          int      q = (a == f) ? blk->rlen[i] : blk->mask_pts[a];
          int      x;
          uint64_t c = 0;
          if (p + kmer <= q)
            { for (x = 1; x < kmer; x++)
                c = (c << 2) | s[p++];
              while (p < q)
                { c = ((c << 2) | s[p]) & kmask;
                  src[n].read = i;
                  src[n].rpos = p++;
                  src[n].code = c;
                  n += 1;
                }
            }
        }
    }
  /* reads shorter than k contribute a negative count to `kmers` in the reference; it rejects
     such blocks up front (damapper.c:403-410), so n == kmers here.  With a mask the reference pads
     the missing slots with ~0 records that sort to the end and are cut off (map.c:524-532,
     :706-724): the list is the n real k-mers. */
  if (blk->mask_off == NULL && n != kmers)
    { fprintf(stderr,"oracle: block holds reads shorter than k\n"); exit (1); }
  kmers = (int) n;
  if (kmers == 0)
    { free(src); free(trg); return (NULL); }

  for (i = 0; i < 16; i++)
    mersort[i] = 0;
  for (i = 0; i < 2*kmer; i += 8)
    mersort[i>>3] = 1;
  rez = (orc_kmer *) lsd_sort(mersort,src,trg,kmers);
  if (rez == trg)
    { trg = src; src = rez; }

  if (suppress > 0 && kmers > 0)
    { int j, p, m = 0;
      j = 0;
      while (j < kmers)
        { p = j++;
          while (j < kmers && rez[j].code == rez[p].code)
            j += 1;
          if (j-p < suppress)
            while (p < j)
              trg[m++] = rez[p++];
        }
      kmers = m;
      src = rez; rez = trg; trg = src;
    }

  rez[kmers].code   = 0xffffffffffffffffull;
  rez[kmers].rpos   = 0; rez[kmers].read = 0;
  rez[kmers+1].code = 0;
  rez[kmers+1].rpos = 0; rez[kmers+1].read = 0;
  free(trg);

  if (kmers <= 0)
    { free(rez);
      return (NULL);
    }
  *len = kmers;
  return (rez);
}

/* `limit` from the run-product histogram, map.c:2992-3015.  0 = no cap (-M0) -> INT32_MAX */
static int compute_limit(const int64_t *histo, uint64_t mem_limit, int64_t asize, int64_t bsize,
                         int alen, int blen)
{ int64_t tom, avail;
  int     j;

  if (mem_limit == 0)
    return (0x7fffffff);
  avail = (int64_t) ((uint64_t) (int64_t) (mem_limit - (uint64_t) (asize + bsize)) / 16u);
  if (avail > alen + 2*(int64_t) blen)
    avail = (avail - alen) / 2;
  else
    avail = avail - (alen + (int64_t) blen);
  avail = (int64_t) (avail * .98);
  tom = 0;
  for (j = 0; j < MAXGRAM; j++)
    { tom += j*histo[j];
      if (tom > avail)
        break;
    }
  return (j);
}

/* count_thread (map.c:881-934), limit (:2992-3055), merge_thread (:939-1002) for one thread
 * (results are thread-count invariant), then the seed sort on the
 * pairsort bytes (:2917-2936, :3121) and the sentinel (:3123-3126). */
orc_seed *orc_merge_join(const orc_kmer *asort, int alen, const orc_kmer *bsort, int blen,
                         uint64_t mem_limit, int64_t asize, int64_t bsize,
                         int amaxlen, int anreads, int bnreads,
                         int64_t *nhits_out, int *limit_out, int64_t *histo_out)
{ int64_t *gram = (int64_t *) calloc(MAXGRAM,sizeof(int64_t));
  int64_t  nhits, ct, powr;
  int      ia, ib, ja, jb, limit, pass, a, b, i, nbyte;
  int      pairsort[16];
  orc_seed *hits, *work, *rez;

  *nhits_out = 0;
  if (limit_out) *limit_out = 0;
  if (alen == 0 || blen == 0)
    { free(gram);
      return (NULL);
    }

  hits = NULL;
  nhits = 0;
  limit = 0;
  for (pass = 0; pass < 2; pass++)
    { int64_t n = 0;
      ia = ib = 0;
      while (ia < alen && ib < blen)
        { uint64_t ca = asort[ia].code, cb = bsort[ib].code;
          if (cb < ca)
            ib += 1;
          else if (cb > ca)
            ia += 1;
          else
            { ja = ia++;
              while (ia < alen && asort[ia].code == ca)
                ia += 1;
              jb = ib++;
              while (ib < blen && bsort[ib].code == cb)
                ib += 1;
              ct = ((int64_t) (ia-ja)) * (ib-jb);
              if (pass == 0)
                { if (ct < MAXGRAM)
                    gram[ct] += 1;
                }
              else if (ct < limit)
                { for (a = ja; a < ia; a++)
                    { int ap = asort[a].rpos;
                      for (b = jb; b < ib; b++)
                        { hits[n].bread = bsort[b].read;
                          hits[n].aread = asort[a].read;
                          hits[n].apos  = ap;
                          hits[n].diag  = ap - bsort[b].rpos;
                          n += 1;
                        }
                    }
                }
            }
        }
      if (pass == 0)
        { limit = compute_limit(gram,mem_limit,asize,bsize,alen,blen);
          if (mem_limit > 0)
            { for (i = 1; i < limit; i++)
                nhits += i*gram[i];
            }
          else
            { /* -M0: every run pair is kept; recount without the MAXGRAM cap (map.c:922) */
              int xa = 0, xb = 0;
              while (xa < alen && xb < blen)
                { uint64_t ca = asort[xa].code, cb = bsort[xb].code;
                  if (cb < ca) xb += 1;
                  else if (cb > ca) xa += 1;
                  else
                    { int ya = xa, yb = xb;
                      while (xa < alen && asort[xa].code == ca) xa += 1;
                      while (xb < blen && bsort[xb].code == cb) xb += 1;
                      nhits += ((int64_t) (xa-ya)) * (xb-yb);
                    }
                }
            }
          hits = (orc_seed *) malloc(sizeof(orc_seed)*((size_t) nhits+1));
          if (hits == NULL)
            { fprintf(stderr,"oracle: out of memory\n"); exit (1); }
        }
      else if (n != nhits)
        { fprintf(stderr,"oracle: hit count mismatch %lld vs %lld\n",(long long) n,(long long) nhits);
          exit (1);
        }
    }

  for (i = 0; i < 16; i++)
    pairsort[i] = 0;
  powr = 1;
  for (nbyte = 0; powr < amaxlen; nbyte += 1)
    powr <<= 8;
  for (i = 4; i < 4+nbyte; i++)
    pairsort[i] = 1;
  powr = 1;
  for (nbyte = 0; powr < bnreads; nbyte += 1)
    powr <<= 8;
  for (i = 8; i < 8+nbyte; i++)
    pairsort[i] = 1;
  powr = 1;
  for (nbyte = 0; powr < anreads; nbyte += 1)
    powr <<= 8;
  for (i = 12; i < 12+nbyte; i++)
    pairsort[i] = 1;

  work = (orc_seed *) malloc(sizeof(orc_seed)*((size_t) nhits+1));
  rez  = (orc_seed *) lsd_sort(pairsort,hits,work,nhits);
  if (rez == work)
    free(hits);
  else
    free(work);
  rez[nhits].aread = 0x7fffffff;
  rez[nhits].bread = 0x7fffffff;
  rez[nhits].diag  = 0x7fffffff;
  rez[nhits].apos  = 0;

  if (histo_out)
    memcpy(histo_out,gram,sizeof(int64_t)*MAXGRAM);
  free(gram);
  *nhits_out = nhits;
  if (limit_out) *limit_out = limit;
  return (rez);
}
