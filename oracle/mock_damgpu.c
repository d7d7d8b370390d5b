/* TEST INFRASTRUCTURE ONLY -- a CPU stand-in for libdamgpu.so, built on the oracle.
 *
 * tests/test_host_driver_cpu.py compiles the host driver's own sources (damapper_b200/host/damapper.c,
 * dazz_db.c, las_post.c) against THIS file instead of the CUDA library, so that the driver's logic -- the
 * reads-block loop (damapper.c:825-914 of the reference), the prefetch thread, packed block loading, mask
 * tracks, the resident reference blocks, -G worker processes, the sort directory, LAsort/LAcat or the
 * built-in stand-in -- runs in the CPU test suite against the unmodified reference binary.  It implements
 * exactly the entry points of include/libdamgpu.h that host/damapper.c calls, each by handing the work to
 * the oracle (oracle/damapper_oracle.h).  Nothing in the product builds, links or loads this file: the shipped
 * driver links libdamgpu.so, which has no CPU path (tests/test_abi.py::test_no_cpu_fallback). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#include "../include/libdamgpu.h"
#include "damapper_oracle.h"
#include "orc_internal.h"

struct damgpu_dblock
  { uint8_t  *raw;                 /* raw[0] = 4, bases = raw+1 (Load_All_Reads image, DB.c:1389-1441) */
    int64_t  *boff;
    int32_t  *rlen;
    int64_t  *moff;
    int32_t  *mpts;
    orc_block b;
  };

struct damgpu_index
  { orc_kmer *list;                /* snapshot of the block's sorted list at build time */
    int       len;
  };

struct damgpu_mapper
  { orc_mapper *m; };

struct damgpu_report
  { const uint8_t *buf[3];         /* owned by the oracle mapper */
    int64_t        len[3], nrec[2];
    int            nreads, tfirst;       /* records carry the read's index in the whole DB (map.c:2540,2558) */
  };

static damgpu_options g_opt = { 0, 0, 100, 1.0, "/tmp", 0, 0 };
static int    g_kmer = 0, g_suppress = 0;
static void (*g_fatal)(int) = NULL;

static void *need(void *p)
{ if (p == NULL)
    { fprintf(stderr,"mock libdamgpu: out of memory\n");
      if (g_fatal != NULL) g_fatal(1);
      exit (1);
    }
  return (p);
}

int damgpu_init(int device) { (void) device; return (0); }
const char *damgpu_last_error(void) { return (""); }
void damgpu_set_options(const damgpu_options *o) { g_opt = *o; }
void damgpu_set_fatal(void (*f)(int)) { g_fatal = f; }

int damgpu_device_memory(uint64_t *free_bytes, uint64_t *total_bytes)
{ *free_bytes = *total_bytes = (uint64_t) 8 << 30;
  return (0);
}

int damgpu_Set_Filter_Params(int kmer, int suppress, int nthreads)
{ (void) nthreads;
  if (kmer <= 1)
    return (1);
  g_kmer = kmer; g_suppress = suppress;
  return (0);
}

damgpu_dblock *damgpu_block_upload_packed(const damgpu_block *v, const uint8_t *packed,
                                          const int64_t *poff, int64_t packed_bytes)
{ damgpu_dblock *d = (damgpu_dblock *) need(calloc(1,sizeof(damgpu_dblock)));
  const int n = v->nreads;
  int i, j;
  (void) packed_bytes;
  d->boff = (int64_t *) need(malloc(sizeof(int64_t)*(n+1)));
  d->rlen = (int32_t *) need(malloc(sizeof(int32_t)*(n+1)));
  memcpy(d->boff,v->boff,sizeof(int64_t)*(n+1));
  memcpy(d->rlen,v->rlen,sizeof(int32_t)*n);
  d->raw = (uint8_t *) need(malloc((size_t) d->boff[n]+8));
  d->raw[0] = 4;
  for (i = 0; i < n; i++)                         /* Uncompress_Read, DB.c:342-363 */
    { const uint8_t *s = packed + poff[i];
      uint8_t *t = d->raw + 1 + d->boff[i];
      for (j = 0; j < d->rlen[i]; j++)
        t[j] = (uint8_t) ((s[j >> 2] >> (6 - 2*(j & 3))) & 3);
      t[d->rlen[i]] = 4;
    }
  if (v->mask_off != NULL)
    { const int64_t np = v->mask_off[n];
      d->moff = (int64_t *) need(malloc(sizeof(int64_t)*(n+1)));
      d->mpts = (int32_t *) need(malloc(sizeof(int32_t)*(np+1)));
      memcpy(d->moff,v->mask_off,sizeof(int64_t)*(n+1));
      memcpy(d->mpts,v->mask_pts,sizeof(int32_t)*np);
    }
  d->b.bases = d->raw + 1; d->b.boff = d->boff; d->b.rlen = d->rlen;
  d->b.nreads = n; d->b.tfirst = v->tfirst; d->b.maxlen = v->maxlen; d->b.totlen = v->totlen;
  d->b.sizeof_db = v->sizeof_db;
  d->b.mask_off = d->moff; d->b.mask_pts = d->mpts;
  return (d);
}

void damgpu_block_free(damgpu_dblock *d)
{ if (d == NULL) return;
  free(d->raw); free(d->boff); free(d->rlen); free(d->moff); free(d->mpts);
  free(d);
}

/* complement_DB(block, inplace), damapper.c:433-522: every read reversed and complemented, the mask's
 * point list of a read reversed with x -> rlen - x */
void damgpu_block_complement(damgpu_dblock *d)
{ int i;
  for (i = 0; i < d->b.nreads; i++)
    { uint8_t *s = d->raw + 1 + d->boff[i];
      int a = 0, e = d->rlen[i]-1;
      while (a < e)
        { uint8_t x = (uint8_t) (3 - s[a]);
          s[a] = (uint8_t) (3 - s[e]); s[e] = x;
          a += 1; e -= 1;
        }
      if (a == e)
        s[a] = (uint8_t) (3 - s[a]);
      if (d->moff != NULL)
        { int64_t lo = d->moff[i], hi = d->moff[i+1]-1;
          const int L = d->rlen[i];
          while (lo < hi)
            { int32_t x = L - d->mpts[lo];
              d->mpts[lo] = L - d->mpts[hi]; d->mpts[hi] = x;
              lo += 1; hi -= 1;
            }
          if (lo == hi)
            d->mpts[lo] = L - d->mpts[lo];
        }
    }
}

damgpu_index *damgpu_index_build(const damgpu_dblock *d)
{ damgpu_index *x = (damgpu_index *) need(calloc(1,sizeof(damgpu_index)));
  if (g_kmer <= 1)
    { fprintf(stderr,"mock libdamgpu: Sort_Kmers called before Set_Filter_Params\n");
      if (g_fatal != NULL) g_fatal(1);
      exit (1);
    }
  x->list = orc_sort_kmers(&d->b,g_kmer,g_suppress,&x->len);
  return (x);
}

damgpu_index *damgpu_index_build_deferred(const damgpu_dblock *d)   /* the reads list lives in the oracle mapper */
{ (void) d;
  return ((damgpu_index *) need(calloc(1,sizeof(damgpu_index))));
}

void damgpu_index_free(damgpu_index *x)
{ if (x == NULL) return;
  free(x->list);
  free(x);
}

damgpu_mapper *damgpu_mapper_new(const damgpu_dblock *reads, const damgpu_index *reads_idx)
{ damgpu_mapper *h = (damgpu_mapper *) need(calloc(1,sizeof(damgpu_mapper)));
  orc_params p;
  (void) reads_idx;
  memset(&p,0,sizeof(p));
  p.kmer = g_kmer; p.suppress = g_suppress; p.spacing = g_opt.spacing; p.profile = g_opt.profile;
  p.best_tie = g_opt.best_tie; p.mem_limit = g_opt.mem_limit;
  p.ave_corr = .85; p.do_a = 1;                   /* the Reporter call brings the real ones */
  h->m = orc_mapper_new(&p,&reads->b);
  return (h);
}

void damgpu_mapper_free(damgpu_mapper *h)
{ if (h == NULL) return;
  orc_mapper_free(h->m);
  free(h);
}

/* Match_Filter from resident handles: orc_match_filter with the reference list taken from the index
 * handle (the driver keeps a block's two lists while the block itself sits complemented) */
void damgpu_mapper_match(damgpu_mapper *h, const damgpu_dblock *ref, const damgpu_index *ref_idx,
                         int comp, int start)
{ orc_mapper *m = h->m;
  int64_t nhits;
  int     limit;
  orc_seed *seeds;
  if (ref_idx == NULL || ref_idx->len == 0 || m->blen == 0)      /* map.c:2955-2956 */
    return;
  seeds = orc_merge_join(m->bidx,m->blen,ref_idx->list,ref_idx->len,m->par.mem_limit,
                         m->reads.sizeof_db,ref->b.sizeof_db,
                         m->reads.maxlen,m->reads.nreads,ref->b.nreads,&nhits,&limit,NULL);
  m->last_nhits = nhits;
  m->last_limit = limit;
  orc_chain_seeds(m,seeds,nhits,ref->b.tfirst,comp,start);
  free(seeds);
}

damgpu_report *damgpu_mapper_report(damgpu_mapper *h, const damgpu_dblock *wholeref,
                                    const damgpu_align_spec *spec, int mflag)
{ damgpu_report *r = (damgpu_report *) need(calloc(1,sizeof(damgpu_report)));
  orc_mapper *m = h->m;
  m->par.ave_corr = spec->ave_corr;
  memcpy(m->par.freq,spec->freq,sizeof(m->par.freq));
  m->par.do_a = (mflag & 1) != 0;
  m->par.do_b = (mflag & 2) != 0;
  orc_report(m,&wholeref->b,&r->buf[0],&r->len[0],&r->nrec[0],&r->buf[1],&r->len[1],&r->nrec[1],
             &r->buf[2],&r->len[2]);
  r->nreads = m->reads.nreads; r->tfirst = m->reads.tfirst;
  return (r);
}

void damgpu_report_free(damgpu_report *r) { free(r); }

int64_t damgpu_report_records(const damgpu_report *r, int family)
{ return (family == 0 ? r->nrec[0] : family == 1 ? r->nrec[1] : 0); }

/* per-"thread" files: reads [(i*n)>>shift, ((i+1)*n)>>shift), map.c:3148,3250-3261; the record stream is in
 * read order, the read of a record is aread in the M family and bread in the R family */
int damgpu_report_write_las(const damgpu_report *r, int family, const char *dir, const char *aname,
                            const char *bname, int nfiles, int tspace)
{ const uint8_t *p = r->buf[family], *end = p + r->len[family];
  const int tbytes = (tspace <= 125) ? 1 : 2;            /* TRACE_XOVR, align.h:45 */
  const int64_t n = r->nreads;
  int shift = 0, nf, i;
  while ((2 << shift) <= nfiles) shift++;
  nf = 1 << shift;
  for (i = 0; i < nf; i++)
    { const int64_t r1 = (i == nf-1) ? n : ((((int64_t) i+1)*n) >> shift);
      const uint8_t *q = p;
      int64_t novl = 0;
      char  path[4096];
      FILE *f;
      while (q < end)
        { int32_t h[10];
          memcpy(h,q,40);
          if ((family == 0 ? h[7] : h[8]) - r->tfirst >= r1)
            break;
          q += 40 + (int64_t) h[0]*tbytes;
          novl += 1;
        }
      if (family == 0) snprintf(path,sizeof(path),"%s/%s.%s.M%d.las",dir,aname,bname,i+1);
      else             snprintf(path,sizeof(path),"%s/%s.%s.R%d.las",dir,bname,aname,i+1);
      f = fopen(path,"w");
      if (f == NULL)
        return (1);
      fwrite(&novl,sizeof(int64_t),1,f);
      fwrite(&tspace,sizeof(int),1,f);
      if (q > p)
        fwrite(p,1,(size_t) (q-p),f);
      if (fclose(f) != 0)
        return (1);
      p = q;
    }
  return (0);
}

/* ./.<aname>.prof.anno/.data, map.c:3295-3318 */
int damgpu_report_write_profile(const damgpu_report *r, const damgpu_block *reads, const char *dir,
                                const char *aname, int tspace)
{ char path[4096];
  FILE *af, *df;
  int   size = sizeof(int64_t), a;
  int64_t cnt = 0;
  snprintf(path,sizeof(path),"%s/.%s.prof.anno",dir,aname);
  af = fopen(path,"w");
  snprintf(path,sizeof(path),"%s/.%s.prof.data",dir,aname);
  df = fopen(path,"w");
  if (af == NULL || df == NULL)
    { if (af != NULL) fclose(af);
      if (df != NULL) fclose(df);
      return (1);
    }
  fwrite(&reads->nreads,sizeof(int),1,af);
  fwrite(&size,sizeof(int),1,af);
  for (a = 0; a < reads->nreads; a++)
    { fwrite(&cnt,sizeof(int64_t),1,af);
      cnt += (reads->rlen[a]-1)/tspace + 2;
    }
  fwrite(&cnt,sizeof(int64_t),1,af);
  if (r->len[2] > 0)
    fwrite(r->buf[2],1,(size_t) r->len[2],df);
  fclose(af); fclose(df);
  return (0);
}
