/* TEST INFRASTRUCTURE ONLY -- the reference-side binding of INTEGRATION.md section 2, compiled.
 *
 * oracle/Makefile links the UNMODIFIED reference driver (/root/reference/damapper.c, DB.c, QV.c and
 * align.c for the Align_Spec accessors) with this file INSTEAD of map.c into oracle/_ref/damapper_gpu:
 * the four functions of map.h:25-36 forward to layer 1 of libdamgpu (include/libdamgpu.h).  The GPU
 * tests run it next to oracle/_ref/damapper on the same databases and compare the .las / .prof
 * bytes (tests/test_dropin.py): that is the drop-in claim, shown rather than asserted.
 *
 * damapper.c:879 frees the reads index with free(); an index lives in HBM, so Sort_Kmers returns a
 * small malloc'd stub that carries the handle.  The driver's free() releases the stub; the device
 * side of the reads index is released by Reporter (its last user, damapper.c:870-879), the
 * reference-block index by Match_Filter (which consumes it in the reference too, map.c:3181-3182).
 */
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#include "DB.h"
#include "align.h"      /* Align_Spec accessors, align.c:277-287 */
#include "map.h"
#include "libdamgpu.h"

typedef struct { uint64_t magic; void *handle; } Stub;
#define STUB_MAGIC 0x64616d6770755f31ull

static void *g_reads_index = NULL;         /* device handle of the reads index in use */
static int   g_inited = 0;

static void view(DAZZ_DB *db, damgpu_block *v, int64 **boff, int **rlen)
{ int i;
  *boff = (int64 *) Malloc(sizeof(int64)*(db->nreads+1),"boff");
  *rlen = (int *)   Malloc(sizeof(int)*(db->nreads+1),"rlen");
  for (i = 0; i <= db->nreads; i++) (*boff)[i] = db->reads[i].boff;      /* DB.h:289 */
  for (i = 0; i <  db->nreads; i++) (*rlen)[i] = db->reads[i].rlen;      /* DB.h:287 */
  memset(v,0,sizeof(*v));
  v->bases = (const uint8_t *) db->bases;  v->boff = *boff;  v->rlen = *rlen;
  v->nreads = db->nreads;  v->tfirst = db->tfirst;  v->maxlen = db->maxlen;
  v->totlen = db->totlen;  v->sizeof_db = sizeof_DB(db);                 /* DB.c:1044 */
}

static void push_globals(void)            /* map.h:16-23 -> damgpu_options */
{ damgpu_options o;
  if (!g_inited)
    { const char *dev = getenv("DAMGPU_DEVICE");
      if (damgpu_init(dev ? atoi(dev) : -1) != 0)
        { fprintf(stderr,"damapper_gpu: no usable CUDA device (%s)\n",damgpu_last_error());
          Clean_Exit(1);
        }
      g_inited = 1;
    }
  o.verbose = VERBOSE; o.profile = PROFILE; o.spacing = SPACING; o.best_tie = BEST_TIE;
  o.sort_path = SORT_PATH; o.mem_limit = MEM_LIMIT; o.mem_physical = MEM_PHYSICAL;
  damgpu_set_options(&o);
  damgpu_set_fatal(Clean_Exit);           /* map.h:39 */
}

int Set_Filter_Params(int kmer, int suppress, int nthreads)
{ return damgpu_Set_Filter_Params(kmer,suppress,nthreads); }

void *Sort_Kmers(DAZZ_DB *block, int *len)
{ damgpu_block v; int64 *bo; int *rl; void *idx; Stub *s;
  push_globals(); view(block,&v,&bo,&rl);
  if (block->tracks != NULL)               /* the merged mask, damapper.c:381-399 */
    { v.mask_off = (const int64_t *) block->tracks->anno;   /* already divided by sizeof(int) */
      v.mask_pts = (const int32_t *) block->tracks->data;
    }
  idx = damgpu_Sort_Kmers(&v,len);
  free(bo); free(rl);
  if (idx == NULL)
    return (NULL);
  s = (Stub *) Malloc(sizeof(Stub),"index stub");
  s->magic = STUB_MAGIC; s->handle = idx;
  return (s);
}

void Match_Filter(DAZZ_DB *ablock, DAZZ_DB *bblock, void *atable, int alen,
                  void *btable, int blen, int comp, int start)
{ Stub *a = (Stub *) atable, *b = (Stub *) btable;
  (void) ablock; (void) bblock;            /* both blocks travel with their index handles */
  push_globals();
  if (a != NULL) g_reads_index = a->handle;
  damgpu_Match_Filter(NULL,NULL,a ? a->handle : NULL,alen,b ? b->handle : NULL,blen,comp,start);
  free(b);                                 /* consumed, as map.c:3181-3182 */
}

void Reporter(char *aname, DAZZ_DB *ablock, char *bname, DAZZ_DB *bblock,
              Align_Spec *aspec, int mflag)
{ damgpu_block a, b; int64 *ao, *bo; int *al, *bl;
  damgpu_align_spec s;
  s.ave_corr = Average_Correlation(aspec); s.trace_space = Trace_Spacing(aspec);
  memcpy(s.freq,Base_Frequencies(aspec),sizeof(s.freq));
  push_globals(); view(ablock,&a,&ao,&al); view(bblock,&b,&bo,&bl);
  damgpu_Reporter(aname,&a,bname,&b,&s,mflag);
  ablock->maxlen += SPACING;               /* side effect of map.c:3246, kept */
  free(ao); free(al); free(bo); free(bl);
  if (g_reads_index != NULL)               /* the driver frees only the stub (damapper.c:879) */
    { damgpu_index_free((damgpu_index *) g_reads_index);
      g_reads_index = NULL;
    }
}
