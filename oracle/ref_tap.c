/* TEST INFRASTRUCTURE ONLY -- never linked into the product.
 *
 * Tap shim around the UNMODIFIED reference sources.  It is compiled by oracle/Makefile with
 * -I$(REF) (REF=/root/reference) into oracle/_ref/libref_tap.so; no reference source is copied
 * into this repository -- the two #include lines below pull the reference translation units
 * from where they lie.  Only tests/ and the golden-vector generator (oracle/make_golden.py) load
 * the result.
 *
 * What it adds on top of the reference:
 *   - the driver (damapper.c) is included with main renamed, so that its globals
 *     (VERBOSE, PROFILE, SPACING, BEST_TIE, SORT_PATH, MEM_LIMIT, MEM_PHYSICAL), Clean_Exit,
 *     read_DB and complement_DB are available to a ctypes harness;
 *   - pthread_create is intercepted inside map.c so that the sorted seed array handed to
 *     chain_thread (map.c:3142,3166) can be snapshotted in binary instead of through the
 *     TEST_CSORT text dump (map.c:3128-3134);
 *   - accessors that flatten the per-read candidate lists and Jump chains kept in the
 *     file-static parmr[] (map.c:1386-1397,1441-1461,2885).
 */
#define main ref_damapper_main
#define complement drv_complement      /* damapper.c:417 and map.c:1940 both define one */
#include "damapper.c"
#undef complement
#undef main

#include <pthread.h>

static int tap_pthread_create(pthread_t *t, const pthread_attr_t *a, void *(*f)(void *), void *arg);
#define pthread_create(t,a,f,arg) tap_pthread_create(t,a,f,arg)
#include "map.c"
#undef pthread_create

static SeedPair *tap_seeds = NULL;
static int64     tap_nseeds = 0;
static int       tap_want_seeds = 0;

static int tap_pthread_create(pthread_t *t, const pthread_attr_t *a, void *(*f)(void *), void *arg)
{ if (tap_want_seeds && f == chain_thread && arg == (void *) parmr)
    { int64 n = parmr[NTHREADS-1].hend;
      free(tap_seeds);
      tap_seeds  = (SeedPair *) malloc(sizeof(SeedPair)*(n+1));
      memcpy(tap_seeds,MR_hits,sizeof(SeedPair)*(n+1));
      tap_nseeds = n;
    }
  return pthread_create(t,a,f,arg);
}

void tap_enable_seeds(int on) { tap_want_seeds = on; }

void *tap_get_seeds(int64 *n) { *n = tap_nseeds; return tap_seeds; }

int tap_nthreads(void) { return NTHREADS; }

/* Count candidates of read block `ablock` (walks reads[i].coff, map.c:1875). */
int64 tap_count_candidates(DAZZ_DB *ablock)
{ int64 n = 0;
  int   i, t, c;
  for (t = 0; t < NTHREADS; t++)
    for (i = parmr[t].abeg; i < parmr[t].aend; i++)
      for (c = ablock->reads[i].coff; c >= 0; c = parmr[t].cbase[c].next)
        n += 1;
  return n;
}

/* Flatten candidates in (read, list order): 9 ints per candidate
 *   read, score, length, bread, comp, afirst, alast, bfirst, blast
 * and, per candidate, its Jump displacements: `jcnt[i]` pairs appended to `jumps` as (da,db). */
int64 tap_get_candidates(DAZZ_DB *ablock, int *out, int *jcnt, int *jumps, int64 jmax)
{ int64 n = 0, nj = 0;
  int   i, t, c, j, m, k, lim;
  for (t = 0; t < NTHREADS; t++)
    for (i = parmr[t].abeg; i < parmr[t].aend; i++)
      for (c = ablock->reads[i].coff; c >= 0; c = parmr[t].cbase[c].next)
        { Candidate *cd = parmr[t].cbase + c;
          int *o = out + 9*n;
          o[0] = i; o[1] = cd->score; o[2] = cd->length; o[3] = cd->bread; o[4] = cd->comp;
          o[5] = cd->afirst; o[6] = cd->alast; o[7] = cd->bfirst; o[8] = cd->blast;
          m = cd->length;
          jcnt[n] = m;
          for (j = cd->chain; j >= 0; j = parmr[t].jbase[j].next)
            { lim = (m < 5) ? m : 5;
              for (k = 0; k < lim; k++)
                { if (nj < jmax)
                    { jumps[2*nj]   = parmr[t].jbase[j].adisp[k];
                      jumps[2*nj+1] = parmr[t].jbase[j].bdisp[k];
                    }
                  nj += 1;
                }
              m -= 5;
            }
          n += 1;
        }
  return nj;
}

/* Thin wrappers so that the harness need not know static names. */
int  tap_read_DB(DAZZ_DB *block, char *name, int kmer) { return read_DB(block,name,NULL,NULL,0,kmer); }
void tap_complement_DB(DAZZ_DB *block) { complement_DB(block,1); }
int  tap_sizeof_DAZZ_DB(void) { return (int) sizeof(DAZZ_DB); }

void tap_set_globals(int verbose, int profile, int spacing, double best_tie, char *sort_path,
                     uint64 mem_limit)
{ VERBOSE = verbose; PROFILE = profile; SPACING = spacing; BEST_TIE = best_tie;
  SORT_PATH = sort_path;
  MEM_PHYSICAL = getMemorySize();
  MEM_LIMIT = (mem_limit == (uint64) -1) ? MEM_PHYSICAL : mem_limit;
  Prog_Name = "ref_tap";
}
