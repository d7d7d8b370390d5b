/* TEST INFRASTRUCTURE ONLY -- core-only wall clock of the UNMODIFIED reference.
 *
 * oracle/Makefile links the reference with -Wl,--wrap=Sort_Kmers,--wrap=Match_Filter,--wrap=Reporter
 * and this file into oracle/_ref/damapper_timed: every call the driver makes through map.h:27-36
 * (damapper.c:833-875) is clocked, the sum goes to stderr at exit as "[core] <seconds> s".  This is the
 * core-only CPU figure of SURVEY.md section 8(d)(ii) without touching a reference source file. */
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include "DB.h"
#include "align.h"
#include "map.h"

void *__real_Sort_Kmers(DAZZ_DB *block, int *len);
void  __real_Match_Filter(DAZZ_DB *ablock, DAZZ_DB *bblock, void *atable, int alen,
                          void *btable, int blen, int comp, int start);
void  __real_Reporter(char *aname, DAZZ_DB *ablock, char *bname, DAZZ_DB *bblock,
                      Align_Spec *asettings, int mflag);

static double g_core = 0.;
static int    g_hooked = 0;

static double now(void)
{ struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC,&ts);
  return (ts.tv_sec + 1e-9*ts.tv_nsec);
}

static void report(void) { fprintf(stderr,"[core] %.6f s\n",g_core); }
static void hook(void)   { if (!g_hooked) { atexit(report); g_hooked = 1; } }

void *__wrap_Sort_Kmers(DAZZ_DB *block, int *len)
{ double t = now(); void *r;
  hook();
  r = __real_Sort_Kmers(block,len);
  g_core += now()-t;
  return (r);
}

void __wrap_Match_Filter(DAZZ_DB *ablock, DAZZ_DB *bblock, void *atable, int alen,
                         void *btable, int blen, int comp, int start)
{ double t = now();
  hook();
  __real_Match_Filter(ablock,bblock,atable,alen,btable,blen,comp,start);
  g_core += now()-t;
}

void __wrap_Reporter(char *aname, DAZZ_DB *ablock, char *bname, DAZZ_DB *bblock,
                     Align_Spec *asettings, int mflag)
{ double t = now();
  hook();
  __real_Reporter(aname,ablock,bname,bblock,asettings,mflag);
  g_core += now()-t;
}
