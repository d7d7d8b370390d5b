/* damapper host driver for libdamgpu: keeps the reference command line
 *   damapper [-vpzCN] [-k<int(20)>] [-t<int>] [-M<int>] [-T<int(4)>] [-P<dir(/tmp)>]
 *            [-e<double(.85)] [-s<int(100)>] [-n<double(1.00)>] [-m<track>]+ <reference:dam> <reads:db> ...
 * (reference damapper.c:52-56,556-922), DAZZ_DB .db/.dam input and the per-thread .las output,
 * and calls the CUDA mapping core through the C ABI of include/libdamgpu.h.  Everything that
 * used to sit behind map.h runs on the GPU; this file only parses flags, loads blocks, and
 * runs the same LAsort/LAcat/LAmerge post-processing commands as the reference.
 *
 * The reference block is uploaded once per block, indexed, matched, reverse-complemented ON
 * THE DEVICE and indexed/matched again (the reference complements on the host and re-sorts,
 * damapper.c:847-861).  Where the reference re-reads and re-sorts every reference block for every
 * reads block (damapper.c:839-863), this driver keeps the two indices of a reference block -- and
 * the whole reference the Reporter aligns against -- resident in HBM for the reads blocks that
 * follow, as long as they fit in DAMGPU_REF_CACHE (default 45 %) of the device memory; the next
 * reads block is read from disk by a helper thread while the GPU maps the current one.
 *
 * -G<n> (not in the reference): n GPUs of the box.  Reads blocks are independent (damapper.c:825-914:
 * nothing is carried from one to the next), so the driver forks n workers BEFORE CUDA is touched, worker
 * r takes blocks r, r+n, ... of the command line on the r-th device (DAMGPU_DEVICES="3,5,..." picks
 * them; default 0..n-1) with CUDA_VISIBLE_DEVICES narrowed to it, in its own sort directory; each
 * builds the reference indices for itself (on a B200 a 250 Mbp block is indexed in ~10 ms per strand,
 * about what moving the 4 GB list over NVLink costs).  The parent only waits.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <dirent.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <time.h>
#include <pthread.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include "dazz_db.h"

static const char *Prog_Name = "damapper";
static char *SORT_PATH = "/tmp";

static const char *Usage[] =
  { "[-vpzCN] [-k<int(20)>] [-t<int>] [-M<int>] [-T<int(4)>] [-P<dir(/tmp)>] [-G<int(1)>]",
    "         [-e<double(.85)] [-s<int(100)>] [-n<double(1.00)>]",
    "         [-m<track>]+  <reference:dam> <reads:db> ...",
  };

/* DAMGPU_TIMING=1: wall clock of every phase of the driver on stderr (development aid) */
static int    TIMING = 0;
static double T_last = 0., T_zero = 0.;

static double now_s(void)
{ struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC,&ts);
  return (ts.tv_sec + 1e-9*ts.tv_nsec);
}

static void tick(const char *what)
{ double t;
  if (!TIMING)
    return;
  t = now_s();
  fprintf(stderr,"[timing] %-28s %8.1f ms   (at %8.1f ms)\n",what,1e3*(t-T_last),1e3*(t-T_zero));
  T_last = t;
}

static void Clean_Exit(int val)                       /* damapper.c:543-554 */
{ char command[8192];
  snprintf(command,sizeof(command),"rm -r %s",SORT_PATH);
  if (system(command) != 0)
    { fprintf(stderr,"%s: Command Failed:\n%*s      %s\n",Prog_Name,(int) strlen(Prog_Name),"",command);
      exit (1);
    }
  tick("rm -r of the sort directory");
  fflush(NULL);                     /* everything is on disk: leave without the CUDA runtime's */
  _exit (val);                      /* teardown (the driver reclaims the context either way)    */
}

static uint64_t physical_memory(void)                  /* damapper.c:74-141, Linux branch */
{ long pages = sysconf(_SC_PHYS_PAGES), psize = sysconf(_SC_PAGESIZE);
  if (pages <= 0 || psize <= 0)
    return (0);
  return ((uint64_t) pages * (uint64_t) psize);
}

static int arg_int(const char *arg, const char *what, int positive)
{ char *e;
  long v = strtol(arg+2,&e,10);
  if (*e != '\0' || arg[2] == '\0')
    { fprintf(stderr,"%s: -%c '%s' argument is not an integer\n",Prog_Name,arg[1],arg+2);
      exit (1);
    }
  if (positive ? v <= 0 : v < 0)
    { fprintf(stderr,"%s: %s must be %s (%ld)\n",Prog_Name,what,positive ? "positive" : "non-negative",v);
      exit (1);
    }
  return ((int) v);
}

static double arg_real(const char *arg)
{ char *e;
  double v = strtod(arg+2,&e);
  if (*e != '\0' || arg[2] == '\0')
    { fprintf(stderr,"%s: -%c '%s' argument is not a real number\n",Prog_Name,arg[1],arg+2);
      exit (1);
    }
  return (v);
}

#define SYSTEM_CHECK(command)                                           \
 { if (VERBOSE)                                                         \
     printf("%s\n",command);                                            \
   if (system(command) != 0)                                            \
     { fprintf(stderr,"\n%s: Command Failed:\n%*s      %s\n",           \
                      Prog_Name,(int) strlen(Prog_Name),"",command);    \
       Clean_Exit(1);                                                   \
     }                                                                  \
 }

/* las_post.c: built-in LAsort + LAcat / LAmerge, used when the DALIGNER programs are not installed */
int las_sort_cat(const char *prefix, int nfiles, const char *out, int map_order, int verbose);
int on_path(const char *prog);

/* reference block k, resident between reads blocks */
typedef struct
  { damgpu_dblock *blk;                 /* the block (complemented: only its sizes are used again) */
    damgpu_index  *fwd, *rev;
    int            have;
  } Ref_Cache;

/* the next reads block, loaded by a helper thread */
typedef struct
  { const char *arg;
    Dazz_Block  blk;
    int         rc;                     /* dazz_load_packed status */
    int         mrc[256];               /* dazz_add_mask status per -m track */
    char      **mask;
    int         mtop;
    const char *prog;
    pthread_t   th;
    int         active;
  } Prefetch;

static void *prefetch_main(void *v)
{ Prefetch *p = (Prefetch *) v;
  int j;
  p->rc = dazz_load_packed(p->arg,&p->blk);
  if (p->rc == 0)
    for (j = 0; j < p->mtop; j++)
      { p->mrc[j] = dazz_add_mask(&p->blk,p->mask[j],p->prog);
        if (p->mrc[j] < 0)
          break;
      }
  return (NULL);
}

static void prefetch_start(Prefetch *p, const char *arg, char **mask, int mtop, const char *prog)
{ memset(p,0,sizeof(*p));
  p->arg = arg; p->mask = mask; p->mtop = mtop; p->prog = prog;
  if (pthread_create(&p->th,NULL,prefetch_main,p) == 0)
    p->active = 1;
  else                                  /* no thread: load it here and now */
    { prefetch_main(p);
      p->active = 2;
    }
}

static void prefetch_wait(Prefetch *p)
{ if (p->active == 1)
    pthread_join(p->th,NULL);
  p->active = 0;
}

int main(int argc, char *argv[])
{ int    VERBOSE = 0, PROFILE = 0, COVER = 0, NOMAP = 0, MAP_ORDER = 1;
  int    KMER_LEN = 20, MAX_REPS = 0, NTHREADS = 4, SPACING = 100, MTOP = 0;
  int    NGPUS = 1, RANK = 0;
  int   *MSHARED = NULL;                                  /* -G: mask-use flags of all workers */
  double AVE_ERROR = .85, BEST_TIE = 1.0;
  uint64_t MEM_PHYSICAL, MEM_LIMIT;
  int    mflag, i, j, k;
  char  *MASK[256];                                       /* -m tracks (damapper.c:575-637) */
  int    MSTAT[256];
  DIR   *dirp;
  Dazz_Block refdb, ablock, bblock;
  damgpu_options opts;
  damgpu_align_spec spec;
  Ref_Cache *rcache;
  int    builtin_sort = 0;
  damgpu_dblock *whole = NULL;                            /* whole reference, kept for the Reporter */
  uint64_t cache_budget = 0, cache_used = 0;
  Prefetch pre;

  TIMING = (getenv("DAMGPU_TIMING") != NULL);
  T_zero = T_last = now_s();
  MEM_PHYSICAL = physical_memory();
  MEM_LIMIT    = MEM_PHYSICAL;
  if (MEM_PHYSICAL == 0)
    { fprintf(stderr,"\nWarning: Could not get physical memory size\n");
      fflush(stderr);
    }

  j = 1;
  for (i = 1; i < argc; i++)
    if (argv[i][0] == '-')
      switch (argv[i][1])
      { default:
          for (k = 1; argv[i][k] != '\0'; k++)
            switch (argv[i][k])
            { case 'v': VERBOSE = 1; break;
              case 'p': PROFILE = 1; break;
              case 'z': MAP_ORDER = 0; break;
              case 'C': COVER = 1; break;
              case 'N': NOMAP = 1; break;
              default:
                fprintf(stderr,"%s: -%c is an illegal option\n",Prog_Name,argv[i][k]);
                exit (1);
            }
          break;
        case 'e':
          AVE_ERROR = arg_real(argv[i]);
          if (AVE_ERROR < .7 || AVE_ERROR >= 1.)
            { fprintf(stderr,"%s: Average correlation must be in [.7,1.) (%g)\n",Prog_Name,AVE_ERROR);
              exit (1);
            }
          break;
        case 'k':
          KMER_LEN = arg_int(argv[i],"K-mer length",1);
          if (KMER_LEN > 32)
            { fprintf(stderr,"%s: K-mer length must be 32 or less\n",Prog_Name);
              exit (1);
            }
          break;
        case 'm':
          if (MTOP >= 256)
            { fprintf(stderr,"%s: too many -m tracks\n",Prog_Name);
              exit (1);
            }
          MASK[MTOP++] = argv[i]+2;
          break;
        case 'n':
          BEST_TIE = arg_real(argv[i]);
          if (BEST_TIE < .7 || BEST_TIE > 1.)
            { fprintf(stderr,"%s: Near optimal threshold must be in [.7,1.] (%g)\n",Prog_Name,BEST_TIE);
              exit (1);
            }
          break;
        case 's':
          SPACING = arg_int(argv[i],"Trace spacing",1);
          break;
        case 't':
          MAX_REPS = arg_int(argv[i],"Tuple supression frequency",1);
          break;
        case 'M':
          MEM_LIMIT = (uint64_t) arg_int(argv[i],"Memory allocation (in Gb)",0) * 0x40000000ull;
          break;
        case 'P':
          SORT_PATH = argv[i]+2;
          if ((dirp = opendir(SORT_PATH)) == NULL)
            { fprintf(stderr,"%s: -P option: cannot open directory %s\n",Prog_Name,SORT_PATH);
              exit (1);
            }
          closedir(dirp);
          break;
        case 'T':
          NTHREADS = arg_int(argv[i],"Number of threads",1);
          break;
        case 'G':
          NGPUS = arg_int(argv[i],"Number of GPUs",1);
          break;
      }
    else
      argv[j++] = argv[i];
  argc = j;

  if (argc <= 2)
    { fprintf(stderr,"Usage: %s %s\n",Prog_Name,Usage[0]);
      fprintf(stderr,"       %*s %s\n",(int) strlen(Prog_Name),"",Usage[1]);
      fprintf(stderr,"       %*s %s\n",(int) strlen(Prog_Name),"",Usage[2]);
      fprintf(stderr,"\n");
      fprintf(stderr,"      -k: k-mer size (must be <= 32).\n");
      fprintf(stderr,"      -t: Ignore k-mers that occur >= -t times in a block.\n");
      fprintf(stderr,"      -M: Use only -M GB of memory by ignoring most frequent k-mers.\n");
      fprintf(stderr,"\n");
      fprintf(stderr,"      -e: Look for alignments with -e percent similarity.\n");
      fprintf(stderr,"      -s: Use -s as the trace point spacing for encoding alignments.\n");
      fprintf(stderr,"      -n: Output all matches within this %% of the best\n");
      fprintf(stderr,"\n");
      fprintf(stderr,"      -T: Use -T threads (here: number of per-range .las files).\n");
      fprintf(stderr,"      -P: Do sorts and merges in directory -P.\n");
      fprintf(stderr,"      -G: Map the reads blocks on -G GPUs of this box, round robin.\n");
      fprintf(stderr,"      -m: Soft mask the blocks with the specified mask.\n");
      fprintf(stderr,"\n");
      fprintf(stderr,"      -v: Verbose mode, output statistics as proceed.\n");
      fprintf(stderr,"      -z: sort .las by A,B-read pairs (overlap piles)\n");
      fprintf(stderr,"          off => sort .las by A-read,A-position pairs (default for mapping)\n");
      fprintf(stderr,"      -p: Output repeat profile track\n");
      fprintf(stderr,"      -C: Output reference vs reads .las.\n");
      fprintf(stderr,"      -N: Do not output reads vs reference .las.\n");
      exit (1);
    }

  if (COVER)
    mflag = NOMAP ? 2 : 3;                              /* FLAG_DOB / FLAG_DOA|FLAG_DOB */
  else if (NOMAP)
    { fprintf(stderr,"%s: Cannot specify N flag without C also\n",Prog_Name);
      exit (1);
    }
  else
    mflag = 1;
  if (NOMAP && PROFILE)
    { fprintf(stderr,"%s: Cannot specify both N and p flags together\n",Prog_Name);
      exit (1);
    }
  for (j = 0; j < MTOP; j++)                              /* damapper.c:727-728 */
    MSTAT[j] = 0;

  /* reference: stub, block count, base frequencies (damapper.c:734-797) */
  tick("flags");
  if (dazz_open(argv[1],&refdb) != 0)
    exit (1);
  tick("open reference stub");
  if (refdb.part > 0)
    { fprintf(stderr,"%s: first argument '%s' cannot be a block\n",Prog_Name,argv[1]);
      exit (1);
    }
  if (refdb.nblocks == 0)
    { fprintf(stderr,"%s: DB %s has not yet been partitioned, cannot request a block !\n",
                     Prog_Name,refdb.root);
      exit (1);
    }
  spec.ave_corr = AVE_ERROR;
  spec.trace_space = SPACING;
  memcpy(spec.freq,refdb.freq,sizeof(spec.freq));

  /* -G: one worker process per GPU, forked before anything touches CUDA */
  if (NGPUS > argc-2)
    NGPUS = argc-2;
  if (NGPUS > 1)
    { pid_t kids[64];
      char  devs[64][16];
      const char *list = getenv("DAMGPU_DEVICES");
      int   r, st, bad = 0;
      if (NGPUS > 64) NGPUS = 64;
      for (r = 0; r < NGPUS; r++)
        snprintf(devs[r],sizeof(devs[r]),"%d",r);
      if (list == NULL) list = getenv("CUDA_VISIBLE_DEVICES");
      if (list != NULL)                                   /* r-th entry of the list */
        { const char *p = list;
          for (r = 0; r < NGPUS && *p != '\0'; r++)
            { size_t n = strcspn(p,",");
              if (n >= sizeof(devs[r])) n = sizeof(devs[r])-1;
              memcpy(devs[r],p,n); devs[r][n] = '\0';
              p += n; if (*p == ',') p++;
            }
          if (r < NGPUS)
            { fprintf(stderr,"%s: -G%d but only %d devices in '%s'\n",Prog_Name,NGPUS,r,list);
              exit (1);
            }
        }
      MSHARED = (int *) mmap(NULL,sizeof(int)*256,PROT_READ|PROT_WRITE,MAP_SHARED|MAP_ANONYMOUS,-1,0);
      if (MSHARED == MAP_FAILED)
        MSHARED = NULL;
      else
        memset(MSHARED,0,sizeof(int)*256);
      fflush(NULL);
      for (r = 0; r < NGPUS; r++)
        { kids[r] = fork();
          if (kids[r] < 0)
            { fprintf(stderr,"%s: cannot fork worker %d\n",Prog_Name,r);
              exit (1);
            }
          if (kids[r] == 0)
            { RANK = r;
              setenv("CUDA_VISIBLE_DEVICES",devs[r],1);
              unsetenv("DAMGPU_DEVICE");
              break;
            }
        }
      if (r == NGPUS)                                     /* the parent: wait, report, leave */
        { for (r = 0; r < NGPUS; r++)
            if (waitpid(kids[r],&st,0) < 0 || !WIFEXITED(st) || WEXITSTATUS(st) != 0)
              bad = 1;
          if (!bad && MSHARED != NULL)
            for (j = 0; j < MTOP; j++)
              if (MSHARED[j] == 0)
                printf("%s: Warning: Track %s given but never used.\n",Prog_Name,MASK[j]);
          exit (bad);
        }
    }
  else
    { /* one device: do not make cuInit enumerate the others */
      const char *dev = getenv("DAMGPU_DEVICE");
      if (dev != NULL && getenv("CUDA_VISIBLE_DEVICES") == NULL)
        { setenv("CUDA_VISIBLE_DEVICES",dev,1);
          unsetenv("DAMGPU_DEVICE");
        }
    }

  /* the first reads block comes off the disk while the CUDA context is created */
  memset(&pre,0,sizeof(pre));
  if (2+RANK < argc)
    prefetch_start(&pre,argv[2+RANK],MASK,MTOP,Prog_Name);

  { const char *dev = getenv("DAMGPU_DEVICE");
    if (damgpu_init(dev ? atoi(dev) : -1) != 0)
      { fprintf(stderr,"%s: no usable CUDA device (%s); this build has no CPU path\n",
                       Prog_Name,damgpu_last_error());
        exit (1);
      }
  }
  tick("damgpu_init");
  if (damgpu_Set_Filter_Params(KMER_LEN,MAX_REPS,NTHREADS))
    { fprintf(stderr,"Illegal combination of filter parameters\n");
      exit (1);
    }

  { char *newpath = (char *) malloc(strlen(SORT_PATH)+30);   /* damapper.c:806-817 */
    if (newpath == NULL)
      { fprintf(stderr,"%s: Out of memory (Allocating sort path)\n",Prog_Name);
        exit (1);
      }
    sprintf(newpath,"%s/damapper.%d",SORT_PATH,getpid());
    if (mkdir(newpath,S_IRWXU) != 0)
      { fprintf(stderr,"%s: Could not create directory %s\n",Prog_Name,newpath);
        exit (1);
      }
    SORT_PATH = newpath;
  }
  opts.verbose = VERBOSE; opts.profile = PROFILE; opts.spacing = SPACING; opts.best_tie = BEST_TIE;
  opts.sort_path = SORT_PATH; opts.mem_limit = MEM_LIMIT; opts.mem_physical = MEM_PHYSICAL;
  damgpu_set_options(&opts);
  damgpu_set_fatal(Clean_Exit);

  rcache = (Ref_Cache *) calloc((size_t) refdb.nblocks+1,sizeof(Ref_Cache));
  { uint64_t fr = 0, tot = 0;
    const char *e = getenv("DAMGPU_REF_CACHE");           /* percent of HBM, 0 = as the reference */
    double pct = (e != NULL) ? atof(e) : 45.;
    if (rcache != NULL && 2+RANK+NGPUS < argc && pct > 0. && damgpu_device_memory(&fr,&tot) == 0)
      cache_budget = (uint64_t) (tot * (pct < 90. ? pct : 90.) / 100.);
  }
  /* the DALIGNER post-processing programs, or the built-in stand-in when they are not installed */
  builtin_sort = (getenv("DAMGPU_BUILTIN_SORT") != NULL || !on_path("LAsort"));
  if (builtin_sort && VERBOSE)
    printf("\n  LAsort is not on PATH (or DAMGPU_BUILTIN_SORT is set): built-in sort and merge of the .las files\n");
  else if (builtin_sort && RANK == 0 && getenv("DAMGPU_BUILTIN_SORT") == NULL)
    fprintf(stderr,"%s: LAsort is not on PATH: the .las files are sorted and merged by the built-in stand-in\n"
                   "%*s  (its order follows DALIGNER's documented -a key and is not pinned against LAsort/LAmerge)\n",
                   Prog_Name,(int) strlen(Prog_Name),"");
  for (i = 2+RANK; i < argc; i += NGPUS)                  /* damapper.c:825-914 */
    { char *broot, *aroot = refdb.root, name[4096], command[16384];
      damgpu_block  bview, aview;
      damgpu_dblock *dreads, *dref;
      damgpu_index  *bindex, *aindex;
      damgpu_mapper *mapper;
      damgpu_report *rep;

      prefetch_wait(&pre);                              /* 2 bits per base to the device */
      if (pre.rc != 0)
        Clean_Exit(1);
      bblock = pre.blk;
      for (j = 0; j < MTOP; j++)                        /* read_DB, damapper.c:352-399 */
        { if (pre.mrc[j] < 0) Clean_Exit(1);
          if (pre.mrc[j] > 0) MSTAT[j] = 1;
        }
      for (k = 0; k < bblock.nreads; k++)
        if (bblock.rlen[k] < KMER_LEN)
          { fprintf(stderr,"%s: Block %s contains reads < %dbp long !  Run DBsplit -x%d\n",
                           Prog_Name,argv[i],KMER_LEN,KMER_LEN);
            Clean_Exit(1);
          }
      if (bblock.part > 0)
        { broot = (char *) malloc(strlen(bblock.root)+20);
          if (broot != NULL)
            sprintf(broot,"%s.%d",bblock.root,bblock.part);
        }
      else
        broot = strdup(bblock.root);
      if (broot == NULL)
        { fprintf(stderr,"%s: Out of memory (Allocating block name)\n",Prog_Name);
          Clean_Exit(1);
        }
      dazz_view(&bblock,&bview);
      tick("load reads block");
      if (VERBOSE)
        printf("\nBuilding index for %s\n",broot);
      dreads = damgpu_block_upload_packed(&bview,bblock.packed,bblock.poff,bblock.packed_bytes);
      tick("upload reads block");
      if (i+NGPUS < argc)                                /* the disk works while the GPU does */
        prefetch_start(&pre,argv[i+NGPUS],MASK,MTOP,Prog_Name);
      /* one or two reference blocks: the reads list is built per block from the k-mers that occur in
         it (deferred); more: sorted once in full */
      bindex = (refdb.nblocks <= 2) ? damgpu_index_build_deferred(dreads) : damgpu_index_build(dreads);
      tick("index reads block");
      mapper = damgpu_mapper_new(dreads,bindex);
      tick("mapper_new");

      for (k = 1; k <= refdb.nblocks; k++)
        { Ref_Cache *rc = rcache+k;
          if (rc->have)                                  /* both indices are still in HBM */
            { if (VERBOSE)
                printf("\nComparing %s to %s.%d\n",broot,aroot,k);
              damgpu_mapper_match(mapper,rc->blk,rc->fwd,0,(k == 1));
              if (VERBOSE)
                printf("\nComparing %s to c(%s.%d)\n",broot,aroot,k);
              damgpu_mapper_match(mapper,rc->blk,rc->rev,1,0);
              tick("match, both strands (cached)");
              continue;
            }
          snprintf(name,sizeof(name),"%s/%s.%d.%s",refdb.pwd,aroot,k,refdb.isdam ? "dam" : "db");
          if (dazz_load_packed(name,&ablock) != 0)
            Clean_Exit(1);
          for (j = 0; j < MTOP; j++)
            { int st = dazz_add_mask(&ablock,MASK[j],Prog_Name);
              if (st < 0) Clean_Exit(1);
              if (st > 0) MSTAT[j] = 1;
            }
          dazz_view(&ablock,&aview);
          tick("load reference block");
          if (VERBOSE)
            printf("\nBuilding index for %s.%d\n",aroot,k);
          dref = damgpu_block_upload_packed(&aview,ablock.packed,ablock.poff,ablock.packed_bytes);
          aindex = damgpu_index_build(dref);
          if (VERBOSE)
            printf("\nComparing %s to %s.%d\n",broot,aroot,k);
          damgpu_mapper_match(mapper,dref,aindex,0,(k == 1));
          { /* a block image + two lists of 16-byte records, one per base and strand */
            uint64_t need = (uint64_t) ablock.totlen*33 + (uint64_t) ablock.nreads*48 + (1u << 20);
            if (cache_budget > 0 && cache_used + need <= cache_budget)
              { cache_used += need;
                rc->have = 1; rc->blk = dref; rc->fwd = aindex;
              }
            else
              damgpu_index_free(aindex);
          }

          damgpu_block_complement(dref);
          if (VERBOSE)
            printf("\nBuilding index for c(%s.%d)\n",aroot,k);
          aindex = damgpu_index_build(dref);
          if (VERBOSE)
            printf("\nComparing %s to c(%s.%d)\n",broot,aroot,k);
          damgpu_mapper_match(mapper,dref,aindex,1,0);
          if (rc->have)
            rc->rev = aindex;
          else
            { damgpu_index_free(aindex);
              damgpu_block_free(dref);
            }
          dazz_close(&ablock);
          tick("index + match, both strands");
        }

      if (whole != NULL)
        dref = whole;
      else
        { snprintf(name,sizeof(name),"%s/%s.%s",refdb.pwd,aroot,refdb.isdam ? "dam" : "db");
          if (dazz_load_packed(name,&ablock) != 0)
            Clean_Exit(1);
          dazz_view(&ablock,&aview);
          tick("load whole reference");
          dref = damgpu_block_upload_packed(&aview,ablock.packed,ablock.poff,ablock.packed_bytes);
          { uint64_t need = (uint64_t) ablock.totlen + (uint64_t) ablock.nreads*16 + (1u << 20);
            if (cache_budget > 0 && cache_used + need <= cache_budget)
              { cache_used += need;
                whole = dref;
              }
          }
          dazz_close(&ablock);
        }
      if (VERBOSE)
        printf("\nFinding best matches for block %s\n",broot);
      rep  = damgpu_mapper_report(mapper,dref,&spec,mflag);
      tick("upload reference + Reporter");
      { int nfiles = 1;
        while (2*nfiles <= NTHREADS) nfiles *= 2;
        if ((mflag & 1) && damgpu_report_write_las(rep,0,SORT_PATH,broot,aroot,nfiles,SPACING))
          Clean_Exit(1);
        if ((mflag & 2) && damgpu_report_write_las(rep,1,SORT_PATH,broot,aroot,nfiles,SPACING))
          Clean_Exit(1);
        if (PROFILE && damgpu_report_write_profile(rep,&bview,".",broot,SPACING))
          Clean_Exit(1);
      }
      tick("write .las / -p track");
      if (VERBOSE)
        { long long nseg = (long long) damgpu_report_records(rep,(mflag & 1) ? 0 : 1);
          char digits[32];                                 /* map.c:3289-3293: commas between thousands */
          int  nd = snprintf(digits,sizeof(digits),"%lld",nseg), d;
          printf("      ");
          for (d = 0; d < nd; d++)
            { if (d > 0 && (nd-d) % 3 == 0 && nd-d <= 9)
                putchar(',');
              putchar(digits[d]);
            }
          printf(" mapped segments\n");
          fflush(stdout);
        }
      damgpu_report_free(rep);
      if (dref != whole)
        damgpu_block_free(dref);
      damgpu_mapper_free(mapper);
      damgpu_index_free(bindex);
      damgpu_block_free(dreads);
      dazz_close(&bblock);
      tick("release block");

      if (builtin_sort)                                  /* no DALIGNER programs: las_post.c */
        { int nfiles = 1;
          while (2*nfiles <= NTHREADS) nfiles *= 2;
          if ((mflag & 1) != 0)
            { sprintf(command,"%s/%s.%s.M",SORT_PATH,broot,aroot);
              sprintf(command+strlen(command)+1,"%s.%s.las",broot,aroot);
              if (las_sort_cat(command,nfiles,command+strlen(command)+1,MAP_ORDER,VERBOSE))
                Clean_Exit(1);
            }
          if ((mflag & 2) != 0)
            { sprintf(command,"%s/%s.%s.R",SORT_PATH,aroot,broot);
              sprintf(command+strlen(command)+1,"%s.%s.las",aroot,broot);
              if (las_sort_cat(command,nfiles,command+strlen(command)+1,MAP_ORDER,VERBOSE))
                Clean_Exit(1);
            }
          tick("built-in LAsort / LAcat / LAmerge");
          free(broot);
          continue;
        }
      if ((mflag & 1) != 0)                              /* damapper.c:893-901 */
        { sprintf(command,"LAsort %s %s %s/%s.%s.M%c.las",VERBOSE?"-v":"",MAP_ORDER?"-a":"",
                          SORT_PATH,broot,aroot,'@');
          SYSTEM_CHECK(command)
          sprintf(command,"LAcat %s %s/%s.%s.M%c.S >%s.%s.las",VERBOSE?"-v":"",
                          SORT_PATH,broot,aroot,'@',broot,aroot);
          SYSTEM_CHECK(command)
        }
      if ((mflag & 2) != 0)                              /* damapper.c:903-911 */
        { sprintf(command,"LAsort %s %s %s/%s.%s.R%c.las",VERBOSE?"-v":"",MAP_ORDER?"-a":"",
                          SORT_PATH,aroot,broot,'@');
          SYSTEM_CHECK(command)
          sprintf(command,"LAmerge %s %s %s.%s %s/%s.%s.R%c.S.las",VERBOSE?"-v":"",MAP_ORDER?"-a":"",
                          aroot,broot,SORT_PATH,aroot,broot,'@');
          SYSTEM_CHECK(command)
        }
      tick("LAsort / LAcat / LAmerge");
      free(broot);
    }

  for (j = 0; j < MTOP; j++)                              /* damapper.c:916-918 */
    if (MSHARED != NULL)
      { if (MSTAT[j]) MSHARED[j] = 1; }                   /* -G: the parent prints what no worker used */
    else if (MSTAT[j] == 0)
      printf("%s: Warning: Track %s given but never used.\n",Prog_Name,MASK[j]);

  Clean_Exit(0);
  return (0);
}
