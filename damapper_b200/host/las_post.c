/* Built-in stand-in for the DALIGNER programs damapper shells out to after every reads block
 * (reference damapper.c:893-911: `LAsort [-a] <dir>/X.Y.M@.las`, `LAcat <dir>/X.Y.M@.S >X.Y.las`,
 * `LAsort [-a] <dir>/Y.X.R@.las`, `LAmerge [-a] Y.X <dir>/Y.X.R@.S.las`).  SURVEY section 8 row (f)1.
 *
 * The DALIGNER sources are not part of the reference tree, so the order below follows their
 * documented behaviour and is NOT pinned by a run of the real programs ("parity unpinned"):
 *   - a chain (a record flagged START and the NEXT records that follow it, align.h:127-135) is a unit
 *     and takes the key of its first record;
 *   - with -a (damapper's default, MAP_ORDER) the key is (aread, abpos); without it
 *     (aread, bread, COMP flag, abpos);
 *   - the sort is stable: equal keys keep damapper's emission order (best chain first).
 * The driver uses this only when no `LAsort` is on PATH or DAMGPU_BUILTIN_SORT is set; with the real
 * programs installed it runs them exactly as the reference does.
 *
 * .las layout (align.c:3115-3142, map.c:2421-2428): int64 novl, int tspace, then per record 40 bytes
 * {int tlen, diffs, abpos, bbpos, aepos, bepos; uint32 flags; int aread, bread; 4 pad} + tlen trace
 * values of 1 byte (tspace <= 125) or 2. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <unistd.h>

typedef struct
  { int64_t off, len;            /* bytes of the chain inside the pool */
    int     aread, bread, comp, abpos;
    int     nrec;
    int64_t seq;                 /* emission order (stable tie-break) */
  } Chain;

static int MAP_KEY;

static int chain_cmp(const void *x, const void *y)
{ const Chain *a = (const Chain *) x, *b = (const Chain *) y;
  if (a->aread != b->aread) return (a->aread < b->aread ? -1 : 1);
  if (!MAP_KEY)
    { if (a->bread != b->bread) return (a->bread < b->bread ? -1 : 1);
      if (a->comp != b->comp) return (a->comp < b->comp ? -1 : 1);
    }
  if (a->abpos != b->abpos) return (a->abpos < b->abpos ? -1 : 1);
  return (a->seq < b->seq ? -1 : (a->seq > b->seq));
}

/* Reads the nfiles per-thread files `<prefix><i>.las` (i = 1..nfiles), sorts their chains and writes
 * `out`.  Returns 0, or 1 with a message on stderr. */
int las_sort_cat(const char *prefix, int nfiles, const char *out, int map_order, int verbose)
{ uint8_t *pool = NULL;
  int64_t  plen = 0, pmax = 0, novl = 0;
  Chain   *ch = NULL;
  int64_t  nch = 0, cmax = 0;
  int      tspace = -1, i;
  char     name[4096];

  for (i = 1; i <= nfiles; i++)
    { FILE *f;
      int64_t n, k;
      int ts, tbytes;

      snprintf(name,sizeof(name),"%s%d.las",prefix,i);
      f = fopen(name,"r");
      if (f == NULL)
        { fprintf(stderr,"las_sort_cat: cannot open %s\n",name);
          free(pool); free(ch);
          return (1);
        }
      if (fread(&n,sizeof(int64_t),1,f) != 1 || fread(&ts,sizeof(int),1,f) != 1)
        { fprintf(stderr,"las_sort_cat: %s has no header\n",name);
          fclose(f); free(pool); free(ch);
          return (1);
        }
      if (tspace >= 0 && ts != tspace)
        { fprintf(stderr,"las_sort_cat: %s has trace spacing %d, the others %d\n",name,ts,tspace);
          fclose(f); free(pool); free(ch);
          return (1);
        }
      tspace = ts;
      tbytes = (ts <= 125) ? 1 : 2;                      /* TRACE_XOVR, align.h:45 */
      for (k = 0; k < n; k++)
        { int32_t h[10];
          int64_t tl;
          if (fread(h,40,1,f) != 1)
            { fprintf(stderr,"las_sort_cat: %s is truncated\n",name);
              fclose(f); free(pool); free(ch);
              return (1);
            }
          tl = (int64_t) h[0]*tbytes;
          if (plen + 40 + tl > pmax)
            { pmax = (plen + 40 + tl)*2 + (1 << 20);
              pool = (uint8_t *) realloc(pool,pmax);
              if (pool == NULL)
                { fprintf(stderr,"las_sort_cat: out of memory\n");
                  fclose(f); free(ch);
                  return (1);
                }
            }
          memcpy(pool+plen,h,40);
          if (tl > 0 && fread(pool+plen+40,1,tl,f) != (size_t) tl)
            { fprintf(stderr,"las_sort_cat: %s is truncated\n",name);
              fclose(f); free(pool); free(ch);
              return (1);
            }
          if ((((uint32_t) h[6]) & 0x8) == 0 || nch == 0)  /* not NEXT: a new unit */
            { if (nch >= cmax)
                { cmax = cmax*2 + 1024;
                  ch = (Chain *) realloc(ch,sizeof(Chain)*cmax);
                  if (ch == NULL)
                    { fprintf(stderr,"las_sort_cat: out of memory\n");
                      fclose(f); free(pool);
                      return (1);
                    }
                }
              ch[nch].off = plen; ch[nch].len = 0; ch[nch].nrec = 0;
              ch[nch].aread = h[7]; ch[nch].bread = h[8];
              ch[nch].comp = (int) (((uint32_t) h[6]) & 0x1);
              ch[nch].abpos = h[2];
              ch[nch].seq = nch;
              nch += 1;
            }
          ch[nch-1].len += 40 + tl;
          ch[nch-1].nrec += 1;
          plen += 40 + tl;
          novl += 1;
        }
      fclose(f);
    }
  if (tspace < 0)
    tspace = 100;

  MAP_KEY = map_order;
  if (nch > 1)
    qsort(ch,nch,sizeof(Chain),chain_cmp);               /* total order: seq breaks every tie */

  { FILE *o = fopen(out,"w");
    int64_t c;
    if (o == NULL)
      { fprintf(stderr,"las_sort_cat: cannot create %s\n",out);
        free(pool); free(ch);
        return (1);
      }
    fwrite(&novl,sizeof(int64_t),1,o);
    fwrite(&tspace,sizeof(int),1,o);
    for (c = 0; c < nch; c++)
      fwrite(pool+ch[c].off,1,ch[c].len,o);
    if (fclose(o) != 0)
      { fprintf(stderr,"las_sort_cat: error writing %s\n",out);
        free(pool); free(ch);
        return (1);
      }
  }
  if (verbose)
    printf("  built-in sort: %lld records in %lld chains -> %s\n",(long long) novl,(long long) nch,out);
  free(pool); free(ch);
  return (0);
}

/* is `prog` an executable on PATH? */
int on_path(const char *prog)
{ const char *path = getenv("PATH");
  char  buf[4096];
  if (path == NULL)
    return (0);
  while (*path)
    { const char *e = strchr(path,':');
      size_t n = e ? (size_t) (e-path) : strlen(path);
      if (n > 0 && n + strlen(prog) + 2 < sizeof(buf))
        { memcpy(buf,path,n);
          buf[n] = '/';
          strcpy(buf+n+1,prog);
          if (access(buf,X_OK) == 0)
            return (1);
        }
      path += n;
      if (*path == ':') path += 1;
    }
  return (0);
}
