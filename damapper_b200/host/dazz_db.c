#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "dazz_db.h"

#define DB_BEST 0x800      /* DB.h:278 */
#define DB_ALL  0x1

/* raw .idx records, LP64 layout (DB.h:285-295,390-420) */
typedef struct
  { int32_t ureads, treads, cutoff, allarr;
    float   freq[4];
    int32_t maxlen, pad0;
    int64_t totlen;
    int32_t nreads, trimmed, part, ufirst, tfirst, pad1;
    int64_t path;
    int32_t loaded, pad2;
    int64_t bases, reads, tracks;
  } Idx_Header;

typedef struct
  { int32_t origin, rlen, fpulse, pad0;
    int64_t boff, coff;
    int32_t flags, pad1;
  } Idx_Read;

static char *dup_range(const char *s, size_t n)
{ char *r = (char *) malloc(n+1);
  memcpy(r,s,n);
  r[n] = '\0';
  return (r);
}

/* split "dir/root[.N][.db|.dam]" */
static void split_name(const char *name, char **pwd, char **root, int *part, int *ext)
{ const char *slash = strrchr(name,'/');
  const char *base = slash ? slash+1 : name;
  size_t len = strlen(base);
  *ext = 0;
  if (len > 4 && strcmp(base+len-4,".dam") == 0) { len -= 4; *ext = 2; }
  else if (len > 3 && strcmp(base+len-3,".db") == 0) { len -= 3; *ext = 1; }
  *pwd = slash ? dup_range(name,(size_t) (slash-name)) : dup_range(".",1);
  *part = 0;
  { size_t i = len;
    while (i > 0 && base[i-1] >= '0' && base[i-1] <= '9') i--;
    if (i < len && i > 1 && base[i-1] == '.')
      { *part = atoi(base+i);
        if (*part > 0) len = i-1; else *part = 0;
      }
  }
  *root = dup_range(base,len);
}

static FILE *open_stub(Dazz_Block *db, int ext)
{ char path[4096];
  FILE *f = NULL;
  if (ext != 2)
    { snprintf(path,sizeof(path),"%s/%s.db",db->pwd,db->root);
      if ((f = fopen(path,"r")) != NULL) { db->isdam = 0; return (f); }
    }
  snprintf(path,sizeof(path),"%s/%s.dam",db->pwd,db->root);
  if ((f = fopen(path,"r")) != NULL) { db->isdam = 1; return (f); }
  return (NULL);
}

int dazz_open(const char *name, Dazz_Block *db)
{ int  ext, nfiles, i, x;
  char a[2048], b[2048];
  FILE *f;

  memset(db,0,sizeof(*db));
  split_name(name,&db->pwd,&db->root,&db->part,&ext);
  if ((f = open_stub(db,ext)) == NULL)
    { fprintf(stderr,"damapper: Could not open database %s\n",name);
      return (-1);
    }
  if (fscanf(f,"files = %9d\n",&nfiles) != 1)
    { fprintf(stderr,"damapper: Stub file (.db) of %s is junk\n",db->root);
      fclose(f);
      return (-1);
    }
  for (i = 0; i < nfiles; i++)
    if (fscanf(f,"  %9d %2047s %2047s\n",&x,a,b) != 3)
      { fprintf(stderr,"damapper: Stub file (.db) of %s is junk\n",db->root);
        fclose(f);
        return (-1);
      }
  if (fscanf(f,"blocks = %9d\n",&db->nblocks) != 1)
    db->nblocks = 0;
  fclose(f);

  { char path[4096];
    Idx_Header h;
    snprintf(path,sizeof(path),"%s/.%s.idx",db->pwd,db->root);
    if ((f = fopen(path,"r")) == NULL || fread(&h,sizeof(h),1,f) != 1)
      { fprintf(stderr,"damapper: Index file (.idx) of %s is junk\n",db->root);
        if (f) fclose(f);
        return (-1);
      }
    fclose(f);
    memcpy(db->freq,h.freq,sizeof(db->freq));
  }
  db->path_len = (int64_t) (strlen(db->pwd) + 2 + strlen(db->root));
  return (0);
}

static int load_block(const char *name, Dazz_Block *db, int keep_packed)
{ char  path[4096], a[2048], b[2048];
  FILE *f;
  int   ext, nfiles, nblocks = 0, cutoff = 0, all = 1, i, x;
  long long size;
  int   ufirst = 0, ulast, tfirst = 0;
  Idx_Header h;
  Idx_Read  *recs;

  if (dazz_open(name,db) != 0)
    return (-1);
  split_name(name,&db->pwd,&db->root,&db->part,&ext);

  if ((f = open_stub(db,ext)) == NULL)
    return (-1);
  if (fscanf(f,"files = %9d\n",&nfiles) != 1) { fclose(f); return (-1); }
  for (i = 0; i < nfiles; i++)
    if (fscanf(f,"  %9d %2047s %2047s\n",&x,a,b) != 3) { fclose(f); return (-1); }
  if (fscanf(f,"blocks = %9d\n",&nblocks) == 1)
    { if (fscanf(f,"size = %11lld cutoff = %9d all = %1d\n",&size,&cutoff,&all) != 3)
        { fprintf(stderr,"damapper: Stub file (.db) of %s is junk\n",db->root);
          fclose(f);
          return (-1);
        }
    }
  else if (db->part > 0)
    { fprintf(stderr,"damapper: DB %s has not yet been partitioned, cannot request a block !\n",db->root);
      fclose(f);
      return (-1);
    }

  snprintf(path,sizeof(path),"%s/.%s.idx",db->pwd,db->root);
  { FILE *g = fopen(path,"r");
    if (g == NULL || fread(&h,sizeof(h),1,g) != 1)
      { fprintf(stderr,"damapper: Index file (.idx) of %s is junk\n",db->root);
        fclose(f);
        return (-1);
      }
    recs = (Idx_Read *) malloc(sizeof(Idx_Read)*(size_t) (h.ureads+1));
    if (fread(recs,sizeof(Idx_Read),(size_t) h.ureads,g) != (size_t) h.ureads)
      { fprintf(stderr,"damapper: Index file (.idx) of %s is junk\n",db->root);
        fclose(g); fclose(f);
        return (-1);
      }
    fclose(g);
  }
  if (nblocks == 0)
    { cutoff = h.cutoff < 0 ? 0 : h.cutoff;
      all = (h.allarr & DB_ALL) != 0;
    }
  ulast = h.ureads;
  if (db->part > 0)
    { int uf = 0, tf = 0, ul = 0, tl = 0;
      if (db->part > nblocks)
        { fprintf(stderr,"damapper: DB %s has only %d blocks\n",db->root,nblocks);
          fclose(f);
          return (-1);
        }
      for (i = 0; i <= db->part; i++)
        { uf = ul; tf = tl;
          if (fscanf(f," %9d %9d\n",&ul,&tl) != 2)
            { fprintf(stderr,"damapper: Stub file (.db) of %s is junk\n",db->root);
              fclose(f);
              return (-1);
            }
        }
      ufirst = uf; ulast = ul; tfirst = tf;
    }
  fclose(f);
  db->cutoff = cutoff; db->all = all; db->nblocks = nblocks;

  /* trim (Trim_DB semantics: keep reads >= cutoff that are the best of their well unless all) */
  { int     n = 0, u;
    int64_t tot = 0, o = 0;
    int     maxlen = 0;
    FILE   *bps;

    for (u = ufirst; u < ulast; u++)
      if ((all || (recs[u].flags & DB_BEST)) && recs[u].rlen >= cutoff)
        { n += 1; tot += recs[u].rlen;
          if (recs[u].rlen > maxlen) maxlen = recs[u].rlen;
        }
    db->nreads = n; db->tfirst = tfirst; db->totlen = tot; db->maxlen = maxlen;
    db->ufirst = ufirst; db->ulast = ulast; db->db_ureads = h.ureads; db->db_treads = h.treads;
    db->kept = (uint8_t *) malloc((size_t) (ulast-ufirst) + 1);
    for (u = ufirst; u < ulast; u++)
      db->kept[u-ufirst] = (uint8_t) ((all || (recs[u].flags & DB_BEST)) && recs[u].rlen >= cutoff);
    db->mask_off = NULL; db->mask_pts = NULL;
    db->raw = db->packed = NULL; db->poff = NULL; db->packed_bytes = 0;
    if (keep_packed)
      { db->packed = (uint8_t *) malloc((size_t) (tot/4 + n + 8));
        db->poff   = (int64_t *) malloc(sizeof(int64_t)*(size_t) (n+1));
      }
    else
      db->raw  = (uint8_t *) malloc((size_t) (tot + n + 8));
    db->boff = (int64_t *) malloc(sizeof(int64_t)*(size_t) (n+1));
    db->rlen = (int32_t *) malloc(sizeof(int32_t)*(size_t) (n+1));
    if ((keep_packed ? (db->packed == NULL || db->poff == NULL) : db->raw == NULL) ||
        db->boff == NULL || db->rlen == NULL)
      { fprintf(stderr,"damapper: Out of memory (Allocating All Sequence Reads)\n");
        return (-1);
      }
    snprintf(path,sizeof(path),"%s/.%s.bps",db->pwd,db->root);
    if ((bps = fopen(path,"r")) == NULL)
      { fprintf(stderr,"damapper: Cannot open %s\n",path);
        return (-1);
      }
    if (!keep_packed)
      db->raw[0] = 4;
    n = 0;
    for (u = ufirst; u < ulast; u++)
      if ((all || (recs[u].flags & DB_BEST)) && recs[u].rlen >= cutoff)
        { int      len = recs[u].rlen, clen = (len+3) >> 2, j;
          uint8_t *s = keep_packed ? NULL : db->raw + 1 + o;
          uint8_t *c = keep_packed ? db->packed + db->packed_bytes
                                   : s + (len - clen);   /* read the packed bytes into the tail */
          if (len < 0 || fseeko(bps,recs[u].boff,SEEK_SET) != 0 ||
              (clen > 0 && fread(c,1,(size_t) clen,bps) != (size_t) clen))
            { fprintf(stderr,"damapper: Read of .bps file failed\n");
              fclose(bps);
              return (-1);
            }
          if (keep_packed)
            { db->poff[n] = db->packed_bytes;
              db->packed_bytes += clen;
            }
          else
            { for (j = 0; j < len; j++)            /* first base in the two top bits */
                s[j] = (uint8_t) ((c[j >> 2] >> (6 - 2*(j & 3))) & 3);
              s[len] = 4;
            }
          db->boff[n] = o;
          db->rlen[n] = len;
          o += len+1;
          n += 1;
        }
    db->boff[n] = o;
    fclose(bps);
  }
  free(recs);
  return (0);
}

int dazz_load(const char *name, Dazz_Block *db)        { return (load_block(name,db,0)); }
int dazz_load_packed(const char *name, Dazz_Block *db) { return (load_block(name,db,1)); }

void dazz_close(Dazz_Block *db)
{ free(db->kept); free(db->mask_off); free(db->mask_pts);
  free(db->raw); free(db->packed); free(db->poff); free(db->boff); free(db->rlen); free(db->root); free(db->pwd);
  memset(db,0,sizeof(*db));
}

void dazz_complement(Dazz_Block *db)
{ int i;
  float x;
  x = db->freq[0]; db->freq[0] = db->freq[3]; db->freq[3] = x;
  x = db->freq[1]; db->freq[1] = db->freq[2]; db->freq[2] = x;
  for (i = 0; i < db->nreads; i++)
    { uint8_t *s = db->raw + 1 + db->boff[i], *t = s + db->rlen[i] - 1;
      while (s < t)
        { uint8_t c = *s;
          *s++ = (uint8_t) (3 - *t);
          *t-- = (uint8_t) (3 - c);
        }
      if (s == t)
        *s = (uint8_t) (3 - *s);
    }
}

void dazz_view(const Dazz_Block *db, damgpu_block *v)
{ v->bases = (db->raw != NULL) ? db->raw + 1 : NULL;
  v->boff = db->boff;
  v->rlen = db->rlen;
  v->nreads = db->nreads;
  v->tfirst = db->tfirst;
  v->maxlen = db->maxlen;
  v->totlen = db->totlen;
  v->mask_off = db->mask_off; v->mask_pts = db->mask_pts;
  v->packed = NULL; v->poff = NULL; v->packed_bytes = 0;
  /* sizeof_DB, DB.c:1044-1051: sizeof(DAZZ_DB)=112, sizeof(DAZZ_READ)=40 */
  v->sizeof_db = 112 + 40*((int64_t) db->nreads+2) + db->path_len + 1 + (db->totlen + db->nreads + 4);
}

/* ---- -m mask tracks ------------------------------------------------------------------------ */

typedef struct { int32_t b, e; } Ival;

static int ival_cmp(const void *x, const void *y)
{ const Ival *a = (const Ival *) x, *b = (const Ival *) y;
  if (a->b != b->b) return (a->b < b->b ? -1 : 1);
  return (a->e < b->e ? -1 : (a->e > b->e));
}

int dazz_add_mask(Dazz_Block *db, const char *track, const char *prog)
{ char     path[4096];
  FILE    *af = NULL, *df;
  int      ispart = 0, tracklen, size, trimmed, first, count, i, n;
  int      ureads, treads;
  int64_t *anno, *noff;
  int32_t *data, *npts;
  int64_t  dbytes, total;

  if (db->part > 0)
    { snprintf(path,sizeof(path),"%s/.%s.%d.%s.anno",db->pwd,db->root,db->part,track);
      if ((af = fopen(path,"r")) != NULL)
        ispart = 1;
    }
  if (af == NULL)
    { snprintf(path,sizeof(path),"%s/.%s.%s.anno",db->pwd,db->root,track);
      af = fopen(path,"r");
    }
  if (af == NULL)
    return (0);                                        /* Check_Track == -2: not for this DB */
  if (fread(&tracklen,sizeof(int),1,af) != 1 || fread(&size,sizeof(int),1,af) != 1 || size < 0)
    { fprintf(stderr,"%s: track files for %s are corrupted\n",prog,track);
      fclose(af);
      return (-1);
    }
  if (size != 0)
    { fprintf(stderr,"%s: %s track is not a mask track.\n",prog,track);
      fclose(af);
      return (-1);
    }
  ureads = ispart ? db->ulast - db->ufirst : db->db_ureads;
  treads = ispart ? db->nreads : db->db_treads;
  if (tracklen == ureads)
    trimmed = 0;
  else if (tracklen == treads)
    trimmed = 1;
  else
    { printf("%s: Warning: %s track not sync'd with db %s, ignored.\n",prog,track,db->root);
      fclose(af);
      return (0);
    }
  first = ispart ? 0 : (trimmed ? db->tfirst : db->ufirst);
  count = trimmed ? db->nreads : db->ulast - db->ufirst;

  anno = (int64_t *) malloc(sizeof(int64_t)*((size_t) count+1));
  if (fseeko(af,(off_t) (8 + 8*(int64_t) first),SEEK_SET) != 0 ||
      fread(anno,sizeof(int64_t),(size_t) count+1,af) != (size_t) count+1)
    { fprintf(stderr,"%s: Track '%s' annotation file is junk\n",prog,track);
      fclose(af); free(anno);
      return (-1);
    }
  fclose(af);
  if (ispart)
    snprintf(path,sizeof(path),"%s/.%s.%d.%s.data",db->pwd,db->root,db->part,track);
  else
    snprintf(path,sizeof(path),"%s/.%s.%s.data",db->pwd,db->root,track);
  dbytes = anno[count] - anno[0];
  data = (int32_t *) malloc((size_t) dbytes + 8);
  df = fopen(path,"r");
  if (df == NULL || fseeko(df,(off_t) anno[0],SEEK_SET) != 0 ||
      (dbytes > 0 && fread(data,1,(size_t) dbytes,df) != (size_t) dbytes))
    { fprintf(stderr,"%s: Track '%s' data file is junk\n",prog,track);
      if (df) fclose(df);
      free(anno); free(data);
      return (-1);
    }
  fclose(df);

  /* union with what is there, read by read (damapper.c:181-343 computes the same point set up to
     zero-length gaps between abutting intervals, which hold no k-mer) */
  total = dbytes/4 + (db->mask_off ? db->mask_off[db->nreads] : 0);
  noff = (int64_t *) malloc(sizeof(int64_t)*((size_t) db->nreads+1));
  npts = (int32_t *) malloc(sizeof(int32_t)*((size_t) total+2));
  { Ival *iv = (Ival *) malloc(sizeof(Ival)*((size_t) total/2+2));
    int64_t top = 0;
    n = 0;
    for (i = 0; i < count; i++)
      { int64_t j, m = 0;
        if (!trimmed && !db->kept[i])
          continue;
        { const int64_t jb = (anno[i]-anno[0])/4, je = (anno[i+1]-anno[0])/4;
          for (j = jb; j+1 < je; j += 2)
            { iv[m].b = data[j]; iv[m].e = data[j+1]; m += 1; }
        }
        if (db->mask_off != NULL)
          for (j = db->mask_off[n]; j+1 < db->mask_off[n+1]; j += 2)
            { iv[m].b = db->mask_pts[j]; iv[m].e = db->mask_pts[j+1]; m += 1; }
        qsort(iv,(size_t) m,sizeof(Ival),ival_cmp);
        noff[n] = top;
        for (j = 0; j < m; )
          { int32_t b = iv[j].b, e = iv[j].e;
            for (j += 1; j < m && iv[j].b <= e; j++)
              if (iv[j].e > e) e = iv[j].e;
            npts[top++] = b; npts[top++] = e;
          }
        n += 1;
      }
    noff[n] = top;
    free(iv);
  }
  free(anno); free(data);
  free(db->mask_off); free(db->mask_pts);
  db->mask_off = noff; db->mask_pts = npts;
  return (1);
}
