/* Minimal reader of DAZZ_DB databases (.db / .dam stub, .<root>.idx, .<root>.bps) for the
 * damgpu host driver.  Written from the on-disk layout (reference DB.h:285-295,390-435 and
 * DB.c:319-363): it loads one block, trimmed, as one byte per base with 4-terminators -- the
 * image Load_All_Reads builds (DB.c:1389-1441) -- ready to hand to libdamgpu. */
#ifndef DAMGPU_DAZZ_DB_H
#define DAMGPU_DAZZ_DB_H

#include <stdint.h>
#include "../../include/libdamgpu.h"

typedef struct
  { char    *root;        /* root name without extension / block suffix */
    char    *pwd;         /* directory */
    int      isdam;
    int      nblocks;     /* 0 if the DB has not been split */
    int      part;        /* block that was loaded, 0 = whole DB */
    float    freq[4];
    int      cutoff, all;
    /* loaded (trimmed) block */
    int      nreads, tfirst, maxlen;
    int64_t  totlen;
    uint8_t *raw;         /* allocation; bases = raw+1, raw[0] = 4 (NULL when loaded packed) */
    uint8_t *packed;      /* the reads as .bps holds them, 2 bits per base (packed loads only) */
    int64_t *poff;        /* nreads byte offsets into packed */
    int64_t  packed_bytes;
    int64_t *boff;        /* nreads+1 */
    int32_t *rlen;
    int64_t  path_len;    /* strlen(db->path) of the reference, enters sizeof_DB (DB.c:1050) */
    /* what the mask loader needs of the untrimmed block */
    int      ufirst, ulast;      /* untrimmed read range of the block in the whole DB */
    int      db_ureads, db_treads;
    uint8_t *kept;               /* ulast-ufirst flags: read survives the trim */
    /* -m: union of the mask tracks found for this block (damapper.c:352-399), or NULL */
    int64_t *mask_off;           /* nreads+1 offsets into mask_pts, counted in ints */
    int32_t *mask_pts;
  } Dazz_Block;

/* Reads the stub: fills root/pwd/isdam/nblocks/freq.  Returns 0, or -1 after printing why. */
int  dazz_open(const char *name, Dazz_Block *db);
/* Loads block `part` (0 = all, or the .N suffix given in `name`) with all reads in memory. */
int  dazz_load(const char *name, Dazz_Block *db);
/* The same, but the reads stay 2-bit packed (db->packed/poff): the device expands them. */
int  dazz_load_packed(const char *name, Dazz_Block *db);
void dazz_close(Dazz_Block *db);
/* Reads mask track `track` of the loaded block (block-level files first, then the whole-DB track;
   trimmed or untrimmed, DB.c:1649-1702,1804-1990) and merges it into db->mask_off/mask_pts.
   Returns 1 if the track was used, 0 if the DB has no such track or it is not in sync (a warning
   is printed, damapper.c:364-367), -1 on error (after printing why). */
int  dazz_add_mask(Dazz_Block *db, const char *track, const char *prog);
/* In-place reverse complement of every read (complement_DB, damapper.c:433-469). */
void dazz_complement(Dazz_Block *db);
void dazz_view(const Dazz_Block *db, damgpu_block *view);

#endif
