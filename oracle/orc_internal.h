/* TEST INFRASTRUCTURE ONLY -- private state of the oracle mapper (see damapper_oracle.h). */
#ifndef ORC_INTERNAL_H
#define ORC_INTERNAL_H

#include "damapper_oracle.h"

typedef struct
  { orc_candidate c;
    int           next;     /* next candidate of the same read, -1 = end, -2 = freed */
    int32_t      *jumps;    /* c.length (da,db) pairs, chain end towards chain start */
  } orc_cand;

struct orc_mapper
  { orc_params par;
    orc_block  reads;
    orc_kmer  *bidx;        /* reads index (damapper.c:833) */
    int        blen;

    orc_cand  *cand;        /* pool; lists are newest-first (map.c:1722-1723) */
    int        ncand, cmax, nlive;
    int       *head;        /* per read: DAZZ_READ.coff (map.c:1875) */

    int16_t   *cover;       /* -p difference array (map.c:1580-1587) */
    int64_t   *coff;        /* per read offset into cover, nreads+1 */

    int64_t    last_nhits;
    int        last_limit;

    uint8_t   *abuf, *bbuf, *prof;
    int64_t    alen, amax, anrec, blen_out, bmax, bnrec, proflen;
    int64_t    nalign, nwaves, ncells, h2_events;
  };

#endif
