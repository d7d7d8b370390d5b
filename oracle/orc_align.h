/* TEST INFRASTRUCTURE ONLY -- private types of the oracle's wave aligner (orc_align.c). */
#ifndef ORC_ALIGN_H
#define ORC_ALIGN_H

#include "damapper_oracle.h"

typedef struct { int ptr, diag, diff, mark; } orc_pebble;     /* Pebble, align.c:344-349 */

typedef struct                                                /* Path, align.h:89-95 */
  { int abpos, bbpos, aepos, bepos, diffs, tlen;
    uint16_t *trace;
  } orc_path;

typedef struct
  { int spacing, ave_path;
    const int16_t *score, *table;
  } orc_aspec;

typedef struct                                                /* _Work_Data, align.c:52-63 */
  { int *V, *M, *HA, *HB, *NA, *NB;
    uint64_t *T;
    orc_pebble *cells;
    int cmax;
    uint16_t *tbuf;
    int tmax;
    int64_t nalign, nwaves, ncells, empty_band;
  } orc_work;

orc_work *orc_work_new(void);
void      orc_work_free(orc_work *w);
void      orc_local_align(orc_work *work, const orc_aspec *spec, const uint8_t *aseq, int alen,
                          const uint8_t *bseq, int blen, int acomp, int low, int hgh, int anti,
                          orc_path *apath, orc_path *bpath);

#endif
