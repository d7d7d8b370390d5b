/* libdamgpu -- C ABI of the B200-native DAMAPPER mapping core.
 *
 * The library replaces everything behind the reference's map.h seam (reference map.h:16-39):
 * Set_Filter_Params, Sort_Kmers, Match_Filter and Reporter, with all compute in sm_100a CUDA
 * kernels.  There is no CPU fallback: every entry point fails through the fatal callback
 * (the reference's Clean_Exit, map.h:39) when no CUDA device is usable.
 *
 * Two layers are exported:
 *   1. damgpu_Set_Filter_Params / damgpu_Sort_Kmers / damgpu_Match_Filter / damgpu_Reporter
 *      mirror the four map.h functions call for call (same argument order and meaning), taking a
 *      plain-C view of the DAZZ_DB fields the core reads (damgpu_block) instead of DAZZ_DB*.
 *   2. handle-based entry points (damgpu_block_*, damgpu_index_*, damgpu_seeds_*, ...) that keep
 *      blocks and indices resident in HBM between calls; layer 1 is written on top of them and
 *      the parity tests use them to compare every intermediate array with the oracle.
 *
 * All functions are called from one host thread (the reference's core is non-reentrant as
 * well: file-scope statics, map.c:164-171,453-455,867-870,1011-1017,2885).
 */
#ifndef LIBDAMGPU_H
#define LIBDAMGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* KmerPos / SeedPair, reference map.c:78-89 (little-endian layout) */
typedef struct { uint64_t code; int32_t rpos; int32_t read; } damgpu_kmer;
typedef struct { int32_t diag, apos, bread, aread; } damgpu_seed;

/* The fields of a loaded DAZZ_DB block that the core reads (reference DB.h:390-420 and
 * Load_All_Reads DB.c:1389-1441): bases[-1] == 4, read i occupies
 * bases[boff[i] .. boff[i]+rlen[i]) and is followed by a 4; boff has nreads+1 entries
 * (DAZZ_READ.boff), rlen nreads entries (DAZZ_READ.rlen). */
typedef struct {
  const uint8_t *bases;
  const int64_t *boff;
  const int32_t *rlen;
  int32_t  nreads;
  int32_t  tfirst;          /* DAZZ_DB.tfirst */
  int32_t  maxlen;          /* DAZZ_DB.maxlen */
  int64_t  totlen;          /* DAZZ_DB.totlen */
  int64_t  sizeof_db;       /* sizeof_DB(block), DB.c:1044 -- enters the k-mer hit cap */
  /* -m: the single merged mask track of the block (block->tracks after damapper.c:381-399): read i
   * is masked over [mask_pts[j], mask_pts[j+1]) for j = mask_off[i], mask_off[i]+2, .. below
   * mask_off[i+1] (offsets count ints, as after the division at damapper.c:385-387); k-mers that
   * touch a masked base are not indexed (tuple_thread, map.c:481-543).  NULL = no mask. */
  const int64_t *mask_off;  /* nreads+1 entries, or NULL */
  const int32_t *mask_pts;
  /* Optional: the reads as the .bps file holds them (Compress_Read, DB.c:319-340: four bases per byte,
   * first base in the two top bits; read i starts at packed[poff[i]]).  When `packed` is not NULL the
   * library uploads these bytes (a quarter of the volume) and expands them on the device, and `bases`
   * may be NULL; boff/rlen still describe the Load_All_Reads image (DB.c:1389-1441). */
  const uint8_t *packed;
  const int64_t *poff;      /* nreads entries */
  int64_t  packed_bytes;
} damgpu_block;

/* The globals map.h:16-23 declares extern and damapper.c:58-65 defines. */
typedef struct {
  int32_t     verbose;      /* VERBOSE: the core's statistics on stdout, as map.c:692-697,792-814,2990-3071,3185-3208 */
  int32_t     profile;      /* PROFILE  (-p) */
  int32_t     spacing;      /* SPACING  (-s) */
  double      best_tie;     /* BEST_TIE (-n) */
  const char *sort_path;    /* SORT_PATH */
  uint64_t    mem_limit;    /* MEM_LIMIT, bytes; 0 = no cap (-M0) */
  uint64_t    mem_physical; /* MEM_PHYSICAL */
} damgpu_options;

/* What New_Align_Spec(ave_corr, spacing, freq, reach=1) is given (align.c:222, damapper.c:796).
 * The opaque Align_Spec cannot cross the ABI, the library rebuilds the tables itself. */
typedef struct {
  double  ave_corr;         /* -e */
  int32_t trace_space;      /* -s */
  float   freq[4];          /* base frequencies of the reference DB */
} damgpu_align_spec;

typedef struct damgpu_dblock damgpu_dblock;   /* a DB block resident in HBM      */
typedef struct damgpu_index  damgpu_index;    /* a sorted k-mer list in HBM      */
typedef struct damgpu_seeds  damgpu_seeds;    /* sorted seed hits in HBM         */

/* ---- library ------------------------------------------------------------------------ */
int  damgpu_init(int device);                         /* 0 = ok, else no usable CUDA device */
void damgpu_set_options(const damgpu_options *opts);
void damgpu_set_fatal(void (*clean_exit)(int));       /* Clean_Exit, map.h:39 / damapper.c:543 */
const char *damgpu_last_error(void);
uint64_t damgpu_launch_count(void);                   /* kernels launched so far            */
/* HBM of the device in use, so a driver can decide what to keep resident between blocks
 * (the reference sizes its working set from physical memory the same way, damapper.c:74-141) */
int  damgpu_device_memory(uint64_t *free_bytes, uint64_t *total_bytes);
void damgpu_time_kernels(int on);                     /* record per-phase CUDA-event times  */
/* first tier of the alignment phase: 1 = two candidates per warp, one per 16-lane half (default),
   0 = a warp per candidate (the tier that also re-runs what outgrows the first); `slots` is
   ignored; both yield the same records (Local_Alignment, align.c:1727-1946) */
void damgpu_set_align_tier(int tier, int slots);
/* every radix sort since the last reset (with damgpu_time_kernels on): [0]=algorithmic bytes (32 per
 * record and pass), [1]=ms of the passes (CUDA events on the launching stream), [2]=pass
 * launches, [3]=sorts */
void damgpu_radix_totals(double out[4], int reset);
/* ms of the last Sort_Kmers: [0]=extraction kernel, [1]=all radix passes, [2]=#passes */
void damgpu_last_sort_times(float out[3]);
/* of the last merge-join (with damgpu_time_kernels on): [0]=ms of the prefix table build (0 when the
 * cached table of the reads index was used), [1]=ms of the match kernel, [2]=alen, [3]=blen */
void damgpu_last_join_times(float out[4]);

/* ---- layer 1: the map.h quartet ------------------------------------------------------- */
/* Set_Filter_Params, map.h:25 / map.c:124-150.  Returns 1 if kmer <= 1. */
int   damgpu_Set_Filter_Params(int kmer, int suppress, int nthreads);
/* Sort_Kmers, map.h:27 / map.c:655-822.  Returns an opaque handle (NULL and *len = 0 when the
 * block has no k-mers).  Release with damgpu_index_free (replaces free(), damapper.c:879). */
void *damgpu_Sort_Kmers(const damgpu_block *block, int *len);
/* Match_Filter, map.h:29-30 / map.c:2889-3209.  ablock/atable = reads, bblock/btable =
 * reference block (consumed: btable is released, as the reference frees it, map.c:3181-3182). */
void  damgpu_Match_Filter(const damgpu_block *ablock, const damgpu_block *bblock,
                          void *atable, int alen, void *btable, int blen, int comp, int start);
/* Reporter, map.h:35-36 / map.c:3227-3319.  mflag: bit0 FLAG_DOA, bit1 FLAG_DOB. */
void  damgpu_Reporter(const char *aname, const damgpu_block *ablock, const char *bname,
                      const damgpu_block *bblock, const damgpu_align_spec *spec, int mflag);

/* ---- layer 2: resident handles -------------------------------------------------------- */
damgpu_dblock *damgpu_block_upload(const damgpu_block *block);
/* The same block from the .bps image as it is on disk: read i is
 * rlen[i] bases packed four per byte, first base in the two top bits (Compress_Read /
 * Uncompress_Read, DB.c:319-363), starting at packed[poff[i]]; the Load_All_Reads image
 * (DB.c:1389-1441) is built on the device, so a quarter of the bytes cross PCIe and the host
 * never expands a base.  block->bases is ignored (may be NULL), the other fields are as above. */
damgpu_dblock *damgpu_block_upload_packed(const damgpu_block *block, const uint8_t *packed,
                                          const int64_t *poff, int64_t packed_bytes);
void           damgpu_block_free(damgpu_dblock *blk);
/* complement_DB(block, inplace=1), damapper.c:433-469, on the device */
void           damgpu_block_complement(damgpu_dblock *blk);
void           damgpu_block_download_bases(const damgpu_dblock *blk, uint8_t *bases);

damgpu_index  *damgpu_index_build(const damgpu_dblock *blk);               /* Sort_Kmers */
/* Sort_Kmers for the READS side of Match_Filter, deferred: Match_Filter uses the reads list only
 * through the merge-join with one reference block's list (map.c:881-1002), where a record whose code
 * does not occur in the reference list contributes nothing.  The call records the block (which must
 * outlive the index); the first damgpu_mapper_match / damgpu_seeds_build / damgpu_Match_Filter
 * against a reference list extracts again with a membership test (blocked Bloom filter of the reference
 * codes in both orientations), compacts in extraction order and sorts the survivors -- a sub-list of the
 * reference's sorted list holding every record that can match, so run pairs, `gram` histogram and
 * seeds are identical.  The filtered list is reused for the complemented block.  With -t or masks, on
 * small blocks, or when the whole list is asked for (download / export / device_ptr, a third
 * reference block) the full list is built as by damgpu_index_build.  damgpu_index_len is the full
 * count either way. */
damgpu_index  *damgpu_index_build_deferred(const damgpu_dblock *blk);
int            damgpu_index_is_deferred(const damgpu_index *idx);
/* mode 0 = never filter, 1 = automatic (default: filter when the reads block has >= 4 M k-mers, the
 * reference list is at most 1.5x as long and fills at most 35 % of a bitmap that stays in L2 (64 MB),
 * and for at most two reference blocks per reads block), 2 = always; log2_bits = size of the bitmap
 * (0 = 32 bits per reference k-mer).  Environment: DAMGPU_FILTER=off|auto|always, DAMGPU_FILTER_BITS */
void           damgpu_set_reads_filter(int mode, int log2_bits);
/* [0]=ms bitmap, [1]=ms filtered extraction, [2]=ms radix passes, [3]=survivors of the last filtered
 * build (with damgpu_time_kernels on) */
void           damgpu_last_filter_times(float out[4]);
int            damgpu_index_len(const damgpu_index *idx);
void           damgpu_index_download(const damgpu_index *idx, damgpu_kmer *out); /* len+2 recs */
void           damgpu_index_free(damgpu_index *idx);
/* raw device pointer of the list (len+2 records), for NCCL broadcast by the caller */
void          *damgpu_index_device_ptr(const damgpu_index *idx);
/* device-to-device copies out of / into a caller-owned device buffer of (len+2)*16 bytes, so a
 * collective library (NCCL through torch.distributed) can move an index between GPUs */
void           damgpu_index_export(const damgpu_index *idx, void *device_dst);
damgpu_index  *damgpu_index_import(const void *device_src, int len);

/* merge-join + seed sort of Match_Filter (map.c:2958-3126) as a separate stage */
damgpu_seeds  *damgpu_seeds_build(const damgpu_index *reads_idx, const damgpu_dblock *reads,
                                  const damgpu_index *ref_idx, const damgpu_dblock *ref);
int64_t        damgpu_seeds_count(const damgpu_seeds *s);
int            damgpu_seeds_limit(const damgpu_seeds *s);
void           damgpu_seeds_histogram(const damgpu_seeds *s, int64_t *histo /*[10000]*/);
void           damgpu_seeds_download(const damgpu_seeds *s, damgpu_seed *out); /* count+1 recs */
void           damgpu_seeds_free(damgpu_seeds *s);

/* ---- the per-reads-block mapper: state that persists from Match_Filter calls to Reporter
 *      (static Report_Arg *parmr, reference map.c:1441-1461,2885) ------------------------------ */
typedef struct damgpu_mapper damgpu_mapper;
typedef struct damgpu_report damgpu_report;

/* Candidate chain, reference map.c:1386-1397 (read = index of the read in its block) */
typedef struct { int32_t read, score, length, bread, comp, afirst, alast, bfirst, blast; } damgpu_candidate;

damgpu_mapper *damgpu_mapper_new(const damgpu_dblock *reads, const damgpu_index *reads_idx);
void           damgpu_mapper_free(damgpu_mapper *m);
/* Match_Filter for one reference block in one orientation, from resident handles.  `ref` must
 * already be complemented when comp != 0; start != 0 resets all candidate lists. */
void           damgpu_mapper_match(damgpu_mapper *m, const damgpu_dblock *ref,
                                   const damgpu_index *ref_idx, int comp, int start);
/* chaining alone from a resident seed set (stage-level parity) */
void           damgpu_mapper_chain(damgpu_mapper *m, const damgpu_seeds *s, int bstart, int comp,
                                   int start);
int64_t        damgpu_mapper_last_hits(const damgpu_mapper *m);
int            damgpu_mapper_last_limit(const damgpu_mapper *m);   /* the k-mer hit cap of the last match (map.c:2992-3052) */
int64_t        damgpu_mapper_num_candidates(const damgpu_mapper *m);
/* candidates in (read, list order); jcnt[i] pairs per candidate, jumps = (da,db) int32 pairs;
 * returns the total number of pairs (call with jumps = NULL to size) */
int64_t        damgpu_mapper_get_candidates(const damgpu_mapper *m, damgpu_candidate *out,
                                            int32_t *jcnt, int32_t *jumps, int64_t jmax);
int64_t        damgpu_mapper_get_cover(const damgpu_mapper *m, int16_t *out, int64_t max);

/* Reporter on resident handles: returns the canonical record streams (40-byte records with the
 * 4 padding bytes zeroed + trace bytes, no file header) of the M and R families and the -p track */
damgpu_report *damgpu_mapper_report(damgpu_mapper *m, const damgpu_dblock *wholeref,
                                    const damgpu_align_spec *spec, int mflag);
void           damgpu_report_free(damgpu_report *r);
int64_t        damgpu_report_bytes(const damgpu_report *r, int family /*0=M,1=R,2=prof*/);
int64_t        damgpu_report_records(const damgpu_report *r, int family);
void           damgpu_report_copy(const damgpu_report *r, int family, uint8_t *out);
/* nalign, nwaves, ncells (furthest-reaching cell updates, align.c:887-893), H2 events,
 * overflow-kernel jobs, empty-band events, alignment-kernel ms (when timing is on, as int us),
 * records of both families that fail Check_Trace_Points (align.c:3194-3236; every record is checked
 * on the device before it is copied back -- the LAcheck step of HPC.damapper.c:453-498; must be 0) */
void           damgpu_report_stats(const damgpu_report *r, int64_t out[8]);
/* write <dir>/<aname>.<bname>.M<i>.las (family 0) or <dir>/<bname>.<aname>.R<i>.las (family 1),
 * i = 1..nfiles, reads split as (i*nreads)>>log2(nfiles) (map.c:3148,3250-3261); returns 0 */
int            damgpu_report_write_las(const damgpu_report *r, int family, const char *dir,
                                       const char *aname, const char *bname, int nfiles,
                                       int tspace);
/* write ./.<aname>.prof.anno/.data (map.c:3295-3318) into dir; returns 0 */
int            damgpu_report_write_profile(const damgpu_report *r, const damgpu_block *reads,
                                           const char *dir, const char *aname, int tspace);

#ifdef __cplusplus
}
#endif
#endif
