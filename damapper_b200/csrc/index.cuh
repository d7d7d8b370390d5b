// Device-resident DB block and k-mer index (internal).
#pragma once
#include <vector>
#include "common.cuh"

namespace damgpu {

constexpr int BLOCK_SLACK = 256;   // readable bytes before bases[-1] and after the last terminator

// Device image of a loaded DB block (Load_All_Reads, DB.c:1389-1441): bases[-1] == 4, read i at
// bases[boff[i] .. boff[i]+rlen[i]) followed by a 4.
struct DeviceBlock
{ uint8_t *raw = nullptr;       // allocation
  uint8_t *bases = nullptr;     // raw+BLOCK_SLACK, 16-byte aligned
  int64_t *boff = nullptr;      // nreads+1
  int32_t *rlen = nullptr;
  int64_t *mask_off = nullptr;  // -m: merged mask track (nreads+1 offsets into mask_pts), or null
  int32_t *mask_pts = nullptr;  //     interval end points, pairs [begin, end)
  int64_t  nmask = 0;           //     number of points
  mutable int32_t *tile_tab = nullptr;   // first / last read of every 4096-position tile (kmer_filter.cu)
  int      nreads = 0, tfirst = 0, maxlen = 0;
  int64_t  totlen = 0, total = 0, sizeof_db = 0;
  uint64_t uid = 0;             // identity of the block (kept by complement_block): key of the filtered reads list
  std::vector<int64_t> h_boff;
  std::vector<int32_t> h_rlen;
};

// Result of Sort_Kmers: len records + the 2 sentinels (map.c:772-773), on the device.
struct KmerIndex
{ KmerPos *list = nullptr;
  int      len = 0;
  float    ms_extract = 0.f, ms_sort = 0.f;
  int      npass = 0;
  DeviceBlock *block = nullptr;   // block the list was built from, owned when set (layer 1)
  mutable uint32_t *lut = nullptr;   // prefix table over the code, built by the first merge-join that
                                  // searches this list (seed_join.cu), released with the index
  int      limit_len = -1;        // filtered view: the length of the whole list (`alen` of the hit cap,
                                  // map.c:3002-3007); -1 = len
  // deferred reads index (kmer_filter.cu): list == nullptr, len = the count, until something needs it
  bool     deferred = false;
  const DeviceBlock *src = nullptr;  // block to extract from (must outlive the index)
  int      K = 0;
  KmerIndex *filt = nullptr;      // sub-list of the records whose code occurs in one reference block
  // what `filt` was built for.  Exact key: the identity of the reference block (the driver complements
  // a block in place, so both orientations carry the same uid).  A list of another block object (layer 1
  // uploads the two orientations separately; an imported list has no block) is recognised by a 128-bit
  // orientation-invariant multiset hash of its codes plus its length.
  uint64_t filt_uid = 0, src_uid = 0;   // src_uid: uid of the block this list was built from (0: unknown)
  unsigned long long *filt_dsig = nullptr;   // device: the two 64-bit sums of the list `filt` was built for
  unsigned long long filt_sig[2] = { 0, 0 }; // host copy, fetched when first needed
  bool     filt_sig_known = false;
  int      filt_blen = 0, nfilt = 0;
};

DeviceBlock *upload_block(const uint8_t *bases, const int64_t *boff, const int32_t *rlen,
                          int nreads, int tfirst, int maxlen, int64_t totlen, int64_t sizeof_db,
                          cudaStream_t stream);
// attach the block's merged mask track (host arrays; no-op when mask_off is null)
void         set_block_mask(DeviceBlock *blk, const int64_t *mask_off, const int32_t *mask_pts,
                            cudaStream_t stream);
// the same block from its .bps image (2 bits per base), expanded on the device
DeviceBlock *upload_block_packed(const uint8_t *packed, const int64_t *poff, int64_t packed_bytes,
                                 const int64_t *boff, const int32_t *rlen, int nreads, int tfirst,
                                 int maxlen, int64_t totlen, int64_t sizeof_db, cudaStream_t stream);
void         free_block(DeviceBlock *blk);
// complement_DB(block, inplace) of the reference driver (damapper.c:433-469), on the device
void         complement_block(DeviceBlock *blk, cudaStream_t stream);
// reverse-complemented copy of every read into dst_bases (same layout as blk->bases)
void         revcomp_copy_block(const DeviceBlock *blk, uint8_t *dst_bases, cudaStream_t stream);
KmerIndex   *sort_kmers(const DeviceBlock *blk, int K, int suppress, cudaStream_t stream);
void         free_index(KmerIndex *idx);
// kmer_filter.cu: Sort_Kmers of a reads block with the list left unbuilt; the list the merge-join of
// reads index a against reference index b runs on (a, or its filtered view); build the whole list
KmerIndex   *sort_kmers_deferred(const DeviceBlock *blk, int K, int suppress, cudaStream_t stream);
const KmerIndex *reads_view(const KmerIndex *a, const KmerIndex *b, cudaStream_t stream);
void         materialize_index(KmerIndex *idx, cudaStream_t stream);
void         ensure_tile_tab(const DeviceBlock *blk, cudaStream_t stream);   // blk->tile_tab (4096-position tiles)
extern int   g_filter_mode, g_filter_log2;
extern float g_filter_times[4];

}  // namespace damgpu
