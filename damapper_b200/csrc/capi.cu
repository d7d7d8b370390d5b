// C ABI of libdamgpu (include/libdamgpu.h): thin extern "C" layer over the CUDA stages.
#include <stdarg.h>
#include <string.h>
#include <math.h>
#include <time.h>
#include <string>
#include "../../include/libdamgpu.h"
#include "common.cuh"
#include "index.cuh"
#include "seeds.cuh"
#include "mapper.cuh"
#include "report.cuh"
#include <vector>
#include <map>
#include <unordered_map>

namespace damgpu {

unsigned long long g_launches = 0;
bool               g_time_kernels = false;
static void      (*g_clean_exit)(int) = nullptr;
static std::string g_last_error;
static int         g_sms = 0;
static bool        g_ready = false;
static float       g_sort_times[3] = { 0, 0, 0 };

Params g_par;          // filter parameters + the map.h globals
bool   g_trace = false;
bool   g_debug_sync = false;
int    g_align_tier = 1;
bool   g_chain_async = true;

void trace_mark(const char *name)
{ static double last = 0;
  cudaDeviceSynchronize();
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  const double now = ts.tv_sec * 1e3 + ts.tv_nsec / 1e6;
  if (name != nullptr && last != 0)
    fprintf(stderr, "[trace] %-28s %8.3f ms\n", name, now - last);
  last = now;
}

void fatal(const char *fmt, ...)
{ char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  fprintf(stderr, "damgpu: %s\n", buf);
  fflush(stderr);
  if (g_clean_exit != nullptr)
    g_clean_exit(1);
  exit(1);
}

int sm_count() { return g_sms > 0 ? g_sms : 148; }

// ---- caching device allocator -----------------------------------------------------------
static std::multimap<size_t, void *> g_free_blocks;      // size -> block
static std::unordered_map<void *, size_t> g_block_size;  // every block ever handed out
static size_t g_cached_bytes = 0;

static size_t round_size(size_t b)
{ if (b < (1u << 20)) return (b + 511) & ~(size_t) 511;
  if (b < (64u << 20)) return (b + (1u << 20) - 1) & ~(size_t) ((1u << 20) - 1);
  size_t step = 1;                                        // 1/16 of the enclosing power of two
  while ((step << 5) <= b) step <<= 1;
  return (b + step - 1) & ~(step - 1);
}

void *cache_alloc(size_t bytes)
{ const size_t want = round_size(bytes);
  auto it = g_free_blocks.lower_bound(want);
  if (it != g_free_blocks.end() && it->first <= want + want / 4)
    { void *p = it->second;
      g_cached_bytes -= it->first;
      g_free_blocks.erase(it);
      return p;
    }
  void *p = nullptr;
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess)                                   // give cached memory back and retry
    { cudaGetLastError();
      cudaDeviceSynchronize();
      for (auto &kv : g_free_blocks)
        { cudaFree(kv.second);
          g_block_size.erase(kv.second);
        }
      g_free_blocks.clear();
      g_cached_bytes = 0;
      e = cudaMalloc(&p, want);
    }
  if (e != cudaSuccess)
    fatal("out of device memory allocating %zu bytes (%s)", want, cudaGetErrorString(e));
  g_block_size[p] = want;
  return p;
}

void cache_free(void *p)
{ auto it = g_block_size.find(p);
  if (it == g_block_size.end())
    fatal("cache_free: unknown device pointer");
  g_free_blocks.emplace(it->second, p);
  g_cached_bytes += it->second;
}

// ---- pool of page-locked host buffers (report.cuh: ByteVec) -------------------------------------
static std::multimap<size_t, void *> g_pin_free;
static std::unordered_map<void *, size_t> g_pin_size;     // 0 = plain malloc (page-locking failed)

void *pinned_get(size_t bytes)
{ if (bytes == 0) bytes = 1;
  auto it = g_pin_free.lower_bound(bytes);
  if (it != g_pin_free.end() && it->first <= 2 * bytes + (1u << 20))
    { void *p = it->second;
      g_pin_free.erase(it);
      return p;
    }
  const size_t cap = (bytes + (1u << 20) - 1) & ~(size_t) ((1u << 20) - 1);
  void *p = nullptr;
  if (cudaHostAlloc(&p, cap, cudaHostAllocDefault) == cudaSuccess)
    { g_pin_size[p] = cap;
      return p;
    }
  cudaGetLastError();
  p = malloc(cap);
  if (p == nullptr)
    fatal("out of host memory allocating %zu bytes", cap);
  g_pin_size[p] = 0;
  return p;
}

void pinned_put(void *p)
{ if (p == nullptr) return;
  auto it = g_pin_size.find(p);
  if (it == g_pin_size.end())
    fatal("pinned_put: unknown host pointer");
  if (it->second == 0)
    { free(p);
      g_pin_size.erase(it);
      return;
    }
  g_pin_free.emplace(it->second, p);
}

static void need_gpu()
{ if (!g_ready)
    { if (damgpu_init(-1) != 0)
        fatal("no usable CUDA device (%s); libdamgpu has no CPU path", g_last_error.c_str());
    }
}

static DeviceBlock *upload(const damgpu_block *b)
{ DeviceBlock *blk = (b->packed != nullptr && b->poff != nullptr)
      ? upload_block_packed(b->packed, b->poff, b->packed_bytes, b->boff, b->rlen, b->nreads,
                            b->tfirst, b->maxlen, b->totlen, b->sizeof_db, 0)
      : upload_block(b->bases, b->boff, b->rlen, b->nreads, b->tfirst, b->maxlen,
                     b->totlen, b->sizeof_db, 0);
  set_block_mask(blk, b->mask_off, b->mask_pts, 0);
  return blk;
}

}  // namespace damgpu

using namespace damgpu;

extern "C" {

int damgpu_init(int device)
{ static int ready_dev = -1;
  if (g_ready && (device < 0 || device == ready_dev))    // already set up on this device: nothing to ask the driver
    return 0;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    { g_last_error = (e != cudaSuccess) ? cudaGetErrorString(e) : "no CUDA devices";
      return 1;
    }
  if (device >= 0)
    { e = cudaSetDevice(device);
      if (e != cudaSuccess) { g_last_error = cudaGetErrorString(e); return 1; }
    }
  int dev = 0;
  cudaGetDevice(&dev);
  int major = 0, sms = 0;                                // two attributes, not the whole property record
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) { g_last_error = cudaGetErrorString(e); return 1; }
  if (major != 10)
    { g_last_error = "device is not sm_100 (Blackwell B200)";
      return 1;
    }
  g_sms = sms;
  ready_dev = dev;
  g_trace = (getenv("DAMGPU_TRACE") != nullptr);
  g_debug_sync = (getenv("DAMGPU_DEBUG_SYNC") != nullptr);
  if (const char *t = getenv("DAMGPU_ALIGN"))
    g_align_tier = !strcmp(t, "warp") ? 0 : 1;
  g_chain_async = (getenv("DAMGPU_SYNC_CHAIN") == nullptr);
  if (const char *t = getenv("DAMGPU_RADIX"))
    g_radix_reload = !strcmp(t, "reload");
  g_radix_pf = g_sms;                                    // one tile per SM ahead (flat between 64 and 200 on B200)
  if (const char *t = getenv("DAMGPU_RADIX_PF"))
    g_radix_pf = atoi(t);
  if (const char *t = getenv("DAMGPU_FILTER"))
    g_filter_mode = !strcmp(t, "off") ? 0 : !strcmp(t, "always") ? 2 : 1;
  if (const char *t = getenv("DAMGPU_FILTER_BITS"))
    g_filter_log2 = (atoi(t) >= 10 && atoi(t) <= 32) ? atoi(t) : 0;
  g_ready = true;
  return 0;
}

void damgpu_set_options(const damgpu_options *o)
{ g_par.verbose = o->verbose; g_par.profile = o->profile; g_par.spacing = o->spacing;
  g_par.best_tie = o->best_tie; g_par.sort_path = o->sort_path ? o->sort_path : "/tmp";
  g_par.mem_limit = o->mem_limit; g_par.mem_physical = o->mem_physical;
}

void damgpu_set_fatal(void (*clean_exit)(int)) { g_clean_exit = clean_exit; }
const char *damgpu_last_error(void) { return g_last_error.c_str(); }
uint64_t damgpu_launch_count(void) { return g_launches; }
int damgpu_device_memory(uint64_t *free_bytes, uint64_t *total_bytes)
{ need_gpu();
  size_t f = 0, t = 0;
  if (cudaMemGetInfo(&f, &t) != cudaSuccess)
    return 1;
  if (free_bytes)  *free_bytes = (uint64_t) f + g_cached_bytes;     // the allocator's cache is reusable
  if (total_bytes) *total_bytes = (uint64_t) t;
  return 0;
}
void damgpu_time_kernels(int on) { g_time_kernels = (on != 0); }
void damgpu_set_align_tier(int tier, int slots)
{ g_align_tier = (tier == 0) ? 0 : 1;
  (void) slots;
}
void damgpu_last_join_times(float out[4]) { join_times(out); }
void damgpu_radix_totals(double out[4], int reset) { radix_totals(out, reset); }
void damgpu_last_sort_times(float out[3]) { out[0] = g_sort_times[0]; out[1] = g_sort_times[1]; out[2] = g_sort_times[2]; }

int damgpu_Set_Filter_Params(int kmer, int suppress, int nthreads)   // map.c:124-150
{ if (kmer <= 1)
    return 1;
  g_par.kmer = kmer;
  g_par.suppress = suppress;
  g_par.nthreads = 1;
  g_par.nshift = 0;
  while (2 * g_par.nthreads <= nthreads)
    { g_par.nthreads *= 2;
      g_par.nshift += 1;
    }
  return 0;
}

damgpu_dblock *damgpu_block_upload(const damgpu_block *block)
{ need_gpu();
  return reinterpret_cast<damgpu_dblock *>(upload(block));
}

damgpu_dblock *damgpu_block_upload_packed(const damgpu_block *b, const uint8_t *packed,
                                          const int64_t *poff, int64_t packed_bytes)
{ need_gpu();
  if (b->nreads > 0 && (packed == nullptr || poff == nullptr))
    fatal("damgpu_block_upload_packed: no packed image");
  DeviceBlock *blk = upload_block_packed(packed, poff, packed_bytes, b->boff, b->rlen, b->nreads,
                                         b->tfirst, b->maxlen, b->totlen, b->sizeof_db, 0);
  set_block_mask(blk, b->mask_off, b->mask_pts, 0);
  return reinterpret_cast<damgpu_dblock *>(blk);
}

void damgpu_block_free(damgpu_dblock *blk) { free_block(reinterpret_cast<DeviceBlock *>(blk)); }

void damgpu_block_complement(damgpu_dblock *blk)
{ complement_block(reinterpret_cast<DeviceBlock *>(blk), 0); }

void damgpu_block_download_bases(const damgpu_dblock *b, uint8_t *bases)
{ const DeviceBlock *blk = reinterpret_cast<const DeviceBlock *>(b);
  CUDA_CHECK(cudaMemcpy(bases - 1, blk->bases - 1, (size_t) blk->total + 1, cudaMemcpyDeviceToHost));
}

// Print_Number of the reference (DB.c:253-295): thousands separated by commas (at most three groups are
// split off, as there), right-justified in `width` columns; width 0 = no padding
static void print_number_w(long long num, int width, FILE *out)
{ char digits[32], text[48];
  const int nd = snprintf(digits, sizeof(digits), "%lld", num);
  int groups = (nd - 1) / 3;
  if (groups > 3) groups = 3;
  const int lead = nd - 3 * groups;
  int t = 0;
  for (int i = 0; i < nd; i++)
    { if (i >= lead && (i - lead) % 3 == 0)
        text[t++] = ',';
      text[t++] = digits[i];
    }
  text[t] = 0;
  fprintf(out, "%*s", width, text);
}

static void print_number(long long num, FILE *out)
{ print_number_w(num, 0, out); }

// what Sort_Kmers tells the user (map.c:692-697,792-814): -v statistics and the block-size warning
static void sort_kmers_report(const DeviceBlock *blk, const KmerIndex *idx)
{ const long long raw = (long long) blk->total - (long long) g_par.kmer * blk->nreads;
  const long long kmers = idx->len;
  if (raw <= 0)                                          // no_mers (map.c:678-679): nothing is said
    return;
  if (g_par.verbose)
    { printf("\n   Kmer count = ");
      print_number(raw, stdout);
      printf("\n   Using %.2fGb of space\n", (1. * raw) / 33554432);
      if (g_par.suppress > 0 || blk->mask_off != nullptr)
        { printf("   Revised kmer count = ");
          print_number(kmers, stdout);
          printf("\n");
        }
      printf("   Index occupies %.2fGb\n", (1. * kmers) / 67108864);
      fflush(stdout);
    }
  if (kmers > (long long) (g_par.mem_limit / (4 * sizeof(KmerPos))))   // also with -M0, as the reference
    { fprintf(stderr, "Warning: Block size too big, index occupies more than 1/4 of");
      if (g_par.mem_limit == g_par.mem_physical)
        fprintf(stderr, " physical memory (%.1fGb)\n", (1. * g_par.mem_limit) / 0x40000000ll);
      else
        fprintf(stderr, " desired memory allocation (%.1fGb)\n", (1. * g_par.mem_limit) / 0x40000000ll);
      fflush(stderr);
    }
}

damgpu_index *damgpu_index_build(const damgpu_dblock *blk)
{ need_gpu();
  if (g_par.kmer <= 1)
    fatal("Sort_Kmers called before Set_Filter_Params");
  KmerIndex *idx = sort_kmers(reinterpret_cast<const DeviceBlock *>(blk), g_par.kmer,
                              g_par.suppress, 0);
  g_sort_times[0] = idx->ms_extract; g_sort_times[1] = idx->ms_sort; g_sort_times[2] = (float) idx->npass;
  sort_kmers_report(reinterpret_cast<const DeviceBlock *>(blk), idx);
  return reinterpret_cast<damgpu_index *>(idx);
}

/* Sort_Kmers for the reads side of Match_Filter, list left unbuilt (kmer_filter.cu) */
damgpu_index *damgpu_index_build_deferred(const damgpu_dblock *blk)
{ need_gpu();
  if (g_par.kmer <= 1)
    fatal("Sort_Kmers called before Set_Filter_Params");
  KmerIndex *idx = sort_kmers_deferred(reinterpret_cast<const DeviceBlock *>(blk), g_par.kmer,
                                       g_par.suppress, 0);
  g_sort_times[0] = idx->ms_extract; g_sort_times[1] = idx->ms_sort; g_sort_times[2] = (float) idx->npass;
  sort_kmers_report(reinterpret_cast<const DeviceBlock *>(blk), idx);
  return reinterpret_cast<damgpu_index *>(idx);
}

int damgpu_index_is_deferred(const damgpu_index *idx)
{ return idx != nullptr && reinterpret_cast<const KmerIndex *>(idx)->deferred; }

void damgpu_set_reads_filter(int mode, int log2_bits)
{ g_filter_mode = (mode < 0 || mode > 2) ? 1 : mode;
  g_filter_log2 = (log2_bits < 10 || log2_bits > 32) ? 0 : log2_bits;
}

void damgpu_last_filter_times(float out[4])
{ for (int i = 0; i < 4; i++) out[i] = g_filter_times[i]; }

int damgpu_index_len(const damgpu_index *idx)
{ return idx ? reinterpret_cast<const KmerIndex *>(idx)->len : 0; }

void damgpu_index_download(const damgpu_index *i, damgpu_kmer *out)
{ const KmerIndex *idx = reinterpret_cast<const KmerIndex *>(i);
  materialize_index(const_cast<KmerIndex *>(idx), 0);
  if (idx->len > 0)
    CUDA_CHECK(cudaMemcpy(out, idx->list, sizeof(KmerPos) * ((size_t) idx->len + 2),
                          cudaMemcpyDeviceToHost));
}

void damgpu_index_free(damgpu_index *idx) { free_index(reinterpret_cast<KmerIndex *>(idx)); }

void *damgpu_index_device_ptr(const damgpu_index *idx)
{ materialize_index(const_cast<KmerIndex *>(reinterpret_cast<const KmerIndex *>(idx)), 0);
  return reinterpret_cast<const KmerIndex *>(idx)->list;
}

void damgpu_index_export(const damgpu_index *i, void *dst)
{ const KmerIndex *idx = reinterpret_cast<const KmerIndex *>(i);
  materialize_index(const_cast<KmerIndex *>(idx), 0);
  if (idx->len > 0)
    CUDA_CHECK(cudaMemcpy(dst, idx->list, sizeof(KmerPos) * ((size_t) idx->len + 2),
                          cudaMemcpyDeviceToDevice));
}

damgpu_index *damgpu_index_import(const void *src, int len)
{ need_gpu();
  KmerIndex *idx = new KmerIndex();
  if (len > 0)
    { idx->list = dalloc<KmerPos>((size_t) len + 2);
      CUDA_CHECK(cudaMemcpy(idx->list, src, sizeof(KmerPos) * ((size_t) len + 2), cudaMemcpyDeviceToDevice));
      idx->len = len;
    }
  return reinterpret_cast<damgpu_index *>(idx);
}

damgpu_seeds *damgpu_seeds_build(const damgpu_index *ai, const damgpu_dblock *ab,
                                 const damgpu_index *bi, const damgpu_dblock *bb)
{ need_gpu();
  SeedSet *ss = merge_join(reinterpret_cast<const KmerIndex *>(ai),
                           reinterpret_cast<const DeviceBlock *>(ab),
                           reinterpret_cast<const KmerIndex *>(bi),
                           reinterpret_cast<const DeviceBlock *>(bb), g_par.kmer, g_par.mem_limit, 0);
  return reinterpret_cast<damgpu_seeds *>(ss);
}

int64_t damgpu_seeds_count(const damgpu_seeds *s) { return reinterpret_cast<const SeedSet *>(s)->nhits; }
int     damgpu_seeds_limit(const damgpu_seeds *s) { return reinterpret_cast<const SeedSet *>(s)->limit; }

void damgpu_seeds_histogram(const damgpu_seeds *s, int64_t *histo)
{ const SeedSet *ss = reinterpret_cast<const SeedSet *>(s);
  for (int i = 0; i < 10000; i++)
    histo[i] = ss->histo.empty() ? 0 : (int64_t) ss->histo[i];
}

void damgpu_seeds_download(const damgpu_seeds *s, damgpu_seed *out)
{ const SeedSet *ss = reinterpret_cast<const SeedSet *>(s);
  if (ss->hits != nullptr)
    CUDA_CHECK(cudaMemcpy(out, ss->hits, sizeof(SeedPair) * ((size_t) ss->nhits + 1),
                          cudaMemcpyDeviceToHost));
}

void damgpu_seeds_free(damgpu_seeds *s) { free_seeds(reinterpret_cast<SeedSet *>(s)); }

// ---- mapper -------------------------------------------------------------------------------

struct MapperH { Mapper *m; const KmerIndex *ridx; long long tfilt = 0; };

// the closing lines of Match_Filter under -v (map.c:3185-3208)
static void match_filter_epilogue(MapperH *h, double atot, double btot, long long nhits, int start)
{ long long nfilt = 0, none = 0, &tfilt = (h != nullptr) ? h->tfilt : none;
  if (start) tfilt = 0;                                  // map.c:2949-2951
  if (nhits > 0 && h != nullptr)
    { const long long live = damgpu_mapper_num_candidates(reinterpret_cast<damgpu_mapper *>(h));
      nfilt = live - tfilt;                              // added minus removed by this call (map.c:1684-1764)
      tfilt = live;
    }
  int width = nhits <= 0 ? 1 : ((int) log10((double) nhits)) + 1;
  width += (width - 1) / 3;
  printf("\n     ");
  print_number_w(nhits, width, stdout);
  printf(" %d-mers (%e of matrix)\n     ", g_par.kmer, (1. * nhits / atot) / btot);
  if (nfilt < 0)
    { print_number_w(-nfilt, width, stdout);
      printf(" candidates removed\n     ");
    }
  else
    { print_number_w(nfilt, width, stdout);
      printf(" candidates added\n     ");
    }
  print_number_w(tfilt, width, stdout);
  printf(" candidates (%e of matrix)\n", (1. * tfilt / atot) / btot);
  fflush(stdout);
}

damgpu_mapper *damgpu_mapper_new(const damgpu_dblock *reads, const damgpu_index *reads_idx)
{ need_gpu();
  MapperH *h = new MapperH();
  h->m = mapper_new(reinterpret_cast<const DeviceBlock *>(reads));
  h->ridx = reinterpret_cast<const KmerIndex *>(reads_idx);
  return reinterpret_cast<damgpu_mapper *>(h);
}

void damgpu_mapper_free(damgpu_mapper *mm)
{ MapperH *h = reinterpret_cast<MapperH *>(mm);
  if (h == nullptr) return;
  mapper_free(h->m);
  delete h;
}

void damgpu_mapper_chain(damgpu_mapper *mm, const damgpu_seeds *s, int bstart, int comp, int start)
{ MapperH *h = reinterpret_cast<MapperH *>(mm);
  if (start) mapper_reset(h->m);
  chain_seeds(h->m, const_cast<SeedSet *>(reinterpret_cast<const SeedSet *>(s)), bstart, comp, 0, false);
}

void damgpu_mapper_match(damgpu_mapper *mm, const damgpu_dblock *ref, const damgpu_index *ref_idx,
                         int comp, int start)
{ MapperH *h = reinterpret_cast<MapperH *>(mm);
  const DeviceBlock *rb = reinterpret_cast<const DeviceBlock *>(ref);
  const KmerIndex *gi = reinterpret_cast<const KmerIndex *>(ref_idx);
  h->m->last_nhits = 0;
  if (h->ridx == nullptr || h->ridx->len == 0 || gi == nullptr || gi->len == 0)   // map.c:2955-2956
    { if (g_par.verbose)
        match_filter_epilogue(h, (double) h->m->reads->totlen, rb != nullptr ? (double) rb->totlen : 0., 0, start);
      return;
    }
  SeedSet *ss = merge_join(h->ridx, h->m->reads, gi, rb, g_par.kmer, g_par.mem_limit, 0);
  h->m->last_nhits = ss->nhits; h->m->last_limit = ss->limit;
  if (g_par.mem_limit > 0 && ss->limit < 10)             // map.c:3029-3039 (limit <= 1 is fatal in merge_join)
    { fprintf(stderr, "\nWarning: Sensitivity hampered by low ");
      if (g_par.mem_limit == g_par.mem_physical)
        fprintf(stderr, " physical memory (%.1fGb), reduce block size\n", (1. * g_par.mem_limit) / 0x40000000ll);
      else
        { fprintf(stderr, " memory allocation (%.1fGb),", (1. * g_par.mem_limit) / 0x40000000ll);
          fprintf(stderr, " reduce block size or increase allocation\n");
        }
      fflush(stderr);
    }
  if (g_par.verbose)                                     // map.c:3040-3071
    { const long long alen = h->ridx->len, blen = gi->len, nh = (long long) ss->nhits;
      printf("\n");
      if (g_par.mem_limit > 0)
        printf("   Capping mutual k-mer matches over %d (effectively -t%d)\n", ss->limit,
               (int) sqrt(1. * ss->limit));
      printf("   Hit count = ");
      print_number(nh, stdout);
      if (nh >= blen)
        printf("\n   Highwater of %.2fGb space\n", (1. * (alen + 2 * nh)) / 67108864);
      else
        printf("\n   Highwater of %.2fGb space\n", (1. * (alen + blen + nh)) / 67108864);
      fflush(stdout);
    }
  if (start) mapper_reset(h->m);
  chain_seeds(h->m, ss, rb->tfirst, comp, 0, g_chain_async);
  if (g_par.verbose)
    match_filter_epilogue(h, (double) h->m->reads->totlen, (double) rb->totlen, (long long) ss->nhits, start);
  free_seeds(ss);
}

int64_t damgpu_mapper_last_hits(const damgpu_mapper *mm)
{ return reinterpret_cast<const MapperH *>(mm)->m->last_nhits; }
int damgpu_mapper_last_limit(const damgpu_mapper *mm)
{ return reinterpret_cast<const MapperH *>(mm)->m->last_limit; }


// host copies of the candidate pool, walked per read (newest first, as the lists are linked)
static void fetch_pool(const Mapper *m, std::vector<Candidate> &cand, std::vector<int> &head,
                       std::vector<uint32_t> &jumps)
{ int ctop = 0; unsigned long long jtop = 0;
  chain_sync();
  CUDA_CHECK(cudaMemcpy(&ctop, m->cand_top, sizeof(int), cudaMemcpyDeviceToHost));
  CUDA_CHECK(cudaMemcpy(&jtop, m->jump_top, sizeof(jtop), cudaMemcpyDeviceToHost));
  cand.resize(ctop); head.resize(m->reads->nreads); jumps.resize(jtop);
  if (ctop) CUDA_CHECK(cudaMemcpy(cand.data(), m->cand, sizeof(Candidate) * ctop, cudaMemcpyDeviceToHost));
  if (jtop) CUDA_CHECK(cudaMemcpy(jumps.data(), m->jumps, sizeof(uint32_t) * jtop, cudaMemcpyDeviceToHost));
  CUDA_CHECK(cudaMemcpy(head.data(), m->head, sizeof(int) * head.size(), cudaMemcpyDeviceToHost));
}

int64_t damgpu_mapper_num_candidates(const damgpu_mapper *mm)
{ const Mapper *m = reinterpret_cast<const MapperH *>(mm)->m;
  std::vector<Candidate> cand; std::vector<int> head; std::vector<uint32_t> jumps;
  fetch_pool(m, cand, head, jumps);
  int64_t n = 0;
  for (size_t r = 0; r < head.size(); r++)
    for (int c = head[r]; c >= 0; c = cand[c].next) n++;
  return n;
}

int64_t damgpu_mapper_get_candidates(const damgpu_mapper *mm, damgpu_candidate *out, int32_t *jcnt,
                                     int32_t *jout, int64_t jmax)
{ const Mapper *m = reinterpret_cast<const MapperH *>(mm)->m;
  std::vector<Candidate> cand; std::vector<int> head; std::vector<uint32_t> jumps;
  fetch_pool(m, cand, head, jumps);
  int64_t n = 0, nj = 0;
  for (size_t r = 0; r < head.size(); r++)
    for (int c = head[r]; c >= 0; c = cand[c].next)
      { const Candidate &C = cand[c];
        if (out)
          { damgpu_candidate &o = out[n];
            o.read = (int) r; o.score = C.score; o.length = C.length; o.bread = C.bread; o.comp = C.comp;
            o.afirst = C.afirst; o.alast = C.alast; o.bfirst = C.bfirst; o.blast = C.blast;
          }
        if (jcnt) jcnt[n] = C.length;
        for (int k = 0; k < C.length; k++)
          { if (jout && nj < jmax)
              { const uint32_t j = jumps[C.chain + k];
                jout[2 * nj] = (int32_t) (j & 0xffff); jout[2 * nj + 1] = (int32_t) (j >> 16);
              }
            nj++;
          }
        n++;
      }
  return nj;
}

int64_t damgpu_mapper_get_cover(const damgpu_mapper *mm, int16_t *out, int64_t max)
{ const Mapper *m = reinterpret_cast<const MapperH *>(mm)->m;
  const int64_t tot = m->h_coff[m->reads->nreads];
  chain_sync();
  if (out != nullptr)
    CUDA_CHECK(cudaMemcpy(out, m->cover, sizeof(int16_t) * (size_t) (tot < max ? tot : max),
                          cudaMemcpyDeviceToHost));
  return tot;
}

damgpu_report *damgpu_mapper_report(damgpu_mapper *mm, const damgpu_dblock *wholeref,
                                    const damgpu_align_spec *spec, int mflag)
{ MapperH *h = reinterpret_cast<MapperH *>(mm);
  if (spec->trace_space != g_par.spacing)
    fatal("Reporter: align spec trace spacing %d differs from SPACING %d", spec->trace_space,
          g_par.spacing);
  ReportOut *r = reporter(h->m, reinterpret_cast<const DeviceBlock *>(wholeref), spec->ave_corr,
                          spec->freq, (mflag & 1) != 0, (mflag & 2) != 0, 0);
  if (r->trace_fails != 0)                               // what LAcheck would reject (align.c:3194)
    fatal("Reporter: %lld records fail Check_Trace_Points", (long long) r->trace_fails);
  return reinterpret_cast<damgpu_report *>(r);
}

void damgpu_report_free(damgpu_report *r) { delete reinterpret_cast<ReportOut *>(r); }

int64_t damgpu_report_bytes(const damgpu_report *rr, int family)
{ const ReportOut *r = reinterpret_cast<const ReportOut *>(rr);
  return family == 0 ? (int64_t) r->a.size() : family == 1 ? (int64_t) r->b.size() : (int64_t) r->prof.size();
}

int64_t damgpu_report_records(const damgpu_report *rr, int family)
{ const ReportOut *r = reinterpret_cast<const ReportOut *>(rr);
  return family == 0 ? r->nrec_a : family == 1 ? r->nrec_b : 0;
}

void damgpu_report_copy(const damgpu_report *rr, int family, uint8_t *out)
{ const ReportOut *r = reinterpret_cast<const ReportOut *>(rr);
  const ByteVec &v = family == 0 ? r->a : family == 1 ? r->b : r->prof;
  if (!v.empty()) memcpy(out, v.data(), v.size());
}

void damgpu_report_stats(const damgpu_report *rr, int64_t out[8])
{ const ReportOut *r = reinterpret_cast<const ReportOut *>(rr);
  out[0] = r->nalign; out[1] = r->nwaves; out[2] = r->ncells; out[3] = r->h2_events;
  out[4] = r->overflow_jobs; out[5] = r->empty_band; out[6] = (int64_t) (r->ms_align * 1000.f); out[7] = r->trace_fails;
}

// per-"thread" files: reads [ (i*n)>>shift, ((i+1)*n)>>shift ), map.c:3148,3250-3261,2421-2428
int damgpu_report_write_las(const damgpu_report *rr, int family, const char *dir, const char *aname,
                            const char *bname, int nfiles, int tspace)
{ const ReportOut *r = reinterpret_cast<const ReportOut *>(rr);
  const ByteVec &v = family == 0 ? r->a : r->b;
  const std::vector<int64_t> &off = family == 0 ? r->read_off_a : r->read_off_b;
  const std::vector<int> &nrec = family == 0 ? r->read_nrec_a : r->read_nrec_b;
  if (off.empty()) return 1;
  const int64_t n = (int64_t) off.size() - 1;
  int shift = 0;
  while ((2 << shift) <= nfiles) shift++;
  const int nf = 1 << shift;
  for (int i = 0; i < nf; i++)
    { const int64_t r0 = (i * n) >> shift, r1 = (i == nf - 1) ? n : (((i + 1) * n) >> shift);
      std::string path = std::string(dir) + "/";
      if (family == 0) path += std::string(aname) + "." + bname + ".M" + std::to_string(i + 1) + ".las";
      else             path += std::string(bname) + "." + aname + ".R" + std::to_string(i + 1) + ".las";
      FILE *f = fopen(path.c_str(), "w");
      if (f == nullptr) return 1;
      int64_t novl = 0;
      for (int64_t x = r0; x < r1; x++) novl += nrec[x];
      int ts = tspace;
      bool ok = (fwrite(&novl, sizeof(int64_t), 1, f) == 1) && (fwrite(&ts, sizeof(int), 1, f) == 1);
      if (ok && off[r1] > off[r0])
        { const size_t nb = (size_t) (off[r1] - off[r0]);
          ok = (fwrite(v.data() + off[r0], 1, nb, f) == nb);
        }
      if (fclose(f) != 0) ok = false;
      if (!ok) return 1;                                 // short write (full disk): the caller stops
    }
  return 0;
}

int damgpu_report_write_profile(const damgpu_report *rr, const damgpu_block *reads, const char *dir,
                                const char *aname, int tspace)
{ const ReportOut *r = reinterpret_cast<const ReportOut *>(rr);
  std::string base = std::string(dir) + "/." + aname + ".prof";
  FILE *af = fopen((base + ".anno").c_str(), "w"), *df = fopen((base + ".data").c_str(), "w");
  if (af == nullptr || df == nullptr)
    { if (af != nullptr) fclose(af);
      if (df != nullptr) fclose(df);
      return 1;
    }
  int size = sizeof(int64_t);
  bool ok = (fwrite(&reads->nreads, sizeof(int), 1, af) == 1) && (fwrite(&size, sizeof(int), 1, af) == 1);
  int64_t cnt = 0;
  for (int a = 0; a < reads->nreads && ok; a++)
    { ok = (fwrite(&cnt, sizeof(int64_t), 1, af) == 1);
      cnt += (reads->rlen[a] - 1) / tspace + 2;
    }
  ok = ok && (fwrite(&cnt, sizeof(int64_t), 1, af) == 1);
  if (ok && !r->prof.empty()) ok = (fwrite(r->prof.data(), 1, r->prof.size(), df) == r->prof.size());
  if (fclose(af) != 0) ok = false;
  if (fclose(df) != 0) ok = false;
  if (!ok) return 1;
  return 0;
}

// ---- layer 1 ----------------------------------------------------------------------------

static MapperH *g_mapper = nullptr;                    // static parmr of the reference (map.c:2885)
static const void *g_mapper_key = nullptr;

void *damgpu_Sort_Kmers(const damgpu_block *block, int *len)
{ need_gpu();
  DeviceBlock *blk = upload(block);
  // reads or reference block is not known yet: deferred, Match_Filter builds what it needs
  KmerIndex *idx = reinterpret_cast<KmerIndex *>(damgpu_index_build_deferred(reinterpret_cast<damgpu_dblock *>(blk)));
  *len = idx->len;
  if (idx->len == 0)
    { free_index(idx);
      free_block(blk);
      return nullptr;
    }
  idx->block = blk;                                    // kept resident with the list
  return idx;
}

void damgpu_Match_Filter(const damgpu_block *ablock, const damgpu_block *bblock, void *atable,
                         int alen, void *btable, int blen, int comp, int start)
{ KmerIndex *ai = reinterpret_cast<KmerIndex *>(atable), *bi = reinterpret_cast<KmerIndex *>(btable);
  if (ai != nullptr && ai->block == nullptr)
    fatal("Match_Filter: the reads index was not made by damgpu_Sort_Kmers (it carries no block)");
  if (alen == 0 || blen == 0 || ai == nullptr || bi == nullptr)     // map.c:2955-2956
    { // the block sizes come from whichever side knows them (the shim of INTEGRATION.md passes no blocks)
      const double atot = (ai != nullptr) ? (double) ai->block->totlen : (ablock != nullptr) ? (double) ablock->totlen : 0.;
      const double btot = (bi != nullptr && bi->block != nullptr) ? (double) bi->block->totlen
                                                                  : (bblock != nullptr) ? (double) bblock->totlen : 0.;
      free_index(bi);
      if (ai != nullptr && (g_mapper == nullptr || g_mapper_key != atable || start))
        { // nothing to match, but the reads block is under way: Reporter writes its (empty) files
          if (g_mapper) damgpu_mapper_free(reinterpret_cast<damgpu_mapper *>(g_mapper));
          g_mapper = reinterpret_cast<MapperH *>(damgpu_mapper_new(reinterpret_cast<damgpu_dblock *>(ai->block),
                                                                   reinterpret_cast<damgpu_index *>(ai)));
          g_mapper_key = atable;
        }
      if (g_par.verbose)                                 // the reference still prints its closing lines
        match_filter_epilogue(ai != nullptr ? g_mapper : nullptr, atot, btot, 0, start);
      return;
    }
  if (g_mapper == nullptr || g_mapper_key != atable)
    { if (g_mapper) damgpu_mapper_free(reinterpret_cast<damgpu_mapper *>(g_mapper));
      g_mapper = reinterpret_cast<MapperH *>(damgpu_mapper_new(reinterpret_cast<damgpu_dblock *>(ai->block),
                                                               reinterpret_cast<damgpu_index *>(ai)));
      g_mapper_key = atable;
    }
  damgpu_mapper_match(reinterpret_cast<damgpu_mapper *>(g_mapper),
                      reinterpret_cast<damgpu_dblock *>(bi->block),
                      reinterpret_cast<damgpu_index *>(bi), comp, start);
  free_index(bi);                                      // map.c:3181-3182
}

void damgpu_Reporter(const char *aname, const damgpu_block *ablock, const char *bname,
                     const damgpu_block *bblock, const damgpu_align_spec *spec, int mflag)
{ need_gpu();
  DeviceBlock *ablk = nullptr;
  KmerIndex   *aidx = nullptr;
  if (g_mapper == nullptr)
    { // no Match_Filter call reached the core for this reads block (no k-mers on one side of every
      // call, map.c:2955-2956): the reference goes on and writes empty files, so does this
      if (ablock == nullptr)
        fatal("Reporter called before Match_Filter");
      ablk = upload(ablock);
      aidx = sort_kmers_deferred(ablk, g_par.kmer, g_par.suppress, 0);
      g_mapper = reinterpret_cast<MapperH *>(damgpu_mapper_new(reinterpret_cast<damgpu_dblock *>(ablk),
                                                               reinterpret_cast<damgpu_index *>(aidx)));
    }
  DeviceBlock *ref = upload(bblock);
  damgpu_report *rep = damgpu_mapper_report(reinterpret_cast<damgpu_mapper *>(g_mapper),
                                            reinterpret_cast<damgpu_dblock *>(ref), spec, mflag);
  free_block(ref);
  if (mflag & 1)
    if (damgpu_report_write_las(rep, 0, g_par.sort_path.c_str(), aname, bname, g_par.nthreads, g_par.spacing))
      fatal("Cannot open .las files in %s for writing", g_par.sort_path.c_str());
  if (mflag & 2)
    if (damgpu_report_write_las(rep, 1, g_par.sort_path.c_str(), aname, bname, g_par.nthreads, g_par.spacing))
      fatal("Cannot open .las files in %s for writing", g_par.sort_path.c_str());
  if (g_par.verbose)
    { printf("      ");                               // map.c:3289-3293
      print_number((long long) damgpu_report_records(rep, (mflag & 1) ? 0 : 1), stdout);
      printf(" mapped segments\n");
      fflush(stdout);
    }
  if (g_par.profile)
    if (damgpu_report_write_profile(rep, ablock, ".", aname, g_par.spacing))
      fatal("Cannot write the .prof track");
  damgpu_report_free(rep);
  damgpu_mapper_free(reinterpret_cast<damgpu_mapper *>(g_mapper));
  g_mapper = nullptr; g_mapper_key = nullptr;
  free_index(aidx);
  free_block(ablk);
}

}  // extern "C"
