// C ABI of libdamgpu (include/libdamgpu.h): thin extern "C" layer over the CUDA stages.
#include <stdarg.h>
#include <string.h>
#include <string>
#include "../../include/libdamgpu.h"
#include "common.cuh"
#include "index.cuh"
#include "seeds.cuh"
#include "mapper.cuh"

namespace damgpu {

unsigned long long g_launches = 0;
bool               g_time_kernels = false;
static void      (*g_clean_exit)(int) = nullptr;
static std::string g_last_error;
static int         g_sms = 0;
static bool        g_ready = false;
static float       g_sort_times[3] = { 0, 0, 0 };

Params g_par;          // filter parameters + the map.h globals

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

static void need_gpu()
{ if (!g_ready)
    { if (damgpu_init(-1) != 0)
        fatal("no usable CUDA device (%s); libdamgpu has no CPU path", g_last_error.c_str());
    }
}

static DeviceBlock *upload(const damgpu_block *b)
{ return upload_block(b->bases, b->boff, b->rlen, b->nreads, b->tfirst, b->maxlen, b->totlen,
                      b->sizeof_db, 0);
}

}  // namespace damgpu

using namespace damgpu;

extern "C" {

int damgpu_init(int device)
{ int n = 0;
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
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) { g_last_error = cudaGetErrorString(e); return 1; }
  if (prop.major != 10)
    { g_last_error = "device is not sm_100 (Blackwell B200)";
      return 1;
    }
  g_sms = prop.multiProcessorCount;
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
void damgpu_time_kernels(int on) { g_time_kernels = (on != 0); }
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

void damgpu_block_free(damgpu_dblock *blk) { free_block(reinterpret_cast<DeviceBlock *>(blk)); }

void damgpu_block_complement(damgpu_dblock *blk)
{ complement_block(reinterpret_cast<DeviceBlock *>(blk), 0); }

void damgpu_block_download_bases(const damgpu_dblock *b, uint8_t *bases)
{ const DeviceBlock *blk = reinterpret_cast<const DeviceBlock *>(b);
  CUDA_CHECK(cudaMemcpy(bases - 1, blk->bases - 1, (size_t) blk->total + 1, cudaMemcpyDeviceToHost));
}

damgpu_index *damgpu_index_build(const damgpu_dblock *blk)
{ need_gpu();
  if (g_par.kmer <= 1)
    fatal("Sort_Kmers called before Set_Filter_Params");
  KmerIndex *idx = sort_kmers(reinterpret_cast<const DeviceBlock *>(blk), g_par.kmer,
                              g_par.suppress, 0);
  g_sort_times[0] = idx->ms_extract; g_sort_times[1] = idx->ms_sort; g_sort_times[2] = (float) idx->npass;
  return reinterpret_cast<damgpu_index *>(idx);
}

int damgpu_index_len(const damgpu_index *idx)
{ return idx ? reinterpret_cast<const KmerIndex *>(idx)->len : 0; }

void damgpu_index_download(const damgpu_index *i, damgpu_kmer *out)
{ const KmerIndex *idx = reinterpret_cast<const KmerIndex *>(i);
  if (idx->len > 0)
    CUDA_CHECK(cudaMemcpy(out, idx->list, sizeof(KmerPos) * ((size_t) idx->len + 2),
                          cudaMemcpyDeviceToHost));
}

void damgpu_index_free(damgpu_index *idx) { free_index(reinterpret_cast<KmerIndex *>(idx)); }

void *damgpu_index_device_ptr(const damgpu_index *idx)
{ return reinterpret_cast<const KmerIndex *>(idx)->list; }

damgpu_index *damgpu_index_adopt(void *device_list, int len)
{ KmerIndex *idx = new KmerIndex();
  idx->list = (KmerPos *) device_list;
  idx->len = len;
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

// ---- layer 1 ----------------------------------------------------------------------------

void *damgpu_Sort_Kmers(const damgpu_block *block, int *len)
{ need_gpu();
  DeviceBlock *blk = upload(block);
  damgpu_index *idx = damgpu_index_build(reinterpret_cast<damgpu_dblock *>(blk));
  free_block(blk);
  *len = damgpu_index_len(idx);
  if (*len == 0)
    { damgpu_index_free(idx);
      return nullptr;
    }
  return idx;
}

}  // extern "C"
