// Context, scratch arena, device-memory helpers and per-kernel event profiling of the C ABI.
#include "common.cuh"

thread_local char g_pob_cuda_err[256] = "";

static const char* k_names[POB_K_COUNT] = {"viterbi_ctc", "viterbi_flipflop", "nw_band_fill", "nw_traceback",
                                           "envelope",    "beam_pair",        "beam_single",  "backtrace",
                                           "forward",     "acceptor",         "prefix_search", "pair_gamma",
                                           "pair_prefix_search"};

extern "C" {

int pob_abi_version(void) { return POB_ABI_VERSION; }

const char* pob_strerror(int s) {
  switch (s) {
    case POB_OK: return "ok";
    case POB_EINVAL: return "invalid argument";
    case POB_ECUDA: return "CUDA error";
    case POB_ENOMEM: return "out of memory";
    case POB_EALIGN: return "float32 read does not start on a 4-row boundary";
    case POB_EUNSUPPORTED: return "unsupported configuration";
  }
  return "unknown status";
}
const char* pob_last_cuda_error(void) { return g_pob_cuda_err; }
const char* pob_kernel_name(int id) { return (id >= 0 && id < POB_K_COUNT) ? k_names[id] : "?"; }

int pob_device_count(int* n) {
  if (!n) return POB_EINVAL;
  cudaError_t e = cudaGetDeviceCount(n);
  if (e != cudaSuccess) {
    *n = 0;
    snprintf(g_pob_cuda_err, sizeof(g_pob_cuda_err), "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return POB_ECUDA;
  }
  return POB_OK;
}

int pob_ctx_create(int device, pob_ctx** out) {
  if (!out) return POB_EINVAL;
  *out = nullptr;
  POB_CUDA(cudaSetDevice(device));
  pob_ctx* c = new pob_ctx();
  c->device = device;
  c->pinned = nullptr;
  c->pinned_size = 0;
  c->prof_on = 0;
  memset(c->prof_ms, 0, sizeof(c->prof_ms));
  memset(c->prof_n, 0, sizeof(c->prof_n));
  memset(c->counters, 0, sizeof(c->counters));
  c->d_counters = nullptr;
  c->t0 = c->t1 = nullptr;
  cudaDeviceProp p;
  POB_CUDA(cudaGetDeviceProperties(&p, device));
  c->sm_count = p.multiProcessorCount;
  POB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  POB_CUDA(cudaMalloc(&c->d_counters, 4 * sizeof(unsigned long long)));
  POB_CUDA(cudaMemsetAsync(c->d_counters, 0, 4 * sizeof(unsigned long long), c->stream));
  *out = c;
  return POB_OK;
}

int pob_ctx_destroy(pob_ctx* c) {
  if (!c) return POB_EINVAL;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (auto& r : c->prof_pending) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  for (auto e : c->prof_pool) cudaEventDestroy(e);
  for (auto& b : c->blocks) cudaFree(b.ptr);
  if (c->pinned) cudaFreeHost(c->pinned);
  if (c->d_counters) cudaFree(c->d_counters);
  cudaStreamDestroy(c->stream);
  delete c;
  return POB_OK;
}

int pob_ctx_sync(pob_ctx* c) {
  if (!c) return POB_EINVAL;
  POB_CUDA(cudaStreamSynchronize(c->stream));
  return POB_OK;
}
void* pob_ctx_stream(pob_ctx* c) { return c ? (void*)c->stream : nullptr; }
int pob_ctx_device(pob_ctx* c) { return c ? c->device : -1; }

int pob_malloc(pob_ctx* c, size_t bytes, void** d) {
  if (!c || !d) return POB_EINVAL;
  POB_CUDA(cudaSetDevice(c->device));
  cudaError_t e = cudaMalloc(d, bytes ? bytes : 1);
  if (e != cudaSuccess) {
    snprintf(g_pob_cuda_err, sizeof(g_pob_cuda_err), "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
    return POB_ENOMEM;
  }
  return POB_OK;
}
int pob_free(pob_ctx* c, void* d) {
  if (!c) return POB_EINVAL;
  POB_CUDA(cudaSetDevice(c->device));
  POB_CUDA(cudaFree(d));
  return POB_OK;
}
int pob_malloc_host(size_t bytes, void** h) {
  if (!h) return POB_EINVAL;
  POB_CUDA(cudaMallocHost(h, bytes ? bytes : 1));
  return POB_OK;
}
int pob_free_host(void* h) {
  POB_CUDA(cudaFreeHost(h));
  return POB_OK;
}
int pob_memcpy_h2d(pob_ctx* c, void* dst, const void* src, size_t bytes) {
  if (!c) return POB_EINVAL;
  POB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
  return POB_OK;
}
int pob_memcpy_d2h(pob_ctx* c, void* dst, const void* src, size_t bytes) {
  if (!c) return POB_EINVAL;
  POB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
  return POB_OK;
}

int pob_timer_start(pob_ctx* c) {
  if (!c) return POB_EINVAL;
  if (!c->t0) {
    POB_CUDA(cudaEventCreate(&c->t0));
    POB_CUDA(cudaEventCreate(&c->t1));
  }
  POB_CUDA(cudaEventRecord(c->t0, c->stream));
  return POB_OK;
}
int pob_timer_stop(pob_ctx* c, double* ms) {
  if (!c || !ms || !c->t0) return POB_EINVAL;
  POB_CUDA(cudaEventRecord(c->t1, c->stream));
  POB_CUDA(cudaEventSynchronize(c->t1));
  float f = 0;
  POB_CUDA(cudaEventElapsedTime(&f, c->t0, c->t1));
  *ms = f;
  return POB_OK;
}

int pob_profile_enable(pob_ctx* c, int on) {
  if (!c) return POB_EINVAL;
  c->prof_on = on;
  return POB_OK;
}

static int prof_collect(pob_ctx* c) {
  POB_CUDA(cudaStreamSynchronize(c->stream));
  for (auto& r : c->prof_pending) {
    float ms = 0;
    POB_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
    c->prof_ms[r.id] += ms;
    c->prof_n[r.id] += 1;
    c->prof_pool.push_back(r.a);
    c->prof_pool.push_back(r.b);
  }
  c->prof_pending.clear();
  return POB_OK;
}

int pob_profile_reset(pob_ctx* c) {
  if (!c) return POB_EINVAL;
  POB_TRY(prof_collect(c));
  memset(c->prof_ms, 0, sizeof(c->prof_ms));
  memset(c->prof_n, 0, sizeof(c->prof_n));
  c->counters[2] = 0;
  return POB_OK;
}

int pob_profile_get(pob_ctx* c, int id, double* ms, int64_t* n) {
  if (!c || id < 0 || id >= POB_K_COUNT) return POB_EINVAL;
  POB_TRY(prof_collect(c));
  if (ms) *ms = c->prof_ms[id];
  if (n) *n = c->prof_n[id];
  return POB_OK;
}

int pob_counters(pob_ctx* c, int64_t* out3) {
  if (!c || !out3) return POB_EINVAL;
  unsigned long long h[4];
  POB_CUDA(cudaMemcpyAsync(h, c->d_counters, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  POB_CUDA(cudaStreamSynchronize(c->stream));
  out3[0] = (int64_t)h[0];
  out3[1] = (int64_t)h[1];
  out3[2] = c->counters[2];
  return POB_OK;
}

}  // extern "C"

pob_prof_scope::pob_prof_scope(pob_ctx* c, int id) : ctx(c), idx(-1) {
  c->counters[2] += 1;
  if (!c->prof_on) return;
  pob_prof_rec r;
  r.id = id;
  cudaEvent_t ev[2];
  for (int i = 0; i < 2; ++i) {
    if (!c->prof_pool.empty()) {
      ev[i] = c->prof_pool.back();
      c->prof_pool.pop_back();
    } else {
      cudaEventCreate(&ev[i]);
    }
  }
  r.a = ev[0];
  r.b = ev[1];
  cudaEventRecord(r.a, c->stream);
  c->prof_pending.push_back(r);
  idx = (int)c->prof_pending.size() - 1;
}
pob_prof_scope::~pob_prof_scope() {
  if (idx >= 0) cudaEventRecord(ctx->prof_pending[idx].b, ctx->stream);
}

// Start a new API call: every block becomes free again.  Blocks are kept as they are: a call makes the same requests
// in the same order as the previous one of its size, so first-fit over the existing blocks serves it without touching
// the allocator; only a larger call adds a block.  (Blocks used to be merged into one here: cudaFree synchronises the
// whole device -- stalling the other GPU call in flight -- and re-allocating tens of GB cost the command line about a
// second per context.)  When the blocks have become many small ones, they are merged once.
int pob_arena_reset(pob_ctx* c) {
  if (c->blocks.size() > 24) {
    size_t total = 0;
    for (auto& b : c->blocks) total += b.size;
    POB_CUDA(cudaStreamSynchronize(c->stream));
    for (auto& b : c->blocks) cudaFree(b.ptr);
    c->blocks.clear();
    char* p = nullptr;
    if (cudaMalloc(&p, total) == cudaSuccess) {
      c->blocks.push_back({p, total, 0});
    } else {
      cudaGetLastError();
    }
  }
  for (auto& b : c->blocks) b.used = 0;
  return POB_OK;
}

void* pob_arena_take(pob_ctx* c, size_t bytes) {
  bytes = pob_align_up(bytes ? bytes : 1, 256);
  for (auto& b : c->blocks) {
    if (b.used + bytes <= b.size) {
      void* p = b.ptr + b.used;
      b.used += bytes;
      return p;
    }
  }
  size_t want = pob_align_up(bytes + bytes / 8, (size_t)8 << 20);
  char* p = nullptr;
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    want = bytes;
    e = cudaMalloc(&p, want);
  }
  if (e != cudaSuccess) {
    snprintf(g_pob_cuda_err, sizeof(g_pob_cuda_err), "arena cudaMalloc(%zu): %s", want, cudaGetErrorString(e));
    cudaGetLastError();
    return nullptr;
  }
  c->blocks.push_back({p, want, bytes});
  return p;
}
