// Shared plumbing of the poreover_b200 CUDA library: context, scratch arena, status, profiling.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "../../include/poreover_b200.h"

extern thread_local char g_pob_cuda_err[256];

#define POB_CUDA(expr)                                                                           \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      snprintf(g_pob_cuda_err, sizeof(g_pob_cuda_err), "%s:%d %s: %s", __FILE__, __LINE__, #expr, \
               cudaGetErrorString(_e));                                                          \
      return POB_ECUDA;                                                                          \
    }                                                                                            \
  } while (0)

#define POB_TRY(expr)            \
  do {                           \
    int _s = (expr);             \
    if (_s != POB_OK) return _s; \
  } while (0)

struct pob_prof_rec {
  int id;
  cudaEvent_t a, b;
};

struct pob_arena_block {
  char* ptr;
  size_t size, used;
};

struct pob_ctx {
  int device;
  int sm_count;
  cudaStream_t stream;
  // bump arena for per-call device scratch: a list of blocks, consolidated into one between calls
  std::vector<pob_arena_block> blocks;
  // pinned staging for small host<->device metadata
  char* pinned;
  size_t pinned_size;
  // profiling
  int prof_on;
  std::vector<pob_prof_rec> prof_pending;
  std::vector<cudaEvent_t> prof_pool;
  double prof_ms[POB_K_COUNT];
  int64_t prof_n[POB_K_COUNT];
  // counters
  int64_t counters[3];
  unsigned long long* d_counters;  // device side (2 x u64)
  cudaEvent_t t0, t1;              // pob_timer_*
};

static inline size_t pob_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Start a new API call: every block becomes free again; several blocks are merged into one.
int pob_arena_reset(pob_ctx* ctx);
// Bump-allocate 256-B aligned device scratch; grows by adding a block. nullptr on out-of-memory.
void* pob_arena_take(pob_ctx* ctx, size_t bytes);
template <typename T>
static inline int pob_take(pob_ctx* ctx, size_t count, T** out) {
  *out = (T*)pob_arena_take(ctx, count * sizeof(T));
  return *out ? POB_OK : POB_ENOMEM;
}

// profiling scope: records events around a kernel launch when enabled
struct pob_prof_scope {
  pob_ctx* ctx;
  int idx;
  pob_prof_scope(pob_ctx* c, int id);
  ~pob_prof_scope();
};

template <typename T>
static inline T* pob_ptr(void* base, size_t off) {
  return reinterpret_cast<T*>(reinterpret_cast<char*>(base) + off);
}

// Kernels see the same packed-batch descriptor as the ABI (pob_reads_t), with device pointers.
typedef pob_reads_t pob_reads;

__device__ __forceinline__ int pob_read_len(const int64_t* row_off, const int32_t* row_len, int r) {
  return row_len ? row_len[r] : (int)(row_off[r + 1] - row_off[r]);
}

// logical column k (A C G T blank) of a read -> physical column, for layout and rc view
__host__ __device__ static inline int pob_col(int k, int S, int layout, int rc) {
  int nb = S - 1;  // number of bases
  if (rc && k < nb) k = nb - 1 - k;
  if (layout == POB_BLANK_FIRST) return (k == nb) ? 0 : k + 1;
  return k;
}

static __device__ __forceinline__ double pob_ninf() { return __longlong_as_double(0xfff0000000000000LL); }
