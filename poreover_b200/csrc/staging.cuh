// Host <-> device staging helpers shared by the API translation units.
#pragma once
#include <algorithm>
#include <numeric>
#include <vector>

#include "common.cuh"
#include "launch.cuh"

namespace {


size_t elt_size(int dtype) { return dtype == POB_F64 ? 8 : (dtype == POB_U8_TRACE ? 1 : 4); }

// bump-allocate + H2D copy of a host array (NULL stays NULL)
template <typename T>
int stage_in(pob_ctx* ctx, const T* host, size_t count, const T** dev, size_t pad_bytes = 0) {
  *dev = nullptr;
  if (!host) return POB_OK;
  T* d = (T*)pob_arena_take(ctx, count * sizeof(T) + pad_bytes);
  if (!d) return POB_ENOMEM;
  if (count) POB_CUDA(cudaMemcpyAsync(d, host, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  *dev = d;
  return POB_OK;
}
// device twin of a host output buffer (NULL host pointer -> no device buffer unless force)
template <typename T>
int stage_out(pob_ctx* ctx, const T* host_out, size_t count, T** dev, bool force = false) {
  *dev = nullptr;
  if (!host_out && !force) return POB_OK;
  T* d = (T*)pob_arena_take(ctx, count * sizeof(T));
  if (!d) return POB_ENOMEM;
  *dev = d;
  return POB_OK;
}
template <typename T>
int copy_back(pob_ctx* ctx, T* host, const T* dev, size_t count) {
  if (!host || !dev || !count) return POB_OK;
  POB_CUDA(cudaMemcpyAsync(host, dev, count * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
  return POB_OK;
}

int check_reads(const pob_reads_t* r, int min_states, int max_states) {
  if (!r || r->n < 0) return POB_EINVAL;
  if (r->n > 0 && (!r->data || !r->row_off)) return POB_EINVAL;
  if (r->n_states < min_states || r->n_states > max_states) return POB_EINVAL;
  if (r->layout != POB_BLANK_LAST && r->layout != POB_BLANK_FIRST) return POB_EINVAL;
  return POB_OK;
}

// total packed rows of a HOST descriptor
size_t total_rows(const pob_reads_t* r) { return r->n > 0 ? (size_t)r->row_off[r->n] : 0; }

// copy a host descriptor's arrays into the arena, produce the device descriptor
int stage_reads(pob_ctx* ctx, const pob_reads_t* h, pob_reads_t* d) {
  *d = *h;
  size_t rows = total_rows(h);
  const char* data;
  POB_TRY(stage_in(ctx, (const char*)h->data, rows * h->n_states * elt_size(h->dtype), &data, 64));
  d->data = data;
  POB_TRY(stage_in(ctx, h->row_off, (size_t)h->n + 1, &d->row_off));
  POB_TRY(stage_in(ctx, h->row_len, (size_t)h->n, &d->row_len));
  POB_TRY(stage_in(ctx, h->rc, (size_t)h->n, &d->rc));
  return POB_OK;
}

// fetch an int64 offsets array (n+1) to the host regardless of where it lives
int fetch_i64(pob_ctx* ctx, int where, const int64_t* p, size_t count, std::vector<int64_t>& out) {
  out.resize(count);
  if (where == POB_HOST) {
    memcpy(out.data(), p, count * 8);
  } else {
    POB_CUDA(cudaMemcpyAsync(out.data(), p, count * 8, cudaMemcpyDeviceToHost, ctx->stream));
    POB_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return POB_OK;
}
int fetch_i32(pob_ctx* ctx, int where, const int32_t* p, size_t count, std::vector<int32_t>& out) {
  out.resize(count);
  if (where == POB_HOST) {
    memcpy(out.data(), p, count * 4);
  } else {
    POB_CUDA(cudaMemcpyAsync(out.data(), p, count * 4, cudaMemcpyDeviceToHost, ctx->stream));
    POB_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return POB_OK;
}
template <typename T>
int upload(pob_ctx* ctx, const std::vector<T>& h, const T** dev) {
  return stage_in(ctx, h.data(), h.size(), dev);
}


}  // namespace
