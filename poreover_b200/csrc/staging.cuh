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

// first packed row of a HOST descriptor: a batch may be a slice [lo, hi) of a larger packed batch (row_off + lo, n =
// hi - lo, same data pointer), so its rows start at row_off[0], not at 0
size_t first_row(const pob_reads_t* r) { return r->n > 0 ? (size_t)r->row_off[0] : 0; }

// copy a host descriptor's arrays into the arena, produce the device descriptor.  Only the rows [row_off[0],
// row_off[n]) travel; the device data pointer is rebased so that the (absolute) row offsets keep addressing them.
int stage_reads(pob_ctx* ctx, const pob_reads_t* h, pob_reads_t* d) {
  *d = *h;
  const size_t r0 = first_row(h), rows = total_rows(h) - r0;
  const size_t rowb = (size_t)h->n_states * elt_size(h->dtype);
  const char* data;
  // 256-byte aligned staging keeps the 16-byte alignment of the reads whatever r0 is (r0 * rowb is a multiple of 16
  // for every read the packer aligned)
  POB_TRY(stage_in(ctx, (const char*)h->data + r0 * rowb, rows * rowb, &data, 64));
  d->data = data - r0 * rowb;
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

// Banded NW + traceback over n pairs with at most NW_CELL_BUDGET score cells (int32) resident at a time: the
// anti-diagonal-major matrix of a T~5000 pair is 16 MB, so a 10k-pair batch is aligned in chunks that reuse one
// scratch matrix (same stream, so a chunk's traceback finishes before the next fill overwrites it).
// cells[p] = matrix size of pair p (0 = skipped); all other per-pair arrays are device pointers indexed by pair.
constexpr int64_t NW_CELL_BUDGET = (int64_t)4 << 30;  // 16 GB = 1024 pairs of T ~ 5000 per fill + traceback launch (32 GB
                                                      // made the scratch arena of every context twice as large for nothing)
int nw_run_chunked(pob_ctx* ctx, const uint8_t* s1, const int64_t* off1, const int32_t* len1, const uint8_t* s2,
                   const int64_t* off2, const int32_t* len2, const int32_t* skip, int n, int band, int match,
                   int mismatch, int gap, int SZ, const std::vector<int64_t>& cells, const int64_t* d_rboff,
                   int32_t* rowband, const int64_t* d_alnoff, uint8_t* a1, uint8_t* a2, int32_t* alen,
                   int32_t* matches) {
  std::vector<int64_t> m_off((size_t)n + 1);
  std::vector<int> starts(1, 0);
  int64_t cur = 0, widest = 0;
  for (int p = 0; p < n; ++p) {
    if (cur > 0 && cur + cells[p] > NW_CELL_BUDGET) { starts.push_back(p); cur = 0; }
    m_off[p] = cur;
    cur += cells[p];
    widest = std::max(widest, cur);
  }
  m_off[n] = cur;
  starts.push_back(n);
  const int64_t* d_moff;
  POB_TRY(upload(ctx, m_off, &d_moff));
  int32_t* M;
  POB_TRY(pob_take(ctx, (size_t)widest + 1, &M));
  for (size_t c = 0; c + 1 < starts.size(); ++c) {
    const int p0 = starts[c], cnt = starts[c + 1] - p0;
    POB_TRY(pob_nw_launch(ctx, s1, off1 + p0, len1 ? len1 + p0 : nullptr, s2, off2 + p0, len2 ? len2 + p0 : nullptr,
                          skip ? skip + p0 : nullptr, cnt, band, match, mismatch, gap, SZ, d_moff + p0, M,
                          d_rboff + p0, rowband, d_alnoff + p0, a1, a2, alen ? alen + p0 : nullptr,
                          matches ? matches + p0 : nullptr));
  }
  return POB_OK;
}

}  // namespace
