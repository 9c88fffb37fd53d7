// extern "C" entry points: argument validation, host <-> device staging, kernel dispatch.
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

size_t reads_bytes(const pob_reads_t* r) {
  size_t rows = total_rows(r);
  return pob_align_up(rows * r->n_states * elt_size(r->dtype) + 64, 256) + pob_align_up(((size_t)r->n + 1) * 8, 256) +
         2 * pob_align_up((size_t)r->n * 4 + 4, 256) + 1024;
}

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

}  // namespace

extern "C" {

int pob_viterbi(pob_ctx* ctx, int where, const pob_reads_t* reads, int kind, uint8_t* out_seq, int32_t* out_s2s,
                int8_t* out_path, int32_t* out_len, int32_t* out_status) {
  if (!ctx) return POB_EINVAL;
  POB_TRY(check_reads(reads, 2, 9));
  if (reads->dtype != POB_F32 && reads->dtype != POB_F64) return POB_EINVAL;
  if (kind != POB_KIND_POREOVER && kind != POB_KIND_BONITO) return POB_EINVAL;
  const int n = reads->n;
  if (n == 0) return POB_OK;
  if (!out_seq || !out_len) return POB_EINVAL;
  POB_CUDA(cudaSetDevice(ctx->device));
  if (where == POB_DEVICE)
    return pob_viterbi_launch(ctx, *reads, kind, out_seq, out_s2s, out_path, out_len, out_status);
  const size_t rows = total_rows(reads);
  pob_arena_plan pl;
  pl.add(reads_bytes(reads)); pl.add(rows + 4); pl.add(rows * 4 + 4); pl.add(rows + 4); pl.add(n * 4); pl.add(n * 4);
  POB_TRY(pob_arena_reserve(ctx, pl.total));
  pob_reads_t d;
  POB_TRY(stage_reads(ctx, reads, &d));
  uint8_t* d_seq; int32_t* d_s2s; int8_t* d_path; int32_t *d_len, *d_st;
  POB_TRY(stage_out(ctx, out_seq, rows, &d_seq));
  POB_TRY(stage_out(ctx, out_s2s, rows, &d_s2s));
  POB_TRY(stage_out(ctx, out_path, rows, &d_path));
  POB_TRY(stage_out(ctx, out_len, (size_t)n, &d_len));
  POB_TRY(stage_out(ctx, out_status, (size_t)n, &d_st, true));
  POB_TRY(pob_viterbi_launch(ctx, d, kind, d_seq, d_s2s, d_path, d_len, d_st));
  POB_TRY(copy_back(ctx, out_seq, d_seq, rows));
  POB_TRY(copy_back(ctx, out_s2s, d_s2s, rows));
  POB_TRY(copy_back(ctx, out_path, d_path, rows));
  POB_TRY(copy_back(ctx, out_len, d_len, (size_t)n));
  POB_TRY(copy_back(ctx, out_status, d_st, (size_t)n));
  POB_CUDA(cudaStreamSynchronize(ctx->stream));
  return POB_OK;
}

// ---- not yet built in this revision: every symbol of the header exists and fails loudly ----
int pob_viterbi_flipflop(pob_ctx*, int, const pob_reads_t*, const double*, uint8_t*, int32_t*, int8_t*, int32_t*) {
  return POB_EUNSUPPORTED;
}
int pob_align_banded(pob_ctx*, int, const uint8_t*, const int64_t*, const uint8_t*, const int64_t*, int, int, int, int,
                     int, uint8_t*, uint8_t*, int32_t*, int32_t*) {
  return POB_EUNSUPPORTED;
}
int pob_build_envelope(pob_ctx*, int, const uint8_t*, const uint8_t*, const int64_t*, const int32_t*, const int32_t*,
                       const int64_t*, const int32_t*, const int32_t*, const int64_t*, const int32_t*, const int32_t*,
                       const int32_t*, const int64_t*, int, int, int32_t*) {
  return POB_EUNSUPPORTED;
}
int pob_beam_search(pob_ctx*, int, const pob_reads_t*, int, int, uint8_t*, int32_t*, double*, int32_t*) {
  return POB_EUNSUPPORTED;
}
int pob_beam_search_2d(pob_ctx*, int, const pob_reads_t*, const pob_reads_t*, const int32_t*, const int64_t*, int, int,
                       int, const int64_t*, uint8_t*, int32_t*, double*, int32_t*) {
  return POB_EUNSUPPORTED;
}
int pob_forward(pob_ctx*, int, const pob_reads_t*, const uint8_t*, const int64_t*, int, double*) {
  return POB_EUNSUPPORTED;
}
int pob_pair_decode(pob_ctx*, int, const pob_reads_t*, const pob_reads_t*, int, int, int, int, int, uint8_t*, int32_t*,
                    uint8_t*, int32_t*, uint8_t*, int32_t*, double*, int32_t*, int32_t*) {
  return POB_EUNSUPPORTED;
}

}  // extern "C"
