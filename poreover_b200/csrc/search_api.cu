// extern "C" entry points of the beam searches and the fused pair-decode pipeline.
#include "staging.cuh"

#include <chrono>
#include <stdlib.h>

namespace {

// host-side copy of the row geometry of a reads descriptor
struct Geometry {
  std::vector<int64_t> off;
  std::vector<int32_t> len;
  int maxlen = 0;
};

int fetch_geometry(pob_ctx* ctx, int where, const pob_reads_t* r, Geometry& g) {
  POB_TRY(fetch_i64(ctx, where, r->row_off, (size_t)r->n + 1, g.off));
  if (r->row_len) {
    POB_TRY(fetch_i32(ctx, where, r->row_len, (size_t)r->n, g.len));
  } else {
    g.len.resize(r->n);
    for (int i = 0; i < r->n; ++i) g.len[i] = (int32_t)(g.off[i + 1] - g.off[i]);
  }
  g.maxlen = 0;
  for (int i = 0; i < r->n; ++i) g.maxlen = std::max(g.maxlen, (int)g.len[i]);
  return POB_OK;
}

__global__ void identity_skip_kernel(const int32_t* __restrict__ alen, const int32_t* __restrict__ matches, int n,
                                     int32_t* __restrict__ skip, int32_t* __restrict__ status) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n || skip[p]) return;
  // sequence_identity = matches / columns < 0.5  (pair_decode.py:391-395); exact in integers
  if (alen[p] <= 0 || 2 * matches[p] < alen[p]) {
    skip[p] = 1;
    status[p] |= POB_ST_SKIPPED_IDENTITY;
  }
}

// longest-first processing order over the items that are actually searched
void make_order(const std::vector<int64_t>& cost, const std::vector<int32_t>* skip, std::vector<int32_t>& order) {
  order.clear();
  for (int i = 0; i < (int)cost.size(); ++i)
    if (!skip || !(*skip)[i]) order.push_back(i);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
}

// Shared body of pob_beam_search / pob_beam_search_2d once everything is on the device.
// d_env may be NULL.  g1/g2: host geometry.  Scratch (trace, top, envt) comes from the arena.
int run_search(pob_ctx* ctx, const pob_reads_t& d1, const pob_reads_t* d2, const Geometry& g1, const Geometry* g2,
               const int32_t* d_env, const int64_t* d_env_off, const int32_t* d_skip,
               const std::vector<int32_t>* h_skip, int W, int model, int mode, const int64_t* d_out_off,
               uint8_t* d_seq, int32_t* d_len, double* d_score, int32_t* d_status) {
  const int n = d1.n;
  std::vector<int64_t> trace_off(n + 1), cost(n);
  trace_off[0] = 0;
  for (int i = 0; i < n; ++i) {
    const int64_t U = g1.len[i], V = g2 ? g2->len[i] : 0;
    const bool sk = h_skip && (*h_skip)[i];
    trace_off[i + 1] = trace_off[i] + (sk ? 8 : (int64_t)W * (U + 2) + 8);
    cost[i] = U + V;
  }
  std::vector<int32_t> order;
  make_order(cost, h_skip, order);
  const int64_t* d_trace_off;
  const int32_t* d_order;
  POB_TRY(upload(ctx, trace_off, &d_trace_off));
  if (order.empty()) order.push_back(0);
  POB_TRY(upload(ctx, order, &d_order));
  uint32_t* d_trace;
  int32_t* d_top;
  POB_TRY(pob_take(ctx, (size_t)trace_off[n] + 1, &d_trace));
  POB_TRY(pob_take(ctx, (size_t)n * 4 + 4, &d_top));
  POB_CUDA(cudaMemsetAsync(d_top, 0, ((size_t)n * 4 + 4) * 4, ctx->stream));
  int span0 = 1, span1 = 1;
  const int32_t* d_envt = nullptr;
  const int64_t* d_envt_off = nullptr;
  if (mode != POB_MODE_1D && d_env) {
    // transpose (needed by row_col) + widest row / column band of the batch
    std::vector<int64_t> envt_off(n + 1);
    envt_off[0] = 0;
    for (int i = 0; i < n; ++i) envt_off[i + 1] = envt_off[i] + g2->len[i];
    POB_TRY(upload(ctx, envt_off, &d_envt_off));
    int32_t *envt, *span;
    const int32_t *dU, *dV;
    POB_TRY(pob_take(ctx, (size_t)envt_off[n] * 2 + 2, &envt));
    POB_TRY(pob_take(ctx, (size_t)n * 2 + 2, &span));
    POB_CUDA(cudaMemsetAsync(span, 0, ((size_t)n * 2 + 2) * 4, ctx->stream));
    POB_TRY(upload(ctx, g1.len, &dU));
    POB_TRY(upload(ctx, g2->len, &dV));
    POB_TRY(pob_envelope_transpose_launch(ctx, d_env, d_env_off, dU, dV, d_envt_off, d_skip, n, envt, span));
    std::vector<int32_t> hspan((size_t)n * 2);
    POB_CUDA(cudaMemcpyAsync(hspan.data(), span, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    POB_CUDA(cudaStreamSynchronize(ctx->stream));
    int wmax = 1, cmax = 1;
    for (int i = 0; i < n; ++i) {
      if (h_skip && (*h_skip)[i]) continue;
      wmax = std::max(wmax, hspan[2 * i]);
      cmax = std::max(cmax, hspan[2 * i + 1]);
    }
    d_envt = envt;
    if (mode == POB_MODE_ROWCOL) { span0 = cmax; span1 = wmax; }
    else { span0 = 1; span1 = wmax; }
  } else if (mode == POB_MODE_ROW) {
    span0 = 1;
    span1 = g2->maxlen;  // no envelope: every row sweeps all of read 2
  }
  const int nsearch = (int)((h_skip) ? std::count(h_skip->begin(), h_skip->end(), 0) : n);
  {
    POB_TRY(pob_beam_launch(ctx, d1, d2, d_env, d_env_off, d_envt, d_envt_off, d_order, d_skip, nsearch, n, W, model,
                            mode, span0, span1, g1.maxlen, g2 ? g2->maxlen : 0, d_trace_off, d_trace, d_top,
                            d_out_off, d_seq, d_len, d_score, d_status));
  }
  return POB_OK;
}

int search_entry(pob_ctx* ctx, int where, const pob_reads_t* r1, const pob_reads_t* r2, const int32_t* env,
                 const int64_t* env_off, int W, int model, int mode, const int64_t* out_off, uint8_t* out_seq,
                 int32_t* out_len, double* out_score, int32_t* out_status) {
  if (!ctx) return POB_EINVAL;
  POB_TRY(check_reads(r1, 2, 5));
  if (r2) {
    POB_TRY(check_reads(r2, 2, 5));
    if (r2->n != r1->n || r2->dtype != r1->dtype) return POB_EINVAL;
  }
  if (r1->dtype != POB_F32 && r1->dtype != POB_F64) return POB_EINVAL;
  if (model != POB_MODEL_CTC && model != POB_MODEL_CTC_MERGE_REPEATS) return POB_EINVAL;
  if (W < 1) return POB_EINVAL;
  const int n = r1->n;
  if (n == 0) return POB_OK;
  if (!out_seq || !out_len || (mode != POB_MODE_1D && !out_off)) return POB_EINVAL;
  if (mode == POB_MODE_ROWCOL && (!env || !env_off)) return POB_EINVAL;
  if (env && !env_off) return POB_EINVAL;
  POB_CUDA(cudaSetDevice(ctx->device));
  POB_TRY(pob_arena_reset(ctx));
  Geometry g1, g2;
  POB_TRY(fetch_geometry(ctx, where, r1, g1));
  if (r2) POB_TRY(fetch_geometry(ctx, where, r2, g2));
  pob_reads_t d1 = *r1, d2;
  if (r2) d2 = *r2;
  const int32_t* d_env = env;
  const int64_t *d_env_off = env_off, *d_out_off = out_off ? out_off : r1->row_off;
  uint8_t* d_seq = out_seq;
  int32_t *d_len = out_len, *d_status = out_status;
  double* d_score = out_score;
  std::vector<int64_t> h_out_off;
  size_t out_bytes = 0;
  if (where == POB_HOST) {
    POB_TRY(stage_reads(ctx, r1, &d1));
    if (r2) POB_TRY(stage_reads(ctx, r2, &d2));
    if (env) {
      POB_TRY(stage_in(ctx, env, (size_t)env_off[n] * 2, &d_env));
      POB_TRY(stage_in(ctx, env_off, (size_t)n + 1, &d_env_off));
    }
    const int64_t* oo = out_off ? out_off : r1->row_off;
    out_bytes = (size_t)oo[n];
    POB_TRY(stage_in(ctx, oo, (size_t)n + 1, &d_out_off));
    POB_TRY(stage_out(ctx, out_seq, out_bytes + 4, &d_seq));
    POB_TRY(stage_out(ctx, out_len, (size_t)n, &d_len));
    POB_TRY(stage_out(ctx, out_score, (size_t)n, &d_score, true));
    POB_TRY(stage_out(ctx, out_status, (size_t)n, &d_status, true));
  } else {
    if (!d_score) POB_TRY(pob_take(ctx, (size_t)n, &d_score));
    if (!d_status) POB_TRY(pob_take(ctx, (size_t)n, &d_status));
  }
  POB_CUDA(cudaMemsetAsync(d_status, 0, (size_t)n * 4, ctx->stream));
  POB_TRY(run_search(ctx, d1, r2 ? &d2 : nullptr, g1, r2 ? &g2 : nullptr, d_env, d_env_off, nullptr, nullptr, W,
                     model, mode, d_out_off, d_seq, d_len, d_score, d_status));
  if (where == POB_HOST) {
    POB_TRY(copy_back(ctx, out_seq, d_seq, out_bytes));
    POB_TRY(copy_back(ctx, out_len, d_len, (size_t)n));
    POB_TRY(copy_back(ctx, out_score, d_score, (size_t)n));
    POB_TRY(copy_back(ctx, out_status, d_status, (size_t)n));
  }
  POB_CUDA(cudaStreamSynchronize(ctx->stream));
  return POB_OK;
}

}  // namespace

extern "C" {

int pob_beam_search(pob_ctx* ctx, int where, const pob_reads_t* reads, int beam_width, int model, uint8_t* out_seq,
                    int32_t* out_len, double* out_score, int32_t* out_status) {
  return search_entry(ctx, where, reads, nullptr, nullptr, nullptr, beam_width, model, POB_MODE_1D, nullptr, out_seq,
                      out_len, out_score, out_status);
}

int pob_beam_search_2d(pob_ctx* ctx, int where, const pob_reads_t* reads1, const pob_reads_t* reads2,
                       const int32_t* env, const int64_t* env_off, int beam_width, int model, int method,
                       const int64_t* out_off, uint8_t* out_seq, int32_t* out_len, double* out_score,
                       int32_t* out_status) {
  if (!reads2) return POB_EINVAL;
  if (method != POB_METHOD_ROW && method != POB_METHOD_ROW_COL) return POB_EINVAL;
  return search_entry(ctx, where, reads1, reads2, env, env_off, beam_width, model,
                      method == POB_METHOD_ROW ? POB_MODE_ROW : POB_MODE_ROWCOL, out_off, out_seq, out_len, out_score,
                      out_status);
}

int pob_pair_decode(pob_ctx* ctx, int where, const pob_reads_t* reads1, const pob_reads_t* reads2, int kind,
                    int beam_width, int padding, int band_width, int method, uint8_t* out_seq1, int32_t* out_len1,
                    uint8_t* out_seq2, int32_t* out_len2, uint8_t* out_cons, int32_t* out_cons_len,
                    double* out_score, int32_t* out_stats, int32_t* out_status) {
  if (!ctx) return POB_EINVAL;
  POB_TRY(check_reads(reads1, 5, 5));
  POB_TRY(check_reads(reads2, 5, 5));
  if (reads1->n != reads2->n || reads1->dtype != reads2->dtype) return POB_EINVAL;
  if (reads1->dtype != POB_F32 && reads1->dtype != POB_F64) return POB_EINVAL;
  if (kind != POB_KIND_POREOVER && kind != POB_KIND_BONITO) return POB_EINVAL;
  if (method != POB_METHOD_ROW && method != POB_METHOD_ROW_COL) return POB_EINVAL;
  if (beam_width < 1 || band_width < 0 || padding < 0) return POB_EINVAL;
  const int n = reads1->n;
  if (n == 0) return POB_OK;
  if (!out_seq1 || !out_len1 || !out_seq2 || !out_len2 || !out_cons || !out_cons_len || !out_status)
    return POB_EINVAL;
  const int model = kind == POB_KIND_BONITO ? POB_MODEL_CTC_MERGE_REPEATS : POB_MODEL_CTC;
  // POB_DEBUG_TIMING=1: wall-clock milliseconds of the call's host-visible stages on stderr
  static const bool timing = getenv("POB_DEBUG_TIMING") != nullptr;
  auto tnow = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double tmark = timing ? tnow() : 0.0;
  auto stage_done = [&](const char* what) {
    if (!timing) return;
    const double t = tnow();
    fprintf(stderr, "[pob timing] %-22s %8.3f ms\n", what, t - tmark);
    tmark = t;
  };
  POB_CUDA(cudaSetDevice(ctx->device));
  POB_TRY(pob_arena_reset(ctx));
  stage_done("arena reset");
  Geometry g1, g2;
  POB_TRY(fetch_geometry(ctx, where, reads1, g1));
  POB_TRY(fetch_geometry(ctx, where, reads2, g2));
  // the batch may be a slice of a larger packed batch: its rows are [off[0], off[n]) and every packed output /
  // scratch array below is rebased by off[0], so that the absolute row offsets address it
  const size_t base1 = (size_t)g1.off[0], base2 = (size_t)g2.off[0];
  const size_t rows1 = (size_t)g1.off[n] - base1, rows2 = (size_t)g2.off[n] - base2;
  pob_reads_t d1 = *reads1, d2 = *reads2;
  uint8_t *d_seq1 = out_seq1, *d_seq2 = out_seq2, *d_cons = out_cons;
  int32_t *d_len1 = out_len1, *d_len2 = out_len2, *d_clen = out_cons_len, *d_stats = out_stats, *d_status = out_status;
  double* d_score = out_score;
  if (where == POB_HOST) {
    POB_TRY(stage_reads(ctx, reads1, &d1));
    POB_TRY(stage_reads(ctx, reads2, &d2));
    POB_TRY(stage_out(ctx, out_seq1, rows1 + 4, &d_seq1));
    POB_TRY(stage_out(ctx, out_seq2, rows2 + 4, &d_seq2));
    POB_TRY(stage_out(ctx, out_cons, rows1 + rows2 + 4, &d_cons));
    d_seq1 -= base1; d_seq2 -= base2; d_cons -= base1 + base2;
    POB_TRY(stage_out(ctx, out_len1, (size_t)n, &d_len1));
    POB_TRY(stage_out(ctx, out_len2, (size_t)n, &d_len2));
    POB_TRY(stage_out(ctx, out_cons_len, (size_t)n, &d_clen));
    POB_TRY(stage_out(ctx, out_status, (size_t)n, &d_status));
  }
  stage_done("geometry + staging");
  if (where == POB_HOST || !d_score) POB_TRY(pob_take(ctx, (size_t)n, &d_score));
  if (where == POB_HOST || !d_stats) POB_TRY(pob_take(ctx, (size_t)n * 4, &d_stats));
  // ---- stage 1: best-path decode + base->timestep mapping of both reads (transducer.py, pair_decode.py:361-382)
  int32_t *d_s2s1, *d_s2s2, *d_st1, *d_st2;
  POB_TRY(pob_take(ctx, rows1 + 4, &d_s2s1));
  POB_TRY(pob_take(ctx, rows2 + 4, &d_s2s2));
  d_s2s1 -= base1; d_s2s2 -= base2;
  POB_TRY(pob_take(ctx, (size_t)n, &d_st1));
  POB_TRY(pob_take(ctx, (size_t)n, &d_st2));
  POB_TRY(pob_viterbi_launch(ctx, d1, kind, d_seq1, d_s2s1, nullptr, d_len1, d_st1));
  POB_TRY(pob_viterbi_launch(ctx, d2, kind, d_seq2, d_s2s2, nullptr, d_len2, d_st2));
  std::vector<int32_t> len1(n), len2(n), st1(n), st2(n);
  POB_CUDA(cudaMemcpyAsync(len1.data(), d_len1, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  POB_CUDA(cudaMemcpyAsync(len2.data(), d_len2, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  POB_CUDA(cudaMemcpyAsync(st1.data(), d_st1, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  POB_CUDA(cudaMemcpyAsync(st2.data(), d_st2, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  POB_CUDA(cudaStreamSynchronize(ctx->stream));
  stage_done("viterbi + read-back");
  // ---- host: who is skipped before alignment, scratch geometry of the rest
  std::vector<int32_t> skip(n, 0), status(n, 0), one_empty(n, 0);
  const int SZ = pob_nw_slots(band_width);
  std::vector<int64_t> m_off(n + 1), rb_off(n + 1), aln_off(n + 1), env_off(n + 1), cons_off(n + 1);
  m_off[0] = rb_off[0] = aln_off[0] = env_off[0] = 0;
  for (int p = 0; p < n; ++p) {
    int st = (st1[p] | st2[p]) & (POB_ST_MAPPING_WRAP | POB_ST_EMPTY);
    const int64_t l1 = len1[p], l2 = len2[p];
    const int64_t dl = l1 > l2 ? l1 - l2 : l2 - l1;
    if (dl > 1000) st |= POB_ST_SKIPPED_LENGTH;  // pair_decode.py:372-375
    // An empty basecall: the banded aligner's row loop never runs (align.pyx:120) and its traceback only drains the
    // other sequence, so the alignment is all gaps, the identity 0 and the pair is skipped by identity
    // (pair_decode.py:391-398).  Two empty basecalls divide 0 by 0: that pool task dies (POB_ST_EMPTY).
    if (l1 == 0 && l2 == 0) st |= POB_ST_EMPTY;
    else if ((l1 == 0 || l2 == 0) && !(st & POB_ST_SKIPPED_LENGTH)) { st |= POB_ST_SKIPPED_IDENTITY; one_empty[p] = 1; }
    status[p] = st;
    skip[p] = st != 0;
    const int64_t D = skip[p] ? 0 : l1 + l2 - 1;
    m_off[p + 1] = m_off[p] + D * SZ;
    rb_off[p + 1] = rb_off[p] + (skip[p] ? 0 : 2 * l1);
    aln_off[p + 1] = aln_off[p] + (skip[p] ? 0 : l1 + l2 + 8);
    env_off[p + 1] = env_off[p] + g1.len[p];
    cons_off[p] = g1.off[p] + g2.off[p];
  }
  cons_off[n] = g1.off[n] + g2.off[n];
  const int64_t *d_rboff, *d_alnoff, *d_envoff, *d_consoff;
  const int32_t *d_skip_c, *d_status_c, *dU, *dV;
  POB_TRY(upload(ctx, rb_off, &d_rboff));
  POB_TRY(upload(ctx, aln_off, &d_alnoff));
  POB_TRY(upload(ctx, env_off, &d_envoff));
  POB_TRY(upload(ctx, cons_off, &d_consoff));
  POB_TRY(upload(ctx, skip, &d_skip_c));
  POB_TRY(upload(ctx, status, &d_status_c));
  POB_TRY(upload(ctx, g1.len, &dU));
  POB_TRY(upload(ctx, g2.len, &dV));
  int32_t* d_skip = const_cast<int32_t*>(d_skip_c);
  POB_CUDA(cudaMemcpyAsync(d_status, d_status_c, (size_t)n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
  // ---- stage 2: banded alignment of the two basecalls (align.pyx:100-178)
  int32_t *rowband, *d_alen, *d_matches;
  uint8_t *d_a1, *d_a2;
  POB_TRY(pob_take(ctx, (size_t)rb_off[n] + 1, &rowband));
  POB_TRY(pob_take(ctx, (size_t)aln_off[n] + 1, &d_a1));
  POB_TRY(pob_take(ctx, (size_t)aln_off[n] + 1, &d_a2));
  POB_TRY(pob_take(ctx, (size_t)n, &d_alen));
  POB_TRY(pob_take(ctx, (size_t)n, &d_matches));
  {
    std::vector<int64_t> cells(n);
    for (int p = 0; p < n; ++p) cells[p] = m_off[p + 1] - m_off[p];
    POB_TRY(nw_run_chunked(ctx, d_seq1, d1.row_off, d_len1, d_seq2, d2.row_off, d_len2, d_skip, n, band_width, 2, -1,
                           -1, SZ, cells, d_rboff, rowband, d_alnoff, d_a1, d_a2, d_alen, d_matches));
  }
  identity_skip_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(d_alen, d_matches, n, d_skip, d_status);
  POB_CUDA(cudaGetLastError());
  // ---- stage 3: alignment columns -> envelope (envelope.py:26-87)
  int32_t* d_env;
  POB_TRY(pob_take(ctx, (size_t)env_off[n] * 2 + 2, &d_env));
  POB_TRY(pob_envelope_launch(ctx, d_a1, d_a2, d_alnoff, d_alen, d_s2s1, d1.row_off, d_len1, d_s2s2, d2.row_off,
                              d_len2, dU, dV, d_envoff, d_skip, n, padding, d_env));
  // final skip decisions are needed on the host to order and size the search
  std::vector<int32_t> alen(n), matches(n);
  POB_CUDA(cudaMemcpyAsync(skip.data(), d_skip, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  POB_CUDA(cudaMemcpyAsync(alen.data(), d_alen, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  POB_CUDA(cudaMemcpyAsync(matches.data(), d_matches, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  POB_CUDA(cudaStreamSynchronize(ctx->stream));
  stage_done("NW + envelope + read-back");
  // ---- stage 4: joint beam search inside the envelope (BeamSearch.h:262-397 / :110-172)
  POB_CUDA(cudaMemsetAsync(d_clen, 0, (size_t)n * 4, ctx->stream));
  POB_CUDA(cudaMemsetAsync(d_score, 0, (size_t)n * 8, ctx->stream));
  POB_TRY(run_search(ctx, d1, &d2, g1, &g2, d_env, d_envoff, d_skip, &skip, beam_width, model,
                     method == POB_METHOD_ROW ? POB_MODE_ROW : POB_MODE_ROWCOL, d_consoff, d_cons, d_clen, d_score,
                     d_status));
  stage_done("search enqueued");
  // stats: len1, len2, matches, columns
  std::vector<int32_t> stats((size_t)n * 4);
  for (int p = 0; p < n; ++p) {
    stats[4 * p] = len1[p]; stats[4 * p + 1] = len2[p];
    const bool aligned = !(status[p] != 0);
    stats[4 * p + 2] = aligned ? matches[p] : 0;
    stats[4 * p + 3] = aligned ? alen[p] : (one_empty[p] ? len1[p] + len2[p] : 0);  // all-gap alignment: 0 matches of l columns
  }
  if (where == POB_HOST) {
    if (out_stats) memcpy(out_stats, stats.data(), stats.size() * 4);
    POB_TRY(copy_back(ctx, out_seq1 + base1, d_seq1 + base1, rows1));
    POB_TRY(copy_back(ctx, out_seq2 + base2, d_seq2 + base2, rows2));
    POB_TRY(copy_back(ctx, out_cons + base1 + base2, d_cons + base1 + base2, rows1 + rows2));
    POB_TRY(copy_back(ctx, out_len1, d_len1, (size_t)n));
    POB_TRY(copy_back(ctx, out_len2, d_len2, (size_t)n));
    POB_TRY(copy_back(ctx, out_cons_len, d_clen, (size_t)n));
    POB_TRY(copy_back(ctx, out_score, d_score, (size_t)n));
    POB_TRY(copy_back(ctx, out_status, d_status, (size_t)n));
  } else if (out_stats) {
    POB_CUDA(cudaMemcpyAsync(out_stats, stats.data(), stats.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  }
  POB_CUDA(cudaStreamSynchronize(ctx->stream));
  stage_done("copy back + sync");
  return POB_OK;
}

}  // extern "C"
