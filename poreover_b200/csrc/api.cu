// extern "C" entry points: argument validation, host <-> device staging, kernel dispatch.
#include "staging.cuh"

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
  POB_TRY(pob_arena_reset(ctx));
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

// ---- flip-flop Viterbi (transducer.py:35-59): uint8 traces through the host-computed table, or float64 ----
int pob_viterbi_flipflop(pob_ctx* ctx, int where, const pob_reads_t* reads, const double* lut, uint8_t* out_seq,
                         int32_t* out_s2s, int8_t* out_path, int32_t* out_len) {
  if (!ctx) return POB_EINVAL;
  if (reads && reads->n == 0) return POB_OK;  // an empty batch has no geometry to validate
  POB_TRY(check_reads(reads, 8, 8));
  if (reads->dtype != POB_F64 && reads->dtype != POB_U8_TRACE) return POB_EINVAL;
  if (reads->dtype == POB_U8_TRACE && !lut) return POB_EINVAL;
  const int n = reads->n;
  if (n == 0) return POB_OK;
  if (!out_seq || !out_len) return POB_EINVAL;
  POB_CUDA(cudaSetDevice(ctx->device));
  POB_TRY(pob_arena_reset(ctx));
  std::vector<int64_t> off;
  POB_TRY(fetch_i64(ctx, where, reads->row_off, (size_t)n + 1, off));
  const size_t rows = (size_t)off[n];
  pob_reads_t d = *reads;
  const double* d_lut = lut;
  uint8_t* d_seq = out_seq; int32_t* d_s2s = out_s2s; int8_t* d_path = out_path; int32_t* d_len = out_len;
  if (where == POB_HOST) {
    POB_TRY(stage_reads(ctx, reads, &d));
    POB_TRY(stage_in(ctx, lut, (size_t)(lut ? 256 : 0), &d_lut));
    POB_TRY(stage_out(ctx, out_seq, rows + 4, &d_seq));
    POB_TRY(stage_out(ctx, out_s2s, rows + 4, &d_s2s));
    POB_TRY(stage_out(ctx, out_path, rows + 4, &d_path));
    POB_TRY(stage_out(ctx, out_len, (size_t)n, &d_len));
  }
  uint32_t* bp;
  POB_TRY(pob_take(ctx, rows + 4, &bp));
  if (!d_path) POB_TRY(pob_take(ctx, rows + 4, &d_path));
  POB_TRY(pob_flipflop_launch(ctx, d, d_lut, bp, d_path, d_seq, d_s2s, d_len));
  if (where == POB_HOST) {
    POB_TRY(copy_back(ctx, out_seq, d_seq, rows));
    POB_TRY(copy_back(ctx, out_s2s, d_s2s, rows));
    POB_TRY(copy_back(ctx, out_path, d_path, rows));
    POB_TRY(copy_back(ctx, out_len, d_len, (size_t)n));
  }
  POB_CUDA(cudaStreamSynchronize(ctx->stream));
  return POB_OK;
}
int pob_align_banded(pob_ctx* ctx, int where, const uint8_t* seq1, const int64_t* off1, const uint8_t* seq2,
                     const int64_t* off2, int n, int band, int match, int mismatch, int gap, uint8_t* out_a1,
                     uint8_t* out_a2, int32_t* out_alen, int32_t* out_matches) {
  if (!ctx || n < 0 || band < 0) return POB_EINVAL;
  if (n == 0) return POB_OK;
  if (!seq1 || !seq2 || !off1 || !off2 || !out_a1 || !out_a2 || !out_alen) return POB_EINVAL;
  POB_CUDA(cudaSetDevice(ctx->device));
  POB_TRY(pob_arena_reset(ctx));
  std::vector<int64_t> o1, o2;
  POB_TRY(fetch_i64(ctx, where, off1, (size_t)n + 1, o1));
  POB_TRY(fetch_i64(ctx, where, off2, (size_t)n + 1, o2));
  const int SZ = pob_nw_slots(band);
  std::vector<int64_t> m_off(n + 1), rb_off(n + 1), aln_off(n + 1);
  m_off[0] = rb_off[0] = 0;
  for (int p = 0; p < n; ++p) {
    int64_t l1 = o1[p + 1] - o1[p], l2 = o2[p + 1] - o2[p];
    int64_t D = (l1 > 0 && l2 > 0) ? l1 + l2 - 1 : 0;
    m_off[p + 1] = m_off[p] + D * SZ;
    rb_off[p + 1] = rb_off[p] + 2 * l1;
    aln_off[p] = o1[p] + o2[p] + 8 * (int64_t)p;
  }
  aln_off[n] = o1[n] + o2[n] + 8 * (int64_t)n;
  const uint8_t *d_s1 = seq1, *d_s2 = seq2;
  const int64_t *d_o1 = off1, *d_o2 = off2;
  uint8_t *d_a1 = out_a1, *d_a2 = out_a2;
  int32_t *d_alen = out_alen, *d_match = out_matches;
  if (where == POB_HOST) {
    POB_TRY(stage_in(ctx, seq1, (size_t)o1[n], &d_s1));
    POB_TRY(stage_in(ctx, seq2, (size_t)o2[n], &d_s2));
    POB_TRY(stage_in(ctx, off1, (size_t)n + 1, &d_o1));
    POB_TRY(stage_in(ctx, off2, (size_t)n + 1, &d_o2));
    POB_TRY(stage_out(ctx, out_a1, (size_t)aln_off[n], &d_a1));
    POB_TRY(stage_out(ctx, out_a2, (size_t)aln_off[n], &d_a2));
    POB_TRY(stage_out(ctx, out_alen, (size_t)n, &d_alen));
    POB_TRY(stage_out(ctx, out_matches, (size_t)n, &d_match));
  }
  const int64_t *d_rboff, *d_alnoff;
  POB_TRY(upload(ctx, rb_off, &d_rboff));
  POB_TRY(upload(ctx, aln_off, &d_alnoff));
  int32_t* rowband;
  POB_TRY(pob_take(ctx, (size_t)rb_off[n] + 1, &rowband));
  std::vector<int64_t> cells(n);
  for (int p = 0; p < n; ++p) cells[p] = m_off[p + 1] - m_off[p];
  POB_TRY(nw_run_chunked(ctx, d_s1, d_o1, nullptr, d_s2, d_o2, nullptr, nullptr, n, band, match, mismatch, gap, SZ, cells,
                         d_rboff, rowband, d_alnoff, d_a1, d_a2, d_alen, d_match));
  if (where == POB_HOST) {
    POB_TRY(copy_back(ctx, out_a1, d_a1, (size_t)aln_off[n]));
    POB_TRY(copy_back(ctx, out_a2, d_a2, (size_t)aln_off[n]));
    POB_TRY(copy_back(ctx, out_alen, d_alen, (size_t)n));
    POB_TRY(copy_back(ctx, out_matches, d_match, (size_t)n));
  }
  // scratch sizes came from host-side offsets, so the stream is drained before the arena can be reused
  POB_CUDA(cudaStreamSynchronize(ctx->stream));
  return POB_OK;
}

int pob_build_envelope(pob_ctx* ctx, int where, const uint8_t* a1, const uint8_t* a2, const int64_t* aln_off,
                       const int32_t* alen, const int32_t* s2s1, const int64_t* soff1, const int32_t* slen1,
                       const int32_t* s2s2, const int64_t* soff2, const int32_t* slen2, const int32_t* U,
                       const int32_t* V, const int64_t* env_off, int n, int padding, int32_t* out_env) {
  if (!ctx || n < 0) return POB_EINVAL;
  if (n == 0) return POB_OK;
  if (!a1 || !a2 || !aln_off || !alen || !s2s1 || !soff1 || !slen1 || !s2s2 || !soff2 || !slen2 || !U || !V ||
      !env_off || !out_env)
    return POB_EINVAL;
  POB_CUDA(cudaSetDevice(ctx->device));
  if (where == POB_DEVICE)
    return pob_envelope_launch(ctx, a1, a2, aln_off, alen, s2s1, soff1, slen1, s2s2, soff2, slen2, U, V, env_off,
                               nullptr, n, padding, out_env);
  POB_TRY(pob_arena_reset(ctx));
  const uint8_t *d_a1, *d_a2;
  const int64_t *d_aoff, *d_so1, *d_so2, *d_eoff;
  const int32_t *d_alen, *d_m1, *d_m2, *d_l1, *d_l2, *d_U, *d_V;
  int32_t* d_env;
  POB_TRY(stage_in(ctx, a1, (size_t)aln_off[n], &d_a1));
  POB_TRY(stage_in(ctx, a2, (size_t)aln_off[n], &d_a2));
  POB_TRY(stage_in(ctx, aln_off, (size_t)n + 1, &d_aoff));
  POB_TRY(stage_in(ctx, alen, (size_t)n, &d_alen));
  POB_TRY(stage_in(ctx, s2s1, (size_t)soff1[n], &d_m1));
  POB_TRY(stage_in(ctx, soff1, (size_t)n + 1, &d_so1));
  POB_TRY(stage_in(ctx, slen1, (size_t)n, &d_l1));
  POB_TRY(stage_in(ctx, s2s2, (size_t)soff2[n], &d_m2));
  POB_TRY(stage_in(ctx, soff2, (size_t)n + 1, &d_so2));
  POB_TRY(stage_in(ctx, slen2, (size_t)n, &d_l2));
  POB_TRY(stage_in(ctx, U, (size_t)n, &d_U));
  POB_TRY(stage_in(ctx, V, (size_t)n, &d_V));
  POB_TRY(stage_in(ctx, env_off, (size_t)n + 1, &d_eoff));
  POB_TRY(stage_out(ctx, out_env, (size_t)env_off[n] * 2, &d_env));
  POB_TRY(pob_envelope_launch(ctx, d_a1, d_a2, d_aoff, d_alen, d_m1, d_so1, d_l1, d_m2, d_so2, d_l2, d_U, d_V,
                              d_eoff, nullptr, n, padding, d_env));
  POB_TRY(copy_back(ctx, out_env, d_env, (size_t)env_off[n] * 2));
  POB_CUDA(cudaStreamSynchronize(ctx->stream));
  return POB_OK;
}
int pob_forward(pob_ctx* ctx, int where, const pob_reads_t* reads, const uint8_t* labels, const int64_t* lab_off,
                int model, double* out_logp) {
  if (!ctx) return POB_EINVAL;
  POB_TRY(check_reads(reads, 2, 9));
  if (reads->dtype != POB_F32 && reads->dtype != POB_F64) return POB_EINVAL;
  if (model != POB_MODEL_CTC && model != POB_MODEL_CTC_MERGE_REPEATS) return POB_EINVAL;
  const int n = reads->n;
  if (n == 0) return POB_OK;
  if (!labels || !lab_off || !out_logp) return POB_EINVAL;
  POB_CUDA(cudaSetDevice(ctx->device));
  POB_TRY(pob_arena_reset(ctx));
  std::vector<int64_t> off, loff;
  std::vector<int32_t> len;
  POB_TRY(fetch_i64(ctx, where, reads->row_off, (size_t)n + 1, off));
  POB_TRY(fetch_i64(ctx, where, lab_off, (size_t)n + 1, loff));
  if (reads->row_len) POB_TRY(fetch_i32(ctx, where, reads->row_len, (size_t)n, len));
  std::vector<int64_t> scr_off(n + 1);
  scr_off[0] = 0;
  for (int i = 0; i < n; ++i) {
    const int64_t T = reads->row_len ? len[i] : off[i + 1] - off[i];
    scr_off[i + 1] = scr_off[i] + 5 * T + 8;
  }
  pob_reads_t d = *reads;
  const uint8_t* d_lab = labels;
  const int64_t* d_loff = lab_off;
  double* d_out = out_logp;
  if (where == POB_HOST) {
    POB_TRY(stage_reads(ctx, reads, &d));
    POB_TRY(stage_in(ctx, labels, (size_t)loff[n], &d_lab, 8));
    POB_TRY(stage_in(ctx, lab_off, (size_t)n + 1, &d_loff));
    POB_TRY(stage_out(ctx, out_logp, (size_t)n, &d_out));
  }
  const int64_t* d_scr_off;
  POB_TRY(upload(ctx, scr_off, &d_scr_off));
  double* scratch;
  POB_TRY(pob_take(ctx, (size_t)scr_off[n] + 8, &scratch));
  POB_TRY(pob_forward_launch(ctx, d, d_lab, d_loff, model, scratch, d_scr_off, d_out));
  if (where == POB_HOST) POB_TRY(copy_back(ctx, out_logp, d_out, (size_t)n));
  POB_CUDA(cudaStreamSynchronize(ctx->stream));
  return POB_OK;
}

int pob_viterbi_acceptor(pob_ctx* ctx, int where, const pob_reads_t* reads, const uint8_t* labels,
                         const int64_t* lab_off, int band_size, int8_t* out_path, int32_t* out_status) {
  if (!ctx) return POB_EINVAL;
  POB_TRY(check_reads(reads, 2, 9));
  if (reads->dtype != POB_F32 && reads->dtype != POB_F64) return POB_EINVAL;
  if (band_size < 1) return POB_EINVAL;
  const int n = reads->n;
  if (n == 0) return POB_OK;
  if (!labels || !lab_off || !out_path) return POB_EINVAL;
  POB_CUDA(cudaSetDevice(ctx->device));
  POB_TRY(pob_arena_reset(ctx));
  std::vector<int64_t> off, loff;
  std::vector<int32_t> len;
  POB_TRY(fetch_i64(ctx, where, reads->row_off, (size_t)n + 1, off));
  POB_TRY(fetch_i64(ctx, where, lab_off, (size_t)n + 1, loff));
  if (reads->row_len) POB_TRY(fetch_i32(ctx, where, reads->row_len, (size_t)n, len));
  const int SZ = pob_acceptor_slots(band_size);
  std::vector<int64_t> bp_off(n + 1);
  bp_off[0] = 0;
  for (int i = 0; i < n; ++i) {
    const int64_t T = reads->row_len ? len[i] : off[i + 1] - off[i];
    const int64_t L = loff[i + 1] - loff[i];
    if (L > T + 1 && T > 0) return POB_EINVAL;  // more bases than timesteps can never be placed
    bp_off[i + 1] = bp_off[i] + (T + L + 1) * (SZ / 32);
  }
  pob_reads_t d = *reads;
  const uint8_t* d_lab = labels;
  const int64_t* d_loff = lab_off;
  int8_t* d_path = out_path;
  int32_t* d_status = out_status;
  if (where == POB_HOST) {
    POB_TRY(stage_reads(ctx, reads, &d));
    POB_TRY(stage_in(ctx, labels, (size_t)loff[n], &d_lab, 8));
    POB_TRY(stage_in(ctx, lab_off, (size_t)n + 1, &d_loff));
    POB_TRY(stage_out(ctx, out_path, (size_t)off[n] + 4, &d_path));
    POB_TRY(stage_out(ctx, out_status, (size_t)n, &d_status, true));
  } else if (!d_status) {
    POB_TRY(pob_take(ctx, (size_t)n, &d_status));
  }
  const int64_t* d_bp_off;
  POB_TRY(upload(ctx, bp_off, &d_bp_off));
  uint32_t* bp;
  double* cum;
  POB_TRY(pob_take(ctx, (size_t)bp_off[n] + 8, &bp));
  POB_TRY(pob_take(ctx, (size_t)off[n] + 8, &cum));
  POB_TRY(pob_acceptor_launch(ctx, d, d_lab, d_loff, band_size, SZ, d_bp_off, bp, cum, d_path, d_status));
  if (where == POB_HOST) {
    POB_TRY(copy_back(ctx, out_path, d_path, (size_t)off[n]));
    POB_TRY(copy_back(ctx, out_status, d_status, (size_t)n));
  }
  POB_CUDA(cudaStreamSynchronize(ctx->stream));
  return POB_OK;
}

int pob_align_global(pob_ctx* ctx, int where, const uint8_t* seq1, const int64_t* off1, const uint8_t* seq2,
                     const int64_t* off2, int n, int match, int mismatch, int gap, uint8_t* out_a1, uint8_t* out_a2,
                     int32_t* out_alen, int32_t* out_matches, const int64_t* dp_off, int32_t* out_dp) {
  if (!ctx || n < 0) return POB_EINVAL;
  if (n == 0) return POB_OK;
  if (!seq1 || !seq2 || !off1 || !off2 || !out_a1 || !out_a2 || !out_alen) return POB_EINVAL;
  if (out_dp && !dp_off) return POB_EINVAL;
  POB_CUDA(cudaSetDevice(ctx->device));
  POB_TRY(pob_arena_reset(ctx));
  std::vector<int64_t> o1, o2;
  POB_TRY(fetch_i64(ctx, where, off1, (size_t)n + 1, o1));
  POB_TRY(fetch_i64(ctx, where, off2, (size_t)n + 1, o2));
  std::vector<int64_t> h_dp_off(n + 1), aln_off(n + 1);
  h_dp_off[0] = 0;
  for (int p = 0; p < n; ++p) {
    const int64_t l1 = o1[p + 1] - o1[p], l2 = o2[p + 1] - o2[p];
    h_dp_off[p + 1] = h_dp_off[p] + (l1 + 1) * (l2 + 1);
    aln_off[p] = o1[p] + o2[p] + 8 * (int64_t)p;
  }
  aln_off[n] = o1[n] + o2[n] + 8 * (int64_t)n;
  const uint8_t *d_s1 = seq1, *d_s2 = seq2;
  const int64_t *d_o1 = off1, *d_o2 = off2;
  uint8_t *d_a1 = out_a1, *d_a2 = out_a2;
  int32_t *d_alen = out_alen, *d_match = out_matches;
  if (where == POB_HOST) {
    POB_TRY(stage_in(ctx, seq1, (size_t)o1[n], &d_s1, 8));
    POB_TRY(stage_in(ctx, seq2, (size_t)o2[n], &d_s2, 8));
    POB_TRY(stage_in(ctx, off1, (size_t)n + 1, &d_o1));
    POB_TRY(stage_in(ctx, off2, (size_t)n + 1, &d_o2));
    POB_TRY(stage_out(ctx, out_a1, (size_t)aln_off[n], &d_a1));
    POB_TRY(stage_out(ctx, out_a2, (size_t)aln_off[n], &d_a2));
    POB_TRY(stage_out(ctx, out_alen, (size_t)n, &d_alen));
    POB_TRY(stage_out(ctx, out_matches, (size_t)n, &d_match));
  }
  const int64_t *d_dpoff, *d_alnoff;
  POB_TRY(upload(ctx, h_dp_off, &d_dpoff));
  POB_TRY(upload(ctx, aln_off, &d_alnoff));
  int32_t* dp;
  if (where == POB_DEVICE && out_dp) dp = out_dp;  // caller's packing must equal (l1+1)*(l2+1) per pair
  else POB_TRY(pob_take(ctx, (size_t)h_dp_off[n] + 1, &dp));
  POB_TRY(pob_global_pair_launch(ctx, d_s1, d_o1, d_s2, d_o2, n, match, mismatch, gap, d_dpoff, dp, d_alnoff, d_a1,
                                 d_a2, d_alen, d_match));
  if (where == POB_HOST) {
    POB_TRY(copy_back(ctx, out_a1, d_a1, (size_t)aln_off[n]));
    POB_TRY(copy_back(ctx, out_a2, d_a2, (size_t)aln_off[n]));
    POB_TRY(copy_back(ctx, out_alen, d_alen, (size_t)n));
    POB_TRY(copy_back(ctx, out_matches, d_match, (size_t)n));
    if (out_dp) POB_TRY(copy_back(ctx, out_dp, dp, (size_t)h_dp_off[n]));
  }
  POB_CUDA(cudaStreamSynchronize(ctx->stream));
  return POB_OK;
}

}  // extern "C"
