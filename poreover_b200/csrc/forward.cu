// Exact label forward log-probability and full (unbanded) Needleman-Wunsch: the two remaining native entry
// points of the reference's Cython modules that are off the default pair-decode path.
//
//   forward   : decoding_cpp.cpp_forward (decoding_cpp.pyx:49-65) -> forward_ (PrefixTree.h:710-749): for every
//               prefix of the label, update_prob over all t; returns last_probability() of the full label.
//   global_pair: align.global_pair (align.pyx:29-98): dense DP with boundary gaps, recompute-style traceback
//               that applies every tied move.
#include "common.cuh"
#include "launch.cuh"

namespace {

__device__ __forceinline__ double f_ninf() { return __longlong_as_double(0xfff0000000000000LL); }

__device__ __forceinline__ double f_lae(double a, double b) {  // Log.h:27-33, bounded term in FP32
  const double m = fmax(a, b);
  if (m == f_ninf()) return m;
  const float d = (float)(fmin(a, b) - m);
  return m + (double)log1pf(expf(d));
}

constexpr int FW_THREADS = 256;

// One CTA per (read, label).  Label positions are processed in tiles of FW_THREADS; inside a tile thread i
// owns position s0+i and walks time; at wavefront step k it is at t = k - i, so its parent's value at t-1 was
// published two steps earlier (triple-buffered shared memory, one barrier per step).  The last position of a
// tile streams its (prob, gap) column to global memory for the first thread of the next tile.
template <int MODEL>
__global__ void __launch_bounds__(FW_THREADS)
forward_kernel(pob_reads rd, const uint8_t* __restrict__ labels, const int64_t* __restrict__ lab_off,
               double* __restrict__ scratch, const int64_t* __restrict__ scr_off, double* __restrict__ out) {
  __shared__ double2 pubv[3][FW_THREADS];
  const int item = blockIdx.x;
  const int tid = threadIdx.x;
  const int64_t ro = rd.row_off[item];
  const int T = pob_read_len(rd.row_off, rd.row_len, item);
  const int L = (int)(lab_off[item + 1] - lab_off[item]);
  const uint8_t* lab = labels + lab_off[item];
  const bool rc = rd.rc ? rd.rc[item] != 0 : false;
  const bool f64 = rd.dtype == POB_F64;
  const int S = rd.n_states;
  const char* base = (const char*)rd.data + (size_t)ro * S * (f64 ? 8 : 4);
  auto y_at = [&](int t, int k) -> double {
    const size_t idx = (size_t)(rc ? (T - 1 - t) : t) * S + pob_col(k, S, rd.layout, rc);
    return f64 ? ((const double*)base)[idx] : (double)((const float*)base)[idx];
  };
  double* bnd = scratch + scr_off[item];  // [2][2][T] ping-pong columns (prob, gap) of a tile's last position
  double* cum = bnd + 4 * (size_t)T;      // [T] ctc: running blank sum (the root's values)
  if (L <= 0 || T <= 0) {
    if (tid == 0) out[item] = 0.0;
    return;
  }
  if (MODEL == POB_MODEL_CTC && tid == 0) {
    double s = 0;
    for (int t = 0; t < T; ++t) { s += y_at(t, S - 1); cum[t] = s; }  // PrefixTree.h:470-475
  }
  __syncthreads();
  double result = f_ninf();
  for (int s0 = 0, tile = 0; s0 < L; s0 += FW_THREADS, ++tile) {
    const int len = min(FW_THREADS, L - s0);
    const double* rd_prob = bnd + (size_t)((tile + 1) & 1) * 2 * T;  // written by the previous tile
    const double* rd_gap = rd_prob + T;
    double* wr_prob = bnd + (size_t)(tile & 1) * 2 * T;
    double* wr_gap = wr_prob + T;
    const int s = s0 + tid;
    const bool mine = tid < len;
    const int last = mine ? lab[s] : 0;
    const bool same = mine && s > 0 && lab[s - 1] == last;
    double p_prev = f_ninf(), ng_prev = f_ninf();
    const int steps = T + len - 1;
    for (int k = 0; k < steps; ++k) {
      const int t = k - tid;
      double2 pb = make_double2(f_ninf(), f_ninf());
      if (mine && t >= 0 && t < T) {
        // parent values at t-1
        double pprob, pgap;
        if (tid > 0) {
          if (t - 1 >= 0) { const double2 q = pubv[(k + 1) % 3][tid - 1]; pprob = q.x; pgap = q.y; }  // step k-2
          else { pprob = f_ninf(); pgap = f_ninf(); }
        } else if (s == 0) {
          // the root: prob(-1) = gap(-1) = 0; ctc keeps the blank prefix sums, merge-repeats nothing else
          if (t == 0) { pprob = 0.0; pgap = 0.0; }
          else { pprob = (MODEL == POB_MODEL_CTC) ? cum[t - 1] : f_ninf(); pgap = f_ninf(); }
        } else {
          if (t - 1 >= 0) { pprob = rd_prob[t - 1]; pgap = rd_gap[t - 1]; }
          else { pprob = f_ninf(); pgap = f_ninf(); }
        }
        const double yl = y_at(t, last), yb = y_at(t, S - 1);
        double prob, gp = f_ninf();
        if (MODEL == POB_MODEL_CTC_MERGE_REPEATS) {
          gp = p_prev + yb;
          const double ng = f_lae((same ? pgap : pprob) + yl, ng_prev + yl);
          prob = f_lae(gp, ng);
          ng_prev = ng;
        } else {
          prob = f_lae(pprob + yl, p_prev + yb);
        }
        p_prev = prob;
        pb = make_double2(prob, gp);
        if (tid == len - 1) { wr_prob[t] = prob; wr_gap[t] = gp; }  // read by the next tile (after the loop)
        if (s == L - 1 && t == T - 1) result = prob;
      }
      pubv[k % 3][tid] = pb;
      __syncthreads();
    }
    __syncthreads();
  }
  // the thread that owned (L-1, T-1) holds the answer
  if (tid == (L - 1) % FW_THREADS) out[item] = result;
}

// ---------------------------------------------------------------------------------------------
// align.pyx:29-98.  dp is (l1+1) x (l2+1) int32 row-major.
constexpr int GP_THREADS = 512;

__global__ void __launch_bounds__(GP_THREADS)
global_pair_fill_kernel(const uint8_t* __restrict__ seq1, const int64_t* __restrict__ off1,
                        const uint8_t* __restrict__ seq2, const int64_t* __restrict__ off2, int match, int mismatch,
                        int gap, const int64_t* __restrict__ dp_off, int32_t* __restrict__ dp_all) {
  const int p = blockIdx.x;
  const int l1 = (int)(off1[p + 1] - off1[p]), l2 = (int)(off2[p + 1] - off2[p]);
  const uint8_t* s1 = seq1 + off1[p];
  const uint8_t* s2 = seq2 + off2[p];
  int32_t* dp = dp_all + dp_off[p];
  const size_t Wd = (size_t)l2 + 1;
  for (int i = threadIdx.x; i <= l1; i += GP_THREADS) dp[(size_t)i * Wd] = gap * i;
  for (int j = threadIdx.x; j <= l2; j += GP_THREADS) dp[j] = gap * j;
  __syncthreads();
  for (int d = 2; d <= l1 + l2; ++d) {  // cells with i + j == d, 1 <= i <= l1, 1 <= j <= l2
    const int ilo = max(1, d - l2), ihi = min(l1, d - 1);
    for (int i = ilo + threadIdx.x; i <= ihi; i += GP_THREADS) {
      const int j = d - i;
      const int sc = (s1[i - 1] == s2[j - 1]) ? match : mismatch;
      const int dg = dp[(size_t)(i - 1) * Wd + j - 1] + sc;
      const int up = dp[(size_t)(i - 1) * Wd + j] + gap;
      const int lf = dp[(size_t)i * Wd + j - 1] + gap;
      dp[(size_t)i * Wd + j] = max(dg, max(up, lf));
    }
    __syncthreads();
  }
}

__device__ __forceinline__ uint8_t gp_wrap(const uint8_t* s, int n, int i) { return s[i < 0 ? i + n : i]; }

__global__ void __launch_bounds__(128)
global_pair_trace_kernel(const uint8_t* __restrict__ seq1, const int64_t* __restrict__ off1,
                         const uint8_t* __restrict__ seq2, const int64_t* __restrict__ off2, int n, int gap,
                         const int64_t* __restrict__ dp_off, const int32_t* __restrict__ dp_all,
                         const int64_t* __restrict__ aln_off, uint8_t* __restrict__ out_a1,
                         uint8_t* __restrict__ out_a2, int32_t* __restrict__ out_alen,
                         int32_t* __restrict__ out_matches) {
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= n) return;
  const int l1 = (int)(off1[p + 1] - off1[p]), l2 = (int)(off2[p + 1] - off2[p]);
  const uint8_t* s1 = seq1 + off1[p];
  const uint8_t* s2 = seq2 + off2[p];
  const int32_t* dp = dp_all + dp_off[p];
  const size_t Wd = (size_t)l2 + 1;
  uint8_t* a1 = out_a1 + aln_off[p];
  uint8_t* a2 = out_a2 + aln_off[p];
  const int cap = (int)(aln_off[p + 1] - aln_off[p]);
  int n_col = 0;
  if (lane == 0) {
    int i = l1, j = l2, w = cap;
    while (i > 0 && j > 0) {
      const int sc = (s1[i - 1] == s2[j - 1]) ? 2 : -1;  // default scores, whatever was passed (align.pyx:64)
      const int c0 = dp[(size_t)(i - 1) * Wd + j - 1] + sc;
      const int c1 = dp[(size_t)(i - 1) * Wd + j] + gap;
      const int c2 = dp[(size_t)i * Wd + j - 1] + gap;
      const int mx = max(c0, max(c1, c2));
      if (c0 == mx) { --i; --j; --w; a1[w] = gp_wrap(s1, l1, i); a2[w] = gp_wrap(s2, l2, j); }
      if (c1 == mx) { --i; --w; a1[w] = gp_wrap(s1, l1, i); a2[w] = '-'; }
      if (c2 == mx) { --j; --w; a1[w] = '-'; a2[w] = gp_wrap(s2, l2, j); }
    }
    while (i > 0 || j > 0) {
      if (i > 0) { --i; --w; a1[w] = s1[i]; a2[w] = '-'; }
      else { --j; --w; a1[w] = '-'; a2[w] = s2[j]; }
    }
    n_col = cap - w;
  }
  n_col = __shfl_sync(0xffffffffu, n_col, 0);
  __syncwarp();
  const int src = cap - n_col;
  int matches = 0;
  for (int c0 = 0; c0 < n_col; c0 += 32) {
    const int c = c0 + lane;
    uint8_t x = 0, y = 1;
    if (c < n_col) { x = a1[src + c]; y = a2[src + c]; }
    __syncwarp();
    if (c < n_col) { a1[c] = x; a2[c] = y; matches += (x == y); }
    __syncwarp();
  }
  matches = __reduce_add_sync(0xffffffffu, matches);
  if (lane == 0) {
    out_alen[p] = n_col;
    if (out_matches) out_matches[p] = matches;
  }
}

}  // namespace

int pob_forward_launch(pob_ctx* ctx, const pob_reads& rd, const uint8_t* labels, const int64_t* lab_off, int model,
                       double* scratch, const int64_t* scr_off, double* out) {
  if (rd.n <= 0) return POB_OK;
  pob_prof_scope ps(ctx, POB_K_FORWARD);
  if (model == POB_MODEL_CTC)
    forward_kernel<POB_MODEL_CTC><<<rd.n, FW_THREADS, 0, ctx->stream>>>(rd, labels, lab_off, scratch, scr_off, out);
  else
    forward_kernel<POB_MODEL_CTC_MERGE_REPEATS><<<rd.n, FW_THREADS, 0, ctx->stream>>>(rd, labels, lab_off, scratch,
                                                                                       scr_off, out);
  POB_CUDA(cudaGetLastError());
  return POB_OK;
}

int pob_global_pair_launch(pob_ctx* ctx, const uint8_t* seq1, const int64_t* off1, const uint8_t* seq2,
                           const int64_t* off2, int n, int match, int mismatch, int gap, const int64_t* dp_off,
                           int32_t* dp, const int64_t* aln_off, uint8_t* out_a1, uint8_t* out_a2, int32_t* out_alen,
                           int32_t* out_matches) {
  if (n <= 0) return POB_OK;
  {
    pob_prof_scope ps(ctx, POB_K_NW_FILL);
    global_pair_fill_kernel<<<n, GP_THREADS, 0, ctx->stream>>>(seq1, off1, seq2, off2, match, mismatch, gap, dp_off, dp);
  }
  POB_CUDA(cudaGetLastError());
  {
    pob_prof_scope ps(ctx, POB_K_NW_TRACE);
    global_pair_trace_kernel<<<(n + 3) / 4, 128, 0, ctx->stream>>>(seq1, off1, seq2, off2, n, gap, dp_off, dp, aln_off,
                                                                  out_a1, out_a2, out_alen, out_matches);
  }
  POB_CUDA(cudaGetLastError());
  return POB_OK;
}
