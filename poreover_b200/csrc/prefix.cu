// Legacy prefix search (Graves-style best-first search over a single growing prefix), 1D and dense 2D, and the
// dense gamma matrix the 2D search scores against.  Replaces (reference tree):
//   prefix_search.prefix_search_log / prefix_search_log_cy          decoding/prefix_search.py:116-174 / :176-238
//   prefix_search.pair_gamma_log / decoding_cy.pair_gamma_log       decoding/prefix_search.py:35-65 / decoding_cy.pyx:177-220
//   prefix_search.pair_prefix_search_log / ..._cy                   decoding/prefix_search.py:247-310 / :312-385
// The C++ envelope variant (PairPrefixSearch.cpp) is not reproduced: it copies its SparseMatrix arguments by value
// and double-frees (Gamma.h:100); the dense Python functions above define the results.
//
// Everything is FP64 in the log domain, with the reference's own formulas.  Two arithmetic flavours, because the
// reference has two: POB_PREFIX_NUMPY (np.logaddexp, scipy logsumexp, LOG_0 = -inf) for the plain functions, and
// POB_PREFIX_CY (log(exp(a) + exp(b)) without a shift, forward vectors initialised to -9999, decoding_cy.pyx:18,
// :127-156) for the *_cy functions.  The search itself (which prefix is extended, when it stops, which label is
// returned) follows the Python loops statement by statement, including their differences between 1D and 2D
// (1D: the top label is updated while the candidates are evaluated; 2D: after the stop test).
#include "staging.cuh"

namespace {

__device__ __forceinline__ double px_ninf() { return __longlong_as_double(0xfff0000000000000LL); }

// np.logaddexp (npy_logaddexp, numpy/core/src/npymath/npy_math_internal.h.src)
__device__ __forceinline__ double np_logaddexp(double x, double y) {
  if (x == y) return x + 0.6931471805599453;  // also covers equal infinities
  const double d = x - y;
  if (d > 0) return x + log1p(exp(-d));
  if (d <= 0) return y + log1p(exp(d));
  return d;  // NaN
}
// decoding_cy.pyx:154, :215-216
__device__ __forceinline__ double cy_logaddexp(double x, double y) { return log(exp(x) + exp(y)); }

template <int FL>
__device__ __forceinline__ double px_lae(double x, double y) {
  return FL == POB_PREFIX_CY ? cy_logaddexp(x, y) : np_logaddexp(x, y);
}
// what a forward vector is initialised with: prefix_search.py:83 (LOG_0 = -inf) / decoding_cy.pyx:141 (-9999)
template <int FL>
__device__ __forceinline__ double px_fw0() { return FL == POB_PREFIX_CY ? -9999.0 : px_ninf(); }

// ---------------------------------------------------------------------------------------------------------------
// 1D search: four lanes per window (lane c evaluates the extension by letter c), eight windows per warp.
// Scratch per window: five vectors of T doubles (the current prefix's forward vector + one per candidate).
// ---------------------------------------------------------------------------------------------------------------
constexpr int PX1_THREADS = 128;

template <int FL>
__global__ void __launch_bounds__(PX1_THREADS)
prefix1d_kernel(const double* __restrict__ y, const int64_t* __restrict__ row_off, int n, int S,
                double* __restrict__ scratch, const int64_t* __restrict__ scr_off, const int64_t* __restrict__ lab_off,
                uint8_t* __restrict__ work_lab, uint8_t* __restrict__ out_lab, int32_t* __restrict__ out_len,
                double* __restrict__ out_score, int32_t* __restrict__ out_status) {
  const int gid = blockIdx.x * (PX1_THREADS / 4) + (threadIdx.x >> 2);
  const int c = threadIdx.x & 3;
  const int lane = threadIdx.x & 31;
  const unsigned gmask = 0xFu << (lane & ~3);
  const int gbase = lane & ~3;
  const bool live = gid < n;
  const int w = live ? gid : 0;
  const int A = S - 1;
  const int64_t r0 = row_off[w];
  const int T = live ? (int)(row_off[w + 1] - r0) : 0;
  const double* yw = y + (size_t)r0 * S;
  double* buf = scratch + scr_off[w];
  uint8_t* lab = work_lab + lab_off[w];
  uint8_t* olab = out_lab + lab_off[w];
  if (T <= 0) {
    if (live && c == 0) { out_len[w] = 0; out_score[w] = 0.0; if (out_status) out_status[w] = POB_ST_EMPTY; }
    return;  // the whole group leaves together
  }
  // forward vector of the empty prefix (forward_vec_log(-1, 0, y): blanks only) and its label probability
  int slot_prev = 0, slot_mine = c + 1;
  double gap_prob = 0.0;
  if (c == 0) {
    double f = 0.0;
    for (int t = 0; t < T; ++t) {
      const double b = yw[(size_t)t * S + A];
      f = (t == 0) ? b : b + f;
      buf[t] = f;
      gap_prob += b;  // np.sum(y[:, -1]) (sequential here, pairwise in numpy: last-bit differences)
    }
  }
  gap_prob = __shfl_sync(gmask, gap_prob, gbase);
  __syncwarp(gmask);
  double top_prob = gap_prob;
  int top_len = 0, top_last = 0;
  int cur_len = 0;
  int status = 0;
  const int max_levels = T + 2;
  for (int level = 1;; ++level) {
    const double* prev = buf + (size_t)slot_prev * T;
    double* mine = buf + (size_t)slot_mine * T;
    double prefix_prob = px_ninf(), label_prob = px_ninf();
    if (c < A) {
      // alpha_ast = insert(prev[:-1], 0, LOG_1 or LOG_0) + y[:, c]  (forward_vec_no_gap_log, prefix_search.py:67-79);
      // scipy logsumexp: shift by the maximum (0 when it is not finite)
      double m = px_ninf();
      for (int t = 0; t < T; ++t) {
        const double a = ((t == 0) ? (level == 1 ? 0.0 : px_ninf()) : prev[t - 1]) + yw[(size_t)t * S + c];
        m = fmax(m, a);
      }
      const double shift = isfinite(m) ? m : 0.0;
      double s = 0.0;
      // forward_vec_log(c, level, y, previous) in the same pass (prefix_search.py:81-97 / decoding_cy.pyx:127-156)
      double f = 0.0;
      for (int t = 0; t < T; ++t) {
        const double yc = yw[(size_t)t * S + c];
        const double a = ((t == 0) ? (level == 1 ? 0.0 : px_ninf()) : prev[t - 1]) + yc;
        s += exp(a - shift);
        if (t == 0) f = (level == 1) ? yc : px_fw0<FL>();
        else f = px_lae<FL>(yw[(size_t)t * S + A] + f, yc + prev[t - 1]);
        mine[t] = f;
      }
      prefix_prob = log(s) + shift;
      label_prob = f;
    }
    // the candidates in alphabet order (prefix_search.py:136-157): top label while evaluating, first best prefix
    int best = 0;
    double best_prob = __shfl_sync(gmask, prefix_prob, gbase);
    for (int k = 0; k < A; ++k) {
      const double lp = __shfl_sync(gmask, label_prob, gbase + k);
      const double pp = __shfl_sync(gmask, prefix_prob, gbase + k);
      if (lp > top_prob) { top_prob = lp; top_len = cur_len + 1; top_last = k; }
      if (pp > best_prob) { best_prob = pp; best = k; }
    }
    if (best_prob < top_prob) break;
    if (level >= max_levels) { status |= POB_ST_MAX_DEPTH; break; }
    // move to the best prefix: its forward vector becomes `prev`, the old one becomes that lane's scratch
    if (c == 0) lab[cur_len] = (uint8_t)best;
    cur_len++;
    const int best_slot = __shfl_sync(gmask, slot_mine, gbase + best);
    if (c == best) slot_mine = slot_prev;
    slot_prev = best_slot;
    __syncwarp(gmask);
  }
  __syncwarp(gmask);
  if (c == 0) {
    // top label = the prefix that was current when it was found + its last letter
    for (int i = 0; i + 1 < top_len; ++i) olab[i] = lab[i];
    if (top_len > 0) olab[top_len - 1] = (uint8_t)top_last;
    out_len[w] = top_len;
    out_score[w] = top_prob;
    if (out_status) out_status[w] = status;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// dense gamma (prefix_search.py:35-65): one CTA per pair, anti-diagonal wavefront from (U-1, V-1) to (0, 0).
// gamma and gamma_ast are (U+1) x (V+1) row-major.
// ---------------------------------------------------------------------------------------------------------------
constexpr int PX2_THREADS = 256;

template <int FL>
__device__ void gamma_fill(const double* __restrict__ y1, const double* __restrict__ y2, int U, int V, int S,
                           double* __restrict__ g, double* __restrict__ ga) {
  const int tid = threadIdx.x, NT = blockDim.x;
  const int A = S - 1;
  const size_t W = (size_t)V + 1;
  const double L0 = px_fw0<FL>();  // np.zeros + LOG_0: -inf (prefix_search.py:44-45) / -9999 (decoding_cy.pyx:187-190)
  for (size_t i = tid; i < ((size_t)U + 1) * W; i += NT) { g[i] = L0; ga[i] = L0; }
  __syncthreads();
  if (tid == 0) { g[(size_t)U * W + V] = 0.0; ga[(size_t)U * W + V] = 0.0; }
  // boundaries: sums of the blank column from v (u) to the end, summed forward as the reference does
  for (int v = tid; v < V; v += NT) {
    double s = 0.0;
    for (int k = v; k < V; ++k) s += y2[(size_t)k * S + A];
    g[(size_t)U * W + v] = s;
  }
  for (int u = tid; u < U; u += NT) {
    double s = 0.0;
    for (int k = u; k < U; ++k) s += y1[(size_t)k * S + A];
    g[(size_t)u * W + V] = s;
  }
  __syncthreads();
  for (int d = U + V - 2; d >= 0; --d) {
    const int ulo = max(0, d - (V - 1)), uhi = min(U - 1, d);
    for (int u = ulo + tid; u <= uhi; u += NT) {
      const int v = d - u;
      const double* a = y1 + (size_t)u * S;
      const double* b = y2 + (size_t)v * S;
      const double gamma_eps = g[(size_t)(u + 1) * W + v] + a[A];
      const double gamma_ast_eps = ga[(size_t)u * W + v + 1] + b[A];
      double tot;
      if (FL == POB_PREFIX_CY) {
        double s = 0.0;
        for (int k = 0; k < A; ++k) s += exp(a[k] + b[k]);
        tot = log(s);
      } else {
        double m = px_ninf();
        for (int k = 0; k < A; ++k) m = fmax(m, a[k] + b[k]);
        const double shift = isfinite(m) ? m : 0.0;
        double s = 0.0;
        for (int k = 0; k < A; ++k) s += exp(a[k] + b[k] - shift);
        tot = log(s) + shift;
      }
      const double gamma_ast_ast = g[(size_t)(u + 1) * W + v + 1] + tot;
      const double x = px_lae<FL>(gamma_ast_eps, gamma_ast_ast);
      ga[(size_t)u * W + v] = x;
      g[(size_t)u * W + v] = px_lae<FL>(gamma_eps, x);
    }
    __syncthreads();
  }
}

template <int FL>
__global__ void __launch_bounds__(PX2_THREADS)
pair_gamma_kernel(const double* __restrict__ y1, const int64_t* __restrict__ off1, const double* __restrict__ y2,
                  const int64_t* __restrict__ off2, int S, double* __restrict__ gamma, double* __restrict__ gamma_ast,
                  const int64_t* __restrict__ g_off) {
  const int p = blockIdx.x;
  const int U = (int)(off1[p + 1] - off1[p]), V = (int)(off2[p + 1] - off2[p]);
  gamma_fill<FL>(y1 + (size_t)off1[p] * S, y2 + (size_t)off2[p] * S, U, V, S, gamma + g_off[p], gamma_ast + g_off[p]);
}

// block-wide reductions over doubles (PX2_THREADS threads)
__device__ double block_max(double v, double* red) {
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = red[0];
  for (int k = 1; k < PX2_THREADS / 32; ++k) r = fmax(r, red[k]);
  __syncthreads();
  return r;
}
__device__ double block_sum(double v, double* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = red[0];
  for (int k = 1; k < PX2_THREADS / 32; ++k) r += red[k];
  __syncthreads();
  return r;
}

// ---------------------------------------------------------------------------------------------------------------
// dense 2D search (prefix_search.py:247-310): one CTA per pair.  Scratch per pair: gamma, gamma_ast, five forward
// vectors per read.  The two forward chains of every candidate run on the first lanes of eight warps; the prefix
// probability of a candidate is logsumexp over the whole U x V grid of alpha*_1[u] + alpha*_2[v] + gamma[u+1, v+1].
// ---------------------------------------------------------------------------------------------------------------
template <int FL>
__global__ void __launch_bounds__(PX2_THREADS)
pair_prefix_kernel(const double* __restrict__ y1g, const int64_t* __restrict__ off1, const double* __restrict__ y2g,
                   const int64_t* __restrict__ off2, int S, double* __restrict__ scratch,
                   const int64_t* __restrict__ scr_off, const int64_t* __restrict__ lab_off,
                   uint8_t* __restrict__ work_lab, uint8_t* __restrict__ out_lab, int32_t* __restrict__ out_len,
                   double* __restrict__ out_score, int32_t* __restrict__ out_status) {
  __shared__ double red[PX2_THREADS / 32];
  __shared__ double s_pp[4], s_lp1[4], s_lp2[4];
  __shared__ int s_slot_prev[2], s_slot_c[2][4];
  const int p = blockIdx.x, tid = threadIdx.x;
  const int U = (int)(off1[p + 1] - off1[p]), V = (int)(off2[p + 1] - off2[p]);
  const int A = S - 1;
  const double* y1 = y1g + (size_t)off1[p] * S;
  const double* y2 = y2g + (size_t)off2[p] * S;
  uint8_t* lab = work_lab + lab_off[p];
  uint8_t* olab = out_lab + lab_off[p];
  if (U <= 0 || V <= 0) {
    if (tid == 0) { out_len[p] = 0; out_score[p] = 0.0; if (out_status) out_status[p] = POB_ST_EMPTY; }
    return;
  }
  const size_t W = (size_t)V + 1;
  double* g = scratch + scr_off[p];
  double* ga = g + ((size_t)U + 1) * W;
  double* f1 = ga + ((size_t)U + 1) * W;  // [5][U]
  double* f2 = f1 + 5 * (size_t)U;        // [5][V]
  gamma_fill<FL>(y1, y2, U, V, S, g, ga);
  const double g00 = g[0];
  // empty prefix: forward vectors of blanks only; label_prob[''] = sum of both blank columns (prefix_search.py:264-265:
  // NOT divided by gamma[0,0], as in the reference)
  __shared__ double s_gap[2];
  if (tid == 0 || tid == 32) {
    const int r = tid >> 5;
    const double* y = r ? y2 : y1;
    double* f = r ? f2 : f1;
    const int T = r ? V : U;
    double acc = 0.0, s = 0.0;
    for (int t = 0; t < T; ++t) {
      const double b = y[(size_t)t * S + A];
      acc = (t == 0) ? b : b + acc;
      f[t] = acc;
      s += b;
    }
    s_gap[r] = s;
    s_slot_prev[r] = 0;
    for (int k = 0; k < 4; ++k) s_slot_c[r][k] = k + 1;
  }
  __syncthreads();
  double top_prob = s_gap[0] + s_gap[1];
  int top_len = 0, top_last = 0, cur_len = 0, status = 0;
  bool stop = false;
  const int max_len = max(U, V);
  for (int level = 1; !stop; ++level) {
    if (cur_len > max_len) { stop = true; status |= POB_ST_MAX_DEPTH; }  // prefix_search.py:279-281: the level still runs
    // forward vectors of the four candidates, both reads: warp w -> candidate w >> 1, read w & 1
    {
      const int wid = tid >> 5, k = wid >> 1, r = wid & 1;
      if ((tid & 31) == 0 && k < A) {
        const double* y = r ? y2 : y1;
        const int T = r ? V : U;
        double* fb = r ? f2 : f1;
        const double* prev = fb + (size_t)s_slot_prev[r] * T;
        double* mine = fb + (size_t)s_slot_c[r][k] * T;
        double f = 0.0;
        for (int t = 0; t < T; ++t) {
          const double yc = y[(size_t)t * S + k];
          if (t == 0) f = (level == 1) ? yc : px_fw0<FL>();
          else f = px_lae<FL>(y[(size_t)t * S + A] + f, yc + prev[t - 1]);
          mine[t] = f;
        }
        if (r) s_lp2[k] = f; else s_lp1[k] = f;
      }
    }
    __syncthreads();
    const double* prev1 = f1 + (size_t)s_slot_prev[0] * U;
    const double* prev2 = f2 + (size_t)s_slot_prev[1] * V;
    for (int k = 0; k < A; ++k) {
      // alpha*_r[t] = insert(prev_r[:-1], 0, LOG_1 / LOG_0)[t] + y_r[t, k]
      double m = px_ninf();
      for (size_t i = tid; i < (size_t)U * V; i += PX2_THREADS) {
        const int u = (int)(i / V), v = (int)(i - (size_t)u * V);
        const double a1 = ((u == 0) ? (level == 1 ? 0.0 : px_ninf()) : prev1[u - 1]) + y1[(size_t)u * S + k];
        const double a2 = ((v == 0) ? (level == 1 ? 0.0 : px_ninf()) : prev2[v - 1]) + y2[(size_t)v * S + k];
        m = fmax(m, (a1 + a2) + g[(size_t)(u + 1) * W + v + 1]);
      }
      m = block_max(m, red);
      const double shift = isfinite(m) ? m : 0.0;
      double s = 0.0;
      for (size_t i = tid; i < (size_t)U * V; i += PX2_THREADS) {
        const int u = (int)(i / V), v = (int)(i - (size_t)u * V);
        const double a1 = ((u == 0) ? (level == 1 ? 0.0 : px_ninf()) : prev1[u - 1]) + y1[(size_t)u * S + k];
        const double a2 = ((v == 0) ? (level == 1 ? 0.0 : px_ninf()) : prev2[v - 1]) + y2[(size_t)v * S + k];
        s += exp(((a1 + a2) + g[(size_t)(u + 1) * W + v + 1]) - shift);
      }
      s = block_sum(s, red);
      if (tid == 0) s_pp[k] = (log(s) + shift) - g00;
    }
    __syncthreads();
    // every thread runs the (tiny) decision with the same inputs
    int best = 0;
    double best_prob = s_pp[0];
    for (int k = 1; k < A; ++k) if (s_pp[k] > best_prob) { best_prob = s_pp[k]; best = k; }  // first maximum
    if (best_prob < top_prob) {
      stop = true;
    } else {
      // highest label probability over everything evaluated so far (first maximum in insertion order)
      for (int k = 0; k < A; ++k) {
        const double lp = (s_lp1[k] + s_lp2[k]) - g00;
        if (lp > top_prob) { top_prob = lp; top_len = cur_len + 1; top_last = k; }
      }
      if (tid == 0) lab[cur_len] = (uint8_t)best;
      cur_len++;
      __syncthreads();
      if (tid == 0) {
        for (int r = 0; r < 2; ++r) {
          const int t = s_slot_c[r][best];
          s_slot_c[r][best] = s_slot_prev[r];
          s_slot_prev[r] = t;
        }
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    for (int i = 0; i + 1 < top_len; ++i) olab[i] = lab[i];
    if (top_len > 0) olab[top_len - 1] = (uint8_t)top_last;
    out_len[p] = top_len;
    out_score[p] = top_prob;
    if (out_status) out_status[p] = status;
  }
}

// forward_vec_log(s, i, y, previous) (prefix_search.py:81-97 / decoding_cy.pyx:127-156): one column of the 1D forward
// algorithm, a chain over t (one thread: this is the building block the reference's tests call, not a hot path)
template <int FL>
__global__ void forward_vec_kernel(const double* __restrict__ y, int T, int S, int s, int i,
                                   const double* __restrict__ prev, double* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int A = S - 1;
  double f = 0.0;
  for (int t = 0; t < T; ++t) {
    if (i == 0) f = (t == 0) ? y[s] : y[(size_t)t * S + A] + f;
    else if (t == 0) f = (i == 1) ? y[s] : px_fw0<FL>();
    else f = px_lae<FL>(y[(size_t)t * S + A] + f, y[(size_t)t * S + s] + prev[t - 1]);
    out[t] = f;
  }
}

int check_offsets(const std::vector<int64_t>& off) {
  for (size_t i = 0; i + 1 < off.size(); ++i)
    if (off[i + 1] < off[i]) return POB_EINVAL;
  return POB_OK;
}

}  // namespace

extern "C" {

int pob_prefix_search(pob_ctx* ctx, int where, const double* y, const int64_t* row_off, int n, int n_states,
                      int flavour, const int64_t* lab_off, uint8_t* out_label, int32_t* out_len, double* out_score,
                      int32_t* out_status) {
  if (!ctx || n < 0) return POB_EINVAL;
  if (n_states < 2 || n_states > 5) return POB_EUNSUPPORTED;
  if (flavour != POB_PREFIX_NUMPY && flavour != POB_PREFIX_CY) return POB_EINVAL;
  if (n == 0) return POB_OK;
  if (!y || !row_off || !lab_off || !out_label || !out_len || !out_score) return POB_EINVAL;
  POB_CUDA(cudaSetDevice(ctx->device));
  POB_TRY(pob_arena_reset(ctx));
  std::vector<int64_t> off, loff;
  POB_TRY(fetch_i64(ctx, where, row_off, (size_t)n + 1, off));
  POB_TRY(fetch_i64(ctx, where, lab_off, (size_t)n + 1, loff));
  POB_TRY(check_offsets(off));
  if (off[0] != 0 || loff[0] != 0) return POB_EINVAL;
  std::vector<int64_t> scr_off((size_t)n + 1);
  scr_off[0] = 0;
  for (int i = 0; i < n; ++i) {
    const int64_t T = off[i + 1] - off[i];
    if (loff[i + 1] - loff[i] < T + 2) return POB_EINVAL;  // a label can be as long as the window (+ the search's slack)
    scr_off[i + 1] = scr_off[i] + 5 * T;
  }
  const double* d_y = y;
  const int64_t *d_off = row_off, *d_loff = lab_off;
  uint8_t* d_lab = out_label;
  int32_t *d_len = out_len, *d_st = out_status;
  double* d_sc = out_score;
  if (where == POB_HOST) {
    POB_TRY(stage_in(ctx, y, (size_t)off[n] * n_states, &d_y));
    POB_TRY(stage_in(ctx, row_off, (size_t)n + 1, &d_off));
    POB_TRY(stage_in(ctx, lab_off, (size_t)n + 1, &d_loff));
    POB_TRY(stage_out(ctx, out_label, (size_t)loff[n], &d_lab));
    POB_TRY(stage_out(ctx, out_len, (size_t)n, &d_len));
    POB_TRY(stage_out(ctx, out_score, (size_t)n, &d_sc));
    POB_TRY(stage_out(ctx, out_status, (size_t)n, &d_st));
  }
  const int64_t* d_scr_off;
  POB_TRY(upload(ctx, scr_off, &d_scr_off));
  double* scratch;
  POB_TRY(pob_take(ctx, (size_t)scr_off[n] + 8, &scratch));
  uint8_t* work;
  POB_TRY(pob_take(ctx, (size_t)loff[n] + 8, &work));
  const int per_block = PX1_THREADS / 4;
  const int grid = (n + per_block - 1) / per_block;
  {
    pob_prof_scope prof(ctx, POB_K_PREFIX_1D);
    if (flavour == POB_PREFIX_CY)
      prefix1d_kernel<POB_PREFIX_CY><<<grid, PX1_THREADS, 0, ctx->stream>>>(d_y, d_off, n, n_states, scratch, d_scr_off,
                                                                          d_loff, work, d_lab, d_len, d_sc, d_st);
    else
      prefix1d_kernel<POB_PREFIX_NUMPY><<<grid, PX1_THREADS, 0, ctx->stream>>>(d_y, d_off, n, n_states, scratch,
                                                                             d_scr_off, d_loff, work, d_lab, d_len,
                                                                             d_sc, d_st);
    POB_CUDA(cudaGetLastError());
  }
  if (where == POB_HOST) {
    POB_TRY(copy_back(ctx, out_label, d_lab, (size_t)loff[n]));
    POB_TRY(copy_back(ctx, out_len, d_len, (size_t)n));
    POB_TRY(copy_back(ctx, out_score, d_sc, (size_t)n));
    POB_TRY(copy_back(ctx, out_status, d_st, (size_t)n));
    POB_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return POB_OK;
}

int pob_forward_vec(pob_ctx* ctx, int where, const double* y, int rows, int n_states, int flavour, int s, int i,
                    const double* previous, double* out) {
  if (!ctx || rows < 0 || i < 0) return POB_EINVAL;
  if (n_states < 2 || n_states > 5) return POB_EUNSUPPORTED;
  if (flavour != POB_PREFIX_NUMPY && flavour != POB_PREFIX_CY) return POB_EINVAL;
  if (s < 0) s += n_states;  // the reference indexes y[t, s] with Python semantics: -1 is the blank
  if (s < 0 || s >= n_states) return POB_EINVAL;
  if (rows == 0) return POB_OK;
  if (!y || !out || (i > 0 && !previous)) return POB_EINVAL;
  POB_CUDA(cudaSetDevice(ctx->device));
  POB_TRY(pob_arena_reset(ctx));
  const double *d_y = y, *d_prev = previous;
  double* d_out = out;
  if (where == POB_HOST) {
    POB_TRY(stage_in(ctx, y, (size_t)rows * n_states, &d_y));
    POB_TRY(stage_in(ctx, previous, (size_t)rows, &d_prev));
    POB_TRY(stage_out(ctx, out, (size_t)rows, &d_out));
  }
  {
    pob_prof_scope prof(ctx, POB_K_PREFIX_1D);
    if (flavour == POB_PREFIX_CY) forward_vec_kernel<POB_PREFIX_CY><<<1, 32, 0, ctx->stream>>>(d_y, rows, n_states, s, i, d_prev, d_out);
    else forward_vec_kernel<POB_PREFIX_NUMPY><<<1, 32, 0, ctx->stream>>>(d_y, rows, n_states, s, i, d_prev, d_out);
    POB_CUDA(cudaGetLastError());
  }
  if (where == POB_HOST) {
    POB_TRY(copy_back(ctx, out, d_out, (size_t)rows));
    POB_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return POB_OK;
}

// shared front end of the two dense 2D entry points: offsets to the host, sizes checked, inputs staged
static int pair_inputs(pob_ctx* ctx, int where, const double* y1, const int64_t* off1, const double* y2,
                       const int64_t* off2, int n, int n_states, std::vector<int64_t>& o1, std::vector<int64_t>& o2,
                       const double** d_y1, const int64_t** d_o1, const double** d_y2, const int64_t** d_o2) {
  POB_TRY(fetch_i64(ctx, where, off1, (size_t)n + 1, o1));
  POB_TRY(fetch_i64(ctx, where, off2, (size_t)n + 1, o2));
  POB_TRY(check_offsets(o1));
  POB_TRY(check_offsets(o2));
  if (o1[0] != 0 || o2[0] != 0) return POB_EINVAL;
  for (int i = 0; i < n; ++i) {
    const int64_t U = o1[i + 1] - o1[i], V = o2[i + 1] - o2[i];
    if ((U + 1) * (V + 1) * 8 > (int64_t)1000000000) return POB_EUNSUPPORTED;  // MEM_LIMIT, pair_decode.py:189
  }
  *d_y1 = y1; *d_o1 = off1; *d_y2 = y2; *d_o2 = off2;
  if (where == POB_HOST) {
    POB_TRY(stage_in(ctx, y1, (size_t)o1[n] * n_states, d_y1));
    POB_TRY(stage_in(ctx, off1, (size_t)n + 1, d_o1));
    POB_TRY(stage_in(ctx, y2, (size_t)o2[n] * n_states, d_y2));
    POB_TRY(stage_in(ctx, off2, (size_t)n + 1, d_o2));
  }
  return POB_OK;
}

int pob_pair_gamma(pob_ctx* ctx, int where, const double* y1, const int64_t* off1, const double* y2,
                   const int64_t* off2, int n, int n_states, int flavour, const int64_t* gamma_off, double* out_gamma) {
  if (!ctx || n < 0) return POB_EINVAL;
  if (n_states < 2 || n_states > 5) return POB_EUNSUPPORTED;
  if (flavour != POB_PREFIX_NUMPY && flavour != POB_PREFIX_CY) return POB_EINVAL;
  if (n == 0) return POB_OK;
  if (!y1 || !off1 || !y2 || !off2 || !gamma_off || !out_gamma) return POB_EINVAL;
  POB_CUDA(cudaSetDevice(ctx->device));
  POB_TRY(pob_arena_reset(ctx));
  std::vector<int64_t> o1, o2, go;
  const double *d_y1, *d_y2;
  const int64_t *d_o1, *d_o2;
  POB_TRY(pair_inputs(ctx, where, y1, off1, y2, off2, n, n_states, o1, o2, &d_y1, &d_o1, &d_y2, &d_o2));
  POB_TRY(fetch_i64(ctx, where, gamma_off, (size_t)n + 1, go));
  if (go[0] != 0) return POB_EINVAL;
  for (int i = 0; i < n; ++i)
    if (go[i + 1] - go[i] != (o1[i + 1] - o1[i] + 1) * (o2[i + 1] - o2[i] + 1)) return POB_EINVAL;
  const int64_t* d_go = gamma_off;
  double* d_g = out_gamma;
  if (where == POB_HOST) {
    POB_TRY(stage_in(ctx, gamma_off, (size_t)n + 1, &d_go));
    POB_TRY(stage_out(ctx, out_gamma, (size_t)go[n], &d_g));
  }
  double* d_ga;
  POB_TRY(pob_take(ctx, (size_t)go[n] + 8, &d_ga));
  {
    pob_prof_scope prof(ctx, POB_K_PAIR_GAMMA);
    if (flavour == POB_PREFIX_CY)
      pair_gamma_kernel<POB_PREFIX_CY><<<n, PX2_THREADS, 0, ctx->stream>>>(d_y1, d_o1, d_y2, d_o2, n_states, d_g, d_ga, d_go);
    else
      pair_gamma_kernel<POB_PREFIX_NUMPY><<<n, PX2_THREADS, 0, ctx->stream>>>(d_y1, d_o1, d_y2, d_o2, n_states, d_g, d_ga, d_go);
    POB_CUDA(cudaGetLastError());
  }
  if (where == POB_HOST) {
    POB_TRY(copy_back(ctx, out_gamma, d_g, (size_t)go[n]));
    POB_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return POB_OK;
}

int pob_pair_prefix_search(pob_ctx* ctx, int where, const double* y1, const int64_t* off1, const double* y2,
                           const int64_t* off2, int n, int n_states, int flavour, const int64_t* lab_off,
                           uint8_t* out_label, int32_t* out_len, double* out_score, int32_t* out_status) {
  if (!ctx || n < 0) return POB_EINVAL;
  if (n_states < 2 || n_states > 5) return POB_EUNSUPPORTED;
  if (flavour != POB_PREFIX_NUMPY && flavour != POB_PREFIX_CY) return POB_EINVAL;
  if (n == 0) return POB_OK;
  if (!y1 || !off1 || !y2 || !off2 || !lab_off || !out_label || !out_len || !out_score) return POB_EINVAL;
  POB_CUDA(cudaSetDevice(ctx->device));
  POB_TRY(pob_arena_reset(ctx));
  std::vector<int64_t> o1, o2, loff;
  const double *d_y1, *d_y2;
  const int64_t *d_o1, *d_o2;
  POB_TRY(pair_inputs(ctx, where, y1, off1, y2, off2, n, n_states, o1, o2, &d_y1, &d_o1, &d_y2, &d_o2));
  POB_TRY(fetch_i64(ctx, where, lab_off, (size_t)n + 1, loff));
  if (loff[0] != 0) return POB_EINVAL;
  std::vector<int64_t> scr_off((size_t)n + 1);
  scr_off[0] = 0;
  for (int i = 0; i < n; ++i) {
    const int64_t U = o1[i + 1] - o1[i], V = o2[i + 1] - o2[i];
    if (loff[i + 1] - loff[i] < std::max(U, V) + 3) return POB_EINVAL;
    scr_off[i + 1] = scr_off[i] + 2 * (U + 1) * (V + 1) + 5 * (U + V);
  }
  const int64_t* d_loff = lab_off;
  uint8_t* d_lab = out_label;
  int32_t *d_len = out_len, *d_st = out_status;
  double* d_sc = out_score;
  if (where == POB_HOST) {
    POB_TRY(stage_in(ctx, lab_off, (size_t)n + 1, &d_loff));
    POB_TRY(stage_out(ctx, out_label, (size_t)loff[n], &d_lab));
    POB_TRY(stage_out(ctx, out_len, (size_t)n, &d_len));
    POB_TRY(stage_out(ctx, out_score, (size_t)n, &d_sc));
    POB_TRY(stage_out(ctx, out_status, (size_t)n, &d_st));
  }
  const int64_t* d_scr_off;
  POB_TRY(upload(ctx, scr_off, &d_scr_off));
  double* scratch;
  POB_TRY(pob_take(ctx, (size_t)scr_off[n] + 8, &scratch));
  uint8_t* work;
  POB_TRY(pob_take(ctx, (size_t)loff[n] + 8, &work));
  {
    pob_prof_scope prof(ctx, POB_K_PREFIX_2D);
    if (flavour == POB_PREFIX_CY)
      pair_prefix_kernel<POB_PREFIX_CY><<<n, PX2_THREADS, 0, ctx->stream>>>(d_y1, d_o1, d_y2, d_o2, n_states, scratch,
                                                                          d_scr_off, d_loff, work, d_lab, d_len, d_sc,
                                                                          d_st);
    else
      pair_prefix_kernel<POB_PREFIX_NUMPY><<<n, PX2_THREADS, 0, ctx->stream>>>(d_y1, d_o1, d_y2, d_o2, n_states, scratch,
                                                                             d_scr_off, d_loff, work, d_lab, d_len,
                                                                             d_sc, d_st);
    POB_CUDA(cudaGetLastError());
  }
  if (where == POB_HOST) {
    POB_TRY(copy_back(ctx, out_label, d_lab, (size_t)loff[n]));
    POB_TRY(copy_back(ctx, out_len, d_len, (size_t)n));
    POB_TRY(copy_back(ctx, out_score, d_sc, (size_t)n));
    POB_TRY(copy_back(ctx, out_status, d_st, (size_t)n));
    POB_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return POB_OK;
}

}  // extern "C"
