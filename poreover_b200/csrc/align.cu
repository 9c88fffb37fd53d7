// Banded Needleman-Wunsch (fill + traceback) and the alignment -> envelope construction.
//
// Bit-exact with the reference, including the behaviours that look unintended:
//   * SparseMatrix<int> default value is 0, out-of-band / missing-row reads return 0  (SparseMatrix.h:70, :110)
//   * the boundary initialisation is a no-op, rows 0..L1-1 only                        (align.pyx:112-116)
//   * column `end` of every row is never written and stays 0                           (align.pyx:127)
//   * seq[i-1] at i == 0 wraps to the last character                                   (align.pyx:129)
//   * the traceback applies EVERY tied move, re-checks nothing in between, and scores
//     with the default 2/-1 regardless of the arguments                                (align.pyx:142-162)
//
// Fill: one CTA per pair, anti-diagonal wavefront.  Three rotating anti-diagonals live in shared
// memory (slot = row & (SZ-1)); every computed cell is also streamed to global memory in
// anti-diagonal-major order (coalesced) for the recompute-style traceback.  Integer pipe + shared memory
// bound; algorithmic work = sum_i (end_i - start_i) cells.
#include <stdlib.h>

#include "common.cuh"
#include "launch.cuh"

namespace {

constexpr int NW_THREADS = 512;

struct NwPair {
  const uint8_t* s1;
  const uint8_t* s2;
  int l1, l2;
};

__device__ __forceinline__ uint8_t wrap_char(const uint8_t* s, int n, int i) { return s[i < 0 ? i + n : i]; }

// align.pyx:122-124: center = int(np.round(l2/l1*i)); np.round is round-half-even == rint
__device__ __forceinline__ void row_band(int i, int l1, int l2, int band, int& start, int& end) {
  double c = rint(((double)l2 / (double)l1) * (double)i);
  int center = (int)c;
  start = max(center - band, 0);
  end = min(center + band, l2 - 1);
}

__global__ void __launch_bounds__(NW_THREADS)
nw_fill_kernel(const uint8_t* __restrict__ seq1, const int64_t* __restrict__ off1, const int32_t* __restrict__ len1,
               const uint8_t* __restrict__ seq2, const int64_t* __restrict__ off2, const int32_t* __restrict__ len2,
               const int32_t* __restrict__ skip, int band, int match, int mismatch, int gap, int SZ,
               const int64_t* __restrict__ m_off, int32_t* __restrict__ M, const int64_t* __restrict__ rb_off,
               int32_t* __restrict__ rowband) {
  extern __shared__ int32_t sm[];
  const int p = blockIdx.x;
  if (skip && skip[p]) return;
  const int l1 = len1 ? len1[p] : (int)(off1[p + 1] - off1[p]);
  const int l2 = len2 ? len2[p] : (int)(off2[p + 1] - off2[p]);
  if (l1 <= 0) return;
  const uint8_t* s1 = seq1 + off1[p];
  const uint8_t* s2 = seq2 + off2[p];
  int32_t* Mp = M + m_off[p];
  int32_t* rs = rowband + rb_off[p];  // [l1] start, then [l1] end
  int32_t* re = rs + l1;
  const int mask = SZ - 1;
  for (int i = threadIdx.x; i < l1; i += NW_THREADS) {
    int s, e;
    row_band(i, l1, l2, band, s, e);
    rs[i] = s;
    re[i] = e;
  }
  for (int k = threadIdx.x; k < 3 * SZ; k += NW_THREADS) sm[k] = 0;
  __syncthreads();
  if (l2 <= 0) return;
  int ilo = 0, ihi = -1;
  const int D = l1 + l2 - 1;
  for (int d = 0; d < D; ++d) {
    while (ihi + 1 < l1 && (ihi + 1) + rs[ihi + 1] <= d) ++ihi;
    while (ilo < l1 && ilo + re[ilo] <= d) ++ilo;
    int32_t* cur = sm + (d % 3) * SZ;
    const int32_t* p1 = sm + ((d + 2) % 3) * SZ;  // d-1
    const int32_t* p2 = sm + ((d + 1) % 3) * SZ;  // d-2
    for (int i = ilo + threadIdx.x; i <= ihi; i += NW_THREADS) {
      const int j = d - i;
      const int si = rs[i], ei = re[i];
      if (j < si || j >= ei) continue;
      int sp = 0, ep = -1;
      if (i > 0) { sp = rs[i - 1]; ep = re[i - 1]; }
      const int sc = (wrap_char(s1, l1, i - 1) == wrap_char(s2, l2, j - 1)) ? match : mismatch;
      const int dg = (i > 0 && j - 1 >= sp && j - 1 < ep) ? p2[(i - 1) & mask] : 0;
      const int up = (i > 0 && j >= sp && j < ep) ? p1[(i - 1) & mask] : 0;
      const int lf = (j - 1 >= si) ? p1[i & mask] : 0;
      int v = max(max(dg + sc, up + gap), lf + gap);
      cur[i & mask] = v;
      Mp[(size_t)d * SZ + (i & mask)] = v;
    }
    __syncthreads();
  }
}

// Fill, row-owner variant (the one that runs whenever SZ <= 1024, i.e. band <= 511): thread t owns the rows
// i = t, t + NT, t + 2 NT, ... (NT == SZ).  The active rows of an anti-diagonal are a contiguous range of fewer than
// SZ rows (i + start_i and i + end_i both increase strictly with i and a row is active for at most 2 band + 1
// diagonals), so a thread has at most one active row per diagonal and stays on a row for its whole band:
//   * the band of rows i and i-1 and the character of row i live in registers (recomputed when the thread moves on),
//   * the left neighbour M(i, j-1) is the thread's own previous value, the diagonal neighbour M(i-1, j-1) is the
//     `up` value it read one diagonal earlier,
// which leaves one shared-memory read (M(i-1, j) from the row above), one character of read 2, one shared and one
// global store per cell, and no per-diagonal search for the active range.  Same cells, same values, same layout of M
// as nw_fill_kernel below (kept for wider bands).
template <int NT>
__global__ void __launch_bounds__(NT, 2048 / NT)
nw_fill_rows_kernel(const uint8_t* __restrict__ seq1, const int64_t* __restrict__ off1,
                    const int32_t* __restrict__ len1, const uint8_t* __restrict__ seq2,
                    const int64_t* __restrict__ off2, const int32_t* __restrict__ len2,
                    const int32_t* __restrict__ skip, int band, int match, int mismatch, int gap,
                    const int64_t* __restrict__ m_off, int32_t* __restrict__ M, const int64_t* __restrict__ rb_off,
                    int32_t* __restrict__ rowband) {
  extern __shared__ int32_t sm[];
  constexpr int SZ = NT, mask = NT - 1;
  const int p = blockIdx.x;
  if (skip && skip[p]) return;
  const int l1 = len1 ? len1[p] : (int)(off1[p + 1] - off1[p]);
  const int l2 = len2 ? len2[p] : (int)(off2[p + 1] - off2[p]);
  if (l1 <= 0) return;
  const uint8_t* s1 = seq1 + off1[p];
  const uint8_t* s2 = seq2 + off2[p];
  int32_t* Mp = M + m_off[p];
  int32_t* rs = rowband + rb_off[p];  // [l1] start, then [l1] end: read by the traceback
  int32_t* re = rs + l1;
  const int t = threadIdx.x;
  for (int i = t; i < l1; i += NT) {
    int s, e;
    row_band(i, l1, l2, band, s, e);
    rs[i] = s;
    re[i] = e;
  }
  for (int k = t; k < 3 * SZ; k += NT) sm[k] = 0;
  __syncthreads();
  if (l2 <= 0) return;
  // state of the row this thread is on
  int i = t, si = 0, ei = -1, sp = 0, ep = -1;
  uint8_t ch = 0;
  bool fresh = true;        // no `up` value of the previous diagonal yet (first diagonal on this row)
  int dg_next = 0, v_prev = 0;
  if (i < l1) {
    row_band(i, l1, l2, band, si, ei);
    if (i > 0) row_band(i - 1, l1, l2, band, sp, ep);
    ch = wrap_char(s1, l1, i - 1);
  }
  const int D = l1 + l2 - 1;
  const int slot_up = (t + NT - 1) & mask;  // (i & mask) == t for every row this thread owns
  int32_t* cur = sm;                        // diagonal d; p1 = d-1, p2 = d-2 (rotated at the end of every iteration)
  const int32_t* p1 = sm + 2 * SZ;
  const int32_t* p2 = sm + SZ;
  int32_t* mrow = Mp + t;                   // &M[d][t]
  for (int d = 0; d < D; ++d, mrow += SZ) {
    if (i < l1) {
      int j = d - i;
      if (j >= ei) {
        // row finished (or empty and passed): move on to the next row of this thread
        do {
          i += NT;
          if (i < l1) {
            row_band(i, l1, l2, band, si, ei);
            row_band(i - 1, l1, l2, band, sp, ep);
            ch = wrap_char(s1, l1, i - 1);
          }
        } while (i < l1 && d - i >= ei);
        fresh = true;
        j = d - i;
      }
      if (i < l1 && j >= si - 1) {
        // M(i-1, j): read from the diagonal before the row's first cell on, it is the next diagonal's M(i-1, j-1)
        const int up = (i > 0 && j >= sp && j < ep) ? p1[slot_up] : 0;
        if (j >= si) {
          int dg;
          if (fresh) dg = (i > 0 && j - 1 >= sp && j - 1 < ep) ? p2[slot_up] : 0;
          else dg = dg_next;
          const int sc = (ch == wrap_char(s2, l2, j - 1)) ? match : mismatch;
          const int lf = (j - 1 >= si) ? v_prev : 0;
          const int v = max(max(dg + sc, up + gap), lf + gap);
          cur[t] = v;
          *mrow = v;
          v_prev = v;
        }
        dg_next = up;
        fresh = false;
      }
    }
    __syncthreads();
    int32_t* const nxt = const_cast<int32_t*>(p2);
    p2 = p1; p1 = cur; cur = nxt;
  }
}

__device__ __forceinline__ int m_get(const int32_t* Mp, const int32_t* rs, const int32_t* re, int l1, int SZ, int i,
                                     int j) {
  if (i < 0 || i >= l1) return 0;
  if (j < rs[i] || j >= re[i]) return 0;  // column `end` itself is in range of get() but holds 0
  return Mp[(size_t)(i + j) * SZ + (i & (SZ - 1))];
}

// One warp per pair: lane 0 walks the path (writing the gapped rows backwards from the end of the
// pair's output slot), then the warp moves them to the front and counts equal columns.
__global__ void __launch_bounds__(128)
nw_traceback_kernel(const uint8_t* __restrict__ seq1, const int64_t* __restrict__ off1,
                    const int32_t* __restrict__ len1, const uint8_t* __restrict__ seq2,
                    const int64_t* __restrict__ off2, const int32_t* __restrict__ len2,
                    const int32_t* __restrict__ skip, int n, int gap, int SZ, const int64_t* __restrict__ m_off,
                    const int32_t* __restrict__ M, const int64_t* __restrict__ rb_off,
                    const int32_t* __restrict__ rowband, const int64_t* __restrict__ aln_off,
                    uint8_t* __restrict__ out_a1, uint8_t* __restrict__ out_a2, int32_t* __restrict__ out_alen,
                    int32_t* __restrict__ out_matches) {
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= n) return;
  if (skip && skip[p]) {
    if (lane == 0) { out_alen[p] = 0; if (out_matches) out_matches[p] = 0; }
    return;
  }
  const int l1 = len1 ? len1[p] : (int)(off1[p + 1] - off1[p]);
  const int l2 = len2 ? len2[p] : (int)(off2[p + 1] - off2[p]);
  if (l1 <= 0) {  // ZeroDivisionError in the reference (align.pyx:122)
    if (lane == 0) { out_alen[p] = -1; if (out_matches) out_matches[p] = 0; }
    return;
  }
  const uint8_t* s1 = seq1 + off1[p];
  const uint8_t* s2 = seq2 + off2[p];
  const int32_t* Mp = M + m_off[p];
  const int32_t* rs = rowband + rb_off[p];
  const int32_t* re = rs + l1;
  uint8_t* a1 = out_a1 + aln_off[p];
  uint8_t* a2 = out_a2 + aln_off[p];
  const int cap = (int)(aln_off[p + 1] - aln_off[p]);
  int n_col = 0;
  if (lane == 0) {
    int i = l1, j = l2, w = cap;  // write position moves down from cap
    while (i > 0 && j > 0) {
      const int sc = (wrap_char(s1, l1, i - 1) == wrap_char(s2, l2, j - 1)) ? 2 : -1;
      const int c0 = m_get(Mp, rs, re, l1, SZ, i - 1, j - 1) + sc;
      const int c1 = m_get(Mp, rs, re, l1, SZ, i - 1, j) + gap;
      const int c2 = m_get(Mp, rs, re, l1, SZ, i, j - 1) + gap;
      const int mx = max(c0, max(c1, c2));
      if (c0 == mx) { --i; --j; --w; a1[w] = wrap_char(s1, l1, i); a2[w] = wrap_char(s2, l2, j); }
      if (c1 == mx) { --i; --w; a1[w] = wrap_char(s1, l1, i); a2[w] = '-'; }
      if (c2 == mx) { --j; --w; a1[w] = '-'; a2[w] = wrap_char(s2, l2, j); }
    }
    while (i > 0 || j > 0) {
      if (i > 0) { --i; --w; a1[w] = wrap_char(s1, l1, i); a2[w] = '-'; }
      else { --j; --w; a1[w] = '-'; a2[w] = wrap_char(s2, l2, j); }
    }
    n_col = cap - w;
  }
  n_col = __shfl_sync(0xffffffffu, n_col, 0);
  __syncwarp();
  const int src = cap - n_col;
  int matches = 0;
  for (int c0 = 0; c0 < n_col; c0 += 32) {  // ascending chunks: safe forward move of overlapping ranges
    int c = c0 + lane;
    uint8_t x = 0, y = 1;
    if (c < n_col) { x = a1[src + c]; y = a2[src + c]; }
    __syncwarp();
    if (c < n_col) {
      a1[c] = x; a2[c] = y;
      matches += (x == y);
    }
    __syncwarp();
  }
  matches = __reduce_add_sync(0xffffffffu, matches);
  if (lane == 0) {
    out_alen[p] = n_col;
    if (out_matches) out_matches[p] = matches;
  }
}

// ---------------------------------------------------------------------------------------------
// envelope.py:26-44 (columns) + :46-87 (blocks, padding, repair), one CTA per pair, plus the
// per-column transpose that beam_search_2d_by_row_col builds (BeamSearch.h:269-284).
constexpr int ENV_THREADS = 256;

__global__ void __launch_bounds__(ENV_THREADS)
envelope_kernel(const uint8_t* __restrict__ a1, const uint8_t* __restrict__ a2, const int64_t* __restrict__ aln_off,
                const int32_t* __restrict__ alen, const int32_t* __restrict__ s2s1,
                const int64_t* __restrict__ soff1, const int32_t* __restrict__ slen1,
                const int32_t* __restrict__ s2s2, const int64_t* __restrict__ soff2,
                const int32_t* __restrict__ slen2, const int32_t* __restrict__ Uarr,
                const int32_t* __restrict__ Varr, const int64_t* __restrict__ env_off,
                const int32_t* __restrict__ skip, int padding, int32_t* __restrict__ env) {
  __shared__ int warp_x[ENV_THREADS / 32], warp_y[ENV_THREADS / 32];
  __shared__ int run_x, run_y;
  const int p = blockIdx.x;
  if (skip && skip[p]) return;
  const int C = alen[p];
  const int U = Uarr[p], V = Varr[p];
  int32_t* e = env + 2 * env_off[p];
  const int L1 = slen1[p], L2 = slen2[p];
  const int32_t* m1 = s2s1 + soff1[p];
  const int32_t* m2 = s2s2 + soff2[p];
  const uint8_t* r1 = a1 + aln_off[p];
  const uint8_t* r2 = a2 + aln_off[p];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int i = tid; i < U; i += ENV_THREADS) { e[2 * i] = 0x7fffffff; e[2 * i + 1] = -1; }
  if (tid == 0) { run_x = -1; run_y = -1; }
  __syncthreads();
  if (C > 0 && L1 > 0 && L2 > 0) {
    for (int c0 = 0; c0 < C; c0 += ENV_THREADS) {
      const int c = c0 + tid;
      int fx = 0, fy = 0;
      if (c < C) { fx = r1[c] != '-'; fy = r2[c] != '-'; }
      // block inclusive scan of the non-gap flags
      int sx = fx, sy = fy;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int tx = __shfl_up_sync(0xffffffffu, sx, d), ty = __shfl_up_sync(0xffffffffu, sy, d);
        if (lane >= d) { sx += tx; sy += ty; }
      }
      if (lane == 31) { warp_x[wid] = sx; warp_y[wid] = sy; }
      __syncthreads();
      int bx = run_x, by = run_y;
      for (int w = 0; w < wid; ++w) { bx += warp_x[w]; by += warp_y[w]; }
      const int xi = bx + sx, yi = by + sy;  // index of the current base in each read (envelope.py:31-34)
      if (c < C) {
        const int a = min(max(xi, 0), L1 - 1), b = min(max(yi, 0), L2 - 1);
        const int bsx = m1[a], bex = (a + 1 < L1) ? m1[a + 1] : U;
        const int bsy = m2[b], bey = (b + 1 < L2) ? m2[b + 1] : V;
        for (int i = bsx; i < bex && i < U; ++i) {  // add_block (envelope.py:5-17)
          atomicMin(&e[2 * i], bsy);
          atomicMax(&e[2 * i + 1], bey);
        }
      }
      __syncthreads();
      if (tid == ENV_THREADS - 1) { run_x = xi; run_y = yi; }
      __syncthreads();
    }
  }
  __syncthreads();
  // padding (envelope.py:73-75); untouched rows hold -1/-1
  for (int i = tid; i < U; i += ENV_THREADS) {
    int lo = e[2 * i], hi = e[2 * i + 1];
    if (lo == 0x7fffffff) lo = -1;
    lo = max(0, lo - padding);
    hi = min(V, hi + padding);
    if (lo > hi) lo = 0;  // first half of the repair pass (envelope.py:79-80), independent per row
    e[2 * i] = lo;
    e[2 * i + 1] = hi;
  }
  __syncthreads();
  // second half of the repair pass: prev_end only moves when a row is clamped (envelope.py:82-85)
  if (wid == 0) {
    int prev_end = 0;
    for (int i0 = 0; i0 < U; i0 += 32) {
      const int i = i0 + lane;
      int lo = 0, hi = 0;
      if (i < U) { lo = e[2 * i]; hi = e[2 * i + 1]; }
      const int nk = min(32, U - i0);
      bool changed = false;
      for (int k = 0; k < nk; ++k) {
        const int lk = __shfl_sync(0xffffffffu, lo, k), hk = __shfl_sync(0xffffffffu, hi, k);
        if (lk > prev_end) {
          if (lane == k) { lo = prev_end; changed = true; }
          prev_end = hk;
        }
      }
      if (changed) e[2 * i] = lo;
    }
  }
}

// BeamSearch.h:269-284: per column x of read 2, (first row whose range contains x, that row + number of
// rows containing x); (-1,-1) when no row does.  Also reports the widest row / column span of the pair.
__global__ void __launch_bounds__(ENV_THREADS)
envelope_transpose_kernel(const int32_t* __restrict__ env, const int64_t* __restrict__ env_off,
                          const int32_t* __restrict__ Uarr, const int32_t* __restrict__ Varr,
                          const int64_t* __restrict__ envt_off, const int32_t* __restrict__ skip,
                          int32_t* __restrict__ envt, int32_t* __restrict__ span) {
  const int p = blockIdx.x;
  if (skip && skip[p]) return;
  const int U = Uarr[p], V = Varr[p];
  const int32_t* e = env + 2 * env_off[p];
  int32_t* t = envt + 2 * envt_off[p];
  const int tid = threadIdx.x;
  for (int x = tid; x < V; x += ENV_THREADS) { t[2 * x] = 0x7fffffff; t[2 * x + 1] = 0; }
  __syncthreads();
  int wmax = 0;
  for (int u = tid; u < U; u += ENV_THREADS) {
    const int lo = e[2 * u], hi = e[2 * u + 1];
    wmax = max(wmax, hi - lo);
    for (int x = max(lo, 0); x < hi && x < V; ++x) {
      atomicMin(&t[2 * x], u);
      atomicAdd(&t[2 * x + 1], 1);
    }
  }
  __syncthreads();
  int cmax = 0;
  for (int x = tid; x < V; x += ENV_THREADS) {
    const int first = t[2 * x], cnt = t[2 * x + 1];
    cmax = max(cmax, cnt);
    if (cnt == 0) { t[2 * x] = -1; t[2 * x + 1] = -1; }
    else t[2 * x + 1] = first + cnt;
  }
  if (span) {
    atomicMax(&span[2 * p], wmax);
    atomicMax(&span[2 * p + 1], cmax);
  }
}

}  // namespace

int pob_nw_smem_bytes(int SZ) { return 3 * SZ * (int)sizeof(int32_t); }

int pob_nw_slots(int band) {
  int need = 2 * band + 2, sz = 64;
  while (sz < need) sz <<= 1;
  return sz;
}

int pob_nw_launch(pob_ctx* ctx, const uint8_t* seq1, const int64_t* off1, const int32_t* len1, const uint8_t* seq2,
                  const int64_t* off2, const int32_t* len2, const int32_t* skip, int n, int band, int match,
                  int mismatch, int gap, int SZ, const int64_t* m_off, int32_t* M, const int64_t* rb_off,
                  int32_t* rowband, const int64_t* aln_off, uint8_t* out_a1, uint8_t* out_a2, int32_t* out_alen,
                  int32_t* out_matches) {
  if (n <= 0) return POB_OK;
  const int smem = pob_nw_smem_bytes(SZ);
  if (smem > 200 * 1024) return POB_EUNSUPPORTED;
  // a limit shared by all host threads of the process: always the same (largest) value (see acceptor.cu)
  POB_CUDA(cudaFuncSetAttribute(nw_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  {
    pob_prof_scope ps(ctx, POB_K_NW_FILL);
    static const bool generic = getenv("POB_DEBUG_NW_GENERIC") != nullptr;
#define POB_NW_ROWS(NT)                                                                                          \
  nw_fill_rows_kernel<NT><<<n, NT, smem, ctx->stream>>>(seq1, off1, len1, seq2, off2, len2, skip, band, match, \
                                                        mismatch, gap, m_off, M, rb_off, rowband)
    if (generic || SZ > 1024) {
      nw_fill_kernel<<<n, NW_THREADS, smem, ctx->stream>>>(seq1, off1, len1, seq2, off2, len2, skip, band, match,
                                                           mismatch, gap, SZ, m_off, M, rb_off, rowband);
    } else if (SZ == 1024) POB_NW_ROWS(1024);
    else if (SZ == 512) POB_NW_ROWS(512);
    else if (SZ == 256) POB_NW_ROWS(256);
    else if (SZ == 128) POB_NW_ROWS(128);
    else if (SZ == 64) POB_NW_ROWS(64);
    else return POB_EINVAL;
#undef POB_NW_ROWS
  }
  POB_CUDA(cudaGetLastError());
  {
    pob_prof_scope ps(ctx, POB_K_NW_TRACE);
    nw_traceback_kernel<<<(n + 3) / 4, 128, 0, ctx->stream>>>(seq1, off1, len1, seq2, off2, len2, skip, n, gap, SZ,
                                                              m_off, M, rb_off, rowband, aln_off, out_a1, out_a2,
                                                              out_alen, out_matches);
  }
  POB_CUDA(cudaGetLastError());
  return POB_OK;
}

int pob_envelope_launch(pob_ctx* ctx, const uint8_t* a1, const uint8_t* a2, const int64_t* aln_off,
                        const int32_t* alen, const int32_t* s2s1, const int64_t* soff1, const int32_t* slen1,
                        const int32_t* s2s2, const int64_t* soff2, const int32_t* slen2, const int32_t* U,
                        const int32_t* V, const int64_t* env_off, const int32_t* skip, int n, int padding,
                        int32_t* env) {
  if (n <= 0) return POB_OK;
  pob_prof_scope ps(ctx, POB_K_ENVELOPE);
  envelope_kernel<<<n, ENV_THREADS, 0, ctx->stream>>>(a1, a2, aln_off, alen, s2s1, soff1, slen1, s2s2, soff2, slen2,
                                                      U, V, env_off, skip, padding, env);
  POB_CUDA(cudaGetLastError());
  return POB_OK;
}

int pob_envelope_transpose_launch(pob_ctx* ctx, const int32_t* env, const int64_t* env_off, const int32_t* U,
                                  const int32_t* V, const int64_t* envt_off, const int32_t* skip, int n,
                                  int32_t* envt, int32_t* span) {
  if (n <= 0) return POB_OK;
  pob_prof_scope ps(ctx, POB_K_ENVELOPE);
  envelope_transpose_kernel<<<n, ENV_THREADS, 0, ctx->stream>>>(env, env_off, U, V, envt_off, skip, envt, span);
  POB_CUDA(cudaGetLastError());
  return POB_OK;
}
