// Flip-flop Viterbi: 8-state max-sum DP in FP64 with the reference's additive 0/1 transition matrix,
// backpointer traceback, remove_repeated + upper-casing, and the flip-flop sequence mapping.
//
// Reference semantics:
//   v[0] = lp[0];  prev[j][i] = transition[i][j] + v[t-1][i];  ptr[t][j] = argmax_i (first index);
//   v[t][j] = lp[t][j] + max_i prev[j][i]                                    transducer.py:35-52
//   transition rows k and k+4 are [1 1 1 1 e_k]: the matrix is ADDED, not a mask      :94-103
//   sequence = remove_repeated(8-letter path).upper()                          transducer.py:4-9, :55
//   mapping  = every t with path[t] != path[t-1], plus t = 0                   pair_decode.py:124-133
// Only FP64 adds and compares are used, in the reference's order, so the result is bit-exact.  The 256
// possible log((x+1e-7)/(255+1e-7)) values of a uint8 trace come from a host-computed table (decode.py:92-93).
//
// Four reads per warp: the 8 lanes of a read own its 8 states and exchange v[t-1] by 8-wide shuffles; the time loop
// is a dependent chain (the kernel is bound by instruction issue, so all 32 lanes do useful work), backpointers are
// packed 8 x 3 bits per timestep and walked back 8 timesteps per coalesced load.
#include "common.cuh"
#include "launch.cuh"

namespace {

constexpr int FF_WARPS = 4;
constexpr int FF_READS_PER_WARP = 4;

template <bool U8>
__global__ void __launch_bounds__(FF_WARPS * 32)
flipflop_kernel(const void* __restrict__ data, const double* __restrict__ lut, const int64_t* __restrict__ row_off,
                const int32_t* __restrict__ row_len, const uint8_t* __restrict__ rcflag, int n,
                uint32_t* __restrict__ bp, int8_t* __restrict__ path_buf, uint8_t* __restrict__ out_seq,
                int32_t* __restrict__ out_s2s, int32_t* __restrict__ out_len) {
  __shared__ double s_lut[256];
  if (U8) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = lut[i];
    __syncthreads();
  }
  // four reads per warp: lanes 8g..8g+7 own the 8 states of read g; every shuffle below is 8 lanes wide
  const int lane = threadIdx.x & 31;
  const int grp = lane >> 3, j = lane & 7;
  const int r = (blockIdx.x * FF_WARPS + (threadIdx.x >> 5)) * FF_READS_PER_WARP + grp;
  const bool valid = r < n;
  const int64_t ro = valid ? row_off[r] : 0;
  const int T = valid ? pob_read_len(row_off, row_len, r) : 0;
  int Tmax = T;
  Tmax = max(Tmax, __shfl_xor_sync(0xffffffffu, Tmax, 8));
  Tmax = max(Tmax, __shfl_xor_sync(0xffffffffu, Tmax, 16));
  if (valid && T <= 0 && j == 0) out_len[r] = 0;
  if (Tmax <= 0) return;  // warp-uniform
  const bool rc = (valid && rcflag) ? rcflag[r] != 0 : false;
  // logical state j -> physical column (transducer.py:104-106: [3,2,1,0,7,6,5,4])
  const int pc = rc ? ((j < 4) ? 3 - j : 11 - j) : j;
  auto lp_at = [&](int t) -> double {
    const size_t idx = (size_t)(ro + (rc ? (T - 1 - t) : t)) * 8 + pc;
    if (U8) return s_lut[((const uint8_t*)data)[idx]];
    return ((const double*)data)[idx];
  };
  uint32_t* mybp = bp + ro;
  double v = (T > 0) ? lp_at(0) : 0.0;
  double nxt = (T > 1) ? lp_at(1) : 0.0;
  for (int t = 1; t < Tmax; ++t) {
    const bool act = t < T;
    const double lp = nxt;
    if (t + 1 < T) nxt = lp_at(t + 1);  // prefetch off the dependent chain
    double best = 0;
    int bi = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const double vi = __shfl_sync(0xffffffffu, v, i, 8);
      // transition[i][j]: 1 for every i when j is a flip state; for flop state j only i == j-4 and i == j
      const double tr = (j < 4 || i == j - 4 || i == j) ? 1.0 : 0.0;
      const double c = tr + vi;
      if (i == 0 || c > best) { best = c; bi = i; }
    }
    if (act) v = lp + best;
    uint32_t pk = (uint32_t)bi << (3 * j);
    pk |= __shfl_xor_sync(0xffffffffu, pk, 1, 8);
    pk |= __shfl_xor_sync(0xffffffffu, pk, 2, 8);
    pk |= __shfl_xor_sync(0xffffffffu, pk, 4, 8);
    if (j == 0 && act) mybp[t] = pk;
  }
  // argmax of the last row (first index wins), then walk the backpointers
  int state = 0;
  {
    double bv = __shfl_sync(0xffffffffu, v, 0, 8);
    for (int i = 1; i < 8; ++i) {
      const double vi = __shfl_sync(0xffffffffu, v, i, 8);
      if (vi > bv) { bv = vi; state = i; }
    }
  }
  __syncwarp();
  int8_t* path = path_buf + ro;
  if (j == 0 && T > 0) path[T - 1] = (int8_t)state;
  // walk the back-pointers 8 timesteps at a time: one coalesced load of 8 packed words per read, then the dependent
  // chain runs on shuffles (the 8 lanes of a read follow the same walk; each keeps the state of "its" timestep)
  for (int c = 0; 1 + 8 * c <= Tmax - 1; ++c) {
    const int th = T - 1 - 8 * c;   // newest timestep of this chunk for this read (may be < 1: nothing left)
    const int tm = th - j;
    const uint32_t w = (tm >= 1) ? mybp[tm] : 0u;
    int mine = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t wk = __shfl_sync(0xffffffffu, w, k, 8);
      if (th - k >= 1) state = (wk >> (3 * state)) & 7;
      if (j == k) mine = state;
    }
    if (tm >= 1) path[tm - 1] = (int8_t)mine;
  }
  __syncwarp();
  // collapse runs of identical states (A != a), upper-case, record the timestep of each emitted base
  int nout = 0, carry = -1;
  uint8_t* oseq = out_seq + ro;
  int32_t* os2s = out_s2s ? out_s2s + ro : nullptr;
  for (int t0 = 0; t0 < Tmax; t0 += 8) {
    const int t = t0 + j;
    const int p = (t < T) ? path[t] : -2;
    int prev = __shfl_up_sync(0xffffffffu, p, 1, 8);
    if (j == 0) prev = carry;
    carry = __shfl_sync(0xffffffffu, p, 7, 8);
    const bool e = (t < T) && (t == 0 || p != prev);
    const unsigned m = (__ballot_sync(0xffffffffu, e) >> (8 * grp)) & 0xffu;
    if (e) {
      const int off = nout + __popc(m & ((1u << j) - 1));
      oseq[off] = (uint8_t)("ACGT"[p & 3]);
      if (os2s) os2s[off] = t;
    }
    nout += __popc(m);
  }
  if (valid && T > 0 && j == 0) out_len[r] = nout;
}

}  // namespace

int pob_flipflop_launch(pob_ctx* ctx, const pob_reads& rd, const double* lut, uint32_t* bp, int8_t* path,
                        uint8_t* out_seq, int32_t* out_s2s, int32_t* out_len) {
  if (rd.n <= 0) return POB_OK;
  const int per_block = FF_WARPS * FF_READS_PER_WARP;
  dim3 block(FF_WARPS * 32), grid((rd.n + per_block - 1) / per_block);
  pob_prof_scope ps(ctx, POB_K_FLIPFLOP);
  if (rd.dtype == POB_U8_TRACE)
    flipflop_kernel<true><<<grid, block, 0, ctx->stream>>>(rd.data, lut, rd.row_off, rd.row_len, rd.rc, rd.n, bp, path,
                                                           out_seq, out_s2s, out_len);
  else
    flipflop_kernel<false><<<grid, block, 0, ctx->stream>>>(rd.data, lut, rd.row_off, rd.row_len, rd.rc, rd.n, bp,
                                                            path, out_seq, out_s2s, out_len);
  POB_CUDA(cudaGetLastError());
  return POB_OK;
}
