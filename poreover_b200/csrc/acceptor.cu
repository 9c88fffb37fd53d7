// Banded Viterbi acceptor ("re-squiggle"): best alignment of a known base sequence to a read's per-timestep
// probabilities, used by `pair-decode --single beam` to turn a beam-search basecall into a path.
//
// Reference semantics (Forward.h:14-121 over SparseMatrix.h:9-116), reproduced exactly:
//   v(0,t) = running sum of the blank column for t in [0, band];  v(1,0) = y[0][label[0]], ptr(1,0) = 1
//   for l = 1..L, c = int(l*T/L), t in [max(1, c-band), min(T, c+band)), t >= l-1:
//       emit = y[t][label[l-1]] + v(l-1,t-1);  stay = y[t][blank] + v(l,t-1);  emit >= stay -> (emit, ptr 1) else (stay, 0)
//   a cell is only STORED when t lies in the inclusive range its row was pushed with -- [0, band] for rows 0 and 1,
//   the range of iteration l-1 for row l >= 2 (the push of iteration l creates row l+1); everything else reads
//   as -inf / 0.  Traceback from (L, T-1): ptr > 0 -> emit label[l-1] at t, l--; t-- every step.
// FP64 adds and compares only, same operands in the same order: bit-exact.
//
// One CTA per read, anti-diagonal wavefront over d = l + t: cell (l,t) needs (l-1,t-1) from diagonal d-2 and
// (l,t-1) from diagonal d-1, so three rotating diagonals of values live in shared memory, indexed by l mod SZ
// (the cells of one diagonal that can exist span fewer than SZ = pow2 >= 2*band+8 labels).  Back-pointers are
// written one bit per cell, one coalesced 32-bit word per warp and diagonal (warp ballot), anti-diagonal-major,
// so the traceback addresses them in O(1).
#include "common.cuh"
#include "launch.cuh"

namespace {

constexpr int ACC_THREADS = 512;

struct AccGeom {
  int T, L, band;
  __device__ __forceinline__ int center(int l) const { return (int)(l * (double)T / (double)L); }
  __device__ __forceinline__ int rs(int l) const { return max(1, center(l) - band); }
  __device__ __forceinline__ int re(int l) const { return min(T, center(l) + band); }
  // inclusive stored range of row l (l >= 1)
  __device__ __forceinline__ bool stored(int l, int t) const {
    if (l == 1) return t >= 0 && t <= band;
    return t >= rs(l - 1) && t <= re(l - 1);
  }
  // cell computed by the main loop and kept
  __device__ __forceinline__ bool exists(int l, int t) const {
    return l >= 1 && l <= L && t >= rs(l) && t < re(l) && t >= l - 1 && stored(l, t);
  }
};

template <typename TIn>
__global__ void __launch_bounds__(ACC_THREADS)
acceptor_kernel(const TIn* __restrict__ data, const int64_t* __restrict__ row_off, const int32_t* __restrict__ row_len,
                const uint8_t* __restrict__ rcflag, int S, int layout, const uint8_t* __restrict__ labels,
                const int64_t* __restrict__ lab_off, int band, int SZ, const int64_t* __restrict__ bp_off,
                uint32_t* __restrict__ bp, double* __restrict__ cum_all, int8_t* __restrict__ out_path,
                int32_t* __restrict__ out_status) {
  extern __shared__ __align__(16) double s_val[];  // [3][SZ]
  const int r = blockIdx.x;
  const int tid = threadIdx.x;
  const int64_t ro = row_off[r];
  AccGeom g;
  g.T = pob_read_len(row_off, row_len, r);
  g.L = (int)(lab_off[r + 1] - lab_off[r]);
  g.band = band;
  const int T = g.T, L = g.L;
  const bool rc = rcflag ? rcflag[r] != 0 : false;
  const uint8_t* lab = labels + lab_off[r];
  const TIn* base = data + ro * S;
  int8_t* path = out_path + ro;
  const int blank = S - 1;
  const int cblank = pob_col(blank, S, layout, rc);
  auto yat = [&](int t, int k) -> double {
    return (double)base[(size_t)(rc ? T - 1 - t : t) * S + (k == blank ? cblank : pob_col(k, S, layout, rc))];
  };
  for (int t = tid; t < T; t += ACC_THREADS) path[t] = (int8_t)blank;
  if (T <= 0 || L <= 0) {
    if (tid == 0) out_status[r] = (T <= 0) ? POB_ST_EMPTY : 0;
    return;
  }
  // row 0: running sum of the blank column (sequential, as the reference), kept for t <= band
  double* cum = cum_all + ro;
  if (tid == 0) {
    double s = 0;
    const int lim = min(T - 1, band);
    for (int t = 0; t <= lim; ++t) { s += yat(t, blank); cum[t] = s; }
  }
  const double NINF = pob_ninf();
  for (int i = tid; i < 3 * SZ; i += ACC_THREADS) s_val[i] = NINF;
  __syncthreads();
  uint32_t* mybp = bp + bp_off[r];
  const int words = SZ >> 5;
  const int passes = SZ / ACC_THREADS > 0 ? SZ / ACC_THREADS : 1;
  const double ratio = 1.0 + (double)T / (double)L;
  // diagonal d = l + t, l >= 1, t >= 0: d in [1, L + T - 1]
  for (int d = 1; d <= L + T - 1; ++d) {
    double* cur = s_val + (d % 3) * SZ;
    const double* p1 = s_val + ((d + 2) % 3) * SZ;  // diagonal d-1
    const double* p2 = s_val + ((d + 1) % 3) * SZ;  // diagonal d-2
    // window of SZ labels around where this diagonal crosses the band
    const int mid = (int)((double)d / ratio);
    const int w0 = mid - (SZ >> 1);
    for (int k = 0; k < passes; ++k) {
      const int slot = tid + k * ACC_THREADS;
      if (slot >= SZ) break;
      // the label in [w0, w0 + SZ) whose index is congruent to slot
      int l = w0 + ((slot - w0) & (SZ - 1));
      const int t = d - l;
      double v = NINF;
      bool emit_bit = false;
      if (l == 1 && t == 0) {
        v = yat(0, lab[0]);  // Forward.h:51-53
        emit_bit = true;
      } else if (g.exists(l, t)) {
        double up;
        if (l == 1) up = (t - 1 >= 0 && t - 1 <= band) ? cum[t - 1] : NINF;
        else up = p2[(l - 1) & (SZ - 1)];
        const double emit = yat(t, lab[l - 1]) + up;
        const double stay = yat(t, blank) + p1[l & (SZ - 1)];
        if (emit >= stay) { v = emit; emit_bit = true; } else v = stay;
      }
      cur[slot] = v;
      const unsigned m = __ballot_sync(0xffffffffu, emit_bit);
      if ((tid & 31) == 0) mybp[(size_t)d * words + (slot >> 5)] = m;
    }
    __syncthreads();
  }
  // traceback (Forward.h:104-116)
  if (tid == 0) {
    int l = L, t = T - 1, st = 0;
    while (l > 0) {
      if (t < 0) { st = POB_ST_UNSET_BAND; break; }  // the reference never terminates here
      bool bit = false;
      if (l == 1 && t == 0) bit = true;
      else if (g.exists(l, t)) {
        const int d = l + t, slot = l & (SZ - 1);
        bit = (mybp[(size_t)d * words + (slot >> 5)] >> (slot & 31)) & 1u;
      }
      if (bit) { path[t] = (int8_t)lab[l - 1]; l -= 1; }
      t -= 1;
    }
    out_status[r] = st;
  }
}

}  // namespace

int pob_acceptor_slots(int band) {
  int need = 2 * band + 8, sz = 64;
  while (sz < need) sz <<= 1;
  return sz;
}

int pob_acceptor_launch(pob_ctx* ctx, const pob_reads& rd, const uint8_t* labels, const int64_t* lab_off, int band,
                        int SZ, const int64_t* bp_off, uint32_t* bp, double* cum, int8_t* out_path,
                        int32_t* out_status) {
  if (rd.n <= 0) return POB_OK;
  const size_t smem = (size_t)3 * SZ * sizeof(double);
  if (smem > 200 * 1024) return POB_EUNSUPPORTED;
  pob_prof_scope ps(ctx, POB_K_ACCEPTOR);
  // the attribute is a limit shared by all host threads of the process: always the same (largest) value, so two
  // threads launching at once cannot lower it under each other
  if (rd.dtype == POB_F32) {
    POB_CUDA(cudaFuncSetAttribute(acceptor_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    acceptor_kernel<float><<<rd.n, ACC_THREADS, smem, ctx->stream>>>((const float*)rd.data, rd.row_off, rd.row_len, rd.rc,
                                                                      rd.n_states, rd.layout, labels, lab_off, band, SZ,
                                                                      bp_off, bp, cum, out_path, out_status);
  } else {
    POB_CUDA(cudaFuncSetAttribute(acceptor_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    acceptor_kernel<double><<<rd.n, ACC_THREADS, smem, ctx->stream>>>((const double*)rd.data, rd.row_off, rd.row_len,
                                                                       rd.rc, rd.n_states, rd.layout, labels, lab_off,
                                                                       band, SZ, bp_off, bp, cum, out_path, out_status);
  }
  POB_CUDA(cudaGetLastError());
  return POB_OK;
}
