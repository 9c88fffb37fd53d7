// Best-path decode of CTC reads (poreover / bonito kinds), fused with collapse, blank removal,
// warp-level stream compaction and the base -> timestep mapping.
//
// Reference semantics:
//   path[t] = first-index argmax over logical columns (A C G T blank)       transducer.py:27-33
//   poreover: bases = path[t] != blank, repeats kept                          transducer.py:72-73
//   bonito  : itertools.groupby collapse, then blanks dropped                 transducer.py:83-89
//   mapping : timestep of each emitted base                                   pair_decode.py:114-142
// Compare-only arithmetic, so results are bit-exact by construction.
//
// Fast path: one warp per read, float32, 5 states.  A warp consumes the read in chunks of 128 rows = 2560
// contiguous bytes.  Chunks are staged with cp.async (16-byte pieces, lane i copies bytes [16 i + 512 k), fully
// coalesced: every 32-byte sector is requested exactly once) into a per-warp ring of STAGES chunk buffers in
// shared memory, two chunks ahead of the one being reduced; a lane then reads "its" 4 consecutive rows (20
// floats, five conflict-free LDS.128 at an 80-byte lane stride).  HBM-bound: 20 B in per timestep, ~2.4 B out.
#include "common.cuh"

namespace {

constexpr int WARPS_PER_BLOCK = 4;

__device__ __forceinline__ int warp_excl_scan(int v, int lane, int& total) {
  int x = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= d) x += y;
  }
  total = __shfl_sync(0xffffffffu, x, 31);
  return x - v;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global.L2::128B [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int STAGES = 3;            // chunk buffers per warp
constexpr int CHUNK_FLOATS = 640;    // 32 groups x 4 rows x 5 states

// first-index argmax of one row held in registers; cols are compile-time after inlining
template <int LAYOUT>
__device__ __forceinline__ int argmax5(const float* r, bool rc) {
  // logical k -> physical column
  float v0, v1, v2, v3, v4;
  if (LAYOUT == POB_BLANK_LAST) {
    v0 = rc ? r[3] : r[0];
    v1 = rc ? r[2] : r[1];
    v2 = rc ? r[1] : r[2];
    v3 = rc ? r[0] : r[3];
    v4 = r[4];
  } else {
    v0 = rc ? r[4] : r[1];
    v1 = rc ? r[3] : r[2];
    v2 = rc ? r[2] : r[3];
    v3 = rc ? r[1] : r[4];
    v4 = r[0];
  }
  int b = 0;
  float m = v0;
  if (v1 > m) { m = v1; b = 1; }
  if (v2 > m) { m = v2; b = 2; }
  if (v3 > m) { m = v3; b = 3; }
  if (v4 > m) { m = v4; b = 4; }
  return b;
}

// One read, one warp.  RC / KIND / PATH are compile-time so the per-row work is a handful of compares and
// selects; chunks whose 32 groups are all complete take a path without per-row validity tests.
template <int LAYOUT, bool RC, int KIND, bool PATH>
__device__ __forceinline__ void viterbi5_read(const float* __restrict__ base, int T, int lane, float* __restrict__ ring,
                                              uint8_t* __restrict__ oseq, int32_t* __restrict__ os2s,
                                              int8_t* __restrict__ opath, int& nout_o, int& pfirst_o, int& plast_o) {
  const int G = (T + 3) >> 2;          // groups of 4 rows
  const int Gfull = T >> 2;            // complete groups
  const int nch = (G + 31) >> 5;
  const bool aligned = (reinterpret_cast<uintptr_t>(base) & 15) == 0;
  float cur[20];
  // Stage chunk c (logical groups [32c, 32c+32), a contiguous span of physical groups) into ring slot c % STAGES.
  // Only bytes of this read are touched: whole 16-byte pieces, then the last 1-3 floats of a partial group.
  auto stage_chunk = [&](int c) {
    if (c < nch) {
      const int ng = min(32, G - (c << 5));                         // groups in this chunk
      const int gfirst = RC ? (G - (c << 5) - ng) : (c << 5);       // lowest physical group
      const float* src = base + (size_t)gfirst * 20;
      float* dst = ring + (c % STAGES) * CHUNK_FLOATS;
      const int nf = min(ng * 20, T * 5 - gfirst * 20);             // floats of this read in the span
      if (aligned) {
        const int npieces = nf >> 2;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const int i = lane + 32 * k;
          if (i < npieces) cp_async16(dst + 4 * i, src + 4 * i);
        }
        const int i = 4 * npieces + lane;
        if (i < nf) cp_async4(dst + i, src + i);
      } else {
        for (int i = lane; i < nf; i += 32) cp_async4(dst + i, src + i);
      }
    }
    cp_async_commit();  // one group per call (possibly empty) keeps the wait arithmetic uniform
  };
  int carry = -1, nout = 0, pfirst = -1, plast = -1;
#pragma unroll
  for (int c = 0; c < STAGES - 1; ++c) stage_chunk(c);
  for (int c = 0; c < nch; ++c) {
    stage_chunk(c + STAGES - 1);
    cp_async_wait<STAGES - 1>();  // chunk c has landed (this lane's copies; the other lanes' after the warp sync)
    __syncwarp();
    const int gi = (c << 5) + lane;
    const int g = RC ? (G - 1 - gi) : gi;
    {
      // the lane's group inside the staged span
      const int ng = min(32, G - (c << 5));
      const int gl = RC ? (ng - 1 - lane) : lane;
      if (lane < ng) {
        const float4* q = reinterpret_cast<const float4*>(ring + (c % STAGES) * CHUNK_FLOATS + gl * 20);
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          const float4 v = q[i];
          cur[4 * i] = v.x; cur[4 * i + 1] = v.y; cur[4 * i + 2] = v.z; cur[4 * i + 3] = v.w;
        }
      }
    }
    __syncwarp();  // every lane has read slot c % STAGES before a later iteration refills it
    // the chunk is "full" when all of its 32 groups are complete groups of the read
    const bool full = RC ? ((c << 5) + 31 < G && (G - 1 - (c << 5)) < Gfull) : ((c << 5) + 32 <= Gfull);
    int p[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) p[j] = argmax5<LAYOUT>(cur + 5 * (RC ? 3 - j : j), RC);  // logical order
    const int t0 = RC ? (T - 4 * g - 4) : 4 * g;  // logical time of the lane's first row (full groups)
    unsigned em = 0;
    if (full) {
      int prev = __shfl_up_sync(0xffffffffu, p[3], 1);
      if (lane == 0) prev = carry;
      carry = __shfl_sync(0xffffffffu, p[3], 31);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool e = (p[j] != 4) && (KIND == POB_KIND_POREOVER || p[j] != prev);
        em |= (e ? 1u : 0u) << j;
        prev = p[j];
      }
      if (PATH) {
        // four consecutive logical timesteps: one 32-bit store when the destination is 4-byte aligned
        const unsigned pk = (unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16) | ((unsigned)p[3] << 24);
        if (((reinterpret_cast<uintptr_t>(opath) + t0) & 3) == 0) *reinterpret_cast<unsigned*>(opath + t0) = pk;
        else { opath[t0] = p[0]; opath[t0 + 1] = p[1]; opath[t0 + 2] = p[2]; opath[t0 + 3] = p[3]; }
      }
      if (c == 0 && lane == 0) pfirst = p[0];
      if (t0 + 4 == T) plast = p[3];
    } else {
      const bool have = gi < G;
      int last = -1;
      bool valid[4];
      int tt[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int pr = RC ? (4 * g + 3 - j) : (4 * g + j);
        valid[j] = have && pr < T;
        tt[j] = RC ? (T - 1 - pr) : pr;
        if (valid[j]) last = p[j];
      }
      int prev = __shfl_up_sync(0xffffffffu, last, 1);
      if (lane == 0) prev = carry;
      carry = __shfl_sync(0xffffffffu, last, 31);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (valid[j]) {
          const bool e = (p[j] != 4) && (KIND == POB_KIND_POREOVER || p[j] != prev);
          em |= (e ? 1u : 0u) << j;
          prev = p[j];
          if (tt[j] == 0) pfirst = p[j];
          if (tt[j] == T - 1) plast = p[j];
          if (PATH) opath[tt[j]] = (int8_t)p[j];
        }
      }
    }
    int total;
    int off = nout + warp_excl_scan(__popc(em), lane, total);
    if (full) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (em & (1u << j)) {
          oseq[off] = (uint8_t)("ACGT"[p[j]]);
          if (os2s) os2s[off] = t0 + j;
          ++off;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (em & (1u << j)) {
          const int pr = RC ? (4 * g + 3 - j) : (4 * g + j);
          oseq[off] = (uint8_t)("ACGT"[p[j]]);
          if (os2s) os2s[off] = RC ? (T - 1 - pr) : pr;
          ++off;
        }
      }
    }
    nout += total;
  }
  cp_async_wait<0>();
  nout_o = nout;
  pfirst_o = __reduce_max_sync(0xffffffffu, pfirst);
  plast_o = __reduce_max_sync(0xffffffffu, plast);
}

template <int LAYOUT, int KIND, bool PATH>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
viterbi5_f32_kernel(const float* __restrict__ data, const int64_t* __restrict__ row_off,
                    const int32_t* __restrict__ row_len, const uint8_t* __restrict__ rcflag, int n,
                    uint8_t* __restrict__ out_seq, int32_t* __restrict__ out_s2s, int8_t* __restrict__ out_path,
                    int32_t* __restrict__ out_len, int32_t* __restrict__ out_status) {
  __shared__ __align__(16) float s_ring[WARPS_PER_BLOCK][STAGES * CHUNK_FLOATS];
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (r >= n) return;
  const int64_t ro = row_off[r];
  const int T = pob_read_len(row_off, row_len, r);
  const bool rc = rcflag ? (rcflag[r] != 0) : false;
  const float* base = data + ro * 5;
  uint8_t* oseq = out_seq + ro;
  int32_t* os2s = out_s2s ? out_s2s + ro : nullptr;
  int8_t* opath = PATH ? out_path + ro : nullptr;
  float* ring = s_ring[threadIdx.x >> 5];
  int nout, pfirst, plast;
  if (rc) viterbi5_read<LAYOUT, true, KIND, PATH>(base, T, lane, ring, oseq, os2s, opath, nout, pfirst, plast);
  else viterbi5_read<LAYOUT, false, KIND, PATH>(base, T, lane, ring, oseq, os2s, opath, nout, pfirst, plast);
  if (lane == 0) {
    out_len[r] = nout;
    int st = (T == 0) ? POB_ST_EMPTY : 0;
    // pair_decode.py:136 compares path[0] with path[-1]; equal bases drop the first one
    if (KIND == POB_KIND_BONITO && T > 0 && pfirst != 4 && pfirst == plast) st |= POB_ST_MAPPING_WRAP;
    if (out_status) out_status[r] = st;
  }
}

// Generic path: any dtype, 2 <= S <= 9, one row per lane.
template <typename TIn>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
viterbi_generic_kernel(const TIn* __restrict__ data, const int64_t* __restrict__ row_off,
                       const int32_t* __restrict__ row_len, const uint8_t* __restrict__ rcflag, int n, int S, int layout, int kind,
                       uint8_t* __restrict__ out_seq, int32_t* __restrict__ out_s2s, int8_t* __restrict__ out_path,
                       int32_t* __restrict__ out_len, int32_t* __restrict__ out_status) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (r >= n) return;
  const int64_t ro = row_off[r];
  const int T = pob_read_len(row_off, row_len, r);
  const bool rc = rcflag ? (rcflag[r] != 0) : false;
  const TIn* base = data + ro * S;
  const int blank = S - 1;
  int carry = -1, nout = 0, pfirst = -1, plast = -1;
  uint8_t* oseq = out_seq + ro;
  int32_t* os2s = out_s2s ? out_s2s + ro : nullptr;
  int8_t* opath = out_path ? out_path + ro : nullptr;
  for (int t0 = 0; t0 < T; t0 += 32) {
    int t = t0 + lane;
    bool valid = t < T;
    int p = -1;
    if (valid) {
      const TIn* row = base + (size_t)(rc ? (T - 1 - t) : t) * S;
      TIn m = row[pob_col(0, S, layout, rc)];
      p = 0;
      for (int k = 1; k < S; ++k) {
        TIn v = row[pob_col(k, S, layout, rc)];
        if (v > m) { m = v; p = k; }
      }
      if (opath) opath[t] = (int8_t)p;
      if (t == 0) pfirst = p;
      if (t == T - 1) plast = p;
    }
    int prev = __shfl_up_sync(0xffffffffu, p, 1);
    if (lane == 0) prev = carry;
    carry = __shfl_sync(0xffffffffu, p, 31);
    bool e = valid && p != blank && (kind == POB_KIND_POREOVER || p != prev);
    unsigned m = __ballot_sync(0xffffffffu, e);
    if (e) {
      int off = nout + __popc(m & ((1u << lane) - 1));
      oseq[off] = (uint8_t)("ACGTNNNN"[p]);
      if (os2s) os2s[off] = t;
    }
    nout += __popc(m);
  }
  pfirst = __reduce_max_sync(0xffffffffu, pfirst);
  plast = __reduce_max_sync(0xffffffffu, plast);
  if (lane == 0) {
    out_len[r] = nout;
    int st = (T == 0) ? POB_ST_EMPTY : 0;
    if (kind == POB_KIND_BONITO && T > 0 && pfirst != blank && pfirst == plast) st |= POB_ST_MAPPING_WRAP;
    if (out_status) out_status[r] = st;
  }
}

}  // namespace

// Device-side entry used by the host API and by the fused pair pipeline.
int pob_viterbi_launch(pob_ctx* ctx, const pob_reads& rd, int kind, uint8_t* out_seq, int32_t* out_s2s,
                       int8_t* out_path, int32_t* out_len, int32_t* out_status) {
  if (rd.n <= 0) return POB_OK;
  dim3 block(WARPS_PER_BLOCK * 32);
  dim3 grid((rd.n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
  pob_prof_scope ps(ctx, POB_K_VITERBI);
  if (rd.dtype == POB_F32 && rd.n_states == 5) {
#define POB_VIT(LAY, KND, PTH)                                                                              \
  viterbi5_f32_kernel<LAY, KND, PTH><<<grid, block, 0, ctx->stream>>>((const float*)rd.data, rd.row_off, rd.row_len, \
                                                                      rd.rc, rd.n, out_seq, out_s2s, out_path,      \
                                                                      out_len, out_status)
    const bool bl = rd.layout == POB_BLANK_LAST, bon = kind == POB_KIND_BONITO, pth = out_path != nullptr;
    if (bl && bon && pth) POB_VIT(POB_BLANK_LAST, POB_KIND_BONITO, true);
    else if (bl && bon) POB_VIT(POB_BLANK_LAST, POB_KIND_BONITO, false);
    else if (bl && pth) POB_VIT(POB_BLANK_LAST, POB_KIND_POREOVER, true);
    else if (bl) POB_VIT(POB_BLANK_LAST, POB_KIND_POREOVER, false);
    else if (bon && pth) POB_VIT(POB_BLANK_FIRST, POB_KIND_BONITO, true);
    else if (bon) POB_VIT(POB_BLANK_FIRST, POB_KIND_BONITO, false);
    else if (pth) POB_VIT(POB_BLANK_FIRST, POB_KIND_POREOVER, true);
    else POB_VIT(POB_BLANK_FIRST, POB_KIND_POREOVER, false);
#undef POB_VIT
  } else if (rd.dtype == POB_F32) {
    viterbi_generic_kernel<float><<<grid, block, 0, ctx->stream>>>((const float*)rd.data, rd.row_off, rd.row_len, rd.rc, rd.n,
                                                                   rd.n_states, rd.layout, kind, out_seq, out_s2s, out_path,
                                                                   out_len, out_status);
  } else {
    viterbi_generic_kernel<double><<<grid, block, 0, ctx->stream>>>((const double*)rd.data, rd.row_off, rd.row_len, rd.rc, rd.n,
                                                                    rd.n_states, rd.layout, kind, out_seq, out_s2s, out_path,
                                                                    out_len, out_status);
  }
  POB_CUDA(cudaGetLastError());
  return POB_OK;
}
