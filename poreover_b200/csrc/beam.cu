// CTC prefix beam search engine: single-read search, and the joint two-read search inside an
// alignment envelope ("row_col" and "row" traversals), for the ctc and ctc_merge_repeats trees.
//
// Reference semantics reproduced:
//   beam_search_                BeamSearch.h:18-58      (single read, rank by last_probability)
//   beam_search_2d_by_row       BeamSearch.h:110-172, :175-260 (rank by max_probability)
//   beam_search_2d_by_row_col   BeamSearch.h:262-397    (rank by max_probability_sym)
//   update_prob recurrences     PrefixTree.h:478-488, :518-531 (ctc), :649-663, :690-704 (merge repeats)
//   Beam::prune                 Beam.h:93-108 (dedupe, keep top W; exact ties -> earliest-created node)
//
// The reference keeps every node's forward values in per-node hash maps keyed by time, never frees a
// node, and relies on "missing key reads as -inf" plus stale entries.  Here one CTA owns one item
// (read or pair).  Every node has a home slot in a global-memory pool: a 128-byte header and, per read,
// a time-indexed ring window [lo, hi) of FP64 values -- exactly the entries the reference could still
// read, because every read is at t-1 >= (current u)-1.  The nodes of the *expanded beam* (beam + children,
// <= 5W) additionally hold their header in shared memory in a stable "active slot" for as long as they stay
// in the expanded beam, so a search step touches global memory only for window entries and for nodes that
// enter or leave.  Nodes that leave are retired, not dropped: a later revival (parent re-enters the beam) or
// a child reading its frozen parent sees the same stale values the reference's hash maps would return.  A
// retired node is recycled only once its windows are dead and it cannot hand retained children to a revival
// (or the pool overflows, which is flagged per item).
//
// Arithmetic: the reference works on log-probabilities in double with log-add-exp (Log.h:27-33).  Here every
// forward value is kept in the LINEAR domain, in FP64, scaled by a power of two per (read, timestep):
// value = exp(reference value) * 2^-K(r, t).  A forward cell is then two multiplications and two additions
// (no exp/log on the dependent chain); a missing hash-map key (-inf) is 0.  The probability rows enter through
// "column records" (exp of the five log-probabilities of a timestep, computed once, times the power of two that
// moves from K(t-1) to K(t)); K only changes when the top of the beam has drifted 64 binades from 1, so the
// rescaling is exact and the result does not depend on when it happens.  Measured against the log-domain
// reference: identical consensus strings and |score difference| ~ 1e-11 (bar: 1e-4).  No tensor cores: nothing
// here is a contraction.  Per step the dependent chain is the time-major band sweep: threads own
// (node, read) items, carry their own t-1 values in registers and exchange parent values through
// double-buffered shared memory, one block barrier per time sub-step; about seven more barriers per step cover
// the sweep's setup and keys, ranking, expansion, retirement and allocation.  Shared-memory hazards are checked
// with compute-sanitizer racecheck / synccheck on tools/sanitize_case.py (all traversals, trees and widths).
#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "launch.cuh"

#pragma nv_diag_suppress 177
#pragma nv_diag_suppress 550

namespace {

enum { MODE_1D = 0, MODE_ROW = 1, MODE_ROWCOL = 2 };
#ifndef POB_SCAN_WIDTH
#define POB_SCAN_WIDTH 4  // independent loads in flight when a clean band has to be scanned (8: no change, 16: spills)
#endif
enum { PS_ROOT = 0, PS_INE = 1, PS_FROZEN = 2, PS_DEAD = 3 };  // where a node's parent values come from
enum { KID_ACTIVE = 0, KID_REVIVE = 1, KID_FRESH = 2 };

struct __align__(16) NodeHdr {  // 128 bytes, home record of a node in the global pool
  uint32_t order;               // creation order (>= 1); 0 = slot is free / link invalid
  int32_t state;                // 0 active (in the expanded beam), > 0 retire stamp, -1 free
  int32_t parent_slot;          // -1: parent is the root
  uint32_t parent_order;
  int32_t parent_tid;
  int32_t tid;                  // trace id, assigned at first expansion; -1 before
  int32_t depth;
  int32_t last;
  int32_t kid_slot[4];
  uint32_t kid_order[4];
  int32_t lo[2], hi[2];         // window of valid time keys per read: lo <= t < hi
  double maxp[2];               // max_prob[] of the reference nodes (linear domain) ...
  int32_t maxk[2];              // ... and the scale (power of two) they are expressed in
  int32_t aslot;                // active slot while state == 0
};

template <int MODEL>
struct Entry;
template <>
struct __align__(16) Entry<POB_MODEL_CTC_MERGE_REPEATS> {
  double gap, nogap;  // prob = gap + nogap (one rounding, PrefixTree.h:128): recomputed by the reader, bit for bit
};
template <>
struct __align__(8) Entry<POB_MODEL_CTC> {
  double prob;
};

// exp(log-probabilities) of one timestep of a read, times 2^(K(t-1) - K(t)); K = scale of the column (value =
// stored * 2^K); nb = number of scale changes up to this column; root = value of the root node (ctc only)
struct __align__(64) Col {
  double y[5];   // bases 0..S-2, blank at [4]
  double root;
  int32_t K, nb;
  double pad;
};
enum { COL_LOOK = 8 };  // columns are created this many at a time

struct BeamParams {
  pob_reads r[2];
  const int32_t* env;       // rows x 2 (may be NULL: ROW without envelope, or 1D)
  const int64_t* env_off;
  const int32_t* envt;      // cols x 2 (ROWCOL only)
  const int64_t* envt_off;
  const int32_t* order;     // item processing order or NULL
  const int32_t* skip;      // per item != 0 -> not searched
  int n_items, W, mode, NP, CAP0, CAP1, CAPC0, CAPC1, RQ, EMAX;
  int dbg_noreclaim, dbg_noreuse, dbg_long;
  int inspect_every;        // the retire queue is inspected every this many expansions (its headers are cold)
  int col_off;              // column records in shared memory: byte offset, or -1 (global workspace)
  int prefetch;             // pull the probability rows / envelope entries of coming steps towards the SM
  double* dbg_trace;        // optional: [step][2] = (top score, sum of beam scores) after each prune
  char* ws;                 // workspace, one stride per resident CTA
  size_t ws_stride;
  uint32_t* trace;
  const int64_t* trace_off;
  int32_t* out_top;         // per item: start tid, extra last (-1), depth, status
  double* out_score;
  int* work_counter;
  unsigned long long* counters;
};

__device__ __forceinline__ double no_nan(double x) { return (x == x) ? x : 0.0; }  // NaN cannot be ranked

// x * 2^dk: exact while the result stays in the normal range (the factor is built from its exponent field)
__device__ __forceinline__ double scale2(double x, int dk) {
  dk = max(-1000, min(1000, dk));
  return x * __longlong_as_double((long long)(1023 + dk) << 52);
}

// a_kid entries: the child's pool slot in the low 16 bits and, while the child is in the expanded beam, its active slot
// + 1 above them (the pool has at most 65536 slots); -1 = the child was never created
__device__ __forceinline__ int kid_slot(int v) { return v < 0 ? -1 : (v & 0xffff); }
__device__ __forceinline__ int kid_act(int v) { return v < 0 ? -1 : (v >> 16) - 1; }
__device__ __forceinline__ int kid_pack(int slot, int act) { return slot | ((act + 1) << 16); }

__device__ __forceinline__ unsigned sign_acc(unsigned d, unsigned acc) {
  unsigned r;
  asm("mad.hi.u32 %0, %1, 2, %2;" : "=r"(r) : "r"(d), "r"(acc));
  return r;
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// physical address helpers for one read of the batch
struct ReadView {
  const void* base;  // first row of the read
  int T, S, layout, rc, f64;
  int cblank;        // physical column of blank
  __device__ __forceinline__ int prow(int t) const { return rc ? (T - 1 - t) : t; }
  __device__ __forceinline__ int pcol(int k) const { return pob_col(k, S, layout, rc); }
  __device__ __forceinline__ double at(int t, int pc) const {
    size_t i = (size_t)prow(t) * S + pc;
    return f64 ? __ldg((const double*)base + i) : (double)__ldg((const float*)base + i);
  }
};

__device__ __forceinline__ ReadView make_view(const pob_reads& r, int item) {
  ReadView v;
  const int64_t ro = r.row_off[item];
  v.T = pob_read_len(r.row_off, r.row_len, item);
  v.S = r.n_states;
  v.layout = r.layout;
  v.rc = r.rc ? (r.rc[item] != 0) : 0;
  v.f64 = r.dtype == POB_F64;
  const size_t es = v.f64 ? 8 : 4;
  v.base = (const char*)r.data + (size_t)ro * v.S * es;
  v.cblank = v.pcol(v.S - 1);
  return v;
}

enum { SH_NB = 0, SH_NUSED, SH_FQH, SH_FQT, SH_AFREE, SH_ORDER, SH_TID, SH_STAMP, SH_RQH, SH_RQT, SH_STATUS, SH_FIRSTALIVE,
       SH_TOTALLOC, SH_TOTFIRST, SH_DMIN, SH_TB0, SH_TB1, SH_XSTEP, SH_NL0, SH_NL1, SH_SPARE0,
       SH_SPARE1, SH_PCNT, SH_PM0, SH_PM1, SH_PM2, SH_PM3,
       SH_CDEF0 /* columns [0, cdef) of read 0 exist */, SH_CDEF1,
       SH_ENV = 32 /* [2][2] band of row u */, SH_ENVT = 36 /* [2][2] band of column v */,
       SH_KLAST0 = 40 /* scale of the newest column */, SH_KLAST1, SH_PEND0 /* scale change asked for */, SH_PEND1,
       SH_NBUMP0, SH_NBUMP1, SH_KCUR0 /* scale of the column of the last single update */, SH_KCUR1,
       SH_KREF0 /* scale of the last band (debug trace) */, SH_KREF1, SH_EBASE /* exponent base of the 31-bit keys */,
       SH_COUNT = 52 };

// Per-CTA engine state.  Scalars and global-memory views live in a static __shared__ struct; the per-active-slot
// arrays live in dynamic shared memory at offsets that depend only on EMAX / W / NP, so every access compiles
// to LDS/STS (a pointer loaded from memory would be a generic pointer and cost a second, slower load).
struct EngState {
  int W, NP, RQ, EMAX, mode, noreclaim, inspect_every, longq;
  int nbase;  // letters of the alphabet (states - 1): children per node
  NodeHdr* hdr;
  char* win[2];
  int32_t* freelist;
  int2* retq;
  Col* col[2];      // column records of each read: ring of cmask + 1 timesteps
  int cmask[2];
  int32_t* sufmin;  // ROW: min over rows >= u of the envelope's band start (band starts are NOT monotone:
                    // build_envelope's repair pass clamps some rows far back, envelope.py:82-85)
  int cap[2], mask[2];
  ReadView rv[2];
  uint32_t* trace;
};
__shared__ EngState g_es;

// Optional per-phase cycle attribution (tools/phase_clocks.py builds a second library with -DPOB_PHASE_CLOCKS):
// thread 0 accumulates the cycles between consecutive marks into g_phase_clk[mark id].
#ifdef POB_PHASE_CLOCKS
__device__ unsigned long long g_phase_clk[32];
__shared__ long long g_pclk_last;
#define PCLK(i)                                                                  \
  do {                                                                           \
    if (threadIdx.x == 0) {                                                      \
      const long long c_ = clock64();                                            \
      atomicAdd(&g_phase_clk[i], (unsigned long long)(c_ - g_pclk_last));        \
      g_pclk_last = c_;                                                          \
    }                                                                            \
  } while (0)
#else
#define PCLK(i) do { } while (0)
#endif
__device__ unsigned long long g_exact_prunes;  // how often the ranking needed its exact pass (diagnostic)
#ifdef POB_COUNT_RESCAN
__device__ unsigned long long g_dbg[8];  // why clean ranges were scanned (diagnostic build)
#endif
extern __shared__ __align__(128) char pob_smem[];

// Shared-memory arrays of the active slots (index a in [0, EMAX)); a_slot[a] < 0 = unused.
//   a_slot            pool slot                 a_order / a_porder   creation order of the node / its parent
//   a_par             active slot of the parent when a_pstat == PS_INE
//   a_pslot           pool slot of the parent (window base when frozen)
//   a_kid/a_kido [4]  children: pool slot | (active slot + 1) << 16 (kid_slot() / kid_act()), -1 = never created; orders
//   a_lo/a_hi [2]     window bounds per read;   a_plo/a_phi [2] bounds of a frozen parent
//   a_che [2]         clean end per read (see sweep)
//   a_maxp [2]        max_prob[] of the reference nodes;  a_last0 value of read 0 at its last written t
//   key               ranking score (-1: unused slot);  a_maxk [2] scale of a_maxp;  pub [2][EMAX][2] (prob, gap) parent -> child exchange
//   a_same            parent's last base == own last base (merge-repeats reads the parent's gap value)
//   beam [W]          active slots in rank order;  a_free stack of unused active slots;  sh scalars SH_*
#define POB_VIEWS                                                                                   \
  const int EMAX = EM_CT ? EM_CT : g_es.EMAX, W = W_CT ? W_CT : g_es.W;                             \
  const int NP = g_es.NP, RQ = g_es.RQ, mode = MODE_CT >= 0 ? MODE_CT : g_es.mode;                  \
  char* const sm_ = pob_smem;                                                                       \
  double2* const pub = (double2*)sm_;                                                               \
  double* const key = (double*)(sm_ + 64 * EMAX);                                                   \
  int32_t* const a_maxk = (int32_t*)(sm_ + 72 * EMAX);                                              \
  double* const a_maxp = (double*)(sm_ + 80 * EMAX);                                                \
  double* const a_last0 = (double*)(sm_ + 96 * EMAX);                                               \
  int32_t* const a_slot = (int32_t*)(sm_ + 104 * EMAX);                                             \
  uint32_t* const a_order = (uint32_t*)(sm_ + 108 * EMAX);                                          \
  int32_t* const a_par = (int32_t*)(sm_ + 112 * EMAX);                                              \
  int32_t* const a_pslot = (int32_t*)(sm_ + 116 * EMAX);                                            \
  uint32_t* const a_porder = (uint32_t*)(sm_ + 120 * EMAX);                                         \
  int32_t* const a_depth = (int32_t*)(sm_ + 124 * EMAX);                                            \
  int32_t* const a_tid = (int32_t*)(sm_ + 128 * EMAX);                                              \
  int32_t* const a_ptid = (int32_t*)(sm_ + 132 * EMAX);                                             \
  int32_t* const a_kid = (int32_t*)(sm_ + 136 * EMAX);                                              \
  uint32_t* const a_kido = (uint32_t*)(sm_ + 152 * EMAX);                                           \
  int32_t* const a_lo = (int32_t*)(sm_ + 168 * EMAX);                                               \
  int32_t* const a_hi = (int32_t*)(sm_ + 176 * EMAX);                                               \
  int32_t* const a_plo = (int32_t*)(sm_ + 184 * EMAX);                                              \
  int32_t* const a_phi = (int32_t*)(sm_ + 192 * EMAX);                                              \
  int32_t* const a_free = (int32_t*)(sm_ + 200 * EMAX);                                             \
  int32_t* const tmpa = (int32_t*)(sm_ + 204 * EMAX);                                               \
  int32_t* const tmpb = (int32_t*)(sm_ + 208 * EMAX);                                               \
  int32_t* const tmpc = (int32_t*)(sm_ + 212 * EMAX);                                               \
  int32_t* const a_che = (int32_t*)(sm_ + 216 * EMAX);                                              \
  int32_t* const beam = (int32_t*)(sm_ + 224 * EMAX);                                               \
  int32_t* const sh = beam + ((W + 3) & ~3);                                                        \
  uint8_t* const a_last = (uint8_t*)(sh + SH_COUNT);                                                \
  const int eb_ = (EMAX + 15) & ~15;                                                                \
  uint8_t* const a_pstat = a_last + eb_;                                                            \
  uint8_t* const a_same = a_pstat + eb_;                                                            \
  uint8_t* const a_inbeam = a_same + eb_;                                                           \
  uint8_t* const a_needed = a_inbeam + eb_;                                                         \
  const int E4 = (EMAX + 3) & ~3;                                                                   \
  uint32_t* const k32 = (uint32_t*)(a_needed + eb_);                                                \
  int32_t* const a_maxt = (int32_t*)(k32 + E4);   /* [2] timestep that holds a_maxp, -1 = unknown */ \
  NodeHdr* const hdr = g_es.hdr;                                                                    \
  int32_t* const freelist = g_es.freelist;                                                          \
  int2* const retq = g_es.retq;                                                                     \
  uint32_t* const trace = g_es.trace;

// EM_CT / W_CT: active-slot count and beam width known at compile time (0 = read from the engine state): the
// benchmark's configuration (beam width 25, 128 slots) gets its own instantiation, in which the offsets of the
// shared-memory arrays and the trip counts of the ranking loop are constants.
template <int MODEL, int EM_CT = 0, int W_CT = 0, int MODE_CT = -1>
struct Engine {
  typedef Entry<MODEL> Ent;

  // letters of the alphabet: the specialised instantiations are only launched on five-state reads
  __device__ __forceinline__ static int nbase() { return EM_CT ? 4 : g_es.nbase; }

  // window entries: what is stored and what a reader gets back
  struct EV { double prob, gap, nogap; };
  __device__ __forceinline__ static EV ld_ent(const Ent* e) {
    EV v;
    if constexpr (MODEL == POB_MODEL_CTC_MERGE_REPEATS) {
      const double2 q = *reinterpret_cast<const double2*>(e);
      v.gap = q.x; v.nogap = q.y; v.prob = __dadd_rn(q.x, q.y);
    } else {
      v.prob = e->prob; v.gap = 0.0; v.nogap = 0.0;
    }
    return v;
  }
  __device__ __forceinline__ static double ld_prob(const Ent* e) { return ld_ent(e).prob; }
  __device__ __forceinline__ static void st_ent(Ent* o, double prob, double gp, double ng) {
    if constexpr (MODEL == POB_MODEL_CTC_MERGE_REPEATS) {
      double2 q; q.x = gp; q.y = ng;
      *reinterpret_cast<double2*>(o) = q;
    } else {
      o->prob = prob;
    }
  }
  __device__ __forceinline__ Ent* wbase(int slot, int r) const {
    return reinterpret_cast<Ent*>(g_es.win[r]) + (size_t)slot * g_es.cap[r];
  }

  // value of the root at time t (parent of depth-1 nodes), in the scale of column t (column -1 has scale 0)
  __device__ __forceinline__ double root_prob(int r, int t) const {
    if (t == -1) return 1.0;
    if (MODEL == POB_MODEL_CTC) return (t >= 0 && t < g_es.rv[r].T) ? colp(r, t)->root : 0.0;
    return 0.0;
  }

  __device__ __forceinline__ Col* colp(int r, int t) const { return g_es.col[r] + (t & g_es.cmask[r]); }

  // ---- one forward cell (PrefixTree.h:518-531 ctc, :690-704 merge repeats) in the scaled linear domain ----
  // Every site that computes a cell goes through this function with explicitly rounded operations, so that a
  // recomputation from the same inputs reproduces the stored value bit for bit (the incremental sweep relies on it).
  __device__ __forceinline__ static void cell(double p_prev, double ng_prev, double pv, double yl, double yb, double& prob,
                                              double& gp, double& ng) {
    if constexpr (MODEL == POB_MODEL_CTC_MERGE_REPEATS) {
      gp = __dmul_rn(p_prev, yb);                     // gap    = prob(t-1) * y[blank]
      ng = __dmul_rn(__dadd_rn(pv, ng_prev), yl);     // no_gap = (parent(t-1) + no_gap(t-1)) * y[last]
      prob = __dadd_rn(gp, ng);
    } else {
      gp = 0.0; ng = 0.0;
      prob = __dadd_rn(__dmul_rn(pv, yl), __dmul_rn(p_prev, yb));
    }
  }

  // ---- column records: exp(log-probabilities) of a timestep, scaled by a power of two ----
  // All threads call this with the same arguments (the columns [0, need_r) of read r are about to be used); it does
  // nothing (and has no barrier) when they exist already.  New columns are created COL_LOOK at a time.  The first new
  // column absorbs the pending change of scale of its read (see prune()): its factors are multiplied by 2^-pend.
  __device__ __noinline__ void define_cols(int need0, int need1) {
    POB_VIEWS
    const int tid = threadIdx.x, NT = blockDim.x;
    for (int r = 0; r < 2; ++r) {
      const int need = r ? need1 : need0;
      const int c0 = sh[SH_CDEF0 + r];
      const ReadView& v = g_es.rv[r];
      if (need <= c0 || c0 >= v.T) continue;  // uniform
      const int c1 = min(v.T, max(need, c0 + COL_LOOK));
      const int pend = sh[SH_PEND0 + r];
      const int K = sh[SH_KLAST0 + r] + pend;
      const int nb = sh[SH_NBUMP0 + r] + (pend != 0);
      const int nbase = v.S - 1;
      for (int i = tid; i < (c1 - c0) * 8; i += NT) {
        const int t = c0 + (i >> 3), k = i & 7;
        Col* c = colp(r, t);
        if (k < 5) {
          double y = 0.0;
          if (k == 4) y = exp(v.at(t, v.cblank));
          else if (k < nbase) y = exp(v.at(t, v.pcol(k)));
          if (t == c0 && pend) y = scale2(y, -pend);
          c->y[k] = y;
        } else if (k == 5) {
          c->K = K; c->nb = nb;
        }
      }
      __syncthreads();
      if (tid == 0) {
        if constexpr (MODEL == POB_MODEL_CTC) {
          // PrefixTree.h:508-514: the root's value is the running product of the blank column
          double s = (c0 == 0) ? 1.0 : colp(r, c0 - 1)->root;
          for (int t = c0; t < c1; ++t) { s = __dmul_rn(s, colp(r, t)->y[4]); colp(r, t)->root = s; }
        }
        sh[SH_CDEF0 + r] = c1; sh[SH_KLAST0 + r] = K; sh[SH_PEND0 + r] = 0; sh[SH_NBUMP0 + r] = nb;
      }
      __syncthreads();
    }
  }
  __device__ __forceinline__ void need_cols(int need0, int need1) {
    POB_VIEWS
    if (need0 > sh[SH_CDEF0] || need1 > sh[SH_CDEF1]) define_cols(need0, need1);
  }

  // bookkeeping that must not race with the phase that produced it: run by thread 0 at the start of a
  // compute phase (separated by barriers from every reader)
  __device__ __forceinline__ void deferred_finalize() {
    POB_VIEWS
    if (threadIdx.x == 0) {
      sh[SH_ORDER] += sh[SH_TOTALLOC]; sh[SH_TOTALLOC] = 0;
      sh[SH_TID] += sh[SH_TOTFIRST]; sh[SH_TOTFIRST] = 0;
      sh[SH_FIRSTALIVE] = 0x7fffffff;
    }
  }

  // ---- one update_prob(n, r, t) for a set of nodes at the same t (1D steps, ROW's read 0, skip steps) ----
  // Gather and commit are separated by a block barrier: when a node and its parent are updated at the same
  // t by different threads, the child must see the parent's window bounds and t-1 entry as they were before
  // this phase (the reference reads t-1, writes t).
  struct UpdIn {
    double p_prev, ng_prev, pv, ylast, yblank;
    int lo, hi, kt;
  };

  __device__ __forceinline__ void update_gather(int a, int r, int t, UpdIn& in) const {
    POB_VIEWS
    const int slot = a_slot[a];
    const int last = a_last[a];
    in.lo = a_lo[2 * a + r]; in.hi = a_hi[2 * a + r];
    const bool self_ok = (t - 1 >= in.lo && t - 1 < in.hi);
    const Ent* se = wbase(slot, r) + (t & g_es.mask[r]);
    in.p_prev = 0.0; in.ng_prev = 0.0;
    if (self_ok) { const EV v = ld_ent(se); in.p_prev = v.prob; in.ng_prev = v.nogap; }
    const Col* c = colp(r, t);
    in.ylast = c->y[last];
    in.yblank = c->y[4];
    in.kt = c->K;
    const int ps = a_pstat[a];
    if (ps == PS_ROOT) {
      in.pv = root_prob(r, t - 1);
    } else if (ps == PS_DEAD) {
      in.pv = 0.0;
    } else {
      int plo, phi;
      if (ps == PS_INE) { const int pa = a_par[a]; plo = a_lo[2 * pa + r]; phi = a_hi[2 * pa + r]; }
      else { plo = a_plo[2 * a + r]; phi = a_phi[2 * a + r]; }
      if (t - 1 >= plo && t - 1 < phi) {
        const Ent* pe = wbase(a_pslot[a], r) + (t & g_es.mask[r]);
        const EV v = ld_ent(pe);
        in.pv = (MODEL == POB_MODEL_CTC_MERGE_REPEATS && a_same[a]) ? v.gap : v.prob;
      } else {
        in.pv = 0.0;
      }
    }
  }

  __device__ __forceinline__ double update_commit(int a, int r, int t, const UpdIn& in) {
    POB_VIEWS
    double prob, gp, ng;
    cell(in.p_prev, in.ng_prev, in.pv, in.ylast, in.yblank, prob, gp, ng);
    st_ent(wbase(a_slot[a], r) + ((t + 1) & g_es.mask[r]), prob, gp, ng);
    a_maxt[2 * a + r] = -1;  // a_maxp becomes a running maximum: the next sweep scans its clean entries
    int lo = in.lo, hi = in.hi;
    if (t >= hi) { if (t > hi) lo = t; hi = t + 1; }
    else if (t < lo) { lo = t; hi = t + 1; }
    if (hi - lo > g_es.cap[r]) lo = hi - g_es.cap[r];
    a_lo[2 * a + r] = lo; a_hi[2 * a + r] = hi;
    if (t >= a_che[2 * a + r]) a_che[2 * a + r] = t + 1;  // a single update extends (or restarts) the clean range
    {
      // max_prob (PrefixTree.h:134, :420): the stored maximum carries the scale of the column it came from
      const double mp = a_maxp[2 * a + r];
      const int mk = a_maxk[2 * a + r];
      const double mine = (mk == in.kt) ? prob : scale2(prob, in.kt - mk);
      if (mine > mp) { a_maxp[2 * a + r] = prob; a_maxk[2 * a + r] = in.kt; }
    }
    if (r == 0) a_last0[a] = prob;
    return prob;
  }

  // all threads call this; `mine` selects the threads that own an active slot this phase
  __device__ __noinline__ double update_all(bool mine, int a, int r, int t) {
    UpdIn in;
    double p = 0;
    need_cols(r == 0 ? t + 1 : 0, r == 1 ? t + 1 : 0);
    if (threadIdx.x == 0) { POB_VIEWS sh[SH_KCUR0 + r] = colp(r, t)->K; }
    deferred_finalize();
    if (mine) update_gather(a, r, t, in);
    __syncthreads();
    if (mine) p = update_commit(a, r, t, in);
    __syncthreads();
    PCLK(11);
    return p;
  }

  // value a child reads from its frozen parent for an update at time t (the parent's entry at t-1)
  __device__ __forceinline__ double frozen_at(const Ent* pwb, int t, int wmask, int plo, int phi, bool same) const {
    if (t - 1 < plo || t - 1 >= phi) return 0.0;
    const Ent* q = pwb + (t & wmask);
    const EV v = ld_ent(q);
    return (MODEL == POB_MODEL_CTC_MERGE_REPEATS && same) ? v.gap : v.prob;
  }

  // Everything a thread needs to recompute cells of one (active slot, read): views of the node's window, of its
  // parent's window and of the read's column records.  Filled from the shared-memory arrays, so that any thread can
  // take over any item.
  struct SwItem {
    Ent* wb;
    const Ent* pwb;
    const Col* cb;   // column ring of the read
    int lo, hi, plo, phi, pstat, pa, wmask, cmask, last, kref;
    bool same, mixed;
  };

  __device__ __forceinline__ void load_item(int a, int r, int ts, int te, SwItem& I) const {
    POB_VIEWS
    I.lo = a_lo[2 * a + r]; I.hi = a_hi[2 * a + r];
    I.wmask = g_es.mask[r];
    I.wb = wbase(a_slot[a], r);
    I.cb = g_es.col[r]; I.cmask = g_es.cmask[r];
    I.last = a_last[a];
    // scale of the band: that of its newest column; `mixed` when the scale changes inside the band (rare)
    I.kref = colp(r, te - 1)->K;
    I.mixed = colp(r, te - 1)->nb != colp(r, ts)->nb;
    I.pstat = a_pstat[a];
    I.same = a_same[a] != 0;
    I.pa = 0; I.plo = 0; I.phi = 0; I.pwb = nullptr;
    if (I.pstat == PS_INE) {
      I.pa = a_par[a];
      I.plo = a_lo[2 * I.pa + r]; I.phi = a_hi[2 * I.pa + r]; I.pwb = wbase(a_pslot[a], r);
    } else if (I.pstat == PS_FROZEN) {
      I.plo = a_plo[2 * a + r]; I.phi = a_phi[2 * a + r]; I.pwb = wbase(a_pslot[a], r);
    }
  }

  __device__ __forceinline__ void load_y(const SwItem& I, int t, double& yl, double& yb) const {
    const Col* c = I.cb + (t & I.cmask);
    yl = c->y[I.last]; yb = c->y[4];
  }

  __device__ __forceinline__ double parent_at(const SwItem& I, int r, int t) const {
    if (I.pstat == PS_INE || I.pstat == PS_FROZEN) return frozen_at(I.pwb, t, I.wmask, I.plo, I.phi, I.same);
    if (I.pstat == PS_ROOT) return root_prob(r, t - 1);
    return 0.0;
  }

  // value of column t expressed in the scale of the band (only needed when the scale changes inside the band)
  __device__ __forceinline__ double in_band_scale(const SwItem& I, double x, int t) const {
    if (!I.mixed) return x;
    return scale2(x, (I.cb + (t & I.cmask))->K - I.kref);
  }

  // Inputs of the first two cells of a private recomputation starting at cs: the node's own values at cs - 1, the
  // column factors and the parent's values.  None of them is written during the private phase of a sweep, so the
  // loads can be issued long before their use (before the scan and the barrier that precede that phase).
  // (Four cells ahead was measured slower: 4432 against 4833 pairs/s.)
  struct ChainIn {
    double p_prev, ng_prev, yl, yb, pv, yl_n, yb_n, pv_n;
  };
  __device__ __forceinline__ void chain_preload(const SwItem& I, int r, int cs, int te, ChainIn& C) const {
    C.p_prev = 0.0; C.ng_prev = 0.0;
    if (cs - 1 >= I.lo && cs - 1 < I.hi) {
      const Ent* se = I.wb + (cs & I.wmask);
      const EV v = ld_ent(se);
      C.p_prev = v.prob; C.ng_prev = v.nogap;
    }
    load_y(I, cs, C.yl, C.yb);
    C.pv = parent_at(I, r, cs);
    C.yl_n = 0; C.yb_n = 0; C.pv_n = 0;
    if (cs + 1 < te) { load_y(I, cs + 1, C.yl_n, C.yb_n); C.pv_n = parent_at(I, r, cs + 1); }
  }

  // band maximum and the (latest) timestep that holds it
  __device__ __forceinline__ static void fold_max(double& maxv, int& maxt, double x, int t) {
    if (x >= maxv) { maxv = x; maxt = t; }
  }


  // Private recomputation of the cells [cs, lim) of one (node, read), cs < lim: every parent entry read here is
  // final.  Leaves the node's values at lim - 1 in p_prev / ng_prev / g_prev and folds the new values into maxv.
  __device__ __forceinline__ void chain(const SwItem& I, int a, int r, int cs, int lim, const ChainIn& C,
                                        double& p_prev, double& ng_prev, double& g_prev, double& maxv, int& maxt) const {
    p_prev = C.p_prev; ng_prev = C.ng_prev;
    // inputs are requested two timesteps ahead of their use
    double yl = C.yl, yb = C.yb, pv = C.pv, yl_n = C.yl_n, yb_n = C.yb_n, pv_n = C.pv_n;
    for (int t = cs; t < lim; ++t) {
      double yl_n2 = 0, yb_n2 = 0, pv_n2 = 0;
      if (t + 2 < lim) { load_y(I, t + 2, yl_n2, yb_n2); pv_n2 = parent_at(I, r, t + 2); }
      double prob, gp, ng;
      cell(p_prev, ng_prev, pv, yl, yb, prob, gp, ng);
      Ent* o = I.wb + ((t + 1) & I.wmask);
      st_ent(o, prob, gp, ng);
      if constexpr (MODEL == POB_MODEL_CTC_MERGE_REPEATS) { ng_prev = ng; g_prev = gp; }
      fold_max(maxv, maxt, in_band_scale(I, prob, t), t);
      p_prev = prob;
      yl = yl_n; yb = yb_n; pv = pv_n;
      yl_n = yl_n2; yb_n = yb_n2; pv_n = pv_n2;
    }
  }

  // ---- band sweep over the expanded beam (BeamSearch.h:361-375, :146-156), incremental ----------------
  // reads_mask bit r: read r swept over [s_r, e_r).  Thread (2a + r) owns (active slot a, read r).
  //
  // The reference recomputes every node over the whole band at every step.  Recomputing an entry whose
  // inputs did not change reproduces the same value, so only the entries that can differ are computed:
  //   * a_che[a][r] ("clean end"): entries [.., che) of the node's window were produced by the previous
  //     sweep / single updates from inputs that are still current; cs = clamp(che, s, e) is where new work
  //     starts (new and revived nodes: cs = s);
  //   * phase A (no barrier): each item computes [cs, min(e, Tb)) on its own; all parent entries it reads
  //     there are final because Tb = 1 + min over live-parent items of the parent's cs;
  //   * phase B (one barrier per timestep): the time-major loop over [Tb, e) with parent values exchanged
  //     through shared memory.  A node whose live parent produced a new value at t-1 recomputes from t on
  //     even inside its own clean range (dirtiness propagates down the tree);
  //   * clean entries only contribute to the band maximum (max_prob is reset every step in the reference).
  // `full` disables the reuse (every node recomputed from s): the literal reference schedule.
  // Band maxima are kept in the scale of the band's newest column (SwItem::kref), the same for every node.
  __device__ __noinline__ void sweep(int reads_mask_, int s0, int e0, int s1, int e1, bool full_,
                                     unsigned long long& n_updates) {
    // the row_col instantiation sweeps both reads incrementally: constants (the launcher keeps the debug switch that
    // turns the reuse off on the generic instantiation)
    const int reads_mask = MODE_CT == MODE_ROWCOL ? 3 : reads_mask_;
    const bool full = MODE_CT == MODE_ROWCOL ? false : full_;
    // algorithmic count (what the reference evaluates): every used slot over every swept band; kept by thread 0 alone
    if (threadIdx.x == 0) {
      POB_VIEWS
      n_updates += (unsigned long long)sh[SH_NUSED] *
                   (unsigned long long)(((reads_mask & 1) ? max(e0 - s0, 0) : 0) + ((reads_mask & 2) ? max(e1 - s1, 0) : 0));
    }
    POB_VIEWS
    const int tid = threadIdx.x;
    const int a = tid >> 1, r = tid & 1;
    need_cols((reads_mask & 1) ? e0 : 0, (reads_mask & 2) ? e1 : 0);
    const bool used = a < EMAX && a_slot[a] >= 0;
    const bool on = used && ((reads_mask >> r) & 1);
    int cs = 0;
    double p_prev = 0.0, ng_prev = 0.0, g_prev = 0.0, maxv = 0.0;
    int maxt = -1;  // timestep of maxv
#ifdef POB_COUNT_RESCAN
    bool dbg_rescan = false;
    int dbg_c0 = -1;
#endif
    const int ts = r ? s1 : s0, te = r ? e1 : e0;
    SwItem I;
    ChainIn C;
    C.p_prev = C.ng_prev = C.yl = C.yb = C.pv = C.yl_n = C.yb_n = C.pv_n = 0.0;
    I.lo = I.hi = I.plo = I.phi = I.pa = I.wmask = I.cmask = I.last = I.kref = 0; I.pstat = PS_DEAD; I.same = I.mixed = false;
    I.wb = nullptr; I.pwb = nullptr; I.cb = nullptr;
    PCLK(12);
    deferred_finalize();
    PCLK(13);
    if (on && te > ts) {
      load_item(a, r, ts, te, I);
      const int che = a_che[2 * a + r];
      cs = full ? ts : min(max(che, ts), te);
      if (I.pstat == PS_INE) {
        const int pche = a_che[2 * I.pa + r];
        const int pcs = full ? ts : min(max(pche, ts), te);
        atomicMin(&sh[SH_TB0 + r], pcs + 1);
      }
      // (a node that entered the expanded beam in this step recomputes its whole band in phase A, as one dependent
      // chain on its own thread.  Queueing those chains for the last warps of the block, one chain per lane, lowered
      // the instruction count and raised the latency of the step -- 4.2 k against 5.9 k pairs/s, its noinline call
      // cost more registers than it saved -- and loading the chain's first inputs before the barrier below was +2 % at
      // two CTAs per SM and -4 % at three: both are gone from the source.)
      PCLK(14);
      // Clean part of the band, [c0, c1): its entries are final and only their maximum is needed.  The band maximum of
      // the node's previous sweep (a_maxp, in the scale a_maxk, found at timestep a_maxt) is that maximum whenever its
      // timestep is still inside: the entries that left the band since (band starts only move forward) were not larger,
      // and the rest are the same values in the same scale.  Otherwise -- the maximum has left the band, the scale of
      // the band changed, single updates came in between -- the entries are scanned (four independent loads in flight).
      // Measured (444 pairs): 9 % of the (node, read) steps still scan, nearly all of them decaying prefixes whose
      // maximum is the first entry of the band at every step.  Scanning at every step (with the newest 16 values
      // mirrored in shared memory) was 33 % of the kernel's instructions; this is 7.6 % fewer instructions at the same
      // speed (the step waits for the warps that recompute whole bands, not for the scans).  Also tracking where the
      // non-increasing tail of the values starts (then the maximum is the first entry: one load) halves the scans left
      // but costs more than it saves: 6363 against 6665 pairs/s.
      {
        const int c0 = max(ts, I.lo), c1 = min(cs, I.hi);
        const int pt = a_maxt[2 * a + r];
        if (pt >= c0 && pt < c1 && a_maxk[2 * a + r] == I.kref) {
          maxv = a_maxp[2 * a + r]; maxt = pt;
        } else {
#ifdef POB_COUNT_RESCAN
          dbg_rescan = c1 - c0 >= 4;
          if (dbg_rescan) {
            const int why = pt < 0 ? 0 : (a_maxk[2 * a + r] != I.kref ? 1 : (pt < c0 ? 2 : 3));
            atomicAdd(&g_dbg[why], 1ULL);
            dbg_c0 = c0;
          }
#endif
          // POB_SCAN_WIDTH independent loads per round trip to L1 / L2
          for (int t = c0; t < c1; t += POB_SCAN_WIDTH) {
            double v[POB_SCAN_WIDTH];
#pragma unroll
            for (int q = 0; q < POB_SCAN_WIDTH; ++q) {
              v[q] = 0.0;
              if (t + q < c1) v[q] = ld_prob(I.wb + ((t + q + 1) & I.wmask));
            }
#pragma unroll
            for (int q = 0; q < POB_SCAN_WIDTH; ++q)
              if (t + q < c1) fold_max(maxv, maxt, in_band_scale(I, v[q], t + q), t + q);
          }
        }
      }
    }
#ifdef POB_COUNT_RESCAN
    if (dbg_rescan) {
      // where the scan found the maximum: first entry (4), first four (5), elsewhere (6)
      atomicAdd(&g_dbg[maxt == dbg_c0 ? 4 : (maxt < dbg_c0 + 4 ? 5 : 6)], 1ULL);
      if (maxv == 0.0) atomicAdd(&g_dbg[7], 1ULL);
    }
    {  // diagnostic build: lanes (low 40 bits) and warps (high bits) that scanned a clean range of >= 4 entries
      const unsigned m = __ballot_sync(0xffffffffu, dbg_rescan);
      if ((threadIdx.x & 31) == 0 && m)
        atomicAdd(reinterpret_cast<unsigned long long*>(&g_exact_prunes), (1ULL << 40) + (unsigned long long)__popc(m));
    }
#endif
    PCLK(15);
    __syncthreads();
    PCLK(1);
    const int Tb0 = sh[SH_TB0], Tb1 = sh[SH_TB1];
    const int Tb = r ? Tb1 : Tb0;  // first timestep of this read that needs the synchronised loop
    bool computing = false;        // p_prev / ng_prev hold the node's values at the previous timestep
    const int limA = min(te, Tb);
    // ---- phase A: private work [cs, min(te, Tb))
    if (on && te > ts && cs < limA) {
      computing = true;
      chain_preload(I, r, cs, te, C);
      chain(I, a, r, cs, limA, C, p_prev, ng_prev, g_prev, maxv, maxt);
    }
    PCLK(2);
    // ---- phase B: synchronised time-major loop over [Tb, te)
    const int iters = max(max((reads_mask & 1) ? e0 - Tb0 : 0, (reads_mask & 2) ? e1 - Tb1 : 0), 0);
    if (iters > 0) {
      uint8_t* const pchg = reinterpret_cast<uint8_t*>(tmpa);  // [2][EMAX*2] "parent value changed" flags
      const bool inB = on && te > ts && Tb < te;
      double ylast = 0, yblank = 0;
      if (inB) {
        // publish the node's value at Tb-1: computed in phase A, or a stored clean entry
        double2 pb; pb.x = 0.0; pb.y = 0.0;
        const int tp = Tb - 1;
        if (computing) { pb.x = p_prev; pb.y = g_prev; }
        else if (tp >= I.lo && tp < I.hi) {
          const Ent* se = I.wb + ((tp + 1) & I.wmask);
          const EV v = ld_ent(se);
          pb.x = v.prob; pb.y = v.gap;
        }
        pub[a * 2 + r] = pb;
        pchg[a * 2 + r] = computing;
        load_y(I, Tb, ylast, yblank);
      }
      __syncthreads();
      PCLK(3);
      const double2* pub_rd = pub + (size_t)I.pa * 2 + r;
      double2* pub_wr = pub + (size_t)(a < EMAX ? a : 0) * 2 + r;
      const int pstride = EMAX * 2;
      double fz_next = 0.0;
      if (inB && I.pstat == PS_FROZEN) fz_next = frozen_at(I.pwb, Tb, I.wmask, I.plo, I.phi, I.same);
      for (int it = 0; it < iters; ++it) {
        const int t = Tb + it;
        const bool go = inB && t < te;
        if (go) {
          double pv;
          bool pchanged = false;
          if (I.pstat == PS_INE) {
            const double2 pb = pub_rd[(it & 1) * pstride];
            pv = (MODEL == POB_MODEL_CTC_MERGE_REPEATS && I.same) ? pb.y : pb.x;
            pchanged = pchg[(it & 1) * pstride + I.pa * 2 + r] != 0;
          } else if (I.pstat == PS_FROZEN) {
            pv = fz_next;
            if (t + 1 < te) fz_next = frozen_at(I.pwb, t + 1, I.wmask, I.plo, I.phi, I.same);
          } else if (I.pstat == PS_ROOT) pv = root_prob(r, t - 1);
          else pv = 0.0;
          const double yl = ylast, yb = yblank;
          if (t + 1 < te) load_y(I, t + 1, ylast, yblank);
          const bool must = computing || t >= cs || pchanged;
          double2 pb;
          if (must) {
            if (!computing) {
              // first recomputed timestep of a node that was clean so far: fetch its own values at t-1
              computing = true;
              p_prev = 0.0; ng_prev = 0.0;
              if (t - 1 >= I.lo && t - 1 < I.hi) {
                const Ent* se = I.wb + (t & I.wmask);
                const EV v = ld_ent(se);
                p_prev = v.prob; ng_prev = v.nogap;
              }
              if (t < cs) {
                // dirtied inside its clean range: the clean maximum may include entries that change now
                cs = t;
                maxv = 0.0; maxt = -1;
                for (int q = max(ts, I.lo); q < min(t, I.hi); ++q)
                  fold_max(maxv, maxt, in_band_scale(I, ld_prob(I.wb + ((q + 1) & I.wmask)), q), q);
              }
            }
            double prob, gp, ng;
            cell(p_prev, ng_prev, pv, yl, yb, prob, gp, ng);
            Ent* o = I.wb + ((t + 1) & I.wmask);
            st_ent(o, prob, gp, ng);
            if constexpr (MODEL == POB_MODEL_CTC_MERGE_REPEATS) ng_prev = ng;
            pb.x = prob; pb.y = gp;
            fold_max(maxv, maxt, in_band_scale(I, prob, t), t);
            p_prev = prob;
          } else {
            // still clean at t: hand the stored value to the children
            pb.x = 0.0; pb.y = 0.0;
            if (t >= I.lo && t < I.hi) {
              const Ent* se = I.wb + ((t + 1) & I.wmask);
              const EV v = ld_ent(se);
              pb.x = v.prob; pb.y = v.gap;
            }
          }
          pub_wr[((it + 1) & 1) * pstride] = pb;
          pchg[((it + 1) & 1) * pstride + a * 2 + r] = must;
        }
        __syncthreads();
      }
      PCLK(4);
    }
    int kmax = I.kref;  // scale of maxv
    if (on) {
      if (te > ts) {
        // window bookkeeping: after this sweep every entry of [ts, te) is current
        int lo = I.lo, hi = I.hi;
        if (ts > hi || ts < lo) { lo = ts; hi = te; }
        else hi = max(hi, te);
        if (hi - lo > g_es.cap[r]) lo = hi - g_es.cap[r];
        a_lo[2 * a + r] = lo; a_hi[2 * a + r] = hi;
        a_che[2 * a + r] = te;
        a_maxp[2 * a + r] = maxv;  // reset + max over the band
        a_maxk[2 * a + r] = kmax;
        a_maxt[2 * a + r] = maxt;
        if (a == 0) sh[SH_KREF0 + r] = kmax;  // for the debug trace only
      } else {
        // empty band: max_prob left stale (A.6b); brought to a scale every node shares (that of the newest column)
        kmax = sh[SH_KLAST0 + r];
        maxv = scale2(a_maxp[2 * a + r], a_maxk[2 * a + r] - kmax);
        if (a == 0) sh[SH_KREF0 + r] = kmax;
      }
    }
    // ranking keys: row_col = max0 * max1 (PrefixTree.h:111, :397); row = value(0, u) * max1 (:107, :393); every
    // factor is in a scale shared by all nodes of this step
    const double other = __shfl_xor_sync(0xffffffffu, maxv, 1);
    if (used) {
      if (mode == MODE_ROWCOL) {
        if (r == 0) set_key(a, no_nan(__dmul_rn(maxv, other)));
      } else if (r == 1) {
        set_key(a, no_nan(__dmul_rn(a_last0[a], maxv)));
      }
    }
    pre_prune();
    __syncthreads();
    PCLK(5);
  }

  // scalars of the coming prune / sweep, reset by thread 0 BEFORE the barrier that precedes prune()
  __device__ __forceinline__ void pre_prune() {
    POB_VIEWS
    if (threadIdx.x == 0) {
      sh[SH_NB] = min(sh[SH_NUSED], W); sh[SH_DMIN] = 0x7fffffff;
      sh[SH_NL0] = 0; sh[SH_NL1] = 0;
      sh[SH_PCNT] = 0; sh[SH_PM0] = 0; sh[SH_PM1] = 0; sh[SH_PM2] = 0; sh[SH_PM3] = 0;
    }
  }

  // Ranking key of an active slot: the score itself (a non-negative double in the scale shared by all nodes of the
  // step: its bit pattern orders like its value) and a 31-bit image of it: exponent relative to the top score of the
  // previous prune (+-127 binades, saturating) and 23 mantissa bits.  The map score -> k32 is monotone (not strict), so
  // ranks computed on k32 are exact whenever they come out distinct; prune() checks that and falls back to the exact
  // keys otherwise.
  __device__ __forceinline__ void set_key(int a, double score) {
    POB_VIEWS
    key[a] = score;
    long long k = (__double_as_longlong(score) >> 29) - ((long long)sh[SH_EBASE] << 23);
    k = max(0LL, min(k, 0x7fffffffLL));
    k32[a] = (unsigned)k;
  }
  __device__ __forceinline__ void clear_key(int a) {
    POB_VIEWS
    key[a] = -1.0;  // below every score: an unused slot never counts in a ranking
    k32[a] = 0;
  }
  // the top of the new beam sets the exponent base of the next step's 31-bit keys and asks for a change of scale of
  // the coming columns of a read when its values have drifted 64 binades away from 1
  __device__ __forceinline__ void top_feedback(int a) {
    POB_VIEWS
    const double sc = key[a];
    if (!(sc > 0.0)) return;
    // atomic stores: when the fast ranking ties at the top, two threads come here (the exact pass then runs, and its
    // single rank-0 thread has the last word)
    atomicExch(&sh[SH_EBASE], max(0, (int)((__double_as_longlong(sc) >> 52) & 0x7ff) - 127));
    for (int r = 0; r < 2; ++r) {
      double m; int k;
      if (mode == MODE_ROWCOL || (mode == MODE_ROW && r == 1)) { m = a_maxp[2 * a + r]; k = a_maxk[2 * a + r]; }
      else if (r == 0) { m = a_last0[a]; k = sh[SH_KCUR0]; }
      else continue;
      if (!(m > 0.0)) continue;
      const int e = (int)((__double_as_longlong(m) >> 52) & 0x7ff) - 1023 + (k - sh[SH_KLAST0 + r]);
      atomicExch(&sh[SH_PEND0 + r], (e >= 64 || e <= -64) ? e : 0);
    }
  }

  // exact ranking on the scores' bit patterns, ties by creation order: only when the fast pass saw equal 31-bit keys
  // inside the beam (two scores within 2^-23 relative, more than 127 binades from the reference, or fewer positive
  // candidates than W)
  __device__ __noinline__ void prune_exact() {
    POB_VIEWS
    const int tid = threadIdx.x;
    const bool two = (int)blockDim.x >= 2 * EMAX;
    const int a = two ? (tid >> 1) : tid;
    const int half = two ? (tid & 1) : 0;
    const bool cand = a < EMAX && a_slot[a] >= 0;
    const int mid = two ? (EMAX >> 1) : EMAX;
    const int j0 = half ? mid : 0, j1 = half ? EMAX : mid;
    int rank = 0;
    if (tid == 0) sh[SH_DMIN] = 0x7fffffff;
    if (cand) {
      const long long ks = __double_as_longlong(key[a]);
      const uint32_t ko = a_order[a];
      for (int j = j0; j < j1; ++j) {
        const long long s = __double_as_longlong(key[j]);
        rank += (s > ks) || (s == ks && a_order[j] < ko);
      }
    }
    if (two) rank += __shfl_xor_sync(0xffffffffu, rank, 1);
    __syncthreads();
    if (a < EMAX && half == 0) {
      const bool inb = cand && rank < W;
      a_inbeam[a] = inb;
      if (inb) {
        beam[rank] = a;
        atomicMin(&sh[SH_DMIN], a_depth[a]);
        if (rank == 0) top_feedback(a);
      }
    }
    if (tid == 0) atomicAdd(reinterpret_cast<unsigned long long*>(&g_exact_prunes), 1ULL);
    __syncthreads();
  }

  // ---- Beam::prune (Beam.h:93-108): rank by score desc, exact ties by creation order ----------
  // one barrier at the end (three more when the exact pass is needed); also clears the per-step flags
  __device__ __noinline__ void prune() {
    POB_VIEWS
    const int tid = threadIdx.x;
    const bool two = (int)blockDim.x >= 2 * EMAX;
    const int a = two ? (tid >> 1) : tid;
    const int half = two ? (tid & 1) : 0;
    const bool cand = a < EMAX && a_slot[a] >= 0;
    int rank = 0;
    // phase-split bounds of the NEXT sweep: every thread read this sweep's values before the barrier that precedes
    // this call (they must not be reset earlier: a sweep without a synchronised phase has no barrier between its
    // readers and its end), and the next sweep's atomicMin comes after the barriers of the expansion
    if (tid == 0) { sh[SH_TB0] = 0x7fffffff; sh[SH_TB1] = 0x7fffffff; }
    // all-pairs rank on the 31-bit keys, four per shared-memory load; two threads share a candidate's key range
    const int nq = E4 >> 2;
    const int qmid = two ? ((nq + 1) >> 1) : nq;
    const int q0 = half ? qmid : 0, q1 = half ? nq : qmid;
    if (cand) {
      const unsigned kh = k32[a];
      const uint4* const kv = reinterpret_cast<const uint4*>(k32);
      unsigned r0 = 0, r1 = 0, r2 = 0, r3 = 0;
#pragma unroll 4
      for (int q = q0; q < q1; ++q) {
        const uint4 v = kv[q];
        // keys are < 2^31: kh - kj has its top bit set exactly when kj > kh; mad.hi(d, 2, r) = r + (d >> 31) is one
        // instruction on the FMA pipe next to the subtraction on the ALU pipe
        r0 = sign_acc(kh - v.x, r0); r1 = sign_acc(kh - v.y, r1);
        r2 = sign_acc(kh - v.z, r2); r3 = sign_acc(kh - v.w, r3);
      }
      rank = (int)(r0 + r1 + r2 + r3);
    }
    if (two) rank += __shfl_xor_sync(0xffffffffu, rank, 1);
    // no barrier here: the ranking loop only reads k32[], the block below only writes other arrays, and the scalars
    // it updates were reset by pre_prune() before the barrier that precedes this call
    int dmin_mine = 0x7fffffff;
    bool inb = false;
    if (a < EMAX && half == 0) {
      a_needed[a] = 0;
      inb = cand && rank < W;
      a_inbeam[a] = inb;
      if (inb) {
        atomicExch(&beam[rank], a);  // two candidates may share a rank here (a tie in the 31-bit keys): the exact pass follows
        dmin_mine = a_depth[a];
        if (rank == 0) top_feedback(a);
      }
    }
    // the ranks inside the beam must be distinct: collect them as bit masks (and their count) per warp
    {
      const unsigned nin = __popc(__ballot_sync(0xffffffffu, inb));
      const int nw = (W + 31) >> 5;
      for (int w = 0; w < nw; ++w) {
        const unsigned m = __reduce_or_sync(0xffffffffu, (inb && (rank >> 5) == w) ? (1u << (rank & 31)) : 0u);
        if ((tid & 31) == 0 && m) atomicOr(reinterpret_cast<unsigned*>(&sh[SH_PM0 + w]), m);
      }
      if ((tid & 31) == 0 && nin) atomicAdd(&sh[SH_PCNT], (int)nin);
    }
    // minimum depth of the new beam: one shared-memory atomic per warp instead of one per beam node
    dmin_mine = __reduce_min_sync(0xffffffffu, dmin_mine);
    if ((tid & 31) == 0 && dmin_mine != 0x7fffffff) atomicMin(&sh[SH_DMIN], dmin_mine);
    __syncthreads();
    {
      const int cnt = sh[SH_PCNT];
      const int bits = __popc(sh[SH_PM0]) + __popc(sh[SH_PM1]) + __popc(sh[SH_PM2]) + __popc(sh[SH_PM3]);
      if (cnt != sh[SH_NB] || bits != cnt) prune_exact();
    }
    PCLK(6);
  }

  // ---- retire an active slot: write the header home, queue the node for reclamation ----
  __device__ void retire(int a) {
    POB_VIEWS
    const int slot = a_slot[a];
    NodeHdr& h = hdr[slot];
    h.tid = a_tid[a];
    for (int c = 0; c < 4; ++c) { h.kid_slot[c] = kid_slot(a_kid[4 * a + c]); h.kid_order[c] = a_kido[4 * a + c]; }
    h.lo[0] = a_lo[2 * a]; h.lo[1] = a_lo[2 * a + 1]; h.hi[0] = a_hi[2 * a]; h.hi[1] = a_hi[2 * a + 1];
    h.maxp[0] = a_maxp[2 * a]; h.maxp[1] = a_maxp[2 * a + 1]; h.maxk[0] = a_maxk[2 * a]; h.maxk[1] = a_maxk[2 * a + 1];
    const int stamp = atomicAdd(&sh[SH_STAMP], 1) + 1;
    h.state = stamp;
    const int pos = atomicAdd(&sh[SH_RQT], 1);
    retq[pos % RQ] = make_int2(slot, stamp);
    h.aslot = -1;
    // a live parent that stays in the expanded beam forgets where this child was active (a parent that retires in
    // this same phase writes only the pool slots home, see above)
    if (a_pstat[a] == PS_INE) {
      const int pa = a_par[a];
      if (a_needed[pa]) a_kid[4 * pa + a_last[a]] = slot;
    }
    a_slot[a] = -1;
    clear_key(a);
    a_free[atomicAdd(&sh[SH_AFREE], 1)] = a;
    atomicSub(&sh[SH_NUSED], 1);
  }

  // fill an active slot for a node that did not exist (fresh) or comes back from retirement (revive)
  __device__ void activate_fresh(int a, int slot, uint32_t order, int pa, int last, int parent_tid) {
    POB_VIEWS
    NodeHdr n;
    n.order = order; n.state = 0; n.parent_slot = a_slot[pa]; n.parent_order = a_order[pa];
    n.parent_tid = parent_tid; n.tid = -1; n.depth = a_depth[pa] + 1; n.last = last;
    for (int q = 0; q < 4; ++q) { n.kid_slot[q] = -1; n.kid_order[q] = 0; }
    n.lo[0] = n.lo[1] = 0; n.hi[0] = n.hi[1] = 0; n.maxp[0] = n.maxp[1] = 0.0; n.maxk[0] = n.maxk[1] = 0;
    n.aslot = a;
    hdr[slot] = n;
    a_slot[a] = slot; a_order[a] = order; a_par[a] = pa; a_pslot[a] = n.parent_slot; a_porder[a] = n.parent_order;
    a_depth[a] = n.depth; a_tid[a] = -1; a_ptid[a] = n.parent_tid;
    for (int q = 0; q < 4; ++q) { a_kid[4 * a + q] = -1; a_kido[4 * a + q] = 0; }
    a_lo[2 * a] = a_lo[2 * a + 1] = 0; a_hi[2 * a] = a_hi[2 * a + 1] = 0;
    a_maxp[2 * a] = a_maxp[2 * a + 1] = 0.0; a_maxk[2 * a] = a_maxk[2 * a + 1] = 0; a_last0[a] = 0.0;
    a_maxt[2 * a] = a_maxt[2 * a + 1] = -1;
    a_che[2 * a] = a_che[2 * a + 1] = -1;
    a_last[a] = (uint8_t)last; a_pstat[a] = PS_INE; a_same[a] = (a_last[pa] == last);
    a_inbeam[a] = 0; a_needed[a] = 1;
    clear_key(a);
  }

  __device__ void activate_revived(int a, int slot, int pa) {
    POB_VIEWS
    NodeHdr& h = hdr[slot];
    h.state = 0;
    h.aslot = a;
    a_slot[a] = slot; a_order[a] = h.order; a_par[a] = pa; a_pslot[a] = h.parent_slot; a_porder[a] = h.parent_order;
    a_depth[a] = h.depth; a_tid[a] = h.tid; a_ptid[a] = h.parent_tid;
    for (int q = 0; q < 4; ++q) {
      const int ks = h.kid_slot[q];
      int link = ks;
      a_kido[4 * a + q] = h.kid_order[q];
      if (ks >= 0) {
        // a child that stayed in the expanded beam (it is a beam member) reads this node live again; whether it did is
        // in its home record (state 0 = active, with its active slot)
        const NodeHdr& hk = hdr[ks];
        if (hk.order == h.kid_order[q] && hk.state == 0) {
          const int ka = hk.aslot;
          a_par[ka] = a; a_pstat[ka] = PS_INE;
          link = kid_pack(ks, ka);
        }
      }
      a_kid[4 * a + q] = link;
    }
    a_lo[2 * a] = h.lo[0]; a_lo[2 * a + 1] = h.lo[1]; a_hi[2 * a] = h.hi[0]; a_hi[2 * a + 1] = h.hi[1];
    a_maxp[2 * a] = h.maxp[0]; a_maxp[2 * a + 1] = h.maxp[1]; a_maxk[2 * a] = h.maxk[0]; a_maxk[2 * a + 1] = h.maxk[1];
    a_last0[a] = 0.0;
    a_maxt[2 * a] = a_maxt[2 * a + 1] = -1;
    a_che[2 * a] = a_che[2 * a + 1] = -1;  // retained entries are stale with respect to the live parent
    a_last[a] = (uint8_t)h.last; a_pstat[a] = PS_INE; a_same[a] = (a_last[pa] == h.last);
    a_inbeam[a] = 0; a_needed[a] = 1;
    clear_key(a);
  }

  // ---- expansion of the new beam + retirement + reclamation: three phases, three barriers ----------
  // dead0/dead1: every later read of read r is at an index >= dead_r.  The pool's free slots form a FIFO ring
  // (pops at SH_FQH, pushes at SH_FQT), so recycling and allocation can share a phase.  Thread 4b+c owns child
  // c of beam[b].
  __device__ __noinline__ void expand_and_retire(int dead0, int dead1) {
    POB_VIEWS
    const int tid = threadIdx.x;
    const int nb = sh[SH_NB];
    // -- phase X1: classify the children of the beam.  A retired child that comes back is marked active right
    //    away so that the queue inspection of the next phase sees its queue entry as stale.
    const int xb = tid >> 2, xc = tid & 3;
    const bool xmine = tid < 4 * nb && xc < nbase();  // alphabets of fewer than four letters leave child threads idle
    int a = -1, kind = KID_ACTIVE;
    if (xmine) {
      a = beam[xb];
      kind = KID_FRESH;
      const int link = a_kid[4 * a + xc];
      const int ks = kid_slot(link);
      if (ks >= 0) {
        const int ka = kid_act(link);
        if (ka >= 0) { kind = KID_ACTIVE; a_needed[ka] = 1; }
        else if (hdr[ks].order == a_kido[4 * a + xc]) { kind = KID_REVIVE; hdr[ks].state = 0; }  // line now in L1 for X3
      }
      if (xc == 0) a_needed[a] = 1;
    }
    // deterministic creation-order / trace-id offsets: prefix counts over the child threads (warp ballots + the
    // per-warp totals in shared memory).  Fresh children are numbered in thread order; a beam node gets its trace
    // id at its first expansion, numbered in beam order.
    const unsigned lane = tid & 31, wid = tid >> 5;
    const unsigned m_fresh = __ballot_sync(0xffffffffu, xmine && kind == KID_FRESH);
    const unsigned m_first = __ballot_sync(0xffffffffu, xmine && xc == 0 && a_tid[xmine ? a : 0] < 0);
    if (lane == 0) { tmpc[2 * wid] = __popc(m_fresh); tmpc[2 * wid + 1] = __popc(m_first); }
    const int head = sh[SH_RQH], tail = sh[SH_RQT];
    const int fq_head = sh[SH_FQH], fq_tail = sh[SH_FQT];
    const int nfree = fq_tail - fq_head;
    // Retired nodes' headers are cold (HBM): look at the queue only every few expansions, unless slots run short
    const int xstep = sh[SH_XSTEP];
    const bool inspect = !g_es.noreclaim && (xstep % g_es.inspect_every == 0 || nfree < 16 * W + 32);
    const int navail = inspect ? min(tail - head, (int)blockDim.x) : 0;
    const int dmin = sh[SH_DMIN];
    // pool pressure: recycle live retirees too (flagged; the reference never frees anything)
    const bool force = nfree < 8 * W + 16 && navail > 0;
    __syncthreads();
    PCLK(7);
    // -- phase X2: retire what the next expanded beam does not contain, freeze orphaned children, inspect the
    //    retire queue (entries queued in earlier steps only: a node retired now is still readable); the child
    //    threads compute their deterministic creation-order / trace-id offsets
    if (tid < EMAX && a_slot[tid] >= 0) {
      if (!a_needed[tid]) retire(tid);
      else if (a_pstat[tid] == PS_INE && !a_needed[a_par[tid]]) {
        const int pa = a_par[tid];
        a_plo[2 * tid] = a_lo[2 * pa]; a_plo[2 * tid + 1] = a_lo[2 * pa + 1];
        a_phi[2 * tid] = a_hi[2 * pa]; a_phi[2 * tid + 1] = a_hi[2 * pa + 1];
        a_pstat[tid] = PS_FROZEN;
      }
    }
    int st = 0, rslot = -1, rstamp = 0;  // 1 stale, 2 dead + freeable, 3 dead but kept, 4 alive
    if (tid < navail) {
      const int2 q = retq[(head + tid) % RQ];
      rslot = q.x; rstamp = q.y;
      const NodeHdr& h = hdr[rslot];
      if (h.state != rstamp) st = 1;
      else if (!(h.hi[0] <= dead0 && h.hi[1] <= dead1)) st = 4;
      else if (h.depth <= dmin) st = 2;
      else {
        bool kids = false;
        for (int c = 0; c < 4; ++c) {
          const int ks = h.kid_slot[c];
          if (ks >= 0 && hdr[ks].order == h.kid_order[c]) kids = true;
        }
        st = kids ? 3 : 2;
      }
      if (st == 4) atomicMin(&sh[SH_FIRSTALIVE], tid);
    }
    int obase = 0, fbase = 0, first = 0;
    if (xmine) {
      obase = __popc(m_fresh & ((1u << lane) - 1u));                 // fresh children created before mine
      fbase = __popc(m_first & ((1u << (lane & ~3u)) - 1u));         // beam nodes first expanded before mine
      first = (m_first >> (lane & ~3u)) & 1u;
      for (unsigned w = 0; w < wid; ++w) { obase += tmpc[2 * w]; fbase += tmpc[2 * w + 1]; }
    }
    __syncthreads();
    PCLK(8);
    // -- phase X3: consume the inspected queue entries (strictly from the head, up to the first live one);
    //    child threads create / revive the missing children
    {
      int fa = min(sh[SH_FIRSTALIVE], navail);
      if (force) { fa = navail; if (tid == 0) sh[SH_STATUS] |= POB_ST_POOL_OVERFLOW; }
      if (tid < fa) {
        if (st == 2 || (force && st >= 2)) {
          NodeHdr& h = hdr[rslot];
          h.order = 0; h.state = -1;
          freelist[atomicAdd(&sh[SH_FQT], 1) % NP] = rslot;
        } else if (st == 3 || st == 4) {
          retq[atomicAdd(&sh[SH_RQT], 1) % RQ] = make_int2(rslot, rstamp);  // look again one queue cycle later
        }
      }
      if (tid == 0) { sh[SH_RQH] = head + fa; sh[SH_XSTEP] = xstep + 1; }
    }
    if (xmine) {
      const int my_tid = first ? sh[SH_TID] + fbase : a_tid[a];  // trace id of the beam node (my parent)
      if (tid == 4 * (nb - 1) + nbase() - 1) { sh[SH_TOTALLOC] = obase + (kind == KID_FRESH); sh[SH_TOTFIRST] = fbase + first; }
      if (kind != KID_ACTIVE) {
        const int ai = atomicSub(&sh[SH_AFREE], 1) - 1;
        int pi = -1;
        if (kind == KID_FRESH) pi = atomicAdd(&sh[SH_FQH], 1);
        if (ai < 0 || (kind == KID_FRESH && pi >= fq_tail)) {
          // cannot happen while the reclamation keeps its margin; refuse to corrupt memory if it does
          atomicOr(&sh[SH_STATUS], POB_ST_POOL_OVERFLOW);
          atomicAdd(&sh[SH_AFREE], 1);
          if (kind == KID_FRESH) atomicSub(&sh[SH_FQH], 1);
        } else {
          const int na = a_free[ai];
          if (kind == KID_FRESH) {
            const int slot = freelist[pi % NP];
            const uint32_t order = (uint32_t)(sh[SH_ORDER] + obase);
            activate_fresh(na, slot, order, a, xc, my_tid);
            a_kid[4 * a + xc] = kid_pack(slot, na); a_kido[4 * a + xc] = order;
          } else {
            const int ks = kid_slot(a_kid[4 * a + xc]);
            activate_revived(na, ks, a);
            a_kid[4 * a + xc] = kid_pack(ks, na);
          }
          atomicAdd(&sh[SH_NUSED], 1);
        }
      }
      // the first expansion of a beam node gives it its trace id; its sibling threads took the id from `first` and
      // fbase above and do not read a_tid[a] in that case, so it can be written in this phase
      if (xc == 0 && first) {
        a_tid[a] = my_tid;
        trace[my_tid] = ((uint32_t)a_ptid[a] << 2) | (uint32_t)a_last[a];
      }
    }
    __syncthreads();
    PCLK(9);
  }

  // ROW traversal while the beam is shorter than W (first row): the reference walks b < beam_width over a
  // list that grows as children are pushed (BeamSearch.h:132-144), i.e. a breadth-first closure.  Sequential,
  // runs once per item.
  __device__ __noinline__ void expand_bfs() {
    POB_VIEWS
    if (threadIdx.x == 0) {
      int* list = tmpa;  // EMAX >= 4 + 4W entries
      int n = 0;
      const int nb = sh[SH_NB];
      for (int b = 0; b < nb; ++b) list[n++] = beam[b];
      for (int k = 0; k < W && k < n; ++k) {
        const int a = list[k];
        if (a_tid[a] < 0) {
          a_tid[a] = sh[SH_TID]++;
          trace[a_tid[a]] = ((uint32_t)a_ptid[a] << 2) | (uint32_t)a_last[a];
        }
        for (int c = 0; c < nbase(); ++c) {
          const int ks = kid_slot(a_kid[4 * a + c]);
          int ka = kid_act(a_kid[4 * a + c]);
          if (ka < 0 && sh[SH_AFREE] > 0 && sh[SH_FQT] - sh[SH_FQH] > 0) {
            ka = a_free[--sh[SH_AFREE]];
            if (ks >= 0 && hdr[ks].order == a_kido[4 * a + c]) {
              activate_revived(ka, ks, a);
              a_kid[4 * a + c] = kid_pack(ks, ka);
            } else {
              const int slot = freelist[(sh[SH_FQH]++) % NP];
              const uint32_t order = (uint32_t)sh[SH_ORDER]++;
              activate_fresh(ka, slot, order, a, c, a_tid[a]);
              a_kid[4 * a + c] = kid_pack(slot, ka); a_kido[4 * a + c] = order;
            }
            sh[SH_NUSED]++;
          }
          if (ka >= 0 && n < EMAX) list[n++] = ka;
        }
      }
    }
    __syncthreads();
  }

  __device__ void dbg_record(const BeamParams& G, long step) {
    POB_VIEWS
    if (!G.dbg_trace) return;
    if (threadIdx.x == 0 && step < 50000) {
      double sum = 0;
      for (int b = 0; b < sh[SH_NB]; ++b) {
        // back to the reference's log domain: the key is a product of factors in the scales KREF0 / KREF1 (KCUR0 for
        // the single-column factor of the 1D and ROW searches)
        const int kk = (mode == MODE_ROWCOL) ? sh[SH_KREF0] + sh[SH_KREF1]
                                             : sh[SH_KCUR0] + (mode == MODE_ROW ? sh[SH_KREF1] : 0);
        const double sc = log(key[beam[b]]) + 0.6931471805599453 * kk;
        if (b == 0) G.dbg_trace[2 * step] = sc;
        if (sc > -1e300) sum += sc;
      }
      G.dbg_trace[2 * step + 1] = sum;
    }
  }

  __device__ void run_item(const BeamParams& G, int item, char* ws);
};

template <int MODEL, int EM_CT, int W_CT, int MODE_CT>
__device__ void Engine<MODEL, EM_CT, W_CT, MODE_CT>::run_item(const BeamParams& G, int item, char* ws) {
  const int tid = threadIdx.x, NT = blockDim.x;
  unsigned long long n_updates = 0;
#ifdef POB_PHASE_CLOCKS
  if (tid == 0) g_pclk_last = clock64();
#endif
  // ---- engine state of this item: scalars + views of the CTA's global workspace
  if (tid == 0) {
    g_es.W = G.W; g_es.NP = G.NP; g_es.RQ = G.RQ; g_es.EMAX = G.EMAX; g_es.mode = G.mode;
    g_es.noreclaim = G.dbg_noreclaim; g_es.inspect_every = G.inspect_every; g_es.longq = G.dbg_long;
    g_es.nbase = G.r[0].n_states - 1;
    g_es.cap[0] = G.CAP0; g_es.cap[1] = G.CAP1; g_es.mask[0] = G.CAP0 - 1; g_es.mask[1] = G.CAP1 - 1;
    g_es.rv[0] = make_view(G.r[0], item);
    if (G.mode != MODE_1D) g_es.rv[1] = make_view(G.r[1], item); else { g_es.rv[1] = g_es.rv[0]; g_es.rv[1].T = 0; }
    char* q = ws;
    g_es.hdr = (NodeHdr*)q; q += sizeof(NodeHdr) * (size_t)G.NP;
    g_es.win[0] = q; q += sizeof(Ent) * (size_t)G.NP * G.CAP0;
    g_es.win[1] = q; q += sizeof(Ent) * (size_t)G.NP * G.CAP1;
    g_es.freelist = (int32_t*)q; q += 4 * (size_t)G.NP;
    g_es.retq = (int2*)q; q += 8 * (size_t)G.RQ;
    if (G.col_off >= 0) {
      g_es.col[0] = (Col*)(pob_smem + G.col_off);
      g_es.col[1] = g_es.col[0] + G.CAPC0;
    } else {
      g_es.col[0] = (Col*)q; q += sizeof(Col) * (size_t)G.CAPC0;
      g_es.col[1] = (Col*)q; q += sizeof(Col) * (size_t)G.CAPC1;
    }
    g_es.cmask[0] = G.CAPC0 - 1; g_es.cmask[1] = G.CAPC1 - 1;
    g_es.sufmin = (int32_t*)q;
    g_es.trace = G.trace + G.trace_off[item];
  }
  __syncthreads();
  POB_VIEWS
  const int U = g_es.rv[0].T, V = g_es.rv[1].T;
  int32_t* otop = G.out_top + 4 * (size_t)item;

  // ---- init pool and active slots
  for (int s = tid; s < NP; s += NT) {
    hdr[s].order = 0; hdr[s].state = -1;
    freelist[s] = s;  // FIFO ring of free pool slots
  }
  for (int a = tid; a < EMAX; a += NT) {
    a_slot[a] = -1; a_free[a] = EMAX - 1 - a; clear_key(a);
    a_inbeam[a] = 0; a_needed[a] = 0;
  }
  for (int i = EMAX + tid; i < E4; i += NT) k32[i] = 0;  // padding of the last key quad
  if (tid == 0) {
    for (int k = 0; k < SH_COUNT; ++k) sh[k] = 0;
    sh[SH_FQH] = 0; sh[SH_FQT] = NP; sh[SH_AFREE] = EMAX; sh[SH_ORDER] = 1; sh[SH_TID] = 1;
    sh[SH_DMIN] = 0x7fffffff; sh[SH_FIRSTALIVE] = 0x7fffffff;
    sh[SH_TB0] = 0x7fffffff; sh[SH_TB1] = 0x7fffffff;
  }
  __syncthreads();
  if (U <= 0 || (mode != MODE_1D && V <= 0)) {
    if (tid == 0) { otop[0] = 0; otop[1] = -1; otop[2] = 0; otop[3] = POB_ST_EMPTY; G.out_score[item] = 0; }
    __syncthreads();
    return;
  }
  // ---- seed: the 4 children of the root, updated at t = 0 (BeamSearch.h:24-30, :287-293)
  const int nbase = g_es.rv[0].S - 1;
  if (tid < nbase) {
    const int a = tid, slot = tid;  // the first pool slots and active slots go to the root's children
    NodeHdr n;
    n.order = 1 + tid; n.state = 0; n.parent_slot = -1; n.parent_order = 0; n.parent_tid = 0; n.tid = -1;
    n.depth = 1; n.last = tid;
    for (int q = 0; q < 4; ++q) { n.kid_slot[q] = -1; n.kid_order[q] = 0; }
    n.lo[0] = n.lo[1] = 0; n.hi[0] = n.hi[1] = 0; n.maxp[0] = n.maxp[1] = 0.0; n.maxk[0] = n.maxk[1] = 0;
    n.aslot = a;
    hdr[slot] = n;
    a_slot[a] = slot; a_order[a] = n.order; a_par[a] = -1; a_pslot[a] = -1; a_porder[a] = 0; a_depth[a] = 1;
    a_tid[a] = -1; a_ptid[a] = 0;
    for (int q = 0; q < 4; ++q) { a_kid[4 * a + q] = -1; a_kido[4 * a + q] = 0; }
    a_lo[2 * a] = a_lo[2 * a + 1] = 0; a_hi[2 * a] = a_hi[2 * a + 1] = 0;
    a_maxp[2 * a] = a_maxp[2 * a + 1] = 0.0; a_maxk[2 * a] = a_maxk[2 * a + 1] = 0; a_last0[a] = 0.0;
    a_maxt[2 * a] = a_maxt[2 * a + 1] = -1;
    a_che[2 * a] = a_che[2 * a + 1] = -1;
    a_last[a] = (uint8_t)tid; a_pstat[a] = PS_ROOT; a_same[a] = 0; a_inbeam[a] = 1; a_needed[a] = 1;
    beam[tid] = a;
    n_updates += (mode != MODE_1D) ? 2 : 1;
  }
  if (tid == 0) {
    sh[SH_FQH] = nbase; sh[SH_AFREE] = EMAX - nbase; sh[SH_ORDER] = 1 + nbase; sh[SH_NB] = nbase;
    sh[SH_NUSED] = nbase;
  }
  __syncthreads();
  {
    const double p = update_all(tid < nbase, tid, 0, 0);
    if (mode == MODE_1D && tid < nbase) set_key(tid, no_nan(p));
    if (mode != MODE_1D) update_all(tid < nbase, tid, 1, 0);
  }

  PCLK(0);  // item setup + seed
  const int32_t* env = G.env ? G.env + 2 * G.env_off[item] : nullptr;
  const int32_t* envt = G.envt ? G.envt + 2 * G.envt_off[item] : nullptr;
  long nsteps = 0;
  if (mode == MODE_ROW && env && tid < 32) {
    // suffix minimum of the band starts, warp-chunked from the end
    int carry = 0x7fffffff;
    for (int base = ((U - 1) / 32) * 32; base >= 0; base -= 32) {
      const int i = base + tid;
      int v = (i < U) ? max(env[2 * i], 0) : 0x7fffffff;
      for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_down_sync(0xffffffffu, v, d);
        if (tid + d < 32) v = min(v, o);
      }
      v = min(v, carry);
      if (i < U) g_es.sufmin[i] = v;
      carry = __shfl_sync(0xffffffffu, v, 0);
    }
  }
  __syncthreads();

  if (mode == MODE_1D) {
    // BeamSearch.h:33-53
    expand_and_retire(-1, 0x7fffffff);
    for (int t = 1; t < U; ++t) {
      const bool mine = tid < EMAX && a_slot[tid] >= 0;
      const double p = update_all(mine, tid, 0, t);
      if (mine) {
        n_updates++;
        set_key(tid, no_nan(p));  // last_probability(): value at the last t
      }
      pre_prune();
      __syncthreads();
      prune();
      dbg_record(G, nsteps);
      expand_and_retire(t, 0x7fffffff);  // the next step reads index t
      ++nsteps;
    }
  } else if (mode == MODE_ROW) {
    // BeamSearch.h:127-167 (envelope) / :196-247 (no envelope: rows start at 1, band = [0, V))
    if (sh[SH_NB] < W) expand_bfs(); else expand_and_retire(-1, -1);
    for (int u = env ? 0 : 1; u < U; ++u) {
      int rs = env ? env[2 * u] : 0, re = env ? env[2 * u + 1] : V;
      rs = max(rs, 0); re = min(re, V);
      {
        const bool mine = tid < EMAX && a_slot[tid] >= 0;
        update_all(mine, tid, 0, u);
        if (mine) n_updates++;
      }
      sweep(2, u, u + 1, rs, re, true, n_updates);
      prune();
      dbg_record(G, nsteps);
      // read 0 is next read at index u; read 1 at >= (smallest band start of any later row) - 1
      const int nrs = (u + 1 < U) ? (env ? g_es.sufmin[u + 1] : 0) : 0x7ffffffe;
      if (sh[SH_NB] < W) expand_bfs(); else expand_and_retire(u, nrs - 1);
      ++nsteps;
    }
  } else {
    // BeamSearch.h:295-392
    int u = 0, v = 0;
    bool have_E = false;
    // The band of row u / column v is read by every thread at the top of every step.  Thread 0 copies the entries
    // of the next row and the next column into shared memory with cp.async while the current step runs (two slots
    // each, indexed by parity), so that the next step starts from shared memory instead of a global load.
    int* const senv = sh + SH_ENV;    // [2][2] band of row u  (slot u & 1)
    int* const senvt = sh + SH_ENVT;  // [2][2] band of column v (slot v & 1)
    if (tid == 0) { senv[0] = env[0]; senv[1] = env[1]; senvt[0] = envt[0]; senvt[1] = envt[1]; }
    __syncthreads();
    // the staging and prefetching below is done by the last threads of the block: they own no (node, read) item, so
    // the extra instructions stay off the warps whose work the step's barriers wait for
    const int aux = NT - 1;
    while (u <= U - 1 && v <= V - 1) {
      const int ers = senv[2 * (u & 1)], ere = senv[2 * (u & 1) + 1], ecs = senvt[2 * (v & 1)], ece = senvt[2 * (v & 1) + 1];
      if (tid == aux) {
        if (u + 1 < U) cp_async8(senv + 2 * ((u + 1) & 1), env + 2 * (u + 1));
        if (v + 1 < V) cp_async8(senvt + 2 * ((v + 1) & 1), envt + 2 * (v + 1));
        cp_async_commit();
      }
      int row_start = v, row_end = v, col_start = u, col_end = u;
      bool rset = false, cset = false;
      if (v >= ers && v < ere) { row_end = ere; rset = true; }
      else if (v < ers) {
        const int nb = sh[SH_NB];
        if (nb < W && tid == 0) sh[SH_STATUS] |= POB_ST_SHORT_BEAM_SKIP;
        const bool mine = tid < EMAX && a_slot[tid] >= 0 && a_inbeam[tid];
        if (tid == aux) cp_async_wait_all();  // visible to everyone after the barriers inside update_all
        update_all(mine, tid, 1, v);
        if (mine) n_updates++;
        ++v; ++nsteps;
        continue;
      }
      if (u >= ecs && u < ece) { col_end = ece; cset = true; }
      else if (u < ecs) {
        const int nb = sh[SH_NB];
        if (nb < W && tid == 0) sh[SH_STATUS] |= POB_ST_SHORT_BEAM_SKIP;
        const bool mine = tid < EMAX && a_slot[tid] >= 0 && a_inbeam[tid];
        if (tid == aux) cp_async_wait_all();
        update_all(mine, tid, 0, u);
        if (mine) n_updates++;
        ++u; ++nsteps;
        continue;
      }
      if ((!rset || !cset) && tid == 0) sh[SH_STATUS] |= POB_ST_UNSET_BAND;
      row_end = min(row_end, V); col_end = min(col_end, U);
      if (!have_E) { expand_and_retire(-1, -1); have_E = true; }
      // The probability rows and envelope entries of the coming steps are pulled towards the SM ahead of time: every
      // step touches one new row per read, and without this each thread of the sweep waits for it to come from HBM.
      if (G.prefetch) {
        const int k = tid - (NT - 8);  // threads NT-8 .. NT-3
        if (k >= 0 && k < 4) {
          const int r = k & 1;
          const ReadView& pv = g_es.rv[r];
          const int t = (r ? row_end : col_end) + ((k & 2) ? 64 : 6);
          if (t < pv.T) {
            const char* q = (const char*)pv.base + (size_t)pv.prow(t) * pv.S * (pv.f64 ? 8 : 4);
            if (k & 2) prefetch_l2(q); else prefetch_l1(q);
          }
        } else if (k >= 4 && k < 6) {
          const int ahead = 32;
          if (k == 4 && u + ahead < U) prefetch_l1(env + 2 * (u + ahead));
          if (k == 5 && v + ahead < V) prefetch_l1(envt + 2 * (v + ahead));
        }
      }
      sweep(3, col_start, col_end, row_start, row_end, G.dbg_noreuse != 0, n_updates);
      prune();
      dbg_record(G, nsteps);
      if (tid == aux) cp_async_wait_all();  // next row / column bands: visible after the barriers of the expansion
      expand_and_retire(u, v);  // later reads are at t-1 >= u (read 0) and >= v (read 1)
      ++u; ++v; ++nsteps;
    }
  }
  __syncthreads();
  if (tid == 0) {
    const int a = beam[0];
    // back to the reference's log domain: log(value) + K ln 2 per factor (log(0) = -inf, as in the reference)
    const double LN2 = 0.6931471805599453;
    double sc;
    if (mode == MODE_1D) sc = log(a_last0[a]) + LN2 * sh[SH_KCUR0];
    else if (mode == MODE_ROW) sc = (log(a_last0[a]) + LN2 * sh[SH_KCUR0]) + (log(a_maxp[2 * a + 1]) + LN2 * a_maxk[2 * a + 1]);
    else sc = (log(a_maxp[2 * a]) + LN2 * a_maxk[2 * a]) + (log(a_maxp[2 * a + 1]) + LN2 * a_maxk[2 * a + 1]);
    G.out_score[item] = sc;
    if (a_tid[a] >= 0) { otop[0] = a_tid[a]; otop[1] = -1; }
    else { otop[0] = a_ptid[a]; otop[1] = a_last[a]; }
    otop[2] = a_depth[a];
    otop[3] = sh[SH_STATUS];
  }
  // per-item counters
  for (int o = 16; o > 0; o >>= 1) n_updates += __shfl_down_sync(0xffffffffu, n_updates, o);
  if ((tid & 31) == 0 && n_updates) atomicAdd(&G.counters[0], n_updates);
  if (tid == 0) atomicAdd(&G.counters[1], (unsigned long long)nsteps);
  __syncthreads();
}

template <int MODEL, int MAXT, int MINB, int EM_CT = 0, int W_CT = 0, int MODE_CT = -1>
__global__ void __launch_bounds__(MAXT, MINB) beam_kernel(BeamParams P) {
  __shared__ int s_item;
  char* ws = P.ws + (size_t)blockIdx.x * P.ws_stride;
  Engine<MODEL, EM_CT, W_CT, MODE_CT> eng;
  for (;;) {
    if (threadIdx.x == 0) s_item = atomicAdd(P.work_counter, 1);
    __syncthreads();
    const int k = s_item;
    __syncthreads();
    if (k >= P.n_items) break;
    const int item = P.order ? P.order[k] : k;
    if (P.skip && P.skip[item]) {
      if (threadIdx.x == 0) {
        int32_t* otop = P.out_top + 4 * (size_t)item;
        otop[0] = 0; otop[1] = -1; otop[2] = 0; otop[3] = 0;
        P.out_score[item] = 0;
      }
      continue;
    }
    eng.run_item(P, item, ws);
  }
}

// label of the returned node: walk the trace ids back to the root (PrefixTree.h:449-457)
__global__ void backtrace_kernel(const uint32_t* __restrict__ trace, const int64_t* __restrict__ trace_off,
                                 const int32_t* __restrict__ top, const int64_t* __restrict__ out_off, int n,
                                 uint8_t* __restrict__ out_seq, int32_t* __restrict__ out_len,
                                 int32_t* __restrict__ out_status) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t* tr = trace + trace_off[i];
  int tid = top[4 * i], extra = top[4 * i + 1], depth = top[4 * i + 2];
  uint8_t* out = out_seq + out_off[i];
  int pos = depth - 1;
  if (extra >= 0 && pos >= 0) out[pos--] = (uint8_t)("ACGT"[extra & 3]);
  while (tid > 0 && pos >= 0) {
    const uint32_t w = tr[tid];
    out[pos--] = (uint8_t)("ACGT"[w & 3]);
    tid = (int)(w >> 2);
  }
  out_len[i] = depth;
  if (out_status) out_status[i] |= top[4 * i + 3];
}

}  // namespace
// debug export (not part of the ABI): cycles per engine phase, only in builds with -DPOB_PHASE_CLOCKS
extern "C" int pob_debug_exact_prunes(unsigned long long* out, int reset) {
#ifdef POB_COUNT_RESCAN
  {
    unsigned long long d[8];
    POB_CUDA(cudaMemcpyFromSymbol(d, g_dbg, sizeof(d)));
    fprintf(stderr, "[pob] scans: unknown %llu, scale %llu, left the band %llu, other %llu; maximum found at first entry %llu, "
            "first four %llu, elsewhere %llu; all-zero %llu\n", d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7]);
  }
#endif
  POB_CUDA(cudaMemcpyFromSymbol(out, g_exact_prunes, sizeof(unsigned long long)));
  if (reset) {
    unsigned long long z = 0;
    POB_CUDA(cudaMemcpyToSymbol(g_exact_prunes, &z, sizeof(z)));
  }
  return POB_OK;
}
extern "C" int pob_debug_phase_clocks(unsigned long long* out32, int reset) {
#ifdef POB_PHASE_CLOCKS
  POB_CUDA(cudaMemcpyFromSymbol(out32, g_phase_clk, 32 * sizeof(unsigned long long)));
  if (reset) {
    unsigned long long z[32] = {0};
    POB_CUDA(cudaMemcpyToSymbol(g_phase_clk, z, sizeof(z)));
  }
  return POB_OK;
#else
  return POB_EUNSUPPORTED;
#endif
}
double* g_pob_dbg_trace = nullptr;
extern "C" int pob_debug_trace(pob_ctx* ctx, double* out, int n) {
  if (!g_pob_dbg_trace) return POB_EINVAL;
  POB_CUDA(cudaMemcpy(out, g_pob_dbg_trace, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
  return POB_OK;
}
namespace {

size_t ws_bytes(int model, int NP, int CAP0, int CAP1, int RQ, int CAPC0, int CAPC1, int Umax) {
  const size_t es = model == POB_MODEL_CTC ? sizeof(Entry<POB_MODEL_CTC>) : sizeof(Entry<POB_MODEL_CTC_MERGE_REPEATS>);
  size_t b = sizeof(NodeHdr) * (size_t)NP + es * (size_t)NP * ((size_t)CAP0 + CAP1) + 4 * (size_t)NP + 8 * (size_t)RQ;
  b += sizeof(Col) * ((size_t)CAPC0 + CAPC1);
  b += 4 * ((size_t)Umax + 2);
  return pob_align_up(b, 256);
}

size_t smem_bytes(int W, int NP, int EMAX) {
  // must match POB_VIEWS
  size_t b = 224 * (size_t)EMAX + 4 * ((W + 3) & ~3) + 4 * SH_COUNT + 5 * ((EMAX + 15) & ~15);
  b += 4 * (size_t)((EMAX + 3) & ~3);  // k32
  b += 8 * (size_t)EMAX;               // a_maxt
  return pob_align_up(b, 16);
}

int pow2_at_least(int x) {
  int p = 4;
  while (p < x) p <<= 1;
  return p;
}

}  // namespace

// Search launcher on device pointers.  Host-side knowledge needed: the widest band of any item
// (max_span0 over read 0 time, max_span1 over read 1 time) and the longest reads, to size the windows.
int pob_beam_launch(pob_ctx* ctx, const pob_reads& r1, const pob_reads* r2, const int32_t* env,
                    const int64_t* env_off, const int32_t* envt, const int64_t* envt_off, const int32_t* order,
                    const int32_t* skip, int n_items, int n_total, int W, int model, int mode, int max_span0,
                    int max_span1, int Umax, int Vmax, const int64_t* trace_off, uint32_t* trace, int32_t* top,
                    const int64_t* out_off, uint8_t* out_seq, int32_t* out_len, double* out_score,
                    int32_t* out_status) {
  if (n_total <= 0) return POB_OK;
  if (W < 1) return POB_EINVAL;
  if (W > 100) return POB_EUNSUPPORTED;  // one thread per (node, read) of the expanded beam: 2 * 5 W <= 1024
  if (r1.n_states < 2 || r1.n_states > 5 || (r2 && r2->n_states != r1.n_states)) return POB_EUNSUPPORTED;
  BeamParams P;
  memset(&P, 0, sizeof(P));
  P.r[0] = r1;
  if (r2) P.r[1] = *r2;
  P.env = env; P.env_off = env_off; P.envt = envt; P.envt_off = envt_off; P.order = order; P.skip = skip;
  P.n_items = n_items; P.W = W; P.mode = mode;
  // active slots: the expanded beam (W nodes and their 4 W children), rounded so that the (node, read) threads fill
  // whole warps (W = 25: 128 slots, 256 threads, 85 registers per thread at three CTAs per SM)
  // (the first step expands all four seeds whatever W is: at least 4 + 16 slots)
  P.EMAX = ((5 * W > 20 ? 5 * W : 20) + 15) / 16 * 16;
  // Node pool: the expanded beam (5W) plus retired nodes whose windows can still be read.  About W nodes
  // retire per step and stay readable for one band width, so the pool scales with W x widest band; an
  // overflow is flagged per item (POB_ST_POOL_OVERFLOW), never silent.
  {
    const int span = (mode == MODE_1D) ? 1 : (max_span0 > max_span1 ? max_span0 : max_span1);
    long want = 64L * W;
    const long by_span = 3L * W * (span + 8) + 16L * W;
    if (by_span > want) want = by_span;
    if (want > 65536) want = 65536;
    P.NP = (int)((want + 255) / 256 * 256);  // the pool's rings are indexed modulo NP: no power of two needed
    if (P.NP < 1024) P.NP = 1024;
  }
  if (const char* e = getenv("POB_DEBUG_NP")) P.NP = atoi(e);
  if (const char* e = getenv("POB_DEBUG_NORECLAIM")) P.dbg_noreclaim = atoi(e);
  if (const char* e = getenv("POB_DEBUG_NOREUSE")) P.dbg_noreuse = atoi(e);
  if (const char* e = getenv("POB_DEBUG_LONG")) P.dbg_long = atoi(e);
  P.inspect_every = 0;  // set below, once the block size is known
  if (const char* e = getenv("POB_DEBUG_INSPECT_EVERY")) P.inspect_every = atoi(e) > 0 ? atoi(e) : 1;
  P.prefetch = 1;
  if (const char* e = getenv("POB_DEBUG_PREFETCH")) P.prefetch = atoi(e);
  if (getenv("POB_DEBUG_TRACE")) {
    static double* dbg = nullptr;
    if (!dbg) cudaMalloc(&dbg, 200000 * sizeof(double));
    cudaMemsetAsync(dbg, 0, 200000 * sizeof(double), ctx->stream);
    P.dbg_trace = dbg;
    g_pob_dbg_trace = dbg;
  }
  P.CAP0 = pow2_at_least(max_span0 + 3);
  P.CAP1 = pow2_at_least(max_span1 + 3);
  // column records: a ring over the widest band plus the look-ahead; the ROW traversal may come back to any earlier
  // timestep of read 2 (its band starts are not monotone), so there it holds the whole read
  P.CAPC0 = pow2_at_least(max_span0 + COL_LOOK + 4);
  P.CAPC1 = pow2_at_least((mode == MODE_ROW ? Vmax + 1 : max_span1) + COL_LOOK + 4);
  if (P.CAPC0 < 16) P.CAPC0 = 16;
  if (P.CAPC1 < 16) P.CAPC1 = 16;
  P.RQ = P.NP * 2;
  // the band sweep wants one thread per (node, read); the single-read search one per node
  int threads = (((mode == MODE_1D ? 1 : 2) * P.EMAX + 31) / 32) * 32;
  if (threads > 1024) return POB_EUNSUPPORTED;
  if (threads < 64) threads = 64;
  if (P.inspect_every == 0) {
    // about W nodes retire per step and one inspection looks at up to `threads` queue entries: keep the interval
    // below threads / W so that the queue does not back up (a short pool falls back to every step anyway)
    P.inspect_every = threads / (W + 8);
    // measured: 16 instead of 8 is +1.4 % (the queue then backs up a little and is worked off when free slots run
    // short, where the inspection runs every step anyway); 32 and 64 add nothing
    P.inspect_every *= 2;
    if (P.inspect_every > 16) P.inspect_every = 16;
    if (P.inspect_every < 1) P.inspect_every = 1;
  }
  size_t smem = smem_bytes(W, P.NP, P.EMAX);
  int max_cta_sm = 0;  // 0 = as many as fit (three CTAs of 256 / 288 threads per SM)
  if (const char* e = getenv("POB_DEBUG_CTA_PER_SM")) max_cta_sm = atoi(e);
  // column records (64 B per timestep of the band ring) in shared memory when they fit next to the rest
  P.col_off = -1;
  {
    const size_t budget_cta = (max_cta_sm == 2 ? 227 * 1024 / 2 : max_cta_sm == 1 ? 200 * 1024 : 227 * 1024 / 3) - 1024;
    const size_t cb = sizeof(Col) * ((size_t)P.CAPC0 + P.CAPC1);
    bool on = pob_align_up(smem, 64) + cb <= budget_cta;
    on = false;  // measured slower (the larger carve-out leaves no L1 for the windows)
    if (const char* e = getenv("POB_DEBUG_COLSMEM")) on = atoi(e) != 0 && pob_align_up(smem, 64) + cb <= budget_cta;
    if (on) { smem = pob_align_up(smem, 64); P.col_off = (int)smem; smem += cb; }
  }
  if (smem > 200 * 1024) return POB_EUNSUPPORTED;
  void (*kern)(BeamParams);
  const bool ctc = model == POB_MODEL_CTC;
  constexpr int M0 = POB_MODEL_CTC, M1 = POB_MODEL_CTC_MERGE_REPEATS;
  if (threads <= 64) {
    threads = 64;
    kern = ctc ? beam_kernel<M0, 64, 12> : beam_kernel<M1, 64, 12>;
    if (W == 5 && P.EMAX == 32 && r1.n_states == 5 && !getenv("POB_DEBUG_NO_CT")) {
      kern = ctc ? beam_kernel<M0, 64, 12, 32, 5> : beam_kernel<M1, 64, 12, 32, 5>;
      if (mode == MODE_ROWCOL && !P.dbg_noreuse)
        kern = ctc ? beam_kernel<M0, 64, 12, 32, 5, MODE_ROWCOL> : beam_kernel<M1, 64, 12, 32, 5, MODE_ROWCOL>;
    }
  }
  else if (threads <= 128) { kern = ctc ? beam_kernel<M0, 128, 6> : beam_kernel<M1, 128, 6>; }
  else if (threads <= 256) {
    kern = ctc ? beam_kernel<M0, 256, 3> : beam_kernel<M1, 256, 3>;
    if (W == 25 && P.EMAX == 128 && r1.n_states == 5 && !getenv("POB_DEBUG_NO_CT")) {
      kern = ctc ? beam_kernel<M0, 256, 3, 128, 25> : beam_kernel<M1, 256, 3, 128, 25>;
      if (mode == MODE_ROWCOL && !P.dbg_noreuse && !getenv("POB_DEBUG_NO_CTMODE"))
        kern = ctc ? beam_kernel<M0, 256, 3, 128, 25, MODE_ROWCOL> : beam_kernel<M1, 256, 3, 128, 25, MODE_ROWCOL>;
    }
    if (max_cta_sm == 2) kern = ctc ? beam_kernel<M0, 256, 2> : beam_kernel<M1, 256, 2>;
    if (max_cta_sm == 4) kern = ctc ? beam_kernel<M0, 256, 4> : beam_kernel<M1, 256, 4>;
  }
  else if (threads <= 288) {
    kern = ctc ? beam_kernel<M0, 288, 3> : beam_kernel<M1, 288, 3>;
    if (max_cta_sm == 2) kern = ctc ? beam_kernel<M0, 288, 2> : beam_kernel<M1, 288, 2>;
    if (max_cta_sm == 1) kern = ctc ? beam_kernel<M0, 288, 1> : beam_kernel<M1, 288, 1>;
  }
  else if (threads <= 512) { kern = ctc ? beam_kernel<M0, 512, 1> : beam_kernel<M1, 512, 1>; }
  else { kern = ctc ? beam_kernel<M0, 1024, 1> : beam_kernel<M1, 1024, 1>; }
  // The dynamic-shared-memory size and carve-out are attributes of the kernel FUNCTION, shared by every host thread of
  // the process (the command line keeps two GPU calls in flight): set them and launch under one lock, so that another
  // thread's smaller setting cannot land between this call's cudaFuncSetAttribute and its launch.
  static std::mutex launch_mu;
  std::unique_lock<std::mutex> launch_lock(launch_mu);
  POB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // leave most of the unified L1/shared array to L1 (window entries are served from it) but make sure the
  // shared-memory carve-out does not cap residency
  {
    int want_blocks = 2048 / threads;
    if (want_blocks > 12) want_blocks = 12;
    if (max_cta_sm > 0 && want_blocks > max_cta_sm) want_blocks = max_cta_sm;
    // the carve-out comes in steps (0, 8, 16, 32, 64, 100, 132, 164, 196, 228 KB): ask for the smallest one that
    // holds the wanted blocks (1 KB per block is reserved by the system), as a percentage that maps back onto it
    static const int steps_kb[] = {0, 8, 16, 32, 64, 100, 132, 164, 196, 228};
    const size_t need = (size_t)want_blocks * (smem + 1024);
    int kb = 228;
    for (int s : steps_kb) if ((size_t)s * 1024 >= need) { kb = s; break; }
    int pct = kb * 100 / 228;
    if (pct > 100) pct = 100;
    POB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
  }
  int per_sm = 0;
  POB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
  if (per_sm < 1) return POB_EUNSUPPORTED;
  if (max_cta_sm > 0 && per_sm > max_cta_sm) per_sm = max_cta_sm;
  // Workspace: one stride per resident CTA.  Prefer a full wave of CTAs; when the pool of a wide-band batch
  // makes that too large for HBM, first run fewer CTAs, then (only if a single CTA still does not fit) shrink
  // the pool and rely on the overflow flag.
  int grid = per_sm * ctx->sm_count;
  if (grid > n_items) grid = n_items;
  if (grid < 1) grid = 1;
  // Budget: what is free on the device right now (a second call in flight on another context sees what this one
  // left), at most 120 GB.  When a batch of wide-band items does not fit at full residency, the number of CTAs gives
  // way first (down to one per two SMs): a smaller pool would recycle nodes that are still readable
  // (POB_ST_POOL_OVERFLOW on every item of a config-4 style batch); only below that the pool is halved, and the
  // overflow flag says so per item.
  size_t free_b = 0, total_b = 0;
  POB_CUDA(cudaMemGetInfo(&free_b, &total_b));
  for (auto& b : ctx->blocks) free_b += b.size;  // this context's own scratch arena is reused by this call
  size_t budget = free_b / 10 * 6;
  if (budget > ((size_t)120 << 30)) budget = (size_t)120 << 30;
  if (budget < ((size_t)1 << 30)) budget = (size_t)1 << 30;
  size_t stride = ws_bytes(model, P.NP, P.CAP0, P.CAP1, P.RQ, P.CAPC0, P.CAPC1, Umax);
  const int floor_grid = grid < ctx->sm_count / 2 ? grid : ctx->sm_count / 2;
  while (grid > floor_grid && (size_t)grid * stride > budget) { grid = grid * 3 / 4; if (grid < floor_grid) grid = floor_grid; }
  while ((size_t)grid * stride > budget && P.NP > 8192) {
    P.NP = (P.NP / 2 + 255) / 256 * 256; P.RQ = P.NP * 2;
    stride = ws_bytes(model, P.NP, P.CAP0, P.CAP1, P.RQ, P.CAPC0, P.CAPC1, Umax);
  }
  while (grid > 1 && (size_t)grid * stride > budget) grid = grid * 3 / 4;
  while (P.NP > 1024 && stride > budget) {
    P.NP = (P.NP / 2 + 255) / 256 * 256; P.RQ = P.NP * 2;
    stride = ws_bytes(model, P.NP, P.CAP0, P.CAP1, P.RQ, P.CAPC0, P.CAPC1, Umax);
  }
  P.ws_stride = stride;
  P.ws = (char*)pob_arena_take(ctx, (size_t)grid * stride);
  if (!P.ws) return POB_ENOMEM;
  int* counter;
  POB_TRY(pob_take(ctx, 64, &counter));
  POB_CUDA(cudaMemsetAsync(counter, 0, 4, ctx->stream));
  POB_CUDA(cudaMemsetAsync(ctx->d_counters, 0, 2 * sizeof(unsigned long long), ctx->stream));
  P.work_counter = counter;
  P.counters = ctx->d_counters;
  P.trace = trace; P.trace_off = trace_off; P.out_top = top; P.out_score = out_score;
  if (getenv("POB_DEBUG_VERBOSE"))
    fprintf(stderr, "[pob] beam launch: items %d W %d mode %d NP %d CAP %d/%d span %d/%d threads %d grid %d (%d/SM) smem %zu ws/CTA %.1f MB\n",
            n_items, W, mode, P.NP, P.CAP0, P.CAP1, max_span0, max_span1, threads, grid, per_sm, smem,
            stride / 1048576.0);
  if (n_items > 0) {
    pob_prof_scope ps(ctx, mode == MODE_1D ? POB_K_BEAM_1D : POB_K_BEAM_2D);
    kern<<<grid, threads, smem, ctx->stream>>>(P);
  }
  POB_CUDA(cudaGetLastError());
  launch_lock.unlock();
  {
    pob_prof_scope ps(ctx, POB_K_BACKTRACE);
    backtrace_kernel<<<(n_total + 127) / 128, 128, 0, ctx->stream>>>(trace, trace_off, top, out_off, n_total, out_seq,
                                                                    out_len, out_status);
  }
  POB_CUDA(cudaGetLastError());
  return POB_OK;
}
