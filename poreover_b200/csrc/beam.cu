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
// (read or pair).  Nodes live in a pool in global memory (L2 resident); each (node, read) has a
// time-indexed ring window [lo, hi) of FP64 values -- exactly the entries the reference could still
// read, because every read is at t-1 >= (current u)-1.  Nodes that leave the expanded beam are retired,
// not dropped: a later revival (parent re-enters the beam) or a child reading its frozen parent sees the
// same stale values the reference's hash maps would return.  A retired node is recycled only once its
// windows are dead (or the pool overflows, which is flagged per item).
//
// Arithmetic: forward values are accumulated in FP64 (the reference is all double and scores reach
// -2.5e3..-5e4); only the bounded log1p(exp(d)) term, d <= 0, is evaluated in FP32.  No tensor cores:
// nothing here is a contraction.  Per step the dependent chain is the time-major band sweep: threads
// own (node, read) items, carry their own t-1 values in registers and exchange parent values through
// double-buffered shared memory, one barrier per time sub-step.
#include <stdlib.h>

#include "common.cuh"
#include "launch.cuh"

namespace {

enum { MODE_1D = 0, MODE_ROW = 1, MODE_ROWCOL = 2 };

struct __align__(16) NodeHdr {  // 128 bytes
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
  double maxp[2];               // max_prob[] of the reference nodes
  int32_t pad[2];
};
static_assert(sizeof(NodeHdr) == 112 || sizeof(NodeHdr) == 128, "hdr size");

template <int MODEL>
struct Entry;
template <>
struct __align__(32) Entry<POB_MODEL_CTC_MERGE_REPEATS> {
  double prob, gap, nogap, pad;
};
template <>
struct __align__(8) Entry<POB_MODEL_CTC> {
  double prob;
};

struct BeamParams {
  pob_reads r[2];
  const int32_t* env;       // rows x 2 (may be NULL: ROW without envelope, or 1D)
  const int64_t* env_off;
  const int32_t* envt;      // cols x 2 (ROWCOL only)
  const int64_t* envt_off;
  const int32_t* order;     // item processing order or NULL
  const int32_t* skip;      // per item != 0 -> not searched
  int n_items, W, mode, NP, CAP0, CAP1, RQ, EMAX;
  int dbg_noreclaim, dbg_step;
  double* dbg_trace;  // optional: [step][2] = (top score, sum of beam scores) after each prune of item 0
  char* ws;                 // workspace, one stride per resident CTA
  size_t ws_stride;
  uint32_t* trace;
  const int64_t* trace_off;
  int32_t* out_top;         // per item: start tid, extra last (-1), depth, status
  double* out_score;
  int* work_counter;
  unsigned long long* counters;
};

__device__ __forceinline__ double ninf() { return __longlong_as_double(0xfff0000000000000LL); }

// Log.h:27-33 with the bounded term in FP32: max + log1p(exp(min - max))
__device__ __forceinline__ double lae(double a, double b) {
  const double m = fmax(a, b);
  if (m == ninf()) return m;
  const float d = (float)(fmin(a, b) - m);
  return m + (double)log1pf(expf(d));
}

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

template <int MODEL>
struct Engine {
  typedef Entry<MODEL> Ent;
  // copied launch parameters (the engine object lives in shared memory: one copy per CTA)
  struct { int W, NP, RQ, EMAX, mode, noreclaim; } P;
  // workspace views
  NodeHdr* hdr;
  Ent* win[2];
  int32_t* freelist;
  int2* retq;
  double* cum[2];  // ctc: blank prefix sums of each read (root node values, PrefixTree.h:508-514)
  int32_t* sufmin; // ROW: min over rows >= u of the envelope's band start (band starts are NOT monotone:
                   // build_envelope's repair pass clamps some rows far back, envelope.py:82-85)
  int cap[2], mask[2];
  ReadView rv[2];
  uint32_t* trace;
  // shared state
  int16_t* slot2e;   // [NP]
  int32_t* E;        // [EMAX] pool slots of the expanded beam
  int32_t* Eold;     // [EMAX]
  uint8_t* act;      // [EMAX] active (not a duplicate)
  uint8_t* actold;
  int32_t* beam;     // [W]
  double* score;     // [EMAX]
  uint32_t* eorder;  // [EMAX]
  double2* pub;      // [2][EMAX][2]
  double* smax;      // [EMAX][2]
  int32_t* tmpa;     // [EMAX] scratch ints
  int32_t* tmpb;
  int32_t* sh;       // scalars: see SH_*
  enum { SH_NB = 0, SH_NE, SH_NFREE, SH_ORDER, SH_TID, SH_STAMP, SH_RQH, SH_RQT, SH_STATUS, SH_FIRSTALIVE,
         SH_FREED, SH_NEOLD, SH_TOTALLOC, SH_TOTFIRST, SH_DMIN, SH_REPUSH, SH_COUNT };

  __device__ __forceinline__ Ent* wptr(int slot, int r, int t) const {
    return win[r] + (size_t)slot * cap[r] + ((t + 1) & mask[r]);
  }

  // value of the root at time t (parent of depth-1 nodes)
  __device__ __forceinline__ double root_prob(int r, int t) const {
    if (t == -1) return 0.0;
    if (MODEL == POB_MODEL_CTC) return (t >= 0 && t < rv[r].T) ? cum[r][t] : ninf();
    return ninf();
  }

  // ---- one update_prob(n, r, t) with every input read from the stored windows -------------------
  // Split in two halves with a block barrier in between (update_all): when a node and its parent are
  // updated at the same t by different threads, the child must see the parent's window bounds and t-1
  // entry as they were BEFORE this phase (the reference reads t-1, writes t).  Reading lo/hi while the
  // parent's thread rewrites them can otherwise pair a new hi with an old lo and admit a stale entry.
  struct UpdIn {
    double p_prev, ng_prev, pv, ylast, yblank;
    int lo, hi;
  };

  __device__ __forceinline__ void update_gather(int slot, int r, int t, UpdIn& in) const {
    const NodeHdr& h = hdr[slot];
    const int last = h.last;
    in.lo = h.lo[r]; in.hi = h.hi[r];
    const bool self_ok = (t - 1 >= in.lo && t - 1 < in.hi);
    const Ent* se = wptr(slot, r, t - 1);
    in.p_prev = self_ok ? se->prob : ninf();
    in.ng_prev = ninf();
    if constexpr (MODEL == POB_MODEL_CTC_MERGE_REPEATS) in.ng_prev = self_ok ? se->nogap : ninf();
    in.ylast = rv[r].at(t, rv[r].pcol(last));
    in.yblank = rv[r].at(t, rv[r].cblank);
    const int ps = h.parent_slot;
    if (ps < 0) {
      in.pv = root_prob(r, t - 1);
    } else {
      const NodeHdr& ph = hdr[ps];
      const bool same = (ph.last == last);
      const int plo = ph.lo[r], phi = ph.hi[r];
      if (ph.order == h.parent_order && t - 1 >= plo && t - 1 < phi) {
        const Ent* pe = wptr(ps, r, t - 1);
        if constexpr (MODEL == POB_MODEL_CTC_MERGE_REPEATS) in.pv = same ? pe->gap : pe->prob;
        else in.pv = pe->prob;
      } else {
        in.pv = ninf();
      }
    }
  }

  __device__ __forceinline__ void update_commit(int slot, int r, int t, const UpdIn& in) {
    NodeHdr& h = hdr[slot];
    Ent out;
    double prob;
    if constexpr (MODEL == POB_MODEL_CTC_MERGE_REPEATS) {
      const double gp = in.p_prev + in.yblank;
      const double ng = lae(in.pv + in.ylast, in.ng_prev + in.ylast);
      prob = lae(gp, ng);
      out.prob = prob; out.gap = gp; out.nogap = ng; out.pad = 0;
    } else {
      prob = lae(in.pv + in.ylast, in.p_prev + in.yblank);
      out.prob = prob;
    }
    *wptr(slot, r, t) = out;
    int lo = in.lo, hi = in.hi;
    if (t >= hi) { if (t > hi) lo = t; hi = t + 1; }
    else if (t < lo) { lo = t; hi = t + 1; }
    if (hi - lo > cap[r]) lo = hi - cap[r];
    h.lo[r] = lo; h.hi[r] = hi;
    if (prob > h.maxp[r]) h.maxp[r] = prob;
  }

  // all threads call this; `mine` selects the threads that own a node (slot) this phase
  __device__ void update_all(bool mine, int slot, int r, int t) {
    UpdIn in;
    if (mine) update_gather(slot, r, t, in);
    __syncthreads();
    if (mine) update_commit(slot, r, t, in);
    __syncthreads();
  }

  // ---- time-major band sweep over the expanded beam (BeamSearch.h:361-375, :146-156) ----------------
  // reads_mask bit r: read r swept over [t0[r], t1[r]).  Resets max_prob of the swept reads first.
  __device__ void sweep(int nE, int reads_mask, const int* t0, const int* t1, bool reset_other,
                        unsigned long long& n_updates) {
    const int tid = threadIdx.x;
    const int EMAX = P.EMAX;
    const int nItems = nE * 2;
    const int len0 = (reads_mask & 1) ? t1[0] - t0[0] : 0, len1 = (reads_mask & 2) ? t1[1] - t0[1] : 0;
    const int maxlen = max(len0, len1);
    // one item per thread (launcher guarantees blockDim >= 2*EMAX)
    const int item = tid;
    const bool have = item < nItems;
    const int e = item >> 1, r = item & 1;
    const bool live = have && act[e];
    const bool on = live && ((reads_mask >> r) & 1);
    int slot = 0, pe = -1, lo = 0, hi = 0, ps = -1, plo = 0, phi = 0;
    bool same = false, proot = false, pfrozen = false;
    double p_prev = ninf(), ng_prev = ninf(), maxv = ninf();
    int ts = 0, te = 0;
    // hoisted addressing: own window, parent window, the two probability columns of this read
    Ent* wbase = nullptr;
    const Ent* pwbase = nullptr;
    int wmask = 0;
    const char* ylast_p = nullptr;
    const char* yblank_p = nullptr;
    long ystep = 0;
    bool f64 = false;
    if (live) slot = E[e];
    if (on) {
      NodeHdr& h = hdr[slot];
      const int last = h.last;
      lo = h.lo[r]; hi = h.hi[r];
      ts = t0[r]; te = t1[r];
      wmask = mask[r];
      wbase = win[r] + (size_t)slot * cap[r];
      {
        const ReadView v = rv[r];
        f64 = v.f64;
        const long es = f64 ? 8 : 4;
        const long rowb = (long)v.S * es;
        const char* row0 = (const char*)v.base + (long)v.prow(ts) * rowb;
        ystep = v.rc ? -rowb : rowb;
        ylast_p = row0 + (long)v.pcol(last) * es;
        yblank_p = row0 + (long)v.cblank * es;
      }
      if (ts - 1 >= lo && ts - 1 < hi) {
        const Ent* se = wbase + (ts & wmask);
        p_prev = se->prob;
        if constexpr (MODEL == POB_MODEL_CTC_MERGE_REPEATS) ng_prev = se->nogap;
      }
      ps = h.parent_slot;
      if (ps < 0) proot = true;
      else {
        const NodeHdr& ph = hdr[ps];
        same = (ph.last == last);
        if (ph.order != h.parent_order) { ps = -2; }  // recycled parent: every read is -inf
        else {
          pe = slot2e[ps];
          if (pe < 0) { pfrozen = true; plo = ph.lo[r]; phi = ph.hi[r]; pwbase = win[r] + (size_t)ps * cap[r]; }
        }
      }
      // publish the stored values at ts-1 for children whose parent is swept too
      double2 pb; pb.x = p_prev; pb.y = ninf();
      if constexpr (MODEL == POB_MODEL_CTC_MERGE_REPEATS) {
        if (ts - 1 >= lo && ts - 1 < hi) pb.y = (wbase + (ts & wmask))->gap;
      }
      pub[(0 * EMAX + e) * 2 + r] = pb;
    }
    __syncthreads();
    const double2* pub_rd = pub + (size_t)(pe < 0 ? 0 : pe) * 2 + r;
    double2* pub_wr = pub + (size_t)e * 2 + r;
    const int pstride = EMAX * 2;
    for (int it = 0; it < maxlen; ++it) {
      const int t = ts + it;
      const bool go = on && t < te;
      if (go) {
        double pv;
        if (proot) pv = root_prob(r, t - 1);
        else if (ps == -2) pv = ninf();
        else if (!pfrozen) {
          const double2 pb = pub_rd[(it & 1) * pstride];
          pv = (MODEL == POB_MODEL_CTC_MERGE_REPEATS && same) ? pb.y : pb.x;
        } else if (t - 1 >= plo && t - 1 < phi) {
          const Ent* q = pwbase + (t & wmask);
          if constexpr (MODEL == POB_MODEL_CTC_MERGE_REPEATS) pv = same ? q->gap : q->prob;
          else pv = q->prob;
        } else pv = ninf();
        double ylast, yblank;
        if (f64) { ylast = __ldg((const double*)ylast_p); yblank = __ldg((const double*)yblank_p); }
        else { ylast = (double)__ldg((const float*)ylast_p); yblank = (double)__ldg((const float*)yblank_p); }
        ylast_p += ystep; yblank_p += ystep;
        double prob;
        double2 pb;
        Ent* o = wbase + ((t + 1) & wmask);
        if constexpr (MODEL == POB_MODEL_CTC_MERGE_REPEATS) {
          const double gp = p_prev + yblank;
          const double ng = lae(pv + ylast, ng_prev + ylast);
          prob = lae(gp, ng);
          double4 v4; v4.x = prob; v4.y = gp; v4.z = ng; v4.w = 0;
          *reinterpret_cast<double4*>(o) = v4;
          ng_prev = ng;
          pb.x = prob; pb.y = gp;
        } else {
          prob = lae(pv + ylast, p_prev + yblank);
          o->prob = prob;
          pb.x = prob; pb.y = ninf();
        }
        p_prev = prob;
        if (prob > maxv) maxv = prob;
        pub_wr[((it + 1) & 1) * pstride] = pb;
      }
      __syncthreads();
    }
    if (on) {
      NodeHdr& h = hdr[slot];
      if (te > ts) {
        // window bookkeeping for the contiguous write [ts, te)
        if (ts > hi || ts < lo) { lo = ts; hi = te; }
        else hi = max(hi, te);
        if (hi - lo > cap[r]) lo = hi - cap[r];
        h.lo[r] = lo; h.hi[r] = hi;
        h.maxp[r] = maxv;  // reset + max over the band
        smax[e * 2 + r] = maxv;
      } else {
        smax[e * 2 + r] = h.maxp[r];  // empty band: max_prob left stale (A.6b)
      }
      n_updates += (unsigned long long)(te - ts);
    } else if (live && reset_other) {
      smax[e * 2 + r] = hdr[slot].maxp[r];
    }
    __syncthreads();
  }

  // ---- expansion: E = beam + children(beam) (PrefixTree.h:439-446), children created / revived ----
  // bfs_first_row: ROW traversal while the beam is shorter than W
  __device__ void build_expanded(int item, bool bfs) {
    const int tid = threadIdx.x;
    const int nb = sh[SH_NB];
    const int W = P.W;
    int nexp = bfs ? W : nb;  // number of nodes expanded this step
    // In BFS mode element k >= nb is child (k-nb)%4 of element (k-nb)/4, all freshly created (first row).
    if (tid < nb) { E[tid] = beam[tid]; act[tid] = 1; }
    __syncthreads();
    if (!bfs) {
      // per beam node: which children are missing, is this its first expansion
      int miss = 0, first = 0;
      if (tid < nb) {
        NodeHdr& h = hdr[beam[tid]];
        first = h.tid < 0;
        for (int c = 0; c < 4; ++c) {
          const int ks = h.kid_slot[c];
          const bool ok = !first && ks >= 0 && hdr[ks].order == h.kid_order[c];
          if (!ok) miss |= 1 << c;
        }
        tmpa[tid] = __popc(miss);
        tmpb[tid] = first;
      }
      __syncthreads();
      if (tid < nb) {
        int abase = 0, fbase = 0;
        for (int b = 0; b < tid; ++b) { abase += tmpa[b]; fbase += tmpb[b]; }
        if (tid == nb - 1) { sh[SH_TOTALLOC] = abase + tmpa[tid]; sh[SH_TOTFIRST] = fbase + tmpb[tid]; }
        const int slot = beam[tid];
        NodeHdr& h = hdr[slot];
        if (first) {
          h.tid = sh[SH_TID] + fbase;
          trace[h.tid] = ((uint32_t)h.parent_tid << 2) | (uint32_t)h.last;
        }
        int k = 0;
        for (int c = 0; c < 4; ++c) {
          int ks;
          if (miss & (1 << c)) {
            const int fi = sh[SH_NFREE] - 1 - (abase + k);
            ks = (fi >= 0) ? freelist[fi] : -1;
            if (ks >= 0) {
              NodeHdr n;
              n.order = (uint32_t)(sh[SH_ORDER] + abase + k);
              n.state = 0; n.parent_slot = slot; n.parent_order = h.order; n.parent_tid = h.tid; n.tid = -1;
              n.depth = h.depth + 1; n.last = c;
              for (int q = 0; q < 4; ++q) { n.kid_slot[q] = -1; n.kid_order[q] = 0; }
              n.lo[0] = n.lo[1] = 0; n.hi[0] = n.hi[1] = 0;
              n.maxp[0] = n.maxp[1] = ninf(); n.pad[0] = n.pad[1] = 0;
              hdr[ks] = n;
              h.kid_slot[c] = ks; h.kid_order[c] = n.order;
            }
            ++k;
          } else {
            ks = h.kid_slot[c];
            if (hdr[ks].state != 0) hdr[ks].state = 0;  // revived with its retained windows
          }
          E[nb + 4 * tid + c] = ks;
          act[nb + 4 * tid + c] = ks >= 0;
        }
      }
      __syncthreads();
      if (tid == 0) {
        if (sh[SH_TOTALLOC] > sh[SH_NFREE]) sh[SH_STATUS] |= POB_ST_POOL_OVERFLOW;
        sh[SH_NFREE] = max(0, sh[SH_NFREE] - sh[SH_TOTALLOC]);
        sh[SH_ORDER] += sh[SH_TOTALLOC];
        sh[SH_TID] += sh[SH_TOTFIRST];
        sh[SH_NE] = 5 * nb;
      }
      __syncthreads();
      // children that are themselves beam members are duplicates (Beam.h:96-99)
      const int nE = 5 * nb;
      if (tid >= nb && tid < nE && act[tid]) {
        const int s = E[tid];
        for (int b = 0; b < nb; ++b) if (beam[b] == s) { act[tid] = 0; break; }
      }
    } else {
      // first-row BFS: sequential dependency only through indices, so do it level by level on thread 0..:
      // element k expanded for k < W; all expansions are first expansions with fresh children.
      if (tid == 0) {
        int ne = nb;
        for (int k = 0; k < nexp && k < ne; ++k) {
          const int slot = E[k];
          NodeHdr& h = hdr[slot];
          if (h.tid < 0) {
            h.tid = sh[SH_TID]++;
            trace[h.tid] = ((uint32_t)h.parent_tid << 2) | (uint32_t)h.last;
          }
          for (int c = 0; c < 4; ++c) {
            int ks = h.kid_slot[c];
            if (!(ks >= 0 && hdr[ks].order == h.kid_order[c])) {
              const int fi = --sh[SH_NFREE];
              ks = freelist[fi];
              NodeHdr n;
              n.order = (uint32_t)(sh[SH_ORDER]++);
              n.state = 0; n.parent_slot = slot; n.parent_order = h.order; n.parent_tid = h.tid; n.tid = -1;
              n.depth = h.depth + 1; n.last = c;
              for (int q = 0; q < 4; ++q) { n.kid_slot[q] = -1; n.kid_order[q] = 0; }
              n.lo[0] = n.lo[1] = 0; n.hi[0] = n.hi[1] = 0;
              n.maxp[0] = n.maxp[1] = ninf(); n.pad[0] = n.pad[1] = 0;
              hdr[ks] = n;
              h.kid_slot[c] = ks; h.kid_order[c] = n.order;
            } else if (hdr[ks].state != 0) hdr[ks].state = 0;
            bool dupl = false;
            for (int q = 0; q < ne; ++q) if (E[q] == ks) { dupl = true; break; }
            E[ne] = ks; act[ne] = !dupl; ++ne;
          }
        }
        sh[SH_NE] = ne;
      }
      __syncthreads();
    }
    __syncthreads();
    const int nE = sh[SH_NE];
    if (tid < nE && act[tid]) slot2e[E[tid]] = (int16_t)tid;
    __syncthreads();
  }

  // ---- Beam::prune (Beam.h:93-108): rank by score desc, exact ties by creation order ----------
  __device__ void prune(int nE) {
    const int tid = threadIdx.x;
    int rank = 0;
    bool a = tid < nE && act[tid];
    if (a && !(score[tid] == score[tid])) score[tid] = ninf();  // NaN cannot be ranked
    if (tid == 0) sh[SH_DMIN] = 0x7fffffff;
    __syncthreads();
    if (a) {
      const double s = score[tid];
      const uint32_t o = eorder[tid];
      for (int j = 0; j < nE; ++j) {
        if (!act[j]) continue;
        const double sj = score[j];
        rank += (sj > s) || (sj == s && eorder[j] < o);
      }
    }
    __syncthreads();
    if (a && rank < P.W) {
      beam[rank] = E[tid];
      atomicMin(&sh[SH_DMIN], hdr[E[tid]].depth);
    }
    if (tid == 0) {
      int n = 0;
      for (int j = 0; j < nE; ++j) n += act[j];
      sh[SH_NB] = min(n, P.W);
    }
    __syncthreads();
  }

  // ---- retire nodes that left the expanded beam; recycle retired nodes that can never matter again ----
  // A retired node is freed when both windows are dead (every later read is at an index >= dead_r, so
  // hi <= dead_r means all reads miss) AND it can no longer hand retained children to a revival: either
  // all its child links are already invalid, or it is unreachable.  The beam's minimum depth never
  // decreases (the next beam is drawn from beam + children), so a non-active node of depth <= that minimum
  // has no ancestor that can ever be expanded again: unreachable.
  __device__ void retire_and_reclaim(int nEold, int dead0, int dead1) {
    const int tid = threadIdx.x;
    int myslot = -1;
    if (tid < nEold && actold[tid]) {
      myslot = Eold[tid];
      if (slot2e[myslot] < 0) {
        const int pos = atomicAdd(&sh[SH_RQT], 1);
        const int stamp = atomicAdd(&sh[SH_STAMP], 1) + 1;
        hdr[myslot].state = stamp;
        retq[pos % P.RQ] = make_int2(myslot, stamp);
      }
    }
    if (tid == 0) { sh[SH_FIRSTALIVE] = 0x7fffffff; sh[SH_FREED] = 0; sh[SH_REPUSH] = 0; }
    __syncthreads();
    const int head = sh[SH_RQH], tail = sh[SH_RQT];
    const int navail = P.noreclaim ? 0 : min(tail - head, (int)blockDim.x);
    const int dmin = sh[SH_DMIN];
    int st = 0;  // 1 stale, 2 dead+freeable, 3 dead but must be kept, 4 alive
    int slot = -1, stamp = 0;
    if (tid < navail) {
      const int2 q = retq[(head + tid) % P.RQ];
      slot = q.x; stamp = q.y;
      const NodeHdr& h = hdr[slot];
      if (h.state != stamp) st = 1;
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
    __syncthreads();
    int fa = min(sh[SH_FIRSTALIVE], navail);  // entries before the first live one are consumed
    bool force = false;
    if (sh[SH_NFREE] + fa < 4 * P.W + 8 && navail > 0) {
      // pool pressure: recycle live retirees too (flagged; the reference never frees anything)
      fa = navail; force = true;
      if (tid == 0) sh[SH_STATUS] |= POB_ST_POOL_OVERFLOW;
    }
    if (tid < fa) {
      if (st == 2 || (force && st >= 2)) {
        NodeHdr& h = hdr[slot];
        h.order = 0; h.state = -1;
        const int pos = atomicAdd(&sh[SH_FREED], 1);
        freelist[sh[SH_NFREE] + pos] = slot;
      } else if (st == 3 || st == 4) {
        const int pos = atomicAdd(&sh[SH_REPUSH], 1);
        retq[(tail + pos) % P.RQ] = make_int2(slot, stamp);  // look again one queue cycle later
      }
    }
    __syncthreads();
    if (tid == 0) { sh[SH_NFREE] += sh[SH_FREED]; sh[SH_RQH] = head + fa; sh[SH_RQT] = tail + sh[SH_REPUSH]; }
    __syncthreads();
  }

  __device__ void dbg_record(const BeamParams& G, long step) {
    if (!G.dbg_trace) return;
    __syncthreads();
    if (threadIdx.x == 0 && step < 100000) {
      double sum = 0;
      for (int b = 0; b < sh[SH_NB]; ++b) {
        const NodeHdr& h = hdr[beam[b]];
        double sc = (P.mode == MODE_1D) ? wptr(beam[b], 0, h.hi[0] - 1)->prob
                  : (P.mode == MODE_ROW) ? wptr(beam[b], 0, h.hi[0] - 1)->prob + h.maxp[1] : h.maxp[0] + h.maxp[1];
        if (b == 0) G.dbg_trace[2 * step] = sc;
        if (sc > -1e300) sum += sc;
      }
      G.dbg_trace[2 * step + 1] = sum;
    }
    __syncthreads();
  }

  __device__ void save_old(int nE) {
    const int tid = threadIdx.x;
    if (tid < nE) {
      Eold[tid] = E[tid]; actold[tid] = act[tid];
      if (act[tid]) slot2e[E[tid]] = -1;
    }
    if (tid == 0) sh[SH_NEOLD] = nE;
    __syncthreads();
  }

  __device__ void run_item(const BeamParams& G, int item, char* ws, char* smem);
};

template <int MODEL>
__device__ void Engine<MODEL>::run_item(const BeamParams& G, int item, char* ws, char* smem) {
  const int tid = threadIdx.x, NT = blockDim.x;
  const int W = G.W, NP = G.NP, EMAX = G.EMAX;
  const int mode = G.mode;
  unsigned long long n_updates = 0;
  // ---- carve shared memory (the engine object itself sits at the front)
  if (tid == 0) {
    P.W = W; P.NP = NP; P.RQ = G.RQ; P.EMAX = EMAX; P.mode = mode; P.noreclaim = G.dbg_noreclaim;
    char* p = smem + ((sizeof(Engine<MODEL>) + 15) & ~(size_t)15);
    pub = (double2*)p; p += sizeof(double2) * 2 * EMAX * 2;
    score = (double*)p; p += sizeof(double) * EMAX;
    smax = (double*)p; p += sizeof(double) * EMAX * 2;
    E = (int32_t*)p; p += 4 * EMAX;
    Eold = (int32_t*)p; p += 4 * EMAX;
    eorder = (uint32_t*)p; p += 4 * EMAX;
    tmpa = (int32_t*)p; p += 4 * EMAX;
    tmpb = (int32_t*)p; p += 4 * EMAX;
    beam = (int32_t*)p; p += 4 * ((W + 3) & ~3);
    sh = (int32_t*)p; p += 4 * 32;
    slot2e = (int16_t*)p; p += 2 * NP;
    act = (uint8_t*)p; p += (EMAX + 15) & ~15;
    actold = (uint8_t*)p; p += (EMAX + 15) & ~15;
  }
  // ---- carve the global workspace
  if (tid == 0) {
  cap[0] = G.CAP0; cap[1] = G.CAP1; mask[0] = G.CAP0 - 1; mask[1] = G.CAP1 - 1;
  rv[0] = make_view(G.r[0], item);
  if (mode != MODE_1D) rv[1] = make_view(G.r[1], item); else { rv[1] = rv[0]; rv[1].T = 0; }
  {
    char* p = ws;
    hdr = (NodeHdr*)p; p += sizeof(NodeHdr) * (size_t)NP;
    win[0] = (Ent*)p; p += sizeof(Ent) * (size_t)NP * cap[0];
    win[1] = (Ent*)p; p += sizeof(Ent) * (size_t)NP * cap[1];
    freelist = (int32_t*)p; p += 4 * (size_t)NP;
    retq = (int2*)p; p += 8 * (size_t)G.RQ;
    cum[0] = (double*)p; p += 8 * (size_t)(MODEL == POB_MODEL_CTC ? rv[0].T : 0);
    cum[1] = (double*)p; p += 8 * (size_t)(MODEL == POB_MODEL_CTC ? rv[1].T : 0);
    sufmin = (int32_t*)p;
  }
  trace = G.trace + G.trace_off[item];
  }
  __syncthreads();
  const int U = rv[0].T, V = rv[1].T;
  int32_t* otop = G.out_top + 4 * (size_t)item;

  // ---- init pool
  for (int s = tid; s < NP; s += NT) {
    hdr[s].order = 0; hdr[s].state = -1;
    freelist[s] = NP - 1 - s;  // pops come from the end: slot 0 first
    slot2e[s] = -1;
  }
  if (tid == 0) {
    for (int k = 0; k < SH_COUNT; ++k) sh[k] = 0;
    sh[SH_NFREE] = NP; sh[SH_ORDER] = 1; sh[SH_TID] = 1;
  }
  if (MODEL == POB_MODEL_CTC && tid < 2 && (tid == 0 || mode != MODE_1D)) {
    // PrefixTree.h:508-514: sequential running sum of the blank column
    const ReadView& v = rv[tid];
    double s = 0;
    for (int t = 0; t < v.T; ++t) { s += v.at(t, v.cblank); cum[tid][t] = s; }
  }
  __syncthreads();
  if (U <= 0 || (mode != MODE_1D && V <= 0)) {
    if (tid == 0) { otop[0] = 0; otop[1] = -1; otop[2] = 0; otop[3] = POB_ST_EMPTY; G.out_score[item] = 0; }
    return;
  }
  // ---- seed: the 4 children of the root, updated at t = 0 (BeamSearch.h:24-30, :287-293)
  const int nbase = rv[0].S - 1;
  if (tid < nbase) {
    const int slot = freelist[NP - 1 - tid];
    NodeHdr n;
    n.order = 1 + tid; n.state = 0; n.parent_slot = -1; n.parent_order = 0; n.parent_tid = 0; n.tid = -1;
    n.depth = 1; n.last = tid;
    for (int q = 0; q < 4; ++q) { n.kid_slot[q] = -1; n.kid_order[q] = 0; }
    n.lo[0] = n.lo[1] = 0; n.hi[0] = n.hi[1] = 0; n.maxp[0] = n.maxp[1] = ninf(); n.pad[0] = n.pad[1] = 0;
    hdr[slot] = n;
    beam[tid] = slot;
    n_updates += (mode != MODE_1D) ? 2 : 1;
  }
  __syncthreads();
  update_all(tid < nbase, tid < nbase ? beam[tid] : 0, 0, 0);
  if (mode != MODE_1D) update_all(tid < nbase, tid < nbase ? beam[tid] : 0, 1, 0);
  if (tid == 0) { sh[SH_NFREE] = NP - nbase; sh[SH_ORDER] = 1 + nbase; sh[SH_NB] = nbase; }
  __syncthreads();

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
      if (i < U) sufmin[i] = v;
      carry = __shfl_sync(0xffffffffu, v, 0);
    }
  }
  __syncthreads();

  if (mode == MODE_1D) {
    // BeamSearch.h:33-53
    build_expanded(item, false);
    for (int t = 1; t < U; ++t) {
      const int nE = sh[SH_NE];
      const bool mine = tid < nE && act[tid];
      update_all(mine, mine ? E[tid] : 0, 0, t);
      if (mine) {
        n_updates++;
        score[tid] = wptr(E[tid], 0, t)->prob;  // last_probability(): value at the last written t
        eorder[tid] = hdr[E[tid]].order;
      }
      __syncthreads();
      save_old(nE);
      prune(nE);
      build_expanded(item, false);
      retire_and_reclaim(nE, t, 0x7fffffff);  // the next step reads index t
      ++nsteps;
    }
  } else if (mode == MODE_ROW) {
    // BeamSearch.h:127-167 (envelope) / :196-247 (no envelope: rows start at 1, band = [0, V))
    build_expanded(item, sh[SH_NB] < W);
    for (int u = env ? 0 : 1; u < U; ++u) {
      int rs = env ? env[2 * u] : 0, re = env ? env[2 * u + 1] : V;
      rs = max(rs, 0); re = min(re, V);
      const int nE = sh[SH_NE];
      {
        const bool mine = tid < nE && act[tid];
        update_all(mine, mine ? E[tid] : 0, 0, u);
        if (mine) n_updates++;
      }
      int t0[2] = {u, rs}, t1[2] = {u + 1, re};
      sweep(nE, 2, t0, t1, false, n_updates);
      if (tid < nE && act[tid]) {
        score[tid] = wptr(E[tid], 0, u)->prob + smax[tid * 2 + 1];  // max_probability() (PrefixTree.h:107, :393)
        eorder[tid] = hdr[E[tid]].order;
      }
      __syncthreads();
      if (G.dbg_trace && nsteps >= G.dbg_step - 2 && nsteps <= G.dbg_step + 1 && tid < nE) {
        // rows of 10 doubles at offset 4000 + ((step - (dbg_step-2)) * 160 + tid) * 10
        double* o = G.dbg_trace + 4000 + ((nsteps - (G.dbg_step - 2)) * 160 + tid) * 10;
        const NodeHdr& h = hdr[E[tid]];
        o[0] = act[tid] ? (double)h.order : -1.0; o[1] = h.depth; o[2] = h.last; o[3] = h.parent_order;
        o[4] = act[tid] ? score[tid] : 0; o[5] = wptr(E[tid], 0, u)->prob; o[6] = smax[tid * 2 + 1];
        o[7] = h.lo[1]; o[8] = h.hi[1]; o[9] = E[tid];
        if (h.parent_slot >= 0) { const NodeHdr& ph = hdr[h.parent_slot]; o[2] = ph.order; o[6] = ph.hi[0]; o[7] = ph.state; o[8] = slot2e[h.parent_slot]; o[1] = h.parent_slot; }
      }
      save_old(nE);
      prune(nE);
      dbg_record(G, nsteps);
      build_expanded(item, sh[SH_NB] < W);
      // read 0 is next read at index u; read 1 at >= (smallest band start of any later row) - 1
      const int nrs = (u + 1 < U) ? (env ? sufmin[u + 1] : 0) : 0x7ffffffe;
      retire_and_reclaim(nE, u, nrs - 1);
      ++nsteps;
    }
  } else {
    // BeamSearch.h:295-392
    int u = 0, v = 0;
    bool have_E = false;
    while (u <= U - 1 && v <= V - 1) {
      const int ers = env[2 * u], ere = env[2 * u + 1], ecs = envt[2 * v], ece = envt[2 * v + 1];
      int row_start = v, row_end = v, col_start = u, col_end = u;
      bool rset = false, cset = false;
      if (v >= ers && v < ere) { row_end = ere; rset = true; }
      else if (v < ers) {
        const int nb = sh[SH_NB];
        if (nb < W && tid == 0) sh[SH_STATUS] |= POB_ST_SHORT_BEAM_SKIP;
        update_all(tid < nb, tid < nb ? beam[tid] : 0, 1, v);
        if (tid < nb) n_updates++;
        ++v; ++nsteps;
        continue;
      }
      if (u >= ecs && u < ece) { col_end = ece; cset = true; }
      else if (u < ecs) {
        const int nb = sh[SH_NB];
        if (nb < W && tid == 0) sh[SH_STATUS] |= POB_ST_SHORT_BEAM_SKIP;
        update_all(tid < nb, tid < nb ? beam[tid] : 0, 0, u);
        if (tid < nb) n_updates++;
        ++u; ++nsteps;
        continue;
      }
      if ((!rset || !cset) && tid == 0) sh[SH_STATUS] |= POB_ST_UNSET_BAND;
      row_end = min(row_end, V); col_end = min(col_end, U);
      if (!have_E) { build_expanded(item, false); have_E = true; }
      const int nE = sh[SH_NE];
      int t0[2] = {col_start, row_start}, t1[2] = {col_end, row_end};
      sweep(nE, 3, t0, t1, false, n_updates);
      if (tid < nE && act[tid]) {
        score[tid] = smax[tid * 2] + smax[tid * 2 + 1];  // max_probability_sym() (PrefixTree.h:111, :397)
        eorder[tid] = hdr[E[tid]].order;
      }
      __syncthreads();
      save_old(nE);
      prune(nE);
      dbg_record(G, nsteps);
      build_expanded(item, false);  // next step's expansion, done eagerly so retirement knows who stays
      retire_and_reclaim(nE, u, v);  // later reads are at t-1 >= u (read 0) and >= v (read 1)
      ++u; ++v; ++nsteps;
    }
  }
  __syncthreads();
  if (tid == 0) {
    const NodeHdr& top = hdr[beam[0]];
    double sc;
    if (mode == MODE_1D) sc = wptr(beam[0], 0, top.hi[0] - 1)->prob;
    else if (mode == MODE_ROW) sc = wptr(beam[0], 0, top.hi[0] - 1)->prob + top.maxp[1];
    else sc = top.maxp[0] + top.maxp[1];
    G.out_score[item] = sc;
    if (top.tid >= 0) { otop[0] = top.tid; otop[1] = -1; }
    else { otop[0] = top.parent_tid; otop[1] = top.last; }
    otop[2] = top.depth;
    otop[3] = sh[SH_STATUS];
  }
  // per-item counters
  for (int o = 16; o > 0; o >>= 1) n_updates += __shfl_down_sync(0xffffffffu, n_updates, o);
  if ((tid & 31) == 0 && n_updates) atomicAdd(&G.counters[0], n_updates);
  if (tid == 0) atomicAdd(&G.counters[1], (unsigned long long)nsteps);
  __syncthreads();
}

template <int MODEL, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) beam_kernel(BeamParams P) {
  extern __shared__ __align__(16) char smem[];
  __shared__ int s_item;
  char* ws = P.ws + (size_t)blockIdx.x * P.ws_stride;
  Engine<MODEL>& eng = *reinterpret_cast<Engine<MODEL>*>(smem);
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
    eng.run_item(P, item, ws, smem);
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
double* g_pob_dbg_trace = nullptr;
extern "C" int pob_debug_trace(pob_ctx* ctx, double* out, int n) {
  if (!g_pob_dbg_trace) return POB_EINVAL;
  POB_CUDA(cudaMemcpy(out, g_pob_dbg_trace, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
  return POB_OK;
}
namespace {

size_t ws_bytes(int model, int NP, int CAP0, int CAP1, int RQ, int Umax, int Vmax) {
  const size_t es = model == POB_MODEL_CTC ? 8 : 32;
  size_t b = sizeof(NodeHdr) * (size_t)NP + es * (size_t)NP * ((size_t)CAP0 + CAP1) + 4 * (size_t)NP + 8 * (size_t)RQ;
  if (model == POB_MODEL_CTC) b += 8 * ((size_t)Umax + Vmax + 2);
  b += 4 * ((size_t)Umax + 2);
  return pob_align_up(b, 256);
}

size_t smem_bytes(int W, int NP, int EMAX) {
  size_t b = 512 + sizeof(double2) * 2 * EMAX * 2 + 8 * EMAX + 16 * EMAX + 4 * EMAX * 5 + 4 * ((W + 3) & ~3) + 4 * 32 +
             2 * (size_t)NP + 2 * ((EMAX + 15) & ~15);
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
  if (n_items <= 0) return POB_OK;
  if (W < 4 || W > 100) return POB_EUNSUPPORTED;
  if (r1.n_states != 5 || (r2 && r2->n_states != 5)) return POB_EUNSUPPORTED;
  BeamParams P;
  memset(&P, 0, sizeof(P));
  P.r[0] = r1;
  if (r2) P.r[1] = *r2;
  P.env = env; P.env_off = env_off; P.envt = envt; P.envt_off = envt_off; P.order = order; P.skip = skip;
  P.n_items = n_items; P.W = W; P.mode = mode;
  P.EMAX = 5 * W + 4;
  // Node pool: the expanded beam (5W) plus retired nodes whose windows can still be read.  About W nodes
  // retire per step and stay readable for one band width, so the pool scales with W x widest band; an
  // overflow is flagged per item (POB_ST_POOL_OVERFLOW), never silent.
  {
    const int span = (mode == MODE_1D) ? 1 : (max_span0 > max_span1 ? max_span0 : max_span1);
    long want = 64L * W;
    const long by_span = 3L * W * (span + 8) + 16L * W;
    if (by_span > want) want = by_span;
    if (want > 65536) want = 65536;
    P.NP = pow2_at_least((int)want);
    if (P.NP < 1024) P.NP = 1024;
  }
  if (const char* e = getenv("POB_DEBUG_NP")) P.NP = atoi(e);
  if (const char* e = getenv("POB_DEBUG_NORECLAIM")) P.dbg_noreclaim = atoi(e);
  P.dbg_step = -1000;
  if (const char* e = getenv("POB_DEBUG_STEP")) P.dbg_step = atoi(e);
  if (getenv("POB_DEBUG_TRACE")) {
    static double* dbg = nullptr;
    if (!dbg) cudaMalloc(&dbg, 200000 * sizeof(double));
    cudaMemsetAsync(dbg, 0, 200000 * sizeof(double), ctx->stream);
    P.dbg_trace = dbg;
    g_pob_dbg_trace = dbg;
  }
  P.CAP0 = pow2_at_least(max_span0 + 3);
  P.CAP1 = pow2_at_least(max_span1 + 3);
  // keep one CTA's workspace under ~160 MB so that a full wave of CTAs fits in HBM (wide-band batches then
  // run with a smaller pool and rely on the overflow flag)
  while (P.NP > 1024 && ws_bytes(model, P.NP, P.CAP0, P.CAP1, 2 * P.NP, Umax, Vmax) > ((size_t)160 << 20)) P.NP >>= 1;
  P.RQ = P.NP * 2;
  // the band sweep wants one thread per (node, read); the single-read search one per node
  int threads = (((mode == MODE_1D ? 1 : 2) * P.EMAX + 31) / 32) * 32;
  if (threads > 1024) return POB_EUNSUPPORTED;
  if (threads < 64) threads = 64;
  const size_t smem = smem_bytes(W, P.NP, P.EMAX);
  if (smem > 200 * 1024) return POB_EUNSUPPORTED;
  void (*kern)(BeamParams);
  const bool ctc = model == POB_MODEL_CTC;
  constexpr int M0 = POB_MODEL_CTC, M1 = POB_MODEL_CTC_MERGE_REPEATS;
  if (threads <= 64) { threads = 64; kern = ctc ? beam_kernel<M0, 64, 12> : beam_kernel<M1, 64, 12>; }
  else if (threads <= 128) { kern = ctc ? beam_kernel<M0, 128, 6> : beam_kernel<M1, 128, 6>; }
  else if (threads <= 288) { kern = ctc ? beam_kernel<M0, 288, 3> : beam_kernel<M1, 288, 3>; }
  else if (threads <= 512) { kern = ctc ? beam_kernel<M0, 512, 1> : beam_kernel<M1, 512, 1>; }
  else { kern = ctc ? beam_kernel<M0, 1024, 1> : beam_kernel<M1, 1024, 1>; }
  POB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // leave most of the unified L1/shared array to L1 (node headers and windows are served from it) but make
  // sure the shared-memory carve-out does not cap residency
  {
    int want_blocks = 2048 / threads;
    if (want_blocks > 12) want_blocks = 12;
    int pct = (int)((want_blocks * (smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
    if (pct > 100) pct = 100;
    POB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
  }
  int per_sm = 0;
  POB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
  if (per_sm < 1) return POB_EUNSUPPORTED;
  const size_t stride = ws_bytes(model, P.NP, P.CAP0, P.CAP1, P.RQ, Umax, Vmax);
  int grid = per_sm * ctx->sm_count;
  if (grid > n_items) grid = n_items;
  // keep the workspace within a sane share of HBM
  const size_t budget = (size_t)96 << 30;
  while (grid > 1 && (size_t)grid * stride > budget) grid = grid * 3 / 4;
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
  {
    pob_prof_scope ps(ctx, mode == MODE_1D ? POB_K_BEAM_1D : POB_K_BEAM_2D);
    kern<<<grid, threads, smem, ctx->stream>>>(P);
  }
  POB_CUDA(cudaGetLastError());
  {
    pob_prof_scope ps(ctx, POB_K_BACKTRACE);
    backtrace_kernel<<<(n_total + 127) / 128, 128, 0, ctx->stream>>>(trace, trace_off, top, out_off, n_total, out_seq,
                                                                    out_len, out_status);
  }
  POB_CUDA(cudaGetLastError());
  return POB_OK;
}
