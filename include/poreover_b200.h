/* poreover_b200 -- C ABI of the B200-native PoreOver decoding backend.
 *
 * This is the drop-in boundary for the reference's native decoding path.  Every entry point replaces
 * one function the reference reaches through Cython / numpy today; the citation after "replaces:" is
 * the reference interface (file:line relative to the reference tree).  All entry points are batched:
 * item r of a batch lives at offsets[r] of a packed buffer.  A single read / pair is a batch of one.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ / torch types.
 *   - `where` selects the memory space of EVERY data pointer of that call:
 *       POB_HOST   : host memory.  The library stages H2D, runs, stages D2H and synchronises.
 *       POB_DEVICE : device memory of the context's GPU.  Work is enqueued on the context's stream; use
 *                    pob_ctx_sync() before reading results.  The single-stage calls (pob_viterbi, pob_align_*,
 *                    pob_build_envelope, pob_forward, pob_viterbi_acceptor) only enqueue.  The searches and
 *                    pob_pair_decode also read small per-item arrays back (read lengths, skip decisions, the widest
 *                    envelope band, which size the search's windows) and therefore synchronise the stream up to
 *                    three times inside the call; run two calls on two contexts to keep the GPU busy across them.
 *   - log-probability matrices ("reads") travel as a pob_reads_t descriptor (below).
 *   - `layout` tells where the blank column is stored: POB_BLANK_LAST (reference in-memory order,
 *     A C G T blank) or POB_BLANK_FIRST (bonito .npy file order; replaces decode.py:79).
 *   - rc[r] != 0 asks for the reverse-complement VIEW of read r (time reversed, A<->T, C<->G;
 *     replaces transducer.py:68-70, :79-81, :104-106) without materialising it.  rc may be NULL.
 *   - a slice [lo, hi) of a packed batch is itself a batch: row_off + lo, row_len + lo, rc + lo, n = hi - lo, the
 *     same data pointer (row offsets stay absolute).  Packed outputs are addressed by the same absolute offsets, so
 *     the caller passes the full-size output buffers and the per-item arrays offset by lo.
 *   - the searches take alphabets of one to four letters (n_states = letters + blank, 2..5) and beam widths 1..100.
 *   - every function returns POB_OK (0) or a negative POB_E* status; pob_strerror() names it.
 *   - the caller owns all inputs and pre-sized outputs; device scratch is owned by the context and
 *     reused across calls.
 */
#ifndef POREOVER_B200_H
#define POREOVER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define POB_ABI_VERSION 1

enum { POB_HOST = 0, POB_DEVICE = 1 };
enum { POB_F32 = 0, POB_F64 = 1 };
enum { POB_BLANK_LAST = 0, POB_BLANK_FIRST = 1 };
/* transducer kinds (transducer.py:64, :75, :91) */
enum { POB_KIND_POREOVER = 0, POB_KIND_BONITO = 1, POB_KIND_FLIPFLOP = 2 };
/* prefix-tree models (decode.py:172, pair_decode.py:147) */
enum { POB_MODEL_CTC = 0, POB_MODEL_CTC_MERGE_REPEATS = 1 };
/* 2D traversal (BeamSearch.h:411-437) */
enum { POB_METHOD_ROW = 0, POB_METHOD_ROW_COL = 1 };

enum {
  POB_OK = 0,
  POB_EINVAL = -1,      /* bad argument (NULL pointer, negative size, unknown enum) */
  POB_ECUDA = -2,       /* CUDA runtime error; pob_last_cuda_error() has the text */
  POB_ENOMEM = -3,      /* device or host allocation failed */
  POB_EALIGN = -4,      /* reserved */
  POB_EUNSUPPORTED = -5 /* legal in the reference but outside this build (beam width > 100, more than 5 states) */
};

/* per-item status bits written by the searches (out_status) */
enum {
  POB_ST_SHORT_BEAM_SKIP = 1, /* a skip step ran with fewer than W nodes: undefined behaviour in the
                                 reference (BeamSearch.h:317, :331); we update the nodes that exist */
  POB_ST_UNSET_BAND = 2,      /* band start left uninitialised by the reference (BeamSearch.h:309) */
  POB_ST_POOL_OVERFLOW = 4,   /* node pool recycled a live node: result may deviate from the reference */
  POB_ST_MAPPING_WRAP = 8,    /* get_sequence_mapping's path[-1] wrap dropped the first base
                                 (pair_decode.py:136): the reference asserts and drops the pair */
  POB_ST_SKIPPED_LENGTH = 16, /* |len1-len2| > 1000 (pair_decode.py:372-375) */
  POB_ST_SKIPPED_IDENTITY = 32, /* alignment identity < 0.5 (pair_decode.py:395-398) */
  POB_ST_EMPTY = 64,          /* zero-length input */
  POB_ST_MAX_DEPTH = 128      /* legacy prefix search: the prefix outgrew the reads ("Max search depth exceeded",
                                 prefix_search.py:279-281) */
};

typedef struct pob_ctx pob_ctx;

/* A packed batch of reads (log-probability matrices), all pointers in the call's memory space.
 *   data    : rows x n_states, row-major, dtype POB_F32 / POB_F64
 *   row_off : [n+1] int64 first row of each read (units: rows)
 *   row_len : [n] int32 number of rows of each read, or NULL meaning row_off[r+1]-row_off[r].
 *             Lets the packer start every read on a 4-row (16-byte) boundary for the float4 fast path;
 *             a read that does not start 16-byte aligned is still decoded, on a slower load path.
 *   rc      : [n] non-zero = decode the reverse-complement view of the read, or NULL */
typedef struct pob_reads_t {
  const void* data;
  const int64_t* row_off;
  const int32_t* row_len;
  const uint8_t* rc;
  int32_t n;
  int32_t n_states;
  int32_t dtype;
  int32_t layout;
} pob_reads_t;

int pob_abi_version(void);
const char* pob_strerror(int status);
const char* pob_last_cuda_error(void);
int pob_device_count(int* n);

/* One context per GPU: owns a stream and a growable device scratch arena. */
int pob_ctx_create(int device, pob_ctx** ctx);
int pob_ctx_destroy(pob_ctx* ctx);
int pob_ctx_sync(pob_ctx* ctx);
/* raw cudaStream_t of the context, for callers that order their own work against it */
void* pob_ctx_stream(pob_ctx* ctx);
int pob_ctx_device(pob_ctx* ctx);

/* Device memory helpers for callers without their own allocator (bench, tests). */
int pob_malloc(pob_ctx* ctx, size_t bytes, void** dptr);
int pob_free(pob_ctx* ctx, void* dptr);
int pob_malloc_host(size_t bytes, void** hptr); /* pinned */
int pob_free_host(void* hptr);
int pob_memcpy_h2d(pob_ctx* ctx, void* dst, const void* src, size_t bytes); /* async on ctx stream */
int pob_memcpy_d2h(pob_ctx* ctx, void* dst, const void* src, size_t bytes); /* async on ctx stream */

/* Per-kernel device timing with CUDA events on the context's stream.  ids: POB_K_* */
enum {
  POB_K_VITERBI = 0, POB_K_FLIPFLOP = 1, POB_K_NW_FILL = 2, POB_K_NW_TRACE = 3, POB_K_ENVELOPE = 4,
  POB_K_BEAM_2D = 5, POB_K_BEAM_1D = 6, POB_K_BACKTRACE = 7, POB_K_FORWARD = 8, POB_K_ACCEPTOR = 9,
  POB_K_PREFIX_1D = 10, POB_K_PAIR_GAMMA = 11, POB_K_PREFIX_2D = 12, POB_K_COUNT = 13
};
/* whole-region device timing: CUDA events recorded on the context's stream */
int pob_timer_start(pob_ctx* ctx);
int pob_timer_stop(pob_ctx* ctx, double* ms); /* records the stop event, synchronises, returns elapsed ms */
int pob_profile_enable(pob_ctx* ctx, int on);
int pob_profile_reset(pob_ctx* ctx);
/* synchronises, then returns accumulated device milliseconds and launch count of kernel `id` */
int pob_profile_get(pob_ctx* ctx, int id, double* ms, int64_t* launches);
const char* pob_kernel_name(int id);

/* ---------------------------------------------------------------------------------------------
 * Best-path ("Viterbi") decode of CTC reads, fused with the base -> timestep mapping.
 * replaces: transducer.poreover.viterbi_decode (transducer.py:72-73, :27-33),
 *           transducer.bonito.viterbi_decode   (transducer.py:83-89),
 *           get_sequence_mapping               (pair_decode.py:114-142)
 * Outputs are packed by the INPUT row offsets (a read of T rows decodes to at most T bases):
 *   out_seq  [total_rows] ASCII bases of read r at out_seq + row_off[r]
 *   out_s2s  [total_rows] int32 timestep of each base (sequence_to_signal), same packing; may be NULL
 *   out_path [total_rows] int8 argmax state per timestep (blank = n_states-1), may be NULL
 *   out_len  [n]          decoded length
 *   out_status[n]         POB_ST_MAPPING_WRAP if the reference's mapping would drop the first base
 */
int pob_viterbi(pob_ctx* ctx, int where, const pob_reads_t* reads, int kind, uint8_t* out_seq, int32_t* out_s2s,
                int8_t* out_path, int32_t* out_len, int32_t* out_status);

/* Flip-flop Viterbi: 8-state FP64 max-sum DP with the additive 0/1 transition matrix.
 * replaces: transducer.viterbi_decode (transducer.py:35-59, :94-103) + remove_repeated (:4-9) + the
 *           flipflop branch of get_sequence_mapping (pair_decode.py:124-133)
 * data: POB_F64 log-probabilities (rows x 8), or (dtype == POB_U8_TRACE) raw uint8 traces with
 * lut[256] = log((x+1e-7)/(255+1e-7)) computed by the host (decode.py:92-93) so no log runs on device. */
#define POB_U8_TRACE 2
int pob_viterbi_flipflop(pob_ctx* ctx, int where, const pob_reads_t* reads, const double* lut, uint8_t* out_seq,
                         int32_t* out_s2s, int8_t* out_path, int32_t* out_len);

/* ---------------------------------------------------------------------------------------------
 * Banded Needleman-Wunsch, bit-exact with the reference including its boundary quirks.
 * replaces: align.global_pair_banded (align.pyx:100-178) over SparseMatrix<int> (SparseMatrix.h:61-117)
 * seq1/seq2: packed ASCII, off1/off2 [n+1] int64.  Pair p's gapped rows are written at
 * out_a1/out_a2 + aln_off[p] where aln_off[p] = off1[p] + off2[p] + 8*p (capacity l1+l2+8);
 * out_alen[p] = alignment length, out_matches[p] = number of columns with equal characters.
 * A pair with l1 == 0 gets out_alen = -1 (ZeroDivisionError in the reference, align.pyx:122). */
int pob_align_banded(pob_ctx* ctx, int where, const uint8_t* seq1, const int64_t* off1, const uint8_t* seq2,
                     const int64_t* off2, int n_pairs, int band_width, int match, int mismatch, int gap_cost,
                     uint8_t* out_a1, uint8_t* out_a2, int32_t* out_alen, int32_t* out_matches);

/* Full (unbanded) Needleman-Wunsch, the `--alignment full` path.
 * replaces: align.global_pair (align.pyx:29-98).  Same output packing as pob_align_banded.  out_dp (may be
 * NULL) receives each pair's (l1+1) x (l2+1) int32 DP matrix at dp_off[p] (the third return value of the
 * reference function). */
int pob_align_global(pob_ctx* ctx, int where, const uint8_t* seq1, const int64_t* off1, const uint8_t* seq2,
                     const int64_t* off2, int n_pairs, int match, int mismatch, int gap_cost, uint8_t* out_a1,
                     uint8_t* out_a2, int32_t* out_alen, int32_t* out_matches, const int64_t* dp_off,
                     int32_t* out_dp);

/* Alignment columns -> per-timestep envelope over read 2.
 * replaces: envelope.get_alignment_columns (envelope.py:26-44) + envelope.build_envelope (:46-87)
 * a1/a2 packed gapped rows at aln_off[p] with length alen[p]; s2s1/s2s2 int32 packed at soff1/soff2
 * with slen1/slen2 entries; U[p], V[p] timesteps.  out_env: int32 (sum U) x 2 at env_off[p] rows. */
int pob_build_envelope(pob_ctx* ctx, int where, const uint8_t* a1, const uint8_t* a2, const int64_t* aln_off,
                       const int32_t* alen, const int32_t* s2s1, const int64_t* soff1, const int32_t* slen1,
                       const int32_t* s2s2, const int64_t* soff2, const int32_t* slen2, const int32_t* U,
                       const int32_t* V, const int64_t* env_off, int n_pairs, int padding, int32_t* out_env);

/* ---------------------------------------------------------------------------------------------
 * CTC prefix beam search on single reads.
 * replaces: decoding_cpp.cpp_beam_search (decoding_cpp.pyx:88-103) -> beam_search (BeamSearch.h:400)
 *           -> beam_search_ (:18-58) with PoreOverPrefixTree (PrefixTree.h:461) / BonitoPrefixTree (:635)
 * out_seq packed by row_off (ASCII), out_len[n], out_score[n] = last_probability() of the returned node. */
int pob_beam_search(pob_ctx* ctx, int where, const pob_reads_t* reads, int beam_width, int model,
                    uint8_t* out_seq, int32_t* out_len, double* out_score, int32_t* out_status);

/* Joint two-read prefix beam search inside an alignment envelope.
 * replaces: decoding_cpp.cpp_beam_search_2d (decoding_cpp.pyx:107-139) -> beam_search (BeamSearch.h:411,
 *           :440) -> beam_search_2d_by_row_col (:262-397) / beam_search_2d_by_row (:110-172, :175-260)
 * env: int32 rows x 2 packed at env_off[p] (rows = U[p]); env == NULL means "no envelope" (method row only).
 * out_seq: pair p at out_off[p] (capacity >= U+V+1, caller-chosen packing), out_score = ranking score of
 * the returned node (max_probability_sym for row_col, max_probability for row). */
int pob_beam_search_2d(pob_ctx* ctx, int where, const pob_reads_t* reads1, const pob_reads_t* reads2,
                       const int32_t* env, const int64_t* env_off, int beam_width, int model, int method,
                       const int64_t* out_off, uint8_t* out_seq, int32_t* out_len, double* out_score,
                       int32_t* out_status);

/* Exact label forward log-probability.
 * replaces: decoding_cpp.cpp_forward (decoding_cpp.pyx:49-65) -> forward (PrefixTree.h:710-759)
 * labels: packed base indices 0..3 at lab_off[n+1]. */
int pob_forward(pob_ctx* ctx, int where, const pob_reads_t* reads, const uint8_t* labels, const int64_t* lab_off,
                int model, double* out_logp);

/* Banded Viterbi acceptor ("re-squiggle"): best alignment of a known base sequence to a read.
 * replaces: decoding_cpp.cpp_viterbi_acceptor (decoding_cpp.pyx:69-84) -> viterbi_acceptor_poreover (Forward.h:14-121),
 *           called by pair-decode --single beam (pair_decode.py:363-370) with band_size = 1000.
 * labels: packed base indices 0..3 at lab_off[n+1].  out_path: int8 per timestep, packed by reads->row_off:
 * n_states-1 (blank) or the base index emitted there.  out_status[n]: POB_ST_UNSET_BAND where the reference's
 * traceback leaves the matrix with labels unplaced (it does not terminate there); POB_ST_EMPTY for empty reads. */
int pob_viterbi_acceptor(pob_ctx* ctx, int where, const pob_reads_t* reads, const uint8_t* labels,
                         const int64_t* lab_off, int band_size, int8_t* out_path, int32_t* out_status);

/* ---------------------------------------------------------------------------------------------
 * The whole pair-decode hot path on the device, no host round trip between stages.
 * replaces: pair_decode_helper's envelope path (pair_decode.py:361-398, :495-511):
 *   viterbi x2 -> get_sequence_mapping x2 -> length check -> global_pair_banded -> identity check ->
 *   get_alignment_columns -> build_envelope -> cpp_beam_search_2d(method).
 * Read 1 of pair p is read p of batch 1, read 2 is read p of batch 2 (rc2 gives --reverse_complement).
 * Outputs: basecalls packed by their row offsets (as pob_viterbi), consensus of pair p at
 * out_cons + row_off1[p] + row_off2[p] (capacity U+V); out_stats[p*4..] = {len1, len2, matches, columns};
 * out_status carries POB_ST_SKIPPED_* exactly where the reference skips the pair. */
int pob_pair_decode(pob_ctx* ctx, int where, const pob_reads_t* reads1, const pob_reads_t* reads2, int kind,
                    int beam_width, int padding, int band_width, int method, uint8_t* out_seq1, int32_t* out_len1,
                    uint8_t* out_seq2, int32_t* out_len2, uint8_t* out_cons, int32_t* out_cons_len,
                    double* out_score, int32_t* out_stats, int32_t* out_status);

/* ---------------------------------------------------------------------------------------------
 * Host-side batched reader of .npy probability tables (no GPU work; native threads instead of one Python-level
 * np.load per file).
 * replaces: the np.load of decode.load_logits (decode.py:41-51) / model_from_trace's bonito branch (decode.py:76-80)
 *           for plain 2-D little-endian float32 C-order files; the logarithm (decode.py:45) stays with the caller.
 * pob_npy_probe: per file, rows / cols / byte offset of the data and a flag (POB_NPY_*); first_row_sum[i] (may be NULL)
 *   is the float32 left-to-right sum of the first row, what np.isclose(np.sum(row), 1) looks at (decode.py:43).
 * pob_npy_read : copies file i's rows[i] x cols float32 payload to dst + row_off[i] * cols and zeroes the alignment
 *   rows up to row_off[i + 1]; ok[i] (may be NULL) = 1 when the whole payload was read. */
enum { POB_NPY_F32_2D = 0, POB_NPY_OTHER = 1, POB_NPY_UNREADABLE = 2 };
int pob_npy_probe(const char* const* paths, int n, int threads, int64_t* rows, int64_t* cols, int64_t* data_off,
                  int32_t* flags, float* first_row_sum);
int pob_npy_read(const char* const* paths, int n, int threads, const int64_t* rows, int64_t cols, const int64_t* data_off,
                 const int64_t* row_off, float* dst, int32_t* ok);

/* ---------------------------------------------------------------------------------------------
 * Legacy prefix search (decode --algorithm prefix; the functions of decoding/prefix_search.py).  Inputs are float64
 * log-probabilities, row-major rows x n_states with the blank LAST, packed at row_off (row_off[0] = 0); letters are
 * indices 0..n_states-2.  `flavour` selects the reference's arithmetic: POB_PREFIX_NUMPY (np.logaddexp, scipy
 * logsumexp, LOG_0 = -inf: prefix_search_log, pair_gamma_log, pair_prefix_search_log) or POB_PREFIX_CY (the Cython
 * helpers' log(exp(a) + exp(b)) and -9999: prefix_search_log_cy, decoding_cy.pair_gamma_log, pair_prefix_search_log_cy).
 * Labels come back as letter indices packed at lab_off (capacity per item >= rows + 2, for pairs >= max(U, V) + 3).
 *
 * pob_prefix_search      replaces: prefix_search.prefix_search_log / _cy (prefix_search.py:116-174 / :176-238), one
 *                        window per item; out_score = label probability of the returned label.
 * pob_pair_gamma         replaces: prefix_search.pair_gamma_log (prefix_search.py:35-65) /
 *                        decoding_cy.pair_gamma_log (decoding_cy.pyx:177-220): dense (U+1) x (V+1) matrix of pair p at
 *                        gamma_off[p] (gamma_off[p+1] - gamma_off[p] = (U+1)(V+1)).
 * pob_pair_prefix_search replaces: prefix_search.pair_prefix_search_log / _cy (prefix_search.py:247-310 / :312-385);
 *                        POB_EUNSUPPORTED when a pair's dense matrix would exceed 1 GB (MEM_LIMIT, pair_decode.py:189).
 * The C++ envelope variant (PairPrefixSearch.cpp:79-229) is not reproduced: it double-frees (Gamma.h:100).
 * pob_forward_vec        replaces: prefix_search.forward_vec_log (prefix_search.py:81-97) /
 *                        decoding_cy.forward_vec_log (decoding_cy.pyx:127-156): one column of the 1D forward algorithm
 *                        for letter s (negative = counted from the end, -1 the blank) at label position i, from the
 *                        previous column (NULL when i == 0). */
enum { POB_PREFIX_NUMPY = 0, POB_PREFIX_CY = 1 };
int pob_forward_vec(pob_ctx* ctx, int where, const double* y, int rows, int n_states, int flavour, int s, int i,
                    const double* previous, double* out);
int pob_prefix_search(pob_ctx* ctx, int where, const double* y, const int64_t* row_off, int n, int n_states,
                      int flavour, const int64_t* lab_off, uint8_t* out_label, int32_t* out_len, double* out_score,
                      int32_t* out_status);
int pob_pair_gamma(pob_ctx* ctx, int where, const double* y1, const int64_t* off1, const double* y2,
                   const int64_t* off2, int n, int n_states, int flavour, const int64_t* gamma_off, double* out_gamma);
int pob_pair_prefix_search(pob_ctx* ctx, int where, const double* y1, const int64_t* off1, const double* y2,
                           const int64_t* off2, int n, int n_states, int flavour, const int64_t* lab_off,
                           uint8_t* out_label, int32_t* out_len, double* out_score, int32_t* out_status);

/* Counters of the last pob_pair_decode / pob_beam_search_2d call on this context (for roofline math):
 * [0] forward cell updates (update_prob calls), [1] search steps, [2] kernels launched since reset. */
int pob_counters(pob_ctx* ctx, int64_t* out3);

#ifdef __cplusplus
}
#endif
#endif
