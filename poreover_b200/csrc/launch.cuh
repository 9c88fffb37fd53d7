// Internal device-pointer launchers shared by the API layer and the fused pair pipeline.
#pragma once
#include "common.cuh"

int pob_viterbi_launch(pob_ctx* ctx, const pob_reads& rd, int kind, uint8_t* out_seq, int32_t* out_s2s,
                       int8_t* out_path, int32_t* out_len, int32_t* out_status);

int pob_nw_slots(int band);
int pob_nw_launch(pob_ctx* ctx, const uint8_t* seq1, const int64_t* off1, const int32_t* len1, const uint8_t* seq2,
                  const int64_t* off2, const int32_t* len2, const int32_t* skip, int n, int band, int match,
                  int mismatch, int gap, int SZ, const int64_t* m_off, int32_t* M, const int64_t* rb_off,
                  int32_t* rowband, const int64_t* aln_off, uint8_t* out_a1, uint8_t* out_a2, int32_t* out_alen,
                  int32_t* out_matches);
int pob_envelope_launch(pob_ctx* ctx, const uint8_t* a1, const uint8_t* a2, const int64_t* aln_off,
                        const int32_t* alen, const int32_t* s2s1, const int64_t* soff1, const int32_t* slen1,
                        const int32_t* s2s2, const int64_t* soff2, const int32_t* slen2, const int32_t* U,
                        const int32_t* V, const int64_t* env_off, const int32_t* skip, int n, int padding,
                        int32_t* env);
int pob_envelope_transpose_launch(pob_ctx* ctx, const int32_t* env, const int64_t* env_off, const int32_t* U,
                                  const int32_t* V, const int64_t* envt_off, const int32_t* skip, int n,
                                  int32_t* envt, int32_t* span);

enum { POB_MODE_1D = 0, POB_MODE_ROW = 1, POB_MODE_ROWCOL = 2 };
int pob_beam_launch(pob_ctx* ctx, const pob_reads& r1, const pob_reads* r2, const int32_t* env,
                    const int64_t* env_off, const int32_t* envt, const int64_t* envt_off, const int32_t* order,
                    const int32_t* skip, int n_items, int n_total, int W, int model, int mode, int max_span0,
                    int max_span1, int Umax, int Vmax, const int64_t* trace_off, uint32_t* trace, int32_t* top,
                    const int64_t* out_off, uint8_t* out_seq, int32_t* out_len, double* out_score,
                    int32_t* out_status);

int pob_flipflop_launch(pob_ctx* ctx, const pob_reads& rd, const double* lut, uint32_t* bp, int8_t* path,
                        uint8_t* out_seq, int32_t* out_s2s, int32_t* out_len);

int pob_forward_launch(pob_ctx* ctx, const pob_reads& rd, const uint8_t* labels, const int64_t* lab_off, int model,
                       double* scratch, const int64_t* scr_off, double* out);
int pob_acceptor_slots(int band);
int pob_acceptor_launch(pob_ctx* ctx, const pob_reads& rd, const uint8_t* labels, const int64_t* lab_off, int band,
                        int SZ, const int64_t* bp_off, uint32_t* bp, double* cum, int8_t* out_path,
                        int32_t* out_status);
int pob_global_pair_launch(pob_ctx* ctx, const uint8_t* seq1, const int64_t* off1, const uint8_t* seq2,
                           const int64_t* off2, int n, int match, int mismatch, int gap, const int64_t* dp_off,
                           int32_t* dp, const int64_t* aln_off, uint8_t* out_a1, uint8_t* out_a2, int32_t* out_alen,
                           int32_t* out_matches);
