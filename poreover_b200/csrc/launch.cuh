// Internal device-pointer launchers shared by the API layer and the fused pair pipeline.
#pragma once
#include "common.cuh"

int pob_viterbi_launch(pob_ctx* ctx, const pob_reads& rd, int kind, uint8_t* out_seq, int32_t* out_s2s,
                       int8_t* out_path, int32_t* out_len, int32_t* out_status);
