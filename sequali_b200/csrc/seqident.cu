// seqident.cu -- _seqident.sequence_identity for many (target, query) pairs at once
// (reference _seqidentmodule.c:32-101: Smith-Waterman that carries, next to each cell's score, how many query
// letters the best path into the cell matched; :126-270 is the same recurrence walked along anti-diagonals).
//
// One warp per pair.  A query has at most 31 letters, so lane j owns query letter j and the warp sweeps the
// anti-diagonals d = i + j of the matrix: a cell needs (i-1, j) = this lane one step ago, (i, j-1) = the lane
// below one step ago and (i-1, j-1) = the lane below two steps ago -- two register pairs per lane and two
// shuffles per step, no shared memory.  Target letters are fetched one step ahead.  Cells outside the matrix
// are zero, which is the recurrence's boundary.  The answer is the largest match count among the cells that
// hold the highest score (the reference's running maximum does not depend on the visiting order).
#include "common.cuh"

constexpr int SI_TPB = 128;

__global__ void __launch_bounds__(SI_TPB)
k_seqident(const uint8_t *__restrict__ targets, const uint64_t *__restrict__ target_off, const uint8_t *__restrict__ queries,
           const uint32_t *__restrict__ query_off, uint32_t n_pairs, int match, int mismatch, int deletion, int insertion,
           int32_t *__restrict__ matches_out) {
    const uint32_t warps = gridDim.x * (SI_TPB / 32), lane = lane_id();
    for (uint32_t p = blockIdx.x * (SI_TPB / 32) + (threadIdx.x >> 5); p < n_pairs; p += warps) {
        const uint8_t *target = targets + target_off[p];
        const int64_t t_len = (int64_t)(target_off[p + 1] - target_off[p]);
        const int q_len = (int)(query_off[p + 1] - query_off[p]);
        const int q_char = (int)lane < q_len ? queries[query_off[p] + lane] : -1;
        int s1 = 0, m1 = 0, s2 = 0, m2 = 0, best_s = 0, best_m = 0;
        const int64_t steps = q_len ? t_len + q_len - 1 : 0;
        int64_t i = -(int64_t)lane;  // target index of this lane's cell on diagonal d
        int t_next = i >= 0 && i < t_len ? target[i] : -2;
        for (int64_t d = 0; d < steps; d++, i++) {
            const int t_char = t_next;
            t_next = i + 1 >= 0 && i + 1 < t_len ? target[i + 1] : -2;
            int del_s = __shfl_up_sync(0xffffffffu, s1, 1), del_m = __shfl_up_sync(0xffffffffu, m1, 1);
            int dia_s = __shfl_up_sync(0xffffffffu, s2, 1), dia_m = __shfl_up_sync(0xffffffffu, m2, 1);
            if (lane == 0) del_s = del_m = dia_s = dia_m = 0;  // column 0
            const bool same = t_char == q_char;
            const int lin_s = dia_s + (same ? match : mismatch), lin_m = dia_m + (same ? 1 : 0);
            const int ins_s = s1 + insertion;
            del_s += deletion;
            int s, m;
            if (lin_s >= ins_s && lin_s >= del_s) {
                s = lin_s;
                m = lin_m;
            }
            else if (ins_s >= del_s) {
                s = ins_s;
                m = m1 - 1;  // an inserted letter costs one of the matches that were still possible (:71-75)
            }
            else {
                s = del_s;
                m = del_m;
            }
            if (s < 0) s = 0, m = 0;
            if (t_char < 0 || q_char < 0) s = 0, m = 0;  // outside the matrix
            else if (s == best_s && m > best_m) best_m = m;
            else if (s > best_s) best_s = s, best_m = m;
            s2 = s1;
            m2 = m1;
            s1 = s;
            m1 = m;
        }
        int top = best_s;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) top = max(top, __shfl_xor_sync(0xffffffffu, top, o));
        int most = best_s == top ? best_m : INT_MIN;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) most = max(most, __shfl_xor_sync(0xffffffffu, most, o));
        if (lane == 0) matches_out[p] = most;
    }
}

// Host buffers in, host buffer out (synchronous).  target_off / query_off have n + 1 entries.
extern "C" int sq_sequence_identity_batch(sq_ctx *ctx, const uint8_t *targets, const uint64_t *target_off,
                                          const uint8_t *queries, const uint32_t *query_off, uint64_t n, int match_score,
                                          int mismatch_penalty, int deletion_penalty, int insertion_penalty,
                                          int32_t *matches_out) {
    if (n == 0) return SQ_OK;
    if (n > 0x7fffffffULL) {
        sq_set_error("too many sequence pairs in one call: %llu", (unsigned long long)n);
        return SQ_E_LIMIT;
    }
    for (uint64_t p = 0; p < n; p++)
        if (query_off[p + 1] - query_off[p] > 31) {  // :331-337
            sq_set_error("query %llu has %u letters, at most 31 are supported", (unsigned long long)p,
                         query_off[p + 1] - query_off[p]);
            return SQ_E_ARG;
        }
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = sq_cur_stream(ctx);
    const uint64_t t_bytes = target_off[n], q_bytes = query_off[n];
    uint8_t *d_t = nullptr, *d_q = nullptr;
    uint64_t *d_to = nullptr;
    uint32_t *d_qo = nullptr;
    int32_t *d_out = nullptr;
    auto body = [&]() -> int {
        SQ_TRY(sq_dalloc(ctx, (void **)&d_t, t_bytes + 8, false));
        SQ_TRY(sq_dalloc(ctx, (void **)&d_q, q_bytes + 8, false));
        SQ_TRY(sq_dalloc(ctx, (void **)&d_to, (n + 1) * 8, false));
        SQ_TRY(sq_dalloc(ctx, (void **)&d_qo, (n + 1) * 4, false));
        SQ_TRY(sq_dalloc(ctx, (void **)&d_out, n * 4, false));
        if (t_bytes) CUDA_TRY(cudaMemcpyAsync(d_t, targets, t_bytes, cudaMemcpyHostToDevice, st));
        if (q_bytes) CUDA_TRY(cudaMemcpyAsync(d_q, queries, q_bytes, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_to, target_off, (n + 1) * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_qo, query_off, (n + 1) * 4, cudaMemcpyHostToDevice, st));
        SQ_LAUNCH(ctx, k_seqident, sq_grid_for(ctx, n * 32, SI_TPB, 16), SI_TPB, 0, d_t, d_to, d_q, d_qo, (uint32_t)n,
                  match_score, mismatch_penalty, deletion_penalty, insertion_penalty, d_out);
        CUDA_TRY(cudaMemcpyAsync(matches_out, d_out, n * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        return SQ_OK;
    };
    const int rc = body();
    for (void *p : {(void *)d_t, (void *)d_q, (void *)d_to, (void *)d_qo, (void *)d_out}) sq_dfree(ctx, p);
    return rc;
}
