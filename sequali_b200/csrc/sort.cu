// sort.cu -- stable LSD radix sort of (u32 key, u32 value) pairs, 8 bits per pass.
// Used to group the reads of a record array by tile while keeping their order
// (PerTileQuality's floating-point sums are order dependent).
#include "common.cuh"

constexpr int RS_TPB = 256;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_TPB * RS_ITEMS;

__global__ void __launch_bounds__(RS_TPB)
k_rs_hist(const uint32_t *__restrict__ keys, uint32_t n, uint32_t shift, uint32_t *__restrict__ hist,
          uint32_t n_blocks) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * RS_TILE;
#pragma unroll
    for (int k = 0; k < RS_ITEMS; k++) {
        uint32_t i = base + k * RS_TPB + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 0xFF], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * n_blocks + blockIdx.x] = h[threadIdx.x];  // digit-major
}

__global__ void __launch_bounds__(RS_TPB)
k_rs_scatter(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
             uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, uint32_t n, uint32_t shift,
             const uint32_t *__restrict__ offsets, uint32_t n_blocks) {
    __shared__ uint32_t digit_base[256];
    digit_base[threadIdx.x] = offsets[(size_t)threadIdx.x * n_blocks + blockIdx.x];
    __syncthreads();
    const uint32_t base = blockIdx.x * RS_TILE, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane_id()) - 1;
    for (int k = 0; k < RS_ITEMS; k++) {
        const uint32_t i = base + k * RS_TPB + threadIdx.x;
        const bool active = i < n;
        uint32_t key = 0, val = 0, digit = 0x100;  // inactive lanes share a digit nobody writes
        if (active) {
            key = keys_in[i];
            val = vals_in[i];
            digit = (key >> shift) & 0xFF;
        }
        const uint32_t peers = __match_any_sync(0xffffffffu, digit);
        const uint32_t rank = __popc(peers & lt_mask);
        // warps take their turn in order so that earlier elements get earlier slots
        for (uint32_t w = 0; w < RS_TPB / 32; w++) {
            if (warp == w && active) {
                const uint32_t b = digit_base[digit];
                const uint32_t dst = b + rank;
                keys_out[dst] = key;
                vals_out[dst] = val;
                __syncwarp(peers);
                if (rank == 0) digit_base[digit] = b + __popc(peers);
            }
            __syncthreads();
        }
    }
}

// Sorts by the low `key_bits` bits of the keys.  keys/vals are overwritten with
// the sorted sequence (tmp_* are scratch of the same size).
int sq_radix_sort_pairs(sq_ctx *ctx, uint32_t *keys, uint32_t *vals, uint32_t *tmp_keys, uint32_t *tmp_vals,
                        uint32_t n, uint32_t key_bits) {
    if (n < 2) return SQ_OK;
    const uint32_t n_blocks = (n + RS_TILE - 1) / RS_TILE;
    uint32_t *hist = nullptr;
    SQ_TRY(sq_dalloc(ctx, (void **)&hist, (size_t)256 * n_blocks * 4, false));
    uint32_t *ki = keys, *vi = vals, *ko = tmp_keys, *vo = tmp_vals;
    uint32_t passes = (key_bits + 7) / 8;
    if (passes == 0) passes = 1;
    for (uint32_t p = 0; p < passes; p++) {
        SQ_LAUNCH(ctx, k_rs_hist, n_blocks, RS_TPB, 0, ki, n, p * 8, hist, n_blocks);
        SQ_TRY(sq_scan_exclusive_u32(ctx, hist, hist, 256 * n_blocks, nullptr));
        SQ_LAUNCH(ctx, k_rs_scatter, n_blocks, RS_TPB, 0, ki, vi, ko, vo, n, p * 8, hist, n_blocks);
        uint32_t *t = ki; ki = ko; ko = t;
        t = vi; vi = vo; vo = t;
    }
    if (ki != keys) {
        CUDA_TRY(cudaMemcpyAsync(keys, ki, (size_t)n * 4, cudaMemcpyDeviceToDevice, sq_cur_stream(ctx)));
        CUDA_TRY(cudaMemcpyAsync(vals, vi, (size_t)n * 4, cudaMemcpyDeviceToDevice, sq_cur_stream(ctx)));
    }
    sq_dfree(ctx, hist);
    return SQ_OK;
}
