// scan.cu -- device-wide exclusive prefix sum over u32 (three small kernels).
// Used wherever a result depends on the global record order (first-come table
// admission, the dedup escalation point, per-tile ordering).
#include "common.cuh"

constexpr int SCAN_TPB = 256;
constexpr int SCAN_ITEMS = 8;  // per thread
constexpr int SCAN_TILE = SCAN_TPB * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_TPB)
k_scan_tile_sums(const uint32_t *__restrict__ in, uint32_t n, uint32_t *__restrict__ tile_sums) {
    __shared__ uint32_t wsum[SCAN_TPB / 32];
    uint32_t base = blockIdx.x * SCAN_TILE;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        uint32_t i = base + k * SCAN_TPB + threadIdx.x;
        if (i < n) s += in[i];
    }
    s = warp_sum_u32(s);
    if (lane_id() == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < SCAN_TPB / 32; w++) t += wsum[w];
        tile_sums[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024)
k_scan_tile_offsets(uint32_t *tile_sums, uint32_t n_tiles, uint32_t *total) {
    __shared__ uint32_t warp_pref[32];
    __shared__ uint32_t carry_s, chunk_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_tiles; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < n_tiles ? tile_sums[i] : 0, wt;
        uint32_t ex = warp_excl_scan_u32(v, &wt);
        if (lane_id() == 0) warp_pref[threadIdx.x >> 5] = wt;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t tt;
            uint32_t e = warp_excl_scan_u32(warp_pref[threadIdx.x], &tt);
            warp_pref[threadIdx.x] = e;
            if (threadIdx.x == 0) chunk_s = tt;
        }
        __syncthreads();
        uint32_t carry = carry_s;
        if (i < n_tiles) tile_sums[i] = carry + warp_pref[threadIdx.x >> 5] + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + chunk_s;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry_s;
}

// element order inside a tile: index = base + k*TPB + tid  (k-major), so the
// per-thread items are NOT contiguous; scan k-slices one after another.
__global__ void __launch_bounds__(SCAN_TPB)
k_scan_apply(const uint32_t *__restrict__ in, uint32_t n, const uint32_t *__restrict__ tile_offsets,
             uint32_t *__restrict__ out) {
    __shared__ uint32_t warp_pref[SCAN_TPB / 32];
    __shared__ uint32_t slice_total;
    uint32_t base = blockIdx.x * SCAN_TILE;
    uint32_t carry = tile_offsets[blockIdx.x];
    for (int k = 0; k < SCAN_ITEMS; k++) {
        uint32_t i = base + k * SCAN_TPB + threadIdx.x;
        uint32_t v = i < n ? in[i] : 0, wt;
        uint32_t ex = warp_excl_scan_u32(v, &wt);
        if (lane_id() == 0) warp_pref[threadIdx.x >> 5] = wt;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t t = threadIdx.x < SCAN_TPB / 32 ? warp_pref[threadIdx.x] : 0, tt;
            uint32_t e = warp_excl_scan_u32(t, &tt);
            if (threadIdx.x < SCAN_TPB / 32) warp_pref[threadIdx.x] = e;
            if (threadIdx.x == 0) slice_total = tt;
        }
        __syncthreads();
        if (i < n) out[i] = carry + warp_pref[threadIdx.x >> 5] + ex;
        carry += slice_total;
        __syncthreads();
    }
}

// out[i] = sum(in[0..i)); *total_dev (device, optional) = sum of all.  in != out allowed to alias.
int sq_scan_exclusive_u32(sq_ctx *ctx, const uint32_t *in, uint32_t *out, uint32_t n, uint32_t *total_dev) {
    if (n == 0) {
        if (total_dev) CUDA_TRY(cudaMemsetAsync(total_dev, 0, 4, sq_cur_stream(ctx)));
        return SQ_OK;
    }
    uint32_t n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    uint32_t *tile_sums = nullptr;
    SQ_TRY(sq_dalloc(ctx, (void **)&tile_sums, (size_t)n_tiles * 4, false));
    SQ_LAUNCH(ctx, k_scan_tile_sums, n_tiles, SCAN_TPB, 0, in, n, tile_sums);
    SQ_LAUNCH(ctx, k_scan_tile_offsets, 1, 1024, 0, tile_sums, n_tiles, total_dev);
    SQ_LAUNCH(ctx, k_scan_apply, n_tiles, SCAN_TPB, 0, in, n, tile_sums, out);
    sq_dfree(ctx, tile_sums);
    return SQ_OK;
}
