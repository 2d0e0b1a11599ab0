// synth.cu -- synthetic NovaSeq-style FASTQ generated on the device (SURVEY.md 8d,
// recipe C2), so that the 100 M-read bench input never crosses PCIe.  Test and
// bench infrastructure: no collector depends on it.
//
// Record r (global index) is a pure function of (seed, r):
//   @SIM:1:FCX:1:<tile>:<x>:<y> 1:N:0:ATCACG      tile = run of reads_per_tile reads
//   <L bases>   uniform ACGT, 0.1 % N; 5 % carry the Illumina adapter from a random
//               3' position; 2 % repeat the bases of an earlier read
//   +
//   <L quals>   read mean ~N(34,4), per-base -U{0..8} and a 3' decay, clamped [2,41]
#include "common.cuh"

constexpr int SY_TPB = 128;

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t rnd(uint64_t seed, uint64_t rec, uint64_t what) {
    return mix64(mix64(seed ^ (rec * 0xD1342543DE82EF95ULL)) + what);
}
__device__ __forceinline__ uint32_t ndigits(uint32_t v) {
    return v >= 100000 ? 6 : v >= 10000 ? 5 : v >= 1000 ? 4 : v >= 100 ? 3 : v >= 10 ? 2 : 1;
}
__device__ __forceinline__ uint8_t *put_dec(uint8_t *p, uint32_t v) {
    uint32_t n = ndigits(v);
    for (uint32_t i = n; i-- > 0;) {
        p[i] = (uint8_t)('0' + v % 10);
        v /= 10;
    }
    return p + n;
}
__device__ __forceinline__ void rec_fields(uint64_t seed, uint64_t g, uint64_t reads_per_tile, uint32_t *tile,
                                           uint32_t *x, uint32_t *y) {
    uint64_t t = (g / reads_per_tile) % 936;  // swaths: a tile comes back after 936 runs
    // {1,2}{1..6}{01..78}
    *tile = (uint32_t)((t / 468 + 1) * 1000 + ((t % 468) / 78 + 1) * 100 + (t % 78) + 1);
    uint64_t r = rnd(seed, g, 1);
    *x = 1000 + (uint32_t)(r % 31000);
    *y = 1000 + (uint32_t)((r >> 32) % 199000);
}

__global__ void __launch_bounds__(SY_TPB)
k_synth_sizes(uint64_t seed, uint64_t first, uint32_t n, uint32_t L, uint64_t reads_per_tile,
              uint32_t *__restrict__ sizes) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t tile, x, y;
        rec_fields(seed, first + i, reads_per_tile, &tile, &x, &y);
        // "@SIM:1:FCX:1:" 13 + tile + ":" + x + ":" + y + " 1:N:0:ATCACG" 13 + "\n" + L + "\n+\n" + L + "\n"
        sizes[i] = 13 + ndigits(tile) + 1 + ndigits(x) + 1 + ndigits(y) + 13 + 1 + L + 3 + L + 1;
    }
}

__device__ __forceinline__ uint8_t base_of(uint64_t seed, uint64_t src, uint32_t pos) {
    uint64_t r = rnd(seed, src, 16 + (pos >> 4));
    uint32_t v = (uint32_t)(r >> ((pos & 15) * 4)) & 15;  // 4 bits per base
    uint64_t n_roll = rnd(seed, src, 1000 + (pos >> 2));
    if (((n_roll >> ((pos & 3) * 16)) & 0xFFFF) < 66) return 'N';  // ~0.1 %
    return (uint8_t)("ACGT"[v & 3]);
}

__global__ void __launch_bounds__(SY_TPB)
k_synth_write(uint64_t seed, uint64_t first, uint32_t n, uint32_t L, uint64_t reads_per_tile,
              const uint32_t *__restrict__ offs, uint8_t *__restrict__ out) {
    const char adapter[] = "AGATCGGAAGAGCACACGTCTGAACTCCAGTCA";
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint64_t g = first + i;
        uint32_t tile, x, y;
        rec_fields(seed, g, reads_per_tile, &tile, &x, &y);
        uint8_t *p = out + offs[i];
        const char h0[] = "@SIM:1:FCX:1:";
        for (int k = 0; k < 13; k++) *p++ = (uint8_t)h0[k];
        p = put_dec(p, tile);
        *p++ = ':';
        p = put_dec(p, x);
        *p++ = ':';
        p = put_dec(p, y);
        const char h1[] = " 1:N:0:ATCACG";
        for (int k = 0; k < 13; k++) *p++ = (uint8_t)h1[k];
        *p++ = '\n';
        const uint64_t roll = rnd(seed, g, 2);
        uint64_t src = g;
        if ((roll & 0xFFFF) < 1311 && g > 0) src = rnd(seed, g, 3) % g;  // 2 % duplicates
        const bool has_adapter = ((roll >> 16) & 0xFFFF) < 3277 && L > 40;  // 5 %
        const uint32_t ad_pos = has_adapter ? 20 + (uint32_t)((roll >> 32) % (L - 32)) : L;
        for (uint32_t k = 0; k < L; k++) {
            uint8_t b = base_of(seed, src, k);
            if (k >= ad_pos && k - ad_pos < 33) b = (uint8_t)adapter[k - ad_pos];
            *p++ = b;
        }
        *p++ = '\n';
        *p++ = '+';
        *p++ = '\n';
        // mean ~ N(34,4): sum of four uniforms, scaled
        const uint64_t m = rnd(seed, g, 4);
        const int mean = 34 + (int)(((m & 0xFF) + ((m >> 8) & 0xFF) + ((m >> 16) & 0xFF) + ((m >> 24) & 0xFF)) / 37) - 13;
        const uint32_t decay = (uint32_t)((m >> 32) & 7);
        for (uint32_t k = 0; k < L; k++) {
            uint64_t r = rnd(seed, g, 2000 + (k >> 4));
            int q = mean - (int)((r >> ((k & 15) * 4)) & 15) % 9 - (int)(decay * k / L);
            q = q < 2 ? 2 : q > 41 ? 41 : q;
            *p++ = (uint8_t)(33 + q);
        }
        *p++ = '\n';
    }
}

// Writes n_reads records starting at global index `first_read` to dev_text.
extern "C" int sq_synth_illumina(sq_ctx *ctx, uint8_t *dev_text, uint64_t cap, uint64_t n_reads,
                                 uint32_t read_length, uint64_t seed, uint64_t first_read,
                                 uint64_t reads_per_tile, uint64_t *nbytes) {
    const uint64_t s = seed & 0xFFFF, first = first_read;
    if (reads_per_tile == 0) reads_per_tile = 1;
    *nbytes = 0;
    if (n_reads == 0) return SQ_OK;
    if (n_reads > (1u << 23) || n_reads * (2ULL * read_length + 64) >= 0xFFFFFF00ULL) {
        sq_set_error("generate at most 8 Mi reads / 4 GiB of text per call");
        return SQ_E_ARG;
    }
    CUDA_TRY(cudaSetDevice(ctx->device));
    const uint32_t n = (uint32_t)n_reads;
    uint32_t *sizes = nullptr, *offs = nullptr;
    SQ_TRY(sq_dalloc(ctx, (void **)&sizes, (size_t)n * 4, false));
    SQ_TRY(sq_dalloc(ctx, (void **)&offs, (size_t)n * 4 + 4, false));
    const int grid = sq_grid_for(ctx, n, SY_TPB, 16);
    SQ_LAUNCH(ctx, k_synth_sizes, grid, SY_TPB, 0, s, first, n, read_length, reads_per_tile, sizes);
    SQ_TRY(sq_scan_exclusive_u32(ctx, sizes, offs, n, offs + n));
    uint32_t *h_total = (uint32_t *)((char *)ctx->h_scratch + 3700);
    CUDA_TRY(cudaMemcpyAsync(h_total, offs + n, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    int rc = SQ_OK;
    if (*h_total > cap) {
        sq_set_error("synthetic text needs %u bytes, buffer has %llu", *h_total, (unsigned long long)cap);
        rc = SQ_E_ARG;
    }
    else {
        SQ_LAUNCH(ctx, k_synth_write, grid, SY_TPB, 0, s, first, n, read_length, reads_per_tile, offs, dev_text);
        *nbytes = *h_total;
    }
    sq_dfree(ctx, sizes);
    sq_dfree(ctx, offs);
    return rc;
}
