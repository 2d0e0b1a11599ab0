// bam.cu -- BAM alignment records -> packed name|seq|qual|tags record array
// (reference BamParser__next__, _qcmodule.c:1623-1694; decode_bam_sequence
// :1266-1334; decode_bam_qualities :1354-1360).
//
// The host walks the block_size chain (one u32 per record) and hands over the
// offsets of the records to keep.  On the device:
//   k_bam_sizes   one thread per record reads the 36-byte header -> field sizes
//   exclusive scan of the packed sizes -> output offsets
//   k_bam_decode  one warp per record: name copy, 4-bit -> ASCII nucleotides
//                 (two bases per input byte), quality + 33 (or '!' when the
//                 qualities are absent, i.e. the first byte is 0xff), raw tags
#include "common.cuh"

constexpr int BAM_TPB = 256;

__device__ __forceinline__ uint32_t brd16(const uint8_t *p) { return p[0] | (uint32_t)p[1] << 8; }
__device__ __forceinline__ uint32_t brd32(const uint8_t *p) {
    return p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24;
}

__global__ void __launch_bounds__(BAM_TPB)
k_bam_sizes(const uint8_t *__restrict__ bam, const uint64_t *__restrict__ rec_off, uint32_t n,
            uint32_t *__restrict__ sizes, uint32_t *name_len, uint32_t *seq_len, uint32_t *tags_len,
            unsigned int *max_len) {
    uint32_t lmax = 0;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const uint8_t *h = bam + rec_off[r];
        const uint32_t block = brd32(h), l_name = h[12], n_cigar = brd16(h + 16), l_seq = brd32(h + 20);
        const uint32_t nl = l_name ? l_name - 1 : 0;  // drop the terminating NUL
        const uint32_t fixed = 36 + l_name + 4 * n_cigar + (l_seq + 1) / 2 + l_seq;
        const uint32_t tl = 4 + block - fixed;
        name_len[r] = nl;
        seq_len[r] = l_seq;
        tags_len[r] = tl;
        sizes[r] = nl + 2 * l_seq + tl;
        lmax = max(lmax, l_seq);
    }
    lmax = warp_max_u32(lmax);
    if (lane_id() == 0 && lmax) atomicMax(max_len, lmax);
}

__global__ void __launch_bounds__(BAM_TPB)
k_bam_decode(const uint8_t *__restrict__ bam, const uint64_t *__restrict__ rec_off, uint32_t n,
             const uint32_t *__restrict__ out_off, const uint32_t *__restrict__ name_len,
             const uint32_t *__restrict__ seq_len, const uint32_t *__restrict__ tags_len, uint8_t *__restrict__ out,
             uint32_t *name_off, uint32_t *seq_off, uint32_t *qual_off, uint32_t *tags_off) {
    const uint32_t warps = gridDim.x * (BAM_TPB / 32);
    const uint32_t lane = lane_id();
    for (uint32_t r = blockIdx.x * (BAM_TPB / 32) + (threadIdx.x >> 5); r < n; r += warps) {
        const uint8_t *h = bam + rec_off[r];
        const uint32_t l_name = h[12], n_cigar = brd16(h + 16);
        const uint32_t nl = name_len[r], sl = seq_len[r], tl = tags_len[r];
        const uint8_t *name = h + 36, *seq = name + l_name + 4 * n_cigar;
        const uint8_t *qual = seq + (sl + 1) / 2, *tags = qual + sl;
        uint8_t *o = out + out_off[r];
        if (lane == 0) {
            name_off[r] = out_off[r];
            seq_off[r] = out_off[r] + nl;
            qual_off[r] = out_off[r] + nl + sl;
            tags_off[r] = out_off[r] + nl + 2 * sl;
        }
        for (uint32_t i = lane; i < nl; i += 32) o[i] = name[i];
        o += nl;
        // "=ACMGRSVTWYHKDBN": code -> letter through two 8-byte tables
        for (uint32_t i = lane; i < sl; i += 32) {
            const uint32_t b = seq[i >> 1];
            const uint32_t code = (i & 1) ? (b & 15) : (b >> 4);
            o[i] = (uint8_t)("=ACMGRSVTWYHKDBN"[code]);
        }
        o += sl;
        const bool missing = sl && qual[0] == 0xff;  // :1658
        for (uint32_t i = lane; i < sl; i += 32) o[i] = missing ? (uint8_t)'!' : (uint8_t)(qual[i] + 33);
        o += sl;
        for (uint32_t i = lane; i < tl; i += 32) o[i] = tags[i];
    }
}

// The block_size chain of BamParser__next__ (_qcmodule.c:1623-1637) over the bytes the caller has read so
// far: offsets of the complete records to keep, how many were dropped for being secondary or
// supplementary alignments (flag & 0x900, :1633), and how many bytes the complete records cover.
extern "C" int sq_bam_walk(const uint8_t *bam, uint64_t nbytes, uint64_t *rec_off, uint64_t cap, uint64_t *n_kept,
                           uint64_t *n_skipped, uint64_t *consumed) {
    uint64_t pos = 0, kept = 0, skipped = 0;
    while (pos + 4 < nbytes) {
        const uint64_t block = (uint64_t)bam[pos] | (uint64_t)bam[pos + 1] << 8 | (uint64_t)bam[pos + 2] << 16 |
                               (uint64_t)bam[pos + 3] << 24;
        const uint64_t end = pos + 4 + block;
        if (end > nbytes) break;
        uint32_t flag = 0;
        if (pos + 18 < nbytes) flag = bam[pos + 18];
        if (pos + 19 < nbytes) flag |= (uint32_t)bam[pos + 19] << 8;
        if (flag & (0x100u | 0x800u)) skipped++;
        else {
            if (kept == cap) {
                sq_set_error("BAM records smaller than their fixed fields");
                return SQ_E_FORMAT;
            }
            rec_off[kept++] = pos;
        }
        pos = end;
    }
    *n_kept = kept;
    *n_skipped = skipped;
    *consumed = pos;
    return SQ_OK;
}

extern "C" int sq_batch_from_bam(sq_ctx *ctx, const uint8_t *bam, uint64_t nbytes, const uint64_t *rec_off,
                                 uint64_t n, sq_batch **out, uint64_t *packed_len) {
    *out = nullptr;
    *packed_len = 0;
    if (nbytes >= 0xC0000000ULL) {  // packed output is at most 4/3 of the input
        sq_set_error("BAM chunk of %llu bytes is too large for one record array", (unsigned long long)nbytes);
        return SQ_E_LIMIT;
    }
    CUDA_TRY(cudaSetDevice(ctx->device));
    SqParserScope on_parser_stream(ctx);
    // validate the chain the host walked before trusting the headers on the device
    for (uint64_t i = 0; i < n; i++) {
        if (rec_off[i] + 36 > nbytes) {
            sq_set_error("BAM record %llu starts outside the buffer", (unsigned long long)i);
            return SQ_E_ARG;
        }
        const uint8_t *h = bam + rec_off[i];
        uint64_t block = (uint64_t)h[0] | (uint64_t)h[1] << 8 | (uint64_t)h[2] << 16 | (uint64_t)h[3] << 24;
        uint64_t l_name = h[12], n_cigar = h[16] | (uint64_t)h[17] << 8;
        uint64_t l_seq = (uint64_t)h[20] | (uint64_t)h[21] << 8 | (uint64_t)h[22] << 16 | (uint64_t)h[23] << 24;
        uint64_t fixed = 36 + l_name + 4 * n_cigar + (l_seq + 1) / 2 + l_seq;
        if (rec_off[i] + 4 + block > nbytes || fixed > 4 + block) {
            sq_set_error("BAM record %llu is inconsistent with its block_size", (unsigned long long)i);
            return SQ_E_FORMAT;
        }
    }
    sq_batch *b = new sq_batch();
    b->ctx = ctx;
    b->n = n;
    uint8_t *d_bam = nullptr;
    uint64_t *d_off = nullptr;
    uint32_t *sizes = nullptr, *offs = nullptr;
    const size_t n4 = (size_t)((n + 3) & ~3ULL);
    const size_t meta_bytes = n4 * 4 * 7 + n4 * 8;
    int rc = sq_dalloc(ctx, (void **)&d_bam, nbytes + 64, false);
    if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&d_off, (n + 1) * 8, false);
    if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&sizes, (n + 1) * 4, false);
    if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&offs, (n + 1) * 4, false);
    if (rc == SQ_OK) rc = sq_dalloc(ctx, &b->meta_block, meta_bytes, true);
    unsigned int *d_max = (unsigned int *)((char *)ctx->d_scratch + 3600);
    uint32_t *d_total = (uint32_t *)((char *)ctx->d_scratch + 3604);
    uint32_t *h_res = (uint32_t *)((char *)ctx->h_scratch + 3600);
    if (rc == SQ_OK && n) {
        uint32_t *d = (uint32_t *)b->meta_block;
        b->name_off = d;
        b->seq_off = d + n4;
        b->seq_len = d + 2 * n4;
        b->qual_off = d + 3 * n4;
        b->name_len = d + 4 * n4;
        b->tags_off = d + 5 * n4;
        b->tags_len = d + 6 * n4;
        b->err_sum = (double *)(d + 7 * n4);
        CUDA_TRY(cudaMemcpyAsync(d_bam, bam, nbytes, cudaMemcpyHostToDevice, sq_cur_stream(ctx)));
        CUDA_TRY(cudaMemcpyAsync(d_off, rec_off, n * 8, cudaMemcpyHostToDevice, sq_cur_stream(ctx)));
        CUDA_TRY(cudaMemsetAsync(d_max, 0, 8, sq_cur_stream(ctx)));
        const int grid = sq_grid_for(ctx, n, BAM_TPB, 16);
        SQ_LAUNCH(ctx, k_bam_sizes, grid, BAM_TPB, 0, d_bam, d_off, (uint32_t)n, sizes, b->name_len, b->seq_len,
                  b->tags_len, d_max);
        rc = sq_scan_exclusive_u32(ctx, sizes, offs, (uint32_t)n, d_total);
        if (rc == SQ_OK) {
            CUDA_TRY(cudaMemcpyAsync(h_res, d_max, 8, cudaMemcpyDeviceToHost, sq_cur_stream(ctx)));
            CUDA_TRY(cudaStreamSynchronize(sq_cur_stream(ctx)));
            b->max_len = h_res[0];
            b->nbytes = h_res[1];
            rc = sq_dalloc(ctx, (void **)&b->text, b->nbytes + 64, false);
        }
        if (rc == SQ_OK) {
            CUDA_TRY(cudaMemsetAsync(b->text + b->nbytes, 0, 64, sq_cur_stream(ctx)));
            const int wgrid = sq_grid_for(ctx, n * 32, BAM_TPB, 16);
            SQ_LAUNCH(ctx, k_bam_decode, wgrid, BAM_TPB, 0, d_bam, d_off, (uint32_t)n, offs, b->name_len, b->seq_len,
                      b->tags_len, b->text, b->name_off, b->seq_off, b->qual_off, b->tags_off);
        }
    }
    sq_dfree(ctx, d_bam);
    sq_dfree(ctx, d_off);
    sq_dfree(ctx, sizes);
    sq_dfree(ctx, offs);
    // the record array is used on the launch stream from here on: the decode (parser stream) is done first
    if (rc == SQ_OK && cudaStreamSynchronize(sq_cur_stream(ctx)) != cudaSuccess)
        rc = sq_cuda_fail(cudaGetLastError(), "BAM decode", __FILE__, __LINE__);
    if (rc != SQ_OK) {
        sq_batch_free(b);
        return rc;
    }
    *packed_len = b->nbytes;
    *out = b;
    return SQ_OK;
}
