// inflate.cu -- BGZF blocks inflated on the device.
//
// The reference leaves decompression to xopen's host threads (src/sequali/util.py:108-123) and
// documents it as the bottleneck (README.rst:168-171).  A BGZF stream (bgzip'd FASTQ, every BAM) is a
// chain of independent gzip members of at most 64 KiB of text, each announcing its compressed size in
// a 'BC' extra field and its text size in the trailer: the host only hops over the headers
// (sq_bgzf_scan), the compressed bytes cross PCIe (2.5-4x fewer than the text), and one WARP inflates
// one block: all 32 lanes run the bit-serial decoder in lockstep on the same data (uniform control
// flow, broadcast loads -- the cost of one lane), which lets every LZ77 match be copied by the
// whole warp; the Huffman tables of the block live in the warp's slice of shared memory.
// CRC-32 of the members is not checked (the text size of the trailer and the decoder's own
// consistency checks are); a corrupt block fails the record array it belongs to.
#include "inflate_core.cuh"
#include "modules.cuh"

constexpr int INF_WARPS = 8;

struct InfSmem {
    InfTables t;
};

__global__ void __launch_bounds__(INF_WARPS * 32)
k_bgzf_inflate(const uint8_t *__restrict__ comp, uint64_t comp_base, const sq_bgzf_block *__restrict__ blocks, uint32_t n_blocks,
               uint64_t text_base, uint8_t *__restrict__ out, unsigned long long *first_bad) {
    __shared__ InfSmem sm[INF_WARPS];
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t warps = gridDim.x * INF_WARPS;
    for (uint32_t bi = blockIdx.x * INF_WARPS + warp; bi < n_blocks; bi += warps) {
        const sq_bgzf_block blk = blocks[bi];
        uint8_t *dst = out + (blk.text_off - text_base);
        uint32_t produced = 0;
        auto copy_match = [&](uint32_t op, uint32_t dist, uint32_t len) {
            // out[op + i] = out[op - dist + (i mod dist)]: the source lies in front of `op` and is
            // complete, so the lanes can take the bytes of the match in any order
            __syncwarp();
            const uint8_t *src = dst + op - dist;
            if (dist >= len)
                for (uint32_t i = lane; i < len; i += 32) dst[op + i] = src[i];
            else
                for (uint32_t i = lane; i < len; i += 32) dst[op + i] = src[i % dist];
            __syncwarp();
        };
        int rc = inf_inflate(comp + (blk.comp_off - comp_base), blk.comp_len, dst, blk.text_len, &produced, sm[warp].t, copy_match);
        if (rc == INF_OK && produced != blk.text_len) rc = INF_E_SIZE;
        if (rc != INF_OK && lane == 0) atomicMin(first_bad, (unsigned long long)bi << 8 | (unsigned)rc);
        __syncwarp();
    }
}

// ---- host: hop over the member headers ---------------------------------------------------------------
// RFC 1952 member with the BGZF extra subfield (SAM spec 4.1): ID1 ID2 CM FLG MTIME(4) XFL OS XLEN(2)
// [SI1='B' SI2='C' SLEN=2 BSIZE(2)] ... CDATA CRC32(4) ISIZE(4); BSIZE = total member size - 1.
extern "C" int sq_bgzf_scan(const uint8_t *host, uint64_t nbytes, sq_bgzf_block *blocks, uint64_t cap, uint64_t *n_blocks,
                            uint64_t *consumed, uint64_t *text_bytes) {
    uint64_t pos = 0, n = 0, text = 0;
    while (pos + 18 <= nbytes) {
        const uint8_t *h = host + pos;
        if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) {
            sq_set_error("not a BGZF member at byte %llu (plain gzip streams cannot be inflated in parallel)",
                         (unsigned long long)pos);
            return SQ_E_FORMAT;
        }
        const uint32_t xlen = h[10] | (uint32_t)h[11] << 8;
        if (pos + 12 + xlen > nbytes) break;
        uint32_t bsize = 0;
        bool have = false;
        for (uint32_t x = 0; x + 4 <= xlen;) {
            const uint8_t *sf = h + 12 + x;
            const uint32_t slen = sf[2] | (uint32_t)sf[3] << 8;
            if (sf[0] == 'B' && sf[1] == 'C' && slen == 2 && x + 6 <= xlen) {
                bsize = sf[4] | (uint32_t)sf[5] << 8;
                have = true;
            }
            x += 4 + slen;
        }
        if (!have) {
            sq_set_error("gzip member at byte %llu has no BGZF 'BC' field", (unsigned long long)pos);
            return SQ_E_FORMAT;
        }
        const uint64_t total = (uint64_t)bsize + 1;
        if (total < 12 + (uint64_t)xlen + 8) {
            sq_set_error("BGZF member at byte %llu is smaller than its header", (unsigned long long)pos);
            return SQ_E_FORMAT;
        }
        if (pos + total > nbytes) break;  // incomplete member: the caller reads on
        const uint8_t *tr = h + total - 8;
        const uint32_t isize = tr[4] | (uint32_t)tr[5] << 8 | (uint32_t)tr[6] << 16 | (uint32_t)tr[7] << 24;
        if (isize > 65536) {
            sq_set_error("BGZF member at byte %llu claims %u bytes of text (limit 65536)", (unsigned long long)pos, isize);
            return SQ_E_FORMAT;
        }
        if (n == cap) break;
        blocks[n].comp_off = pos + 12 + xlen;
        blocks[n].comp_len = (uint32_t)(total - 12 - xlen - 8);
        blocks[n].text_len = isize;
        blocks[n].text_off = text;
        text += isize;
        n++;
        pos += total;
    }
    *n_blocks = n;
    *consumed = pos;
    *text_bytes = text;
    return SQ_OK;
}

// blocks[0 .. n): descriptors in HOST memory; their compressed bytes are at dev_comp + (comp_off - comp_base),
// the text goes to dev_out + (text_off - text_base).  Enqueued on the calling thread's stream (parser stream
// inside a parser entry point); *dev_first_bad (device, 8 bytes, preset to ~0) receives block << 8 | code of
// the first block that failed.
int bgzf_inflate_async(sq_ctx *ctx, const uint8_t *dev_comp, uint64_t comp_base, const sq_bgzf_block *dev_blocks, uint64_t n,
                       uint64_t text_base, uint8_t *dev_out, unsigned long long *dev_first_bad) {
    if (n == 0) return SQ_OK;
    uint64_t grid = (n + INF_WARPS - 1) / INF_WARPS;
    const uint64_t cap = (uint64_t)ctx->num_sms * 8;  // 64 warps per SM: one decoder per warp slot
    if (grid > cap) grid = cap;
    SQ_LAUNCH(ctx, k_bgzf_inflate, (unsigned)grid, INF_WARPS * 32, 0, dev_comp, comp_base, dev_blocks, (uint32_t)n, text_base,
              dev_out, dev_first_bad);
    return SQ_OK;
}

extern "C" int sq_bgzf_inflate(sq_ctx *ctx, const uint8_t *host_comp, uint64_t nbytes, const sq_bgzf_block *blocks, uint64_t n,
                               uint8_t *dev_out, uint64_t *bad_block, int *bad_code) {
    *bad_block = ~0ULL;
    *bad_code = 0;
    if (n == 0) return SQ_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = sq_cur_stream(ctx);
    uint8_t *d_comp = nullptr;
    sq_bgzf_block *d_blocks = nullptr;
    unsigned long long *d_bad = nullptr;
    int rc = sq_dalloc(ctx, (void **)&d_comp, nbytes + 64, false);
    if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&d_blocks, n * sizeof(sq_bgzf_block), false);
    if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&d_bad, 8, false);
    unsigned long long h_bad = ~0ULL;
    if (rc == SQ_OK) {
        CUDA_TRY(cudaMemcpyAsync(d_comp, host_comp, nbytes, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_blocks, blocks, n * sizeof(sq_bgzf_block), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemsetAsync(d_bad, 0xFF, 8, st));
        rc = bgzf_inflate_async(ctx, d_comp, 0, d_blocks, n, blocks[0].text_off, dev_out, d_bad);
        if (rc == SQ_OK) {
            CUDA_TRY(cudaMemcpyAsync(&h_bad, d_bad, 8, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
        }
    }
    sq_dfree(ctx, d_comp);
    sq_dfree(ctx, d_blocks);
    sq_dfree(ctx, d_bad);
    if (rc != SQ_OK) return rc;
    if (h_bad != ~0ULL) {
        *bad_block = h_bad >> 8;
        *bad_code = (int)(h_bad & 0xFF);
        sq_set_error("BGZF block %llu is corrupt (inflate error %d)", (unsigned long long)*bad_block, *bad_code);
        return SQ_E_FORMAT;
    }
    return SQ_OK;
}

// the same decoder on the host, for the unit tests (no GPU needed); not used by any product path
extern "C" int sq_selftest_inflate_host(const uint8_t *deflate, uint32_t len, uint8_t *out, uint32_t cap, uint32_t *out_len) {
    static thread_local InfTables t;
    auto copy_match = [&](uint32_t op, uint32_t dist, uint32_t n) {
        for (uint32_t i = 0; i < n; i++) out[op + i] = out[op - dist + (i % dist)];
    };
    return inf_inflate(deflate, len, out, cap, out_len, t, copy_match);
}
