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

// One warp = one CTA = one member at a time.  The newest INF_WIN bytes of the member's text live in a ring in
// shared memory: literals and matches are written there, matches read their source there (a match that
// reaches further back than the ring reads the bytes this warp already flushed to global memory), and
// the ring is flushed to global memory 1 KiB at a time with coalesced stores.  16 KiB of ring + 3 KiB of tables
// per warp: ten decoders per SM.  (Measured with a ring spanning DEFLATE's whole 32 KiB history, which spares the
// far matches their L2 round trip but leaves six decoders per SM: 6.6 GB/s of text against 8.9 GB/s.)
constexpr uint32_t INF_WIN = 16384, INF_FLUSH = 1024;

struct WarpOut {
    uint8_t *win;      // shared-memory ring
    uint8_t *dst;      // the member's text in global memory
    uint32_t flushed;  // bytes [0, flushed) are in global memory
    uint32_t lane;

    __device__ __forceinline__ void flush_to(uint32_t end) {  // bytes [flushed, end) -> global, all lanes
        __syncwarp();
        for (uint32_t i = flushed + lane; i < end; i += 32) dst[i] = win[i & (INF_WIN - 1)];
        flushed = end;
        __syncwarp();
    }
    __device__ __forceinline__ void room(uint32_t op, uint32_t n) {
        // keep the unflushed part short: the ring then always holds the last INF_WIN - INF_FLUSH - 258 bytes
        if (op + n - flushed > INF_FLUSH) flush_to(op);
    }
    __device__ __forceinline__ void put(uint32_t op, uint8_t c) {
        room(op, 1);
        if (lane == 0) win[op & (INF_WIN - 1)] = c;
    }
    __device__ __forceinline__ void raw(uint32_t op, const uint8_t *src, uint32_t n) {
        flush_to(op);
        for (uint32_t i = lane; i < n; i += 32) dst[op + i] = src[i];  // stored block: straight to global ...
        const uint32_t keep = n < INF_WIN ? n : INF_WIN;                // ... and its tail into the ring
        for (uint32_t i = n - keep + lane; i < n; i += 32) win[(op + i) & (INF_WIN - 1)] = src[i];
        flushed = op + n;
        __syncwarp();
    }
    __device__ __forceinline__ void match(uint32_t op, uint32_t dist, uint32_t n) {
        room(op, n);
        __syncwarp();
        // out[op + i] = out[op - dist + (i mod dist)]: the source lies in front of `op` and is complete, so
        // the lanes take the bytes of the match in any order
        if (dist <= INF_WIN - INF_FLUSH - 258) {
            for (uint32_t i = lane; i < n; i += 32) {
                const uint32_t k = dist >= n ? i : i % dist;
                win[(op + i) & (INF_WIN - 1)] = win[(op - dist + k) & (INF_WIN - 1)];
            }
        }
        else {
            // further back than the ring reaches: those bytes were flushed by this warp (dist >= n here)
            flush_to(op);
            for (uint32_t i = lane; i < n; i += 32) win[(op + i) & (INF_WIN - 1)] = dst[op - dist + i];
        }
        __syncwarp();
    }
};

__global__ void __launch_bounds__(32)
k_bgzf_inflate(const uint8_t *__restrict__ comp, uint64_t comp_base, const sq_bgzf_block *__restrict__ blocks, uint32_t n_blocks,
               uint64_t text_base, uint8_t *__restrict__ out, unsigned long long *first_bad) {
    __shared__ InfTables tables;
    __shared__ __align__(16) uint8_t win[INF_WIN];
    for (uint32_t bi = blockIdx.x; bi < n_blocks; bi += gridDim.x) {
        const sq_bgzf_block blk = blocks[bi];
        WarpOut o;
        o.win = win;
        o.dst = out + (blk.text_off - text_base);
        o.flushed = 0;
        o.lane = lane_id();
        uint32_t produced = 0;
        int rc = inf_inflate(comp + (blk.comp_off - comp_base), blk.comp_len, blk.text_len, &produced, tables, o);
        if (rc == INF_OK && produced != blk.text_len) rc = INF_E_SIZE;
        if (rc == INF_OK) o.flush_to(produced);
        else if (o.lane == 0) atomicMin(first_bad, (unsigned long long)bi << 8 | (unsigned)rc);
        __syncwarp();
    }
}

// (Measured and dropped: one THREAD per member -- 32 members per warp, tables in local memory, text straight to
// global memory.  Lanes of a warp take different branches at almost every symbol and serialise; with the few
// thousand members a 256 MiB window holds it ran at 0.8 GB/s against 8.5 GB/s for the warp-per-member kernel.)

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
        if (blocks) {  // (blocks == NULL: count only)
            if (n == cap) break;
            blocks[n].comp_off = pos + 12 + xlen;
            blocks[n].comp_len = (uint32_t)(total - 12 - xlen - 8);
            blocks[n].text_len = isize;
            blocks[n].text_off = text;
        }
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
    uint64_t grid = n;
    const uint64_t cap = (uint64_t)ctx->num_sms * 10;  // ten one-warp CTAs fit an SM (shared memory)
    if (grid > cap) grid = cap;
    SQ_LAUNCH(ctx, k_bgzf_inflate, (unsigned)grid, 32, 0, dev_comp, comp_base, dev_blocks, (uint32_t)n, text_base, dev_out,
              dev_first_bad);
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
    struct HostOut {
        uint8_t *out;
        // (__host__ __device__ only because inf_inflate is: this instantiation runs on the host)
        __host__ __device__ void put(uint32_t op, uint8_t c) { out[op] = c; }
        __host__ __device__ void raw(uint32_t op, const uint8_t *src, uint32_t n) {
            for (uint32_t i = 0; i < n; i++) out[op + i] = src[i];
        }
        __host__ __device__ void match(uint32_t op, uint32_t dist, uint32_t n) {
            for (uint32_t i = 0; i < n; i++) out[op + i] = out[op - dist + (i % dist)];
        }
    } o{out};
    return inf_inflate(deflate, len, cap, out_len, t, o);
}
