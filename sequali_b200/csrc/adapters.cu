// adapters.cu -- AdapterCounter (reference _qcmodule.c:2465-2823).
//
// The reference packs the adapters into 64-bit shift-AND automata and walks
// every read base by base.  What its tables record is, per adapter, the FIRST
// position where the adapter occurs in the read (letters compared by class:
// A/C/G/T case-insensitive, everything else one class).  Here the bit
// parallelism runs along the read instead of along the pattern: a thread turns
// up to 256 read positions into four bit planes (valid / bit1 / bit2 /
// present), and an adapter matches where the AND of its letters' planes,
// shifted by the letter's offset, leaves a bit.  Mismatches kill all bits
// after a few letters, so most adapters cost ~4 plane operations per chunk.
#include "modules.cuh"



// bit i of the result = bit 0 of byte i of x (x has only bit 0 of each byte set)
__device__ __forceinline__ uint32_t pack_bytes_lsb(uint32_t x) { return (x * 0x00204081u) >> 21 & 0xFu; }

__device__ __forceinline__ uint32_t ad_load_word(const uint8_t *p) {
    uintptr_t a = (uintptr_t)p;
    const uint32_t *w = (const uint32_t *)(a & ~(uintptr_t)3);
    return __funnelshift_r(__ldg(w), __ldg(w + 1), (uint32_t)(a & 3) * 8);
}

// 0x01 per byte that is ACGTacgt (same construction as qc.cu)
__device__ __forceinline__ uint32_t ad_acgt_bytes(uint32_t w) {
    uint32_t sel = w & 0x07070707u;
    uint32_t t = sel | (sel >> 4);
    uint32_t nib = __byte_perm(t, 0, 0x4420);
    uint32_t expect = __byte_perm(0x40FF40FFu, 0x40FFFF50u, nib);
    return zero_bytes80((w & 0xD8D8D8D8u) ^ expect) >> 7;
}

__device__ __forceinline__ uint64_t shr_pair(uint64_t lo, uint64_t hi, uint32_t s) {
    return s == 0 ? lo : (lo >> s) | (hi << (64 - s));
}

// One chunk of a read: positions [c0, c0 + 256) as four 64-bit words per plane; every adapter of
// [a0, a1) whose bit in `skip` is clear is matched, emit(a, p) gets the first match position inside the
// chunk (absolute, p + adapter length <= L guaranteed by the planes being empty past the read).
template <typename Emit>
__device__ __forceinline__ void ad_scan_chunk(const uint8_t *seq, uint32_t L, uint32_t c0, const uint8_t *__restrict__ pat,
                                              const uint32_t *__restrict__ plen, uint32_t a0, uint32_t a1, uint64_t skip,
                                              Emit emit) {
    uint64_t V[5], H[5], G[5], P[5];
    V[4] = H[4] = G[4] = P[4] = 0;  // zero word past the end for the shifts
    const uint32_t span = min(256u, L - c0);
#pragma unroll
    for (int wi = 0; wi < 4; wi++) {
        uint64_t v = 0, h = 0, g = 0, p = 0;
        if ((uint32_t)wi * 64 < span) {
#pragma unroll 4
            for (int k = 0; k < 16; k++) {
                uint32_t off = wi * 64 + k * 4;
                if (off >= span) break;
                uint32_t w = ad_load_word(seq + c0 + off);
                uint32_t nvalid = min(4u, span - off);
                uint32_t pm = 0x01010101u >> (8 * (4 - nvalid));
                uint32_t vb = ad_acgt_bytes(w) & pm;
                uint32_t hb = (w >> 1) & vb, gb = (w >> 2) & vb;
                v |= (uint64_t)pack_bytes_lsb(vb) << (k * 4);
                h |= (uint64_t)pack_bytes_lsb(hb) << (k * 4);
                g |= (uint64_t)pack_bytes_lsb(gb) << (k * 4);
                p |= (uint64_t)pack_bytes_lsb(pm) << (k * 4);
            }
        }
        V[wi] = v; H[wi] = h; G[wi] = g; P[wi] = p;
    }
    const int nw = (span + 63) / 64;
    for (uint32_t a = a0; a < a1; a++) {
        if (skip >> (a - a0) & 1) continue;
        const uint32_t m = plen[a];
        if (m == 0 || m > L - c0) continue;
        uint64_t M[4] = {~0ULL, ~0ULL, ~0ULL, ~0ULL};
        const uint8_t *pa = pat + (size_t)a * AD_MAXLEN;
        bool alive = true;
        for (uint32_t j = 0; j < m && alive; j++) {
            const uint32_t c = pa[j];
            // letter class -> plane polarity: class<4 needs V=1,H=h,G=g; class 4 needs V=0 (and present)
            const uint64_t xv = c == 4 ? ~0ULL : 0ULL;
            const uint64_t xh = (c == 4 || !(c == 1 || c == 2)) ? ~0ULL : 0ULL;
            const uint64_t xg = (c == 4 || !(c == 2 || c == 3)) ? ~0ULL : 0ULL;
            uint64_t any = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (i < nw) {
                    // letter plane of word i and i+1, shifted right by j (< 64)
                    uint64_t x0 = (V[i] ^ xv) & (H[i] ^ xh) & (G[i] ^ xg) & P[i];
                    uint64_t x1 = (V[i + 1] ^ xv) & (H[i + 1] ^ xh) & (G[i + 1] ^ xg) & P[i + 1];
                    M[i] &= shr_pair(x0, x1, j);
                    any |= M[i];
                }
            }
            alive = any != 0;
        }
        if (!alive) continue;
        // first set bit = first match inside this chunk
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (i < nw && M[i]) {
                emit(a, c0 + i * 64 + (uint32_t)(__ffsll((long long)M[i]) - 1));
                break;
            }
        }
    }
}

// short reads: a thread per read walks its (few) chunks in order and stops at the first match
__global__ void __launch_bounds__(AD_TPB)
k_adapters(BatchView bv, const uint8_t *__restrict__ pat, const uint32_t *__restrict__ plen,
           uint32_t n_adapters, uint32_t max_pat_len, uint64_t *counts, uint64_t cap_len) {
    const uint32_t step = 256 - (max_pat_len ? max_pat_len - 1 : 0);
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < bv.n; r += gridDim.x * blockDim.x) {
        const uint32_t L = bv.seq_len[r];
        const uint8_t *seq = bv.text + bv.seq_off[r];
        for (uint32_t a0 = 0; a0 < n_adapters; a0 += 64) {
            uint64_t found = 0;  // bit a - a0
            const uint32_t a1 = min(n_adapters, a0 + 64);
            const uint64_t all = a1 - a0 == 64 ? ~0ULL : (1ULL << (a1 - a0)) - 1;
            for (uint32_t c0 = 0; c0 < L && found != all; c0 += step) {
                ad_scan_chunk(seq, L, c0, pat, plen, a0, a1, found, [&](uint32_t a, uint32_t p) {
                    uint64_t *fwd = counts + (size_t)a * 2 * cap_len;
                    atomic_add_u64(fwd + p, 1);
                    atomic_add_u64(fwd + cap_len + (L - 1 - p), 1);
                    found |= 1ULL << (a - a0);
                });
            }
        }
    }
}

// long reads (nanopore: mean 20 kb, up to 1 Mb): a warp per read, a lane per chunk -- the chunks of
// one read are walked 32 at a time, the first occurrence per adapter (:2644-2672: only the first match
// of an adapter in a read counts) is the minimum over the chunks, kept per (read, adapter) in `first`;
// k_adapters_finish turns it into the forward / reverse counts.  A chunk whose 32-chunk group starts
// behind the best match so far of every adapter is skipped.
constexpr uint32_t AD_NONE = 0xFFFFFFFFu;
__global__ void __launch_bounds__(AD_TPB)
k_adapters_chunks(BatchView bv, const uint8_t *__restrict__ pat, const uint32_t *__restrict__ plen, uint32_t n_adapters,
                  uint32_t max_pat_len, uint32_t *__restrict__ first /* [n][n_adapters] */) {
    const uint32_t step = 256 - (max_pat_len ? max_pat_len - 1 : 0);
    const uint32_t warps = gridDim.x * (AD_TPB / 32), lane = lane_id();
    for (uint32_t r = blockIdx.x * (AD_TPB / 32) + (threadIdx.x >> 5); r < bv.n; r += warps) {
        const uint32_t L = bv.seq_len[r];
        const uint8_t *seq = bv.text + bv.seq_off[r];
        uint32_t *fr = first + (size_t)r * n_adapters;
        const uint32_t n_chunks = L ? (L + step - 1) / step : 0;
        for (uint32_t a0 = 0; a0 < n_adapters; a0 += 64) {
            const uint32_t a1 = min(n_adapters, a0 + 64);
            uint64_t found = 0;  // adapters matched in an earlier group of chunks (warp uniform)
            const uint64_t all = a1 - a0 == 64 ? ~0ULL : (1ULL << (a1 - a0)) - 1;
            for (uint32_t cb = 0; cb < n_chunks && found != all; cb += 32) {
                const uint32_t c = cb + lane;
                uint64_t mine = 0;
                if (c < n_chunks)
                    ad_scan_chunk(seq, L, c * step, pat, plen, a0, a1, found, [&](uint32_t a, uint32_t p) {
                        atomicMin(fr + a, p);
                        mine |= 1ULL << (a - a0);
                    });
                // matches of this group settle those adapters: every later chunk starts behind them
                for (int o = 16; o > 0; o >>= 1) mine |= __shfl_xor_sync(0xffffffffu, mine, o);
                found |= mine;
            }
        }
    }
}
__global__ void __launch_bounds__(256)
k_adapters_finish(BatchView bv, const uint32_t *__restrict__ first, uint32_t n_adapters, uint64_t *counts, uint64_t cap_len) {
    const uint64_t total = (uint64_t)bv.n * n_adapters;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t p = first[i];
        if (p == AD_NONE) continue;
        const uint32_t r = (uint32_t)(i / n_adapters), a = (uint32_t)(i % n_adapters);
        uint64_t *fwd = counts + (size_t)a * 2 * cap_len;
        atomic_add_u64(fwd + p, 1);
        atomic_add_u64(fwd + cap_len + (bv.seq_len[r] - 1 - p), 1);
    }
}

extern "C" int sq_adapters_create(sq_ctx *ctx, const char *const *adapters, uint64_t n, sq_adapters **out) {
    *out = nullptr;
    if (n < 1) {
        sq_set_error("At least one adapter is expected");
        return SQ_E_ARG;
    }
    CUDA_TRY(cudaSetDevice(ctx->device));
    std::vector<uint8_t> pat(n * AD_MAXLEN, 0);
    std::vector<uint32_t> plen(n, 0);
    uint32_t max_pat = 0;
    for (uint64_t i = 0; i < n; i++) {
        size_t len = strlen(adapters[i]);
        if (len > AD_MAXLEN) {
            sq_set_error("Maximum adapter size is %d, got %zu", AD_MAXLEN, len);
            return SQ_E_ARG;
        }
        plen[i] = (uint32_t)len;
        if (len > max_pat) max_pat = (uint32_t)len;
        for (size_t j = 0; j < len; j++) {
            uint8_t c = (uint8_t)adapters[i][j] | 0x20;
            pat[i * AD_MAXLEN + j] = c == 'a' ? 0 : c == 'c' ? 1 : c == 'g' ? 2 : c == 't' ? 3 : 4;
        }
    }
    sq_adapters *a = new sq_adapters();
    a->ctx = ctx;
    a->n_adapters = (uint32_t)n;
    a->max_pat_len = max_pat;
    int rc = sq_dalloc(ctx, (void **)&a->pat, pat.size(), false);
    if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&a->plen, plen.size() * 4, false);
    if (rc != SQ_OK) {
        sq_adapters_destroy(a);
        return rc;
    }
    CUDA_TRY(cudaMemcpyAsync(a->pat, pat.data(), pat.size(), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(a->plen, plen.data(), plen.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    *out = a;
    return SQ_OK;
}

extern "C" void sq_adapters_destroy(sq_adapters *a) {
    if (!a) return;
    cudaSetDevice(a->ctx->device);
    sq_dfree(a->ctx, a->pat);
    sq_dfree(a->ctx, a->plen);
    sq_dfree(a->ctx, a->counts);
    delete a;
}

int adapters_grow(sq_adapters *a, uint64_t len) {
    if (len <= a->cap_len) return SQ_OK;
    uint64_t cap = a->cap_len * 2 > len ? a->cap_len * 2 : len;
    if (cap < 256) cap = 256;
    uint64_t *nc = nullptr;
    SQ_TRY(sq_dalloc(a->ctx, (void **)&nc, (size_t)a->n_adapters * 2 * cap * 8, true));
    if (a->cap_len)
        CUDA_TRY(cudaMemcpy2DAsync(nc, cap * 8, a->counts, a->cap_len * 8, a->cap_len * 8,
                                   (size_t)a->n_adapters * 2, cudaMemcpyDeviceToDevice, a->ctx->stream));
    sq_dfree(a->ctx, a->counts);
    a->counts = nc;
    a->cap_len = cap;
    return SQ_OK;
}

extern "C" int sq_adapters_add(sq_adapters *a, sq_batch *b) {
    sq_ctx *ctx = a->ctx;
    if (b->ctx != ctx) {
        sq_set_error("record array belongs to another context");
        return SQ_E_ARG;
    }
    if (b->n == 0) return SQ_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    SQ_TRY(adapters_grow(a, b->max_len));
    if (b->max_len > 2048) {
        // long reads: (read, chunk) parallel, first occurrence by atomicMin, then the counts
        uint32_t *first = nullptr;
        const size_t cells = (size_t)b->n * a->n_adapters;
        SQ_TRY(sq_dalloc(ctx, (void **)&first, cells * 4, false));
        CUDA_TRY(cudaMemsetAsync(first, 0xFF, cells * 4, ctx->stream));
        SQ_LAUNCH(ctx, k_adapters_chunks, sq_grid_for(ctx, b->n * 32, AD_TPB, 32), AD_TPB, 0, b->view(), a->pat, a->plen,
                  a->n_adapters, a->max_pat_len, first);
        SQ_LAUNCH(ctx, k_adapters_finish, sq_grid_for(ctx, cells, 256, 8), 256, 0, b->view(), first, a->n_adapters,
                  a->counts, a->cap_len);
        sq_dfree(ctx, first);
    }
    else if (b->max_len > 0) {
        int grid = sq_grid_for(ctx, b->n, AD_TPB, 16);
        SQ_LAUNCH(ctx, k_adapters, grid, AD_TPB, 0, b->view(), a->pat, a->plen, a->n_adapters,
                  a->max_pat_len, a->counts, a->cap_len);
    }
    a->n_seqs += b->n;
    if (b->max_len > a->max_len) a->max_len = b->max_len;
    return SQ_OK;
}

extern "C" int sq_adapters_sync(sq_adapters *a, uint64_t *number_of_sequences, uint64_t *max_length) {
    CUDA_TRY(cudaSetDevice(a->ctx->device));
    CUDA_TRY(cudaStreamSynchronize(a->ctx->stream));
    *number_of_sequences = a->n_seqs;
    *max_length = a->max_len;
    return SQ_OK;
}

extern "C" int sq_adapters_read(sq_adapters *a, uint64_t index, uint64_t *forward, uint64_t *reverse) {
    if (index >= a->n_adapters) {
        sq_set_error("adapter index out of range");
        return SQ_E_ARG;
    }
    CUDA_TRY(cudaSetDevice(a->ctx->device));
    if (a->max_len) {
        const uint64_t *src = a->counts + index * 2 * a->cap_len;
        CUDA_TRY(cudaMemcpyAsync(forward, src, a->max_len * 8, cudaMemcpyDeviceToHost, a->ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(reverse, src + a->cap_len, a->max_len * 8, cudaMemcpyDeviceToHost, a->ctx->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(a->ctx->stream));
    return SQ_OK;
}
