// common.cuh -- shared host/device definitions of libsqgpu (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/sqgpu.h"

#define SQ_WARP 32
constexpr int SQ_NUM_SMS_HINT = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
void sq_set_error(const char *fmt, ...);
int sq_cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define CUDA_TRY(expr)                                                        \
    do {                                                                      \
        cudaError_t _e = (expr);                                              \
        if (_e != cudaSuccess) return sq_cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

#define SQ_TRY(expr)              \
    do {                          \
        int _rc = (expr);         \
        if (_rc != SQ_OK) return _rc; \
    } while (0)

// ---------------------------------------------------------------------------
// context and record arrays
// ---------------------------------------------------------------------------
struct sq_ctx {
    int device = 0;
    int num_sms = SQ_NUM_SMS_HINT;
    cudaStream_t stream = nullptr;
    // The record-boundary scan of the NEXT record array may run on a helper thread while this thread
    // feeds the collectors with the current one: parser entry points put their work on `pstream`
    // (SqParserScope below), so their host synchronisations do not wait for the collectors' kernels.
    // One collector thread + one parser thread per context is the supported concurrency.
    cudaStream_t pstream = nullptr;
    std::recursive_mutex parse_mutex;  // one parser entry point at a time (SqParserScope)
    // Table stream: the hash-table modules of a record array (OverrepresentedSequences, NanoStats, DedupEstimator:
    // latency and atomics bound kernels with host waits in between) run here from the end of k_fused_reads on, beside
    // the per-position pass and PerTileQuality's chain kernel on the launch stream; sq_fused_add forks and joins, so
    // everything else stays ordered on the launch stream.
    cudaStream_t tstream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // cache of large scratch blocks (sq_dalloc / sq_dfree)
    struct BigBlock {
        void *p = nullptr;
        size_t cls = 0;
        cudaEvent_t ev = nullptr;        // recorded where the block was freed
        cudaStream_t freed_on = nullptr;
    };
    std::mutex big_mutex;
    std::vector<BigBlock> big_free;      // oldest first
    size_t big_free_bytes = 0, big_free_cap = (size_t)16 << 30;
    std::atomic<uint64_t> big_from_driver{0}, big_from_cache{0};  // requests of >= 8 MiB served by cudaMallocAsync / by the cache
    std::unordered_map<void *, BigBlock> big_live;
    std::atomic<uint64_t> launches{0};
    std::mutex prof_mutex;
    // pinned scratch for small device->host result structs
    void *h_scratch = nullptr;  // 4 KiB pinned
    void *d_scratch = nullptr;  // 4 KiB device
    double *d_err_table = nullptr;      // [94]  10^-(q/10), host libm generated
    double *d_phred_thresholds = nullptr;  // [94] bucket edges derived from host log10
    void *parse_masks = nullptr;        // newline bit masks of the record array being parsed (grow-only)
    size_t parse_masks_cap = 0;
    void *parse_fields = nullptr;       // one-pass parser: descriptor fields before the record count is known
    size_t parse_fields_cap = 0;
    void *parse_status = nullptr;       // one-pass parser: look-back status words
    size_t parse_status_cap = 0;
    void *h_bounce = nullptr;           // pinned bounce buffer for large read-outs into pageable caller memory
    size_t h_bounce_cap = 0;
    // staging ring of the host reader (sq_fastq_stream), kept between readers
    void *stage_slot[3] = {nullptr, nullptr, nullptr};
    size_t stage_cap = 0;
    bool stage_in_use = false;
    uint32_t func_attr_done = 0;           // bit per kernel family whose smem opt-in was set
    // optional per-kernel timing (CUDA events on the launch stream), see sq_ctx_profile
    bool profile = false;
    struct ProfEvent { const char *name; cudaEvent_t start, stop; };
    std::vector<ProfEvent> prof_events;
    cudaEvent_t timer_start = nullptr, timer_stop = nullptr;
};
void sq_prof_begin(sq_ctx *ctx, const char *name);
void sq_prof_end(sq_ctx *ctx);

// stream of the calling thread's work: the context's launch stream, or the parser stream inside a
// parser entry point
extern thread_local cudaStream_t sq_tls_stream;
inline cudaStream_t sq_cur_stream(const sq_ctx *ctx) { return sq_tls_stream ? sq_tls_stream : ctx->stream; }
// work of the calling thread goes to stream `s` inside the scope (nullptr: no change)
struct SqStreamScope {
    cudaStream_t prev;
    explicit SqStreamScope(cudaStream_t s) : prev(sq_tls_stream) {
        if (s) sq_tls_stream = s;
    }
    ~SqStreamScope() { sq_tls_stream = prev; }
};
// A parser entry point: its work goes to the parser stream, and it has the context's parser scratch (field
// buffers, look-back words, result slots) to itself -- two parsers of one process may be driven from two threads
// (the paired loop: one parser reads ahead on a helper thread while the caller asks the other for read(n)); their
// scans take turns.
struct SqParserScope {
    cudaStream_t prev;
    std::unique_lock<std::recursive_mutex> turn;
    // (while per-kernel profiling is on everything stays on the launch stream: its events are ordered)
    explicit SqParserScope(sq_ctx *ctx) : prev(sq_tls_stream), turn(ctx->parse_mutex) {
        sq_tls_stream = ctx->profile ? nullptr : ctx->pstream;
    }
    ~SqParserScope() { sq_tls_stream = prev; }
};

// Device-side view of a record array.  Offsets index `text`.
struct BatchView {
    const uint8_t *text;
    const uint32_t *name_off, *seq_off, *seq_len, *qual_off;
    const uint32_t *name_len, *tags_off, *tags_len;  // nullptr for FASTQ text batches
    double *err_sum;
    uint32_t n;
};

struct sq_batch {
    sq_ctx *ctx = nullptr;
    uint8_t *text = nullptr;
    bool owns_text = true;
    uint64_t nbytes = 0;
    uint64_t n = 0;
    uint32_t max_len = 0;
    uint32_t max_rec_bytes = 0;  // FASTQ text arrays: longest record in bytes (0 = unknown)
    uint64_t text_end = 0;       // FASTQ text arrays: byte after the last complete record
    uint32_t *name_off = nullptr, *seq_off = nullptr, *seq_len = nullptr, *qual_off = nullptr;
    uint32_t *name_len = nullptr, *tags_off = nullptr, *tags_len = nullptr;
    double *err_sum = nullptr;
    void *meta_block = nullptr;  // single allocation behind the arrays above
    uint64_t meta_stride = 0;    // elements per array inside meta_block (0: (n + 3) & ~3)
    bool err_sum_valid = false;  // QCMetrics ran on this array
    // tile ids PerTileQuality met in this array (sq_batch_select_tiles skips arrays that cannot hold a tile)
    bool tile_range_valid = false;
    uint64_t tile_lo = 0, tile_hi = 0;

    BatchView view() const {
        BatchView v;
        v.text = text;
        v.name_off = name_off; v.seq_off = seq_off; v.seq_len = seq_len; v.qual_off = qual_off;
        v.name_len = name_len; v.tags_off = tags_off; v.tags_len = tags_len;
        v.err_sum = err_sum;
        v.n = (uint32_t)n;
        return v;
    }
};

// device -> pageable host memory through the context's pinned bounce buffer (DMA at full speed,
// then a plain memcpy), synchronous
int sq_d2h_bounced(sq_ctx *ctx, void *dst, const void *dev_src, size_t nbytes);

// name bytes of record r of a record array (host copy; rare paths only)
int sq_batch_get_name(sq_batch *b, uint64_t r, std::vector<uint8_t> &out);

// stream-ordered allocation helpers (cudaMallocAsync on the calling thread's stream; blocks of 8 MiB and more are
// cached by the context, see core.cu)
int sq_dalloc(sq_ctx *ctx, void **p, size_t nbytes, bool zero);
void sq_dfree(sq_ctx *ctx, void *p);

// Stream-ordered scratch that goes back to the pool on EVERY way out of a function (early error returns
// included); keep(p) hands a block over to the caller / a longer-lived owner.
struct SqScratch {
    sq_ctx *ctx;
    std::vector<void *> owned;
    explicit SqScratch(sq_ctx *c) : ctx(c) {}
    SqScratch(const SqScratch &) = delete;
    SqScratch &operator=(const SqScratch &) = delete;
    ~SqScratch() {
        for (void *p : owned) sq_dfree(ctx, p);
    }
    template <class T> int get(T **p, size_t nbytes, bool zero = false) {
        const int rc = sq_dalloc(ctx, (void **)p, nbytes, zero);
        if (rc == SQ_OK && *p) owned.push_back((void *)*p);
        return rc;
    }
    void keep(const void *p) {
        for (size_t i = 0; i < owned.size(); i++)
            if (owned[i] == p) {
                owned.erase(owned.begin() + i);
                return;
            }
    }
};

// device-wide exclusive scan (scan.cu)
int sq_scan_exclusive_u32(sq_ctx *ctx, const uint32_t *in, uint32_t *out, uint32_t n, uint32_t *total_dev);

// stable radix sort of (key, value) pairs by the low key_bits of the key (sort.cu)
int sq_radix_sort_pairs(sq_ctx *ctx, uint32_t *keys, uint32_t *vals, uint32_t *tmp_keys, uint32_t *tmp_vals,
                        uint32_t n, uint32_t key_bits);

inline int sq_grid_for(sq_ctx *ctx, uint64_t work_items, int per_block, int max_waves = 8) {
    uint64_t blocks = (work_items + per_block - 1) / per_block;
    uint64_t cap = (uint64_t)ctx->num_sms * max_waves;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

#define SQ_LAUNCH(ctx, kernel, grid, block, smem, ...)                         \
    do {                                                                       \
        const bool _prof = (ctx)->profile;                                     \
        if (_prof) {                                                           \
            (ctx)->prof_mutex.lock();                                          \
            sq_prof_begin((ctx), #kernel);                                     \
        }                                                                      \
        kernel<<<(grid), (block), (smem), sq_cur_stream(ctx)>>>(__VA_ARGS__);  \
        if (_prof) {                                                           \
            sq_prof_end((ctx));                                                \
            (ctx)->prof_mutex.unlock();                                        \
        }                                                                      \
        (ctx)->launches++;                                                     \
        CUDA_TRY(cudaGetLastError());                                          \
    } while (0)

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

// A/a C/c G/g T/t -> 0..3, everything else 4 (reference LUT _qcmodule.c:1748-1763)
__device__ __forceinline__ uint32_t nuc5(uint32_t c) {
    uint32_t u = c | 0x20u;
    return u == 'a' ? 0u : u == 'c' ? 1u : u == 'g' ? 2u : u == 't' ? 3u : 4u;
}

// 0x80 in every byte of x that is zero (exact, no cross-byte borrow)
__device__ __forceinline__ uint32_t zero_bytes80(uint32_t x) {
    return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;
}

// 32-bit little-endian load from an arbitrary byte address (two aligned loads
// + funnel shift).  Reads up to 7 bytes past `p`; buffers are padded for that.
__device__ __forceinline__ uint32_t load_u32_unaligned(const uint8_t *p) {
    uintptr_t a = (uintptr_t)p;
    const uint32_t *w = (const uint32_t *)(a & ~(uintptr_t)3);
    uint32_t sh = (uint32_t)(a & 3) * 8;
    uint32_t lo = __ldg(w);
    if (sh == 0) return lo;
    uint32_t hi = __ldg(w + 1);
    return __funnelshift_r(lo, hi, sh);
}

__device__ __forceinline__ uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

__device__ __forceinline__ uint64_t fmix64(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

// MurmurHash3 x64 128, second 64-bit half (reference murmur3.h:49-156), bytes
// fetched through `get(i)`.
template <typename GetByte>
__device__ __forceinline__ uint64_t murmur3_h2(GetByte get, uint64_t len, uint64_t seed) {
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    uint64_t h1 = seed, h2 = seed;
    uint64_t nb = len >> 4;
    for (uint64_t b = 0; b < nb; b++) {
        uint64_t k1 = 0, k2 = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            k1 |= (uint64_t)get(b * 16 + i) << (8 * i);
            k2 |= (uint64_t)get(b * 16 + 8 + i) << (8 * i);
        }
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    uint64_t rem = len & 15, base = nb << 4, k1 = 0, k2 = 0;
    for (uint64_t i = 0; i < rem; i++) {
        uint64_t v = get(base + i);
        if (i < 8) k1 |= v << (8 * i);
        else k2 |= v << (8 * (i - 8));
    }
    if (rem > 8) { k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2; }
    if (rem > 0) { k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1; }
    h1 ^= len; h2 ^= len;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2; h2 += h1;
    return h2;
}

// Thomas Wang's 64-bit mix (reference wanghash.h:14-25)
__device__ __forceinline__ uint64_t wang64(uint64_t k) {
    k = ~k + (k << 21);
    k ^= k >> 24;
    k *= 265;
    k ^= k >> 14;
    k *= 21;
    k ^= k >> 28;
    k += k << 31;
    return k;
}

__host__ __device__ inline uint64_t unxorshift64(uint64_t v, int s) {
    uint64_t x = v;
    for (int i = s; i < 64; i += s) x = v ^ (x >> s);
    return x;
}
// inverse of wang64 (each step undone by its modular inverse / xorshift inverse)
__host__ __device__ inline uint64_t wang64_inverse(uint64_t k) {
    k *= 0x3fffffff80000001ULL;  // (1 + 2^31)^-1
    k = unxorshift64(k, 28);
    k *= 14933078535860113213ULL;  // 21^-1
    k = unxorshift64(k, 14);
    k *= 15244667743933553977ULL;  // 265^-1
    k = unxorshift64(k, 24);
    return (k + 1) * 0x7ffffbffffdfffffULL;  // (2^21 - 1)^-1
}

__device__ __forceinline__ uint32_t warp_sum_u32(uint32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ uint32_t warp_max_u32(uint32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ uint32_t warp_excl_scan_u32(uint32_t v, uint32_t *total) {
    uint32_t x = v;
    uint32_t l = lane_id();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (l >= (uint32_t)o) x += y;
    }
    *total = __shfl_sync(0xffffffffu, x, 31);
    return x - v;
}

__device__ __forceinline__ void atomic_add_u64(uint64_t *p, uint64_t v) {
    atomicAdd((unsigned long long *)p, (unsigned long long)v);
}
__device__ __forceinline__ void atomic_min_u64(uint64_t *p, uint64_t v) {
    atomicMin((unsigned long long *)p, (unsigned long long)v);
}
__device__ __forceinline__ void atomic_max_u64(uint64_t *p, uint64_t v) {
    atomicMax((unsigned long long *)p, (unsigned long long)v);
}

#endif  // __CUDACC__
