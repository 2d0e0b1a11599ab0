// core.cu -- context, memory, record arrays and the FASTQ record-boundary scan.
//
// Replaces FastqParser_create_record_array's memchr loop
// (reference _qcmodule.c:1093-1171) and string_is_ascii (:204-237) with three
// kernels over text staged in HBM:
//   k_count_newlines   16-byte vector loads, SWAR newline count per CTA, first
//                      non-ASCII byte by atomicMin
//   sq_scan_exclusive  exclusive scan of the per-CTA counts (scan.cu)
//   k_scatter_fields   same traversal, warp-shuffle scans give every newline
//                      its rank k; the newline writes the descriptor field it
//                      closes (record k/4, line k%4) and checks '@' / '+'
//   k_finish_records   one thread per record: sequence length, equal-length
//                      check, longest sequence / record of the array
#include <ctype.h>
#include <math.h>
#include <stdarg.h>

#include <chrono>

#include "common.cuh"

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local char g_err[512] = "";
thread_local cudaStream_t sq_tls_stream = nullptr;

void sq_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int sq_cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    sq_set_error("CUDA error %s (%s) at %s:%d in %s", cudaGetErrorName(e), cudaGetErrorString(e),
                 file, line, what);
    return e == cudaErrorMemoryAllocation ? SQ_E_NOMEM : SQ_E_CUDA;
}
extern "C" const char *sq_last_error(void) { return g_err; }

extern "C" int sq_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// ---------------------------------------------------------------------------
// allocation
// ---------------------------------------------------------------------------
// Large scratch blocks (descriptor blocks, hashes, tile ids, segment histograms: 16-170 MB each, several per
// record array) are kept by the context instead of going back to the stream-ordered pool.  The pool serves a large
// request by re-mapping free physical granules into one address range when none of its fragments fits, and on a
// virtualised box that step takes 10-800 ms of host time inside cudaMallocAsync, at random, for dozens of passes
// (`SQ_TRACE_ALLOC=1`, tools/step_times.py: single passes of 160-1350 ms among 88 ms ones, with the pool's
// reserved size unchanged).  Here a freed block of >= 8 MiB keeps its size class and the event of its last use;
// the next request of that class takes it and, on another stream, waits for that event on the device.  After
// the first record arrays no allocation reaches the driver any more.
constexpr size_t SQ_BIG_MIN = (size_t)8 << 20;
static size_t sq_big_class(size_t nbytes) {
    const size_t step = nbytes < ((size_t)256 << 20) ? (size_t)16 << 20 : (size_t)64 << 20;
    return (nbytes + step - 1) / step * step;
}

static int sq_dalloc_traced(sq_ctx *ctx, void **p, size_t nbytes) {
    static const bool trace = getenv("SQ_TRACE_ALLOC") != nullptr;
    if (!trace) {
        CUDA_TRY(cudaMallocAsync(p, nbytes, sq_cur_stream(ctx)));
        return SQ_OK;
    }
    const auto t0 = std::chrono::steady_clock::now();
    cudaError_t e = cudaMallocAsync(p, nbytes, sq_cur_stream(ctx));
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (ms > 2.0) {
        cudaMemPool_t pool;
        uint64_t reserved = 0, used = 0;
        cudaDeviceGetDefaultMemPool(&pool, ctx->device);
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved);
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used);
        fprintf(stderr, "[sq] cudaMallocAsync(%zu bytes) took %.1f ms on the %s stream; pool: reserved %.2f GB, in use %.2f GB\n",
                nbytes, ms, sq_cur_stream(ctx) == ctx->stream ? "launch" : sq_cur_stream(ctx) == ctx->pstream ? "parser" : "table",
                reserved / 1e9, used / 1e9);
    }
    if (e != cudaSuccess) return sq_cuda_fail(e, "cudaMallocAsync", __FILE__, __LINE__);
    return SQ_OK;
}

int sq_dalloc(sq_ctx *ctx, void **p, size_t nbytes, bool zero) {
    *p = nullptr;
    if (nbytes == 0) nbytes = 16;
    static const bool no_cache = getenv("SQ_NO_BLOCK_CACHE") != nullptr;
    cudaStream_t cur = sq_cur_stream(ctx);
    if (nbytes >= SQ_BIG_MIN && !no_cache) {
        const size_t cls = sq_big_class(nbytes);
        sq_ctx::BigBlock blk;
        bool found = false;
        {
            std::lock_guard<std::mutex> lk(ctx->big_mutex);
            int pick = -1;
            for (int i = (int)ctx->big_free.size() - 1; i >= 0; i--)
                if (ctx->big_free[i].cls == cls) {
                    if (pick < 0) pick = i;
                    if (ctx->big_free[i].freed_on == cur) {  // last used on this very stream: nothing to wait for
                        pick = i;
                        break;
                    }
                }
            if (pick >= 0) {
                blk = ctx->big_free[pick];
                ctx->big_free.erase(ctx->big_free.begin() + pick);
                ctx->big_free_bytes -= blk.cls;
                found = true;
            }
        }
        if (found) {
            if (blk.freed_on != cur) CUDA_TRY(cudaStreamWaitEvent(cur, blk.ev, 0));
            ctx->big_from_cache++;
        }
        else {
            SQ_TRY(sq_dalloc_traced(ctx, &blk.p, cls));
            ctx->big_from_driver++;
            blk.cls = cls;
            blk.freed_on = nullptr;
            CUDA_TRY(cudaEventCreateWithFlags(&blk.ev, cudaEventDisableTiming));
        }
        {
            std::lock_guard<std::mutex> lk(ctx->big_mutex);
            ctx->big_live[blk.p] = blk;
        }
        *p = blk.p;
    }
    else SQ_TRY(sq_dalloc_traced(ctx, p, nbytes));
    if (zero) CUDA_TRY(cudaMemsetAsync(*p, 0, nbytes, cur));
    return SQ_OK;
}

void sq_dfree(sq_ctx *ctx, void *p) {
    if (!p) return;
    cudaStream_t cur = sq_cur_stream(ctx);
    {
        std::lock_guard<std::mutex> lk(ctx->big_mutex);
        auto it = ctx->big_live.find(p);
        if (it != ctx->big_live.end()) {
            sq_ctx::BigBlock blk = it->second;
            ctx->big_live.erase(it);
            blk.freed_on = cur;
            cudaEventRecord(blk.ev, cur);  // everything that used the block was enqueued on `cur` before this point
            ctx->big_free.push_back(blk);
            ctx->big_free_bytes += blk.cls;
            // idle blocks beyond the cap go back to the driver, oldest first (inputs whose array sizes drift)
            while (ctx->big_free_bytes > ctx->big_free_cap && ctx->big_free.size() > 1) {
                sq_ctx::BigBlock old = ctx->big_free.front();
                ctx->big_free.erase(ctx->big_free.begin());
                ctx->big_free_bytes -= old.cls;
                cudaFreeAsync(old.p, old.freed_on);
                cudaEventDestroy(old.ev);
            }
            return;
        }
    }
    cudaFreeAsync(p, cur);
}

extern "C" int sq_ctx_block_cache_stats(sq_ctx *ctx, uint64_t *from_driver, uint64_t *from_cache, uint64_t *idle_bytes) {
    std::lock_guard<std::mutex> lk(ctx->big_mutex);
    *from_driver = ctx->big_from_driver;
    *from_cache = ctx->big_from_cache;
    *idle_bytes = ctx->big_free_bytes;
    return SQ_OK;
}

// blocks of the cache back to the driver (context teardown)
static void sq_big_release(sq_ctx *ctx) {
    std::lock_guard<std::mutex> lk(ctx->big_mutex);
    for (auto &blk : ctx->big_free) {
        cudaFreeAsync(blk.p, ctx->stream);
        cudaEventDestroy(blk.ev);
    }
    ctx->big_free.clear();
    ctx->big_free_bytes = 0;
    for (auto &kv : ctx->big_live) cudaEventDestroy(kv.second.ev);  // (their owners never freed them)
    ctx->big_live.clear();
}

// Largest x with floor(-10*log10(x)) >= k, found on the bit pattern of the
// double with the HOST libm, so that the device never has to reproduce glibc's
// log10 (reference bucket: _qcmodule.c:2127-2137).
static double phred_bucket_edge(int k) {
    auto ok = [k](double x) { return floor(-10.0 * log10(x)) >= (double)k; };
    uint64_t lo_bits, hi_bits;
    double lo = 1e-300, hi = 1.0;  // ok(lo) is true for every k <= 93, ok(hi) false for k >= 1
    memcpy(&lo_bits, &lo, 8);
    memcpy(&hi_bits, &hi, 8);
    while (hi_bits - lo_bits > 1) {
        uint64_t mid = lo_bits + (hi_bits - lo_bits) / 2;
        double x;
        memcpy(&x, &mid, 8);
        if (ok(x)) lo_bits = mid;
        else hi_bits = mid;
    }
    double r;
    memcpy(&r, &lo_bits, 8);
    return r;
}

// NUMA node the device's PCIe root hangs off (sysfs), -1 when unknown.  Pinned staging memory should live
// there: host->device copies that cross the socket interconnect run at half speed or less.
extern "C" int sq_device_numa_node(int device) {
    char bus[32] = "";
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    for (char *c = bus; *c; c++) *c = (char)tolower(*c);
    char path[128];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}

extern "C" int sq_ctx_create(int device, sq_ctx **out) {
    *out = nullptr;
    int n = sq_device_count();
    if (n <= 0) {
        sq_set_error("no CUDA device visible: libsqgpu has no CPU fallback");
        return SQ_E_NODEVICE;
    }
    if (device < 0 || device >= n) {
        sq_set_error("device %d out of range (%d devices)", device, n);
        return SQ_E_ARG;
    }
    CUDA_TRY(cudaSetDevice(device));
    sq_ctx *ctx = new sq_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    ctx->num_sms = prop.multiProcessorCount;
    // (equal priorities on purpose: with the parser stream below the launch stream the scan of the next record
    // array never overlaps the collectors' kernels and the read-ahead hides nothing -- measured)
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->pstream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->tstream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    // keep freed blocks in the pool: record arrays come and go every batch
    cudaMemPool_t pool;
    CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t threshold = UINT64_MAX;
    CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    ctx->big_free_cap = std::min<size_t>(total_b / 8, (size_t)16 << 30);
    CUDA_TRY(cudaMallocHost(&ctx->h_scratch, 4096));
    CUDA_TRY(cudaMalloc(&ctx->d_scratch, 4096));
    double tab[94], edges[94];
    for (int q = 0; q < 94; q++) tab[q] = pow(10.0, -((double)q / 10.0));
    edges[0] = 1.0;  // bucket 0 catches everything down to edge 1
    for (int k = 1; k < 94; k++) edges[k] = phred_bucket_edge(k);
    CUDA_TRY(cudaMalloc(&ctx->d_err_table, sizeof(tab)));
    CUDA_TRY(cudaMalloc(&ctx->d_phred_thresholds, sizeof(edges)));
    CUDA_TRY(cudaMemcpy(ctx->d_err_table, tab, sizeof(tab), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(ctx->d_phred_thresholds, edges, sizeof(edges), cudaMemcpyHostToDevice));
    *out = ctx;
    return SQ_OK;
}

extern "C" void sq_ctx_destroy(sq_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (getenv("SQ_TRACE_ALLOC")) {
        cudaMemPool_t pool;
        uint64_t reserved = 0, used = 0;
        if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess &&
            cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemHigh, &reserved) == cudaSuccess &&
            cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemHigh, &used) == cudaSuccess)
            fprintf(stderr, "[sq] memory pool high water marks: reserved %.2f GB, in use %.2f GB\n", reserved / 1e9, used / 1e9);
    }
    cudaStreamSynchronize(ctx->pstream);
    cudaStreamSynchronize(ctx->tstream);
    cudaStreamSynchronize(ctx->stream);
    sq_big_release(ctx);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_err_table);
    cudaFree(ctx->d_phred_thresholds);
    cudaFree(ctx->d_scratch);
    cudaFree(ctx->parse_masks);
    cudaFree(ctx->parse_fields);
    cudaFree(ctx->parse_status);
    if (ctx->h_bounce) cudaFreeHost(ctx->h_bounce);
    for (int k = 0; k < 3; k++) cudaFree(ctx->stage_slot[k]);
    cudaFreeHost(ctx->h_scratch);
    cudaEventDestroy(ctx->ev_fork);
    cudaEventDestroy(ctx->ev_join);
    cudaStreamDestroy(ctx->pstream);
    cudaStreamDestroy(ctx->tstream);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

// ---- per-kernel timing: CUDA events around every launch on the launch stream ----
void sq_prof_begin(sq_ctx *ctx, const char *name) {
    sq_ctx::ProfEvent e;
    e.name = name;
    cudaEventCreate(&e.start);
    cudaEventCreate(&e.stop);
    cudaEventRecord(e.start, sq_cur_stream(ctx));
    ctx->prof_events.push_back(e);
}
void sq_prof_end(sq_ctx *ctx) { cudaEventRecord(ctx->prof_events.back().stop, sq_cur_stream(ctx)); }

extern "C" int sq_ctx_profile(sq_ctx *ctx, int enable) {
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    for (auto &e : ctx->prof_events) {
        cudaEventDestroy(e.start);
        cudaEventDestroy(e.stop);
    }
    ctx->prof_events.clear();
    ctx->profile = enable != 0;
    return SQ_OK;
}
// "kernel launches total_ms\n" per kernel, most expensive first
extern "C" int sq_ctx_profile_report(sq_ctx *ctx, char *buf, size_t cap) {
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    struct Row { std::string name; uint64_t n; double ms; };
    std::vector<Row> rows;
    auto add = [&rows](const std::string &name, float ms) {
        for (auto &r : rows)
            if (r.name == name) { r.n++; r.ms += ms; return; }
        rows.push_back({name, 1, ms});
    };
    for (size_t i = 0; i < ctx->prof_events.size(); i++) {
        auto &e = ctx->prof_events[i];
        float ms = 0;
        cudaEventElapsedTime(&ms, e.start, e.stop);
        add(e.name, ms);
        // idle stream time in front of this launch (host work, syncs, copies, memsets), as "gap>kernel"
        if (i) {
            float gap = 0;
            cudaEventElapsedTime(&gap, ctx->prof_events[i - 1].stop, e.start);
            add(std::string("gap>") + e.name, gap);
        }
    }
    for (size_t i = 0; i < rows.size(); i++)
        for (size_t j = i + 1; j < rows.size(); j++)
            if (rows[j].ms > rows[i].ms) std::swap(rows[i], rows[j]);
    std::string out;
    char line[256];
    for (auto &r : rows) {
        snprintf(line, sizeof(line), "%s %llu %.6f\n", r.name.c_str(), (unsigned long long)r.n, r.ms);
        out += line;
    }
    if (cap) {
        size_t k = out.size() < cap - 1 ? out.size() : cap - 1;
        memcpy(buf, out.data(), k);
        buf[k] = 0;
    }
    return SQ_OK;
}

// device-side stopwatch on the launch stream (bench.py times steps with it)
extern "C" int sq_timer_start(sq_ctx *ctx) {
    if (!ctx->timer_start) {
        CUDA_TRY(cudaEventCreate(&ctx->timer_start));
        CUDA_TRY(cudaEventCreate(&ctx->timer_stop));
    }
    CUDA_TRY(cudaEventRecord(ctx->timer_start, ctx->stream));
    return SQ_OK;
}
extern "C" int sq_timer_stop(sq_ctx *ctx, double *ms) {
    CUDA_TRY(cudaEventRecord(ctx->timer_stop, ctx->stream));
    CUDA_TRY(cudaEventSynchronize(ctx->timer_stop));
    float f = 0;
    CUDA_TRY(cudaEventElapsedTime(&f, ctx->timer_start, ctx->timer_stop));
    *ms = f;
    return SQ_OK;
}

extern "C" int sq_ctx_sync(sq_ctx *ctx) {
    CUDA_TRY(cudaStreamSynchronize(ctx->pstream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return SQ_OK;
}
extern "C" void *sq_ctx_stream(sq_ctx *ctx) { return (void *)ctx->stream; }
extern "C" uint64_t sq_ctx_launch_count(sq_ctx *ctx) { return ctx->launches.load(); }

extern "C" void *sq_pinned_alloc(sq_ctx *ctx, size_t nbytes) {
    void *p = nullptr;
    cudaSetDevice(ctx->device);
    if (cudaMallocHost(&p, nbytes ? nbytes : 1) != cudaSuccess) {
        cudaGetLastError();
        sq_set_error("cudaMallocHost(%zu) failed", nbytes);
        return nullptr;
    }
    return p;
}
extern "C" void sq_pinned_free(sq_ctx *ctx, void *p) {
    (void)ctx;
    if (p) cudaFreeHost(p);
}
extern "C" void *sq_device_alloc(sq_ctx *ctx, size_t nbytes) {
    void *p = nullptr;
    cudaSetDevice(ctx->device);
    if (cudaMalloc(&p, nbytes + 64) != cudaSuccess) {  // +64: readable tail for vector loads
        cudaGetLastError();
        sq_set_error("cudaMalloc(%zu) failed", nbytes);
        return nullptr;
    }
    cudaMemsetAsync((uint8_t *)p + nbytes, 0, 64, sq_cur_stream(ctx));
    return p;
}
extern "C" void sq_device_free(sq_ctx *ctx, void *p) {
    if (!p) return;
    cudaStreamSynchronize(sq_cur_stream(ctx));
    cudaFree(p);
}
extern "C" void *sq_stream_alloc(sq_ctx *ctx, uint64_t nbytes) {
    void *p = nullptr;
    cudaSetDevice(ctx->device);
    if (sq_dalloc(ctx, &p, nbytes + 64, false) != SQ_OK) return nullptr;
    cudaMemsetAsync((uint8_t *)p + nbytes, 0, 64, sq_cur_stream(ctx));  // readable tail for vector loads
    return p;
}
extern "C" void sq_stream_free(sq_ctx *ctx, void *p) { sq_dfree(ctx, p); }
extern "C" int sq_stream_memset(sq_ctx *ctx, void *p, int value, uint64_t nbytes) {
    if (nbytes) CUDA_TRY(cudaMemsetAsync(p, value, nbytes, sq_cur_stream(ctx)));
    return SQ_OK;
}
extern "C" int sq_memcpy_h2d(sq_ctx *ctx, void *dst, const void *src, size_t n) {
    CUDA_TRY(cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, sq_cur_stream(ctx)));
    CUDA_TRY(cudaStreamSynchronize(sq_cur_stream(ctx)));
    return SQ_OK;
}
extern "C" int sq_memcpy_d2h(sq_ctx *ctx, void *dst, const void *src, size_t n) {
    CUDA_TRY(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, sq_cur_stream(ctx)));
    CUDA_TRY(cudaStreamSynchronize(sq_cur_stream(ctx)));
    return SQ_OK;
}

// ---------------------------------------------------------------------------
// record-boundary scan
// ---------------------------------------------------------------------------
constexpr int PARSE_THREADS = 256;
constexpr int PARSE_ITERS = 4;                                      // 16-byte vectors per lane
constexpr int PARSE_WARP_BYTES = 32 * 16 * PARSE_ITERS;             // 2 KiB per warp
constexpr int PARSE_CTA_BYTES = (PARSE_THREADS / 32) * PARSE_WARP_BYTES;  // 16 KiB per CTA

struct ParseState {               // lives in device memory, copied back once
    unsigned long long n_newlines;
    unsigned long long first_non_ascii;  // byte offset, ULLONG_MAX if none
    unsigned long long err_key;          // (record << 3) | code, ULLONG_MAX if none
    unsigned int max_seq_len;
    unsigned int max_rec_bytes;  // longest record, '@' through the final newline
};

// 16 bytes at vector index v (text is 16-byte aligned); bytes past nbytes read as 0
__device__ __forceinline__ uint4 load_vec16(const uint8_t *text, uint64_t nbytes, uint64_t v) {
    uint64_t off = v * 16;
    if (off + 16 <= nbytes) return __ldg((const uint4 *)(text + off));
    uint32_t w[4] = {0, 0, 0, 0};
    for (uint64_t i = off; i < nbytes; i++) w[(i - off) >> 2] |= (uint32_t)text[i] << (8 * ((i - off) & 3));
    return make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ uint32_t newline_mask(uint32_t w) { return zero_bytes80(w ^ 0x0A0A0A0Au); }

// bit i of the result = bit 7 of byte i of m (m has at most bit 7 of each byte set)
__device__ __forceinline__ uint32_t pack_msb4(uint32_t m) { return ((m >> 7) * 0x00204081u) >> 21 & 0xFu; }

__global__ void __launch_bounds__(PARSE_THREADS)
k_count_newlines(const uint8_t *__restrict__ text, uint64_t nbytes, uint32_t *__restrict__ cta_counts,
                 uint16_t *__restrict__ vec_masks, ParseState *st) {
    __shared__ uint32_t warp_tot[PARSE_THREADS / 32];
    uint64_t warp_base = (uint64_t)blockIdx.x * PARSE_CTA_BYTES + (uint64_t)(threadIdx.x >> 5) * PARSE_WARP_BYTES;
    uint32_t cnt = 0;
#pragma unroll
    for (int it = 0; it < PARSE_ITERS; it++) {
        uint64_t off = warp_base + (uint64_t)(it * 32 + lane_id()) * 16;
        if (off >= nbytes) continue;
        uint4 v = load_vec16(text, nbytes, off >> 4);
        // one bit per byte: the second pass ranks newlines from these masks (1/8 of the text)
        const uint32_t bits = pack_msb4(newline_mask(v.x)) | pack_msb4(newline_mask(v.y)) << 4 |
                              pack_msb4(newline_mask(v.z)) << 8 | pack_msb4(newline_mask(v.w)) << 12;
        vec_masks[off >> 4] = (uint16_t)bits;
        cnt += __popc(bits);
        uint32_t hi = (v.x | v.y | v.z | v.w) & 0x80808080u;
        if (hi) {  // rare: locate the first byte >= 0x80 (reference :1056-1061)
            uint32_t w[4] = {v.x, v.y, v.z, v.w};
            for (int j = 0; j < 16; j++)
                if ((w[j >> 2] >> (8 * (j & 3))) & 0x80u) {
                    atomicMin(&st->first_non_ascii, (unsigned long long)(off + j));
                    break;
                }
        }
    }
    cnt = warp_sum_u32(cnt);
    if (lane_id() == 0) warp_tot[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = 0;
        for (int i = 0; i < PARSE_THREADS / 32; i++) s += warp_tot[i];
        cta_counts[blockIdx.x] = s;
    }
}

// Second pass over the text: every newline learns its global rank k (CTA offset
// from the scan + warp/CTA prefix) and writes the descriptor field it closes
// straight into the record arrays: record k/4, line k%4.
//   line 0 (name)  -> seq_off[rec]  = p + 1
//   line 1 (seq)   -> seq_len[rec]  = p   (end of the sequence; turned into a length by k_finish_records),
//                     and the '+' check on the byte that follows (:1119-1127)
//   line 2 ('+')   -> qual_off[rec] = p + 1
//   line 3 (qual)  -> name_off[rec + 1] = p + 2, and the '@' check on the next record (:1097)
// Loads stay coalesced; a lane ranks 64 consecutive bytes, so one warp scan covers 2 KiB of text.
__global__ void __launch_bounds__(PARSE_THREADS)
k_scatter_fields(const uint8_t *__restrict__ text, uint64_t nbytes, const uint16_t *__restrict__ vec_masks,
                 const uint32_t *__restrict__ cta_offsets, uint64_t n_rec, int check_partial,
                 uint32_t *__restrict__ name_off, uint32_t *__restrict__ seq_off, uint32_t *__restrict__ seq_len,
                 uint32_t *__restrict__ qual_off, ParseState *st) {
    __shared__ uint32_t warp_tot[PARSE_THREADS / 32];
    const uint32_t warp = threadIdx.x >> 5;
    const uint64_t warp_base = (uint64_t)blockIdx.x * PARSE_CTA_BYTES + (uint64_t)warp * PARSE_WARP_BYTES;
    const uint64_t lane_base = warp_base + (uint64_t)lane_id() * 64;
    // four 16-bit vector masks = the newline bits of this lane's 64 consecutive bytes
    uint64_t mask = 0;
    if (lane_base < nbytes) {
        const uint64_t n_vec = (nbytes + 15) >> 4, v0 = lane_base >> 4;
        if (v0 + 4 <= n_vec) mask = *(const uint64_t *)(vec_masks + v0);
        else
            for (uint64_t v = v0; v < n_vec; v++) mask |= (uint64_t)vec_masks[v] << (16 * (v - v0));
    }
    const uint32_t mine = __popcll(mask);
    uint32_t wsum;
    const uint32_t ex = warp_excl_scan_u32(mine, &wsum);
    if (lane_id() == 0) warp_tot[warp] = wsum;
    __syncthreads();
    if (mine == 0) return;
    uint64_t k = (uint64_t)cta_offsets[blockIdx.x] + ex;
    for (uint32_t i = 0; i < warp; i++) k += warp_tot[i];
    while (mask) {
        const uint64_t p = lane_base + (uint32_t)(__ffsll((long long)mask) - 1);
        mask &= mask - 1;
        const uint64_t rec = k >> 2;
        const uint32_t line = (uint32_t)k & 3;
        k++;
        if (rec > n_rec || (rec == n_rec && !check_partial)) continue;
        if (line == 0) seq_off[rec] = (uint32_t)p + 1;
        else if (line == 1) {
            seq_len[rec] = (uint32_t)p;
            if (p + 1 < nbytes && text[p + 1] != '+')
                atomicMin(&st->err_key, (unsigned long long)(rec << 3 | SQ_PARSE_NO_PLUS));
        }
        else if (line == 2) qual_off[rec] = (uint32_t)p + 1;
        else if (rec < n_rec) {
            name_off[rec + 1] = (uint32_t)p + 2;
            // the record that starts behind this newline: complete, or the partial tail (:1094)
            const uint64_t start = p + 1;
            const bool look = rec + 1 < n_rec || (check_partial && start + 2 < nbytes);
            if (look && text[start] != '@')
                atomicMin(&st->err_key, (unsigned long long)((rec + 1) << 3 | SQ_PARSE_NO_AT));
        }
    }
}

// ---------------------------------------------------------------------------
// One-pass variant: count, rank and scatter in a single walk over the text.
// A CTA takes the 64 KiB tile of its block index (blocks are handed out in index order, so
// every predecessor is already running), counts its newlines, publishes the count and
// finds its global rank offset by decoupled look-back over the predecessors'
// published counts (status word = flag << 62 | value; flag 1: tile count, flag 2:
// inclusive prefix).  Newline positions are then compacted per warp so that the
// lanes turn consecutive ranks into descriptor fields: four consecutive ranks are
// the four fields of one record, and a warp's stores fall into full sectors of the
// four arrays.  The '+' / '@' probes read bytes this CTA has just pulled through
// L1.  Fields go to a scratch of `cap` slots per array (the number of records is
// not known yet); k_finish_records moves them into the record array's own block.
// ---------------------------------------------------------------------------
constexpr int OP_ROUND = 128;  // compacted newline positions per warp and round
constexpr unsigned long long OP_FLAG_COUNT = 1ULL << 62, OP_FLAG_PREFIX = 2ULL << 62, OP_VALUE = (1ULL << 62) - 1;

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// bits 7/15/23/31 of m -> bits 28..31 of the product (every partial product lands on its own bit)
__device__ __forceinline__ uint32_t msb_nibble(uint32_t m) { return (m * 0x00204081u) >> 28; }
// 0x80 in every byte of w that is a newline, for ASCII text (bytes < 0x80: the addition cannot carry
// into the next byte).  A text with a byte >= 0x80 fails with the ASCII error before any field is used.
__device__ __forceinline__ uint32_t newline_mask_ascii(uint32_t w) {
    return ~((w ^ 0x0A0A0A0Au) + 0x7F7F7F7Fu) & 0x80808080u;
}

#ifndef SQ_OP_SUB
#define SQ_OP_SUB 4
#endif
constexpr int OP_SUB = SQ_OP_SUB;                          // 16 KiB sub-tiles per CTA: fewer, longer tiles keep
constexpr int OP_TILE_BYTES = OP_SUB * PARSE_CTA_BYTES;    // the look-back short (64 KiB per CTA)
constexpr int OP_WARPS = PARSE_THREADS / 32;
constexpr int OP_PARTS = OP_SUB * OP_WARPS;  // (sub-tile, warp) counts, scanned by one warp
static_assert(OP_PARTS == 32 || OP_PARTS == 64, "one or two parts per lane of the scanning warp");

__global__ void __launch_bounds__(PARSE_THREADS, OP_SUB <= 4 ? 6 : 4)
k_parse_onepass(const uint8_t *__restrict__ text, uint64_t nbytes, uint32_t n_tiles, uint64_t max_records,
                uint32_t cap, unsigned long long *status,
                uint32_t *__restrict__ fields /* [4][cap]: seq_off, seq_end, qual_off, name_off */, ParseState *st) {
    __shared__ __align__(8) uint16_t s_nl[OP_TILE_BYTES / 16];
    __shared__ uint16_t s_pos[OP_WARPS][OP_ROUND];
    __shared__ uint32_t part_tot[OP_SUB * OP_WARPS], part_excl[OP_SUB * OP_WARPS];  // [sub-tile][warp]
    __shared__ unsigned long long s_prefix;
    // tiles in block-index order: the hardware hands out blocks in that order, so every predecessor
    // of a running block is running or done (what CUB's decoupled look-back relies on as well)
    const uint32_t tile = blockIdx.x, warp = threadIdx.x >> 5, lane = lane_id();
    const uint64_t cta_base = (uint64_t)tile * OP_TILE_BYTES;
    const bool full_tile = cta_base + OP_TILE_BYTES <= nbytes;  // no bounds checks, loads issued back to back
#pragma unroll
    for (int sub = 0; sub < OP_SUB; sub++) {
        uint32_t cnt = 0;
        const uint32_t v0 = sub * (PARSE_CTA_BYTES / 16) + warp * (PARSE_WARP_BYTES / 16) + lane;
        uint4 v[PARSE_ITERS];
        if (full_tile) {
#pragma unroll
            for (int it = 0; it < PARSE_ITERS; it++)
                v[it] = __ldg((const uint4 *)(text + cta_base) + v0 + it * 32);
        }
        else {
#pragma unroll
            for (int it = 0; it < PARSE_ITERS; it++) {
                const uint64_t off = cta_base + (uint64_t)(v0 + it * 32) * 16;
                v[it] = off < nbytes ? load_vec16(text, nbytes, off >> 4) : make_uint4(0, 0, 0, 0);
            }
        }
#pragma unroll
        for (int it = 0; it < PARSE_ITERS; it++) {
            const uint32_t vi = v0 + it * 32;
            const uint32_t bits = msb_nibble(newline_mask_ascii(v[it].x)) | msb_nibble(newline_mask_ascii(v[it].y)) << 4 |
                                  msb_nibble(newline_mask_ascii(v[it].z)) << 8 |
                                  msb_nibble(newline_mask_ascii(v[it].w)) << 12;
            if ((v[it].x | v[it].y | v[it].z | v[it].w) & 0x80808080u) {  // rare: first byte >= 0x80 (:1056-1061)
                const uint32_t w[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
                for (int j = 0; j < 16; j++)
                    if ((w[j >> 2] >> (8 * (j & 3))) & 0x80u) {
                        atomicMin(&st->first_non_ascii, (unsigned long long)(cta_base + (uint64_t)vi * 16 + j));
                        break;
                    }
            }
            s_nl[vi] = (uint16_t)bits;
            cnt += __popc(bits);
        }
        cnt = warp_sum_u32(cnt);
        if (lane == 0) part_tot[sub * OP_WARPS + warp] = cnt;
    }
    // before the prefix is known: this lane's 64 consecutive bytes per sub-tile and its rank inside the warp
    // (the warp wrote these masks itself)
    __syncwarp();
    uint64_t lmask[OP_SUB];
    uint32_t lrank[OP_SUB];
    if (warp != 0) {
#pragma unroll
        for (int sub = 0; sub < OP_SUB; sub++) {
            uint32_t unused;
            lmask[sub] = *(const uint64_t *)(s_nl + sub * (PARSE_CTA_BYTES / 16) + warp * (PARSE_WARP_BYTES / 16) + lane * 4);
            lrank[sub] = warp_excl_scan_u32(__popcll(lmask[sub]), &unused);
        }
    }
    __syncthreads();
    if (warp == 0) {
        uint32_t total;
        if (OP_PARTS == 32) {
            const uint32_t ex = warp_excl_scan_u32(part_tot[lane], &total);
            part_excl[lane] = ex;
        }
        else {
            const uint32_t x0 = part_tot[(2 * lane) % OP_PARTS], x1 = part_tot[(2 * lane + 1) % OP_PARTS];
            const uint32_t ex = warp_excl_scan_u32(x0 + x1, &total);
            part_excl[(2 * lane) % OP_PARTS] = ex;
            part_excl[(2 * lane + 1) % OP_PARTS] = ex + x0;
        }
        if (lane == 0) st_status(status + tile, (tile == 0 ? OP_FLAG_PREFIX : OP_FLAG_COUNT) | total);
        unsigned long long excl = 0;
        if (tile > 0) {
            long long j = (long long)tile - 1 - lane;
            while (true) {
                unsigned long long v = j >= 0 ? ld_status(status + j) : OP_FLAG_PREFIX;
                while (__any_sync(0xffffffffu, (v >> 62) == 0)) {
                    if ((v >> 62) == 0) v = ld_status(status + j);
                }
                const uint32_t pm = __ballot_sync(0xffffffffu, (v >> 62) == 2);
                const uint32_t first = pm ? (uint32_t)__ffs(pm) - 1 : 31u;  // nearest predecessor with a prefix
                unsigned long long add = lane <= first ? (v & OP_VALUE) : 0ULL;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) add += __shfl_xor_sync(0xffffffffu, add, o);
                excl += add;
                if (pm) break;
                j -= 32;
            }
            if (lane == 0) st_status(status + tile, OP_FLAG_PREFIX | (excl + total));
        }
        if (lane == 0) {
            s_prefix = excl;
            if (tile == n_tiles - 1) st->n_newlines = excl + total;
            if (tile == 0) fields[3 * (size_t)cap] = 1;  // name_off[0]: the first record starts at byte 0
        }
#pragma unroll
        for (int sub = 0; sub < OP_SUB; sub++) {  // warp 0's own share of the preparation, off the look-back's path
            uint32_t unused;
            lmask[sub] = *(const uint64_t *)(s_nl + sub * (PARSE_CTA_BYTES / 16) + lane * 4);
            lrank[sub] = warp_excl_scan_u32(__popcll(lmask[sub]), &unused);
        }
    }
    __syncthreads();
    // a record array is < 4 GiB (check_size): byte offsets, newline ranks and record numbers fit 32 bits
    const uint32_t prefix = (uint32_t)s_prefix, nb32 = (uint32_t)nbytes;
    const uint32_t max_rec = (uint32_t)min(max_records, (uint64_t)0xFFFFFFFEu);
#pragma unroll
    for (int sub = 0; sub < OP_SUB; sub++) {
        const uint32_t wsum = part_tot[sub * OP_WARPS + warp];
        if (wsum == 0) continue;
        const uint32_t warp_base = (uint32_t)cta_base + sub * PARSE_CTA_BYTES + warp * PARSE_WARP_BYTES;
        const uint32_t kbase = prefix + part_excl[sub * OP_WARPS + warp];
        uint64_t mask = lmask[sub];
        uint32_t myrank = lrank[sub];  // warp-local rank of this lane's next newline
        for (uint32_t base = 0; base < wsum; base += OP_ROUND) {
            while (mask && myrank < base + OP_ROUND) {
                s_pos[warp][myrank - base] = (uint16_t)(lane * 64 + (uint32_t)(__ffsll((long long)mask) - 1));
                mask &= mask - 1;
                myrank++;
            }
            __syncwarp();
            const uint32_t n_here = min((uint32_t)OP_ROUND, wsum - base);
            for (uint32_t i = lane; i < n_here; i += 32) {
                const uint32_t p = warp_base + s_pos[warp][i];
                const uint32_t k = kbase + base + i;
                const uint32_t rec = k >> 2, line = k & 3;
                if (rec >= max_rec) continue;
                const uint32_t l3 = line == 3, l1 = line == 1;
                const uint32_t slot = rec + l3;
                if (slot < cap) fields[(size_t)line * cap + slot] = p + 1 + l3 - l1;
                if (line & 1) {
                    // line 1: the byte behind the sequence's newline must be '+' (:1119-1127)
                    // line 3: the record that starts behind this newline must start with '@' (:1097):
                    //         complete, or the partial tail when it has at least 3 bytes
                    const bool look = l1 ? p + 1 < nb32 : (rec + 1 < max_rec && (uint64_t)p + 3 < nbytes);
                    if (look && __ldg(text + p + 1) != (l1 ? '+' : '@'))
                        atomicMin(&st->err_key, (unsigned long long)((uint64_t)(rec + l3) << 3 |
                                                                     (l1 ? SQ_PARSE_NO_PLUS : SQ_PARSE_NO_AT)));
                }
            }
            __syncwarp();
        }
    }
}

// One thread per record: sequence length, the equal-length check (:1140), longest
// sequence / record.  Thread 0 also looks at the very first byte of the text.
__global__ void __launch_bounds__(256)
k_finish_records(const uint8_t *__restrict__ text, uint64_t nbytes, uint64_t n_rec, int check_partial,
                 const uint32_t *s_name_off, const uint32_t *s_seq_off, const uint32_t *s_seq_end,
                 const uint32_t *s_qual_off,  // where the scatter left the fields (may be the arrays below)
                 uint32_t *name_off, uint32_t *seq_off, uint32_t *seq_len, uint32_t *qual_off, ParseState *st) {
    uint32_t local_max = 0, local_rec = 0;
    const bool move = s_name_off != name_off;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= n_rec;
         r += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t no = s_name_off[r], so = s_seq_off[r], se = s_seq_end[r], qo = s_qual_off[r];
        if (r == n_rec) {  // slot n_rec: consumed offset, and what a partial tail left (error wording only)
            if (move) {
                name_off[r] = no;
                seq_off[r] = so;
                seq_len[r] = se;
                qual_off[r] = qo;
            }
            break;
        }
        const uint32_t start = no - 1;
        const uint32_t e4 = s_name_off[r + 1] - 2;
        const uint32_t L = se - so;
        if (L != e4 - qo) atomicMin(&st->err_key, (unsigned long long)(r << 3 | SQ_PARSE_LEN));
        if (move) {
            name_off[r] = no;
            seq_off[r] = so;
            qual_off[r] = qo;
        }
        seq_len[r] = L;
        local_max = max(local_max, L);
        local_rec = max(local_rec, e4 + 1 - start);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const bool look = n_rec > 0 || (check_partial && 2 < nbytes);
        if (look && text[0] != '@') atomicMin(&st->err_key, (unsigned long long)(0ULL << 3 | SQ_PARSE_NO_AT));
    }
    local_max = warp_max_u32(local_max);
    local_rec = warp_max_u32(local_rec);
    if (lane_id() == 0 && local_max) atomicMax(&st->max_seq_len, local_max);
    if (lane_id() == 0 && local_rec) atomicMax(&st->max_rec_bytes, local_rec);
}

static int alloc_fastq_metas(sq_batch *b, uint64_t n) {
    // name_off | seq_off | seq_len | qual_off | err_sum
    size_t n4 = (size_t)((n + 1 + 3) & ~3ULL);
    b->meta_stride = n4;
    SQ_TRY(sq_dalloc(b->ctx, &b->meta_block, n4 * 4 * 4 + n4 * 8, false));
    uint32_t *p = (uint32_t *)b->meta_block;
    b->name_off = p;
    b->seq_off = p + n4;
    b->seq_len = p + 2 * n4;
    b->qual_off = p + 3 * n4;
    b->err_sum = (double *)(p + 4 * n4);
    CUDA_TRY(cudaMemsetAsync(b->err_sum, 0, n4 * 8, sq_cur_stream(b->ctx)));
    return SQ_OK;
}

int sq_d2h_bounced(sq_ctx *ctx, void *dst, const void *dev_src, size_t nbytes) {
    const size_t chunk = (size_t)8 << 20;
    if (!ctx->h_bounce) {
        CUDA_TRY(cudaMallocHost(&ctx->h_bounce, 2 * chunk));
        ctx->h_bounce_cap = 2 * chunk;
    }
    // two halves: the copy of chunk i + 1 runs while chunk i is moved to the caller's memory
    size_t done = 0, issued = 0;
    int slot = 0;
    cudaEvent_t ev[2];
    CUDA_TRY(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
    size_t len[2] = {0, 0};
    auto issue = [&](int sl) -> int {
        len[sl] = std::min(chunk, nbytes - issued);
        CUDA_TRY(cudaMemcpyAsync((char *)ctx->h_bounce + sl * chunk, (const char *)dev_src + issued, len[sl],
                                 cudaMemcpyDeviceToHost, sq_cur_stream(ctx)));
        CUDA_TRY(cudaEventRecord(ev[sl], sq_cur_stream(ctx)));
        issued += len[sl];
        return SQ_OK;
    };
    int rc = SQ_OK;
    if (nbytes) rc = issue(0);
    while (rc == SQ_OK && done < nbytes) {
        if (issued < nbytes) rc = issue(slot ^ 1);
        if (rc != SQ_OK) break;
        if (cudaEventSynchronize(ev[slot]) != cudaSuccess) {
            sq_set_error("cudaEventSynchronize failed in sq_d2h_bounced");
            rc = SQ_E_CUDA;
            break;
        }
        memcpy((char *)dst + done, (char *)ctx->h_bounce + slot * chunk, len[slot]);
        done += len[slot];
        slot ^= 1;
    }
    cudaEventDestroy(ev[0]);
    cudaEventDestroy(ev[1]);
    return rc;
}

// grow-only scratch of the context
static int ensure_scratch(sq_ctx *ctx, void **ptr, size_t *cap, size_t need) {
    if (need <= *cap) return SQ_OK;
    CUDA_TRY(cudaStreamSynchronize(sq_cur_stream(ctx)));
    if (*ptr) CUDA_TRY(cudaFree(*ptr));
    *ptr = nullptr;
    *cap = 0;
    need += need / 8;
    CUDA_TRY(cudaMalloc(ptr, need));
    // (slot n_rec of the parser's field scratch is read back even when no partial record wrote it)
    CUDA_TRY(cudaMemset(*ptr, 0, need));
    *cap = need;
    return SQ_OK;
}

static int parse_device_text(sq_ctx *ctx, sq_batch *b, uint64_t max_records, sq_parse_info *info) {
    memset(info, 0, sizeof(*info));
    uint64_t nbytes = b->nbytes;
    if (nbytes == 0) return SQ_OK;
    uint32_t n_cta = (uint32_t)((nbytes + PARSE_CTA_BYTES - 1) / PARSE_CTA_BYTES);
    ParseState *st = (ParseState *)ctx->d_scratch;
    ParseState init;
    init.n_newlines = 0;
    init.first_non_ascii = ~0ULL;
    init.err_key = ~0ULL;
    init.max_seq_len = 0;
    init.max_rec_bytes = 0;
    memcpy(ctx->h_scratch, &init, sizeof(init));
    CUDA_TRY(cudaMemcpyAsync(st, ctx->h_scratch, sizeof(init), cudaMemcpyHostToDevice, sq_cur_stream(ctx)));
    ParseState *hst = (ParseState *)ctx->h_scratch;

    // ---- one pass: count + rank (decoupled look-back) + scatter into a scratch of `cap` slots ----
    // (records shorter than 32 bytes on average do not fit the scratch and take the two-pass path)
    static const bool force_two_pass = getenv("SQ_PARSE_TWO_PASS") != nullptr;
    const uint32_t cap = (uint32_t)(nbytes / 32 + 1024);
    uint32_t *fields = nullptr;
    if (!force_two_pass) {
        SQ_TRY(ensure_scratch(ctx, &ctx->parse_fields, &ctx->parse_fields_cap, (size_t)cap * 16));
        SQ_TRY(ensure_scratch(ctx, &ctx->parse_status, &ctx->parse_status_cap, ((size_t)n_cta + 1) * 8));
        fields = (uint32_t *)ctx->parse_fields;
        unsigned long long *status = (unsigned long long *)ctx->parse_status;
        const uint32_t n_op = (uint32_t)((nbytes + OP_TILE_BYTES - 1) / OP_TILE_BYTES);
        CUDA_TRY(cudaMemsetAsync(status, 0, ((size_t)n_op + 1) * 8, sq_cur_stream(ctx)));
        SQ_LAUNCH(ctx, k_parse_onepass, n_op, PARSE_THREADS, 0, b->text, nbytes, n_op, max_records, cap, status,
                  fields, st);
        CUDA_TRY(cudaMemcpyAsync(hst, st, sizeof(ParseState), cudaMemcpyDeviceToHost, sq_cur_stream(ctx)));
        CUDA_TRY(cudaStreamSynchronize(sq_cur_stream(ctx)));
    }
    // ---- two passes: count (+ bit masks), device-wide scan of the per-CTA counts, scatter ----------
    uint32_t *cta_counts = nullptr;
    uint16_t *vec_masks = nullptr;
    auto count_pass = [&]() -> int {
        SQ_TRY(sq_dalloc(ctx, (void **)&cta_counts, (size_t)n_cta * 4, false));
        // one bit per text byte; a grow-only scratch of the context (a fresh 100+ MB stream-ordered
        // allocation per record array makes the pool re-map memory every time)
        SQ_TRY(ensure_scratch(ctx, &ctx->parse_masks, &ctx->parse_masks_cap, (size_t)(((nbytes + 15) >> 4) + 4) * 2));
        vec_masks = (uint16_t *)ctx->parse_masks;
        SQ_LAUNCH(ctx, k_count_newlines, n_cta, PARSE_THREADS, 0, b->text, nbytes, cta_counts, vec_masks, st);
        // exclusive scan of the per-CTA counts (in place); the total is the number of newlines
        uint32_t *d_total = (uint32_t *)((char *)ctx->d_scratch + 128);
        uint32_t *h_total = (uint32_t *)((char *)ctx->h_scratch + 128);
        SQ_TRY(sq_scan_exclusive_u32(ctx, cta_counts, cta_counts, n_cta, d_total));
        CUDA_TRY(cudaMemcpyAsync(hst, st, sizeof(ParseState), cudaMemcpyDeviceToHost, sq_cur_stream(ctx)));
        CUDA_TRY(cudaMemcpyAsync(h_total, d_total, 4, cudaMemcpyDeviceToHost, sq_cur_stream(ctx)));
        CUDA_TRY(cudaStreamSynchronize(sq_cur_stream(ctx)));
        hst->n_newlines = *h_total;
        return SQ_OK;
    };
    if (force_two_pass) SQ_TRY(count_pass());
    uint64_t n_newlines = hst->n_newlines;
    info->n_newlines = n_newlines;
    if (hst->first_non_ascii != ~0ULL) {  // checked before any record is looked at (:1055)
        info->err_code = SQ_PARSE_ASCII;
        info->err_pos = hst->first_non_ascii;
        sq_dfree(ctx, cta_counts);
        return SQ_E_FORMAT;
    }
    uint64_t n_rec = n_newlines / 4;
    int check_partial = 1;
    if (n_rec >= max_records) {
        n_rec = max_records;
        check_partial = 0;
    }
    SQ_TRY(alloc_fastq_metas(b, n_rec));
    int grid = sq_grid_for(ctx, n_rec + 1, 256);
    if (fields && n_rec + 1 <= cap) {
        SQ_LAUNCH(ctx, k_finish_records, grid, 256, 0, b->text, nbytes, n_rec, check_partial, fields + 3 * (size_t)cap,
                  fields, fields + (size_t)cap, fields + 2 * (size_t)cap, b->name_off, b->seq_off, b->seq_len,
                  b->qual_off, st);
    }
    else {
        if (!cta_counts) SQ_TRY(count_pass());
        // name_off[0]: the first record starts at byte 0; slot n_rec of the other arrays is only
        // written when a partial record follows (and only read to word an error message)
        uint32_t *h_one = (uint32_t *)((char *)ctx->h_scratch + 256);
        *h_one = 1;
        CUDA_TRY(cudaMemcpyAsync(b->name_off, h_one, 4, cudaMemcpyHostToDevice, sq_cur_stream(ctx)));
        if (n_newlines)
            SQ_LAUNCH(ctx, k_scatter_fields, n_cta, PARSE_THREADS, 0, b->text, nbytes, vec_masks, cta_counts, n_rec,
                      check_partial, b->name_off, b->seq_off, b->seq_len, b->qual_off, st);
        SQ_LAUNCH(ctx, k_finish_records, grid, 256, 0, b->text, nbytes, n_rec, check_partial, b->name_off, b->seq_off,
                  b->seq_len, b->qual_off, b->name_off, b->seq_off, b->seq_len, b->qual_off, st);
    }
    // the consumed offset is the byte after the last record's 4th newline = name_off[n_rec] - 1
    uint32_t *h_last = (uint32_t *)((char *)ctx->h_scratch + 260);
    *h_last = 1;
    if (n_rec) CUDA_TRY(cudaMemcpyAsync(h_last, b->name_off + n_rec, 4, cudaMemcpyDeviceToHost, sq_cur_stream(ctx)));
    CUDA_TRY(cudaMemcpyAsync(hst, st, sizeof(ParseState), cudaMemcpyDeviceToHost, sq_cur_stream(ctx)));
    CUDA_TRY(cudaStreamSynchronize(sq_cur_stream(ctx)));
    info->n_records = n_rec;
    info->consumed = n_rec ? (uint64_t)*h_last - 1 : 0;
    info->max_seq_len = hst->max_seq_len;
    int rc = SQ_OK;
    if (hst->err_key != ~0ULL) {
        uint64_t r = hst->err_key >> 3;
        info->err_code = (int32_t)(hst->err_key & 7);
        info->err_record = r;
        // locate the offending byte for the message (rare path, tiny copies)
        uint32_t v[3] = {1, 0, 0};  // name_off, seq_off, seq_len (or, for the partial tail, the sequence end)
        if (r > 0) SQ_TRY(sq_memcpy_d2h(ctx, &v[0], b->name_off + r, 4));
        const uint64_t start = (uint64_t)v[0] - 1;
        if (info->err_code == SQ_PARSE_NO_AT) info->err_pos = start;
        else if (info->err_code == SQ_PARSE_NO_PLUS) {
            SQ_TRY(sq_memcpy_d2h(ctx, &v[1], b->seq_off + r, 4));
            SQ_TRY(sq_memcpy_d2h(ctx, &v[2], b->seq_len + r, 4));
            info->err_pos = (r < n_rec ? (uint64_t)v[1] + v[2] : (uint64_t)v[2]) + 1;
        }
        else info->err_pos = start + 1;
        rc = SQ_E_FORMAT;
    }
    b->n = n_rec;
    b->max_len = info->max_seq_len;
    b->max_rec_bytes = hst->max_rec_bytes;
    b->text_end = info->consumed;
    sq_dfree(ctx, cta_counts);
    return rc;
}

static int check_size(uint64_t nbytes) {
    if (nbytes >= 0xFFFFFF00ULL) {
        sq_set_error("record array of %llu bytes exceeds the 4 GiB offset range",
                     (unsigned long long)nbytes);
        return SQ_E_LIMIT;
    }
    return SQ_OK;
}

extern "C" void sq_batch_free(sq_batch *b) {
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    if (b->owns_text) sq_dfree(b->ctx, b->text);
    sq_dfree(b->ctx, b->meta_block);
    delete b;
}

extern "C" int sq_batch_from_fastq(sq_ctx *ctx, const uint8_t *text, uint64_t nbytes,
                                   uint64_t max_records, sq_batch **out, sq_parse_info *info) {
    *out = nullptr;
    SQ_TRY(check_size(nbytes));
    CUDA_TRY(cudaSetDevice(ctx->device));
    SqParserScope on_parser_stream(ctx);
    sq_batch *b = new sq_batch();
    b->ctx = ctx;
    b->nbytes = nbytes;
    int rc = sq_dalloc(ctx, (void **)&b->text, nbytes + 64, false);
    if (rc == SQ_OK && nbytes)
        rc = cudaMemcpyAsync(b->text, text, nbytes, cudaMemcpyHostToDevice, sq_cur_stream(ctx)) == cudaSuccess
                 ? SQ_OK : sq_cuda_fail(cudaGetLastError(), "H2D text", __FILE__, __LINE__);
    if (rc == SQ_OK) cudaMemsetAsync(b->text + nbytes, 0, 64, sq_cur_stream(ctx));
    if (rc == SQ_OK) rc = parse_device_text(ctx, b, max_records, info);
    if (rc != SQ_OK) {
        sq_batch_free(b);
        return rc;
    }
    *out = b;
    return SQ_OK;
}

extern "C" int sq_batch_from_device_fastq(sq_ctx *ctx, const uint8_t *dev_text, uint64_t nbytes,
                                          uint64_t max_records, sq_batch **out, sq_parse_info *info) {
    *out = nullptr;
    SQ_TRY(check_size(nbytes));
    if (((uintptr_t)dev_text & 15) != 0) {
        sq_set_error("device text must be 16-byte aligned");
        return SQ_E_ARG;
    }
    CUDA_TRY(cudaSetDevice(ctx->device));
    SqParserScope on_parser_stream(ctx);
    sq_batch *b = new sq_batch();
    b->ctx = ctx;
    b->nbytes = nbytes;
    b->text = (uint8_t *)dev_text;
    b->owns_text = false;
    int rc = parse_device_text(ctx, b, max_records, info);
    if (rc != SQ_OK) {
        sq_batch_free(b);
        return rc;
    }
    *out = b;
    return SQ_OK;
}

extern "C" int sq_batch_from_packed(sq_ctx *ctx, const uint8_t *buf, uint64_t nbytes,
                                    const sq_meta *metas, uint64_t n, sq_batch **out) {
    *out = nullptr;
    SQ_TRY(check_size(nbytes));
    CUDA_TRY(cudaSetDevice(ctx->device));
    sq_batch *b = new sq_batch();
    b->ctx = ctx;
    b->nbytes = nbytes;
    b->n = n;
    size_t n4 = (size_t)((n + 3) & ~3ULL);
    size_t meta_bytes = n4 * 4 * 7 + n4 * 8;
    std::vector<uint8_t> host(meta_bytes ? meta_bytes : 16, 0);
    uint32_t *p = (uint32_t *)host.data();
    uint32_t max_len = 0;
    for (uint64_t i = 0; i < n; i++) {
        const sq_meta &m = metas[i];
        if ((uint64_t)m.name_off + m.name_len > nbytes || (uint64_t)m.seq_off + m.seq_len > nbytes ||
            (uint64_t)m.qual_off + m.seq_len > nbytes || (uint64_t)m.tags_off + m.tags_len > nbytes) {
            delete b;
            sq_set_error("record %llu points outside the buffer", (unsigned long long)i);
            return SQ_E_ARG;
        }
        p[i] = m.name_off;
        p[n4 + i] = m.seq_off;
        p[2 * n4 + i] = m.seq_len;
        p[3 * n4 + i] = m.qual_off;
        p[4 * n4 + i] = m.name_len;
        p[5 * n4 + i] = m.tags_off;
        p[6 * n4 + i] = m.tags_len;
        ((double *)(p + 7 * n4))[i] = m.err_sum;
        if (m.seq_len > max_len) max_len = m.seq_len;
    }
    b->max_len = max_len;
    int rc = sq_dalloc(ctx, (void **)&b->text, nbytes + 64, false);
    if (rc == SQ_OK) rc = sq_dalloc(ctx, &b->meta_block, meta_bytes, false);
    if (rc != SQ_OK) {
        sq_batch_free(b);
        return rc;
    }
    // pageable sources: cudaMemcpyAsync stages them before returning
    if (nbytes) CUDA_TRY(cudaMemcpyAsync(b->text, buf, nbytes, cudaMemcpyHostToDevice, sq_cur_stream(ctx)));
    CUDA_TRY(cudaMemsetAsync(b->text + nbytes, 0, 64, sq_cur_stream(ctx)));
    if (meta_bytes)
        CUDA_TRY(cudaMemcpyAsync(b->meta_block, host.data(), meta_bytes, cudaMemcpyHostToDevice, sq_cur_stream(ctx)));
    CUDA_TRY(cudaStreamSynchronize(sq_cur_stream(ctx)));
    uint32_t *d = (uint32_t *)b->meta_block;
    b->name_off = d;
    b->seq_off = d + n4;
    b->seq_len = d + 2 * n4;
    b->qual_off = d + 3 * n4;
    b->name_len = d + 4 * n4;
    b->tags_off = d + 5 * n4;
    b->tags_len = d + 6 * n4;
    b->err_sum = (double *)(d + 7 * n4);
    *out = b;
    return SQ_OK;
}

// ---------------------------------------------------------------------------
// Double-buffered host reader: the device-side stand-in for the reference's
// read loop (FastqParser__next__ -> readinto -> leftover copy,
// _qcmodule.c:985-1029, 1175-1180) when the uncompressed text sits in (pinned)
// host memory.  Fixed-size windows of raw bytes go host -> device on a COPY
// stream into a ring of staging slots, ahead of the parser; the parser's
// stream only waits for the window it is about to scan, then places
// [leftover of the previous array | window] in a fresh 16-byte aligned text
// buffer with two device-to-device copies.  H2D of the next windows overlaps
// with the boundary scan and the collectors of the current record array.
// ---------------------------------------------------------------------------
// dst[0..n) = src[0..n) for any alignment of either side, on the SMs: the copy engines
// stay free for the host->device windows that are in flight on the copy stream
__global__ void __launch_bounds__(256)
k_copy_bytes(uint8_t *__restrict__ dst, const uint8_t *__restrict__ src, uint64_t n) {
    const uint64_t head = min(n, (uint64_t)((16 - ((uintptr_t)dst & 15)) & 15));
    const uint64_t n_vec = (n - head) >> 4, tail0 = head + (n_vec << 4);
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (uint64_t)gridDim.x * blockDim.x;
    if (tid < head) dst[tid] = src[tid];
    if (tid < n - tail0) dst[tail0 + tid] = src[tail0 + tid];
    const uint8_t *s0 = src + head;
    const uint32_t sh = (uint32_t)((uintptr_t)s0 & 3) * 8;
    const uint32_t *sw = (const uint32_t *)((uintptr_t)s0 & ~(uintptr_t)3);
    uint4 *dv = (uint4 *)(dst + head);
    for (uint64_t v = tid; v < n_vec; v += nthr) {
        const uint32_t *w = sw + v * 4;
        const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2), w3 = __ldg(w + 3);
        const uint32_t w4 = sh ? __ldg(w + 4) : 0u;  // unshifted copies never read past the source
        dv[v] = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh),
                           __funnelshift_r(w3, w4, sh));
    }
}
static int copy_bytes(sq_ctx *ctx, uint8_t *dst, const uint8_t *src, uint64_t n) {
    if (n == 0) return SQ_OK;
    SQ_LAUNCH(ctx, k_copy_bytes, sq_grid_for(ctx, (n >> 4) + 32, 256, 16), 256, 0, dst, src, n);
    return SQ_OK;
}

struct sq_fastq_stream {
    static constexpr int SLOTS = 3;
    sq_ctx *ctx = nullptr;
    const uint8_t *host = nullptr;
    uint64_t nbytes = 0, window = 0;
    cudaStream_t copy = nullptr;
    uint8_t *slot[SLOTS] = {nullptr, nullptr, nullptr};
    cudaEvent_t filled[SLOTS] = {nullptr, nullptr, nullptr}, drained[SLOTS] = {nullptr, nullptr, nullptr};
    uint64_t slot_len[SLOTS] = {0, 0, 0};
    uint64_t issue_pos = 0;            // host offset of the next window to copy
    uint64_t n_issued = 0, n_taken = 0;
    uint8_t *tail = nullptr;           // bytes behind the last complete record of the previous array
    uint64_t tail_len = 0, tail_cap = 0;
    bool slots_from_ctx = false;
    // BGZF input (sq_fastq_stream_create_bgzf): windows of whole members travel compressed
    bool bgzf = false;
    std::vector<sq_bgzf_block> blocks;  // every member of the stream (host)
    sq_bgzf_block *d_blocks = nullptr;  // the same on the device
    struct Win { uint64_t b0, b1, comp0, comp1, text; };  // members [b0, b1), their bytes in the stream, text bytes
    std::vector<Win> wins;
    uint64_t slot_win[SLOTS] = {0, 0, 0};
    unsigned long long *d_bad = nullptr;  // first member that failed to inflate (block << 8 | code)
};

int bgzf_inflate_async(sq_ctx *ctx, const uint8_t *dev_comp, uint64_t comp_base, const sq_bgzf_block *dev_blocks, uint64_t n,
                       uint64_t text_base, uint8_t *dev_out, unsigned long long *dev_first_bad);

static int stream_issue(sq_fastq_stream *s) {
    if (s->bgzf) {
        while (s->n_issued - s->n_taken < (uint64_t)sq_fastq_stream::SLOTS && s->n_issued < s->wins.size()) {
            const int k = (int)(s->n_issued % sq_fastq_stream::SLOTS);
            const sq_fastq_stream::Win &w = s->wins[s->n_issued];
            if (s->n_issued >= (uint64_t)sq_fastq_stream::SLOTS) CUDA_TRY(cudaStreamWaitEvent(s->copy, s->drained[k], 0));
            CUDA_TRY(cudaMemcpyAsync(s->slot[k], s->host + w.comp0, w.comp1 - w.comp0, cudaMemcpyHostToDevice, s->copy));
            CUDA_TRY(cudaEventRecord(s->filled[k], s->copy));
            s->slot_len[k] = w.text;
            s->slot_win[k] = s->n_issued;
            s->n_issued++;
        }
        return SQ_OK;
    }
    while (s->n_issued - s->n_taken < (uint64_t)sq_fastq_stream::SLOTS && s->issue_pos < s->nbytes) {
        const int k = (int)(s->n_issued % sq_fastq_stream::SLOTS);
        const uint64_t len = s->nbytes - s->issue_pos < s->window ? s->nbytes - s->issue_pos : s->window;
        if (s->n_issued >= (uint64_t)sq_fastq_stream::SLOTS) CUDA_TRY(cudaStreamWaitEvent(s->copy, s->drained[k], 0));
        CUDA_TRY(cudaMemcpyAsync(s->slot[k], s->host + s->issue_pos, len, cudaMemcpyHostToDevice, s->copy));
        CUDA_TRY(cudaEventRecord(s->filled[k], s->copy));
        s->slot_len[k] = len;
        s->issue_pos += len;
        s->n_issued++;
    }
    return SQ_OK;
}

extern "C" void sq_fastq_stream_destroy(sq_fastq_stream *s) {
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    if (s->copy) cudaStreamSynchronize(s->copy);
    cudaStreamSynchronize(s->ctx->pstream);
    cudaStreamSynchronize(s->ctx->stream);
    if (s->slots_from_ctx) s->ctx->stage_in_use = false;
    for (int k = 0; k < sq_fastq_stream::SLOTS; k++) {
        if (s->slot[k] && !s->slots_from_ctx) cudaFree(s->slot[k]);
        if (s->filled[k]) cudaEventDestroy(s->filled[k]);
        if (s->drained[k]) cudaEventDestroy(s->drained[k]);
    }
    if (s->tail) cudaFree(s->tail);
    if (s->d_blocks) cudaFree(s->d_blocks);
    if (s->d_bad) cudaFree(s->d_bad);
    if (s->copy) cudaStreamDestroy(s->copy);
    delete s;
}

// staging ring + events of a reader whose slots hold `slot_bytes` each; starts the first copies
static int stream_setup(sq_ctx *ctx, sq_fastq_stream *s, uint64_t slot_bytes) {
    int rc = SQ_OK;
    auto fail = [&](cudaError_t e, const char *what) {
        if (e != cudaSuccess && rc == SQ_OK) rc = sq_cuda_fail(e, what, __FILE__, __LINE__);
    };
    fail(cudaStreamCreateWithFlags(&s->copy, cudaStreamNonBlocking), "copy stream");
    // the staging ring lives in the context between readers (page-mapping 3 x window per pass is slow)
    if (!ctx->stage_in_use) {
        if (ctx->stage_cap < slot_bytes + 64) {
            cudaStreamSynchronize(ctx->pstream);
            cudaStreamSynchronize(ctx->stream);
            for (int k = 0; k < 3; k++) {
                if (ctx->stage_slot[k]) cudaFree(ctx->stage_slot[k]);
                ctx->stage_slot[k] = nullptr;
            }
            ctx->stage_cap = 0;
            for (int k = 0; k < 3; k++) fail(cudaMalloc(&ctx->stage_slot[k], slot_bytes + 64), "staging slot");
            if (rc == SQ_OK) ctx->stage_cap = slot_bytes + 64;
        }
        if (rc == SQ_OK) {
            for (int k = 0; k < 3; k++) s->slot[k] = (uint8_t *)ctx->stage_slot[k];
            s->slots_from_ctx = true;
            ctx->stage_in_use = true;
        }
    }
    for (int k = 0; k < sq_fastq_stream::SLOTS && rc == SQ_OK; k++) {
        if (!s->slots_from_ctx) fail(cudaMalloc(&s->slot[k], slot_bytes + 64), "staging slot");
        fail(cudaEventCreateWithFlags(&s->filled[k], cudaEventDisableTiming), "event");
        fail(cudaEventCreateWithFlags(&s->drained[k], cudaEventDisableTiming), "event");
    }
    if (rc == SQ_OK) rc = stream_issue(s);
    return rc;
}

extern "C" int sq_fastq_stream_create(sq_ctx *ctx, const uint8_t *host_text, uint64_t nbytes, uint64_t window,
                                      sq_fastq_stream **out) {
    *out = nullptr;
    if (window < 4096 || window >= 0xF0000000ULL) {
        sq_set_error("window must be between 4 KiB and 3.75 GiB");
        return SQ_E_ARG;
    }
    CUDA_TRY(cudaSetDevice(ctx->device));
    sq_fastq_stream *s = new sq_fastq_stream();
    s->ctx = ctx;
    s->host = host_text;
    s->nbytes = nbytes;
    s->window = window;
    int rc = stream_setup(ctx, s, window);
    if (rc != SQ_OK) {
        sq_fastq_stream_destroy(s);
        return rc;
    }
    *out = s;
    return SQ_OK;
}

extern "C" int sq_fastq_stream_create_bgzf(sq_ctx *ctx, const uint8_t *host_bgzf, uint64_t nbytes, uint64_t window,
                                           sq_fastq_stream **out) {
    *out = nullptr;
    if (window < 65536 || window >= 0xF0000000ULL) {
        sq_set_error("window must be between 64 KiB and 3.75 GiB of text");
        return SQ_E_ARG;
    }
    CUDA_TRY(cudaSetDevice(ctx->device));
    sq_fastq_stream *s = new sq_fastq_stream();
    s->ctx = ctx;
    s->host = host_bgzf;
    s->nbytes = nbytes;
    s->window = window;
    s->bgzf = true;
    // every member header of the stream (a hop per member: ~64 KiB of text each)
    uint64_t n_blocks = 0, consumed = 0, text = 0;
    int rc = sq_bgzf_scan(host_bgzf, nbytes, nullptr, 0, &n_blocks, &consumed, &text);  // count, then list
    if (rc == SQ_OK) {
        s->blocks.resize(n_blocks + 1);
        rc = sq_bgzf_scan(host_bgzf, nbytes, s->blocks.data(), s->blocks.size(), &n_blocks, &consumed, &text);
    }
    if (rc == SQ_OK && consumed != nbytes) {
        sq_set_error("truncated BGZF stream: %llu bytes behind the last complete member",
                     (unsigned long long)(nbytes - consumed));
        rc = SQ_E_FORMAT;
    }
    if (rc != SQ_OK) {
        delete s;
        return rc;
    }
    s->blocks.resize(n_blocks);
    uint64_t max_comp = 0;
    for (uint64_t b0 = 0; b0 < n_blocks;) {
        uint64_t b1 = b0, t = 0;
        while (b1 < n_blocks && (b1 == b0 || t + s->blocks[b1].text_len <= window)) t += s->blocks[b1++].text_len;
        sq_fastq_stream::Win w;
        w.b0 = b0;
        w.b1 = b1;
        w.comp0 = s->blocks[b0].comp_off;
        w.comp1 = s->blocks[b1 - 1].comp_off + s->blocks[b1 - 1].comp_len;
        w.text = t;
        if (t) s->wins.push_back(w);  // (members without text -- the BGZF end marker -- carry nothing)
        if (w.comp1 - w.comp0 > max_comp) max_comp = w.comp1 - w.comp0;
        b0 = b1;
    }
    auto fail = [&](cudaError_t e, const char *what) {
        if (e != cudaSuccess && rc == SQ_OK) rc = sq_cuda_fail(e, what, __FILE__, __LINE__);
    };
    if (n_blocks) {
        fail(cudaMalloc(&s->d_blocks, n_blocks * sizeof(sq_bgzf_block)), "member table");
        if (rc == SQ_OK)
            fail(cudaMemcpy(s->d_blocks, s->blocks.data(), n_blocks * sizeof(sq_bgzf_block), cudaMemcpyHostToDevice), "member table");
    }
    fail(cudaMalloc(&s->d_bad, 8), "status word");
    if (rc == SQ_OK) fail(cudaMemset(s->d_bad, 0xFF, 8), "status word");
    if (rc == SQ_OK) rc = stream_setup(ctx, s, max_comp ? max_comp : 4096);
    if (rc != SQ_OK) {
        sq_fastq_stream_destroy(s);
        return rc;
    }
    *out = s;
    return SQ_OK;
}

extern "C" uint64_t sq_fastq_stream_leftover(const sq_fastq_stream *s) { return s->tail_len; }

// The next record array, or *out == NULL at the end of the text (bytes of a
// trailing partial record: sq_fastq_stream_leftover).  A window that holds no
// complete record is carried over whole and joined with the next one.
extern "C" int sq_fastq_stream_next(sq_fastq_stream *s, sq_batch **out, sq_parse_info *info) {
    *out = nullptr;
    memset(info, 0, sizeof(*info));
    sq_ctx *ctx = s->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    SqParserScope on_parser_stream(ctx);
    while (s->n_taken < s->n_issued) {
        const int k = (int)(s->n_taken % sq_fastq_stream::SLOTS);
        const uint64_t len = s->slot_len[k], total = s->tail_len + len;
        SQ_TRY(check_size(total));
        sq_batch *b = new sq_batch();
        b->ctx = ctx;
        b->nbytes = total;
        int rc = sq_dalloc(ctx, (void **)&b->text, total + 64, false);
        if (rc != SQ_OK) {
            delete b;
            return rc;
        }
        CUDA_TRY(cudaStreamWaitEvent(sq_cur_stream(ctx), s->filled[k], 0));
        SQ_TRY(copy_bytes(ctx, b->text, s->tail, s->tail_len));
        if (s->bgzf) {
            // the members of this window inflate straight behind the leftover of the previous array
            const sq_fastq_stream::Win &w = s->wins[s->slot_win[k]];
            SQ_TRY(bgzf_inflate_async(ctx, s->slot[k], w.comp0, s->d_blocks + w.b0, w.b1 - w.b0, s->blocks[w.b0].text_off,
                                      b->text + s->tail_len, s->d_bad));
        }
        else SQ_TRY(copy_bytes(ctx, b->text + s->tail_len, s->slot[k], len));
        CUDA_TRY(cudaMemsetAsync(b->text + total, 0, 64, sq_cur_stream(ctx)));
        CUDA_TRY(cudaEventRecord(s->drained[k], sq_cur_stream(ctx)));
        s->n_taken++;
        SQ_TRY(stream_issue(s));  // the slot refills as soon as the copy above has run
        rc = parse_device_text(ctx, b, UINT64_MAX, info);
        if (s->bgzf) {  // a member that failed to inflate explains (and outranks) whatever the parser saw
            unsigned long long bad = ~0ULL;
            cudaMemcpyAsync(&bad, s->d_bad, 8, cudaMemcpyDeviceToHost, sq_cur_stream(ctx));
            cudaStreamSynchronize(sq_cur_stream(ctx));
            if (bad != ~0ULL) {
                sq_set_error("BGZF member %llu is corrupt (inflate error %d)", (unsigned long long)(bad >> 8), (int)(bad & 0xFF));
                sq_batch_free(b);
                memset(info, 0, sizeof(*info));
                info->err_code = SQ_PARSE_INFLATE;
                info->err_record = bad >> 8;
                return SQ_E_FORMAT;
            }
        }
        if (rc != SQ_OK) {
            sq_batch_free(b);
            return rc;
        }
        const uint64_t left = total - info->consumed;
        if (left > s->tail_cap) {
            CUDA_TRY(cudaStreamSynchronize(sq_cur_stream(ctx)));
            if (s->tail) CUDA_TRY(cudaFree(s->tail));
            s->tail = nullptr;
            s->tail_cap = 0;
            CUDA_TRY(cudaMalloc(&s->tail, left + left / 2 + 4096));
            s->tail_cap = left + left / 2 + 4096;
        }
        SQ_TRY(copy_bytes(ctx, s->tail, b->text + info->consumed, left));
        // (the caller may free the record array on the launch stream at any time: the copy out of it is done first)
        if (left) CUDA_TRY(cudaStreamSynchronize(sq_cur_stream(ctx)));
        s->tail_len = left;
        if (info->n_records == 0) {  // no complete record yet: join with the next window
            sq_batch_free(b);
            continue;
        }
        *out = b;
        return SQ_OK;
    }
    return SQ_OK;
}

extern "C" uint64_t sq_batch_size(const sq_batch *b) { return b->n; }
extern "C" uint64_t sq_batch_nbytes(const sq_batch *b) { return b->nbytes; }
extern "C" uint32_t sq_batch_max_seq_len(const sq_batch *b) { return b->max_len; }

extern "C" int sq_batch_get_bytes(sq_batch *b, uint8_t *out) {
    CUDA_TRY(cudaSetDevice(b->ctx->device));
    if (b->nbytes) CUDA_TRY(cudaMemcpyAsync(out, b->text, b->nbytes, cudaMemcpyDeviceToHost, sq_cur_stream(b->ctx)));
    CUDA_TRY(cudaStreamSynchronize(sq_cur_stream(b->ctx)));
    return SQ_OK;
}

extern "C" int sq_batch_get_metas(sq_batch *b, sq_meta *out) {
    CUDA_TRY(cudaSetDevice(b->ctx->device));
    uint64_t n = b->n;
    if (n == 0) return SQ_OK;
    size_t n4 = b->meta_stride ? (size_t)b->meta_stride : (size_t)((n + 3) & ~3ULL);
    bool aux = b->name_len != nullptr;
    size_t bytes = n4 * 4 * (aux ? 7 : 4) + n4 * 8;
    std::vector<uint8_t> host(bytes);
    CUDA_TRY(cudaMemcpyAsync(host.data(), b->meta_block, bytes, cudaMemcpyDeviceToHost, sq_cur_stream(b->ctx)));
    CUDA_TRY(cudaStreamSynchronize(sq_cur_stream(b->ctx)));
    const uint32_t *p = (const uint32_t *)host.data();
    const double *es = (const double *)(p + (aux ? 7 : 4) * n4);
    for (uint64_t i = 0; i < n; i++) {
        sq_meta &m = out[i];
        m.name_off = p[i];
        m.seq_off = p[n4 + i];
        m.seq_len = p[2 * n4 + i];
        m.qual_off = p[3 * n4 + i];
        if (aux) {
            m.name_len = p[4 * n4 + i];
            m.tags_off = p[5 * n4 + i];
            m.tags_len = p[6 * n4 + i];
        }
        else {  // FASTQ text: "@name\nSEQ\n+...\nQUAL\n"
            m.name_len = m.seq_off - 1 - m.name_off;
            m.tags_off = m.qual_off + m.seq_len;
            m.tags_len = 0;
        }
        m.reserved = 0;
        m.err_sum = es[i];
    }
    return SQ_OK;
}

// name of record r (for skipped_reason messages): three 4-byte reads + the name bytes
int sq_batch_get_name(sq_batch *b, uint64_t r, std::vector<uint8_t> &out) {
    sq_ctx *ctx = b->ctx;
    uint32_t name_off = 0, name_len = 0, seq_off = 0;
    SQ_TRY(sq_memcpy_d2h(ctx, &name_off, b->name_off + r, 4));
    if (b->name_len) SQ_TRY(sq_memcpy_d2h(ctx, &name_len, b->name_len + r, 4));
    else {
        SQ_TRY(sq_memcpy_d2h(ctx, &seq_off, b->seq_off + r, 4));
        name_len = seq_off - 1 - name_off;
    }
    out.resize(name_len);
    if (name_len) SQ_TRY(sq_memcpy_d2h(ctx, out.data(), b->text + name_off, name_len));
    return SQ_OK;
}

// ---------------------------------------------------------------------------
// mate check (reference _qcmodule.c:778-850)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t view_name_len(const BatchView &v, uint32_t r) {
    return v.name_len ? v.name_len[r] : v.seq_off[r] - 1 - v.name_off[r];
}

__global__ void __launch_bounds__(256) k_is_mate(BatchView a, BatchView b, unsigned long long *first_bad) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < a.n; r += gridDim.x * blockDim.x) {
        const uint8_t *n1 = a.text + a.name_off[r], *n2 = b.text + b.name_off[r];
        uint32_t l1 = view_name_len(a, r), l2 = view_name_len(b, r);
        uint32_t id = 0;
        while (id < l1 && n1[id] != ' ' && n1[id] != '\t') id++;
        bool ok = l2 >= id;
        if (ok && l2 > id) ok = n2[id] == ' ' || n2[id] == '\t';
        if (ok) {
            uint32_t cmp = id;
            if (id > 0) {
                uint8_t c1 = n1[id - 1], c2 = n2[id - 1];
                if ((c1 == '1' || c1 == '2') && (c2 == '1' || c2 == '2')) cmp--;
            }
            for (uint32_t i = 0; i < cmp && ok; i++) ok = n1[i] == n2[i];
        }
        if (!ok) atomicMin(first_bad, (unsigned long long)r);
    }
}

extern "C" int sq_batch_is_mate(sq_batch *a, sq_batch *b, uint64_t *first_mismatch) {
    if (a->n != b->n) {
        sq_set_error("record arrays differ in length: %llu vs %llu", (unsigned long long)a->n,
                     (unsigned long long)b->n);
        return SQ_E_ARG;
    }
    sq_ctx *ctx = a->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    *first_mismatch = a->n;
    if (a->n == 0) return SQ_OK;
    unsigned long long *d = (unsigned long long *)((char *)ctx->d_scratch + 512);
    unsigned long long *h = (unsigned long long *)((char *)ctx->h_scratch + 512);
    *h = a->n;
    CUDA_TRY(cudaMemcpyAsync(d, h, 8, cudaMemcpyHostToDevice, sq_cur_stream(ctx)));
    SQ_LAUNCH(ctx, k_is_mate, sq_grid_for(ctx, a->n, 256), 256, 0, a->view(), b->view(), d);
    CUDA_TRY(cudaMemcpyAsync(h, d, 8, cudaMemcpyDeviceToHost, sq_cur_stream(ctx)));
    CUDA_TRY(cudaStreamSynchronize(sq_cur_stream(ctx)));
    *first_mismatch = *h;
    return SQ_OK;
}
