// nanostats.cu -- NanoStats (reference _qcmodule.c:248-322, 5006-5052, 5078-5259, 5269-5324).
//
// One thread per record turns the guppy-style FASTQ header ("... ch=<n>
// start_time=<iso8601> ...") or the BAM aux fields (ch, st, du, pi) into one
// 40-byte NanoInfo, in record order, in a device array that grows with the
// input.  The first FASTQ header that cannot be parsed switches the module off:
// that record index is taken by atomicMin and everything from it on is ignored
// at read-out.  The per-read error sum is the one QCMetrics stored on the
// record array (same stream, so it is already there when this kernel runs).
#include "common.cuh"
#include "modules.cuh"

constexpr int NS_TPB = 128;

__device__ __forceinline__ long long dec_field(const uint8_t *s, const uint8_t *end, uint32_t len) {
    if (len < 1 || len > 18 || s + len > end) return -1;
    long long v = 0;
    for (uint32_t i = 0; i < len; i++) {
        uint32_t d = (uint32_t)s[i] - '0';
        if (d > 9) return -1;
        v = v * 10 + d;
    }
    return v;
}

// "YYYY-MM-DDThh:mm:ss[.fff](Z|+hh:mm|-hh:mm)" -> seconds since the epoch, -1 on error (:272-322)
__device__ long long nanopore_time(const uint8_t *s, const uint8_t *end) {
    if (s + 20 > end) return -1;
    long long Y = dec_field(s, end, 4), M = dec_field(s + 5, end, 2), D = dec_field(s + 8, end, 2);
    long long h = dec_field(s + 11, end, 2), mi = dec_field(s + 14, end, 2), sec = dec_field(s + 17, end, 2);
    if ((Y | M | D | h | mi | sec) < 0 || s[4] != '-' || s[7] != '-' || s[10] != 'T' || s[13] != ':' ||
        s[16] != ':')
        return -1;
    const uint8_t *tz = s + 19;
    if (*tz == '.') {
        tz++;
        while (tz < end && *tz >= '0' && *tz <= '9') tz++;
    }
    if (tz >= end) return -1;
    if (*tz == '+' || *tz == '-') {
        long long oh = dec_field(tz + 1, end, 2), om = dec_field(tz + 4, end, 2);
        if ((oh | om) < 0 || tz[3] != ':') return -1;
        if (*tz == '+') { h += oh; mi += om; }
        else { h -= oh; mi -= om; }
    }
    else if (*tz != 'Z') return -1;
    if (Y < 1970 || M < 1 || M > 12) return -1;  // :252
    const int cum[12] = {0, 31, 59, 90, 120, 151, 181, 212, 243, 273, 304, 334};
    long long y = Y - 1900, yday = cum[M - 1] + D - 1;
    return sec + mi * 60 + h * 3600 + yday * 86400 + (y - 70) * 31536000 + ((y - 69) / 4) * 86400 -
           ((y - 1) / 100) * 86400 + ((y + 299) / 400) * 86400;
}

__device__ __forceinline__ const uint8_t *find_byte(const uint8_t *p, const uint8_t *end, uint8_t c) {
    while (p < end && *p != c) p++;
    return p < end ? p : nullptr;
}

__device__ int nano_header(const uint8_t *h, uint32_t n, int32_t *ch, long long *st) {
    const uint8_t *end = h + n, *p = find_byte(h, end, ' ');
    if (!p) return -1;
    p++;
    long long channel = -1, start = -1;
    while (p < end) {
        const uint8_t *eq = find_byte(p, end, '=');
        if (!eq) return -1;
        const uint8_t *val = eq + 1, *ve = find_byte(val, end, ' ');
        if (!ve) ve = end;
        uint32_t kl = (uint32_t)(eq - p);
        if (kl == 2 && p[0] == 'c' && p[1] == 'h')
            channel = (long long)(int32_t)dec_field(val, end, (uint32_t)(ve - val));
        else if (kl == 10 && p[0] == 's' && p[1] == 't' && p[2] == 'a' && p[3] == 'r' && p[4] == 't' &&
                 p[5] == '_' && p[6] == 't' && p[7] == 'i' && p[8] == 'm' && p[9] == 'e')
            start = nanopore_time(val, end);
        p = ve + 1;
    }
    if (channel == -1 || start == -1) return -1;
    *ch = (int32_t)channel;
    *st = start;
    return 0;
}

__device__ __forceinline__ uint32_t rd16(const uint8_t *p) { return p[0] | (uint32_t)p[1] << 8; }
__device__ __forceinline__ uint32_t rd32(const uint8_t *p) {
    return p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24;
}

// what went wrong in an aux block, as -(kind << 16 | detail): the reference's exceptions (:5080-5137, 5188-5195)
constexpr int NS_TAG_TRUNCATED = 1;   // ValueError("truncated tags")
constexpr int NS_TAG_ARRAY_TYPE = 2;  // ValueError("Invalid type for array %c"), detail = the type
constexpr int NS_TAG_UNKNOWN = 3;     // ValueError("Unknown tag type %c"), detail = the type
constexpr int NS_TAG_WRONG_TYPE = 4;  // RuntimeError("Wrong tag type for 'st' ..."), detail = tag (0 st, 1 du, 2 pi) << 8 | type
constexpr int NS_TAG_CH_TYPE = 5;     // ch with a non-integer type: the reference returns -1 without an exception set
__device__ __forceinline__ long long ns_tag_error(int kind, uint32_t detail = 0) { return -(long long)(kind << 16 | detail); }

__device__ long long aux_len(const uint8_t *t, uint64_t avail) {  // :5078-5140
    if (avail < 4) return ns_tag_error(NS_TAG_TRUNCATED);
    uint8_t ty = t[2];
    uint64_t head = 3, count = 1, width;
    bool is_array = false;
    if (ty == 'B') {
        if (avail < 8) return ns_tag_error(NS_TAG_TRUNCATED);
        ty = t[3];
        count = rd32(t + 4);
        head = 8;
        is_array = true;
    }
    switch (ty) {
        case 'A': case 'c': case 'C': width = 1; break;
        case 's': case 'S': width = 2; break;
        case 'i': case 'I': case 'f': width = 4; break;
        case 'Z': case 'H': {
            if (is_array) return ns_tag_error(NS_TAG_ARRAY_TYPE, ty);
            const uint8_t *z = find_byte(t + 3, t + avail, 0);
            if (!z) return ns_tag_error(NS_TAG_TRUNCATED);
            width = (uint64_t)(z - (t + 3)) + 1;
            break;
        }
        default: return ns_tag_error(NS_TAG_UNKNOWN, ty);
    }
    uint64_t len = head + count * width;
    return len > avail ? ns_tag_error(NS_TAG_TRUNCATED) : (long long)len;
}

// strtoull(s, &end, 16) consuming exactly 8 bytes (blanks, sign, 0x prefix accepted like libc)
__device__ bool hex8_like_strtoull(const uint8_t *s, uint64_t *out) {
    int i = 0;
    bool neg = false;
    while (i < 8 && (s[i] == ' ' || (s[i] >= 9 && s[i] <= 13))) i++;
    if (i < 8 && (s[i] == '+' || s[i] == '-')) neg = s[i++] == '-';
    int digits_at = i;
    if (i + 1 < 8 && s[i] == '0' && (s[i + 1] | 0x20) == 'x') {
        uint8_t c = i + 2 < 8 ? s[i + 2] : 0;
        bool hexd = (c >= '0' && c <= '9') || ((c | 0x20) >= 'a' && (c | 0x20) <= 'f');
        if (hexd) digits_at = i + 2;
    }
    uint64_t v = 0;
    int j = digits_at;
    for (; j < 8; j++) {
        uint8_t c = s[j];
        int d = (c >= '0' && c <= '9') ? c - '0' : ((c | 0x20) >= 'a' && (c | 0x20) <= 'f') ? (c | 0x20) - 'a' + 10 : -1;
        if (d < 0) break;
        v = v << 4 | (uint64_t)d;
    }
    if (j != 8 || j == digits_at) return false;
    *out = neg ? 0ULL - v : v;
    return true;
}
__device__ uint64_t uuid4_hash(const uint8_t *u) {  // :5153-5179
    if (u[8] != '-' || u[13] != '-' || u[14] != '4' || u[18] != '-' || u[23] != '-' || u[36] != 0) return 0;
    uint64_t a, b;
    if (!hex8_like_strtoull(u, &a) || !hex8_like_strtoull(u + 28, &b)) return 0;
    return a << 32 | (b & 0xffffffffULL);
}

// 0, or kind << 16 | detail of the first defect
__device__ int nano_tags(const uint8_t *t, uint64_t n, sq_nanoinfo *o, NsState *st, uint64_t record) {
    o->channel_id = -1;
    o->duration = 0.0f;
    o->start_time = 0;
    o->parent_id_hash = 0;
    while (n) {
        long long len = aux_len(t, n);
        if (len < 0) return (int)-len;
        uint8_t ty = t[2];
        if (t[0] == 'c' && t[1] == 'h') {
            const uint8_t *v = t + 3;
            switch (ty) {
                case 'c': o->channel_id = (int8_t)v[0]; break;
                case 'C': o->channel_id = v[0]; break;
                case 's': o->channel_id = (int16_t)rd16(v); break;
                case 'S': o->channel_id = (int32_t)rd16(v); break;
                case 'i': case 'I': o->channel_id = (int32_t)rd32(v); break;
                default: return NS_TAG_CH_TYPE << 16;
            }
        }
        else if (t[0] == 's' && t[1] == 't') {
            if (ty != 'Z') return NS_TAG_WRONG_TYPE << 16 | 0 << 8 | ty;
            o->start_time = nanopore_time(t + 3, t + len);
        }
        else if (t[0] == 'd' && t[1] == 'u') {
            if (ty != 'f') return NS_TAG_WRONG_TYPE << 16 | 1 << 8 | ty;
            o->duration = __uint_as_float(rd32(t + 3));
        }
        else if (t[0] == 'p' && t[1] == 'i') {
            if (ty != 'Z') return NS_TAG_WRONG_TYPE << 16 | 2 << 8 | ty;
            if (len - 4 != 36) {
                atomicAdd(&st->pi_warnings, 1ULL);
                atomicMin(&st->pi_first, (unsigned long long)record << 24 | (unsigned long long)min((long long)0xffffff, len - 4));
            }
            else o->parent_id_hash = uuid4_hash(t + 3);
        }
        t += len;
        n -= (uint64_t)len;
    }
    return 0;
}

__global__ void __launch_bounds__(NS_TPB)
k_ns_parse(BatchView bv, sq_nanoinfo *out, uint64_t base, NsState *st) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < bv.n; r += gridDim.x * blockDim.x) {
        // records behind an unparsable header are never looked at again (the module switches
        // itself off, :5302-5313): once one is known, later records need no work
        if (*(volatile unsigned long long *)&st->fail_idx < base + r) continue;
        sq_nanoinfo info;
        info.start_time = 0;
        info.duration = 0.0f;
        info.channel_id = 0;
        info.length = bv.seq_len[r];
        info.reserved = 0;
        info.parent_id_hash = 0;
        const uint32_t tl = bv.tags_len ? bv.tags_len[r] : 0;
        if (tl) {
            const int defect = nano_tags(bv.text + bv.tags_off[r], tl, &info, st, base + r);
            if (defect) atomicMin(&st->tag_err_idx, (unsigned long long)(base + r) << 24 | (unsigned)defect);
        }
        else {
            const uint32_t nl = bv.name_len ? bv.name_len[r] : bv.seq_off[r] - 1 - bv.name_off[r];
            int32_t ch;
            long long t0;
            if (nano_header(bv.text + bv.name_off[r], nl, &ch, &t0)) {
                atomicMin(&st->fail_idx, (unsigned long long)(base + r));
            }
            else {
                info.channel_id = ch;
                info.start_time = t0;
            }
        }
        info.cumulative_error_rate = bv.err_sum[r];
        out[base + r] = info;
    }
}

// the header that switched the module off, if it lies in this record array: copied aside for skipped_reason
__global__ void __launch_bounds__(256)
k_ns_capture_name(BatchView bv, uint64_t base, const NsState *st, uint8_t *name, uint32_t *name_len) {
    const unsigned long long fail = st->fail_idx;
    if (fail < base || fail >= base + bv.n) return;
    const uint32_t r = (uint32_t)(fail - base);
    const uint32_t nl = min(NS_NAME_CAP, bv.name_len ? bv.name_len[r] : bv.seq_off[r] - 1 - bv.name_off[r]);
    const uint8_t *src = bv.text + bv.name_off[r];
    for (uint32_t i = threadIdx.x; i < nl; i += blockDim.x) name[i] = src[i];
    if (threadIdx.x == 0) *name_len = nl;
}

// min/max start time over the first n kept records
__global__ void __launch_bounds__(256) k_ns_minmax(const sq_nanoinfo *infos, uint64_t n, NsState *st) {
    long long lo = 0x7fffffffffffffffLL, hi = 0;
    unsigned int nonpos = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        long long t = infos[i].start_time;
        lo = min(lo, t);
        hi = max(hi, t);
        nonpos |= t <= 0;
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        nonpos |= __shfl_xor_sync(0xffffffffu, nonpos, o);
    }
    if (lane_id() == 0) {
        atomicMin(&st->min_time, lo);
        atomicMax(&st->max_time, hi);
        if (nonpos) atomicOr(&st->nonpositive_time, 1u);
    }
}
// The reference treats min_time == 0 as "unset" (:5319), which makes the fold
// order-dependent once a timestamp <= 0 shows up.  Replay it in order (rare).
__global__ void k_ns_minmax_ordered(const sq_nanoinfo *infos, uint64_t n, NsState *st) {
    if (blockIdx.x || threadIdx.x) return;
    long long lo = 0, hi = 0;
    for (uint64_t i = 0; i < n; i++) {
        long long t = infos[i].start_time;
        if (t > hi) hi = t;
        if (lo == 0 || t < lo) lo = t;
    }
    st->min_time = lo;
    st->max_time = hi;
}

extern "C" int sq_nanostats_create(sq_ctx *ctx, sq_nanostats **out) {
    *out = nullptr;
    CUDA_TRY(cudaSetDevice(ctx->device));
    sq_nanostats *s = new sq_nanostats();
    s->ctx = ctx;
    int rc = sq_dalloc(ctx, (void **)&s->st, sizeof(NsState), true);
    if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&s->d_name, NS_NAME_CAP, false);
    if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&s->d_name_len, 4, true);
    if (rc == SQ_OK && cudaMallocHost((void **)&s->h_peek, sizeof(NsState)) != cudaSuccess) rc = SQ_E_CUDA;
    if (rc == SQ_OK && cudaEventCreateWithFlags(&s->peek_ev, cudaEventDisableTiming) != cudaSuccess) rc = SQ_E_CUDA;
    if (rc == SQ_OK) {
        NsState init;
        memset(&init, 0, sizeof(init));
        init.fail_idx = ~0ULL;
        init.tag_err_idx = ~0ULL;
        init.pi_first = ~0ULL;
        // same stream as the zero-fill above, so the two cannot swap
        rc = cudaMemcpyAsync(s->st, &init, sizeof(init), cudaMemcpyHostToDevice, sq_cur_stream(ctx)) == cudaSuccess &&
                     cudaStreamSynchronize(sq_cur_stream(ctx)) == cudaSuccess
                 ? SQ_OK : SQ_E_CUDA;
    }
    if (rc != SQ_OK) {
        sq_nanostats_destroy(s);
        return rc;
    }
    *out = s;
    return SQ_OK;
}

extern "C" void sq_nanostats_destroy(sq_nanostats *s) {
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    sq_dfree(s->ctx, s->infos);
    sq_dfree(s->ctx, s->st);
    sq_dfree(s->ctx, s->d_name);
    sq_dfree(s->ctx, s->d_name_len);
    if (s->peek_ev) {
        cudaEventSynchronize(s->peek_ev);
        cudaEventDestroy(s->peek_ev);
    }
    if (s->h_peek) cudaFreeHost(s->h_peek);
    sq_dfree(s->ctx, s->rp_channel);
    sq_dfree(s->ctx, s->rp_bases);
    sq_dfree(s->ctx, s->rp_error);
    delete s;
}

// the device has met a header it cannot parse: record index and the header itself to the host (synchronises)
static int ns_learn_skip(sq_nanostats *s, unsigned long long fail_idx) {
    sq_ctx *ctx = s->ctx;
    uint32_t len = 0;
    SQ_TRY(sq_memcpy_d2h(ctx, &len, s->d_name_len, 4));
    s->skipped_name.resize(len);
    if (len) SQ_TRY(sq_memcpy_d2h(ctx, s->skipped_name.data(), s->d_name, len));
    s->skipped = true;
    s->skipped_record = fail_idx;
    return SQ_OK;
}

// everything enqueued so far is done and the host knows whether the module switched itself off
static int ns_settle(sq_nanostats *s) {
    sq_ctx *ctx = s->ctx;
    NsState *h = (NsState *)((char *)ctx->h_scratch + 2560);
    CUDA_TRY(cudaMemcpyAsync(h, s->st, sizeof(NsState), cudaMemcpyDeviceToHost, sq_cur_stream(ctx)));
    CUDA_TRY(cudaStreamSynchronize(sq_cur_stream(ctx)));
    s->peek_pending = false;
    if (!s->skipped && h->fail_idx != ~0ULL) SQ_TRY(ns_learn_skip(s, h->fail_idx));
    return SQ_OK;
}

extern "C" int sq_nanostats_add(sq_nanostats *s, sq_batch *b) {
    sq_ctx *ctx = s->ctx;
    if (b->ctx != ctx) {
        sq_set_error("record array belongs to another context");
        return SQ_E_ARG;
    }
    if (s->skipped || b->n == 0) return SQ_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (s->peek_pending && cudaEventQuery(s->peek_ev) == cudaSuccess) {  // what an earlier add found out
        s->peek_pending = false;
        if (s->h_peek->fail_idx != ~0ULL) {
            SQ_TRY(ns_learn_skip(s, s->h_peek->fail_idx));
            return SQ_OK;
        }
    }
    if (s->n_added + b->n > s->cap) {
        uint64_t cap = s->cap ? s->cap : 16384;
        while (cap < s->n_added + b->n) cap *= 2;
        sq_nanoinfo *ni = nullptr;
        SQ_TRY(sq_dalloc(ctx, (void **)&ni, cap * sizeof(sq_nanoinfo), false));
        if (s->n_added)
            CUDA_TRY(cudaMemcpyAsync(ni, s->infos, s->n_added * sizeof(sq_nanoinfo), cudaMemcpyDeviceToDevice, sq_cur_stream(ctx)));
        sq_dfree(ctx, s->infos);
        s->infos = ni;
        s->cap = cap;
    }
    SQ_LAUNCH(ctx, k_ns_parse, sq_grid_for(ctx, b->n, NS_TPB, 16), NS_TPB, 0, b->view(), s->infos, s->n_added, s->st);
    // the name of the record that switches the module off is wanted for skipped_reason: copied aside on the
    // device while the array is alive; a copy of the state trails behind so that a later add can tell, without
    // waiting, that there is nothing left to do
    SQ_LAUNCH(ctx, k_ns_capture_name, 1, 256, 0, b->view(), s->n_added, s->st, s->d_name, s->d_name_len);
    if (!s->peek_pending) {
        CUDA_TRY(cudaMemcpyAsync(s->h_peek, s->st, sizeof(NsState), cudaMemcpyDeviceToHost, sq_cur_stream(ctx)));
        CUDA_TRY(cudaEventRecord(s->peek_ev, sq_cur_stream(ctx)));
        s->peek_pending = true;
    }
    s->n_added += b->n;
    return SQ_OK;
}

extern "C" int sq_nanostats_sync(sq_nanostats *s, sq_nanostats_info *info) {
    sq_ctx *ctx = s->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    memset(info, 0, sizeof(*info));
    SQ_TRY(ns_settle(s));
    const uint64_t n = s->skipped ? s->skipped_record : s->n_added;
    NsState *h = (NsState *)((char *)ctx->h_scratch + 2560);
    // recompute min/max over the kept prefix
    CUDA_TRY(cudaMemsetAsync(&s->st->max_time, 0, 8, sq_cur_stream(ctx)));
    long long big = 0x7fffffffffffffffLL;
    CUDA_TRY(cudaMemcpyAsync(&s->st->min_time, &big, 8, cudaMemcpyHostToDevice, sq_cur_stream(ctx)));
    CUDA_TRY(cudaMemsetAsync(&s->st->nonpositive_time, 0, 4, sq_cur_stream(ctx)));
    if (n) SQ_LAUNCH(ctx, k_ns_minmax, sq_grid_for(ctx, n, 256, 8), 256, 0, s->infos, n, s->st);
    CUDA_TRY(cudaMemcpyAsync(h, s->st, sizeof(NsState), cudaMemcpyDeviceToHost, sq_cur_stream(ctx)));
    CUDA_TRY(cudaStreamSynchronize(sq_cur_stream(ctx)));
    if (n && h->nonpositive_time) {
        SQ_LAUNCH(ctx, k_ns_minmax_ordered, 1, 32, 0, s->infos, n, s->st);
        CUDA_TRY(cudaMemcpyAsync(h, s->st, sizeof(NsState), cudaMemcpyDeviceToHost, sq_cur_stream(ctx)));
        CUDA_TRY(cudaStreamSynchronize(sq_cur_stream(ctx)));
    }
    info->number_of_reads = n;
    info->minimum_time = n ? h->min_time : 0;
    info->maximum_time = n ? h->max_time : 0;
    info->skipped = s->skipped;
    info->skipped_record = s->skipped_record;
    info->tag_error = h->tag_err_idx != ~0ULL ? (int32_t)(h->tag_err_idx >> 16 & 0xff) : 0;
    info->tag_error_record = h->tag_err_idx >> 24;
    info->tag_error_detail = (uint32_t)(h->tag_err_idx & 0xffff);
    info->pi_warnings = h->pi_warnings;
    info->pi_first_length = h->pi_first != ~0ULL ? (uint32_t)(h->pi_first & 0xffffff) : 0;
    return SQ_OK;
}

extern "C" int sq_nanostats_skipped_name(sq_nanostats *s, uint8_t *out, uint64_t cap, uint64_t *len) {
    uint64_t n = s->skipped_name.size() < cap ? s->skipped_name.size() : cap;
    if (n) memcpy(out, s->skipped_name.data(), n);
    *len = n;
    return SQ_OK;
}

extern "C" int sq_nanostats_read(sq_nanostats *s, sq_nanoinfo *out) {
    sq_ctx *ctx = s->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    SQ_TRY(ns_settle(s));
    const uint64_t n = s->skipped ? s->skipped_record : s->n_added;
    if (n) CUDA_TRY(cudaMemcpyAsync(out, s->infos, n * sizeof(sq_nanoinfo), cudaMemcpyDeviceToHost, sq_cur_stream(ctx)));
    CUDA_TRY(cudaStreamSynchronize(sq_cur_stream(ctx)));
    return SQ_OK;
}

// Sharded runs: NanoStats_add_meta appends one record per read, in read order (:5314-5322), and the first
// header that cannot be parsed switches the module off for everything behind it (:5302-5313).  Rank g
// holds the reads [first_record, first_record + n_added) of the stream: the reads in front of the first
// failing header of ANY rank are kept, and their records are gathered on every rank in rank (= read)
// order; sq_nanostats_sync recomputes the time range over the merged array.
extern "C" int sq_nanostats_allgather(sq_nanostats *s, sq_comm *c, uint64_t first_record) {
    sq_ctx *ctx = s->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    SQ_TRY(ns_settle(s));
    const int rank = sq_comm_rank(c), world = sq_comm_world(c);
    const uint64_t NONE = 1ULL << 62;
    uint64_t fail = s->skipped ? first_record + s->skipped_record : NONE;
    uint64_t F = fail;
    SQ_TRY(sq_comm_allreduce_host_u64(c, &F, 1, 2));  // min
    uint64_t kept = s->skipped ? s->skipped_record : s->n_added;
    if (F < first_record) kept = 0;
    else if (F - first_record < kept) kept = F - first_record;
    std::vector<uint64_t> counts((size_t)world);
    SQ_TRY(sq_comm_allgather_host(c, &kept, counts.data(), 8));
    uint64_t total = 0, my_off = 0;
    for (int g = 0; g < world; g++) {
        if (g == rank) my_off = total;
        total += counts[g];
    }
    sq_nanoinfo *merged = nullptr;
    SQ_TRY(sq_dalloc(ctx, (void **)&merged, (total ? total : 1) * sizeof(sq_nanoinfo), false));
    if (kept)
        CUDA_TRY(cudaMemcpyAsync(merged + my_off, s->infos, kept * sizeof(sq_nanoinfo), cudaMemcpyDeviceToDevice,
                                 sq_cur_stream(ctx)));
    uint64_t off = 0;
    SQ_TRY(sq_comm_group_start(c));
    for (int g = 0; g < world; g++) {
        SQ_TRY(sq_comm_bcast(c, merged + off, counts[g] * sizeof(sq_nanoinfo), g));
        off += counts[g];
    }
    SQ_TRY(sq_comm_group_end(c));
    // the header that switched the module off travels from the rank that met it
    uint64_t owner = fail == F && F != NONE ? (uint64_t)rank : NONE;
    SQ_TRY(sq_comm_allreduce_host_u64(c, &owner, 1, 2));
    if (F != NONE) {
        uint64_t len = s->skipped_name.size();
        SQ_TRY(sq_comm_bcast_host(c, &len, 8, (int)owner));
        s->skipped_name.resize(len);
        SQ_TRY(sq_comm_bcast_host(c, s->skipped_name.data(), len, (int)owner));
    }
    // pi warnings / tag errors: counters of all ranks
    NsState *h = (NsState *)((char *)ctx->h_scratch + 2560);
    CUDA_TRY(cudaMemcpyAsync(h, s->st, sizeof(NsState), cudaMemcpyDeviceToHost, sq_cur_stream(ctx)));
    CUDA_TRY(cudaStreamSynchronize(sq_cur_stream(ctx)));
    auto global_key = [&](unsigned long long key) {  // record index of this shard -> of the whole stream
        return key == ~0ULL ? NONE : ((first_record + (key >> 24)) << 24 | (key & 0xffffff));
    };
    uint64_t pi = h->pi_warnings, tag_err = global_key(h->tag_err_idx), pi_first = global_key(h->pi_first);
    SQ_TRY(sq_comm_allreduce_host_u64(c, &pi, 1, 0));
    SQ_TRY(sq_comm_allreduce_host_u64(c, &tag_err, 1, 2));
    SQ_TRY(sq_comm_allreduce_host_u64(c, &pi_first, 1, 2));
    h->pi_warnings = pi;
    h->tag_err_idx = tag_err == NONE ? ~0ULL : tag_err;
    h->pi_first = pi_first == NONE ? ~0ULL : pi_first;
    CUDA_TRY(cudaMemcpyAsync(&s->st->tag_err_idx, &h->tag_err_idx, 16, cudaMemcpyHostToDevice, sq_cur_stream(ctx)));
    CUDA_TRY(cudaMemcpyAsync(&s->st->pi_first, &h->pi_first, 8, cudaMemcpyHostToDevice, sq_cur_stream(ctx)));
    CUDA_TRY(cudaStreamSynchronize(sq_cur_stream(ctx)));
    sq_dfree(ctx, s->infos);
    s->infos = merged;
    s->cap = total ? total : 1;
    s->n_added = total;
    s->skipped = F != NONE;
    s->skipped_record = total;  // every merged record lies in front of the failing header
    return SQ_OK;
}
