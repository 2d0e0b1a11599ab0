// report.cu -- report aggregation on the tables and per-read records that already live on the device
// (reference report_modules.py: aggregate_count_matrix :307-322, SequenceLengthDistribution
// .from_base_count_tables :575-637, qc_metrics_modules :2537-2572, NanoStatsReport.from_nanostats :1952-2046).
//
// The reference's report pulls every table to Python (136 MB for 1 Mb reads) and loops over rows, and over
// every read for NanoStats.  Here only the aggregated numbers leave the device.
//
// QCMetrics (sq_qc_aggregate):
//   k_rp_ranges    a CTA per data range sums the rows [start, stop) of the base and phred tables
//   k_rp_lengths   one CTA: r[i] = reads longer than i (row sums of the base table), then everything the
//                  reference derives by walking the length histogram comes out of r by telescoping --
//                  reads of a length range = r[start] - r[stop]; reads up to a length = r[0] - r[length];
//                  bases in reads up to a length = sum(r[0..length)) - length * r[length] -- and each answer is
//                  the first length where a monotone quantity passes a threshold: a parallel minimum
// NanoStats (sq_nanostats_report):
//   k_nr_reads     a thread per read: time slot, quality class (11 thresholds on error / length, found on the
//                  host with the host's log10 so that the classes are the host's), translocation speed, distinct
//                  (slot, channel) pairs through a hash set; integer atomics only
//   per-channel error sums are floating point and the reference adds them in read order: stable sort of
//   (channel, read) + a thread per channel adding its reads front to back (k_nr_channels)
#include <cmath>

#include "common.cuh"
#include "modules.cuh"

// ---- QCMetrics tables --------------------------------------------------------------------------------
constexpr int RP_TPB = 256;

__global__ void __launch_bounds__(RP_TPB)
k_rp_ranges(const uint64_t *__restrict__ base, const uint64_t *__restrict__ phred, uint64_t max_len,
            const uint64_t *__restrict__ starts, const uint64_t *__restrict__ stops, uint64_t *__restrict__ base_out,
            uint64_t *__restrict__ phred_out, uint64_t *__restrict__ length_counts) {
    __shared__ unsigned long long acc[17];
    if (threadIdx.x < 17) acc[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t start = min(starts[blockIdx.x], max_len), stop = min(stops[blockIdx.x], max_len);
    uint64_t a[17];
#pragma unroll
    for (int c = 0; c < 17; c++) a[c] = 0;
    for (uint64_t r = start + threadIdx.x; r < stop; r += RP_TPB) {
#pragma unroll
        for (int c = 0; c < 5; c++) a[c] += base[r * 5 + c];
#pragma unroll
        for (int c = 0; c < 12; c++) a[5 + c] += phred[r * 12 + c];
    }
#pragma unroll
    for (int c = 0; c < 17; c++) {
        uint64_t v = a[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane_id() == 0 && v) atomicAdd(&acc[c], (unsigned long long)v);
    }
    __syncthreads();
    if (threadIdx.x < 5) base_out[blockIdx.x * 5 + threadIdx.x] = acc[threadIdx.x];
    else if (threadIdx.x < 17) phred_out[blockIdx.x * 12 + threadIdx.x - 5] = acc[threadIdx.x];
    if (threadIdx.x == 32) {  // reads with start < length <= stop
        auto longer_than = [&](uint64_t i) {
            uint64_t s = 0;
            if (i < max_len)
                for (int c = 0; c < 5; c++) s += base[i * 5 + c];
            return s;
        };
        length_counts[blockIdx.x] = start < stop ? longer_than(start) - longer_than(stop) : 0;
    }
}

constexpr int RL_TPB = 1024;
constexpr int RL_MAX_THRESHOLDS = 16;

struct RlOut {  // == sq_qc_length_summary
    unsigned long long total_bases, minimum_length, n50, n90;
    unsigned long long threshold_lengths[RL_MAX_THRESHOLDS];
};

__global__ void __launch_bounds__(RL_TPB)
k_rp_lengths(const uint64_t *__restrict__ base, uint64_t max_len, uint64_t total_sequences,
             const uint64_t *__restrict__ count_thresholds, uint32_t n_thresholds, uint64_t *__restrict__ r, RlOut *out) {
    __shared__ unsigned long long chunk_sum[RL_TPB];
    __shared__ unsigned long long total_s, found[3 + RL_MAX_THRESHOLDS];
    const uint64_t per = (max_len + RL_TPB - 1) / RL_TPB;
    const uint64_t a = min((uint64_t)threadIdx.x * per, max_len), b = min(a + per, max_len);
    unsigned long long s = 0;
    for (uint64_t i = a; i < b; i++) {
        unsigned long long v = 0;
        for (int c = 0; c < 5; c++) v += base[i * 5 + c];
        r[i] = v;
        s += v;
    }
    chunk_sum[threadIdx.x] = s;
    if (threadIdx.x < 3 + RL_MAX_THRESHOLDS) found[threadIdx.x] = ~0ULL;
    __syncthreads();
    if (threadIdx.x == 0) {  // exclusive prefix of the chunk sums
        unsigned long long run = 0;
        for (int t = 0; t < RL_TPB; t++) {
            const unsigned long long v = chunk_sum[t];
            chunk_sum[t] = run;
            run += v;
        }
        total_s = run;
    }
    __syncthreads();
    const unsigned long long total = total_s, half = total / 2;
    const unsigned long long tenth = (unsigned long long)((double)total * 0.1);  // int(total_bases * 0.1), :623
    const unsigned long long r0 = max_len ? r[0] : 0;
    unsigned long long prefix = chunk_sum[threadIdx.x];  // sum of r[0 .. a)
    unsigned long long mine[3 + RL_MAX_THRESHOLDS];
#pragma unroll
    for (int k = 0; k < 3 + RL_MAX_THRESHOLDS; k++) mine[k] = ~0ULL;
    auto at_length = [&](uint64_t length, unsigned long long pre, unsigned long long r_len) {
        const unsigned long long bases_upto = pre - length * r_len;  // bases in reads of at most `length`
        const unsigned long long reads_upto = r0 - r_len;            // reads of 1 .. length letters
        if (bases_upto >= half) mine[0] = min(mine[0], (unsigned long long)length);
        if (bases_upto >= tenth) mine[1] = min(mine[1], (unsigned long long)length);
        for (uint32_t k = 0; k < n_thresholds; k++)
            if (reads_upto > count_thresholds[k]) mine[3 + k] = min(mine[3 + k], (unsigned long long)length);
    };
    if (threadIdx.x == 0) at_length(0, 0, r0);  // the walk starts at length 0, where nothing has been counted yet
    for (uint64_t i = a; i < b; i++) {
        const unsigned long long ri = r[i];
        if (ri < total_sequences) mine[2] = min(mine[2], (unsigned long long)i);  // first position not every read reaches
        prefix += ri;
        at_length(i + 1, prefix, i + 1 < max_len ? r[i + 1] : 0);
    }
    for (int k = 0; k < 3 + (int)n_thresholds; k++)
        if (mine[k] != ~0ULL) atomicMin(&found[k], mine[k]);
    __syncthreads();
    if (threadIdx.x == 0) {
        out->total_bases = total;
        out->n50 = found[0];  // always found: every base is in a read of at most max_len
        out->n90 = found[1];
        out->minimum_length = found[2] == ~0ULL ? max_len : found[2];
        for (uint32_t k = 0; k < n_thresholds; k++) out->threshold_lengths[k] = found[3 + k] == ~0ULL ? 0 : found[3 + k];
    }
}

extern "C" int sq_qc_aggregate(sq_qc *m, const uint64_t *starts, const uint64_t *stops, uint64_t n_ranges,
                               uint64_t *base_matrix, uint64_t *phred_matrix, uint64_t *length_counts,
                               const uint64_t *count_thresholds, uint64_t n_thresholds, uint64_t total_sequences,
                               sq_qc_length_summary *summary) {
    static_assert(sizeof(RlOut) == sizeof(sq_qc_length_summary), "layout");
    sq_ctx *ctx = m->ctx;
    if (n_thresholds > RL_MAX_THRESHOLDS) {
        sq_set_error("at most %d count thresholds", RL_MAX_THRESHOLDS);
        return SQ_E_ARG;
    }
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t in_words = 2 * n_ranges + n_thresholds, out_words = n_ranges * 18 + sizeof(RlOut) / 8;
    uint64_t *d_in = nullptr, *d_out = nullptr, *d_r = nullptr;
    std::vector<uint64_t> h_in(in_words + 1), h_out(out_words);
    if (n_ranges) {
        memcpy(h_in.data(), starts, n_ranges * 8);
        memcpy(h_in.data() + n_ranges, stops, n_ranges * 8);
    }
    if (n_thresholds) memcpy(h_in.data() + 2 * n_ranges, count_thresholds, n_thresholds * 8);
    auto body = [&]() -> int {
        SQ_TRY(sq_dalloc(ctx, (void **)&d_in, (in_words + 1) * 8, false));
        SQ_TRY(sq_dalloc(ctx, (void **)&d_out, out_words * 8, true));
        SQ_TRY(sq_dalloc(ctx, (void **)&d_r, (m->max_len + 1) * 8, false));
        CUDA_TRY(cudaMemcpyAsync(d_in, h_in.data(), (in_words + 1) * 8, cudaMemcpyHostToDevice, st));
        uint64_t *d_base = d_out, *d_phred = d_out + n_ranges * 5, *d_len = d_out + n_ranges * 17;
        RlOut *d_sum = (RlOut *)(d_out + n_ranges * 18);
        if (n_ranges)
            SQ_LAUNCH(ctx, k_rp_ranges, (unsigned)n_ranges, RP_TPB, 0, m->base, m->phred, m->max_len, d_in, d_in + n_ranges,
                      d_base, d_phred, d_len);
        SQ_LAUNCH(ctx, k_rp_lengths, 1, RL_TPB, 0, m->base, m->max_len, total_sequences, d_in + 2 * n_ranges,
                  (uint32_t)n_thresholds, d_r, d_sum);
        CUDA_TRY(cudaMemcpyAsync(h_out.data(), d_out, out_words * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        return SQ_OK;
    };
    const int rc = body();
    sq_dfree(ctx, d_in);
    sq_dfree(ctx, d_out);
    sq_dfree(ctx, d_r);
    if (rc != SQ_OK) return rc;
    if (n_ranges) {
        memcpy(base_matrix, h_out.data(), n_ranges * 5 * 8);
        memcpy(phred_matrix, h_out.data() + n_ranges * 5, n_ranges * 12 * 8);
        memcpy(length_counts, h_out.data() + n_ranges * 17, n_ranges * 8);
    }
    memcpy(summary, h_out.data() + n_ranges * 18, sizeof(RlOut));
    return SQ_OK;
}

// ---- NanoStats ---------------------------------------------------------------------------------------
constexpr int NR_TPB = 256;

struct NrParams {
    long long run_start, interval;
    unsigned long long n_slots;
    double class_edge[11];  // error / length <= class_edge[k - 1]  <=>  quality class >= k
};

struct NrCounters {
    unsigned long long reads_with_parent, first_error;  // first_error = record << 2 | kind, min
    unsigned int n_channels, pad;
};

__global__ void __launch_bounds__(NR_TPB)
k_nr_reads(const sq_nanoinfo *__restrict__ infos, uint64_t n, NrParams p, unsigned long long *__restrict__ time_bases,
           unsigned long long *__restrict__ time_reads, unsigned long long *__restrict__ time_active,
           unsigned long long *__restrict__ time_quals, unsigned long long *__restrict__ speeds,
           unsigned long long *__restrict__ pair_set, uint64_t set_mask, uint32_t *__restrict__ keys,
           uint32_t *__restrict__ vals, NrCounters *cnt) {
    unsigned int parents = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * NR_TPB + threadIdx.x; i < n; i += (uint64_t)gridDim.x * NR_TPB) {
        const sq_nanoinfo ni = infos[i];
        parents += ni.parent_id_hash != 0;
        keys[i] = (uint32_t)ni.channel_id ^ 0x80000000u;  // unsigned order == signed order
        vals[i] = (uint32_t)i;
        int cls = 0;
        if (ni.length) {
            const double ratio = ni.cumulative_error_rate / (double)ni.length;
#pragma unroll
            for (int k = 0; k < 11; k++) cls += ratio <= p.class_edge[k];
        }
        if (ni.start_time) {
            const long long rel = ni.start_time - p.run_start;
            long long slot = rel / p.interval;
            if (rel % p.interval < 0) slot--;                       // Python's floor division
            if (slot < 0) slot += (long long)p.n_slots;             // ... and its negative indices
            if (slot < 0 || slot >= (long long)p.n_slots) atomicMin(&cnt->first_error, (unsigned long long)i << 2 | 3);
            else {
                atomicAdd(&time_bases[slot], (unsigned long long)ni.length);
                atomicAdd(&time_reads[slot], 1ULL);
                atomicAdd(&time_quals[slot * 12 + cls], 1ULL);
                const unsigned long long key = ((unsigned long long)slot << 32 | (uint32_t)ni.channel_id) + 1;
                uint64_t h = fmix64(key) & set_mask;
                for (;;) {
                    const unsigned long long old = atomicCAS(&pair_set[h], 0ULL, key);
                    if (old == 0) {
                        atomicAdd(&time_active[slot], 1ULL);
                        break;
                    }
                    if (old == key) break;
                    h = (h + 1) & set_mask;
                }
            }
        }
        if (ni.duration != 0.0f) {  // NaN counts as set, like Python's truth value of a float
            const double v = rint((double)ni.length / (double)ni.duration);  // round(): half to even
            if (isnan(v)) atomicMin(&cnt->first_error, (unsigned long long)i << 2 | 2);
            else if (isinf(v)) atomicMin(&cnt->first_error, (unsigned long long)i << 2 | 1);
            else {
                long long bin = (long long)floor(fmin(v, 800.0) / 10.0);
                if (bin < 0) bin += 81;
                if (bin < 0) atomicMin(&cnt->first_error, (unsigned long long)i << 2 | 3);
                else atomicAdd(&speeds[bin], 1ULL);
            }
        }
    }
    parents = warp_sum_u32(parents);
    if (lane_id() == 0 && parents) atomicAdd(&cnt->reads_with_parent, (unsigned long long)parents);
}

__global__ void __launch_bounds__(NR_TPB)
k_nr_heads(const uint32_t *__restrict__ keys, uint32_t n, uint32_t *__restrict__ head) {
    for (uint32_t i = blockIdx.x * NR_TPB + threadIdx.x; i < n; i += gridDim.x * NR_TPB) head[i] = i == 0 || keys[i] != keys[i - 1];
}

// a thread per channel: its reads front to back, one rounded addition each (:2012-2013)
__global__ void __launch_bounds__(NR_TPB)
k_nr_channels(const sq_nanoinfo *__restrict__ infos, const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals,
              const uint32_t *__restrict__ head, const uint32_t *__restrict__ head_rank, uint32_t n, int32_t *__restrict__ channel,
              unsigned long long *__restrict__ bases, double *__restrict__ error) {
    for (uint32_t i = blockIdx.x * NR_TPB + threadIdx.x; i < n; i += gridDim.x * NR_TPB) {
        if (!head[i]) continue;
        const uint32_t key = keys[i];
        unsigned long long b = 0;
        double e = 0.0;
        for (uint32_t j = i; j < n && keys[j] == key; j++) {
            const sq_nanoinfo ni = infos[vals[j]];
            b += ni.length;
            e += ni.cumulative_error_rate;
        }
        const uint32_t c = head_rank[i];
        channel[c] = (int32_t)(key ^ 0x80000000u);
        bases[c] = b;
        error[c] = e;
    }
}

// quality class edges with the HOST's log10: class k (1..11) <=> round(-10 * log10(ratio)) >= 4k (:1998-2002),
// the largest such ratio found by bisection on the bit pattern (log10 is monotone)
static void nr_class_edges(double *edge) {
    for (int k = 1; k <= 11; k++) {
        auto in_class = [&](double ratio) { return nearbyint(-10.0 * log10(ratio)) >= 4.0 * k; };
        uint64_t lo = 1, hi = 0x7ff0000000000000ULL;  // smallest subnormal (in class) .. +inf (not)
        auto as_double = [](uint64_t u) {
            double d;
            memcpy(&d, &u, 8);
            return d;
        };
        while (hi - lo > 1) {
            const uint64_t mid = lo + (hi - lo) / 2;
            if (in_class(as_double(mid))) lo = mid;
            else hi = mid;
        }
        edge[k - 1] = as_double(lo);
    }
}

// NanoStatsReport.from_nanostats over the device-resident records.  The caller derives run_start / interval /
// n_slots from minimum_time / maximum_time exactly like :1968-1980.  Returns the number of channels; the
// per-channel arrays stay with the collector until sq_nanostats_report_channels fetches them.
extern "C" int sq_nanostats_report(sq_nanostats *s, int64_t run_start, int64_t interval, uint64_t n_slots,
                                   uint64_t *time_bases, uint64_t *time_reads, uint64_t *time_active_channels,
                                   uint64_t *time_qualities, uint64_t *translocation_speeds, uint64_t *reads_with_parent,
                                   uint64_t *n_channels, sq_nano_report_error *error) {
    sq_ctx *ctx = s->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint64_t n = s->skipped ? s->skipped_record : s->n_added;
    memset(error, 0, sizeof(*error));
    *reads_with_parent = *n_channels = 0;
    if (interval < 1 || n_slots < 1 || n > 0xfffffff0ULL) {
        sq_set_error("sq_nanostats_report: interval %lld, %llu slots, %llu reads", (long long)interval,
                     (unsigned long long)n_slots, (unsigned long long)n);
        return SQ_E_ARG;
    }
    NrParams p;
    p.run_start = run_start;
    p.interval = interval;
    p.n_slots = n_slots;
    static double edges[11];
    static bool have_edges = false;
    if (!have_edges) {
        nr_class_edges(edges);
        have_edges = true;
    }
    memcpy(p.class_edge, edges, sizeof(edges));
    uint64_t set_cap = 1024;
    while (set_cap < 2 * n) set_cap *= 2;
    const size_t slot_words = n_slots * 15 + 81;
    unsigned long long *d_slots = nullptr, *d_set = nullptr;
    uint32_t *keys = nullptr, *vals = nullptr, *tk = nullptr, *tv = nullptr, *head = nullptr, *head_rank = nullptr;
    NrCounters *d_cnt = nullptr;
    sq_dfree(ctx, s->rp_channel);
    sq_dfree(ctx, s->rp_bases);
    sq_dfree(ctx, s->rp_error);
    s->rp_channel = nullptr, s->rp_bases = nullptr, s->rp_error = nullptr, s->rp_n = 0;
    std::vector<uint64_t> h_slots(slot_words);
    NrCounters h_cnt;
    auto body = [&]() -> int {
        SQ_TRY(sq_dalloc(ctx, (void **)&d_slots, slot_words * 8, true));
        SQ_TRY(sq_dalloc(ctx, (void **)&d_set, set_cap * 8, true));
        SQ_TRY(sq_dalloc(ctx, (void **)&d_cnt, sizeof(NrCounters), true));
        for (uint32_t **q : {&keys, &vals, &tk, &tv, &head, &head_rank}) SQ_TRY(sq_dalloc(ctx, (void **)q, (n + 1) * 4, false));
        CUDA_TRY(cudaMemsetAsync(&d_cnt->first_error, 0xff, 8, st));
        if (n) {
            SQ_LAUNCH(ctx, k_nr_reads, sq_grid_for(ctx, n, NR_TPB, 8), NR_TPB, 0, s->infos, n, p, d_slots, d_slots + n_slots,
                      d_slots + 2 * n_slots, d_slots + 3 * n_slots, d_slots + 15 * n_slots, d_set, set_cap - 1, keys, vals, d_cnt);
            SQ_TRY(sq_radix_sort_pairs(ctx, keys, vals, tk, tv, (uint32_t)n, 32));
            SQ_LAUNCH(ctx, k_nr_heads, sq_grid_for(ctx, n, NR_TPB, 8), NR_TPB, 0, keys, (uint32_t)n, head);
            SQ_TRY(sq_scan_exclusive_u32(ctx, head, head_rank, (uint32_t)n, &d_cnt->n_channels));
        }
        CUDA_TRY(cudaMemcpyAsync(&h_cnt, d_cnt, sizeof(NrCounters), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(h_slots.data(), d_slots, slot_words * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        s->rp_n = n ? h_cnt.n_channels : 0;
        if (s->rp_n) {
            SQ_TRY(sq_dalloc(ctx, (void **)&s->rp_channel, s->rp_n * 4, false));
            SQ_TRY(sq_dalloc(ctx, (void **)&s->rp_bases, s->rp_n * 8, false));
            SQ_TRY(sq_dalloc(ctx, (void **)&s->rp_error, s->rp_n * 8, false));
            SQ_LAUNCH(ctx, k_nr_channels, sq_grid_for(ctx, n, NR_TPB, 8), NR_TPB, 0, s->infos, keys, vals, head, head_rank,
                      (uint32_t)n, s->rp_channel, s->rp_bases, s->rp_error);
        }
        return SQ_OK;
    };
    const int rc = body();
    for (void *q : {(void *)d_slots, (void *)d_set, (void *)d_cnt, (void *)keys, (void *)vals, (void *)tk, (void *)tv, (void *)head,
                    (void *)head_rank})
        sq_dfree(ctx, q);
    if (rc != SQ_OK) return rc;
    memcpy(time_bases, h_slots.data(), n_slots * 8);
    memcpy(time_reads, h_slots.data() + n_slots, n_slots * 8);
    memcpy(time_active_channels, h_slots.data() + 2 * n_slots, n_slots * 8);
    memcpy(time_qualities, h_slots.data() + 3 * n_slots, n_slots * 12 * 8);
    memcpy(translocation_speeds, h_slots.data() + 15 * n_slots, 81 * 8);
    *reads_with_parent = h_cnt.reads_with_parent;
    *n_channels = s->rp_n;
    if (h_cnt.first_error != ~0ULL) {
        error->kind = (int32_t)(h_cnt.first_error & 3);
        error->record = h_cnt.first_error >> 2;
    }
    return SQ_OK;
}

extern "C" int sq_nanostats_report_channels(sq_nanostats *s, int32_t *channel_ids, uint64_t *bases, double *cumulative_error,
                                            uint64_t cap) {
    sq_ctx *ctx = s->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const uint64_t n = s->rp_n;
    if (cap < n) {
        sq_set_error("room for %llu channels, %llu needed", (unsigned long long)cap, (unsigned long long)n);
        return SQ_E_ARG;
    }
    if (n) {
        CUDA_TRY(cudaMemcpyAsync(channel_ids, s->rp_channel, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(bases, s->rp_bases, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(cumulative_error, s->rp_error, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return SQ_OK;
}
