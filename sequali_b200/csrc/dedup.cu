// dedup.cu -- DedupEstimator (reference _qcmodule.c:4383-4517).
//
// The reference is a sequential open-addressing table with content-based
// sampling: a fingerprint hash h is looked at only if its low `m` bits are
// zero; when the table holds `max` entries the next sampled add raises m,
// rebuilds the table and then inserts at the index computed with the OLD m.
// Which entry sits in which slot is decided by arrival order, and the getter
// returns counts in slot order.  To reproduce that exactly in parallel the
// table is built by *priority* linear probing (Shun & Blelloch, "Phase-
// concurrent hash tables for determinism", SPAA 2014): inserting keys with
// atomicMin on a priority word yields the layout of a sequential insertion in
// priority order.  Priorities are the global record index of a key's first
// occurrence (for rebuilds: the old slot index), so the device table is, slot
// for slot, the table the reference would hold.
//
// Per record array:
//   k_dd_hash       fingerprint -> MurmurHash3 (second half), one thread/read
//   k_dd_classify   sampled hashes probe the table: existing slot, or "new"
//                   -> collected (first index, count) in a scratch table
//   k_dd_flags + scan + k_dd_trigger*   find, in record order, the add that
//                   meets a full table (the escalation point), if any
//   k_dd_apply / k_dd_insert_new / k_dd_place   commit counts, place new keys
//   k_dd_rebuild_* / k_dd_insert_one    escalation (rare: ~log2(reads/max))
#include <math.h>

#include "modules.cuh"





// ---------------------------------------------------------------------------
// hashing (reference :4463-4517)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(DD_TPB)
k_dd_hash(BatchView bv, uint64_t front_len, uint64_t back_len, uint64_t front_off, uint64_t back_off,
          uint64_t *__restrict__ hashes) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < bv.n; r += gridDim.x * blockDim.x) {
        const uint8_t *s = bv.text + bv.seq_off[r];
        const uint64_t L = bv.seq_len[r], fl = front_len + back_len;
        uint64_t h;
        if (L <= fl) {
            h = murmur3_h2([&](uint64_t i) { return s[i]; }, L, 0);
        }
        else {
            const uint64_t rem = L - fl;
            const uint64_t fo = min(rem / 2, front_off), bo = min(rem / 2, back_off);
            const uint8_t *f = s + fo, *b = s + L - (bo + back_len);
            h = murmur3_h2([&](uint64_t i) { return i < front_len ? f[i] : b[i - front_len]; }, fl, L >> 6);
        }
        hashes[r] = h;
    }
}

// Pair fingerprint.  When a read is shorter than the configured length the
// reference hashes stale bytes of its scratch buffer (:4503-4516); that only
// concerns reads shorter than 8 bases and is serialised through `stale` here:
// such pairs are rare enough to be handled by one thread in record order.
__global__ void __launch_bounds__(DD_TPB)
k_dd_hash_pair(BatchView b1, BatchView b2, uint64_t front_len, uint64_t back_len, uint64_t front_off,
               uint64_t back_off, uint64_t *__restrict__ hashes, uint32_t *__restrict__ short_flag,
               uint32_t *range) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < b1.n; r += gridDim.x * blockDim.x) {
        const uint8_t *s1 = b1.text + b1.seq_off[r], *s2 = b2.text + b2.seq_off[r];
        const uint64_t L1 = b1.seq_len[r], L2 = b2.seq_len[r], fl = front_len + back_len;
        const uint64_t f = min(front_len, L1), b = min(back_len, L2);
        if (f + b < fl) {  // would read stale scratch bytes: defer to the ordered kernel
            short_flag[r] = 1;
            hashes[r] = 0;
            atomicMin(range, r);
            atomicMax(range + 1, r);
            continue;
        }
        short_flag[r] = 0;
        const uint64_t fo = min(front_off, L1 - f), bo = min(back_off, L2 - b);
        const uint8_t *pf = s1 + fo, *pb = s2 + bo;
        hashes[r] = murmur3_h2([&](uint64_t i) { return i < f ? pf[i] : pb[i - f]; }, fl, (L1 + L2) >> 6);
    }
}

// The scratch buffer of the reference (fingerprint_store, :4500-4516) as it stands after record r: a pair
// that is long enough overwrites all of it, a short pair only its first f and the b bytes behind them.
__device__ __forceinline__ void dd_pair_into_stale(const BatchView &b1, const BatchView &b2, uint32_t r,
                                                   uint64_t front_len, uint64_t back_len, uint64_t front_off,
                                                   uint64_t back_off, uint8_t *stale, uint64_t *f_out) {
    const uint8_t *s1 = b1.text + b1.seq_off[r], *s2 = b2.text + b2.seq_off[r];
    const uint64_t L1 = b1.seq_len[r], L2 = b2.seq_len[r];
    const uint64_t f = min(front_len, L1), b = min(back_len, L2);
    const uint64_t fo = min(front_off, L1 - f), bo = min(back_off, L2 - b);
    for (uint64_t i = 0; i < f; i++) stale[i] = s1[fo + i];
    for (uint64_t i = 0; i < b; i++) stale[f + i] = s2[bo + i];
    *f_out = f;
}

// One CTA walks the short-pair flags between the first and the last flagged record (`range`, kept by
// k_dd_hash_pair), 1024 records per step; thread 0 re-hashes the flagged pairs in record order with the
// scratch buffer exactly as the reference leaves it: the record in front of a short pair (or the last
// pair of an earlier batch, kept in `stale`) supplies the bytes the short pair does not overwrite.
// The last pair of the batch is left in `stale` for the next batch.
__global__ void __launch_bounds__(1024)
k_dd_hash_pair_ordered(BatchView b1, BatchView b2, uint64_t front_len, uint64_t back_len, uint64_t front_off,
                       uint64_t back_off, uint64_t *hashes, const uint32_t *__restrict__ short_flag,
                       uint32_t *range, uint8_t *stale) {
    __shared__ uint32_t s_flag[1024];
    const uint64_t fl = front_len + back_len;
    const uint32_t n = b1.n, tid = threadIdx.x;
    const uint32_t lo = range[0], hi = range[1];
    uint32_t applied = 0xFFFFFFFFu;  // thread 0: last record whose bytes are in `stale`
    if (lo <= hi) {
        for (uint32_t base = lo & ~1023u; base <= hi; base += 1024) {
            const uint32_t r = base + tid;
            const uint32_t fl_here = r < n ? short_flag[r] : 0u;
            s_flag[tid] = fl_here;
            if (__syncthreads_or((int)fl_here) && tid == 0) {
                for (uint32_t i = 0; i < 1024; i++) {
                    if (!s_flag[i]) continue;
                    const uint32_t rr = base + i;
                    uint64_t f;
                    if (rr > 0 && applied != rr - 1)
                        dd_pair_into_stale(b1, b2, rr - 1, front_len, back_len, front_off, back_off, stale, &f);
                    dd_pair_into_stale(b1, b2, rr, front_len, back_len, front_off, back_off, stale, &f);
                    applied = rr;
                    hashes[rr] = murmur3_h2([&](uint64_t k) { return stale[k]; }, fl,
                                            ((uint64_t)b1.seq_len[rr] + b2.seq_len[rr]) >> 6);
                }
            }
            __syncthreads();
        }
    }
    if (tid == 0) {
        if (n && applied != n - 1) {
            uint64_t f;
            dd_pair_into_stale(b1, b2, n - 1, front_len, back_len, front_off, back_off, stale, &f);
        }
        range[0] = 0xFFFFFFFFu;  // ready for the next batch
        range[1] = 0;
    }
}

// ---------------------------------------------------------------------------
// classification against the table as it is at the start of a segment
// ---------------------------------------------------------------------------
struct Scratch {  // distinct new keys of a segment
    uint64_t *key;    // DD_EMPTY = free
    uint32_t *first;  // smallest record index holding the key
    uint32_t *cnt;    // occurrences in the segment
    uint32_t mask;
};

__device__ __forceinline__ uint32_t scratch_slot(const Scratch &S, uint64_t h, bool insert) {
    uint32_t i = (uint32_t)(h ^ (h >> 32)) & S.mask;
    for (;;) {
        uint64_t k = S.key[i];
        if (k == h) return i;
        if (k == DD_EMPTY) {
            if (!insert) return 0xFFFFFFFFu;
            uint64_t old = atomicCAS((unsigned long long *)&S.key[i], DD_EMPTY, h);
            if (old == DD_EMPTY || old == h) return i;
        }
        i = (i + 1) & S.mask;
    }
}

__global__ void __launch_bounds__(DD_TPB)
k_dd_classify(const uint64_t *__restrict__ hashes, uint32_t lo, uint32_t hi, uint32_t m, DdTable T,
              uint64_t tmask, uint32_t *__restrict__ cls, Scratch S) {
    for (uint32_t r = lo + blockIdx.x * blockDim.x + threadIdx.x; r < hi; r += gridDim.x * blockDim.x) {
        const uint64_t h = hashes[r];
        if (h & ((1ULL << m) - 1)) {
            cls[r] = DD_NOPASS;
            continue;
        }
        uint64_t i = (h >> m) & tmask;
        uint32_t c = DD_NEW;
        for (;;) {
            if (T.count[i] == 0) break;
            if (T.hash[i] == h) {
                c = (uint32_t)i;
                break;
            }
            i = (i + 1) & tmask;
        }
        cls[r] = c;
        if (c == DD_NEW) {
            uint32_t s = scratch_slot(S, h, true);
            atomicMin(&S.first[s], r);
            atomicAdd(&S.cnt[s], 1u);
        }
    }
}

// flag[r] = 1 when record r is the first occurrence of a new key
__global__ void __launch_bounds__(DD_TPB)
k_dd_flags(const uint64_t *__restrict__ hashes, const uint32_t *__restrict__ cls, uint32_t lo, uint32_t hi,
           Scratch S, uint32_t *__restrict__ flag, DdCounters *cnt) {
    uint32_t local = 0;
    for (uint32_t r = lo + blockIdx.x * blockDim.x + threadIdx.x; r < hi; r += gridDim.x * blockDim.x) {
        uint32_t f = 0;
        if (cls[r] == DD_NEW) {
            uint32_t s = scratch_slot(S, hashes[r], false);
            f = S.first[s] == r;
        }
        flag[r - lo] = f;
        local += f;
    }
    local = warp_sum_u32(local);
    if (lane_id() == 0 && local) atomicAdd(&cnt->n_new, local);
}

// r_full = record holding the K-th first occurrence (rank is the exclusive scan of flag)
__global__ void __launch_bounds__(DD_TPB)
k_dd_trigger_full(const uint32_t *__restrict__ flag, const uint32_t *__restrict__ rank, uint32_t lo,
                  uint32_t hi, uint32_t K, DdCounters *cnt) {
    for (uint32_t r = lo + blockIdx.x * blockDim.x + threadIdx.x; r < hi; r += gridDim.x * blockDim.x)
        if (flag[r - lo] && rank[r - lo] + 1 == K) cnt->r_full = r;
}
// r_star = first sampled record after r_full (or the first sampled record when the table is already full)
__global__ void __launch_bounds__(DD_TPB)
k_dd_trigger_star(const uint32_t *__restrict__ cls, uint32_t lo, uint32_t hi, int already_full,
                  DdCounters *cnt) {
    const unsigned long long r_full = cnt->r_full;
    if (!already_full && r_full == ~0ULL) return;
    unsigned long long best = ~0ULL;
    for (uint32_t r = lo + blockIdx.x * blockDim.x + threadIdx.x; r < hi; r += gridDim.x * blockDim.x)
        if (cls[r] != DD_NOPASS && (already_full || r > r_full)) {
            best = r;
            break;  // indices grow along the stride: the first hit is this thread's minimum
        }
    if (best != ~0ULL) atomicMin(&cnt->r_star, best);
}

// ---------------------------------------------------------------------------
// commit
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(DD_TPB)
k_dd_apply_existing(const uint32_t *__restrict__ cls, uint32_t lo, uint32_t hi, DdTable T) {
    for (uint32_t r = lo + blockIdx.x * blockDim.x + threadIdx.x; r < hi; r += gridDim.x * blockDim.x) {
        uint32_t c = cls[r];
        if (c < DD_NOPASS) atomicAdd(&T.count[c], 1u);
    }
}

// priority linear probing: the smaller word wins the slot, the loser moves on
__device__ __forceinline__ void prio_insert(uint64_t *prio, uint64_t tmask, uint64_t home, uint64_t word) {
    uint64_t i = home;
    for (;;) {
        uint64_t old = atomicMin((unsigned long long *)&prio[i], (unsigned long long)word);
        if (old == DD_EMPTY) return;
        if (old > word) word = old;  // displaced a later arrival: carry it onwards
        i = (i + 1) & tmask;
    }
}
__device__ __forceinline__ uint64_t prio_locate(const uint64_t *prio, uint64_t tmask, uint64_t home, uint64_t word) {
    uint64_t i = home;
    while (prio[i] != word) i = (i + 1) & tmask;
    return i;
}

__global__ void __launch_bounds__(DD_TPB)
k_dd_insert_new(const uint64_t *__restrict__ hashes, const uint32_t *__restrict__ flag, uint32_t lo,
                uint32_t hi, uint32_t m, DdTable T, uint64_t tmask, uint64_t prio_base) {
    for (uint32_t r = lo + blockIdx.x * blockDim.x + threadIdx.x; r < hi; r += gridDim.x * blockDim.x)
        if (flag[r - lo]) prio_insert(T.prio, tmask, (hashes[r] >> m) & tmask, prio_base + r);
}
__global__ void __launch_bounds__(DD_TPB)
k_dd_place_new(const uint64_t *__restrict__ hashes, const uint32_t *__restrict__ flag, uint32_t lo,
               uint32_t hi, uint32_t m, DdTable T, uint64_t tmask, uint64_t prio_base, Scratch S) {
    for (uint32_t r = lo + blockIdx.x * blockDim.x + threadIdx.x; r < hi; r += gridDim.x * blockDim.x)
        if (flag[r - lo]) {
            const uint64_t h = hashes[r];
            uint64_t pos = prio_locate(T.prio, tmask, (h >> m) & tmask, prio_base + r);
            T.hash[pos] = h;
            T.count[pos] = S.cnt[scratch_slot(S, h, false)];
        }
}

// rebuild (reference :4383-4423): survivors re-enter in old slot order, no equality test
__global__ void __launch_bounds__(DD_TPB)
k_dd_rebuild_insert(DdTable oldT, DdTable newT, uint64_t size, uint32_t new_bits, DdCounters *cnt) {
    const uint64_t tmask = size - 1, drop = (1ULL << new_bits) - 1;
    uint32_t local = 0;
    for (uint64_t i = blockIdx.x * blockDim.x + threadIdx.x; i < size; i += (uint64_t)gridDim.x * blockDim.x) {
        if (oldT.count[i] == 0) continue;
        const uint64_t h = oldT.hash[i];
        if (h & drop) continue;
        prio_insert(newT.prio, tmask, (h >> new_bits) & tmask, i + 1);
        local++;
    }
    local = warp_sum_u32(local);
    if (lane_id() == 0 && local) atomicAdd(&cnt->kept, local);
}
__global__ void __launch_bounds__(DD_TPB)
k_dd_rebuild_place(DdTable oldT, DdTable newT, uint64_t size, uint32_t new_bits) {
    const uint64_t tmask = size - 1, drop = (1ULL << new_bits) - 1;
    for (uint64_t i = blockIdx.x * blockDim.x + threadIdx.x; i < size; i += (uint64_t)gridDim.x * blockDim.x) {
        if (oldT.count[i] == 0) continue;
        const uint64_t h = oldT.hash[i];
        if (h & drop) continue;
        uint64_t pos = prio_locate(newT.prio, tmask, (h >> new_bits) & tmask, i + 1);
        newT.hash[pos] = h;
        newT.count[pos] = oldT.count[i];
    }
}
// the add that triggered the escalation: probes from the index of the OLD bit count (:4442)
__global__ void k_dd_insert_one(const uint64_t *hashes, uint32_t r, uint32_t old_bits, DdTable T, uint64_t tmask,
                                uint64_t prio_base, DdCounters *cnt) {
    if (blockIdx.x || threadIdx.x) return;
    const uint64_t h = hashes[r];
    uint64_t i = (h >> old_bits) & tmask;
    for (;;) {
        if (T.count[i] == 0) {
            T.hash[i] = h;
            T.count[i] = 1;
            T.prio[i] = prio_base + r;
            cnt->inserted_one = 1;
            return;
        }
        if (T.hash[i] == h) {
            T.count[i] += 1;
            cnt->inserted_one = 0;
            return;
        }
        i = (i + 1) & tmask;
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static int table_alloc(sq_ctx *ctx, DdTable *t, uint64_t size) {
    SQ_TRY(sq_dalloc(ctx, (void **)&t->hash, size * 8, true));
    SQ_TRY(sq_dalloc(ctx, (void **)&t->count, size * 4, true));
    SQ_TRY(sq_dalloc(ctx, (void **)&t->prio, size * 8, false));
    CUDA_TRY(cudaMemsetAsync(t->prio, 0xFF, size * 8, sq_cur_stream(ctx)));
    return SQ_OK;
}
static int table_clear(sq_ctx *ctx, DdTable *t, uint64_t size) {
    CUDA_TRY(cudaMemsetAsync(t->hash, 0, size * 8, sq_cur_stream(ctx)));
    CUDA_TRY(cudaMemsetAsync(t->count, 0, size * 4, sq_cur_stream(ctx)));
    CUDA_TRY(cudaMemsetAsync(t->prio, 0xFF, size * 8, sq_cur_stream(ctx)));
    return SQ_OK;
}
static void table_free(sq_ctx *ctx, DdTable *t) {
    sq_dfree(ctx, t->hash);
    sq_dfree(ctx, t->count);
    sq_dfree(ctx, t->prio);
}

extern "C" int sq_dedup_create(sq_ctx *ctx, uint64_t max_stored_fingerprints, uint64_t front_len,
                               uint64_t back_len, uint64_t front_off, uint64_t back_off, sq_dedup **out) {
    *out = nullptr;
    if (max_stored_fingerprints < 100 || front_len + back_len == 0) {
        sq_set_error("invalid DedupEstimator parameters");
        return SQ_E_ARG;
    }
    CUDA_TRY(cudaSetDevice(ctx->device));
    sq_dedup *d = new sq_dedup();
    d->ctx = ctx;
    d->max_stored = max_stored_fingerprints;
    d->front_len = front_len;
    d->back_len = back_len;
    d->front_off = front_off;
    d->back_off = back_off;
    uint64_t bits = (uint64_t)(log2((double)max_stored_fingerprints * 1.5) + 1);  // :4327
    d->table_size = 1ULL << bits;
    int rc = table_alloc(ctx, &d->tab, d->table_size);
    if (rc == SQ_OK) rc = table_alloc(ctx, &d->spare, d->table_size);
    if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&d->cnt, sizeof(DdCounters), true);
    if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&d->stale_fp, front_len + back_len + 16, true);
    if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&d->pair_range, 8, true);
    if (rc == SQ_OK) rc = cudaMemsetAsync(d->pair_range, 0xFF, 4, sq_cur_stream(ctx)) == cudaSuccess ? SQ_OK : SQ_E_CUDA;
    if (rc != SQ_OK) {
        sq_dedup_destroy(d);
        return rc;
    }
    *out = d;
    return SQ_OK;
}

extern "C" void sq_dedup_destroy(sq_dedup *d) {
    if (!d) return;
    cudaSetDevice(d->ctx->device);
    table_free(d->ctx, &d->tab);
    table_free(d->ctx, &d->spare);
    for (auto &k : d->kept) sq_dfree(d->ctx, k.hashes);
    sq_dfree(d->ctx, d->compact);
    sq_dfree(d->ctx, d->cnt);
    sq_dfree(d->ctx, d->stale_fp);
    sq_dfree(d->ctx, d->pair_range);
    delete d;
}

__global__ void __launch_bounds__(DD_TPB)
k_dd_pass_flags(const uint64_t *__restrict__ hashes, uint32_t n, uint64_t mask, uint32_t *__restrict__ flag) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        flag[i] = (hashes[i] & mask) == 0;
}
__global__ void __launch_bounds__(DD_TPB)
k_dd_pass_scatter(const uint64_t *__restrict__ hashes, const uint32_t *__restrict__ flag,
                  const uint32_t *__restrict__ rank, uint32_t n, uint64_t *__restrict__ out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (flag[i]) out[rank[i]] = hashes[i];
}

static int dedup_consume_range(sq_dedup *d, const uint64_t *hashes, uint32_t n);

// Process hashes[0..n) in record order.
int dedup_consume(sq_dedup *d, const uint64_t *hashes, uint32_t n) {
    sq_ctx *ctx = d->ctx;
    if (d->deferred) {
        d->kept.push_back({(uint64_t *)hashes, n});
        d->n_records += n;
        return SQ_OK;
    }
    // A hash that fails the mask of m bits fails every later mask (:4429-4431 tests the low bits and m
    // only grows), so once m > 0 only n / 2^m hashes can touch the table: keep those, in record order,
    // and run the table kernels over the short list (their order-dependent steps only compare positions).
    int rc;
    if (d->mod_bits >= 2 && n >= 4096) {
        uint32_t *flag = nullptr, *rank = nullptr;
        uint64_t *kept = nullptr;
        SqScratch scratch(ctx);
        SQ_TRY(scratch.get(&flag, (size_t)n * 4));
        SQ_TRY(scratch.get(&rank, (size_t)n * 4 + 4));
        const int grid = sq_grid_for(ctx, n, DD_TPB, 16);
        SQ_LAUNCH(ctx, k_dd_pass_flags, grid, DD_TPB, 0, hashes, n, (1ULL << d->mod_bits) - 1, flag);
        SQ_TRY(sq_scan_exclusive_u32(ctx, flag, rank, n, rank + n));
        uint32_t *h_total = (uint32_t *)((char *)ctx->h_scratch + 3088);
        CUDA_TRY(cudaMemcpyAsync(h_total, rank + n, 4, cudaMemcpyDeviceToHost, sq_cur_stream(ctx)));
        CUDA_TRY(cudaStreamSynchronize(sq_cur_stream(ctx)));
        const uint32_t n_kept = *h_total;
        SQ_TRY(scratch.get(&kept, ((size_t)n_kept + 1) * 8));
        SQ_LAUNCH(ctx, k_dd_pass_scatter, grid, DD_TPB, 0, hashes, flag, rank, n, kept);
        rc = n_kept ? dedup_consume_range(d, kept, n_kept) : SQ_OK;
    }
    else rc = dedup_consume_range(d, hashes, n);
    d->n_records += n;
    return rc;
}

static int dedup_consume_range(sq_dedup *d, const uint64_t *hashes, uint32_t n) {
    sq_ctx *ctx = d->ctx;
    const uint64_t tmask = d->table_size - 1;
    const uint64_t prio_base = d->table_size + 1 + d->n_records;  // above every rebuild priority
    uint32_t *cls = nullptr, *flag = nullptr, *rank = nullptr;
    SqScratch scratch(ctx);
    SQ_TRY(scratch.get(&cls, (size_t)n * 4));
    SQ_TRY(scratch.get(&flag, (size_t)n * 4));
    SQ_TRY(scratch.get(&rank, (size_t)n * 4));
    // scratch sized for the worst case (every record a distinct new key)
    uint32_t scap = 1024;
    while (scap < 2 * (uint64_t)n) scap <<= 1;
    Scratch S;
    S.mask = scap - 1;
    SQ_TRY(scratch.get(&S.key, (size_t)scap * 8));
    SQ_TRY(scratch.get(&S.first, (size_t)scap * 4));
    SQ_TRY(scratch.get(&S.cnt, (size_t)scap * 4));
    DdCounters *hc = (DdCounters *)((char *)ctx->h_scratch + 2048);
    int rc = SQ_OK;
    uint32_t lo = 0;
    while (lo < n && rc == SQ_OK) {
        uint32_t hi = n;
        const uint32_t m = (uint32_t)d->mod_bits;
        bool truncated = false;
        for (;;) {  // at most two passes: whole segment, then the part before the trigger
            const uint32_t len = hi - lo;
            const int grid = sq_grid_for(ctx, len, DD_TPB, 16);
            CUDA_TRY(cudaMemsetAsync(S.key, 0xFF, (size_t)scap * 8, sq_cur_stream(ctx)));
            CUDA_TRY(cudaMemsetAsync(S.first, 0xFF, (size_t)scap * 4, sq_cur_stream(ctx)));
            CUDA_TRY(cudaMemsetAsync(S.cnt, 0, (size_t)scap * 4, sq_cur_stream(ctx)));
            DdCounters init;
            init.r_full = ~0ULL;
            init.r_star = ~0ULL;
            init.n_new = init.kept = init.inserted_one = init.pad = 0;
            *hc = init;
            CUDA_TRY(cudaMemcpyAsync(d->cnt, hc, sizeof(init), cudaMemcpyHostToDevice, sq_cur_stream(ctx)));
            SQ_LAUNCH(ctx, k_dd_classify, grid, DD_TPB, 0, hashes, lo, hi, m, d->tab, tmask, cls, S);
            SQ_LAUNCH(ctx, k_dd_flags, grid, DD_TPB, 0, hashes, cls, lo, hi, S, flag, d->cnt);
            if (truncated) break;  // range ends before the trigger: no escalation inside
            const bool already_full = d->stored >= d->max_stored;
            if (!already_full) {
                SQ_TRY(sq_scan_exclusive_u32(ctx, flag, rank, len, nullptr));
                uint64_t K = d->max_stored - d->stored;
                if (K <= len)
                    SQ_LAUNCH(ctx, k_dd_trigger_full, grid, DD_TPB, 0, flag, rank, lo, hi, (uint32_t)K, d->cnt);
            }
            SQ_LAUNCH(ctx, k_dd_trigger_star, grid, DD_TPB, 0, cls, lo, hi, already_full ? 1 : 0, d->cnt);
            CUDA_TRY(cudaMemcpyAsync(hc, d->cnt, sizeof(DdCounters), cudaMemcpyDeviceToHost, sq_cur_stream(ctx)));
            CUDA_TRY(cudaStreamSynchronize(sq_cur_stream(ctx)));
            if (hc->r_star == ~0ULL) break;  // no add meets a full table
            hi = (uint32_t)hc->r_star;
            truncated = true;
            if (hi == lo) break;  // the very first record of the segment triggers
        }
        const uint32_t r_star = truncated ? hi : n;
        if (hi > lo) {
            const uint32_t len = hi - lo;
            const int grid = sq_grid_for(ctx, len, DD_TPB, 16);
            if (truncated) {  // n_new of the truncated range
                CUDA_TRY(cudaMemcpyAsync(hc, d->cnt, sizeof(DdCounters), cudaMemcpyDeviceToHost, sq_cur_stream(ctx)));
                CUDA_TRY(cudaStreamSynchronize(sq_cur_stream(ctx)));
            }
            SQ_LAUNCH(ctx, k_dd_apply_existing, grid, DD_TPB, 0, cls, lo, hi, d->tab);
            SQ_LAUNCH(ctx, k_dd_insert_new, grid, DD_TPB, 0, hashes, flag, lo, hi, m, d->tab, tmask, prio_base);
            SQ_LAUNCH(ctx, k_dd_place_new, grid, DD_TPB, 0, hashes, flag, lo, hi, m, d->tab, tmask, prio_base, S);
            d->stored += hc->n_new;
        }
        if (!truncated) break;
        // ---- escalation at record r_star ---------------------------------------------------
        SQ_TRY(table_clear(ctx, &d->spare, d->table_size));
        CUDA_TRY(cudaMemsetAsync(&d->cnt->kept, 0, 8, sq_cur_stream(ctx)));
        const int tgrid = sq_grid_for(ctx, d->table_size, DD_TPB, 16);
        SQ_LAUNCH(ctx, k_dd_rebuild_insert, tgrid, DD_TPB, 0, d->tab, d->spare, d->table_size, m + 1, d->cnt);
        SQ_LAUNCH(ctx, k_dd_rebuild_place, tgrid, DD_TPB, 0, d->tab, d->spare, d->table_size, m + 1);
        DdTable t = d->tab;
        d->tab = d->spare;
        d->spare = t;
        SQ_LAUNCH(ctx, k_dd_insert_one, 1, 32, 0, hashes, r_star, m, d->tab, tmask, prio_base, d->cnt);
        CUDA_TRY(cudaMemcpyAsync(hc, d->cnt, sizeof(DdCounters), cudaMemcpyDeviceToHost, sq_cur_stream(ctx)));
        CUDA_TRY(cudaStreamSynchronize(sq_cur_stream(ctx)));
        d->stored = (uint64_t)hc->kept + hc->inserted_one;
        d->mod_bits = m + 1;
        lo = r_star + 1;
    }
    return rc;
}

extern "C" int sq_dedup_add(sq_dedup *d, sq_batch *b) {
    sq_ctx *ctx = d->ctx;
    if (b->ctx != ctx) {
        sq_set_error("record array belongs to another context");
        return SQ_E_ARG;
    }
    if (b->n == 0) return SQ_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    uint64_t *hashes = nullptr;
    SqScratch scratch(ctx);
    SQ_TRY(scratch.get(&hashes, b->n * 8));
    SQ_LAUNCH(ctx, k_dd_hash, sq_grid_for(ctx, b->n, DD_TPB, 16), DD_TPB, 0, b->view(), d->front_len,
              d->back_len, d->front_off, d->back_off, hashes);
    int rc = dedup_consume(d, hashes, (uint32_t)b->n);
    if (d->deferred && rc == SQ_OK) scratch.keep(hashes);  // a deferred estimator holds on to them
    return rc;
}

extern "C" int sq_dedup_add_pair(sq_dedup *d, sq_batch *b1, sq_batch *b2) {
    sq_ctx *ctx = d->ctx;
    if (b1->ctx != ctx || b2->ctx != ctx) {
        sq_set_error("record array belongs to another context");
        return SQ_E_ARG;
    }
    if (b1->n != b2->n) {
        sq_set_error("record_array1 and record_array2 must be of the same size");
        return SQ_E_ARG;
    }
    if (b1->n == 0) return SQ_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const uint32_t n = (uint32_t)b1->n;
    uint64_t *hashes = nullptr;
    uint32_t *short_flag = nullptr;
    SqScratch scratch(ctx);
    SQ_TRY(scratch.get(&hashes, (size_t)n * 8));
    SQ_TRY(scratch.get(&short_flag, (size_t)n * 4));
    SQ_LAUNCH(ctx, k_dd_hash_pair, sq_grid_for(ctx, n, DD_TPB, 16), DD_TPB, 0, b1->view(), b2->view(),
              d->front_len, d->back_len, d->front_off, d->back_off, hashes, short_flag, d->pair_range);
    // short pairs hash what the previous pair left in the reference's scratch buffer (:4503-4516): one warp
    // re-hashes them in record order and leaves the batch's last pair in stale_fp for the next batch
    SQ_LAUNCH(ctx, k_dd_hash_pair_ordered, 1, 1024, 0, b1->view(), b2->view(), d->front_len, d->back_len,
              d->front_off, d->back_off, hashes, short_flag, d->pair_range, d->stale_fp);
    int rc = dedup_consume(d, hashes, n);
    if (d->deferred && rc == SQ_OK) scratch.keep(hashes);
    return rc;
}

extern "C" int sq_dedup_sync(sq_dedup *d, sq_dedup_info *info) {
    CUDA_TRY(cudaSetDevice(d->ctx->device));
    CUDA_TRY(cudaStreamSynchronize(sq_cur_stream(d->ctx)));
    info->modulo_bits = d->mod_bits;
    info->hash_table_size = d->table_size;
    info->tracked_sequences = d->stored;
    return SQ_OK;
}

// counts of the occupied slots, in slot order (the reference getter, :4736-4744), compacted on the device
__global__ void __launch_bounds__(DD_TPB)
k_dd_occupied(const uint32_t *__restrict__ count, uint32_t size, uint32_t *__restrict__ flag) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < size; i += gridDim.x * blockDim.x)
        flag[i] = count[i] != 0;
}
__global__ void __launch_bounds__(DD_TPB)
k_dd_compact(const uint32_t *__restrict__ count, const uint32_t *__restrict__ rank, uint32_t size,
             uint64_t *__restrict__ out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < size; i += gridDim.x * blockDim.x)
        if (count[i]) out[rank[i]] = count[i];
}

extern "C" int sq_dedup_read(sq_dedup *d, uint64_t *counts, uint64_t *n) {
    sq_ctx *ctx = d->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    *n = 0;
    const uint32_t size = (uint32_t)d->table_size;
    uint32_t *flag = nullptr, *rank = nullptr, *total = nullptr;
    uint64_t *out = nullptr;
    SqScratch scratch(ctx);
    SQ_TRY(scratch.get(&flag, (size_t)size * 4));
    SQ_TRY(scratch.get(&rank, (size_t)size * 4 + 4));
    total = rank + size;
    SQ_TRY(scratch.get(&out, (size_t)(d->stored + 1) * 8));
    const int grid = sq_grid_for(ctx, size, DD_TPB, 16);
    SQ_LAUNCH(ctx, k_dd_occupied, grid, DD_TPB, 0, d->tab.count, size, flag);
    SQ_TRY(sq_scan_exclusive_u32(ctx, flag, rank, size, total));
    SQ_LAUNCH(ctx, k_dd_compact, grid, DD_TPB, 0, d->tab.count, rank, size, out);
    uint32_t *h_total = (uint32_t *)((char *)ctx->h_scratch + 3080);
    CUDA_TRY(cudaMemcpyAsync(h_total, total, 4, cudaMemcpyDeviceToHost, sq_cur_stream(ctx)));
    CUDA_TRY(cudaStreamSynchronize(sq_cur_stream(ctx)));
    const uint64_t got = *h_total;
    if (got > d->stored) {
        sq_set_error("dedup table holds %llu entries, expected at most %llu", (unsigned long long)got,
                     (unsigned long long)d->stored);
        return SQ_E_CUDA;
    }
    if (got) SQ_TRY(sq_d2h_bounced(ctx, counts, out, got * 8));
    *n = got;
    return SQ_OK;
}


// ---------------------------------------------------------------------------
// Sharded runs.  The estimator's table is order dependent
// (escalation point, stale-index quirk), so one rank owns it.  Every other rank
// only hashes its reads (deferred mode).  A hash that fails the mask of m bits
// fails every later mask too (:4429-4431 tests the low bits, m only grows), so
// once the owner has finished its own shard with m0 bits the other ranks can
// drop everything that fails mask(m0) and hand over the rest in record order:
// about n / 2^m0 hashes per rank.
// ---------------------------------------------------------------------------
extern "C" int sq_dedup_set_deferred(sq_dedup *d, int deferred) {
    if (d->n_records != 0 && (deferred != 0) != d->deferred) {
        sq_set_error("deferred mode must be chosen before the first record array");
        return SQ_E_ARG;
    }
    d->deferred = deferred != 0;
    return SQ_OK;
}

// Compacts the kept hashes that pass mask(mod_bits) into one device buffer (record order kept)
// and releases the per-array buffers.  *n = number of surviving hashes.
extern "C" int sq_dedup_deferred_compact(sq_dedup *d, uint64_t mod_bits, uint64_t *n) {
    sq_ctx *ctx = d->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    *n = 0;
    const uint64_t mask = mod_bits >= 64 ? ~0ULL : (1ULL << mod_bits) - 1;
    const size_t n_arr = d->kept.size();
    std::vector<uint32_t *> flags(n_arr, nullptr), ranks(n_arr, nullptr);
    uint32_t *totals = nullptr;
    SqScratch scratch(ctx);
    SQ_TRY(scratch.get(&totals, (n_arr + 1) * 4, true));
    for (size_t a = 0; a < n_arr; a++) {
        const uint32_t len = d->kept[a].n;
        SQ_TRY(scratch.get(&flags[a], (size_t)len * 4));
        SQ_TRY(scratch.get(&ranks[a], (size_t)len * 4));
        SQ_LAUNCH(ctx, k_dd_pass_flags, sq_grid_for(ctx, len, DD_TPB, 16), DD_TPB, 0, d->kept[a].hashes, len, mask,
                  flags[a]);
        SQ_TRY(sq_scan_exclusive_u32(ctx, flags[a], ranks[a], len, totals + a));
    }
    std::vector<uint32_t> h_tot(n_arr + 1, 0);
    if (n_arr) CUDA_TRY(cudaMemcpyAsync(h_tot.data(), totals, n_arr * 4, cudaMemcpyDeviceToHost, sq_cur_stream(ctx)));
    CUDA_TRY(cudaStreamSynchronize(sq_cur_stream(ctx)));
    uint64_t total = 0;
    for (size_t a = 0; a < n_arr; a++) total += h_tot[a];
    sq_dfree(ctx, d->compact);
    d->compact = nullptr;
    SQ_TRY(sq_dalloc(ctx, (void **)&d->compact, (total + 1) * 8, false));
    uint64_t off = 0;
    for (size_t a = 0; a < n_arr; a++) {
        const uint32_t len = d->kept[a].n;
        SQ_LAUNCH(ctx, k_dd_pass_scatter, sq_grid_for(ctx, len, DD_TPB, 16), DD_TPB, 0, d->kept[a].hashes, flags[a],
                  ranks[a], len, d->compact + off);
        off += h_tot[a];
        sq_dfree(ctx, d->kept[a].hashes);
    }
    d->kept.clear();
    d->compact_n = total;
    *n = total;
    return SQ_OK;
}

// Copies the compacted hashes to a caller-owned DEVICE buffer of compact_n entries.
extern "C" int sq_dedup_deferred_fetch(sq_dedup *d, uint64_t *dev_out) {
    sq_ctx *ctx = d->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (d->compact_n)
        CUDA_TRY(cudaMemcpyAsync(dev_out, d->compact, d->compact_n * 8, cudaMemcpyDeviceToDevice, sq_cur_stream(ctx)));
    CUDA_TRY(cudaStreamSynchronize(sq_cur_stream(ctx)));
    sq_dfree(ctx, d->compact);
    d->compact = nullptr;
    d->compact_n = 0;
    return SQ_OK;
}

// DedupEstimator_add_fingerprint (:4426-4460) for n hashes that already sit in device memory,
// in record order (the hashes another rank handed over).
extern "C" int sq_dedup_add_hashes(sq_dedup *d, const uint64_t *dev_hashes, uint64_t n) {
    if (d->deferred) {
        sq_set_error("a deferred DedupEstimator owns no table");
        return SQ_E_ARG;
    }
    CUDA_TRY(cudaSetDevice(d->ctx->device));
    const uint64_t step = 1ULL << 26;
    for (uint64_t lo = 0; lo < n; lo += step) {
        const uint64_t len = n - lo < step ? n - lo : step;
        SQ_TRY(dedup_consume(d, dev_hashes + lo, (uint32_t)len));
    }
    CUDA_TRY(cudaStreamSynchronize(sq_cur_stream(d->ctx)));
    return SQ_OK;
}
