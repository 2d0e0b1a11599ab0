// overrep.cu -- OverrepresentedSequences (reference _qcmodule.c:3543-3568, 3589-3608,
// 3635-3694, 3830-3942).
//
// Every `sample_every`-th read (by running read count) contributes the
// canonical 2-bit k-mers of a few non-overlapping fragments from both ends,
// de-duplicated within the read, to a capped open-addressing count table keyed
// by Thomas Wang's hash of the k-mer.
//
//   k_ov_fragments   one thread per sampled read: k-mers, reverse complement,
//                    hash, and the reference's per-read staging table (its slot
//                    order is the tie-break when the cap is hit mid-read)
//   k_ov_count       insert-or-increment with atomicCAS (table below the cap:
//                    content is order-free), or lookup-only (table full)
//   k_ov_classify / k_ov_flags / scan / k_ov_admit   the one batch that crosses
//                    the cap: the first (max - unique) new hashes in (read,
//                    staging slot) order are admitted, the rest dropped
#include <math.h>

#include "common.cuh"
#include "modules.cuh"

constexpr int OV_TPB = 128;
constexpr int OV_STAGE_MAX = 128;  // per-read staging slots kept in local memory

struct OvCounters {
    unsigned long long total_frags;
    unsigned long long warn_records;
    unsigned long long first_warn;
    unsigned int n_unique;
    unsigned int n_new;
};

struct sq_overrep {
    sq_ctx *ctx = nullptr;
    uint64_t max_unique = 0, k = 0, sample_every = 0, frags_front = 0, frags_back = 0;
    uint64_t n_seqs = 0, n_sampled = 0, table_size = 0;
    uint64_t unique_known = 0;   // exact unique count at the last synchronisation
    uint64_t unique_upper = 0;   // upper bound since then
    bool full = false;
    uint64_t *keys = nullptr;    // wang hash, 0 = empty
    uint32_t *counts = nullptr;
    OvCounters *cnt = nullptr;
    // Once the table is full its key set is frozen (:3553): a presence bitmap (one bit per value of
    // 27 hash bits the table index does not use, 16 MiB: stays in L2) answers most lookups of unknown
    // hashes without touching the table in HBM.
    uint32_t *filter = nullptr;
    bool filter_valid = false;
    // sharded runs: fragments of the sampled reads are kept, the table work waits for the
    // table state of the ranks before this one (sq_overrep_apply_deferred)
    bool deferred = false;
    struct Kept { uint64_t *frag_hash; uint32_t *frag_n; uint64_t n_sampled, total; uint32_t fcap; };
    std::vector<Kept> kept;
};

// 0x01 per byte that is ACGTacgt (same construction as qc.cu / fused.cu)
__device__ __forceinline__ uint32_t ov_acgt_bytes(uint32_t w) {
    uint32_t sel = w & 0x07070707u;
    uint32_t t = sel | (sel >> 4);
    uint32_t nib = __byte_perm(t, 0, 0x4420);
    uint32_t expect = __byte_perm(0x40FF40FFu, 0x40FFFF50u, nib);
    return zero_bytes80((w & 0xD8D8D8D8u) ^ expect) >> 7;
}

// canonical k-mer of s[0..k): 0 ok, 1 holds N/n, 2 holds another non-ACGT letter (:3612-3694).
// The k <= 31 letters come in with aligned word loads issued together (the text has 64 readable
// bytes of padding) and are encoded four at a time: the 2-bit code A=0 C=1 G=2 T=3 is
// ((c >> 1) & 3) ^ ((c >> 2) & 1) for either case, one multiplication gathers four codes into a
// byte, and the reverse complement is the complemented forward k-mer with its 2-bit groups in
// reverse order (bit reversal + a swap inside the pairs).  A fragment with any other letter takes
// the letter-by-letter path (it only has to tell N from the rest).
__device__ __forceinline__ int canonical_kmer(const uint8_t *s, uint32_t k, uint64_t *out) {
    const uint32_t *wp = (const uint32_t *)((uintptr_t)s & ~(uintptr_t)3);
    const uint32_t sh = ((uint32_t)(uintptr_t)s & 3u) * 8;
    const uint32_t nw = (k + 3) / 4;  // words after alignment; one more is read for the shift
    uint32_t W[9];
#pragma unroll
    for (int j = 0; j < 9; j++) W[j] = (uint32_t)j <= nw ? __ldg(wp + j) : 0u;
    uint64_t fw = 0;
    uint32_t bad = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        if ((uint32_t)j < nw) {
            const uint32_t a = __funnelshift_r(W[j], W[j + 1], sh);
            const uint32_t left = k - 4 * j;  // letters of the fragment in this word (>= 1)
            const uint32_t pm = left >= 4 ? 0x01010101u : 0x01010101u >> (8 * (4 - left));
            bad |= pm & ~ov_acgt_bytes(a);
            const uint32_t code = (((a >> 1) & 0x03030303u) ^ ((a >> 2) & 0x01010101u)) & (pm * 3u);
            fw = (fw << 8) | ((code * 0x40100401u) >> 24);  // first letter in the top two bits
        }
    }
    if (bad) {  // rare: which kind of letter is it
        uint32_t flags = 0;
        for (uint32_t i = 0; i < k; i++) {
            const uint32_t ch = s[i] | 0x20u;
            if (ch != 'a' && ch != 'c' && ch != 'g' && ch != 't') flags |= ch == 'n' ? 1u : 2u;
        }
        return flags & 2 ? 2 : 1;
    }
    fw >>= 2 * (4 * nw - k);  // the padding letters of the last word were encoded as zeros
    uint64_t r = __brevll(~fw);  // complement, then reverse the order of the 2-bit groups
    r = ((r >> 1) & 0x5555555555555555ULL) | ((r & 0x5555555555555555ULL) << 1);
    const uint64_t rc = r >> (64 - 2 * k);
    *out = rc < fw ? rc : fw;
    return 0;
}

// per-read staging set of fragment hashes (open addressing, read out in slot order: the order the
// reference's own staging table is walked in)
template <typename Stage>
__device__ __forceinline__ uint32_t ov_stage_read(Stage stage, uint32_t ssize, const uint8_t *seq, uint64_t L, uint32_t k,
                                                  uint64_t nf, uint64_t nb, uint64_t *__restrict__ out, bool *warn,
                                                  uint32_t *valid) {
    const uint32_t total = (uint32_t)(nf + nb);
    for (uint32_t i = 0; i < ssize; i++) stage(i) = 0;
    for (uint32_t f = 0; f < total; f++) {
        const uint64_t off = f < nf ? (uint64_t)f * k : L - (nb - (f - nf)) * k;
        uint64_t kmer;
        int rc = canonical_kmer(seq + off, k, &kmer);
        if (rc) {
            *warn |= rc == 2;
            continue;
        }
        (*valid)++;
        const uint64_t h = wang64(kmer);
        uint32_t i = (uint32_t)h & (ssize - 1);
        while (stage(i) != 0 && stage(i) != h) i = (i + 1) & (ssize - 1);
        stage(i) = h;
    }
    uint32_t emitted = 0;
    for (uint32_t i = 0; i < ssize; i++)
        if (stage(i)) out[emitted++] = stage(i);
    return emitted;
}

constexpr int OV_STAGE_SMEM = 16;  // staging sets up to this size live in shared memory (short reads)

__global__ void __launch_bounds__(OV_TPB)
k_ov_fragments(BatchView bv, uint32_t first_sampled, uint32_t sample_every, uint32_t n_sampled, uint32_t k,
               uint64_t frags_front, uint64_t frags_back, uint32_t fcap, uint64_t *__restrict__ frag_hash,
               uint32_t *__restrict__ frag_n, OvCounters *cnt, uint64_t record_base) {
    __shared__ uint64_t s_stage[OV_STAGE_SMEM][OV_TPB];
    unsigned long long valid_total = 0;
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n_sampled; s += gridDim.x * blockDim.x) {
        const uint32_t r = first_sampled + s * sample_every;
        const uint64_t L = bv.seq_len[r];
        uint32_t emitted = 0;
        if (L >= k) {
            const uint8_t *seq = bv.text + bv.seq_off[r];
            const uint64_t maxf = (L + k - 1) / k, back_cap = maxf / 2, front_cap = maxf - back_cap;
            const uint64_t nf = min(frags_front, front_cap), nb = min(frags_back, back_cap);
            const uint32_t total = (uint32_t)(nf + nb);
            if (total) {
                uint32_t ssize = 1;  // 2^ceil(log2(1.5*total)); 3*total is never a power of two
                while (2 * ssize < 3 * total) ssize <<= 1;
                bool warn = false;
                uint32_t valid = 0;
                uint64_t *out = frag_hash + (size_t)s * fcap;
                if (ssize <= OV_STAGE_SMEM) {
                    const uint32_t tid = threadIdx.x;
                    emitted = ov_stage_read([&](uint32_t i) -> uint64_t & { return s_stage[i][tid]; }, ssize, seq, L, k,
                                            nf, nb, out, &warn, &valid);
                }
                else if (ssize <= OV_STAGE_MAX) {
                    uint64_t stage[OV_STAGE_MAX];
                    emitted = ov_stage_read([&](uint32_t i) -> uint64_t & { return stage[i]; }, ssize, seq, L, k, nf, nb,
                                            out, &warn, &valid);
                }
                else {
                    // whole-read fragment sets of long reads (bases_from_start / bases_from_end < 0 or large,
                    // :3499-3504): the read's own row of frag_hash (fcap >= ssize slots) is the staging table,
                    // compacted in place in slot order afterwards (the write index never passes the read index)
                    emitted = ov_stage_read([&](uint32_t i) -> uint64_t & { return out[i]; }, ssize, seq, L, k, nf, nb,
                                            out, &warn, &valid);
                }
                valid_total += valid;
                if (warn) {
                    atomicAdd(&cnt->warn_records, 1ULL);
                    atomicMin(&cnt->first_warn, (unsigned long long)(record_base + r));
                }
            }
        }
        frag_n[s] = emitted;
    }
    // block-level aggregation of the (non de-duplicated) fragment count
    for (int o = 16; o > 0; o >>= 1) valid_total += __shfl_xor_sync(0xffffffffu, valid_total, o);
    if (lane_id() == 0 && valid_total) atomicAdd(&cnt->total_frags, valid_total);
}

__device__ __forceinline__ void ov_insert_or_count(uint64_t *keys, uint32_t *counts, uint64_t mask, uint64_t h,
                                                   bool may_insert, unsigned int *n_unique) {
    uint64_t i = h & mask;
    for (;;) {
        uint64_t kk = keys[i];
        if (kk == h) {
            atomicAdd(&counts[i], 1u);
            return;
        }
        if (kk == 0) {
            if (!may_insert) return;  // table full: unknown hashes are dropped (:3553)
            uint64_t old = atomicCAS((unsigned long long *)&keys[i], 0ULL, (unsigned long long)h);
            if (old == 0) {
                atomicAdd(n_unique, 1u);
                atomicAdd(&counts[i], 1u);
                return;
            }
            if (old == h) {
                atomicAdd(&counts[i], 1u);
                return;
            }
        }
        i = (i + 1) & mask;
    }
}

constexpr uint32_t OV_FILTER_BITS = 27, OV_FILTER_SHIFT = 23;  // table index = low bits of the hash
__device__ __forceinline__ uint32_t ov_filter_slot(uint64_t h) {
    return (uint32_t)(h >> OV_FILTER_SHIFT) & ((1u << OV_FILTER_BITS) - 1);
}
__global__ void __launch_bounds__(256)
k_ov_filter_build(const uint64_t *__restrict__ keys, uint64_t table_size, uint32_t *__restrict__ filter) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < table_size;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t h = keys[i];
        if (h) {
            const uint32_t b = ov_filter_slot(h);
            atomicOr(filter + (b >> 5), 1u << (b & 31));
        }
    }
}

// filter != nullptr: the table is full (lookup only) and `filter` holds a bit for every stored key
__global__ void __launch_bounds__(256)
k_ov_count(const uint64_t *__restrict__ frag_hash, const uint32_t *__restrict__ frag_n, uint32_t n_sampled,
           uint32_t fcap, uint64_t *keys, uint32_t *counts, uint64_t mask, int may_insert, OvCounters *cnt,
           const uint32_t *__restrict__ filter) {
    const uint64_t total = (uint64_t)n_sampled * fcap;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t s = (uint32_t)(t / fcap), i = (uint32_t)(t % fcap);
        if (i >= frag_n[s]) continue;
        const uint64_t h = frag_hash[t];
        if (filter) {
            const uint32_t b = ov_filter_slot(h);
            if (!((filter[b >> 5] >> (b & 31)) & 1u)) continue;  // not a stored key: dropped (:3553)
        }
        ov_insert_or_count(keys, counts, mask, h, may_insert != 0, &cnt->n_unique);
    }
}

// ---- the batch that crosses the cap ---------------------------------------------------------
struct OvScratch {
    uint64_t *key;    // ~0 = free
    uint32_t *first;  // smallest occurrence index t holding the key
    uint32_t mask;
};
__device__ __forceinline__ uint32_t ov_scratch_slot(const OvScratch &S, uint64_t h, bool insert) {
    uint32_t i = (uint32_t)(h ^ (h >> 29)) & S.mask;
    for (;;) {
        uint64_t kk = S.key[i];
        if (kk == h) return i;
        if (kk == ~0ULL) {
            if (!insert) return 0xFFFFFFFFu;
            uint64_t old = atomicCAS((unsigned long long *)&S.key[i], ~0ULL, (unsigned long long)h);
            if (old == ~0ULL || old == h) return i;
        }
        i = (i + 1) & S.mask;
    }
}
// cls[t]: 0 = slot unused, 1 = hash already in the table, 2 = new hash
__global__ void __launch_bounds__(256)
k_ov_classify(const uint64_t *__restrict__ frag_hash, const uint32_t *__restrict__ frag_n, uint32_t n_sampled,
              uint32_t fcap, const uint64_t *__restrict__ keys, uint64_t mask, uint8_t *__restrict__ cls,
              OvScratch S) {
    const uint64_t total = (uint64_t)n_sampled * fcap;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t s = (uint32_t)(t / fcap), i = (uint32_t)(t % fcap);
        if (i >= frag_n[s]) {
            cls[t] = 0;
            continue;
        }
        const uint64_t h = frag_hash[t];
        uint64_t p = h & mask;
        uint8_t c = 2;
        for (;;) {
            uint64_t kk = keys[p];
            if (kk == 0) break;
            if (kk == h) {
                c = 1;
                break;
            }
            p = (p + 1) & mask;
        }
        cls[t] = c;
        if (c == 2) atomicMin(&S.first[ov_scratch_slot(S, h, true)], (uint32_t)t);
    }
}
__global__ void __launch_bounds__(256)
k_ov_flags(const uint64_t *__restrict__ frag_hash, const uint8_t *__restrict__ cls, uint64_t total, OvScratch S,
           uint32_t *__restrict__ flag) {
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t f = 0;
        if (cls[t] == 2) f = S.first[ov_scratch_slot(S, frag_hash[t], false)] == (uint32_t)t;
        flag[t] = f;
    }
}
// admit the first K new hashes in occurrence order; count every occurrence of
// table hashes and of admitted hashes
__global__ void __launch_bounds__(256)
k_ov_admit(const uint64_t *__restrict__ frag_hash, const uint8_t *__restrict__ cls, uint64_t total,
           OvScratch S, const uint32_t *__restrict__ rank, uint32_t K, uint64_t *keys, uint32_t *counts,
           uint64_t mask, OvCounters *cnt) {
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint8_t c = cls[t];
        if (c == 0) continue;
        const uint64_t h = frag_hash[t];
        bool ok = c == 1;
        if (c == 2) {
            uint32_t first = S.first[ov_scratch_slot(S, h, false)];
            ok = rank[first] < K;  // rank of the hash's first occurrence among new hashes
        }
        if (ok) ov_insert_or_count(keys, counts, mask, h, true, &cnt->n_unique);
    }
}

// ---------------------------------------------------------------------------
extern "C" int sq_overrep_create(sq_ctx *ctx, uint64_t max_unique_fragments, uint32_t fragment_length,
                                 uint64_t sample_every, int64_t bases_from_start, int64_t bases_from_end,
                                 sq_overrep **out) {
    *out = nullptr;
    if (max_unique_fragments < 1 || (fragment_length & 1) == 0 || fragment_length > 31 ||
        fragment_length < 3 || sample_every < 1) {
        sq_set_error("invalid OverrepresentedSequences parameters");
        return SQ_E_ARG;
    }
    CUDA_TRY(cudaSetDevice(ctx->device));
    sq_overrep *o = new sq_overrep();
    o->ctx = ctx;
    o->max_unique = max_unique_fragments;
    o->k = fragment_length;
    o->sample_every = sample_every;
    if (bases_from_start < 0) bases_from_start = UINT32_MAX;  // :3499-3504
    if (bases_from_end < 0) bases_from_end = UINT32_MAX;
    o->frags_front = ((uint64_t)bases_from_start + fragment_length - 1) / fragment_length;
    o->frags_back = ((uint64_t)bases_from_end + fragment_length - 1) / fragment_length;
    uint64_t bits = (uint64_t)(log2((double)max_unique_fragments * 1.5) + 1);  // :3508
    o->table_size = 1ULL << bits;
    int rc = sq_dalloc(ctx, (void **)&o->keys, o->table_size * 8, true);
    if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&o->counts, o->table_size * 4, true);
    if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&o->cnt, sizeof(OvCounters), true);
    if (rc == SQ_OK)
        rc = cudaMemsetAsync(&o->cnt->first_warn, 0xFF, 8, sq_cur_stream(ctx)) == cudaSuccess ? SQ_OK : SQ_E_CUDA;
    if (rc != SQ_OK) {
        sq_overrep_destroy(o);
        return rc;
    }
    *out = o;
    return SQ_OK;
}

extern "C" void sq_overrep_destroy(sq_overrep *o) {
    if (!o) return;
    cudaSetDevice(o->ctx->device);
    for (auto &k : o->kept) {
        sq_dfree(o->ctx, k.frag_hash);
        sq_dfree(o->ctx, k.frag_n);
    }
    sq_dfree(o->ctx, o->keys);
    sq_dfree(o->ctx, o->counts);
    sq_dfree(o->ctx, o->cnt);
    sq_dfree(o->ctx, o->filter);
    delete o;
}

static int ov_refresh_unique(sq_overrep *o) {
    sq_ctx *ctx = o->ctx;
    OvCounters *h = (OvCounters *)((char *)ctx->h_scratch + 1536);
    CUDA_TRY(cudaMemcpyAsync(h, o->cnt, sizeof(OvCounters), cudaMemcpyDeviceToHost, sq_cur_stream(ctx)));
    CUDA_TRY(cudaStreamSynchronize(sq_cur_stream(ctx)));
    o->unique_known = o->unique_upper = h->n_unique;
    o->full = o->unique_known >= o->max_unique;
    return SQ_OK;
}

static int ov_refresh_unique(sq_overrep *o);
// Table maintenance for the fragments of one record array (staging layout of k_ov_fragments):
// Sequence_duplication_insert_hash (:3543-3568) for every staged hash in (read, slot) order.
static int ov_apply(sq_overrep *o, const uint64_t *frag_hash, const uint32_t *frag_n, uint64_t n_sampled,
                    uint64_t total, uint32_t fcap) {
    sq_ctx *ctx = o->ctx;
    const uint64_t occ = n_sampled * fcap;
    const uint64_t mask = o->table_size - 1;
    const int grid = sq_grid_for(ctx, occ, 256, 16);
    int rc = SQ_OK;
    if (!o->full && o->unique_upper + n_sampled * total > o->max_unique) rc = ov_refresh_unique(o);
    if (rc == SQ_OK) {
        if (o->full) {
            if (!o->filter_valid) {
                const size_t fbytes = (size_t)1 << (OV_FILTER_BITS - 3);
                if (!o->filter) SQ_TRY(sq_dalloc(ctx, (void **)&o->filter, fbytes, false));
                CUDA_TRY(cudaMemsetAsync(o->filter, 0, fbytes, sq_cur_stream(ctx)));
                SQ_LAUNCH(ctx, k_ov_filter_build, sq_grid_for(ctx, o->table_size, 256, 16), 256, 0, o->keys,
                          o->table_size, o->filter);
                o->filter_valid = true;
            }
            SQ_LAUNCH(ctx, k_ov_count, grid, 256, 0, frag_hash, frag_n, (uint32_t)n_sampled, fcap, o->keys,
                      o->counts, mask, 0, o->cnt, o->filter_valid ? o->filter : nullptr);
        }
        else if (o->unique_upper + n_sampled * total <= o->max_unique) {
            SQ_LAUNCH(ctx, k_ov_count, grid, 256, 0, frag_hash, frag_n, (uint32_t)n_sampled, fcap, o->keys,
                      o->counts, mask, 1, o->cnt, (const uint32_t *)nullptr);
            o->unique_upper += n_sampled * total;
        }
        else {
            // this array may cross the cap: admission in (read, staging slot) order
            uint8_t *cls = nullptr;
            uint32_t *flag = nullptr, *rank = nullptr;
            OvScratch S;
            uint32_t scap = 1024;
            while (scap < 2 * occ) scap <<= 1;
            S.mask = scap - 1;
            rc = sq_dalloc(ctx, (void **)&cls, occ, false);
            if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&flag, occ * 4, false);
            if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&rank, occ * 4, false);
            if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&S.key, (size_t)scap * 8, false);
            if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&S.first, (size_t)scap * 4, false);
            if (rc == SQ_OK) {
                CUDA_TRY(cudaMemsetAsync(S.key, 0xFF, (size_t)scap * 8, sq_cur_stream(ctx)));
                CUDA_TRY(cudaMemsetAsync(S.first, 0xFF, (size_t)scap * 4, sq_cur_stream(ctx)));
                SQ_LAUNCH(ctx, k_ov_classify, grid, 256, 0, frag_hash, frag_n, (uint32_t)n_sampled, fcap, o->keys,
                          mask, cls, S);
                SQ_LAUNCH(ctx, k_ov_flags, grid, 256, 0, frag_hash, cls, occ, S, flag);
                rc = sq_scan_exclusive_u32(ctx, flag, rank, (uint32_t)occ, nullptr);
            }
            if (rc == SQ_OK) {
                const uint32_t K = (uint32_t)(o->max_unique - o->unique_known);
                SQ_LAUNCH(ctx, k_ov_admit, grid, 256, 0, frag_hash, cls, occ, S, rank, K, o->keys, o->counts, mask,
                          o->cnt);
                rc = ov_refresh_unique(o);
            }
            sq_dfree(ctx, cls);
            sq_dfree(ctx, flag);
            sq_dfree(ctx, rank);
            sq_dfree(ctx, S.key);
            sq_dfree(ctx, S.first);
        }
    }
    return rc;
}


// First half of an add: the sampled reads' fragment hashes (no host wait); ov_add_end applies them to the table.
// Split so that sq_fused_add can put the fragment kernel on the device right behind k_fused_reads, in front of the
// host wait of PerTileQuality's planning.
int ov_add_begin(sq_overrep *o, sq_batch *b, OvPendingAdd *pa) {
    *pa = OvPendingAdd();
    sq_ctx *ctx = o->ctx;
    if (b->ctx != ctx) {
        sq_set_error("record array belongs to another context");
        return SQ_E_ARG;
    }
    const uint64_t n = b->n;
    if (n == 0) return SQ_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    // reads whose running index is a multiple of sample_every (:3833)
    const uint64_t se = o->sample_every;
    const uint64_t first = (se - o->n_seqs % se) % se;
    const uint64_t n_sampled = first < n ? (n - first + se - 1) / se : 0;
    const uint64_t record_base = o->n_seqs;
    o->n_seqs += n;
    o->n_sampled += n_sampled;
    if (n_sampled == 0 || b->max_len < o->k) return SQ_OK;
    // staging capacity needed by the longest read of this array
    const uint64_t maxf = (b->max_len + o->k - 1) / o->k;
    const uint64_t nf = std::min(o->frags_front, maxf - maxf / 2), nb = std::min(o->frags_back, maxf / 2);
    const uint64_t total = nf + nb;
    if (total == 0) return SQ_OK;
    uint32_t fcap = 1;
    while (2 * (uint64_t)fcap < 3 * total) fcap <<= 1;
    const uint64_t occ = n_sampled * fcap;
    if (occ >= 0xFFFFFFFFULL) {
        sq_set_error("record array too large for the fragment index");
        return SQ_E_LIMIT;
    }
    uint64_t *frag_hash = nullptr;
    uint32_t *frag_n = nullptr;
    SqScratch scratch(ctx);
    SQ_TRY(scratch.get(&frag_hash, occ * 8));
    SQ_TRY(scratch.get(&frag_n, n_sampled * 4));
    SQ_LAUNCH(ctx, k_ov_fragments, sq_grid_for(ctx, n_sampled, OV_TPB, 16), OV_TPB, 0, b->view(), (uint32_t)first,
              (uint32_t)se, (uint32_t)n_sampled, (uint32_t)o->k, o->frags_front, o->frags_back, fcap, frag_hash,
              frag_n, o->cnt, record_base);
    scratch.keep(frag_hash);
    scratch.keep(frag_n);
    pa->frag_hash = frag_hash;
    pa->frag_n = frag_n;
    pa->n_sampled = n_sampled;
    pa->total = total;
    pa->fcap = fcap;
    return SQ_OK;
}

// Second half: table update (or, deferred, the hashes are kept).  Always releases what the first half allocated.
int ov_add_end(sq_overrep *o, OvPendingAdd *pa) {
    if (!pa->frag_hash) return SQ_OK;
    sq_ctx *ctx = o->ctx;
    if (o->deferred) {
        o->kept.push_back({pa->frag_hash, pa->frag_n, pa->n_sampled, pa->total, pa->fcap});
        *pa = OvPendingAdd();
        return SQ_OK;
    }
    const int rc = ov_apply(o, pa->frag_hash, pa->frag_n, pa->n_sampled, pa->total, pa->fcap);
    sq_dfree(ctx, pa->frag_hash);
    sq_dfree(ctx, pa->frag_n);
    *pa = OvPendingAdd();
    return rc;
}

void ov_add_abandon(sq_overrep *o, OvPendingAdd *pa) {
    sq_dfree(o->ctx, pa->frag_hash);
    sq_dfree(o->ctx, pa->frag_n);
    *pa = OvPendingAdd();
}

extern "C" int sq_overrep_add(sq_overrep *o, sq_batch *b) {
    OvPendingAdd pa;
    SQ_TRY(ov_add_begin(o, b, &pa));
    return ov_add_end(o, &pa);
}

extern "C" int sq_overrep_sync(sq_overrep *o, sq_overrep_info *info) {
    sq_ctx *ctx = o->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    OvCounters *h = (OvCounters *)((char *)ctx->h_scratch + 1536);
    CUDA_TRY(cudaMemcpyAsync(h, o->cnt, sizeof(OvCounters), cudaMemcpyDeviceToHost, sq_cur_stream(ctx)));
    CUDA_TRY(cudaStreamSynchronize(sq_cur_stream(ctx)));
    o->unique_known = o->unique_upper = h->n_unique;
    o->full = o->unique_known >= o->max_unique;
    info->number_of_sequences = o->n_seqs;
    info->sampled_sequences = o->n_sampled;
    info->collected_unique_fragments = h->n_unique;
    info->total_fragments = h->total_frags;
    info->max_unique_fragments = o->max_unique;
    info->table_size = o->table_size;
    info->warn_records = h->warn_records;
    info->first_warn_record = h->first_warn;
    return SQ_OK;
}

// table entries with count >= min_count, compacted on the device (slot order is
// not observable: the getters return a dict / a sorted list, :4020-4062, :4169-4175)
__global__ void __launch_bounds__(256)
k_ov_compact(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ counts, uint64_t table_size,
             uint32_t min_count, uint64_t cap, uint64_t *__restrict__ out_kmer, uint32_t *__restrict__ out_count,
             unsigned long long *n_out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < table_size + 31;
         i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t h = 0;
        uint32_t c = 0;
        if (i < table_size) {
            h = keys[i];
            c = counts[i];
        }
        const bool take = h != 0 && c >= min_count;
        const uint32_t m = __ballot_sync(0xffffffffu, take);
        if (!m) continue;
        unsigned long long base = 0;
        if (lane_id() == 0) base = atomicAdd(n_out, (unsigned long long)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (take) {
            const uint64_t w = base + __popc(m & ((1u << lane_id()) - 1));
            if (w < cap) {
                out_kmer[w] = wang64_inverse(h);  // key -> sequence, as the getter does (:4042)
                out_count[w] = c;
            }
        }
    }
}

extern "C" int sq_overrep_read_min(sq_overrep *o, uint32_t min_count, uint64_t *kmers, uint32_t *counts,
                                   uint64_t cap, uint64_t *n) {
    sq_ctx *ctx = o->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    *n = 0;
    uint64_t *dk = nullptr;
    uint32_t *dc = nullptr;
    unsigned long long *dn = nullptr;
    SQ_TRY(sq_dalloc(ctx, (void **)&dk, cap * 8, false));
    SQ_TRY(sq_dalloc(ctx, (void **)&dc, cap * 4, false));
    SQ_TRY(sq_dalloc(ctx, (void **)&dn, 8, true));
    SQ_LAUNCH(ctx, k_ov_compact, sq_grid_for(ctx, o->table_size + 31, 256, 16), 256, 0, o->keys, o->counts,
              o->table_size, min_count, cap, dk, dc, dn);
    unsigned long long *hn = (unsigned long long *)((char *)ctx->h_scratch + 1792);
    CUDA_TRY(cudaMemcpyAsync(hn, dn, 8, cudaMemcpyDeviceToHost, sq_cur_stream(ctx)));
    CUDA_TRY(cudaStreamSynchronize(sq_cur_stream(ctx)));
    const uint64_t got = *hn < cap ? *hn : cap;
    if (got) {
        CUDA_TRY(cudaMemcpyAsync(kmers, dk, got * 8, cudaMemcpyDeviceToHost, sq_cur_stream(ctx)));
        CUDA_TRY(cudaMemcpyAsync(counts, dc, got * 4, cudaMemcpyDeviceToHost, sq_cur_stream(ctx)));
        CUDA_TRY(cudaStreamSynchronize(sq_cur_stream(ctx)));
    }
    *n = *hn;  // > cap tells the caller to retry with a larger buffer
    sq_dfree(ctx, dk);
    sq_dfree(ctx, dc);
    sq_dfree(ctx, dn);
    return SQ_OK;
}

extern "C" int sq_overrep_read(sq_overrep *o, uint64_t *kmers, uint32_t *counts, uint64_t *n) {
    // every stored fragment; the caller sized the buffers from collected_unique_fragments
    sq_overrep_info info;
    SQ_TRY(sq_overrep_sync(o, &info));
    return sq_overrep_read_min(o, 0, kmers, counts, info.collected_unique_fragments, n);
}


// ---------------------------------------------------------------------------
// Sharded runs.  The table is "the first
// max_unique_fragments distinct hashes in (sampled read, staging slot) order,
// each with all of its occurrences": order matters only until the table is
// full.  Ranks behind the first one keep their fragments (deferred mode) and
//   - while the table is not full, take it over from the rank before them and
//     apply their fragments to it (the sequential semantics, rank by rank);
//   - once it is full the key set is frozen: every remaining rank loads the
//     keys with zero counts, counts its own fragments (lookup only, any order)
//     and the count arrays are summed.
// ---------------------------------------------------------------------------
extern "C" int sq_overrep_set_deferred(sq_overrep *o, int deferred, uint64_t first_record) {
    if (o->n_seqs != 0 || !o->kept.empty()) {
        sq_set_error("deferred mode must be chosen before the first record array");
        return SQ_E_ARG;
    }
    o->deferred = deferred != 0;
    o->n_seqs = first_record;  // global index of the shard's first read: sampling phase (:3833) and warnings
    return SQ_OK;
}

extern "C" int sq_overrep_apply_deferred(sq_overrep *o) {
    sq_ctx *ctx = o->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    int rc = SQ_OK;
    for (auto &k : o->kept) {
        if (rc == SQ_OK) rc = ov_apply(o, k.frag_hash, k.frag_n, k.n_sampled, k.total, k.fcap);
        sq_dfree(ctx, k.frag_hash);
        sq_dfree(ctx, k.frag_n);
    }
    o->kept.clear();
    o->deferred = false;
    return rc;
}

// keys[table_size] (u64) and counts[table_size] (u32) into caller-owned DEVICE buffers
extern "C" int sq_overrep_copy_table(sq_overrep *o, uint64_t *dev_keys, uint32_t *dev_counts) {
    sq_ctx *ctx = o->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (dev_keys) CUDA_TRY(cudaMemcpyAsync(dev_keys, o->keys, o->table_size * 8, cudaMemcpyDeviceToDevice, sq_cur_stream(ctx)));
    if (dev_counts)
        CUDA_TRY(cudaMemcpyAsync(dev_counts, o->counts, o->table_size * 4, cudaMemcpyDeviceToDevice, sq_cur_stream(ctx)));
    CUDA_TRY(cudaStreamSynchronize(sq_cur_stream(ctx)));
    return SQ_OK;
}

// Replace the table: keys (nullptr: keep), counts (nullptr: zero) from DEVICE buffers, and the
// number of stored keys.
extern "C" int sq_overrep_load_table(sq_overrep *o, const uint64_t *dev_keys, const uint32_t *dev_counts,
                                     uint64_t n_unique) {
    sq_ctx *ctx = o->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (dev_keys) CUDA_TRY(cudaMemcpyAsync(o->keys, dev_keys, o->table_size * 8, cudaMemcpyDeviceToDevice, sq_cur_stream(ctx)));
    if (dev_counts)
        CUDA_TRY(cudaMemcpyAsync(o->counts, dev_counts, o->table_size * 4, cudaMemcpyDeviceToDevice, sq_cur_stream(ctx)));
    else CUDA_TRY(cudaMemsetAsync(o->counts, 0, o->table_size * 4, sq_cur_stream(ctx)));
    unsigned int *h = (unsigned int *)((char *)ctx->h_scratch + 1664);
    *h = (unsigned int)n_unique;
    CUDA_TRY(cudaMemcpyAsync(&o->cnt->n_unique, h, 4, cudaMemcpyHostToDevice, sq_cur_stream(ctx)));
    CUDA_TRY(cudaStreamSynchronize(sq_cur_stream(ctx)));
    o->unique_known = o->unique_upper = n_unique;
    o->full = n_unique >= o->max_unique;
    o->filter_valid = false;  // another key set
    return SQ_OK;
}

// Overwrite the additive counters with merged values (what the members report afterwards).
extern "C" int sq_overrep_set_counters(sq_overrep *o, uint64_t number_of_sequences, uint64_t sampled_sequences,
                                       uint64_t total_fragments, uint64_t warn_records, uint64_t first_warn_record) {
    sq_ctx *ctx = o->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    o->n_seqs = number_of_sequences;
    o->n_sampled = sampled_sequences;
    OvCounters *h = (OvCounters *)((char *)ctx->h_scratch + 1536);
    CUDA_TRY(cudaMemcpyAsync(h, o->cnt, sizeof(OvCounters), cudaMemcpyDeviceToHost, sq_cur_stream(ctx)));
    CUDA_TRY(cudaStreamSynchronize(sq_cur_stream(ctx)));
    h->total_frags = total_fragments;
    h->warn_records = warn_records;
    h->first_warn = first_warn_record;
    CUDA_TRY(cudaMemcpyAsync(o->cnt, h, sizeof(OvCounters), cudaMemcpyHostToDevice, sq_cur_stream(ctx)));
    CUDA_TRY(cudaStreamSynchronize(sq_cur_stream(ctx)));
    return SQ_OK;
}
