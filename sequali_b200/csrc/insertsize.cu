// insertsize.cu -- InsertSizeMetrics (reference _qcmodule.c:5571-5611, 5634-5744).
//
//   k_is_overlap   one thread per pair: reverse-complement the first and last
//                  16 bases of read 2 and slide a 16-byte window over read 1
//                  (two 64-bit halves; a half must match case-insensitively,
//                  then <= 1 mismatch on the raw bytes) -> insert size, CTA
//                  histogram, and the adapter remainder of each read (<= 31 B)
//                  hashed with MurmurHash3.
//   adapter tables capped first-come tables (10 000 entries by default) whose
//                  list order is slot order.  Same machinery as dedup.cu:
//                  classify against the table, collect new keys with their
//                  first pair index, admit the first (cap - entries) of them in
//                  pair order, and place them by priority linear probing so the
//                  slot layout equals the sequential one.
#include <math.h>

#include "common.cuh"

constexpr int IS_TPB = 128;
constexpr int IS_HIST = 1024;  // insert sizes below this are counted in shared memory
constexpr uint64_t IS_EMPTY = ~0ULL;

struct IsTable {
    uint64_t *hash = nullptr;   // [size]
    uint64_t *count = nullptr;  // [size], 0 = empty
    uint8_t *key = nullptr;     // [size][32]: length byte + up to 31 adapter bytes
    uint64_t *prio = nullptr;   // [size]
};

struct IsCounters {
    unsigned long long n_ad[2];
    unsigned long long max_insert;
    unsigned int n_new[2];
    unsigned int admitted[2];
};

struct sq_insert {
    sq_ctx *ctx = nullptr;
    uint64_t max_adapters = 0, table_size = 0, total = 0, n_pairs_base = 0;
    uint64_t entries[2] = {0, 0};
    uint64_t sizes_cap = 0;
    uint64_t *sizes = nullptr;
    IsTable tab[2];
    IsCounters *cnt = nullptr;
    // sharded runs (rank > 0): the adapter tables admit first come, first served (:5588-5610), so they live
    // on the first rank; the other ranks keep their adapter occurrences (32-byte key + hash, pair order)
    bool deferred = false;
    struct Kept { uint8_t *keys; uint64_t *hashes; uint64_t n; };
    std::vector<Kept> kept[2];
};

__device__ __forceinline__ uint8_t comp_upper(uint8_t c) {
    switch (c | 0x20) {
        case 'a': return 'T';
        case 'c': return 'G';
        case 'g': return 'C';
        case 't': return 'A';
    }
    return 0;  // never equals a read letter (:5614-5630)
}
__device__ __forceinline__ uint32_t nonzero_bytes(uint64_t x) {
    uint64_t t = (((x & 0x7F7F7F7F7F7F7F7FULL) + 0x7F7F7F7F7F7F7F7FULL) | x) & 0x8080808080808080ULL;
    return __popcll(t);
}

__global__ void __launch_bounds__(IS_TPB)
k_is_overlap(BatchView b1, BatchView b2, uint32_t *__restrict__ ins_out, uint64_t *__restrict__ h_out /*[2][n]*/,
             uint8_t *__restrict__ len_out /*[2][n]*/, uint64_t *g_sizes, IsCounters *cnt) {
    __shared__ uint32_t s_hist[IS_HIST];
    for (uint32_t i = threadIdx.x; i < IS_HIST; i += IS_TPB) s_hist[i] = 0;
    __syncthreads();
    const uint32_t n = b1.n;
    uint32_t local_max = 0, ad1 = 0, ad2 = 0;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const uint8_t *s1 = b1.text + b1.seq_off[r], *s2 = b2.text + b2.seq_off[r];
        const uint32_t L1 = b1.seq_len[r], L2 = b2.seq_len[r];
        uint32_t ins = 0;
        if (L1 >= 16 && L2 >= 16) {
            uint64_t head_lo = 0, head_hi = 0, tail_lo = 0, tail_hi = 0;
#pragma unroll
            for (int i = 0; i < 16; i++) {  // needle byte 15-i = complement of read-2 byte i
                uint64_t hb = comp_upper(s2[i]), tb = comp_upper(s2[L2 - 16 + i]);
                int pos = 15 - i;
                if (pos < 8) { head_lo |= hb << (8 * pos); tail_lo |= tb << (8 * pos); }
                else { head_hi |= hb << (8 * (pos - 8)); tail_hi |= tb << (8 * (pos - 8)); }
            }
            uint64_t w_lo = 0, w_hi = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                w_lo |= (uint64_t)s1[i] << (8 * i);
                w_hi |= (uint64_t)s1[8 + i] << (8 * i);
            }
            const uint64_t UP = 0xDFDFDFDFDFDFDFDFULL;
            for (uint32_t i = 0;; i++) {
                const uint64_t u_lo = w_lo & UP, u_hi = w_hi & UP;
                if ((u_lo == head_lo || u_hi == head_hi) &&
                    nonzero_bytes(w_lo ^ head_lo) + nonzero_bytes(w_hi ^ head_hi) <= 1) {
                    ins = i + 16;
                    break;
                }
                if ((u_lo == tail_lo || u_hi == tail_hi) &&
                    nonzero_bytes(w_lo ^ tail_lo) + nonzero_bytes(w_hi ^ tail_hi) <= 1) {
                    ins = i + L2;
                    break;
                }
                if (i + 16 >= L1) break;
                w_lo = (w_lo >> 8) | (w_hi << 56);
                w_hi = (w_hi >> 8) | ((uint64_t)s1[i + 16] << 56);
            }
        }
        ins_out[r] = ins;
        if (ins < IS_HIST) atomicAdd(&s_hist[ins], 1u);
        else atomic_add_u64(g_sizes + ins, 1);
        local_max = max(local_max, ins);
        uint8_t l1 = 0, l2 = 0;
        uint64_t h1 = 0, h2 = 0;
        if (ins) {
            if (L1 > ins) {
                l1 = (uint8_t)min(L1 - ins, 31u);
                const uint8_t *a = s1 + ins;
                h1 = murmur3_h2([&](uint64_t i) { return a[i]; }, l1, 0);
                ad1++;
            }
            if (L2 > ins) {
                l2 = (uint8_t)min(L2 - ins, 31u);
                const uint8_t *a = s2 + ins;
                h2 = murmur3_h2([&](uint64_t i) { return a[i]; }, l2, 0);
                ad2++;
            }
        }
        h_out[r] = h1;
        h_out[n + r] = h2;
        len_out[r] = l1;
        len_out[n + r] = l2;
    }
    local_max = warp_max_u32(local_max);
    ad1 = warp_sum_u32(ad1);
    ad2 = warp_sum_u32(ad2);
    if (lane_id() == 0) {
        if (local_max) atomicMax(&cnt->max_insert, (unsigned long long)local_max);
        if (ad1) atomicAdd(&cnt->n_ad[0], (unsigned long long)ad1);
        if (ad2) atomicAdd(&cnt->n_ad[1], (unsigned long long)ad2);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < IS_HIST; i += IS_TPB)
        if (s_hist[i]) atomic_add_u64(g_sizes + i, s_hist[i]);
}

// ---- adapter tables -----------------------------------------------------------------------
struct IsScratch {
    uint64_t *key;    // IS_EMPTY free
    uint32_t *first;  // first pair index
    uint32_t *cnt;    // occurrences
    uint32_t mask;
};
__device__ __forceinline__ uint32_t is_scratch_slot(const IsScratch &S, uint64_t h, bool insert) {
    uint32_t i = (uint32_t)(h ^ (h >> 31)) & S.mask;
    for (;;) {
        uint64_t k = S.key[i];
        if (k == h) return i;
        if (k == IS_EMPTY) {
            if (!insert) return 0xFFFFFFFFu;
            uint64_t old = atomicCAS((unsigned long long *)&S.key[i], IS_EMPTY, (unsigned long long)h);
            if (old == IS_EMPTY || old == h) return i;
        }
        i = (i + 1) & S.mask;
    }
}

// adapter bytes of pair r for read `which`
__device__ __forceinline__ const uint8_t *adapter_ptr(const BatchView &b, uint32_t r, uint32_t ins) {
    // (a list of 32-byte keys instead of a record array: sq_insert_add_keys)
    return b.seq_off ? b.text + b.seq_off[r] + ins : b.text + (size_t)r * 32 + 1;
}

// cls[r]: 0xFFFFFFFF none / 0xFFFFFFFE new / slot of the existing entry (:5588-5610)
__global__ void __launch_bounds__(256)
k_is_classify(BatchView b, const uint32_t *__restrict__ ins, const uint64_t *__restrict__ hs,
              const uint8_t *__restrict__ lens, IsTable T, uint64_t tmask, uint32_t *__restrict__ cls,
              IsScratch S, int collect_new) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < b.n; r += gridDim.x * blockDim.x) {
        const uint32_t len = lens[r];
        if (!len) {
            cls[r] = 0xFFFFFFFFu;
            continue;
        }
        const uint64_t h = hs[r];
        const uint8_t *a = adapter_ptr(b, r, ins[r]);
        uint64_t i = h & tmask;
        uint32_t c = 0xFFFFFFFEu;
        for (;;) {
            if (T.hash[i] == h && T.count[i] != 0) {
                const uint8_t *k = T.key + i * 32;
                bool same = k[0] == len;
                for (uint32_t j = 0; j < len && same; j++) same = k[1 + j] == a[j];
                if (same) {
                    c = (uint32_t)i;
                    break;
                }
            }
            else if (T.count[i] == 0) break;
            i = (i + 1) & tmask;
        }
        cls[r] = c;
        if (c == 0xFFFFFFFEu && collect_new) {
            uint32_t s = is_scratch_slot(S, h, true);
            atomicMin(&S.first[s], r);
            atomicAdd(&S.cnt[s], 1u);
        }
    }
}
__global__ void __launch_bounds__(256)
k_is_count_existing(const uint32_t *__restrict__ cls, uint32_t n, IsTable T) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x)
        if (cls[r] < 0xFFFFFFFEu) atomic_add_u64(&T.count[cls[r]], 1);
}
__global__ void __launch_bounds__(256)
k_is_flags(const uint64_t *__restrict__ hs, const uint32_t *__restrict__ cls, uint32_t n, IsScratch S,
           uint32_t *__restrict__ flag, unsigned int *n_new) {
    uint32_t local = 0;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        uint32_t f = 0;
        if (cls[r] == 0xFFFFFFFEu) f = S.first[is_scratch_slot(S, hs[r], false)] == r;
        flag[r] = f;
        local += f;
    }
    local = warp_sum_u32(local);
    if (lane_id() == 0 && local) atomicAdd(n_new, local);
}
__device__ __forceinline__ void is_prio_insert(uint64_t *prio, uint64_t tmask, uint64_t home, uint64_t word) {
    uint64_t i = home;
    for (;;) {
        uint64_t old = atomicMin((unsigned long long *)&prio[i], (unsigned long long)word);
        if (old == IS_EMPTY) return;
        if (old > word) word = old;
        i = (i + 1) & tmask;
    }
}
// first occurrences ranked below K enter the table
__global__ void __launch_bounds__(256)
k_is_admit(const uint64_t *__restrict__ hs, const uint32_t *__restrict__ flag, const uint32_t *__restrict__ rank,
           uint32_t n, uint32_t K, IsTable T, uint64_t tmask, uint64_t prio_base, unsigned int *admitted) {
    uint32_t local = 0;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x)
        if (flag[r] && rank[r] < K) {
            is_prio_insert(T.prio, tmask, hs[r] & tmask, prio_base + r);
            local++;
        }
    local = warp_sum_u32(local);
    if (lane_id() == 0 && local) atomicAdd(admitted, local);
}
__global__ void __launch_bounds__(256)
k_is_place(BatchView b, const uint32_t *__restrict__ ins, const uint64_t *__restrict__ hs,
           const uint8_t *__restrict__ lens, const uint32_t *__restrict__ flag, const uint32_t *__restrict__ rank,
           uint32_t K, IsTable T, uint64_t tmask, uint64_t prio_base, IsScratch S) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < b.n; r += gridDim.x * blockDim.x)
        if (flag[r] && rank[r] < K) {
            const uint64_t h = hs[r], word = prio_base + r;
            uint64_t i = h & tmask;
            while (T.prio[i] != word) i = (i + 1) & tmask;
            T.hash[i] = h;
            T.count[i] = S.cnt[is_scratch_slot(S, h, false)];
            uint8_t *k = T.key + i * 32;
            const uint8_t *a = adapter_ptr(b, r, ins[r]);
            const uint32_t len = lens[r];
            k[0] = (uint8_t)len;
            for (uint32_t j = 0; j < 31; j++) k[1 + j] = j < len ? a[j] : 0;
        }
}

// ---------------------------------------------------------------------------
static int is_table_alloc(sq_ctx *ctx, IsTable *t, uint64_t size) {
    SQ_TRY(sq_dalloc(ctx, (void **)&t->hash, size * 8, true));
    SQ_TRY(sq_dalloc(ctx, (void **)&t->count, size * 8, true));
    SQ_TRY(sq_dalloc(ctx, (void **)&t->key, size * 32, true));
    SQ_TRY(sq_dalloc(ctx, (void **)&t->prio, size * 8, false));
    CUDA_TRY(cudaMemsetAsync(t->prio, 0xFF, size * 8, ctx->stream));
    return SQ_OK;
}

extern "C" int sq_insert_create(sq_ctx *ctx, uint64_t max_adapters, sq_insert **out) {
    *out = nullptr;
    if (max_adapters < 1) {
        sq_set_error("max_adapters must be at least 1");
        return SQ_E_ARG;
    }
    CUDA_TRY(cudaSetDevice(ctx->device));
    sq_insert *m = new sq_insert();
    m->ctx = ctx;
    m->max_adapters = max_adapters;
    uint64_t bits = (uint64_t)(log2((double)max_adapters * 1.5) + 1);  // :5525
    m->table_size = 1ULL << bits;
    int rc = is_table_alloc(ctx, &m->tab[0], m->table_size);
    if (rc == SQ_OK) rc = is_table_alloc(ctx, &m->tab[1], m->table_size);
    if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&m->cnt, sizeof(IsCounters), true);
    m->sizes_cap = IS_HIST + 1;
    if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&m->sizes, m->sizes_cap * 8, true);
    if (rc != SQ_OK) {
        sq_insert_destroy(m);
        return rc;
    }
    *out = m;
    return SQ_OK;
}

static void is_free_kept(sq_insert *m) {
    for (int w = 0; w < 2; w++) {
        for (auto &k : m->kept[w]) {
            sq_dfree(m->ctx, k.keys);
            sq_dfree(m->ctx, k.hashes);
        }
        m->kept[w].clear();
    }
}

extern "C" void sq_insert_destroy(sq_insert *m) {
    if (!m) return;
    is_free_kept(m);
    cudaSetDevice(m->ctx->device);
    for (int w = 0; w < 2; w++) {
        sq_dfree(m->ctx, m->tab[w].hash);
        sq_dfree(m->ctx, m->tab[w].count);
        sq_dfree(m->ctx, m->tab[w].key);
        sq_dfree(m->ctx, m->tab[w].prio);
    }
    sq_dfree(m->ctx, m->cnt);
    sq_dfree(m->ctx, m->sizes);
    delete m;
}

// deferred mode: occurrences (len > 0) of one read side, packed in pair order
__global__ void __launch_bounds__(256)
k_is_has_adapter(const uint8_t *__restrict__ lens, uint32_t n, uint32_t *__restrict__ flag) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) flag[r] = lens[r] != 0;
}
__global__ void __launch_bounds__(256)
k_is_pack_keys(BatchView b, const uint32_t *__restrict__ ins, const uint64_t *__restrict__ hs, const uint8_t *__restrict__ lens,
               const uint32_t *__restrict__ rank, uint8_t *__restrict__ keys, uint64_t *__restrict__ hashes) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < b.n; r += gridDim.x * blockDim.x) {
        const uint32_t len = lens[r];
        if (!len) continue;
        const uint32_t j = rank[r];
        const uint8_t *a = adapter_ptr(b, r, ins[r]);
        uint8_t *k = keys + (size_t)j * 32;
        k[0] = (uint8_t)len;
        for (uint32_t i = 0; i < 31; i++) k[1 + i] = i < len ? a[i] : 0;
        hashes[j] = hs[r];
    }
}
__global__ void __launch_bounds__(256)
k_is_key_lens(const uint8_t *__restrict__ keys, uint32_t n, uint8_t *__restrict__ lens) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) lens[r] = keys[(size_t)r * 32];
}

// Table maintenance of read side `w` for n occurrences in arrival order (:5571-5611): count what is stored,
// admit the first (cap - entries) new keys in order, place them by priority probing.
static int is_table_update(sq_insert *m, int w, const BatchView &bv, uint32_t n, const uint32_t *ins, const uint64_t *hs,
                           const uint8_t *lens, uint32_t *cls, uint32_t *flag, uint32_t *rank, IsScratch &S, uint32_t scap) {
    sq_ctx *ctx = m->ctx;
    const uint64_t tmask = m->table_size - 1;
    const uint64_t prio_base = 1 + m->n_pairs_base;
    const int grid = sq_grid_for(ctx, n, 256, 16);
    IsCounters *hc = (IsCounters *)((char *)ctx->h_scratch + 3200);
    const bool full = m->entries[w] >= m->max_adapters;
    if (!full && !S.key) {
        SQ_TRY(sq_dalloc(ctx, (void **)&S.key, (size_t)scap * 8, false));
        SQ_TRY(sq_dalloc(ctx, (void **)&S.first, (size_t)scap * 4, false));
        SQ_TRY(sq_dalloc(ctx, (void **)&S.cnt, (size_t)scap * 4, false));
    }
    if (!full) {
        CUDA_TRY(cudaMemsetAsync(S.key, 0xFF, (size_t)scap * 8, ctx->stream));
        CUDA_TRY(cudaMemsetAsync(S.first, 0xFF, (size_t)scap * 4, ctx->stream));
        CUDA_TRY(cudaMemsetAsync(S.cnt, 0, (size_t)scap * 4, ctx->stream));
    }
    SQ_LAUNCH(ctx, k_is_classify, grid, 256, 0, bv, ins, hs, lens, m->tab[w], tmask, cls, S, full ? 0 : 1);
    SQ_LAUNCH(ctx, k_is_count_existing, grid, 256, 0, cls, n, m->tab[w]);
    if (full) return SQ_OK;
    CUDA_TRY(cudaMemsetAsync(&m->cnt->n_new[w], 0, 4, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(&m->cnt->admitted[w], 0, 4, ctx->stream));
    SQ_LAUNCH(ctx, k_is_flags, grid, 256, 0, hs, cls, n, S, flag, &m->cnt->n_new[w]);
    SQ_TRY(sq_scan_exclusive_u32(ctx, flag, rank, n, nullptr));
    const uint32_t K = (uint32_t)(m->max_adapters - m->entries[w]);
    SQ_LAUNCH(ctx, k_is_admit, grid, 256, 0, hs, flag, rank, n, K, m->tab[w], tmask, prio_base, &m->cnt->admitted[w]);
    SQ_LAUNCH(ctx, k_is_place, grid, 256, 0, bv, ins, hs, lens, flag, rank, K, m->tab[w], tmask, prio_base, S);
    CUDA_TRY(cudaMemcpyAsync(hc, m->cnt, sizeof(IsCounters), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    m->entries[w] += hc->admitted[w];
    return SQ_OK;
}

extern "C" int sq_insert_add_pair(sq_insert *m, sq_batch *b1, sq_batch *b2) {
    sq_ctx *ctx = m->ctx;
    if (b1->ctx != ctx || b2->ctx != ctx) {
        sq_set_error("record array belongs to another context");
        return SQ_E_ARG;
    }
    if (b1->n != b2->n) {
        sq_set_error("record_array1 and record_array2 must be of the same size");
        return SQ_E_ARG;
    }
    const uint32_t n = (uint32_t)b1->n;
    if (n == 0) return SQ_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    // insert sizes are at most max(L1) + max(L2)
    uint64_t need = (uint64_t)b1->max_len + b2->max_len + 1;
    if (need < IS_HIST + 1) need = IS_HIST + 1;
    if (need > m->sizes_cap) {
        uint64_t *ns = nullptr;
        SQ_TRY(sq_dalloc(ctx, (void **)&ns, need * 8, true));
        CUDA_TRY(cudaMemcpyAsync(ns, m->sizes, m->sizes_cap * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        sq_dfree(ctx, m->sizes);
        m->sizes = ns;
        m->sizes_cap = need;
    }
    uint32_t *ins = nullptr, *cls = nullptr, *flag = nullptr, *rank = nullptr;
    uint64_t *hs = nullptr;
    uint8_t *lens = nullptr;
    SQ_TRY(sq_dalloc(ctx, (void **)&ins, (size_t)n * 4, false));
    SQ_TRY(sq_dalloc(ctx, (void **)&hs, (size_t)n * 16, false));
    SQ_TRY(sq_dalloc(ctx, (void **)&lens, (size_t)n * 2, false));
    SQ_TRY(sq_dalloc(ctx, (void **)&cls, (size_t)n * 4, false));
    SQ_TRY(sq_dalloc(ctx, (void **)&flag, (size_t)n * 4, false));
    SQ_TRY(sq_dalloc(ctx, (void **)&rank, (size_t)n * 4, false));
    SQ_LAUNCH(ctx, k_is_overlap, sq_grid_for(ctx, n, IS_TPB, 16), IS_TPB, 0, b1->view(), b2->view(), ins, hs, lens,
              m->sizes, m->cnt);
    IsScratch S;
    uint32_t scap = 1024;
    while (scap < 2 * (uint64_t)n) scap <<= 1;
    S.mask = scap - 1;
    S.key = nullptr;
    S.first = S.cnt = nullptr;
    int rc = SQ_OK;
    for (int w = 0; w < 2 && rc == SQ_OK; w++) {
        sq_batch *b = w == 0 ? b1 : b2;
        if (m->deferred) {
            // keep the occurrences of this side, packed in pair order, for the rank that owns the tables
            const int grid = sq_grid_for(ctx, n, 256, 16);
            uint32_t *d_total = (uint32_t *)((char *)ctx->d_scratch + 3300);
            uint32_t *h_total = (uint32_t *)((char *)ctx->h_scratch + 3300);
            SQ_LAUNCH(ctx, k_is_has_adapter, grid, 256, 0, lens + (size_t)w * n, n, flag);
            rc = sq_scan_exclusive_u32(ctx, flag, rank, n, d_total);
            if (rc != SQ_OK) break;
            CUDA_TRY(cudaMemcpyAsync(h_total, d_total, 4, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            const uint64_t k = *h_total;
            if (k) {
                sq_insert::Kept kp = {nullptr, nullptr, k};
                rc = sq_dalloc(ctx, (void **)&kp.keys, k * 32, false);
                if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&kp.hashes, k * 8, false);
                if (rc != SQ_OK) break;
                SQ_LAUNCH(ctx, k_is_pack_keys, grid, 256, 0, b->view(), ins, hs + (size_t)w * n, lens + (size_t)w * n, rank,
                          kp.keys, kp.hashes);
                m->kept[w].push_back(kp);
            }
            continue;
        }
        rc = is_table_update(m, w, b->view(), n, ins, hs + (size_t)w * n, lens + (size_t)w * n, cls, flag, rank, S, scap);
    }
    m->total += n;
    m->n_pairs_base += n;
    sq_dfree(ctx, ins);
    sq_dfree(ctx, hs);
    sq_dfree(ctx, lens);
    sq_dfree(ctx, cls);
    sq_dfree(ctx, flag);
    sq_dfree(ctx, rank);
    sq_dfree(ctx, S.key);
    sq_dfree(ctx, S.first);
    sq_dfree(ctx, S.cnt);
    return rc;
}

extern "C" int sq_insert_sync(sq_insert *m, sq_insert_info *info) {
    sq_ctx *ctx = m->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    IsCounters *hc = (IsCounters *)((char *)ctx->h_scratch + 3200);
    CUDA_TRY(cudaMemcpyAsync(hc, m->cnt, sizeof(IsCounters), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    info->total_reads = m->total;
    info->number_of_adapters_read1 = hc->n_ad[0];
    info->number_of_adapters_read2 = hc->n_ad[1];
    info->max_insert_size = hc->max_insert;
    info->entries_read1 = m->entries[0];
    info->entries_read2 = m->entries[1];
    return SQ_OK;
}

extern "C" int sq_insert_read_sizes(sq_insert *m, uint64_t *sizes) {
    sq_ctx *ctx = m->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    IsCounters *hc = (IsCounters *)((char *)ctx->h_scratch + 3200);
    CUDA_TRY(cudaMemcpyAsync(hc, m->cnt, sizeof(IsCounters), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(sizes, m->sizes, (hc->max_insert + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return SQ_OK;
}

extern "C" int sq_insert_read_adapters(sq_insert *m, int which, uint8_t *seqs, uint64_t *counts, uint64_t *n) {
    sq_ctx *ctx = m->ctx;
    if (which < 0 || which > 1) {
        sq_set_error("which must be 0 or 1");
        return SQ_E_ARG;
    }
    CUDA_TRY(cudaSetDevice(ctx->device));
    std::vector<uint64_t> hc(m->table_size);
    std::vector<uint8_t> hk(m->table_size * 32);
    CUDA_TRY(cudaMemcpyAsync(hc.data(), m->tab[which].count, m->table_size * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(hk.data(), m->tab[which].key, m->table_size * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    uint64_t w = 0;
    for (uint64_t i = 0; i < m->table_size; i++)  // slot order (:5894-5910)
        if (hc[i]) {
            memcpy(seqs + w * 32, hk.data() + i * 32, 32);
            counts[w++] = hc[i];
        }
    *n = w;
    return SQ_OK;
}

// ---------------------------------------------------------------------------
// sharded runs: the insert size histogram and the counters are sums; the two adapter
// tables admit first come, first served, so the first rank owns them and the other ranks hand their
// adapter occurrences over in pair order (32-byte key + hash each; a few percent of the pairs).
// ---------------------------------------------------------------------------
extern "C" int sq_insert_set_deferred(sq_insert *m, int deferred) {
    m->deferred = deferred != 0;
    return SQ_OK;
}
extern "C" int sq_insert_deferred_count(sq_insert *m, int which, uint64_t *n) {
    uint64_t t = 0;
    for (auto &k : m->kept[which & 1]) t += k.n;
    *n = t;
    return SQ_OK;
}
// the kept occurrences of read side `which`, concatenated in pair order, into caller-owned DEVICE buffers
extern "C" int sq_insert_deferred_fetch(sq_insert *m, int which, uint8_t *dev_keys, uint64_t *dev_hashes) {
    sq_ctx *ctx = m->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    uint64_t at = 0;
    for (auto &k : m->kept[which & 1]) {
        CUDA_TRY(cudaMemcpyAsync(dev_keys + at * 32, k.keys, k.n * 32, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(dev_hashes + at, k.hashes, k.n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        at += k.n;
    }
    return SQ_OK;
}
// the owner: n occurrences (arrival order) for the table of read side `which`
extern "C" int sq_insert_add_keys(sq_insert *m, int which, const uint8_t *dev_keys, const uint64_t *dev_hashes, uint64_t n) {
    sq_ctx *ctx = m->ctx;
    if (n == 0) return SQ_OK;
    if (n >= 0xFFFFFFFFULL || which < 0 || which > 1) {
        sq_set_error("sq_insert_add_keys: bad arguments");
        return SQ_E_ARG;
    }
    CUDA_TRY(cudaSetDevice(ctx->device));
    const uint32_t n32 = (uint32_t)n;
    uint32_t *ins = nullptr, *cls = nullptr, *flag = nullptr, *rank = nullptr;
    uint8_t *lens = nullptr;
    int rc = sq_dalloc(ctx, (void **)&ins, n * 4, true);
    if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&cls, n * 4, false);
    if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&flag, n * 4, false);
    if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&rank, n * 4, false);
    if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&lens, n, false);
    IsScratch S;
    uint32_t scap = 1024;
    while (scap < 2 * n) scap <<= 1;
    S.mask = scap - 1;
    S.key = nullptr;
    S.first = S.cnt = nullptr;
    if (rc == SQ_OK) {
        BatchView bv;
        memset(&bv, 0, sizeof(bv));
        bv.text = dev_keys;  // seq_off == nullptr: adapter_ptr reads the 32-byte keys
        bv.n = n32;
        SQ_LAUNCH(ctx, k_is_key_lens, sq_grid_for(ctx, n, 256, 16), 256, 0, dev_keys, n32, lens);
        rc = is_table_update(m, which, bv, n32, ins, dev_hashes, lens, cls, flag, rank, S, scap);
        m->n_pairs_base += n;
    }
    void *ptrs[] = {ins, cls, flag, rank, lens, S.key, S.first, S.cnt};
    for (void *q : ptrs) sq_dfree(ctx, q);
    return rc;
}
// histogram, counters: sums over the ranks, in place on the device
extern "C" int sq_insert_allreduce(sq_insert *m, sq_comm *c) {
    sq_ctx *ctx = m->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    IsCounters *hc = (IsCounters *)((char *)ctx->h_scratch + 3200);
    CUDA_TRY(cudaMemcpyAsync(hc, m->cnt, sizeof(IsCounters), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    uint64_t mx[2] = {hc->max_insert, m->sizes_cap};
    SQ_TRY(sq_comm_allreduce_host_u64(c, mx, 2, 1));
    if (mx[1] > m->sizes_cap) {
        uint64_t *ns = nullptr;
        SQ_TRY(sq_dalloc(ctx, (void **)&ns, mx[1] * 8, true));
        CUDA_TRY(cudaMemcpyAsync(ns, m->sizes, m->sizes_cap * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        sq_dfree(ctx, m->sizes);
        m->sizes = ns;
        m->sizes_cap = mx[1];
    }
    uint64_t sums[3] = {hc->n_ad[0], hc->n_ad[1], m->total};
    SQ_TRY(sq_comm_allreduce_host_u64(c, sums, 3, 0));
    SQ_TRY(sq_comm_allreduce_u64(c, m->sizes, mx[0] + 1, 0));
    hc->n_ad[0] = sums[0];
    hc->n_ad[1] = sums[1];
    hc->max_insert = mx[0];
    CUDA_TRY(cudaMemcpyAsync(m->cnt, hc, 24, cudaMemcpyHostToDevice, ctx->stream));  // n_ad[2], max_insert
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    m->total = sums[2];
    return SQ_OK;
}
