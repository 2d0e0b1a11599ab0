// pertile.cu -- PerTileQuality (reference _qcmodule.c:3089-3222, 3307-3359).
//
// total_errors[tile][pos] is a plain double sum over the reads of a tile IN READ
// ORDER, so it cannot be reduced in arbitrary order.  Per record array:
//   k_pt_tile        tile id (decimal between the 4th and 5th ':') per read, or
//                    the first unparsable header by atomicMin (module "skipped")
//   k_pt_map_*       tile id -> dense slot through a small device hash map
//   stable radix sort of (slot, read index)   groups reads by tile, order kept
//   k_pt_accumulate  one thread per (tile, position) walks that tile's reads in
//                    order and extends the chain that lives in HBM; consecutive
//                    threads read consecutive quality bytes of the same read
//   k_pt_lengths     length_counts[tile][L-1]++ (order free)
#include <algorithm>

#include "common.cuh"

constexpr int PT_TPB = 256;
constexpr uint64_t PT_EMPTY = ~0ULL;
constexpr uint32_t PT_MAP_CAP = 1u << 18;   // open addressing, <= 2^17 distinct tiles
constexpr uint32_t PT_NONE = 0xFFFFFFFFu;

struct PtState {  // device
    unsigned long long fail_idx;   // global index of the first header without a tile id
    unsigned long long err_key;    // (global record << 8 | byte) of an invalid phred
    unsigned int n_slots;
    unsigned int max_len;          // longest kept read
    unsigned long long n_kept;     // kept reads (before fail_idx)
};

struct sq_pertile {
    sq_ctx *ctx = nullptr;
    uint64_t n_added = 0;
    uint64_t slot_cap = 0, len_cap = 0;
    uint64_t *map_keys = nullptr;  // [PT_MAP_CAP]
    uint32_t *map_vals = nullptr;
    uint64_t *slot_tile = nullptr;  // [slot_cap]
    double *errors = nullptr;       // [slot_cap][len_cap]
    uint64_t *lengths = nullptr;    // [slot_cap][len_cap]
    PtState *st = nullptr;
    bool skipped = false;
    uint64_t skipped_record = 0;
    std::vector<uint8_t> skipped_name;
    // host mirror after the last sync
    uint64_t n_slots = 0, max_len = 0;
};

__device__ long long tile_id_of(const uint8_t *h, uint32_t n) {  // :3089-3121, :160-180
    uint32_t i = 0, colons = 0;
    for (; i < n; i++)
        if (h[i] == ':' && ++colons == 4) break;
    uint32_t start = i + 1, j = start;
    for (; j < n; j++)
        if (h[j] == ':') break;
    if (j >= n) return -1;
    uint32_t len = j - start;
    if (len < 1 || len > 18) return -1;
    long long v = 0;
    for (uint32_t k = start; k < j; k++) {
        uint32_t d = (uint32_t)h[k] - '0';
        if (d > 9) return -1;
        v = v * 10 + d;
    }
    return v;
}

__global__ void __launch_bounds__(PT_TPB)
k_pt_tile(BatchView bv, long long *__restrict__ tile, uint64_t base, PtState *st) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < bv.n; r += gridDim.x * blockDim.x) {
        const uint32_t nl = bv.name_len ? bv.name_len[r] : bv.seq_off[r] - 1 - bv.name_off[r];
        long long t = tile_id_of(bv.text + bv.name_off[r], nl);
        tile[r] = t;
        if (t < 0) atomicMin(&st->fail_idx, (unsigned long long)(base + r));
    }
}

__device__ __forceinline__ uint32_t pt_hash(uint64_t t) {
    t *= 0x9E3779B97F4A7C15ULL;
    return (uint32_t)(t >> 40);
}
__global__ void __launch_bounds__(PT_TPB)
k_pt_map_insert(const long long *__restrict__ tile, uint32_t n, uint64_t base, uint64_t *map_keys,
                uint32_t *map_vals, uint64_t *slot_tile, uint64_t slot_cap, PtState *st) {
    const unsigned long long fail = st->fail_idx;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        if (base + r >= fail) continue;
        const uint64_t t = (uint64_t)tile[r];
        uint32_t i = pt_hash(t) & (PT_MAP_CAP - 1);
        for (;;) {
            uint64_t k = map_keys[i];
            if (k == t) break;
            if (k == PT_EMPTY) {
                uint64_t old = atomicCAS((unsigned long long *)&map_keys[i], PT_EMPTY, (unsigned long long)t);
                if (old == PT_EMPTY) {
                    uint32_t s = atomicAdd(&st->n_slots, 1u);
                    map_vals[i] = s;
                    if (s < slot_cap) slot_tile[s] = t;
                    break;
                }
                if (old == t) break;
            }
            i = (i + 1) & (PT_MAP_CAP - 1);
        }
    }
}
__global__ void __launch_bounds__(PT_TPB)
k_pt_map_lookup(BatchView bv, const long long *__restrict__ tile, uint64_t base, const uint64_t *map_keys,
                const uint32_t *map_vals, uint32_t *__restrict__ slot, uint32_t *__restrict__ idx, PtState *st) {
    const unsigned long long fail = st->fail_idx;
    uint32_t lmax = 0, kept = 0;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < bv.n; r += gridDim.x * blockDim.x) {
        idx[r] = r;
        if (base + r >= fail) {
            slot[r] = PT_NONE;
            continue;
        }
        const uint64_t t = (uint64_t)tile[r];
        uint32_t i = pt_hash(t) & (PT_MAP_CAP - 1);
        while (map_keys[i] != t) i = (i + 1) & (PT_MAP_CAP - 1);
        slot[r] = map_vals[i];
        lmax = max(lmax, bv.seq_len[r]);
        kept++;
    }
    lmax = warp_max_u32(lmax);
    kept = warp_sum_u32(kept);
    if (lane_id() == 0) {
        if (lmax) atomicMax(&st->max_len, lmax);
        if (kept) atomicAdd(&st->n_kept, (unsigned long long)kept);
    }
}

__global__ void __launch_bounds__(PT_TPB)
k_pt_lengths(BatchView bv, const uint32_t *__restrict__ slot, uint64_t *lengths, uint64_t len_cap) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < bv.n; r += gridDim.x * blockDim.x) {
        const uint32_t s = slot[r], L = bv.seq_len[r];
        if (s != PT_NONE && L) atomic_add_u64(lengths + (uint64_t)s * len_cap + (L - 1), 1);
    }
}

// after the sort: seg_lo/seg_hi[slot] = range of that tile's reads in `order`
__global__ void __launch_bounds__(PT_TPB)
k_pt_segments(const uint32_t *__restrict__ sorted_slot, uint32_t n, uint32_t *seg_lo, uint32_t *seg_hi) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t s = sorted_slot[i];
        if (s == PT_NONE) continue;
        if (i == 0 || sorted_slot[i - 1] != s) seg_lo[s] = i;
        if (i + 1 == n || sorted_slot[i + 1] != s) seg_hi[s] = i + 1;
    }
}

__global__ void __launch_bounds__(PT_TPB)
k_pt_accumulate(BatchView bv, const uint32_t *__restrict__ order, const uint32_t *__restrict__ seg_lo,
                const uint32_t *__restrict__ seg_hi, uint32_t n_slots, uint32_t width, double *errors,
                uint64_t len_cap, const double *__restrict__ err_tab, uint64_t base, PtState *st) {
    __shared__ double s_err[94];
    for (uint32_t i = threadIdx.x; i < 94; i += PT_TPB) s_err[i] = err_tab[i];
    __syncthreads();
    const uint64_t total = (uint64_t)n_slots * width;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t s = (uint32_t)(t / width), pos = (uint32_t)(t % width);
        const uint32_t lo = seg_lo[s], hi = seg_hi[s];
        if (lo >= hi) continue;
        double *cell = errors + (uint64_t)s * len_cap + pos;
        double acc = *cell;
        bool touched = false;
        for (uint32_t i = lo; i < hi; i++) {
            const uint32_t r = order[i];
            if (pos >= bv.seq_len[r]) continue;
            const uint8_t c = bv.text[bv.qual_off[r] + pos];
            const uint32_t q = (uint8_t)(c - 33);
            if (q > 93) {
                atomicMin(&st->err_key, (unsigned long long)((base + r) << 8 | c));
                continue;
            }
            acc += s_err[q];  // read order within the tile: the reference's chain (:3199-3219)
            touched = true;
        }
        if (touched) *cell = acc;
    }
}

// ---------------------------------------------------------------------------
extern "C" int sq_pertile_create(sq_ctx *ctx, sq_pertile **out) {
    *out = nullptr;
    CUDA_TRY(cudaSetDevice(ctx->device));
    sq_pertile *p = new sq_pertile();
    p->ctx = ctx;
    int rc = sq_dalloc(ctx, (void **)&p->map_keys, (size_t)PT_MAP_CAP * 8, false);
    if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&p->map_vals, (size_t)PT_MAP_CAP * 4, true);
    if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&p->st, sizeof(PtState), true);
    if (rc == SQ_OK) {
        CUDA_TRY(cudaMemsetAsync(p->map_keys, 0xFF, (size_t)PT_MAP_CAP * 8, ctx->stream));
        CUDA_TRY(cudaMemsetAsync(&p->st->fail_idx, 0xFF, 16, ctx->stream));
    }
    if (rc != SQ_OK) {
        sq_pertile_destroy(p);
        return rc;
    }
    *out = p;
    return SQ_OK;
}

extern "C" void sq_pertile_destroy(sq_pertile *p) {
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    sq_dfree(p->ctx, p->map_keys);
    sq_dfree(p->ctx, p->map_vals);
    sq_dfree(p->ctx, p->slot_tile);
    sq_dfree(p->ctx, p->errors);
    sq_dfree(p->ctx, p->lengths);
    sq_dfree(p->ctx, p->st);
    delete p;
}

// grow [slot_cap][len_cap] tables, keeping content
static int pt_grow(sq_pertile *p, uint64_t slots, uint64_t len) {
    if (slots <= p->slot_cap && len <= p->len_cap) return SQ_OK;
    sq_ctx *ctx = p->ctx;
    uint64_t nsc = p->slot_cap ? p->slot_cap : 1024, nlc = p->len_cap ? p->len_cap : 256;
    while (nsc < slots) nsc *= 2;
    while (nlc < len) nlc *= 2;
    double *ne = nullptr;
    uint64_t *nl = nullptr, *nt = nullptr;
    SQ_TRY(sq_dalloc(ctx, (void **)&ne, nsc * nlc * 8, true));
    SQ_TRY(sq_dalloc(ctx, (void **)&nl, nsc * nlc * 8, true));
    SQ_TRY(sq_dalloc(ctx, (void **)&nt, nsc * 8, true));
    if (p->slot_cap) {
        CUDA_TRY(cudaMemcpy2DAsync(ne, nlc * 8, p->errors, p->len_cap * 8, p->len_cap * 8, p->slot_cap,
                                   cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpy2DAsync(nl, nlc * 8, p->lengths, p->len_cap * 8, p->len_cap * 8, p->slot_cap,
                                   cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(nt, p->slot_tile, p->slot_cap * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    sq_dfree(ctx, p->errors);
    sq_dfree(ctx, p->lengths);
    sq_dfree(ctx, p->slot_tile);
    p->errors = ne;
    p->lengths = nl;
    p->slot_tile = nt;
    p->slot_cap = nsc;
    p->len_cap = nlc;
    return SQ_OK;
}

extern "C" int sq_pertile_add(sq_pertile *p, sq_batch *b) {
    sq_ctx *ctx = p->ctx;
    if (b->ctx != ctx) {
        sq_set_error("record array belongs to another context");
        return SQ_E_ARG;
    }
    if (p->skipped || b->n == 0) return SQ_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const uint32_t n = (uint32_t)b->n;
    const uint64_t base = p->n_added;
    const int grid = sq_grid_for(ctx, n, PT_TPB, 16);
    long long *tile = nullptr;
    uint32_t *slot = nullptr, *idx = nullptr, *tmpk = nullptr, *tmpv = nullptr, *seg = nullptr;
    SQ_TRY(sq_dalloc(ctx, (void **)&tile, (size_t)n * 8, false));
    SQ_TRY(sq_dalloc(ctx, (void **)&slot, (size_t)n * 4, false));
    SQ_TRY(sq_dalloc(ctx, (void **)&idx, (size_t)n * 4, false));
    SQ_TRY(sq_dalloc(ctx, (void **)&tmpk, (size_t)n * 4, false));
    SQ_TRY(sq_dalloc(ctx, (void **)&tmpv, (size_t)n * 4, false));
    SQ_TRY(pt_grow(p, p->n_slots ? p->n_slots : 1, b->max_len ? b->max_len : 1));
    SQ_LAUNCH(ctx, k_pt_tile, grid, PT_TPB, 0, b->view(), tile, base, p->st);
    // slot ids of new tiles may exceed slot_cap: slot_tile writes are guarded, and the
    // map is re-read after growing
    SQ_LAUNCH(ctx, k_pt_map_insert, grid, PT_TPB, 0, tile, n, base, p->map_keys, p->map_vals, p->slot_tile,
              p->slot_cap, p->st);
    SQ_LAUNCH(ctx, k_pt_map_lookup, grid, PT_TPB, 0, b->view(), tile, base, p->map_keys, p->map_vals, slot, idx,
              p->st);
    PtState *h = (PtState *)((char *)ctx->h_scratch + 3400);
    CUDA_TRY(cudaMemcpyAsync(h, p->st, sizeof(PtState), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    int rc = SQ_OK;
    if (h->n_slots > PT_MAP_CAP / 2) {
        sq_set_error("more than %u distinct tiles are not supported", PT_MAP_CAP / 2);
        rc = SQ_E_LIMIT;
    }
    if (rc == SQ_OK && h->n_slots > p->slot_cap) {
        // tile ids of slots beyond the old capacity were not recorded: rebuild slot_tile from the map
        rc = pt_grow(p, h->n_slots, h->max_len ? h->max_len : 1);
        if (rc == SQ_OK) {
            std::vector<uint64_t> mk(PT_MAP_CAP);
            std::vector<uint32_t> mv(PT_MAP_CAP);
            std::vector<uint64_t> st(p->slot_cap, 0);
            SQ_TRY(sq_memcpy_d2h(ctx, mk.data(), p->map_keys, (size_t)PT_MAP_CAP * 8));
            SQ_TRY(sq_memcpy_d2h(ctx, mv.data(), p->map_vals, (size_t)PT_MAP_CAP * 4));
            for (uint32_t i = 0; i < PT_MAP_CAP; i++)
                if (mk[i] != PT_EMPTY) st[mv[i]] = mk[i];
            SQ_TRY(sq_memcpy_h2d(ctx, p->slot_tile, st.data(), p->slot_cap * 8));
        }
    }
    if (rc == SQ_OK) rc = pt_grow(p, h->n_slots ? h->n_slots : 1, h->max_len ? h->max_len : 1);
    p->n_slots = h->n_slots;
    p->max_len = h->max_len;
    if (rc == SQ_OK && h->n_kept && h->n_slots) {
        uint32_t key_bits = 1;
        while ((1ull << key_bits) < h->n_slots) key_bits++;
        // PT_NONE (skipped reads) carries all-ones low bits, so it sorts last within those bits
        SQ_LAUNCH(ctx, k_pt_lengths, grid, PT_TPB, 0, b->view(), slot, p->lengths, p->len_cap);
        rc = sq_radix_sort_pairs(ctx, slot, idx, tmpk, tmpv, n, h->fail_idx != ~0ULL ? 32 : key_bits);
        if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&seg, (size_t)h->n_slots * 8, true);
        if (rc == SQ_OK) {
            uint32_t *seg_lo = seg, *seg_hi = seg + h->n_slots;
            SQ_LAUNCH(ctx, k_pt_segments, grid, PT_TPB, 0, slot, n, seg_lo, seg_hi);
            const uint32_t width = b->max_len;
            if (width) {
                const uint64_t work = (uint64_t)h->n_slots * width;
                SQ_LAUNCH(ctx, k_pt_accumulate, sq_grid_for(ctx, work, PT_TPB, 32), PT_TPB, 0, b->view(), idx,
                          seg_lo, seg_hi, h->n_slots, width, p->errors, p->len_cap, ctx->d_err_table, base,
                          p->st);
            }
        }
    }
    if (rc == SQ_OK && h->fail_idx != ~0ULL) {
        p->skipped = true;
        p->skipped_record = h->fail_idx;
        const uint64_t r = h->fail_idx - base;
        std::vector<sq_meta> metas(b->n);
        rc = sq_batch_get_metas(b, metas.data());
        if (rc == SQ_OK) {
            p->skipped_name.resize(metas[r].name_len);
            if (metas[r].name_len)
                rc = sq_memcpy_d2h(ctx, p->skipped_name.data(), b->text + metas[r].name_off, metas[r].name_len);
        }
    }
    p->n_added += n;
    sq_dfree(ctx, tile);
    sq_dfree(ctx, slot);
    sq_dfree(ctx, idx);
    sq_dfree(ctx, tmpk);
    sq_dfree(ctx, tmpv);
    sq_dfree(ctx, seg);
    return rc;
}

extern "C" int sq_pertile_sync(sq_pertile *p, sq_pertile_info *info) {
    sq_ctx *ctx = p->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    PtState *h = (PtState *)((char *)ctx->h_scratch + 3400);
    CUDA_TRY(cudaMemcpyAsync(h, p->st, sizeof(PtState), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    memset(info, 0, sizeof(*info));
    info->number_of_reads = h->n_kept;
    info->max_length = h->max_len;
    info->n_tiles = h->n_slots;
    info->skipped = p->skipped;
    info->skipped_record = p->skipped_record;
    if (h->err_key != ~0ULL) {
        info->bad_phred = 1;
        info->bad_phred_char = (uint8_t)(h->err_key & 0xFF);
    }
    return SQ_OK;
}

extern "C" int sq_pertile_skipped_name(sq_pertile *p, uint8_t *out, uint64_t cap, uint64_t *len) {
    uint64_t n = p->skipped_name.size() < cap ? p->skipped_name.size() : cap;
    if (n) memcpy(out, p->skipped_name.data(), n);
    *len = n;
    return SQ_OK;
}

extern "C" int sq_pertile_read(sq_pertile *p, uint64_t *tile_ids, double *errors, uint64_t *counts) {
    sq_ctx *ctx = p->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    const uint64_t nt = p->n_slots, ml = p->max_len;
    if (nt == 0) return SQ_OK;
    std::vector<uint64_t> ids(nt), len(nt * (ml ? ml : 1));
    std::vector<double> err(nt * (ml ? ml : 1));
    CUDA_TRY(cudaMemcpyAsync(ids.data(), p->slot_tile, nt * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (ml) {
        CUDA_TRY(cudaMemcpy2DAsync(err.data(), ml * 8, p->errors, p->len_cap * 8, ml * 8, nt, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaMemcpy2DAsync(len.data(), ml * 8, p->lengths, p->len_cap * 8, ml * 8, nt, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    // tiles ascending; counts[j] = reads longer than j (suffix sum of length_counts, :3336-3347)
    std::vector<uint64_t> order(nt);
    for (uint64_t i = 0; i < nt; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) { return ids[a] < ids[b]; });
    for (uint64_t o = 0; o < nt; o++) {
        const uint64_t s = order[o];
        tile_ids[o] = ids[s];
        uint64_t run = 0;
        for (uint64_t j = ml; j-- > 0;) {
            run += len[s * ml + j];
            counts[o * ml + j] = run;
            errors[o * ml + j] = err[s * ml + j];
        }
    }
    return SQ_OK;
}
