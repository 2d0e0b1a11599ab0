// fused.cu -- one walk over the text for every per-read quantity of a short-read
// record array (Illumina-sized records; long reads keep the per-module kernels).
//
// The reference visits each record once per collector (__main__.py:279-306):
// QCMetrics_add_meta (_qcmodule.c:1966), PerTileQuality_add_meta (:3124),
// AdapterCounter_add_meta (:2786), DedupEstimator_add_sequence_ptr (:4463).
// Their per-read parts are instruction bound, not byte bound, so here they
// share one kernel:
//
//   k_fused_reads   persistent CTAs; each pulls a tile of whole records into
//                   shared memory with one TMA bulk copy (cp.async.bulk +
//                   mbarrier), then one thread per record:
//                     - sequence -> three bit planes (ACGT-valid, bit1, bit2),
//                       32 positions per register
//                     - GC% bucket from plane popcounts              (:2045-2058)
//                     - adapters: AND of shifted letter planes, first set
//                       bit = first occurrence                       (:2786-2823)
//                     - the four-chain ordered error sum, its tail, the
//                       mean-phred bucket; err_sum written back      (:2059-2137)
//                     - fingerprint + MurmurHash3 for DedupEstimator (:4463-4485)
//                     - tile id from the header                      (:3089-3121)
//
// Per-position tables (base / phred histograms, per-tile sums) stay in the
// vertical kernels of qc.cu / pertile.cu, table maintenance (dedup, overrep)
// in their own files; sq_fused_add() strings them together in the reference's
// module order.
#include "modules.cuh"

constexpr int FH_TPB = 128;
constexpr int FH_BUF = 48 * 1024;     // text tile in shared memory
constexpr int FH_MAX_ADAPTERS = 64;   // patterns kept in shared memory
constexpr int FH_MAX_PAT = 32;        // longest adapter the plane matcher shifts by

struct FusedArgs {
    BatchView bv;
    uint32_t recs_per_tile, n_tiles;
    uint64_t text_end;
    // QCMetrics
    int do_qc;
    const double *err_tab, *edges;
    uint64_t *gc, *mean_phred;
    unsigned long long *qc_err_key;
    uint64_t qc_base;
    // AdapterCounter
    int do_ad;
    const uint8_t *pat;
    const uint32_t *plen;
    uint32_t n_adapters;
    uint64_t *ad_counts, ad_cap_len;
    // DedupEstimator
    int do_dd;
    uint64_t front_len, back_len, front_off, back_off;
    uint64_t *hashes;
    // PerTileQuality
    int do_pt;
    long long *tile;
    uint64_t pt_base;
    PtState *pt_st;
};

// ---- mbarrier / bulk-copy wrappers (sm_90+ PTX) --------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// 0x01 per byte that is ACGTacgt (same construction as qc.cu)
__device__ __forceinline__ uint32_t fh_acgt_bytes(uint32_t w) {
    uint32_t sel = w & 0x07070707u;
    uint32_t t = sel | (sel >> 4);
    uint32_t nib = __byte_perm(t, 0, 0x4420);
    uint32_t expect = __byte_perm(0x40FF40FFu, 0x40FFFF50u, nib);
    return zero_bytes80((w & 0xD8D8D8D8u) ^ expect) >> 7;
}
// bit i of the result = bit 0 of byte i of x (x has only bit 0 of each byte set)
__device__ __forceinline__ uint32_t fh_pack4(uint32_t x) { return (x * 0x00204081u) >> 21 & 0xFu; }

// word k (4 bytes) of a byte string that starts at shared-memory offset `off`
__device__ __forceinline__ uint32_t fh_word(const uint8_t *buf, uint32_t off, uint32_t k) {
    const uint32_t *w = (const uint32_t *)(buf + (off & ~3u)) + k;
    return __funnelshift_r(w[0], w[1], (off & 3u) * 8);
}

template <int NW>
__global__ void __launch_bounds__(FH_TPB)
k_fused_reads(const FusedArgs A) {
    extern __shared__ __align__(128) uint8_t buf[];  // FH_BUF + 16
    __shared__ __align__(8) uint64_t bar;
    __shared__ double s_err[128];  // indexed by the raw quality byte; 0.0 outside '!'..'~'
    __shared__ double s_edge[94];
    __shared__ uint32_t s_gc[101], s_mp[94];
    __shared__ uint8_t s_pat[FH_MAX_ADAPTERS * FH_MAX_PAT];
    __shared__ uint32_t s_plen[FH_MAX_ADAPTERS];
    const uint32_t tid = threadIdx.x;
    for (uint32_t i = tid; i < 128; i += FH_TPB) s_err[i] = (i >= 33 && i < 127) ? A.err_tab[i - 33] : 0.0;
    for (uint32_t i = tid; i < 94; i += FH_TPB) {
        s_edge[i] = A.edges[i];
        s_mp[i] = 0;
    }
    for (uint32_t i = tid; i < 101; i += FH_TPB) s_gc[i] = 0;
    if (A.do_ad) {
        for (uint32_t i = tid; i < A.n_adapters * FH_MAX_PAT; i += FH_TPB)
            s_pat[i] = A.pat[(i / FH_MAX_PAT) * AD_MAXLEN + (i % FH_MAX_PAT)];
        for (uint32_t i = tid; i < A.n_adapters; i += FH_TPB) s_plen[i] = A.plen[i];
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const BatchView &bv = A.bv;
    uint32_t parity = 0;
    for (uint32_t t = blockIdx.x; t < A.n_tiles; t += gridDim.x) {
        const uint32_t r0 = t * A.recs_per_tile, r1 = min(r0 + A.recs_per_tile, bv.n);
        const uint64_t start = (uint64_t)bv.name_off[r0] - 1;
        const uint64_t end = r1 < bv.n ? (uint64_t)bv.name_off[r1] - 1 : A.text_end;
        const uint64_t gstart = start & ~15ULL;
        const uint32_t bytes = (uint32_t)(((end + 15) & ~15ULL) - gstart);
        if (tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&bar, bytes);
            bulk_g2s(buf, bv.text + gstart, bytes, &bar);
        }
        const uint32_t r = r0 + tid;
        const bool active = r < r1;
        uint32_t so = 0, qo = 0, no = 0, L = 0, name_len = 0;
        if (active) {
            const uint32_t noff = bv.name_off[r], soff = bv.seq_off[r];
            L = bv.seq_len[r];
            no = noff - (uint32_t)gstart;
            so = soff - (uint32_t)gstart;
            qo = bv.qual_off[r] - (uint32_t)gstart;
            name_len = soff - 1 - noff;
        }
        mbar_wait(&bar, parity);
        parity ^= 1;
        if (active) {
            // ---- sequence planes ------------------------------------------------------------------
            uint32_t V[NW], H[NW], G[NW];
            if (A.do_qc | A.do_ad) {
#pragma unroll
                for (int pw = 0; pw < NW; pw++) {
                    uint32_t v = 0, h = 0, g = 0;
                    if ((uint32_t)pw * 32 < L) {
#pragma unroll
                        for (int k = 0; k < 8; k++) {
                            const uint32_t pos = pw * 32 + k * 4;
                            if (pos < L) {
                                const uint32_t w = fh_word(buf, so, pos >> 2);
                                const uint32_t nvalid = min(4u, L - pos);
                                const uint32_t pm = 0x01010101u >> (8 * (4 - nvalid));
                                const uint32_t vb = fh_acgt_bytes(w) & pm;
                                const uint32_t hb = (w >> 1) & vb, gb = (w >> 2) & vb;
                                v |= fh_pack4(vb) << (k * 4);
                                h |= fh_pack4(hb) << (k * 4);
                                g |= fh_pack4(gb) << (k * 4);
                            }
                        }
                    }
                    V[pw] = v;
                    H[pw] = h;
                    G[pw] = g;
                }
            }
            // ---- GC bucket (:2045-2058) ---------------------------------------------------------------
            if (A.do_qc) {
                uint32_t gc = 0, valid = 0;
#pragma unroll
                for (int pw = 0; pw < NW; pw++) {
                    gc += __popc(H[pw]);
                    valid += __popc(V[pw]);
                }
                if (valid) {
                    const double pct = (double)gc * 100.0 / (double)valid;
                    atomicAdd(&s_gc[(uint32_t)round(pct)], 1u);
                }
            }
            // ---- adapters: first occurrence per adapter (:2786-2823) ---------------------------------
            if (A.do_ad) {
                uint32_t P[NW];
#pragma unroll
                for (int pw = 0; pw < NW; pw++)
                    P[pw] = L >= (uint32_t)(pw + 1) * 32 ? 0xFFFFFFFFu : (L > (uint32_t)pw * 32 ? (1u << (L - pw * 32)) - 1 : 0u);
                for (uint32_t a = 0; a < A.n_adapters; a++) {
                    const uint32_t m = s_plen[a];
                    if (m == 0 || m > L) continue;
                    uint32_t M[NW];
#pragma unroll
                    for (int pw = 0; pw < NW; pw++) M[pw] = 0xFFFFFFFFu;
                    bool alive = true;
                    for (uint32_t j = 0; j < m && alive; j++) {
                        const uint32_t c = s_pat[a * FH_MAX_PAT + j];
                        uint32_t X[NW + 1];
                        X[NW] = 0;
                        // letter class -> plane of read positions holding that class
                        switch (c) {
                        case 0:
#pragma unroll
                            for (int pw = 0; pw < NW; pw++) X[pw] = V[pw] & ~H[pw] & ~G[pw];
                            break;
                        case 1:
#pragma unroll
                            for (int pw = 0; pw < NW; pw++) X[pw] = V[pw] & H[pw] & ~G[pw];
                            break;
                        case 2:
#pragma unroll
                            for (int pw = 0; pw < NW; pw++) X[pw] = V[pw] & H[pw] & G[pw];
                            break;
                        case 3:
#pragma unroll
                            for (int pw = 0; pw < NW; pw++) X[pw] = V[pw] & ~H[pw] & G[pw];
                            break;
                        default:
#pragma unroll
                            for (int pw = 0; pw < NW; pw++) X[pw] = ~V[pw] & P[pw];
                            break;
                        }
                        uint32_t any = 0;
#pragma unroll
                        for (int pw = 0; pw < NW; pw++) {
                            M[pw] &= __funnelshift_r(X[pw], X[pw + 1], j);
                            any |= M[pw];
                        }
                        alive = any != 0;
                    }
                    if (!alive) continue;
                    uint32_t p = 0xFFFFFFFFu;
#pragma unroll
                    for (int pw = NW - 1; pw >= 0; pw--)
                        if (M[pw]) p = pw * 32 + (__ffs(M[pw]) - 1);
                    uint64_t *fwd = A.ad_counts + (size_t)a * 2 * A.ad_cap_len;
                    atomic_add_u64(fwd + p, 1);
                    atomic_add_u64(fwd + A.ad_cap_len + (L - 1 - p), 1);
                }
            }
            // ---- ordered error sum, mean-phred bucket (:2059-2137) -----------------------------------
            if (A.do_qc) {
                const uint32_t nit = L >= 5 ? (L - 1) / 4 : 0;
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                uint32_t badw = 0;
#pragma unroll 4
                for (uint32_t i = 0; i < nit; i++) {
                    const uint32_t w = fh_word(buf, qo, i);
                    badw |= (w - 0x21212121u) | (w + 0x01010101u);
                    a0 += s_err[w & 0xFF];
                    a1 += s_err[(w >> 8) & 0xFF];
                    a2 += s_err[(w >> 16) & 0xFF];
                    a3 += s_err[w >> 24];
                }
                double sum = ((a0 + a1) + a2) + a3;  // :2098-2099
                if (L) {
                    uint32_t w = fh_word(buf, qo, nit);
                    const uint32_t ntail = L - 4 * nit;  // 1..4, in order (:2100-2112)
                    w &= 0xFFFFFFFFu >> (8 * (4 - ntail));
                    w |= ntail < 4 ? 0x21212121u << (8 * ntail) : 0u;  // padding reads as '!' for the check only
                    badw |= (w - 0x21212121u) | (w + 0x01010101u);
                    sum += s_err[w & 0xFF];
                    if (ntail > 1) sum += s_err[(w >> 8) & 0xFF];
                    if (ntail > 2) sum += s_err[(w >> 16) & 0xFF];
                    if (ntail > 3) sum += s_err[w >> 24];
                }
                if (badw & 0x80808080u) {
                    // a byte outside '!'..'~': report the first one, as the reference's scan would
                    for (uint32_t i = 0; i < L; i++) {
                        const uint32_t c = buf[qo + i];
                        if ((uint32_t)(c - 33) > 93u) {
                            atomicMin(A.qc_err_key, (unsigned long long)((A.qc_base + r) << 8 | c));
                            break;
                        }
                    }
                }
                else {
                    bv.err_sum[r] = sum;
                    if (L) {  // floor(-10*log10(sum/L)) through host-derived bucket edges
                        const double avg = sum / (double)L;
                        uint32_t lo = 0, hi = 93;
                        while (lo < hi) {
                            const uint32_t mid = (lo + hi + 1) >> 1;
                            if (avg <= s_edge[mid]) lo = mid;
                            else hi = mid - 1;
                        }
                        atomicAdd(&s_mp[lo], 1u);
                    }
                }
            }
            // ---- fingerprint hash (:4463-4485) -------------------------------------------------------
            if (A.do_dd) {
                const uint8_t *s = buf + so;
                const uint64_t fl = A.front_len + A.back_len;
                uint64_t h;
                if (L <= fl) h = murmur3_h2([&](uint64_t i) { return s[i]; }, L, 0);
                else {
                    const uint64_t rem = L - fl;
                    const uint64_t fo = min(rem / 2, A.front_off), bo = min(rem / 2, A.back_off);
                    const uint8_t *f = s + fo, *b = s + L - (bo + A.back_len);
                    const uint64_t front_len = A.front_len;
                    h = murmur3_h2([&](uint64_t i) { return i < front_len ? f[i] : b[i - front_len]; }, fl,
                                   (uint64_t)L >> 6);
                }
                A.hashes[r] = h;
            }
            // ---- tile id: decimal between the 4th and 5th ':' of the header (:3089-3121) -------------
            if (A.do_pt) {
                const uint8_t *hname = buf + no;
                uint32_t i = 0, colons = 0;
                for (; i < name_len; i++)
                    if (hname[i] == ':' && ++colons == 4) break;
                const uint32_t first = i + 1;
                uint32_t j = first;
                long long v = 0;
                bool ok = true;
                for (; j < name_len; j++) {
                    const uint32_t d = (uint32_t)hname[j] - '0';
                    if (hname[j] == ':') break;
                    ok &= d <= 9;
                    v = v * 10 + d;
                }
                const uint32_t len = j - first;
                if (j >= name_len || len < 1 || len > 18 || !ok) v = -1;
                A.tile[r] = v;
                if (v < 0) atomicMin(&A.pt_st->fail_idx, (unsigned long long)(A.pt_base + r));
            }
        }
        __syncthreads();  // everyone is done with the tile before the next copy lands
    }
    if (A.do_qc) {
        for (uint32_t i = tid; i < 101; i += FH_TPB)
            if (s_gc[i]) atomic_add_u64(A.gc + i, s_gc[i]);
        for (uint32_t i = tid; i < 94; i += FH_TPB)
            if (s_mp[i]) atomic_add_u64(A.mean_phred + i, s_mp[i]);
    }
}

template <int NW>
static int launch_fused(sq_ctx *ctx, const FusedArgs &A) {
    CUDA_TRY(cudaFuncSetAttribute(k_fused_reads<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, FH_BUF + 16));
    int grid = ctx->num_sms * 4;
    if ((uint32_t)grid > A.n_tiles) grid = (int)A.n_tiles;
    SQ_LAUNCH(ctx, k_fused_reads<NW>, grid, FH_TPB, FH_BUF + 16, A);
    return SQ_OK;
}

// Can this record array take the fused pass?  FASTQ text straight from the
// parser (records contiguous, ASCII checked), short records, adapters the
// plane matcher can shift.
static bool fused_eligible(const sq_batch *b, const sq_adapters *ad) {
    if (b->name_len != nullptr || b->max_rec_bytes == 0) return false;  // packed (BAM / constructed) arrays
    if (b->max_len > 320) return false;
    if ((uint64_t)b->max_rec_bytes * 8 + 32 > FH_BUF) return false;  // at least 8 records per tile
    if (ad && (ad->n_adapters > FH_MAX_ADAPTERS || ad->max_pat_len > FH_MAX_PAT)) return false;
    return true;
}

extern "C" int sq_fused_add(sq_ctx *ctx, sq_batch *b, sq_qc *qc, sq_pertile *pt, sq_overrep *ov,
                            sq_nanostats *ns, sq_adapters *ad, sq_dedup *dd) {
    if (b->ctx != ctx || (qc && qc->ctx != ctx) || (pt && pt->ctx != ctx) || (ad && ad->ctx != ctx) ||
        (dd && dd->ctx != ctx)) {
        sq_set_error("record array belongs to another context");
        return SQ_E_ARG;
    }
    if (b->n == 0) return SQ_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (pt && pt->skipped) pt = nullptr;
    if (!fused_eligible(b, ad) || !(qc || pt || ad || dd)) {
        // the reference's module order (__main__.py:280-306)
        if (qc) SQ_TRY(sq_qc_add(qc, b));
        if (pt) SQ_TRY(sq_pertile_add(pt, b));
        if (ov) SQ_TRY(sq_overrep_add(ov, b));
        if (ns) SQ_TRY(sq_nanostats_add(ns, b));
        if (ad) SQ_TRY(sq_adapters_add(ad, b));
        if (dd) SQ_TRY(sq_dedup_add(dd, b));
        return SQ_OK;
    }
    const uint32_t n = (uint32_t)b->n;
    FusedArgs A;
    memset(&A, 0, sizeof(A));
    A.bv = b->view();
    uint32_t rpt = (FH_BUF - 32) / b->max_rec_bytes;
    if (rpt > FH_TPB) rpt = FH_TPB;
    A.recs_per_tile = rpt;
    A.n_tiles = (n + rpt - 1) / rpt;
    A.text_end = b->text_end;
    A.err_tab = ctx->d_err_table;
    A.edges = ctx->d_phred_thresholds;
    long long *tile = nullptr;
    uint64_t *hashes = nullptr;
    if (qc) {
        A.do_qc = 1;
        A.gc = qc->gc;
        A.mean_phred = qc->mean_phred;
        A.qc_err_key = qc->err_key;
        A.qc_base = qc->n_reads;
    }
    if (ad) {
        SQ_TRY(adapters_grow(ad, b->max_len));
        A.do_ad = b->max_len > 0;
        A.pat = ad->pat;
        A.plen = ad->plen;
        A.n_adapters = ad->n_adapters;
        A.ad_counts = ad->counts;
        A.ad_cap_len = ad->cap_len;
    }
    if (dd) {
        SQ_TRY(sq_dalloc(ctx, (void **)&hashes, (size_t)n * 8, false));
        A.do_dd = 1;
        A.front_len = dd->front_len;
        A.back_len = dd->back_len;
        A.front_off = dd->front_off;
        A.back_off = dd->back_off;
        A.hashes = hashes;
    }
    if (pt) {
        SQ_TRY(sq_dalloc(ctx, (void **)&tile, (size_t)n * 8, false));
        A.do_pt = 1;
        A.tile = tile;
        A.pt_base = pt->n_added;
        A.pt_st = pt->st;
    }
    int rc;
    if (b->max_len <= 96) rc = launch_fused<3>(ctx, A);
    else if (b->max_len <= 160) rc = launch_fused<5>(ctx, A);
    else if (b->max_len <= 256) rc = launch_fused<8>(ctx, A);
    else rc = launch_fused<10>(ctx, A);
    // per-position tables and table maintenance, in the reference's module order
    if (rc == SQ_OK && qc) {
        rc = qc_add_vertical(qc, b);
        qc->n_reads += n;
        if (b->max_len > qc->max_len) qc->max_len = b->max_len;
        b->err_sum_valid = true;
    }
    if (rc == SQ_OK && pt) rc = pt_add_with_tiles(pt, b, tile);
    if (rc == SQ_OK && ov) rc = sq_overrep_add(ov, b);
    if (rc == SQ_OK && ns) rc = sq_nanostats_add(ns, b);
    if (ad) {
        ad->n_seqs += n;
        if (b->max_len > ad->max_len) ad->max_len = b->max_len;
    }
    if (rc == SQ_OK && dd) rc = dedup_consume(dd, hashes, n);
    sq_dfree(ctx, tile);
    sq_dfree(ctx, hashes);
    return rc;
}
