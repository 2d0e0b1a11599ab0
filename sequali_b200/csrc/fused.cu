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

// ===========================================================================
// Per-position pass ("columns"): a thread owns four read positions and walks
// the records of a tile that a TMA bulk copy staged in shared memory.
//
//   QCMetrics       base / phred-bin histograms (:2004-2031, :2068-2124) with the
//                   counters of qc.cu's vertical kernel (byte-sliced registers,
//                   byte counters in shared memory, no atomics in the loop)
//   PerTileQuality  for tiles whose records belong to one flow-cell tile: the
//                   exact in-binade integer sums r_k(e) of the error rates for
//                   BOTH binades hinted by pertile.cu's k_pt_guess (k, k + 1),
//                   from rows of the precomputed increment table copied into
//                   shared memory per tile; the chain kernel picks the one that
//                   matches the exact state, or replays the tile read by read
// ===========================================================================
constexpr int FC_BINS = 17;       // 5 base classes + 12 phred bins
constexpr int FC_LUT_WINDOW = 8;  // binade pairs tabulated per tile: rows kmin .. kmin + 8

struct ColumnArgs {
    BatchView bv;
    uint32_t recs_per_tile, n_tiles;
    uint64_t text_end;
    uint32_t CG, RG, W;  // column groups (4 positions each), row groups, W = 4*CG >= longest read
    uint32_t buf_bytes;  // shared-memory tile buffer
    // QCMetrics tables
    int do_qc;
    uint64_t *base, *phred, *ea_base, *ea_phred;
    uint32_t ea_len;
    uint8_t *cta_mixed;  // [grid] CTAs that met reads of different lengths
    // PerTileQuality
    int do_pt;
    const uint64_t *lut;      // [PT_LUT_NK][94] r_k(10^-(q/10)), k = PT_LUT_KMIN + row
    const uint16_t *kguess;   // [W][n_tiles]  (position-major: a chain reads consecutive tiles)
    uint64_t *incr, *incr_hi; // [W][n_tiles] sums for binade kguess / kguess + 1
    const uint8_t *tile_uniform;  // [n_tiles] 1: all records of the tile belong to one flow-cell tile
    PtState *pt_st;
    uint64_t pt_base;
};

template <int TPB>
__global__ void __launch_bounds__(TPB)
k_fused_columns(const ColumnArgs A) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_lmin, s_lmax, s_kmin;
    const uint32_t tid = threadIdx.x, R = A.recs_per_tile, W = A.W;
    uint8_t *buf = smem_raw;
    uint32_t *s_qo = (uint32_t *)(smem_raw + A.buf_bytes + 16);
    uint32_t *s_so = s_qo + R;
    uint32_t *s_L = s_so + R;
    uint64_t *s_lut = (uint64_t *)(s_L + R + (R & 1));    // [WINDOW + 1][128] by raw byte; zero outside '!'..'~'
    uint64_t *parti = s_lut + (FC_LUT_WINDOW + 1) * 128;  // [W][2]
    uint32_t *hist = (uint32_t *)(parti + 2 * W);         // [W][17]
    uint32_t *priv = hist + W * FC_BINS;                  // [16][TPB] words; row 12 = padding, 13..15 only
                                                          // reachable through an invalid quality byte

    if (A.do_qc)
        for (uint32_t i = tid; i < W * FC_BINS + 16 * TPB; i += TPB) hist[i] = 0;
    if (A.do_pt)
        for (uint32_t i = tid; i < (FC_LUT_WINDOW + 1) * 128; i += TPB) s_lut[i] = 0;
    if (tid == 0) {
        s_lmin = 0xFFFFFFFFu;
        s_lmax = 0;
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const BatchView &bv = A.bv;
    const uint32_t CG = A.CG, RG = A.RG;
    const uint32_t rg = tid / CG, cg = tid - rg * CG;
    const bool worker = rg < RG;
    const uint32_t col0 = cg * 4;
    uint8_t *priv8 = (uint8_t *)priv + tid * 4;
    uint32_t acc_v = 0, acc_h = 0, acc_g = 0, acc_hg = 0, acc_n = 0, rows = 0;
    uint32_t lmin = 0xFFFFFFFFu, lmax = 0;

    auto spill = [&]() {
        // registers -> CTA histogram; per column: A = v-h-g+hg, C = h-hg, G = hg, T = g-hg
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t v = (acc_v >> (8 * j)) & 0xFF, h = (acc_h >> (8 * j)) & 0xFF;
            const uint32_t g = (acc_g >> (8 * j)) & 0xFF, hg = (acc_hg >> (8 * j)) & 0xFF;
            const uint32_t nn = (acc_n >> (8 * j)) & 0xFF;
            uint32_t *hrow = hist + (col0 + j) * FC_BINS;
            if (v | nn) {
                const uint32_t a = v - h - g + hg, c = h - hg, t = g - hg;
                if (a) atomicAdd(hrow + 0, a);
                if (c) atomicAdd(hrow + 1, c);
                if (hg) atomicAdd(hrow + 2, hg);
                if (t) atomicAdd(hrow + 3, t);
                if (nn) atomicAdd(hrow + 4, nn);
            }
        }
        acc_v = acc_h = acc_g = acc_hg = acc_n = 0;
#pragma unroll
        for (int b = 0; b < 12; b++) {
            const uint32_t wv = priv[b * TPB + tid];
            if (wv) {
                priv[b * TPB + tid] = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t c = (wv >> (8 * j)) & 0xFF;
                    if (c) atomicAdd(hist + (col0 + j) * FC_BINS + 5 + b, c);
                }
            }
        }
        rows = 0;
    };

    uint32_t parity = 0;
    for (uint32_t t = blockIdx.x; t < A.n_tiles; t += gridDim.x) {
        const uint32_t r0 = t * R, r1 = min(r0 + R, bv.n), nrec = r1 - r0;
        const bool pt_here = A.do_pt && A.tile_uniform[t];  // other tiles are replayed read by read in the chain
        if (!A.do_qc && !pt_here) continue;
        {
            const uint64_t start = (uint64_t)bv.name_off[r0] - 1;
            const uint64_t end = r1 < bv.n ? (uint64_t)bv.name_off[r1] - 1 : A.text_end;
            const uint64_t gstart = start & ~15ULL;
            const uint32_t bytes = (uint32_t)(((end + 15) & ~15ULL) - gstart);
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&bar, bytes);
                bulk_g2s(buf, bv.text + gstart, bytes, &bar);
                s_kmin = 0xFFFFFFFFu;
            }
            for (uint32_t i = tid; i < nrec; i += TPB) {
                const uint32_t L = bv.seq_len[r0 + i];
                s_so[i] = bv.seq_off[r0 + i] - (uint32_t)gstart;
                s_qo[i] = bv.qual_off[r0 + i] - (uint32_t)gstart;
                s_L[i] = L;
                lmin = min(lmin, L);
                lmax = max(lmax, L);
            }
        }
        // ---- PerTileQuality: this tile's window of binades -------------------------------------
        uint32_t kg[4] = {0, 0, 0, 0};
        if (pt_here) {
            uint32_t kmin = 0xFFFFFFFFu;
            if (worker) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    kg[j] = col0 + j < W ? A.kguess[(uint64_t)(col0 + j) * A.n_tiles + t] : 0;
                    if (kg[j]) kmin = min(kmin, kg[j]);
                }
            }
            kmin = ~warp_max_u32(~kmin);
            __syncthreads();  // s_kmin is reset, nobody reads the previous tile's table or sums any more
            if (lane_id() == 0 && kmin != 0xFFFFFFFFu) atomicMin(&s_kmin, kmin);
            __syncthreads();
            kmin = s_kmin;
            for (uint32_t i = tid; i < (FC_LUT_WINDOW + 1) * 94; i += TPB) {
                const uint32_t row = i / 94, q = i - row * 94;
                const uint32_t k = kmin + row - (uint32_t)PT_LUT_KMIN;
                s_lut[row * 128 + 33 + q] = k < (uint32_t)PT_LUT_NK ? A.lut[k * 94 + q] : PT_HARD;
            }
            for (uint32_t i = tid; i < 2 * W; i += TPB) parti[i] = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) kg[j] = (kg[j] && kg[j] - kmin < FC_LUT_WINDOW) ? kg[j] - kmin + 1 : 0;  // 0: no table
        }
        __syncthreads();
        mbar_wait(&bar, parity);
        parity ^= 1;
        uint64_t ia[4] = {0, 0, 0, 0}, ib[4] = {0, 0, 0, 0};
        if (worker) {
            // kg[j] == 0 (no table for that column): any row will do, the sums are marked PT_HARD below
            const uint64_t *lrow0 = s_lut + (kg[0] ? kg[0] - 1 : 0) * 128;
            const uint64_t *lrow1 = s_lut + (kg[1] ? kg[1] - 1 : 0) * 128;
            const uint64_t *lrow2 = s_lut + (kg[2] ? kg[2] - 1 : 0) * 128;
            const uint64_t *lrow3 = s_lut + (kg[3] ? kg[3] - 1 : 0) * 128;
            uint32_t badw = 0;
            for (uint32_t i = rg; i < nrec; i += RG) {
                const uint32_t L = s_L[i];
                if (L <= col0) continue;
                const uint32_t nvalid = min(4u, L - col0);
                // bytes past the end of the read become 0x7F: zero in every table, own trash bin
                const uint32_t keep = 0xFFFFFFFFu >> (8 * (4 - nvalid));
                const uint32_t raw = fh_word(buf, s_qo[i], cg);
                const uint32_t q = (raw & keep) | (0x7F7F7F7Fu & ~keep);
                if (A.do_qc) {
                    const uint32_t w = fh_word(buf, s_so[i], cg);
                    const uint32_t pm = keep & 0x01010101u;
                    const uint32_t vb = fh_acgt_bytes(w) & pm;
                    const uint32_t hb = (w >> 1) & vb, gb = (w >> 2) & vb;
                    acc_v += vb;
                    acc_h += hb;
                    acc_g += gb;
                    acc_hg += hb & gb;
                    acc_n += pm & ~vb;
                    // phred bins min(q,47)>>2 as byte counters at bin*TPB*4 + tid*4 + j, four bins at
                    // once: 4*(min(q,47)>>2) per byte (q >= 48 saturates to bin 11, the padding byte is
                    // counted in the spare 13th row that nobody reads)
                    const uint32_t t4 = q - 0x21212121u;
                    const uint32_t sat = (((t4 + 0x50505050u) >> 7) & 0x01010101u) * 0xFFu;
                    uint32_t bin4 = ((t4 & 0x3C3C3C3Cu) & ~sat) | (0x2C2C2C2Cu & sat);
                    bin4 = (bin4 & keep) | (0x30303030u & ~keep);
                    priv8[(bin4 & 0xFF) * TPB + 0] += 1;
                    priv8[((bin4 >> 8) & 0xFF) * TPB + 1] += 1;
                    priv8[((bin4 >> 16) & 0xFF) * TPB + 2] += 1;
                    priv8[(bin4 >> 24) * TPB + 3] += 1;
                    if (++rows == 255) spill();
                }
                if (pt_here) {
                    badw |= ((raw - 0x21212121u) | (raw + 0x01010101u)) & keep;
                    const uint64_t *e0 = lrow0 + (q & 0xFF), *e1 = lrow1 + ((q >> 8) & 0xFF);
                    const uint64_t *e2 = lrow2 + ((q >> 16) & 0xFF), *e3 = lrow3 + (q >> 24);
                    ia[0] += e0[0];
                    ib[0] += e0[128];
                    ia[1] += e1[0];
                    ib[1] += e1[128];
                    ia[2] += e2[0];
                    ib[2] += e2[128];
                    ia[3] += e3[0];
                    ib[3] += e3[128];
                }
            }
            if (pt_here && (badw & 0x80808080u)) {
                // a quality byte outside '!'..'~' (PerTileQuality raises for it, :3213): find it
                for (uint32_t i = rg; i < nrec; i += RG) {
                    const uint32_t L = s_L[i];
                    for (uint32_t j = 0; j < 4 && col0 + j < L; j++) {
                        const uint32_t c = buf[s_qo[i] + col0 + j];
                        if (c - 33u > 93u)
                            atomicMin(&A.pt_st->err_key, (unsigned long long)((A.pt_base + r0 + i) << 8 | c));
                    }
                }
            }
        }
        // ---- per-tile column sums: combine the row groups -------------------------------------
        if (pt_here) {
            if (worker) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (col0 + j < W) {
                        const uint64_t a = kg[j] && ia[j] < PT_HARD ? ia[j] : PT_HARD;
                        const uint64_t b = kg[j] && ib[j] < PT_HARD ? ib[j] : PT_HARD;
                        atomicAdd((unsigned long long *)(parti + 2 * (col0 + j)), (unsigned long long)a);
                        atomicAdd((unsigned long long *)(parti + 2 * (col0 + j) + 1), (unsigned long long)b);
                    }
                }
            }
            __syncthreads();
            for (uint32_t c = tid; c < W; c += TPB) {
                const uint64_t a = parti[2 * c], b = parti[2 * c + 1];
                A.incr[(uint64_t)c * A.n_tiles + t] = a < PT_HARD ? a : PT_HARD;
                A.incr_hi[(uint64_t)c * A.n_tiles + t] = b < PT_HARD ? b : PT_HARD;
            }
        }
        __syncthreads();  // tile buffer, offsets, table and partial sums are free again
    }
    if (!A.do_qc) return;
    if (worker) spill();
    lmax = warp_max_u32(lmax);
    lmin = ~warp_max_u32(~lmin);
    if (lane_id() == 0) {
        atomicMin(&s_lmin, lmin);
        atomicMax(&s_lmax, lmax);
    }
    __syncthreads();
    // CTA histogram -> global tables
    for (uint32_t i = tid; i < W * FC_BINS; i += TPB) {
        const uint32_t c = hist[i];
        if (!c) continue;
        const uint32_t pos = i / FC_BINS, k = i % FC_BINS;
        if (k < 5) atomic_add_u64(A.base + (uint64_t)pos * 5 + k, c);
        else atomic_add_u64(A.phred + (uint64_t)pos * 12 + (k - 5), c);
    }
    if (s_lmax == 0 && s_lmin == 0xFFFFFFFFu) return;  // this CTA had no tile
    if (s_lmin == s_lmax) {
        // every record had length L0: the end-anchored rows are a shifted window of hist
        const uint32_t L0 = s_lmin, ea_n = min(L0, A.ea_len);
        const uint32_t lo = L0 - ea_n, hi = L0;
        const uint32_t span = (hi - lo) * FC_BINS;
        for (uint32_t i = tid; i < span; i += TPB) {
            const uint32_t pos = lo + i / FC_BINS, k = i % FC_BINS;
            const uint32_t c = hist[pos * FC_BINS + k];
            if (!c) continue;
            const uint64_t row = (uint64_t)A.ea_len - (L0 - pos);
            if (k < 5) atomic_add_u64(A.ea_base + row * 5 + k, c);
            else atomic_add_u64(A.ea_phred + row * 12 + (k - 5), c);
        }
    }
    else if (tid == 0) A.cta_mixed[blockIdx.x] = 1;
}

// end-anchored tables for the CTAs of k_fused_columns that met reads of different
// lengths (:2034-2043, :2115-2124): same tile assignment, one warp per read
constexpr int FC_TPB = 128;  // block size of the fallback (the columns kernel picks its own)
__global__ void __launch_bounds__(FC_TPB)
k_fused_ea_fallback(BatchView bv, uint32_t R, uint32_t n_tiles, const uint8_t *cta_mixed, uint64_t *g_ea_base,
                    uint64_t *g_ea_phred, uint32_t ea_len, int smem_hist) {
    extern __shared__ uint32_t ea_hist[];  // [ea_len][17] when smem_hist
    if (!cta_mixed[blockIdx.x]) return;
    if (smem_hist) {
        for (uint32_t i = threadIdx.x; i < ea_len * FC_BINS; i += FC_TPB) ea_hist[i] = 0;
        __syncthreads();
    }
    const uint32_t warp = threadIdx.x >> 5, nwarps = FC_TPB / 32;
    for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const uint32_t r0 = t * R, r1 = min(r0 + R, bv.n);
        for (uint32_t r = r0 + warp; r < r1; r += nwarps) {
            const uint32_t L = bv.seq_len[r], ea_n = min(L, ea_len);
            const uint8_t *s = bv.text + bv.seq_off[r] + (L - ea_n);
            const uint8_t *q = bv.text + bv.qual_off[r] + (L - ea_n);
            const uint32_t row0 = ea_len - ea_n;
            for (uint32_t k = lane_id(); k < ea_n; k += 32) {
                const uint32_t bcls = nuc5(s[k]);
                const uint32_t p = min((uint32_t)(uint8_t)(q[k] - 33), 47u) >> 2;
                if (smem_hist) {
                    atomicAdd(ea_hist + (row0 + k) * FC_BINS + bcls, 1u);
                    atomicAdd(ea_hist + (row0 + k) * FC_BINS + 5 + p, 1u);
                }
                else {
                    atomic_add_u64(g_ea_base + (uint64_t)(row0 + k) * 5 + bcls, 1);
                    atomic_add_u64(g_ea_phred + (uint64_t)(row0 + k) * 12 + p, 1);
                }
            }
        }
    }
    if (smem_hist) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < ea_len * FC_BINS; i += FC_TPB) {
            const uint32_t c = ea_hist[i];
            if (!c) continue;
            const uint32_t row = i / FC_BINS, k = i % FC_BINS;
            if (k < 5) atomic_add_u64(g_ea_base + (uint64_t)row * 5 + k, c);
            else atomic_add_u64(g_ea_phred + (uint64_t)row * 12 + (k - 5), c);
        }
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
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

// shared-memory plan of k_fused_columns.  The block size is the one of 128 / 192 / 256 threads
// that wastes the fewest lanes for this read length (150 bp: 38 column groups, 5 x 38 = 190 of
// 192 threads work); 4 / 3 / 2 CTAs share an SM.
struct ColGeom {
    uint32_t R, n_tiles, CG, RG, W, buf_bytes, grid, tpb;
    size_t smem;
    bool ok;
};
static ColGeom col_geometry(sq_ctx *ctx, const sq_batch *b) {
    ColGeom g;
    memset(&g, 0, sizeof(g));
    if (b->max_len == 0) return g;
    g.CG = (b->max_len + 3) / 4;
    if (g.CG > 256) return g;
    g.W = g.CG * 4;
    uint32_t best_used = 0;
    for (uint32_t tpb = 128; tpb <= 256; tpb += 64) {
        if (tpb < g.CG) continue;
        const uint32_t used = (tpb / g.CG) * g.CG * 1024 / tpb;  // working lanes per 1024
        if (used > best_used + 16) {
            best_used = used;
            g.tpb = tpb;
        }
    }
    g.RG = g.tpb / g.CG;
    const uint32_t ctas = g.tpb == 128 ? 4 : g.tpb == 192 ? 3 : 2;
    const uint32_t budget = (228u * 1024u) / ctas - 1024u - 128u;
    const uint32_t fixed = (FC_LUT_WINDOW + 1) * 128 * 8 + 2 * g.W * 8 + g.W * FC_BINS * 4 + 16 * g.tpb * 4 + 64;
    if (fixed + 8 * (b->max_rec_bytes + 12) + 64 > budget) return g;
    g.R = (budget - fixed - 64) / (b->max_rec_bytes + 12);
    if (g.R > 255) g.R = 255;
    g.buf_bytes = (g.R * b->max_rec_bytes + 32 + 15) & ~15u;
    g.smem = (size_t)g.buf_bytes + 16 + (size_t)(3 * g.R + 1) * 4 + fixed;
    g.n_tiles = (uint32_t)((b->n + g.R - 1) / g.R);
    g.grid = (uint32_t)ctx->num_sms * ctas;
    if (g.grid > g.n_tiles) g.grid = g.n_tiles;
    g.ok = true;
    return g;
}

template <int TPB>
static int launch_columns_t(sq_ctx *ctx, const ColGeom &g, const ColumnArgs &C) {
    CUDA_TRY(cudaFuncSetAttribute(k_fused_columns<TPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
    SQ_LAUNCH(ctx, k_fused_columns<TPB>, g.grid, TPB, g.smem, C);
    return SQ_OK;
}
static int launch_columns(sq_ctx *ctx, const ColGeom &g, const ColumnArgs &C) {
    if (g.tpb == 128) return launch_columns_t<128>(ctx, g, C);
    if (g.tpb == 192) return launch_columns_t<192>(ctx, g, C);
    return launch_columns_t<256>(ctx, g, C);
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
    uint8_t *cta_mixed = nullptr;
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

    // ---- PerTileQuality: slots, segments and binade hints (needs the tile ids) ---------------
    const ColGeom g = col_geometry(ctx, b);
    PtPlan plan;
    const uint64_t pt_base = pt ? pt->n_added : 0;
    if (rc == SQ_OK && pt) rc = pt_prepare(pt, b, tile, g.ok ? g.R : 0, g.n_tiles, g.W, &plan);

    // ---- per-position pass: QCMetrics histograms + PerTileQuality in-binade sums --------------
    if (rc == SQ_OK && g.ok && (qc || plan.runs)) {
        ColumnArgs C;
        memset(&C, 0, sizeof(C));
        C.bv = b->view();
        C.recs_per_tile = g.R;
        C.n_tiles = g.n_tiles;
        C.text_end = b->text_end;
        C.CG = g.CG;
        C.RG = g.RG;
        C.W = g.W;
        C.buf_bytes = g.buf_bytes;
        if (qc) {
            rc = qc_grow(qc, b->max_len);
            if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&cta_mixed, g.grid, true);
            C.do_qc = 1;
            C.base = qc->base;
            C.phred = qc->phred;
            C.ea_base = qc->ea_base;
            C.ea_phred = qc->ea_phred;
            C.ea_len = (uint32_t)qc->ea_len;
            C.cta_mixed = cta_mixed;
        }
        if (plan.runs) {
            C.do_pt = 1;
            C.lut = pt->lut;
            C.kguess = plan.kguess;
            C.incr = plan.incr;
            C.incr_hi = plan.incr_hi;
            C.tile_uniform = plan.uniform;
            C.pt_st = pt->st;
            C.pt_base = pt_base;
        }
        if (rc == SQ_OK) rc = launch_columns(ctx, g, C);
        if (rc == SQ_OK && qc && qc->ea_len) {
            const size_t ea_smem = (size_t)qc->ea_len * FC_BINS * 4;
            const int use_smem = ea_smem <= 96 * 1024;
            if (use_smem)
                CUDA_TRY(cudaFuncSetAttribute(k_fused_ea_fallback, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              96 * 1024));
            SQ_LAUNCH(ctx, k_fused_ea_fallback, g.grid, FC_TPB, use_smem ? ea_smem : 0, b->view(), g.R, g.n_tiles,
                      cta_mixed, qc->ea_base, qc->ea_phred, (uint32_t)qc->ea_len, use_smem);
        }
    }
    else if (rc == SQ_OK && qc) rc = qc_add_vertical(qc, b);
    if (qc) {
        qc->n_reads += n;
        if (b->max_len > qc->max_len) qc->max_len = b->max_len;
        b->err_sum_valid = true;
    }
    // ---- table maintenance, in the reference's module order ------------------------------------
    if (pt) {
        if (rc == SQ_OK) rc = pt_finish(pt, b, &plan);
        else pt_plan_free(ctx, &plan);
    }
    if (rc == SQ_OK && ov) rc = sq_overrep_add(ov, b);
    if (rc == SQ_OK && ns) rc = sq_nanostats_add(ns, b);
    if (ad) {
        ad->n_seqs += n;
        if (b->max_len > ad->max_len) ad->max_len = b->max_len;
    }
    if (rc == SQ_OK && dd) rc = dedup_consume(dd, hashes, n);
    sq_dfree(ctx, tile);
    if (!(dd && dd->deferred && rc == SQ_OK)) sq_dfree(ctx, hashes);  // a deferred estimator keeps them
    sq_dfree(ctx, cta_mixed);
    return rc;
}
