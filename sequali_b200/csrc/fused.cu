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

constexpr int FH_RECS = 128;          // records per text tile, one per thread of a role
constexpr int FH_TPB = 2 * FH_RECS;   // two roles (warp specialised): sequence side, quality / header side
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
    uint8_t *all_acgt;  // [n] 1: every base of the read is ACGTacgt (k_fused_columns skips the letter test)
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
// the same test with the flag left in bit 7 of each byte
__device__ __forceinline__ uint32_t fh_acgt_bytes80(uint32_t w) {
    uint32_t sel = w & 0x07070707u;
    uint32_t t = sel | (sel >> 4);
    uint32_t nib = __byte_perm(t, 0, 0x4420);
    uint32_t expect = __byte_perm(0x40FF40FFu, 0x40FFFF50u, nib);
    return zero_bytes80((w & 0xD8D8D8D8u) ^ expect);
}
// bit i of the result = bit 0 of byte i of x (x has only bit 0 of each byte set)
__device__ __forceinline__ uint32_t fh_pack4(uint32_t x) { return (x * 0x00204081u) >> 21 & 0xFu; }

// word k (4 bytes) of a byte string that starts at shared-memory offset `off`
__device__ __forceinline__ uint32_t fh_word(const uint8_t *buf, uint32_t off, uint32_t k) {
    const uint32_t *w = (const uint32_t *)(buf + (off & ~3u)) + k;
    return __funnelshift_r(w[0], w[1], (off & 3u) * 8);
}

// the same for an offset packed as (off >> 2) << 5 | (off & 3) * 8: the funnel shift takes the low five
// bits as they are, the word index is one shift away (k_fused_columns packs once per record and tile)
__device__ __forceinline__ uint32_t fh_pack_off(uint32_t off) { return (off >> 2) << 5 | (off & 3u) * 8; }
__device__ __forceinline__ uint32_t fh_word_packed(const uint8_t *buf, uint32_t packed, uint32_t k) {
    const uint32_t *w = (const uint32_t *)buf + (packed >> 5) + k;
    return __funnelshift_r(w[0], w[1], packed);
}

template <int NW>
__global__ void __launch_bounds__(FH_TPB, NW <= 5 ? 4 : 3)
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
    uint32_t qmn = 0xFFu, qmx = 0;  // quality bytes seen in the sampled words
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
        // warps 0-3 take the sequence side of a record (planes, GC, adapters), warps 4-7 its quality
        // string and header (error sum, fingerprint hash, tile id): twice the warps per tile, and
        // neither half carries the other's registers
        const uint32_t role = tid / FH_RECS;
        const uint32_t r = r0 + (tid - role * FH_RECS);
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
        if (active && role == 0) {
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
                                // flags live in bit 7 of their byte: one multiplication lands the four of a
                                // word in bits 28..31 (every partial product on its own bit), one shift and one
                                // masked OR put them at positions 4k .. 4k+3 of the plane
                                const uint32_t w = fh_word(buf, so, pos >> 2);
                                const uint32_t nvalid = min(4u, L - pos);
                                const uint32_t pm = 0x80808080u >> (8 * (4 - nvalid));
                                const uint32_t vb = fh_acgt_bytes80(w) & pm;
                                const uint32_t hb = (w << 6) & vb, gb = (w << 5) & vb;  // bit 1 / bit 2 of the letter
                                const uint32_t nib = 0xFu << (k * 4);
                                v |= ((vb * 0x00204081u) >> (28 - k * 4)) & nib;
                                h |= ((hb * 0x00204081u) >> (28 - k * 4)) & nib;
                                g |= ((gb * 0x00204081u) >> (28 - k * 4)) & nib;
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
                A.all_acgt[r] = valid == L;
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
                    auto letter = [&](uint32_t j) {
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
#pragma unroll
                        for (int pw = 0; pw < NW; pw++) M[pw] &= __funnelshift_r(X[pw], X[pw + 1], j);
                    };
                    auto any_left = [&]() {
                        uint32_t any = 0;
#pragma unroll
                        for (int pw = 0; pw < NW; pw++) any |= M[pw];
                        return any != 0;
                    };
                    // two letters per look at the candidates (a letter too many now and then costs less
                    // than testing after every one)
                    uint32_t j = 0;
                    for (; j + 1 < m && alive; j += 2) {
                        letter(j);
                        letter(j + 1);
                        alive = any_left();
                    }
                    if (alive && j < m) {
                        letter(j);
                        alive = any_left();
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
        }
        else if (active) {
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
                // range of the quality values, from two words per read at positions that sweep the read
                // length over consecutive reads: sizes the per-quality counters of k_fused_columns (a value
                // the sample misses only costs that kernel its slow path)
                if (L) {
                    const uint32_t nw = L >> 2;
                    uint32_t w0, w1;
                    if (nw) {
                        const uint32_t i0 = r % nw, i1 = i0 + (nw >> 1) >= nw ? i0 + (nw >> 1) - nw : i0 + (nw >> 1);
                        w0 = fh_word(buf, qo, i0);
                        w1 = fh_word(buf, qo, i1);
                    }
                    else {
                        w0 = fh_word(buf, qo, 0);
                        const uint32_t b0 = w0 & 0xFF;
                        w0 = L == 1 ? b0 * 0x01010101u : L == 2 ? (w0 & 0xFFFF) * 0x00010001u : (w0 & 0xFFFFFF) | b0 << 24;
                        w1 = w0;
                    }
                    const uint32_t mn = __vminu4(w0, w1), mx = __vmaxu4(w0, w1);
                    qmn = min(qmn, min(min(mn & 0xFF, (mn >> 8) & 0xFF), min((mn >> 16) & 0xFF, mn >> 24)));
                    qmx = max(qmx, max(max(mx & 0xFF, (mx >> 8) & 0xFF), max((mx >> 16) & 0xFF, mx >> 24)));
                }
            }
        }
        __syncthreads();  // everyone is done with the tile before the next copy lands
    }
    if (A.do_pt) {
        qmx = warp_max_u32(qmx);
        qmn = ~warp_max_u32(~qmn);
        if (lane_id() == 0 && qmn <= qmx) {
            atomicMin(&A.pt_st->qmin, qmn);
            atomicMax(&A.pt_st->qmax, qmx);
        }
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
//   QCMetrics       base / phred-bin histograms (:2004-2031, :2068-2124): byte-sliced
//                   registers for the bases, private byte counters in shared memory
//                   for the qualities, no atomics in the loop
//   PerTileQuality  (PT) the private byte counters are indexed by the quality VALUE
//                   (rows qbase .. qbase + qrows of the range k_fused_reads sampled),
//                   one set per segment of `tiles_per_seg` text tiles.  At the end of
//                   a segment the row groups are combined: four consecutive rows are
//                   one phred bin of QCMetrics, and the per-(position, quality)
//                   counts leave through a TMA bulk store as the segment's quality
//                   histogram (PtHistGeom).  k_pt_chain_hist turns a histogram into
//                   the exact in-binade integer sum of that segment's error rates for
//                   whatever binade the chain is in -- no table lookups per base here.
//                   A quality outside the sampled rows takes a slow path (QCMetrics
//                   counted directly, the segment flagged for a read-by-read replay).
// ===========================================================================
constexpr int FC_BINS = 17;  // 5 base classes + 12 phred bins

struct ColumnArgs {
    BatchView bv;
    uint32_t recs_per_tile, n_tiles;  // text tiles: one bulk copy each
    uint32_t tiles_per_seg, n_segs;   // PT: a segment = tiles_per_seg text tiles = one fixed tile of PerTileQuality
    uint64_t text_end;
    uint32_t CG, RG, W;  // column groups (4 positions each), row groups, W = 4*CG >= longest read
    uint32_t buf_bytes;  // shared-memory tile buffer
    // QCMetrics tables
    int do_qc;
    uint64_t *base, *phred, *ea_base, *ea_phred;
    uint32_t ea_len;
    uint8_t *cta_mixed;  // [grid] CTAs that met reads of different lengths
    const uint8_t *all_acgt;  // [n] from k_fused_reads
    // PerTileQuality
    PtHistGeom hg;
    uint8_t *qh;                 // [n_segs][hg.seg_bytes]
    const uint8_t *seg_uniform;  // [n_segs] 1: all records of the segment belong to one flow-cell tile
    uint8_t *seg_oob;            // [n_segs] set to 1 when the histogram misses a read
    PtState *pt_st;
    uint64_t pt_base;
};

__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

template <int TPB, bool PT>
__global__ void __launch_bounds__(TPB)
k_fused_columns(const ColumnArgs A) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_lmin, s_lmax, s_oob;
    const uint32_t tid = threadIdx.x, R = A.recs_per_tile, W = A.W;
    const uint32_t prows = PT ? A.hg.qrows + 1 : 16;  // counter rows: PT qualities + trash; else 12 bins + spare
    uint8_t *buf = smem_raw;
    uint32_t *s_qo = (uint32_t *)(smem_raw + A.buf_bytes + 16);
    uint32_t *s_so = s_qo + R;
    uint32_t *s_L = s_so + R;
    uint32_t *hist = s_L + R;                  // [W][17]
    uint32_t *priv = hist + W * FC_BINS;       // [prows][TPB] words = private byte counters of 4 positions
    uint32_t *stage = priv + prows * TPB;      // PT: [4][CG][QW] words, 16-byte aligned (host pads)
    stage = (uint32_t *)(((uintptr_t)stage + 15) & ~(uintptr_t)15);

    for (uint32_t i = tid; i < W * FC_BINS + prows * TPB; i += TPB) hist[i] = 0;
    if (PT)
        for (uint32_t i = tid; i < 4 * A.CG * A.hg.QW; i += TPB) stage[i] = 0;  // padding words stay 0
    if (tid == 0) {
        s_lmin = 0xFFFFFFFFu;
        s_lmax = 0;
        s_oob = 0;
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
    // PT: quality byte -> counter row, four at a time
    const uint32_t base4 = (A.hg.qbase + 33u) * 0x01010101u;
    const uint32_t hi4 = (0x80u - A.hg.qrows) * 0x01010101u;
    const uint32_t trash4 = A.hg.qrows * 0x01010101u;

    auto spill_bases = [&]() {
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
        rows = 0;
    };
    auto spill = [&]() {
        spill_bases();
        if (!PT) {
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
        }
    };

    uint32_t parity = 0;
    const uint32_t TS = PT ? A.tiles_per_seg : 1, n_segs = PT ? A.n_segs : A.n_tiles;
    for (uint32_t seg = blockIdx.x; seg < n_segs; seg += gridDim.x) {
        const bool pt_here = PT && A.seg_uniform[seg];  // other segments are replayed read by read in the chain
        if (!A.do_qc && !pt_here) continue;
        const uint32_t t_end = min((seg + 1) * TS, A.n_tiles);
        for (uint32_t t = seg * TS; t < t_end; t++) {
            const uint32_t r0 = t * R, r1 = min(r0 + R, bv.n), nrec = r1 - r0;
            {
                const uint64_t start = (uint64_t)bv.name_off[r0] - 1;
                const uint64_t end = r1 < bv.n ? (uint64_t)bv.name_off[r1] - 1 : A.text_end;
                const uint64_t gstart = start & ~15ULL;
                const uint32_t bytes = (uint32_t)(((end + 15) & ~15ULL) - gstart);
                if (tid == 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_expect_tx(&bar, bytes);
                    bulk_g2s(buf, bv.text + gstart, bytes, &bar);
                }
                for (uint32_t i = tid; i < nrec; i += TPB) {
                    const uint32_t L = bv.seq_len[r0 + i];
                    s_so[i] = fh_pack_off(bv.seq_off[r0 + i] - (uint32_t)gstart);
                    s_qo[i] = fh_pack_off(bv.qual_off[r0 + i] - (uint32_t)gstart);
                    // bit 31: the read holds nothing but ACGT (k_fused_reads looked at every base already)
                    s_L[i] = L | (A.do_qc && A.all_acgt[r0 + i] ? 0x80000000u : 0u);
                    lmin = min(lmin, L);
                    lmax = max(lmax, L);
                }
            }
            __syncthreads();
            mbar_wait(&bar, parity);
            parity ^= 1;
            if (worker && rg < nrec) {
                // the byte counters (registers and shared memory) hold 255 rows: settle that once per text
                // tile, not once per row
                const uint32_t my_rows = (nrec - rg + RG - 1) / RG;
                if (rows + my_rows > 255) {
                    if (PT) {
                        if (A.do_qc) spill_bases();
                    }
                    else spill();
                }
                rows += my_rows;
                // software pipeline: the words of the next record are on their way while this one is counted
                // (the last row prefetches itself again: no branch)
                uint32_t nL = s_L[rg], nraw = fh_word_packed(buf, s_qo[rg], cg), nw = 0;
                if (A.do_qc) nw = fh_word_packed(buf, s_so[rg], cg);
                for (uint32_t i = rg; i < nrec; i += RG) {
                    const uint32_t L = nL & 0x7FFFFFFFu, raw = nraw, w = nw;
                    const bool acgt_only = nL >> 31;
                    {
                        const uint32_t ip = min(i + RG, nrec - 1);
                        nL = s_L[ip];
                        nraw = fh_word_packed(buf, s_qo[ip], cg);
                        if (A.do_qc) nw = fh_word_packed(buf, s_so[ip], cg);
                    }
                    if (L <= col0) continue;
                    const uint32_t nvalid = min(4u, L - col0);
                    const uint32_t keep = 0xFFFFFFFFu >> (8 * (4 - nvalid));
                    if (A.do_qc) {
                        const uint32_t pm = keep & 0x01010101u;
                        uint32_t vb = pm;
                        if (!acgt_only) {  // (a warp mostly walks one record: the branch is nearly uniform)
                            vb = fh_acgt_bytes(w) & pm;
                            acc_n += pm & ~vb;
                        }
                        const uint32_t hb = (w >> 1) & vb, gb = (w >> 2) & vb;
                        acc_v += vb;
                        acc_h += hb;
                        acc_g += gb;
                        acc_hg += hb & gb;
                    }
                    uint32_t row4;  // per byte: counter row
                    if (PT) {
                        // row = quality - (qbase + 33); bytes past the end of the read and qualities outside
                        // [qbase, qbase + qrows) go to the trash row, the latter through the slow path
                        const uint32_t d = (raw | 0x80808080u) - base4;  // bit 7: quality >= first row
                        const uint32_t idx = d & 0x7F7F7F7Fu;
                        const uint32_t e = idx + hi4;                     // bit 7: row >= qrows
                        const uint32_t bad = (~d | e) & 0x80808080u & keep;
                        row4 = (idx & keep) | (trash4 & ~keep);
                        if (bad) {
                            s_oob = 1;
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                if (!((bad >> (8 * j)) & 0x80u)) continue;
                                const uint32_t c = (raw >> (8 * j)) & 0xFF;
                                row4 = (row4 & ~(0xFFu << (8 * j))) | (A.hg.qrows << (8 * j));
                                if (c - 33u > 93u)  // PerTileQuality raises for it (:3213), QCMetrics found it already
                                    atomicMin(&A.pt_st->err_key,
                                              (unsigned long long)((A.pt_base + r0 + i) << 8 | c));
                                else if (A.do_qc)
                                    atomicAdd(hist + (col0 + j) * FC_BINS + 5 + (min(c - 33u, 47u) >> 2), 1u);
                            }
                        }
                        {
                            // the four counters are different bytes (byte lane j): load all, then store all,
                            // one shared-memory round trip instead of four in a row
                            uint8_t *p0 = priv8 + (row4 & 0xFF) * (TPB * 4) + 0;
                            uint8_t *p1 = priv8 + ((row4 >> 8) & 0xFF) * (TPB * 4) + 1;
                            uint8_t *p2 = priv8 + ((row4 >> 16) & 0xFF) * (TPB * 4) + 2;
                            uint8_t *p3 = priv8 + (row4 >> 24) * (TPB * 4) + 3;
                            const uint32_t c0 = *p0, c1 = *p1, c2 = *p2, c3 = *p3;
                            *p0 = (uint8_t)(c0 + 1);
                            *p1 = (uint8_t)(c1 + 1);
                            *p2 = (uint8_t)(c2 + 1);
                            *p3 = (uint8_t)(c3 + 1);
                        }
                    }
                    else {
                        // phred bins min(q,47)>>2 as byte counters at bin*TPB*4 + tid*4 + j, four bins at
                        // once: 4*(min(q,47)>>2) per byte (q >= 48 saturates to bin 11, the padding byte is
                        // counted in the spare 13th row that nobody reads)
                        const uint32_t q = (raw & keep) | (0x7F7F7F7Fu & ~keep);
                        const uint32_t t4 = q - 0x21212121u;
                        const uint32_t sat = (((t4 + 0x50505050u) >> 7) & 0x01010101u) * 0xFFu;
                        uint32_t bin4 = ((t4 & 0x3C3C3C3Cu) & ~sat) | (0x2C2C2C2Cu & sat);
                        bin4 = (bin4 & keep) | (0x30303030u & ~keep);
                        {
                            uint8_t *p0 = priv8 + (bin4 & 0xFF) * TPB + 0;
                            uint8_t *p1 = priv8 + ((bin4 >> 8) & 0xFF) * TPB + 1;
                            uint8_t *p2 = priv8 + ((bin4 >> 16) & 0xFF) * TPB + 2;
                            uint8_t *p3 = priv8 + (bin4 >> 24) * TPB + 3;
                            const uint32_t c0 = *p0, c1 = *p1, c2 = *p2, c3 = *p3;
                            *p0 = (uint8_t)(c0 + 1);
                            *p1 = (uint8_t)(c1 + 1);
                            *p2 = (uint8_t)(c2 + 1);
                            *p3 = (uint8_t)(c3 + 1);
                        }
                    }
                }
            }
            if (PT && tid == 0) bulk_wait_read();  // the previous segment's histogram has left `stage`
            __syncthreads();                       // tile buffer and offsets are free again
        }
        if (!PT) continue;
        // ---- end of a segment: combine the row groups' private counters --------------------------
        // thread (rg, cg) owns the quads of rows 4m .. 4m+3 with m = rg, rg + RG, ...: one phred bin each
        if (worker) {
            const uint32_t nq = A.hg.qrows >> 2, QW = A.hg.QW;
            for (uint32_t m = rg; m < nq; m += RG) {
                uint32_t w[4];
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    uint32_t *pr = priv + (4 * m + r) * TPB + cg;
                    uint32_t acc = 0;
                    for (uint32_t g = 0; g < RG; g++) {  // counts of a segment fit a byte: packed adds
                        acc += pr[g * CG];
                        pr[g * CG] = 0;
                    }
                    w[r] = acc;
                }
                const uint32_t qsum = w[0] + w[1] + w[2] + w[3];
                if (qsum == 0) {
                    if (pt_here) {
#pragma unroll
                        for (int j = 0; j < 4; j++) stage[(j * CG + cg) * QW + m] = 0;
                    }
                    continue;
                }
                if (A.do_qc) {
                    const uint32_t bin = min((A.hg.qbase >> 2) + m, 11u);
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint32_t c = (qsum >> (8 * j)) & 0xFF;
                        if (c) atomicAdd(hist + (col0 + j) * FC_BINS + 5 + bin, c);
                    }
                }
                if (pt_here) {
                    // 4 x 4 byte transpose: word j = counts of rows 4m .. 4m+3 at position col0 + j
                    const uint32_t t0 = __byte_perm(w[0], w[1], 0x5140), t1 = __byte_perm(w[2], w[3], 0x5140);
                    const uint32_t t2 = __byte_perm(w[0], w[1], 0x7362), t3 = __byte_perm(w[2], w[3], 0x7362);
                    stage[(0 * CG + cg) * QW + m] = __byte_perm(t0, t1, 0x5410);
                    stage[(1 * CG + cg) * QW + m] = __byte_perm(t0, t1, 0x7632);
                    stage[(2 * CG + cg) * QW + m] = __byte_perm(t2, t3, 0x5410);
                    stage[(3 * CG + cg) * QW + m] = __byte_perm(t2, t3, 0x7632);
                }
            }
            // the trash row only ever counts padding and slow-path bytes
            priv[A.hg.qrows * TPB + tid] = 0;
        }
        __syncthreads();
        if (tid == 0) {
            if (pt_here) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                bulk_s2g(A.qh + (uint64_t)seg * A.hg.seg_bytes, stage, A.hg.seg_bytes);
                if (s_oob) A.seg_oob[seg] = 1;
            }
            s_oob = 0;
        }
    }
    if (PT && tid == 0) bulk_wait_read();
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

// ===========================================================================
// k_onewalk: per-read AND per-position work on one TMA tile (round 2).
//
// k_fused_reads and k_fused_columns above each pull the text through shared
// memory once, and the columns kernel spends most of its instructions on
// re-finding the bytes (two loads + a funnel shift per unaligned word, the
// letter test, length masks).  Here a CTA of 192 threads does both on the same
// tile of <= 96 whole records:
//
//   phase 1  (one thread per record and role, as in k_fused_reads)
//     sequence role  letters -> one-hot class planes A / C / G / T / other (32 positions per
//                    register); GC bucket from popcounts; adapters = AND of shifted class planes,
//                    the shifts shared by up to six adapters (:2786-2823); the planes are left in
//                    shared memory for the base histogram
//     quality role   four-chain ordered error sum + mean-phred bucket (:2059-2137), and every
//                    quality word turned into four COUNTER ROWS (quality value, or phred bin when
//                    PerTileQuality is not fed) stored word-aligned in a staging matrix; fingerprint
//                    hash (:4463-4485); tile id (:3089-3121)
//   base histogram   bit-sliced vertical counters: a thread owns one plane word (32 positions of one
//                    letter class) of ~14 records, adds them with carry-save adders (two LOP3 per
//                    input word) into 12 registers that count up to 4095 reads per position, and
//                    expands them once per launch -- ~4 warp instructions per read instead of ~36
//   phase 2  (a thread per 4 positions and row group, as in k_fused_columns) one aligned load of
//                    the staged rows per record, then the four private byte counters
//   the next tile's bulk copy is issued as soon as phase 1 is done with the text (phase 2 only
//   reads the staging matrix), so the copy overlaps phase 2.
//
// Segments (TS tiles, <= 255 records) end exactly as in k_fused_columns: row groups combined, four
// consecutive rows = one phred bin of QCMetrics, the per-(position, quality) byte counts leave
// through a TMA bulk store as the segment's histogram for k_pt_chain_hist (PtHistGeom).  Every
// segment gets its histogram; pt_prepare decides afterwards which ones are usable (one flow-cell
// tile per segment), so the tile ids can come from the same kernel.
// ===========================================================================
constexpr int MG_TPB = 192;
constexpr int MG_RMAX = 96;         // records per text tile (a multiple of 32: the roles split at warp borders)
constexpr int MG_BUF = 36 * 1024;   // text tile in shared memory
constexpr int MG_GA = 6;            // adapters matched together (they share the shifted class planes)
constexpr int MG_QBINS = 12;

struct WalkArgs {
    BatchView bv;
    uint32_t R, n_tiles;      // text tiles: one bulk copy each
    uint32_t TS, n_segs;      // a segment = TS text tiles (<= 255 records)
    uint64_t text_end;
    uint32_t CG, RG, W, QS;   // column groups (4 positions), row groups, W = 4 * CG, row stride of the staging matrix
    uint32_t PARTS, PR;       // base histogram: record parts per plane word, records per part (<= 16)
    uint32_t buf_bytes, xstage_words, prows;
    int do_hist;              // phase 2 runs (QCMetrics and / or PerTileQuality histograms)
    int pt_rows;              // counter rows are quality values (hg), else the 12 phred bins
    // QCMetrics
    int do_qc;
    const double *err_tab, *edges;
    uint64_t *gc, *mean_phred;
    unsigned long long *qc_err_key;
    uint64_t qc_base;
    uint64_t *base, *phred, *ea_base, *ea_phred;
    uint32_t ea_len;
    uint8_t *cta_mixed;
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
    PtHistGeom hg;
    uint8_t *qh;       // [n_segs][hg.seg_bytes]
    uint8_t *seg_oob;  // [n_segs]
};

// carry-save adder: (h, l) = a + b + c per bit position
__device__ __forceinline__ void csa(uint32_t &h, uint32_t &l, uint32_t a, uint32_t b, uint32_t c) {
    const uint32_t u = a ^ b;
    h = (a & b) | (u & c);
    l = u ^ c;
}

template <int NW>
__global__ void __launch_bounds__(MG_TPB, 2)
k_onewalk(const WalkArgs A) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ double s_err[128];  // indexed by the raw quality byte; 0.0 outside '!'..'~'
    __shared__ double s_edge[94];
    __shared__ uint32_t s_gc[101], s_mp[94];
    __shared__ uint8_t s_pat[FH_MAX_ADAPTERS * FH_MAX_PAT];
    __shared__ uint32_t s_plen[FH_MAX_ADAPTERS];
    __shared__ uint32_t s_lmin, s_lmax, s_oob;
    const uint32_t tid = threadIdx.x, R = A.R, W = A.W, CG = A.CG, RG = A.RG, QS = A.QS, XS = R + 1;
    const uint32_t prows = A.prows;
    uint8_t *buf = smem_raw;
    uint32_t *qstage = (uint32_t *)(smem_raw + A.buf_bytes + 16);  // [R][QS] counter rows, four per word
    uint32_t *xstage = qstage + R * QS;                            // [5 * NW][XS] class planes; also the segment's
    uint32_t *stage = xstage;                                      //   outgoing histogram [4][CG][QW]
    uint32_t *priv = xstage + A.xstage_words;                      // [prows][MG_TPB] private byte counters
    uint32_t *hist = priv + prows * MG_TPB;                        // [W][12] phred bins of this CTA
    uint32_t *bhist = hist + W * MG_QBINS;                         // [NW * 32][5] base counts of this CTA

    for (uint32_t i = tid; i < 128; i += MG_TPB) s_err[i] = (i >= 33 && i < 127) ? A.err_tab[i - 33] : 0.0;
    for (uint32_t i = tid; i < 94; i += MG_TPB) {
        s_edge[i] = A.edges[i];
        s_mp[i] = 0;
    }
    for (uint32_t i = tid; i < 101; i += MG_TPB) s_gc[i] = 0;
    if (A.do_ad) {
        for (uint32_t i = tid; i < A.n_adapters * FH_MAX_PAT; i += MG_TPB)
            s_pat[i] = A.pat[(i / FH_MAX_PAT) * AD_MAXLEN + (i % FH_MAX_PAT)];
        for (uint32_t i = tid; i < A.n_adapters; i += MG_TPB) s_plen[i] = A.plen[i];
    }
    for (uint32_t i = tid; i < prows * MG_TPB + W * MG_QBINS + NW * 32 * 5; i += MG_TPB) priv[i] = 0;
    if (tid == 0) {
        s_lmin = 0xFFFFFFFFu;
        s_lmax = 0;
        s_oob = 0;
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const BatchView &bv = A.bv;
    const uint32_t TS = A.TS, n_segs = A.n_segs, n_tiles = A.n_tiles;
    // roles of phase 1
    const uint32_t role = tid / R, ri = tid - role * R;  // role 0: sequence side, 1: quality side, >= 2: idle
    // base histogram: plane word j = class * NW + word, records [part * PR, (part + 1) * PR)
    const bool reducer = A.do_qc && tid < 5 * NW * A.PARTS;
    const uint32_t red_j = tid / A.PARTS, red_part = tid - red_j * A.PARTS;
    uint32_t acc[12];
#pragma unroll
    for (int l = 0; l < 12; l++) acc[l] = 0;
    uint32_t acc_tiles = 0;
    // phase 2
    const uint32_t rg = tid / CG, cg = tid - rg * CG, col0 = cg * 4;
    const bool worker = A.do_hist && rg < RG;
    uint8_t *priv8 = (uint8_t *)priv + tid * 4;
    // quality byte -> counter row, four at a time
    const uint32_t base4 = (A.hg.qbase + 33u) * 0x01010101u;
    const uint32_t hi4 = (0x80u - A.hg.qrows) * 0x01010101u;
    const uint32_t trash4 = (A.pt_rows ? A.hg.qrows : (uint32_t)MG_QBINS) * 0x01010101u;
    uint32_t lmin = 0xFFFFFFFFu, lmax = 0;
    uint32_t qmn = 0xFFu, qmx = 0;  // quality bytes seen in the sampled words

    auto flush_bases = [&]() {
        // bit-sliced counters -> CTA histogram (class c of positions 32 * pw .. 32 * pw + 31)
        if (reducer) {
            const uint32_t c = red_j / NW, pw = red_j - c * NW;
            for (uint32_t b = 0; b < 32; b++) {
                uint32_t cnt = 0;
#pragma unroll
                for (int l = 0; l < 12; l++) cnt |= ((acc[l] >> b) & 1u) << l;
                if (cnt) atomicAdd(bhist + (pw * 32 + b) * 5 + c, cnt);
            }
#pragma unroll
            for (int l = 0; l < 12; l++) acc[l] = 0;
        }
        acc_tiles = 0;
    };

    // descriptors of the tile about to be processed (record threads), fetched one tile ahead
    uint32_t d_no = 0, d_so = 0, d_qo = 0, d_L = 0, d_gs = 0;
    auto tile_range = [&](uint32_t t, uint64_t *gstart, uint32_t *bytes) {
        const uint32_t r0 = t * R, r1 = min(r0 + R, bv.n);
        const uint64_t start = (uint64_t)bv.name_off[r0] - 1;
        const uint64_t end = r1 < bv.n ? (uint64_t)bv.name_off[r1] - 1 : A.text_end;
        *gstart = start & ~15ULL;
        *bytes = (uint32_t)(((end + 15) & ~15ULL) - *gstart);
    };
    auto issue_tile = [&](uint32_t t) {  // thread 0
        uint64_t gstart;
        uint32_t bytes;
        tile_range(t, &gstart, &bytes);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&bar, bytes);
        bulk_g2s(buf, bv.text + gstart, bytes, &bar);
    };
    auto fetch_desc = [&](uint32_t t) {  // record threads of both roles
        const uint32_t r0 = t * R, r1 = min(r0 + R, bv.n);
        const uint32_t r = r0 + ri;
        d_L = 0;
        if (role < 2 && r < r1) {
            const uint32_t gs = (uint32_t)(((uint64_t)bv.name_off[r0] - 1) & ~15ULL);
            d_gs = gs;
            d_no = bv.name_off[r];
            d_so = bv.seq_off[r];
            d_qo = bv.qual_off[r];
            d_L = bv.seq_len[r];
        }
    };

    uint32_t parity = 0;
    if (blockIdx.x < n_segs) {
        if (tid == 0) issue_tile(blockIdx.x * TS);
        fetch_desc(blockIdx.x * TS);
    }
    for (uint32_t seg = blockIdx.x; seg < n_segs; seg += gridDim.x) {
        const uint32_t t_end = min((seg + 1) * TS, n_tiles);
        for (uint32_t t = seg * TS; t < t_end; t++) {
            const uint32_t r0 = t * R, r1 = min(r0 + R, bv.n), nrec = r1 - r0;
            const uint32_t r = r0 + ri;
            const bool active = role < 2 && ri < nrec;
            const uint32_t L = d_L;
            const uint32_t no = d_no - d_gs, so = d_so - d_gs, qo = d_qo - d_gs, name_len = d_so - 1 - d_no;
            if (active && role == 0) {
                lmin = min(lmin, L);
                lmax = max(lmax, L);
            }
            mbar_wait(&bar, parity);
            parity ^= 1;
            if (role == 0) {
                // ---- sequence planes ------------------------------------------------------------------
                uint32_t X[5][NW + 1];
#pragma unroll
                for (int c = 0; c < 5; c++) X[c][NW] = 0;
                uint32_t gc = 0, valid = 0;
#pragma unroll
                for (int pw = 0; pw < NW; pw++) {
                    uint32_t v = 0, h = 0, g = 0;
                    if (active && (uint32_t)pw * 32 < L) {
#pragma unroll
                        for (int k = 0; k < 8; k++) {
                            const uint32_t pos = pw * 32 + k * 4;
                            if (pos < L) {
                                const uint32_t w = fh_word(buf, so, pos >> 2);
                                const uint32_t nvalid = min(4u, L - pos);
                                const uint32_t pm = 0x80808080u >> (8 * (4 - nvalid));
                                const uint32_t vb = fh_acgt_bytes80(w) & pm;
                                const uint32_t hb = (w << 6) & vb, gb = (w << 5) & vb;  // bit 1 / bit 2 of the letter
                                const uint32_t nib = 0xFu << (k * 4);
                                v |= ((vb * 0x00204081u) >> (28 - k * 4)) & nib;
                                h |= ((hb * 0x00204081u) >> (28 - k * 4)) & nib;
                                g |= ((gb * 0x00204081u) >> (28 - k * 4)) & nib;
                            }
                        }
                    }
                    const uint32_t P = !active ? 0u
                                       : L >= (uint32_t)(pw + 1) * 32 ? 0xFFFFFFFFu
                                       : (L > (uint32_t)pw * 32 ? (1u << (L - pw * 32)) - 1 : 0u);
                    X[0][pw] = v & ~h & ~g;  // A 0x41: bit 1 = 0, bit 2 = 0
                    X[1][pw] = h & ~g;       // C 0x43: bit 1 = 1, bit 2 = 0   (h, g are subsets of v)
                    X[2][pw] = h & g;        // G 0x47
                    X[3][pw] = g & ~h;       // T 0x54
                    X[4][pw] = ~v & P;       // anything else inside the read
                    gc += __popc(h);
                    valid += __popc(v);
                }
                if (A.do_qc) {
                    // GC bucket (:2045-2058)
                    if (active && valid) {
                        const double pct = (double)gc * 100.0 / (double)valid;
                        atomicAdd(&s_gc[(uint32_t)round(pct)], 1u);
                    }
#pragma unroll
                    for (int c = 0; c < 5; c++)
#pragma unroll
                        for (int pw = 0; pw < NW; pw++) xstage[(c * NW + pw) * XS + ri] = X[c][pw];
                }
                // ---- adapters: first occurrence per adapter (:2786-2823) ---------------------------------
                if (A.do_ad) {
                    for (uint32_t a0 = 0; a0 < A.n_adapters; a0 += MG_GA) {
                        uint32_t M[MG_GA][NW];
                        uint32_t mlen[MG_GA];
                        uint32_t maxm = 0;
#pragma unroll
                        for (int q = 0; q < MG_GA; q++) {
                            const uint32_t m = a0 + q < A.n_adapters ? s_plen[a0 + q] : 0u;
                            mlen[q] = m;
                            maxm = max(maxm, m);
                            const uint32_t init = (m == 0 || !active) ? 0u : 0xFFFFFFFFu;
#pragma unroll
                            for (int pw = 0; pw < NW; pw++) M[q][pw] = init;
                        }
                        for (uint32_t j = 0; j < maxm; j++) {
                            uint32_t S[5][NW];
#pragma unroll
                            for (int c = 0; c < 5; c++)
#pragma unroll
                                for (int pw = 0; pw < NW; pw++) S[c][pw] = __funnelshift_r(X[c][pw], X[c][pw + 1], j);
#pragma unroll
                            for (int q = 0; q < MG_GA; q++) {
                                if (j >= mlen[q]) continue;  // (uniform: the adapter set is the same for every lane)
                                const uint32_t c = s_pat[(a0 + q) * FH_MAX_PAT + j];
                                switch (c) {
                                case 0:
#pragma unroll
                                    for (int pw = 0; pw < NW; pw++) M[q][pw] &= S[0][pw];
                                    break;
                                case 1:
#pragma unroll
                                    for (int pw = 0; pw < NW; pw++) M[q][pw] &= S[1][pw];
                                    break;
                                case 2:
#pragma unroll
                                    for (int pw = 0; pw < NW; pw++) M[q][pw] &= S[2][pw];
                                    break;
                                case 3:
#pragma unroll
                                    for (int pw = 0; pw < NW; pw++) M[q][pw] &= S[3][pw];
                                    break;
                                default:
#pragma unroll
                                    for (int pw = 0; pw < NW; pw++) M[q][pw] &= S[4][pw];
                                    break;
                                }
                            }
                        }
#pragma unroll
                        for (int q = 0; q < MG_GA; q++) {
                            uint32_t any = 0;
#pragma unroll
                            for (int pw = 0; pw < NW; pw++) any |= M[q][pw];
                            if (!any) continue;
                            uint32_t p = 0xFFFFFFFFu;
#pragma unroll
                            for (int pw = NW - 1; pw >= 0; pw--)
                                if (M[q][pw]) p = pw * 32 + (__ffs(M[q][pw]) - 1);
                            uint64_t *fwd = A.ad_counts + (size_t)(a0 + q) * 2 * A.ad_cap_len;
                            atomic_add_u64(fwd + p, 1);
                            atomic_add_u64(fwd + A.ad_cap_len + (L - 1 - p), 1);
                        }
                    }
                }
            }
            else if (role == 1) {
                // ---- ordered error sum, mean-phred bucket (:2059-2137) + counter rows of every word ----
                uint32_t *qrow = qstage + ri * QS;
                if (A.do_qc | A.do_hist) {
                    const uint32_t nit = (active && L >= 5) ? (L - 1) / 4 : 0;
                    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                    uint32_t badw = 0;
                    auto rows_of = [&](uint32_t raw, uint32_t keep, uint32_t k) -> uint32_t {
                        // four quality bytes -> four counter rows; bytes past the end of the read (keep = 0)
                        // and, for PerTileQuality, qualities outside the tabulated rows go to the trash row
                        if (A.pt_rows) {
                            const uint32_t d = (raw | 0x80808080u) - base4;  // bit 7: quality >= first row
                            const uint32_t idx = d & 0x7F7F7F7Fu;
                            const uint32_t e = idx + hi4;                     // bit 7: row >= qrows
                            const uint32_t bad = (~d | e) & 0x80808080u & keep;
                            uint32_t row4 = (idx & keep) | (trash4 & ~keep);
                            if (bad) {
                                s_oob = 1;
#pragma unroll
                                for (int j = 0; j < 4; j++) {
                                    if (!((bad >> (8 * j)) & 0x80u)) continue;
                                    const uint32_t c = (raw >> (8 * j)) & 0xFF;
                                    row4 = (row4 & ~(0xFFu << (8 * j))) | (A.hg.qrows << (8 * j));
                                    if (c - 33u > 93u) {  // PerTileQuality raises for it (:3213)
                                        if (A.do_pt)
                                            atomicMin(&A.pt_st->err_key,
                                                      (unsigned long long)((A.pt_base + r) << 8 | c));
                                    }
                                    else if (A.do_qc)
                                        atomicAdd(hist + (k * 4 + j) * MG_QBINS + (min(c - 33u, 47u) >> 2), 1u);
                                }
                            }
                            return row4;
                        }
                        // phred bins min(q, 47) >> 2, four at once (q >= 48 saturates to bin 11)
                        const uint32_t q = (raw & keep) | (0x7F7F7F7Fu & ~keep);
                        const uint32_t t4 = q - 0x21212121u;
                        const uint32_t sat = (((t4 + 0x50505050u) >> 7) & 0x01010101u) * 0xFFu;
                        const uint32_t bin4 = (((t4 >> 2) & 0x0F0F0F0Fu) & ~sat) | (0x0B0B0B0Bu & sat);
                        return (bin4 & keep) | (trash4 & ~keep);
                    };
#pragma unroll 4
                    for (uint32_t i = 0; i < nit; i++) {
                        const uint32_t w = fh_word(buf, qo, i);
                        badw |= (w - 0x21212121u) | (w + 0x01010101u);
                        a0 += s_err[w & 0xFF];
                        a1 += s_err[(w >> 8) & 0xFF];
                        a2 += s_err[(w >> 16) & 0xFF];
                        a3 += s_err[w >> 24];
                        if (A.do_hist) qrow[i] = rows_of(w, 0xFFFFFFFFu, i);
                    }
                    double sum = ((a0 + a1) + a2) + a3;  // :2098-2099
                    uint32_t nwords = 0;
                    if (active && L) {
                        uint32_t w = fh_word(buf, qo, nit);
                        const uint32_t ntail = L - 4 * nit;  // 1..4, in order (:2100-2112)
                        const uint32_t keep = 0xFFFFFFFFu >> (8 * (4 - ntail));
                        w &= keep;
                        if (A.do_hist) qrow[nit] = rows_of(w, keep, nit);
                        nwords = nit + 1;
                        w |= ntail < 4 ? 0x21212121u << (8 * ntail) : 0u;  // padding reads as '!' for the check only
                        badw |= (w - 0x21212121u) | (w + 0x01010101u);
                        sum += s_err[w & 0xFF];
                        if (ntail > 1) sum += s_err[(w >> 8) & 0xFF];
                        if (ntail > 2) sum += s_err[(w >> 16) & 0xFF];
                        if (ntail > 3) sum += s_err[w >> 24];
                    }
                    if (A.do_hist && ri < R)
                        for (uint32_t k = nwords; k < CG; k++) qrow[k] = trash4;  // shorter reads: nothing to count
                    if (active && A.do_qc) {
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
                }
                // ---- fingerprint hash (:4463-4485) -------------------------------------------------------
                if (active && A.do_dd) {
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
                if (active && A.do_pt) {
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
                    // range of the quality values, from two words per read at positions that sweep the read
                    // length over consecutive reads: sizes the counter rows of the NEXT record array
                    if (L) {
                        const uint32_t nw = L >> 2;
                        uint32_t w0, w1;
                        if (nw) {
                            const uint32_t i0 = r % nw, i1 = i0 + (nw >> 1) >= nw ? i0 + (nw >> 1) - nw : i0 + (nw >> 1);
                            w0 = fh_word(buf, qo, i0);
                            w1 = fh_word(buf, qo, i1);
                        }
                        else {
                            w0 = fh_word(buf, qo, 0);
                            const uint32_t b0 = w0 & 0xFF;
                            w0 = L == 1 ? b0 * 0x01010101u : L == 2 ? (w0 & 0xFFFF) * 0x00010001u : (w0 & 0xFFFFFF) | b0 << 24;
                            w1 = w0;
                        }
                        const uint32_t mn = __vminu4(w0, w1), mx = __vmaxu4(w0, w1);
                        qmn = min(qmn, min(min(mn & 0xFF, (mn >> 8) & 0xFF), min((mn >> 16) & 0xFF, mn >> 24)));
                        qmx = max(qmx, max(max(mx & 0xFF, (mx >> 8) & 0xFF), max((mx >> 16) & 0xFF, mx >> 24)));
                    }
                }
            }
            __syncthreads();  // planes and counter rows are staged; nobody reads the text tile any more
            {
                // the next tile's copy and descriptors travel while phase 2 runs
                uint32_t nt = t + 1 < t_end ? t + 1 : (seg + gridDim.x < n_segs ? (seg + gridDim.x) * TS : 0xFFFFFFFFu);
                if (nt != 0xFFFFFFFFu) {
                    if (tid == 0) issue_tile(nt);
                    fetch_desc(nt);
                }
            }
            // ---- base histogram: bit-sliced vertical counters over the staged class planes ----------------
            if (reducer) {
                const uint32_t *src = xstage + red_j * XS + red_part * A.PR;
                const uint32_t left = red_part * A.PR < R ? min(A.PR, R - red_part * A.PR) : 0u;
                uint32_t ones = acc[0], twos = acc[1], fours = acc[2];
#pragma unroll
                for (int half = 0; half < 2; half++) {
                    uint32_t w[8];
#pragma unroll
                    for (int k = 0; k < 8; k++) w[k] = (uint32_t)(half * 8 + k) < left ? src[half * 8 + k] : 0u;
                    uint32_t t2a, t2b, t4a, t4b, e8;
                    csa(t2a, ones, ones, w[0], w[1]);
                    csa(t2b, ones, ones, w[2], w[3]);
                    csa(t4a, twos, twos, t2a, t2b);
                    csa(t2a, ones, ones, w[4], w[5]);
                    csa(t2b, ones, ones, w[6], w[7]);
                    csa(t4b, twos, twos, t2a, t2b);
                    csa(e8, fours, fours, t4a, t4b);
#pragma unroll
                    for (int l = 3; l < 12; l++) {  // ripple the carry into the eights and above
                        const uint32_t nv = acc[l] ^ e8;
                        e8 &= acc[l];
                        acc[l] = nv;
                    }
                }
                acc[0] = ones;
                acc[1] = twos;
                acc[2] = fours;
            }
            if (++acc_tiles == 255) flush_bases();  // 16 reads per tile and counter at most: 4080 < 4096
            // ---- phase 2: private byte counters, one aligned load per record ------------------------------
            if (worker && rg < nrec) {
                uint32_t nrow = qstage[rg * QS + cg];
                for (uint32_t i = rg; i < nrec; i += RG) {
                    const uint32_t row4 = nrow;
                    nrow = qstage[min(i + RG, nrec - 1) * QS + cg];
                    // the four counters are different bytes (byte lane j): load all, then store all
                    uint8_t *p0 = priv8 + (row4 & 0xFF) * (MG_TPB * 4) + 0;
                    uint8_t *p1 = priv8 + ((row4 >> 8) & 0xFF) * (MG_TPB * 4) + 1;
                    uint8_t *p2 = priv8 + ((row4 >> 16) & 0xFF) * (MG_TPB * 4) + 2;
                    uint8_t *p3 = priv8 + (row4 >> 24) * (MG_TPB * 4) + 3;
                    const uint32_t c0 = *p0, c1 = *p1, c2 = *p2, c3 = *p3;
                    *p0 = (uint8_t)(c0 + 1);
                    *p1 = (uint8_t)(c1 + 1);
                    *p2 = (uint8_t)(c2 + 1);
                    *p3 = (uint8_t)(c3 + 1);
                }
            }
            __syncthreads();  // staging matrices are free again
        }
        if (!A.do_hist) continue;
        // ---- end of a segment: combine the row groups' private counters ------------------------------
        if (worker) {
            if (A.pt_rows) {
                // thread (rg, cg) owns the quads of rows 4m .. 4m+3 with m = rg, rg + RG, ...: one phred bin each
                const uint32_t nq = A.hg.qrows >> 2, QW = A.hg.QW;
                for (uint32_t m = rg; m < nq; m += RG) {
                    uint32_t w[4];
#pragma unroll
                    for (int rr = 0; rr < 4; rr++) {
                        uint32_t *pr = priv + (4 * m + rr) * MG_TPB + cg;
                        uint32_t a = 0;
                        for (uint32_t g = 0; g < RG; g++) {  // counts of a segment fit a byte: packed adds
                            a += pr[g * CG];
                            pr[g * CG] = 0;
                        }
                        w[rr] = a;
                    }
                    const uint32_t qsum = w[0] + w[1] + w[2] + w[3];
                    if (A.do_qc && qsum) {
                        const uint32_t bin = min((A.hg.qbase >> 2) + m, 11u);
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const uint32_t c = (qsum >> (8 * j)) & 0xFF;
                            if (c) atomicAdd(hist + (col0 + j) * MG_QBINS + bin, c);
                        }
                    }
                    if (A.do_pt) {
                        // 4 x 4 byte transpose: word j = counts of rows 4m .. 4m+3 at position col0 + j
                        const uint32_t t0 = __byte_perm(w[0], w[1], 0x5140), t1 = __byte_perm(w[2], w[3], 0x5140);
                        const uint32_t t2 = __byte_perm(w[0], w[1], 0x7362), t3 = __byte_perm(w[2], w[3], 0x7362);
                        stage[(0 * CG + cg) * QW + m] = __byte_perm(t0, t1, 0x5410);
                        stage[(1 * CG + cg) * QW + m] = __byte_perm(t0, t1, 0x7632);
                        stage[(2 * CG + cg) * QW + m] = __byte_perm(t2, t3, 0x5410);
                        stage[(3 * CG + cg) * QW + m] = __byte_perm(t2, t3, 0x7632);
                    }
                }
                if (A.do_pt && rg == 0 && QW > nq) {  // padding word of every position: zero for the reader
#pragma unroll
                    for (int j = 0; j < 4; j++) stage[(j * CG + cg) * QW + nq] = 0;
                }
                // the trash row only ever counts padding and slow-path bytes
                priv[A.hg.qrows * MG_TPB + tid] = 0;
            }
            else {
                // rows are the 12 phred bins: bin b of column group cg has one owner, plain adds
                for (uint32_t b = rg; b < (uint32_t)MG_QBINS; b += RG) {
                    uint32_t *pr = priv + b * MG_TPB + cg;
                    uint32_t a = 0;
                    for (uint32_t g = 0; g < RG; g++) {
                        a += pr[g * CG];
                        pr[g * CG] = 0;
                    }
                    if (a) {
#pragma unroll
                        for (int j = 0; j < 4; j++) hist[(col0 + j) * MG_QBINS + b] += (a >> (8 * j)) & 0xFF;
                    }
                }
                priv[MG_QBINS * MG_TPB + tid] = 0;
            }
        }
        if (A.pt_rows && A.do_pt) {
            __syncthreads();
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                bulk_s2g(A.qh + (uint64_t)seg * A.hg.seg_bytes, stage, A.hg.seg_bytes);
                if (s_oob) A.seg_oob[seg] = 1;
                s_oob = 0;
                bulk_wait_read();  // the class planes of the next tile land in the same shared memory
            }
        }
        else if (tid == 0) s_oob = 0;
        __syncthreads();
    }
    if (A.do_pt) {
        qmx = warp_max_u32(qmx);
        qmn = ~warp_max_u32(~qmn);
        if (lane_id() == 0 && qmn <= qmx) {
            atomicMin(&A.pt_st->qmin, qmn);
            atomicMax(&A.pt_st->qmax, qmx);
        }
    }
    if (!A.do_qc) return;
    flush_bases();
    lmax = warp_max_u32(lmax);
    lmin = ~warp_max_u32(~lmin);
    if (lane_id() == 0) {
        atomicMin(&s_lmin, lmin);
        atomicMax(&s_lmax, lmax);
    }
    __syncthreads();
    for (uint32_t i = tid; i < 101; i += MG_TPB)
        if (s_gc[i]) atomic_add_u64(A.gc + i, s_gc[i]);
    for (uint32_t i = tid; i < 94; i += MG_TPB)
        if (s_mp[i]) atomic_add_u64(A.mean_phred + i, s_mp[i]);
    // CTA histograms -> global tables
    for (uint32_t i = tid; i < W * MG_QBINS; i += MG_TPB) {
        const uint32_t c = hist[i];
        if (c) atomic_add_u64(A.phred + i, c);  // [pos][12]
    }
    for (uint32_t i = tid; i < W * 5; i += MG_TPB) {
        const uint32_t c = bhist[i];
        if (c) atomic_add_u64(A.base + i, c);  // [pos][5]
    }
    if (s_lmax == 0 && s_lmin == 0xFFFFFFFFu) return;  // this CTA had no tile
    if (s_lmin == s_lmax) {
        // every record had length L0: the end-anchored rows are a shifted window of the histograms
        const uint32_t L0 = s_lmin, ea_n = min(L0, A.ea_len);
        const uint32_t lo = L0 - ea_n;
        for (uint32_t i = tid; i < ea_n * 17; i += MG_TPB) {
            const uint32_t pos = lo + i / 17, k = i % 17;
            const uint32_t c = k < 5 ? bhist[pos * 5 + k] : hist[pos * MG_QBINS + (k - 5)];
            if (!c) continue;
            const uint64_t row = (uint64_t)A.ea_len - (L0 - pos);
            if (k < 5) atomic_add_u64(A.ea_base + row * 5 + k, c);
            else atomic_add_u64(A.ea_phred + row * 12 + (k - 5), c);
        }
    }
    else if (tid == 0) A.cta_mixed[blockIdx.x] = 1;
}

// quality byte range of the first record array PerTileQuality sees (two words per read, as the
// fused kernels sample it for the arrays that follow)
__global__ void __launch_bounds__(256)
k_sample_qrange(BatchView bv, PtState *st) {
    uint32_t qmn = 0xFFu, qmx = 0;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < bv.n; r += gridDim.x * blockDim.x) {
        const uint32_t L = bv.seq_len[r];
        if (!L) continue;
        const uint8_t *q = bv.text + bv.qual_off[r];
        const uint32_t nw = L >> 2;
        uint32_t w0, w1;
        if (nw) {
            const uint32_t i0 = r % nw, i1 = i0 + (nw >> 1) >= nw ? i0 + (nw >> 1) - nw : i0 + (nw >> 1);
            w0 = load_u32_unaligned(q + 4 * i0);
            w1 = load_u32_unaligned(q + 4 * i1);
        }
        else {
            w0 = load_u32_unaligned(q);
            const uint32_t b0 = w0 & 0xFF;
            w0 = L == 1 ? b0 * 0x01010101u : L == 2 ? (w0 & 0xFFFF) * 0x00010001u : (w0 & 0xFFFFFF) | b0 << 24;
            w1 = w0;
        }
        const uint32_t mn = __vminu4(w0, w1), mx = __vmaxu4(w0, w1);
        qmn = min(qmn, min(min(mn & 0xFF, (mn >> 8) & 0xFF), min((mn >> 16) & 0xFF, mn >> 24)));
        qmx = max(qmx, max(max(mx & 0xFF, (mx >> 8) & 0xFF), max((mx >> 16) & 0xFF, mx >> 24)));
    }
    qmx = warp_max_u32(qmx);
    qmn = ~warp_max_u32(~qmn);
    if (lane_id() == 0 && qmn <= qmx) {
        atomicMin(&st->qmin, qmn);
        atomicMax(&st->qmax, qmx);
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
// 192 threads work); as many CTAs share an SM as leave room for a useful text tile.
struct ColGeom {
    uint32_t R, n_tiles, CG, RG, W, buf_bytes, grid, tpb;
    uint32_t TS, n_segs;  // PT: text tiles per segment, segments (R * TS <= 255 records each)
    PtHistGeom hg;        // PT: qrows != 0
    size_t smem;
    bool ok;
};
// qmin / qmax: sampled quality byte range for the PerTileQuality histograms, qmin > qmax: QCMetrics only
static ColGeom col_geometry(sq_ctx *ctx, const sq_batch *b, uint32_t qmin, uint32_t qmax) {
    ColGeom g;
    memset(&g, 0, sizeof(g));
    g.hg = PtHistGeom();
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
    const bool pt = qmin <= qmax;
    uint32_t prows = 16, stage = 0;
    if (pt) {
        const uint32_t lo = (qmin < 33 ? 33 : qmin) - 33, hi = (qmax > 126 ? 126 : qmax) - 33;
        if (lo > hi) return g;
        g.hg.qbase = lo & ~3u;
        g.hg.qrows = (hi - g.hg.qbase + 4) & ~3u;
        g.hg.QW = (g.hg.qrows / 4) | 1u;  // odd: the transposed stores of a warp hit 32 banks
        g.hg.CG = g.CG;
        g.hg.seg_bytes = g.W * g.hg.QW * 4;
        prows = g.hg.qrows + 1;
        stage = g.hg.seg_bytes + 16;
    }
    const uint32_t fixed = g.W * FC_BINS * 4 + prows * g.tpb * 4 + stage + 64;
    const uint32_t per_rec = b->max_rec_bytes + 12;
    uint32_t ctas = g.tpb == 128 ? 4 : g.tpb == 192 ? 3 : 2;
    uint32_t budget = 0;
    for (; ctas >= 1; ctas--) {  // fewer CTAs per SM when the counters leave no room for 24 records
        budget = (228u * 1024u) / ctas - 1024u - 128u;
        if (budget > 227u * 1024u) budget = 227u * 1024u;
        if (fixed + (ctas > 1 ? 24 : 8) * per_rec + 64 <= budget) break;
    }
    if (ctas == 0) return g;
    g.R = (budget - fixed - 64) / per_rec;
    if (g.R > 255) g.R = 255;
    g.TS = pt ? 255 / g.R : 1;
    g.buf_bytes = (g.R * b->max_rec_bytes + 32 + 15) & ~15u;
    g.smem = (size_t)g.buf_bytes + 16 + (size_t)3 * g.R * 4 + fixed;
    g.n_tiles = (uint32_t)((b->n + g.R - 1) / g.R);
    g.n_segs = (g.n_tiles + g.TS - 1) / g.TS;
    g.grid = (uint32_t)ctx->num_sms * ctas;
    if (g.grid > g.n_segs) g.grid = g.n_segs;
    g.ok = true;
    return g;
}

template <int TPB, bool PT>
static int launch_columns_t(sq_ctx *ctx, const ColGeom &g, const ColumnArgs &C) {
    // (named for the profiler: with / without the per-segment quality histograms)
    if constexpr (PT) {
        auto k_fused_columns_qhist = k_fused_columns<TPB, true>;
        CUDA_TRY(cudaFuncSetAttribute(k_fused_columns_qhist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
        SQ_LAUNCH(ctx, k_fused_columns_qhist, g.grid, TPB, g.smem, C);
    }
    else {
        auto k_fused_columns_bins = k_fused_columns<TPB, false>;
        CUDA_TRY(cudaFuncSetAttribute(k_fused_columns_bins, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
        SQ_LAUNCH(ctx, k_fused_columns_bins, g.grid, TPB, g.smem, C);
    }
    return SQ_OK;
}
static int launch_columns(sq_ctx *ctx, const ColGeom &g, const ColumnArgs &C) {
    if (g.hg.qrows) {
        if (g.tpb == 128) return launch_columns_t<128, true>(ctx, g, C);
        if (g.tpb == 192) return launch_columns_t<192, true>(ctx, g, C);
        return launch_columns_t<256, true>(ctx, g, C);
    }
    if (g.tpb == 128) return launch_columns_t<128, false>(ctx, g, C);
    if (g.tpb == 192) return launch_columns_t<192, false>(ctx, g, C);
    return launch_columns_t<256, false>(ctx, g, C);
}


// ---- one-walk plan --------------------------------------------------------------------------
struct WalkGeom {
    uint32_t R, n_tiles, TS, n_segs, CG, RG, W, QS, PARTS, PR, buf_bytes, xstage_words, prows, grid, NW;
    PtHistGeom hg;  // qrows != 0: counter rows are quality values
    size_t smem;
    bool ok;
};
// qmin / qmax: sampled quality byte range for the PerTileQuality histograms, qmin > qmax: phred bins only
static WalkGeom walk_geometry(sq_ctx *ctx, const sq_batch *b, uint32_t qmin, uint32_t qmax) {
    WalkGeom g;
    memset(&g, 0, sizeof(g));
    g.hg = PtHistGeom();
    if (b->max_len == 0 || b->max_len > 160 || b->max_rec_bytes == 0) return g;
    g.NW = b->max_len <= 96 ? 3 : 5;
    g.CG = (b->max_len + 3) / 4;
    g.W = g.CG * 4;
    g.QS = g.CG | 1u;
    g.RG = MG_TPB / g.CG;
    uint32_t R = (MG_BUF - 48) / b->max_rec_bytes;
    if (R > MG_RMAX) R = MG_RMAX;
    R &= ~31u;
    if (R == 0) return g;
    g.R = R;
    g.TS = 255 / R;
    g.PARTS = MG_TPB / (5 * g.NW);
    if (g.PARTS > R) g.PARTS = R;
    g.PR = (R + g.PARTS - 1) / g.PARTS;
    if (g.PR > 16) return g;
    g.prows = 16;  // 12 phred bins + trash (+ spare rows: a byte below '!' cannot index outside)
    uint32_t seg_words = 0;
    if (qmin <= qmax) {
        const uint32_t lo = (qmin < 33 ? 33 : qmin) - 33, hi = (qmax > 126 ? 126 : qmax) - 33;
        if (lo > hi) return g;
        g.hg.qbase = lo & ~3u;
        g.hg.qrows = (hi - g.hg.qbase + 4) & ~3u;
        g.hg.QW = (g.hg.qrows / 4) | 1u;  // odd: the transposed stores of a warp hit 32 banks
        g.hg.CG = g.CG;
        g.hg.seg_bytes = g.W * g.hg.QW * 4;
        g.prows = g.hg.qrows + 1;
        seg_words = g.W * g.hg.QW;
    }
    g.buf_bytes = (R * b->max_rec_bytes + 32 + 15) & ~15u;
    g.xstage_words = 5 * g.NW * (R + 1);
    if (g.xstage_words < seg_words) g.xstage_words = seg_words;
    g.xstage_words = (g.xstage_words + 3) & ~3u;
    g.smem = (size_t)g.buf_bytes + 16 + (size_t)R * g.QS * 4 + (size_t)g.xstage_words * 4 +
             (size_t)g.prows * MG_TPB * 4 + (size_t)g.W * MG_QBINS * 4 + (size_t)g.NW * 32 * 5 * 4;
    if (g.smem > 220 * 1024) return g;  // e.g. a quality range of 90 values: the two-kernel path has room for it
    g.n_tiles = (uint32_t)((b->n + R - 1) / R);
    g.n_segs = (g.n_tiles + g.TS - 1) / g.TS;
    const uint32_t ctas = g.smem <= 112 * 1024 ? 2 : 1;
    g.grid = (uint32_t)ctx->num_sms * ctas;
    if (g.grid > g.n_segs) g.grid = g.n_segs;
    g.ok = true;
    return g;
}

static int launch_onewalk(sq_ctx *ctx, const WalkGeom &g, const WalkArgs &A) {
    if (g.NW == 3) {
        CUDA_TRY(cudaFuncSetAttribute(k_onewalk<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
        SQ_LAUNCH(ctx, k_onewalk<3>, g.grid, MG_TPB, g.smem, A);
    }
    else {
        CUDA_TRY(cudaFuncSetAttribute(k_onewalk<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
        SQ_LAUNCH(ctx, k_onewalk<5>, g.grid, MG_TPB, g.smem, A);
    }
    return SQ_OK;
}

// The whole short-read loop body on one walk over the text (k_onewalk), then the table maintenance
// of each collector in the reference's module order.  Returns SQ_OK with *done = false when the
// record array does not fit the kernel's plan (the caller takes the two-kernel path).
static int fused_add_onewalk(sq_ctx *ctx, sq_batch *b, sq_qc *qc, sq_pertile *pt, sq_overrep *ov, sq_nanostats *ns,
                             sq_adapters *ad, sq_dedup *dd, bool *done) {
    *done = false;
    // Opt-in (SQ_ONEWALK=1).  Measured on B200 (profiles/r2d_onewalk_*): 258 warp instructions per read
    // against 306 for the two kernels, but its shared-memory footprint (text tile + staging + private
    // counters) leaves 12 warps per SM and the issue slots 42 % busy: 2.23 ms per 4 Mi reads against
    // 0.90 + 0.86 ms.  Kept as the starting point for a version with the counters split off.
    static const bool enabled = getenv("SQ_ONEWALK") != nullptr;
    if (!enabled) return SQ_OK;
    const uint32_t n = (uint32_t)b->n;
    // quality range for PerTileQuality's counter rows: sampled by the kernels over the arrays seen so
    // far; the first array is sampled on its own (one host synchronisation per collector)
    uint32_t qmin = 1, qmax = 0;
    if (pt) {
        if (pt->qmin == 0xFFFFFFFFu)
            SQ_LAUNCH(ctx, k_sample_qrange, sq_grid_for(ctx, n, 256, 8), 256, 0, b->view(), pt->st);
        SQ_TRY(pt_quality_range(pt, &qmin, &qmax));
    }
    WalkGeom g = walk_geometry(ctx, b, qmin, qmax);
    if (!g.ok && pt && qmin <= qmax) g = walk_geometry(ctx, b, 1, 0);  // too many rows: bins here, sort-based PerTileQuality
    if (!g.ok) return SQ_OK;
    *done = true;

    WalkArgs A;
    memset(&A, 0, sizeof(A));
    A.bv = b->view();
    A.R = g.R;
    A.n_tiles = g.n_tiles;
    A.TS = g.TS;
    A.n_segs = g.n_segs;
    A.text_end = b->text_end;
    A.CG = g.CG;
    A.RG = g.RG;
    A.W = g.W;
    A.QS = g.QS;
    A.PARTS = g.PARTS;
    A.PR = g.PR;
    A.buf_bytes = g.buf_bytes;
    A.xstage_words = g.xstage_words;
    A.prows = g.prows;
    A.hg = g.hg;
    A.pt_rows = g.hg.qrows != 0;
    A.err_tab = ctx->d_err_table;
    A.edges = ctx->d_phred_thresholds;
    long long *tile = nullptr;
    uint64_t *hashes = nullptr;
    uint8_t *cta_mixed = nullptr, *qh = nullptr, *oob = nullptr;
    SqScratch scratch(ctx);
    int rc = SQ_OK;
    if (qc) {
        rc = qc_grow(qc, b->max_len);
        if (rc == SQ_OK) rc = scratch.get(&cta_mixed, g.grid, true);
        A.do_qc = 1;
        A.gc = qc->gc;
        A.mean_phred = qc->mean_phred;
        A.qc_err_key = qc->err_key;
        A.qc_base = qc->n_reads;
        A.base = qc->base;
        A.phred = qc->phred;
        A.ea_base = qc->ea_base;
        A.ea_phred = qc->ea_phred;
        A.ea_len = (uint32_t)qc->ea_len;
        A.cta_mixed = cta_mixed;
    }
    if (rc == SQ_OK && ad) {
        rc = adapters_grow(ad, b->max_len);
        A.do_ad = 1;
        A.pat = ad->pat;
        A.plen = ad->plen;
        A.n_adapters = ad->n_adapters;
        A.ad_counts = ad->counts;
        A.ad_cap_len = ad->cap_len;
    }
    if (rc == SQ_OK && dd) {
        rc = scratch.get(&hashes, (size_t)n * 8);
        A.do_dd = 1;
        A.front_len = dd->front_len;
        A.back_len = dd->back_len;
        A.front_off = dd->front_off;
        A.back_off = dd->back_off;
        A.hashes = hashes;
    }
    const uint64_t pt_base = pt ? pt->n_added : 0;
    if (rc == SQ_OK && pt) {
        rc = scratch.get(&tile, (size_t)n * 8);
        A.do_pt = 1;
        A.tile = tile;
        A.pt_base = pt_base;
        A.pt_st = pt->st;
        if (rc == SQ_OK && A.pt_rows) {
            rc = scratch.get(&qh, (size_t)g.n_segs * g.hg.seg_bytes);
            if (rc == SQ_OK) rc = scratch.get(&oob, g.n_segs, true);
            A.qh = qh;
            A.seg_oob = oob;
        }
    }
    A.do_hist = (qc || (pt && A.pt_rows)) ? 1 : 0;
    if (rc == SQ_OK) rc = launch_onewalk(ctx, g, A);
    if (rc == SQ_OK && qc && qc->ea_len) {
        const size_t ea_smem = (size_t)qc->ea_len * FC_BINS * 4;
        const int use_smem = ea_smem <= 96 * 1024;
        if (use_smem)
            CUDA_TRY(cudaFuncSetAttribute(k_fused_ea_fallback, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        // same CTA -> records assignment as k_onewalk: whole segments
        SQ_LAUNCH(ctx, k_fused_ea_fallback, g.grid, FC_TPB, use_smem ? ea_smem : 0, b->view(), g.R * g.TS, g.n_segs,
                  cta_mixed, qc->ea_base, qc->ea_phred, (uint32_t)qc->ea_len, use_smem);
    }
    if (qc) {
        qc->n_reads += n;
        if (b->max_len > qc->max_len) qc->max_len = b->max_len;
        b->err_sum_valid = true;
    }
    // ---- table maintenance, in the reference's module order ------------------------------------
    if (pt) {
        PtPlan plan;
        if (rc == SQ_OK) {  // (the plan takes over qh / oob and frees them)
            scratch.keep(qh);
            scratch.keep(oob);
            rc = pt_prepare(pt, b, tile, A.pt_rows ? g.R * g.TS : 0, g.n_segs, g.W, g.hg, &plan, qh, oob);
        }
        if (rc == SQ_OK) rc = pt_finish(pt, b, &plan);
        else pt_plan_free(ctx, &plan);
    }
    if (rc == SQ_OK && ov) rc = sq_overrep_add(ov, b);
    if (rc == SQ_OK && ns) rc = sq_nanostats_add(ns, b);
    if (ad) {
        ad->n_seqs += n;
        if (b->max_len > ad->max_len) ad->max_len = b->max_len;
    }
    if (rc == SQ_OK && dd) {
        rc = dedup_consume(dd, hashes, n);
        if (dd->deferred && rc == SQ_OK) scratch.keep(hashes);  // a deferred estimator keeps them
    }
    return rc;
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
    {
        bool done = false;
        SQ_TRY(fused_add_onewalk(ctx, b, qc, pt, ov, ns, ad, dd, &done));
        if (done) return SQ_OK;
    }
    const uint32_t n = (uint32_t)b->n;
    FusedArgs A;
    memset(&A, 0, sizeof(A));
    A.bv = b->view();
    uint32_t rpt = (FH_BUF - 32) / b->max_rec_bytes;
    if (rpt > FH_RECS) rpt = FH_RECS;
    A.recs_per_tile = rpt;
    A.n_tiles = (n + rpt - 1) / rpt;
    A.text_end = b->text_end;
    A.err_tab = ctx->d_err_table;
    A.edges = ctx->d_phred_thresholds;
    long long *tile = nullptr;
    uint64_t *hashes = nullptr;
    uint8_t *cta_mixed = nullptr, *all_acgt = nullptr;
    SqScratch scratch(ctx);  // tile / hashes / cta_mixed / all_acgt leave with the function, whichever way
    if (qc) {
        SQ_TRY(scratch.get(&all_acgt, n));
        A.all_acgt = all_acgt;
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
        SQ_TRY(scratch.get(&hashes, (size_t)n * 8));
        A.do_dd = 1;
        A.front_len = dd->front_len;
        A.back_len = dd->back_len;
        A.front_off = dd->front_off;
        A.back_off = dd->back_off;
        A.hashes = hashes;
    }
    if (pt) {
        SQ_TRY(scratch.get(&tile, (size_t)n * 8));
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
    // The hash-table modules (OverrepresentedSequences, NanoStats, DedupEstimator) need nothing but what this kernel
    // leaves (fingerprint hashes, error sums) and the text: from here on they can run on the table stream, beside
    // the per-position pass and PerTileQuality's chain kernel, which are enqueued first (see the end of the function)
    static const bool no_fork = getenv("SQ_NO_TABLE_STREAM") != nullptr;
    const bool forked = rc == SQ_OK && !ctx->profile && !no_fork && sq_cur_stream(ctx) == ctx->stream && (ov || dd || ns);
    OvPendingAdd ov_pending;
    struct OvGuard {  // (an error on the way out must not leave the fragment buffers behind)
        sq_overrep *o;
        OvPendingAdd *pa;
        ~OvGuard() {
            if (o) ov_add_abandon(o, pa);
        }
    } ov_guard{ov, &ov_pending};
    if (forked) {
        CUDA_TRY(cudaEventRecord(ctx->ev_fork, ctx->stream));
        CUDA_TRY(cudaStreamWaitEvent(ctx->tstream, ctx->ev_fork, 0));
        if (ov) {  // the fragment kernel goes to the device now: it runs while the host waits in pt_prepare
            SqStreamScope on_table_stream(ctx->tstream);
            rc = ov_add_begin(ov, b, &ov_pending);
        }
    }

    // ---- PerTileQuality: slots and segments (needs the tile ids) ---------------------------------
    // two plans for the per-position pass: with per-segment quality histograms (reads in tile runs)
    // and without (QCMetrics alone, or tiles in random order -> sort-based PerTileQuality)
    ColGeom g = col_geometry(ctx, b, 1, 0);
    PtPlan plan;
    struct PlanGuard {
        sq_ctx *ctx;
        PtPlan *plan;
        ~PlanGuard() { pt_plan_free(ctx, plan); }  // (a plan that pt_finish consumed is empty by then)
    } plan_guard{ctx, &plan};
    const uint64_t pt_base = pt ? pt->n_added : 0;
    if (rc == SQ_OK && pt) {
        uint32_t qmin = 1, qmax = 0;
        rc = pt_quality_range(pt, &qmin, &qmax);
        ColGeom gp;
        memset(&gp, 0, sizeof(gp));
        if (rc == SQ_OK && qmin <= qmax) gp = col_geometry(ctx, b, qmin, qmax);
        if (rc == SQ_OK)
            rc = pt_prepare(pt, b, tile, gp.ok ? gp.R * gp.TS : 0, gp.n_segs, gp.W, gp.hg, &plan, nullptr, nullptr);
        if (rc == SQ_OK && plan.runs) g = gp;
    }

    // ---- per-position pass: QCMetrics histograms + PerTileQuality quality histograms ----------
    if (rc == SQ_OK && g.ok && (qc || plan.runs)) {
        ColumnArgs C;
        memset(&C, 0, sizeof(C));
        C.bv = b->view();
        C.recs_per_tile = g.R;
        C.n_tiles = g.n_tiles;
        C.tiles_per_seg = g.TS;
        C.n_segs = g.n_segs;
        C.text_end = b->text_end;
        C.CG = g.CG;
        C.RG = g.RG;
        C.W = g.W;
        C.buf_bytes = g.buf_bytes;
        if (qc) {
            rc = qc_grow(qc, b->max_len);
            if (rc == SQ_OK) rc = scratch.get(&cta_mixed, g.grid, true);
            C.do_qc = 1;
            C.base = qc->base;
            C.phred = qc->phred;
            C.ea_base = qc->ea_base;
            C.ea_phred = qc->ea_phred;
            C.ea_len = (uint32_t)qc->ea_len;
            C.cta_mixed = cta_mixed;
            C.all_acgt = all_acgt;
        }
        if (plan.runs) {
            C.hg = g.hg;
            C.qh = plan.qh;
            C.seg_uniform = plan.uniform;
            C.seg_oob = plan.oob;
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
            // same CTA -> records assignment as k_fused_columns: whole segments
            SQ_LAUNCH(ctx, k_fused_ea_fallback, g.grid, FC_TPB, use_smem ? ea_smem : 0, b->view(), g.R * g.TS,
                      g.n_segs, cta_mixed, qc->ea_base, qc->ea_phred, (uint32_t)qc->ea_len, use_smem);
        }
    }
    else if (rc == SQ_OK && qc) rc = qc_add_vertical(qc, b);
    if (qc) {
        qc->n_reads += n;
        if (b->max_len > qc->max_len) qc->max_len = b->max_len;
        b->err_sum_valid = true;
    }
    // ---- table maintenance ---------------------------------------------------------------------------
    // PerTileQuality's ordered sums stay on the launch stream, behind the per-position pass that feeds them
    if (pt) {
        if (rc == SQ_OK) rc = pt_finish(pt, b, &plan);
        else pt_plan_free(ctx, &plan);
    }
    // ... and everything above is only ENQUEUED by now (the last host wait was in pt_prepare): while the host goes
    // through the table modules below, with their waits, on the table stream, the device works on both streams.
    // The launch stream joins before this function returns, so nothing else ever sees the fork.  Module order as in
    // the reference's loop (the modules do not read each other's state).
    {
        SqStreamScope on_table_stream(forked ? ctx->tstream : nullptr);
        if (rc == SQ_OK && ov) {
            if (!forked) rc = ov_add_begin(ov, b, &ov_pending);
            if (rc == SQ_OK) rc = ov_add_end(ov, &ov_pending);
        }
        if (rc == SQ_OK && ns) rc = sq_nanostats_add(ns, b);
        if (rc == SQ_OK && dd) rc = dedup_consume(dd, hashes, n);
        if (forked && cudaEventRecord(ctx->ev_join, ctx->tstream) != cudaSuccess && rc == SQ_OK)
            rc = sq_cuda_fail(cudaGetLastError(), "end of the table stream's part", __FILE__, __LINE__);
    }
    if (ad) {
        ad->n_seqs += n;
        if (b->max_len > ad->max_len) ad->max_len = b->max_len;
    }
    if (dd && dd->deferred && rc == SQ_OK) scratch.keep(hashes);  // a deferred estimator keeps them
    if (forked && cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0) != cudaSuccess && rc == SQ_OK)
        rc = sq_cuda_fail(cudaGetLastError(), "join of the table stream", __FILE__, __LINE__);
    return rc;
}
