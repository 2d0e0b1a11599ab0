// qc.cu -- QCMetrics (reference _qcmodule.c:1966-2139) on the device.
//
// Two passes over the record array (the text stays L2/HBM resident):
//
//  k_qc_vertical    "a thread owns four read positions": every thread walks the
//                   rows (reads) of its CTA and keeps the per-position counters
//                   private -- base classes as five byte-sliced SWAR
//                   accumulators in registers, phred bins as byte counters in a
//                   bank-conflict-free shared-memory slab.  No atomics in the
//                   inner loop; private counters spill into a CTA histogram
//                   every 255 rows, the CTA histogram into the global u64
//                   tables once.  When all rows of a CTA have one length (the
//                   Illumina case) the end-anchored tables are a shifted copy
//                   of that histogram and cost nothing extra.
//  k_qc_end_anchored  fallback for CTAs with mixed lengths (shared atomics).
//  k_qc_horizontal  per-read quantities: GC bucket, the four-chain ordered
//                   error sum (bit-exact evaluation order of :2059-2112), the
//                   mean-phred bucket.  Four lanes per read, one per chain.
#include "modules.cuh"


constexpr int QC_TPB = 256;
constexpr int QC_BINS = 17;  // 5 base classes + 12 phred bins

// --------------------------------------------------------------------------
// vertical pass
// --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t load_word_any(const uint8_t *p) {
    uintptr_t a = (uintptr_t)p;
    const uint32_t *w = (const uint32_t *)(a & ~(uintptr_t)3);
    return __funnelshift_r(__ldg(w), __ldg(w + 1), (uint32_t)(a & 3) * 8);
}

// 0x01 in each byte of w that is one of ACGTacgt
__device__ __forceinline__ uint32_t acgt_bytes(uint32_t w) {
    // low three bits select the expected value of (c & 0xD8): A,C,G -> 0x40, T -> 0x50
    uint32_t sel = w & 0x07070707u;
    uint32_t t = sel | (sel >> 4);
    uint32_t nib = __byte_perm(t, 0, 0x4420);  // nibble j = selector of byte j
    uint32_t expect = __byte_perm(0x40FF40FFu, 0x40FFFF50u, nib);
    uint32_t x = (w & 0xD8D8D8D8u) ^ expect;
    return zero_bytes80(x) >> 7;
}

__global__ void __launch_bounds__(QC_TPB)
k_qc_vertical(BatchView bv, uint32_t rows_per_cta, uint32_t CG, uint32_t RG, uint64_t *g_base,
              uint64_t *g_phred, uint64_t *g_ea_base, uint64_t *g_ea_phred, uint32_t ea_len,
              uint8_t *mixed_flag) {
    extern __shared__ uint32_t smem[];
    uint32_t *priv = smem;                // [12][QC_TPB] words: byte j = phred-bin count of column j
    uint32_t *hist = smem + 12 * QC_TPB;  // [W][17]
    __shared__ uint32_t s_lmin, s_lmax;
    const uint32_t tid = threadIdx.x;
    const uint32_t W = CG * 4;
    const uint32_t win0 = blockIdx.y * W;
    const uint32_t r0 = blockIdx.x * rows_per_cta;
    const uint32_t r1 = min(r0 + rows_per_cta, bv.n);
    for (uint32_t i = tid; i < 12 * QC_TPB + W * QC_BINS; i += QC_TPB) smem[i] = 0;
    if (tid == 0) {
        s_lmin = 0xFFFFFFFFu;
        s_lmax = 0;
    }
    __syncthreads();
    {
        uint32_t lmin = 0xFFFFFFFFu, lmax = 0;
        for (uint32_t r = r0 + tid; r < r1; r += QC_TPB) {
            uint32_t L = bv.seq_len[r];
            lmin = min(lmin, L);
            lmax = max(lmax, L);
        }
        lmax = warp_max_u32(lmax);
        lmin = ~warp_max_u32(~lmin);
        if (lane_id() == 0) {
            atomicMin(&s_lmin, lmin);
            atomicMax(&s_lmax, lmax);
        }
    }
    __syncthreads();

    const uint32_t rg = tid / CG, cg = tid - rg * CG;
    const uint32_t col0 = win0 + cg * 4;
    uint8_t *priv8 = (uint8_t *)priv + tid * 4;
    uint32_t acc_v = 0, acc_h = 0, acc_g = 0, acc_hg = 0, acc_n = 0, rows = 0;

    auto spill = [&]() {
        // registers -> CTA histogram; per column: A = v-h-g+hg, C = h-hg, G = hg, T = g-hg
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t v = (acc_v >> (8 * j)) & 0xFF, h = (acc_h >> (8 * j)) & 0xFF;
            uint32_t g = (acc_g >> (8 * j)) & 0xFF, hg = (acc_hg >> (8 * j)) & 0xFF;
            uint32_t nn = (acc_n >> (8 * j)) & 0xFF;
            uint32_t *hrow = hist + (cg * 4 + j) * QC_BINS;
            if (v | nn) {
                uint32_t a = v - h - g + hg, c = h - hg, t = g - hg;
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
            uint32_t wv = priv[b * QC_TPB + tid];
            if (wv) {
                priv[b * QC_TPB + tid] = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    uint32_t c = (wv >> (8 * j)) & 0xFF;
                    if (c) atomicAdd(hist + (cg * 4 + j) * QC_BINS + 5 + b, c);
                }
            }
        }
        rows = 0;
    };

    if (rg < RG) {
        for (uint32_t r = r0 + rg; r < r1; r += RG) {
            uint32_t L = bv.seq_len[r];
            if (L > col0) {
                uint32_t nvalid = min(4u, L - col0);
                uint32_t w = load_word_any(bv.text + bv.seq_off[r] + col0);
                uint32_t q = load_word_any(bv.text + bv.qual_off[r] + col0);
                uint32_t pm = 0x01010101u >> (8 * (4 - nvalid));
                uint32_t vb = acgt_bytes(w) & pm;
                uint32_t hb = (w >> 1) & vb, gb = (w >> 2) & vb;
                acc_v += vb;
                acc_h += hb;
                acc_g += gb;
                acc_hg += hb & gb;
                acc_n += pm & ~vb;
                // phred bins min(q,47)>>2 as byte counters: address = bin*1024 + tid*4 + j
                uint32_t t = q - 0x21212121u;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if ((uint32_t)j < nvalid) {
                        uint32_t qv = (t >> (8 * j)) & 0xFF;
                        uint32_t bin = min(qv, 47u) >> 2;
                        priv8[(bin << 10) + j] += 1;
                    }
                }
            }
            if (++rows == 255) spill();
        }
        spill();
    }
    __syncthreads();

    // CTA histogram -> global tables
    for (uint32_t i = tid; i < W * QC_BINS; i += QC_TPB) {
        uint32_t c = hist[i];
        if (!c) continue;
        uint32_t pos = win0 + i / QC_BINS, k = i % QC_BINS;
        if (k < 5) atomic_add_u64(g_base + (uint64_t)pos * 5 + k, c);
        else atomic_add_u64(g_phred + (uint64_t)pos * 12 + (k - 5), c);
    }
    if (r0 >= r1) return;
    if (s_lmin == s_lmax) {
        // every row has length L0: the end-anchored rows are a shifted window of hist
        uint32_t L0 = s_lmin, ea_n = min(L0, ea_len);
        uint32_t lo = max(L0 - ea_n, win0), hi = min(L0, win0 + W);
        if (lo < hi) {
            uint32_t span = (hi - lo) * QC_BINS;
            for (uint32_t i = tid; i < span; i += QC_TPB) {
                uint32_t pos = lo + i / QC_BINS, k = i % QC_BINS;
                uint32_t c = hist[(pos - win0) * QC_BINS + k];
                if (!c) continue;
                uint64_t row = (uint64_t)ea_len - (L0 - pos);
                if (k < 5) atomic_add_u64(g_ea_base + row * 5 + k, c);
                else atomic_add_u64(g_ea_phred + row * 12 + (k - 5), c);
            }
        }
    }
    else if (blockIdx.y == 0 && tid == 0) mixed_flag[blockIdx.x] = 1;
}

// end-anchored counts for the row chunks whose lengths are mixed (:2034-2043, :2115-2124)
template <bool SMEM_HIST>
__global__ void __launch_bounds__(QC_TPB)
k_qc_end_anchored(BatchView bv, uint32_t rows_per_cta, const uint8_t *mixed_flag, uint64_t *g_ea_base,
                  uint64_t *g_ea_phred, uint32_t ea_len) {
    extern __shared__ uint32_t hist[];  // [ea_len][17] when SMEM_HIST
    if (!mixed_flag[blockIdx.x]) return;
    const uint32_t r0 = blockIdx.x * rows_per_cta, r1 = min(r0 + rows_per_cta, bv.n);
    if (SMEM_HIST) {
        for (uint32_t i = threadIdx.x; i < ea_len * QC_BINS; i += QC_TPB) hist[i] = 0;
        __syncthreads();
    }
    const uint32_t warp = threadIdx.x >> 5, nwarps = QC_TPB / 32;
    for (uint32_t r = r0 + warp; r < r1; r += nwarps) {
        uint32_t L = bv.seq_len[r], ea_n = min(L, ea_len);
        const uint8_t *s = bv.text + bv.seq_off[r] + (L - ea_n);
        const uint8_t *q = bv.text + bv.qual_off[r] + (L - ea_n);
        uint32_t row0 = ea_len - ea_n;
        for (uint32_t k = lane_id(); k < ea_n; k += 32) {
            uint32_t b = nuc5(s[k]);
            uint32_t p = min((uint32_t)(uint8_t)(q[k] - 33), 47u) >> 2;
            if (SMEM_HIST) {
                atomicAdd(hist + (row0 + k) * QC_BINS + b, 1u);
                atomicAdd(hist + (row0 + k) * QC_BINS + 5 + p, 1u);
            }
            else {
                atomic_add_u64(g_ea_base + (uint64_t)(row0 + k) * 5 + b, 1);
                atomic_add_u64(g_ea_phred + (uint64_t)(row0 + k) * 12 + p, 1);
            }
        }
    }
    if (SMEM_HIST) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < ea_len * QC_BINS; i += QC_TPB) {
            uint32_t c = hist[i];
            if (!c) continue;
            uint32_t row = i / QC_BINS, k = i % QC_BINS;
            if (k < 5) atomic_add_u64(g_ea_base + (uint64_t)row * 5 + k, c);
            else atomic_add_u64(g_ea_phred + (uint64_t)row * 12 + (k - 5), c);
        }
    }
}

// --------------------------------------------------------------------------
// horizontal pass: four lanes per read, lane k owns error chain k
// --------------------------------------------------------------------------
__global__ void __launch_bounds__(QC_TPB)
k_qc_horizontal(BatchView bv, const double *__restrict__ err_tab, const double *__restrict__ edges,
                uint64_t *g_gc, uint64_t *g_mean_phred, unsigned long long *err_key,
                uint64_t record_base) {
    __shared__ double s_err[94], s_edge[94];
    __shared__ uint32_t s_gc[101], s_mp[94];
    for (uint32_t i = threadIdx.x; i < 94; i += QC_TPB) {
        s_err[i] = err_tab[i];
        s_edge[i] = edges[i];
        s_mp[i] = 0;
    }
    for (uint32_t i = threadIdx.x; i < 101; i += QC_TPB) s_gc[i] = 0;
    __syncthreads();
    const uint32_t sub = threadIdx.x & 3;
    const uint32_t groups = gridDim.x * (QC_TPB / 4);
    // all four lanes of a group run the same trip counts up to the shuffles
    for (uint32_t r = blockIdx.x * (QC_TPB / 4) + (threadIdx.x >> 2); r < bv.n; r += groups) {
        const uint32_t L = bv.seq_len[r];
        const uint8_t *s = bv.text + bv.seq_off[r];
        const uint8_t *q = bv.text + bv.qual_off[r];
        const uint32_t gmask = 0xFu << (lane_id() & ~3u);
        uint32_t gc = 0, at = 0;
        for (uint32_t i = sub; i < L; i += 4) {
            uint32_t c = s[i] | 0x20u;
            gc += (c == 'c') | (c == 'g');
            at += (c == 'a') | (c == 't');
        }
        const uint32_t nit = L >= 5 ? (L - 1) / 4 : 0;  // groups of four while > 4 remain (:2068)
        double acc = 0.0;
        uint32_t bad = 0xFFFFFFFFu;
        for (uint32_t i = 0; i < nit; i++) {
            uint32_t v = (uint8_t)(q[4 * i + sub] - 33);
            if (v > 93) {
                bad = 4 * i + sub;
                break;
            }
            acc += s_err[v];
        }
        __syncwarp(gmask);
        gc += __shfl_xor_sync(gmask, gc, 1);
        gc += __shfl_xor_sync(gmask, gc, 2);
        at += __shfl_xor_sync(gmask, at, 1);
        at += __shfl_xor_sync(gmask, at, 2);
        bad = min(bad, __shfl_xor_sync(gmask, bad, 1));
        bad = min(bad, __shfl_xor_sync(gmask, bad, 2));
        const uint32_t lead = lane_id() & ~3u;
        double a1 = __shfl_sync(gmask, acc, lead + 1);
        double a2 = __shfl_sync(gmask, acc, lead + 2);
        double a3 = __shfl_sync(gmask, acc, lead + 3);
        if (sub == 0) {
            double sum = ((acc + a1) + a2) + a3;  // :2098-2099
            if (bad == 0xFFFFFFFFu) {
                for (uint32_t i = 4 * nit; i < L; i++) {  // tail, in order (:2100-2112)
                    uint32_t v = (uint8_t)(q[i] - 33);
                    if (v > 93) {
                        bad = i;
                        break;
                    }
                    sum += s_err[v];
                }
            }
            if (at + gc) {  // :2045-2058
                double pct = (double)gc * 100.0 / (double)(at + gc);
                atomicAdd(&s_gc[(uint32_t)round(pct)], 1u);
            }
            if (bad != 0xFFFFFFFFu) {
                atomicMin(err_key, (unsigned long long)((record_base + r) << 8 | q[bad]));
            }
            else {
                bv.err_sum[r] = sum;
                if (L) {  // floor(-10*log10(sum/L)) through host-derived bucket edges (:2127-2137)
                    double avg = sum / (double)L;
                    uint32_t lo = 0, hi = 93;
                    while (lo < hi) {
                        uint32_t mid = (lo + hi + 1) >> 1;
                        if (avg <= s_edge[mid]) lo = mid;
                        else hi = mid - 1;
                    }
                    atomicAdd(&s_mp[lo], 1u);
                }
            }
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < 101; i += QC_TPB)
        if (s_gc[i]) atomic_add_u64(g_gc + i, s_gc[i]);
    for (uint32_t i = threadIdx.x; i < 94; i += QC_TPB)
        if (s_mp[i]) atomic_add_u64(g_mean_phred + i, s_mp[i]);
}

// --------------------------------------------------------------------------
// horizontal pass for long reads (nanopore: mean 20 kb, up to 1 Mb): a warp per read.
// The four error chains of a read (:2059-2099) are serial by construction -- a 1 Mb read is four
// chains of 250 000 dependent additions -- so the point is to keep every chain step down to ONE
// dependent DADD: all 32 lanes fetch the next 128 quality bytes with one coalesced word load each,
// look the error rates up and park them in shared memory; lanes 0..3 then add their chain's 32
// values in order (loads issued ahead, additions back to back).  GC content is counted on the same
// trip with a coalesced word load of the sequence per lane.
// --------------------------------------------------------------------------
constexpr int QL_WARPS = QC_TPB / 32;
__global__ void __launch_bounds__(QC_TPB)
k_qc_horizontal_long(BatchView bv, const double *__restrict__ err_tab, const double *__restrict__ edges,
                     uint64_t *g_gc, uint64_t *g_mean_phred, unsigned long long *err_key, uint64_t record_base) {
    __shared__ double s_err[128], s_edge[94];
    __shared__ uint32_t s_gc[101], s_mp[94];
    __shared__ double s_val[QL_WARPS][2][128];  // error rates of 128 consecutive positions, double buffered
    for (uint32_t i = threadIdx.x; i < 128; i += QC_TPB) s_err[i] = (i >= 33 && i < 127) ? err_tab[i - 33] : 0.0;
    for (uint32_t i = threadIdx.x; i < 94; i += QC_TPB) {
        s_edge[i] = edges[i];
        s_mp[i] = 0;
    }
    for (uint32_t i = threadIdx.x; i < 101; i += QC_TPB) s_gc[i] = 0;
    __syncthreads();
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t warps = gridDim.x * QL_WARPS;
    for (uint32_t r = blockIdx.x * QL_WARPS + warp; r < bv.n; r += warps) {
        const uint32_t L = bv.seq_len[r];
        const uint8_t *s = bv.text + bv.seq_off[r];
        const uint8_t *q = bv.text + bv.qual_off[r];
        // ---- GC content: SWAR over coalesced words -----------------------------------------------------
        uint32_t gc = 0, at = 0;
        for (uint32_t i = lane * 4; i < L; i += 128) {
            uint32_t w = load_u32_unaligned(s + i);
            const uint32_t nvalid = min(4u, L - i);
            if (nvalid < 4) w &= 0xFFFFFFFFu >> (8 * (4 - nvalid));
            const uint32_t u = w | 0x20202020u;
            gc += __popc(zero_bytes80(u ^ 0x63636363u) | zero_bytes80(u ^ 0x67676767u));
            at += __popc(zero_bytes80(u ^ 0x61616161u) | zero_bytes80(u ^ 0x74747474u));
        }
        gc = warp_sum_u32(gc);
        at = warp_sum_u32(at);
        // ---- the four chains over positions [0, 4 * nit) ----------------------------------------------------
        const uint32_t nit = L >= 5 ? (L - 1) / 4 : 0;  // groups of four while > 4 remain (:2068)
        const uint32_t n_main = 4 * nit;
        double acc = 0.0;
        uint32_t bad = 0xFFFFFFFFu;  // first position holding a byte outside '!'..'~'
        auto stage_block = [&](uint32_t base, int buf) {
            // lane l: positions base + 4l .. base + 4l + 3 -> s_val[..][4l .. 4l+3] (0.0 past n_main)
            const uint32_t i = base + lane * 4;
            double e[4] = {0.0, 0.0, 0.0, 0.0};
            if (i < n_main) {
                const uint32_t w = load_u32_unaligned(q + i);  // n_main is a multiple of 4: whole words
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t c = (w >> (8 * k)) & 0xFF;
                    if (c - 33u > 93u) bad = min(bad, i + k);
                    e[k] = s_err[c & 0x7F];
                }
            }
            double2 *dst = (double2 *)&s_val[warp][buf][lane * 4];
            dst[0] = make_double2(e[0], e[1]);
            dst[1] = make_double2(e[2], e[3]);
        };
        if (n_main) stage_block(0, 0);
        __syncwarp();
        int buf = 0;
        for (uint32_t base = 0; base < n_main; base += 128) {
            if (base + 128 < n_main) stage_block(base + 128, buf ^ 1);  // the next block travels meanwhile
            if (lane < 4) {
                const double *v = &s_val[warp][buf][lane];
                const uint32_t steps = min(32u, (n_main - base) / 4);
                if (steps == 32) {
#pragma unroll
                    for (int g = 0; g < 32; g++) acc += v[4 * g];
                }
                else
                    for (uint32_t g = 0; g < steps; g++) acc += v[4 * g];
            }
            __syncwarp();
            buf ^= 1;
        }
        bad = ~warp_max_u32(~bad);
        const double a1 = __shfl_sync(0xffffffffu, acc, 1), a2 = __shfl_sync(0xffffffffu, acc, 2),
                     a3 = __shfl_sync(0xffffffffu, acc, 3);
        if (lane == 0) {
            double sum = ((acc + a1) + a2) + a3;  // :2098-2099
            if (bad == 0xFFFFFFFFu) {
                for (uint32_t i = n_main; i < L; i++) {  // tail, in order (:2100-2112)
                    const uint32_t v = (uint8_t)(q[i] - 33);
                    if (v > 93) {
                        bad = i;
                        break;
                    }
                    sum += s_err[v + 33];
                }
            }
            if (at + gc) {  // :2045-2058
                const double pct = (double)gc * 100.0 / (double)(at + gc);
                atomicAdd(&s_gc[(uint32_t)round(pct)], 1u);
            }
            if (bad != 0xFFFFFFFFu) atomicMin(err_key, (unsigned long long)((record_base + r) << 8 | q[bad]));
            else {
                bv.err_sum[r] = sum;
                if (L) {  // floor(-10*log10(sum/L)) through host-derived bucket edges (:2127-2137)
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
        __syncwarp();
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < 101; i += QC_TPB)
        if (s_gc[i]) atomic_add_u64(g_gc + i, s_gc[i]);
    for (uint32_t i = threadIdx.x; i < 94; i += QC_TPB)
        if (s_mp[i]) atomic_add_u64(g_mean_phred + i, s_mp[i]);
}

// --------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------
extern "C" int sq_qc_create(sq_ctx *ctx, uint64_t end_anchor_length, sq_qc **out) {
    *out = nullptr;
    if (end_anchor_length > 0xFFFFFFFFULL) {
        sq_set_error("end_anchor_length must be between 0 and 4294967295");
        return SQ_E_ARG;
    }
    CUDA_TRY(cudaSetDevice(ctx->device));
    sq_qc *m = new sq_qc();
    m->ctx = ctx;
    m->ea_len = end_anchor_length;
    int rc = sq_dalloc(ctx, (void **)&m->ea_base, end_anchor_length * 5 * 8, true);
    if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&m->ea_phred, end_anchor_length * 12 * 8, true);
    if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&m->gc, 101 * 8, true);
    if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&m->mean_phred, 94 * 8, true);
    if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&m->err_key, 8, false);
    if (rc == SQ_OK)
        rc = cudaMemsetAsync(m->err_key, 0xFF, 8, ctx->stream) == cudaSuccess ? SQ_OK : SQ_E_CUDA;
    if (rc != SQ_OK) {
        sq_qc_destroy(m);
        return rc;
    }
    *out = m;
    return SQ_OK;
}

extern "C" void sq_qc_destroy(sq_qc *m) {
    if (!m) return;
    cudaSetDevice(m->ctx->device);
    sq_dfree(m->ctx, m->base);
    sq_dfree(m->ctx, m->phred);
    sq_dfree(m->ctx, m->ea_base);
    sq_dfree(m->ctx, m->ea_phred);
    sq_dfree(m->ctx, m->gc);
    sq_dfree(m->ctx, m->mean_phred);
    sq_dfree(m->ctx, m->err_key);
    delete m;
}

int qc_grow(sq_qc *m, uint64_t len) {
    if (len <= m->cap_len) return SQ_OK;
    uint64_t cap = m->cap_len * 2 > len ? m->cap_len * 2 : len;
    if (cap < 256) cap = 256;
    uint64_t *nb = nullptr, *np = nullptr;
    SQ_TRY(sq_dalloc(m->ctx, (void **)&nb, cap * 5 * 8, true));
    SQ_TRY(sq_dalloc(m->ctx, (void **)&np, cap * 12 * 8, true));
    if (m->cap_len) {
        CUDA_TRY(cudaMemcpyAsync(nb, m->base, m->cap_len * 5 * 8, cudaMemcpyDeviceToDevice, m->ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(np, m->phred, m->cap_len * 12 * 8, cudaMemcpyDeviceToDevice, m->ctx->stream));
    }
    sq_dfree(m->ctx, m->base);
    sq_dfree(m->ctx, m->phred);
    m->base = nb;
    m->phred = np;
    m->cap_len = cap;
    return SQ_OK;
}

int qc_add_vertical(sq_qc *m, sq_batch *b) {
    sq_ctx *ctx = m->ctx;
    SQ_TRY(qc_grow(m, b->max_len));
    BatchView bv = b->view();
    uint32_t n = (uint32_t)b->n;
    if (b->max_len > 0) {
        uint32_t CG = (b->max_len + 3) / 4;
        if (CG > QC_TPB) CG = QC_TPB;
        uint32_t RG = QC_TPB / CG;
        uint32_t W = CG * 4;
        uint32_t windows = (b->max_len + W - 1) / W;
        // enough CTAs for ~4 per SM, but at least 512 rows each so the final
        // histogram flush (W*17 atomics) stays amortised
        uint32_t rows_per_cta = (n + ctx->num_sms * 4 - 1) / (ctx->num_sms * 4);
        if (rows_per_cta < 512) rows_per_cta = 512;
        uint32_t chunks = (n + rows_per_cta - 1) / rows_per_cta;
        uint8_t *mixed = nullptr;
        SQ_TRY(sq_dalloc(ctx, (void **)&mixed, chunks, true));
        size_t smem = (size_t)(12 * QC_TPB + W * QC_BINS) * 4;
        if (!(ctx->func_attr_done & 1u)) {
            CUDA_TRY(cudaFuncSetAttribute(k_qc_vertical, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            CUDA_TRY(cudaFuncSetAttribute(k_qc_end_anchored<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            ctx->func_attr_done |= 1u;
        }
        dim3 grid(chunks, windows);
        SQ_LAUNCH(ctx, k_qc_vertical, grid, QC_TPB, smem, bv, rows_per_cta, CG, RG, m->base, m->phred,
                  m->ea_base, m->ea_phred, (uint32_t)m->ea_len, mixed);
        if (m->ea_len) {
            size_t ea_smem = (size_t)m->ea_len * QC_BINS * 4;
            if (ea_smem <= 96 * 1024)
                SQ_LAUNCH(ctx, k_qc_end_anchored<true>, chunks, QC_TPB, ea_smem, bv, rows_per_cta, mixed,
                          m->ea_base, m->ea_phred, (uint32_t)m->ea_len);
            else
                SQ_LAUNCH(ctx, k_qc_end_anchored<false>, chunks, QC_TPB, 0, bv, rows_per_cta, mixed,
                          m->ea_base, m->ea_phred, (uint32_t)m->ea_len);
        }
        sq_dfree(ctx, mixed);
    }
    return SQ_OK;
}

extern "C" int sq_qc_add(sq_qc *m, sq_batch *b) {
    sq_ctx *ctx = m->ctx;
    if (b->ctx != ctx) {
        sq_set_error("record array belongs to another context");
        return SQ_E_ARG;
    }
    if (b->n == 0) return SQ_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    SQ_TRY(qc_add_vertical(m, b));
    const uint32_t n = (uint32_t)b->n;
    if (b->max_len > 2048) {  // long reads: a warp per read
        SQ_LAUNCH(ctx, k_qc_horizontal_long, sq_grid_for(ctx, (uint64_t)n * 32, QC_TPB, 16), QC_TPB, 0, b->view(),
                  ctx->d_err_table, ctx->d_phred_thresholds, m->gc, m->mean_phred, m->err_key, m->n_reads);
    }
    else {
        int grid_h = sq_grid_for(ctx, (uint64_t)n * 4, QC_TPB, 8);
        SQ_LAUNCH(ctx, k_qc_horizontal, grid_h, QC_TPB, 0, b->view(), ctx->d_err_table, ctx->d_phred_thresholds, m->gc,
                  m->mean_phred, m->err_key, m->n_reads);
    }
    m->n_reads += n;
    if (b->max_len > m->max_len) m->max_len = b->max_len;
    b->err_sum_valid = true;
    return SQ_OK;
}

extern "C" int sq_qc_sync(sq_qc *m, sq_qc_info *info) {
    sq_ctx *ctx = m->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    unsigned long long *h = (unsigned long long *)((char *)ctx->h_scratch + 1024);
    CUDA_TRY(cudaMemcpyAsync(h, m->err_key, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    memset(info, 0, sizeof(*info));
    info->number_of_reads = m->n_reads;
    info->max_length = m->max_len;
    info->end_anchor_length = m->ea_len;
    if (*h != ~0ULL) {
        info->bad_phred = 1;
        info->bad_phred_char = (uint8_t)(*h & 0xFF);
        info->bad_phred_record = *h >> 8;
    }
    return SQ_OK;
}

extern "C" int sq_qc_read(sq_qc *m, uint64_t *base_counts, uint64_t *phred_counts, uint64_t *ea_base,
                          uint64_t *ea_phred, uint64_t *gc_content, uint64_t *phred_scores) {
    sq_ctx *ctx = m->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    if (base_counts && m->max_len)
        CUDA_TRY(cudaMemcpyAsync(base_counts, m->base, m->max_len * 5 * 8, cudaMemcpyDeviceToHost, s));
    if (phred_counts && m->max_len)
        CUDA_TRY(cudaMemcpyAsync(phred_counts, m->phred, m->max_len * 12 * 8, cudaMemcpyDeviceToHost, s));
    if (ea_base && m->ea_len)
        CUDA_TRY(cudaMemcpyAsync(ea_base, m->ea_base, m->ea_len * 5 * 8, cudaMemcpyDeviceToHost, s));
    if (ea_phred && m->ea_len)
        CUDA_TRY(cudaMemcpyAsync(ea_phred, m->ea_phred, m->ea_len * 12 * 8, cudaMemcpyDeviceToHost, s));
    if (gc_content) CUDA_TRY(cudaMemcpyAsync(gc_content, m->gc, 101 * 8, cudaMemcpyDeviceToHost, s));
    if (phred_scores) CUDA_TRY(cudaMemcpyAsync(phred_scores, m->mean_phred, 94 * 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return SQ_OK;
}
