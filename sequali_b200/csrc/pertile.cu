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
#include <math.h>

#include <algorithm>

#include "modules.cuh"




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
        // reads come in runs of one tile: only the first read of a run has to look at the map
        if (r > 0 && base + r - 1 < fail && (uint64_t)tile[r - 1] == t) continue;
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
    uint32_t lmax = 0, kept = 0, changes = 0;
    unsigned long long tlo = ~0ULL, thi = 0;
    uint64_t tprev = ~0ULL;
    const uint32_t n_round = (bv.n + 31) & ~31u;  // whole warps stay in the loop for the match below
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_round; r += gridDim.x * blockDim.x) {
        const bool in = r < bv.n;
        const bool keep = in && base + r < fail;
        if (in) idx[r] = r;
        const uint64_t t = keep ? (uint64_t)tile[r] : ~0ULL;
        // one probe per distinct tile in the warp (reads in tile runs: usually one tile, settled by a vote)
        const uint64_t t0 = __shfl_sync(0xffffffffu, t, 0);
        const uint32_t peers = __all_sync(0xffffffffu, t == t0) ? 0xffffffffu : __match_any_sync(0xffffffffu, t);
        const uint32_t leader = (uint32_t)__ffs(peers) - 1;
        uint32_t sl = PT_NONE;
        if (keep && lane_id() == leader) {
            uint32_t i = pt_hash(t) & (PT_MAP_CAP - 1);
            while (map_keys[i] != t) i = (i + 1) & (PT_MAP_CAP - 1);
            sl = map_vals[i];
        }
        sl = __shfl_sync(0xffffffffu, sl, leader);
        if (in) slot[r] = keep ? sl : PT_NONE;
        if (keep) {
            changes += r == 0 || tile[r - 1] != tile[r];
            lmax = max(lmax, bv.seq_len[r]);
            kept++;
            if (t != tprev) {  // reads come in tile runs: rarely taken
                tlo = min(tlo, (unsigned long long)t);
                thi = max(thi, (unsigned long long)t);
                tprev = t;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        tlo = min(tlo, __shfl_xor_sync(0xffffffffu, tlo, o));
        thi = max(thi, __shfl_xor_sync(0xffffffffu, thi, o));
    }
    if (lane_id() == 0 && tlo <= thi) {
        atomicMin(&st->tile_lo, tlo);
        atomicMax(&st->tile_hi, thi);
    }
    lmax = warp_max_u32(lmax);
    kept = warp_sum_u32(kept);
    changes = warp_sum_u32(changes);
    if (lane_id() == 0) {
        if (lmax) atomicMax(&st->max_len, lmax);
        if (kept) atomicAdd(&st->n_kept, (unsigned long long)kept);
        if (changes) atomicAdd(&st->n_changes, changes);
    }
}

__global__ void __launch_bounds__(PT_TPB)
k_pt_lengths(BatchView bv, const uint32_t *__restrict__ slot, uint64_t *lengths, uint64_t len_cap) {
    // neighbouring reads share tile and length: one atomic per distinct (tile, length) in a warp
    const uint32_t n_round = (bv.n + 31) & ~31u;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_round; r += gridDim.x * blockDim.x) {
        uint32_t s = PT_NONE, L = 0;
        if (r < bv.n) {
            s = slot[r];
            L = bv.seq_len[r];
        }
        const bool take = s != PT_NONE && L;
        const uint64_t key = take ? ((uint64_t)s << 32 | L) : ~0ULL;
        const uint32_t peers = __match_any_sync(0xffffffffu, key);
        if (take && lane_id() == (uint32_t)__ffs(peers) - 1)
            atomic_add_u64(lengths + (uint64_t)s * len_cap + (L - 1), __popc(peers));
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

// ---------------------------------------------------------------------------
// Ordered double sums without a serial pass over the reads.
//
// The reference extends total_errors[tile][pos] by one rounded double addition
// per read (:3199-3219).  While the running sum s stays inside one binade
// [2^k, 2^(k+1)) every such addition moves s by a whole number of ulps:
// fl(s + e) = s + round_to_nearest(e / ulp_k) * ulp_k, and on the IEEE bit
// pattern that is the integer addition bits(s) += r_k(e).  Integer additions
// are associative, so a run of reads that keeps s inside its binade (and meets
// no round-half tie, whose direction would depend on the parity of s) can be
// summed in any order and applied at once; only the few additions that cross
// a power of two or tie are executed as real, ordered double additions.
//
//   k_pt_approx   per (segment of consecutive reads of one tile, position):
//                 approximate (float) sum of the error rates
//   k_pt_guess    per chain: prefix of those sums from the chain's exact state
//                 -> the binade each segment is expected to start in
//   k_pt_incr     per (segment, position): exact integer sum of r_k(e) for the
//                 guessed binade k (poisoned by a tie or an increment that
//                 cannot stay in the binade)
//   k_pt_chain    one warp per chain walks its segments 32 at a time, verifies
//                 the guess against the exact state and applies the integer
//                 sums; a segment that fails the check is replayed read by read
//                 (32 reads per step, same integer rule, real additions at the
//                 crossing points).  Guesses are hints: every accepted step is
//                 verified exactly, so the result is bit-identical to the serial
//                 chain whatever the guess was.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(PT_TPB)
k_pt_seg_counts(const uint32_t *__restrict__ seg_lo, const uint32_t *__restrict__ seg_hi, uint32_t n_slots,
                uint32_t seg_rows, uint32_t *__restrict__ nseg) {
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n_slots; s += gridDim.x * blockDim.x)
        nseg[s] = seg_hi[s] > seg_lo[s] ? (seg_hi[s] - seg_lo[s] + seg_rows - 1) / seg_rows : 0;
}
__global__ void __launch_bounds__(PT_TPB)
k_pt_seg_fill(const uint32_t *__restrict__ seg_lo, const uint32_t *__restrict__ seg_hi,
              const uint32_t *__restrict__ seg_first, uint32_t n_slots, uint32_t seg_rows, PtSeg *__restrict__ segs) {
    // one warp per tile: lanes write that tile's segment descriptors
    const uint32_t warps = gridDim.x * (PT_TPB / 32);
    for (uint32_t s = blockIdx.x * (PT_TPB / 32) + (threadIdx.x >> 5); s < n_slots; s += warps) {
        const uint32_t lo = seg_lo[s], hi = seg_hi[s], first = seg_first[s];
        const uint32_t cnt = (hi - lo + seg_rows - 1) / seg_rows;
        for (uint32_t j = lane_id(); j < cnt; j += 32) {
            PtSeg g;
            g.lo = lo + j * seg_rows;
            g.hi = min(hi, g.lo + seg_rows);
            g.slot = s;
            g.data = first + j;
            segs[first + j] = g;
        }
    }
}

// Vertical walk shared by k_pt_approx / k_pt_incr: a thread owns four positions
// (col0..col0+3) of one segment; the CG threads of a row group read consecutive
// words of the same quality line.
template <bool EXACT>
__global__ void __launch_bounds__(PT_TPB)
k_pt_segment_sums(BatchView bv, const uint32_t *__restrict__ order, const PtSeg *__restrict__ segs,
                  const uint32_t *__restrict__ n_segs_p, uint32_t CG, uint32_t RG, uint32_t width4, uint32_t stride,
                  const double *__restrict__ err_tab, const uint64_t *__restrict__ lut,
                  const uint16_t *__restrict__ kguess, float *__restrict__ approx, uint64_t *__restrict__ incr,
                  uint64_t base, PtState *st) {
    __shared__ float s_errf[96];
    extern __shared__ uint64_t s_lut[];  // [PT_LUT_NK][94] when EXACT
    if (EXACT) {
        for (uint32_t i = threadIdx.x; i < PT_LUT_NK * 94; i += PT_TPB) s_lut[i] = lut[i];
    }
    else {
        for (uint32_t i = threadIdx.x; i < 94; i += PT_TPB) s_errf[i] = (float)err_tab[i];
    }
    __syncthreads();
    const uint32_t n_segs = *n_segs_p;
    const uint32_t rg = threadIdx.x / CG, cg = threadIdx.x - rg * CG;
    if (rg >= RG) return;
    const uint32_t col0 = blockIdx.y * CG * 4 + cg * 4;
    if (col0 >= width4) return;
    for (uint32_t g = blockIdx.x * RG + rg; g < n_segs; g += gridDim.x * RG) {
        const PtSeg sg = segs[g];
        float fa[4] = {0.f, 0.f, 0.f, 0.f};
        uint64_t ia[4] = {0, 0, 0, 0};
        const uint64_t *lrow[4];
        if (EXACT) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int k = (int)kguess[(uint64_t)(col0 + j) * stride + g] - PT_LUT_KMIN;
                if (k < 0 || k >= PT_LUT_NK) {
                    lrow[j] = s_lut;
                    ia[j] = PT_HARD;
                }
                else lrow[j] = s_lut + k * 94;
            }
        }
        for (uint32_t i = sg.lo; i < sg.hi; i++) {
            const uint32_t r = order[i];
            const uint32_t L = bv.seq_len[r];
            if (L <= col0) continue;
            const uint32_t nvalid = min(4u, L - col0);
            const uint32_t w = load_u32_unaligned(bv.text + bv.qual_off[r] + col0) - 0x21212121u;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if ((uint32_t)j < nvalid) {
                    const uint32_t q = (w >> (8 * j)) & 0xFF;
                    if (q > 93) {
                        if (!EXACT)
                            atomicMin(&st->err_key, (unsigned long long)((base + r) << 8 | ((q + 33) & 0xFF)));
                    }
                    else if (EXACT) ia[j] += lrow[j][q];
                    else fa[j] += s_errf[q];
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (EXACT) incr[(uint64_t)(col0 + j) * stride + g] = ia[j] < PT_HARD ? ia[j] : PT_HARD;
            else approx[(uint64_t)(col0 + j) * stride + g] = fa[j];
        }
    }
}

__device__ __forceinline__ uint64_t warp_incl_scan_u64(uint64_t v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint64_t y = __shfl_up_sync(0xffffffffu, v, o);
        if (lane_id() >= (uint32_t)o) v += y;
    }
    return v;
}

// float sum of the error rates of rows [lo, hi) at `pos` (segments without precomputed sums)
__device__ double pt_rows_sum(const BatchView &bv, const uint32_t *__restrict__ order, uint32_t lo, uint32_t hi,
                              uint32_t pos, const double *__restrict__ err_tab) {
    double s = 0.0;
    for (uint32_t i = lo; i < hi; i++) {
        const uint32_t r = order ? order[i] : i;
        if (pos >= bv.seq_len[r]) continue;
        const uint32_t q = (uint8_t)(bv.text[bv.qual_off[r] + pos] - 33);
        if (q <= 93) s += err_tab[q];
    }
    return s;
}

// one warp per chain (tile slot, position): expected binade at the start of every segment.
// A tile's segments are those of segs[first .. first+cnt) that carry its slot (the general
// path groups them; the tile-run path leaves foreign segments of a mixed tile in between).
__global__ void __launch_bounds__(PT_TPB)
k_pt_guess(BatchView bv, const uint32_t *__restrict__ order, const PtSeg *__restrict__ segs,
           const uint32_t *__restrict__ seg_first, const uint32_t *__restrict__ nseg, uint32_t n_slots,
           uint32_t width, uint32_t stride, const float *__restrict__ approx, const double *__restrict__ errors,
           uint64_t len_cap, const double *__restrict__ err_tab, uint16_t *__restrict__ kguess, int dual) {
    const uint64_t chains = (uint64_t)n_slots * width;
    const uint64_t warps = (uint64_t)gridDim.x * (PT_TPB / 32);
    for (uint64_t c = (uint64_t)blockIdx.x * (PT_TPB / 32) + (threadIdx.x >> 5); c < chains; c += warps) {
        const uint32_t s = (uint32_t)(c / width), pos = (uint32_t)(c % width);
        const uint32_t cnt = nseg[s];
        if (cnt == 0) continue;
        const uint32_t first = seg_first[s];
        double run = errors[(uint64_t)s * len_cap + pos];
        for (uint32_t j0 = 0; j0 < cnt; j0 += 32) {
            const uint32_t j = j0 + lane_id();
            double v = 0.0;
            uint32_t data = PT_NONE, lo = 0, hi = 0;
            bool replayed = false;
            if (j < cnt) {
                const PtSeg sg = segs[first + j];
                if (sg.slot == s) {
                    data = sg.data;
                    lo = sg.lo;
                    hi = sg.hi;
                    if (data != PT_NONE) v = (double)approx[(uint64_t)pos * stride + data];
                    else replayed = true;
                }
            }
            // segments without precomputed sums (pieces of a mixed tile): the warp sums their rows together
            uint32_t todo = __ballot_sync(0xffffffffu, replayed);
            while (todo) {
                const uint32_t l = (uint32_t)__ffs(todo) - 1;
                todo &= todo - 1;
                const uint32_t rlo = __shfl_sync(0xffffffffu, lo, l), rhi = __shfl_sync(0xffffffffu, hi, l);
                double part = 0.0;
                for (uint32_t i = rlo + lane_id(); i < rhi; i += 32) part += pt_rows_sum(bv, order, i, i + 1, pos, err_tab);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                if (lane_id() == l) v = part;
            }
            double x = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                double y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane_id() >= (uint32_t)o) x += y;
            }
            const double start = run + (x - v);
            if (data != PT_NONE) {
                uint32_t kg;
                if (dual) {
                    // the sums are estimates from a sample of the reads: tabulate the two binades around
                    // the middle of the segment, (k-1, k) below sqrt(2) * 2^k and (k, k+1) above
                    const uint64_t mb = (uint64_t)__double_as_longlong(start + 0.5 * v);
                    kg = (uint32_t)(mb >> 52);
                    if ((mb & PT_MANT) < 0x6A09E667F3BCDULL && kg > 1) kg--;
                }
                else kg = (uint32_t)((uint64_t)__double_as_longlong(start) >> 52);
                kguess[(uint64_t)pos * stride + data] = (uint16_t)kg;
            }
            run += __shfl_sync(0xffffffffu, x, 31);
        }
    }
}

// Replay rows [lo, hi) of `order` (or of the array) on chain `pos`, bit-exact.  128 rows
// are gathered at a time (four per lane, all loads in flight together); within a
// block of 32 rows the in-binade integer rule is applied to as many rows as stay
// in the binade, the row that does not gets a real addition, and the block is
// resumed behind it.
template <int NB = 4>  // blocks of 32 rows gathered together (8: a whole segment of the run path in one go)
__device__ uint64_t pt_replay_rows(uint64_t sbits, const BatchView &bv, const uint32_t *__restrict__ order,
                                   uint32_t lo, uint32_t hi, uint32_t pos, const double *s_err) {
    for (uint32_t c0 = lo; c0 < hi; c0 += NB * 32) {
        uint32_t qv[NB];  // phred of this lane's row in block b; >= 94: the row adds nothing
        {
            uint32_t rr[NB], off[NB];
            bool ok[NB];
#pragma unroll
            for (int b = 0; b < NB; b++) {
                const uint32_t ii = c0 + b * 32 + lane_id();
                ok[b] = ii < hi;
                rr[b] = ok[b] ? (order ? order[ii] : ii) : 0;
            }
#pragma unroll
            for (int b = 0; b < NB; b++) {
                off[b] = 0;
                if (ok[b]) {
                    ok[b] = pos < bv.seq_len[rr[b]];
                    off[b] = bv.qual_off[rr[b]];
                }
            }
#pragma unroll
            for (int b = 0; b < NB; b++) qv[b] = ok[b] ? (uint32_t)(uint8_t)(bv.text[off[b] + pos] - 33) : 0xFFu;
        }
#pragma unroll
        for (int b = 0; b < NB; b++) {
            const uint32_t base = c0 + b * 32;
            if (base >= hi) break;
            const uint32_t nv = min(32u, hi - base);
            const uint64_t eb = qv[b] <= 93 ? (uint64_t)__double_as_longlong(s_err[qv[b]]) : 0ULL;
            uint32_t s0 = 0;  // rows of this block already applied
            while (s0 < nv) {
                const uint32_t k = (uint32_t)(sbits >> 52);
                const bool mine = lane_id() >= s0 && lane_id() < nv && eb != 0;
                uint64_t inc = mine ? pt_increment(k, eb) : 0;
                const bool hard = inc >= PT_HARD;
                if (hard) inc = 0;
                const uint64_t incl = warp_incl_scan_u64(inc);
                const bool stays = (uint32_t)((sbits + incl) >> 52) == k;
                const uint32_t bad = __ballot_sync(0xffffffffu, mine && (hard || !stays));
                const uint32_t f = bad ? (uint32_t)__ffs(bad) - 1 : 32u;
                if (f >= nv) {  // every remaining row of the block stays in the binade
                    sbits += __shfl_sync(0xffffffffu, incl, 31);
                    break;
                }
                if (f > 0) sbits += __shfl_sync(0xffffffffu, incl, f - 1);
                // a real, ordered addition: crosses a power of two, ties, or starts from 0.0
                const uint64_t e = __shfl_sync(0xffffffffu, eb, f);
                const double sum = __longlong_as_double((long long)sbits) + __longlong_as_double((long long)e);
                sbits = (uint64_t)__double_as_longlong(sum);
                s0 = f + 1;
            }
        }
    }
    return sbits;
}

__global__ void __launch_bounds__(PT_TPB)
k_pt_chain(BatchView bv, const uint32_t *__restrict__ order, const PtSeg *__restrict__ segs,
           const uint32_t *__restrict__ seg_first, const uint32_t *__restrict__ nseg, uint32_t n_slots,
           uint32_t width, uint32_t stride, const uint16_t *__restrict__ kguess,
           const uint64_t *__restrict__ incr, const uint64_t *__restrict__ incr_hi, double *errors,
           uint64_t len_cap, const double *__restrict__ err_tab) {
    // incr_hi != nullptr: incr holds the sums for binade kguess, incr_hi those for kguess + 1
    __shared__ double s_err[94];
    for (uint32_t i = threadIdx.x; i < 94; i += PT_TPB) s_err[i] = err_tab[i];
    __syncthreads();
    const uint64_t chains = (uint64_t)n_slots * width;
    const uint64_t warps = (uint64_t)gridDim.x * (PT_TPB / 32);
    for (uint64_t c = (uint64_t)blockIdx.x * (PT_TPB / 32) + (threadIdx.x >> 5); c < chains; c += warps) {
        const uint32_t s = (uint32_t)(c / width), pos = (uint32_t)(c % width);
        const uint32_t cnt = nseg[s];
        if (cnt == 0) continue;
        const uint32_t first = seg_first[s];
        double *cell = errors + (uint64_t)s * len_cap + pos;
        uint64_t sbits = (uint64_t)__double_as_longlong(*cell);
        uint32_t j = 0;
        while (j < cnt) {
            const uint32_t jj = j + lane_id();
            uint64_t inc = 0;
            uint32_t kg = 0, lo = 0, hi = 0;
            bool mine = false;
            const uint32_t k = (uint32_t)(sbits >> 52);
            if (jj < cnt) {
                const PtSeg sg = segs[first + jj];
                if (sg.slot == s) {  // a foreign segment adds nothing
                    mine = true;
                    lo = sg.lo;
                    hi = sg.hi;
                    inc = PT_HARD;
                    if (sg.data != PT_NONE) {
                        kg = kguess[(uint64_t)pos * stride + sg.data];
                        if (incr_hi && kg + 1 == k) {
                            inc = incr_hi[(uint64_t)pos * stride + sg.data];
                            kg = k;
                        }
                        else if (kg == k) inc = incr[(uint64_t)pos * stride + sg.data];
                    }
                }
            }
            const bool hard = mine && (inc >= PT_HARD || kg != k);
            if (hard) inc = 0;
            const uint64_t incl = warp_incl_scan_u64(inc);
            const bool stays = (uint32_t)((sbits + incl) >> 52) == k;
            const uint32_t nv = min(32u, cnt - j);
            const uint32_t bad = __ballot_sync(0xffffffffu, mine && (hard || !stays));
            const uint32_t f = bad ? (uint32_t)__ffs(bad) - 1 : 32u;
            const uint32_t nacc = min(f, nv);
            if (nacc) sbits += __shfl_sync(0xffffffffu, incl, nacc - 1);
            j += nacc;
            if (f < nv) {
                const uint32_t rlo = __shfl_sync(0xffffffffu, lo, f), rhi = __shfl_sync(0xffffffffu, hi, f);
                sbits = pt_replay_rows(sbits, bv, order, rlo, rhi, pos, s_err);
                j += 1;
            }
        }
        if (lane_id() == 0) *cell = __longlong_as_double((long long)sbits);
    }
}

// Run path: the chain over fixed tiles whose per-position quality histograms k_fused_columns
// wrote (PtHistGeom).  The binade k of the running sum is known exactly at every step, so the
// integer sum of a tile is the dot product of its histogram with row k of the increment table;
// 32 tiles are tried per step and accepted up to the first one that leaves the binade, ties,
// or has no usable histogram -- that one is replayed read by read.
__global__ void __launch_bounds__(PT_TPB)
k_pt_chain_hist(BatchView bv, const PtSeg *__restrict__ segs, const uint32_t *__restrict__ seg_first,
                const uint32_t *__restrict__ nseg, uint32_t n_slots, uint32_t width, const uint8_t *__restrict__ qh,
                PtHistGeom hg, const uint8_t *__restrict__ oob, const uint64_t *__restrict__ lut, double *errors,
                uint64_t len_cap, const double *__restrict__ err_tab) {
    extern __shared__ uint64_t s_lut[];  // [PT_LUT_NK][96]; entries past phred 93 are PT_HARD
    __shared__ double s_err[94];
    for (uint32_t i = threadIdx.x; i < 94; i += PT_TPB) s_err[i] = err_tab[i];
    for (uint32_t i = threadIdx.x; i < PT_LUT_NK * 96; i += PT_TPB) {
        const uint32_t k = i / 96, q = i - k * 96;
        s_lut[i] = q < 94 ? lut[k * 94 + q] : PT_HARD;
    }
    __syncthreads();
    const uint64_t chains = (uint64_t)n_slots * width;
    const uint64_t warps = (uint64_t)gridDim.x * (PT_TPB / 32);
    const uint32_t nq = hg.qrows / 4;
    for (uint64_t c = (uint64_t)blockIdx.x * (PT_TPB / 32) + (threadIdx.x >> 5); c < chains; c += warps) {
        const uint32_t s = (uint32_t)(c / width), pos = (uint32_t)(c % width);
        const uint32_t cnt = nseg[s];
        if (cnt == 0) continue;
        const uint32_t first = seg_first[s];
        const uint32_t colw = ((pos & 3u) * hg.CG + (pos >> 2)) * hg.QW;
        double *cell = errors + (uint64_t)s * len_cap + pos;
        uint64_t sbits = (uint64_t)__double_as_longlong(*cell);
        // A window = 32 consecutive segments, one per lane.  Its loads (segment descriptor, then the
        // segment's histogram words) do not depend on the running sum, so the next window is fetched
        // before the current one's replay: the chain is a string of dependent memory round trips,
        // this takes most of them off the critical path.
        constexpr int HW = 12;  // histogram words held in registers (48 quality values); more are read in the loop
        uint32_t lo = 0, hi = 0, hw[HW];
        bool mine = false, nodata = true;
        const uint32_t *hrow = nullptr;
        auto load_window = [&](uint32_t j0) {
            const uint32_t jj = j0 + lane_id();
            mine = false;
            nodata = true;
            hrow = nullptr;
#pragma unroll
            for (int u = 0; u < HW; u++) hw[u] = 0;
            if (jj < cnt) {
                const PtSeg sg = segs[first + jj];
                if (sg.slot == s) {  // a foreign segment adds nothing
                    mine = true;
                    lo = sg.lo;
                    hi = sg.hi;
                    if (sg.data != PT_NONE && !oob[sg.data]) {
                        nodata = false;
                        hrow = (const uint32_t *)(qh + (uint64_t)sg.data * hg.seg_bytes) + colw;
#pragma unroll
                        for (int u = 0; u < HW; u++) hw[u] = (uint32_t)u < nq ? __ldg(hrow + u) : 0u;
                    }
                }
            }
        };
        uint32_t j = 0;
        load_window(0);
        while (j < cnt) {
            const uint32_t k = (uint32_t)(sbits >> 52), krow = k - (uint32_t)PT_LUT_KMIN;
            uint64_t inc = 0;
            bool hard = false;
            if (mine) {
                if (nodata || krow >= (uint32_t)PT_LUT_NK) hard = true;
                else {
                    const uint64_t *lr = s_lut + krow * 96 + hg.qbase;
#pragma unroll
                    for (int u = 0; u < HW; u++) {
                        const uint32_t w = hw[u];
                        if (w == 0) continue;
                        const uint64_t *l4 = lr + 4 * u;
                        inc += (uint64_t)(w & 0xFF) * l4[0] + (uint64_t)((w >> 8) & 0xFF) * l4[1] +
                               (uint64_t)((w >> 16) & 0xFF) * l4[2] + (uint64_t)(w >> 24) * l4[3];
                    }
                    for (uint32_t m = HW; m < nq; m++) {  // wide quality ranges only
                        const uint32_t w = __ldg(hrow + m);
                        if (w == 0) continue;
                        const uint64_t *l4 = lr + 4 * m;
                        inc += (uint64_t)(w & 0xFF) * l4[0] + (uint64_t)((w >> 8) & 0xFF) * l4[1] +
                               (uint64_t)((w >> 16) & 0xFF) * l4[2] + (uint64_t)(w >> 24) * l4[3];
                    }
                    // a tabulated PT_HARD (tie, increment that cannot stay in the binade) times a non-zero
                    // count lifts the sum to >= 2^53; true sums of <= 255 in-binade increments that large
                    // leave the binade anyway.  (0 * PT_HARD adds nothing: unused rows do not poison.)
                    hard = inc >= PT_HARD;
                }
            }
            if (hard) inc = 0;
            const uint64_t incl = warp_incl_scan_u64(inc);
            const bool stays = (uint32_t)((sbits + incl) >> 52) == k;
            const uint32_t nv = min(32u, cnt - j);
            const uint32_t bad = __ballot_sync(0xffffffffu, mine && (hard || !stays));
            const uint32_t f = bad ? (uint32_t)__ffs(bad) - 1 : 32u;
            const uint32_t nacc = min(f, nv);
            if (nacc) sbits += __shfl_sync(0xffffffffu, incl, nacc - 1);
            const bool replay = f < nv;
            const uint32_t rlo = __shfl_sync(0xffffffffu, lo, f & 31), rhi = __shfl_sync(0xffffffffu, hi, f & 31);
            j += nacc + (replay ? 1u : 0u);
            if (j < cnt) load_window(j);  // in flight while the failed segment is replayed
            if (replay) sbits = pt_replay_rows<8>(sbits, bv, nullptr, rlo, rhi, pos, s_err);
        }
        if (lane_id() == 0) *cell = __longlong_as_double((long long)sbits);
    }
}

// ---- segments of the fused path: the record array is cut into fixed tiles of R
// ---- records (the tiles k_fused_columns sums over); a tile whose records all
// ---- belong to one flow-cell tile is one segment with precomputed sums, any
// ---- other tile is split into runs that the chain replays read by read.
// One warp per fixed tile: the lanes read the tile's slots together; a tile whose records all carry
// one slot (nearly all of them when reads arrive in tile runs) is settled by a vote, any other tile
// is walked by lane 0.
__device__ __forceinline__ bool pt_ftile_uniform(const uint32_t *__restrict__ slot, uint32_t r0, uint32_t r1) {
    const uint32_t s0 = slot[r0];
    bool same = s0 != PT_NONE;
    for (uint32_t r = r0 + lane_id(); r < r1; r += 32) same &= slot[r] == s0;
    return __all_sync(0xffffffffu, same);
}
__global__ void __launch_bounds__(PT_TPB)
k_pt_ftile_runs(const uint32_t *__restrict__ slot, uint32_t n, uint32_t R, uint32_t n_ftiles,
                uint32_t *__restrict__ runs, uint8_t *__restrict__ uniform) {
    const uint32_t warps = gridDim.x * (PT_TPB / 32);
    for (uint32_t t = blockIdx.x * (PT_TPB / 32) + (threadIdx.x >> 5); t < n_ftiles; t += warps) {
        const uint32_t r0 = t * R, r1 = min(n, r0 + R);
        const bool uni = pt_ftile_uniform(slot, r0, r1);
        if (lane_id() != 0) continue;
        uint32_t cnt = 1;
        if (!uni) {
            cnt = 0;
            uint32_t prev = PT_NONE;
            for (uint32_t r = r0; r < r1; r++) {
                const uint32_t s = slot[r];
                if (s != prev && s != PT_NONE) cnt++;
                prev = s;
            }
        }
        runs[t] = cnt;
        uniform[t] = uni;
    }
}
__global__ void __launch_bounds__(PT_TPB)
k_pt_ftile_segs(const uint32_t *__restrict__ slot, uint32_t n, uint32_t R, uint32_t n_ftiles,
                const uint32_t *__restrict__ seg_off, const uint8_t *__restrict__ uniform,
                PtSeg *__restrict__ segs, uint32_t *__restrict__ seg_lo, uint32_t *__restrict__ seg_hi) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;  // a thread per tile: uniform tiles are one store
    if (t >= n_ftiles) return;
    const uint32_t r0 = t * R, r1 = min(n, r0 + R);
    uint32_t w = seg_off[t];
    if (uniform[t]) {
        PtSeg g;
        g.lo = r0;
        g.hi = r1;
        g.slot = slot[r0];
        g.data = t;
        segs[w] = g;
        atomicMin(seg_lo + g.slot, w);
        atomicMax(seg_hi + g.slot, w + 1);
        return;
    }
    uint32_t prev = PT_NONE, lo = r0;
    for (uint32_t r = r0; r <= r1; r++) {
        const uint32_t s = r < r1 ? slot[r] : PT_NONE;
        if (s != prev || r == r1) {
            if (prev != PT_NONE) {
                PtSeg g;
                g.lo = lo;
                g.hi = r;
                g.slot = prev;
                g.data = PT_NONE;
                segs[w] = g;
                // segments are numbered in read order: a tile's run is [min, max] of its numbers
                atomicMin(seg_lo + prev, w);
                atomicMax(seg_hi + prev, w + 1);
                w++;
            }
            lo = r;
        }
        prev = s;
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
    if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&p->lut, (size_t)PT_LUT_NK * 94 * 8, false);
    if (rc == SQ_OK) {
        std::vector<uint64_t> lut((size_t)PT_LUT_NK * 94);
        for (int k = 0; k < PT_LUT_NK; k++)
            for (int q = 0; q < 94; q++) {
                const double e = pow(10.0, -((double)q / 10.0));  // same table as sq_ctx::d_err_table
                uint64_t eb;
                memcpy(&eb, &e, 8);
                lut[(size_t)k * 94 + q] = pt_increment((uint32_t)(PT_LUT_KMIN + k), eb);
            }
        CUDA_TRY(cudaMemcpyAsync(p->lut, lut.data(), lut.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        CUDA_TRY(cudaMemsetAsync(p->map_keys, 0xFF, (size_t)PT_MAP_CAP * 8, ctx->stream));
        CUDA_TRY(cudaMemsetAsync(&p->st->fail_idx, 0xFF, 16, ctx->stream));
        CUDA_TRY(cudaMemsetAsync(&p->st->qmin, 0xFF, 4, ctx->stream));
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
    sq_dfree(p->ctx, p->lut);
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

// ordered per-(tile, position) sums for one tile-sorted record array (see the note above k_pt_approx)
static int pt_accumulate(sq_pertile *p, sq_batch *b, const uint32_t *order, const uint32_t *seg_lo,
                         const uint32_t *seg_hi, uint32_t n_slots, uint32_t width, uint64_t base) {
    sq_ctx *ctx = p->ctx;
    const uint32_t n = (uint32_t)b->n;
    uint32_t CG = (width + 3) / 4;
    const uint32_t width4 = CG * 4;
    if (CG > PT_TPB) CG = PT_TPB;
    const uint32_t RG = PT_TPB / CG;
    const uint32_t windows = (width4 + CG * 4 - 1) / (CG * 4);
    // segment length: enough (segment, 4-position) threads to fill the machine twice over
    uint64_t want = (uint64_t)n * (width4 / 4) / ((uint64_t)ctx->num_sms * 4096);
    uint32_t seg_rows = 32;
    while (seg_rows < 1024 && seg_rows * 2 <= want) seg_rows *= 2;
    const uint32_t seg_cap = n / seg_rows + n_slots + 1;
    uint32_t *nseg = nullptr, *seg_first = nullptr, *n_segs_dev = nullptr;
    PtSeg *segs = nullptr;
    float *approx = nullptr;
    uint64_t *incr = nullptr;
    uint16_t *kguess = nullptr;
    SQ_TRY(sq_dalloc(ctx, (void **)&nseg, (size_t)n_slots * 4, false));
    SQ_TRY(sq_dalloc(ctx, (void **)&seg_first, (size_t)n_slots * 4, false));
    SQ_TRY(sq_dalloc(ctx, (void **)&n_segs_dev, 4, false));
    SQ_TRY(sq_dalloc(ctx, (void **)&segs, (size_t)seg_cap * sizeof(PtSeg), false));
    SQ_TRY(sq_dalloc(ctx, (void **)&approx, (size_t)seg_cap * width4 * 4, false));
    SQ_TRY(sq_dalloc(ctx, (void **)&incr, (size_t)seg_cap * width4 * 8, false));
    SQ_TRY(sq_dalloc(ctx, (void **)&kguess, (size_t)seg_cap * width4 * 2, false));
    const int slot_grid = sq_grid_for(ctx, n_slots, PT_TPB, 8);
    SQ_LAUNCH(ctx, k_pt_seg_counts, slot_grid, PT_TPB, 0, seg_lo, seg_hi, n_slots, seg_rows, nseg);
    SQ_TRY(sq_scan_exclusive_u32(ctx, nseg, seg_first, n_slots, n_segs_dev));
    SQ_LAUNCH(ctx, k_pt_seg_fill, sq_grid_for(ctx, (uint64_t)n_slots * 32, PT_TPB, 8), PT_TPB, 0, seg_lo, seg_hi,
              seg_first, n_slots, seg_rows, segs);
    dim3 sgrid((unsigned)sq_grid_for(ctx, (uint64_t)seg_cap, (int)RG, 64), windows);
    const BatchView bv = b->view();
    SQ_LAUNCH(ctx, k_pt_segment_sums<false>, sgrid, PT_TPB, 0, bv, order, segs, n_segs_dev, CG, RG, width4, seg_cap,
              ctx->d_err_table, p->lut, kguess, approx, incr, base, p->st);
    const int chain_grid = sq_grid_for(ctx, (uint64_t)n_slots * width * 32, PT_TPB, 32);
    SQ_LAUNCH(ctx, k_pt_guess, chain_grid, PT_TPB, 0, bv, order, segs, seg_first, nseg,
              n_slots, width, seg_cap, approx, p->errors, p->len_cap, ctx->d_err_table, kguess, 0);
    SQ_LAUNCH(ctx, k_pt_segment_sums<true>, sgrid, PT_TPB, (size_t)PT_LUT_NK * 94 * 8, bv, order, segs, n_segs_dev,
              CG, RG, width4, seg_cap, ctx->d_err_table, p->lut, kguess, approx, incr, base, p->st);
    SQ_LAUNCH(ctx, k_pt_chain, chain_grid, PT_TPB, 0, bv, order, segs, seg_first, nseg,
              n_slots, width, seg_cap, kguess, incr, (const uint64_t *)nullptr, p->errors, p->len_cap,
              ctx->d_err_table);
    sq_dfree(ctx, nseg);
    sq_dfree(ctx, seg_first);
    sq_dfree(ctx, n_segs_dev);
    sq_dfree(ctx, segs);
    sq_dfree(ctx, approx);
    sq_dfree(ctx, incr);
    sq_dfree(ctx, kguess);
    return SQ_OK;
}

void pt_plan_free(sq_ctx *ctx, PtPlan *pl) {
    void *ptrs[] = {pl->slot, pl->idx, pl->tmpk, pl->tmpv, pl->seg, pl->runs_cnt, pl->seg_off, pl->uniform,
                    pl->oob, pl->segs, pl->nseg, pl->qh};
    for (void *q : ptrs) sq_dfree(ctx, q);
    *pl = PtPlan();
}

// Reads that arrive in tile runs (every real Illumina file): no sort and no extra pass over the
// text.  Fixed tiles of R records (the tiles k_fused_columns walks) whose records belong to one
// flow-cell tile are the segments; k_fused_columns leaves a quality histogram per (tile, position)
// and k_pt_chain_hist turns it into the exact in-binade integer sum for the binade the chain is in.
static int pt_prepare_runs(sq_pertile *p, sq_batch *b, PtPlan *pl, uint32_t seg_cap) {
    sq_ctx *ctx = p->ctx;
    const uint32_t n = (uint32_t)b->n, n_ftiles = pl->n_ftiles, n_slots = pl->n_slots, R = pl->R;
    SQ_TRY(sq_dalloc(ctx, (void **)&pl->runs_cnt, (size_t)n_ftiles * 4, false));
    SQ_TRY(sq_dalloc(ctx, (void **)&pl->seg_off, (size_t)n_ftiles * 4, false));
    SQ_TRY(sq_dalloc(ctx, (void **)&pl->uniform, n_ftiles, false));
    if (!pl->oob) SQ_TRY(sq_dalloc(ctx, (void **)&pl->oob, n_ftiles, true));
    SQ_TRY(sq_dalloc(ctx, (void **)&pl->segs, (size_t)seg_cap * sizeof(PtSeg), false));
    SQ_TRY(sq_dalloc(ctx, (void **)&pl->seg, (size_t)n_slots * 8, false));
    SQ_TRY(sq_dalloc(ctx, (void **)&pl->nseg, (size_t)n_slots * 4, false));
    if (!pl->qh) SQ_TRY(sq_dalloc(ctx, (void **)&pl->qh, (size_t)n_ftiles * pl->hg.seg_bytes, false));
    uint32_t *seg_lo = pl->seg, *seg_hi = pl->seg + n_slots;
    CUDA_TRY(cudaMemsetAsync(seg_lo, 0xFF, (size_t)n_slots * 4, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(seg_hi, 0, (size_t)n_slots * 4, ctx->stream));
    const int wgrid = sq_grid_for(ctx, (uint64_t)n_ftiles * 32, PT_TPB, 8);
    SQ_LAUNCH(ctx, k_pt_ftile_runs, wgrid, PT_TPB, 0, pl->slot, n, R, n_ftiles, pl->runs_cnt, pl->uniform);
    SQ_TRY(sq_scan_exclusive_u32(ctx, pl->runs_cnt, pl->seg_off, n_ftiles, nullptr));
    SQ_LAUNCH(ctx, k_pt_ftile_segs, (n_ftiles + PT_TPB - 1) / PT_TPB, PT_TPB, 0, pl->slot, n, R, n_ftiles, pl->seg_off,
              pl->uniform, pl->segs, seg_lo, seg_hi);
    SQ_LAUNCH(ctx, k_pt_seg_counts, sq_grid_for(ctx, n_slots, PT_TPB, 8), PT_TPB, 0, seg_lo, seg_hi, n_slots, 1u,
              pl->nseg);
    return SQ_OK;
}

// Quality byte range sampled by k_fused_reads over the arrays seen so far (this one included: the
// kernel is already queued).  Known from the second array on without a sync: a stale range only
// sends the odd read through k_fused_columns' slow path.
int pt_quality_range(sq_pertile *p, uint32_t *qmin, uint32_t *qmax) {
    sq_ctx *ctx = p->ctx;
    if (p->qmin == 0xFFFFFFFFu) {
        PtState *h = (PtState *)((char *)ctx->h_scratch + 3400);
        CUDA_TRY(cudaMemcpyAsync(h, p->st, sizeof(PtState), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        p->qmin = h->qmin;
        p->qmax = h->qmax;
    }
    *qmin = p->qmin;
    *qmax = p->qmax;
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
    long long *tile = nullptr;
    SQ_TRY(sq_dalloc(ctx, (void **)&tile, (size_t)n * 8, false));
    SQ_LAUNCH(ctx, k_pt_tile, sq_grid_for(ctx, n, PT_TPB, 16), PT_TPB, 0, b->view(), tile, p->n_added, p->st);
    PtPlan pl;
    int rc = pt_prepare(p, b, tile, 0, 0, 0, PtHistGeom(), &pl, nullptr, nullptr);
    if (rc == SQ_OK) rc = pt_finish(p, b, &pl);
    else pt_plan_free(ctx, &pl);
    sq_dfree(ctx, tile);
    return rc;
}

// Tile ids -> slots, table growth, length counts; for reads in tile runs (R != 0: the caller
// will run k_fused_columns over fixed tiles of R records) also the segments and their hints.
int pt_prepare(sq_pertile *p, sq_batch *b, long long *tile, uint32_t R, uint32_t n_ftiles, uint32_t W,
               const PtHistGeom &hg, PtPlan *pl, uint8_t *qh, uint8_t *oob) {
    sq_ctx *ctx = p->ctx;
    const uint32_t n = (uint32_t)b->n;
    const uint64_t base = p->n_added;
    const int grid = sq_grid_for(ctx, n, PT_TPB, 16);
    *pl = PtPlan();
    pl->qh = qh;
    pl->oob = oob;
    pl->R = R;
    pl->n_ftiles = n_ftiles;
    pl->W = W;
    pl->hg = hg;
    SQ_TRY(sq_dalloc(ctx, (void **)&pl->slot, (size_t)n * 4, false));
    SQ_TRY(sq_dalloc(ctx, (void **)&pl->idx, (size_t)n * 4, false));
    SQ_TRY(pt_grow(p, p->n_slots ? p->n_slots : 1, b->max_len ? b->max_len : 1));
    // slot ids of new tiles may exceed slot_cap: slot_tile writes are guarded, and the
    // map is re-read after growing
    CUDA_TRY(cudaMemsetAsync(&p->st->n_changes, 0, 4, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(&p->st->tile_lo, 0xFF, 8, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(&p->st->tile_hi, 0, 8, ctx->stream));
    SQ_LAUNCH(ctx, k_pt_map_insert, grid, PT_TPB, 0, tile, n, base, p->map_keys, p->map_vals, p->slot_tile,
              p->slot_cap, p->st);
    SQ_LAUNCH(ctx, k_pt_map_lookup, grid, PT_TPB, 0, b->view(), tile, base, p->map_keys, p->map_vals, pl->slot,
              pl->idx, p->st);
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
    if (h->qmin <= h->qmax) {
        p->qmin = h->qmin;
        p->qmax = h->qmax;
    }
    b->tile_range_valid = true;  // tile_lo > tile_hi: no read of this array was kept
    b->tile_lo = h->tile_lo;
    b->tile_hi = h->tile_hi;
    pl->n_slots = h->n_slots;
    pl->width = (uint32_t)b->max_len;
    pl->fail_idx = h->fail_idx;
    pl->work = rc == SQ_OK && h->n_kept && h->n_slots;
    if (pl->work) {
        SQ_LAUNCH(ctx, k_pt_lengths, grid, PT_TPB, 0, b->view(), pl->slot, p->lengths, p->len_cap);
        // reads in tile runs: at most one extra segment per change of tile
        const uint64_t seg_cap = (uint64_t)n_ftiles + h->n_changes + 2;
        if (R && hg.qrows && b->max_len && seg_cap <= n / 16 + 64) {
            pl->runs = true;
            rc = pt_prepare_runs(p, b, pl, (uint32_t)seg_cap);
        }
    }
    if (rc != SQ_OK) pl->runs = false;
    return rc;
}

// After k_fused_columns (run path) or straight after pt_prepare (general path): the ordered sums.
int pt_finish(sq_pertile *p, sq_batch *b, PtPlan *pl) {
    sq_ctx *ctx = p->ctx;
    const uint32_t n = (uint32_t)b->n;
    const uint64_t base = p->n_added;
    const int grid = sq_grid_for(ctx, n, PT_TPB, 16);
    int rc = SQ_OK;
    if (pl->work && pl->runs) {
        // 4 CTAs fit an SM next to the increment table: one wave
        const int chain_grid = sq_grid_for(ctx, (uint64_t)pl->n_slots * pl->width * 32, PT_TPB, 4);
        const size_t lut_smem = (size_t)PT_LUT_NK * 96 * 8;
        CUDA_TRY(cudaFuncSetAttribute(k_pt_chain_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lut_smem));
        SQ_LAUNCH(ctx, k_pt_chain_hist, chain_grid, PT_TPB, lut_smem, b->view(), pl->segs, pl->seg, pl->nseg,
                  pl->n_slots, pl->width, pl->qh, pl->hg, pl->oob, p->lut, p->errors, p->len_cap, ctx->d_err_table);
    }
    else if (pl->work) {
        uint32_t key_bits = 1;
        while ((1ull << key_bits) < pl->n_slots) key_bits++;
        // PT_NONE (skipped reads) carries all-ones low bits, so it sorts last within those bits
        rc = sq_dalloc(ctx, (void **)&pl->tmpk, (size_t)n * 4, false);
        if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&pl->tmpv, (size_t)n * 4, false);
        if (rc == SQ_OK)
            rc = sq_radix_sort_pairs(ctx, pl->slot, pl->idx, pl->tmpk, pl->tmpv, n, pl->fail_idx != ~0ULL ? 32 : key_bits);
        if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&pl->seg, (size_t)pl->n_slots * 8, true);
        if (rc == SQ_OK) {
            uint32_t *seg_lo = pl->seg, *seg_hi = pl->seg + pl->n_slots;
            SQ_LAUNCH(ctx, k_pt_segments, grid, PT_TPB, 0, pl->slot, n, seg_lo, seg_hi);
            if (pl->width) rc = pt_accumulate(p, b, pl->idx, seg_lo, seg_hi, pl->n_slots, pl->width, base);
        }
    }
    if (rc == SQ_OK && pl->fail_idx != ~0ULL) {
        p->skipped = true;
        p->skipped_record = pl->fail_idx;
        const uint64_t r = pl->fail_idx - base;
        rc = sq_batch_get_name(b, r, p->skipped_name);
    }
    p->n_added += n;
    pt_plan_free(ctx, pl);
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


// ---------------------------------------------------------------------------
// Sharded runs.  total_errors[tile][pos] is a chain over the
// reads of a tile in read order, so one rank must see all reads of a tile: the
// lowest rank that holds any read of it.  The other ranks copy their records of
// such tiles, whole and in order, into one FASTQ text (sq_batch_select_tiles)
// and hand it to the owner, which parses it like any other record array and
// adds it after its own reads.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(PT_TPB)
k_pt_select_sizes(BatchView bv, uint32_t limit, const long long *__restrict__ ids, uint32_t n_ids,
                  uint32_t *__restrict__ sizes) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < bv.n; r += gridDim.x * blockDim.x) {
        uint32_t sz = 0;
        if (r < limit) {
            const long long t = tile_id_of(bv.text + bv.name_off[r], bv.seq_off[r] - 1 - bv.name_off[r]);
            uint32_t lo = 0, hi = n_ids;  // ids sorted ascending
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (ids[mid] < t) lo = mid + 1;
                else hi = mid;
            }
            if (t >= 0 && lo < n_ids && ids[lo] == t) sz = bv.name_off[r + 1] - bv.name_off[r];
        }
        sizes[r] = sz;
    }
}
__global__ void __launch_bounds__(PT_TPB)
k_pt_select_copy(BatchView bv, const uint32_t *__restrict__ sizes, const uint32_t *__restrict__ offs,
                 uint8_t *__restrict__ out) {
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < bv.n; r += warps) {
        const uint32_t sz = sizes[r];
        if (!sz) continue;
        const uint8_t *src = bv.text + bv.name_off[r] - 1;
        uint8_t *dst = out + offs[r];
        for (uint32_t i = lane_id(); i < sz; i += 32) dst[i] = src[i];
    }
}

// Records r < limit_records of a FASTQ-text record array whose tile id is in tile_ids[0..n_ids)
// (host array, ascending), copied in order to dev_out (DEVICE, cap bytes).  dev_out == NULL only
// sizes the output.  *nbytes = bytes needed / written.
extern "C" int sq_batch_select_tiles(sq_batch *b, const int64_t *tile_ids, uint64_t n_ids, uint64_t limit_records,
                                     uint8_t *dev_out, uint64_t cap, uint64_t *nbytes) {
    sq_ctx *ctx = b->ctx;
    *nbytes = 0;
    if (b->name_len != nullptr) {
        sq_set_error("sq_batch_select_tiles needs a FASTQ-text record array");
        return SQ_E_ARG;
    }
    if (b->n == 0 || n_ids == 0 || limit_records == 0) return SQ_OK;
    if (b->tile_range_valid) {  // PerTileQuality has seen this array: none of the tiles may lie in its range
        bool any = false;
        for (uint64_t i = 0; i < n_ids && !any; i++)
            any = tile_ids[i] >= 0 && (uint64_t)tile_ids[i] >= b->tile_lo && (uint64_t)tile_ids[i] <= b->tile_hi;
        if (!any) return SQ_OK;
    }
    CUDA_TRY(cudaSetDevice(ctx->device));
    const uint32_t n = (uint32_t)b->n;
    long long *ids = nullptr;
    uint32_t *sizes = nullptr, *offs = nullptr;
    SQ_TRY(sq_dalloc(ctx, (void **)&ids, n_ids * 8, false));
    SQ_TRY(sq_dalloc(ctx, (void **)&sizes, (size_t)n * 4, false));
    SQ_TRY(sq_dalloc(ctx, (void **)&offs, (size_t)n * 4 + 4, false));
    CUDA_TRY(cudaMemcpyAsync(ids, tile_ids, n_ids * 8, cudaMemcpyHostToDevice, ctx->stream));
    const uint32_t limit = limit_records > n ? n : (uint32_t)limit_records;
    const int grid = sq_grid_for(ctx, n, PT_TPB, 16);
    SQ_LAUNCH(ctx, k_pt_select_sizes, grid, PT_TPB, 0, b->view(), limit, ids, (uint32_t)n_ids, sizes);
    SQ_TRY(sq_scan_exclusive_u32(ctx, sizes, offs, n, offs + n));
    uint32_t *h_total = (uint32_t *)((char *)ctx->h_scratch + 3400);
    CUDA_TRY(cudaMemcpyAsync(h_total, offs + n, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // also: the pageable tile_ids have been read
    *nbytes = *h_total;
    int rc = SQ_OK;
    if (dev_out && *h_total) {
        if (*h_total > cap) {
            sq_set_error("selected records need %u bytes, buffer has %llu", *h_total, (unsigned long long)cap);
            rc = SQ_E_ARG;
        }
        else {
            SQ_LAUNCH(ctx, k_pt_select_copy, sq_grid_for(ctx, (uint64_t)n * 32, PT_TPB, 32), PT_TPB, 0, b->view(), sizes,
                      offs, dev_out);
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        }
    }
    sq_dfree(ctx, ids);
    sq_dfree(ctx, sizes);
    sq_dfree(ctx, offs);
    return rc;
}
