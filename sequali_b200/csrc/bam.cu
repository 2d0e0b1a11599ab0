// bam.cu -- BAM alignment records -> packed name|seq|qual|tags record array
// (reference BamParser__next__, _qcmodule.c:1623-1694; decode_bam_sequence
// :1266-1334; decode_bam_qualities :1354-1360).
//
// sq_batch_from_bam_bytes: the whole of BamParser__next__ on the device.  The host copies the bytes it read
// (one cudaMemcpyAsync from pinned memory) and never looks at a record:
//   k_bam_candidates  every byte offset is tested for "a complete, self-consistent record header starts here"
//                     (reference id first: three coalesced word loads decide four offsets; the rest of the
//                     test only runs for the few offsets that survive) -> bitmap, 1 bit per byte
//   scan of the bitmap popcounts -> rank of every candidate
//   k_bam_nodes       per candidate: where its block_size points (rank of the successor / end of data /
//                     an offset that is no candidate)
//   k_bam_hop         pointer doubling from the first record: after round k everything within 2^k hops of
//                     the start is marked; ceil(log2(candidates)) rounds.  False candidates (byte patterns
//                     inside names, bases or tags that look like a header) are simply never reached.
//   k_bam_last / k_bam_pick   end of the chain, records dropped for flag & 0x900, kept offsets compacted
//   k_bam_chase       the plain one-thread walk, only from a point where the chain leaves the candidates
//                     (a record the reference would accept although its header is not self-consistent)
// so the result is the reference's chain exactly; the candidate test is a speculation that is verified.
// Then, for both entry points:
//   k_bam_sizes   one thread per record reads the 36-byte header -> field sizes, header consistency
//   exclusive scan of the packed sizes -> output offsets; scan of the tiles per record
//   k_bam_decode  one warp per TILE of 4096 bases (a 1 Mb read is 245 tiles, not one warp's afternoon):
//                 4-bit -> ASCII nucleotides and quality + 33 (or '!' when the qualities are absent, i.e.
//                 the first byte is 0xff), produced four output bytes per lane at aligned addresses;
//                 tile 0 also copies the name and the raw tags
#include "common.cuh"

constexpr int BAM_TPB = 256;
constexpr uint32_t BAM_TILE = 4096;  // bases per decode tile

__device__ __forceinline__ uint32_t brd16(const uint8_t *p) { return p[0] | (uint32_t)p[1] << 8; }
__device__ __forceinline__ uint32_t brd32(const uint8_t *p) {
    return p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24;
}

// ---- the record chain on the device ------------------------------------------------------------------
struct BamWalk {  // device scalars of one walk (in ctx->d_scratch)
    uint32_t n_cand, head, last, pad;
    unsigned long long n_kept, n_skipped, consumed, broken_at;  // broken_at == ~0: the chain is whole
};
constexpr uint32_t BAM_END = 0xffffffffu;     // successor: the data end here
constexpr uint32_t BAM_BROKEN = 0xfffffffeu;  // successor: a record that is no candidate

// is p the start of a complete record whose header fields agree with its block_size?
__device__ __forceinline__ bool bam_plausible(const uint8_t *bam, uint64_t p, uint64_t nbytes, uint32_t n_ref) {
    if (p + 36 > nbytes) return false;
    const uint8_t *h = bam + p;
    const uint64_t block = brd32(h);
    if (block < 32 || p + 4 + block > nbytes) return false;
    if (brd32(h + 4) + 1u > n_ref || brd32(h + 24) + 1u > n_ref) return false;  // -1 .. n_ref - 1
    const uint64_t l_name = h[12], n_cigar = brd16(h + 16), l_seq = brd32(h + 20);
    if (l_name == 0) return false;
    if (36 + l_name + 4 * n_cigar + (l_seq + 1) / 2 + l_seq > 4 + block) return false;
    return h[36 + l_name - 1] == 0;
}

// thread = four consecutive offsets; eight threads assemble one bitmap word
__global__ void __launch_bounds__(BAM_TPB)
k_bam_candidates(const uint8_t *__restrict__ bam, uint64_t nbytes, uint32_t n_ref, uint32_t *__restrict__ bitmap,
                 uint32_t *__restrict__ counts, uint32_t n_words) {
    const uint32_t *w = (const uint32_t *)bam;  // the buffer is padded: words up to nbytes / 4 + 3 are readable
    // n_words * 8 is a multiple of eight, and the eight threads of a word stay together in the loop
    for (uint64_t t = (uint64_t)blockIdx.x * BAM_TPB + threadIdx.x; t < (uint64_t)n_words * 8; t += (uint64_t)gridDim.x * BAM_TPB) {
        const uint32_t w1 = w[t + 1], w2 = w[t + 2];  // bytes 4t+4 .. 4t+11: the reference ids of offsets 4t .. 4t+3
        uint32_t nib = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t ref_id = __funnelshift_r(w1, w2, 8 * j);
            if (ref_id + 1u <= n_ref && bam_plausible(bam, 4 * t + j, nbytes, n_ref)) nib |= 1u << j;
        }
        uint32_t word = nib << (4 * (threadIdx.x & 7));
        const uint32_t group = 0xffu << (lane_id() & 24);
        word |= __shfl_xor_sync(group, word, 1);
        word |= __shfl_xor_sync(group, word, 2);
        word |= __shfl_xor_sync(group, word, 4);
        if ((threadIdx.x & 7) == 0) {
            bitmap[t >> 3] = word;
            counts[t >> 3] = __popc(word);
        }
    }
}

__device__ __forceinline__ bool bam_is_candidate(const uint32_t *bitmap, uint64_t p) { return bitmap[p >> 5] >> (p & 31) & 1; }
__device__ __forceinline__ uint32_t bam_rank(const uint32_t *bitmap, const uint32_t *rank, uint64_t p) {
    return rank[p >> 5] + __popc(bitmap[p >> 5] & ((1u << (p & 31)) - 1));
}

// what follows the complete record that ends at e (the host loop of sq_bam_walk, one step)
__device__ __forceinline__ uint32_t bam_successor(const uint8_t *bam, uint64_t nbytes, const uint32_t *bitmap,
                                                  const uint32_t *rank, uint64_t e) {
    if (e + 4 >= nbytes) return BAM_END;
    if (e + 4 + (uint64_t)brd32(bam + e) > nbytes) return BAM_END;  // an incomplete record: the caller's leftover
    return bam_is_candidate(bitmap, e) ? bam_rank(bitmap, rank, e) : BAM_BROKEN;
}

// thread per bitmap word: offset, successor and flag bit of its candidates
__global__ void __launch_bounds__(BAM_TPB)
k_bam_nodes(const uint8_t *__restrict__ bam, uint64_t nbytes, const uint32_t *__restrict__ bitmap,
            const uint32_t *__restrict__ rank, uint32_t n_words, uint32_t *__restrict__ node_off,
            uint32_t *__restrict__ succ, uint32_t *__restrict__ jump, uint8_t *__restrict__ dropped, BamWalk *walk) {
    for (uint32_t wi = blockIdx.x * BAM_TPB + threadIdx.x; wi < n_words; wi += gridDim.x * BAM_TPB) {
        uint32_t bits = bitmap[wi], i = rank[wi];
        const uint32_t n_cand = walk->n_cand;
        while (bits) {
            const uint64_t p = (uint64_t)wi * 32 + (__ffs(bits) - 1);
            bits &= bits - 1;
            const uint8_t *h = bam + p;
            const uint32_t s = bam_successor(bam, nbytes, bitmap, rank, p + 4 + brd32(h));
            node_off[i] = (uint32_t)p;
            succ[i] = s;
            jump[i] = s < n_cand ? s : n_cand;  // the sentinel node n_cand points at itself
            dropped[i] = (brd16(h + 18) & 0x900u) != 0;  // secondary / supplementary (:1633)
            i++;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {  // where the chain starts
        const uint32_t n_cand = walk->n_cand;
        jump[n_cand] = n_cand;
        uint32_t head = BAM_END;
        walk->broken_at = ~0ULL;
        if (nbytes > 4 && 4 + (uint64_t)brd32(bam) <= nbytes) {
            if (bam_is_candidate(bitmap, 0)) head = 0;
            else {
                head = BAM_BROKEN;
                walk->broken_at = 0;
            }
        }
        walk->head = head;
        walk->last = 0;
        walk->n_kept = walk->n_skipped = walk->consumed = 0;
    }
}

__global__ void k_bam_mark_head(uint8_t *mark, const BamWalk *walk) {
    if (walk->head == 0) mark[0] = 1;
}

// one round of pointer doubling: marked nodes mark the node `jump` ahead, jumps double
__global__ void __launch_bounds__(BAM_TPB)
k_bam_hop(uint8_t *__restrict__ mark, const uint32_t *__restrict__ jump_in, uint32_t *__restrict__ jump_out, uint32_t n) {
    for (uint32_t i = blockIdx.x * BAM_TPB + threadIdx.x; i <= n; i += gridDim.x * BAM_TPB) {
        const uint32_t j = jump_in[i];
        if (mark[i] && j < n) mark[j] = 1;  // every marked node is on the chain, so is what it reaches
        jump_out[i] = jump_in[j];
    }
}

// records of the chain: kept ones flagged for the scan, the last one found (the chain only moves forward)
__global__ void __launch_bounds__(BAM_TPB)
k_bam_flags(const uint8_t *__restrict__ mark, const uint8_t *__restrict__ dropped, uint32_t n, uint32_t *__restrict__ keep,
            BamWalk *walk) {
    uint32_t last = 0, skipped = 0;
    for (uint32_t i = blockIdx.x * BAM_TPB + threadIdx.x; i < n; i += gridDim.x * BAM_TPB) {
        const bool on = mark[i];
        keep[i] = on && !dropped[i];
        skipped += on && dropped[i];
        if (on) last = i;
    }
    last = warp_max_u32(last);
    skipped = warp_sum_u32(skipped);
    if (lane_id() == 0) {
        if (last) atomicMax(&walk->last, last);
        if (skipped) atomicAdd(&walk->n_skipped, (unsigned long long)skipped);
    }
}

__global__ void __launch_bounds__(BAM_TPB)
k_bam_pick(const uint32_t *__restrict__ keep, const uint32_t *__restrict__ keep_rank, const uint32_t *__restrict__ node_off,
           uint32_t n, uint64_t *__restrict__ rec_off) {
    for (uint32_t i = blockIdx.x * BAM_TPB + threadIdx.x; i < n; i += gridDim.x * BAM_TPB)
        if (keep[i]) rec_off[keep_rank[i]] = node_off[i];
}

__global__ void k_bam_last(const uint8_t *bam, const uint32_t *node_off, const uint32_t *succ, const uint32_t *kept_total,
                           BamWalk *walk) {
    walk->n_kept = *kept_total;
    if (walk->head != 0) return;  // empty, or broken at offset 0
    const uint32_t l = walk->last;
    const unsigned long long e = (unsigned long long)node_off[l] + 4 + brd32(bam + node_off[l]);
    walk->consumed = e;
    if (succ[l] == BAM_BROKEN) walk->broken_at = e;
}

// the plain walk from `walk->broken_at` (sq_bam_walk on the device, one thread): only for chains that leave
// the candidates.  Appends to rec_off behind the records found so far.
__global__ void k_bam_chase(const uint8_t *bam, uint64_t nbytes, uint64_t *rec_off, uint64_t cap, BamWalk *walk, int *overflow) {
    uint64_t pos = walk->broken_at, kept = walk->n_kept, skipped = walk->n_skipped;
    while (pos + 4 < nbytes) {
        const uint64_t end = pos + 4 + brd32(bam + pos);
        if (end > nbytes) break;
        uint32_t flag = 0;
        if (pos + 18 < nbytes) flag = bam[pos + 18];
        if (pos + 19 < nbytes) flag |= (uint32_t)bam[pos + 19] << 8;
        if (flag & 0x900u) skipped++;
        else {
            if (kept == cap) {
                *overflow = 1;
                break;
            }
            rec_off[kept++] = pos;
        }
        pos = end;
    }
    walk->n_kept = kept;
    walk->n_skipped = skipped;
    walk->consumed = pos;
}

// ---- decode -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BAM_TPB)
k_bam_sizes(const uint8_t *__restrict__ bam, uint64_t nbytes, const uint64_t *__restrict__ rec_off, uint32_t n,
            uint32_t *__restrict__ sizes, uint32_t *__restrict__ tiles, uint32_t *name_len, uint32_t *seq_len,
            uint32_t *tags_len, unsigned int *max_len, unsigned long long *first_bad) {
    uint32_t lmax = 0;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const uint64_t p = rec_off[r];
        uint32_t nl = 0, l_seq = 0, tl = 0;
        if (p + 36 > nbytes) atomicMin(first_bad, (unsigned long long)r << 1);
        else {
            const uint8_t *h = bam + p;
            const uint64_t block = brd32(h), l_name = h[12], n_cigar = brd16(h + 16);
            const uint64_t fixed = 36 + l_name + 4 * n_cigar + ((uint64_t)brd32(h + 20) + 1) / 2 + brd32(h + 20);
            if (p + 4 + block > nbytes || fixed > 4 + block) atomicMin(first_bad, (unsigned long long)r << 1 | 1);
            else {
                l_seq = brd32(h + 20);
                nl = l_name ? (uint32_t)l_name - 1 : 0;  // drop the terminating NUL
                tl = (uint32_t)(4 + block - fixed);
            }
        }
        name_len[r] = nl;
        seq_len[r] = l_seq;
        tags_len[r] = tl;
        sizes[r] = nl + 2 * l_seq + tl;
        tiles[r] = l_seq ? (l_seq + BAM_TILE - 1) / BAM_TILE : 1;
        lmax = max(lmax, l_seq);
    }
    lmax = warp_max_u32(lmax);
    if (lane_id() == 0 && lmax) atomicMax(max_len, lmax);
}

// the warp writes gen(i .. i+3) to dst[i .. i+3] for i in [0, n): whole words where dst is aligned
template <class Gen4, class Gen1>
__device__ __forceinline__ void bam_emit(uint8_t *dst, uint32_t n, uint32_t lane, Gen4 gen4, Gen1 gen1) {
    const uint32_t head = min(n, (uint32_t)(-(uintptr_t)dst & 3));
    if (lane < head) dst[lane] = gen1(lane);
    const uint32_t words = (n - head) >> 2;
    for (uint32_t k = lane; k < words; k += 32) *(uint32_t *)(dst + head + 4 * k) = gen4(head + 4 * k);
    const uint32_t done = head + 4 * words;
    if (lane < n - done) dst[done + lane] = gen1(done + lane);
}

__device__ __forceinline__ uint32_t bam_letter(uint32_t code) {  // "=ACMGRSVTWYHKDBN"
    const uint64_t t = code & 8 ? 0x4e42444b48595754ULL : 0x565352474d43413dULL;
    return (uint32_t)(t >> (8 * (code & 7))) & 0xff;
}

__global__ void __launch_bounds__(BAM_TPB)
k_bam_decode(const uint8_t *__restrict__ bam, const uint64_t *__restrict__ rec_off, uint32_t n,
             const uint32_t *__restrict__ out_off, const uint32_t *__restrict__ tile_base,
             const uint32_t *__restrict__ tile_total, const uint32_t *__restrict__ name_len,
             const uint32_t *__restrict__ seq_len, const uint32_t *__restrict__ tags_len, uint8_t *__restrict__ out,
             uint32_t *name_off, uint32_t *seq_off, uint32_t *qual_off, uint32_t *tags_off) {
    const uint32_t warps = gridDim.x * (BAM_TPB / 32);
    const uint32_t lane = lane_id();
    const uint32_t n_tiles = *tile_total;
    for (uint32_t t = blockIdx.x * (BAM_TPB / 32) + (threadIdx.x >> 5); t < n_tiles; t += warps) {
        // record of tile t: the last r with tile_base[r] <= t (32-ary search, the lanes probe together)
        uint32_t lo = 0, hi = n;
        if (n_tiles == n) lo = t;  // one tile per record
        else
            while (hi - lo > 1) {
                const uint32_t step = (hi - lo + 31) / 32, idx = lo + lane * step;
                const uint32_t le = __ballot_sync(0xffffffffu, idx < hi && tile_base[idx] <= t);
                lo += (__popc(le) - 1) * step;
                hi = min(hi, lo + step);
            }
        const uint32_t r = lo, k = t - tile_base[r];
        const uint8_t *h = bam + rec_off[r];
        const uint32_t l_name = h[12], n_cigar = brd16(h + 16);
        const uint32_t nl = name_len[r], sl = seq_len[r], tl = tags_len[r];
        const uint8_t *name = h + 36, *seq = name + l_name + 4 * n_cigar;
        const uint8_t *qual = seq + (sl + 1) / 2, *tags = qual + sl;
        uint8_t *o = out + out_off[r];
        if (k == 0) {
            if (lane == 0) {
                name_off[r] = out_off[r];
                seq_off[r] = out_off[r] + nl;
                qual_off[r] = out_off[r] + nl + sl;
                tags_off[r] = out_off[r] + nl + 2 * sl;
            }
            for (uint32_t i = lane; i < nl; i += 32) o[i] = name[i];
            bam_emit(o + nl + 2 * sl, tl, lane,
                     [&](uint32_t i) { return tags[i] | (uint32_t)tags[i + 1] << 8 | (uint32_t)tags[i + 2] << 16 | (uint32_t)tags[i + 3] << 24; },
                     [&](uint32_t i) { return tags[i]; });
        }
        const uint32_t b0 = k * BAM_TILE, nb = min(BAM_TILE, sl - b0);
        if (sl == 0) continue;
        // bases b0 .. b0+nb: two per input byte, high nibble first
        bam_emit(o + nl + b0, nb, lane,
                 [&](uint32_t i) {
                     const uint32_t g = b0 + i;
                     const uint8_t *s = seq + (g >> 1);
                     const uint32_t three = s[0] << 16 | s[1] << 8 | s[2];          // six codes
                     const uint32_t four = g & 1 ? three >> 4 : three >> 8;        // the four wanted, first one on top
                     return bam_letter(four >> 12 & 15) | bam_letter(four >> 8 & 15) << 8 | bam_letter(four >> 4 & 15) << 16 |
                            bam_letter(four & 15) << 24;
                 },
                 [&](uint32_t i) {
                     const uint32_t g = b0 + i, b = seq[g >> 1];
                     return (uint8_t)bam_letter(g & 1 ? b & 15 : b >> 4);
                 });
        const bool missing = qual[0] == 0xff;  // :1658
        bam_emit(o + nl + sl + b0, nb, lane,
                 [&](uint32_t i) {
                     const uint8_t *q = qual + b0 + i;
                     const uint32_t x = q[0] | (uint32_t)q[1] << 8 | (uint32_t)q[2] << 16 | (uint32_t)q[3] << 24;
                     return missing ? 0x21212121u : ((x & 0x7f7f7f7fu) + 0x21212121u) ^ (x & 0x80808080u);  // bytes + 33
                 },
                 [&](uint32_t i) { return missing ? (uint8_t)'!' : (uint8_t)(qual[b0 + i] + 33); });
    }
}

// The block_size chain of BamParser__next__ (_qcmodule.c:1623-1637) over the bytes the caller has read so
// far: offsets of the complete records to keep, how many were dropped for being secondary or
// supplementary alignments (flag & 0x900, :1633), and how many bytes the complete records cover.
extern "C" int sq_bam_walk(const uint8_t *bam, uint64_t nbytes, uint64_t *rec_off, uint64_t cap, uint64_t *n_kept,
                           uint64_t *n_skipped, uint64_t *consumed) {
    uint64_t pos = 0, kept = 0, skipped = 0;
    while (pos + 4 < nbytes) {
        const uint64_t block = (uint64_t)bam[pos] | (uint64_t)bam[pos + 1] << 8 | (uint64_t)bam[pos + 2] << 16 |
                               (uint64_t)bam[pos + 3] << 24;
        const uint64_t end = pos + 4 + block;
        if (end > nbytes) break;
        uint32_t flag = 0;
        if (pos + 18 < nbytes) flag = bam[pos + 18];
        if (pos + 19 < nbytes) flag |= (uint32_t)bam[pos + 19] << 8;
        if (flag & (0x100u | 0x800u)) skipped++;
        else {
            if (kept == cap) {
                sq_set_error("BAM records smaller than their fixed fields");
                return SQ_E_FORMAT;
            }
            rec_off[kept++] = pos;
        }
        pos = end;
    }
    *n_kept = kept;
    *n_skipped = skipped;
    *consumed = pos;
    return SQ_OK;
}

// sizes -> offsets -> decode of the records at d_off[0..n) of the device copy d_bam
static int bam_decode_records(sq_ctx *ctx, const uint8_t *d_bam, uint64_t nbytes, const uint64_t *d_off, uint64_t n,
                              sq_batch **out, uint64_t *packed_len) {
    sq_batch *b = new sq_batch();
    b->ctx = ctx;
    b->n = n;
    uint32_t *sizes = nullptr, *offs = nullptr, *tiles = nullptr, *tile_base = nullptr;
    const size_t n4 = (size_t)((n + 3) & ~3ULL);
    const size_t meta_bytes = n4 * 4 * 7 + n4 * 8;
    int rc = sq_dalloc(ctx, (void **)&sizes, (n + 1) * 4, false);
    if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&offs, (n + 1) * 4, false);
    if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&tiles, (n + 1) * 4, false);
    if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&tile_base, (n + 1) * 4, false);
    if (rc == SQ_OK) rc = sq_dalloc(ctx, &b->meta_block, meta_bytes, true);
    char *ds = (char *)ctx->d_scratch, *hs = (char *)ctx->h_scratch;
    unsigned int *d_max = (unsigned int *)(ds + 3600);
    uint32_t *d_total = (uint32_t *)(ds + 3604), *d_tiles = (uint32_t *)(ds + 3828);
    unsigned long long *d_bad = (unsigned long long *)(ds + 3832);
    auto fail = [&](int code) {
        sq_dfree(ctx, sizes);
        sq_dfree(ctx, offs);
        sq_dfree(ctx, tiles);
        sq_dfree(ctx, tile_base);
        sq_batch_free(b);
        return code;
    };
    if (rc != SQ_OK) return fail(rc);
    uint32_t *d = (uint32_t *)b->meta_block;
    b->name_off = d;
    b->seq_off = d + n4;
    b->seq_len = d + 2 * n4;
    b->qual_off = d + 3 * n4;
    b->name_len = d + 4 * n4;
    b->tags_off = d + 5 * n4;
    b->tags_len = d + 6 * n4;
    b->err_sum = (double *)(d + 7 * n4);
    cudaStream_t st = sq_cur_stream(ctx);
    if (cudaMemsetAsync(d_max, 0, 8, st) != cudaSuccess || cudaMemsetAsync(d_bad, 0xff, 8, st) != cudaSuccess)
        return fail(sq_cuda_fail(cudaGetLastError(), "BAM decode setup", __FILE__, __LINE__));
    const int grid = sq_grid_for(ctx, n, BAM_TPB, 16);
    auto sizes_pass = [&]() -> int {
        SQ_LAUNCH(ctx, k_bam_sizes, grid, BAM_TPB, 0, d_bam, nbytes, d_off, (uint32_t)n, sizes, tiles, b->name_len,
                  b->seq_len, b->tags_len, d_max, d_bad);
        SQ_TRY(sq_scan_exclusive_u32(ctx, sizes, offs, (uint32_t)n, d_total));
        SQ_TRY(sq_scan_exclusive_u32(ctx, tiles, tile_base, (uint32_t)n, d_tiles));
        CUDA_TRY(cudaMemcpyAsync(hs + 3600, ds + 3600, 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(hs + 3828, ds + 3828, 12, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        return SQ_OK;
    };
    if ((rc = sizes_pass()) != SQ_OK) return fail(rc);
    unsigned long long bad;
    memcpy(&bad, hs + 3832, 8);
    if (bad != ~0ULL) {
        if (bad & 1) sq_set_error("BAM record %llu is inconsistent with its block_size", bad >> 1);
        else sq_set_error("BAM record %llu starts outside the buffer", bad >> 1);
        return fail(bad & 1 ? SQ_E_FORMAT : SQ_E_ARG);
    }
    b->max_len = *(uint32_t *)(hs + 3600);
    b->nbytes = *(uint32_t *)(hs + 3604);
    const uint32_t n_tiles = *(uint32_t *)(hs + 3828);
    if ((rc = sq_dalloc(ctx, (void **)&b->text, b->nbytes + 64, false)) != SQ_OK) return fail(rc);
    auto decode_pass = [&]() -> int {
        CUDA_TRY(cudaMemsetAsync(b->text + b->nbytes, 0, 64, st));
        const int wgrid = sq_grid_for(ctx, (uint64_t)n_tiles * 32, BAM_TPB, 16);
        SQ_LAUNCH(ctx, k_bam_decode, wgrid, BAM_TPB, 0, d_bam, d_off, (uint32_t)n, offs, tile_base, d_tiles, b->name_len,
                  b->seq_len, b->tags_len, b->text, b->name_off, b->seq_off, b->qual_off, b->tags_off);
        // the record array is used on the launch stream from here on: the decode (parser stream) is done first
        CUDA_TRY(cudaStreamSynchronize(st));
        return SQ_OK;
    };
    if ((rc = decode_pass()) != SQ_OK) return fail(rc);
    sq_dfree(ctx, sizes);
    sq_dfree(ctx, offs);
    sq_dfree(ctx, tiles);
    sq_dfree(ctx, tile_base);
    *packed_len = b->nbytes;
    *out = b;
    return SQ_OK;
}

static int bam_upload(sq_ctx *ctx, const uint8_t *bam, uint64_t nbytes, uint8_t **d_bam) {
    if (nbytes >= 0xC0000000ULL) {  // packed output is at most 4/3 of the input
        sq_set_error("BAM chunk of %llu bytes is too large for one record array", (unsigned long long)nbytes);
        return SQ_E_LIMIT;
    }
    CUDA_TRY(cudaSetDevice(ctx->device));
    SQ_TRY(sq_dalloc(ctx, (void **)d_bam, nbytes + 64, false));
    CUDA_TRY(cudaMemcpyAsync(*d_bam, bam, nbytes, cudaMemcpyHostToDevice, sq_cur_stream(ctx)));
    CUDA_TRY(cudaMemsetAsync(*d_bam + nbytes, 0, 64, sq_cur_stream(ctx)));
    return SQ_OK;
}

extern "C" int sq_batch_from_bam(sq_ctx *ctx, const uint8_t *bam, uint64_t nbytes, const uint64_t *rec_off,
                                 uint64_t n, sq_batch **out, uint64_t *packed_len) {
    *out = nullptr;
    *packed_len = 0;
    SqParserScope on_parser_stream(ctx);
    uint8_t *d_bam = nullptr;
    uint64_t *d_off = nullptr;
    int rc = bam_upload(ctx, bam, nbytes, &d_bam);
    if (rc == SQ_OK) rc = sq_dalloc(ctx, (void **)&d_off, (n + 1) * 8, false);
    if (rc == SQ_OK && n &&
        cudaMemcpyAsync(d_off, rec_off, n * 8, cudaMemcpyHostToDevice, sq_cur_stream(ctx)) != cudaSuccess)
        rc = sq_cuda_fail(cudaGetLastError(), "BAM offsets", __FILE__, __LINE__);
    if (rc == SQ_OK) rc = bam_decode_records(ctx, d_bam, nbytes, d_off, n, out, packed_len);  // (the chain is checked there)
    sq_dfree(ctx, d_bam);
    sq_dfree(ctx, d_off);
    return rc;
}

// The chain of the device copy: *d_off_out = offsets of the kept records (device), counters to the host.
static int bam_walk_device(sq_ctx *ctx, const uint8_t *d_bam, uint64_t nbytes, uint32_t n_ref, uint64_t **d_off_out,
                           uint64_t *n_kept, uint64_t *n_skipped, uint64_t *consumed) {
    cudaStream_t st = sq_cur_stream(ctx);
    char *ds = (char *)ctx->d_scratch, *hs = (char *)ctx->h_scratch;
    BamWalk *walk = (BamWalk *)(ds + 3776), *h_walk = (BamWalk *)(hs + 3776);
    uint32_t *d_kept_total = (uint32_t *)(ds + 3824);
    int *d_overflow = (int *)(ds + 3840), *h_overflow = (int *)(hs + 3840);
    const uint32_t n_words = (uint32_t)((nbytes + 31) / 32);
    uint32_t *bitmap = nullptr, *counts = nullptr, *rank = nullptr, *node_off = nullptr, *succ = nullptr, *jump_a = nullptr,
             *jump_b = nullptr, *keep = nullptr, *keep_rank = nullptr;
    uint8_t *dropped = nullptr, *mark = nullptr;
    uint64_t *d_off = nullptr;
    auto body = [&]() -> int {
        SQ_TRY(sq_dalloc(ctx, (void **)&bitmap, (size_t)n_words * 4, false));
        SQ_TRY(sq_dalloc(ctx, (void **)&counts, (size_t)n_words * 4, false));
        SQ_TRY(sq_dalloc(ctx, (void **)&rank, (size_t)n_words * 4, false));
        SQ_LAUNCH(ctx, k_bam_candidates, sq_grid_for(ctx, (uint64_t)n_words * 8, BAM_TPB, 16), BAM_TPB, 0, d_bam, nbytes,
                  n_ref, bitmap, counts, n_words);
        SQ_TRY(sq_scan_exclusive_u32(ctx, counts, rank, n_words, &walk->n_cand));
        CUDA_TRY(cudaMemcpyAsync(&h_walk->n_cand, &walk->n_cand, 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        const uint32_t m = h_walk->n_cand;
        SQ_TRY(sq_dalloc(ctx, (void **)&node_off, (size_t)(m + 1) * 4, false));
        SQ_TRY(sq_dalloc(ctx, (void **)&succ, (size_t)(m + 1) * 4, false));
        SQ_TRY(sq_dalloc(ctx, (void **)&jump_a, (size_t)(m + 1) * 4, false));
        SQ_TRY(sq_dalloc(ctx, (void **)&jump_b, (size_t)(m + 1) * 4, false));
        SQ_TRY(sq_dalloc(ctx, (void **)&keep, (size_t)(m + 1) * 4, false));
        SQ_TRY(sq_dalloc(ctx, (void **)&keep_rank, (size_t)(m + 1) * 4, false));
        SQ_TRY(sq_dalloc(ctx, (void **)&dropped, (size_t)m + 1, false));
        SQ_TRY(sq_dalloc(ctx, (void **)&mark, (size_t)m + 1, true));
        SQ_LAUNCH(ctx, k_bam_nodes, sq_grid_for(ctx, n_words, BAM_TPB, 16), BAM_TPB, 0, d_bam, nbytes, bitmap, rank, n_words,
                  node_off, succ, jump_a, dropped, walk);
        SQ_LAUNCH(ctx, k_bam_mark_head, 1, 1, 0, mark, walk);
        const int grid = sq_grid_for(ctx, (uint64_t)m + 1, BAM_TPB, 16);
        for (uint64_t reach = 1; reach < (uint64_t)m; reach *= 2) {  // after the round: everything within 2 * reach - 1 hops
            SQ_LAUNCH(ctx, k_bam_hop, grid, BAM_TPB, 0, mark, jump_a, jump_b, m);
            std::swap(jump_a, jump_b);
        }
        SQ_LAUNCH(ctx, k_bam_flags, grid, BAM_TPB, 0, mark, dropped, m, keep, walk);
        SQ_TRY(sq_scan_exclusive_u32(ctx, keep, keep_rank, m, d_kept_total));
        SQ_LAUNCH(ctx, k_bam_last, 1, 1, 0, d_bam, node_off, succ, d_kept_total, walk);
        CUDA_TRY(cudaMemcpyAsync(h_walk, walk, sizeof(BamWalk), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        const bool broken = h_walk->broken_at != ~0ULL;
        const uint64_t cap = broken ? nbytes / 36 + 1 : h_walk->n_kept;
        SQ_TRY(sq_dalloc(ctx, (void **)&d_off, (cap + 1) * 8, false));
        if (m) SQ_LAUNCH(ctx, k_bam_pick, grid, BAM_TPB, 0, keep, keep_rank, node_off, m, d_off);
        if (broken) {
            CUDA_TRY(cudaMemsetAsync(d_overflow, 0, 4, st));
            SQ_LAUNCH(ctx, k_bam_chase, 1, 1, 0, d_bam, nbytes, d_off, cap, walk, d_overflow);
            CUDA_TRY(cudaMemcpyAsync(h_walk, walk, sizeof(BamWalk), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(h_overflow, d_overflow, 4, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            if (*h_overflow) {
                sq_set_error("BAM records smaller than their fixed fields");
                return SQ_E_FORMAT;
            }
        }
        return SQ_OK;
    };
    int rc = body();
    for (void *p : {(void *)bitmap, (void *)counts, (void *)rank, (void *)node_off, (void *)succ, (void *)jump_a, (void *)jump_b,
                    (void *)keep, (void *)keep_rank, (void *)dropped, (void *)mark})
        sq_dfree(ctx, p);
    if (rc != SQ_OK) {
        sq_dfree(ctx, d_off);
        return rc;
    }
    *d_off_out = d_off;
    *n_kept = h_walk->n_kept;
    *n_skipped = h_walk->n_skipped;
    *consumed = h_walk->consumed;
    return SQ_OK;
}

extern "C" int sq_batch_from_bam_bytes(sq_ctx *ctx, const uint8_t *bam, uint64_t nbytes, int32_t n_ref, sq_batch **out,
                                       uint64_t *n_kept, uint64_t *n_skipped, uint64_t *consumed, uint64_t *packed_len) {
    *out = nullptr;
    *n_kept = *n_skipped = *consumed = *packed_len = 0;
    if (nbytes <= 4) return SQ_OK;
    SqParserScope on_parser_stream(ctx);
    uint8_t *d_bam = nullptr;
    uint64_t *d_off = nullptr;
    int rc = bam_upload(ctx, bam, nbytes, &d_bam);
    if (rc == SQ_OK) rc = bam_walk_device(ctx, d_bam, nbytes, n_ref < 0 ? 0u : (uint32_t)n_ref, &d_off, n_kept, n_skipped, consumed);
    if (rc == SQ_OK && *n_kept) rc = bam_decode_records(ctx, d_bam, nbytes, d_off, *n_kept, out, packed_len);
    sq_dfree(ctx, d_bam);
    sq_dfree(ctx, d_off);
    return rc;
}

// the chain alone (tests, profiling): offsets of the kept records copied back to rec_off[0..cap)
extern "C" int sq_bam_walk_device(sq_ctx *ctx, const uint8_t *bam, uint64_t nbytes, int32_t n_ref, uint64_t *rec_off,
                                  uint64_t cap, uint64_t *n_kept, uint64_t *n_skipped, uint64_t *consumed) {
    *n_kept = *n_skipped = *consumed = 0;
    if (nbytes <= 4) return SQ_OK;
    SqParserScope on_parser_stream(ctx);
    uint8_t *d_bam = nullptr;
    uint64_t *d_off = nullptr;
    int rc = bam_upload(ctx, bam, nbytes, &d_bam);
    if (rc == SQ_OK) rc = bam_walk_device(ctx, d_bam, nbytes, n_ref < 0 ? 0u : (uint32_t)n_ref, &d_off, n_kept, n_skipped, consumed);
    if (rc == SQ_OK && *n_kept > cap) {
        sq_set_error("BAM records smaller than their fixed fields");
        rc = SQ_E_FORMAT;
    }
    if (rc == SQ_OK && *n_kept) {
        if (cudaMemcpyAsync(rec_off, d_off, *n_kept * 8, cudaMemcpyDeviceToHost, sq_cur_stream(ctx)) != cudaSuccess ||
            cudaStreamSynchronize(sq_cur_stream(ctx)) != cudaSuccess)
            rc = sq_cuda_fail(cudaGetLastError(), "BAM offsets", __FILE__, __LINE__);
    }
    sq_dfree(ctx, d_bam);
    sq_dfree(ctx, d_off);
    return rc;
}
