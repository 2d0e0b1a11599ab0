// inflate_core.cuh -- DEFLATE (RFC 1951) decoder for one BGZF block, written for one decoding lane per
// warp with its Huffman tables in shared memory.  The same code compiles for the host, where the unit
// tests check it against zlib without a GPU (sq_selftest_inflate_host).
//
// in the reference, decompression is xopen's job on host threads
// (src/sequali/util.py:108-123, README.rst:168-171) and is the end-to-end bottleneck.  A BGZF file
// (bgzip'd FASTQ, every BAM) is a chain of independent gzip members of <= 64 KiB of text each, so the
// blocks inflate in parallel: one warp per block, thousands of blocks in flight.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define INF_HD __host__ __device__ __forceinline__
#else
#define INF_HD inline
#endif

constexpr int INF_LIT_BITS = 10;   // primary table of the literal/length code
constexpr int INF_DIST_BITS = 8;   // primary table of the distance code
constexpr int INF_MAXBITS = 15;

// Decode tables of one deflate block.  A primary entry packs (symbol << 4 | code length); length 0
// marks a code longer than the primary index, resolved by the canonical walk over count[] / symbol[].
struct InfTables {
    uint16_t lit[1 << INF_LIT_BITS];
    uint16_t dist[1 << INF_DIST_BITS];
    uint16_t lit_count[INF_MAXBITS + 1], dist_count[INF_MAXBITS + 1];
    uint16_t lit_sym[288], dist_sym[32];
};

enum { INF_OK = 0, INF_E_TRUNCATED = 1, INF_E_BLOCKTYPE = 2, INF_E_STORED = 3, INF_E_LENGTHS = 4, INF_E_CODE = 5,
       INF_E_DISTANCE = 6, INF_E_OVERFLOW = 7, INF_E_SIZE = 8 };

struct InfBits {
    const uint8_t *in;
    uint32_t len, pos;  // bytes
    uint64_t buf;
    uint32_t cnt;       // valid bits in buf
};

INF_HD void inf_refill(InfBits &b) {
    // at least 32 valid bits afterwards (while the input lasts): four bytes at a time, single bytes at the very end
    if (b.cnt > 32) return;
    if (b.pos + 4 <= b.len) {
        const uint8_t *p = b.in + b.pos;
        const uint32_t w = (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24;
        b.buf |= (uint64_t)w << b.cnt;
        b.cnt += 32;
        b.pos += 4;
        return;
    }
    while (b.cnt <= 56 && b.pos < b.len) {
        b.buf |= (uint64_t)b.in[b.pos++] << b.cnt;
        b.cnt += 8;
    }
}
INF_HD uint32_t inf_peek(const InfBits &b, uint32_t n) { return (uint32_t)(b.buf & ((1ULL << n) - 1)); }
INF_HD void inf_drop(InfBits &b, uint32_t n) {
    b.buf >>= n;
    b.cnt -= n;
}
INF_HD uint32_t inf_take(InfBits &b, uint32_t n) {
    const uint32_t v = inf_peek(b, n);
    inf_drop(b, n);
    return v;
}

INF_HD uint32_t inf_reverse(uint32_t code, uint32_t len) {
    uint32_t r = 0;
    for (uint32_t i = 0; i < len; i++) {
        r = (r << 1) | (code & 1);
        code >>= 1;
    }
    return r;
}

// canonical Huffman tables from code lengths; returns false for an over-subscribed set
INF_HD bool inf_build(const uint8_t *lengths, uint32_t n, uint16_t *primary, uint32_t primary_bits, uint16_t *count,
                      uint16_t *symbol) {
    for (int l = 0; l <= INF_MAXBITS; l++) count[l] = 0;
    for (uint32_t s = 0; s < n; s++) count[lengths[s]]++;
    int left = 1;
    for (int l = 1; l <= INF_MAXBITS; l++) {
        left <<= 1;
        left -= count[l];
        if (left < 0) return false;
    }
    uint16_t offs[INF_MAXBITS + 2];
    offs[1] = 0;
    for (int l = 1; l <= INF_MAXBITS; l++) offs[l + 1] = offs[l] + count[l];
    for (uint32_t s = 0; s < n; s++)
        if (lengths[s]) symbol[offs[lengths[s]]++] = (uint16_t)s;
    for (uint32_t i = 0; i < (1u << primary_bits); i++) primary[i] = 0;
    // assign codes in canonical order; fill the primary table for the short ones
    uint32_t code = 0, idx = 0;
    for (uint32_t l = 1; l <= INF_MAXBITS; l++) {
        for (uint32_t k = 0; k < count[l]; k++, code++, idx++) {
            if (l > primary_bits) continue;
            const uint32_t rev = inf_reverse(code, l);
            const uint16_t entry = (uint16_t)(symbol[idx] << 4 | l);
            for (uint32_t i = rev; i < (1u << primary_bits); i += 1u << l) primary[i] = entry;
        }
        code <<= 1;
    }
    count[0] = 0;
    return true;
}

// one symbol: primary lookup, or the canonical bit-by-bit walk for long codes; -1 on a bad code
INF_HD int inf_symbol(InfBits &b, const uint16_t *primary, uint32_t primary_bits, const uint16_t *count, const uint16_t *symbol) {
    const uint16_t e = primary[inf_peek(b, primary_bits)];
    if (e & 15) {
        if ((uint32_t)(e & 15) > b.cnt) return -1;
        inf_drop(b, e & 15);
        return e >> 4;
    }
    int code = 0, first = 0, index = 0;
    uint64_t bits = b.buf;
    for (int l = 1; l <= INF_MAXBITS; l++) {
        code |= (int)(bits & 1);
        bits >>= 1;
        const int c = count[l];
        if (code - c < first) {
            if ((uint32_t)l > b.cnt) return -1;
            inf_drop(b, (uint32_t)l);
            return symbol[index + (code - first)];
        }
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

// Inflates the raw deflate stream in[0 .. in_len) into at most out_cap bytes; *out_len = bytes produced.
// Out: where the text goes -- put(op, byte), raw(op, src, n) for stored blocks, match(op, dist, n) for LZ77
// copies.  The device build keeps a window of the text in shared memory and spreads every copy over the
// lanes of the warp; the host build writes a plain array.
template <typename Tables, typename Out>
INF_HD int inf_inflate(const uint8_t *in, uint32_t in_len, uint32_t out_cap, uint32_t *out_len, Tables &T, Out &out) {
    const uint16_t len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115,
                                   131, 163, 195, 227, 258};
    const uint8_t len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    const uint16_t dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537,
                                    2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    const uint8_t dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    const uint8_t cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    InfBits b;
    b.in = in;
    b.len = in_len;
    b.pos = 0;
    b.buf = 0;
    b.cnt = 0;
    uint32_t op = 0;
    *out_len = 0;
    for (;;) {
        inf_refill(b);
        if (b.cnt < 3) return INF_E_TRUNCATED;
        const uint32_t last = inf_take(b, 1), type = inf_take(b, 2);
        if (type == 0) {
            // stored: skip to the byte boundary, LEN / NLEN, raw bytes
            inf_drop(b, b.cnt & 7);
            inf_refill(b);
            if (b.cnt < 32) return INF_E_TRUNCATED;
            const uint32_t len = inf_take(b, 16), nlen = inf_take(b, 16);
            if ((len ^ 0xFFFFu) != nlen) return INF_E_STORED;
            // bytes still in the bit buffer were already consumed from `in`: step back
            const uint32_t start = b.pos - b.cnt / 8;
            if (start + len > in_len) return INF_E_TRUNCATED;
            if (op + len > out_cap) return INF_E_OVERFLOW;
            out.raw(op, in + start, len);
            op += len;
            b.pos = start + len;
            b.buf = 0;
            b.cnt = 0;
        }
        else if (type == 1 || type == 2) {
            uint8_t lengths[320];
            uint32_t nlen = 288, ndist = 30;
            if (type == 1) {
                for (int s = 0; s < 144; s++) lengths[s] = 8;
                for (int s = 144; s < 256; s++) lengths[s] = 9;
                for (int s = 256; s < 280; s++) lengths[s] = 7;
                for (int s = 280; s < 288; s++) lengths[s] = 8;
                for (int s = 0; s < 30; s++) lengths[288 + s] = 5;
            }
            else {
                inf_refill(b);
                if (b.cnt < 14) return INF_E_TRUNCATED;
                nlen = inf_take(b, 5) + 257;
                ndist = inf_take(b, 5) + 1;
                const uint32_t ncode = inf_take(b, 4) + 4;
                if (nlen > 286 || ndist > 30) return INF_E_LENGTHS;
                uint8_t cl[19];
                for (int i = 0; i < 19; i++) cl[i] = 0;
                for (uint32_t i = 0; i < ncode; i++) {
                    inf_refill(b);
                    if (b.cnt < 3) return INF_E_TRUNCATED;
                    cl[cl_order[i]] = (uint8_t)inf_take(b, 3);
                }
                // the code-length code reuses the distance table's storage (rebuilt right after)
                if (!inf_build(cl, 19, T.dist, 7, T.dist_count, T.dist_sym)) return INF_E_LENGTHS;
                uint32_t idx = 0;
                while (idx < nlen + ndist) {
                    inf_refill(b);
                    const int sym = inf_symbol(b, T.dist, 7, T.dist_count, T.dist_sym);
                    if (sym < 0) return b.pos >= b.len ? INF_E_TRUNCATED : INF_E_CODE;
                    if (sym < 16) lengths[idx++] = (uint8_t)sym;
                    else {
                        uint32_t rep, val = 0;
                        if (b.cnt < 7) return INF_E_TRUNCATED;
                        if (sym == 16) {
                            if (idx == 0) return INF_E_LENGTHS;
                            val = lengths[idx - 1];
                            rep = 3 + inf_take(b, 2);
                        }
                        else if (sym == 17) rep = 3 + inf_take(b, 3);
                        else rep = 11 + inf_take(b, 7);
                        if (idx + rep > nlen + ndist) return INF_E_LENGTHS;
                        while (rep--) lengths[idx++] = (uint8_t)val;
                    }
                }
                if (lengths[256] == 0) return INF_E_LENGTHS;  // no end-of-block code
            }
            if (!inf_build(lengths, nlen, T.lit, INF_LIT_BITS, T.lit_count, T.lit_sym)) return INF_E_LENGTHS;
            uint8_t dl[32];
            for (uint32_t s = 0; s < 32; s++) dl[s] = s < ndist ? lengths[nlen + s] : 0;
            if (!inf_build(dl, ndist, T.dist, INF_DIST_BITS, T.dist_count, T.dist_sym)) return INF_E_LENGTHS;
            for (;;) {
                inf_refill(b);
                int sym = inf_symbol(b, T.lit, INF_LIT_BITS, T.lit_count, T.lit_sym);
                if (sym < 0) return b.pos >= b.len ? INF_E_TRUNCATED : INF_E_CODE;
                if (sym < 256) {
                    if (op >= out_cap) return INF_E_OVERFLOW;
                    out.put(op++, (uint8_t)sym);
                    continue;
                }
                if (sym == 256) break;
                sym -= 257;
                if (sym >= 29) return INF_E_CODE;
                if (b.cnt < len_extra[sym]) return INF_E_TRUNCATED;
                const uint32_t len = len_base[sym] + inf_take(b, len_extra[sym]);
                inf_refill(b);
                const int ds = inf_symbol(b, T.dist, INF_DIST_BITS, T.dist_count, T.dist_sym);
                if (ds < 0) return b.pos >= b.len ? INF_E_TRUNCATED : INF_E_CODE;
                if (ds >= 30) return INF_E_CODE;
                if (b.cnt < dist_extra[ds]) return INF_E_TRUNCATED;
                const uint32_t dist = dist_base[ds] + inf_take(b, dist_extra[ds]);
                if (dist > op) return INF_E_DISTANCE;
                if (op + len > out_cap) return INF_E_OVERFLOW;
                out.match(op, dist, len);
                op += len;
            }
        }
        else return INF_E_BLOCKTYPE;
        if (last) break;
    }
    *out_len = op;
    return INF_OK;
}
