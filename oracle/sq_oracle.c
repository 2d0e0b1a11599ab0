/*
 * sq_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Scalar CPU restatement of the per-record QC hot path of rhpvorderman/sequali
 * (reference commit a82688e, src/sequali/_qcmodule.c).  It exists so that the
 * CUDA path can be checked bit-for-bit on any machine (the GPU box has no
 * /root/reference).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (sequali_b200) never does.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every
 * function below against the reference's own compiled extension
 * (oracle/_ref, built by oracle/build_ref.sh from /root/reference) and
 * against the golden vectors committed under tests/golden/.
 *
 * Everything here is written from the behaviour of the reference, not from
 * its text; each function cites the reference lines it follows.
 *
 * Records are described by `orc_rec` (offsets into one byte buffer), the
 * equivalent of the reference's FastqMeta (_qcmodule.c:337-355).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

typedef struct {
    uint64_t name_off, seq_off, qual_off, tags_off;
    uint32_t name_len, seq_len, tags_len, pad;
    double err_sum; /* accumulated_error_rate, written by orc_qc_add */
} orc_rec;

/* ------------------------------------------------------------------ */
/* constants                                                           */
/* ------------------------------------------------------------------ */

#define PHRED_MAX 93
static double ERR[PHRED_MAX + 1]; /* 10^-(q/10): score_to_error_rate.h:4-99 */
static int ERR_ready = 0;

static void
err_init(void)
{
    if (ERR_ready) return;
    for (int q = 0; q <= PHRED_MAX; q++) ERR[q] = pow(10.0, -((double)q / 10.0));
    ERR_ready = 1;
}

ORC_API void
orc_error_table(double *out94)
{
    err_init();
    memcpy(out94, ERR, sizeof(ERR));
}

/* counting / adapter alphabet: ACGT (either case) -> 0..3, anything else 4
 * (_qcmodule.c:1748-1763) */
static inline int
nuc5(uint8_t c)
{
    switch (c | 0x20) {
        case 'a': return 0;
        case 'c': return 1;
        case 'g': return 2;
        case 't': return 3;
    }
    return 4;
}

/* ------------------------------------------------------------------ */
/* FASTQ record boundary scan  (_qcmodule.c:1093-1171)                 */
/* ------------------------------------------------------------------ */

enum {
    ORC_OK = 0,
    ORC_E_NO_AT = 1,   /* "Record does not start with @ but with %c"      */
    ORC_E_NO_PLUS = 2, /* "Record second header does not start with + ..." */
    ORC_E_LEN = 3,     /* "Record sequence and qualities do not have equal length" */
    ORC_E_ASCII = 4,   /* "Found non-ASCII character in file: %c"          */
    ORC_E_PHRED = 5,   /* "Not a valid phred character: %c"                */
};

/* Parses complete records from buf[0..n).  Returns number of records written
 * (<= cap, <= max_records); *consumed = offset just past the last parsed
 * record; on a format error returns -1 and sets err_code / err_pos (offset of
 * the offending byte, or of the record's name for ORC_E_LEN). */
ORC_API int64_t
orc_parse_fastq(const uint8_t *buf, uint64_t n, uint64_t max_records,
                orc_rec *out, uint64_t cap, uint64_t *consumed, int *err_code,
                uint64_t *err_pos)
{
    uint64_t pos = 0, count = 0;
    *err_code = ORC_OK;
    *err_pos = 0;
    while (count < max_records && count < cap) {
        if (pos + 2 >= n) break; /* :1094 */
        if (buf[pos] != '@') {
            *err_code = ORC_E_NO_AT;
            *err_pos = pos;
            return -1;
        }
        const uint8_t *e1 = memchr(buf + pos + 1, '\n', n - pos - 1);
        if (!e1) break;
        uint64_t seq = (uint64_t)(e1 - buf) + 1;
        const uint8_t *e2 = memchr(buf + seq, '\n', n - seq);
        if (!e2) break;
        uint64_t plus = (uint64_t)(e2 - buf) + 1;
        if (plus < n && buf[plus] != '+') { /* :1119 */
            *err_code = ORC_E_NO_PLUS;
            *err_pos = plus;
            return -1;
        }
        const uint8_t *e3 = memchr(buf + plus, '\n', n - plus);
        if (!e3) break;
        uint64_t qual = (uint64_t)(e3 - buf) + 1;
        const uint8_t *e4 = memchr(buf + qual, '\n', n - qual);
        if (!e4) break;
        uint64_t qend = (uint64_t)(e4 - buf);
        if (plus - 1 - seq != qend - qual) { /* :1140 */
            *err_code = ORC_E_LEN;
            *err_pos = pos + 1;
            return -1;
        }
        orc_rec *r = out + count++;
        r->name_off = pos + 1;
        r->name_len = (uint32_t)(seq - 1 - (pos + 1));
        r->seq_off = seq;
        r->seq_len = (uint32_t)(plus - 1 - seq);
        r->qual_off = qual;
        r->tags_off = qend;
        r->tags_len = 0;
        r->pad = 0;
        r->err_sum = 0.0;
        pos = qend + 1;
    }
    *consumed = pos;
    return (int64_t)count;
}

/* first byte >= 0x80, or n if none (_qcmodule.c:1055-1067) */
ORC_API uint64_t
orc_first_non_ascii(const uint8_t *buf, uint64_t n)
{
    for (uint64_t i = 0; i < n; i++)
        if (buf[i] & 0x80) return i;
    return n;
}

/* ------------------------------------------------------------------ */
/* BAM record decode (_qcmodule.c:1623-1694)                           */
/* ------------------------------------------------------------------ */

static inline uint32_t
rd32(const uint8_t *p)
{
    return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 |
           (uint32_t)p[3] << 24;
}
static inline uint32_t
rd16(const uint8_t *p)
{
    return (uint32_t)p[0] | (uint32_t)p[1] << 8;
}

/* Decodes every complete alignment record in bam[0..n) (the stream after the
 * BAM header) into the packed name|seq|qual|tags layout.  Secondary (0x100)
 * and supplementary (0x800) records are skipped.  Returns the number of
 * records emitted; *consumed = bytes of input used, *skipped = skipped
 * records, *packed_len = bytes written to `packed` (capacity must be at least
 * (n*4+2)/3, the reference's bound, :1590). */
ORC_API int64_t
orc_decode_bam(const uint8_t *bam, uint64_t n, uint8_t *packed, orc_rec *out,
               uint64_t cap, uint64_t *consumed, uint64_t *skipped,
               uint64_t *packed_len)
{
    static const char code[] = "=ACMGRSVTWYHKDBN";
    uint64_t pos = 0, w = 0, count = 0, skip = 0;
    while (count < cap) {
        if (pos + 4 >= n) break; /* :1624 */
        uint64_t block = rd32(bam + pos);
        uint64_t end = pos + 4 + block;
        if (end > n) break;
        const uint8_t *h = bam + pos;
        uint32_t l_name = h[12], n_cigar = rd16(h + 16), flag = rd16(h + 18);
        uint32_t l_seq = rd32(h + 20);
        if (flag & (0x100 | 0x800)) {
            pos = end;
            skip++;
            continue;
        }
        const uint8_t *name = h + 36;
        const uint8_t *seq = name + l_name + 4 * (uint64_t)n_cigar;
        const uint8_t *qual = seq + (l_seq + 1) / 2;
        const uint8_t *tags = qual + l_seq;
        uint64_t tags_len = (uint64_t)((bam + end) - tags);
        uint32_t nl = l_name ? l_name - 1 : 0; /* drop the NUL */
        orc_rec *r = out + count++;
        r->name_off = w;
        r->name_len = nl;
        memcpy(packed + w, name, nl);
        w += nl;
        r->seq_off = w;
        r->seq_len = l_seq;
        for (uint32_t i = 0; i < l_seq; i++) {
            uint8_t b = seq[i >> 1];
            packed[w + i] = (uint8_t)code[(i & 1) ? (b & 15) : (b >> 4)];
        }
        w += l_seq;
        r->qual_off = w;
        if (l_seq && qual[0] == 0xff) /* :1658 */
            memset(packed + w, '!', l_seq);
        else
            for (uint32_t i = 0; i < l_seq; i++) packed[w + i] = (uint8_t)(qual[i] + 33);
        w += l_seq;
        r->tags_off = w;
        r->tags_len = (uint32_t)tags_len;
        memcpy(packed + w, tags, tags_len);
        w += tags_len;
        r->pad = 0;
        r->err_sum = 0.0;
        pos = end;
    }
    *consumed = pos;
    *skipped = skip;
    *packed_len = w;
    return (int64_t)count;
}

/* ------------------------------------------------------------------ */
/* QCMetrics (_qcmodule.c:1966-2139)                                   */
/* ------------------------------------------------------------------ */

typedef struct {
    uint64_t end_anchor, max_len, n_reads;
    uint64_t *base;  /* [max_len][5]  */
    uint64_t *phred; /* [max_len][12] */
    uint64_t *ea_base, *ea_phred;
    uint64_t gc[101], mean_phred[PHRED_MAX + 1];
} orc_qc;

ORC_API orc_qc *
orc_qc_new(uint64_t end_anchor)
{
    err_init();
    orc_qc *m = calloc(1, sizeof(*m));
    m->end_anchor = end_anchor;
    m->ea_base = calloc(end_anchor ? end_anchor * 5 : 1, 8);
    m->ea_phred = calloc(end_anchor ? end_anchor * 12 : 1, 8);
    return m;
}
ORC_API void
orc_qc_free(orc_qc *m)
{
    if (!m) return;
    free(m->base);
    free(m->phred);
    free(m->ea_base);
    free(m->ea_phred);
    free(m);
}

static void
qc_grow(orc_qc *m, uint64_t len)
{
    if (len <= m->max_len) return;
    m->base = realloc(m->base, len * 5 * 8);
    m->phred = realloc(m->phred, len * 12 * 8);
    memset(m->base + m->max_len * 5, 0, (len - m->max_len) * 5 * 8);
    memset(m->phred + m->max_len * 12, 0, (len - m->max_len) * 12 * 8);
    m->max_len = len;
}

static inline int
phred_bin(uint8_t q)
{
    return (q > 47 ? 47 : q) >> 2; /* :1778 */
}

/* Per-read error sum in the reference's evaluation order (:2059-2112): four
 * interleaved chains while more than four values remain, combined
 * ((a0+a1)+a2)+a3, then the tail in order.  Returns -1 and the offending byte
 * when a quality is outside '!'..'~'. */
static int
ordered_error_sum(const uint8_t *q, uint64_t L, double *sum, uint8_t *bad,
                  uint64_t *n_ok)
{
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    uint64_t i = 0;
    while (i + 4 < L) {
        uint8_t q0 = q[i] - 33, q1 = q[i + 1] - 33, q2 = q[i + 2] - 33, q3 = q[i + 3] - 33;
        if (q0 > PHRED_MAX || q1 > PHRED_MAX || q2 > PHRED_MAX || q3 > PHRED_MAX) break;
        a0 += ERR[q0];
        a1 += ERR[q1];
        a2 += ERR[q2];
        a3 += ERR[q3];
        i += 4;
    }
    double s = ((a0 + a1) + a2) + a3;
    for (; i < L; i++) {
        uint8_t v = q[i] - 33;
        if (v > PHRED_MAX) {
            *bad = q[i];
            *n_ok = i;
            *sum = s;
            return -1;
        }
        s += ERR[v];
    }
    *sum = s;
    *n_ok = L;
    return 0;
}

/* returns 0, or -1 with bad_char / bad_rec set (processing stops there, like
 * the reference, with earlier side effects kept) */
ORC_API int
orc_qc_add(orc_qc *m, const uint8_t *buf, orc_rec *recs, uint64_t n,
           uint8_t *bad_char, uint64_t *bad_rec)
{
    for (uint64_t r = 0; r < n; r++) {
        orc_rec *rec = recs + r;
        const uint8_t *s = buf + rec->seq_off, *q = buf + rec->qual_off;
        uint64_t L = rec->seq_len;
        qc_grow(m, L);
        m->n_reads++;
        uint64_t ea = L < m->end_anchor ? L : m->end_anchor;
        uint64_t ea_row0 = m->end_anchor - ea;
        uint64_t at = 0, gc = 0;
        for (uint64_t i = 0; i < L; i++) {
            int k = nuc5(s[i]);
            m->base[i * 5 + k]++;
            if (k == 0 || k == 3) at++;
            if (k == 1 || k == 2) gc++;
            if (i + ea >= L) m->ea_base[(ea_row0 + (i - (L - ea))) * 5 + k]++;
        }
        if (at + gc) { /* :2045-2058 */
            double pct = (double)gc * 100.0 / (double)(at + gc);
            m->gc[(uint64_t)round(pct)]++;
        }
        double sum;
        uint8_t bad = 0;
        uint64_t n_ok;
        int rc = ordered_error_sum(q, L, &sum, &bad, &n_ok);
        for (uint64_t i = 0; i < n_ok; i++) m->phred[i * 12 + phred_bin(q[i] - 33)]++;
        if (rc) {
            *bad_char = bad;
            *bad_rec = r;
            return -1;
        }
        for (uint64_t i = L - ea; i < L; i++)
            m->ea_phred[(ea_row0 + (i - (L - ea))) * 12 + phred_bin(q[i] - 33)]++;
        rec->err_sum = sum;
        if (L) { /* :2127-2137 */
            double avg = sum / (double)L;
            double ph = -10.0 * log10(avg);
            m->mean_phred[(uint64_t)floor(ph)]++;
        }
    }
    return 0;
}

ORC_API void
orc_qc_info(const orc_qc *m, uint64_t *max_len, uint64_t *n_reads)
{
    *max_len = m->max_len;
    *n_reads = m->n_reads;
}
ORC_API void
orc_qc_tables(const orc_qc *m, uint64_t *base, uint64_t *phred, uint64_t *ea_base,
              uint64_t *ea_phred, uint64_t *gc, uint64_t *mean_phred)
{
    if (base) memcpy(base, m->base, m->max_len * 5 * 8);
    if (phred) memcpy(phred, m->phred, m->max_len * 12 * 8);
    if (ea_base) memcpy(ea_base, m->ea_base, m->end_anchor * 5 * 8);
    if (ea_phred) memcpy(ea_phred, m->ea_phred, m->end_anchor * 12 * 8);
    if (gc) memcpy(gc, m->gc, sizeof(m->gc));
    if (mean_phred) memcpy(mean_phred, m->mean_phred, sizeof(m->mean_phred));
}

/* ------------------------------------------------------------------ */
/* AdapterCounter (_qcmodule.c:2644-2823)                              */
/* The observable result of the reference's packed shift-AND automaton */
/* is: for each adapter, the first position where it occurs in the     */
/* read (alphabet classes as nuc5: an adapter letter outside ACGT      */
/* matches any read letter outside ACGT).                              */
/* ------------------------------------------------------------------ */

typedef struct {
    uint64_t n_adapters, max_len, n_seqs;
    uint8_t **pat; /* nuc5 classes */
    uint32_t *plen;
    uint64_t **fwd, **rev;
} orc_ad;

ORC_API orc_ad *
orc_ad_new(const char *const *adapters, uint64_t n)
{
    orc_ad *a = calloc(1, sizeof(*a));
    a->n_adapters = n;
    a->pat = calloc(n, sizeof(*a->pat));
    a->plen = calloc(n, sizeof(*a->plen));
    a->fwd = calloc(n, sizeof(*a->fwd));
    a->rev = calloc(n, sizeof(*a->rev));
    for (uint64_t i = 0; i < n; i++) {
        a->plen[i] = (uint32_t)strlen(adapters[i]);
        a->pat[i] = malloc(a->plen[i] + 1);
        for (uint32_t j = 0; j < a->plen[i]; j++)
            a->pat[i][j] = (uint8_t)nuc5((uint8_t)adapters[i][j]);
    }
    return a;
}
ORC_API void
orc_ad_free(orc_ad *a)
{
    if (!a) return;
    for (uint64_t i = 0; i < a->n_adapters; i++) {
        free(a->pat[i]);
        free(a->fwd[i]);
        free(a->rev[i]);
    }
    free(a->pat);
    free(a->plen);
    free(a->fwd);
    free(a->rev);
    free(a);
}
ORC_API void
orc_ad_add(orc_ad *a, const uint8_t *buf, const orc_rec *recs, uint64_t n)
{
    for (uint64_t r = 0; r < n; r++) {
        const uint8_t *s = buf + recs[r].seq_off;
        uint64_t L = recs[r].seq_len;
        a->n_seqs++;
        if (L > a->max_len) {
            for (uint64_t i = 0; i < a->n_adapters; i++) {
                a->fwd[i] = realloc(a->fwd[i], L * 8);
                a->rev[i] = realloc(a->rev[i], L * 8);
                memset(a->fwd[i] + a->max_len, 0, (L - a->max_len) * 8);
                memset(a->rev[i] + a->max_len, 0, (L - a->max_len) * 8);
            }
            a->max_len = L;
        }
        for (uint64_t i = 0; i < a->n_adapters; i++) {
            uint32_t pl = a->plen[i];
            if (pl == 0 || pl > L) continue;
            for (uint64_t p = 0; p + pl <= L; p++) {
                uint32_t j = 0;
                while (j < pl && nuc5(s[p + j]) == a->pat[i][j]) j++;
                if (j == pl) {
                    a->fwd[i][p]++;
                    a->rev[i][L - 1 - p]++;
                    break;
                }
            }
        }
    }
}
ORC_API void
orc_ad_info(const orc_ad *a, uint64_t *max_len, uint64_t *n_seqs)
{
    *max_len = a->max_len;
    *n_seqs = a->n_seqs;
}
ORC_API void
orc_ad_counts(const orc_ad *a, uint64_t idx, uint64_t *fwd, uint64_t *rev)
{
    if (a->max_len == 0) return;
    memcpy(fwd, a->fwd[idx], a->max_len * 8);
    memcpy(rev, a->rev[idx], a->max_len * 8);
}

/* ------------------------------------------------------------------ */
/* PerTileQuality (_qcmodule.c:3089-3222, 3307-3359)                   */
/* ------------------------------------------------------------------ */

typedef struct {
    uint64_t tile;
    uint64_t *len_counts;
    double *err;
} orc_tile;

typedef struct {
    uint64_t max_len, n_reads, n_tiles, cap;
    orc_tile *tiles; /* unsorted; sorted at read-out */
    int skipped;
    uint64_t skipped_rec; /* index (within the add call) of the unparsable header */
} orc_ptq;

/* decimal tile id between the 4th and 5th ':' of the name, 1..18 digits;
 * -1 otherwise (:3089-3121, :160-180) */
static int64_t
tile_id(const uint8_t *h, uint64_t n)
{
    uint64_t i = 0, colons = 0;
    for (; i < n; i++)
        if (h[i] == ':' && ++colons == 4) break;
    uint64_t start = i + 1, j = start;
    for (; j < n; j++)
        if (h[j] == ':') break;
    if (j >= n) return -1;
    uint64_t len = j - start;
    if (len < 1 || len > 18) return -1;
    int64_t v = 0;
    for (uint64_t k = start; k < j; k++) {
        uint8_t d = h[k] - '0';
        if (d > 9) return -1;
        v = v * 10 + d;
    }
    return v;
}

ORC_API orc_ptq *
orc_ptq_new(void)
{
    err_init();
    return calloc(1, sizeof(orc_ptq));
}
ORC_API void
orc_ptq_free(orc_ptq *p)
{
    if (!p) return;
    for (uint64_t i = 0; i < p->n_tiles; i++) {
        free(p->tiles[i].len_counts);
        free(p->tiles[i].err);
    }
    free(p->tiles);
    free(p);
}
static orc_tile *
ptq_tile(orc_ptq *p, uint64_t id)
{
    for (uint64_t i = 0; i < p->n_tiles; i++)
        if (p->tiles[i].tile == id) return p->tiles + i;
    if (p->n_tiles == p->cap) {
        p->cap = p->cap ? p->cap * 2 : 64;
        p->tiles = realloc(p->tiles, p->cap * sizeof(orc_tile));
    }
    orc_tile *t = p->tiles + p->n_tiles++;
    t->tile = id;
    t->len_counts = calloc(p->max_len ? p->max_len : 1, 8);
    t->err = calloc(p->max_len ? p->max_len : 1, 8);
    return t;
}
/* returns 0; 1 if this call hit an unparsable header (module now skipped,
 * *skip_rec = index of that record); -1 on an invalid phred (bad_char set). */
ORC_API int
orc_ptq_add(orc_ptq *p, const uint8_t *buf, const orc_rec *recs, uint64_t n,
            uint64_t *skip_rec, uint8_t *bad_char)
{
    if (p->skipped) return 0;
    for (uint64_t r = 0; r < n; r++) {
        const orc_rec *rec = recs + r;
        int64_t id = tile_id(buf + rec->name_off, rec->name_len);
        if (id < 0) {
            p->skipped = 1;
            *skip_rec = r;
            return 1;
        }
        uint64_t L = rec->seq_len;
        if (L > p->max_len) {
            for (uint64_t i = 0; i < p->n_tiles; i++) {
                orc_tile *t = p->tiles + i;
                t->len_counts = realloc(t->len_counts, L * 8);
                t->err = realloc(t->err, L * 8);
                memset(t->len_counts + p->max_len, 0, (L - p->max_len) * 8);
                memset(t->err + p->max_len, 0, (L - p->max_len) * 8);
            }
            p->max_len = L;
        }
        orc_tile *t = ptq_tile(p, (uint64_t)id);
        p->n_reads++;
        if (L == 0) continue;
        t->len_counts[L - 1]++;
        const uint8_t *q = buf + rec->qual_off;
        /* plain per-position chain across reads (:3189-3220); groups of four
         * are validated together before any of the four is added */
        uint64_t i = 0;
        while (i + 3 < L) {
            uint8_t q0 = q[i] - 33, q1 = q[i + 1] - 33, q2 = q[i + 2] - 33, q3 = q[i + 3] - 33;
            if (q0 > PHRED_MAX || q1 > PHRED_MAX || q2 > PHRED_MAX || q3 > PHRED_MAX) break;
            t->err[i] += ERR[q0];
            t->err[i + 1] += ERR[q1];
            t->err[i + 2] += ERR[q2];
            t->err[i + 3] += ERR[q3];
            i += 4;
        }
        for (; i < L; i++) {
            uint8_t v = q[i] - 33;
            if (v > PHRED_MAX) {
                *bad_char = q[i];
                return -1;
            }
            t->err[i] += ERR[v];
        }
    }
    return 0;
}
ORC_API void
orc_ptq_info(const orc_ptq *p, uint64_t *max_len, uint64_t *n_reads,
             uint64_t *n_tiles, int *skipped)
{
    *max_len = p->max_len;
    *n_reads = p->n_reads;
    *n_tiles = p->n_tiles;
    *skipped = p->skipped;
}
static int
tile_cmp(const void *a, const void *b)
{
    uint64_t x = ((const orc_tile *)a)->tile, y = ((const orc_tile *)b)->tile;
    return x < y ? -1 : x > y;
}
/* tiles ascending; counts[j] = number of reads with length > j (:3336-3347) */
ORC_API void
orc_ptq_tables(orc_ptq *p, uint64_t *tile_ids, double *err, uint64_t *counts)
{
    qsort(p->tiles, p->n_tiles, sizeof(orc_tile), tile_cmp);
    for (uint64_t i = 0; i < p->n_tiles; i++) {
        tile_ids[i] = p->tiles[i].tile;
        uint64_t run = 0;
        for (uint64_t j = p->max_len; j-- > 0;) {
            run += p->tiles[i].len_counts[j];
            counts[i * p->max_len + j] = run;
            err[i * p->max_len + j] = p->tiles[i].err[j];
        }
    }
}

/* ------------------------------------------------------------------ */
/* hashes: Thomas Wang 64-bit mix (wanghash.h:14-25) and               */
/* MurmurHash3 x64 128, second half (murmur3.h:49-156)                 */
/* ------------------------------------------------------------------ */

static inline uint64_t
wang64(uint64_t k)
{
    k = ~k + (k << 21);
    k ^= k >> 24;
    k *= 265;
    k ^= k >> 14;
    k *= 21;
    k ^= k >> 28;
    k += k << 31;
    return k;
}

static inline uint64_t
unxorshift(uint64_t v, int s)
{
    uint64_t x = v;
    for (int i = s; i < 64; i += s) x = v ^ (x >> s);
    return x;
}

ORC_API uint64_t
orc_wang64_inverse(uint64_t k)
{
    /* k += k<<31  => multiply by (1 + 2^31); invert with modular inverse */
    k *= 0x3fffffff80000001ULL; /* inverse of 2^31+1 mod 2^64 */
    k = unxorshift(k, 28);
    k *= 14933078535860113213ULL; /* 21^-1 */
    k = unxorshift(k, 14);
    k *= 15244667743933553977ULL; /* 265^-1 */
    k = unxorshift(k, 24);
    /* k = ~x + (x<<21) = x*(2^21 - 1) - 1  => x = (k+1) * (2^21-1)^-1 */
    return (k + 1) * 0x7ffffbffffdfffffULL;
}

static inline uint64_t
rotl64(uint64_t x, int r)
{
    return (x << r) | (x >> (64 - r));
}
static inline uint64_t
fmix64(uint64_t k)
{
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}
static uint64_t
murmur3_h2(const uint8_t *d, uint64_t len, uint64_t seed)
{
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    uint64_t h1 = seed, h2 = seed;
    uint64_t nb = len / 16;
    for (uint64_t i = 0; i < nb; i++) {
        uint64_t k1, k2;
        memcpy(&k1, d + 16 * i, 8);
        memcpy(&k2, d + 16 * i + 8, 8);
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    const uint8_t *t = d + nb * 16;
    uint64_t rem = len & 15, k1 = 0, k2 = 0;
    for (uint64_t i = rem; i > 8; i--) k2 ^= (uint64_t)t[i - 1] << (8 * (i - 9));
    if (rem > 8) { k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2; }
    for (uint64_t i = rem < 8 ? rem : 8; i > 0; i--) k1 ^= (uint64_t)t[i - 1] << (8 * (i - 1));
    if (rem > 0) { k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1; }
    h1 ^= len; h2 ^= len;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2; h2 += h1;
    return h2;
}
ORC_API uint64_t
orc_murmur3(const uint8_t *d, uint64_t len, uint64_t seed)
{
    return murmur3_h2(d, len, seed);
}
ORC_API uint64_t
orc_wang64(uint64_t k)
{
    return wang64(k);
}

/* ------------------------------------------------------------------ */
/* OverrepresentedSequences (_qcmodule.c:3543-3568, 3589-3608,         */
/* 3635-3694, 3830-3942)                                               */
/* ------------------------------------------------------------------ */

typedef struct {
    uint64_t k, sample_every, max_unique, frags_front, frags_back;
    uint64_t n_seqs, n_sampled, n_unique, total_frags;
    uint64_t table_size; /* power of two */
    uint64_t *keys;      /* wang hash of the canonical k-mer, 0 = empty */
    uint32_t *counts;
    uint64_t *stage;
    uint64_t stage_cap;
} orc_ov;

ORC_API orc_ov *
orc_ov_new(uint64_t max_unique, uint64_t k, uint64_t sample_every,
           int64_t bases_front, int64_t bases_back)
{
    orc_ov *o = calloc(1, sizeof(*o));
    o->k = k;
    o->sample_every = sample_every;
    o->max_unique = max_unique;
    if (bases_front < 0) bases_front = UINT32_MAX; /* :3499-3504 */
    if (bases_back < 0) bases_back = UINT32_MAX;
    o->frags_front = ((uint64_t)bases_front + k - 1) / k;
    o->frags_back = ((uint64_t)bases_back + k - 1) / k;
    uint64_t bits = (uint64_t)(log2((double)max_unique * 1.5) + 1); /* :3508 */
    o->table_size = 1ULL << bits;
    o->keys = calloc(o->table_size, 8);
    o->counts = calloc(o->table_size, 4);
    return o;
}
ORC_API void
orc_ov_free(orc_ov *o)
{
    if (!o) return;
    free(o->keys);
    free(o->counts);
    free(o->stage);
    free(o);
}

/* k-mer alphabet: ACGT -> 0..3, N/n -> "ignore fragment", other -> "ignore
 * and warn" (:3612-3627).  Returns 0 ok, 1 contains N only, 2 contains an
 * unknown letter. */
static int
canonical_kmer(const uint8_t *s, uint64_t k, uint64_t *out)
{
    uint64_t fw = 0, rc = 0;
    int has_n = 0, has_unknown = 0;
    for (uint64_t i = 0; i < k; i++) {
        int c = nuc5(s[i]);
        if (c == 4) {
            if ((s[i] | 0x20) == 'n') has_n = 1;
            else has_unknown = 1;
            c = 0;
        }
        fw = (fw << 2) | (uint64_t)c;
        rc |= (uint64_t)(3 - c) << (2 * i);
    }
    if (has_unknown) return 2;
    if (has_n) return 1;
    *out = rc < fw ? rc : fw;
    return 0;
}

static void
ov_main_insert(orc_ov *o, uint64_t h)
{
    uint64_t mask = o->table_size - 1, i = h & mask;
    for (;;) {
        if (o->keys[i] == 0) {
            if (o->n_unique < o->max_unique) {
                o->keys[i] = h;
                o->counts[i] = 1;
                o->n_unique++;
            }
            return;
        }
        if (o->keys[i] == h) {
            o->counts[i]++;
            return;
        }
        i = (i + 1) & mask;
    }
}

/* *warn_rec receives the index of the first record that held a non-ACGTN
 * letter in a sampled fragment (or n if none); returns number of such records */
ORC_API uint64_t
orc_ov_add(orc_ov *o, const uint8_t *buf, const orc_rec *recs, uint64_t n,
           uint64_t *warn_rec)
{
    uint64_t warned = 0;
    *warn_rec = n;
    for (uint64_t r = 0; r < n; r++) {
        uint64_t idx = o->n_seqs++;
        if (idx % o->sample_every) continue;
        o->n_sampled++;
        uint64_t L = recs[r].seq_len, k = o->k;
        if (L < k) continue;
        const uint8_t *s = buf + recs[r].seq_off;
        uint64_t maxf = (L + k - 1) / k;
        uint64_t back_cap = maxf / 2, front_cap = maxf - back_cap;
        uint64_t nf = o->frags_front < front_cap ? o->frags_front : front_cap;
        uint64_t nb = o->frags_back < back_cap ? o->frags_back : back_cap;
        uint64_t total = nf + nb;
        if (total == 0) continue; /* reference would take log2(0); unreachable with defaults */
        uint64_t bits = (uint64_t)ceil(log2((double)total * 1.5)); /* :3884 */
        uint64_t ssize = 1ULL << bits, smask = ssize - 1;
        if (ssize > o->stage_cap) {
            o->stage = realloc(o->stage, ssize * 8);
            o->stage_cap = ssize;
        }
        memset(o->stage, 0, ssize * 8);
        uint64_t valid = 0;
        int warn = 0;
        for (uint64_t f = 0; f < total; f++) {
            uint64_t off = f < nf ? f * k : L - (nb - (f - nf)) * k;
            uint64_t kmer;
            int rc = canonical_kmer(s + off, k, &kmer);
            if (rc) {
                if (rc == 2) warn = 1;
                continue;
            }
            valid++;
            uint64_t h = wang64(kmer), i = h & smask;
            while (o->stage[i] != 0 && o->stage[i] != h) i = (i + 1) & smask;
            o->stage[i] = h;
        }
        for (uint64_t i = 0; i < ssize; i++)
            if (o->stage[i]) ov_main_insert(o, o->stage[i]);
        if (warn) {
            if (!warned) *warn_rec = r;
            warned++;
        }
        o->total_frags += valid;
    }
    return warned;
}
ORC_API void
orc_ov_info(const orc_ov *o, uint64_t *out6)
{
    out6[0] = o->n_seqs;
    out6[1] = o->n_sampled;
    out6[2] = o->n_unique;
    out6[3] = o->total_frags;
    out6[4] = o->table_size;
    out6[5] = o->max_unique;
}
/* writes n_unique (kmer, count) pairs in slot order; kmer = inverse hash */
ORC_API uint64_t
orc_ov_entries(const orc_ov *o, uint64_t *kmers, uint32_t *counts)
{
    uint64_t w = 0;
    for (uint64_t i = 0; i < o->table_size; i++)
        if (o->keys[i]) {
            kmers[w] = orc_wang64_inverse(o->keys[i]);
            counts[w++] = o->counts[i];
        }
    return w;
}

/* Table hand-over for the sharded-run protocol tests (tests/test_sharded.py): the
 * state a rank passes to the next one.  No reference counterpart. */
ORC_API void
orc_ov_get_table(const orc_ov *o, uint64_t *keys, uint32_t *counts)
{
    memcpy(keys, o->keys, o->table_size * 8);
    memcpy(counts, o->counts, o->table_size * 4);
}
ORC_API void
orc_ov_set_table(orc_ov *o, const uint64_t *keys, const uint32_t *counts, uint64_t n_unique)
{
    memcpy(o->keys, keys, o->table_size * 8);
    if (counts) memcpy(o->counts, counts, o->table_size * 4);
    else memset(o->counts, 0, o->table_size * 4);
    o->n_unique = n_unique;
}
ORC_API void
orc_ov_set_counters(orc_ov *o, uint64_t n_seqs, uint64_t n_sampled, uint64_t total_frags)
{
    o->n_seqs = n_seqs;
    o->n_sampled = n_sampled;
    o->total_frags = total_frags;
}

/* ------------------------------------------------------------------ */
/* DedupEstimator (_qcmodule.c:4383-4517)                              */
/* ------------------------------------------------------------------ */

typedef struct {
    uint64_t max_stored, table_size, stored, mod_bits;
    uint64_t front_len, back_len, front_off, back_off;
    uint64_t *hash;
    uint32_t *count;
    uint8_t *fp; /* persistent fingerprint scratch: stale bytes are observable
                    in the paired path for reads shorter than the lengths */
} orc_dd;

ORC_API orc_dd *
orc_dd_new(uint64_t max_stored, uint64_t front_len, uint64_t back_len,
           uint64_t front_off, uint64_t back_off)
{
    orc_dd *d = calloc(1, sizeof(*d));
    d->max_stored = max_stored;
    d->front_len = front_len;
    d->back_len = back_len;
    d->front_off = front_off;
    d->back_off = back_off;
    uint64_t bits = (uint64_t)(log2((double)max_stored * 1.5) + 1); /* :4327 */
    d->table_size = 1ULL << bits;
    d->hash = calloc(d->table_size, 8);
    d->count = calloc(d->table_size, 4);
    d->fp = calloc(front_len + back_len + 1, 1);
    return d;
}
ORC_API void
orc_dd_free(orc_dd *d)
{
    if (!d) return;
    free(d->hash);
    free(d->count);
    free(d->fp);
    free(d);
}

static void
dd_escalate(orc_dd *d)
{
    uint64_t nb = d->mod_bits + 1, drop = (1ULL << nb) - 1, mask = d->table_size - 1;
    uint64_t *nh = calloc(d->table_size, 8);
    uint32_t *nc = calloc(d->table_size, 4);
    uint64_t kept = 0;
    for (uint64_t i = 0; i < d->table_size; i++) {
        if (d->count[i] == 0 || (d->hash[i] & drop)) continue;
        uint64_t j = (d->hash[i] >> nb) & mask;
        while (nc[j]) j = (j + 1) & mask; /* no equality test: duplicates survive */
        nh[j] = d->hash[i];
        nc[j] = d->count[i];
        kept++;
    }
    free(d->hash);
    free(d->count);
    d->hash = nh;
    d->count = nc;
    d->mod_bits = nb;
    d->stored = kept;
}

static void
dd_add_hash(orc_dd *d, uint64_t h)
{
    uint64_t m = d->mod_bits;
    if (h & ((1ULL << m) - 1)) return;
    if (d->stored >= d->max_stored) dd_escalate(d);
    /* the slot index still uses the pre-escalation bit count (:4430, :4442) */
    uint64_t mask = d->table_size - 1, i = (h >> m) & mask;
    for (;;) {
        if (d->count[i] == 0) {
            d->hash[i] = h;
            d->count[i] = 1;
            d->stored++;
            return;
        }
        if (d->hash[i] == h) {
            d->count[i]++;
            return;
        }
        i = (i + 1) & mask;
    }
}

ORC_API uint64_t
orc_dd_fingerprint_hash(orc_dd *d, const uint8_t *s, uint64_t L)
{
    uint64_t fl = d->front_len + d->back_len;
    if (L <= fl) return murmur3_h2(s, L, 0); /* :4472 */
    uint64_t rem = L - fl;
    uint64_t fo = rem / 2 < d->front_off ? rem / 2 : d->front_off;
    uint64_t bo = rem / 2 < d->back_off ? rem / 2 : d->back_off;
    memcpy(d->fp, s + fo, d->front_len);
    memcpy(d->fp + d->front_len, s + L - (bo + d->back_len), d->back_len);
    return murmur3_h2(d->fp, fl, L >> 6);
}
ORC_API uint64_t
orc_dd_pair_hash(orc_dd *d, const uint8_t *s1, uint64_t L1, const uint8_t *s2,
                 uint64_t L2)
{
    uint64_t fl = d->front_len + d->back_len;
    uint64_t f = d->front_len < L1 ? d->front_len : L1;
    uint64_t fo = d->front_off < L1 - f ? d->front_off : L1 - f;
    uint64_t b = d->back_len < L2 ? d->back_len : L2;
    uint64_t bo = d->back_off < L2 - b ? d->back_off : L2 - b;
    memcpy(d->fp, s1 + fo, f);
    memcpy(d->fp + f, s2 + bo, b); /* note: placed after the *clamped* front (:4513) */
    return murmur3_h2(d->fp, fl, (L1 + L2) >> 6);
}
ORC_API void
orc_dd_add(orc_dd *d, const uint8_t *buf, const orc_rec *recs, uint64_t n)
{
    for (uint64_t r = 0; r < n; r++)
        dd_add_hash(d, orc_dd_fingerprint_hash(d, buf + recs[r].seq_off, recs[r].seq_len));
}
ORC_API void
orc_dd_add_pair(orc_dd *d, const uint8_t *buf1, const orc_rec *recs1,
                const uint8_t *buf2, const orc_rec *recs2, uint64_t n)
{
    for (uint64_t r = 0; r < n; r++)
        dd_add_hash(d, orc_dd_pair_hash(d, buf1 + recs1[r].seq_off, recs1[r].seq_len,
                                        buf2 + recs2[r].seq_off, recs2[r].seq_len));
}
ORC_API void
orc_dd_add_raw_hash(orc_dd *d, uint64_t h)
{
    dd_add_hash(d, h);
}
ORC_API void
orc_dd_info(const orc_dd *d, uint64_t *out3)
{
    out3[0] = d->mod_bits;
    out3[1] = d->table_size;
    out3[2] = d->stored;
}
/* counts (and hashes) of occupied slots in slot order */
ORC_API uint64_t
orc_dd_counts(const orc_dd *d, uint64_t *counts, uint64_t *hashes)
{
    uint64_t w = 0;
    for (uint64_t i = 0; i < d->table_size; i++)
        if (d->count[i]) {
            if (hashes) hashes[w] = d->hash[i];
            counts[w++] = d->count[i];
        }
    return w;
}

/* ------------------------------------------------------------------ */
/* NanoStats (_qcmodule.c:248-322, 5006-5052, 5078-5259, 5269-5324)    */
/* ------------------------------------------------------------------ */

typedef struct {
    int64_t start_time;
    float duration;
    int32_t channel_id;
    uint32_t length;
    uint32_t pad;
    double cumulative_error_rate;
    uint64_t parent_id_hash;
} orc_nanoinfo; /* 40 bytes, same fields as struct NanoInfo (:4808-4815) */

typedef struct {
    uint64_t n_reads, cap;
    orc_nanoinfo *infos;
    int64_t min_time, max_time;
    int skipped;
} orc_ns;

static int64_t
dec_field(const uint8_t *s, const uint8_t *end, uint64_t len)
{
    if (len < 1 || len > 18 || s + len > end) return -1;
    int64_t v = 0;
    for (uint64_t i = 0; i < len; i++) {
        uint8_t d = s[i] - '0';
        if (d > 9) return -1;
        v = v * 10 + d;
    }
    return v;
}

/* "YYYY-MM-DDThh:mm:ss[.fff](Z|+hh:mm|-hh:mm)" -> unix seconds, -1 on error.
 * `end` bounds the readable bytes (the reference reads unguarded; inside a
 * valid header the outcomes are the same). */
static int64_t
nanopore_time(const uint8_t *s, const uint8_t *end)
{
    if (s + 20 > end) return -1;
    int64_t Y = dec_field(s, end, 4), M = dec_field(s + 5, end, 2), D = dec_field(s + 8, end, 2);
    int64_t h = dec_field(s + 11, end, 2), mi = dec_field(s + 14, end, 2), sec = dec_field(s + 17, end, 2);
    if ((Y | M | D | h | mi | sec) < 0 || s[4] != '-' || s[7] != '-' || s[10] != 'T' ||
        s[13] != ':' || s[16] != ':')
        return -1;
    const uint8_t *tz = s + 19;
    if (*tz == '.') {
        tz++;
        while (tz < end && *tz >= '0' && *tz <= '9') tz++;
    }
    if (tz >= end) return -1;
    if (*tz == '+' || *tz == '-') {
        int64_t oh = dec_field(tz + 1, end, 2), om = dec_field(tz + 4, end, 2);
        if ((oh | om) < 0 || tz[3] != ':') return -1;
        if (*tz == '+') { h += oh; mi += om; } else { h -= oh; mi -= om; }
    } else if (*tz != 'Z')
        return -1;
    if (Y < 1970 || M < 1 || M > 12) return -1;
    static const int cum[12] = {0, 31, 59, 90, 120, 151, 181, 212, 243, 273, 304, 334};
    int64_t y = Y - 1900, yday = cum[M - 1] + D - 1;
    return sec + mi * 60 + h * 3600 + yday * 86400 + (y - 70) * 31536000 +
           ((y - 69) / 4) * 86400 - ((y - 1) / 100) * 86400 + ((y + 299) / 400) * 86400;
}

static int
nano_header(const uint8_t *h, uint64_t n, int32_t *ch, int64_t *st)
{
    const uint8_t *end = h + n, *p = memchr(h, ' ', n);
    if (!p) return -1;
    p++;
    int64_t channel = -1, start = -1;
    while (p < end) {
        const uint8_t *eq = memchr(p, '=', (size_t)(end - p));
        if (!eq) return -1;
        const uint8_t *val = eq + 1, *ve = memchr(val, ' ', (size_t)(end - val));
        if (!ve) ve = end;
        uint64_t kl = (uint64_t)(eq - p);
        if (kl == 2 && p[0] == 'c' && p[1] == 'h') {
            /* the reference stores the parse into an int32 */
            int64_t v = dec_field(val, end, (uint64_t)(ve - val));
            channel = (int64_t)(int32_t)v;
        }
        else if (kl == 10 && memcmp(p, "start_time", 10) == 0)
            start = nanopore_time(val, end);
        p = ve + 1;
    }
    if (channel == -1 || start == -1) return -1;
    *ch = (int32_t)channel;
    *st = start;
    return 0;
}

/* strtoull(s, &end, 16) restricted to the question the reference asks: does
 * the conversion consume exactly the 8 bytes s[0..8) (s[8] is known not to be
 * a hex digit)?  Accepts what strtoull accepts: leading blanks, one sign, an
 * optional 0x/0X prefix, then at least one hex digit. */
static int
hex8_like_strtoull(const uint8_t *s, uint64_t *out)
{
    int i = 0, neg = 0;
    while (i < 8 && (s[i] == ' ' || (s[i] >= 9 && s[i] <= 13))) i++;
    if (i < 8 && (s[i] == '+' || s[i] == '-')) neg = s[i++] == '-';
    int digits_at = i;
    if (i + 2 < 8 + 1 && i + 1 < 8 && s[i] == '0' && (s[i + 1] | 0x20) == 'x') {
        /* prefix only counts when a hex digit follows it */
        uint8_t c = i + 2 < 8 ? s[i + 2] : 0;
        int hexd = (c >= '0' && c <= '9') || ((c | 0x20) >= 'a' && (c | 0x20) <= 'f');
        if (hexd) digits_at = i + 2;
    }
    uint64_t v = 0;
    int j = digits_at;
    for (; j < 8; j++) {
        uint8_t c = s[j];
        int d = (c >= '0' && c <= '9') ? c - '0'
                : ((c | 0x20) >= 'a' && (c | 0x20) <= 'f') ? (c | 0x20) - 'a' + 10 : -1;
        if (d < 0) break;
        v = v << 4 | (uint64_t)d;
    }
    if (j != 8 || j == digits_at) return 0;
    *out = neg ? (uint64_t)0 - v : v;
    return 1;
}
/* first 8 + last 8 hex digits of a 36-char uuid4, 0 when malformed (:5153-5179) */
static uint64_t
uuid4_hash(const uint8_t *u)
{
    if (u[8] != '-' || u[13] != '-' || u[14] != '4' || u[18] != '-' || u[23] != '-' || u[36] != 0)
        return 0;
    uint64_t a, b;
    if (!hex8_like_strtoull(u, &a)) return 0;
    if (!hex8_like_strtoull(u + 28, &b)) return 0;
    return a << 32 | (b & 0xffffffffULL);
}

/* one BAM aux field; returns its length or -1 (truncated / unknown type) */
static int64_t
aux_len(const uint8_t *t, uint64_t avail)
{
    if (avail < 4) return -1;
    uint8_t ty = t[2];
    uint64_t head = 3, count = 1, width;
    if (ty == 'B') {
        if (avail < 8) return -1;
        ty = t[3];
        count = rd32(t + 4);
        head = 8;
        if (ty == 'Z' || ty == 'H') return -1;
    }
    switch (ty) {
        case 'A': case 'c': case 'C': width = 1; break;
        case 's': case 'S': width = 2; break;
        case 'i': case 'I': case 'f': width = 4; break;
        case 'Z': case 'H': {
            const uint8_t *z = memchr(t + 3, 0, avail - 3);
            if (!z) return -1;
            width = (uint64_t)(z - (t + 3)) + 1;
            break;
        }
        default: return -1;
    }
    uint64_t len = head + count * width;
    return len > avail ? -1 : (int64_t)len;
}

/* 0 ok, -1 malformed tags (the reference raises) ; *warn set when a pi tag is
 * not 36 characters (the reference warns and ignores it) */
static int
nano_tags(const uint8_t *t, uint64_t n, orc_nanoinfo *o, int *warn)
{
    o->channel_id = -1;
    o->duration = 0.0f;
    o->start_time = 0;
    o->parent_id_hash = 0;
    while (n) {
        int64_t len = aux_len(t, n);
        if (len < 0) return -1;
        uint8_t ty = t[2];
        if (t[0] == 'c' && t[1] == 'h') {
            const uint8_t *v = t + 3;
            switch (ty) {
                case 'c': o->channel_id = (int8_t)v[0]; break;
                case 'C': o->channel_id = v[0]; break;
                case 's': o->channel_id = (int16_t)rd16(v); break;
                case 'S': o->channel_id = (int32_t)rd16(v); break;
                case 'i': case 'I': o->channel_id = (int32_t)rd32(v); break;
                default: return -1;
            }
        }
        else if (t[0] == 's' && t[1] == 't') {
            if (ty != 'Z') return -1;
            o->start_time = nanopore_time(t + 3, t + len);
        }
        else if (t[0] == 'd' && t[1] == 'u') {
            if (ty != 'f') return -1;
            memcpy(&o->duration, t + 3, 4);
        }
        else if (t[0] == 'p' && t[1] == 'i') {
            if (ty != 'Z') return -1;
            if (len - 4 != 36) *warn = 1;
            else o->parent_id_hash = uuid4_hash(t + 3);
        }
        t += len;
        n -= (uint64_t)len;
    }
    return 0;
}

ORC_API orc_ns *
orc_ns_new(void)
{
    return calloc(1, sizeof(orc_ns));
}
ORC_API void
orc_ns_free(orc_ns *s)
{
    if (!s) return;
    free(s->infos);
    free(s);
}
/* 0 ok; 1 header unparsable at *skip_rec (module now skipped); -1 bad tags */
ORC_API int
orc_ns_add(orc_ns *s, const uint8_t *buf, const orc_rec *recs, uint64_t n,
           uint64_t *skip_rec)
{
    if (s->skipped) return 0;
    for (uint64_t r = 0; r < n; r++) {
        if (s->n_reads == s->cap) {
            uint64_t nc = s->cap ? s->cap * 2 : 16384;
            s->infos = realloc(s->infos, nc * sizeof(orc_nanoinfo));
            memset(s->infos + s->cap, 0, (nc - s->cap) * sizeof(orc_nanoinfo));
            s->cap = nc;
        }
        orc_nanoinfo *o = s->infos + s->n_reads;
        o->length = recs[r].seq_len;
        if (recs[r].tags_len) {
            int warn = 0;
            if (nano_tags(buf + recs[r].tags_off, recs[r].tags_len, o, &warn)) return -1;
        }
        else {
            int32_t ch;
            int64_t st;
            if (nano_header(buf + recs[r].name_off, recs[r].name_len, &ch, &st)) {
                s->skipped = 1;
                *skip_rec = r;
                return 1;
            }
            o->channel_id = ch;
            o->start_time = st;
        }
        o->cumulative_error_rate = recs[r].err_sum;
        if (o->start_time > s->max_time) s->max_time = o->start_time;
        if (s->min_time == 0 || o->start_time < s->min_time) s->min_time = o->start_time;
        s->n_reads++;
    }
    return 0;
}
ORC_API void
orc_ns_info(const orc_ns *s, uint64_t *n_reads, int64_t *min_time, int64_t *max_time,
            int *skipped)
{
    *n_reads = s->n_reads;
    *min_time = s->min_time;
    *max_time = s->max_time;
    *skipped = s->skipped;
}
ORC_API void
orc_ns_infos(const orc_ns *s, orc_nanoinfo *out)
{
    memcpy(out, s->infos, s->n_reads * sizeof(orc_nanoinfo));
}

/* ------------------------------------------------------------------ */
/* InsertSizeMetrics (_qcmodule.c:5571-5611, 5634-5744)                */
/* ------------------------------------------------------------------ */

#define ORC_ADAPTER_STORE 31
typedef struct {
    uint64_t hash, count;
    uint8_t len, seq[ORC_ADAPTER_STORE];
} orc_is_entry;

typedef struct {
    uint64_t max_adapters, table_size, total, n_ad1, n_ad2, max_insert;
    uint64_t n_entries[2];
    orc_is_entry *tab[2];
    uint64_t *sizes;
} orc_is;

ORC_API orc_is *
orc_is_new(uint64_t max_adapters)
{
    orc_is *m = calloc(1, sizeof(*m));
    m->max_adapters = max_adapters;
    uint64_t bits = (uint64_t)(log2((double)max_adapters * 1.5) + 1);
    m->table_size = 1ULL << bits;
    m->tab[0] = calloc(m->table_size, sizeof(orc_is_entry));
    m->tab[1] = calloc(m->table_size, sizeof(orc_is_entry));
    m->sizes = calloc(1, 8);
    return m;
}
ORC_API void
orc_is_free(orc_is *m)
{
    if (!m) return;
    free(m->tab[0]);
    free(m->tab[1]);
    free(m->sizes);
    free(m);
}

static inline uint8_t
comp_upper(uint8_t c)
{
    switch (c | 0x20) {
        case 'a': return 'T';
        case 'c': return 'G';
        case 'g': return 'C';
        case 't': return 'A';
    }
    return 0; /* never equals a read letter (:5614-5630) */
}

/* does the 16-mer at r (raw bytes) hit `needle` with <= 1 mismatch, given the
 * reference's prefilter: one 8-byte half must match after clearing bit 5 */
static int
needle_hit(const uint8_t *r, const uint8_t *needle)
{
    int half0 = 1, half1 = 1, mism = 0;
    for (int i = 0; i < 16; i++) {
        if ((r[i] & 0xDF) != needle[i]) { if (i < 8) half0 = 0; else half1 = 0; }
        if (r[i] != needle[i]) mism++;
    }
    return (half0 || half1) && mism <= 1;
}

ORC_API uint64_t
orc_insert_size(const uint8_t *s1, uint64_t L1, const uint8_t *s2, uint64_t L2)
{
    if (L1 < 16 || L2 < 16) return 0;
    uint8_t head[16], tail[16];
    for (int i = 0; i < 16; i++) {
        head[15 - i] = comp_upper(s2[i]);
        tail[15 - i] = comp_upper(s2[L2 - 16 + i]);
    }
    for (uint64_t i = 0; i + 16 <= L1; i++) {
        if (needle_hit(s1 + i, head)) return i + 16;
        if (needle_hit(s1 + i, tail)) return i + L2;
    }
    return 0;
}

static void
is_add_adapter(orc_is *m, int which, const uint8_t *a, uint64_t len)
{
    uint64_t h = murmur3_h2(a, len, 0), mask = m->table_size - 1, i = h & mask;
    int full = m->n_entries[which] == m->max_adapters;
    orc_is_entry *t = m->tab[which];
    for (;;) {
        orc_is_entry *e = t + i;
        if (e->hash == h) {
            if (e->len == len && memcmp(e->seq, a, len) == 0) {
                e->count++;
                return;
            }
        }
        else if (e->count == 0) {
            if (!full) {
                e->hash = h;
                e->len = (uint8_t)len;
                memcpy(e->seq, a, len);
                e->count = 1;
                m->n_entries[which]++;
            }
            return;
        }
        i = (i + 1) & mask;
    }
}

ORC_API void
orc_is_add_pair_seq(orc_is *m, const uint8_t *s1, uint64_t L1, const uint8_t *s2,
                    uint64_t L2)
{
    uint64_t ins = orc_insert_size(s1, L1, s2, L2);
    if (ins > m->max_insert) {
        m->sizes = realloc(m->sizes, (ins + 1) * 8);
        memset(m->sizes + m->max_insert + 1, 0, (ins - m->max_insert) * 8);
        m->max_insert = ins;
    }
    m->total++;
    m->sizes[ins]++;
    if (ins == 0) return;
    if (L1 > ins) {
        m->n_ad1++;
        uint64_t l = L1 - ins;
        is_add_adapter(m, 0, s1 + ins, l < ORC_ADAPTER_STORE ? l : ORC_ADAPTER_STORE);
    }
    if (L2 > ins) {
        m->n_ad2++;
        uint64_t l = L2 - ins;
        is_add_adapter(m, 1, s2 + ins, l < ORC_ADAPTER_STORE ? l : ORC_ADAPTER_STORE);
    }
}
ORC_API void
orc_is_add_pair(orc_is *m, const uint8_t *buf1, const orc_rec *r1, const uint8_t *buf2,
                const orc_rec *r2, uint64_t n)
{
    for (uint64_t i = 0; i < n; i++)
        orc_is_add_pair_seq(m, buf1 + r1[i].seq_off, r1[i].seq_len, buf2 + r2[i].seq_off,
                            r2[i].seq_len);
}
ORC_API void
orc_is_info(const orc_is *m, uint64_t *out6)
{
    out6[0] = m->total;
    out6[1] = m->n_ad1;
    out6[2] = m->n_ad2;
    out6[3] = m->max_insert;
    out6[4] = m->n_entries[0];
    out6[5] = m->n_entries[1];
}
ORC_API void
orc_is_sizes(const orc_is *m, uint64_t *out)
{
    memcpy(out, m->sizes, (m->max_insert + 1) * 8);
}
/* entries of table `which` in slot order: seqs is n*32 bytes (len byte + 31) */
ORC_API uint64_t
orc_is_adapters(const orc_is *m, int which, uint8_t *seqs, uint64_t *counts)
{
    uint64_t w = 0;
    for (uint64_t i = 0; i < m->table_size; i++) {
        const orc_is_entry *e = m->tab[which] + i;
        if (!e->count) continue;
        seqs[w * 32] = e->len;
        memcpy(seqs + w * 32 + 1, e->seq, ORC_ADAPTER_STORE);
        counts[w++] = e->count;
    }
    return w;
}

/* mate check (_qcmodule.c:778-800): ids up to the first blank are equal,
 * ignoring a final 1/2 on both */
ORC_API int
orc_names_are_mates(const uint8_t *n1, uint64_t l1, const uint8_t *n2, uint64_t l2)
{
    uint64_t id = 0;
    while (id < l1 && n1[id] != ' ' && n1[id] != '\t') id++;
    if (l2 < id) return 0;
    if (l2 > id && n2[id] != ' ' && n2[id] != '\t') return 0;
    if (id > 0) {
        uint8_t a = n1[id - 1], b = n2[id - 1];
        if ((a == '1' || a == '2') && (b == '1' || b == '2')) id--;
    }
    return memcmp(n1, n2, id) == 0;
}
