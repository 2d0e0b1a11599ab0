// modules.cuh -- collector state shared between the per-module entry points
// (qc.cu, adapters.cu, pertile.cu, dedup.cu) and the fused short-read pass
// (fused.cu), which feeds several collectors from one walk over the text.
#pragma once
#include "common.cuh"

// ---- QCMetrics (qc.cu) -------------------------------------------------------
struct sq_qc {
    sq_ctx *ctx = nullptr;
    uint64_t ea_len = 0, n_reads = 0, max_len = 0, cap_len = 0;
    uint64_t *base = nullptr, *phred = nullptr;        // [cap_len][5], [cap_len][12]
    uint64_t *ea_base = nullptr, *ea_phred = nullptr;  // [ea_len][5], [ea_len][12]
    uint64_t *gc = nullptr, *mean_phred = nullptr;     // [101], [94]
    unsigned long long *err_key = nullptr;             // (global record << 8 | byte), min
};

int qc_grow(sq_qc *m, uint64_t len);
// per-position tables only (base / phred / end-anchored); the per-read part is the caller's
int qc_add_vertical(sq_qc *m, sq_batch *b);

// ---- AdapterCounter (adapters.cu) ---------------------------------------------
constexpr int AD_TPB = 128;
constexpr int AD_MAXLEN = 64;

struct sq_adapters {
    sq_ctx *ctx = nullptr;
    uint32_t n_adapters = 0, max_pat_len = 0;
    uint64_t n_seqs = 0, max_len = 0, cap_len = 0;
    uint8_t *pat = nullptr;      // device [n_adapters][64] letter classes 0..4
    uint32_t *plen = nullptr;    // device [n_adapters]
    uint64_t *counts = nullptr;  // device [n_adapters][2][cap_len]: forward, reverse
};

int adapters_grow(sq_adapters *a, uint64_t len);

// ---- PerTileQuality (pertile.cu) ----------------------------------------------
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
    unsigned int n_changes;        // this array: reads whose tile differs from the previous read's
    unsigned int qmin, qmax;       // smallest / largest quality byte among the sampled ones (all arrays so far)
    unsigned int pad;
    unsigned long long tile_lo, tile_hi;  // this array: smallest / largest tile id among the kept reads
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
    uint64_t *lut = nullptr;        // [PT_LUT_NK][94] in-binade increments r_k(10^-(q/10))
    bool skipped = false;
    uint64_t skipped_record = 0;
    std::vector<uint8_t> skipped_name;
    // host mirror after the last sync
    uint64_t n_slots = 0, max_len = 0;
    uint32_t qmin = 0xFFFFFFFFu, qmax = 0;  // sampled quality byte range (k_fused_reads), 0xFFFFFFFF: not known yet
};

constexpr uint64_t PT_HARD = 1ULL << 53;          // no valid in-binade increment reaches this
constexpr uint64_t PT_MANT = (1ULL << 52) - 1;
constexpr int PT_LUT_KMIN = 1023 - 30;            // binades 2^-30 .. 2^25 are tabulated
constexpr int PT_LUT_NK = 56;

// r_k(e): e in ulps of binade k (biased exponent), round to nearest; PT_HARD when
// the addition cannot be expressed that way (tie, e >= 2^k, s == 0)
__host__ __device__ inline uint64_t pt_increment(uint32_t k, uint64_t ebits) {
    const int d = (int)k - (int)(ebits >> 52);
    if (k == 0 || d < 1) return PT_HARD;
    if (d >= 64) return 0;
    const uint64_t m = (ebits & PT_MANT) | (1ULL << 52);
    uint64_t r = m >> d;
    const uint64_t rem = m & ((1ULL << d) - 1), half = 1ULL << (d - 1);
    if (rem > half) r++;
    else if (rem == half) return PT_HARD;
    return r;
}

struct PtSeg {  // a run of consecutive reads of one tile: rows [lo, hi) of `order` (or of the array)
    uint32_t lo, hi, slot;
    uint32_t data;  // general path: row of approx / kguess / incr holding this segment's sums; run path: the
                    // fixed tile whose quality histograms cover exactly these reads; PT_NONE: replay the reads
};

// Layout of the per-(fixed tile, position) quality histograms k_fused_columns writes for the run path:
// qh[tile][seg_bytes]; position c = 4 * cg + j owns the QW words at word ((j * CG + cg) * QW); byte i of
// those words = number of the tile's reads with phred qbase + i at that position (0 <= i < qrows).
struct PtHistGeom {
    uint32_t qbase = 0, qrows = 0, QW = 0, CG = 0, seg_bytes = 0;
};


// One record array's way through PerTileQuality once the tile ids are known (tile[r] < 0:
// unparsable header, already folded into st->fail_idx):
//   pt_prepare   tile ids -> slots, table growth, length counts.  R != 0 announces that the
//                caller runs k_fused_columns over fixed tiles of R records: when the reads arrive
//                in tile runs the plan then holds the segments (plan.runs) and the caller must
//                have k_fused_columns fill plan.qh (quality histogram per fixed tile and position)
//                and plan.oob (tiles that met a quality outside the tabulated rows)
//   pt_finish    the ordered chains (run path), or the sort-based general path; frees the plan
struct PtPlan {
    bool work = false, runs = false;
    uint32_t R = 0, n_ftiles = 0, W = 0, n_slots = 0, width = 0;
    unsigned long long fail_idx = ~0ULL;
    uint32_t *slot = nullptr, *idx = nullptr, *tmpk = nullptr, *tmpv = nullptr, *seg = nullptr;
    uint32_t *runs_cnt = nullptr, *seg_off = nullptr, *nseg = nullptr;
    uint8_t *uniform = nullptr;   // [n_ftiles] 1: all records of the fixed tile belong to one flow-cell tile
    uint8_t *oob = nullptr;       // [n_ftiles] 1: the histogram of the tile is incomplete, replay its reads
    PtSeg *segs = nullptr;
    PtHistGeom hg;
    uint8_t *qh = nullptr;        // [n_ftiles][hg.seg_bytes]
};
// qh / oob != nullptr: the histograms were already written (k_onewalk computes the tile ids and the
// histograms in the same pass); the plan owns and frees them either way
int pt_prepare(sq_pertile *p, sq_batch *b, long long *tile, uint32_t R, uint32_t n_ftiles, uint32_t W,
               const PtHistGeom &hg, PtPlan *pl, uint8_t *qh, uint8_t *oob);
// sampled quality byte range for the next k_fused_columns launch (syncs once, for the first array)
int pt_quality_range(sq_pertile *p, uint32_t *qmin, uint32_t *qmax);
int pt_finish(sq_pertile *p, sq_batch *b, PtPlan *pl);
void pt_plan_free(sq_ctx *ctx, PtPlan *pl);

// ---- DedupEstimator (dedup.cu) -------------------------------------------------
constexpr int DD_TPB = 256;
constexpr uint32_t DD_NEW = 0xFFFFFFFFu;     // sampled, not in the table
constexpr uint32_t DD_NOPASS = 0xFFFFFFFEu;  // not sampled at this level
constexpr uint64_t DD_EMPTY = ~0ULL;

struct DdTable {
    uint64_t *hash = nullptr;   // [size]
    uint32_t *count = nullptr;  // [size], 0 = empty
    uint64_t *prio = nullptr;   // [size], DD_EMPTY = empty; arrival priority of the occupant
};

struct DdCounters {  // device
    unsigned long long r_full;   // record index of the K-th first occurrence
    unsigned long long r_star;   // the add that triggers the escalation
    unsigned int n_new;          // distinct new keys in the segment
    unsigned int kept;           // entries surviving a rebuild
    unsigned int inserted_one;   // k_dd_insert_one created an entry
    unsigned int pad;
};

struct sq_dedup {
    sq_ctx *ctx = nullptr;
    uint64_t max_stored = 0, table_size = 0, stored = 0, mod_bits = 0;
    uint64_t front_len = 0, back_len = 0, front_off = 0, back_off = 0;
    uint64_t n_records = 0;  // records added so far (global index base)
    DdTable tab, spare;
    DdCounters *cnt = nullptr;
    uint8_t *stale_fp = nullptr;  // pair path: persistent fingerprint scratch of the reference
    uint32_t *pair_range = nullptr;  // pair path: {first, last} short pair of the batch being hashed
    // sharded runs (rank > 0): keep the fingerprint hashes, the table lives on the first rank
    bool deferred = false;
    struct Kept { uint64_t *hashes; uint32_t n; };
    std::vector<Kept> kept;
    uint64_t *compact = nullptr;  // hashes that pass the requested mask, in record order
    uint64_t compact_n = 0;
};

// Takes `hashes` in record order.  In deferred mode the buffer is KEPT (the caller must not free it).
int dedup_consume(sq_dedup *d, const uint64_t *hashes, uint32_t n);

// ---- NanoStats (nanostats.cu, report.cu) -----------------------------------------
constexpr uint32_t NS_NAME_CAP = 1u << 16;

struct NsState {  // device
    unsigned long long fail_idx;      // global index of the first unparsable header
    unsigned long long tag_err_idx;   // first malformed aux block: record << 24 | kind << 16 | detail (NS_TAG_*), min
    unsigned long long pi_warnings;
    long long min_time, max_time;
    unsigned int nonpositive_time;    // a timestamp <= 0 exists (order-dependent min, see k_ns_minmax_ordered)
    unsigned int pad;
    unsigned long long pi_first;      // first pi:Z tag that is not 36 characters long: record << 24 | its length, min
};

struct sq_nanostats {
    sq_ctx *ctx = nullptr;
    uint64_t n_added = 0, cap = 0;
    sq_nanoinfo *infos = nullptr;
    NsState *st = nullptr;
    bool skipped = false;          // known on the host after a sync
    uint64_t skipped_record = 0;
    std::vector<uint8_t> skipped_name;
    // the header that switches the module off is copied aside on the device while its record array is alive
    // (k_ns_capture_name); the host learns about it without waiting: a copy of the state trails every add
    uint8_t *d_name = nullptr;          // [NS_NAME_CAP]
    uint32_t *d_name_len = nullptr;
    NsState *h_peek = nullptr;          // pinned
    cudaEvent_t peek_ev = nullptr;
    bool peek_pending = false;
    // per-channel results of the last sq_nanostats_report, until sq_nanostats_report_channels fetches them
    uint64_t rp_n = 0;
    int32_t *rp_channel = nullptr;
    unsigned long long *rp_bases = nullptr;
    double *rp_error = nullptr;
};

// ---- OverrepresentedSequences: an add in two halves (overrep.cu; sq_fused_add puts other work between them) ----
struct OvPendingAdd {
    uint64_t *frag_hash = nullptr;
    uint32_t *frag_n = nullptr;
    uint64_t n_sampled = 0, total = 0;
    uint32_t fcap = 0;
};
int ov_add_begin(sq_overrep *o, sq_batch *b, OvPendingAdd *pa);
int ov_add_end(sq_overrep *o, OvPendingAdd *pa);
void ov_add_abandon(sq_overrep *o, OvPendingAdd *pa);
