/*
 * sqgpu.h -- C ABI of libsqgpu.so, the B200 (sm_100a) implementation of the
 * per-record QC hot path of rhpvorderman/sequali.
 *
 * This is the drop-in boundary.  The reference's Python extension
 * (`sequali._qc`, src/sequali/_qcmodule.c) is the only caller of the hot path;
 * every entry point below replaces one piece of that file and says which.
 * All functions are `extern "C"`, take plain pointers and sizes, return 0 on
 * success and a negative SQ_E_* code on failure (text via sq_last_error()).
 * No torch / Python types cross this boundary.
 *
 * Execution model: everything a collector does is enqueued on the context's
 * CUDA stream; `*_add` calls return immediately.  Results become observable
 * through the `*_sync` / `*_read_*` calls, which synchronise first -- the same
 * observability the reference offers (getters only).
 */
#ifndef SQGPU_H
#define SQGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SQ_API __attribute__((visibility("default")))

/* ---- error codes ------------------------------------------------------- */
enum {
    SQ_OK = 0,
    SQ_E_CUDA = -1,      /* CUDA runtime error (see sq_last_error)           */
    SQ_E_ARG = -2,       /* invalid argument                                  */
    SQ_E_NOMEM = -3,     /* host or device allocation failed                  */
    SQ_E_FORMAT = -4,    /* malformed input (details in the out struct)       */
    SQ_E_NODEVICE = -5,  /* no CUDA device: there is no CPU fallback          */
    SQ_E_LIMIT = -6,     /* an implementation limit was exceeded              */
};

/* FASTQ format errors, in the reference's check order (_qcmodule.c:1097-1150) */
enum {
    SQ_PARSE_OK = 0,
    SQ_PARSE_NO_AT = 1,    /* "Record does not start with @ but with %c"  :1098 */
    SQ_PARSE_NO_PLUS = 2,  /* "Record second header does not start with + ..." :1121 */
    SQ_PARSE_LEN = 3,      /* "Record sequence and qualities do not have equal length" :1143 */
    SQ_PARSE_ASCII = 4,    /* "Found non-ASCII character in file: %c" :1062 */
    SQ_PARSE_INFLATE = 5,  /* sq_fastq_stream over BGZF: a member failed to inflate (err_record = member index) */
};

typedef struct sq_ctx sq_ctx;     /* one CUDA device + stream + scratch      */
typedef struct sq_batch sq_batch; /* a device-resident record array          */

/* Host view of one record: the reference's struct FastqMeta
 * (_qcmodule.c:337-355) with the pointer replaced by offsets into the batch's
 * byte buffer (offsets are absolute, not relative to the name). */
typedef struct {
    uint32_t name_off, name_len;
    uint32_t seq_off, seq_len;
    uint32_t qual_off;
    uint32_t tags_off, tags_len;
    uint32_t reserved;
    double err_sum; /* accumulated_error_rate, valid after sq_qc_add + sync */
} sq_meta;

typedef struct {
    uint64_t n_records;   /* complete records parsed (<= max_records)         */
    uint64_t consumed;    /* bytes of input covered by those records          */
    uint64_t n_newlines;  /* newlines seen in the whole input                 */
    uint32_t max_seq_len; /* longest sequence among the parsed records        */
    int32_t err_code;     /* SQ_PARSE_*                                       */
    uint64_t err_record;  /* record index of the first format error           */
    uint64_t err_pos;     /* byte offset of the offending byte / record name  */
} sq_parse_info;

/* ---- context ----------------------------------------------------------- */
SQ_API int sq_device_count(void);
SQ_API const char *sq_last_error(void);
/* NUMA node of the device's PCIe root (-1: unknown): where pinned staging memory should be allocated */
SQ_API int sq_device_numa_node(int device);
SQ_API int sq_ctx_create(int device, sq_ctx **out);
SQ_API void sq_ctx_destroy(sq_ctx *ctx);
/* Scratch blocks of 8 MiB and more are kept by the context after their first use instead of going back to the
 * stream-ordered pool (size classes; reuse ordered by the event of the last use): how many such requests went to
 * cudaMallocAsync, how many were served from the cache, and the bytes idle in the cache right now. */
SQ_API int sq_ctx_block_cache_stats(sq_ctx *ctx, uint64_t *from_driver, uint64_t *from_cache,
                                    uint64_t *idle_bytes);
SQ_API int sq_ctx_sync(sq_ctx *ctx);
/* raw stream handle (cudaStream_t) for callers that time with CUDA events */
SQ_API void *sq_ctx_stream(sq_ctx *ctx);
/* number of kernels this library has launched on ctx so far */
SQ_API uint64_t sq_ctx_launch_count(sq_ctx *ctx);
/* per-kernel device time (CUDA events around every launch on the context
 * stream); report = "kernel launches total_ms" lines, most expensive first */
SQ_API int sq_ctx_profile(sq_ctx *ctx, int enable);
SQ_API int sq_ctx_profile_report(sq_ctx *ctx, char *buf, size_t cap);
/* stopwatch: CUDA events recorded on the context stream */
SQ_API int sq_timer_start(sq_ctx *ctx);
SQ_API int sq_timer_stop(sq_ctx *ctx, double *elapsed_ms);
/* pinned host staging memory for the parsers' read buffers */
SQ_API void *sq_pinned_alloc(sq_ctx *ctx, size_t nbytes);
SQ_API void sq_pinned_free(sq_ctx *ctx, void *p);
/* plain device memory for callers that keep inputs HBM-resident (bench) */
SQ_API void *sq_device_alloc(sq_ctx *ctx, size_t nbytes);
SQ_API void sq_device_free(sq_ctx *ctx, void *p);
SQ_API int sq_memcpy_h2d(sq_ctx *ctx, void *dst, const void *src, size_t n);
SQ_API int sq_memcpy_d2h(sq_ctx *ctx, void *dst, const void *src, size_t n);

/* ---- record arrays ------------------------------------------------------ */
/* FastqParser_create_record_array (_qcmodule.c:965-1184): copy `nbytes` of
 * FASTQ text to the device (cudaMemcpyAsync; `text` may be pinned), find the
 * record boundaries there and validate them.  Synchronises, because the
 * caller needs n_records / consumed to carry the leftover.  On a format error
 * returns SQ_E_FORMAT with *out == NULL and info filled in. */
SQ_API int sq_batch_from_fastq(sq_ctx *ctx, const uint8_t *text, uint64_t nbytes,
                               uint64_t max_records, sq_batch **out, sq_parse_info *info);
/* Same, for text that already lives in device memory (no copy; the batch
 * borrows `dev_text`, which must outlive it). */
SQ_API int sq_batch_from_device_fastq(sq_ctx *ctx, const uint8_t *dev_text, uint64_t nbytes,
                                      uint64_t max_records, sq_batch **out,
                                      sq_parse_info *info);
/* FastqRecordArrayView.__new__ / FastqRecordView (_qcmodule.c:373-482,
 * 610-687): a packed name|seq|qual|tags buffer with host-built descriptors. */
SQ_API int sq_batch_from_packed(sq_ctx *ctx, const uint8_t *buf, uint64_t nbytes,
                                const sq_meta *metas, uint64_t n, sq_batch **out);
/* BamParser__next__ record walk (_qcmodule.c:1623-1637) on the HOST bytes read so far: rec_off[0..*n_kept)
 * = offsets of the complete records to keep (cap slots; nbytes / 36 + 1 always suffices for well
 * formed records), *n_skipped = complete records dropped for flag & (0x100 | 0x800) (:1633),
 * *consumed = bytes covered by complete records (the rest is the caller's leftover). */
SQ_API int sq_bam_walk(const uint8_t *bam, uint64_t nbytes, uint64_t *rec_off, uint64_t cap, uint64_t *n_kept,
                       uint64_t *n_skipped, uint64_t *consumed);
/* BamParser__next__ record decode (_qcmodule.c:1623-1694): `bam` holds
 * alignment records; rec_off[i] is the offset of the i-th record to keep (the
 * host walks the block_size chain and drops secondary/supplementary records).
 * The 4-bit sequence -> ASCII and quality +33 decode runs on the device. */
SQ_API int sq_batch_from_bam(sq_ctx *ctx, const uint8_t *bam, uint64_t nbytes,
                             const uint64_t *rec_off, uint64_t n, sq_batch **out,
                             uint64_t *packed_len);
/* BamParser__next__ (_qcmodule.c:1506-1703) entirely on the device: `bam` = nbytes of alignment records starting
 * at a record boundary (host memory, pinned for full copy speed), n_ref = the reference count of the BAM header.
 * The bytes are copied once; the block_size chain (:1623-1637), the flag & 0x900 drop (:1633) and the decode
 * run on the device, the host never reads a record.  *out = the record array of the kept complete records (NULL
 * when there is none), *n_skipped = complete records dropped, *consumed = bytes covered by complete records (the
 * rest is the caller's leftover).  The chain is found by testing every byte offset for a self-consistent record
 * header and pointer doubling over the survivors; a chain that leaves them (a record the reference accepts
 * although its header is not self-consistent) is followed by a plain one-thread walk from there, so the result
 * is sq_bam_walk's for any input.  n_ref only sharpens the test. */
SQ_API int sq_batch_from_bam_bytes(sq_ctx *ctx, const uint8_t *bam, uint64_t nbytes, int32_t n_ref,
                                   sq_batch **out, uint64_t *n_kept, uint64_t *n_skipped,
                                   uint64_t *consumed, uint64_t *packed_len);
/* The device chain alone: same results as sq_bam_walk, offsets copied back to rec_off[0..cap). */
SQ_API int sq_bam_walk_device(sq_ctx *ctx, const uint8_t *bam, uint64_t nbytes, int32_t n_ref,
                              uint64_t *rec_off, uint64_t cap, uint64_t *n_kept, uint64_t *n_skipped,
                              uint64_t *consumed);
/* FastqParser's read loop (_qcmodule.c:985-1029: readinto + leftover carry,
 * :1187 __next__) for uncompressed text in (pinned) HOST memory: windows of
 * `window` bytes are copied host->device on a separate copy stream ahead of
 * the boundary scan (triple-buffered), each record array is [leftover of the
 * previous array | next window].  sq_fastq_stream_next returns the next record
 * array, or *out == NULL when the text is exhausted; a trailing partial record
 * is reported by sq_fastq_stream_leftover (the caller raises EOFError :1073). */
typedef struct sq_fastq_stream sq_fastq_stream;
SQ_API int sq_fastq_stream_create(sq_ctx *ctx, const uint8_t *host_text, uint64_t nbytes,
                                  uint64_t window, sq_fastq_stream **out);
SQ_API int sq_fastq_stream_next(sq_fastq_stream *s, sq_batch **out, sq_parse_info *info);
SQ_API uint64_t sq_fastq_stream_leftover(const sq_fastq_stream *s);
SQ_API void sq_fastq_stream_destroy(sq_fastq_stream *s);
/* ---- BGZF on the device ------------------------------------------------------
 * In the reference decompression is xopen's job on host threads (src/sequali/util.py:108-123) and the
 * documented bottleneck (README.rst:168-171).  A BGZF stream (bgzip'd FASTQ, every BAM) is a chain of
 * independent gzip members of <= 64 KiB of text: the host hops over the member headers, the compressed
 * bytes cross PCIe, one warp inflates one member. */
typedef struct {
    uint64_t comp_off; /* offset of the member's DEFLATE payload in the compressed stream */
    uint64_t text_off; /* offset of its text in the inflated stream                         */
    uint32_t comp_len, text_len;
} sq_bgzf_block;
/* member headers of host[0 .. nbytes): blocks[0 .. *n_blocks) (at most cap), *consumed = bytes covered by
 * the complete members listed, *text_bytes = their inflated size.  SQ_E_FORMAT for anything that is not a
 * BGZF member (a plain gzip stream has no block index and cannot be inflated in parallel).  blocks == NULL
 * counts the members only. */
SQ_API int sq_bgzf_scan(const uint8_t *host, uint64_t nbytes, sq_bgzf_block *blocks, uint64_t cap, uint64_t *n_blocks,
                        uint64_t *consumed, uint64_t *text_bytes);
/* inflate blocks[0 .. n) of the HOST stream `host_comp` into DEVICE memory: the text of block i lands at
 * dev_out + (blocks[i].text_off - blocks[0].text_off).  SQ_E_FORMAT + *bad_block / *bad_code for a corrupt
 * member.  CRC-32 is not checked; the trailer's text size and the decoder's consistency checks are. */
SQ_API int sq_bgzf_inflate(sq_ctx *ctx, const uint8_t *host_comp, uint64_t nbytes, const sq_bgzf_block *blocks, uint64_t n,
                           uint8_t *dev_out, uint64_t *bad_block, int *bad_code);
/* sq_fastq_stream over BGZF-compressed FASTQ in (pinned) host memory: windows of whole members holding
 * ~`window` bytes of TEXT travel compressed, are inflated straight into the record array's text buffer
 * behind the leftover of the previous array, and parsed there.  Use with sq_fastq_stream_next / _leftover /
 * _destroy. */
SQ_API int sq_fastq_stream_create_bgzf(sq_ctx *ctx, const uint8_t *host_bgzf, uint64_t nbytes, uint64_t window,
                                       sq_fastq_stream **out);
/* the device decoder's code run on the host, for unit tests without a GPU (no product path calls it) */
SQ_API int sq_selftest_inflate_host(const uint8_t *deflate, uint32_t len, uint8_t *out, uint32_t cap, uint32_t *out_len);
SQ_API uint64_t sq_batch_size(const sq_batch *b);
SQ_API uint64_t sq_batch_nbytes(const sq_batch *b);
SQ_API uint32_t sq_batch_max_seq_len(const sq_batch *b);
/* copy descriptors (incl. err_sum) / the byte buffer back to the host */
SQ_API int sq_batch_get_metas(sq_batch *b, sq_meta *out);
SQ_API int sq_batch_get_bytes(sq_batch *b, uint8_t *out);
/* FastqRecordArrayView_is_mate (_qcmodule.c:778-850); *first_mismatch = n when mates */
SQ_API int sq_batch_is_mate(sq_batch *a, sq_batch *b, uint64_t *first_mismatch);
SQ_API void sq_batch_free(sq_batch *b);

/* ---- QCMetrics (_qcmodule.c:1966-2139) ---------------------------------- */
typedef struct sq_qc sq_qc;
typedef struct {
    uint64_t number_of_reads;
    uint64_t max_length;
    uint64_t end_anchor_length;
    int32_t bad_phred;       /* 1 when a quality byte outside '!'..'~' was met */
    uint8_t bad_phred_char;  /* that byte ("Not a valid phred character: %c")  */
    uint64_t bad_phred_record; /* global index of the record holding it        */
} sq_qc_info;
SQ_API int sq_qc_create(sq_ctx *ctx, uint64_t end_anchor_length, sq_qc **out);
SQ_API void sq_qc_destroy(sq_qc *m);
SQ_API int sq_qc_add(sq_qc *m, sq_batch *b);
SQ_API int sq_qc_sync(sq_qc *m, sq_qc_info *info);
/* tables sized as the reference's getters (_qcmodule.c:2214-2334); NULL skips */
SQ_API int sq_qc_read(sq_qc *m, uint64_t *base_counts /*max_len*5*/,
                      uint64_t *phred_counts /*max_len*12*/, uint64_t *ea_base /*ea*5*/,
                      uint64_t *ea_phred /*ea*12*/, uint64_t *gc_content /*101*/,
                      uint64_t *phred_scores /*94*/);

/* ---- AdapterCounter (_qcmodule.c:2465-2823) ------------------------------ */
typedef struct sq_adapters sq_adapters;
SQ_API int sq_adapters_create(sq_ctx *ctx, const char *const *adapters, uint64_t n,
                              sq_adapters **out);
SQ_API void sq_adapters_destroy(sq_adapters *a);
SQ_API int sq_adapters_add(sq_adapters *a, sq_batch *b);
SQ_API int sq_adapters_sync(sq_adapters *a, uint64_t *number_of_sequences, uint64_t *max_length);
SQ_API int sq_adapters_read(sq_adapters *a, uint64_t index, uint64_t *forward, uint64_t *reverse);

/* ---- PerTileQuality (_qcmodule.c:3089-3222, 3307-3359) ------------------- */
typedef struct sq_pertile sq_pertile;
typedef struct {
    uint64_t number_of_reads;
    uint64_t max_length;
    uint64_t n_tiles;
    int32_t skipped;         /* header without a tile id was met               */
    uint64_t skipped_record; /* global index of that record                    */
    int32_t bad_phred;
    uint8_t bad_phred_char;
} sq_pertile_info;
SQ_API int sq_pertile_create(sq_ctx *ctx, sq_pertile **out);
SQ_API void sq_pertile_destroy(sq_pertile *p);
SQ_API int sq_pertile_add(sq_pertile *p, sq_batch *b);
SQ_API int sq_pertile_sync(sq_pertile *p, sq_pertile_info *info);
/* name of the record that switched the module off (for skipped_reason) */
SQ_API int sq_pertile_skipped_name(sq_pertile *p, uint8_t *out, uint64_t cap, uint64_t *len);
/* tiles ascending; errors[t*max_len+j] is the ordered sum, counts[..] = reads longer than j */
SQ_API int sq_pertile_read(sq_pertile *p, uint64_t *tile_ids, double *errors, uint64_t *counts);
/* Sharded runs: PerTileQuality's sums (_qcmodule.c:3199-3219) are a chain over the reads of a
 * tile in read order, so the lowest rank holding a tile owns it.  The other ranks copy their
 * records r < limit_records whose tile id is in tile_ids[0..n_ids) (host, ascending) whole and
 * in order to dev_out (DEVICE; NULL = size only); the owner parses that text with
 * sq_batch_from_device_fastq and adds it behind its own reads. */
SQ_API int sq_batch_select_tiles(sq_batch *b, const int64_t *tile_ids, uint64_t n_ids,
                                 uint64_t limit_records, uint8_t *dev_out, uint64_t cap,
                                 uint64_t *nbytes);

/* ---- OverrepresentedSequences (_qcmodule.c:3543-3568, 3830-3942) --------- */
typedef struct sq_overrep sq_overrep;
typedef struct {
    uint64_t number_of_sequences, sampled_sequences, collected_unique_fragments,
        total_fragments, max_unique_fragments, table_size;
    uint64_t warn_records;      /* sampled records holding a non-ACGTN letter   */
    uint64_t first_warn_record; /* global index of the first of them            */
} sq_overrep_info;
SQ_API int sq_overrep_create(sq_ctx *ctx, uint64_t max_unique_fragments, uint32_t fragment_length,
                             uint64_t sample_every, int64_t bases_from_start,
                             int64_t bases_from_end, sq_overrep **out);
SQ_API void sq_overrep_destroy(sq_overrep *o);
SQ_API int sq_overrep_add(sq_overrep *o, sq_batch *b);
SQ_API int sq_overrep_sync(sq_overrep *o, sq_overrep_info *info);
/* the stored fragments as 2-bit k-mers (wanghash64_inverse applied) + counts */
SQ_API int sq_overrep_read(sq_overrep *o, uint64_t *kmers, uint32_t *counts, uint64_t *n);
/* Sharded runs: Sequence_duplication_insert_hash (_qcmodule.c:3543-3568) admits the first
 * max_unique_fragments distinct hashes in read order.  A deferred collector (ranks behind the
 * first; `first_record` = global index of its first read, which fixes the sampling phase :3833)
 * keeps the fragment hashes of its sampled reads; sq_overrep_apply_deferred runs the table
 * maintenance on them once the table of the ranks before it has been loaded
 * (sq_overrep_load_table: DEVICE keys[table_size] u64, counts[table_size] u32 or NULL = zero).
 * sq_overrep_copy_table copies the table out to DEVICE buffers; sq_overrep_set_counters
 * installs the merged additive counters. */
SQ_API int sq_overrep_set_deferred(sq_overrep *o, int deferred, uint64_t first_record);
SQ_API int sq_overrep_apply_deferred(sq_overrep *o);
SQ_API int sq_overrep_copy_table(sq_overrep *o, uint64_t *dev_keys, uint32_t *dev_counts);
SQ_API int sq_overrep_load_table(sq_overrep *o, const uint64_t *dev_keys, const uint32_t *dev_counts,
                                 uint64_t n_unique);
SQ_API int sq_overrep_set_counters(sq_overrep *o, uint64_t number_of_sequences,
                                   uint64_t sampled_sequences, uint64_t total_fragments,
                                   uint64_t warn_records, uint64_t first_warn_record);
/* only fragments seen at least min_count times (the filter of
 * OverrepresentedSequences_overrepresented_sequences, _qcmodule.c:4100-4180),
 * compacted on the device; *n > cap means "call again with room for *n" */
SQ_API int sq_overrep_read_min(sq_overrep *o, uint32_t min_count, uint64_t *kmers, uint32_t *counts,
                               uint64_t cap, uint64_t *n);

/* ---- DedupEstimator (_qcmodule.c:4383-4517) ------------------------------- */
typedef struct sq_dedup sq_dedup;
typedef struct {
    uint64_t modulo_bits, hash_table_size, tracked_sequences;
} sq_dedup_info;
SQ_API int sq_dedup_create(sq_ctx *ctx, uint64_t max_stored_fingerprints, uint64_t front_len,
                           uint64_t back_len, uint64_t front_off, uint64_t back_off,
                           sq_dedup **out);
SQ_API void sq_dedup_destroy(sq_dedup *d);
SQ_API int sq_dedup_add(sq_dedup *d, sq_batch *b);
SQ_API int sq_dedup_add_pair(sq_dedup *d, sq_batch *b1, sq_batch *b2);
SQ_API int sq_dedup_sync(sq_dedup *d, sq_dedup_info *info);
/* counts of the occupied slots in slot order, like duplication_counts() */
SQ_API int sq_dedup_read(sq_dedup *d, uint64_t *counts, uint64_t *n);
/* Sharded runs (one process per GPU, contiguous shards of the read stream): the table of
 * DedupEstimator_add_fingerprint (_qcmodule.c:4426-4460) is order dependent, so the first rank
 * owns it.  A deferred estimator only hashes (sq_dedup_add / sq_fused_add keep the hashes);
 * sq_dedup_deferred_compact keeps those passing the mask of `mod_bits` bits (the owner's
 * _modulo_bits after its own shard -- the mask only grows, :4429-4431) in record order,
 * sq_dedup_deferred_fetch copies them to a caller-owned DEVICE buffer, and the owner feeds what
 * it received through sq_dedup_add_hashes, rank by rank. */
SQ_API int sq_dedup_set_deferred(sq_dedup *d, int deferred);
SQ_API int sq_dedup_deferred_compact(sq_dedup *d, uint64_t mod_bits, uint64_t *n);
SQ_API int sq_dedup_deferred_fetch(sq_dedup *d, uint64_t *dev_out);
SQ_API int sq_dedup_add_hashes(sq_dedup *d, const uint64_t *dev_hashes, uint64_t n);

/* ---- NanoStats (_qcmodule.c:5006-5324) ------------------------------------ */
typedef struct sq_nanostats sq_nanostats;
typedef struct {
    int64_t start_time;
    float duration;
    int32_t channel_id;
    uint32_t length;
    uint32_t reserved;
    double cumulative_error_rate;
    uint64_t parent_id_hash;
} sq_nanoinfo; /* struct NanoInfo, _qcmodule.c:4808-4815 */
typedef struct {
    uint64_t number_of_reads;
    int64_t minimum_time, maximum_time;
    int32_t skipped;
    uint64_t skipped_record;
    int32_t tag_error;        /* malformed BAM aux data (the reference raises, _qcmodule.c:5078-5259): 0 none,
                                 1 "truncated tags", 2 "Invalid type for array %c", 3 "Unknown tag type %c"
                                 (ValueError), 4 "Wrong tag type for '%s' expected '%c' got '%c'" (RuntimeError),
                                 5 ch with a non-integer type (the reference returns NULL with no exception set) */
    uint64_t tag_error_record;
    uint64_t pi_warnings;     /* pi:Z tags that are not 36 characters long        */
    uint32_t tag_error_detail; /* kinds 2, 3: the type character; kind 4: tag (0 st, 1 du, 2 pi) << 8 | type found */
    uint32_t pi_first_length;  /* length of the first such pi tag (the "Counted %zu" of the warning, :5245) */
} sq_nanostats_info;
SQ_API int sq_nanostats_create(sq_ctx *ctx, sq_nanostats **out);
SQ_API void sq_nanostats_destroy(sq_nanostats *s);
SQ_API int sq_nanostats_add(sq_nanostats *s, sq_batch *b);
SQ_API int sq_nanostats_sync(sq_nanostats *s, sq_nanostats_info *info);
SQ_API int sq_nanostats_skipped_name(sq_nanostats *s, uint8_t *out, uint64_t cap, uint64_t *len);
SQ_API int sq_nanostats_read(sq_nanostats *s, sq_nanoinfo *out);

/* ---- InsertSizeMetrics (_qcmodule.c:5571-5744) ---------------------------- */
typedef struct sq_insert sq_insert;
typedef struct {
    uint64_t total_reads, number_of_adapters_read1, number_of_adapters_read2;
    uint64_t max_insert_size, entries_read1, entries_read2;
} sq_insert_info;
SQ_API int sq_insert_create(sq_ctx *ctx, uint64_t max_adapters, sq_insert **out);
SQ_API void sq_insert_destroy(sq_insert *m);
SQ_API int sq_insert_add_pair(sq_insert *m, sq_batch *b1, sq_batch *b2);
SQ_API int sq_insert_sync(sq_insert *m, sq_insert_info *info);
SQ_API int sq_insert_read_sizes(sq_insert *m, uint64_t *sizes /* max_insert_size+1 */);
/* adapters of read `which` (0/1): seqs[i*32] = length, seqs[i*32+1..] = bytes */
SQ_API int sq_insert_read_adapters(sq_insert *m, int which, uint8_t *seqs, uint64_t *counts,
                                   uint64_t *n);

/* ---- fused add ------------------------------------------------------------ */
/* One call for the whole hot loop body of src/sequali/__main__.py:280-306:
 * equivalent to calling the *_add of every non-NULL collector on `b` in the
 * reference's order (QCMetrics, PerTileQuality, OverrepresentedSequences,
 * NanoStats, AdapterCounter, DedupEstimator), but short-read FASTQ arrays are
 * walked once for all per-read quantities instead of once per collector. */
SQ_API int sq_fused_add(sq_ctx *ctx, sq_batch *b, sq_qc *qc, sq_pertile *pt, sq_overrep *ov,
                        sq_nanostats *ns, sq_adapters *ad, sq_dedup *dd);

/* ---- sharded runs: one process per GPU, NCCL over NVLink ---- */
/* The reference is single process; these entry points have no counterpart in it.  Every rank owns a
 * contiguous shard of the read stream; after the pass the tables are merged so that they equal one
 * sequential pass over all shards (protocol: sequali_b200/sharded.py).  All collectives run on the
 * context's launch stream, behind the collectors' kernels.  NCCL is dlopen'ed on first use. */
typedef struct sq_comm sq_comm;
SQ_API int sq_comm_unique_id(uint8_t *out128); /* rank 0 makes it, the other ranks get it out of band */
SQ_API int sq_comm_create(sq_ctx *ctx, const uint8_t *id128, int rank, int world, sq_comm **out);
SQ_API void sq_comm_destroy(sq_comm *c);
SQ_API int sq_comm_rank(const sq_comm *c);
SQ_API int sq_comm_world(const sq_comm *c);
/* device buffers; op: 0 sum, 1 max, 2 min */
SQ_API int sq_comm_allreduce_u64(sq_comm *c, uint64_t *dev, uint64_t n, int op);
SQ_API int sq_comm_allreduce_u32(sq_comm *c, uint32_t *dev, uint64_t n, int op);
SQ_API int sq_comm_bcast(sq_comm *c, void *dev, uint64_t nbytes, int root);
SQ_API int sq_comm_send(sq_comm *c, const void *dev, uint64_t nbytes, int dst);
SQ_API int sq_comm_recv(sq_comm *c, void *dev, uint64_t nbytes, int src);
SQ_API int sq_comm_group_start(sq_comm *c);
SQ_API int sq_comm_group_end(sq_comm *c);
/* small host payloads, staged through pinned memory; these synchronise */
SQ_API int sq_comm_allreduce_host_u64(sq_comm *c, uint64_t *host, uint64_t n, int op);
SQ_API int sq_comm_bcast_host(sq_comm *c, void *host, uint64_t nbytes, int root);
SQ_API int sq_comm_allgather_host(sq_comm *c, const void *host_in, void *host_out, uint64_t nbytes);
SQ_API int sq_comm_barrier(sq_comm *c);
/* additive tables (QCMetrics _qcmodule.c:1786-1803, AdapterCounter :2912-2914): ncclAllReduce(sum, u64) in
 * place on the device tables, max_length by all-reduce(max); afterwards every rank's collector answers
 * its getters with the merged tables */
SQ_API int sq_qc_allreduce(sq_qc *m, sq_comm *c);
SQ_API int sq_adapters_allreduce(sq_adapters *a, sq_comm *c);
/* NanoStats (:5314-5322): per-read records of all ranks in rank (= read) order, cut at the first header
 * of any rank that cannot be parsed; first_record = global index of this rank's first read */
SQ_API int sq_nanostats_allgather(sq_nanostats *s, sq_comm *c, uint64_t first_record);
/* InsertSizeMetrics (:5571-5611, 5883-5885): histogram and counters are sums (sq_insert_allreduce); the two
 * capped adapter tables admit first come, first served, so the first rank owns them: a deferred collector
 * keeps its adapter occurrences (32-byte key = length byte + <= 31 adapter bytes, and the hash) in pair
 * order, the owner feeds them through sq_insert_add_keys rank by rank. */
SQ_API int sq_insert_set_deferred(sq_insert *m, int deferred);
SQ_API int sq_insert_deferred_count(sq_insert *m, int which, uint64_t *n);
SQ_API int sq_insert_deferred_fetch(sq_insert *m, int which, uint8_t *dev_keys, uint64_t *dev_hashes);
SQ_API int sq_insert_add_keys(sq_insert *m, int which, const uint8_t *dev_keys, const uint64_t *dev_hashes, uint64_t n);
SQ_API int sq_insert_allreduce(sq_insert *m, sq_comm *c);
/* stream-ordered device memory for exchange buffers (hash lists, table copies, border-tile text) */
SQ_API void *sq_stream_alloc(sq_ctx *ctx, uint64_t nbytes);
SQ_API void sq_stream_free(sq_ctx *ctx, void *p);
SQ_API int sq_stream_memset(sq_ctx *ctx, void *p, int value, uint64_t nbytes);

/* ---- synthetic input (bench / tests; SURVEY.md 8d recipe C2) -------------- */
/* Fill dev_text with records `first_read .. first_read + n_reads` of a NovaSeq-style stream generated on
 * the device: read g belongs to tile number (g / reads_per_tile) % 936 of the flow cell (tiles in runs,
 * as in real files).  Returns the number of bytes written (<= cap) in *nbytes. */
SQ_API int sq_synth_illumina(sq_ctx *ctx, uint8_t *dev_text, uint64_t cap, uint64_t n_reads,
                             uint32_t read_length, uint64_t seed, uint64_t first_read,
                             uint64_t reads_per_tile, uint64_t *nbytes);

/* ---- report aggregation on device-resident tables (reference report_modules.py) ----------------------- */
typedef struct {
    uint64_t total_bases;            /* sum(base_count_tables), :620 */
    uint64_t minimum_length;         /* qc_metrics_modules :2563-2566 */
    uint64_t n50, n90;               /* :621-634 */
    uint64_t threshold_lengths[16];  /* the percentile walk of :598-617, one length per count threshold */
} sq_qc_length_summary;
/* aggregate_count_matrix (report_modules.py:307-322) of base_count_table and phred_count_table over the data
 * ranges [starts[k], stops[k]) (positions; clipped to max_length like a Python slice), and
 * SequenceLengthDistribution.from_base_count_tables (:575-637) without moving the tables to the host:
 * base_matrix[n_ranges][5], phred_matrix[n_ranges][12], length_counts[k] = reads with starts[k] < length <=
 * stops[k]; summary->threshold_lengths[j] = the first length at which more than count_thresholds[j] reads (of
 * at least one letter) are at most that long, 0 if there is none (the caller passes int(p * total / 100) for
 * its percentiles; n_thresholds <= 16); total_sequences = QCMetrics.number_of_reads. */
SQ_API int sq_qc_aggregate(sq_qc *m, const uint64_t *starts, const uint64_t *stops, uint64_t n_ranges,
                           uint64_t *base_matrix, uint64_t *phred_matrix, uint64_t *length_counts,
                           const uint64_t *count_thresholds, uint64_t n_thresholds,
                           uint64_t total_sequences, sq_qc_length_summary *summary);
typedef struct {
    int32_t kind;     /* 0 none; 1 / 2: length / duration is infinite / NaN (Python: OverflowError / ValueError
                         from round()); 3: a list index out of range (IndexError) */
    uint64_t record;  /* the first read it happens at */
} sq_nano_report_error;
/* NanoStatsReport.from_nanostats (report_modules.py:1952-2046) over the per-read records on the device.  The
 * caller computes run_start = minimum_time, interval and n_slots like :1968-1980.  time_bases / time_reads /
 * time_active_channels [n_slots], time_qualities [n_slots][12] (class = min(round(-10 log10(error / length)),
 * 47) >> 2 with the host's log10), translocation_speeds [81], *reads_with_parent; *n_channels = distinct
 * channel ids, whose sorted ids / base totals / error sums (added in read order) are then fetched with
 * sq_nanostats_report_channels. */
SQ_API int sq_nanostats_report(sq_nanostats *s, int64_t run_start, int64_t interval, uint64_t n_slots,
                               uint64_t *time_bases, uint64_t *time_reads, uint64_t *time_active_channels,
                               uint64_t *time_qualities, uint64_t *translocation_speeds,
                               uint64_t *reads_with_parent, uint64_t *n_channels,
                               sq_nano_report_error *error);
SQ_API int sq_nanostats_report_channels(sq_nanostats *s, int32_t *channel_ids, uint64_t *bases,
                                        double *cumulative_error, uint64_t cap);

/* ---- _seqident (reference _seqidentmodule.c) ---------------------------------------------------------- */
/* sequence_identity (_seqidentmodule.c:32-101, 279-343) for n (target, query) pairs in one launch: pair p is
 * targets[target_off[p] .. target_off[p+1]) against queries[query_off[p] .. query_off[p+1]) (at most 31 letters,
 * else SQ_E_ARG; host buffers).  matches_out[p] = the reference's most_matches: the largest count of matched query
 * letters among the Smith-Waterman cells with the highest score; identity = matches / len(query).  Scores are
 * added as 32-bit integers (the reference's portable loop uses Py_ssize_t, its AVX2 loop 8-bit lanes: equal
 * whenever the 8-bit lanes do not overflow, e.g. for the default scores).  Synchronous. */
SQ_API int sq_sequence_identity_batch(sq_ctx *ctx, const uint8_t *targets, const uint64_t *target_off,
                                      const uint8_t *queries, const uint32_t *query_off, uint64_t n,
                                      int match_score, int mismatch_penalty, int deletion_penalty,
                                      int insertion_penalty, int32_t *matches_out);

#ifdef __cplusplus
}
#endif
#endif /* SQGPU_H */
