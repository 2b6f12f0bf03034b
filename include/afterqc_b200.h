/*
 * afterqc_b200.h -- C-ABI of the B200-native AfterQC per-read hot path.
 *
 * This is the drop-in boundary.  The reference has exactly one FFI precedent:
 * util.py:4-24 loads editdistance/libed.so with ctypes and calls two extern "C" symbols with
 * caller-owned buffers, POD arguments and sentinel/int returns (_editdistance.h:16,23).  This
 * header keeps those conventions (extern "C", plain pointers and sizes, int return codes, no
 * exceptions, no allocation ownership crossing the boundary except the explicit pinned-host
 * helpers) and replaces the *Python* per-read loop instead of only seek_overlap:
 *
 *   aqc_stat_reads     replaces QualityControl.statFile/statRead      qualitycontrol.py:73-122,331-357
 *   aqc_filter_pairs   replaces the seqFilter.run() loop body          preprocesser.py:411-631
 *                      (trim :455-466, length :476-479, hasPolyX :30-51/:482-490,
 *                       lowQualityNum :61-68/:493-501, nNumber :70-76/:504-512,
 *                       util.overlap -> overlap_hm util.py:88-89,158-212,
 *                       adapter trim + rescan :516-541, correction walk :542-617,
 *                       postfilter statRead :624-627, counters :378-409)
 *   aqc_ops_pairs      the bare per-read operators (hasPolyX / lowQualityNum / nNumber /
 *                      util.overlap) for operator-level parity tests
 *   aqc_get_*          fetch the integer counter blocks the JSON writer (preprocesser.py:660-778)
 *                      and QualityControl.qc() (qualitycontrol.py:124-156) consume
 *
 * The same structs and enums are used by the CPU oracle (oracle/aqc_oracle.c, symbols aqo_*),
 * which is test infrastructure and never linked into the product.
 *
 * Reads are passed as packed SoA batches: one byte column for bases, one for qualities and an
 * n+1 offsets column per mate (variable length).  Qualities share the offsets of the bases
 * (a FASTQ record whose quality line length differs from its sequence length is rejected by
 * the host parser; the reference tolerates it only inside statRead, qualitycontrol.py:81-87).
 */
#ifndef AFTERQC_B200_H
#define AFTERQC_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AQC_ABI_VERSION 2
#define AQC_MAX_LEN 1000      /* qualitycontrol.py:23  MAX_LEN; longer reads raise IndexError there */
#define AQC_MAX_KMER 8        /* dense 4^k table + 64-bit raw-byte keys for non-ACGT k-mers; --qc_kmer outside 1..8 is rejected */
#define AQC_NUM_QC 4          /* r1 prefilter, r2 prefilter, r1 postfilter, r2 postfilter */

/* error codes (0 = ok); aqc_last_error() gives text */
enum {
    AQC_OK = 0,
    AQC_ERR_INVALID = 1,        /* bad argument / parameter out of the supported domain */
    AQC_ERR_CUDA = 2,           /* CUDA runtime failure */
    AQC_ERR_TOO_LONG = 3,       /* a read longer than AQC_MAX_LEN (reference: IndexError) */
    AQC_ERR_KMER_TABLE_FULL = 4,/* non-ACGT k-mer side table overflow; raise kmer_side_log2 */
    AQC_ERR_NOMEM = 5,
    AQC_ERR_TOO_SHORT_STAT = 6  /* a read of 1..4 bases reached statRead (reference: IndexError,
                                   qualitycontrol.py:97-108 indexes seq[j+1] up to 4) */
};

/* where the batch columns and the result array live */
enum { AQC_MEM_HOST = 0, AQC_MEM_DEVICE = 1 };

/* pair classes in the reference's priority order (preprocesser.py:436-614); the numeric
 * value is what aqc_result.cls holds.  BADBCD1/2 and BADBBL are out of scope (barcode, debubble). */
enum {
    AQC_GOOD = 0,
    AQC_BADTRIM1 = 1,     /* :457-460 */
    AQC_BADTRIM2 = 2,     /* :463-466 */
    AQC_BADLEN = 3,       /* :476-479 and :529-532 */
    AQC_BADPOL = 4,       /* :487-490 */
    AQC_BADLQC = 5,       /* :498-501 */
    AQC_BADNCT = 6,       /* :509-512 */
    AQC_BADDIFF = 7,      /* :538-541 */
    AQC_BADMISMATCH = 8,  /* :611-614 */
    AQC_NUM_CLASSES = 9
};

/* Filter parameters = the resolved option values the loop reads (after.py:14-93; trim values
 * already resolved by autoTrim, preprocesser.py:260-280, so they are >= 0 here). */
typedef struct aqc_params {
    int32_t paired;                  /* read2_file != None */
    int32_t trim_front, trim_tail;   /* -f / -t (R1) */
    int32_t trim_front2, trim_tail2; /* R2 values (trim_pair_same copies R1's) */
    int32_t seq_len_req;             /* -s 35 */
    int32_t poly_size_limit;         /* -p 35 */
    int32_t allow_mismatch_in_poly;  /* -a 2 */
    int32_t qualified_quality_phred; /* -q 15 */
    int32_t unqualified_base_limit;  /* -u 60 */
    int32_t n_base_limit;            /* -n 5 */
    int32_t no_overlap;              /* --no_overlap */
    int32_t no_correction;           /* --no_correction */
    int32_t mask_mismatch;           /* --mask_mismatch */
    int32_t qc_sample;               /* --qc_sample 200000 (postfilter gate, preprocesser.py:624) */
    int32_t qc_kmer;                 /* --qc_kmer 8 */
    int32_t kmer_side_log2;          /* log2 capacity of the non-ACGT k-mer side table (0 = default 20) */
    int32_t filter_kernel;           /* which kernel aqc_filter_pairs launches: 0 = default and 2 = lane-per-pair (lane_kernel, then
                                        pair_kernel's list mode for pairs holding a byte outside A,C,G,T,N, then stat_kernel for the
                                        sampled statistics) for batches whose reads are <= 256 bases, pair_kernel otherwise;
                                        1 = warp-per-pair (pair_kernel) for every batch.  Results are identical; a performance knob. */
    int32_t stat_kernel;             /* how statRead (qualitycontrol.py:73-122) is executed in aqc_stat_reads (and after lane_kernel):
                                        0 = stat_kernel (one warp per read, every histogram in shared memory, aqc_stat_kernel.cuh) for
                                        batches whose reads are <= 256 bases, stat_read inside pair_kernel otherwise; 1 = stat_read
                                        always (with filter_kernel = 1 this is the whole round-1 path).  Results are identical. */
    int32_t reserved[5];
} aqc_params;

/* One packed batch.  off*[i]..off*[i+1] delimit record i in seq* and qual*.  seq2/qual2/off2
 * are NULL for single-end input.  first_index is the 0-based global record index of record 0
 * (the reference's TOTAL_READS for record i is first_index + i + 1).  Device-side columns need
 * 16 readable bytes of slack after the last base (bulk copies are 16-byte granular). */
typedef struct aqc_batch {
    uint64_t first_index;
    uint32_t n;
    uint32_t flags;             /* bits 0-15: optional hint, the longest read of the batch (0 = unknown);
                                   AQC_BATCH_QUAL2_IN_PLACE: see below; other bits 0 */
    const uint8_t *seq1, *qual1;
    const uint32_t *off1;
    const uint8_t *seq2, *qual2;
    const uint32_t *off2;
} aqc_batch;

/* aqc_batch.flags, AQC_MEM_HOST batches of aqc_filter_pairs only: qual2 lives in page-locked host memory that the device
 * can address (aqc_host_alloc / cudaHostAlloc / cudaHostRegister).  The loop reads qualities of mate 2 only in the correction
 * walk (two bytes per visited mismatch, preprocesser.py:566-567) and in the sampled statRead (:624-627), so with the
 * lane-per-pair kernel the engine may leave that column where it is and let the kernel fetch those bytes over PCIe instead of
 * copying a quarter of the batch.  The engine checks the pointer (cudaPointerGetAttributes) and silently copies as usual when
 * the condition does not hold.  Results are identical either way. */
#define AQC_BATCH_QUAL2_IN_PLACE (1u << 16)
/* Per-pair outcome, 32 bytes.  start/len are the final coordinates into the ORIGINAL read after
 * front/tail trim and adapter cut (what the good/bad writer must emit).  edits are the byte
 * changes of the correction walk (preprocesser.py:563-598), applied in order:
 *   bits 0-9 pos (index into the ORIGINAL read), bits 10-11 kind, bits 16-23 base, bits 24-31 qual
 *   kind 0: r1[pos] := base, r1q[pos] := qual     (:583-584)
 *   kind 1: r2[pos] := base, r2q[pos] := qual     (:575-576)
 *   kind 2: mask_mismatch -- r1q[pos] := '!' and the mate position r2q[pos2] := '!' where pos2 is
 *           carried in bits 16-25 instead of base/qual                     (:590-592)
 *   kind 3: mismatch skipped (no byte change; kept so the host can count)  (:595)
 * ov_* is the result of the last util.overlap call for the pair (after the adapter rescan). */
typedef struct aqc_result {
    uint8_t cls;
    uint8_t n_edits;
    uint16_t start1, len1, start2, len2;
    int16_t ov_offset;
    uint16_t ov_len, ov_diff;
    uint32_t edits[4];
} aqc_result;

#define AQC_EDIT_POS(e) ((e) & 0x3FFu)
#define AQC_EDIT_KIND(e) (((e) >> 10) & 3u)
#define AQC_EDIT_BASE(e) (((e) >> 16) & 0xFFu)
#define AQC_EDIT_QUAL(e) (((e) >> 24) & 0xFFu)
#define AQC_EDIT_POS2(e) (((e) >> 16) & 0x3FFu)

/* Bare operator outputs on the (front/tail-trimmed) mates, 32 bytes; no classification. */
typedef struct aqc_ops {
    uint8_t poly1, poly2;       /* hasPolyX return char, 0 = None            preprocesser.py:30-51 */
    uint16_t lowq1, lowq2;      /* lowQualityNum                             :61-68 */
    uint16_t n1, n2;            /* nNumber                                   :70-76 */
    uint16_t len1, len2;        /* trimmed lengths                           :19-28 */
    int16_t ov_offset;          /* util.overlap(r1, r2) first call           util.py:158-212 */
    uint16_t ov_len, ov_diff;
    uint8_t pad[12];
} aqc_ops;

/* Scalar counter block (int64), indices; names follow preprocesser.py:378-409. */
enum {
    AQC_C_TOTAL_READS = 0,
    AQC_C_TOTAL_BASES_R1, AQC_C_TOTAL_BASES_R2,   /* raw lengths; host applies quirk Q1 (:416,:431) */
    AQC_C_GOOD_READS,
    AQC_C_GOOD_BASES_R1, AQC_C_GOOD_BASES_R2,     /* final lengths of good pairs (:621-623) */
    AQC_C_BADTRIM1, AQC_C_BADTRIM2, AQC_C_BADLEN, AQC_C_BADPOL, AQC_C_BADLQC, AQC_C_BADNCT,
    AQC_C_BADDIFF, AQC_C_BADMISMATCH,
    AQC_C_READ_CORRECTED, AQC_C_BASE_CORRECTED, AQC_C_BASE_SKIPPED_CORRECTION, AQC_C_BASE_ZERO_QUAL_MASKED,
    AQC_C_OVERLAPPED, AQC_C_OVERLAP_LEN_SUM, AQC_C_OVERLAP_BASE_SUM, AQC_C_OVERLAP_BASE_ERR,
    AQC_C_TRIMMED_ADAPTER_BASE, AQC_C_TRIMMED_ADAPTER_READ,
    AQC_C_ERR_MATRIX = 32,      /* 16 cells [correct][error], base order A,T,C,G (ALL_BASES, qualitycontrol.py:24) */
    AQC_C_SCALARS = 64,
    AQC_C_OVERLAP_HIST = 64,                          /* [AQC_MAX_LEN+1]  overlap_histgram  :517 */
    AQC_C_DISTANCE_HIST = 64 + AQC_MAX_LEN + 1,       /* [AQC_MAX_LEN+1]  distance_histgram :536 */
    AQC_C_TOTAL = 64 + 2 * (AQC_MAX_LEN + 1)
};

/* QC slots */
enum { AQC_QC_R1_PRE = 0, AQC_QC_R2_PRE = 1, AQC_QC_R1_POST = 2, AQC_QC_R2_POST = 3 };

/* Per-cycle integer counters of one QualityControl object (qualitycontrol.py:33-57), each
 * AQC_MAX_LEN long, base order A,T,C,G. */
typedef struct aqc_qc_counters {
    int64_t totalNum[AQC_MAX_LEN];
    int64_t totalQual[AQC_MAX_LEN];
    int64_t baseCounts[4][AQC_MAX_LEN];
    int64_t baseTotalQual[4][AQC_MAX_LEN];
    int64_t totalDiscontinuity[AQC_MAX_LEN];
    int64_t gcHistogram[AQC_MAX_LEN + 1];
    int64_t totalKmer;
    int64_t reads;              /* number of reads stat'd into this object */
} aqc_qc_counters;

/* K-mer first-seen key: (order_index << 11) | (pos << 1) | seeded_by_revcomp.  Sorting the
 * entries that have first != AQC_KMER_NEVER by (count desc, first asc) reproduces
 * sorted(kmerCount.items(), key=count, reverse=True) over an insertion-ordered dict
 * (qualitycontrol.py:113-122,155-156; quirk Q12). */
#define AQC_KMER_NEVER 0xFFFFFFFFFFFFFFFFull

typedef struct aqc_ctx aqc_ctx;

/* ---- lifecycle ---- */
int aqc_abi_version(void);
int aqc_device_count(void);                        /* CUDA devices visible to the process (directory mode spreads its jobs over them, after.py:168-171) */
/* device < 0: the current CUDA device. */
int aqc_create(int device, const aqc_params *params, aqc_ctx **out);
void aqc_destroy(aqc_ctx *ctx);
int aqc_set_params(aqc_ctx *ctx, const aqc_params *params);
int aqc_reset(aqc_ctx *ctx);                       /* zero all counters / QC objects */
int aqc_reset_filter(aqc_ctx *ctx);                /* zero the scalar block, both histograms and the two
                                                      postfilter QC slots; prefilter slots are kept */
const char *aqc_last_error(const aqc_ctx *ctx);    /* ctx may be NULL: last create error */

/* pinned host memory for async copies (optional; pageable host buffers also work) */
int aqc_host_alloc(size_t bytes, void **out);
void aqc_host_free(void *p);
/* plain device buffers (so callers can keep batches resident in HBM) */
int aqc_device_alloc(aqc_ctx *ctx, size_t bytes, void **out);
void aqc_device_free(aqc_ctx *ctx, void *p);
int aqc_memcpy_h2d(aqc_ctx *ctx, void *dst, const void *src, size_t bytes);
int aqc_memcpy_d2h(aqc_ctx *ctx, void *dst, const void *src, size_t bytes);

/* ---- the hot path ---- */
/* Prefilter statistics: statRead for every record i with stat_lo <= first_index+i < stat_hi;
 * mate 1 goes to QC slot qc1, mate 2 (if present) to qc2 (pass -1 to skip a mate).  The k-mer
 * order index of record i is order_base + (first_index + i - stat_lo). */
int aqc_stat_reads(aqc_ctx *ctx, const aqc_batch *batch, int mem, int qc1, int qc2,
                   uint64_t stat_lo, uint64_t stat_hi, uint64_t order_base);

/* The fused per-pair loop body.  results: n entries in the same memory space as the batch.
 * Accumulates the scalar counters, both histograms, the error matrix and the postfilter QC
 * slots (gated by qc_sample exactly as preprocesser.py:624). */
int aqc_filter_pairs(aqc_ctx *ctx, const aqc_batch *batch, int mem, aqc_result *results);

/* Operator-level outputs (no counters touched). */
int aqc_ops_pairs(aqc_ctx *ctx, const aqc_batch *batch, int mem, aqc_ops *out);

/* Use an externally owned CUDA stream (cudaStream_t passed as void*) for all kernels and device-side
 * work of this context (e.g. the caller's current stream, so that the caller's events time it).
 * NULL restores the context's own stream. */
int aqc_set_stream(aqc_ctx *ctx, void *cuda_stream);

/* Device addresses of the accumulator blocks, for in-place collectives (NCCL) across shards.
 * what: 0 = scalar/histogram counter block (int64[AQC_C_TOTAL]);
 *       1..10 = arrays of QC slot `slot`: 1 cls_cnt[5][MAX_LEN], 2 cls_qsum[5][MAX_LEN] (raw byte sums),
 *       3 disc[MAX_LEN], 4 gchist[MAX_LEN+1], 5 scal[2], 6 dense k-mer counts[4^k], 7 dense k-mer first[4^k] (MIN-reduce),
 *       8 side keys, 9 side counts, 10 side first-direct, 11 side first-seed   (side tables are not reducible in place)
 * *n_out receives the number of 64-bit elements. */
int aqc_device_ptr(aqc_ctx *ctx, int what, int slot, void **ptr_out, uint64_t *n_out);

/* Wait for all queued work of this context. */
int aqc_sync(aqc_ctx *ctx);

/* ---- fetch (host pointers; these synchronise) ---- */
int aqc_get_counters(aqc_ctx *ctx, int64_t *out /* AQC_C_TOTAL */);
int aqc_add_counters(aqc_ctx *ctx, const int64_t *in /* AQC_C_TOTAL */);   /* merge another shard's block */
int aqc_get_qc(aqc_ctx *ctx, int slot, aqc_qc_counters *out);
/* dense table: 4^k entries, index = sum code(base_j) << 2*(k-1-j), code A=0 C=1 G=2 T=3 */
int aqc_get_kmer_dense(aqc_ctx *ctx, int slot, uint64_t *counts, uint64_t *first);
/* side table of k-mers with a non-ACGT byte: keys are the k raw bytes, first byte in the
 * most-significant position of a k-byte big-endian integer. *n_out receives the entry count
 * (AQC_ERR_INVALID if cap is too small; call with cap = 0 to query). */
int aqc_get_kmer_side(aqc_ctx *ctx, int slot, uint64_t *keys, uint64_t *counts, uint64_t *first,
                      uint32_t cap, uint32_t *n_out);

/* The side table as the device holds it, unresolved: per key the count, the first DIRECT sighting and the first seeding by a
 * k-mer that holds a byte outside util.COMP (AQC_KMER_NEVER = none).  Shards merge these by key (sum, min, min) and then apply
 * the insertion rule of aqc_get_kmer_side on the merged table (afterqc_b200/multigpu.py: resolve_side), which makes the
 * first-seen order of sharded runs exact for every byte (quirk Q12).  Entries include keys that were only ever seeded. */
int aqc_get_kmer_side_raw(aqc_ctx *ctx, int slot, uint64_t *keys, uint64_t *counts, uint64_t *first_direct, uint64_t *first_seed,
                          uint32_t cap, uint32_t *n_out);

/* ---- host-side FASTQ ingest / egress on the packed columns (no GPU work; replaces fastq.py:17-104) ---- */
/* Parse complete 4-line records of buf[0..n): lines are rstrip()'d, the first empty line ends the file (its record is
 * dropped), a trailing partial record stays unconsumed unless `final`.  Column k (0 names, 1 bases, 2 '+' lines,
 * 3 qualities) is appended to out_bytes[k] (capacity >= n) with offsets out_off[k][0..records] (caller sets [0]).
 * AQC_ERR_INVALID + *bad_record when a quality line's length differs from its sequence line's. */
int aqc_fastq_parse(const uint8_t *buf, uint64_t n, int final, uint64_t max_records,
                    uint8_t *const out_bytes[4], uint64_t *const out_off[4],
                    uint64_t *n_records, uint64_t *consumed, int *hit_eof, uint64_t *bad_record);
/* FASTQ text of one mate's records selected by `which` (0 good, 1 bad with the "@BADxxx" name prefix of
 * preprocesser.py:212-213, 2 overlapped tails for --store_overlap :615-617): final slice + correction edits of
 * results[i] applied to column record rec_base + i; mate 0 = index read (-7/-5 files), passed through whole.
 * AQC_ERR_NOMEM when out_cap is too small. */
int aqc_fastq_emit(int mate, int which,
                   const uint8_t *names, const uint64_t *name_off, const uint8_t *seqs, const uint64_t *seq_off,
                   const uint8_t *plus, const uint64_t *plus_off, const uint8_t *quals,
                   uint64_t rec_base, const aqc_result *results, uint64_t n,
                   uint8_t *out, uint64_t out_cap, uint64_t *out_len);

/* ---- the same parse on the device (fastq.Reader.nextRead, fastq.py:37-49, for a whole buffer of text): newline index, line
 * table, validation and the packed base / quality columns of one mate, built in HBM by csrc/aqc_parse.cuh.  Semantics are
 * aqc_fastq_parse's (rstrip()ped lines, the first empty line ends the file, a trailing partial record stays unconsumed,
 * AQC_ERR_INVALID + bad_record when a quality line's length differs from its sequence line's; records before it are valid).
 * text: AQC_MEM_HOST (copied inside) or AQC_MEM_DEVICE; below 4 GiB.  The pointers of *out are DEVICE pointers owned by the
 * context, valid until the next call with the same slot (0 or 1: the two mates of a pair can be held at once); seq / qual /
 * off are directly the columns of one mate of an aqc_batch for AQC_MEM_DEVICE calls; line_start / line_len index `text`
 * (names and '+' lines stay text: only the writers need them). */
typedef struct aqc_parsed {
    uint64_t n_records;            /* complete records in the columns */
    uint64_t consumed;             /* bytes of the text they cover: the next call continues here */
    uint64_t bad_record;           /* with AQC_ERR_INVALID */
    uint64_t seq_bytes;            /* bytes in each of seq / qual */
    int32_t hit_eof, reserved;
    const uint8_t *seq, *qual;     /* packed, 16-byte aligned, 16 bytes of slack */
    const uint32_t *off;           /* [n_records + 1] */
    const uint32_t *line_start, *line_len;   /* [4 * n_records]: name, bases, '+', qualities of every record */
    const uint8_t *text;           /* the device copy of the text (or the caller's device pointer) the line table refers to */
} aqc_parsed;
int aqc_fastq_parse_device(aqc_ctx *ctx, int slot, const uint8_t *text, uint64_t n, int mem, int final, uint64_t max_records,
                           aqc_parsed *out);

/* aqc_fastq_emit for records that were parsed on the device: the four lines of record r are text[line_start[4r + k] ..
 * + line_len[4r + k]) (host copies of the text and of aqc_parsed's line table). */
int aqc_fastq_emit_lines(int mate, int which, const uint8_t *text, const uint32_t *line_start, const uint32_t *line_len,
                         uint64_t rec_base, const aqc_result *results, uint64_t n,
                         uint8_t *out, uint64_t out_cap, uint64_t *out_len);

/* ---- barcode (UMI) pre-pass on packed columns (host only; barcodeprocesser.py, preprocesser.py:435-452) ---- */
#define AQC_HOST_BADBCD1 16   /* host-only pseudo classes for aqc_fastq_emit: pairs rejected before the device loop */
#define AQC_HOST_BADBCD2 17
typedef struct aqc_columns {       /* one mate's records; qualities share seq_off; offsets may be absolute */
    const uint8_t *names; const uint64_t *name_off;
    const uint8_t *seqs; const uint64_t *seq_off;
    const uint8_t *quals;
} aqc_columns;
typedef struct aqc_columns_out {
    uint8_t *names; uint64_t *name_off;
    uint8_t *seqs; uint64_t *seq_off;
    uint8_t *quals;
} aqc_columns_out;
/* detectBarcode on both mates, barcode moved into the name, barcode + verify bases and (pairs) the read-through tail of
 * cleanBarcodeTail removed.  Records keep the input order; status[i] = 0 ok, 1 BADBCD1, 2 BADBCD2 (copied unchanged).
 * in2/out2 NULL = single-end (strips the design length, preprocesser.py:443-444).  removed[m] = bases dropped from mate
 * m+1 of the status-0 records.  Capacity: names in-bytes + n*(barcode_length+3); bases/qualities in-bytes. */
int aqc_barcode_pairs(int barcode_length, const char *verify, uint64_t n, const aqc_columns *in1, const aqc_columns *in2,
                      aqc_columns_out *out1, aqc_columns_out *out2, uint8_t *status, uint64_t removed[2]);

/* ---- streaming reader: one FASTQ file -> packed record batches, parsed by a background thread into a ring of
 * reusable buffers (replaces fastq.Reader, fastq.py:17-55; plain and .gz by file extension, :23-28; concatenated gzip
 * members are read through).  Batches hold exactly `batch_records` records except the last; the columns of a batch stay
 * valid until aqc_reader_release(slot).  bytes[1]/bytes[3] (bases/qualities) share the offsets off[1] (equal lengths are
 * checked), carry >= 64 readable bytes after the data, and seq_off32 is off[1] as uint32: together they are the
 * aqc_batch columns of one mate, without a copy. */
typedef struct aqc_reader aqc_reader;
typedef struct aqc_records {
    uint64_t n;               /* records in this batch; 0 = end of file (no slot to release) */
    uint64_t first_index;     /* index in the file of record 0 of the batch */
    uint32_t slot;            /* pass to aqc_reader_release */
    uint32_t max_len;         /* longest sequence in the batch */
    const uint8_t *bytes[4];  /* names, bases, '+' lines, qualities */
    const uint64_t *off[4];   /* n + 1 offsets per column */
    const uint32_t *seq_off32;
} aqc_records;
int aqc_reader_open(const char *path, uint64_t batch_records, uint32_t slots, aqc_reader **out);
/* blocks until the next batch is parsed.  AQC_ERR_INVALID (text in aqc_reader_error) on an unreadable / corrupt file or a
 * record whose quality length differs from its sequence length; batches before the bad record are delivered first. */
int aqc_reader_next(aqc_reader *r, aqc_records *out);
int aqc_reader_release(aqc_reader *r, uint32_t slot);
const char *aqc_reader_error(const aqc_reader *r);
/* the reader's own gzip/DEFLATE decoder on a memory buffer (all members; every member's CRC-32 and length verified):
 * AQC_ERR_INVALID + text in err on a corrupt or truncated stream, AQC_ERR_NOMEM when out_cap is too small. */
int aqc_gunzip_buffer(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t out_cap, uint64_t *out_len, char *err, uint64_t err_cap);
/* ... and its multi-threaded form (speculative block search + symbolic windows, exact chaining; csrc/aqc_pinflate.hpp).
 * stats receives rounds, accepted pieces, rejected search hits. */
int aqc_gunzip_buffer_mt(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t out_cap, uint64_t *out_len, int threads,
                         uint64_t stats[3], char *err, uint64_t err_cap);
void aqc_reader_close(aqc_reader *r);
/* The reader's sources without its parser (plain files, .gz through the multi-threaded decoder): text for callers that parse
 * elsewhere -- the device parser.  aqc_text_read fills up to cap bytes (cap >= 4 MiB; whole 4 MiB blocks): 0 = end of file,
 * -1 = error (aqc_reader_error).  Close with aqc_reader_close. */
int aqc_text_open(const char *path, aqc_reader **out);
int64_t aqc_text_read(aqc_reader *r, uint8_t *dst, uint64_t cap);

/* ---- Levenshtein distance of n string pairs, bit-parallel on the GPU (one lane per pair): replaces util.editDistance
 * (util.py:65-83) -> edit_distance (editdistance/_editdistance.cpp:100-126) for batches.  a / b: byte columns with n + 1
 * uint32 offsets each; any byte is a valid character; out[i] = -1 when both strings of pair i exceed AQC_MAX_LEN. ---- */
int aqc_edit_distance_batch(aqc_ctx *ctx, const uint8_t *a, const uint32_t *a_off, const uint8_t *b, const uint32_t *b_off,
                            uint32_t n, int mem, int32_t *out);

/* ---- the two symbols of the reference's editdistance/libed.so (_editdistance.h:16,23), same signatures, so that the
 * untouched reference can load this library through util.py:15-18.  edit_distance: as the reference's.  seek_overlap: the
 * semantics of util.overlap_hm (the Python the reference actually runs), not those of the C function it replaces; returns
 * (offset << 8) + min(diff, 255), or 0x7FFFFFFF for "not matched" / unsupported constants.  One pair per call, served by a
 * lazily created context on the current device (AQC_DEVICE overrides). ---- */
unsigned int edit_distance(const char *a, const unsigned int asize, const char *b, const unsigned int bsize);
int seek_overlap(const char *r1, const int len1, const char *rc_r2, const int len2, const int limit_distance,
                 const int complete_compare_require, const int overlap_require);

/* instrumentation for bench.py: kernels launched by this context so far, and the device
 * time in ms of the last filter/stat call's kernels (CUDA events on the launching stream) */
uint64_t aqc_launch_count(const aqc_ctx *ctx);
float aqc_last_kernel_ms(const aqc_ctx *ctx);
/* ... split by phase: 0 = the filter kernel (lane_kernel or pair_kernel; aqc_ops_pairs), 1 = pair_kernel's list mode after
 * lane_kernel, 2 = the statistics launches (stamp_bits_kernel + stat_kernel, or pair_kernel in stat mode); -1 = all */
float aqc_last_phase_ms(const aqc_ctx *ctx, int phase);

#ifdef __cplusplus
}
#endif
#endif /* AFTERQC_B200_H */
