/*
 * aqc_oracle.c -- CPU oracle for the AfterQC per-read hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is a plain-C restatement of the reference's Python
 * algorithm (byte loops in the reference's own order, no bit tricks) with the same structs as
 * include/afterqc_b200.h.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product (afterqc_b200/) never links or calls it.
 *
 * Parity pin: oracle/make_golden.py + tests/test_oracle_vs_reference.py run the UNMODIFIED
 * reference (oracle/ref_loader.py) in the build container and compare it with this file on
 * testdata/ and on seeded synthetic reads; the resulting vectors are committed under
 * tests/golden/ so the pin also holds where /root/reference does not exist.
 *
 * Behaviour outside the reference's defined domain (it would raise there):
 *   - a byte not in util.COMP reaching util.complement (KeyError, preprocesser.py:565,573)
 *     is treated like reverseComplement treats it: it complements to 'N' (util.py:47-50);
 *   - an error-matrix update whose bases are not both in ACGT (KeyError, :573,:581) is skipped;
 *   - reads of 1..4 bases reaching statRead (IndexError, qualitycontrol.py:106-108) and reads
 *     longer than MAX_LEN set a sticky error code instead.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "../include/afterqc_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---------- k-mer side table (non-ACGT k-mers), open addressing, grows on demand ---------- */
typedef struct {
    uint64_t *keys, *counts, *first;
    uint8_t *used;
    uint32_t cap, n;
} side_t;

static void side_init(side_t *t, uint32_t cap) {
    t->cap = cap; t->n = 0;
    t->keys = (uint64_t *)calloc(cap, 8); t->counts = (uint64_t *)calloc(cap, 8);
    t->first = (uint64_t *)calloc(cap, 8); t->used = (uint8_t *)calloc(cap, 1);
}
static void side_free(side_t *t) { free(t->keys); free(t->counts); free(t->first); free(t->used); }
static uint32_t side_hash(uint64_t k) { k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; return (uint32_t)k; }
static uint32_t side_find(side_t *t, uint64_t key, int insert);
static void side_grow(side_t *t) {
    side_t o = *t;
    side_init(t, o.cap * 2);
    for (uint32_t i = 0; i < o.cap; i++)
        if (o.used[i]) { uint32_t j = side_find(t, o.keys[i], 1); t->counts[j] = o.counts[i]; t->first[j] = o.first[i]; }
    side_free(&o);
}
static uint32_t side_find(side_t *t, uint64_t key, int insert) {
    if (insert && (t->n + 1) * 2 > t->cap) side_grow(t);
    uint32_t m = t->cap - 1, i = side_hash(key) & m;
    while (t->used[i]) { if (t->keys[i] == key) return i; i = (i + 1) & m; }
    if (!insert) return 0xFFFFFFFFu;
    t->used[i] = 1; t->keys[i] = key; t->counts[i] = 0; t->first[i] = AQC_KMER_NEVER; t->n++;
    return i;
}

/* ---------- one QualityControl object (qualitycontrol.py:33-57) ---------- */
typedef struct {
    aqc_qc_counters c;
    uint64_t *kmer_count;   /* dense 4^k */
    uint64_t *kmer_first;
    side_t side;
} qc_t;

typedef struct aqo_ctx {
    aqc_params p;
    int64_t counters[AQC_C_TOTAL];
    qc_t qc[AQC_NUM_QC];
    int err;
    char errmsg[256];
} aqo_ctx;

/* util.COMP (util.py:27); everything else complements to 'N' as in reverseComplement (util.py:47-50) */
static uint8_t COMP[256];
static int8_t BASE_IDX[256];   /* ALL_BASES order A,T,C,G (qualitycontrol.py:24), -1 otherwise */
static int8_t KCODE[256];      /* dense k-mer code A=0 C=1 G=2 T=3, -1 otherwise */
static int tables_ready = 0;
static void init_tables(void) {
    if (tables_ready) return;
    for (int i = 0; i < 256; i++) { COMP[i] = 'N'; BASE_IDX[i] = -1; KCODE[i] = -1; }
    COMP['A'] = 'T'; COMP['T'] = 'A'; COMP['C'] = 'G'; COMP['G'] = 'C';
    COMP['a'] = 't'; COMP['t'] = 'a'; COMP['c'] = 'g'; COMP['g'] = 'c';
    COMP['N'] = 'N'; COMP['\n'] = '\n';
    BASE_IDX['A'] = 0; BASE_IDX['T'] = 1; BASE_IDX['C'] = 2; BASE_IDX['G'] = 3;
    KCODE['A'] = 0; KCODE['C'] = 1; KCODE['G'] = 2; KCODE['T'] = 3;
    tables_ready = 1;
}

/* ---------- per-read operators ---------- */

/* hasPolyX(seq, maxPoly, mismatch)  preprocesser.py:30-51.  Returns the base or 0 for None. */
static int has_polyx(const uint8_t *seq, int len, int maxPoly, int mismatch) {
    static const char polyArray[9] = {'A', 'T', 'C', 'G', 'a', 't', 'c', 'g', 'N'};
    int polyCount[256];
    if (len < maxPoly) return 0;                                  /* :31-32 */
    memset(polyCount, 0, sizeof polyCount);
    for (int x = 0; x < len; x++) {                               /* :38 */
        uint8_t frontbase = seq[x];
        int in_array = 0;
        for (int k = 0; k < 9; k++) if (frontbase == (uint8_t)polyArray[k]) in_array = 1;
        if (!in_array) return 0;                                  /* :41-42 */
        if (x >= maxPoly) polyCount[seq[x - maxPoly]] -= 1;       /* :44-46 */
        polyCount[frontbase] += 1;                                /* :48 */
        if (polyCount[frontbase] >= maxPoly - mismatch) return frontbase; /* :49-50 */
    }
    return 0;
}

/* lowQualityNum(read, qual)  preprocesser.py:61-68 */
static int low_quality_num(const uint8_t *q, int len, int qual) {
    int n = 0;
    qual += 33;
    for (int i = 0; i < len; i++) if ((int)q[i] < qual) n++;
    return n;
}

/* nNumber(read)  preprocesser.py:70-76 */
static int n_number(const uint8_t *s, int len) {
    int n = 0;
    for (int i = 0; i < len; i++) if (s[i] == 'N') n++;
    return n;
}

/* util.reverseComplement  util.py:42-51 */
static void reverse_complement(const uint8_t *s, int len, uint8_t *out) {
    for (int i = 0; i < len; i++) out[i] = COMP[s[len - i - 1]];
}

/* util.overlap -> overlap_hm(r1, r2)  util.py:88-89,158-212.  The acceptance test uses the
 * loop variable i after the loop exactly as the reference does (quirk Q6). */
static void overlap_hm(const uint8_t *r1, int len1, const uint8_t *r2, int len2, int *o_off, int *o_len, int *o_diff) {
    uint8_t rc[AQC_MAX_LEN + 8];
    const int limit_distance = 3, overlap_require = 30, complete_compare_require = 50;   /* :163-165 */
    int overlap_len = 0, offset = 0, i = 0, diff;
    reverse_complement(r2, len2, rc);                              /* :161 */
    while (offset < len1 - overlap_require) {                      /* :172 */
        overlap_len = (len1 - offset < len2) ? len1 - offset : len2;   /* :174 */
        diff = 0;
        for (i = 0; i < overlap_len; i++) {                        /* :177 */
            if (r1[offset + i] != rc[i]) {
                diff += 1;
                if (diff >= limit_distance && i < complete_compare_require) break;   /* :180-181 */
            }
        }
        /* python's `for i in xrange(n)` leaves i = n-1 after a complete loop, C leaves n */
        if (i == overlap_len && overlap_len > 0) i = overlap_len - 1;
        if (diff < limit_distance || (diff >= limit_distance && i > complete_compare_require)) {   /* :183 */
            *o_off = offset; *o_len = overlap_len; *o_diff = diff; return;
        }
        offset += 1;
    }
    offset = 0;                                                    /* :194 */
    while (offset > -(len2 - overlap_require)) {                   /* :195 */
        int ao = -offset;
        overlap_len = (len1 < len2 - ao) ? len1 : len2 - ao;       /* :197 */
        diff = 0;
        for (i = 0; i < overlap_len; i++) {                        /* :200 */
            if (r1[i] != rc[ao + i]) {
                diff += 1;
                if (diff >= limit_distance && i < complete_compare_require) break;
            }
        }
        if (i == overlap_len && overlap_len > 0) i = overlap_len - 1;
        if (diff < limit_distance || (diff >= limit_distance && i > complete_compare_require)) {   /* :206 */
            *o_off = offset; *o_len = overlap_len; *o_diff = diff; return;
        }
        offset -= 1;
    }
    *o_off = 0; *o_len = 0; *o_diff = 0;                            /* :212 */
}

/* ---------- QualityControl.statRead  qualitycontrol.py:73-122 ---------- */
static uint64_t pack_key(const uint8_t *s, int k) {
    uint64_t v = 0;
    for (int j = 0; j < k; j++) v = (v << 8) | s[j];
    return v;
}

static void kmer_touch(qc_t *q, const uint8_t *kmer, int k, uint64_t when, int add) {
    /* kmerCount[kmer] (+= add), inserting with insertion stamp `when` if absent */
    int dense = 1; uint32_t idx = 0;
    for (int j = 0; j < k; j++) { int c = KCODE[kmer[j]]; if (c < 0) { dense = 0; break; } idx = (idx << 2) | (uint32_t)c; }
    if (dense) {
        if (q->kmer_first[idx] == AQC_KMER_NEVER) q->kmer_first[idx] = when;
        q->kmer_count[idx] += (uint64_t)add;
    } else {
        uint32_t i = side_find(&q->side, pack_key(kmer, k), 1);
        if (q->side.first[i] == AQC_KMER_NEVER) q->side.first[i] = when;
        q->side.counts[i] += (uint64_t)add;
    }
}
static int kmer_present(qc_t *q, const uint8_t *kmer, int k) {
    int dense = 1; uint32_t idx = 0;
    for (int j = 0; j < k; j++) { int c = KCODE[kmer[j]]; if (c < 0) { dense = 0; break; } idx = (idx << 2) | (uint32_t)c; }
    if (dense) return q->kmer_first[idx] != AQC_KMER_NEVER;
    return side_find(&q->side, pack_key(kmer, k), 0) != 0xFFFFFFFFu;
}

static int stat_read(aqo_ctx *ctx, qc_t *q, const uint8_t *seq, const uint8_t *qual, int seqlen, uint64_t order) {
    int gc = 0, k = ctx->p.qc_kmer;
    if (seqlen > AQC_MAX_LEN) return AQC_ERR_TOO_LONG;
    if (seqlen > 0 && seqlen < 5) return AQC_ERR_TOO_SHORT_STAT;
    for (int i = 0; i < seqlen; i++) {                              /* :79 */
        int qnum, left, right, discontinuity = 0;
        uint8_t b;
        q->c.totalNum[i] += 1;                                      /* :80 */
        qnum = (int)qual[i] - 33;                                   /* :82 util.qualNum */
        q->c.totalQual[i] += qnum;                                  /* :88 */
        b = seq[i];
        if (b == 'G' || b == 'C') gc += 1;                          /* :90-91 */
        if (BASE_IDX[b] >= 0) {                                     /* :92-94 */
            q->c.baseCounts[BASE_IDX[b]][i] += 1;
            q->c.baseTotalQual[BASE_IDX[b]][i] += qnum;
        }
        left = i - 2; right = i + 3;                                /* :97-104 */
        if (left < 0) { left = 0; right = 5; }
        else if (right >= seqlen) { right = seqlen; left = seqlen - 5; }
        for (int j = left; j < right - 1; j++) if (seq[j] != seq[j + 1]) discontinuity += 1;   /* :106-108 */
        q->c.totalDiscontinuity[i] += discontinuity;                /* :109 */
    }
    q->c.gcHistogram[gc] += 1;                                      /* :112 */
    for (int i = 0; i < seqlen - k; i++) {                          /* :113 (drops the last k-mer, quirk Q11) */
        const uint8_t *kmer = seq + i;
        uint64_t when = (order << 11) | ((uint64_t)i << 1);
        q->c.totalKmer += 1;                                        /* :114 */
        if (kmer_present(q, kmer, k)) {                             /* :116-117 */
            kmer_touch(q, kmer, k, when, 1);
        } else {                                                    /* :118-122 */
            uint8_t rck[AQC_MAX_KMER];
            kmer_touch(q, kmer, k, when, 1);
            reverse_complement(kmer, k, rck);
            if (!kmer_present(q, rck, k)) kmer_touch(q, rck, k, when | 1, 0);
        }
    }
    q->c.reads += 1;
    return 0;
}

/* ---------- helpers ---------- */
/* trim(read, front, tail)  preprocesser.py:19-28, python slice semantics s[front:-tail] / s[front:] */
static void py_trim(int len, int front, int tail, int *start, int *newlen) {
    int s = front < len ? front : len;
    int e = tail > 0 ? len - tail : len;
    if (e < 0) e = 0;
    if (e < s) e = s;
    *start = s; *newlen = e - s;
}

static void set_err(aqo_ctx *ctx, int code, const char *msg) {
    if (!ctx->err) { ctx->err = code; snprintf(ctx->errmsg, sizeof ctx->errmsg, "%s", msg); }
}

/* ---------- API ---------- */
int aqo_create(const aqc_params *p, aqo_ctx **out) {
    init_tables();
    if (!p || !out) return AQC_ERR_INVALID;
    if (p->qc_kmer < 1 || p->qc_kmer > AQC_MAX_KMER) return AQC_ERR_INVALID;
    aqo_ctx *ctx = (aqo_ctx *)calloc(1, sizeof(aqo_ctx));
    if (!ctx) return AQC_ERR_NOMEM;
    ctx->p = *p;
    size_t nk = (size_t)1 << (2 * p->qc_kmer);
    for (int s = 0; s < AQC_NUM_QC; s++) {
        ctx->qc[s].kmer_count = (uint64_t *)calloc(nk, 8);
        ctx->qc[s].kmer_first = (uint64_t *)malloc(nk * 8);
        memset(ctx->qc[s].kmer_first, 0xFF, nk * 8);
        side_init(&ctx->qc[s].side, 1024);
    }
    *out = ctx;
    return 0;
}

void aqo_destroy(aqo_ctx *ctx) {
    if (!ctx) return;
    for (int s = 0; s < AQC_NUM_QC; s++) { free(ctx->qc[s].kmer_count); free(ctx->qc[s].kmer_first); side_free(&ctx->qc[s].side); }
    free(ctx);
}

int aqo_set_params(aqo_ctx *ctx, const aqc_params *p) {
    if (p->qc_kmer != ctx->p.qc_kmer) return AQC_ERR_INVALID;
    ctx->p = *p;
    return 0;
}

const char *aqo_last_error(const aqo_ctx *ctx) { return ctx ? ctx->errmsg : ""; }

int aqo_stat_reads(aqo_ctx *ctx, const aqc_batch *b, int qc1, int qc2, uint64_t stat_lo, uint64_t stat_hi, uint64_t order_base) {
    for (uint32_t i = 0; i < b->n; i++) {
        uint64_t g = b->first_index + i;
        int rc;
        if (g < stat_lo || g >= stat_hi) continue;
        uint64_t order = order_base + (g - stat_lo);
        if (qc1 >= 0) {
            rc = stat_read(ctx, &ctx->qc[qc1], b->seq1 + b->off1[i], b->qual1 + b->off1[i], (int)(b->off1[i + 1] - b->off1[i]), order);
            if (rc) { set_err(ctx, rc, "statRead domain error (mate 1)"); return rc; }
        }
        if (qc2 >= 0 && b->seq2) {
            rc = stat_read(ctx, &ctx->qc[qc2], b->seq2 + b->off2[i], b->qual2 + b->off2[i], (int)(b->off2[i + 1] - b->off2[i]), order);
            if (rc) { set_err(ctx, rc, "statRead domain error (mate 2)"); return rc; }
        }
    }
    return 0;
}

int aqo_ops_pairs(aqo_ctx *ctx, const aqc_batch *b, aqc_ops *out) {
    const aqc_params *p = &ctx->p;
    for (uint32_t i = 0; i < b->n; i++) {
        aqc_ops r; memset(&r, 0, sizeof r);
        int len1 = (int)(b->off1[i + 1] - b->off1[i]), s1 = 0, s2 = 0, len2 = 0;
        const uint8_t *q1 = b->qual1 + b->off1[i], *r1 = b->seq1 + b->off1[i], *r2 = NULL, *q2 = NULL;
        if (len1 > AQC_MAX_LEN) return AQC_ERR_TOO_LONG;
        if (b->seq2) { len2 = (int)(b->off2[i + 1] - b->off2[i]); r2 = b->seq2 + b->off2[i]; q2 = b->qual2 + b->off2[i]; if (len2 > AQC_MAX_LEN) return AQC_ERR_TOO_LONG; }
        if (p->trim_front > 0 || p->trim_tail > 0) {                /* gate keyed on R1 only, quirk Q4 */
            py_trim(len1, p->trim_front, p->trim_tail, &s1, &len1);
            if (r2) py_trim(len2, p->trim_front2, p->trim_tail2, &s2, &len2);
        }
        r1 += s1; q1 += s1; if (r2) { r2 += s2; q2 += s2; }
        r.len1 = (uint16_t)len1; r.len2 = (uint16_t)len2;
        r.poly1 = (uint8_t)has_polyx(r1, len1, p->poly_size_limit, p->allow_mismatch_in_poly);
        r.lowq1 = (uint16_t)low_quality_num(q1, len1, p->qualified_quality_phred);
        r.n1 = (uint16_t)n_number(r1, len1);
        if (r2) {
            int o, l, d;
            r.poly2 = (uint8_t)has_polyx(r2, len2, p->poly_size_limit, p->allow_mismatch_in_poly);
            r.lowq2 = (uint16_t)low_quality_num(q2, len2, p->qualified_quality_phred);
            r.n2 = (uint16_t)n_number(r2, len2);
            overlap_hm(r1, len1, r2, len2, &o, &l, &d);
            r.ov_offset = (int16_t)o; r.ov_len = (uint16_t)l; r.ov_diff = (uint16_t)d;
        }
        out[i] = r;
    }
    return 0;
}

/* The loop body of seqFilter.run()  preprocesser.py:411-631 */
int aqo_filter_pairs(aqo_ctx *ctx, const aqc_batch *b, aqc_result *results) {
    const aqc_params *p = &ctx->p;
    int64_t *C = ctx->counters;
    uint8_t s1buf[AQC_MAX_LEN + 8], q1buf[AQC_MAX_LEN + 8], s2buf[AQC_MAX_LEN + 8], q2buf[AQC_MAX_LEN + 8];
    for (uint32_t i = 0; i < b->n; i++) {
        aqc_result r; memset(&r, 0, sizeof r);
        uint64_t total_reads = b->first_index + i + 1;              /* TOTAL_READS after :433 */
        int paired = b->seq2 != NULL;
        int olen1 = (int)(b->off1[i + 1] - b->off1[i]);
        int olen2 = paired ? (int)(b->off2[i + 1] - b->off2[i]) : 0;
        int start1 = 0, len1 = olen1, start2 = 0, len2 = olen2;
        int cls = AQC_GOOD;
        if (olen1 > AQC_MAX_LEN || olen2 > AQC_MAX_LEN) { set_err(ctx, AQC_ERR_TOO_LONG, "read longer than MAX_LEN"); return AQC_ERR_TOO_LONG; }
        /* working copies: the loop mutates reads in place (corrections) */
        memcpy(s1buf, b->seq1 + b->off1[i], (size_t)olen1); memcpy(q1buf, b->qual1 + b->off1[i], (size_t)olen1);
        if (paired) { memcpy(s2buf, b->seq2 + b->off2[i], (size_t)olen2); memcpy(q2buf, b->qual2 + b->off2[i], (size_t)olen2); }

        C[AQC_C_TOTAL_READS] += 1;                                  /* :433 */
        C[AQC_C_TOTAL_BASES_R1] += olen1;                           /* :416 */
        C[AQC_C_TOTAL_BASES_R2] += olen2;                           /* :431 (host applies the index2 quirk) */

        do {
            uint8_t *r1, *r1q, *r2 = NULL, *r2q = NULL;
            /* trim  :455-466 */
            if (p->trim_front > 0 || p->trim_tail > 0) {
                py_trim(olen1, p->trim_front, p->trim_tail, &start1, &len1);
                if (len1 < 5) { cls = AQC_BADTRIM1; break; }
                if (paired) {
                    py_trim(olen2, p->trim_front2, p->trim_tail2, &start2, &len2);
                    if (len2 < 5) { cls = AQC_BADTRIM2; break; }
                }
            }
            r1 = s1buf + start1; r1q = q1buf + start1;
            if (paired) { r2 = s2buf + start2; r2q = q2buf + start2; }
            /* length  :476-479 (R2 never checked, quirk Q3) */
            if (len1 < p->seq_len_req) { cls = AQC_BADLEN; break; }
            /* polyX  :482-490 */
            if (p->poly_size_limit > 0) {
                int poly1 = has_polyx(r1, len1, p->poly_size_limit, p->allow_mismatch_in_poly);
                int poly2 = paired ? has_polyx(r2, len2, p->poly_size_limit, p->allow_mismatch_in_poly) : 0;
                if (poly1 || poly2) { cls = AQC_BADPOL; break; }
            }
            /* low quality  :493-501 (only lowQual1 is tested, quirk Q2) */
            if (p->unqualified_base_limit > 0) {
                int lowQual1 = low_quality_num(r1q, len1, p->qualified_quality_phred);
                if (lowQual1 > p->unqualified_base_limit || lowQual1 > p->unqualified_base_limit) { cls = AQC_BADLQC; break; }
            }
            /* N count  :504-512 */
            if (p->n_base_limit > 0) {
                int nNum1 = n_number(r1, len1);
                int nNum2 = paired ? n_number(r2, len2) : 0;
                if (nNum1 > p->n_base_limit || nNum2 > p->n_base_limit) { cls = AQC_BADNCT; break; }
            }
            /* overlap  :515-617 */
            if (paired && !p->no_overlap) {
                int offset, overlap_len, distance;
                overlap_hm(r1, len1, r2, len2, &offset, &overlap_len, &distance);   /* :516 */
                C[AQC_C_OVERLAP_HIST + overlap_len] += 1;             /* :517 */
                if (offset < 0 && overlap_len > 30) {                 /* :520 adapter trimming */
                    len1 = overlap_len; len2 = overlap_len;           /* :522-525 */
                    C[AQC_C_TRIMMED_ADAPTER_BASE] += 2 * (-offset);   /* :526 */
                    C[AQC_C_TRIMMED_ADAPTER_READ] += 1;
                    if (len1 < p->seq_len_req) {                      /* :529-532 */
                        r.ov_offset = (int16_t)offset; r.ov_len = (uint16_t)overlap_len; r.ov_diff = (uint16_t)distance;
                        cls = AQC_BADLEN; break;
                    }
                    overlap_hm(r1, len1, r2, len2, &offset, &overlap_len, &distance);   /* :534 */
                }
                r.ov_offset = (int16_t)offset; r.ov_len = (uint16_t)overlap_len; r.ov_diff = (uint16_t)distance;
                C[AQC_C_DISTANCE_HIST + distance] += 1;               /* :536 */
                if (distance > 3) { cls = AQC_BADDIFF; break; }       /* :538-541 */
                if (overlap_len > 30) {                               /* :542 */
                    int corrected = 0, zero_qual_masked = 0, skipped_mismatch = 0;
                    int64_t err_mtx[16];
                    C[AQC_C_OVERLAPPED] += 1;
                    C[AQC_C_OVERLAP_LEN_SUM] += overlap_len;
                    C[AQC_C_OVERLAP_BASE_SUM] += overlap_len * 2;
                    C[AQC_C_OVERLAP_BASE_ERR] += distance;
                    if (distance > 0) {                               /* :551 */
                        memset(err_mtx, 0, sizeof err_mtx);
                        for (int o = 0; o < overlap_len; o++) {       /* :563 */
                            int p1 = len1 - overlap_len + o, p2 = len2 - 1 - o;
                            uint8_t b1 = r1[p1];                      /* :564 */
                            uint8_t b2 = COMP[r2[p2]];                /* :565 */
                            uint8_t qa = r1q[p1], qb = r2q[p2];       /* :566-567 */
                            if (b1 != b2) {
                                int this_is_corrected = 0;
                                int Q1 = (int)qa - 33, Q2 = (int)qb - 33;
                                uint32_t e = 0; int have_edit = 0;
                                if (Q1 >= 30 && Q2 <= 14) {            /* :571 */
                                    if (b1 != 'N' && b2 != 'N') {
                                        int a = BASE_IDX[COMP[b1]], c = BASE_IDX[COMP[b2]];
                                        if (a >= 0 && c >= 0) err_mtx[a * 4 + c] += 1;     /* :573 */
                                    }
                                    if (!p->no_correction) {          /* :574-578 */
                                        r2[p2] = COMP[b1]; r2q[p2] = qa;
                                        corrected += 1; this_is_corrected = 1;
                                        e = (uint32_t)(start2 + p2) | (1u << 10) | ((uint32_t)COMP[b1] << 16) | ((uint32_t)qa << 24);
                                        have_edit = 1;
                                    }
                                } else if (Q2 >= 30 && Q1 <= 14) {     /* :579 */
                                    if (b1 != 'N' && b2 != 'N') {
                                        int a = BASE_IDX[b2], c = BASE_IDX[b1];
                                        if (a >= 0 && c >= 0) err_mtx[a * 4 + c] += 1;     /* :581 */
                                    }
                                    if (!p->no_correction) {          /* :582-586 */
                                        r1[p1] = b2; r1q[p1] = qb;
                                        corrected += 1; this_is_corrected = 1;
                                        e = (uint32_t)(start1 + p1) | (0u << 10) | ((uint32_t)b2 << 16) | ((uint32_t)qb << 24);
                                        have_edit = 1;
                                    }
                                }
                                if (!this_is_corrected) {             /* :587-595 */
                                    if (p->mask_mismatch) {
                                        r2q[p2] = '!'; r1q[p1] = '!';
                                        zero_qual_masked += 1;
                                        e = (uint32_t)(start1 + p1) | (2u << 10) | ((uint32_t)(start2 + p2) << 16);
                                    } else {
                                        skipped_mismatch += 1;
                                        e = (uint32_t)(start1 + p1) | (3u << 10) | ((uint32_t)(start2 + p2) << 16);
                                    }
                                    have_edit = 1;
                                }
                                if (have_edit && r.n_edits < 4) r.edits[r.n_edits++] = e;
                                if (corrected + zero_qual_masked + skipped_mismatch >= distance) break;   /* :597-598 */
                            }
                        }
                        if (corrected + zero_qual_masked + skipped_mismatch == distance) {   /* :603-610 */
                            for (int k = 0; k < 16; k++) C[AQC_C_ERR_MATRIX + k] += err_mtx[k];
                            if (corrected > 0) C[AQC_C_READ_CORRECTED] += 1;
                            C[AQC_C_BASE_CORRECTED] += corrected;
                            C[AQC_C_BASE_ZERO_QUAL_MASKED] += zero_qual_masked * 2;
                            C[AQC_C_BASE_SKIPPED_CORRECTION] += skipped_mismatch * 2;
                        } else {                                       /* :611-614 */
                            cls = AQC_BADMISMATCH; break;
                        }
                    }
                }
            }
        } while (0);

        r.cls = (uint8_t)cls;
        r.start1 = (uint16_t)start1; r.len1 = (uint16_t)len1; r.start2 = (uint16_t)start2; r.len2 = (uint16_t)len2;
        if (cls == AQC_GOOD) {
            C[AQC_C_GOOD_READS] += 1;                                 /* :629 */
            C[AQC_C_GOOD_BASES_R1] += len1;                           /* :621 */
            C[AQC_C_GOOD_BASES_R2] += len2;                           /* :623 (host applies the index2 quirk) */
            if (p->qc_sample <= 0 || total_reads < (uint64_t)p->qc_sample) {   /* :624 */
                int rc = stat_read(ctx, &ctx->qc[AQC_QC_R1_POST], s1buf + start1, q1buf + start1, len1, total_reads - 1);
                if (!rc && paired) rc = stat_read(ctx, &ctx->qc[AQC_QC_R2_POST], s2buf + start2, q2buf + start2, len2, total_reads - 1);
                if (rc) { set_err(ctx, rc, "statRead domain error (postfilter)"); return rc; }
            }
        } else {
            C[AQC_C_BADTRIM1 + (cls - AQC_BADTRIM1)] += 1;
        }
        if (results) results[i] = r;
    }
    return 0;
}

int aqo_get_counters(aqo_ctx *ctx, int64_t *out) { memcpy(out, ctx->counters, sizeof ctx->counters); return ctx->err; }
int aqo_get_qc(aqo_ctx *ctx, int slot, aqc_qc_counters *out) {
    if (slot < 0 || slot >= AQC_NUM_QC) return AQC_ERR_INVALID;
    *out = ctx->qc[slot].c; return ctx->err;
}
int aqo_get_kmer_dense(aqo_ctx *ctx, int slot, uint64_t *counts, uint64_t *first) {
    size_t nk = (size_t)1 << (2 * ctx->p.qc_kmer);
    if (slot < 0 || slot >= AQC_NUM_QC) return AQC_ERR_INVALID;
    memcpy(counts, ctx->qc[slot].kmer_count, nk * 8); memcpy(first, ctx->qc[slot].kmer_first, nk * 8);
    return 0;
}
int aqo_get_kmer_side(aqo_ctx *ctx, int slot, uint64_t *keys, uint64_t *counts, uint64_t *first, uint32_t cap, uint32_t *n_out) {
    if (slot < 0 || slot >= AQC_NUM_QC) return AQC_ERR_INVALID;
    side_t *t = &ctx->qc[slot].side;
    *n_out = t->n;
    if (cap < t->n) return AQC_ERR_INVALID;
    uint32_t k = 0;
    for (uint32_t i = 0; i < t->cap; i++) if (t->used[i]) { keys[k] = t->keys[i]; counts[k] = t->counts[i]; first[k] = t->first[i]; k++; }
    return 0;
}
static int reset_from(aqo_ctx *ctx, int first_slot) {
    size_t nk = (size_t)1 << (2 * ctx->p.qc_kmer);
    memset(ctx->counters, 0, sizeof ctx->counters);
    for (int s = first_slot; s < AQC_NUM_QC; s++) {
        memset(&ctx->qc[s].c, 0, sizeof ctx->qc[s].c);
        memset(ctx->qc[s].kmer_count, 0, nk * 8); memset(ctx->qc[s].kmer_first, 0xFF, nk * 8);
        side_free(&ctx->qc[s].side); side_init(&ctx->qc[s].side, 1024);
    }
    ctx->err = 0; ctx->errmsg[0] = 0;
    return 0;
}
int aqo_reset(aqo_ctx *ctx) { return reset_from(ctx, 0); }
int aqo_reset_filter(aqo_ctx *ctx) { return reset_from(ctx, AQC_QC_R1_POST); }

#ifdef __cplusplus
}
#endif

/* Levenshtein distance, the plain dynamic programme (the number util.editDistance returns, util.py:65-83; the reference
 * computes it with the editdistance module or editdistance/_editdistance.cpp:100-126, Myers' bit-vector algorithm with a DP
 * fall-back :64-75 -- all the same function of the two strings).  Checker of aqc_edit_distance_batch; the reference's own C++
 * is compiled next to it (oracle/Makefile, _ref/libed_ref.so) where /root/reference exists and pins this restatement. */
int aqo_edit_distance(const uint8_t *a, uint32_t la, const uint8_t *b, uint32_t lb) {
    if (la == 0) return (int)lb;
    if (lb == 0) return (int)la;
    uint32_t *row = (uint32_t *)malloc(((size_t)lb + 1) * sizeof(uint32_t));
    if (!row) return -1;
    for (uint32_t j = 0; j <= lb; j++) row[j] = j;
    for (uint32_t i = 1; i <= la; i++) {
        uint32_t diag = row[0];
        row[0] = i;
        for (uint32_t j = 1; j <= lb; j++) {
            uint32_t up = row[j];
            uint32_t v = up + 1;
            if (row[j - 1] + 1 < v) v = row[j - 1] + 1;
            if (diag + (a[i - 1] != b[j - 1]) < v) v = diag + (a[i - 1] != b[j - 1]);
            diag = up;
            row[j] = v;
        }
    }
    int d = (int)row[lb];
    free(row);
    return d;
}

