// aqc_fastq.cpp -- host-side FASTQ ingest/egress of libafterqc_b200.so.
// Replaces the python str-list I/O of the reference (fastq.py:17-104) on the packed-column layout: no CUDA here.
#include <cstdint>
#include <cstring>
#include "../../include/afterqc_b200.h"

namespace {

inline bool is_ws(uint8_t c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == 0x0b || c == 0x0c; }

const char *const kFlags[AQC_NUM_CLASSES] = {"", "BADTRIM1", "BADTRIM2", "BADLEN", "BADPOL", "BADLQC", "BADNCT", "BADDIFF", "BADMISMATCH"};

}  // namespace

extern "C" {

// Parse complete 4-line records out of buf[0..n).  Every line is rstrip()'d; the first line that is empty after the
// strip ends the file and its record is dropped (fastq.py:41-48, quirk Q13).  A trailing partial record is left
// unconsumed (or dropped when `final`).  Column k (0 names, 1 bases, 2 '+' lines, 3 qualities) is appended to
// out_bytes[k] with offsets out_off[k][0..records] (out_off[k][0] must be set by the caller, normally 0).
// Returns 0, or AQC_ERR_INVALID when a record's quality length differs from its sequence length (*bad_record).
int aqc_fastq_parse(const uint8_t *buf, uint64_t n, int final, uint64_t max_records,
                    uint8_t *const out_bytes[4], uint64_t *const out_off[4],
                    uint64_t *n_records, uint64_t *consumed, int *hit_eof, uint64_t *bad_record) {
    if (!buf && n) return AQC_ERR_INVALID;
    uint64_t pos = 0, rec = 0, done_to = 0;
    uint64_t w[4] = {out_off[0][0], out_off[1][0], out_off[2][0], out_off[3][0]};
    *hit_eof = 0;
    *bad_record = 0;
    while (rec < max_records) {
        uint64_t ls[4], ll[4];
        uint64_t p = pos;
        int k = 0;
        bool partial = false, empty = false;
        for (; k < 4; k++) {
            if (p >= n) { partial = true; break; }
            const uint8_t *nl = (const uint8_t *)memchr(buf + p, '\n', n - p);
            uint64_t e;
            if (nl) e = (uint64_t)(nl - buf);
            else if (final) e = n;                      // last line without newline
            else { partial = true; break; }
            uint64_t s = p, t = e;
            while (t > s && is_ws(buf[t - 1])) t--;
            if (t == s) { empty = true; break; }
            ls[k] = s; ll[k] = t - s;
            p = nl ? e + 1 : n;
        }
        if (empty) { *hit_eof = 1; break; }
        if (partial) { if (final) *hit_eof = 1; break; }
        if (ll[1] != ll[3]) { *bad_record = rec; *n_records = rec; *consumed = done_to; return AQC_ERR_INVALID; }
        for (int c = 0; c < 4; c++) {
            memcpy(out_bytes[c] + w[c], buf + ls[c], ll[c]);
            w[c] += ll[c];
            out_off[c][rec + 1] = w[c];
        }
        rec++;
        pos = p;
        done_to = p;
    }
    if (final && pos >= n) *hit_eof = 1;
    *n_records = rec;
    *consumed = done_to;
    return 0;
}

// FASTQ text of the records of one mate selected by `which`: 0 = good, 1 = bad (name becomes "@" FLAG name[1:],
// preprocesser.py:212-213), 2 = overlapped tails of good pairs (--store_overlap, :615-617).  results[i] belongs to
// column record rec_base + i.  Slices and edits come from the aqc_result records (see include/afterqc_b200.h).
int aqc_fastq_emit(int mate, int which,
                   const uint8_t *names, const uint64_t *name_off, const uint8_t *seqs, const uint64_t *seq_off,
                   const uint8_t *plus, const uint64_t *plus_off, const uint8_t *quals,
                   uint64_t rec_base, const aqc_result *results, uint64_t n,
                   uint8_t *out, uint64_t out_cap, uint64_t *out_len) {
    if (mate < 0 || mate > 2) return AQC_ERR_INVALID;          // mate 0: index read, passed through whole and unedited
    uint64_t w = 0;
    for (uint64_t i = 0; i < n; i++) {
        const aqc_result &r = results[i];
        const bool good = r.cls == AQC_GOOD;
        if ((which == 1) == good) continue;
        const uint64_t rec0 = rec_base + i;
        uint32_t start = mate == 1 ? r.start1 : r.start2, len = mate == 1 ? r.len1 : r.len2;
        if (mate == 0) { start = 0; len = (uint32_t)(seq_off[rec0 + 1] - seq_off[rec0]); }
        int corrected = 0;
        for (int e = 0; e < r.n_edits && e < 4; e++) if (AQC_EDIT_KIND(r.edits[e]) < 2) corrected++;
        if (which == 2) {
            if (!(r.ov_len > 30 && (r.ov_diff == 0 || (int)r.ov_diff == corrected))) continue;
            if (mate != 0) { start += len - r.ov_len; len = r.ov_len; }
        }
        const uint64_t rec = rec_base + i;
        const uint64_t nl = name_off[rec + 1] - name_off[rec], pl = plus_off[rec + 1] - plus_off[rec];
        if (w + nl + pl + 2ull * len + 4 + 16 > out_cap) return AQC_ERR_NOMEM;
        const uint8_t *nm = names + name_off[rec];
        if (which == 1) {
            out[w++] = '@';
            const char *f = kFlags[r.cls < AQC_NUM_CLASSES ? r.cls : 0];
            size_t fl = strlen(f);
            memcpy(out + w, f, fl); w += fl;
            if (nl > 1) { memcpy(out + w, nm + 1, nl - 1); w += nl - 1; }
        } else { memcpy(out + w, nm, nl); w += nl; }
        out[w++] = '\n';
        const uint8_t *s = seqs + seq_off[rec], *q = quals + seq_off[rec];
        uint8_t *so = out + w;
        memcpy(so, s + start, len); w += len;
        out[w++] = '\n';
        memcpy(out + w, plus + plus_off[rec], pl); w += pl;
        out[w++] = '\n';
        uint8_t *qo = out + w;
        memcpy(qo, q + start, len); w += len;
        out[w++] = '\n';
        for (int e = 0; mate != 0 && e < r.n_edits && e < 4; e++) {       // apply the correction-walk edits that fall inside the slice
            const uint32_t ed = r.edits[e], kind = AQC_EDIT_KIND(ed);
            uint32_t pos; bool has_base = false;
            if (kind == 0 && mate == 1) { pos = AQC_EDIT_POS(ed); has_base = true; }
            else if (kind == 1 && mate == 2) { pos = AQC_EDIT_POS(ed); has_base = true; }
            else if (kind == 2) pos = mate == 1 ? AQC_EDIT_POS(ed) : AQC_EDIT_POS2(ed);
            else continue;
            if (pos < start || pos >= start + len) continue;
            if (has_base) { so[pos - start] = (uint8_t)AQC_EDIT_BASE(ed); qo[pos - start] = (uint8_t)AQC_EDIT_QUAL(ed); }
            else qo[pos - start] = '!';
        }
    }
    *out_len = w;
    return 0;
}

}  // extern "C"
