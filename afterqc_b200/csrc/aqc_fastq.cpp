// aqc_fastq.cpp -- host-side FASTQ ingest/egress of libafterqc_b200.so.
// Replaces the python str-list I/O of the reference (fastq.py:17-104) on the packed-column layout: no CUDA here.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../include/afterqc_b200.h"

namespace {

inline bool is_ws(uint8_t c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == 0x0b || c == 0x0c; }

const char *const kFlags[AQC_NUM_CLASSES] = {"", "BADTRIM1", "BADTRIM2", "BADLEN", "BADPOL", "BADLQC", "BADNCT", "BADDIFF", "BADMISMATCH"};

const char *flag_of(uint32_t cls) {
    if (cls == AQC_HOST_BADBCD1) return "BADBCD1";
    if (cls == AQC_HOST_BADBCD2) return "BADBCD2";
    return kFlags[cls < AQC_NUM_CLASSES ? cls : 0];
}

// ---- barcode (UMI) helpers, barcodeprocesser.py ----
inline int diff_number(const uint8_t *a, const uint8_t *b, int n) {          // diffNumber :9-14
    int d = 0;
    for (int i = 0; i < n; i++) d += a[i] != b[i];
    return d;
}

// detectBarcode :19-32: where the barcode ends (design length, or one off), 0 = not found
int detect_barcode(const uint8_t *seq, int64_t len, int blen, const uint8_t *verify, int vlen) {
    if (len <= (int64_t)vlen + blen + 1) return 0;
    if (diff_number(seq + blen, verify, vlen) <= 1) return blen;
    if (diff_number(seq + blen - 1, verify, vlen) == 0) return blen - 1;
    if (diff_number(seq + blen + 1, verify, vlen) == 0) return blen + 1;
    return 0;
}

inline uint8_t comp_or_n(uint8_t c) {                                          // util.reverseComplement :42-51
    switch (c) {
        case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C';
        case 'a': return 't'; case 't': return 'a'; case 'c': return 'g'; case 'g': return 'c';
        case 'N': return 'N'; case '\n': return '\n';
        default: return 'N';
    }
}

// Levenshtein distance (util.editDistance :65-83; the editdistance module and libed.so compute the same number)
int edit_distance(const uint8_t *a, int m, const uint8_t *b, int n, std::vector<int> &row) {
    row.resize((size_t)n + 1);
    for (int j = 0; j <= n; j++) row[j] = j;
    for (int i = 1; i <= m; i++) {
        int diag = row[0];
        row[0] = i;
        for (int j = 1; j <= n; j++) {
            int up = row[j];
            int v = std::min(std::min(up + 1, row[j - 1] + 1), diag + (a[i - 1] != b[j - 1]));
            diag = up;
            row[j] = v;
        }
    }
    return row[n];
}

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
// record rec_base + i.  Slices and edits come from the aqc_result records (see include/afterqc_b200.h).  `Rec` says where
// the four lines of a record lie: in packed columns (aqc_fastq_emit) or in the FASTQ text itself (aqc_fastq_emit_lines).
extern "C++" {
namespace {
struct RecLines { const uint8_t *name, *seq, *plus, *qual; uint64_t name_len, seq_len, plus_len; };

template <class Rec>
int emit_impl(int mate, int which, const Rec &at, uint64_t rec_base, const aqc_result *results, uint64_t n,
              uint8_t *out, uint64_t out_cap, uint64_t *out_len) {
    if (mate < 0 || mate > 2) return AQC_ERR_INVALID;          // mate 0: index read, passed through whole and unedited
    uint64_t w = 0;
    for (uint64_t i = 0; i < n; i++) {
        const aqc_result &r = results[i];
        const bool good = r.cls == AQC_GOOD;
        if ((which == 1) == good) continue;
        const RecLines L = at(rec_base + i);
        uint32_t start = mate == 1 ? r.start1 : r.start2, len = mate == 1 ? r.len1 : r.len2;
        if (mate == 0) { start = 0; len = (uint32_t)L.seq_len; }
        int corrected = 0;
        for (int e = 0; e < r.n_edits && e < 4; e++) if (AQC_EDIT_KIND(r.edits[e]) < 2) corrected++;
        if (which == 2) {
            if (!(r.ov_len > 30 && (r.ov_diff == 0 || (int)r.ov_diff == corrected))) continue;
            if (mate != 0) { start += len - r.ov_len; len = r.ov_len; }
        }
        const uint64_t nl = L.name_len, pl = L.plus_len;
        if (w + nl + pl + 2ull * len + 4 + 16 > out_cap) return AQC_ERR_NOMEM;
        const uint8_t *nm = L.name;
        if (which == 1) {
            out[w++] = '@';
            const char *f = flag_of(r.cls);
            size_t fl = strlen(f);
            memcpy(out + w, f, fl); w += fl;
            if (nl > 1) { memcpy(out + w, nm + 1, nl - 1); w += nl - 1; }
        } else { memcpy(out + w, nm, nl); w += nl; }
        out[w++] = '\n';
        uint8_t *so = out + w;
        memcpy(so, L.seq + start, len); w += len;
        out[w++] = '\n';
        memcpy(out + w, L.plus, pl); w += pl;
        out[w++] = '\n';
        uint8_t *qo = out + w;
        memcpy(qo, L.qual + start, len); w += len;
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
}  // namespace
}  // extern "C++"

int aqc_fastq_emit(int mate, int which,
                   const uint8_t *names, const uint64_t *name_off, const uint8_t *seqs, const uint64_t *seq_off,
                   const uint8_t *plus, const uint64_t *plus_off, const uint8_t *quals,
                   uint64_t rec_base, const aqc_result *results, uint64_t n,
                   uint8_t *out, uint64_t out_cap, uint64_t *out_len) {
    auto at = [&](uint64_t rec) {
        RecLines L;
        L.name = names + name_off[rec]; L.name_len = name_off[rec + 1] - name_off[rec];
        L.seq = seqs + seq_off[rec]; L.qual = quals + seq_off[rec]; L.seq_len = seq_off[rec + 1] - seq_off[rec];
        L.plus = plus + plus_off[rec]; L.plus_len = plus_off[rec + 1] - plus_off[rec];
        return L;
    };
    return emit_impl(mate, which, at, rec_base, results, n, out, out_cap, out_len);
}

// The same from the FASTQ text and the line table of aqc_fastq_parse_device (line 4r + k of record r: name, bases, '+',
// qualities; start and length of the rstrip()ped line in `text`).
int aqc_fastq_emit_lines(int mate, int which, const uint8_t *text, const uint32_t *line_start, const uint32_t *line_len,
                         uint64_t rec_base, const aqc_result *results, uint64_t n,
                         uint8_t *out, uint64_t out_cap, uint64_t *out_len) {
    if (!text || !line_start || !line_len) return AQC_ERR_INVALID;
    auto at = [&](uint64_t rec) {
        RecLines L;
        const uint32_t *s = line_start + 4 * rec, *l = line_len + 4 * rec;
        L.name = text + s[0]; L.name_len = l[0];
        L.seq = text + s[1]; L.seq_len = l[1];
        L.plus = text + s[2]; L.plus_len = l[2];
        L.qual = text + s[3];
        return L;
    };
    return emit_impl(mate, which, at, rec_base, results, n, out, out_cap, out_len);
}

// Barcode (UMI) pre-pass of the per-read loop (preprocesser.py:435-452, barcodeprocesser.py): for every pair detect the
// barcode of both mates, move it into the name ('@' + barcode + name[name.find(':'):], :34-46), drop barcode + verify
// bases from the front and, for pairs, the read-through tail found by cleanBarcodeTail (:48-74).  Records are written
// to `out*` in the input order; pairs without a barcode are copied unchanged with status 1 (BADBCD1) / 2 (BADBCD2).
// Single-end input strips the DESIGN length whatever detectBarcode returned (:443-444).  removed[m] receives the bases
// dropped from mate m+1 of the status-0 records.  Output capacity: names in.bytes + n * (barcode_length + 3); bases and
// qualities in.bytes.
int aqc_barcode_pairs(int barcode_length, const char *verify, uint64_t n, const aqc_columns *in1, const aqc_columns *in2,
                      aqc_columns_out *out1, aqc_columns_out *out2, uint8_t *status, uint64_t removed[2]) {
    if (!verify || !in1 || !out1 || !status || !removed || barcode_length < 1) return AQC_ERR_INVALID;
    if ((in2 == nullptr) != (out2 == nullptr)) return AQC_ERR_INVALID;
    const int vlen = (int)strlen(verify);
    const uint8_t *vf = (const uint8_t *)verify;
    const aqc_columns *in[2] = {in1, in2};
    aqc_columns_out *out[2] = {out1, out2};
    const int mates = in2 ? 2 : 1;
    uint64_t wn[2] = {0, 0}, ws[2] = {0, 0};
    removed[0] = removed[1] = 0;
    for (int m = 0; m < mates; m++) { out[m]->name_off[0] = 0; out[m]->seq_off[0] = 0; }
    std::vector<int> row;
    std::vector<uint8_t> start[2], rev[2];
    for (uint64_t i = 0; i < n; i++) {
        const uint8_t *seq[2], *qual[2], *name[2];
        int64_t slen[2], nlen[2];
        for (int m = 0; m < mates; m++) {
            seq[m] = in[m]->seqs + in[m]->seq_off[i]; qual[m] = in[m]->quals + in[m]->seq_off[i];
            slen[m] = (int64_t)(in[m]->seq_off[i + 1] - in[m]->seq_off[i]);
            name[m] = in[m]->names + in[m]->name_off[i];
            nlen[m] = (int64_t)(in[m]->name_off[i + 1] - in[m]->name_off[i]);
        }
        int bl[2] = {0, 0};
        uint8_t st = 0;
        bl[0] = detect_barcode(seq[0], slen[0], barcode_length, vf, vlen);
        if (bl[0] == 0) st = 1;
        else if (mates == 2) { bl[1] = detect_barcode(seq[1], slen[1], barcode_length, vf, vlen); if (bl[1] == 0) st = 2; }
        else bl[0] = barcode_length;
        status[i] = st;
        int64_t front[2] = {0, 0}, tail = 0;
        if (st == 0) {
            for (int m = 0; m < mates; m++) front[m] = std::min<int64_t>(slen[m], (int64_t)bl[m] + vlen);
            if (mates == 2) {                                                  // cleanBarcodeTail on the stripped reads
                for (int m = 0; m < 2; m++) {                                 // readStart = barcode + verify (the DESIGN verify)
                    start[m].assign(seq[m], seq[m] + bl[m]);
                    start[m].insert(start[m].end(), vf, vf + vlen);
                    const size_t L = start[m].size();
                    rev[m].resize(L);
                    for (size_t k = 0; k < L; k++) rev[m][k] = comp_or_n(start[m][L - 1 - k]);
                }
                const int64_t L = (int64_t)std::min(start[0].size(), start[1].size());
                const int64_t r1len = slen[0] - front[0], r2len = slen[1] - front[1];
                for (int64_t k = 0; k < L; k++) {
                    const int64_t comp = L - k;
                    if (comp >= r1len || comp >= r2len) continue;
                    const int d1 = edit_distance(seq[0] + slen[0] - comp, (int)comp, rev[1].data() + k, (int)(rev[1].size() - k), row);
                    const int d2 = edit_distance(seq[1] + slen[1] - comp, (int)comp, rev[0].data() + k, (int)(rev[0].size() - k), row);
                    const int thr = (int)(comp / 5);
                    if (d1 <= thr && d2 <= thr) { tail = comp; break; }
                }
            }
        }
        for (int m = 0; m < mates; m++) {
            uint8_t *nm = out[m]->names + wn[m];
            if (st == 0) {
                int64_t colon = -1;
                for (int64_t k = 0; k < nlen[m]; k++) if (name[m][k] == ':') { colon = k; break; }
                if (colon < 0) colon = nlen[m] > 0 ? nlen[m] - 1 : 0;          // name[-1:] when there is no ':' (python slice)
                const int64_t bc = std::min<int64_t>(bl[m], slen[m]);
                *nm++ = '@';
                memcpy(nm, seq[m], (size_t)bc); nm += bc;
                memcpy(nm, name[m] + colon, (size_t)(nlen[m] - colon)); nm += nlen[m] - colon;
            } else { memcpy(nm, name[m], (size_t)nlen[m]); nm += nlen[m]; }
            wn[m] = (uint64_t)(nm - out[m]->names);
            out[m]->name_off[i + 1] = wn[m];
            const int64_t keep = slen[m] - front[m] - tail;
            memcpy(out[m]->seqs + ws[m], seq[m] + front[m], (size_t)keep);
            memcpy(out[m]->quals + ws[m], qual[m] + front[m], (size_t)keep);
            ws[m] += (uint64_t)keep;
            out[m]->seq_off[i + 1] = ws[m];
            if (st == 0) removed[m] += (uint64_t)(front[m] + tail);
        }
    }
    return 0;
}

}  // extern "C"
