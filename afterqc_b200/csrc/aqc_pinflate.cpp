// aqc_pinflate.cpp -- see aqc_pinflate.hpp.
#include "aqc_pinflate.hpp"

#include <zlib.h>   // crc32(), crc32_combine()

#include <algorithm>
#include <atomic>
#include <cstring>
#include <functional>
#include <thread>

#include "aqc_inflate.hpp"

namespace aqc {
namespace {

// table entry format of GzipInflater::build_table (bits 0-7 code bits, 8-11 extra bits, 12-15 flags, 16-31 value)
constexpr uint32_t F_LIT = 0x8000, F_SUB = 0x4000, F_EOB = 0x2000, F_BAD = 0x1000;
constexpr int kLitBits = GzipInflater::kLitBits, kDistBits = GzipInflater::kDistBits;
constexpr size_t kWin = 32768;
constexpr size_t kChunk = 1u << 20;             // compressed bytes between two search targets
constexpr uint16_t kPlaceholder = 256;          // symbol 256 + i = byte i of the unknown 32 KB window

const uint8_t kPreOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

inline uint64_t load64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }

struct Tables {
    std::vector<uint32_t> lit, dist, pre, scratch;
    bool lit_complete = false, dist_complete = false;
};

struct Bits {
    const uint8_t *base, *end, *ip;
    uint64_t bb = 0;
    int bc = 0;
    Bits(const uint8_t *b, size_t n, uint64_t bit) : base(b), end(b + n), ip(b + (bit >> 3)) {
        const int sh = (int)(bit & 7);
        if (sh && ip < end) { bb = (uint64_t)*ip++ >> sh; bc = 8 - sh; }
    }
    uint64_t tell() const { return (uint64_t)(ip - base) * 8 - (uint64_t)bc; }
    bool need(int n) {
        while (bc < n) {
            if (ip >= end) return false;
            bb |= (uint64_t)*ip++ << bc;
            bc += 8;
        }
        return true;
    }
    void careful() { while (bc <= 56 && ip < end) { bb |= (uint64_t)*ip++ << bc; bc += 8; } }
    void fast() { bb |= load64(ip) << bc; ip += (63 - bc) >> 3; bc |= 56; }
    uint32_t take(int n) { uint32_t v = (uint32_t)(bb & ((1ull << n) - 1)); bb >>= n; bc -= n; return v; }
    void align() { int d = bc & 7; bb >>= d; bc -= d; ip -= bc >> 3; bb = 0; bc = 0; }
};

bool kraft_complete(const uint8_t *lens, int n, bool *empty_or_single) {
    uint32_t sum = 0;
    int used = 0;
    for (int i = 0; i < n; i++) if (lens[i]) { sum += 1u << (15 - lens[i]); used++; }
    if (empty_or_single) *empty_or_single = used <= 1;
    return sum == (1u << 15);
}

// One block header at the reader's position.  type 0: stored (stored_len set, reader aligned to the data), 1/2: tables built.
const char *parse_header(Bits &b, Tables &t, bool &final, int &type, uint32_t &stored_len) {
    if (!b.need(3)) return "truncated deflate stream";
    final = b.take(1);
    type = (int)b.take(2);
    if (type == 0) {
        b.align();
        if (b.end - b.ip < 4) return "truncated stored block";
        uint32_t len = b.ip[0] | (b.ip[1] << 8), nlen = b.ip[2] | (b.ip[3] << 8);
        if ((len ^ 0xFFFF) != nlen) return "stored block length check failed";
        b.ip += 4;
        stored_len = len;
        return nullptr;
    }
    if (type == 3) return "invalid deflate block type";
    uint8_t lens[288 + 32];
    int nlit, ndist;
    if (type == 1) {
        nlit = 288; ndist = 32;
        for (int i = 0; i < 144; i++) lens[i] = 8;
        for (int i = 144; i < 256; i++) lens[i] = 9;
        for (int i = 256; i < 280; i++) lens[i] = 7;
        for (int i = 280; i < 288; i++) lens[i] = 8;
        for (int i = 0; i < 32; i++) lens[288 + i] = 5;
    } else {
        if (!b.need(14)) return "truncated deflate stream";
        nlit = 257 + (int)b.take(5);
        ndist = 1 + (int)b.take(5);
        const int npre = 4 + (int)b.take(4);
        if (nlit > 286 || ndist > 30) return "too many length or distance symbols";
        uint8_t plen[19] = {0};
        for (int i = 0; i < npre; i++) {
            if (!b.need(3)) return "truncated deflate stream";
            plen[kPreOrder[i]] = (uint8_t)b.take(3);
        }
        if (const char *e = GzipInflater::build_table(plen, 19, 7, 2, t.pre)) return e;
        const int total = nlit + ndist;
        int i = 0;
        while (i < total) {
            b.need(14);
            const uint32_t e = t.pre[b.bb & 127];
            const int cl = (int)(e & 0xff);
            if (cl == 0 || cl > b.bc) return "invalid code lengths set";
            b.take(cl);
            const int sym = (int)(e >> 16);
            if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
            int rep, xb;
            uint8_t val = 0;
            if (sym == 16) {
                if (i == 0) return "invalid bit length repeat";
                val = lens[i - 1]; xb = 2; rep = 3;
            } else if (sym == 17) { xb = 3; rep = 3; }
            else { xb = 7; rep = 11; }
            if (xb > b.bc) return "truncated deflate stream";
            rep += (int)b.take(xb);
            if (i + rep > total) return "invalid bit length repeat";
            while (rep--) lens[i++] = val;
        }
        if (lens[256] == 0) return "missing end-of-block code";
        memmove(lens + 288, lens + nlit, (size_t)ndist);
        for (int k = nlit; k < 288; k++) lens[k] = 0;
    }
    bool single = false;
    t.lit_complete = kraft_complete(lens, nlit, nullptr);
    t.dist_complete = kraft_complete(lens + 288, ndist, &single) || single;
    if (const char *e = GzipInflater::build_table(lens, nlit, kLitBits, 0, t.lit)) return e;
    if (const char *e = GzipInflater::build_table(lens + 288, ndist, kDistBits, 1, t.dist)) return e;
    GzipInflater::pair_literals(t.lit, t.scratch);
    return nullptr;
}

inline bool text_byte(uint32_t c) { return (c >= 32 && c < 127) || c == '\n' || c == '\r' || c == '\t'; }

// Body of one Huffman block.  OUT = true: append 16-bit symbols to out (grown as needed, at most max_out);
// OUT = false (block search): nothing is stored, literals must be text and at most max_out symbols are accepted.
template <bool OUT>
const char *decode_body(Bits &b, const Tables &t, std::vector<uint16_t> &out, size_t &op_io, size_t max_out) {
    const uint32_t *const lit = t.lit.data();
    const uint32_t *const dst = t.dist.data();
    const uint32_t lit_mask = (1u << kLitBits) - 1, dist_mask = (1u << kDistBits) - 1;
    size_t op = op_io;
    uint16_t *o16 = OUT ? out.data() : nullptr;
    size_t cap = OUT ? out.size() : 0;
    const char *why = nullptr;
    for (;;) {
        if (OUT) {
            if (op + 600 > cap) {
                if (cap >= max_out) { why = "runaway output"; break; }
                out.resize(std::min(max_out + 600, cap * 2 + 65536));
                o16 = out.data(); cap = out.size();
            }
        } else if (op > max_out) { why = "block too long"; break; }
        const bool fast = (size_t)(b.end - b.ip) >= 16;
        if (fast) b.fast(); else b.careful();
        uint32_t e = lit[b.bb & lit_mask];
        if (e & F_LIT) {
            if (!fast && (int)(e & 0xff) > b.bc) { why = "truncated deflate stream"; break; }
            int more = fast ? 3 : 0;
            for (;;) {
                b.bb >>= (e & 0xff); b.bc -= (int)(e & 0xff);
                if (OUT) { o16[op] = (uint16_t)((e >> 16) & 0xff); o16[op + 1] = (uint16_t)(e >> 24); }
                else if (!text_byte((e >> 16) & 0xff) || ((e & F_SUB) && !text_byte(e >> 24))) { why = "binary literal"; break; }
                op += 1 + ((e >> 14) & 1);
                if (!more--) break;
                e = lit[b.bb & lit_mask];
                if (!(e & F_LIT)) break;
            }
            if (why) break;
            continue;
        }
        if (e & F_SUB) {
            b.bb >>= kLitBits; b.bc -= kLitBits;
            e = lit[(e >> 16) + (b.bb & ((1u << ((e >> 8) & 15)) - 1))];
            if (e & F_LIT) {
                if (!fast && (int)(e & 0xff) > b.bc) { why = "truncated deflate stream"; break; }
                b.bb >>= (e & 0xff); b.bc -= (int)(e & 0xff);
                if (OUT) o16[op] = (uint16_t)((e >> 16) & 0xff);
                else if (!text_byte((e >> 16) & 0xff)) { why = "binary literal"; break; }
                op++;
                continue;
            }
        }
        if ((e & 0xff) == 0 || (e & F_BAD)) { why = "invalid literal/length code"; break; }
        if (!fast && (int)(e & 0xff) > b.bc) { why = "truncated deflate stream"; break; }
        b.bb >>= (e & 0xff); b.bc -= (int)(e & 0xff);
        if (e & F_EOB) break;
        uint32_t xb = (e >> 8) & 15;
        if (!fast) { b.careful(); if ((int)xb > b.bc) { why = "truncated deflate stream"; break; } }
        const uint32_t len = (e >> 16) + (uint32_t)(b.bb & ((1u << xb) - 1));
        b.bb >>= xb; b.bc -= (int)xb;
        if (!fast) b.careful();
        uint32_t d = dst[b.bb & dist_mask];
        if (d & F_SUB) {
            b.bb >>= kDistBits; b.bc -= kDistBits;
            d = dst[(d >> 16) + (b.bb & ((1u << ((d >> 8) & 15)) - 1))];
        }
        if ((d & 0xff) == 0 || (d & F_BAD)) { why = "invalid distance code"; break; }
        if (!fast && (int)(d & 0xff) > b.bc) { why = "truncated deflate stream"; break; }
        b.bb >>= (d & 0xff); b.bc -= (int)(d & 0xff);
        xb = (d >> 8) & 15;
        if (!fast) { b.careful(); if ((int)xb > b.bc) { why = "truncated deflate stream"; break; } }
        const uint32_t dist = (d >> 16) + (uint32_t)(b.bb & ((1u << xb) - 1));
        b.bb >>= xb; b.bc -= (int)xb;
        if (OUT) {                                                       // dist <= 32768 <= op: the placeholder prefix covers it
            uint16_t *o = o16 + op;
            const uint16_t *s = o - dist;
            uint16_t *const oend = o + len;
            if (dist >= 4) { do { memcpy(o, s, 8); o += 4; s += 4; } while (o < oend); }
            else { do { *o++ = *s++; } while (o < oend); }
        }
        op += len;
    }
    op_io = op;
    return why;
}

struct Piece {
    uint64_t start_bit = 0, end_bit = 0;
    std::vector<uint16_t> sym;              // [32768 placeholders | output symbols]
    size_t n_out = 0;
    bool member_end = false;
    const char *err = nullptr;
    uint32_t crc = 0;
};

// Decode blocks from a block header at start_bit until (a) a block boundary equal to one of `stops` (ascending; the
// ones passed over were false hits), (b) the first boundary at or beyond force_stop_bit, or (c) the member's last block.
void decode_piece(const uint8_t *data, size_t size, uint64_t start_bit, const uint64_t *stops, size_t n_stops,
                  uint64_t force_stop_bit, Tables &t, Piece &p) {
    p.start_bit = start_bit; p.member_end = false; p.err = nullptr;
    if (p.sym.size() < kWin + (4u << 20)) p.sym.resize(kWin + (4u << 20));
    for (size_t i = 0; i < kWin; i++) p.sym[i] = (uint16_t)(kPlaceholder + i);
    size_t op = kWin;
    const size_t soft_cap = kWin + (48u << 20);                          // end the piece at the next block boundary beyond this
    const size_t max_out = soft_cap + (24u << 20);                       // one block cannot add more than ~8 M symbols
    Bits b(data, size, start_bit);
    size_t si = 0;
    bool first = true;
    for (;;) {
        const uint64_t at = b.tell();
        if (!first) {
            while (si < n_stops && stops[si] < at) si++;
            if ((si < n_stops && stops[si] == at) || at >= force_stop_bit || op >= soft_cap) { p.end_bit = at; break; }
        }
        first = false;
        bool final = false;
        int type = 0;
        uint32_t stored = 0;
        if ((p.err = parse_header(b, t, final, type, stored))) break;
        if (type == 0) {
            if ((size_t)(b.end - b.ip) < stored) { p.err = "truncated stored block"; break; }
            if (op + stored + 600 > p.sym.size()) p.sym.resize(op + stored + (4u << 20));
            for (uint32_t k = 0; k < stored; k++) p.sym[op + k] = b.ip[k];
            b.ip += stored; op += stored;
        } else if ((p.err = decode_body<true>(b, t, p.sym, op, max_out))) break;
        if (final) { p.member_end = true; p.end_bit = b.tell(); break; }
    }
    p.n_out = op - kWin;
}

// Does a block start at `bit`?  Trial parse: dynamic-Huffman header with complete codes, a whole block of text
// literals, then a plausible header again.  (Stored / fixed blocks are not looked for: rare in FASTQ streams.)
bool block_starts_at(const uint8_t *data, size_t size, uint64_t bit, Tables &t, std::vector<uint16_t> &none) {
    // cheap rejects on the first 17 bits: BFINAL = 0, BTYPE = 2, HLIT <= 29, HDIST <= 29
    const uint64_t w = load64(data + (bit >> 3)) >> (bit & 7);
    if ((w & 7) != 4) return false;
    if (((w >> 3) & 31) > 29 || ((w >> 8) & 31) > 29) return false;
    const int npre = 4 + (int)((w >> 13) & 15);
    // the code-length code must be complete
    {
        uint64_t pb = bit + 17;
        uint32_t sum = 0;
        for (int i = 0; i < npre; i++, pb += 3) {
            const uint32_t l = (uint32_t)(load64(data + (pb >> 3)) >> (pb & 7)) & 7;
            if (l) sum += 1u << (7 - l);
        }
        if (sum != 128) return false;
    }
    Bits b(data, size, bit);
    bool final;
    int type;
    uint32_t stored;
    if (parse_header(b, t, final, type, stored)) return false;
    if (!t.lit_complete || !t.dist_complete) return false;
    size_t n = 0;
    if (decode_body<false>(b, t, none, n, 4u << 20)) return false;
    if (n < 64) return false;                                            // zlib does not emit tiny dynamic blocks mid-stream
    // what follows must look like a block header again
    if (!b.need(3)) return false;
    const uint32_t h = (uint32_t)(b.bb & 7);
    const int ntype = (int)(h >> 1);
    if (ntype == 3) return false;
    if (ntype == 2) {
        Bits nb(data, size, b.tell());
        Tables &t2 = t;                                                  // tables are scratch here
        bool f2; int ty2; uint32_t st2;
        if (parse_header(nb, t2, f2, ty2, st2)) return false;
        if (!t2.lit_complete || !t2.dist_complete) return false;
    } else if (ntype == 0) {
        Bits nb(data, size, b.tell());
        bool f2; int ty2; uint32_t st2;
        if (parse_header(nb, t, f2, ty2, st2)) return false;
    }
    return true;
}

int64_t find_block(const uint8_t *data, size_t size, uint64_t from_bit, uint64_t to_bit, Tables &t) {
    std::vector<uint16_t> none;
    const uint64_t last = size > 64 ? (uint64_t)(size - 64) * 8 : 0;     // 8-byte peeks stay inside the buffer
    to_bit = std::min(to_bit, last);
    for (uint64_t bit = from_bit; bit < to_bit; bit++)
        if (block_starts_at(data, size, bit, t, none)) return (int64_t)bit;
    return -1;
}

void run_parallel(int threads, size_t n, const std::function<void(size_t, int)> &fn) {
    if (n == 0) return;
    const int k = (int)std::min<size_t>((size_t)std::max(threads, 1), n);
    if (k == 1) { for (size_t i = 0; i < n; i++) fn(i, 0); return; }
    std::atomic<size_t> next{0};
    std::vector<std::thread> th;
    for (int w = 0; w < k; w++)
        th.emplace_back([&, w] { for (size_t i; (i = next.fetch_add(1)) < n;) fn(i, w); });
    for (auto &x : th) x.join();
}

// bytes of out symbols [a, b) of a piece, `win` = the (<= 32 KB) history before the piece; false = a reference beyond it
bool translate(const Piece &p, size_t a, size_t b, const uint8_t *win, size_t win_len, uint8_t *dst) {
    const uint16_t *s = p.sym.data() + kWin;
    bool ok = true;
    size_t i = a;
    auto one = [&](size_t k) {
        const uint16_t v = s[k];
        if (v < 256) dst[k - a] = (uint8_t)v;
        else {
            const size_t back = kWin - (size_t)(v - kPlaceholder);       // distance before the piece start
            if (back > win_len) { ok = false; dst[k - a] = 0; }
            else dst[k - a] = win[win_len - back];
        }
    };
    if (win_len == kWin) {
        // full window: every placeholder resolves, so the piece is one table gather (about 1 in 6 symbols of a FASTQ
        // piece is a placeholder: copies of copies keep them alive, a branch per symbol would mispredict constantly)
        std::vector<uint8_t> lut(256 + kWin);
        for (int v = 0; v < 256; v++) lut[(size_t)v] = (uint8_t)v;
        memcpy(lut.data() + 256, win, kWin);
        const uint8_t *const L = lut.data();
        uint8_t *d = dst - a;
        for (; i + 4 <= b; i += 4) { d[i] = L[s[i]]; d[i + 1] = L[s[i + 1]]; d[i + 2] = L[s[i + 2]]; d[i + 3] = L[s[i + 3]]; }
        for (; i < b; i++) d[i] = L[s[i]];
        return true;
    }
    for (; i < b; i++) one(i);
    return ok;
}

}  // namespace

struct ParallelGunzipState {                 // buffers reused from round to round
    std::vector<Piece> pieces;
    std::vector<Tables> tables;
    std::vector<std::vector<uint8_t>> windows;
};

ParallelGunzip::ParallelGunzip(const uint8_t *data, size_t size, int threads) : data_(data), size_(size), threads_(std::max(threads, 1)) {
    state_ = new ParallelGunzipState();
    state_->tables.resize((size_t)threads_);
}

ParallelGunzip::~ParallelGunzip() {
    delete seq_;
    delete state_;
}

bool ParallelGunzip::fail(const std::string &m) {
    if (st_ != FAILED) err_ = m;
    st_ = FAILED;
    return false;
}

bool ParallelGunzip::begin_member() {
    const uint8_t *ip = data_ + pos_byte_, *end = data_ + size_;
    while (ip < end && *ip == 0 && any_member_) ip++;                    // zero padding
    if (ip == end) {
        if (!any_member_) return fail("empty gzip file");
        st_ = DONE;
        return true;
    }
    if (end - ip < 10) return fail("truncated gzip header");
    if (ip[0] != 0x1f || ip[1] != 0x8b) return fail(any_member_ ? "trailing garbage after gzip member" : "not a gzip file");
    if (ip[2] != 8) return fail("unknown gzip compression method");
    const int flg = ip[3];
    if (flg & 0xE0) return fail("reserved gzip flag bits set");
    ip += 10;
    if (flg & 4) {
        if (end - ip < 2) return fail("truncated gzip header");
        size_t xlen = ip[0] | (ip[1] << 8);
        ip += 2;
        if ((size_t)(end - ip) < xlen) return fail("truncated gzip header");
        ip += xlen;
    }
    for (int bit = 8; bit <= 16; bit <<= 1)
        if (flg & bit) {
            while (ip < end && *ip) ip++;
            if (ip == end) return fail("truncated gzip header");
            ip++;
        }
    if (flg & 2) {
        if (end - ip < 2) return fail("truncated gzip header");
        ip += 2;
    }
    any_member_ = true;
    crc_ = 0; member_out_ = 0;
    window_.clear();
    cur_bit_ = (uint64_t)(ip - data_) * 8;
    st_ = IN_MEMBER;
    return true;
}

bool ParallelGunzip::end_member() {
    const size_t at = (size_t)((cur_bit_ + 7) / 8);
    if (size_ - at < 8) return fail("truncated gzip trailer");
    const uint8_t *ip = data_ + at;
    const uint32_t want_crc = ip[0] | (ip[1] << 8) | (ip[2] << 16) | ((uint32_t)ip[3] << 24);
    const uint32_t want_len = ip[4] | (ip[5] << 8) | (ip[6] << 16) | ((uint32_t)ip[7] << 24);
    if (want_crc != crc_) return fail("gzip CRC-32 mismatch");
    if (want_len != (uint32_t)member_out_) return fail("gzip length mismatch");
    pos_byte_ = at + 8;
    st_ = AT_MEMBER;
    small_members_ = member_out_ < (4u << 20) ? small_members_ + 1 : 0;
    if (small_members_ >= 4) {                                           // bgzip-like file: rounds cannot pay off, go sequential
        while (pos_byte_ < size_ && data_[pos_byte_] == 0) pos_byte_++;
        if (pos_byte_ == size_) { st_ = DONE; return true; }
        seq_ = new GzipInflater(data_ + pos_byte_, size_ - pos_byte_);
        st_ = SEQUENTIAL;
    }
    return true;
}

bool ParallelGunzip::next_round() {
    round_len_ = 0; rd_ = 0;
    if (st_ == AT_MEMBER && !begin_member()) return false;
    if (st_ != IN_MEMBER) return st_ != FAILED;
    ParallelGunzipState &S = *state_;
    stats_.rounds++;
    // ---- 1. search ----
    const size_t byte0 = (size_t)(cur_bit_ / 8);
    const size_t n_targets = (size_t)threads_ * 2;
    std::vector<int64_t> found(n_targets, -1);
    std::vector<size_t> target;
    for (size_t k = 1; k <= n_targets; k++) {
        const size_t tb = byte0 + k * kChunk;
        if (tb + kChunk / 4 >= size_) break;
        target.push_back(tb);
    }
    run_parallel(threads_, target.size(), [&](size_t i, int w) {
        found[i] = find_block(data_, size_, (uint64_t)target[i] * 8, (uint64_t)(target[i] + kChunk) * 8, S.tables[(size_t)w]);
    });
    std::vector<uint64_t> starts{cur_bit_};
    for (size_t i = 0; i < target.size(); i++)
        if (found[i] >= 0 && (uint64_t)found[i] > starts.back()) starts.push_back((uint64_t)found[i]);
    const uint64_t force_stop = (uint64_t)(byte0 + (target.size() + 1) * kChunk) * 8;
    // ---- 2. speculative decode ----
    if (S.pieces.size() < starts.size()) S.pieces.resize(starts.size());
    run_parallel(threads_, starts.size(), [&](size_t i, int w) {
        decode_piece(data_, size_, starts[i], starts.data() + i + 1, starts.size() - i - 1, force_stop, S.tables[(size_t)w], S.pieces[i]);
    });
    // ---- 3. chain the true pieces ----
    std::vector<size_t> chain;
    size_t i = 0;
    for (;;) {
        Piece &p = S.pieces[i];
        if (p.err) {
            if (chain.empty()) return fail(p.err);                       // decoded from a known-true header: a real stream error
            break;                                                       // a later piece failed: decode it again as the first piece of the next round
        }
        chain.push_back(i);
        if (p.member_end) break;
        size_t j = i + 1;
        while (j < starts.size() && starts[j] < p.end_bit) { j++; stats_.false_starts++; }
        if (j < starts.size() && starts[j] == p.end_bit) { i = j; continue; }
        break;                                                           // ended on a forced stop: next round starts there
    }
    stats_.pieces += chain.size();
    // windows in front of every piece (sequential, 32 KB each), then translation + CRC in parallel
    if (S.windows.size() < chain.size() + 1) S.windows.resize(chain.size() + 1);
    S.windows[0] = window_;
    std::vector<size_t> offs(chain.size() + 1, 0);
    std::vector<uint8_t> tail;
    for (size_t c = 0; c < chain.size(); c++) {
        const Piece &p = S.pieces[chain[c]];
        offs[c + 1] = offs[c] + p.n_out;
        const std::vector<uint8_t> &w = S.windows[c];
        std::vector<uint8_t> &nw = S.windows[c + 1];
        const size_t take = std::min(p.n_out, kWin);
        tail.resize(take);
        if (!translate(p, p.n_out - take, p.n_out, w.data(), w.size(), tail.data())) return fail("invalid distance too far back");
        if (take == kWin) nw = tail;
        else {
            const size_t keep = std::min(w.size(), kWin - take);
            nw.assign(w.end() - (long)keep, w.end());
            nw.insert(nw.end(), tail.begin(), tail.end());
        }
    }
    if (round_.size() < offs.back()) round_.resize(offs.back() + offs.back() / 4);
    round_len_ = offs.back();
    std::atomic<bool> bad{false};
    run_parallel(threads_, chain.size(), [&](size_t c, int) {
        Piece &p = S.pieces[chain[c]];
        if (!translate(p, 0, p.n_out, S.windows[c].data(), S.windows[c].size(), round_.data() + offs[c])) bad = true;
        p.crc = (uint32_t)crc32(0, round_.data() + offs[c], (uInt)p.n_out);
    });
    if (bad) return fail("invalid distance too far back");
    for (size_t c = 0; c < chain.size(); c++) {
        const Piece &p = S.pieces[chain[c]];
        crc_ = (uint32_t)crc32_combine(crc_, p.crc, (z_off_t)p.n_out);
        member_out_ += p.n_out;
    }
    window_ = S.windows[chain.size()];
    const Piece &last = S.pieces[chain.back()];
    cur_bit_ = last.end_bit;
    if (last.member_end && !end_member()) return false;
    return true;
}

long ParallelGunzip::read(uint8_t *dst, size_t n) {
    size_t total = 0;
    while (total < n) {
        if (rd_ < round_len_) {                                          // bytes of the current round first
            const size_t k = std::min(n - total, round_len_ - rd_);
            memcpy(dst + total, round_.data() + rd_, k);
            rd_ += k; total += k;
            continue;
        }
        if (st_ == FAILED) return -1;
        if (st_ == DONE) break;
        if (st_ == SEQUENTIAL) {
            long g = seq_->read(dst + total, n - total);
            if (g < 0) { fail(seq_->error()); return -1; }
            if (g == 0) break;
            total += (size_t)g;
            continue;
        }
        if (!next_round()) return -1;
    }
    return (long)total;
}

}  // namespace aqc
