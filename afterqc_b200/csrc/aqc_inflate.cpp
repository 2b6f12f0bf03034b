// aqc_inflate.cpp -- see aqc_inflate.hpp.  DEFLATE per RFC 1951, gzip framing per RFC 1952.
#include "aqc_inflate.hpp"

#include <zlib.h>   // crc32() only

#include <algorithm>
#include <cstring>

namespace aqc {
namespace {

constexpr uint32_t F_LIT = 0x8000, F_SUB = 0x4000, F_EOB = 0x2000, F_BAD = 0x1000;
constexpr size_t kMargin = 320;         // a match (258) + word over-copy may run past the piece limit

const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t kPreOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

inline uint64_t load64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
inline void store64(uint8_t *p, uint64_t v) { memcpy(p, &v, 8); }

inline uint32_t reverse_bits(uint32_t code, int len) {
    uint32_t r = 0;
    for (int i = 0; i < len; i++) { r = (r << 1) | (code & 1); code >>= 1; }
    return r;
}

// table entry of a decoded symbol (bits 0-7: code bits to consume, 8-11: extra bits, 12-15: flags, 16-31: value)
inline uint32_t symbol_entry(int sym, bool is_dist, bool is_pre) {
    if (is_pre) return (uint32_t)sym << 16;
    if (is_dist) {
        if (sym >= 30) return F_BAD;
        return ((uint32_t)kDistBase[sym] << 16) | ((uint32_t)kDistExtra[sym] << 8);
    }
    if (sym < 256) return F_LIT | ((uint32_t)sym << 16);
    if (sym == 256) return F_EOB;
    if (sym >= 286) return F_BAD;
    return ((uint32_t)kLenBase[sym - 257] << 16) | ((uint32_t)kLenExtra[sym - 257] << 8);
}

}  // namespace

GzipInflater::GzipInflater(const uint8_t *data, size_t size) : in_(data), ip_(data), in_end_(data + size) {
    buf_.resize(kWindow + kChunk + kMargin);
    op_ = rd_ = crc_pos_ = lo_ = kWindow;
}

bool GzipInflater::fail(const char *m) {
    if (st_ != FAILED) err_ = m;
    st_ = FAILED;
    return false;
}

// Canonical Huffman decode table: root table of 2^root_bits entries, codes longer than the root go through a
// sub-table per root prefix (sized by the longest code sharing the prefix).  kind: 0 literal/length, 1 distance, 2 precode.
bool GzipInflater::build(const uint8_t *lens, int n, int root_bits, int kind, std::vector<uint32_t> &tab) {
    const char *e = build_table(lens, n, root_bits, kind, tab);
    return e ? fail(e) : true;
}

const char *GzipInflater::build_table(const uint8_t *lens, int n, int root_bits, int kind, std::vector<uint32_t> &tab) {
    int count[16] = {0};
    for (int i = 0; i < n; i++) count[lens[i]]++;
    count[0] = 0;
    int left = 1;
    for (int l = 1; l <= 15; l++) {
        left = left * 2 - count[l];
        if (left < 0) return "over-subscribed Huffman code";
    }
    uint32_t next[16];
    uint32_t code = 0;
    for (int l = 1; l <= 15; l++) { code = (code + (uint32_t)count[l - 1]) << 1; next[l] = code; }
    const uint32_t root_size = 1u << root_bits, root_mask = root_size - 1;
    tab.assign(root_size, 0);
    // pass 1: longest code per root prefix
    std::vector<uint8_t> sub_bits(root_size, 0);
    uint32_t nx[16];
    memcpy(nx, next, sizeof nx);
    for (int s = 0; s < n; s++) {
        int l = lens[s];
        if (l <= root_bits) { if (l) nx[l]++; continue; }
        uint32_t rev = reverse_bits(nx[l]++, l);
        uint8_t &b = sub_bits[rev & root_mask];
        b = std::max<uint8_t>(b, (uint8_t)(l - root_bits));
    }
    for (uint32_t p = 0; p < root_size; p++)
        if (sub_bits[p]) {
            uint32_t start = (uint32_t)tab.size();
            if (start + (1u << sub_bits[p]) > 0xFFFF) return "Huffman table too large";
            tab.resize(start + (1u << sub_bits[p]), 0);
            tab[p] = F_SUB | (start << 16) | ((uint32_t)sub_bits[p] << 8) | (uint32_t)root_bits;
        }
    // pass 2: fill
    for (int s = 0; s < n; s++) {
        int l = lens[s];
        if (!l) continue;
        uint32_t rev = reverse_bits(next[l]++, l);
        uint32_t e = symbol_entry(s, kind == 1, kind == 2);
        if (l <= root_bits) {
            e |= (uint32_t)l;
            for (uint32_t i = rev; i < root_size; i += 1u << l) tab[i] = e;
        } else {
            uint32_t ptr = tab[rev & root_mask];
            uint32_t start = ptr >> 16, sb = (ptr >> 8) & 15;
            e |= (uint32_t)(l - root_bits);
            for (uint32_t i = rev >> root_bits; i < (1u << sb); i += 1u << (l - root_bits)) tab[start + i] = e;
        }
    }
    return nullptr;
}

// Two literals per lookup: when the root index holds a whole literal code AND the whole code of the literal that
// follows, the entry carries both bytes (F_LIT | F_SUB, second byte in bits 24-31, bits 0-7 = both code lengths).
// FASTQ text is literal-dominated with 2-4 bit codes, so most literal lookups emit two bytes.
void GzipInflater::pair_literals(std::vector<uint32_t> &t, std::vector<uint32_t> &scratch) {
    const uint32_t root_size = 1u << kLitBits;
    scratch.assign(t.begin(), t.begin() + root_size);
    for (uint32_t i = 0; i < root_size; i++) {
        const uint32_t e1 = t[i];
        if ((e1 & (F_LIT | F_SUB)) != F_LIT) continue;
        const uint32_t l1 = e1 & 0xff;
        if (l1 >= (uint32_t)kLitBits) continue;
        const uint32_t e2 = t[i >> l1];                                  // the unknown high bits read as zeros: valid iff the code fits
        if ((e2 & (F_LIT | F_SUB)) != F_LIT) continue;
        const uint32_t l2 = e2 & 0xff;
        if (l1 + l2 > (uint32_t)kLitBits) continue;
        scratch[i] = F_LIT | F_SUB | (e1 & 0x00FF0000u) | ((e2 & 0x00FF0000u) << 8) | (l1 + l2);
    }
    memcpy(t.data(), scratch.data(), root_size * sizeof(uint32_t));
}

bool GzipInflater::need_bits(int n) {
    while (bitcnt_ < n) {
        if (ip_ >= in_end_) return false;
        bitbuf_ |= (uint64_t)*ip_++ << bitcnt_;
        bitcnt_ += 8;
    }
    return true;
}

// drop to a byte boundary and give whole unread bytes back to the input pointer
void GzipInflater::align_to_byte() {
    int drop = bitcnt_ & 7;
    bitbuf_ >>= drop; bitcnt_ -= drop;
    ip_ -= bitcnt_ >> 3;
    bitbuf_ = 0; bitcnt_ = 0;
}

bool GzipInflater::parse_member_header() {
    // bit buffer is empty here
    while (ip_ < in_end_ && *ip_ == 0 && any_member_) ip_++;            // zero padding between / after members
    if (ip_ == in_end_) {
        if (!any_member_) return fail("empty gzip file");
        st_ = FINISHED;
        return true;
    }
    if (in_end_ - ip_ < 10) return fail("truncated gzip header");
    if (ip_[0] != 0x1f || ip_[1] != 0x8b) return fail(any_member_ ? "trailing garbage after gzip member" : "not a gzip file");
    if (ip_[2] != 8) return fail("unknown gzip compression method");
    const int flg = ip_[3];
    if (flg & 0xE0) return fail("reserved gzip flag bits set");
    ip_ += 10;
    if (flg & 4) {                                                     // FEXTRA
        if (in_end_ - ip_ < 2) return fail("truncated gzip header");
        size_t xlen = ip_[0] | (ip_[1] << 8);
        ip_ += 2;
        if ((size_t)(in_end_ - ip_) < xlen) return fail("truncated gzip header");
        ip_ += xlen;
    }
    for (int bit = 8; bit <= 16; bit <<= 1)                            // FNAME, FCOMMENT: zero-terminated
        if (flg & bit) {
            while (ip_ < in_end_ && *ip_) ip_++;
            if (ip_ == in_end_) return fail("truncated gzip header");
            ip_++;
        }
    if (flg & 2) {                                                     // FHCRC
        if (in_end_ - ip_ < 2) return fail("truncated gzip header");
        ip_ += 2;
    }
    any_member_ = true;
    crc_ = 0; member_out_ = 0;
    crc_pos_ = op_;
    lo_ = op_;                                                         // back-references stay inside the member
    st_ = BLOCK_HEADER;
    return true;
}

bool GzipInflater::parse_block_header() {
    if (!need_bits(3)) return fail("truncated deflate stream");
    final_ = bitbuf_ & 1;
    int type = (bitbuf_ >> 1) & 3;
    bitbuf_ >>= 3; bitcnt_ -= 3;
    if (type == 0) {
        align_to_byte();
        if (in_end_ - ip_ < 4) return fail("truncated stored block");
        uint32_t len = ip_[0] | (ip_[1] << 8), nlen = ip_[2] | (ip_[3] << 8);
        if ((len ^ 0xFFFF) != nlen) return fail("stored block length check failed");
        ip_ += 4;
        stored_left_ = len;
        st_ = STORED;
        return true;
    }
    if (type == 3) return fail("invalid deflate block type");
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
        if (!need_bits(14)) return fail("truncated deflate stream");
        nlit = 257 + (int)(bitbuf_ & 31);
        ndist = 1 + (int)((bitbuf_ >> 5) & 31);
        int npre = 4 + (int)((bitbuf_ >> 10) & 15);
        bitbuf_ >>= 14; bitcnt_ -= 14;
        if (nlit > 286 || ndist > 30) return fail("too many length or distance symbols");
        uint8_t plen[19] = {0};
        for (int i = 0; i < npre; i++) {
            if (!need_bits(3)) return fail("truncated deflate stream");
            plen[kPreOrder[i]] = (uint8_t)(bitbuf_ & 7);
            bitbuf_ >>= 3; bitcnt_ -= 3;
        }
        if (!build(plen, 19, 7, 2, pre_)) return false;
        int i = 0;
        const int total = nlit + ndist;
        while (i < total) {
            need_bits(7 + 7);                                           // code (<= 7) + repeat count (<= 7); may fall short at EOF
            uint32_t e = pre_[bitbuf_ & 127];
            int cl = (int)(e & 0xff);
            if (cl == 0 || cl > bitcnt_) return fail("invalid code lengths set");
            bitbuf_ >>= cl; bitcnt_ -= cl;
            int sym = (int)(e >> 16);
            if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
            int rep, xb;
            uint8_t val = 0;
            if (sym == 16) {
                if (i == 0) return fail("invalid bit length repeat");
                val = lens[i - 1]; xb = 2; rep = 3;
            } else if (sym == 17) { xb = 3; rep = 3; }
            else { xb = 7; rep = 11; }
            if (xb > bitcnt_) return fail("truncated deflate stream");
            rep += (int)(bitbuf_ & ((1u << xb) - 1));
            bitbuf_ >>= xb; bitcnt_ -= xb;
            if (i + rep > total) return fail("invalid bit length repeat");
            while (rep--) lens[i++] = val;
        }
        if (lens[256] == 0) return fail("missing end-of-block code");
        // distance lengths follow the literal/length lengths directly: move them to the fixed slot
        memmove(lens + 288, lens + nlit, (size_t)ndist);
        for (int k = nlit; k < 288; k++) lens[k] = 0;
    }
    if (!build(lens, nlit, kLitBits, 0, lit_)) return false;
    if (!build(lens + 288, ndist, kDistBits, 1, dist_)) return false;
    pair_literals(lit_, pair_);
    st_ = HUFFMAN;
    return true;
}

bool GzipInflater::parse_trailer() {
    align_to_byte();
    if (in_end_ - ip_ < 8) return fail("truncated gzip trailer");
    crc_ = (uint32_t)crc32(crc_, buf_.data() + crc_pos_, (uInt)(op_ - crc_pos_));
    member_out_ += op_ - crc_pos_;
    crc_pos_ = op_;
    uint32_t want_crc = ip_[0] | (ip_[1] << 8) | (ip_[2] << 16) | ((uint32_t)ip_[3] << 24);
    uint32_t want_len = ip_[4] | (ip_[5] << 8) | (ip_[6] << 16) | ((uint32_t)ip_[7] << 24);
    ip_ += 8;
    if (want_crc != crc_) return fail("gzip CRC-32 mismatch");
    if (want_len != (uint32_t)member_out_) return fail("gzip length mismatch");
    st_ = MEMBER_HEADER;
    return true;
}

// One Huffman-coded block (or the part of it that fits below `limit`).
bool GzipInflater::decode_huffman(size_t limit) {
    uint8_t *const out = buf_.data();
    const uint32_t *const lit = lit_.data();
    const uint32_t *const dst = dist_.data();
    size_t op = op_;
    uint64_t bb = bitbuf_;
    int bc = bitcnt_;
    const uint8_t *ip = ip_;
    const uint8_t *const in_end = in_end_;
    const uint32_t lit_mask = (1u << kLitBits) - 1, dist_mask = (1u << kDistBits) - 1;
    const char *why = nullptr;

    // ---- fast loop: >= 32 input bytes ahead, refills are one unaligned 8-byte load; the next literal/length entry is
    // looked up before the previous match is copied ----
#define AQC_REFILL() do { bb |= load64(ip) << bc; ip += (63 - bc) >> 3; bc |= 56; } while (0)
#define AQC_PUT_LITERALS(e) do { out[op] = (uint8_t)((e) >> 16); out[op + 1] = (uint8_t)((e) >> 24); op += 1 + (((e) >> 14) & 1); } while (0)
    if (op < limit && (size_t)(in_end - ip) >= 32) {
        AQC_REFILL();
        uint32_t e = lit[bb & lit_mask];
        bool leave = false;
        for (;;) {
            if (e & F_LIT) {                                            // up to four lookups (<= 44 bits) per refill
                int k = 0;
                do {
                    bb >>= (e & 0xff); bc -= (int)(e & 0xff);
                    AQC_PUT_LITERALS(e);
                    e = lit[bb & lit_mask];
                } while ((e & F_LIT) && ++k < 4);
                AQC_REFILL();                                           // e stays valid: a refill only adds high bits
                if (e & F_LIT) {
                    if (op >= limit || (size_t)(in_end - ip) < 32) break;
                    continue;
                }
            }
            if (e & F_SUB) {
                bb >>= kLitBits; bc -= kLitBits;
                e = lit[(e >> 16) + (bb & ((1u << ((e >> 8) & 15)) - 1))];
                if (e & F_LIT) {                                        // a long literal code
                    bb >>= (e & 0xff); bc -= (int)(e & 0xff);
                    out[op++] = (uint8_t)(e >> 16);
                    AQC_REFILL();
                    e = lit[bb & lit_mask];
                    if (op >= limit || (size_t)(in_end - ip) < 32) break;
                    continue;
                }
            }
            if ((e & 0xff) == 0 || (e & F_BAD)) { why = "invalid literal/length code"; break; }
            if (e & F_EOB) { bb >>= (e & 0xff); bc -= (int)(e & 0xff); st_ = final_ ? MEMBER_TRAILER : BLOCK_HEADER; leave = true; break; }
            // code + extra bits leave the bit buffer in one shift; the extra bits are read from the saved copy
            uint32_t xb = (e >> 8) & 15, tot = (e & 0xff) + xb;
            uint64_t saved = bb;
            bb >>= tot; bc -= (int)tot;
            const uint32_t len = (e >> 16) + (uint32_t)((saved >> (e & 0xff)) & ((1u << xb) - 1));
            uint32_t d = dst[bb & dist_mask];
            if (d & F_SUB) {
                bb >>= kDistBits; bc -= kDistBits;
                d = dst[(d >> 16) + (bb & ((1u << ((d >> 8) & 15)) - 1))];
            }
            if ((d & 0xff) == 0 || (d & F_BAD)) { why = "invalid distance code"; break; }
            xb = (d >> 8) & 15; tot = (d & 0xff) + xb;
            saved = bb;
            bb >>= tot; bc -= (int)tot;
            const uint32_t dist = (d >> 16) + (uint32_t)((saved >> (d & 0xff)) & ((1u << xb) - 1));
            if (dist > op - lo_) { why = "invalid distance too far back"; break; }
            AQC_REFILL();
            e = lit[bb & lit_mask];                                     // next symbol's entry, in flight during the copy
            uint8_t *o = out + op;
            const uint8_t *sp = o - dist;
            uint8_t *const oend = o + len;
            if (dist >= 8) {
                store64(o, load64(sp)); store64(o + 8, load64(sp + 8));
                if (len > 16) {
                    o += 16; sp += 16;
                    do { store64(o, load64(sp)); o += 8; sp += 8; } while (o < oend);
                }
            } else if (dist == 1) {
                const uint64_t v = 0x0101010101010101ull * sp[0];
                do { store64(o, v); o += 8; } while (o < oend);
            } else {
                do { *o++ = *sp++; } while (o < oend);
            }
            op += len;
            if (op >= limit || (size_t)(in_end - ip) < 32) break;
        }
        if (why || leave) {
            op_ = op; bitbuf_ = bb; bitcnt_ = bc; ip_ = ip;
            if (why) return fail(why);
            return true;
        }
    }
#undef AQC_REFILL
#undef AQC_PUT_LITERALS

    // ---- careful loop: the last bytes of the input, every bit count checked ----
#define AQC_CAREFUL_REFILL() do { while (bc <= 56 && ip < in_end) { bb |= (uint64_t)*ip++ << bc; bc += 8; } } while (0)
    while (op < limit) {
        if ((size_t)(in_end - ip) >= 32) {                          // far from the end again (after a stored block): back to the fast loop
            op_ = op; bitbuf_ = bb; bitcnt_ = bc; ip_ = ip;
            return true;
        }
        const bool fast = false;
        AQC_CAREFUL_REFILL();
        uint32_t e = lit[bb & lit_mask];
        if (e & F_LIT) {                                                // one literal, or two (F_SUB set): both bytes are always stored
            if (!fast && (int)(e & 0xff) > bc) { why = "truncated deflate stream"; break; }
            bb >>= (e & 0xff); bc -= (int)(e & 0xff);
            out[op] = (uint8_t)(e >> 16); out[op + 1] = (uint8_t)(e >> 24);
            op += 1 + ((e >> 14) & 1);
            if (fast) {                                                 // >= 45 bits left: up to three more lookups without a refill
                for (int k = 0; k < 3; k++) {
                    e = lit[bb & lit_mask];
                    if (!(e & F_LIT)) break;
                    bb >>= (e & 0xff); bc -= (int)(e & 0xff);
                    out[op] = (uint8_t)(e >> 16); out[op + 1] = (uint8_t)(e >> 24);
                    op += 1 + ((e >> 14) & 1);
                }
            }
            continue;
        }
        if (e & F_SUB) {
            bb >>= kLitBits; bc -= kLitBits;
            e = lit[(e >> 16) + (bb & ((1u << ((e >> 8) & 15)) - 1))];
            if (e & F_LIT) {                                            // a long literal code
                if (!fast && (int)(e & 0xff) > bc) { why = "truncated deflate stream"; break; }
                bb >>= (e & 0xff); bc -= (int)(e & 0xff);
                out[op++] = (uint8_t)(e >> 16);
                continue;
            }
        }
        if (!fast && (int)(e & 0xff) > bc) { why = "truncated deflate stream"; break; }
        if ((e & 0xff) == 0 || (e & F_BAD)) { why = "invalid literal/length code"; break; }
        bb >>= (e & 0xff); bc -= (int)(e & 0xff);
        if (e & F_EOB) { st_ = final_ ? MEMBER_TRAILER : BLOCK_HEADER; break; }
        // length + distance
        uint32_t xb = (e >> 8) & 15;
        if (!fast) { AQC_CAREFUL_REFILL(); if ((int)xb > bc) { why = "truncated deflate stream"; break; } }
        uint32_t len = (e >> 16) + (uint32_t)(bb & ((1u << xb) - 1));
        bb >>= xb; bc -= (int)xb;
        if (!fast) AQC_CAREFUL_REFILL();
        uint32_t d = dst[bb & dist_mask];
        if (d & F_SUB) {
            bb >>= kDistBits; bc -= kDistBits;
            d = dst[(d >> 16) + (bb & ((1u << ((d >> 8) & 15)) - 1))];
        }
        if ((d & 0xff) == 0 || (d & F_BAD)) { why = "invalid distance code"; break; }
        if (!fast && (int)(d & 0xff) > bc) { why = "truncated deflate stream"; break; }
        bb >>= (d & 0xff); bc -= (int)(d & 0xff);
        xb = (d >> 8) & 15;
        if (!fast) { AQC_CAREFUL_REFILL(); if ((int)xb > bc) { why = "truncated deflate stream"; break; } }
        uint32_t dist = (d >> 16) + (uint32_t)(bb & ((1u << xb) - 1));
        bb >>= xb; bc -= (int)xb;
        if (dist > op - lo_) { why = "invalid distance too far back"; break; }
        uint8_t *o = out + op;
        const uint8_t *s = o - dist;
        uint8_t *const oend = o + len;
        if (dist >= 8) {
            do { store64(o, load64(s)); o += 8; s += 8; } while (o < oend);
        } else if (dist == 1) {
            const uint64_t v = 0x0101010101010101ull * s[0];
            do { store64(o, v); o += 8; } while (o < oend);
        } else {
            do { *o++ = *s++; } while (o < oend);
        }
        op += len;
    }
#undef AQC_CAREFUL_REFILL
    op_ = op; bitbuf_ = bb; bitcnt_ = bc; ip_ = ip;
    if (why) return fail(why);
    return true;
}

// Decode until buf_ holds a full piece (or the stream ends).  false = failed.
bool GzipInflater::decode_piece() {
    const size_t limit = kWindow + kChunk;
    while (op_ < limit) {
        switch (st_) {
            case MEMBER_HEADER: if (!parse_member_header()) return false; break;
            case BLOCK_HEADER: if (!parse_block_header()) return false; break;
            case STORED: {
                size_t k = std::min<size_t>({(size_t)stored_left_, limit - op_, (size_t)(in_end_ - ip_)});
                memcpy(buf_.data() + op_, ip_, k);
                ip_ += k; op_ += k; stored_left_ -= (uint32_t)k;
                if (stored_left_ == 0) st_ = final_ ? MEMBER_TRAILER : BLOCK_HEADER;
                else if (ip_ == in_end_) return fail("truncated stored block");
                break;
            }
            case HUFFMAN: if (!decode_huffman(limit)) return false; break;
            case MEMBER_TRAILER: if (!parse_trailer()) return false; break;
            case FINISHED: return true;
            case FAILED: return false;
        }
    }
    return true;
}

long GzipInflater::read(uint8_t *dstp, size_t n) {
    size_t total = 0;
    while (total < n) {
        if (rd_ == op_) {
            if (st_ == FAILED) return -1;
            if (st_ == FINISHED) break;
            // account the bytes of the running member, slide the window, decode the next piece
            if (op_ > crc_pos_) {
                crc_ = (uint32_t)crc32(crc_, buf_.data() + crc_pos_, (uInt)(op_ - crc_pos_));
                member_out_ += op_ - crc_pos_;
            }
            const size_t shift = op_ - kWindow;
            if (shift) memmove(buf_.data(), buf_.data() + shift, kWindow);
            lo_ = lo_ > shift ? lo_ - shift : 0;
            op_ = rd_ = crc_pos_ = kWindow;
            if (!decode_piece()) return -1;
            if (rd_ == op_) { if (st_ == FINISHED) break; continue; }
        }
        size_t k = std::min(n - total, op_ - rd_);
        memcpy(dstp + total, buf_.data() + rd_, k);
        rd_ += k; total += k;
    }
    return (long)total;
}

}  // namespace aqc
