// aqc_inflate.hpp -- gzip (RFC 1952) / DEFLATE (RFC 1951) decoder of the FASTQ reader.
// zlib's inflate (~240 MB/s on FASTQ text) is what bounds .fq.gz ingest; this decoder keeps a 64-bit bit buffer, decodes
// through one-lookup tables (11-bit literal/length root, 8-bit distance root, sub-tables for longer codes), emits up to
// three literals per refill and copies matches a word at a time.  It reads from a memory range (the mapped file),
// produces output in bounded pieces (resumable between symbols), walks concatenated members and verifies every member's
// CRC-32 and length, so a decoding defect cannot pass silently.  No CUDA here.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace aqc {

class GzipInflater {
  public:
    GzipInflater(const uint8_t *data, size_t size);
    // Decompress up to n bytes into dst.  Returns the number of bytes produced (0 = end of the last member),
    // or -1 on a corrupt / truncated stream (error()).
    long read(uint8_t *dst, size_t n);
    const std::string &error() const { return err_; }

    // shared with the parallel decoder (aqc_pinflate.cpp)
    static constexpr int kLitBits = 11, kDistBits = 8;
    // canonical Huffman decode table (kind 0 literal/length, 1 distance, 2 precode); nullptr or an error text
    static const char *build_table(const uint8_t *lens, int n, int root_bits, int kind, std::vector<uint32_t> &tab);
    // two literals per root entry where both codes fit the root index
    static void pair_literals(std::vector<uint32_t> &lit, std::vector<uint32_t> &scratch);

  private:
    static constexpr size_t kWindow = 32768;
    static constexpr size_t kChunk = 1u << 20;       // bytes decoded per inner call (after the window)

    enum State { MEMBER_HEADER, BLOCK_HEADER, STORED, HUFFMAN, MEMBER_TRAILER, FINISHED, FAILED };

    bool fail(const char *m);
    bool need_bits(int n);                  // careful refill: false = input exhausted
    void align_to_byte();
    bool parse_member_header();
    bool parse_block_header();
    bool parse_trailer();
    bool build(const uint8_t *lens, int n, int root_bits, int kind, std::vector<uint32_t> &tab);
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
    __attribute__((target_clones("bmi2", "default")))     // shrx/shlx/bzhi for the variable shifts where the CPU has them
#endif
    bool decode_huffman(size_t limit);      // false = failed
    bool decode_piece();                    // fill buf_ after the window; false = failed

    const uint8_t *in_, *ip_, *in_end_;
    uint64_t bitbuf_ = 0;
    int bitcnt_ = 0;
    State st_ = MEMBER_HEADER;
    bool final_ = false;
    uint32_t stored_left_ = 0;
    std::vector<uint32_t> lit_, dist_, pre_, pair_;
    std::vector<uint8_t> buf_;              // [32 KB window | piece | margin]
    size_t op_ = 0;                         // write position in buf_
    size_t rd_ = 0;                         // next byte of buf_ to hand out
    size_t crc_pos_ = 0;                    // bytes of the running member before this index are in crc_ / member_out_
    size_t lo_ = 0;                         // lowest index a back-reference may reach (start of the member, if still in buf_)
    uint32_t crc_ = 0;
    uint64_t member_out_ = 0;
    bool any_member_ = false;
    std::string err_;
};

}  // namespace aqc
