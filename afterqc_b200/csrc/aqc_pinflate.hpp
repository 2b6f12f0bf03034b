// aqc_pinflate.hpp -- multi-threaded decoder for ONE gzip stream ("parallel gz inflate").
//
// A gzip member is a chain of DEFLATE blocks whose back-references reach 32 KB into earlier output, so it cannot simply
// be cut into pieces.  The decoder works in rounds (the approach of pugz, Kerbiriou & Chikhi 2019, restated here):
//   1. search: from evenly spaced byte offsets, find the bit position of the next block header by trial parsing
//      (dynamic-Huffman header with complete codes, one whole block of text-only literals, a plausible next header);
//   2. decode: every found position is decoded speculatively on its own thread into 16-bit symbols, the unknown 32 KB
//      of history in front of it represented by placeholder symbols (256 + window index) that propagate through copies;
//   3. chain: starting from the one position known to be true, a piece is accepted only if the previous piece ended
//      EXACTLY on its start bit (otherwise the search hit was false and the previous decoder simply kept going);
//      the 32 KB window is carried from piece to piece, placeholders are resolved (in parallel), CRC-32s are combined.
// The heuristics of step 1 can therefore only cost time, never correctness, and every member's CRC-32 and length are
// verified as in the sequential decoder.  Many small members (bgzip-like files) and small inputs fall back to the
// sequential GzipInflater.  No CUDA here.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace aqc {

class GzipInflater;
struct ParallelGunzipState;

class ParallelGunzip {
  public:
    ParallelGunzip(const uint8_t *data, size_t size, int threads);
    ~ParallelGunzip();
    long read(uint8_t *dst, size_t n);          // bytes produced, 0 at the end, -1 on a corrupt stream (error())
    const std::string &error() const { return err_; }
    static constexpr size_t kMinSize = 8u << 20; // smaller files are not worth the threads

    struct Stats { uint64_t rounds = 0, pieces = 0, false_starts = 0, sequential_bytes = 0; };
    const Stats &stats() const { return stats_; }

  private:
    bool next_round();                          // fills round_ with decoded bytes; false = failed
    bool begin_member();                        // gzip header at pos_byte_ -> cur_bit_
    bool end_member();                          // trailer check at cur_bit_
    bool fail(const std::string &m);

    const uint8_t *data_;
    size_t size_;
    int threads_;
    std::string err_;
    Stats stats_;
    // stream position
    enum { AT_MEMBER, IN_MEMBER, DONE, FAILED, SEQUENTIAL } st_ = AT_MEMBER;
    size_t pos_byte_ = 0;                       // AT_MEMBER: where the next gzip header starts
    uint64_t cur_bit_ = 0;                      // IN_MEMBER: a block header known to be true
    std::vector<uint8_t> window_;               // last <= 32 KB of the member's output
    uint32_t crc_ = 0;
    uint64_t member_out_ = 0;
    bool any_member_ = false;
    int small_members_ = 0;
    GzipInflater *seq_ = nullptr;               // fallback for the rest of the file
    ParallelGunzipState *state_ = nullptr;      // per-round buffers, reused
    // decoded bytes of the current round
    std::vector<uint8_t> round_;
    size_t round_len_ = 0, rd_ = 0;
};

}  // namespace aqc
