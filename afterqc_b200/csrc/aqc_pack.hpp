// aqc_pack.hpp -- transport encoding of the columns for the host-buffer entry (AQC_BATCH_PACK_BASES, AQC_BATCH_PACK_QUALS).
//
// The host-buffer path of aqc_filter_pairs / aqc_stat_reads is PCIe bound: 4 x L bytes per pair cross
// the bus.  Bases are two bits of information: host threads pack a chunk's base columns to 2 bits per base before the copy
// (A0 C1 T2 G3 = bits 1-2 of the ASCII byte, the code the kernels use), every byte that is not A,C,G,T travels in an
// exception list (position, byte), and unpack_bases_kernel / apply_exceptions_kernel restore the byte column in HBM, where
// the unchanged kernels read it.  Lossless for any input; a chunk with too many exceptions is copied as bytes.  This is
// transport only -- nothing of the reference's algorithm is computed on the host.
#pragma once
#include <cstddef>
#include <cstdint>

namespace aqc_pack {

struct Pool;                                   // persistent worker threads (one pool per context)
Pool *pool_create(int threads);                // threads <= 0: hardware concurrency, capped (AQC_PACK_THREADS overrides)
void pool_destroy(Pool *p);
int pool_threads(const Pool *p);

// dst[k] = codes of src[4k..4k+3], base j in bits 2*(j & 3); dst holds (n + 3) / 4 bytes (a partial last byte is zero-filled).
// Returns false (dst and the lists undefined) when more than max_exc bytes are not A,C,G,T.
bool pack_bases(Pool *p, const uint8_t *src, size_t n, uint8_t *dst, uint32_t *exc_pos, uint8_t *exc_val, size_t max_exc, size_t *n_exc);

// Quality bytes, 6 bits each: code = byte - 33 ('!' .. '`', i.e. Phred+33 qualities 0..63); four codes a,b,c,d become the three
// bytes of a | b << 6 | c << 12 | d << 18 (little endian).  dst holds 3 * ((n + 3) / 4) bytes.  Bytes outside that range are
// exceptions, as for the bases.
enum { KIND_BASES = 0, KIND_QUALS = 1 };
constexpr int MAX_COLUMNS = 4;
inline size_t packed_bytes(int kind, size_t n) { return kind == KIND_BASES ? (n + 3) / 4 : 3 * ((n + 3) / 4); }

// up to four columns (bases and qualities of both mates of a chunk) in ONE dispatch of the pool; ok as pack_bases' return value
struct Column {
    int kind;
    const uint8_t *src; size_t n; uint8_t *dst;
    uint32_t *exc_pos; uint8_t *exc_val; size_t max_exc;
    size_t n_exc; bool ok;
};
void pack_columns(Pool *p, Column *cols, int n_cols);

}  // namespace aqc_pack
