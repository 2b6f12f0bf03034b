// aqc_parse.cuh -- FASTQ text -> packed columns on the device (fastq.Reader.nextRead, fastq.py:37-49, for a whole buffer).
//
// The reference reads a record as four readline()s, rstrip()s each and stops at the first empty line.  Here a buffer of text
// that is already in HBM is indexed and packed by five small launches; no byte goes back to the host:
//   newline_count_kernel   64 bytes per thread: SWAR "byte == '\n'" masks, popcount, one count per 16 KB block
//   (exclusive scan of the block counts: scan_* below)
//   newline_write_kernel   the same masks again, a block-wide scan of the per-thread counts, positions written in order
//   record_kernel          one thread per record: its four lines, rstrip()ped; line table (start, length); the first record
//                          with an empty line (end of file for the reference) and the first whose quality line is not as long
//                          as its sequence line (rejected, as aqc_fastq_parse does) by atomicMin
//   (exclusive scan of the sequence lengths -> the batch's offsets column)
//   gather_kernel          one warp per record copies bases and qualities into the packed columns
// All of it is HBM-bound byte work: ~2 reads of the text, one write of the columns.  Positions are 32-bit: a buffer holds
// less than 4 GiB (callers split; the columns of an aqc_batch have the same limit).
#pragma once
#include "aqc_device.cuh"

namespace aqc {

constexpr int PARSE_BLOCK_THREADS = 256;
constexpr uint32_t PARSE_THREAD_BYTES = 64u;       // four 16-byte loads per thread and block step: fewer block-wide scans per byte
constexpr uint32_t PARSE_BLOCK_BYTES = PARSE_BLOCK_THREADS * PARSE_THREAD_BYTES;
constexpr uint32_t SCAN_BLOCK_ELEMS = 1024u;      // 256 threads x 4

struct ParseArgs {
    const uint8_t *text;        // 16-byte aligned, readable up to the next multiple of 16 beyond n
    uint32_t n;                 // bytes (including the '\n' the host appended after a last line without one)
    uint32_t n_blk;             // ceil(n / PARSE_BLOCK_BYTES)
    uint32_t *blk;              // [n_blk + 1]: newline count per block, scanned in place (exclusive; [n_blk] = total)
    uint32_t *nl_pos;           // [n_lines]: byte position of every '\n', ascending
    uint32_t n_rec;             // records to look at (complete groups of four lines, capped by the caller)
    uint32_t *line_start, *line_len;   // [4 * n_rec]: rstrip()ped lines
    uint32_t *rec_len;          // [n_rec + 1]: length of the sequence line; scanned in place into offsets
    uint32_t *flags;            // [0] first record with an empty line, [1] first record with len(seq) != len(qual); init 0xFFFFFFFF
    uint32_t n_keep;            // records that are gathered
    uint8_t *seq, *qual;        // packed columns
};

// bit 7 of byte j of the result: byte j of w equals '\n'
__device__ __forceinline__ uint32_t newline_bytes(uint32_t w) {
    const uint32_t x = w ^ 0x0A0A0A0Au;
    return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);
}

// the sixteen masks of the thread's 64 bytes, bytes at or beyond n cleared; returns the number of newlines
__device__ __forceinline__ uint32_t newline_masks(const ParseArgs &P, uint32_t byte0, uint32_t m[16]) {
    uint32_t c = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const uint32_t q0 = byte0 + 16u * (uint32_t)q;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (q0 < P.n) v = *reinterpret_cast<const uint4 *>(P.text + q0);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t mk = q0 < P.n ? newline_bytes(w[k]) : 0u;
            const uint32_t b = q0 + 4u * (uint32_t)k;
            if (b + 4u > P.n) {                                  // the word straddles the end: keep bytes b .. n-1
                const uint32_t valid = P.n > b ? P.n - b : 0u;   // 0..3
                mk &= valid ? (0xFFFFFFFFu >> (8u * (4u - valid))) : 0u;
            }
            m[4 * q + k] = mk;
            c += (uint32_t)__popc(mk);
        }
    }
    return c;
}

__global__ void __launch_bounds__(PARSE_BLOCK_THREADS) newline_count_kernel(const __grid_constant__ ParseArgs P) {
    __shared__ uint32_t wsum[PARSE_BLOCK_THREADS / 32];
    for (uint32_t b = blockIdx.x; b < P.n_blk; b += gridDim.x) {
        uint32_t m[16];
        uint32_t c = newline_masks(P, b * PARSE_BLOCK_BYTES + threadIdx.x * PARSE_THREAD_BYTES, m);
        c = __reduce_add_sync(FULL, c);
        if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0;
            for (int i = 0; i < PARSE_BLOCK_THREADS / 32; i++) t += wsum[i];
            P.blk[b] = t;
        }
        __syncthreads();
    }
}

// exclusive prefix of the calling thread's value over the block (256 threads); *total = the block's sum
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *wsum /* [8] shared */, uint32_t *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(FULL, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    uint32_t base = 0, t = 0;
    for (int i = 0; i < PARSE_BLOCK_THREADS / 32; i++) { if (i < warp) base += wsum[i]; t += wsum[i]; }
    __syncthreads();
    *total = t;
    return base + inc - v;
}

__global__ void __launch_bounds__(PARSE_BLOCK_THREADS) newline_write_kernel(const __grid_constant__ ParseArgs P) {
    __shared__ uint32_t wsum[PARSE_BLOCK_THREADS / 32];
    for (uint32_t b = blockIdx.x; b < P.n_blk; b += gridDim.x) {
        const uint32_t byte0 = b * PARSE_BLOCK_BYTES + threadIdx.x * PARSE_THREAD_BYTES;
        uint32_t m[16];
        const uint32_t c = newline_masks(P, byte0, m);
        uint32_t total;
        uint32_t at = P.blk[b] + block_exclusive_scan(c, wsum, &total);
#pragma unroll
        for (int k = 0; k < 16; k++) {
            uint32_t mk = m[k];
            while (mk) {
                const int bit = __ffs((int)mk) - 1;          // 7, 15, 23 or 31: byte bit / 8
                mk &= mk - 1u;
                P.nl_pos[at++] = byte0 + 4u * (uint32_t)k + (uint32_t)(bit >> 3);
            }
        }
    }
}

// ---- exclusive scan of a u32 array, in place, n + 1 outputs (out[n] = total): three launches ----
struct ScanArgs { uint32_t *data; uint32_t n; uint32_t *part; uint32_t n_part; };

__global__ void __launch_bounds__(PARSE_BLOCK_THREADS) scan_partial_kernel(const __grid_constant__ ScanArgs S) {
    __shared__ uint32_t wsum[PARSE_BLOCK_THREADS / 32];
    for (uint32_t b = blockIdx.x; b < S.n_part; b += gridDim.x) {
        uint32_t v = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t i = b * SCAN_BLOCK_ELEMS + threadIdx.x * 4u + (uint32_t)k;
            if (i < S.n) v += S.data[i];
        }
        v = __reduce_add_sync(FULL, v);
        if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0;
            for (int i = 0; i < PARSE_BLOCK_THREADS / 32; i++) t += wsum[i];
            S.part[b] = t;
        }
        __syncthreads();
    }
}

// one CTA: exclusive scan of part[0 .. n_part) in place; part[n_part] = total
__global__ void __launch_bounds__(PARSE_BLOCK_THREADS) scan_single_kernel(const __grid_constant__ ScanArgs S) {
    __shared__ uint32_t wsum[PARSE_BLOCK_THREADS / 32];
    const uint32_t per = (S.n_part + PARSE_BLOCK_THREADS - 1) / PARSE_BLOCK_THREADS;
    const uint32_t lo = min(S.n_part, threadIdx.x * per), hi = min(S.n_part, lo + per);
    uint32_t v = 0;
    for (uint32_t i = lo; i < hi; i++) v += S.part[i];
    uint32_t total;
    uint32_t run = block_exclusive_scan(v, wsum, &total);
    for (uint32_t i = lo; i < hi; i++) { const uint32_t x = S.part[i]; S.part[i] = run; run += x; }
    if (threadIdx.x == 0) S.part[S.n_part] = total;
}

__global__ void __launch_bounds__(PARSE_BLOCK_THREADS) scan_final_kernel(const __grid_constant__ ScanArgs S) {
    __shared__ uint32_t wsum[PARSE_BLOCK_THREADS / 32];
    for (uint32_t b = blockIdx.x; b < S.n_part; b += gridDim.x) {
        uint32_t x[4], v = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t i = b * SCAN_BLOCK_ELEMS + threadIdx.x * 4u + (uint32_t)k;
            x[k] = i < S.n ? S.data[i] : 0u;
            v += x[k];
        }
        uint32_t total;
        uint32_t run = S.part[b] + block_exclusive_scan(v, wsum, &total);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t i = b * SCAN_BLOCK_ELEMS + threadIdx.x * 4u + (uint32_t)k;
            if (i < S.n) S.data[i] = run;
            run += x[k];
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) S.data[S.n] = S.part[S.n_part];
}

// str.rstrip()'s whitespace (the bytes aqc_fastq_parse strips)
__device__ __forceinline__ bool parse_is_ws(uint32_t c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == 0x0Bu || c == 0x0Cu; }

__global__ void __launch_bounds__(PARSE_BLOCK_THREADS) record_kernel(const __grid_constant__ ParseArgs P) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < P.n_rec; r += gridDim.x * blockDim.x) {
        bool empty = false;
        uint32_t len1 = 0, len3 = 0;
#pragma unroll
        for (uint32_t k = 0; k < 4; k++) {
            const uint32_t line = 4u * r + k;
            const uint32_t s = line ? P.nl_pos[line - 1] + 1u : 0u;
            uint32_t e = P.nl_pos[line];
            while (e > s && parse_is_ws(P.text[e - 1])) e--;
            P.line_start[line] = s;
            P.line_len[line] = e - s;
            empty = empty || e == s;
            if (k == 1) len1 = e - s;
            if (k == 3) len3 = e - s;
        }
        P.rec_len[r] = len1;
        if (empty) atomicMin(&P.flags[0], r);
        else if (len1 != len3) atomicMin(&P.flags[1], r);
    }
}

// dst[0 .. len) = src[0 .. len) by a group of G lanes: aligned 4-byte stores, each fed by two aligned loads and a funnel shift
// (src may read up to 3 bytes past src + len: the text has 16 bytes of slack); the bytes before the first aligned word and after
// the last one singly
template <uint32_t G>
__device__ __forceinline__ void group_copy(uint8_t *dst, const uint8_t *src, uint32_t len, uint32_t gl) {
    const uint32_t head = min(len, (4u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 3u)) & 3u);
    if (gl < head) dst[gl] = src[gl];
    const uint32_t n_words = (len - head) >> 2;
    for (uint32_t w = gl; w < n_words; w += G) {
        const uint32_t b = head + 4u * w;
        const uintptr_t a = reinterpret_cast<uintptr_t>(src + b);
        const uint32_t *ap = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
        const uint32_t sh = 8u * (uint32_t)(a & 3u);
        const uint32_t lo = ap[0], hi = sh ? ap[1] : 0u;
        *reinterpret_cast<uint32_t *>(dst + b) = __funnelshift_r(lo, hi, sh);
    }
    const uint32_t done = head + 4u * n_words;
    if (gl < len - done) dst[done + gl] = src[done + gl];
}

// eight lanes per record, four records in flight per warp (the kernel waits on memory: two dependent loads, then the copies)
__global__ void __launch_bounds__(PARSE_BLOCK_THREADS) gather_kernel(const __grid_constant__ ParseArgs P) {
    const uint32_t gl = threadIdx.x & 7u;
    const uint32_t W = gridDim.x * (blockDim.x >> 3);
    for (uint32_t r = blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3); r < P.n_keep; r += W) {
        const uint32_t s1 = P.line_start[4u * r + 1u], s3 = P.line_start[4u * r + 3u];
        const uint32_t dst = P.rec_len[r], len = P.rec_len[r + 1u] - dst;
        group_copy<8>(P.seq + dst, P.text + s1, len, gl);
        group_copy<8>(P.qual + dst, P.text + s3, len, gl);
    }
}

}  // namespace aqc
