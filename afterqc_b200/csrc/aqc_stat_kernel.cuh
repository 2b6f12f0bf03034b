// aqc_stat_kernel.cuh -- QualityControl.statRead (qualitycontrol.py:73-122) for batches whose reads are <= 256 bases.
//
// One LANE per read, one CTA of 24 warps per SM, and EVERY histogram of statRead lives in that CTA's shared memory until
// the kernel ends.  A warp takes 32 consecutive records of one mate; each lane streams its own read from HBM in aligned
// 16-byte pieces (bases and qualities) and walks it byte by byte with a few registers of rolling state:
//   * per-cycle counts and quality sums (:76-96): two shared-memory atomics per base into [class][cycle] tables; lanes sit
//     at different cycles (every lane starts at its own 16-byte boundary) and the class stride is odd, so the 32 lanes of
//     an instruction spread over the banks;
//   * discontinuity (:97-108): a 4-bit history of "differs from the previous base"; cycle p-2 is settled when base p
//     arrives, the clamped windows at both ends reuse the first / last complete window;
//   * k-mers (:113-122): the dense index (two K-bit plane windows) and the number of consecutive A,C,G,T bases are rolling
//     registers; a complete k-mer is one shared-memory atomic on a packed 16-bit counter of the CTA's 4^k table (128 KB for
//     k = 8).  Measured on a B200 (tools/ubench_atomics.cu, profiles/r02_ubench_atomics.jsonl): a random increment costs
//     1/161 ns as a global reduction (REDG, bound by the L2 atomic units whatever the grid), 1/3200 ns as a shared-memory
//     atomic and 1/1400 ns as a shared-memory atomic whose old value is inspected.  A 16-bit half never overflows: the lane
//     whose increment takes a half from 0x3FFF to 0x4000 (exactly one lane sees that old value) moves 0x4000 counts to the
//     64-bit global table; a half would need another 49 151 increments before that lane's next instruction to be damaged;
//   * first-seen stamps (quirk Q12: ties of sortKmer resolve by insertion order): the engine hands the kernel a bitmap
//     "k-mer already stamped by an earlier, lower-ordered launch" (stamp_bits_kernel); it is copied to shared memory, a
//     set bit costs one shared load, and only k-mers without it take the global path (load, rare atomicMin).  The engine
//     splits a large launch into a short head and the rest, so that the rest runs with a nearly full bitmap;
//     The head's stamps come from stamp_head_kernel (one thread per k-mer of the first few thousand records, atomicMin
//     only), so this kernel runs once over the whole range and almost never touches the stamp table in HBM;
//   * k-mers with a byte outside A,C,G,T take stat_read's side-table path (the last eight raw bytes are a rolling register) --
//     not where they are met (one N in any of the 32 reads would drag the whole warp through ~500 instructions at that
//     cycle: 24 % of all instructions of the first lane-per-read form, profiles/r02_v2_stat_kernel_hot_lines.txt) but
//     from a per-warp queue in shared memory that is drained 32 entries at a time with every lane busy.
// The first form of this kernel gave a warp to one read (lane = cycle, ballots for the k-mer planes): 1050 warp-instructions
// per 150-base read, ~0.5 G reads/s (profiles/r02_v1_stat_kernel_ncu_full.txt).  Lane-per-read needs no cross-lane work at all.
// Two entry forms: the prefilter window of raw reads (aqc_stat_reads; statFile, qualitycontrol.py:331-357) and, POST,
// the sampled GOOD pairs of a filter launch taken from their 32-byte records (final slices + the correction walk's
// edits; preprocesser.py:624-627) -- the lane-per-pair filter kernel itself carries no statistics code.
#pragma once
#include "aqc_device.cuh"

namespace aqc {

// warps of the one CTA per SM: 24 (85 registers, no spills) measured 10-19 % faster than 32 (64 registers) and 28
#ifndef AQC_STAT_WARPS
#define AQC_STAT_WARPS 24
#endif
constexpr int STAT_WARPS = AQC_STAT_WARPS;
constexpr uint32_t KTAB_SPILL = 0x4000u;

struct SKArgs {
    KArgs k;
    const uint32_t *kbits[2];      // per mate: bit idx = dense k-mer idx holds a stamp below every stamp this launch can produce
    const uint32_t *kbits_set;     // [2]: number of set bits per mate (all 4^k set: the kernel skips the bitmap test)
    uint32_t lo, hi;               // records [lo, hi) of the batch are walked by this launch
    uint32_t head_hi;              // stamp_head_kernel: records [lo, head_hi)
};

constexpr int STAT_QUEUE = 64;     // entries of a warp's queue of k-mers for the side table (drained at > 32)

__host__ __device__ __forceinline__ uint32_t stat_ktab_words(int K) { const uint32_t n = 1u << (2 * K); return n >= 4u ? n >> 1 : 2u; }
__host__ __device__ __forceinline__ uint32_t stat_kbit_words(int K) { const uint32_t n = 1u << (2 * K); return n >= 32u ? n >> 5 : 1u; }

// bit idx of out = first[idx] < min_when, for both mates' dense tables (either may be null)
__global__ void stamp_bits_kernel(const unsigned long long *first0, const unsigned long long *first1, uint32_t n_dense,
                                  unsigned long long min_when, uint32_t *out0, uint32_t *out1, uint32_t *n_set /* [2], zeroed */) {
    const uint32_t n_pad = (n_dense + 31u) & ~31u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x) {
        const bool in = i < n_dense;
        const uint32_t b0 = __ballot_sync(FULL, in && first0 && first0[i] < min_when);
        const uint32_t b1 = __ballot_sync(FULL, in && first1 && first1[i] < min_when);
        if ((threadIdx.x & 31) == 0) {
            if (out0) { out0[i >> 5] = b0; if (b0) atomicAdd(&n_set[0], (uint32_t)__popc(b0)); }
            if (out1) { out1[i >> 5] = b1; if (b1) atomicAdd(&n_set[1], (uint32_t)__popc(b1)); }
        }
    }
}

// dynamic shared memory of one CTA (MAXB = 32*NW, class stride CS = MAXB + 1):
//   ktab [stat_ktab_words] u32 | kbits [stat_kbit_words] u32 | cnt [5][CS] | qsum [5][CS] | disc [MAXB] | gch [MAXB + 1] | scratch [32] | pad to 8
//   | queue keys [nwarps][STAT_QUEUE] u64 | queue stamps [nwarps][STAT_QUEUE] u64
__host__ __device__ __forceinline__ size_t stat_hist_words(int K, int nw) {
    const size_t maxb = 32 * (size_t)nw;
    const size_t w = (size_t)stat_ktab_words(K) + stat_kbit_words(K) + 2 * QC_CLASSES * (maxb + 1) + maxb + maxb + 1 + 32;
    return (w + 1) & ~(size_t)1;
}
__host__ __device__ __forceinline__ size_t stat_smem_bytes(int K, int nw, int nwarps) {
    return stat_hist_words(K, nw) * 4 + (size_t)nwarps * STAT_QUEUE * 16;
}

// COMP (util.py:27) and "byte outside COMP" without tables: the side-table path is rare, the bytes few
__device__ __forceinline__ uint32_t comp_byte(uint32_t b) {
    switch (b) {
        case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C';
        case 'a': return 't'; case 't': return 'a'; case 'c': return 'g'; case 'g': return 'c';
        case 'N': return 'N'; case '\n': return '\n';
        default: return 'N';
    }
}
__device__ __forceinline__ bool outside_comp(uint32_t b) {
    return !(b == 'A' || b == 'T' || b == 'C' || b == 'G' || b == 'a' || b == 't' || b == 'c' || b == 'g' || b == 'N' || b == '\n');
}

// A k-mer that holds a byte outside A,C,G,T: count + first direct sighting in the side table, and the seeding of its reverse
// complement when it holds a byte outside util.COMP (see the comment above stat_read, aqc_device.cuh).  key = the K raw bytes,
// first base most significant.  Out of line: every call site is a rare branch of the unrolled walk (instruction cache).
__device__ __noinline__ void side_kmer(const QcDev &qd, unsigned long long key, unsigned long long when, int K, int *error_flag) {
    unsigned long long rkey = 0;
    bool foreign = false;
    for (int j = 0; j < K; j++) {
        const uint32_t bj = (uint32_t)(key >> (8 * (K - 1 - j))) & 0xFFu;
        rkey |= (unsigned long long)comp_byte(bj) << (8 * j);
        foreign |= outside_comp(bj);
    }
    if (key == AQC_KMER_NEVER || rkey == AQC_KMER_NEVER) { atomicExch(error_flag, AQC_ERR_INVALID); return; }
    const int h = side_slot(qd, key);
    const int hr = side_slot(qd, rkey);
    if (h < 0 || hr < 0) { atomicExch(error_flag, AQC_ERR_KMER_TABLE_FULL); return; }
    atomicAdd(&qd.scnt[h], 1ULL);
    first_min(&qd.sfirst[h], when);
    if (foreign) first_min(&qd.sseed[hr], when | 1ULL);
}

// What a lane needs to know about its record: where the (final) read lies, whether it is stat'd, its k-mer order index and,
// POST, the bytes the correction walk rewrote in this mate (preprocesser.py:575-592; at most three: distance <= 3).
struct StatRead {
    uint32_t a;                    // offset of the read's first base in the column
    int len;                       // 0: not stat'd here
    uint64_t order;
    int ep[3];                     // position in the read (-1: none)
    uint32_t eb[3], eq[3];         // new base (0: unchanged), new quality
};

template <bool POST>
__device__ __forceinline__ bool stat_locate(const KArgs &A, const uint32_t *off, int mate, uint32_t p, StatRead &R) {
    const uint64_t gidx = A.first_index + p;
    R.a = off[p];
    R.len = (int)(off[p + 1] - R.a);
#pragma unroll
    for (int k = 0; k < 3; k++) { R.ep[k] = -1; R.eb[k] = 0; R.eq[k] = 0; }
    if constexpr (POST) {
        R.order = gidx;
        if (!(A.p.qc_sample <= 0 || gidx + 1 < (uint64_t)A.p.qc_sample)) return false;           // preprocesser.py:624 (quirk Q10)
        const uint4 *r = reinterpret_cast<const uint4 *>(A.results + p);
        const uint4 w0 = r[0];
        if ((w0.x & 0xFFu) != (uint32_t)AQC_GOOD) return false;
        const int n_edits = (int)((w0.x >> 8) & 0xFFu);
        const int start1 = (int)(w0.x >> 16), start2 = (int)(w0.y >> 16);
        R.a += (uint32_t)(mate ? start2 : start1);                                                 // the final slice: trim + adapter cut
        R.len = (int)((mate ? w0.z : w0.y) & 0xFFFFu);
        if (n_edits) {
            const uint4 w1 = r[1];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                if (k < n_edits) {
                    const uint32_t e = k == 0 ? w1.x : (k == 1 ? w1.y : w1.z);
                    const int kind = (int)AQC_EDIT_KIND(e);
                    if (kind == 0 && mate == 0) { R.ep[k] = (int)AQC_EDIT_POS(e) - start1; R.eb[k] = AQC_EDIT_BASE(e); R.eq[k] = AQC_EDIT_QUAL(e); }
                    else if (kind == 1 && mate == 1) { R.ep[k] = (int)AQC_EDIT_POS(e) - start2; R.eb[k] = AQC_EDIT_BASE(e); R.eq[k] = AQC_EDIT_QUAL(e); }
                    else if (kind == 2) { R.ep[k] = mate == 0 ? (int)AQC_EDIT_POS(e) - start1 : (int)AQC_EDIT_POS2(e) - start2; R.eq[k] = '!'; }
                }
            }
        }
        return true;
    } else {
        R.order = A.order_base + (gidx - A.stat_lo);
        return gidx >= A.stat_lo && gidx < A.stat_hi;
    }
}

// First-seen stamps of the DENSE k-mers of records [lo, head_hi): one warp per (record, mate), one lane per k-mer, atomicMin
// only.  Run before stat_kernel, it lets that kernel find (nearly) every k-mer already stamped below its own stamps.
template <bool PAIRED, bool POST>
__global__ void __launch_bounds__(256) stamp_head_kernel(const __grid_constant__ SKArgs S) {
    const KArgs &A = S.k;
    const int lane = threadIdx.x & 31;
    const int K = A.p.qc_kmer;
    const uint32_t nm = PAIRED ? 2u : 1u;
    const uint32_t units = (S.head_hi - S.lo) * nm;
    const uint32_t W = gridDim.x * (blockDim.x >> 5);
    for (uint32_t u = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); u < units; u += W) {
        const int mate = PAIRED ? (int)(u & 1u) : 0;
        const QcDev &qd = A.qc[mate];
        if (!qd.valid) continue;
        const uint32_t p = S.lo + (PAIRED ? (u >> 1) : u);
        StatRead R;
        if (!stat_locate<POST>(A, mate ? A.off2 : A.off1, mate, p, R)) continue;
        if (R.len < 5) continue;                              // empty / too short: stat_kernel reports it
        const uint8_t *s = (mate ? A.seq2 : A.seq1) + R.a;
        const int nk = R.len - K;
        for (int i = lane; i < nk; i += 32) {
            uint32_t w0 = 0, w1 = 0;
            bool dense = true;
            for (int j = 0; j < K; j++) {
                uint32_t b = s[i + j];
                if constexpr (POST) {
#pragma unroll
                    for (int k = 0; k < 3; k++) if (R.ep[k] == i + j && R.eb[k]) b = R.eb[k];
                }
                const uint32_t code = (b >> 1) & 3u;
                dense = dense && (((0x47544341u >> (8 * code)) & 0xFFu) == b);
                w0 |= ((code ^ (code >> 1)) & 1u) << j;       // k-mer codes A0 C1 G2 T3: bit 0 = C or T, bit 1 = G or T
                w1 |= (code >> 1) << j;
            }
            if (dense) {
                const uint32_t idx = (w1 << K) | w0;
                const unsigned long long when = (R.order << 11) | ((unsigned long long)i << 1);
                if (__ldcg(&qd.kfirst[idx]) > when) atomicMin(&qd.kfirst[idx], when);
            }
        }
    }
}

template <bool PAIRED, int NW, bool POST>
__global__ void __launch_bounds__(STAT_WARPS * 32, 1) stat_kernel(const __grid_constant__ SKArgs S) {
    const KArgs &A = S.k;
    AQC_DYN_SMEM(smem_raw);
    constexpr int MAXB = 32 * NW;
    constexpr int CS = MAXB + 1;                             // odd class stride: the five classes of a cycle sit in five banks
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = blockDim.x >> 5;
    const int K = A.p.qc_kmer;
    const uint32_t TW = stat_ktab_words(K), BW = stat_kbit_words(K);

    uint32_t *ktab = reinterpret_cast<uint32_t *>(smem_raw);
    uint32_t *kbits = ktab + TW;
    uint32_t *s_cnt = kbits + BW;
    uint32_t *s_qs = s_cnt + QC_CLASSES * CS;
    uint32_t *s_disc = s_qs + QC_CLASSES * CS;
    uint32_t *s_gch = s_disc + MAXB;
    uint32_t *trash = s_gch + MAXB + 1 + lane;           // the lane's scratch word: target of the atomics that do not apply
    unsigned long long *qkey = reinterpret_cast<unsigned long long *>(ktab + stat_hist_words(K, NW)) + (size_t)warp * STAT_QUEUE;
    unsigned long long *qwhen = reinterpret_cast<unsigned long long *>(ktab + stat_hist_words(K, NW)) + (size_t)nwarps * STAT_QUEUE + (size_t)warp * STAT_QUEUE;

    // which mate this CTA works on: both mates wanted -> even CTAs mate 1, odd CTAs mate 2 (the host launches an even grid)
    const bool both = PAIRED && A.qc[0].valid && A.qc[1].valid;
    const int mate = both ? (int)(blockIdx.x & 1u) : ((PAIRED && !A.qc[0].valid) ? 1 : 0);
    const uint32_t cta = both ? blockIdx.x >> 1 : blockIdx.x, nctas = both ? gridDim.x >> 1 : gridDim.x;
    const QcDev &qd = A.qc[mate];
    if (!qd.valid) return;                                   // uniform for the CTA: nothing to do

    for (uint32_t i = tid; i < TW; i += blockDim.x) ktab[i] = 0;
    {
        const uint32_t *gb = S.kbits[mate];
        for (uint32_t i = tid; i < BW; i += blockDim.x) kbits[i] = gb ? gb[i] : 0u;
    }
    for (int i = tid; i < 2 * QC_CLASSES * CS + 2 * MAXB + 1 + 32; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
    // every dense k-mer already stamped below this launch's stamps: no bitmap test per k-mer
    const bool all_stamped = S.kbits_set != nullptr && S.kbits_set[mate] == (1u << (2 * K));

    const uint8_t *seq = mate ? A.seq2 : A.seq1, *qual = mate ? A.qual2 : A.qual1;
    const uint32_t *off = mate ? A.off2 : A.off1;
    const uint32_t kshift = (uint32_t)(K - 1);
    const uint32_t kmask = (1u << K) - 1u;
    const unsigned long long keymask = K >= 8 ? ~0ULL : ((1ULL << (8 * K)) - 1ULL);
    unsigned long long n_kmers = 0, n_reads = 0;             // per lane, added at the end
    unsigned long long *const kfirst = qd.kfirst, *const kcnt = qd.kcnt;
    int qn = 0;                                              // entries in the warp's side-table queue (warp-uniform)

    auto drain = [&](int take) {                             // the last `take` (<= 32) queued k-mers, one per lane
        __syncwarp();
        if (lane < take) side_kmer(qd, qkey[qn - take + lane], qwhen[qn - take + lane], K, A.error_flag);
        qn -= take;
        __syncwarp();
    };

    const uint32_t tiles = (S.hi - S.lo + 31u) >> 5;
    const uint32_t stride = nctas * (uint32_t)nwarps;
#pragma unroll 1
    for (uint32_t t = cta * (uint32_t)nwarps + (uint32_t)warp; t < tiles; t += stride) {
        const uint32_t p = S.lo + t * 32u + (uint32_t)lane;
        StatRead R;
        R.a = 0; R.len = 0; R.order = 0;
#pragma unroll
        for (int k = 0; k < 3; k++) { R.ep[k] = -1; R.eb[k] = 0; R.eq[k] = 0; }
        bool want = p < S.hi && stat_locate<POST>(A, off, mate, p, R);
        if (want) {
            if (R.len <= 0) {        // an empty read runs no loop of statRead but is still counted: gcHistogram[0] += 1 (:112)
                atomicAdd(&s_gch[0], 1u);
                n_reads++;
                want = false;
            } else if (R.len < 5) { atomicExch(A.error_flag, AQC_ERR_TOO_SHORT_STAT); want = false; }  // reference: IndexError (:97-108)
            else if (R.len > MAXB) { atomicExch(A.error_flag, AQC_ERR_TOO_LONG); want = false; }      // the host picks NW; defensive
        }
        const int len = want ? R.len : 0;
        bool any_edit = false;
        if constexpr (POST) any_edit = __any_sync(FULL, want && (R.ep[0] >= 0 || R.ep[1] >= 0 || R.ep[2] >= 0));

        // the lane walks bytes [0, lead + len) from the 16-byte boundary below its read; cycle = byte index - lead
        const uintptr_t sa = reinterpret_cast<uintptr_t>(seq + R.a), qa = reinterpret_cast<uintptr_t>(qual + R.a);
        const int lead = (int)(sa & 15);
        const uint4 *sp = reinterpret_cast<const uint4 *>(sa & ~(uintptr_t)15);
        const uint4 *qp = reinterpret_cast<const uint4 *>(qa & ~(uintptr_t)15);
        const int nch = len > 0 ? (lead + len + 15) >> 4 : 0;
        const int maxch = (int)__reduce_max_sync(FULL, (unsigned)nch);
        const bool same_lead = __all_sync(FULL, len == 0 || lead == (int)(qa & 15));      // columns equally aligned (always, in practice)
        const int nk = len - K;                               // k-mers start at i < len - K (quirk Q11)
        const unsigned long long when0 = R.order << 11;

        // rolling state
        uint32_t kw = 0;                                      // plane windows of the last K bases: bits 0..K-1 plane 0, 16.. plane 1; bit K-1 = newest
        int vrun = 0;                                         // consecutive A,C,G,T bases ending here (saturates at K)
        uint32_t prevb = 0, hist = 0, d4 = 0;                 // previous base; bit j = "base p-j differs from base p-j-1"; window of cycle 2
        unsigned long long raw = 0;                           // the last eight raw bytes, newest in the low byte
        int gc = 0;

        uint4 cb = make_uint4(0, 0, 0, 0), cq = make_uint4(0, 0, 0, 0);
        if (0 < nch) { cb = ldg_stream16(sp); if (same_lead) cq = ldg_stream16(qp); }
#pragma unroll 1
        for (int c = 0; c < maxch; c++) {
            uint4 nb = make_uint4(0, 0, 0, 0), nq = make_uint4(0, 0, 0, 0);
            if (c + 1 < nch) { nb = ldg_stream16(sp + c + 1); if (same_lead) nq = ldg_stream16(qp + c + 1); }      // the next piece is on its way while this one is walked
            uint32_t bw[4] = {cb.x, cb.y, cb.z, cb.w}, qw[4] = {cq.x, cq.y, cq.z, cq.w};
            if (__builtin_expect(!same_lead, 0)) {            // rare layout: quality bytes one by one
                if (c < nch) {
#pragma unroll
                    for (int w = 0; w < 4; w++) {
                        uint32_t v = 0;
#pragma unroll
                        for (int tt = 0; tt < 4; tt++) {
                            const int pos = 16 * c + 4 * w + tt - lead;
                            if (pos >= 0 && pos < len) v |= (uint32_t)qual[R.a + pos] << (8 * tt);
                        }
                        qw[w] = v;
                    }
                }
            }
            if constexpr (POST) {
                if (any_edit) {
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        const int bi = R.ep[k] + lead - 16 * c;       // byte index inside this piece
                        if (R.ep[k] >= 0 && bi >= 0 && bi < 16) {
                            const uint32_t sh = 8u * (uint32_t)(bi & 3), m = 0xFFu << sh;
#pragma unroll
                            for (int w = 0; w < 4; w++) {
                                if ((bi >> 2) == w) {
                                    if (R.eb[k]) bw[w] = (bw[w] & ~m) | (R.eb[k] << sh);
                                    qw[w] = (qw[w] & ~m) | (R.eq[k] << sh);
                                }
                            }
                        }
                    }
                }
            }
#pragma unroll 1
            for (int w = 0; w < 4; w++) {                             // rolled: four copies of the per-base code, not sixteen
                const uint32_t wb_ = bw[0], wq_ = qw[0];
                bw[0] = bw[1]; bw[1] = bw[2]; bw[2] = bw[3];
                qw[0] = qw[1]; qw[1] = qw[2]; qw[2] = qw[3];
#pragma unroll
                for (int tt = 0; tt < 4; tt++) {
                    const int pos = 16 * c + 4 * w + tt - lead;       // the lane's cycle
                    bool side = false;
                    if ((unsigned)pos < (unsigned)len) {
                        // straight-line code: what does not apply to this cycle adds 0 to the lane's scratch word instead of branching
                        const uint32_t b = (wb_ >> (8 * tt)) & 0xFFu, q = (wq_ >> (8 * tt)) & 0xFFu;
                        // A 0x41, C 0x43, T 0x54, G 0x47: bits 1..2 are a 2-bit code (A0 C1 T2 G3); the byte is one of the four
                        // iff it re-encodes to itself (selector nibbles 1..3 pick zero bytes)
                        const uint32_t code = (b >> 1) & 3u;
                        const bool acgt = __byte_perm(0x47544341u, 0u, code | 0x4440u) == b;
                        // ALL_BASES order (qualitycontrol.py:24): A0 T1 C2 G3, anything else 4
                        const uint32_t cls = acgt ? __byte_perm(0x03010200u, 0u, code | 0x4440u) : 4u;
                        uint32_t *cell = s_cnt + cls * CS + pos;
                        atomicAdd(cell, 1u);
                        atomicAdd(cell + QC_CLASSES * CS, q);
                        gc += (int)(acgt ? (code & 1u) : 0u);         // C and G have code bit 0 set (:93-94)
                        // discontinuity: the window of cycle pos-2 is complete now; the clamped windows at the ends reuse the first
                        // and the last complete window (added after the walk)
                        hist = ((hist << 1) | ((pos > 0 && b != prevb) ? 1u : 0u)) & 0xFu;
                        prevb = b;
                        const uint32_t d = (uint32_t)__popc(hist);
                        d4 = pos == 4 ? d : d4;
                        atomicAdd(pos >= 4 ? &s_disc[pos - 2] : trash, d);
                        // k-mer ending here: internal dense index (plane 1 bits << K) | plane 0 bits, bit t = base t of the k-mer,
                        // k-mer codes A0 C1 G2 T3 (lut2 of the warp-per-read path): bit 0 = C or T, bit 1 = G or T
                        const uint32_t kk = ((code ^ (code >> 1)) & 1u) | ((code >> 1) << 16);
                        kw = ((kw >> 1) & 0x7FFF7FFFu) | (kk << kshift);
                        vrun = acgt ? min(vrun + 1, K) : 0;
                        raw = (raw << 8) | b;
                        const int i = pos - (int)kshift;               // its first base
                        const bool whole = i >= 0 && i < nk;           // a k-mer of the read ends here (quirk Q11: not at the last base)
                        const bool dense = whole && vrun == K;
                        side = whole && vrun != K;                     // a byte outside A,C,G,T in the k-mer: queued for the side table
                        const uint32_t idx = ((kw >> 16) << K) | (kw & kmask);
                        const uint32_t sh = (idx & 1u) << 4;
                        const uint32_t old = atomicAdd(dense ? &ktab[idx >> 1] : trash, dense ? (1u << sh) : 0u);
                        if (__builtin_expect(dense && ((old >> sh) & 0xFFFFu) == KTAB_SPILL - 1u, 0)) {
                            atomicSub(&ktab[idx >> 1], KTAB_SPILL << sh);
                            atomicAdd(&kcnt[idx], (unsigned long long)KTAB_SPILL);
                        }
                        // only the DIRECT first sighting is tracked on the device (see stat_read)
                        if (!all_stamped) {
                            if (__builtin_expect(dense && !((kbits[idx >> 5] >> (idx & 31u)) & 1u), 0)) {
                                const unsigned long long when = when0 | ((unsigned long long)i << 1);
                                if (__ldcg(&kfirst[idx]) > when) atomicMin(&kfirst[idx], when);
                            }
                        }
                    }
                    const uint32_t sm = __ballot_sync(FULL, side);
                    if (__builtin_expect(sm != 0u, 0)) {
                        if (side) {
                            const int slot = qn + __popc(sm & ((1u << lane) - 1u));
                            qkey[slot] = raw & keymask;
                            qwhen[slot] = when0 | ((unsigned long long)(pos - (int)kshift) << 1);
                        }
                        qn += __popc(sm);
                        if (qn > 32) drain(32);
                    }
                }
            }
            cb = nb; cq = nq;
        }
        if (want) {
            // the clamped windows (:97-104): cycles 0,1 share the window of cycle 2, the last two cycles that of cycle len-3
            if (d4) { atomicAdd(&s_disc[0], d4); atomicAdd(&s_disc[1], d4); }
            const uint32_t dl = (uint32_t)__popc(hist);
            if (dl) { atomicAdd(&s_disc[len - 2], dl); atomicAdd(&s_disc[len - 1], dl); }
            atomicAdd(&s_gch[gc], 1u);                        // :112
            if (nk > 0) n_kmers += (unsigned long long)nk;    // totalKmer :114
            n_reads++;
        }
    }
    if (qn > 0) drain(qn);

    // ---- epilogue: the CTA's histograms go to the QC object ----
    __syncthreads();
    for (uint32_t i = tid; i < TW; i += blockDim.x) {
        const uint32_t v = ktab[i];
        if (v & 0xFFFFu) atomicAdd(&kcnt[2 * i], (unsigned long long)(v & 0xFFFFu));
        if (v >> 16) atomicAdd(&kcnt[2 * i + 1], (unsigned long long)(v >> 16));
    }
    for (int i = tid; i < QC_CLASSES * MAXB; i += blockDim.x) {
        const int c = i / MAXB, pos = i - c * MAXB;
        const uint32_t n = s_cnt[c * CS + pos], qs = s_qs[c * CS + pos];
        if (n) atomicAdd(&qd.cls_cnt[c * AQC_MAX_LEN + pos], (unsigned long long)n);
        if (qs) atomicAdd(&qd.cls_qsum[c * AQC_MAX_LEN + pos], (unsigned long long)qs);
    }
    for (int i = tid; i < MAXB; i += blockDim.x) { const uint32_t v = s_disc[i]; if (v) atomicAdd(&qd.disc[i], (unsigned long long)v); }
    for (int i = tid; i <= MAXB; i += blockDim.x) { const uint32_t v = s_gch[i]; if (v) atomicAdd(&qd.gchist[i], (unsigned long long)v); }
    if (n_kmers) atomicAdd(&qd.scal[0], n_kmers);
    if (n_reads) atomicAdd(&qd.scal[1], n_reads);
}

}  // namespace aqc
