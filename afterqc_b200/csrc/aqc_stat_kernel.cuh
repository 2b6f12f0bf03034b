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
//     "k-mer already stamped by an earlier, lower-ordered launch" (stamp_bits_kernel).  A k-mer WITHOUT that bit starts
//     with bit 15 of its 16-bit counter set: the atomic that counts a sighting returns the old value anyway, so "flagged"
//     costs one more instruction per k-mer (it shares the test for a full half), and only flagged k-mers take the global
//     path (load, rare atomicMin).  The head's stamps come from stamp_head_kernel (one thread per k-mer of the first few
//     thousand records, atomicMin only), so this kernel runs once over the whole range, nearly every k-mer is unflagged,
//     and the stamp table in HBM is hardly touched;
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
#define AQC_STAT_WARPS 20
#endif
constexpr int STAT_WARPS = AQC_STAT_WARPS;
constexpr uint32_t KTAB_SPILL = 0x4000u;
constexpr uint32_t KTAB_FLAG = 0x8000u;       // bit 15 of a 16-bit counter: the k-mer has no stamp below this launch's stamps yet

struct SKArgs {
    KArgs k;
    const uint32_t *kbits[2];      // per mate: bit idx = dense k-mer idx holds a stamp below every stamp this launch can produce
    const uint32_t *kbits_set;     // [2]: number of set bits per mate (all 4^k set: the kernel skips the bitmap test)
    uint32_t lo, hi;               // records [lo, hi) of the batch are walked by this launch
    uint32_t head_hi;              // stamp_head_kernel: records [lo, head_hi)
};

constexpr int STAT_QUEUE = 64;     // entries of a warp's queue of k-mers for the side table (drained at >= 32; a lane that finds it full works its k-mer off itself)
// Per-cycle tables: four copies, one per eight lanes.  A copy is cnt [5][MAXB] | qsum [5][MAXB] | disc [MAXB] (MAXB = 32 * NW, a
// multiple of 32: the class does not move the bank).  Reads of one length L put the lanes of a warp at few different cycles
// (the lane's 16-byte phase is lane * L mod 16: 8 values for L = 150, 4 for L = 100) and lanes 8 apart at the SAME cycle; an
// ATOMS.ADD on one address is served one lane at a time (only the +1 form, ATOMS.POPC.INC, merges), measured 3.8 wavefronts
// per instruction with one table (profiles/r02_v5_stat_kernel_*).  With the copies 1 and 16 words off the 32-word grid the
// 32 lanes of the common layouts fall into 32 different banks.
constexpr int STAT_ROWS = 2 * QC_CLASSES + 1;
__host__ __device__ __forceinline__ uint32_t stat_copy_words(int maxb) { return (uint32_t)(STAT_ROWS * maxb); }
__host__ __device__ __forceinline__ uint32_t stat_copy_off(int copy, int maxb) {
    return (uint32_t)(copy & 1) * (stat_copy_words(maxb) + 1u) + (uint32_t)(copy >> 1) * (2u * stat_copy_words(maxb) + 48u);
}
__host__ __device__ __forceinline__ uint32_t stat_tables_words(int maxb) { return stat_copy_off(3, maxb) + stat_copy_words(maxb); }

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
//   ktab [stat_ktab_words] u32 | per-cycle tables [stat_tables_words] | gch [MAXB + 1] | queue fill [32] | pad to 8
//   | queue keys [nwarps][STAT_QUEUE] u64 | queue stamps [nwarps][STAT_QUEUE] u64
__host__ __device__ __forceinline__ size_t stat_hist_words(int K, int nw) {
    const size_t maxb = 32 * (size_t)nw;
    const size_t w = (size_t)stat_ktab_words(K) + stat_tables_words((int)maxb) + maxb + 1 + 32;
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
    constexpr int CS = MAXB;                                 // class stride: a multiple of 32 (see STAT_ROWS)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = blockDim.x >> 5;
    const int K = A.p.qc_kmer;
    const uint32_t TW = stat_ktab_words(K);

    uint32_t *ktab = reinterpret_cast<uint32_t *>(smem_raw);
    uint32_t *s_tab = ktab + TW;
    uint32_t *s_cnt = s_tab + stat_copy_off(lane >> 3, MAXB);            // this lane's copy: cnt, then qsum, then disc
    uint32_t *s_disc = s_cnt + 2 * QC_CLASSES * CS;
    uint32_t *s_gch = s_tab + stat_tables_words(MAXB);
    uint32_t *qfill = s_gch + MAXB + 1 + warp;           // entries the warp's lanes have asked for in its side-table queue
    unsigned long long *qkey = reinterpret_cast<unsigned long long *>(ktab + stat_hist_words(K, NW)) + (size_t)warp * STAT_QUEUE;
    unsigned long long *qwhen = reinterpret_cast<unsigned long long *>(ktab + stat_hist_words(K, NW)) + (size_t)nwarps * STAT_QUEUE + (size_t)warp * STAT_QUEUE;

    // which mate this CTA works on: both mates wanted -> even CTAs mate 1, odd CTAs mate 2 (the host launches an even grid)
    const bool both = PAIRED && A.qc[0].valid && A.qc[1].valid;
    const int mate = both ? (int)(blockIdx.x & 1u) : ((PAIRED && !A.qc[0].valid) ? 1 : 0);
    const uint32_t cta = both ? blockIdx.x >> 1 : blockIdx.x, nctas = both ? gridDim.x >> 1 : gridDim.x;
    const QcDev &qd = A.qc[mate];
    if (!qd.valid) return;                                   // uniform for the CTA: nothing to do

    // every dense k-mer already stamped below this launch's stamps (the usual case after stamp_head_kernel): no flags
    const bool all_stamped = S.kbits_set != nullptr && S.kbits_set[mate] == (1u << (2 * K));
    {
        const uint32_t *gb = S.kbits[mate];
        const uint32_t n_dense = 1u << (2 * K);
        for (uint32_t i = tid; i < TW; i += blockDim.x) {
            uint32_t f = 0;
            if (!all_stamped) {
#pragma unroll
                for (uint32_t hf = 0; hf < 2; hf++) {
                    const uint32_t idx = 2u * i + hf;
                    const bool stamped = gb != nullptr && idx < n_dense && ((gb[idx >> 5] >> (idx & 31u)) & 1u);
                    if (!stamped) f |= KTAB_FLAG << (16u * hf);
                }
            }
            ktab[i] = f;
        }
    }
    for (int i = tid; i < (int)stat_tables_words(MAXB) + MAXB + 1 + 32; i += blockDim.x) s_tab[i] = 0;
    __syncthreads();

    const uint8_t *seq = mate ? A.seq2 : A.seq1, *qual = mate ? A.qual2 : A.qual1;
    const uint32_t *off = mate ? A.off2 : A.off1;
    const uint32_t kshift = (uint32_t)(K - 1);
    const uint32_t kmask = (1u << K) - 1u;
    const unsigned long long keymask = K >= 8 ? ~0ULL : ((1ULL << (8 * K)) - 1ULL);
    unsigned long long n_kmers = 0, n_reads = 0;             // per lane, added at the end
    unsigned long long *const kfirst = qd.kfirst, *const kcnt = qd.kcnt;

    // side-table k-mers are queued by whichever lane meets one (a slot from an atomic counter: the lanes need not be converged)
    // and worked off 32 at a time by the whole warp
    auto push_side = [&](unsigned long long key, unsigned long long when) {
        const uint32_t slot = atomicAdd(qfill, 1u);
        if (slot < (uint32_t)STAT_QUEUE) { qkey[slot] = key; qwhen[slot] = when; }
        else side_kmer(qd, key, when, K, A.error_flag);
    };
    auto drain = [&]() {                                     // whole warp
        __syncwarp();
        const uint32_t n = min(*reinterpret_cast<volatile uint32_t *>(qfill), (uint32_t)STAT_QUEUE);
        for (uint32_t b0 = 0; b0 < n; b0 += 32u)
            if (b0 + (uint32_t)lane < n) side_kmer(qd, qkey[b0 + lane], qwhen[b0 + lane], K, A.error_flag);
        __syncwarp();
        if (lane == 0) *reinterpret_cast<volatile uint32_t *>(qfill) = 0u;
        __syncwarp();
    };
    auto maybe_drain = [&]() {
        __syncwarp();
        const uint32_t fill = *reinterpret_cast<volatile uint32_t *>(qfill);
        __syncwarp();                                        // every lane has read it before any lane moves on and queues again
        if (fill >= 32u) drain();
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

        // The read is walked in three stretches.  HEAD: cycles [0, h) byte by byte (the windows that reach before the read,
        // the first discontinuity windows); WORDS: the aligned 4-byte words that lie wholly inside cycles [5, len - 1), four
        // bases per step from 16-byte loads; TAIL: the last 1..4 cycles byte by byte (quirk Q11 lives there).  Every lane
        // has its own 16-byte phase (lead), so the lanes of a warp sit at different cycles and their shared-memory atomics
        // spread over the banks.
        const uintptr_t sa = reinterpret_cast<uintptr_t>(seq + R.a), qa = reinterpret_cast<uintptr_t>(qual + R.a);
        const int lead = (int)(sa & 15);
        const uint4 *sp = reinterpret_cast<const uint4 *>(sa & ~(uintptr_t)15);
        const uint4 *qp = reinterpret_cast<const uint4 *>(qa & ~(uintptr_t)15);
        const bool same_lead = __all_sync(FULL, len == 0 || lead == (int)(qa & 15));      // columns equally aligned (always, in practice)
        int Wlo = (lead + 8) >> 2;                            // first word with cycle(byte 0) >= 5
        int Whi = len >= 5 ? (len - 5 + lead) >> 2 : -1;      // last word with cycle(byte 3) < len - 1
        if (!same_lead || len == 0) Whi = -1;                 // rare layout: the whole read byte by byte
        const bool inner = Wlo <= Whi;
        if (!inner) { Wlo = 1 << 20; Whi = -1; }
        const int h = inner ? 4 * Wlo - lead : len;           // cycles [0, h) are the head
        const int tl = inner ? 4 * (Whi + 1) - lead : len;    // cycles [tl, len) are the tail
        const int hmax = (int)__reduce_max_sync(FULL, (unsigned)h);
        const int maxch = (int)__reduce_max_sync(FULL, (unsigned)(inner ? ((Whi + 1) >> 2) + 1 : 0));
        const int nk = len - K;                               // k-mers start at i < len - K (quirk Q11)
        const unsigned long long when0 = R.order << 11;

        // rolling state
        uint32_t P0 = 0, P1 = 0;                              // k-mer code planes of the last 32 bases, newest at bit 31 (codes A0 C1 G2 T3)
        uint32_t B = 0xFFFFFFFFu;                             // same, "not A,C,G,T" (what lies before the read counts as such)
        uint32_t prevb = 0, Hs = 0, d4 = 0;                   // previous base; bit 7-j = "base p-j differs from base p-j-1"; window of cycle 2
        uint32_t r0 = 0, r1 = 0, r2 = 0;                      // the last twelve raw bytes, newest in the low byte of r0
        int gc = 0;

        // What the old value of a counter says beyond the count: y = (old half) ^ 0x3FFF is 0 or 0x8000 when this increment
        // filled the half to 0x4000 (this lane, and only it, moves 0x4000 counts to the global table), and has bit 15 set when
        // the k-mer is flagged (its first-seen stamp may still be ours to give: load, rare atomicMin).  Both are rare.
        auto kmer_rare = [&](uint32_t idx, int i, uint32_t y) {
            if ((y & 0x7FFFu) == 0u) {
                atomicSub(&ktab[idx >> 1], KTAB_SPILL << ((idx & 1u) << 4));
                atomicAdd(&kcnt[idx], (unsigned long long)KTAB_SPILL);
            }
            if (y & KTAB_FLAG) {                              // only the DIRECT first sighting is tracked on the device (see stat_read)
                const unsigned long long when = when0 | ((unsigned long long)i << 1);
                if (__ldcg(&kfirst[idx]) > when) atomicMin(&kfirst[idx], when);
            }
        };
        auto kmer_hit = [&](uint32_t idx, int i) {            // the dense k-mer idx starts at cycle i
            const uint32_t sh = (idx & 1u) << 4;
            const uint32_t old = atomicAdd(&ktab[idx >> 1], 1u << sh);
            const uint32_t y = ((old >> sh) ^ (KTAB_SPILL - 1u)) & 0xFFFFu;
            if (__builtin_expect(y - 1u >= 0x7FFFu, 0)) kmer_rare(idx, i, y);
        };
        auto step_byte = [&](uint32_t b, uint32_t q, int pos) {
            // A 0x41, C 0x43, T 0x54, G 0x47: bits 1..2 are a 2-bit code (A0 C1 T2 G3); the byte is one of the four iff it
            // re-encodes to itself (selector nibbles 1..3 pick zero bytes)
            const uint32_t code = (b >> 1) & 3u;
            const bool acgt = __byte_perm(0x47544341u, 0u, code | 0x4440u) == b;
            // ALL_BASES order (qualitycontrol.py:24): A0 T1 C2 G3, anything else 4
            const uint32_t cls = acgt ? __byte_perm(0x03010200u, 0u, code | 0x4440u) : 4u;
            uint32_t *cell = s_cnt + cls * CS + pos;
            atomicAdd(cell, 1u);
            atomicAdd(cell + QC_CLASSES * CS, q);
            gc += (int)(acgt ? (code & 1u) : 0u);             // C and G have code bit 0 set (:93-94)
            // discontinuity: the window of cycle pos-2 is complete now; the clamped windows at the ends reuse the first and
            // the last complete window (added after the walk)
            Hs = (Hs >> 1) | ((pos > 0 && b != prevb) ? 0x80u : 0u);
            prevb = b;
            const uint32_t d = (uint32_t)__popc(Hs >> 4);
            if (pos == 4) d4 = d;
            if (pos >= 4) atomicAdd(&s_disc[pos - 2], d);
            P0 = (P0 >> 1) | (((code ^ (code >> 1)) & 1u) << 31);
            P1 = (P1 >> 1) | ((code >> 1) << 31);
            B = (B >> 1) | (acgt ? 0u : 0x80000000u);
            r2 = (r2 << 8) | (r1 >> 24); r1 = (r1 << 8) | (r0 >> 24); r0 = (r0 << 8) | b;
            const int i = pos - (int)kshift;                  // first base of the k-mer that ends here
            if (i >= 0 && i < nk) {                           // a k-mer of the read (quirk Q11: not the one ending at the last base)
                if ((B >> (32 - K)) == 0u) {
                    const uint32_t idx = ((P1 >> (32 - K)) << K) | (P0 >> (32 - K));
                    kmer_hit(idx, i);
                } else {
                    push_side((((unsigned long long)r1 << 32) | r0) & keymask, when0 | ((unsigned long long)i << 1));
                }
            }
        };
        auto edited = [&](uint32_t &b, uint32_t &q, int pos) {
            if constexpr (POST) {
                if (any_edit) {
#pragma unroll
                    for (int k = 0; k < 3; k++) if (R.ep[k] == pos) { if (R.eb[k]) b = R.eb[k]; q = R.eq[k]; }
                }
            }
        };

        // ---- head ----
        // The first two 16-byte pieces hold the head (byte lead + h <= 20) and are the word loop's first pieces anyway; the
        // head's <= 8 bytes are moved to the bottom of two registers per column (word select by lead / 4, funnel shift by
        // lead % 4) -- no byte loads.  Only a read without inner words (shorter than ~13 bases, or the rare layout with
        // differently aligned columns) is read byte by byte from memory, whole.
        uint4 cb = make_uint4(0, 0, 0, 0), cq = make_uint4(0, 0, 0, 0), nb = make_uint4(0, 0, 0, 0), nq = make_uint4(0, 0, 0, 0);
        if (inner) {
            cb = ldg_stream16(sp); cq = ldg_stream16(qp);
            if (lead + len > 16) { nb = ldg_stream16(sp + 1); nq = ldg_stream16(qp + 1); }
        }
        uint32_t hb0, hb1, hq0, hq1;
        {
            const bool s1 = (lead & 4) != 0, s2 = (lead & 8) != 0;
            const uint32_t fs = 8u * (uint32_t)(lead & 3);
            auto sel4 = [&](uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3) { return s2 ? (s1 ? a3 : a2) : (s1 ? a1 : a0); };
            const uint32_t x0 = sel4(cb.x, cb.y, cb.z, cb.w), x1 = sel4(cb.y, cb.z, cb.w, nb.x), x2 = sel4(cb.z, cb.w, nb.x, nb.y);
            const uint32_t y0 = sel4(cq.x, cq.y, cq.z, cq.w), y1 = sel4(cq.y, cq.z, cq.w, nq.x), y2 = sel4(cq.z, cq.w, nq.x, nq.y);
            hb0 = __funnelshift_r(x0, x1, fs); hb1 = __funnelshift_r(x1, x2, fs);
            hq0 = __funnelshift_r(y0, y1, fs); hq1 = __funnelshift_r(y1, y2, fs);
        }
#pragma unroll 1
        for (int p0 = 0; p0 < hmax; p0++) {
            if (p0 < h) {
                const uint32_t fs = 8u * (uint32_t)(p0 & 3);
                uint32_t b = ((p0 & 4) ? hb1 : hb0) >> fs & 0xFFu, q = ((p0 & 4) ? hq1 : hq0) >> fs & 0xFFu;
                if (__builtin_expect(!inner, 0)) { b = seq[R.a + p0]; q = qual[R.a + p0]; }
                edited(b, q, p0);
                step_byte(b, q, p0);
            }
        }
        maybe_drain();

        // ---- words (the word after the last inner one is the tail: kept for the byte steps below) ----
        uint32_t tv = 0, tq = 0;
#pragma unroll 1
        for (int c = 0; c < maxch; c++) {
            uint4 nb2 = nb, nq2 = nq;                                 // piece 1 is here already
            if (c > 0) {
                nb2 = make_uint4(0, 0, 0, 0); nq2 = make_uint4(0, 0, 0, 0);
                if (c + 1 <= ((Whi + 1) >> 2)) { nb2 = ldg_stream16(sp + c + 1); nq2 = ldg_stream16(qp + c + 1); }   // on its way while this piece is walked
            }
            uint32_t bw[4] = {cb.x, cb.y, cb.z, cb.w}, qw[4] = {cq.x, cq.y, cq.z, cq.w};
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
            for (int w = 0; w < 4; w++) {                             // rolled: one copy of the per-word code
                const uint32_t v = bw[0], qv = qw[0];
                bw[0] = bw[1]; bw[1] = bw[2]; bw[2] = bw[3];
                qw[0] = qw[1]; qw[1] = qw[2]; qw[2] = qw[3];
                const int W = 4 * c + w;
                if (W == Whi + 1) { tv = v; tq = qv; }
                if (W >= Wlo && W <= Whi) {
                    const int pos0 = 4 * W - lead;                    // the lane's cycle of byte 0; 5 <= pos0, pos0 + 3 < len - 1
                    // the four 2-bit codes -> selector nibbles -> "re-encodes to itself" and the class of each byte
                    const uint32_t tt = (v >> 1) & 0x03030303u;
                    const uint32_t s16 = prmt_raw(tt + (tt >> 4), 0u, 0x4420u);
                    const uint32_t bad = prmt_raw(0x47544341u, 0u, s16) ^ v;
                    const uint32_t nzb = (((bad & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | bad) & 0x80808080u;   // bit 7 of a byte: not A,C,G,T
                    const uint32_t nzm = nzb >> 7;
                    const uint32_t clsw = (prmt_raw(0x03010200u, 0u, s16) & ~(nzm * 3u)) | (nzm << 2);   // 4: anything else
                    gc += __popc(tt & 0x01010101u & ~nzm);
                    const uint32_t x = v ^ ((v << 8) | prevb);        // byte t != 0: base t differs from the one before
                    prevb = v >> 24;
                    const uint32_t nzx = (((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;
                    // bits 0, 8, 16, 24 -> bits 24..27 of the product (all sixteen partial products fall on different bits)
                    Hs = (Hs >> 4) | ((((nzx >> 7) * 0x01020408u) >> 24) << 4);
                    const uint32_t u = tt ^ ((tt >> 1) & 0x01010101u);                                // k-mer codes A0 C1 G2 T3
                    P0 = __funnelshift_r(P0, ((u & 0x01010101u) * 0x01020408u) >> 24, 4);
                    P1 = __funnelshift_r(P1, (((u >> 1) & 0x01010101u) * 0x01020408u) >> 24, 4);
                    B = __funnelshift_r(B, (nzm * 0x01020408u) >> 24, 4);
                    r2 = r1; r1 = r0; r0 = prmt_raw(v, 0u, 0x0123u);
                    uint32_t *cell0 = s_cnt + pos0;
                    uint32_t *disc0 = s_disc + pos0 - 2;
#pragma unroll
                    for (int tt4 = 0; tt4 < 4; tt4++) {
                        uint32_t *cell = cell0 + ((clsw >> (8 * tt4)) & 0xFFu) * CS + tt4;
                        atomicAdd(cell, 1u);
                        atomicAdd(cell + QC_CLASSES * CS, (qv >> (8 * tt4)) & 0xFFu);
                        atomicAdd(disc0 + tt4, (uint32_t)__popc(Hs & (0xFu << (tt4 + 1))));
                    }
                    // the k-mer that ends at base t: bits [29 + t - K, 28 + t] of the planes
                    const uint32_t ks = 29u - (uint32_t)K;
                    if (__builtin_expect((B >> ks) == 0u, 1)) {       // every byte of the last K + 3 is one of A,C,G,T: four dense k-mers
                        uint32_t idx[4], y[4];
#pragma unroll
                        for (int tt4 = 0; tt4 < 4; tt4++) {
                            idx[tt4] = (((P1 >> (ks + tt4)) & kmask) << K) | ((P0 >> (ks + tt4)) & kmask);
                            const uint32_t sh = (idx[tt4] & 1u) << 4;
                            const uint32_t old = atomicAdd(&ktab[idx[tt4] >> 1], 1u << sh);
                            y[tt4] = ((old >> sh) ^ (KTAB_SPILL - 1u)) & 0xFFFFu;
                        }
                        if (__builtin_expect(max(max(y[0] - 1u, y[1] - 1u), max(y[2] - 1u, y[3] - 1u)) >= 0x7FFFu, 0)) {
#pragma unroll
                            for (int tt4 = 0; tt4 < 4; tt4++)
                                if (y[tt4] - 1u >= 0x7FFFu) kmer_rare(idx[tt4], pos0 + tt4 - (int)kshift, y[tt4]);
                        }
                    } else {
#pragma unroll 1
                        for (int tt4 = 0; tt4 < 4; tt4++) {
                            const int i = pos0 + tt4 - (int)kshift;
                            if (i < 0) continue;
                            if (((B >> (ks + tt4)) & kmask) == 0u) {
                                const uint32_t idx = (((P1 >> (ks + tt4)) & kmask) << K) | ((P0 >> (ks + tt4)) & kmask);
                                kmer_hit(idx, i);
                            } else {                                  // its raw bytes end 3 - t bytes above the newest
                                const uint32_t sh = 8u * (uint32_t)(3 - tt4);
                                const uint32_t lo = __funnelshift_r(r0, r1, sh), hi = __funnelshift_r(r1, r2, sh);
                                push_side((((unsigned long long)hi << 32) | lo) & keymask, when0 | ((unsigned long long)i << 1));
                            }
                        }
                    }
                }
            }
            cb = nb2; cq = nq2;
            maybe_drain();
        }

        // ---- tail ----
#pragma unroll 1
        for (int j = 0; j < 4; j++) {
            const int p0 = tl + j;
            if (p0 < len) {
                uint32_t b = (tv >> (8 * j)) & 0xFFu, q = (tq >> (8 * j)) & 0xFFu;
                edited(b, q, p0);
                step_byte(b, q, p0);
            }
        }
        maybe_drain();
        const uint32_t hist = Hs >> 4;
        if (want) {
            // the clamped windows (:97-104): cycles 0,1 share the window of cycle 2, the last two cycles that of cycle len-3
            if (d4) { atomicAdd(&s_disc[0], d4); atomicAdd(&s_disc[1], d4); }      // (eight lanes per copy meet here)
            const uint32_t dl = (uint32_t)__popc(hist);
            if (dl) { atomicAdd(&s_disc[len - 2], dl); atomicAdd(&s_disc[len - 1], dl); }
            atomicAdd(&s_gch[gc], 1u);                        // :112
            if (nk > 0) n_kmers += (unsigned long long)nk;    // totalKmer :114
            n_reads++;
        }
    }
    drain();

    // ---- epilogue: the CTA's histograms go to the QC object ----
    __syncthreads();
    for (uint32_t i = tid; i < TW; i += blockDim.x) {
        const uint32_t v = ktab[i];
        if (v & 0x7FFFu) atomicAdd(&kcnt[2 * i], (unsigned long long)(v & 0x7FFFu));
        if ((v >> 16) & 0x7FFFu) atomicAdd(&kcnt[2 * i + 1], (unsigned long long)((v >> 16) & 0x7FFFu));
    }
    for (int i = tid; i < QC_CLASSES * MAXB; i += blockDim.x) {
        const int c = i / MAXB, pos = i - c * MAXB;
        uint32_t n = 0, qs = 0;
#pragma unroll
        for (int cp = 0; cp < 4; cp++) {
            const uint32_t *t = s_tab + stat_copy_off(cp, MAXB) + c * CS + pos;
            n += t[0]; qs += t[QC_CLASSES * CS];
        }
        if (n) atomicAdd(&qd.cls_cnt[c * AQC_MAX_LEN + pos], (unsigned long long)n);
        if (qs) atomicAdd(&qd.cls_qsum[c * AQC_MAX_LEN + pos], (unsigned long long)qs);
    }
    for (int i = tid; i < MAXB; i += blockDim.x) {
        uint32_t v = 0;
#pragma unroll
        for (int cp = 0; cp < 4; cp++) v += s_tab[stat_copy_off(cp, MAXB) + 2 * QC_CLASSES * CS + i];
        if (v) atomicAdd(&qd.disc[i], (unsigned long long)v);
    }
    for (int i = tid; i <= MAXB; i += blockDim.x) { const uint32_t v = s_gch[i]; if (v) atomicAdd(&qd.gchist[i], (unsigned long long)v); }
    if (n_kmers) atomicAdd(&qd.scal[0], n_kmers);
    if (n_reads) atomicAdd(&qd.scal[1], n_reads);
}

}  // namespace aqc
