// aqc_stat_kernel.cuh -- QualityControl.statRead (qualitycontrol.py:73-122) for batches whose reads are <= 256 bases.
//
// One WARP per read, lane = cycle (position), one CTA of 32 warps per SM, and EVERY histogram of statRead lives in that
// CTA's shared memory until the kernel ends:
//   * the dense k-mer table (4^k counters, k <= 8) is 128 KB of packed 16-bit counters.  Measured on a B200
//     (tools/ubench_atomics.cu, profiles/r02_ubench_atomics.jsonl): a random increment costs 1/161 ns as a global
//     reduction (REDG, bound by the L2 atomic units whatever the grid), 1/3200 ns as a shared-memory atomic and 1/1400 ns
//     as a shared-memory atomic whose old value is inspected -- the k-mer loop, 142 increments per 150-base read, is what
//     bounded stat_read (aqc_device.cuh) at ~110 G k-mers/s.  A 16-bit half never overflows: the lane whose increment
//     takes a half from 0x3FFF to 0x4000 (exactly one lane sees that old value) moves 0x4000 counts to the 64-bit global
//     table; a half would have to collect another 49 151 increments before that lane's next instruction to be damaged.
//   * per-cycle counts and quality sums, discontinuity and the GC histogram: 32-bit shared-memory atomics, lane = cycle,
//     so a warp instruction touches 32 consecutive words (no bank conflicts, no per-read flush: 32 bits hold any launch).
//   * first-seen stamps (quirk Q12: ties of sortKmer resolve by insertion order): the engine hands the kernel a bitmap
//     "k-mer already stamped by an earlier, lower-ordered launch" (stamp_bits_kernel); it is copied to shared memory, a
//     set bit costs one shared load, and only k-mers without it take the global path (load, rare atomicMin).  The engine
//     splits a large launch into a short head and the rest, so that the rest runs with a nearly full bitmap.
//   * k-mers with a byte outside A,C,G,T take stat_read's side-table path unchanged.
// Two entry forms: the prefilter window of raw reads (aqc_stat_reads; statFile, qualitycontrol.py:331-357) and, POST,
// the sampled GOOD pairs of a filter launch taken from their 32-byte records (final slices + the correction walk's
// edits; preprocesser.py:624-627) -- the lane-per-pair filter kernel itself carries no statistics code.
#pragma once
#include "aqc_device.cuh"

namespace aqc {

constexpr int STAT_WARPS = 32;
constexpr uint32_t KTAB_SPILL = 0x4000u;

struct SKArgs {
    KArgs k;
    const uint32_t *kbits[2];      // per mate: bit idx = dense k-mer idx holds a stamp below every stamp this launch can produce
    uint32_t lo, hi;               // records [lo, hi) of the batch are walked by this launch
};

__host__ __device__ __forceinline__ uint32_t stat_ktab_words(int K) { const uint32_t n = 1u << (2 * K); return n >= 4u ? n >> 1 : 2u; }
__host__ __device__ __forceinline__ uint32_t stat_kbit_words(int K) { const uint32_t n = 1u << (2 * K); return n >= 32u ? n >> 5 : 1u; }

// bit idx of out = first[idx] < min_when, for both mates' dense tables (either may be null)
__global__ void stamp_bits_kernel(const unsigned long long *first0, const unsigned long long *first1, uint32_t n_dense,
                                  unsigned long long min_when, uint32_t *out0, uint32_t *out1) {
    const uint32_t n_pad = (n_dense + 31u) & ~31u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x) {
        const bool in = i < n_dense;
        const uint32_t b0 = __ballot_sync(FULL, in && first0 && first0[i] < min_when);
        const uint32_t b1 = __ballot_sync(FULL, in && first1 && first1[i] < min_when);
        if ((threadIdx.x & 31) == 0) { if (out0) out0[i >> 5] = b0; if (out1) out1[i >> 5] = b1; }
    }
}

// dynamic shared memory of one CTA (MAXB = 32*NW):
//   ktab [stat_ktab_words] u32 | kbits [stat_kbit_words] u32 | cnt [5][MAXB] | qsum [5][MAXB] | disc [MAXB] | gch [MAXB+1 .. pad 4]
//   | luts (768 B) | scratch [nwarps][MAXB] bytes (reads that hold a k-mer with a foreign byte)
__host__ __device__ __forceinline__ size_t stat_smem_bytes(int K, int nw, int nwarps) {
    const size_t maxb = 32 * (size_t)nw;
    return ((size_t)stat_ktab_words(K) + stat_kbit_words(K) + 2 * QC_CLASSES * maxb + maxb + maxb + 4) * 4 + 768 + (size_t)nwarps * maxb;
}

template <bool PAIRED, int NW, bool POST>
__global__ void __launch_bounds__(STAT_WARPS * 32, 1) stat_kernel(const __grid_constant__ SKArgs S) {
    const KArgs &A = S.k;
    AQC_DYN_SMEM(smem_raw);
    constexpr int MAXB = 32 * NW;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = blockDim.x >> 5;
    const int K = A.p.qc_kmer;
    const uint32_t TW = stat_ktab_words(K), BW = stat_kbit_words(K);

    uint32_t *ktab = reinterpret_cast<uint32_t *>(smem_raw);
    uint32_t *kbits = ktab + TW;
    uint32_t *s_cnt = kbits + BW;
    uint32_t *s_qs = s_cnt + QC_CLASSES * MAXB;
    uint32_t *s_disc = s_qs + QC_CLASSES * MAXB;
    uint32_t *s_gch = s_disc + MAXB;
    uint8_t *lutbase = reinterpret_cast<uint8_t *>(s_gch + MAXB + 4);
    const uint8_t *lut1 = lutbase, *lut2 = lutbase + 256, *lut3 = lutbase + 512;
    uint8_t *scratch = lutbase + 768 + (size_t)warp * MAXB;

    // which mate this CTA works on: both mates wanted -> even CTAs mate 1, odd CTAs mate 2 (the host launches an even grid)
    const bool both = PAIRED && A.qc[0].valid && A.qc[1].valid;
    const int mate = both ? (int)(blockIdx.x & 1u) : ((PAIRED && !A.qc[0].valid) ? 1 : 0);
    const uint32_t cta = both ? blockIdx.x >> 1 : blockIdx.x, nctas = both ? gridDim.x >> 1 : gridDim.x;
    const QcDev &qd = A.qc[mate];
    if (!qd.valid) return;                                   // uniform for the CTA: nothing to do (single-end launch of mate 2, ...)

    for (uint32_t i = tid; i < TW; i += blockDim.x) ktab[i] = 0;
    {
        const uint32_t *gb = S.kbits[mate];
        for (uint32_t i = tid; i < BW; i += blockDim.x) kbits[i] = gb ? gb[i] : 0u;
    }
    for (int i = tid; i < 2 * QC_CLASSES * MAXB + 2 * MAXB + 4; i += blockDim.x) s_cnt[i] = 0;
    for (int i = tid; i < 768; i += blockDim.x) lutbase[i] = reinterpret_cast<const uint8_t *>(A.luts)[i];
    __syncthreads();

    const uint8_t *seq = mate ? A.seq2 : A.seq1, *qual = mate ? A.qual2 : A.qual1;
    const uint32_t *off = mate ? A.off2 : A.off1;
    const uint32_t km = (1u << K) - 1u;                      // K <= AQC_MAX_KMER (8)
    unsigned long long n_kmers = 0, n_reads = 0;             // warp-uniform, added to the QC object at the end
    unsigned long long *const kfirst = qd.kfirst, *const kcnt = qd.kcnt;

    const uint32_t stride = nctas * (uint32_t)nwarps;
#pragma unroll 1
    for (uint32_t p = S.lo + cta * (uint32_t)nwarps + (uint32_t)warp; p < S.hi; p += stride) {
        const uint64_t gidx = A.first_index + p;
        uint32_t a = off[p];
        int len = (int)(off[p + 1] - a);
        uint32_t e0 = 0, e1 = 0, e2 = 0, e3 = 0;
        int n_edits = 0, start1 = 0, start2 = 0;
        uint64_t order;
        if constexpr (POST) {
            if (!(A.p.qc_sample <= 0 || gidx + 1 < (uint64_t)A.p.qc_sample)) continue;       // preprocesser.py:624 (quirk Q10)
            const uint4 *r = reinterpret_cast<const uint4 *>(A.results + p);
            const uint4 w0 = r[0];
            if ((w0.x & 0xFFu) != (uint32_t)AQC_GOOD) continue;
            n_edits = (int)((w0.x >> 8) & 0xFFu);
            start1 = (int)(w0.x >> 16); start2 = (int)(w0.y >> 16);
            a += (uint32_t)(mate ? start2 : start1);                                          // the final slice: trim + adapter cut
            len = (int)((mate ? w0.z : w0.y) & 0xFFFFu);
            if (n_edits) { const uint4 w1 = r[1]; e0 = w1.x; e1 = w1.y; e2 = w1.z; e3 = w1.w; }
            order = gidx;
        } else {
            if (gidx < A.stat_lo || gidx >= A.stat_hi) continue;
            order = A.order_base + (gidx - A.stat_lo);
        }
        if (len <= 0) {          // an empty read runs no loop of statRead but is still counted: gcHistogram[0] += 1 (:112)
            if (lane == 0) atomicAdd(&s_gch[0], 1u);
            n_reads++;
            continue;
        }
        if (len < 5) { if (lane == 0) atomicExch(A.error_flag, AQC_ERR_TOO_SHORT_STAT); continue; }     // reference: IndexError (:97-108)
        if (len > MAXB) { if (lane == 0) atomicExch(A.error_flag, AQC_ERR_TOO_LONG); continue; }        // the host picks NW; defensive
        const int chunks = (len + 31) >> 5;
        const int nk = len - K;                              // k-mers start at i < len - K (quirk Q11)

        // ---- the read's bytes: lane l of chunk w holds cycle 32w + l ----
        uint32_t b[NW], q[NW];
        {
            const uint8_t *s = seq + a, *qp = qual + a;
#pragma unroll
            for (int w = 0; w < NW; w++) {
                const int pos = 32 * w + lane;
                b[w] = 0; q[w] = 0;
                if (pos < len) { b[w] = s[pos]; q[w] = qp[pos]; }
            }
        }
        if constexpr (POST) {
            if (n_edits) {                                   // bytes rewritten by the correction walk (preprocesser.py:575-592)
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (k < n_edits) {
                        const uint32_t e = k == 0 ? e0 : (k == 1 ? e1 : (k == 2 ? e2 : e3));
                        const int kind = (int)AQC_EDIT_KIND(e);
                        int pos = -1;
                        uint32_t nb = 0, nq = 0;
                        if (kind == 0 && mate == 0) { pos = (int)AQC_EDIT_POS(e) - start1; nb = AQC_EDIT_BASE(e); nq = AQC_EDIT_QUAL(e); }
                        else if (kind == 1 && mate == 1) { pos = (int)AQC_EDIT_POS(e) - start2; nb = AQC_EDIT_BASE(e); nq = AQC_EDIT_QUAL(e); }
                        else if (kind == 2) { pos = mate == 0 ? (int)AQC_EDIT_POS(e) - start1 : (int)AQC_EDIT_POS2(e) - start2; nq = '!'; }
                        if (pos >= 0 && pos < len && (pos & 31) == lane) {
#pragma unroll
                            for (int w = 0; w < NW; w++)
                                if ((pos >> 5) == w) { if (kind != 2) b[w] = nb; q[w] = nq; }
                        }
                    }
                }
            }
        }

        // ---- per-cycle counters (:76-96), G/C count (:93-94), plane ballots for the k-mers, "differs from the next base" ----
        uint32_t k0[NW + 1], k1[NW + 1], kv[NW + 1], nq[NW + 1];
        int gc = 0;
        bool foreign_kmer = false;
#pragma unroll
        for (int w = 0; w < NW; w++) {
            k0[w] = k1[w] = kv[w] = nq[w] = 0;
            if (w < chunks) {                                // warp-uniform
                const int pos = 32 * w + lane;
                const bool valid = pos < len;
                const uint32_t l2 = valid ? lut2[b[w]] : 4u;
                if (valid) {
                    const uint32_t cls = l2 & 7u;
                    atomicAdd(&s_cnt[cls * MAXB + pos], 1u);
                    atomicAdd(&s_qs[cls * MAXB + pos], q[w]);
                }
                gc += __popc(__ballot_sync(FULL, valid && (l2 & 8u)));
                k0[w] = __ballot_sync(FULL, valid && (l2 & 0x10u));
                k1[w] = __ballot_sync(FULL, valid && (l2 & 0x20u));
                kv[w] = __ballot_sync(FULL, valid && (l2 & 0x40u));
                if (kv[w] != lowmask(len - 32 * w)) foreign_kmer = true;
                uint32_t nx = __shfl_down_sync(FULL, b[w], 1);
                const uint32_t first_of_next = __shfl_sync(FULL, b[w + 1 < NW ? w + 1 : w], 0);
                if (lane == 31) nx = (w + 1 < NW) ? first_of_next : 0u;
                nq[w] = __ballot_sync(FULL, pos + 1 < len && b[w] != nx);
            }
        }
        k0[NW] = k1[NW] = kv[NW] = nq[NW] = 0;

        // ---- discontinuity (:97-108): unequal neighbours inside the 5-base window around the cycle, clamped at both ends ----
#pragma unroll
        for (int w = 0; w < NW; w++) {
            if (w < chunks) {
                const int pos = 32 * w + lane;
                if (pos < len) {
                    int left = pos - 2;
                    if (left < 0) left = 0;
                    else if (pos + 3 >= len) left = len - 5;
                    const bool prev = (left >> 5) < w;       // the window starts in chunk w-1 (never earlier) or in chunk w
                    const uint32_t lo_w = prev ? nq[w > 0 ? w - 1 : 0] : nq[w];
                    const uint32_t hi_w = prev ? nq[w] : nq[w + 1];
                    const uint32_t d = (uint32_t)__popc(__funnelshift_r(lo_w, hi_w, (uint32_t)left & 31u) & 0xFu);
                    if (d) atomicAdd(&s_disc[pos], d);
                }
            }
        }

        // ---- k-mers (:113-122): the K-bit windows of the plane ballots are the dense table index ----
        if (__builtin_expect(foreign_kmer, 0)) {             // warp-uniform: the side-table path reads the bytes of the k-mer
            __syncwarp();
#pragma unroll
            for (int w = 0; w < NW; w++) { const int pos = 32 * w + lane; if (pos < len) scratch[pos] = (uint8_t)b[w]; }
            __syncwarp();
        }
        const unsigned long long when0 = order << 11;
#pragma unroll
        for (int w = 0; w < NW; w++) {
            if (32 * w < nk) {                               // warp-uniform
                const int i = 32 * w + lane;
                if (i < nk) {
                    const uint32_t w0 = __funnelshift_r(k0[w], k0[w + 1], lane) & km;
                    const uint32_t w1 = __funnelshift_r(k1[w], k1[w + 1], lane) & km;
                    const uint32_t wv = __funnelshift_r(kv[w], kv[w + 1], lane) & km;
                    const unsigned long long when = when0 | ((unsigned long long)i << 1);
                    if (__builtin_expect(wv == km, 1)) {
                        const uint32_t idx = (w1 << K) | w0;
                        const uint32_t sh = (idx & 1u) << 4;
                        const uint32_t old = atomicAdd(&ktab[idx >> 1], 1u << sh);
                        if (__builtin_expect(((old >> sh) & 0xFFFFu) == KTAB_SPILL - 1u, 0)) {
                            atomicSub(&ktab[idx >> 1], KTAB_SPILL << sh);
                            atomicAdd(&kcnt[idx], (unsigned long long)KTAB_SPILL);
                        }
                        // only the DIRECT first sighting is tracked on the device (see stat_read)
                        if (__builtin_expect(!((kbits[idx >> 5] >> (idx & 31u)) & 1u), 0)) {
                            if (__ldcg(&kfirst[idx]) > when) atomicMin(&kfirst[idx], when);
                        }
                    } else {                                 // a byte outside A,C,G,T in the k-mer: side table, keyed by its bytes
                        unsigned long long key = 0, rkey = 0;
                        bool foreign = false;
                        for (int j = 0; j < K; j++) {
                            const unsigned long long bj = scratch[i + j];
                            key = (key << 8) | bj;
                            rkey |= (unsigned long long)lut3[bj] << (8 * j);
                            foreign |= (lut1[bj] & 15u) == 15u;
                        }
                        if (key == AQC_KMER_NEVER || rkey == AQC_KMER_NEVER) atomicExch(A.error_flag, AQC_ERR_INVALID);
                        else {
                            const int h = side_slot(qd, key);
                            const int hr = side_slot(qd, rkey);
                            if (h < 0 || hr < 0) atomicExch(A.error_flag, AQC_ERR_KMER_TABLE_FULL);
                            else {
                                atomicAdd(&qd.scnt[h], 1ULL);
                                first_min(&qd.sfirst[h], when);
                                if (foreign) first_min(&qd.sseed[hr], when | 1ULL);
                            }
                        }
                    }
                }
            }
        }
        if (lane == 0) atomicAdd(&s_gch[gc], 1u);            // :112
        if (nk > 0) n_kmers += (unsigned long long)nk;       // totalKmer :114
        n_reads++;
    }

    // ---- epilogue: the CTA's histograms go to the QC object ----
    __syncthreads();
    for (uint32_t i = tid; i < TW; i += blockDim.x) {
        const uint32_t v = ktab[i];
        if (v & 0xFFFFu) atomicAdd(&kcnt[2 * i], (unsigned long long)(v & 0xFFFFu));
        if (v >> 16) atomicAdd(&kcnt[2 * i + 1], (unsigned long long)(v >> 16));
    }
    for (int i = tid; i < QC_CLASSES * MAXB; i += blockDim.x) {
        const int c = i / MAXB, pos = i - c * MAXB;
        const uint32_t n = s_cnt[i], qs = s_qs[i];
        if (n) atomicAdd(&qd.cls_cnt[c * AQC_MAX_LEN + pos], (unsigned long long)n);
        if (qs) atomicAdd(&qd.cls_qsum[c * AQC_MAX_LEN + pos], (unsigned long long)qs);
    }
    for (int i = tid; i < MAXB; i += blockDim.x) { const uint32_t v = s_disc[i]; if (v) atomicAdd(&qd.disc[i], (unsigned long long)v); }
    for (int i = tid; i <= MAXB; i += blockDim.x) { const uint32_t v = s_gch[i]; if (v) atomicAdd(&qd.gchist[i], (unsigned long long)v); }
    if (lane == 0) {
        if (n_kmers) atomicAdd(&qd.scal[0], n_kmers);
        if (n_reads) atomicAdd(&qd.scal[1], n_reads);
    }
}

}  // namespace aqc
