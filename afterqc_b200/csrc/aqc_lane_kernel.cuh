// aqc_lane_kernel.cuh -- the filter path for short reads (every mate <= 32*NW bases, NW <= 8): ONE LANE PER PAIR.
//
// pair_kernel (aqc_kernel.cuh) gives a whole warp to one pair: simple and length-agnostic, but a 150-base mate fills
// only 5 of the 32 lanes of a plane set and every piece of per-pair bookkeeping is executed 32 times.  For Illumina
// lengths the planes of both mates fit in registers (2 x NW words per plane), so here a warp takes a tile of 32 pairs
// and every lane runs the whole reference loop body (preprocesser.py:455-631) for its own pair:
//   * the warp's private shared-memory stage receives the tile's three byte columns (bases 1, qualities 1, bases 2;
//     qualities 2 are only ever touched by the correction walk and the statistics) with 1-D TMA bulk copies; the lane
//     converts its own reads to bit-planes (SWAR, 4 bases per 32-bit word: codes from the ASCII bits, re-encoding
//     check for foreign bytes, multiply-gather of the code bits) and counts low qualities; after that the stage is
//     free and the next tile's copy overlaps the rest of the work;
//   * util.overlap_hm (util.py:158-212): per 32 candidate offsets the lane funnel-shifts two plane words, XORs them with
//     the fixed mate's first word and keeps the offsets with < 3 mismatches in the first 32 positions as a bit mask
//     (a necessary condition of the acceptance rule); the lanes then evaluate their candidates exactly, in scan order,
//     in lock step; the scanned mate's words are rotated one register per round so that the code does not depend on
//     the round;
//   * hasPolyX is screened by a multi-word run-length test per lane; adapter cut, rescan, correction walk and the
//     classifier are per-lane code on registers, reading the few bytes the walk needs from HBM (L2);
//   * rare work that is better done by a whole warp is handed over by ballot: exact hasPolyX of screened reads,
//     statRead of the sampled good pairs (the trimmed, corrected reads are rebuilt in a per-warp scratch from the
//     result record), and pairs holding a byte outside A,C,G,T,N, which are appended to a list that pair_kernel
//     processes in its list mode right after this kernel.
// Counters: per-tile packed warp reductions into lane-owned 64-bit registers; histograms and the error matrix are
// shared-memory atomics.  Results are bit-identical to pair_kernel and the oracle.
#pragma once
#include "aqc_device.cuh"

namespace aqc {

constexpr int LANE_MAX_WARPS = 4;

struct LArgs {
    KArgs k;                      // batch, parameters, outputs (tile_pairs/col_cap as used by this kernel)
    uint32_t *fb_list;            // pairs that need the general (warp-per-pair) path
    uint32_t *fb_count;
    int lane_col_cap;             // bytes reserved per column in a warp's stage
    uint32_t *tile_counter;       // lane2_kernel: next unclaimed tile (zeroed before the launch)
    uint32_t *skip_bits;          // SMODE 2: bit pp set = pair pp went to pair_kernel's list mode (which also does its statistics)
};

template <int NW> struct LanePlanes {
    uint32_t p0[NW], p1[NW], pn[NW];
};

// python slice semantics of trim() (preprocesser.py:19-28)
__device__ __forceinline__ void lane_py_trim(int len, int front, int tail, int &start, int &newlen) {
    int s = front < len ? front : len;
    int e = tail > 0 ? len - tail : len;
    if (e < 0) e = 0;
    if (e < s) e = s;
    start = s; newlen = e - s;
}

// multi-word logical right shift by s bits, 0 <= s < 32*NW (zeros enter at the top)
template <int NW>
__device__ __forceinline__ void shr_bits(uint32_t (&X)[NW], int s) {
    static_assert(NW >= 2 && NW <= 8, "NW in 2..8");
    const int q = s >> 5, sh = s & 31;
    if (q) {
        if (q & 1) {
#pragma unroll
            for (int i = 0; i < NW; i++) X[i] = (i + 1 < NW) ? X[i + 1 < NW ? i + 1 : 0] : 0u;
        }
        if (q & 2) {
#pragma unroll
            for (int i = 0; i < NW; i++) X[i] = (i + 2 < NW) ? X[i + 2 < NW ? i + 2 : 0] : 0u;
        }
        if (q & 4) {
#pragma unroll
            for (int i = 0; i < NW; i++) X[i] = (i + 4 < NW) ? X[i + 4 < NW ? i + 4 : 0] : 0u;
        }
    }
#pragma unroll
    for (int i = 0; i < NW; i++) X[i] = __funnelshift_r(X[i], (i + 1 < NW) ? X[i + 1 < NW ? i + 1 : 0] : 0u, sh);
}

// Planes of one read (bytes in the warp's stage, any alignment).  p0/p1 = bits 1/2 of the ASCII byte
// (A0 C1 T2 G3), pn = 'N' (its code bits are cleared); exotic = some byte is not A,C,G,T,N.
// The chunk loop is rolled (one copy of the SWAR code in the instruction cache): every iteration converts the next 32
// bases into the TOP word of the plane registers and moves the others down one word, so after NW iterations word c
// holds chunk c.
template <int NW>
__device__ __forceinline__ void lane_convert(const uint8_t *s, int len, LanePlanes<NW> &P, bool &exotic, int &n_count) {
    const smem_addr_t a = smem_addr(s);
    const smem_addr_t w = a & ~(smem_addr_t)3;
    const int sh = (int)(a & 3) * 8;
    exotic = false;
    n_count = 0;
    uint32_t prev = len > 0 ? lds_u32(w) : 0u;
#ifdef AQC_LANE_UNROLL_CONVERT
#pragma unroll
#else
#pragma unroll 1
#endif
    for (int c = 0; c < NW; c++) {
        uint32_t p0 = 0, p1 = 0, pn = 0;
        const int nvalid = len - 32 * c;
        if (nvalid > 0) {
            uint32_t v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t cur = lds_u32(w + 4 * (8 * c + j + 1));
                v[j] = __funnelshift_r(prev, cur, sh);
                prev = cur;
            }
            uint32_t rlo = 0, rhi = 0, bad = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t t = v[j] & 0x06060606u;
                const uint32_t tt = t >> 1;                                                     // 2-bit codes
                // codes -> ASCII: the four codes of the word become the four selector nibbles of a table lookup
                const uint32_t e = prmt_raw(0x47544341u, 0u, prmt_raw(tt + (tt >> 4), 0u, 0x4420u));
                bad |= e ^ v[j];
                const uint32_t z = ((t << 2) + tt) & 0x11111111u;                               // bit0 = code bit 0, bit4 = code bit 1
                const uint32_t r = z * 0x01020408u;                     // byte 3 = nibble of plane 0 | nibble of plane 1 << 4
                constexpr uint32_t sel[4] = {0x3217u, 0x3270u, 0x3710u, 0x7210u};               // byte 3 of r -> byte j of the accumulator
                if (j < 4) rlo = __byte_perm(rlo, r, sel[j & 3]); else rhi = __byte_perm(rhi, r, sel[j & 3]);
            }
            // de-interleave the nibbles: bytes of rlo/rhi hold (p1 nibble << 4 | p0 nibble) of 4 bases each
            {
                const uint32_t l0 = rlo & 0x0F0F0F0Fu, h0 = rhi & 0x0F0F0F0Fu;
                const uint32_t l1 = (rlo >> 4) & 0x0F0F0F0Fu, h1 = (rhi >> 4) & 0x0F0F0F0Fu;
                const uint32_t a0 = (l0 | (l0 >> 4)) & 0x00FF00FFu, b0 = (h0 | (h0 >> 4)) & 0x00FF00FFu;
                const uint32_t a1 = (l1 | (l1 >> 4)) & 0x00FF00FFu, b1 = (h1 | (h1 >> 4)) & 0x00FF00FFu;
                p0 = __byte_perm(a0, b0, 0x6420);
                p1 = __byte_perm(a1, b1, 0x6420);
            }
            const uint32_t vm = lowmask(nvalid);
            p0 &= vm; p1 &= vm;
            if (__builtin_expect(bad != 0u, 0)) {           // some byte of the 32 is not A,C,G,T (maybe beyond the read)
                uint32_t nb = 0, xb = 0;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const uint32_t t = v[j] & 0x06060606u;
                    const uint32_t tt = t >> 1;
                    const uint32_t e = __byte_perm(0x47544341u, 0u, __byte_perm(tt | (tt >> 4), 0u, 0x4420));
                    const uint32_t isn = ~hibit_nonzero(v[j] ^ 0x4E4E4E4Eu) & 0x80808080u;
                    const uint32_t isbad = hibit_nonzero(e ^ v[j]);
                    nb |= gather4(isn >> 7) << (4 * j);
                    xb |= gather4((isbad & ~isn) >> 7) << (4 * j);
                }
                nb &= vm; xb &= vm;
                if (xb) exotic = true;
                pn = nb;
                p0 &= ~nb; p1 &= ~nb;
                n_count += __popc(nb);
            }
        }
#ifdef AQC_LANE_UNROLL_CONVERT
        P.p0[c] = p0; P.p1[c] = p1; P.pn[c] = pn;
#else
#pragma unroll
        for (int i = 0; i < NW; i++) {
            P.p0[i] = (i + 1 < NW) ? P.p0[i + 1 < NW ? i + 1 : 0] : p0;
            P.p1[i] = (i + 1 < NW) ? P.p1[i + 1 < NW ? i + 1 : 0] : p1;
            P.pn[i] = (i + 1 < NW) ? P.pn[i + 1 < NW ? i + 1 : 0] : pn;
        }
#endif
    }
}

// lowQualityNum (preprocesser.py:61-68) on the lane's own quality bytes: aligned words, byte-masked at both ends
__device__ __forceinline__ int lane_lowq(const uint8_t *q, int len, int thr) {
    if (len <= 0 || thr <= 0) return 0;
    if (thr >= 128) {                                          // outside the SWAR domain (never with sane -q): byte loop
        int n = 0;
        for (int i = 0; i < len; i++) n += (int)q[i] < thr;
        return n;
    }
    const smem_addr_t a = smem_addr(q);
    const smem_addr_t w = a & ~(smem_addr_t)3;
    const int lead = (int)(a & 3);
    const int total = lead + len;                              // bytes from the aligned start to the end of the read
    const int last = (total - 1) >> 2;                         // index of the last word (<= 64: per-byte sums stay below 256)
    const uint32_t t4 = (uint32_t)thr * 0x01010101u;
    // 0x80 in every byte < thr:  byte >= thr  <=>  high bit of ((byte | 0x80) - thr) | byte   (thr < 128)
    auto low_of = [&](uint32_t v) -> uint32_t { return ~(((v | 0x80808080u) - t4) | v) & 0x80808080u; };
    uint32_t first = low_of(lds_u32(w)) & ~bytemask_lo(lead);
    if (last == 0) first &= bytemask_lo(total);
    uint32_t acc = first >> 7;
    for (int j = 1; j < last; j++) acc += low_of(lds_u32(w + 4 * j)) >> 7;
    if (last > 0) acc += (low_of(lds_u32(w + 4 * last)) & bytemask_lo(total - 4 * last)) >> 7;
    const uint32_t h = (acc & 0x00FF00FFu) + ((acc >> 8) & 0x00FF00FFu);     // the four byte sums can add up to more than 255
    return (int)((h + (h >> 16)) & 0xFFFFu);
}

// Run-length screen of hasPolyX (see polyx_screen_fast): true = the read needs the exact test.
template <int NW>
__device__ __forceinline__ bool lane_polyx_screen(const uint32_t (&p0)[NW], const uint32_t (&p1)[NW], const uint32_t (&pn)[NW],
                                                  int len, int maxPoly, int m) {
    if (len < maxPoly || m < 0) return false;
    if (m == 0 || m > 31) return true;
    uint32_t y[NW];
#pragma unroll
    for (int i = 0; i < NW; i++) {
        const uint32_t u0 = i ? p0[i - 1 >= 0 ? i - 1 : 0] : 0u, u1 = i ? p1[i - 1 >= 0 ? i - 1 : 0] : 0u, un = i ? pn[i - 1 >= 0 ? i - 1 : 0] : 0u;
        const uint32_t d = (p0[i] ^ __funnelshift_l(u0, p0[i], 1)) | (p1[i] ^ __funnelshift_l(u1, p1[i], 1)) | (pn[i] ^ __funnelshift_l(un, pn[i], 1));
        y[i] = ~d & lowmask(len - 32 * i);                     // bit x: base x equals base x-1
    }
    y[0] &= ~1u;
    int t = 1;
    while (t < m) {                                             // y[x] := run of m "same as previous" bits starts at x
        const int step = min(t, m - t);
#pragma unroll
        for (int i = 0; i < NW; i++) y[i] &= __funnelshift_r(y[i], (i + 1 < NW) ? y[i + 1 < NW ? i + 1 : 0] : 0u, step);
        t += step;
    }
    uint32_t any = 0;
#pragma unroll
    for (int i = 0; i < NW; i++) any |= y[i];
    return any != 0u;
}

// util.overlap_hm (util.py:158-212), one direction, lane-per-pair.  S is scanned at offsets 0 .. lenS-31 against the
// fixed read F; `active` lanes take part, the others idle through the warp-uniform loops.
template <int NW>
__device__ __forceinline__ bool lane_scan_dir(uint32_t (&S0)[NW], uint32_t (&S1)[NW], uint32_t (&SN)[NW], int lenS,
                                              const uint32_t (&F0)[NW], const uint32_t (&F1)[NW], const uint32_t (&FN)[NW], int lenF,
                                              bool active, int &o_out, int &ol_out, int &mm_out) {
    const int nOff = active ? lenS - 30 : 0;                    // overlap_require = 30 (util.py:164)
    const int maxOff = (int)__reduce_max_sync(FULL, (unsigned)max(nOff, 0));
    const int rounds = (maxOff + 31) >> 5;
    const bool slow = lenF < 32;                                // first window shorter than 32: every offset is evaluated exactly
    const uint32_t f0 = F0[0];
#ifdef AQC_LANE_TWO_PLANE_FILTER
    const uint32_t f1 = F1[0];
#endif
    bool found = false;
#pragma unroll 1
    for (int r = 0; r < rounds; r++) {
        const int rem = nOff - (r << 5);
        if (!__any_sync(FULL, !found && rem > 0)) break;
        uint32_t cm = 0;
        if (!found && rem > 0) {
            if (!slow) {
#pragma unroll
                for (int b = 0; b < 32; b++) {
#ifdef AQC_LANE_TWO_PLANE_FILTER
                    const uint32_t x = (__funnelshift_r(S0[0], S0[1], b) ^ f0) | (__funnelshift_r(S1[0], S1[1], b) ^ f1);
#else
                    // one code bit is enough for a necessary condition: equal bases have equal bits, and 32 random positions
                    // differ in fewer than 3 of them with probability 1.2e-7
#ifdef AQC_LANE_IMAD_SHIFT
                    // tuning variant: the window as two multiplies (FMA pipe) instead of one funnel shift (ALU pipe, the busy one)
                    const uint32_t x = (b == 0 ? S0[0] : __umulhi(S0[0], 1u << ((32 - b) & 31)) + S0[1] * (1u << ((32 - b) & 31))) ^ f0;
#else
                    const uint32_t x = __funnelshift_r(S0[0], S0[1], b) ^ f0;
#endif
#endif
                    if (__popc(x) < 3) cm |= 1u << b;
                }
                cm &= lowmask(rem);
                if (rem <= 32) cm |= 1u << (rem - 1);           // offset lenS-31 sees only 31 positions: always evaluated exactly
            } else {
                cm = lowmask(rem);
            }
        }
        while (__any_sync(FULL, cm != 0u)) {                    // candidates in scan order, all lanes in lock step
            if (cm) {
                const int b = __ffs(cm) - 1;
                cm &= cm - 1;
                const int oc = (r << 5) + b;
                const int olc = min(lenS - oc, lenF);
                const int l50 = min(50, olc);
                int mm = 0, mm50 = 0;
#pragma unroll
                for (int w = 0; w < NW; w++) {
                    const uint32_t n0 = (w + 1 < NW) ? S0[w + 1 < NW ? w + 1 : 0] : 0u;
                    const uint32_t n1 = (w + 1 < NW) ? S1[w + 1 < NW ? w + 1 : 0] : 0u;
                    const uint32_t nn = (w + 1 < NW) ? SN[w + 1 < NW ? w + 1 : 0] : 0u;
                    uint32_t xw = (__funnelshift_r(S0[w], n0, b) ^ F0[w]) | (__funnelshift_r(S1[w], n1, b) ^ F1[w]) | (__funnelshift_r(SN[w], nn, b) ^ FN[w]);
                    xw &= lowmask(olc - 32 * w);
                    mm += __popc(xw);
                    if (w < 2) mm50 += __popc(xw & lowmask(l50 - 32 * w));
                }
                if (mm50 < 3 && (mm < 3 || olc >= 52)) {        // closed form of the leaked loop variable (quirk Q6)
                    found = true; o_out = oc; ol_out = olc; mm_out = mm; cm = 0;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < NW; i++) {                           // next round: word i+1 becomes word i
            S0[i] = (i + 1 < NW) ? S0[i + 1 < NW ? i + 1 : 0] : 0u;
            S1[i] = (i + 1 < NW) ? S1[i + 1 < NW ? i + 1 : 0] : 0u;
            SN[i] = (i + 1 < NW) ? SN[i + 1 < NW ? i + 1 : 0] : 0u;
        }
    }
    return found;
}

// util.overlap(r1, r2) for the active lanes: forward offsets, then reverse (util.py:172-209), else (0,0,0) (:212)
template <int NW>
__device__ __forceinline__ void lane_overlap(const LanePlanes<NW> &P1, const LanePlanes<NW> &RC, int len1, int len2, bool active,
                                             int &offset, int &ol, int &diff) {
    bool found = false;
    int o = 0;
#pragma unroll 1
    for (int dir = 0; dir < 2; dir++) {
        uint32_t S0[NW], S1[NW], SN[NW], F0[NW], F1[NW], FN[NW];
#pragma unroll
        for (int i = 0; i < NW; i++) {
            S0[i] = dir ? RC.p0[i] : P1.p0[i]; S1[i] = dir ? RC.p1[i] : P1.p1[i]; SN[i] = dir ? RC.pn[i] : P1.pn[i];
            F0[i] = dir ? P1.p0[i] : RC.p0[i]; F1[i] = dir ? P1.p1[i] : RC.p1[i]; FN[i] = dir ? P1.pn[i] : RC.pn[i];
        }
        const bool act = active && !found;
        if (!__any_sync(FULL, act)) break;
        int oo = 0, ool = 0, omm = 0;
        const bool f = lane_scan_dir<NW>(S0, S1, SN, dir ? len2 : len1, F0, F1, FN, dir ? len1 : len2, act, oo, ool, omm);
        if (f) { found = true; o = dir ? -oo : oo; ol = ool; diff = omm; }
    }
    if (active) {
        if (found) offset = o;
        else { offset = 0; ol = 0; diff = 0; }
    }
}

}  // namespace aqc
#include "aqc_stat2.cuh"      // stat_tile: statRead with one lane per read (needs LanePlanes / shr_bits above)
namespace aqc {

// dynamic shared memory of one CTA:
//   [nwarps][ 3 * lane_col_cap ]           per-warp stage: bases 1 | qualities 1 | bases 2 (TMA destinations)
//   [nwarps][ 4 * 32*NW ]                  per-warp scratch of the statistics hand-over
//   luts (768 B)
//   qc acc [2][5][max_len] u32, qc disc [2][max_len] u32, overlap_hist [max_len+1], distance_hist [max_len+1], err matrix [16]
// SMODE: sampled postfilter statistics -- 0 = stat_read hand-over (one warp per read), 1 = stat_tile (one lane per read), both in this kernel;
//        2 = none here: stat_lane_kernel<.., POST> runs after this kernel and pair_kernel's list mode (pairs handed over are marked in skip_bits)
template <bool PAIRED, int NW, int SMODE = 0>
__global__ void __launch_bounds__(LANE_MAX_WARPS * 32, (NW > 5 ? 3 : 4)) lane_kernel(const __grid_constant__ LArgs L) {
    AQC_DYN_SMEM(smem_raw);
    __shared__ __align__(8) uint64_t full_bar[LANE_MAX_WARPS];
    const KArgs &A = L.k;
    constexpr bool paired = PAIRED;
    constexpr int MAXB = 32 * NW;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = blockDim.x >> 5;
    const int col_cap = L.lane_col_cap;
    const int ncols = paired ? 3 : 2;

    uint8_t *stage = smem_raw + (size_t)warp * ncols * col_cap;
    uint8_t *scratch = smem_raw + (size_t)nwarps * ncols * col_cap + (size_t)warp * 4 * MAXB;
    uint8_t *lutbase = smem_raw + (size_t)nwarps * (ncols * col_cap + 4 * MAXB);
    const uint8_t *lut1 = lutbase, *lut2 = lutbase + 256, *lut3 = lutbase + 512;
    uint32_t *s_acc = reinterpret_cast<uint32_t *>(lutbase + 768);
    uint32_t *s_disc = s_acc + 2 * QC_CLASSES * A.max_len;
    uint32_t *s_ovh = s_disc + 2 * A.max_len;
    uint32_t *s_dih = s_ovh + (A.max_len + 1);
    uint32_t *s_em = s_dih + (A.max_len + 1);
    const int n_qc_words = 2 * QC_CLASSES * A.max_len + 2 * A.max_len;
    const int n_acc_words = n_qc_words + 2 * (A.max_len + 1) + 16;

    for (int i = tid; i < 768; i += blockDim.x) lutbase[i] = reinterpret_cast<const uint8_t *>(A.luts)[i];
    for (int i = tid; i < n_acc_words; i += blockDim.x) s_acc[i] = 0;
    if (tid == 0) {
        for (int s = 0; s < nwarps; s++) mbar_init(&full_bar[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    QcSmem qsm; qsm.acc = s_acc; qsm.disc = s_disc; qsm.max_len = A.max_len;
    const uint32_t flush_limit = QC_FLUSH_READS / (uint32_t)nwarps;
    uint32_t stat_since_flush = 0;

    // lane i owns scalar counter i
    unsigned long long wc0 = 0;

    const uint32_t gw = blockIdx.x * (uint32_t)nwarps + (uint32_t)warp;
    const uint32_t W = gridDim.x * (uint32_t)nwarps;
    uint64_t *bar = &full_bar[warp];

    // offsets of the lane's pair in tile t (pairs beyond n: empty records at the end of the columns)
    auto load_offsets = [&](uint32_t t, uint32_t &a1, uint32_t &e1, uint32_t &a2, uint32_t &e2) {
        const uint32_t pp = min(t * 32u + (uint32_t)lane, A.n), pq = min(pp + 1u, A.n);
        a1 = A.off1[pp]; e1 = A.off1[pq];
        a2 = 0; e2 = 0;
        if (paired) { a2 = A.off2[pp]; e2 = A.off2[pq]; }
    };
    // producer (whole warp computes, lane 0 issues): bulk copies of the tile's columns into the warp's stage
    auto issue_tile = [&](uint32_t a1, uint32_t e1, uint32_t a2, uint32_t e2) {
        const uint32_t f1 = __shfl_sync(FULL, a1, 0), l1 = __shfl_sync(FULL, e1, 31);
        const uint32_t f2 = __shfl_sync(FULL, a2, 0), l2 = __shfl_sync(FULL, e2, 31);
        if (lane == 0) {
            const uint32_t g1 = f1 & ~15u, bytes1 = (l1 - g1 + 15u) & ~15u;
            const uint32_t g2 = f2 & ~15u, bytes2 = paired ? ((l2 - g2 + 15u) & ~15u) : 0u;
            mbar_expect_tx(bar, 2 * bytes1 + bytes2);
            if (bytes1) {
                bulk_g2s(stage, A.seq1 + g1, bytes1, bar);
                bulk_g2s(stage + col_cap, A.qual1 + g1, bytes1, bar);
            }
            if (paired && bytes2) bulk_g2s(stage + 2 * col_cap, A.seq2 + g2, bytes2, bar);
        }
    };

    uint32_t a1 = 0, e1 = 0, a2 = 0, e2 = 0;
    uint32_t t = gw;
    if (t < A.num_tiles) {
        load_offsets(t, a1, e1, a2, e2);
        issue_tile(a1, e1, a2, e2);
    }
    uint32_t parity = 0;

#pragma unroll 1
    for (; t < A.num_tiles; t += W) {
        const uint32_t tn = t + W;
        uint32_t na1 = 0, ne1 = 0, na2 = 0, ne2 = 0;
        if (tn < A.num_tiles) load_offsets(tn, na1, ne1, na2, ne2);

        mbar_wait(bar, parity);
        parity ^= 1u;

        const uint32_t pp = t * 32u + (uint32_t)lane;
        const bool valid = pp < A.n;
        const uint64_t gidx = A.first_index + pp;
        const uint32_t g1 = __shfl_sync(FULL, a1, 0) & ~15u;
        const uint32_t g2 = __shfl_sync(FULL, a2, 0) & ~15u;
        const int olen1 = (int)(e1 - a1), olen2 = paired ? (int)(e2 - a2) : 0;

        int start1 = 0, len1 = olen1, start2 = 0, len2 = olen2;
        int cls = AQC_GOOD;
        bool live = valid;                                       // still walking the loop body
        bool fallback = false;
        LanePlanes<NW> P1, RC;
        int n1 = 0, n2 = 0, lowq1 = 0;
#pragma unroll
        for (int i = 0; i < NW; i++) { P1.p0[i] = P1.p1[i] = P1.pn[i] = 0; RC.p0[i] = RC.p1[i] = RC.pn[i] = 0; }
        bool cand1 = false, cand2 = false;

        // ================================ phase A: the lane's bytes in the stage ================================
        if (valid) {
            if (olen1 > MAXB || olen2 > MAXB) {                  // the host picks NW from the longest read; defensive
                atomicExch(A.error_flag, AQC_ERR_TOO_LONG);
                live = false; cls = AQC_NUM_CLASSES;
            }
        }
        if (live) {
            const bool do_trim = (A.p.trim_front > 0 || A.p.trim_tail > 0);   // gate keyed on R1 only (quirk Q4)
            if (do_trim) {                                           // preprocesser.py:455-466
                lane_py_trim(olen1, A.p.trim_front, A.p.trim_tail, start1, len1);
                if (len1 < 5) { cls = AQC_BADTRIM1; live = false; }
                else if (paired) {
                    lane_py_trim(olen2, A.p.trim_front2, A.p.trim_tail2, start2, len2);
                    if (len2 < 5) { cls = AQC_BADTRIM2; live = false; }
                }
            }
            if (live && len1 < A.p.seq_len_req) { cls = AQC_BADLEN; live = false; }   // :476-479 (R2 never checked, quirk Q3)
        }
        if (live) {
            bool ex = false;
#pragma unroll 1
            for (int m = 0; m < (paired ? 2 : 1); m++) {              // one copy of the conversion + screen code for both mates
                const uint8_t *r = m ? stage + 2 * col_cap + (a2 - g2) + start2 : stage + (a1 - g1) + start1;
                const int len = m ? len2 : len1;
                LanePlanes<NW> F;
#pragma unroll
                for (int i = 0; i < NW; i++) F.p0[i] = F.p1[i] = F.pn[i] = 0;
                bool exm = false; int nn = 0;
                lane_convert<NW>(r, len, F, exm, nn);
                ex |= exm;
                const bool cand = A.p.poly_size_limit > 0 && lane_polyx_screen<NW>(F.p0, F.p1, F.pn, len, A.p.poly_size_limit, A.poly_m);
                if (m == 0) {
                    n1 = nn; cand1 = cand;
#pragma unroll
                    for (int i = 0; i < NW; i++) { P1.p0[i] = F.p0[i]; P1.p1[i] = F.p1[i]; P1.pn[i] = F.pn[i]; }
                } else {
                    n2 = nn; cand2 = cand;
                    // reverseComplement (util.py:42-51): reverse the 32*NW-bit strings, shift the read down to bit 0, flip plane 1
#pragma unroll
                    for (int i = 0; i < NW; i++) { RC.p0[i] = __brev(F.p0[NW - 1 - i]); RC.p1[i] = __brev(F.p1[NW - 1 - i]); RC.pn[i] = __brev(F.pn[NW - 1 - i]); }
                    shr_bits<NW>(RC.p0, MAXB - len2); shr_bits<NW>(RC.p1, MAXB - len2); shr_bits<NW>(RC.pn, MAXB - len2);
#pragma unroll
                    for (int i = 0; i < NW; i++) RC.p1[i] ^= lowmask(len2 - 32 * i) & ~RC.pn[i];
                }
            }
            if (A.p.unqualified_base_limit > 0) lowq1 = lane_lowq(stage + col_cap + (a1 - g1) + start1, len1, A.p.qualified_quality_phred + 33);
            if (ex) { fallback = true; live = false; cls = AQC_NUM_CLASSES; }
        }

        // ---- the stage is free: prefetch the next tile while the registers are worked on ----
        fence_proxy_async();
        __syncwarp();
        if (tn < A.num_tiles) issue_tile(na1, ne1, na2, ne2);

        // pairs with foreign bytes go to the general kernel (list mode of pair_kernel, launched right after this one)
        {
            const uint32_t fb = __ballot_sync(FULL, fallback);
            if (fb) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(L.fb_count, (uint32_t)__popc(fb));
                base = __shfl_sync(FULL, base, 0);
                if (fallback) L.fb_list[base + (uint32_t)__popc(fb & lowmask(lane))] = pp;
                if constexpr (SMODE == 2) { if (fallback) atomicOr(&L.skip_bits[pp >> 5], 1u << (pp & 31u)); }     // pair_kernel does their statistics
            }
        }

        // ================================ phase B: registers (+ a few bytes from HBM) ================================
        const uint8_t *G1 = A.seq1 + a1 + start1, *G1q = A.qual1 + a1 + start1;
        const uint8_t *G2 = paired ? A.seq2 + a2 + start2 : nullptr, *G2q = paired ? A.qual2 + a2 + start2 : nullptr;

        // hasPolyX (:482-490): exact window test of the screened reads, one read at a time by the whole warp
        if (A.p.poly_size_limit > 0) {
            uint32_t cb = __ballot_sync(FULL, live && (cand1 || cand2));
            bool poly = false;
            while (cb) {
                const int src = __ffs(cb) - 1;
                cb &= cb - 1;
                const uint32_t c1 = __shfl_sync(FULL, (uint32_t)cand1, src), c2 = __shfl_sync(FULL, (uint32_t)cand2, src);
                const uint32_t lo1 = __shfl_sync(FULL, (uint32_t)(uintptr_t)G1, src), hi1 = __shfl_sync(FULL, (uint32_t)((uintptr_t)G1 >> 32), src);
                const uint32_t lo2 = __shfl_sync(FULL, (uint32_t)(uintptr_t)G2, src), hi2 = __shfl_sync(FULL, (uint32_t)((uintptr_t)G2 >> 32), src);
                const int l1 = __shfl_sync(FULL, len1, src), l2 = __shfl_sync(FULL, len2, src);
                bool hit = false;
#pragma unroll 1
                for (int m = 0; m < 2 && !hit; m++) {
                    if (m ? c2 : c1) {
                        const uint8_t *ptr = reinterpret_cast<const uint8_t *>(((uintptr_t)(m ? hi2 : hi1) << 32) | (uintptr_t)(m ? lo2 : lo1));
                        hit = polyx_exact(ptr, m ? l2 : l1, A.p.poly_size_limit, A.p.allow_mismatch_in_poly, lut2, lane) != 0;
                    }
                }
                if (lane == src) poly = hit;
            }
            if (live && poly) { cls = AQC_BADPOL; live = false; }
        }
        if (live && A.p.unqualified_base_limit > 0 && lowq1 > A.p.unqualified_base_limit) { cls = AQC_BADLQC; live = false; }   // :493-501 (quirk Q2)
        if (live && A.p.n_base_limit > 0 && (n1 > A.p.n_base_limit || n2 > A.p.n_base_limit)) { cls = AQC_BADNCT; live = false; }   // :504-512

        uint32_t edits[4] = {0, 0, 0, 0};
        int n_edits = 0;
        int ov_off = 0, ov_len = 0, ov_diff = 0;
        // per-tile counter contributions of this lane
        uint32_t t_adapter_reads = 0, t_adapter_half = 0, t_overlapped = 0, t_ol = 0, t_dist = 0;
        uint32_t t_read_corr = 0, t_corr = 0, t_masked = 0, t_skipped = 0;

        if (paired && !A.p.no_overlap) {                            // :515-617
            int offset = 0, ol = 0, distance = 0;
            bool scanning = live;
#pragma unroll 1
            for (int pass = 0; pass < 2; pass++) {
                if (!__any_sync(FULL, scanning)) break;
                lane_overlap<NW>(P1, RC, len1, len2, scanning, offset, ol, distance);    // :516 / :534
                bool again = false;
                if (scanning && pass == 0) {
                    atomicAdd(&s_ovh[ol], 1u);                                           // :517
                    if (offset < 0 && ol > 30) {                                         // :520 adapter trimming
                        // rc(r2[0:ol]) = last ol bases of rc(r2): shift the rc planes down by len2-ol; r1 keeps its first ol bases
                        const int sh = len2 - ol;
                        if (sh > 0) { shr_bits<NW>(RC.p0, sh); shr_bits<NW>(RC.p1, sh); shr_bits<NW>(RC.pn, sh); }
#pragma unroll
                        for (int i = 0; i < NW; i++) { const uint32_t m = lowmask(ol - 32 * i); P1.p0[i] &= m; P1.p1[i] &= m; P1.pn[i] &= m; }
                        len1 = ol; len2 = ol;                                            // :522-525
                        t_adapter_half += (uint32_t)(-offset);                           // :526
                        t_adapter_reads += 1;
                        if (len1 < A.p.seq_len_req) {                                    // :529-532
                            ov_off = offset; ov_len = ol; ov_diff = distance;
                            cls = AQC_BADLEN; live = false;
                        } else again = true;
                    }
                }
                scanning = again;
            }
            if (live) {
                ov_off = offset; ov_len = ol; ov_diff = distance;
                atomicAdd(&s_dih[distance], 1u);                                         // :536
                if (distance > 3) { cls = AQC_BADDIFF; live = false; }                   // :538-541
            }
            if (live && ol > 30) {                                                       // :542
                t_overlapped = 1; t_ol = (uint32_t)ol; t_dist = (uint32_t)distance;
                if (distance > 0) {                                                      // :551
                    // mismatch mask of the walk alignment r1[len1-ol+o] vs rc[o] (always this alignment: quirk Q8)
                    uint32_t X0[NW], X1[NW], XN[NW];
#pragma unroll
                    for (int i = 0; i < NW; i++) { X0[i] = P1.p0[i]; X1[i] = P1.p1[i]; XN[i] = P1.pn[i]; }
                    const int oc = len1 - ol;
                    if (oc > 0) { shr_bits<NW>(X0, oc); shr_bits<NW>(X1, oc); shr_bits<NW>(XN, oc); }
                    uint32_t xx[NW];
#pragma unroll
                    for (int i = 0; i < NW; i++) xx[i] = ((X0[i] ^ RC.p0[i]) | (X1[i] ^ RC.p1[i]) | (XN[i] ^ RC.pn[i])) & lowmask(ol - 32 * i);
                    int corrected = 0, masked = 0, skipped = 0;
                    int em_cell[3] = {-1, -1, -1};
                    int done = 0;
#pragma unroll 1
                    while (done < distance) {
                        int o = -1;
#pragma unroll
                        for (int i = 0; i < NW; i++) {
                            if (o < 0 && xx[i]) { o = 32 * i + __ffs(xx[i]) - 1; xx[i] &= xx[i] - 1; }
                        }
                        if (o < 0) break;
                        const int p1 = len1 - ol + o, p2 = len2 - 1 - o;
                        const uint8_t b1 = G1[p1];                                 // :564
                        const uint8_t b2 = lut3[G2[p2]];                           // :565 util.complement
                        const uint8_t qa = G1q[p1], qb = G2q[p2];                  // :566-567
                        const int Qa = (int)qa - 33, Qb = (int)qb - 33;
                        bool fixed = false;
                        uint32_t e = 0;
                        int cell = -1;
                        if (Qa >= 30 && Qb <= 14) {                                // :571
                            if (b1 != 'N' && b2 != 'N') {
                                const uint32_t la = lut2[lut3[b1]], lc = lut2[lut3[b2]];
                                if ((la & 0x40u) && (lc & 0x40u)) cell = (int)((la & 7u) * 4u + (lc & 7u));   // :573
                            }
                            if (!A.p.no_correction) {                              // :574-578
                                const uint8_t nb = lut3[b1];
                                corrected++; fixed = true;
                                e = (uint32_t)(start2 + p2) | (1u << 10) | ((uint32_t)nb << 16) | ((uint32_t)qa << 24);
                            }
                        } else if (Qb >= 30 && Qa <= 14) {                         // :579
                            if (b1 != 'N' && b2 != 'N') {
                                const uint32_t la = lut2[b2], lc = lut2[b1];
                                if ((la & 0x40u) && (lc & 0x40u)) cell = (int)((la & 7u) * 4u + (lc & 7u));   // :581
                            }
                            if (!A.p.no_correction) {                              // :582-586
                                corrected++; fixed = true;
                                e = (uint32_t)(start1 + p1) | (0u << 10) | ((uint32_t)b2 << 16) | ((uint32_t)qb << 24);
                            }
                        }
                        if (!fixed) {                                              // :587-595
                            if (A.p.mask_mismatch) {
                                masked++;
                                e = (uint32_t)(start1 + p1) | (2u << 10) | ((uint32_t)(start2 + p2) << 16);
                            } else {
                                skipped++;
                                e = (uint32_t)(start1 + p1) | (3u << 10) | ((uint32_t)(start2 + p2) << 16);
                            }
                        }
#pragma unroll
                        for (int k = 0; k < 3; k++) if (k == done) { edits[k] = e; em_cell[k] = cell; }   // distance <= 3 here
                        done++;
                    }
                    n_edits = done;
                    if (corrected + masked + skipped == distance) {               // :603-610
#pragma unroll
                        for (int k = 0; k < 3; k++) if (em_cell[k] >= 0) atomicAdd(&s_em[em_cell[k]], 1u);
                        if (corrected > 0) t_read_corr = 1;
                        t_corr = (uint32_t)corrected; t_masked = (uint32_t)masked; t_skipped = (uint32_t)skipped;
                    } else { cls = AQC_BADMISMATCH; live = false; }               // :611-614
                }
            }
        }

        // ---- the 32-byte record ----
        if (valid && cls != AQC_NUM_CLASSES) {
            uint4 w0, w1;
            w0.x = (uint32_t)cls | ((uint32_t)n_edits << 8) | ((uint32_t)start1 << 16);
            w0.y = (uint32_t)len1 | ((uint32_t)start2 << 16);
            w0.z = (uint32_t)len2 | (((uint32_t)ov_off & 0xFFFFu) << 16);
            w0.w = (uint32_t)ov_len | ((uint32_t)ov_diff << 16);
            w1.x = edits[0]; w1.y = edits[1]; w1.z = edits[2]; w1.w = edits[3];
            uint4 *dst = reinterpret_cast<uint4 *>(&A.results[pp]);
            dst[0] = w0; dst[1] = w1;
        }

        // ---- counters: packed warp sums, lane i keeps scalar counter i (preprocesser.py:378-409) ----
        {
            const bool good = valid && cls == AQC_GOOD;
            const bool bad = valid && cls >= AQC_BADTRIM1 && cls <= AQC_BADMISMATCH;
            const uint32_t wa = (good ? 1u : 0u) | (t_overlapped << 8) | (t_read_corr << 16) | (t_adapter_reads << 24);
            const uint32_t wb = t_corr | (t_skipped << 8) | (t_masked << 16) | (t_dist << 24);
            const uint32_t wc = (good ? (uint32_t)len1 : 0u) | ((good ? (uint32_t)len2 : 0u) << 16);
            const uint32_t wd = t_ol | (t_adapter_half << 16);
            const uint32_t we = (bad && cls <= 4) ? (1u << (8 * (cls - 1))) : 0u;
            const uint32_t wf = (bad && cls >= 5) ? (1u << (8 * (cls - 5))) : 0u;
            const uint32_t sa = __reduce_add_sync(FULL, wa), sb = __reduce_add_sync(FULL, wb), sc = __reduce_add_sync(FULL, wc);
            const uint32_t sd = __reduce_add_sync(FULL, wd), se = __reduce_add_sync(FULL, we), sf = __reduce_add_sync(FULL, wf);
            uint32_t add = 0;
            switch (lane) {
                case AQC_C_GOOD_READS: add = sa & 0xFFu; break;
                case AQC_C_GOOD_BASES_R1: add = sc & 0xFFFFu; break;
                case AQC_C_GOOD_BASES_R2: add = sc >> 16; break;
                case AQC_C_BADTRIM1: add = se & 0xFFu; break;
                case AQC_C_BADTRIM2: add = (se >> 8) & 0xFFu; break;
                case AQC_C_BADLEN: add = (se >> 16) & 0xFFu; break;
                case AQC_C_BADPOL: add = se >> 24; break;
                case AQC_C_BADLQC: add = sf & 0xFFu; break;
                case AQC_C_BADNCT: add = (sf >> 8) & 0xFFu; break;
                case AQC_C_BADDIFF: add = (sf >> 16) & 0xFFu; break;
                case AQC_C_BADMISMATCH: add = sf >> 24; break;
                case AQC_C_READ_CORRECTED: add = (sa >> 16) & 0xFFu; break;
                case AQC_C_BASE_CORRECTED: add = sb & 0xFFu; break;
                case AQC_C_BASE_SKIPPED_CORRECTION: add = 2u * ((sb >> 8) & 0xFFu); break;
                case AQC_C_BASE_ZERO_QUAL_MASKED: add = 2u * ((sb >> 16) & 0xFFu); break;
                case AQC_C_OVERLAPPED: add = (sa >> 8) & 0xFFu; break;
                case AQC_C_OVERLAP_LEN_SUM: add = sd & 0xFFFFu; break;
                case AQC_C_OVERLAP_BASE_SUM: add = 2u * (sd & 0xFFFFu); break;
                case AQC_C_OVERLAP_BASE_ERR: add = sb >> 24; break;
                case AQC_C_TRIMMED_ADAPTER_BASE: add = 2u * (sd >> 16); break;
                case AQC_C_TRIMMED_ADAPTER_READ: add = sa >> 24; break;
                default: break;
            }
            wc0 += add;
        }

        // ---- postfilter statistics of the sampled good pairs (:624-627): the warp rebuilds the trimmed, corrected reads
        //      in its scratch from the record and runs statRead on them ----
        if constexpr (SMODE != 2) {
            const bool want = valid && cls == AQC_GOOD && (A.p.qc_sample <= 0 || gidx + 1 < (uint64_t)A.p.qc_sample);
            uint32_t sbm = __ballot_sync(FULL, want);
            if (__builtin_expect(sbm != 0u, 0)) {
                stat_since_flush += (uint32_t)__popc(sbm);
                uint8_t *sc_s1 = scratch, *sc_q1 = scratch + MAXB, *sc_s2 = scratch + 2 * MAXB, *sc_q2 = scratch + 3 * MAXB;
                uint32_t need[2] = {sbm, sbm};                    // per mate: lanes whose read still needs stat_read
                if constexpr (SMODE == 1) {                       // one lane per read for everything made of A,C,G,T,N (aqc_stat2.cuh)
#pragma unroll 1
                    for (int m = 0; m < (paired ? 2 : 1); m++) {
                        MatePatches mp;
                        mate_patches(edits, n_edits, m, start1, start2, mp);
                        const uint8_t *gs = m ? A.seq2 + a2 + start2 : A.seq1 + a1 + start1;
                        const uint8_t *gq = m ? A.qual2 + a2 + start2 : A.qual1 + a1 + start1;
                        const bool done = stat_tile<NW>(want, gs, gq, m ? len2 : len1, gidx, m, mp, qsm, A.qc[m], A.p.qc_kmer, lane, A.error_flag);
                        need[m] = __ballot_sync(FULL, want && !done);
                    }
                    sbm = need[0] | (paired ? need[1] : 0u);
                }
                while (sbm) {
                    const int src = __ffs(sbm) - 1;
                    sbm &= sbm - 1;
                    const uint32_t ba1 = __shfl_sync(FULL, a1, src), ba2 = __shfl_sync(FULL, a2, src);
                    const int bs1 = __shfl_sync(FULL, start1, src), bs2 = __shfl_sync(FULL, start2, src);
                    const int bl1 = __shfl_sync(FULL, len1, src), bl2 = __shfl_sync(FULL, len2, src);
                    const int bne = __shfl_sync(FULL, n_edits, src);
                    const uint32_t be0 = __shfl_sync(FULL, edits[0], src), be1 = __shfl_sync(FULL, edits[1], src);
                    const uint32_t be2 = __shfl_sync(FULL, edits[2], src), be3 = __shfl_sync(FULL, edits[3], src);
                    const uint64_t bg = A.first_index + t * 32u + (uint32_t)src;
                    __syncwarp();
                    for (int x = lane; x < bl1; x += 32) { sc_s1[x] = A.seq1[ba1 + bs1 + x]; sc_q1[x] = A.qual1[ba1 + bs1 + x]; }
                    if (paired)
                        for (int x = lane; x < bl2; x += 32) { sc_s2[x] = A.seq2[ba2 + bs2 + x]; sc_q2[x] = A.qual2[ba2 + bs2 + x]; }
                    __syncwarp();
                    if (lane == 0) {
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            if (k >= bne) break;
                            const uint32_t e = k == 0 ? be0 : (k == 1 ? be1 : (k == 2 ? be2 : be3));
                            const int kind = (int)AQC_EDIT_KIND(e), pos = (int)AQC_EDIT_POS(e);
                            if (kind == 0) { sc_s1[pos - bs1] = (uint8_t)AQC_EDIT_BASE(e); sc_q1[pos - bs1] = (uint8_t)AQC_EDIT_QUAL(e); }
                            else if (kind == 1) { sc_s2[pos - bs2] = (uint8_t)AQC_EDIT_BASE(e); sc_q2[pos - bs2] = (uint8_t)AQC_EDIT_QUAL(e); }
                            else if (kind == 2) { sc_q1[pos - bs1] = '!'; sc_q2[(int)AQC_EDIT_POS2(e) - bs2] = '!'; }
                        }
                    }
                    __syncwarp();
#pragma unroll 1
                    for (int m = 0; m < (paired ? 2 : 1); m++) {
                        if constexpr (SMODE == 1) { if (!((need[m] >> src) & 1u)) continue; }
                        stat_read(m ? sc_s2 : sc_s1, m ? sc_q2 : sc_q1, m ? bl2 : bl1, m, bg, qsm, A.qc[m], lut1, lut2, lut3, A.p.qc_kmer, lane, A.error_flag);
                    }
                }
                __syncwarp();
                if (stat_since_flush + 32u > flush_limit) {      // packed shared accumulators: count field is 12 bits
                    for (int m = 0; m < 2; m++) {
                        const QcDev &qd = A.qc[m];
                        if (!qd.valid) continue;
                        for (int i = lane; i < QC_CLASSES * A.max_len; i += 32) {
                            const uint32_t v = atomicExch(&s_acc[m * QC_CLASSES * A.max_len + i], 0u);
                            if (v) {
                                const int c = i / A.max_len, pos = i - c * A.max_len;
                                atomicAdd(&qd.cls_cnt[c * AQC_MAX_LEN + pos], (unsigned long long)(v >> 20));
                                atomicAdd(&qd.cls_qsum[c * AQC_MAX_LEN + pos], (unsigned long long)(v & 0xFFFFFu));
                            }
                        }
                        for (int i = lane; i < A.max_len; i += 32) {
                            const uint32_t v = atomicExch(&s_disc[m * A.max_len + i], 0u);
                            if (v) atomicAdd(&qd.disc[i], (unsigned long long)v);
                        }
                    }
                    stat_since_flush = 0;
                }
            }
        }

        a1 = na1; e1 = ne1; a2 = na2; e2 = ne2;
    }

    // ---- epilogue: flush everything this CTA accumulated ----
    __syncthreads();
    for (int m = 0; m < 2; m++) {
        const QcDev &qd = A.qc[m];
        if (!qd.valid) continue;
        for (int i = tid; i < QC_CLASSES * A.max_len; i += blockDim.x) {
            const uint32_t v = s_acc[m * QC_CLASSES * A.max_len + i];
            if (v) {
                const int c = i / A.max_len, pos = i - c * A.max_len;
                atomicAdd(&qd.cls_cnt[c * AQC_MAX_LEN + pos], (unsigned long long)(v >> 20));
                atomicAdd(&qd.cls_qsum[c * AQC_MAX_LEN + pos], (unsigned long long)(v & 0xFFFFFu));
            }
        }
        for (int i = tid; i < A.max_len; i += blockDim.x) {
            const uint32_t v = s_disc[m * A.max_len + i];
            if (v) atomicAdd(&qd.disc[i], (unsigned long long)v);
        }
    }
    for (int i = tid; i <= A.max_len; i += blockDim.x) {
        uint32_t v = s_ovh[i]; if (v) atomicAdd(&A.counters[AQC_C_OVERLAP_HIST + i], (unsigned long long)v);
        v = s_dih[i]; if (v) atomicAdd(&A.counters[AQC_C_DISTANCE_HIST + i], (unsigned long long)v);
    }
    if (tid < 16 && s_em[tid]) atomicAdd(&A.counters[AQC_C_ERR_MATRIX + tid], (unsigned long long)s_em[tid]);
    if (wc0) atomicAdd(&A.counters[lane], wc0);
    if (blockIdx.x == 0 && tid == 0) {          // TOTAL_READS / TOTAL_BASES (:416,:431,:433) are sums over the batch
        atomicAdd(&A.counters[AQC_C_TOTAL_READS], (unsigned long long)A.n);
        atomicAdd(&A.counters[AQC_C_TOTAL_BASES_R1], (unsigned long long)(A.off1[A.n] - A.off1[0]));
        if (paired) atomicAdd(&A.counters[AQC_C_TOTAL_BASES_R2], (unsigned long long)(A.off2[A.n] - A.off2[0]));
    }
}

}  // namespace aqc
