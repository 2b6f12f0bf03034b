// aqc_lane_kernel.cuh -- the filter path for short reads (every mate <= 32*NW bases, NW <= 8): ONE LANE PER PAIR.
//
// pair_kernel (aqc_kernel.cuh) gives a whole warp to one pair: simple and length-agnostic, but a 150-base mate fills
// only 5 of the 32 lanes of a plane set and every piece of per-pair bookkeeping is executed 32 times.  For Illumina
// lengths the planes of both mates fit in registers (2 x NW words per plane), so here a warp takes a tile of 32 pairs
// and every lane runs the whole reference loop body (preprocesser.py:455-631) for its own pair:
//   * the warp's private shared-memory stage has TWO columns: bases 1 and bases 2 arrive first (1-D TMA bulk copies);
//     as soon as the lane has converted mate 1 to bit-planes (SWAR, 4 bases per 32-bit word: codes from the ASCII bits,
//     re-encoding check for foreign bytes, multiply-gather of the code bits) the qualities of mate 1 are copied over
//     its bases and that copy overlaps the conversion of mate 2.  Qualities of mate 2 are only ever touched by the
//     correction walk (two bytes per visited mismatch) and are never staged.  After the low-quality count the stage is
//     free and the next tile's copy overlaps the rest of the work;
//   * util.overlap_hm (util.py:158-212): per 32 candidate offsets the lane funnel-shifts one code-bit plane, XORs it with
//     the fixed mate's first word and keeps the offsets with < 3 differing bits in the first 32 positions as a bit mask
//     (a necessary condition of the acceptance rule); the lanes then evaluate their candidates exactly, in scan order,
//     in lock step; the scanned mate's words are rotated one register per round so that the code does not depend on
//     the round;
//   * hasPolyX is screened by a multi-word run-length test per lane; adapter cut, rescan, correction walk and the
//     classifier are per-lane code on registers, reading the few bytes the walk needs from HBM (L2);
//   * work that suits a whole warp is handed over by ballot: exact hasPolyX of screened reads, and pairs holding a byte
//     outside A,C,G,T,N, which are appended to a list that pair_kernel processes in its list mode right after this kernel;
//   * the kernel carries NO statistics code: the sampled good pairs are stat'd from their 32-byte records by
//     stat_kernel<.., POST> (aqc_stat_kernel.cuh) in the next launch -- instruction-cache misses were the first stall
//     reason of the previous generation (profiles/r02_lane2_kernel_ncu_full.txt), the fused statRead hand-over a
//     seventh of its instructions.
// Counters: per-tile packed warp reductions into lane-owned 64-bit registers; histograms and the error matrix are
// shared-memory atomics.  Results are bit-identical to pair_kernel and the oracle.
#pragma once
#include "aqc_device.cuh"

namespace aqc {

// ONE CTA per SM, a multiple of four warps (one scheduler each): 16 warps for paired reads <= 160 bases (120 registers, no
// spills), 12 for paired reads <= 256 bases (161 registers), 24 for single-end reads (72 registers).  Measured on PE150,
// filter launch for 10 M pairs (profiles/r02_notes.md): five 4-warp CTAs 3.75 ms; one CTA of 16 / 18 / 20 / 22 warps
// 3.24 / 3.66 / 3.32 / 3.70 ms.
#ifndef AQC_LANE_WARPS_SHORT
#define AQC_LANE_WARPS_SHORT 16
#endif
#ifndef AQC_LANE_WARPS_LONG
#define AQC_LANE_WARPS_LONG 12
#endif
#ifndef AQC_LANE_WARPS_SINGLE
#define AQC_LANE_WARPS_SINGLE 24
#endif
constexpr int LANE_MAX_WARPS = 24;
__host__ __device__ constexpr int lane_max_warps(int nw, bool paired) {
    return !paired ? AQC_LANE_WARPS_SINGLE : (nw > 5 ? AQC_LANE_WARPS_LONG : AQC_LANE_WARPS_SHORT);
}
static_assert(AQC_LANE_WARPS_SHORT <= LANE_MAX_WARPS && AQC_LANE_WARPS_LONG <= LANE_MAX_WARPS && AQC_LANE_WARPS_SINGLE <= LANE_MAX_WARPS, "barrier array");

struct LArgs {
    KArgs k;                      // batch, parameters, outputs (tile_pairs/col_cap as used by this kernel)
    uint32_t *fb_list;            // pairs that need the general (warp-per-pair) path
    uint32_t *fb_count;
    int lane_col_cap;             // bytes reserved per column in a warp's stage
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
#pragma unroll 1
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
            uint32_t rlo = 0, rhi = 0, nlo = 0, nhi = 0, bad = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                // bits 1..3 of the ASCII byte: A 000, C 001, T 010, G 011, N 111 -- 2-bit code + "is N"
                const uint32_t tt = (v[j] & 0x0E0E0E0Eu) >> 1;
                // codes -> ASCII: the four 3-bit values become the selector nibbles of an 8-entry table lookup
                // (entries 4..6 hold 0, which no byte with those bits re-encodes to)
                const uint32_t e = prmt_raw(0x47544341u, 0x4E000000u, prmt_raw(tt + (tt >> 4), 0u, 0x4420u));
                bad |= e ^ v[j];                                                                // non-zero: a byte outside A,C,G,T,N
                const uint32_t z = (tt + (tt << 3)) & 0x11111111u;                              // bit0 = code bit 0, bit4 = code bit 1
                const uint32_t r = z * 0x01020408u;                     // byte 3 = nibble of plane 0 | nibble of plane 1 << 4
                const uint32_t rn = (tt & 0x04040404u) * 0x00408102u;   // byte 3 = nibble of the N plane
                constexpr uint32_t sel[4] = {0x3217u, 0x3270u, 0x3710u, 0x7210u};               // byte 3 of r -> byte j of the accumulator
                if (j < 4) { rlo = __byte_perm(rlo, r, sel[j & 3]); nlo = __byte_perm(nlo, rn, sel[j & 3]); }
                else { rhi = __byte_perm(rhi, r, sel[j & 3]); nhi = __byte_perm(nhi, rn, sel[j & 3]); }
            }
            // de-interleave the nibbles: bytes of rlo/rhi hold (p1 nibble << 4 | p0 nibble) of 4 bases each
            {
                const uint32_t l0 = rlo & 0x0F0F0F0Fu, h0 = rhi & 0x0F0F0F0Fu;
                const uint32_t l1 = (rlo >> 4) & 0x0F0F0F0Fu, h1 = (rhi >> 4) & 0x0F0F0F0Fu;
                const uint32_t a0 = (l0 | (l0 >> 4)) & 0x00FF00FFu, b0 = (h0 | (h0 >> 4)) & 0x00FF00FFu;
                const uint32_t a1 = (l1 | (l1 >> 4)) & 0x00FF00FFu, b1 = (h1 | (h1 >> 4)) & 0x00FF00FFu;
                const uint32_t an = (nlo | (nlo >> 4)) & 0x00FF00FFu, bn = (nhi | (nhi >> 4)) & 0x00FF00FFu;
                p0 = __byte_perm(a0, b0, 0x6420);
                p1 = __byte_perm(a1, b1, 0x6420);
                pn = __byte_perm(an, bn, 0x6420);
            }
            const uint32_t vm = lowmask(nvalid);
            pn &= vm;
            p0 &= vm & ~pn; p1 &= vm & ~pn;                     // 'N' keeps its own plane, its code bits are cleared
            n_count += __popc(pn);
            if (__builtin_expect(bad != 0u, 0)) {               // a byte outside A,C,G,T,N among the 32 (maybe beyond the read)
                uint32_t xb = 0;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const uint32_t tt = (v[j] & 0x0E0E0E0Eu) >> 1;
                    const uint32_t e = __byte_perm(0x47544341u, 0x4E000000u, __byte_perm(tt | (tt >> 4), 0u, 0x4420));
                    xb |= gather4(hibit_nonzero(e ^ v[j]) >> 7) << (4 * j);
                }
                if (xb & vm) exotic = true;
            }
        }
#pragma unroll
        for (int i = 0; i < NW; i++) {
            P.p0[i] = (i + 1 < NW) ? P.p0[i + 1 < NW ? i + 1 : 0] : p0;
            P.p1[i] = (i + 1 < NW) ? P.p1[i + 1 < NW ? i + 1 : 0] : p1;
            P.pn[i] = (i + 1 < NW) ? P.pn[i + 1 < NW ? i + 1 : 0] : pn;
        }
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
    bool found = false;
#pragma unroll 1
    for (int r = 0; r < rounds; r++) {
        const int rem = nOff - (r << 5);
        if (!__any_sync(FULL, !found && rem > 0)) break;
        uint32_t cm = 0;
        if (!found && rem > 0) {
            if (!slow) {
                // one code bit is enough for a necessary condition: equal bases have equal bits, and 32 random positions
                // differ in fewer than 3 of them with probability 1.2e-7.  The verdicts are shifted into the mask through the
                // sign of popc - 3 (offset 31 first, so offset b ends up in bit b).
#pragma unroll
                for (int b = 31; b >= 0; b--) {
                    const uint32_t x = __funnelshift_r(S0[0], S0[1], b) ^ f0;
                    cm = __funnelshift_l((uint32_t)(__popc(x) - 3), cm, 1);
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

// dynamic shared memory of one CTA:
//   [nwarps][ 2 * lane_col_cap ]     per-warp stage: column 0 = bases 1, later qualities 1 | column 1 = bases 2 (or qualities 1, single-end)
//   luts (768 B) | overlap_hist [max_len+1] | distance_hist [max_len+1] | err matrix [16]
__host__ __device__ __forceinline__ size_t lane_smem_bytes(int nwarps, int col_cap, int max_len) {
    return (size_t)nwarps * 2 * (size_t)col_cap + 768 + (size_t)(2 * (max_len + 1) + 16) * 4 + 64;
}

template <bool PAIRED, int NW>
__global__ void __launch_bounds__(lane_max_warps(NW, PAIRED) * 32, 1) lane_kernel(const __grid_constant__ LArgs L) {
    AQC_DYN_SMEM(smem_raw);
    __shared__ __align__(8) uint64_t full_bar[2 * LANE_MAX_WARPS];      // per warp: [0] bases landed, [1] qualities landed
    const KArgs &A = L.k;
    constexpr bool paired = PAIRED;
    constexpr int MAXB = 32 * NW;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = blockDim.x >> 5;
    const int col_cap = L.lane_col_cap;

    uint8_t *stage = smem_raw + (size_t)warp * 2 * col_cap;
    uint8_t *lutbase = smem_raw + (size_t)nwarps * 2 * col_cap;
    const uint8_t *lut2 = lutbase + 256, *lut3 = lutbase + 512;
    uint32_t *s_ovh = reinterpret_cast<uint32_t *>(lutbase + 768);
    uint32_t *s_dih = s_ovh + (A.max_len + 1);
    uint32_t *s_em = s_dih + (A.max_len + 1);
    const int n_acc_words = 2 * (A.max_len + 1) + 16;

    for (int i = tid; i < 768; i += blockDim.x) lutbase[i] = reinterpret_cast<const uint8_t *>(A.luts)[i];
    for (int i = tid; i < n_acc_words; i += blockDim.x) s_ovh[i] = 0;
    if (tid == 0) {
        for (int s = 0; s < 2 * nwarps; s++) mbar_init(&full_bar[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    // lane i owns scalar counter i
    unsigned long long wc0 = 0;

    uint64_t *bar = &full_bar[2 * warp], *qbar = &full_bar[2 * warp + 1];
    const uint32_t gw = blockIdx.x * (uint32_t)nwarps + (uint32_t)warp;
    const uint32_t W = gridDim.x * (uint32_t)nwarps;

    // offsets of the lane's pair in tile t (pairs beyond n: empty records at the end of the columns)
    auto load_offsets = [&](uint32_t t, uint32_t &a1, uint32_t &e1, uint32_t &a2, uint32_t &e2) {
        const uint32_t pp = min(t * 32u + (uint32_t)lane, A.n), pq = min(pp + 1u, A.n);
        a1 = A.off1[pp]; e1 = A.off1[pq];
        a2 = 0; e2 = 0;
        if (paired) { a2 = A.off2[pp]; e2 = A.off2[pq]; }
    };
    // producer (whole warp computes, lane 0 issues).  Paired input: the bases of both mates first; the qualities of mate 1
    // follow into column 0 once mate 1 has been converted (issue_quals).  Single-end input: bases and qualities together.
    auto issue_tile = [&](uint32_t a1, uint32_t e1, uint32_t a2, uint32_t e2) {
        const uint32_t f1 = __shfl_sync(FULL, a1, 0), l1 = __shfl_sync(FULL, e1, 31);
        const uint32_t f2 = __shfl_sync(FULL, a2, 0), l2 = __shfl_sync(FULL, e2, 31);
        if (lane == 0) {
            const uint32_t g1 = f1 & ~15u, bytes1 = (l1 - g1 + 15u) & ~15u;
            const uint32_t g2 = f2 & ~15u, bytes2 = paired ? ((l2 - g2 + 15u) & ~15u) : 0u;
            mbar_expect_tx(bar, paired ? bytes1 + bytes2 : 2 * bytes1);
            if (bytes1) {
                bulk_g2s(stage, A.seq1 + g1, bytes1, bar);
                if (!paired) bulk_g2s(stage + col_cap, A.qual1 + g1, bytes1, bar);
            }
            if (paired && bytes2) bulk_g2s(stage + col_cap, A.seq2 + g2, bytes2, bar);
        }
    };
    auto issue_quals = [&](uint32_t a1, uint32_t e1) {          // paired only: qualities of mate 1 over the bases of mate 1
        const uint32_t f1 = __shfl_sync(FULL, a1, 0), l1 = __shfl_sync(FULL, e1, 31);
        if (lane == 0) {
            const uint32_t g1 = f1 & ~15u, bytes1 = (l1 - g1 + 15u) & ~15u;
            mbar_expect_tx(qbar, bytes1);
            if (bytes1) bulk_g2s(stage, A.qual1 + g1, bytes1, qbar);
        }
    };

    uint32_t a1 = 0, e1 = 0, a2 = 0, e2 = 0;
    uint32_t t = gw;
    if (t < A.num_tiles) {
        load_offsets(t, a1, e1, a2, e2);
        issue_tile(a1, e1, a2, e2);
    }
    uint32_t parity = 0, qparity = 0;

#pragma unroll 1
    for (; t < A.num_tiles; t += W) {
        const uint32_t tn = t + W;
        uint32_t na1 = 0, ne1 = 0, na2 = 0, ne2 = 0;
        if (tn < A.num_tiles) load_offsets(tn, na1, ne1, na2, ne2);

        mbar_wait(bar, parity);
        parity ^= 1u;

        const uint32_t pp = t * 32u + (uint32_t)lane;
        const bool valid = pp < A.n;
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
        const bool want_lowq = A.p.unqualified_base_limit > 0;     // warp-uniform
        {
            bool ex = false;
#pragma unroll 1
            for (int m = 0; m < (paired ? 2 : 1); m++) {              // one copy of the conversion + screen code for both mates
                if (live) {
                    const uint8_t *r = m ? stage + col_cap + (a2 - g2) + start2 : stage + (a1 - g1) + start1;
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
                if (paired && m == 0 && want_lowq) {     // the bases of mate 1 are in registers: their column takes the qualities
                    fence_proxy_async();
                    __syncwarp();
                    issue_quals(a1, e1);
                }
            }
            if (want_lowq) {
                if (paired) { mbar_wait(qbar, qparity); qparity ^= 1u; }
                if (live) lowq1 = lane_lowq(stage + (paired ? 0 : col_cap) + (a1 - g1) + start1, len1, A.p.qualified_quality_phred + 33);
            }
            if (live && ex) { fallback = true; live = false; cls = AQC_NUM_CLASSES; }
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
                        } else if (sh == -offset) {
                            // The rescan (:534) starts with forward offset 0 on the cut mates: r1[i] against rc(r2[0:ol])[i], i < ol.
                            // When the cut dropped exactly the -offset leading bases of rc(r2) (always for mates of equal
                            // length) that is the alignment just accepted -- same positions, same mismatches, same rule --
                            // so the rescan returns (0, ol, distance) without being run.
                            offset = 0;
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
                    // the first `distance` (<= 3) mismatches of the walk; their four bytes each are fetched together so that
                    // the walk waits for HBM (or, with AQC_BATCH_QUAL2_IN_PLACE, for PCIe) once, not once per mismatch
                    int wo[3] = {-1, -1, -1};
                    uint32_t wb[3] = {0, 0, 0};          // b1 | r2 byte << 8 | qa << 16 | qb << 24
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        if (k < distance) {
                            int o = -1;
#pragma unroll
                            for (int i = 0; i < NW; i++) {
                                if (o < 0 && xx[i]) { o = 32 * i + __ffs(xx[i]) - 1; xx[i] &= xx[i] - 1; }
                            }
                            wo[k] = o;
                            if (o >= 0) {
                                const int p1 = len1 - ol + o, p2 = len2 - 1 - o;
                                wb[k] = (uint32_t)G1[p1] | ((uint32_t)G2[p2] << 8) | ((uint32_t)G1q[p1] << 16) | ((uint32_t)G2q[p2] << 24);
                            }
                        }
                    }
                    int corrected = 0, masked = 0, skipped = 0;
                    int em_cell[3] = {-1, -1, -1};
                    int done = 0;
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        if (k < distance && wo[k] >= 0) {
                            const int o = wo[k];
                            const int p1 = len1 - ol + o, p2 = len2 - 1 - o;
                            const uint8_t b1 = (uint8_t)wb[k];                            // :564
                            const uint8_t b2 = lut3[(wb[k] >> 8) & 0xFFu];                // :565 util.complement
                            const uint8_t qa = (uint8_t)(wb[k] >> 16), qb = (uint8_t)(wb[k] >> 24);   // :566-567
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
                            edits[k] = e; em_cell[k] = cell;
                            done++;
                        }
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

        a1 = na1; e1 = ne1; a2 = na2; e2 = ne2;
    }

    // ---- epilogue: flush everything this CTA accumulated ----
    __syncthreads();
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
