// aqc_edit.cuh -- Levenshtein distance on the GPU, one LANE per string pair, bit-parallel (Myers 1999 as blocked by Hyyro
// 2003): the column of the dynamic-programming matrix lives in the bits of 64-bit words as vertical +1 / -1 deltas, a text
// character costs ~15 word operations per 64 pattern characters instead of 64 cell updates.
//
// Replaces, for batches, the reference's only other native-backed per-read operator: util.editDistance (util.py:65-83) ->
// editdistance/_editdistance.cpp:100-126 (edit_distance), used by the barcode path (barcodeprocesser.py:60-75, 19-base windows).
// Any byte is a valid character: A,C,G,T,N have precomputed match masks, every other byte gets its mask by a scan of the
// pattern.  Patterns up to 64 characters take the single-word form; longer ones the blocked form (AQC_MAX_LEN = 1000 -> 16 blocks).
#pragma once
#include "aqc_device.cuh"

namespace aqc {

constexpr int EDIT_MAX_BLOCKS = (AQC_MAX_LEN + 63) / 64;

__device__ __forceinline__ int edit_class(uint32_t c) {
    switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; case 'N': return 4; default: return -1; }
}

// match mask of character c against pattern p[lo .. lo+len) (bit i = p[lo+i] == c)
__device__ __forceinline__ unsigned long long edit_eq_scan(const uint8_t *p, int lo, int len, uint32_t c) {
    unsigned long long m = 0;
    for (int i = 0; i < len; i++) m |= (unsigned long long)(p[lo + i] == c) << i;
    return m;
}

// one block, one text character; hin / return value = horizontal delta entering at the top / leaving at bit `top`
__device__ __forceinline__ int edit_advance(unsigned long long &Pv, unsigned long long &Mv, unsigned long long Eq, int hin, unsigned long long top) {
    if (hin < 0) Eq |= 1ULL;
    const unsigned long long Xv = Eq | Mv;
    const unsigned long long Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
    unsigned long long Ph = Mv | ~(Xh | Pv);
    unsigned long long Mh = Pv & Xh;
    const int hout = (Ph & top) ? 1 : ((Mh & top) ? -1 : 0);
    Ph <<= 1; Mh <<= 1;
    if (hin < 0) Mh |= 1ULL; else if (hin > 0) Ph |= 1ULL;
    Pv = Mh | ~(Xv | Ph);
    Mv = Ph & Xv;
    return hout;
}

// Levenshtein distance of a[0..la) and b[0..lb); -1 when the shorter string is longer than AQC_MAX_LEN
__device__ int edit_distance_lane(const uint8_t *a, int la, const uint8_t *b, int lb) {
    if (la == 0) return lb;
    if (lb == 0) return la;
    const uint8_t *p = a, *t = b;                     // pattern = the shorter string (fewer blocks)
    int m = la, n = lb;
    if (lb < la) { p = b; t = a; m = lb; n = la; }
    if (m <= 64) {
        unsigned long long peq[5] = {0, 0, 0, 0, 0};
        for (int i = 0; i < m; i++) {
            const int k = edit_class(p[i]);
#pragma unroll
            for (int q = 0; q < 5; q++) if (k == q) peq[q] |= 1ULL << i;
        }
        unsigned long long Pv = m == 64 ? ~0ULL : ((1ULL << m) - 1ULL), Mv = 0;
        const unsigned long long top = 1ULL << (m - 1);
        int score = m;
        for (int j = 0; j < n; j++) {
            const uint32_t c = t[j];
            const int k = edit_class(c);
            unsigned long long Eq = 0;
            if (k >= 0) {
#pragma unroll
                for (int q = 0; q < 5; q++) if (k == q) Eq = peq[q];
            } else Eq = edit_eq_scan(p, 0, m, c);
            score += edit_advance(Pv, Mv, Eq, 1, top);      // the top row of the matrix grows by one per text character
        }
        return score;
    }
    const int nb = (m + 63) >> 6;
    if (nb > EDIT_MAX_BLOCKS) return -1;
    unsigned long long Pv[EDIT_MAX_BLOCKS], Mv[EDIT_MAX_BLOCKS], peq[5][EDIT_MAX_BLOCKS];
    for (int r = 0; r < nb; r++) {
        const int len = min(64, m - 64 * r);
        Pv[r] = len == 64 ? ~0ULL : ((1ULL << len) - 1ULL);
        Mv[r] = 0;
        for (int q = 0; q < 5; q++) peq[q][r] = 0;
        for (int i = 0; i < len; i++) {
            const int k = edit_class(p[64 * r + i]);
            if (k >= 0) peq[k][r] |= 1ULL << i;
        }
    }
    const unsigned long long last_top = 1ULL << ((m - 1) & 63);
    int score = m;
    for (int j = 0; j < n; j++) {
        const uint32_t c = t[j];
        const int k = edit_class(c);
        int h = 1;
        for (int r = 0; r < nb; r++) {
            const unsigned long long Eq = k >= 0 ? peq[k][r] : edit_eq_scan(p, 64 * r, min(64, m - 64 * r), c);
            h = edit_advance(Pv[r], Mv[r], Eq, h, r == nb - 1 ? last_top : (1ULL << 63));
        }
        score += h;
    }
    return score;
}

// out[i] = distance(a[a_off[i] .. a_off[i+1]), b[b_off[i] .. b_off[i+1]))
__global__ void edit_distance_kernel(const uint8_t *a, const uint32_t *a_off, const uint8_t *b, const uint32_t *b_off, uint32_t n, int32_t *out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t a0 = a_off[i], b0 = b_off[i];
        out[i] = edit_distance_lane(a + a0, (int)(a_off[i + 1] - a0), b + b0, (int)(b_off[i + 1] - b0));
    }
}

}  // namespace aqc
