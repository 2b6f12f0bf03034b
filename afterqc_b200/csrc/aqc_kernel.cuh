// aqc_kernel.cuh -- the persistent tile kernel (filter / stat / ops modes).
#pragma once
#include "aqc_device.cuh"

namespace aqc {

// python slice semantics of trim() (preprocesser.py:19-28): s[front:-tail] or s[front:]
__device__ __forceinline__ void py_trim(int len, int front, int tail, int &start, int &newlen) {
    int s = front < len ? front : len;
    int e = tail > 0 ? len - tail : len;
    if (e < 0) e = 0;
    if (e < s) e = s;
    start = s; newlen = e - s;
}


// Planes of both mates for the comparisons: fast ASCII-bit codes when every byte is A,C,G,T,N, else the
// general LUT codes (np = 4).  Also N counts and the hasPolyX screen verdicts.
struct PairPlanes {
    uint32_t P1[4], RC[4];
    int np, n1, n2;
    bool cand1, cand2;      // hasPolyX candidates (screen passed)
};

__device__ __forceinline__ void prepare_pair(const uint8_t *r1, int len1, const uint8_t *r2, int len2, bool paired, bool want_poly,
                                             int maxPoly, int poly_m, const uint8_t *lut1, const uint4 *bytemask16, int lane, PairPlanes &pp) {
    FastPlanes F1, F2;
    fast_build2(r1, len1, r2, paired ? len2 : 0, lane, bytemask16, F1, F2);
    const bool exotic = F1.exotic || F2.exotic;
    pp.cand1 = pp.cand2 = false;
    if (__builtin_expect(!exotic, 1)) {
        pp.np = (F1.hasN || F2.hasN) ? 3 : 2;
#pragma unroll
        for (int k = 0; k < 4; k++) { pp.P1[k] = F1.P[k]; pp.RC[k] = 0; }
        pp.n1 = F1.n_count; pp.n2 = F2.n_count;
        if (want_poly && poly_m >= 0) {
            if (paired && poly_m >= 1 && poly_m <= 31 && len1 <= 512 && len2 <= 512) {
                const uint32_t c = polyx_screen_pair(F1.P, F2.P, pp.np, len1, len2, maxPoly, poly_m, lane);
                pp.cand1 = (c & 1u) != 0; pp.cand2 = (c & 2u) != 0;
            } else {
                pp.cand1 = polyx_screen_fast(F1.P, pp.np, len1, maxPoly, poly_m, lane);
                if (paired) pp.cand2 = polyx_screen_fast(F2.P, pp.np, len2, maxPoly, poly_m, lane);
            }
        }
        if (paired) fast_revcomp(F2, len2, lane, pp.RC);
    } else {                                     // some byte outside A,C,G,T,N: LUT codes, 4 planes (rare)
        pp.np = 4;
        pp.n1 = 0; pp.n2 = 0;
#pragma unroll 1
        for (int m = 0; m < (paired ? 2 : 1); m++) {                // one inlined copy of the LUT build for both mates
            uint32_t Q[4]; int nn = 0; bool ee = false;
            build_planes(m ? r2 : r1, m ? len2 : len1, m != 0, lut1, lane, Q, nn, ee);
            const bool cand = want_poly && polyx_screen_fast(Q, 4, m ? len2 : len1, maxPoly, poly_m, lane);
            if (m) { pp.n2 = nn; pp.cand2 = cand; }
            else { pp.n1 = nn; pp.cand1 = cand; }
#pragma unroll
            for (int k = 0; k < 4; k++) { if (m) pp.RC[k] = Q[k]; else pp.P1[k] = Q[k]; }
        }
    }
}

__device__ __forceinline__ int lowq_any(const uint8_t *q, int len, int thr, int lane) {
    if (thr <= 0) return 0;
    return thr < 128 ? count_lowq_fast(q, len, thr, lane) : count_lowq(q, len, thr, lane);
}

struct StageBuf {
    uint8_t *col[4];        // seq1, qual1, seq2, qual2
    uint32_t *off1, *off2;  // tile_pairs + 4 entries each
};

// dynamic shared memory layout:
//   [NSTAGES][ 4 * col_cap + 2 * off_cap ]   tile staging (TMA destinations, 128-byte aligned)
//   luts (768 B) + byte-mask table (272 B)
//   qc acc  [2][5][max_len] u32, qc disc [2][max_len] u32
//   overlap_hist [max_len+1] u32, distance_hist [max_len+1] u32
#ifndef AQC_MIN_BLOCKS
#define AQC_MIN_BLOCKS 4
#endif
template <int MODE, bool PAIRED>
__global__ void __launch_bounds__(THREADS, AQC_MIN_BLOCKS) pair_kernel(const __grid_constant__ KArgs A) {
    AQC_DYN_SMEM(smem_raw);
    __shared__ __align__(8) uint64_t full_bar[NSTAGES];
    __shared__ uint32_t tile_base[NSTAGES][2];      // 16-byte aligned column origin of the staged tile
    __shared__ uint32_t qc_reads_since_flush;
    __shared__ uint32_t tile_next[NSTAGES];        // AQC_DYNAMIC_CLAIM: next unclaimed pair of the staged tile

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr bool paired = PAIRED;
    const int P = A.tile_pairs;
    const int off_cap = ((P + 8) * 4 + 15) & ~15;        // P+1 entries from a 16-byte aligned start (<= P+4) + slack
    const int stage_bytes = (4 * A.col_cap + 2 * off_cap + 127) & ~127;

    uint8_t *lutbase = smem_raw + (size_t)NSTAGES * stage_bytes;
    const uint8_t *lut1 = lutbase, *lut2 = lutbase + 256, *lut3 = lutbase + 512;
    uint4 *const bytemask16 = reinterpret_cast<uint4 *>(lutbase + 768);          // 17 entries
    uint32_t *s_acc = reinterpret_cast<uint32_t *>(lutbase + 768 + 272);
    uint32_t *s_disc = s_acc + 2 * QC_CLASSES * A.max_len;
    uint32_t *s_ovh = s_disc + 2 * A.max_len;
    uint32_t *s_dih = s_ovh + (A.max_len + 1);
    const int n_acc_words = 2 * QC_CLASSES * A.max_len + 2 * A.max_len + 2 * (A.max_len + 1);

    auto stage_ptr = [&](int st) -> StageBuf {
        StageBuf b;
        uint8_t *p = smem_raw + (size_t)st * stage_bytes;
        b.col[0] = p; b.col[1] = p + A.col_cap; b.col[2] = p + 2 * A.col_cap; b.col[3] = p + 3 * A.col_cap;
        b.off1 = reinterpret_cast<uint32_t *>(p + 4 * A.col_cap);
        b.off2 = reinterpret_cast<uint32_t *>(p + 4 * A.col_cap + off_cap);
        return b;
    };

    // ---- one-time setup ----
    for (int i = tid; i < 768; i += THREADS) lutbase[i] = reinterpret_cast<const uint8_t *>(A.luts)[i];
    for (int i = tid; i < 272; i += THREADS) lutbase[768 + i] = ((i & 15) < (i >> 4)) ? 0xFF : 0x00;   // entry k: first k bytes set
    for (int i = tid; i < n_acc_words; i += THREADS) s_acc[i] = 0;
    if (tid == 0) {
        for (int s = 0; s < NSTAGES; s++) mbar_init(&full_bar[s], 1);
        qc_reads_since_flush = 0;
        for (int s = 0; s < NSTAGES; s++) tile_next[s] = 0;
        fence_mbar_init();
    }
    __syncthreads();

    // list mode: one listed pair per tile, the number of tiles is read from device memory
    constexpr bool listed = (MODE == MODE_LIST);     // its own instantiation: the regular filter kernel carries none of this
    const uint32_t num_tiles = listed ? *A.list_count : A.num_tiles;
    auto tile_first = [&](uint32_t tile) -> uint32_t { return listed ? A.list[tile] : tile * (uint32_t)P; };
    auto tile_end = [&](uint32_t p0) -> uint32_t { return listed ? p0 + 1u : min(A.n, p0 + (uint32_t)P); };

    // producer: issue the bulk copies of one tile into a stage (single thread)
    auto issue_tile = [&](uint32_t tile, int st) {
        StageBuf b = stage_ptr(st);
        uint32_t p0 = tile_first(tile);
        uint32_t p1 = tile_end(p0);
        uint32_t a1 = A.off1[p0], e1 = A.off1[p1];
        uint32_t g1 = a1 & ~15u;
        uint32_t bytes1 = (e1 - g1 + 15u) & ~15u;
        uint32_t o0 = p0 & ~3u;                                   // offsets: 16-byte aligned start
        uint32_t obytes = ((p1 + 1 - o0) * 4u + 15u) & ~15u;
        uint32_t total = 2 * bytes1 + obytes;
        uint32_t a2 = 0, g2 = 0, bytes2 = 0;
        if (paired) {
            a2 = A.off2[p0]; uint32_t e2 = A.off2[p1];
            g2 = a2 & ~15u;
            bytes2 = (e2 - g2 + 15u) & ~15u;
            total += 2 * bytes2 + obytes;
        }
        tile_base[st][0] = g1; tile_base[st][1] = g2;
        mbar_expect_tx(&full_bar[st], total);
        if (bytes1) { bulk_g2s(b.col[0], A.seq1 + g1, bytes1, &full_bar[st]); bulk_g2s(b.col[1], A.qual1 + g1, bytes1, &full_bar[st]); }
        bulk_g2s(b.off1, A.off1 + o0, obytes, &full_bar[st]);
        if (paired) {
            if (bytes2) { bulk_g2s(b.col[2], A.seq2 + g2, bytes2, &full_bar[st]); bulk_g2s(b.col[3], A.qual2 + g2, bytes2, &full_bar[st]); }
            bulk_g2s(b.off2, A.off2 + o0, obytes, &full_bar[st]);
        }
        (void)a1; (void)a2;
    };

    // per-warp scalar counters: lane i owns counter i (wc0) / error-matrix cell i (wc1)
    unsigned long long wc0 = 0, wc1 = 0;
    auto bump = [&](int idx, unsigned long long v) { if (lane == idx) wc0 += v; };

    QcSmem qsm; qsm.acc = s_acc; qsm.disc = s_disc; qsm.max_len = A.max_len;

    auto flush_qc = [&]() {   // all threads; caller syncs before and after
        for (int m = 0; m < 2; m++) {
            const QcDev &qd = A.qc[m];
            if (!qd.valid) continue;
            for (int i = tid; i < QC_CLASSES * A.max_len; i += THREADS) {
                uint32_t v = s_acc[m * QC_CLASSES * A.max_len + i];
                if (v) {
                    int cls = i / A.max_len, pos = i - cls * A.max_len;
                    atomicAdd(&qd.cls_cnt[cls * AQC_MAX_LEN + pos], (unsigned long long)(v >> 20));
                    atomicAdd(&qd.cls_qsum[cls * AQC_MAX_LEN + pos], (unsigned long long)(v & 0xFFFFFu));
                    s_acc[m * QC_CLASSES * A.max_len + i] = 0;
                }
            }
            for (int i = tid; i < A.max_len; i += THREADS) {
                uint32_t v = s_disc[m * A.max_len + i];
                if (v) { atomicAdd(&qd.disc[i], (unsigned long long)v); s_disc[m * A.max_len + i] = 0; }
            }
        }
    };

    // list mode stages its single-pair tiles with plain loads (a rare path; and mate-2 qualities may live in page-locked
    // host memory there, AQC_BATCH_QUAL2_IN_PLACE, which bulk copies are not used on)
    auto stage_listed = [&](uint32_t tile, int st) {     // all threads; ends with a CTA barrier
        StageBuf b = stage_ptr(st);
        const uint32_t p0 = tile_first(tile), p1 = tile_end(p0);
        const uint32_t o0 = p0 & ~3u;
        const uint32_t a1 = A.off1[p0], e1 = A.off1[p1], g1 = a1 & ~15u;
        uint32_t a2 = 0, e2 = 0, g2 = 0;
        if (paired) { a2 = A.off2[p0]; e2 = A.off2[p1]; g2 = a2 & ~15u; }
        for (uint32_t i = a1 - g1 + tid; i < e1 - g1; i += THREADS) { b.col[0][i] = A.seq1[g1 + i]; b.col[1][i] = A.qual1[g1 + i]; }
        for (uint32_t i = tid; i <= p1 - o0; i += THREADS) b.off1[i] = A.off1[o0 + i];
        if (paired) {
            for (uint32_t i = a2 - g2 + tid; i < e2 - g2; i += THREADS) { b.col[2][i] = A.seq2[g2 + i]; b.col[3][i] = A.qual2[g2 + i]; }
            for (uint32_t i = tid; i <= p1 - o0; i += THREADS) b.off2[i] = A.off2[o0 + i];
        }
        if (tid == 0) { tile_base[st][0] = g1; tile_base[st][1] = g2; }
        __syncthreads();
    };

    // ---- prologue: prefetch the first NSTAGES tiles of this CTA ----
    if (tid == 0 && !listed) {
        for (int s = 0; s < NSTAGES; s++) {
            uint32_t t = blockIdx.x + (uint32_t)s * gridDim.x;
            if (t < num_tiles) issue_tile(t, s);
        }
    }

    uint32_t it = 0;
    for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, it++) {
        const int st = it % NSTAGES;
        const uint32_t parity = (it / NSTAGES) & 1u;
        if (listed) stage_listed(tile, st);
        else mbar_wait(&full_bar[st], parity);
        StageBuf sb = stage_ptr(st);
        const uint32_t p0 = tile_first(tile);
        const uint32_t p1 = tile_end(p0);
        const uint32_t o0 = p0 & ~3u;
        const uint32_t g1 = tile_base[st][0], g2 = tile_base[st][1];

#ifdef AQC_DYNAMIC_CLAIM
        for (;;) {
            uint32_t pp = 0;
            if (lane == 0) pp = p0 + atomicAdd(&tile_next[st], 1u);
            pp = __shfl_sync(FULL, pp, 0);
            if (pp >= p1) break;
#else
        for (uint32_t pp = p0 + warp; pp < p1; pp += WARPS) {
#endif
            const uint32_t a1 = sb.off1[pp - o0], e1 = sb.off1[pp + 1 - o0];
            uint8_t *S1 = sb.col[0] + (a1 - g1);
            uint8_t *Q1 = sb.col[1] + (a1 - g1);
            const int olen1 = (int)(e1 - a1);
            uint8_t *S2 = nullptr, *Q2 = nullptr;
            int olen2 = 0;
            if (paired) {
                const uint32_t a2 = sb.off2[pp - o0], e2 = sb.off2[pp + 1 - o0];
                S2 = sb.col[2] + (a2 - g2);
                Q2 = sb.col[3] + (a2 - g2);
                olen2 = (int)(e2 - a2);
            }
            const uint64_t gidx = A.first_index + pp;           // 0-based global record index
            if (olen1 > AQC_MAX_LEN || olen2 > AQC_MAX_LEN) {   // host pre-checks; defensive
                if (lane == 0) atomicExch(A.error_flag, AQC_ERR_TOO_LONG);
                continue;
            }

            // ================================ MODE_STAT ================================
            if (MODE == MODE_STAT) {
                if (gidx >= A.stat_lo && gidx < A.stat_hi) {
                    uint64_t order = A.order_base + (gidx - A.stat_lo);
                    if (A.qc[0].valid) stat_read(S1, Q1, olen1, 0, order, qsm, A.qc[0], lut1, lut2, lut3, A.p.qc_kmer, lane, A.error_flag);
                    if (paired && A.qc[1].valid) stat_read(S2, Q2, olen2, 1, order, qsm, A.qc[1], lut1, lut2, lut3, A.p.qc_kmer, lane, A.error_flag);
                }
                continue;
            }

            // ================================ trim ================================
            int start1 = 0, len1 = olen1, start2 = 0, len2 = olen2;
            int cls = AQC_GOOD;
            const bool do_trim = (A.p.trim_front > 0 || A.p.trim_tail > 0);   // gate keyed on R1 only (quirk Q4)

            if (MODE == MODE_OPS) {
                if (do_trim) {
                    py_trim(olen1, A.p.trim_front, A.p.trim_tail, start1, len1);
                    if (paired) py_trim(olen2, A.p.trim_front2, A.p.trim_tail2, start2, len2);
                }
                const uint8_t *r1 = S1 + start1, *r1q = Q1 + start1;
                const uint8_t *r2 = paired ? S2 + start2 : nullptr, *r2q = paired ? Q2 + start2 : nullptr;
                PairPlanes pl;
                prepare_pair(r1, len1, r2, len2, paired, true, A.p.poly_size_limit, A.poly_m, lut1, bytemask16, lane, pl);
                const int n1 = pl.n1, n2 = pl.n2;
                const int thr = A.p.qualified_quality_phred + 33;
                int lowq1 = lowq_any(r1q, len1, thr, lane), lowq2 = 0;
                int poly1 = 0, poly2 = 0;
                if (pl.cand1) poly1 = polyx_exact(r1, len1, A.p.poly_size_limit, A.p.allow_mismatch_in_poly, lut2, lane);
                int off = 0, ol = 0, diff = 0;
                if (paired) {
                    lowq2 = lowq_any(r2q, len2, thr, lane);
                    if (pl.cand2) poly2 = polyx_exact(r2, len2, A.p.poly_size_limit, A.p.allow_mismatch_in_poly, lut2, lane);
                    overlap_np(pl.np, pl.P1, pl.RC, len1, len2, lane, off, ol, diff);
                }
                if (lane == 0) {
                    aqc_ops r;
                    r.poly1 = (uint8_t)poly1; r.poly2 = (uint8_t)poly2;
                    r.lowq1 = (uint16_t)lowq1; r.lowq2 = (uint16_t)lowq2;
                    r.n1 = (uint16_t)n1; r.n2 = (uint16_t)n2;
                    r.len1 = (uint16_t)len1; r.len2 = (uint16_t)len2;
                    r.ov_offset = (int16_t)off; r.ov_len = (uint16_t)ol; r.ov_diff = (uint16_t)diff;
                    for (int k = 0; k < 12; k++) r.pad[k] = 0;
                    const uint4 *src = reinterpret_cast<const uint4 *>(&r);
                    uint4 *dst = reinterpret_cast<uint4 *>(&A.ops[pp]);
                    dst[0] = src[0]; dst[1] = src[1];
                }
                continue;
            }

            // ================================ MODE_FILTER ================================
            uint32_t edits[4] = {0, 0, 0, 0};
            int n_edits = 0;
            int ov_off = 0, ov_len = 0, ov_diff = 0;

            do {
                if (do_trim) {                                           // preprocesser.py:455-466
                    py_trim(olen1, A.p.trim_front, A.p.trim_tail, start1, len1);
                    if (len1 < 5) { cls = AQC_BADTRIM1; break; }
                    if (paired) {
                        py_trim(olen2, A.p.trim_front2, A.p.trim_tail2, start2, len2);
                        if (len2 < 5) { cls = AQC_BADTRIM2; break; }
                    }
                }
                if (len1 < A.p.seq_len_req) { cls = AQC_BADLEN; break; }   // :476-479 (R2 never checked, quirk Q3)

                uint8_t *r1 = S1 + start1, *r1q = Q1 + start1;
                uint8_t *r2 = paired ? S2 + start2 : nullptr, *r2q = paired ? Q2 + start2 : nullptr;
                PairPlanes pl;
                prepare_pair(r1, len1, r2, len2, paired, A.p.poly_size_limit > 0, A.p.poly_size_limit, A.poly_m, lut1, bytemask16, lane, pl);

                if (A.p.poly_size_limit > 0) {                             // :482-490
                    bool poly = false;
                    if (__builtin_expect(pl.cand1 || pl.cand2, 0)) {        // rare: exact window test of the screened mates
#pragma unroll 1
                        for (int m = 0; m < 2 && !poly; m++)
                            if (m ? pl.cand2 : pl.cand1)
                                poly = polyx_exact(m ? r2 : r1, m ? len2 : len1, A.p.poly_size_limit, A.p.allow_mismatch_in_poly, lut2, lane) != 0;
                    }
                    if (poly) { cls = AQC_BADPOL; break; }
                }
                if (A.p.unqualified_base_limit > 0) {                      // :493-501 (only lowQual1 tested, quirk Q2)
                    int lowq1 = lowq_any(r1q, len1, A.p.qualified_quality_phred + 33, lane);
                    if (lowq1 > A.p.unqualified_base_limit) { cls = AQC_BADLQC; break; }
                }
                if (A.p.n_base_limit > 0) {                                // :504-512
                    if (pl.n1 > A.p.n_base_limit || pl.n2 > A.p.n_base_limit) { cls = AQC_BADNCT; break; }
                }
                if (paired && !A.p.no_overlap) {                           // :515-617
                    const int np = pl.np;
                    uint32_t (&P1)[4] = pl.P1;
                    uint32_t (&RC)[4] = pl.RC;
                    int offset, ol, distance;
                    overlap_np(np, P1, RC, len1, len2, lane, offset, ol, distance);       // :516
                    if (lane == 0) atomicAdd(&s_ovh[ol], 1u);                              // :517
                    if (__builtin_expect(offset < 0 && ol > 30, 0)) {                     // :520 adapter trimming
                        // rc(r2[0:ol]) = last ol bases of rc(r2): shift the rc planes down by len2-ol
                        const int sh = len2 - ol;
                        if (sh > 0) {
#pragma unroll
                            for (int k = 0; k < 4; k++) if (k < np) RC[k] = plane_window(RC[k], lane + (sh >> 5), sh & 31);
                        }
                        len1 = ol; len2 = ol;                                             // :522-525
                        bump(AQC_C_TRIMMED_ADAPTER_BASE, (unsigned long long)(2 * (-offset)));   // :526
                        bump(AQC_C_TRIMMED_ADAPTER_READ, 1);
                        if (len1 < A.p.seq_len_req) {                                     // :529-532
                            ov_off = offset; ov_len = ol; ov_diff = distance;
                            cls = AQC_BADLEN; break;
                        }
                        overlap_np(np, P1, RC, len1, len2, lane, offset, ol, distance);   // :534
                    }
                    ov_off = offset; ov_len = ol; ov_diff = distance;
                    if (lane == 0) atomicAdd(&s_dih[distance], 1u);                        // :536
                    if (distance > 3) { cls = AQC_BADDIFF; break; }                        // :538-541
                    if (ol > 30) {                                                         // :542
                        bump(AQC_C_OVERLAPPED, 1);
                        bump(AQC_C_OVERLAP_LEN_SUM, (unsigned long long)ol);
                        bump(AQC_C_OVERLAP_BASE_SUM, (unsigned long long)(2 * ol));
                        bump(AQC_C_OVERLAP_BASE_ERR, (unsigned long long)distance);
                        if (distance > 0) {                                                // :551
                            // mismatch mask of the walk alignment r1[len1-ol+o] vs rc[o] (always this
                            // alignment, whatever offset was found: quirk Q8)
                            const int oc = len1 - ol;
                            uint32_t xx = 0;
#pragma unroll
                            for (int k = 0; k < 4; k++) if (k < np) xx |= plane_window(P1[k], lane + (oc >> 5), oc & 31) ^ RC[k];
                            xx &= lowmask(ol - (lane << 5));
                            int corrected = 0, masked = 0, skipped = 0;
                            int em_cell[3] = {-1, -1, -1};
                            int done = 0;
                            while (done < distance) {
                                uint32_t hv = __ballot_sync(FULL, xx != 0);
                                if (!hv) break;
                                int jl = __ffs(hv) - 1;
                                uint32_t wj = __shfl_sync(FULL, xx, jl);
                                int bit = __ffs(wj) - 1;
                                if (lane == jl) xx &= xx - 1;
                                const int o = (jl << 5) + bit;
                                const int p1 = len1 - ol + o, p2 = len2 - 1 - o;
                                const uint8_t b1 = r1[p1];                                 // :564
                                const uint8_t b2 = lut3[r2[p2]];                           // :565 util.complement
                                const uint8_t qa = r1q[p1], qb = r2q[p2];                  // :566-567
                                const int Qa = (int)qa - 33, Qb = (int)qb - 33;
                                __syncwarp();      // every lane has read the bytes before lane 0 patches them
                                bool fixed = false;
                                uint32_t e = 0;
                                if (Qa >= 30 && Qb <= 14) {                                // :571
                                    if (b1 != 'N' && b2 != 'N') {
                                        uint32_t la = lut2[lut3[b1]], lc = lut2[lut3[b2]];
                                        if ((la & 0x40u) && (lc & 0x40u)) em_cell[done] = (int)((la & 7u) * 4u + (lc & 7u));   // :573
                                    }
                                    if (!A.p.no_correction) {                              // :574-578
                                        const uint8_t nb = lut3[b1];
                                        if (lane == 0) { r2[p2] = nb; r2q[p2] = qa; }
                                        corrected++; fixed = true;
                                        e = (uint32_t)(start2 + p2) | (1u << 10) | ((uint32_t)nb << 16) | ((uint32_t)qa << 24);
                                    }
                                } else if (Qb >= 30 && Qa <= 14) {                         // :579
                                    if (b1 != 'N' && b2 != 'N') {
                                        uint32_t la = lut2[b2], lc = lut2[b1];
                                        if ((la & 0x40u) && (lc & 0x40u)) em_cell[done] = (int)((la & 7u) * 4u + (lc & 7u));   // :581
                                    }
                                    if (!A.p.no_correction) {                              // :582-586
                                        if (lane == 0) { r1[p1] = b2; r1q[p1] = qb; }
                                        corrected++; fixed = true;
                                        e = (uint32_t)(start1 + p1) | (0u << 10) | ((uint32_t)b2 << 16) | ((uint32_t)qb << 24);
                                    }
                                }
                                if (!fixed) {                                              // :587-595
                                    if (A.p.mask_mismatch) {
                                        if (lane == 0) { r2q[p2] = '!'; r1q[p1] = '!'; }
                                        masked++;
                                        e = (uint32_t)(start1 + p1) | (2u << 10) | ((uint32_t)(start2 + p2) << 16);
                                    } else {
                                        skipped++;
                                        e = (uint32_t)(start1 + p1) | (3u << 10) | ((uint32_t)(start2 + p2) << 16);
                                    }
                                }
                                if (n_edits < 4) edits[n_edits++] = e;
                                done++;
                            }
                            __syncwarp();
                            if (corrected + masked + skipped == distance) {               // :603-610
                                for (int k = 0; k < 3; k++) if (em_cell[k] >= 0 && lane == em_cell[k]) wc1 += 1;
                                if (corrected > 0) bump(AQC_C_READ_CORRECTED, 1);
                                bump(AQC_C_BASE_CORRECTED, (unsigned long long)corrected);
                                bump(AQC_C_BASE_ZERO_QUAL_MASKED, (unsigned long long)(2 * masked));
                                bump(AQC_C_BASE_SKIPPED_CORRECTION, (unsigned long long)(2 * skipped));
                            } else { cls = AQC_BADMISMATCH; break; }                      // :611-614
                        }
                    }
                }
            } while (0);

            if (cls == AQC_GOOD) {
                bump(AQC_C_GOOD_READS, 1);
                bump(AQC_C_GOOD_BASES_R1, (unsigned long long)len1);
                bump(AQC_C_GOOD_BASES_R2, (unsigned long long)len2);
                const uint64_t total_reads = gidx + 1;
                if (__builtin_expect(!A.no_stats && (A.p.qc_sample <= 0 || total_reads < (uint64_t)A.p.qc_sample), 0)) {       // :624
#pragma unroll 1
                    for (int m = 0; m < (paired ? 2 : 1); m++)      // one inlined copy of statRead for both mates
                        stat_read(m ? S2 + start2 : S1 + start1, m ? Q2 + start2 : Q1 + start1, m ? len2 : len1, m, gidx, qsm, A.qc[m],
                                  lut1, lut2, lut3, A.p.qc_kmer, lane, A.error_flag);
                }
            } else {
                bump(AQC_C_BADTRIM1 + (cls - AQC_BADTRIM1), 1);
            }
            if (lane == 0) {
                uint4 w0, w1;
                w0.x = (uint32_t)cls | ((uint32_t)n_edits << 8) | ((uint32_t)start1 << 16);
                w0.y = (uint32_t)len1 | ((uint32_t)start2 << 16);
                w0.z = (uint32_t)len2 | (((uint32_t)ov_off & 0xFFFFu) << 16);
                w0.w = (uint32_t)ov_len | ((uint32_t)ov_diff << 16);
                w1.x = edits[0]; w1.y = edits[1]; w1.z = edits[2]; w1.w = edits[3];
                uint4 *dst = reinterpret_cast<uint4 *>(&A.results[pp]);
                dst[0] = w0; dst[1] = w1;
            }
        }

        // ---- end of tile: release the stage, refill it, flush QC accumulators when due ----
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            uint32_t nxt = tile + (uint32_t)NSTAGES * gridDim.x;
            tile_next[st] = 0;
            if (nxt < num_tiles && !listed) issue_tile(nxt, st);
            qc_reads_since_flush += (p1 - p0);
        }
        __syncthreads();
        if (qc_reads_since_flush + (uint32_t)P > QC_FLUSH_READS) {
            flush_qc();
            __syncthreads();
            if (tid == 0) qc_reads_since_flush = 0;
            __syncthreads();
        }
    }

    // ---- epilogue: flush everything this CTA accumulated ----
    __syncthreads();
    flush_qc();
    if (MODE == MODE_FILTER || MODE == MODE_LIST) {
        for (int i = tid; i <= A.max_len; i += THREADS) {
            uint32_t v = s_ovh[i]; if (v) atomicAdd(&A.counters[AQC_C_OVERLAP_HIST + i], (unsigned long long)v);
            v = s_dih[i]; if (v) atomicAdd(&A.counters[AQC_C_DISTANCE_HIST + i], (unsigned long long)v);
        }
        if (wc0) atomicAdd(&A.counters[lane], wc0);
        if (blockIdx.x == 0 && tid == 0 && !listed) {      // TOTAL_READS / TOTAL_BASES (:416,:431,:433) are sums over the batch
            atomicAdd(&A.counters[AQC_C_TOTAL_READS], (unsigned long long)A.n);
            atomicAdd(&A.counters[AQC_C_TOTAL_BASES_R1], (unsigned long long)(A.off1[A.n] - A.off1[0]));
            if (paired) atomicAdd(&A.counters[AQC_C_TOTAL_BASES_R2], (unsigned long long)(A.off2[A.n] - A.off2[0]));
        }
        if (wc1 && lane < 16) atomicAdd(&A.counters[AQC_C_ERR_MATRIX + lane], wc1);
    }
}

// longest read of a batch (device-resident batches): max over off[i+1]-off[i]
__global__ void maxlen_kernel(const uint32_t *off1, const uint32_t *off2, uint32_t n, uint32_t *out) {
    uint32_t m = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        m = max(m, off1[i + 1] - off1[i]);
        if (off2) m = max(m, off2[i + 1] - off2[i]);
    }
    m = __reduce_max_sync(FULL, m);
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

}  // namespace aqc
