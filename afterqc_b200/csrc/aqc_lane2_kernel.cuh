// aqc_lane2_kernel.cuh -- second generation of the lane-per-pair filter kernel (aqc_params.filter_kernel = 3).
//
// Same per-lane algorithm and device functions as lane_kernel (aqc_lane_kernel.cuh), which is the kernel that has been
// checked on hardware and is therefore left untouched.  Two scheduling changes, both aimed at what its first GPU timing
// showed (31 % of the issue slots with 12 resident warps per SM; the statistics tiles at the head of the batch all landing
// on the first warps of a static round-robin):
//   * two columns per warp instead of three: the qualities of mate 1 are copied over the bases of mate 1 as soon as those
//     are converted (their copy overlaps the conversion of mate 2), so a 4-warp CTA needs 51 KB instead of 70 KB at PE150 and
//     a fourth CTA fits on the SM (16 warps);
//   * tiles are claimed from a counter in HBM, one ahead, instead of a static stride.
// Written without GPU access (verified under the SIMT emulator only); bench.py tries it before lane_kernel and keeps it only
// if it is identical and faster on the GPU it runs on.
#pragma once
#include "aqc_lane_kernel.cuh"

namespace aqc {

// dynamic shared memory of one CTA:
//   [nwarps][ 2 * lane_col_cap ]           per-warp stage: column 0 = bases 1, later qualities 1 | column 1 = bases 2 (or qualities 1, single-end)
//   [nwarps][ 4 * 32*NW ]                  per-warp scratch of the statistics hand-over
//   luts (768 B)
//   qc acc [2][5][max_len] u32, qc disc [2][max_len] u32, overlap_hist [max_len+1], distance_hist [max_len+1], err matrix [16]
// SMODE: sampled postfilter statistics -- 0 = stat_read hand-over (one warp per read), 1 = stat_tile (one lane per read), both in this kernel;
//        2 = none here: stat_lane_kernel<.., POST> runs after this kernel and pair_kernel's list mode (pairs handed over are marked in skip_bits)
template <bool PAIRED, int NW, int SMODE = 0>
__global__ void __launch_bounds__(LANE_MAX_WARPS * 32, (NW > 5 ? 3 : 4)) lane2_kernel(const __grid_constant__ LArgs L) {
    AQC_DYN_SMEM(smem_raw);
    __shared__ __align__(8) uint64_t full_bar[2 * LANE_MAX_WARPS];      // per warp: [0] bases landed, [1] qualities landed
    const KArgs &A = L.k;
    constexpr bool paired = PAIRED;
    constexpr int MAXB = 32 * NW;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = blockDim.x >> 5;
    const int col_cap = L.lane_col_cap;
    const int ncols = 2;

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
        for (int s = 0; s < 2 * nwarps; s++) mbar_init(&full_bar[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    QcSmem qsm; qsm.acc = s_acc; qsm.disc = s_disc; qsm.max_len = A.max_len;
    const uint32_t flush_limit = QC_FLUSH_READS / (uint32_t)nwarps;
    uint32_t stat_since_flush = 0;

    // lane i owns scalar counter i
    unsigned long long wc0 = 0;

    uint64_t *bar = &full_bar[2 * warp], *qbar = &full_bar[2 * warp + 1];
    // tiles are claimed from a counter in HBM, one tile ahead of the one being worked on: warps that meet expensive tiles
    // (the statistics window sits at the head of the batch) simply take fewer
    auto claim_tile = [&]() -> uint32_t {
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(L.tile_counter, 1u);
        return __shfl_sync(FULL, t, 0);
    };

    // offsets of the lane's pair in tile t (pairs beyond n: empty records at the end of the columns)
    auto load_offsets = [&](uint32_t t, uint32_t &a1, uint32_t &e1, uint32_t &a2, uint32_t &e2) {
        const uint32_t pp = min(t * 32u + (uint32_t)lane, A.n), pq = min(pp + 1u, A.n);
        a1 = A.off1[pp]; e1 = A.off1[pq];
        a2 = 0; e2 = 0;
        if (paired) { a2 = A.off2[pp]; e2 = A.off2[pq]; }
    };
    // producer (whole warp computes, lane 0 issues).  Paired input: the bases of both mates first; the qualities of mate 1
    // follow into column 0 once mate 1 has been converted (issue_quals) -- two columns per warp instead of three, so that a
    // fourth CTA fits on the SM.  Single-end input: bases and qualities together, as in lane_kernel.
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
    uint32_t t = claim_tile();
    if (t < A.num_tiles) {
        load_offsets(t, a1, e1, a2, e2);
        issue_tile(a1, e1, a2, e2);
    }
    uint32_t parity = 0, qparity = 0;

#pragma unroll 1
    for (; t < A.num_tiles;) {
        const uint32_t tn = claim_tile();
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
                    qc_flush_warp(A.qc, s_acc, s_disc, A.max_len, lane);
                    stat_since_flush = 0;
                }
            }
        }

        a1 = na1; e1 = ne1; a2 = na2; e2 = ne2;
        t = tn;
    }

    // ---- epilogue: flush everything this CTA accumulated ----
    __syncthreads();
    qc_flush_cta(A.qc, s_acc, s_disc, A.max_len, tid, blockDim.x);
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
