// aqc_stat2.cuh -- QualityControl.statRead (qualitycontrol.py:73-122) for 32 reads at a time: ONE LANE PER READ
// (aqc_params.stat_kernel = 2; written without GPU access -- verified under the SIMT emulator only, opt-in).
//
// stat_read (aqc_device.cuh) gives a warp to one read: lane = position, ~1000 warp-instructions per 150-base read, one
// shared-memory atomic lane-operation per base and one L2 reduction per k-mer.  With lane_kernel at ~150 warp-instructions per
// PAIR the sampled statistics became a quarter of the filter kernel and the prefilter launch a sixth of the bench step.
// Here every lane converts ITS read to code-bit planes (registers) and the warp walks the cycle index i together, so that
// the per-cycle accumulators (which are indexed by i) are reduced ACROSS the lanes before they touch memory:
//   * class counts of A,T,C,G (6 bits each) and the discontinuity sum (8 bits) travel in ONE warp reduction (REDUX), the
//     quality-byte sums of the four classes in two more (2 x 16 bits each); lanes 0..3 add "count << 20 | byte sum" to the
//     CTA's packed accumulators, lane 4 the discontinuity: 5 shared-memory atomic lane-operations per cycle instead of 32+;
//   * discontinuity: "base j != base j+1" is a bit string per lane; the 5-base window is a popcount of 4 bits;
//   * G/C count: popcount of plane 0;
//   * k-mers: the K-bit windows of the planes ARE the dense table index; one RED per lane and cycle, the first-seen stamp is
//     checked with a load issued one cycle ahead (as stat_read does).  K-mers holding an N take the side-table path of
//     stat_read with the bytes rebuilt from the planes.
// Reads of A,C,G,T,N with >= 5 and <= 32*NW bases take this path; the caller hands the others (a byte outside that
// alphabet, empty or too short reads) to stat_read, which stays the single definition of those cases.  Same accumulators,
// same flush rules (the caller counts the reads), same results.
#pragma once
// included by aqc_lane_kernel.cuh after its helpers (LanePlanes, shr_bits) and before the kernels that call stat_tile

namespace aqc {

// planes of a read in HBM (generic word loads, never past the last word that holds a base); exotic = a byte outside A,C,G,T,N
template <int NW>
__device__ __forceinline__ void convert_g(const uint8_t *s, int len, LanePlanes<NW> &P, bool &exotic) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(s);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
    const int lead = (int)(a & 3), sh = lead * 8;
    const int last = len > 0 ? (lead + len - 1) >> 2 : -1;        // last word that holds a base
    exotic = false;
    uint32_t prev = len > 0 ? w[0] : 0u;
#pragma unroll 1
    for (int c = 0; c < NW; c++) {
        uint32_t p0 = 0, p1 = 0, pn = 0;
        const int nvalid = len - 32 * c;
        if (nvalid > 0) {
            uint32_t v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int k = 8 * c + j + 1;
                const uint32_t cur = k <= last ? w[k] : 0u;
                v[j] = __funnelshift_r(prev, cur, sh);
                prev = cur;
            }
            uint32_t rlo = 0, rhi = 0, bad = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {                       // same SWAR conversion as lane_convert
                const uint32_t t = v[j] & 0x06060606u;
                const uint32_t tt = t >> 1;
                const uint32_t e = prmt_raw(0x47544341u, 0u, prmt_raw(tt + (tt >> 4), 0u, 0x4420u));
                bad |= e ^ v[j];
                const uint32_t z = ((t << 2) + tt) & 0x11111111u;
                const uint32_t r = z * 0x01020408u;
                constexpr uint32_t sel[4] = {0x3217u, 0x3270u, 0x3710u, 0x7210u};
                if (j < 4) rlo = __byte_perm(rlo, r, sel[j & 3]); else rhi = __byte_perm(rhi, r, sel[j & 3]);
            }
            const uint32_t l0 = rlo & 0x0F0F0F0Fu, h0 = rhi & 0x0F0F0F0Fu;
            const uint32_t l1 = (rlo >> 4) & 0x0F0F0F0Fu, h1 = (rhi >> 4) & 0x0F0F0F0Fu;
            const uint32_t a0 = (l0 | (l0 >> 4)) & 0x00FF00FFu, b0 = (h0 | (h0 >> 4)) & 0x00FF00FFu;
            const uint32_t a1 = (l1 | (l1 >> 4)) & 0x00FF00FFu, b1 = (h1 | (h1 >> 4)) & 0x00FF00FFu;
            const uint32_t vm = lowmask(nvalid);
            p0 = __byte_perm(a0, b0, 0x6420) & vm;
            p1 = __byte_perm(a1, b1, 0x6420) & vm;
            if (__builtin_expect(bad != 0u, 0)) {               // a byte of the 32 is not A,C,G,T: N, foreign, or beyond the read
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

// byte patches of one mate from the record's edits (positions relative to the final read; -1 = unused)
struct MatePatches {
    int pos[4];
    uint32_t base[4];      // 0 = base unchanged
    uint32_t qual[4];      // 0 = quality unchanged
};

__device__ __forceinline__ void no_patches(MatePatches &mp) {
#pragma unroll
    for (int k = 0; k < 4; k++) { mp.pos[k] = -1; mp.base[k] = 0; mp.qual[k] = 0; }
}

// the edits of the correction walk (preprocesser.py:575-595) that touch mate `mate`; start1/start2 = trim_front offsets
__device__ __forceinline__ void mate_patches(const uint32_t (&edits)[4], int n_edits, int mate, int start1, int start2, MatePatches &mp) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
        mp.pos[k] = -1; mp.base[k] = 0; mp.qual[k] = 0;
        if (k < n_edits) {
            const uint32_t e = edits[k];
            const int kind = (int)AQC_EDIT_KIND(e);
            if (kind == 0 && mate == 0) { mp.pos[k] = (int)AQC_EDIT_POS(e) - start1; mp.base[k] = AQC_EDIT_BASE(e); mp.qual[k] = AQC_EDIT_QUAL(e); }
            else if (kind == 1 && mate == 1) { mp.pos[k] = (int)AQC_EDIT_POS(e) - start2; mp.base[k] = AQC_EDIT_BASE(e); mp.qual[k] = AQC_EDIT_QUAL(e); }
            else if (kind == 2) { mp.pos[k] = mate == 0 ? (int)AQC_EDIT_POS(e) - start1 : (int)AQC_EDIT_POS2(e) - start2; mp.qual[k] = '!'; }
        }
    }
}

// statRead of the lanes' reads (s/q: the final read's bytes in HBM before the walk's edits, len bases; mp: those edits) into
// the CTA's shared accumulators and the QC object qd.  Returns true for the lanes whose read was handled here; a lane with
// want && !returned needs stat_read.  Whole warp calls; `want` may differ per lane.
template <int NW>
__device__ __forceinline__ bool stat_tile(bool want, const uint8_t *s, const uint8_t *q, int len, uint64_t order, int mate,
                                          const MatePatches &mp, const QcSmem &sm, const QcDev &qd, int K, int lane, int *error_flag) {
    LanePlanes<NW> P;
    uint32_t D2[NW];                                          // bit x+2: base x differs from base x+1
#pragma unroll
    for (int i = 0; i < NW; i++) P.p0[i] = P.p1[i] = P.pn[i] = D2[i] = 0;
    bool ok = want && len >= 5 && len <= 32 * NW;
    if (ok) {
        bool exotic = false;
        convert_g<NW>(s, len, P, exotic);
#pragma unroll
        for (int k = 0; k < 4; k++) {                         // corrected bases (preprocesser.py:575,583)
            if (mp.pos[k] >= 0 && mp.base[k]) {
                const uint32_t b = mp.base[k];
                const bool acgt = b == 'A' || b == 'C' || b == 'G' || b == 'T', isn = b == 'N';
                if (!acgt && !isn) exotic = true;
                const uint32_t bit = 1u << (mp.pos[k] & 31);
                const uint32_t c0 = acgt ? (b >> 1) & 1u : 0u, c1 = acgt ? (b >> 2) & 1u : 0u;
                const int wsel = mp.pos[k] >> 5;
#pragma unroll
                for (int w = 0; w < NW; w++) {                 // every word rewritten: keeps the planes in registers
                    const uint32_t mb = (w == wsel) ? bit : 0u;
                    P.p0[w] = (P.p0[w] & ~mb) | (c0 ? mb : 0u);
                    P.p1[w] = (P.p1[w] & ~mb) | (c1 ? mb : 0u);
                    P.pn[w] = (P.pn[w] & ~mb) | (isn ? mb : 0u);
                }
            }
        }
        ok = !exotic;
    }
    const uint32_t okm = __ballot_sync(FULL, ok);
    if (okm == 0u) return false;

    uint32_t *acc = sm.acc + (size_t)mate * QC_CLASSES * sm.max_len;
    uint32_t *dsc = sm.disc + (size_t)mate * sm.max_len;
    int gc = 0, d_head = 0, d_tail = 0;
    {
        uint32_t df[NW];                                      // bit x: base x differs from base x+1, x <= len-2
#pragma unroll
        for (int w = 0; w < NW; w++) {
            const uint32_t n0 = (w + 1 < NW) ? P.p0[w + 1 < NW ? w + 1 : 0] : 0u, n1 = (w + 1 < NW) ? P.p1[w + 1 < NW ? w + 1 : 0] : 0u;
            const uint32_t nn = (w + 1 < NW) ? P.pn[w + 1 < NW ? w + 1 : 0] : 0u;
            df[w] = ((P.p0[w] ^ __funnelshift_r(P.p0[w], n0, 1)) | (P.p1[w] ^ __funnelshift_r(P.p1[w], n1, 1)) | (P.pn[w] ^ __funnelshift_r(P.pn[w], nn, 1)))
                    & lowmask(len - 1 - 32 * w);
            gc += __popc(P.p0[w]);                            // plane 0 is set for C and G
        }
#pragma unroll
        for (int w = 0; w < NW; w++) D2[w] = __funnelshift_l(w ? df[w - 1 >= 0 ? w - 1 : 0] : 0u, df[w], 2);
        if (ok) {                                             // windows clamped at the read ends (qualitycontrol.py:97-104)
            d_head = __popc(df[0] & 0xFu);
            shr_bits<NW>(df, len - 5);
            d_tail = __popc(df[0] & 0xFu);
        }
    }
    const uint32_t km = (1u << K) - 1u;                       // K <= AQC_MAX_KMER (8)
    const int nk = len - K;                                   // k-mers start at i < len - K (quirk Q11)
    const int maxlen = (int)__reduce_max_sync(FULL, ok ? (unsigned)len : 0u);

    // quality bytes: aligned words of the lane's read, funnel-shifted to the read's first byte
    const uintptr_t qa = reinterpret_cast<uintptr_t>(q);
    const uint32_t *qw = reinterpret_cast<const uint32_t *>(qa & ~(uintptr_t)3);
    const int qsh = (int)(qa & 3) * 8;
    const int qlast = ok ? (((int)(qa & 3) + len - 1) >> 2) : -1;
    uint32_t qprev = ok ? qw[0] : 0u;

    unsigned long long pend_val = 0, pend_when = 0;
    uint32_t pend_idx = 0;
    bool pend = false;
    const unsigned long long when0 = order << 11;
    unsigned long long *const kfirst = qd.kfirst, *const kcnt = qd.kcnt;     // registers, not indexed constant-bank loads
    // the lane's share of the per-cycle flush (see the loop)
    uint32_t *const abase = lane < 4 ? acc + lane * sm.max_len : (lane == 4 ? dsc : acc);
    const uint32_t cshift = lane < 4 ? 6u * (uint32_t)lane : 24u, cmask = lane < 4 ? 63u : (lane == 4 ? 0xFFu : 0u);
    const uint32_t qshift = 16u * (uint32_t)(lane & 1), qmask = lane < 4 ? 0xFFFFu : 0u;

#pragma unroll 1
    for (int j = 0; 4 * j < maxlen; j++) {                    // 4 cycles per iteration; word 0 of the queues = current 32 cycles
        const int jj = j & 7;
        if (jj == 0 && j > 0) {
#pragma unroll
            for (int i = 0; i < NW; i++) {
                P.p0[i] = (i + 1 < NW) ? P.p0[i + 1 < NW ? i + 1 : 0] : 0u;
                P.p1[i] = (i + 1 < NW) ? P.p1[i + 1 < NW ? i + 1 : 0] : 0u;
                P.pn[i] = (i + 1 < NW) ? P.pn[i + 1 < NW ? i + 1 : 0] : 0u;
                D2[i] = (i + 1 < NW) ? D2[i + 1 < NW ? i + 1 : 0] : 0u;
            }
        }
        const int sh = 4 * jj;
        uint32_t x0 = __funnelshift_r(P.p0[0], P.p0[1], sh), x1 = __funnelshift_r(P.p1[0], P.p1[1], sh);
        uint32_t xn = __funnelshift_r(P.pn[0], P.pn[1], sh), xd = __funnelshift_r(D2[0], D2[1], sh);
        const uint32_t qcur = (j + 1 <= qlast) ? qw[j + 1] : 0u;
        uint32_t qv = __funnelshift_r(qprev, qcur, qsh);
        qprev = qcur;
#pragma unroll
        for (int k = 0; k < 4; k++)                            // qualities rewritten by the walk (:576,:584,:590-591)
            if (mp.qual[k] && (mp.pos[k] >> 2) == j) {
                const uint32_t by = (uint32_t)(mp.pos[k] & 3);
                qv = __byte_perm(qv, mp.qual[k], 0x3210u ^ ((4u ^ by) << (4 * by)));
            }
        const int iend = min(4 * j + 4, maxlen);
#pragma unroll 1
        for (int i = 4 * j; i < iend; i++) {                   // one copy of the cycle body; bit 0 of the x words = cycle i
            const bool v = ok && i < len;
            if (pend && pend_val > pend_when) atomicMin(&kfirst[pend_idx], pend_when);
            pend = false;
            const uint32_t c0 = x0 & 1u, c1 = x1 & 1u;
            const bool isn = xn & 1u;
            const uint32_t cls = (c0 << 1) | c1;              // A0 T1 C2 G3 (ALL_BASES order, qualitycontrol.py:24)
            const uint32_t qq = qv & 0xFFu;
            uint32_t d = __popc(xd & 0xFu);
            if (i < 2) d = d_head; else if (i + 3 >= len) d = d_tail;
            const bool vb = v && !isn;
            const uint32_t r1 = v ? ((isn ? 0u : (1u << (6 * cls))) | (d << 24)) : 0u;
            const uint32_t ra = (vb && cls < 2u) ? qq << (16 * cls) : 0u;
            const uint32_t rb = (vb && cls >= 2u) ? qq << (16 * (cls - 2u)) : 0u;
            const uint32_t R1 = __reduce_add_sync(FULL, r1), RA = __reduce_add_sync(FULL, ra), RB = __reduce_add_sync(FULL, rb);
            {   // lanes 0..3 own the classes A,T,C,G (count << 20 | byte sum; a zero count has a zero sum), lane 4 the discontinuity
                const uint32_t cnt = (R1 >> cshift) & cmask;
                const uint32_t val = lane == 4 ? cnt : ((cnt << 20) | (((lane < 2 ? RA : RB) >> qshift) & qmask));
                if (val) atomicAdd(abase + i, val);
            }
            const uint32_t nm = __ballot_sync(FULL, v && isn);
            if (__builtin_expect(nm != 0u, 0)) {               // class "other" (here: N), :83-92 count it in totalNum/totalQual only
                const uint32_t qn = __reduce_add_sync(FULL, (v && isn) ? qq : 0u);
                if (lane == 5) atomicAdd(&acc[4 * sm.max_len + i], ((uint32_t)__popc(nm) << 20) | qn);
            }
            if (v && i < nk) {
                // k-mer code planes (A0 C1 G2 T3): bit 0 = C or T = p0 ^ p1, bit 1 = G or T = p1
                const uint32_t w0 = (x0 ^ x1) & km, w1 = x1 & km, wn = xn & km;
                const unsigned long long when = when0 | ((unsigned long long)i << 1);
                if (__builtin_expect(wn == 0u, 1)) {
                    const uint32_t idx = (w1 << K) | w0;
                    pend_val = __ldcg(&kfirst[idx]);
                    pend_idx = idx; pend_when = when; pend = true;
                    atomicAdd(&kcnt[idx], 1ULL);
                } else {                                        // a k-mer with an N: side table, keyed by its bytes (see stat_read)
                    unsigned long long key = 0, rkey = 0;
                    for (int t = 0; t < K; t++) {
                        const uint32_t t0 = (x0 >> t) & 1u, t1 = (x1 >> t) & 1u, tn = (xn >> t) & 1u;
                        const uint32_t code = (t0 << 1) | t1;                 // A0 T1 C2 G3
                        const unsigned long long bj = tn ? 'N' : ((0x47435441u >> (8 * code)) & 0xFFu);          // "ATCG"
                        const unsigned long long cj = tn ? 'N' : ((0x43474154u >> (8 * code)) & 0xFFu);          // complements "TAGC"
                        key = (key << 8) | bj;
                        rkey |= cj << (8 * t);
                    }
                    const int h = side_slot(qd, key);
                    const int hr = side_slot(qd, rkey);
                    if (h < 0 || hr < 0) atomicExch(error_flag, AQC_ERR_KMER_TABLE_FULL);
                    else {
                        atomicAdd(&qd.scnt[h], 1ULL);
                        first_min(&qd.sfirst[h], when);
                    }
                }
            }
            x0 >>= 1; x1 >>= 1; xn >>= 1; xd >>= 1; qv >>= 8;
        }
    }
    if (pend && pend_val > pend_when) atomicMin(&kfirst[pend_idx], pend_when);
    if (ok) atomicAdd(&qd.gchist[gc], 1ULL);                                            // :112
    const uint32_t nks = __reduce_add_sync(FULL, ok ? (unsigned)max(nk, 0) : 0u);
    if (lane == 0) {
        if (nks) atomicAdd(&qd.scal[0], (unsigned long long)nks);                       // totalKmer :114
        atomicAdd(&qd.scal[1], (unsigned long long)__popc(okm));
    }
    return ok;
}


// ------------------------------------------------------------------------------------------------------------------
// Prefilter statistics (statFile's window, qualitycontrol.py:331-357) with one lane per read: the MODE_STAT launch of
// aqc_stat_reads when aqc_params.stat_kernel = 2 and no read of the batch is longer than 256 bases.  Reads go straight from
// HBM to the lanes' registers (no staging: every byte is used once); warps walk tiles of 32 records round-robin.
// dynamic shared memory:  [nwarps][2 * 32*NW] scratch of the stat_read hand-over | luts (768 B) |
//                         qc acc [2][5][max_len] u32 | qc disc [2][max_len] u32
// ------------------------------------------------------------------------------------------------------------------
constexpr int STAT2_WARPS = 4;

struct SArgs {
    KArgs k;
    const uint32_t *skip_bits;    // POST: bit pp set = pair pp was filtered (and stat'd) by pair_kernel's list mode
};

// POST = false: the prefilter window [stat_lo, stat_hi) of raw reads (aqc_stat_reads).
// POST = true : the postfilter statistics of a filter launch that ran with SMODE 2 (preprocesser.py:624-627): the sampled GOOD
//               pairs of k.results, final slices and the correction walk's edits taken from the 32-byte records.
template <bool PAIRED, int NW, bool POST = false>
__global__ void __launch_bounds__(STAT2_WARPS * 32, (POST ? 4 : 5)) stat_lane_kernel(const __grid_constant__ SArgs S) {
    const KArgs &A = S.k;
    AQC_DYN_SMEM(smem_raw);
    constexpr bool paired = PAIRED;
    constexpr int MAXB = 32 * NW;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = blockDim.x >> 5;
    uint8_t *scratch = smem_raw + (size_t)warp * 2 * MAXB;
    uint8_t *lutbase = smem_raw + (size_t)nwarps * 2 * MAXB;
    const uint8_t *lut1 = lutbase, *lut2 = lutbase + 256, *lut3 = lutbase + 512;
    uint32_t *s_acc = reinterpret_cast<uint32_t *>(lutbase + 768);
    uint32_t *s_disc = s_acc + 2 * QC_CLASSES * A.max_len;
    const int n_qc_words = 2 * QC_CLASSES * A.max_len + 2 * A.max_len;

    for (int i = tid; i < 768; i += blockDim.x) lutbase[i] = reinterpret_cast<const uint8_t *>(A.luts)[i];
    for (int i = tid; i < n_qc_words; i += blockDim.x) s_acc[i] = 0;
    __syncthreads();

    QcSmem qsm; qsm.acc = s_acc; qsm.disc = s_disc; qsm.max_len = A.max_len;
    const uint32_t flush_limit = QC_FLUSH_READS / (uint32_t)nwarps;
    uint32_t stat_since_flush = 0;
    MatePatches mp;
    no_patches(mp);

    const uint32_t W = gridDim.x * (uint32_t)nwarps;
#pragma unroll 1
    for (uint32_t t = blockIdx.x * (uint32_t)nwarps + (uint32_t)warp; t < A.num_tiles; t += W) {
        const uint32_t pp = t * 32u + (uint32_t)lane;
        const uint64_t gidx = A.first_index + pp;
        bool want;
        uint32_t rec[4] = {0, 0, 0, 0}, edits[4] = {0, 0, 0, 0};  // POST: the pair's record (class, slices | edits)
        if constexpr (POST) {
            want = pp < A.n && (A.p.qc_sample <= 0 || gidx + 1 < (uint64_t)A.p.qc_sample) && !((S.skip_bits[pp >> 5] >> (pp & 31u)) & 1u);
            if (want) {
                const uint4 *r = reinterpret_cast<const uint4 *>(A.results + pp);
                const uint4 w0 = r[0], w1 = r[1];
                rec[0] = w0.x; rec[1] = w0.y; rec[2] = w0.z;
                edits[0] = w1.x; edits[1] = w1.y; edits[2] = w1.z; edits[3] = w1.w;
                want = (rec[0] & 0xFFu) == (uint32_t)AQC_GOOD;
            }
        } else {
            want = pp < A.n && gidx >= A.stat_lo && gidx < A.stat_hi;
        }
        const uint32_t sbm = __ballot_sync(FULL, want);
        if (sbm == 0u) continue;
        stat_since_flush += (uint32_t)__popc(sbm);
        const uint64_t order = POST ? gidx : A.order_base + (gidx - A.stat_lo);
        const uint32_t pc = min(pp, A.n), pq = min(pp + 1u, A.n);
        const int n_edits = (int)((rec[0] >> 8) & 0xFFu);
        const int start1 = (int)(rec[0] >> 16), start2 = (int)(rec[1] >> 16);
#pragma unroll 1
        for (int m = 0; m < (paired ? 2 : 1); m++) {
            if (!A.qc[m].valid) continue;
            const uint32_t *off = m ? A.off2 : A.off1;
            const uint8_t *seq = m ? A.seq2 : A.seq1, *qual = m ? A.qual2 : A.qual1;
            uint32_t a = off[pc];
            int len = (int)(off[pq] - a);
            if constexpr (POST) {                               // the final slice of the read (trim + adapter cut) and the walk's edits
                a += (uint32_t)(m ? start2 : start1);
                len = (int)((m ? rec[2] : rec[1]) & 0xFFFFu);
                mate_patches(edits, n_edits, m, start1, start2, mp);
            }
            bool handled = false;
            if (len > MAXB) {                                   // the host picks NW from the longest read; defensive
                if (want) atomicExch(A.error_flag, AQC_ERR_TOO_LONG);
                handled = true;
            }
            const bool done = stat_tile<NW>(want && !handled, seq + a, qual + a, len, order, m, mp, qsm, A.qc[m], A.p.qc_kmer, lane, A.error_flag);
            uint32_t need = __ballot_sync(FULL, want && !handled && !done);
            while (need) {                                       // a byte outside A,C,G,T,N, an empty or a very short read: stat_read
                const int src = __ffs(need) - 1;
                need &= need - 1;
                const uint32_t ba = __shfl_sync(FULL, a, src);
                const int bl = __shfl_sync(FULL, len, src);
                const uint64_t bg = A.first_index + t * 32u + (uint32_t)src;
                const uint64_t bo = POST ? bg : A.order_base + (bg - A.stat_lo);
                __syncwarp();
                for (int x = lane; x < bl; x += 32) { scratch[x] = seq[ba + x]; scratch[MAXB + x] = qual[ba + x]; }
                if constexpr (POST) {
                    __syncwarp();
#pragma unroll
                    for (int k = 0; k < 4; k++) {                // the source lane patches its own read
                        if (lane == src && mp.pos[k] >= 0) {
                            if (mp.base[k]) scratch[mp.pos[k]] = (uint8_t)mp.base[k];
                            if (mp.qual[k]) scratch[MAXB + mp.pos[k]] = (uint8_t)mp.qual[k];
                        }
                    }
                }
                __syncwarp();
                stat_read(scratch, scratch + MAXB, bl, m, bo, qsm, A.qc[m], lut1, lut2, lut3, A.p.qc_kmer, lane, A.error_flag);
            }
        }
        __syncwarp();
        if (stat_since_flush + 32u > flush_limit) {              // packed shared accumulators: count field is 12 bits
            qc_flush_warp(A.qc, s_acc, s_disc, A.max_len, lane);
            stat_since_flush = 0;
        }
    }

    __syncthreads();
    qc_flush_cta(A.qc, s_acc, s_disc, A.max_len, tid, blockDim.x);
}

}  // namespace aqc
