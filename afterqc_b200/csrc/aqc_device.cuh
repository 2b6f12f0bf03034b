// aqc_device.cuh -- device-side building blocks of the B200 AfterQC engine (sm_100a).
//
// Execution model: one WARP per read pair.  A persistent CTA (8 warps) walks tiles of <= 32
// pairs; the four byte columns of a tile (bases/quals of both mates) are contiguous in HBM and
// are staged into shared memory with 1-D TMA bulk copies (cp.async.bulk + mbarrier), double
// buffered.  Inside a warp the bases of both mates are turned into BIT-PLANES (lane j holds bits
// 32j..32j+31 of each plane) -- by SWAR arithmetic on 16 bases per lane for A/C/G/T/N reads
// (fast_build2), by a LUT + warp ballots for anything else (build_planes) -- so that
//   * the sliding-offset Hamming scan of util.overlap_hm (util.py:158-212) becomes: each lane
//     scores ONE candidate offset with funnel-shift + XOR + popc on the first 32 positions,
//     a ballot picks the survivors in scan order, and a warp-wide popc/redux evaluates the
//     reference's acceptance rule exactly (closed form of the leaked loop variable, quirk Q6);
//   * hasPolyX (preprocesser.py:30-51) is screened by a bit-parallel run-length test and only
//     candidate reads take the exact sliding-window path;
//   * N counts are popcounts of the N plane, low-quality counts a SWAR byte compare.
// All arithmetic is integer/byte work; there is no tensor-core use by design.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/afterqc_b200.h"

namespace aqc {

// Rare paths (statRead, exact hasPolyX, LUT plane build) stay inlined: out-of-line calls force the ABI register
// convention onto the hot loop and cost 45 % of the throughput (measured, profiles/r01_notes.md).
#ifdef AQC_NOINLINE_RARE
#define AQC_RARE __noinline__
#else
#define AQC_RARE __forceinline__
#endif

constexpr unsigned FULL = 0xffffffffu;
#ifndef AQC_WARPS
#define AQC_WARPS 8
#endif
constexpr int WARPS = AQC_WARPS;
constexpr int THREADS = WARPS * 32;
constexpr int NSTAGES = 2;
constexpr int MAX_TILE_PAIRS = 32;
constexpr int QC_CLASSES = 5;              // A, T, C, G, other  (ALL_BASES order, qualitycontrol.py:24)
constexpr uint32_t QC_FLUSH_READS = 4000;  // packed smem word = count(12 bit) << 20 | byte sum (20 bit)

enum Mode { MODE_FILTER = 0, MODE_STAT = 1, MODE_OPS = 2, MODE_LIST = 3 };   // MODE_LIST: filter over a list of single pairs

// lut1[b]: low nibble = comparison code of b as an R1 byte, high nibble = code of COMP[b] (rc side)
//   codes: A0 C1 G2 T3 N4 a8 c9 g10 t11 '\n'12, any other R1 byte 15 (never equal to an rc code)
// lut2[b]: bits0-2 QC class (A0 T1 C2 G3 other4), bit3 = G|C, bits4-5 k-mer code (A0 C1 G2 T3),
//          bit6 = uppercase ACGT, bit7 = member of hasPolyX's polyArray (preprocesser.py:35)
// lut3[b]: COMP[b] as a byte, unknown -> 'N' (util.py:27,47-50)
struct Luts { uint8_t lut1[256], lut2[256], lut3[256]; };

// global accumulators of one QualityControl object
struct QcDev {
    unsigned long long *cls_cnt;    // [5][AQC_MAX_LEN]
    unsigned long long *cls_qsum;   // [5][AQC_MAX_LEN] raw quality-byte sums (33 not subtracted)
    unsigned long long *disc;       // [AQC_MAX_LEN]
    unsigned long long *gchist;     // [AQC_MAX_LEN+1]
    unsigned long long *scal;       // [0] totalKmer, [1] reads
    unsigned long long *kcnt;       // dense 4^k, internal index (plane1bits << k) | plane0bits
    unsigned long long *kfirst;
    unsigned long long *skeys, *scnt, *sfirst;   // side table (non-ACGT k-mers): key, count, first DIRECT sighting
    unsigned long long *sseed;                   // first seeding by a k-mer holding a byte outside util.COMP (see below)
    uint32_t smask;
    uint32_t valid;                 // 0 = this mate is not stat'd in this launch
};

struct KArgs {
    // batch
    const uint8_t *seq1, *qual1, *seq2, *qual2;
    const uint32_t *off1, *off2;
    uint32_t n;
    uint32_t num_tiles;
    uint64_t first_index;
    // tiling
    int tile_pairs;       // pairs per tile (<= 32)
    int col_cap;          // bytes reserved per column per stage
    int max_len;          // longest read in the batch (smem accumulator extent)
    int mode;
    // parameters
    aqc_params p;
    int poly_m;           // hasPolyX screen: consecutive 'same as previous' bits needed; -1 never flagged, 0 always candidate
    // stat gating (MODE_STAT): records with stat_lo <= global < stat_hi; order = order_base + g - stat_lo
    uint64_t stat_lo, stat_hi, order_base;
    // outputs
    aqc_result *results;          // MODE_FILTER
    aqc_ops *ops;                 // MODE_OPS
    unsigned long long *counters; // AQC_C_TOTAL
    QcDev qc[2];                  // [0] mate 1, [1] mate 2 for this launch
    int *error_flag;
    const Luts *luts;
    // list mode (filter): tile i is the single pair list[i], i < *list_count -- the pairs lane_kernel hands over
    const uint32_t *list;
    const uint32_t *list_count;
    int no_stats;         // filter modes: the postfilter statistics of this launch's good pairs are left to stat_kernel<.., POST>
};

// ------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier + 1-D bulk async copy (TMA) + proxy fence
// ------------------------------------------------------------------------------------------
#ifndef AQC_EMU
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// global -> shared bulk copy; dst, src and bytes are multiples of 16
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
#define AQC_DYN_SMEM(name) extern __shared__ __align__(128) uint8_t name[]
// explicit shared-memory word loads (the compiler falls back to generic LD once a pointer went through integer arithmetic)
typedef uint32_t smem_addr_t;
__device__ __forceinline__ smem_addr_t smem_addr(const void *p) { return smem_u32(p); }
__device__ __forceinline__ uint32_t lds_u32(smem_addr_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
// PRMT without __byte_perm's selector sanitising (selector nibbles must be 0..7)
__device__ __forceinline__ uint32_t prmt_raw(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
#else
// host SIMT emulator build (tests/emu, test infrastructure): same contracts, modelled transaction counts
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) { simt::mbar_init(bar, count); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { simt::mbar_arrive_expect_tx(bar, bytes); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) { return simt::mbar_try_wait(bar, parity); }
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) { simt::bulk_g2s(dst, src, bytes, bar); }
__device__ __forceinline__ void fence_proxy_async() {}
__device__ __forceinline__ void fence_mbar_init() {}
#define AQC_DYN_SMEM(name) uint8_t *const name = simt::dyn_smem()
typedef uintptr_t smem_addr_t;
__device__ __forceinline__ smem_addr_t smem_addr(const void *p) { return reinterpret_cast<uintptr_t>(p); }
__device__ __forceinline__ uint32_t lds_u32(smem_addr_t a) { return *reinterpret_cast<const uint32_t *>(a); }
__device__ __forceinline__ uint32_t prmt_raw(uint32_t a, uint32_t b, uint32_t sel) {
    if (sel & 0x8888u) simt::fail("prmt_raw: selector nibble > 7");
    return __byte_perm(a, b, sel);
}
#endif
// 16-byte global load that asks the L2 to bring in the whole 128-byte line: lanes that walk their own read 16 bytes at a time
// come back for the neighbouring sectors a moment later
__device__ __forceinline__ uint4 ldg_stream16(const uint4 *p) {
#ifndef AQC_EMU
    uint4 v;
    asm volatile("ld.global.L2::128B.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
#else
    return *p;
#endif
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) { }
}

// ------------------------------------------------------------------------------------------
// bit-plane helpers.  A "plane set" is uint32_t P[4]; lane j holds positions 32j..32j+31.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lowmask(int nbits) {   // nbits may be <= 0 or >= 32
#if defined(AQC_NO_BMSK) || defined(AQC_EMU)
    return nbits >= 32 ? 0xffffffffu : (nbits <= 0 ? 0u : ((1u << nbits) - 1u));
#else
    uint32_t m;
    asm("bmsk.clamp.b32 %0, %1, %2;" : "=r"(m) : "r"(0u), "r"((uint32_t)max(nbits, 0)));   // width clamps at 32
    return m;
#endif
}

// 32 bits of plane word array `p` starting at bit (32*word + sh), word may exceed 31 (-> zeros)
__device__ __forceinline__ uint32_t plane_window(uint32_t p, int word, int sh) {
    uint32_t a = __shfl_sync(FULL, p, word & 31);
    uint32_t b = __shfl_sync(FULL, p, (word + 1) & 31);
    if (word > 31) a = 0;
    if (word + 1 > 31) b = 0;
    return __funnelshift_r(a, b, sh);
}

// Build the planes of a read from shared-memory bytes.  rev: position i takes byte len-1-i and the
// rc-side code (this yields the planes of reverseComplement(read), util.py:42-51).
// Also returns the exact count of 'N' bytes (nNumber, preprocesser.py:70-76) and whether any
// position has a code >= 4 (then planes 2,3 are needed by the comparisons).
__device__ AQC_RARE void build_planes(const uint8_t *s, int len, bool rev, const uint8_t *lut1, int lane,
                                             uint32_t (&P)[4], int &n_count, bool &exotic) {
    P[0] = P[1] = P[2] = P[3] = 0;
    n_count = 0;
    uint32_t anyx = 0;
    const int nchunks = (len + 31) >> 5;
    for (int c = 0; c < nchunks; c++) {
        int pos = (c << 5) + lane;
        bool valid = pos < len;
        uint32_t byte = valid ? s[rev ? (len - 1 - pos) : pos] : 0u;
        uint32_t l = lut1[byte];
        uint32_t code = valid ? (rev ? (l >> 4) : (l & 15u)) : 0u;
        uint32_t b0 = __ballot_sync(FULL, code & 1u);
        uint32_t b1 = __ballot_sync(FULL, code & 2u);
        uint32_t bx = __ballot_sync(FULL, code & 12u);
        uint32_t bn = __ballot_sync(FULL, valid && byte == 'N');
        n_count += __popc(bn);
        uint32_t b2 = 0, b3 = 0;
        if (bx) {   // warp-uniform
            b2 = __ballot_sync(FULL, code & 4u);
            b3 = __ballot_sync(FULL, code & 8u);
            anyx |= bx;
        }
        if (lane == c) { P[0] = b0; P[1] = b1; P[2] = b2; P[3] = b3; }
    }
    exotic = anyx != 0;
}

// lowQualityNum (preprocesser.py:61-68): number of quality bytes < qual + 33
__device__ __forceinline__ int count_lowq(const uint8_t *q, int len, int thr, int lane) {
    int n = 0;
    for (int base = 0; base < len; base += 32) {
        int pos = base + lane;
        bool hit = pos < len && (int)q[pos] < thr;
        n += __popc(__ballot_sync(FULL, hit));
    }
    return n;
}

// ------------------------------------------------------------------------------------------
// util.overlap_hm (util.py:158-212) on planes.
// scan_dir scans offsets o = 0 .. lenS-31 of the "shifted" read S against the "fixed" read F:
//   compare S[o+i] with F[i], i < ol = min(lenS-o, lenF)
// forward pass: S = r1, F = rc(r2);  reverse pass: S = rc(r2), F = r1 (offset = -o).
// Acceptance (closed form of the leaked loop variable, validated against the reference):
//   mm50 = mismatches among i < min(50, ol);  mm = mismatches among all i < ol
//   accept  <=>  mm50 < 3  and  (mm < 3  or  ol >= 52);   returned diff = mm
// ------------------------------------------------------------------------------------------
template <int NP>
__device__ __forceinline__ int scan_dir(const uint32_t (&S)[4], const uint32_t (&F)[4], int lenS, int lenF, int lane,
                                        int &ol_out, int &mm_out) {
    const int nOff = lenS - 30;     // overlap_require = 30 (util.py:164)
    if (nOff <= 0) return -1;
    uint32_t f0[NP];
#pragma unroll
    for (int k = 0; k < NP; k++) f0[k] = __shfl_sync(FULL, F[k], 0);
    for (int base = 0; base < nOff; base += 32) {
        const int r = base >> 5;
        const int o = base + lane;
        uint32_t x = 0;
#pragma unroll
        for (int k = 0; k < NP; k++) {
            uint32_t a = __shfl_sync(FULL, S[k], r);
            uint32_t b = (r + 1 < 32) ? __shfl_sync(FULL, S[k], (r + 1) & 31) : 0u;
            x |= __funnelshift_r(a, b, lane) ^ f0[k];
        }
        int ol = min(lenS - o, lenF);
        bool surv = (o < nOff) && (__popc(x & lowmask(ol)) < 3);   // first min(32, ol) positions
        uint32_t sv = __ballot_sync(FULL, surv);
        while (sv) {   // warp-uniform loop over survivors in scan order
            int l = __ffs(sv) - 1;
            sv &= sv - 1;
            int oc = base + l;
            int olc = min(lenS - oc, lenF);
            int q = oc >> 5, sh = oc & 31;
            uint32_t xx = 0;
#pragma unroll
            for (int k = 0; k < NP; k++) xx |= plane_window(S[k], lane + q, sh) ^ F[k];
            int lo = lane << 5;
            xx &= lowmask(olc - lo);
            int mm = __reduce_add_sync(FULL, (unsigned)__popc(xx));
            int mm50 = __reduce_add_sync(FULL, (unsigned)__popc(xx & lowmask(min(50, olc) - lo)));
            if (mm50 < 3 && (mm < 3 || olc >= 52)) { ol_out = olc; mm_out = mm; return oc; }
        }
    }
    return -1;
}

template <int NP>
__device__ __forceinline__ void overlap_hm(const uint32_t (&P1)[4], const uint32_t (&RC)[4], int len1, int len2, int lane,
                                           int &offset, int &ol, int &diff) {
    int o = -1, dir = 0;
#pragma unroll 1
    for (dir = 0; dir < 2; dir++) {                                  // forward util.py:172-186, reverse :194-209
        uint32_t S[4], F[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { S[k] = dir ? RC[k] : P1[k]; F[k] = dir ? P1[k] : RC[k]; }
        o = scan_dir<NP>(S, F, dir ? len2 : len1, dir ? len1 : len2, lane, ol, diff);
        if (o >= 0) break;
    }
    if (o >= 0) offset = dir ? -o : o;
    else { offset = 0; ol = 0; diff = 0; }                           // util.py:212
}

// ------------------------------------------------------------------------------------------
// hasPolyX (preprocesser.py:30-51)
// ------------------------------------------------------------------------------------------
// Exact hasPolyX on the raw bytes (taken only by screened candidates).  Returns the char or 0.
__device__ AQC_RARE int polyx_exact(const uint8_t *s, int len, int maxPoly, int mismatch, const uint8_t *lut2, int lane) {
    if (len < maxPoly) return 0;
    const int T = maxPoly - mismatch;
    // first byte outside polyArray aborts the scan (:41-42)
    int limit = len;
    for (int base = 0; base < len; base += 32) {
        int pos = base + lane;
        bool foreign = pos < len && !(lut2[s[pos]] & 0x80);
        uint32_t fb = __ballot_sync(FULL, foreign);
        if (fb) { limit = base + __ffs(fb) - 1; break; }
    }
    for (int base = 0; base < limit; base += 32) {
        int x = base + lane;
        bool flag = false;
        if (x < limit) {
            uint8_t c = s[x];
            int lo = max(0, x - maxPoly + 1);
            int cnt = 0;
            for (int y = x; y >= lo; y--) cnt += (s[y] == c);
            flag = cnt >= T;
        }
        uint32_t fb = __ballot_sync(FULL, flag);
        if (fb) return s[base + __ffs(fb) - 1];
    }
    return 0;
}

// ==========================================================================================
// FAST PATH (reads made of A,C,G,T,N only -- every Illumina read in practice).
// Codes come straight from the ASCII bits: code = (byte >> 1) & 3  ->  A0 C1 T2 G3, complement = code ^ 2
// (flip plane 1).  'N' gets its own plane 2 with planes 0/1 forced to 0.  Each lane converts 8 consecutive
// bases with SWAR arithmetic (one pass covers 256 bases), validates them with two PRMTs (re-encode the 2-bit
// codes to ASCII and compare), and the 8-bit plane slices are gathered into the lane-per-word layout with 4
// shuffles.  Any other byte (lowercase, IUPAC, ...) sends the PAIR to the general LUT/ballot path above.
// ==========================================================================================
__device__ __forceinline__ uint32_t bytemask_lo(int nbytes) {        // 0xFF for the first nbytes bytes (0..4)
    return nbytes >= 4 ? 0xffffffffu : (nbytes <= 0 ? 0u : ((1u << (8 * nbytes)) - 1u));
}

// bytes s[x0 .. x0+7] (zero beyond len) from shared memory at any alignment; vm* = 0xFF per valid byte
__device__ __forceinline__ void load8(const uint8_t *s, int x0, int len, uint32_t &v0, uint32_t &v1, uint32_t &vm0, uint32_t &vm1) {
    v0 = v1 = vm0 = vm1 = 0;
    const int nv = len - x0;
    if (nv > 0) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(s + x0);
        const uint32_t *w = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
        const int sh = (int)(a & 3) * 8;
        const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
        vm0 = bytemask_lo(nv); vm1 = bytemask_lo(nv - 4);
        v0 = __funnelshift_r(w0, w1, sh) & vm0;
        v1 = __funnelshift_r(w1, w2, sh) & vm1;
    }
}

__device__ __forceinline__ uint32_t gather4(uint32_t x01010101) {    // bit 0 of each byte -> 4-bit value
    return (x01010101 * 0x01020408u) >> 24;
}
__device__ __forceinline__ uint32_t hibit_nonzero(uint32_t x) {      // 0x80 in every byte of x that is non-zero
    return (((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;
}

struct FastPlanes {
    uint32_t P[4];      // P[0], P[1] code planes, P[2] = N plane, P[3] unused (0)
    int n_count;        // exact number of 'N' bytes
    bool hasN;          // warp-uniform
    bool exotic;        // warp-uniform: some byte is not A,C,G,T,N -> general path required
};

// Both mates in ONE instruction stream: lanes 0-15 convert mate 1, lanes 16-31 mate 2, 16 consecutive bases per lane
// (256 bases per mate and pass).  Same arithmetic as fast_build on four words; the 16-bit plane slices of lanes
// 2j, 2j+1 form plane word j, so the gather needs two shuffles per mate.  len2 = 0 for single-end input.
// bytemask16: shared-memory table, entry k (0..16) = 16 bytes with the first k bytes 0xFF (built once per CTA)
__device__ __forceinline__ void fast_build2(const uint8_t *r1, int len1, const uint8_t *r2, int len2, int lane,
                                            const uint4 *bytemask16, FastPlanes &F1, FastPlanes &F2) {
    F1.P[0] = F1.P[1] = F1.P[2] = F1.P[3] = 0; F1.n_count = 0; F1.hasN = false; F1.exotic = false;
    F2.P[0] = F2.P[1] = F2.P[2] = F2.P[3] = 0; F2.n_count = 0; F2.hasN = false; F2.exotic = false;
    const bool hi = lane >= 16;
    const int hl = lane & 15;
    const uint8_t *s = hi ? r2 : r1;
    const int len = hi ? len2 : len1;
    const int npass = (max(len1, len2) + 255) >> 8;
    for (int p = 0; p < npass; p++) {
        const int x0 = (p << 8) + (hl << 4);
        uint32_t v[4] = {0, 0, 0, 0}, vm[4] = {0, 0, 0, 0};
        const int nv = len - x0;
        if (nv > 0) {
            const uintptr_t a = reinterpret_cast<uintptr_t>(s + x0);
            const uint32_t *w = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
            const int sh = (int)(a & 3) * 8;
            const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];
            const uint4 mk = bytemask16[min(nv, 16)];
            vm[0] = mk.x; vm[1] = mk.y; vm[2] = mk.z; vm[3] = mk.w;
            v[0] = __funnelshift_r(w0, w1, sh) & vm[0];
            v[1] = __funnelshift_r(w1, w2, sh) & vm[1];
            v[2] = __funnelshift_r(w2, w3, sh) & vm[2];
            v[3] = __funnelshift_r(w3, w4, sh) & vm[3];
        }
        uint32_t t[4], bad = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            t[i] = (v[i] >> 1) & 0x03030303u;
            const uint32_t e = __byte_perm(0x47544341u, 0u, __byte_perm(t[i] | (t[i] >> 4), 0u, 0x4420));
            bad |= (e ^ v[i]) & vm[i];
        }
        uint32_t nbits = 0;
        const uint32_t anybad = __ballot_sync(FULL, bad != 0u);
        if (anybad) {                                              // warp-uniform: some non-ACGT byte in this pass
            uint32_t ex = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t isN = ~hibit_nonzero(v[i] ^ 0x4E4E4E4Eu) & 0x80808080u & vm[i];
                const uint32_t e = __byte_perm(0x47544341u, 0u, __byte_perm(t[i] | (t[i] >> 4), 0u, 0x4420));
                ex |= hibit_nonzero((e ^ v[i]) & vm[i]) & ~isN;
                const uint32_t n01 = isN >> 7;
                t[i] &= ~(n01 * 3u);
                nbits |= gather4(n01) << (4 * i);
            }
            const uint32_t exb = __ballot_sync(FULL, ex != 0u);
            if (exb & 0xFFFFu) F1.exotic = true;
            if (exb >> 16) F2.exotic = true;
            const uint32_t nb = __ballot_sync(FULL, nbits != 0u);
            if (nb) {
                const unsigned c = (unsigned)__popc(nbits);
                if (nb & 0xFFFFu) { F1.hasN = true; F1.n_count += (int)__reduce_add_sync(FULL, hi ? 0u : c); }
                if (nb >> 16) { F2.hasN = true; F2.n_count += (int)__reduce_add_sync(FULL, hi ? c : 0u); }
            }
        }
        uint32_t p0 = 0, p1 = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            p0 |= gather4(t[i] & 0x01010101u) << (4 * i);
            p1 |= gather4((t[i] >> 1) & 0x01010101u) << (4 * i);
        }
        const uint32_t pv = p0 | (p1 << 16);
        const int jj = (lane & 7) << 1;
        const uint32_t a1 = __shfl_sync(FULL, pv, jj), b1 = __shfl_sync(FULL, pv, jj + 1);
        const uint32_t a2 = __shfl_sync(FULL, pv, 16 + jj), b2 = __shfl_sync(FULL, pv, 17 + jj);
        const bool mine = (lane >> 3) == p;
        if (mine) {
            F1.P[0] = __byte_perm(a1, b1, 0x5410); F1.P[1] = __byte_perm(a1, b1, 0x7632);
            F2.P[0] = __byte_perm(a2, b2, 0x5410); F2.P[1] = __byte_perm(a2, b2, 0x7632);
        }
        if (F1.hasN || F2.hasN) {                                  // warp-uniform
            const uint32_t na1 = __shfl_sync(FULL, nbits, jj), nb1 = __shfl_sync(FULL, nbits, jj + 1);
            const uint32_t na2 = __shfl_sync(FULL, nbits, 16 + jj), nb2 = __shfl_sync(FULL, nbits, 17 + jj);
            if (mine) { F1.P[2] = __byte_perm(na1, nb1, 0x5410); F2.P[2] = __byte_perm(na2, nb2, 0x5410); }
        }
    }
}

// planes of reverseComplement(read) from its forward fast planes: rc[i] = comp(fwd[len-1-i])
__device__ __forceinline__ void fast_revcomp(const FastPlanes &F, int len, int lane, uint32_t (&RC)[4]) {
    const int s = 1024 - len;
    const int q = s >> 5, sh = s & 31;
    const int np = F.hasN ? 3 : 2;
    RC[0] = RC[1] = RC[2] = RC[3] = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        if (k < np) {
            uint32_t r = __brev(__shfl_sync(FULL, F.P[k], 31 - lane));
            RC[k] = plane_window(r, lane + q, sh);
        }
    }
    RC[1] ^= lowmask(len - (lane << 5)) & ~RC[2];      // complement = code ^ 2, N stays 0
}

// lowQualityNum on 8 quality bytes per lane (thr in 1..127)
__device__ __forceinline__ int count_lowq_fast(const uint8_t *q, int len, int thr, int lane) {
    const uint32_t t4 = (uint32_t)thr * 0x01010101u;
    int n = 0;
    const int npass = (len + 255) >> 8;
    for (int p = 0; p < npass; p++) {
        uint32_t v0, v1, vm0, vm1;
        load8(q, (p << 8) + (lane << 3), len, v0, v1, vm0, vm1);
        // byte >= thr  <=>  high bit of ((byte | 0x80) - thr) | byte   (thr < 128)
        const uint32_t ge0 = (((v0 | 0x80808080u) - t4) | v0) & 0x80808080u;
        const uint32_t ge1 = (((v1 | 0x80808080u) - t4) | v1) & 0x80808080u;
        n += __popc(~ge0 & 0x80808080u & vm0) + __popc(~ge1 & 0x80808080u & vm1);
    }
    return (int)__reduce_add_sync(FULL, (unsigned)n);
}

// Run-length screen of hasPolyX (preprocesser.py:30-51) on the planes of either code set (np planes).
// A flagged window holds >= T = maxPoly - mismatch equal bases with <= mismatch interruptions, i.e. a run of
// >= R = ceil(T / (mismatch + 1)) identical bases, i.e. m = R - 1 consecutive "same as previous" bits.
// m is computed once on the host (KArgs.poly_m): -1 = never flagged, 0 = every read is a candidate, > 31 = candidate.
__device__ __forceinline__ uint32_t run_within_word(uint32_t y, int m) {   // bit x set <=> y[x..x+m-1] all ones (inside the word)
    if (m == 10) { y &= y >> 1; y &= y >> 2; y &= y >> 4; y &= y >> 2; return y; }    // default -p 35 -a 2
    int t = 1;
    while (t < m) { const int step = min(t, m - t); y &= y >> step; t += step; }
    return y;
}

__device__ __forceinline__ bool polyx_screen_fast(const uint32_t (&P)[4], int np, int len, int maxPoly, int m, int lane) {
    if (len < maxPoly || m < 0) return false;
    if (m == 0 || m > 31) return true;
    uint32_t d = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (k < np) {
            uint32_t up = __shfl_up_sync(FULL, P[k], 1);
            if (lane == 0) up = 0;
            d |= P[k] ^ __funnelshift_l(up, P[k], 1);
        }
    }
    uint32_t y = ~d & lowmask(len - (lane << 5));
    if (lane == 0) y &= ~1u;
    // straddling runs: ones at the top of word j + ones at the bottom of word j+1
    const int lead = __clz((int)~y);
    int trail = __ffs((int)~y) - 1; if (trail < 0) trail = 32;
    int nxt = __shfl_down_sync(FULL, trail, 1);
    if (lane == 31) nxt = 0;
    bool hit = (lead + nxt) >= m;
    hit |= (run_within_word(y, m) != 0u);
    return __ballot_sync(FULL, hit) != 0u;
}

// Both mates in one instruction stream (reads <= 512 bases): lanes 0-15 hold the words of mate 1, lanes 16-31 those of
// mate 2 (forward fast planes A, B with the same np).  Returns bit 0 = mate 1 candidate, bit 1 = mate 2 candidate.
__device__ __forceinline__ uint32_t polyx_screen_pair(const uint32_t (&A)[4], const uint32_t (&B)[4], int np, int len1, int len2,
                                                      int maxPoly, int m, int lane) {
    const bool hi = lane >= 16;
    const int hl = lane & 15;
    const int len = hi ? len2 : len1;
    uint32_t d = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        if (k < np) {
            const uint32_t b = __shfl_sync(FULL, B[k], hl);          // word hl of mate 2 (meaningful for lanes >= 16)
            const uint32_t w = hi ? b : A[k];
            uint32_t up = __shfl_up_sync(FULL, w, 1);
            if (hl == 0) up = 0;
            d |= w ^ __funnelshift_l(up, w, 1);
        }
    }
    uint32_t y = ~d & lowmask(len - (hl << 5));
    if (hl == 0) y &= ~1u;
    const int lead = __clz((int)~y);
    int trail = __ffs((int)~y) - 1; if (trail < 0) trail = 32;
    int nxt = __shfl_down_sync(FULL, trail, 1);
    if (hl == 15) nxt = 0;
    bool hit = (lead + nxt) >= m;
    hit |= (run_within_word(y, m) != 0u);
    hit &= (len >= maxPoly);
    const uint32_t hb = __ballot_sync(FULL, hit);
    return ((hb & 0xFFFFu) ? 1u : 0u) | ((hb >> 16) ? 2u : 0u);
}

// leaner offset scan (same contract as scan_dir) for lenF >= 32
template <int NP>
__device__ __forceinline__ int scan_dir_fast(const uint32_t (&S)[4], const uint32_t (&F)[4], int lenS, int lenF, int lane,
                                             int &ol_out, int &mm_out) {
    const int nOff = lenS - 30;
    if (nOff <= 0) return -1;
    uint32_t f0[NP], a[NP];
#pragma unroll
    for (int k = 0; k < NP; k++) { f0[k] = __shfl_sync(FULL, F[k], 0); a[k] = __shfl_sync(FULL, S[k], 0); }
    const int last = lenS - 31;                         // the only offset whose first window has 31 positions
    for (int base = 0; base < nOff; base += 32) {
        const int o = base + lane;
        const int r1 = (base >> 5) + 1;
        uint32_t x = 0;
#pragma unroll
        for (int k = 0; k < NP; k++) {
            const uint32_t b = __shfl_sync(FULL, S[k], r1 & 31);     // r1 <= 31 because lenS <= 1000
            x |= __funnelshift_r(a[k], b, lane) ^ f0[k];
            a[k] = b;
        }
        if (o == last) x &= 0x7fffffffu;
        uint32_t sv = __ballot_sync(FULL, (o < nOff) && (__popc(x) < 3));
        while (sv) {
            const int l = __ffs(sv) - 1;
            sv &= sv - 1;
            const int oc = base + l;
            const int olc = min(lenS - oc, lenF);
            const int q = oc >> 5, sh = oc & 31;
            uint32_t xx = 0;
#pragma unroll
            for (int k = 0; k < NP; k++) xx |= plane_window(S[k], lane + q, sh) ^ F[k];
            const int lo = lane << 5;
            xx &= lowmask(olc - lo);
            const int mm = __reduce_add_sync(FULL, (unsigned)__popc(xx));
            const int mm50 = __reduce_add_sync(FULL, (unsigned)__popc(xx & lowmask(min(50, olc) - lo)));
            if (mm50 < 3 && (mm < 3 || olc >= 52)) { ol_out = olc; mm_out = mm; return oc; }
        }
    }
    return -1;
}

template <int NP>
__device__ __forceinline__ void overlap_fast(const uint32_t (&P1)[4], const uint32_t (&RC)[4], int len1, int len2, int lane,
                                             int &offset, int &ol, int &diff) {
    int o = scan_dir_fast<NP>(P1, RC, len1, len2, lane, ol, diff);      // forward  util.py:172-186
    if (o >= 0) { offset = o; return; }
    o = scan_dir_fast<NP>(RC, P1, len2, len1, lane, ol, diff);          // reverse  util.py:194-209
    if (o >= 0) { offset = -o; return; }
    offset = 0; ol = 0; diff = 0;
}

// np = 2 (ACGT only), 3 (fast codes with N plane) or 4 (general LUT codes).  Mates shorter than 32 bases (rare) and
// the LUT-coded pairs share the one general 4-plane scan (unused planes are zero on both sides).
__device__ __forceinline__ void overlap_np(int np, const uint32_t (&P1)[4], const uint32_t (&RC)[4], int len1, int len2,
                                           int lane, int &offset, int &ol, int &diff) {
    if (np == 4 || len1 < 32 || len2 < 32) overlap_hm<4>(P1, RC, len1, len2, lane, offset, ol, diff);
    else if (np == 2) overlap_fast<2>(P1, RC, len1, len2, lane, offset, ol, diff);
    else overlap_fast<3>(P1, RC, len1, len2, lane, offset, ol, diff);
}

// ------------------------------------------------------------------------------------------
// shared-memory QC accumulators of one CTA (flushed to QcDev with 64-bit atomics)
// ------------------------------------------------------------------------------------------
struct QcSmem {
    uint32_t *acc;    // [2 mates][5 classes][max_len]  count << 20 | byte sum
    uint32_t *disc;   // [2 mates][max_len]
    int max_len;
};

// Flush of the packed accumulators by ONE warp while the CTA's other warps keep adding: every word is exchanged with 0.
__device__ __forceinline__ void qc_flush_warp(const QcDev (&qc)[2], uint32_t *s_acc, uint32_t *s_disc, int max_len, int lane) {
    for (int m = 0; m < 2; m++) {
        const QcDev &qd = qc[m];
        if (!qd.valid) continue;
        for (int i = lane; i < QC_CLASSES * max_len; i += 32) {
            const uint32_t v = atomicExch(&s_acc[m * QC_CLASSES * max_len + i], 0u);
            if (v) {
                const int c = i / max_len, pos = i - c * max_len;
                atomicAdd(&qd.cls_cnt[c * AQC_MAX_LEN + pos], (unsigned long long)(v >> 20));
                atomicAdd(&qd.cls_qsum[c * AQC_MAX_LEN + pos], (unsigned long long)(v & 0xFFFFFu));
            }
        }
        for (int i = lane; i < max_len; i += 32) {
            const uint32_t v = atomicExch(&s_disc[m * max_len + i], 0u);
            if (v) atomicAdd(&qd.disc[i], (unsigned long long)v);
        }
    }
}

// Final flush by the whole CTA (after a __syncthreads: nobody adds any more).
__device__ __forceinline__ void qc_flush_cta(const QcDev (&qc)[2], const uint32_t *s_acc, const uint32_t *s_disc, int max_len, int tid, unsigned nthreads) {
    for (int m = 0; m < 2; m++) {
        const QcDev &qd = qc[m];
        if (!qd.valid) continue;
        for (int i = tid; i < QC_CLASSES * max_len; i += nthreads) {
            const uint32_t v = s_acc[m * QC_CLASSES * max_len + i];
            if (v) {
                const int c = i / max_len, pos = i - c * max_len;
                atomicAdd(&qd.cls_cnt[c * AQC_MAX_LEN + pos], (unsigned long long)(v >> 20));
                atomicAdd(&qd.cls_qsum[c * AQC_MAX_LEN + pos], (unsigned long long)(v & 0xFFFFFu));
            }
        }
        for (int i = tid; i < max_len; i += nthreads) {
            const uint32_t v = s_disc[m * max_len + i];
            if (v) atomicAdd(&qd.disc[i], (unsigned long long)v);
        }
    }
}

__device__ __forceinline__ unsigned long long side_hash(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return k;
}

// insert-or-find `key` in the side table; returns slot or -1 on overflow
__device__ __forceinline__ int side_slot(const QcDev &q, unsigned long long key) {
    uint32_t h = (uint32_t)side_hash(key) & q.smask;
    for (uint32_t probe = 0; probe <= q.smask; probe++) {
        unsigned long long prev = atomicCAS(&q.skeys[h], AQC_KMER_NEVER, key);
        if (prev == AQC_KMER_NEVER || prev == key) return (int)h;
        h = (h + 1) & q.smask;
    }
    return -1;
}

__device__ __forceinline__ void first_min(unsigned long long *addr, unsigned long long key) {
    if (*((volatile unsigned long long *)addr) > key) atomicMin(addr, key);
}

// QualityControl.statRead (qualitycontrol.py:73-122) for one read, warp-cooperative.
// s/qv: shared-memory bytes of the (trimmed, corrected) read; mate selects the accumulator half.
//
// k-mer insertion order (quirk Q12).  The reference inserts a k-mer X at its first direct sighting and, only at
// that NEW insertion, seeds reverseComplement(X) with 0.  For k-mers over util.COMP's alphabet revcomp is an
// involution and "first = min(direct sightings of X, sightings of rc(X) | 1)" is exact; the dense table stores
// exactly that.  A k-mer holding a byte outside COMP maps that byte to 'N' (util.py:47-50), so revcomp is not
// invertible there: such a k-mer has no pre-image (it is always newly inserted by its first sighting and always
// seeds), while a k-mer X over COMP's alphabet seeds rc(X) only if X itself was not seeded earlier.  The side
// table therefore keeps the first direct sighting (sfirst) and the first seeding by a foreign-byte k-mer (sseed)
// apart, guarantees a slot for rc(X), and aqc_get_kmer_side resolves the partner rule on the host.
__device__ AQC_RARE void stat_read(const uint8_t *s, const uint8_t *qv, int len, int mate, uint64_t order,
                                          const QcSmem &sm, const QcDev &qd, const uint8_t *lut1, const uint8_t *lut2, const uint8_t *lut3,
                                          int K, int lane, int *error_flag) {
    if (len <= 0) {     // an empty read runs no loop of statRead but is still counted: gcHistogram[0] += 1 (qualitycontrol.py:112)
        if (lane == 0) { atomicAdd(&qd.gchist[0], 1ULL); atomicAdd(&qd.scal[1], 1ULL); }
        return;
    }
    if (len < 5) { if (lane == 0) atomicExch(error_flag, AQC_ERR_TOO_SHORT_STAT); return; }
    uint32_t *acc = sm.acc + (size_t)mate * QC_CLASSES * sm.max_len;
    uint32_t *dsc = sm.disc + (size_t)mate * sm.max_len;
    int gc = 0;
    const int nk = len - K;                         // k-mers start at i < len - K (quirk Q11)
    const uint32_t km = (K >= 32) ? 0xffffffffu : ((1u << K) - 1u);
    // k-mer plane words of the previous chunk (positions base-32 .. base-1)
    uint32_t pk0 = 0, pk1 = 0, pkv = 0;
    // software pipeline of the dense first-seen stamps: the load issued for chunk c-1 is consumed one iteration later
    unsigned long long pend_val = 0, pend_when = 0;
    uint32_t pend_idx = 0;
    bool pend = false;
    const int nchunks = (len + 31) >> 5;
    for (int c = 0; c <= nchunks + 1; c++) {        // +1: k-mers of the last chunk, +1: drain the stamp pipeline
        const int base = c << 5;
        const int pos = base + lane;
        // ---- stage C: finish the stamp update whose load was issued in the previous iteration ----
        if (pend && pend_val > pend_when) atomicMin(&qd.kfirst[pend_idx], pend_when);
        pend = false;
        if (c > nchunks) break;
        uint32_t k0 = 0, k1 = 0, kv = 0;
        if (c < nchunks) {
            bool valid = pos < len;
            uint32_t b = valid ? s[pos] : 0u;
            uint32_t l2 = valid ? lut2[b] : 4u;
            if (valid) {
                uint32_t q = qv[pos];
                // discontinuity window (:97-108)
                int left = pos - 2;
                if (left < 0) left = 0;
                else if (pos + 3 >= len) left = len - 5;
                uint32_t c0 = s[left], c1 = s[left + 1], c2 = s[left + 2], c3 = s[left + 3], c4 = s[left + 4];
                uint32_t d = (c0 != c1) + (c1 != c2) + (c2 != c3) + (c3 != c4);
                atomicAdd(&acc[(l2 & 7u) * sm.max_len + pos], (1u << 20) | q);
                if (d) atomicAdd(&dsc[pos], d);
            }
            gc += __popc(__ballot_sync(FULL, valid && (l2 & 8u)));
            k0 = __ballot_sync(FULL, valid && (l2 & 0x10u));
            k1 = __ballot_sync(FULL, valid && (l2 & 0x20u));
            kv = __ballot_sync(FULL, valid && (l2 & 0x40u));
        }
        if (c > 0) {
            // ---- stage B: k-mers starting in the previous chunk: i = base - 32 + lane ----
            const int i = base - 32 + lane;
            if (i < nk) {
                uint32_t w0 = __funnelshift_r(pk0, k0, lane) & km;
                uint32_t w1 = __funnelshift_r(pk1, k1, lane) & km;
                uint32_t wv = __funnelshift_r(pkv, kv, lane) & km;
                unsigned long long when = (order << 11) | ((unsigned long long)i << 1);
                if (wv == km) {
                    // dense table: count + DIRECT first sighting only; the stamp a k-mer gets from being seeded by its
                    // reverse complement is derived at fetch: first[X] = min(direct[X], direct[rc(X)] | 1)
                    const uint32_t idx = (w1 << K) | w0;
                    pend_val = __ldcg(&qd.kfirst[idx]);
                    pend_idx = idx; pend_when = when; pend = true;
                    atomicAdd(&qd.kcnt[idx], 1ULL);
                } else {
                    unsigned long long key = 0, rkey = 0;
                    bool foreign = false;
                    for (int j = 0; j < K; j++) {
                        unsigned long long bj = s[i + j];
                        key = (key << 8) | bj;
                        rkey |= (unsigned long long)lut3[bj] << (8 * j);
                        foreign |= (lut1[bj] & 15u) == 15u;
                    }
                    if (key == AQC_KMER_NEVER || rkey == AQC_KMER_NEVER) atomicExch(error_flag, AQC_ERR_INVALID);
                    else {
                        int h = side_slot(qd, key);
                        int hr = side_slot(qd, rkey);
                        if (h < 0 || hr < 0) atomicExch(error_flag, AQC_ERR_KMER_TABLE_FULL);
                        else {
                            atomicAdd(&qd.scnt[h], 1ULL);
                            first_min(&qd.sfirst[h], when);
                            if (foreign) first_min(&qd.sseed[hr], when | 1ULL);
                        }
                    }
                }
            }
        }
        pk0 = k0; pk1 = k1; pkv = kv;
    }
    if (lane == 0) {
        atomicAdd(&qd.gchist[gc], 1ULL);            // :112
        if (nk > 0) atomicAdd(&qd.scal[0], (unsigned long long)nk);   // totalKmer :114
        atomicAdd(&qd.scal[1], 1ULL);
    }
}

}  // namespace aqc
