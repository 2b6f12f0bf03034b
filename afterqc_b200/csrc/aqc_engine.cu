// aqc_engine.cu -- C-ABI of libafterqc_b200.so (include/afterqc_b200.h) on top of lane_kernel / stat_kernel / pair_kernel.
//
// Host responsibilities only: context + device buffers, tile sizing, the chunked
// H2D -> kernel -> D2H pipeline for host-resident batches, counter fetch/convert.
// There is NO CPU implementation of the hot path in this library.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <array>
#include <algorithm>
#include "aqc_kernel.cuh"
#include "aqc_lane_kernel.cuh"
#include "aqc_stat_kernel.cuh"
#include "aqc_edit.cuh"
#include "aqc_parse.cuh"
#include <mutex>

using namespace aqc;

namespace {

thread_local char g_create_err[256] = "";

struct QcHost {      // owning handles of one QC slot's device arrays
    QcDev d;
    size_t dense_n;
    uint32_t side_cap;
};

struct Staging {     // device staging of one host chunk
    uint8_t *col[4] = {nullptr, nullptr, nullptr, nullptr};
    uint32_t *off[2] = {nullptr, nullptr};
    void *res = nullptr;
    size_t col_cap = 0, off_cap = 0, res_cap = 0;
    cudaEvent_t h2d_done = nullptr, k_done = nullptr, d2h_done = nullptr;
};

}  // namespace

struct aqc_ctx {
    int device = 0;
    aqc_params p;
    int sm_count = 148;
    size_t max_dyn_smem = 200 * 1024;
    cudaStream_t compute = nullptr, copy_in = nullptr, copy_out = nullptr;
    cudaStream_t own_compute = nullptr;
    unsigned long long *d_counters = nullptr;
    QcHost qc[AQC_NUM_QC];
    Luts *d_luts = nullptr;
    int *d_error = nullptr;
    uint32_t *d_maxlen = nullptr;
    // lane-per-pair filter path (aqc_lane_kernel.cuh): hand-over list of the pairs that need the general kernel
    uint32_t *d_fb_list = nullptr, *d_fb_count = nullptr;
    size_t fb_cap = 0;
    uint32_t *d_kbits = nullptr;   // stat_kernel: "already stamped" bitmaps of the two mates of a launch (aqc_stat_kernel.cuh)
    // aqc_fastq_parse_device (aqc_parse.cuh): per slot (one per mate) the device buffers of the last parse, grown on demand
    struct ParseSlot {
        uint8_t *text = nullptr, *seq = nullptr, *qual = nullptr;
        uint32_t *blk = nullptr, *nl = nullptr, *line_start = nullptr, *line_len = nullptr, *rec = nullptr, *part = nullptr, *flags = nullptr;
        size_t text_cap = 0, col_cap = 0, blk_cap = 0, nl_cap = 0, line_cap = 0, line2_cap = 0, rec_cap = 0, part_cap = 0, flags_cap = 0;
    } parse[2];
    Staging stg[2];
    uint32_t chunk_pairs = 1u << 18;
    uint32_t stat_head = 8192;          // records whose dense k-mers stamp_head_kernel stamps before stat_kernel runs (AQC_STAT_HEAD)
    std::vector<std::array<cudaEvent_t, 4>> ev_pool;    // per timed launch group: start | filter kernel done | list mode done | statistics done
    size_t ev_used = 0;
    uint64_t launches = 0;
    int sticky_err = 0;
    char err[256] = "";
};

namespace {

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            snprintf(ctx->err, sizeof ctx->err, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return AQC_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

int fail(aqc_ctx *ctx, int code, const char *msg) {
    snprintf(ctx->err, sizeof ctx->err, "%s", msg);
    return code;
}

void fill_luts(Luts &L) {
    for (int b = 0; b < 256; b++) {
        uint8_t c1 = 15, crc = 4, comp = 'N';
        switch (b) {
            case 'A': c1 = 0; crc = 3; comp = 'T'; break;
            case 'C': c1 = 1; crc = 2; comp = 'G'; break;
            case 'G': c1 = 2; crc = 1; comp = 'C'; break;
            case 'T': c1 = 3; crc = 0; comp = 'A'; break;
            case 'N': c1 = 4; crc = 4; comp = 'N'; break;
            case 'a': c1 = 8; crc = 11; comp = 't'; break;
            case 'c': c1 = 9; crc = 10; comp = 'g'; break;
            case 'g': c1 = 10; crc = 9; comp = 'c'; break;
            case 't': c1 = 11; crc = 8; comp = 'a'; break;
            case '\n': c1 = 12; crc = 12; comp = '\n'; break;
            default: break;
        }
        L.lut1[b] = (uint8_t)(c1 | (crc << 4));
        L.lut3[b] = comp;
        uint8_t l2 = 4;   // class "other"
        switch (b) {
            case 'A': l2 = 0 | (0 << 4) | 0x40; break;
            case 'T': l2 = 1 | (3 << 4) | 0x40; break;
            case 'C': l2 = 2 | 8 | (1 << 4) | 0x40; break;
            case 'G': l2 = 3 | 8 | (2 << 4) | 0x40; break;
            default: break;
        }
        if (b == 'A' || b == 'T' || b == 'C' || b == 'G' || b == 'a' || b == 't' || b == 'c' || b == 'g' || b == 'N') l2 |= 0x80;
        L.lut2[b] = l2;
    }
}

unsigned long long host_side_hash(unsigned long long k) {   // must match aqc::side_hash
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return k;
}

int check_params(const aqc_params *p, char *err, size_t errn) {
    if (p->qc_kmer < 1 || p->qc_kmer > AQC_MAX_KMER) { snprintf(err, errn, "qc_kmer %d outside 1..%d", p->qc_kmer, AQC_MAX_KMER); return AQC_ERR_INVALID; }
    if (p->filter_kernel < 0 || p->filter_kernel > 2) { snprintf(err, errn, "filter_kernel %d outside 0..2", p->filter_kernel); return AQC_ERR_INVALID; }
    if (p->stat_kernel < 0 || p->stat_kernel > 1) { snprintf(err, errn, "stat_kernel %d outside 0..1", p->stat_kernel); return AQC_ERR_INVALID; }
    if (p->trim_front < 0 || p->trim_tail < 0 || p->trim_front2 < 0 || p->trim_tail2 < 0) { snprintf(err, errn, "negative trim value (resolve auto-trim on the host first)"); return AQC_ERR_INVALID; }
    return 0;
}

size_t smem_bytes_for(int P, int col_cap, int max_len) {
    int off_cap = ((P + 8) * 4 + 15) & ~15;
    size_t stage = ((size_t)4 * col_cap + 2 * off_cap + 127) & ~(size_t)127;
    size_t acc = (size_t)(2 * QC_CLASSES * max_len + 2 * max_len + 2 * (max_len + 1)) * 4;
    return NSTAGES * stage + 768 + 272 + acc + 64;
}

// A launchable kernel handle.  On the GPU this is the __global__ function itself; in the host SIMT emulator build
// (tests/emu, test infrastructure) it is a trampoline that unpacks cudaLaunchKernel's argument array.
#ifndef AQC_EMU
#define AQC_KERNEL_HANDLE(...) ((const void *)(__VA_ARGS__))
#else
template <int MODE, bool PAIRED> void emu_pair_kernel(void **a) { pair_kernel<MODE, PAIRED>(*(const KArgs *)a[0]); }
void emu_edit_distance_kernel(void **a) {
    edit_distance_kernel(*(const uint8_t **)a[0], *(const uint32_t **)a[1], *(const uint8_t **)a[2], *(const uint32_t **)a[3], *(uint32_t *)a[4], *(int32_t **)a[5]);
}
void emu_newline_count_kernel(void **a) { newline_count_kernel(*(const ParseArgs *)a[0]); }
void emu_newline_write_kernel(void **a) { newline_write_kernel(*(const ParseArgs *)a[0]); }
void emu_record_kernel(void **a) { record_kernel(*(const ParseArgs *)a[0]); }
void emu_gather_kernel(void **a) { gather_kernel(*(const ParseArgs *)a[0]); }
void emu_scan_partial_kernel(void **a) { scan_partial_kernel(*(const ScanArgs *)a[0]); }
void emu_scan_single_kernel(void **a) { scan_single_kernel(*(const ScanArgs *)a[0]); }
void emu_scan_final_kernel(void **a) { scan_final_kernel(*(const ScanArgs *)a[0]); }
void emu_maxlen_kernel(void **a) { maxlen_kernel(*(const uint32_t **)a[0], *(const uint32_t **)a[1], *(uint32_t *)a[2], *(uint32_t **)a[3]); }
#define pair_kernel emu_pair_kernel
#define AQC_KERNEL_HANDLE(...) ((const void *)(simt::Entry)(__VA_ARGS__))
#endif

const void *kernel_for(int mode, bool paired) {
    if (mode == MODE_FILTER) return paired ? AQC_KERNEL_HANDLE(pair_kernel<MODE_FILTER, true>) : AQC_KERNEL_HANDLE(pair_kernel<MODE_FILTER, false>);
    if (mode == MODE_LIST) return paired ? AQC_KERNEL_HANDLE(pair_kernel<MODE_LIST, true>) : AQC_KERNEL_HANDLE(pair_kernel<MODE_LIST, false>);
    if (mode == MODE_STAT) return paired ? AQC_KERNEL_HANDLE(pair_kernel<MODE_STAT, true>) : AQC_KERNEL_HANDLE(pair_kernel<MODE_STAT, false>);
    return paired ? AQC_KERNEL_HANDLE(pair_kernel<MODE_OPS, true>) : AQC_KERNEL_HANDLE(pair_kernel<MODE_OPS, false>);
}
#ifdef AQC_EMU
#undef pair_kernel
template <bool PAIRED, int NW> void emu_lane_kernel(void **a) { lane_kernel<PAIRED, NW>(*(const LArgs *)a[0]); }
template <bool PAIRED, int NW, bool POST> void emu_stat_kernel(void **a) { stat_kernel<PAIRED, NW, POST>(*(const SKArgs *)a[0]); }
void emu_stamp_bits_kernel(void **a) {
    stamp_bits_kernel(*(const unsigned long long **)a[0], *(const unsigned long long **)a[1], *(uint32_t *)a[2], *(unsigned long long *)a[3],
                      *(uint32_t **)a[4], *(uint32_t **)a[5], *(uint32_t **)a[6]);
}
template <bool PAIRED, bool POST> void emu_stamp_head_kernel(void **a) { stamp_head_kernel<PAIRED, POST>(*(const SKArgs *)a[0]); }
#define stamp_head_kernel emu_stamp_head_kernel
#define lane_kernel emu_lane_kernel
#define stat_kernel emu_stat_kernel
#endif

// lane-per-pair filter kernel / warp-per-read statistics kernel for mates of at most 32*NW bases
int lane_words_for(int max_len) { return max_len <= 128 ? 4 : (max_len <= 160 ? 5 : (max_len <= 256 ? 8 : 0)); }
const void *lane_kernel_for(bool paired, int nw) {
    if (nw == 4) return paired ? AQC_KERNEL_HANDLE(lane_kernel<true, 4>) : AQC_KERNEL_HANDLE(lane_kernel<false, 4>);
    if (nw == 5) return paired ? AQC_KERNEL_HANDLE(lane_kernel<true, 5>) : AQC_KERNEL_HANDLE(lane_kernel<false, 5>);
    return paired ? AQC_KERNEL_HANDLE(lane_kernel<true, 8>) : AQC_KERNEL_HANDLE(lane_kernel<false, 8>);
}
// post: the sampled good pairs of a filter launch, from their records; else the prefilter window of aqc_stat_reads
const void *stat_kernel_for(bool paired, int nw, bool post) {
#define AQC_STAT_ROW(NWV)                                                                                                         \
    if (nw == NWV) return paired ? (post ? AQC_KERNEL_HANDLE(stat_kernel<true, NWV, true>) : AQC_KERNEL_HANDLE(stat_kernel<true, NWV, false>)) \
                                 : (post ? AQC_KERNEL_HANDLE(stat_kernel<false, NWV, true>) : AQC_KERNEL_HANDLE(stat_kernel<false, NWV, false>));
    AQC_STAT_ROW(4) AQC_STAT_ROW(5)
    return paired ? (post ? AQC_KERNEL_HANDLE(stat_kernel<true, 8, true>) : AQC_KERNEL_HANDLE(stat_kernel<true, 8, false>))
                  : (post ? AQC_KERNEL_HANDLE(stat_kernel<false, 8, true>) : AQC_KERNEL_HANDLE(stat_kernel<false, 8, false>));
#undef AQC_STAT_ROW
}
const void *stamp_head_kernel_for(bool paired, bool post) {
    return paired ? (post ? AQC_KERNEL_HANDLE(stamp_head_kernel<true, true>) : AQC_KERNEL_HANDLE(stamp_head_kernel<true, false>))
                  : (post ? AQC_KERNEL_HANDLE(stamp_head_kernel<false, true>) : AQC_KERNEL_HANDLE(stamp_head_kernel<false, false>));
}
#ifdef AQC_EMU
#undef lane_kernel
#undef stat_kernel
#undef stamp_head_kernel
#endif

int alloc_qc(aqc_ctx *ctx, QcHost &q) {
    q.dense_n = (size_t)1 << (2 * ctx->p.qc_kmer);
    int lg = ctx->p.kmer_side_log2 > 0 ? ctx->p.kmer_side_log2 : 20;
    if (lg < 4 || lg > 30) return fail(ctx, AQC_ERR_INVALID, "kmer_side_log2 out of range");
    q.side_cap = 1u << lg;
    q.d.smask = q.side_cap - 1;
    q.d.valid = 1;
    CK(cudaMalloc(&q.d.cls_cnt, sizeof(unsigned long long) * QC_CLASSES * AQC_MAX_LEN));
    CK(cudaMalloc(&q.d.cls_qsum, sizeof(unsigned long long) * QC_CLASSES * AQC_MAX_LEN));
    CK(cudaMalloc(&q.d.disc, sizeof(unsigned long long) * AQC_MAX_LEN));
    CK(cudaMalloc(&q.d.gchist, sizeof(unsigned long long) * (AQC_MAX_LEN + 1)));
    CK(cudaMalloc(&q.d.scal, sizeof(unsigned long long) * 2));
    CK(cudaMalloc(&q.d.kcnt, sizeof(unsigned long long) * q.dense_n));
    CK(cudaMalloc(&q.d.kfirst, sizeof(unsigned long long) * q.dense_n));
    CK(cudaMalloc(&q.d.skeys, sizeof(unsigned long long) * q.side_cap));
    CK(cudaMalloc(&q.d.scnt, sizeof(unsigned long long) * q.side_cap));
    CK(cudaMalloc(&q.d.sfirst, sizeof(unsigned long long) * q.side_cap));
    CK(cudaMalloc(&q.d.sseed, sizeof(unsigned long long) * q.side_cap));
    return 0;
}

int zero_qc(aqc_ctx *ctx, QcHost &q) {
    cudaStream_t s = ctx->compute;
    CK(cudaMemsetAsync(q.d.cls_cnt, 0, sizeof(unsigned long long) * QC_CLASSES * AQC_MAX_LEN, s));
    CK(cudaMemsetAsync(q.d.cls_qsum, 0, sizeof(unsigned long long) * QC_CLASSES * AQC_MAX_LEN, s));
    CK(cudaMemsetAsync(q.d.disc, 0, sizeof(unsigned long long) * AQC_MAX_LEN, s));
    CK(cudaMemsetAsync(q.d.gchist, 0, sizeof(unsigned long long) * (AQC_MAX_LEN + 1), s));
    CK(cudaMemsetAsync(q.d.scal, 0, sizeof(unsigned long long) * 2, s));
    CK(cudaMemsetAsync(q.d.kcnt, 0, sizeof(unsigned long long) * q.dense_n, s));
    CK(cudaMemsetAsync(q.d.kfirst, 0xFF, sizeof(unsigned long long) * q.dense_n, s));
    CK(cudaMemsetAsync(q.d.skeys, 0xFF, sizeof(unsigned long long) * q.side_cap, s));
    CK(cudaMemsetAsync(q.d.scnt, 0, sizeof(unsigned long long) * q.side_cap, s));
    CK(cudaMemsetAsync(q.d.sfirst, 0xFF, sizeof(unsigned long long) * q.side_cap, s));
    CK(cudaMemsetAsync(q.d.sseed, 0xFF, sizeof(unsigned long long) * q.side_cap, s));
    return 0;
}

void free_qc(QcHost &q) {
    cudaFree(q.d.cls_cnt); cudaFree(q.d.cls_qsum); cudaFree(q.d.disc); cudaFree(q.d.gchist); cudaFree(q.d.scal);
    cudaFree(q.d.kcnt); cudaFree(q.d.kfirst); cudaFree(q.d.skeys); cudaFree(q.d.scnt); cudaFree(q.d.sfirst); cudaFree(q.d.sseed);
}

int poll_error(aqc_ctx *ctx) {   // requires the compute stream to be idle
    int e = 0;
    CK(cudaMemcpy(&e, ctx->d_error, sizeof e, cudaMemcpyDeviceToHost));
    if (e && !ctx->sticky_err) {
        ctx->sticky_err = e;
        const char *m = e == AQC_ERR_TOO_LONG ? "a read is longer than AQC_MAX_LEN" :
                        e == AQC_ERR_KMER_TABLE_FULL ? "non-ACGT k-mer side table full (raise kmer_side_log2)" :
                        e == AQC_ERR_TOO_SHORT_STAT ? "a read of 1..4 bases reached statRead (the reference raises IndexError)" :
                        "device-side domain error";
        snprintf(ctx->err, sizeof ctx->err, "%s", m);
    }
    return ctx->sticky_err;
}

struct DevBatch {    // device-visible view of a (chunk of a) batch
    const uint8_t *seq1, *qual1, *seq2, *qual2;
    const uint32_t *off1, *off2;
    uint32_t n;
    uint64_t first_index;
    int max_len;
    bool qual2_in_sysmem = false;    // qual2 points into page-locked host memory: plain loads only, no bulk copies from it
};

struct LaunchExtra {
    int mode;
    int qc1, qc2;                 // slots (or -1)
    uint64_t stat_lo, stat_hi, order_base;
    void *out;                    // results / ops (device)
};

// tile geometry of pair_kernel for a batch whose longest read is maxl: fills tile_pairs / col_cap / num_tiles, returns smem bytes
size_t pair_tiling(aqc_ctx *ctx, KArgs &A, uint32_t n_tiles_of, int maxl, int cap_pairs) {
    // tile: <= 32 pairs and <= ~24 KB of column bytes per mate column group
    int P = std::min(MAX_TILE_PAIRS, std::max(1, (24 * 1024) / (4 * maxl)));
    P = std::min(P, cap_pairs);
    if (const char *tp = getenv("AQC_TILE_PAIRS")) P = std::max(1, std::min(P, atoi(tp)));   // tuning knob
    // small batches: shrink tiles so that every SM gets work
    while (P > 8 && (n_tiles_of + P - 1) / P < (uint32_t)ctx->sm_count * 2) P >>= 1;
    size_t smem;
    for (;;) {
        A.tile_pairs = P; A.col_cap = (P * maxl + 32 + 15) & ~15; A.num_tiles = (n_tiles_of + P - 1) / P;
        smem = smem_bytes_for(P, A.col_cap, maxl);
        if (smem <= ctx->max_dyn_smem || P == 1) break;
        P >>= 1;
    }
    return smem;
}

// aqc_params.filter_kernel: 0 / 2 = lane-per-pair kernel for batches of short reads (pair_kernel otherwise), 1 = pair_kernel always
bool lane_path(const aqc_ctx *ctx, int mode, int max_len) {
    return mode == MODE_FILTER && ctx->p.filter_kernel != 1 && lane_words_for(std::max(max_len, 8)) != 0;
}
// aqc_params.stat_kernel: 0 = stat_kernel (shared-memory histograms) for batches of short reads, 1 = stat_read inside pair_kernel always
bool stat_path(const aqc_ctx *ctx, int max_len) {
    return ctx->p.stat_kernel != 1 && lane_words_for(std::max(max_len, 8)) != 0;
}

// device-side address of a page-locked host range, or nullptr when the device cannot address all of it.  The kernels read
// the column in aligned 16-byte pieces, so the range is widened to 16-byte boundaries first: a registration that ends (or
// starts) inside such a piece is refused and the caller copies the column as usual.
const uint8_t *device_view_of_host(const uint8_t *p, size_t first, size_t last) {
    if (!p || last < first) return nullptr;
    const uintptr_t lo = (reinterpret_cast<uintptr_t>(p) + first) & ~(uintptr_t)15;
    const uintptr_t hi = ((reinterpret_cast<uintptr_t>(p) + last) | (uintptr_t)15);
    cudaPointerAttributes a0, a1;
    if (cudaPointerGetAttributes(&a0, reinterpret_cast<const void *>(lo)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (cudaPointerGetAttributes(&a1, reinterpret_cast<const void *>(hi)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (a0.type != cudaMemoryTypeHost || a1.type != cudaMemoryTypeHost || !a0.devicePointer || !a1.devicePointer) return nullptr;
    const uint8_t *d0 = (const uint8_t *)a0.devicePointer, *d1 = (const uint8_t *)a1.devicePointer;
    if ((size_t)(d1 - d0) != (size_t)(hi - lo)) return nullptr;      // not one mapping
    return d0 + (reinterpret_cast<uintptr_t>(p) + first - lo) - first;        // so that view + first is the device address of p + first
}

int launch(aqc_ctx *ctx, const DevBatch &b, const LaunchExtra &x, cudaStream_t stream, bool timed) {
    if (b.n == 0) return 0;
    if (b.max_len > AQC_MAX_LEN) { ctx->sticky_err = AQC_ERR_TOO_LONG; return fail(ctx, AQC_ERR_TOO_LONG, "a read is longer than AQC_MAX_LEN"); }
    KArgs A;
    memset(&A, 0, sizeof A);
    A.seq1 = b.seq1; A.qual1 = b.qual1; A.seq2 = b.seq2; A.qual2 = b.qual2; A.off1 = b.off1; A.off2 = b.off2;
    A.n = b.n; A.first_index = b.first_index;
    const int maxl = std::max(b.max_len, 8);
    A.max_len = maxl;
    A.mode = x.mode;
    A.p = ctx->p;
    A.stat_lo = x.stat_lo; A.stat_hi = x.stat_hi; A.order_base = x.order_base;
    A.results = x.mode == MODE_FILTER ? (aqc_result *)x.out : nullptr;
    A.ops = x.mode == MODE_OPS ? (aqc_ops *)x.out : nullptr;
    A.counters = ctx->d_counters;
    A.error_flag = ctx->d_error;
    A.luts = ctx->d_luts;
    for (int m = 0; m < 2; m++) {
        int slot = m == 0 ? x.qc1 : x.qc2;
        if (slot >= 0) { A.qc[m] = ctx->qc[slot].d; A.qc[m].valid = 1; }
        else { A.qc[m] = ctx->qc[0].d; A.qc[m].valid = 0; }
    }
    {   // hasPolyX screen constant (see polyx_screen_fast)
        const int mismatch = ctx->p.allow_mismatch_in_poly, T = ctx->p.poly_size_limit - mismatch;
        if (mismatch < 0) A.poly_m = -1;
        else if (T <= 1) A.poly_m = 0;
        else { int m = (T + mismatch) / (mismatch + 1) - 1; A.poly_m = m < 0 ? 0 : m; }
    }
    const bool pe = b.seq2 != nullptr;
    const int nw = lane_path(ctx, x.mode, maxl) ? lane_words_for(maxl) : 0;

    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    if (timed) {
        if (ctx->ev_used == ctx->ev_pool.size()) {
            std::array<cudaEvent_t, 4> a;
            for (auto &e : a) CK(cudaEventCreate(&e));
            ctx->ev_pool.push_back(a);
        }
        for (int i = 0; i < 4; i++) ev[i] = ctx->ev_pool[ctx->ev_used][i];
        ctx->ev_used++;
    }
    auto mark = [&](int i) -> int { if (timed) CK(cudaEventRecord(ev[i], stream)); return 0; };

    // statRead over records [lo, hi) of the batch with stat_kernel (aqc_stat_kernel.cuh): stamp_head_kernel puts the first-seen
    // stamps of the first records' dense k-mers into the table, stamp_bits_kernel turns the table into the bitmap "stamped below
    // every stamp the rest of the range can produce", and stat_kernel walks the whole range with that bitmap in shared memory.
    auto stat_launches = [&](const KArgs &K0, bool post, uint32_t lo, uint32_t hi) -> int {
        if (hi <= lo) return 0;
        const int snw = lane_words_for(maxl);
        const void *sk = stat_kernel_for(pe, snw, post);
        const size_t ssmem = stat_smem_bytes(ctx->p.qc_kmer, snw, STAT_WARPS);
        if (ssmem > ctx->max_dyn_smem) return fail(ctx, AQC_ERR_INVALID, "statistics tables do not fit shared memory");
        const bool both = pe && K0.qc[0].valid && K0.qc[1].valid;
        const uint32_t n_dense = 1u << (2 * ctx->p.qc_kmer), bw = stat_kbit_words(ctx->p.qc_kmer);
        const uint32_t HEAD = ctx->stat_head;
        SKArgs SA;
        memset(&SA, 0, sizeof SA);
        SA.k = K0;
        uint32_t *o0 = ctx->d_kbits, *o1 = ctx->d_kbits + bw, *nset = ctx->d_kbits + 2 * bw;
        SA.kbits[0] = o0; SA.kbits[1] = o1; SA.kbits_set = nset;
        SA.lo = lo; SA.hi = hi; SA.head_hi = std::min<uint64_t>(hi, (uint64_t)lo + HEAD);
        void *sargs[1] = {(void *)&SA};
        {
            const uint32_t units = (SA.head_hi - lo) * (pe ? 2u : 1u);
            const uint32_t hgrid = std::max<uint32_t>(1u, std::min<uint32_t>((units + 7) / 8, (uint32_t)ctx->sm_count * 8u));
            CK(cudaLaunchKernel(stamp_head_kernel_for(pe, post), dim3(hgrid), dim3(256), sargs, 0, stream));
            CK(cudaGetLastError());
            ctx->launches++;
        }
        // lowest stamp a record beyond the head can produce: order << 11 (see stat_read)
        unsigned long long min_order = post ? K0.first_index + SA.head_hi : K0.order_base + (K0.first_index + SA.head_hi - K0.stat_lo);
        unsigned long long min_when = min_order << 11;
        const unsigned long long *f0 = K0.qc[0].valid ? K0.qc[0].kfirst : nullptr, *f1 = (pe && K0.qc[1].valid) ? K0.qc[1].kfirst : nullptr;
        uint32_t nd = n_dense;
        CK(cudaMemsetAsync(nset, 0, 2 * sizeof(uint32_t), stream));
        void *bargs[7] = {(void *)&f0, (void *)&f1, (void *)&nd, (void *)&min_when, (void *)&o0, (void *)&o1, (void *)&nset};
#ifndef AQC_EMU
        const void *bk = (const void *)stamp_bits_kernel;
#else
        const void *bk = (const void *)(simt::Entry)emu_stamp_bits_kernel;
#endif
        CK(cudaLaunchKernel(bk, dim3(std::max<uint32_t>(1u, std::min<uint32_t>((n_dense + 255) / 256, 64u))), dim3(256), bargs, 0, stream));
        CK(cudaGetLastError());
        ctx->launches++;
        const uint32_t tiles = (hi - lo + 31) / 32;
        const uint32_t per_mate = std::max<uint32_t>(1u, std::min<uint32_t>((tiles + STAT_WARPS - 1) / STAT_WARPS,
                                                                           both ? std::max(1, ctx->sm_count / 2) : ctx->sm_count));
        const uint32_t sgrid = both ? 2 * per_mate : per_mate;
        CK(cudaLaunchKernel(sk, dim3(sgrid), dim3(STAT_WARPS * 32), sargs, ssmem, stream));
        CK(cudaGetLastError());
        ctx->launches++;
        return 0;
    };

    if (nw) {
        // ---- lane-per-pair kernel over the whole batch, pair_kernel (list mode) over the pairs it handed over, then the
        //      postfilter statistics of the sampled good pairs from the records both wrote ----
        if (b.n > ctx->fb_cap) {
            cudaFree(ctx->d_fb_list); ctx->d_fb_list = nullptr; ctx->fb_cap = 0;
            size_t cap = (size_t)b.n + (size_t)b.n / 4 + 1024;
            CK(cudaMalloc(&ctx->d_fb_list, cap * sizeof(uint32_t)));
            ctx->fb_cap = cap;
        }
        CK(cudaMemsetAsync(ctx->d_fb_count, 0, sizeof(uint32_t), stream));
        LArgs L;
        memset(&L, 0, sizeof L);
        L.k = A;
        L.k.tile_pairs = 32;
        L.k.num_tiles = (b.n + 31) / 32;
        L.fb_list = ctx->d_fb_list; L.fb_count = ctx->d_fb_count;
        L.lane_col_cap = (32 * maxl + 96 + 15) & ~15;
        const void *lk = lane_kernel_for(pe, nw);
        // one CTA per SM with as many warps as the kernel is built for and the stage allows, in steps of four (one per scheduler)
        int best_w = 0, best_occ = 0;
        for (int w = lane_max_warps(nw, pe); w >= 1 && best_w == 0; w -= (w > 4 ? 4 : 1)) {
            size_t sm = lane_smem_bytes(w, L.lane_col_cap, maxl);
            if (sm > ctx->max_dyn_smem) continue;
            int occ = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lk, w * 32, sm));
            if (occ > 0) { best_w = w; best_occ = occ; }
        }
        if (const char *fw = getenv("AQC_LANE_WARPS")) {        // tuning knob
            int w = std::max(1, std::min(lane_max_warps(nw, pe), atoi(fw)));
            size_t sm = lane_smem_bytes(w, L.lane_col_cap, maxl);
            int occ = 0;
            if (sm <= ctx->max_dyn_smem) { CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lk, w * 32, sm)); if (occ > 0) { best_w = w; best_occ = occ; } }
        }
        if (best_w == 0) return fail(ctx, AQC_ERR_INVALID, "lane kernel stage does not fit shared memory");
        const size_t lsmem = lane_smem_bytes(best_w, L.lane_col_cap, maxl);
        uint32_t want = (L.k.num_tiles + best_w - 1) / best_w;
        uint32_t lgrid = std::min<uint32_t>(want, (uint32_t)(ctx->sm_count * best_occ));
        if (mark(0)) return AQC_ERR_CUDA;
        void *largs[1] = {(void *)&L};
        CK(cudaLaunchKernel(lk, dim3(lgrid), dim3(best_w * 32), largs, lsmem, stream));
        CK(cudaGetLastError());
        ctx->launches++;
        if (mark(1)) return AQC_ERR_CUDA;
        // general kernel over the hand-over list (usually empty: the count lives in device memory); no statistics there either
        KArgs G = A;
        G.list = ctx->d_fb_list; G.list_count = ctx->d_fb_count;
        G.no_stats = 1;
        size_t smem = pair_tiling(ctx, G, b.n, maxl, 1);
        if (smem > ctx->max_dyn_smem) return fail(ctx, AQC_ERR_INVALID, "tile does not fit shared memory");
        const void *kern = kernel_for(MODE_LIST, pe);
        G.mode = MODE_LIST;
        int occ = 1;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem));
        if (occ < 1) occ = 1;
        uint32_t grid = std::min<uint32_t>(b.n, (uint32_t)(ctx->sm_count * occ));
        void *kargs[1] = {(void *)&G};
        CK(cudaLaunchKernel(kern, dim3(grid), dim3(THREADS), kargs, smem, stream));
        CK(cudaGetLastError());
        ctx->launches++;
        if (mark(2)) return AQC_ERR_CUDA;
        // ---- postfilter statistics of the sampled good pairs (preprocesser.py:624-627) ----
        uint64_t lim = b.n;                                  // pairs [0, lim) of this batch are inside the sample gate (quirk Q10)
        if (ctx->p.qc_sample > 0) {
            const uint64_t gate = (uint64_t)ctx->p.qc_sample - 1;      // global indices below this one are sampled
            lim = gate > b.first_index ? std::min<uint64_t>(b.n, gate - b.first_index) : 0;
        }
        int rc = stat_launches(L.k, true, 0, (uint32_t)lim);
        if (rc) return rc;
        if (mark(3)) return AQC_ERR_CUDA;
        return 0;
    }

    if (x.mode == MODE_STAT && stat_path(ctx, maxl)) {
        // ---- prefilter statistics: only the records inside the window [stat_lo, stat_hi) are walked ----
        uint64_t lo = x.stat_lo > b.first_index ? x.stat_lo - b.first_index : 0;
        uint64_t hi = x.stat_hi > b.first_index ? std::min<uint64_t>(b.n, x.stat_hi - b.first_index) : 0;
        if (mark(0) || mark(1) || mark(2)) return AQC_ERR_CUDA;
        if (lo < hi) { int rc = stat_launches(A, false, (uint32_t)lo, (uint32_t)hi); if (rc) return rc; }
        if (mark(3)) return AQC_ERR_CUDA;
        return 0;
    }

    size_t smem = pair_tiling(ctx, A, b.n, maxl, MAX_TILE_PAIRS);
    if (smem > ctx->max_dyn_smem) return fail(ctx, AQC_ERR_INVALID, "tile does not fit shared memory");
    int occ = 1;
    const void *kern = kernel_for(x.mode, pe);
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem));
    if (occ < 1) occ = 1;
    uint32_t grid = std::min<uint32_t>(A.num_tiles, (uint32_t)(ctx->sm_count * occ));
    // pair_kernel does everything of its mode in one launch: reported as phase 0 (filter / ops) or phase 2 (MODE_STAT)
    const bool as_stat = x.mode == MODE_STAT;
    if (mark(0) || (as_stat && (mark(1) || mark(2)))) return AQC_ERR_CUDA;
    void *kargs[1] = {(void *)&A};
    CK(cudaLaunchKernel(kern, dim3(grid), dim3(THREADS), kargs, smem, stream));
    CK(cudaGetLastError());
    if ((!as_stat && (mark(1) || mark(2))) || mark(3)) return AQC_ERR_CUDA;
    ctx->launches++;
    return 0;
}

int ensure_staging(aqc_ctx *ctx, Staging &s, size_t col_bytes, size_t off_entries, size_t res_bytes) {
    if (!s.h2d_done) {
        CK(cudaEventCreateWithFlags(&s.h2d_done, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&s.k_done, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&s.d2h_done, cudaEventDisableTiming));
    }
    if (col_bytes > s.col_cap) {
        for (int k = 0; k < 4; k++) { cudaFree(s.col[k]); s.col[k] = nullptr; }
        size_t cap = col_bytes + col_bytes / 4 + 256;
        for (int k = 0; k < 4; k++) CK(cudaMalloc(&s.col[k], cap));
        s.col_cap = cap;
    }
    if (off_entries > s.off_cap) {
        for (int k = 0; k < 2; k++) { cudaFree(s.off[k]); s.off[k] = nullptr; }
        size_t cap = off_entries + 64;
        for (int k = 0; k < 2; k++) CK(cudaMalloc(&s.off[k], cap * 4));
        s.off_cap = cap;
    }
    if (res_bytes > s.res_cap) {
        cudaFree(s.res); s.res = nullptr;
        CK(cudaMalloc(&s.res, res_bytes + 256));
        s.res_cap = res_bytes;
    }
    return 0;
}

// host-resident batch: chunked, double-buffered H2D -> kernel -> D2H
int run_host(aqc_ctx *ctx, const aqc_batch *b, const LaunchExtra &x0, void *out_host, size_t out_elem) {
    const bool paired = b->seq2 != nullptr;
    ctx->ev_used = 0;
    uint32_t done = 0;
    int slot = 0;
    bool used[2] = {false, false};
    while (done < b->n) {
        uint32_t lo = done, hi = std::min(b->n, lo + ctx->chunk_pairs);
        uint32_t cn = hi - lo;
        uint32_t a1 = b->off1[lo], e1 = b->off1[hi], g1 = a1 & ~15u;
        uint32_t a2 = 0, e2 = 0, g2 = 0;
        if (paired) { a2 = b->off2[lo]; e2 = b->off2[hi]; g2 = a2 & ~15u; }
        int maxl = (int)(b->flags & 0xFFFFu);               // the caller's hint (longest read of the batch), else look
        if (maxl == 0) {
            for (uint32_t i = lo; i < hi; i++) {
                maxl = std::max<int>(maxl, (int)(b->off1[i + 1] - b->off1[i]));
                if (paired) maxl = std::max<int>(maxl, (int)(b->off2[i + 1] - b->off2[i]));
            }
        }
        Staging &s = ctx->stg[slot];
        if (used[slot]) CK(cudaEventSynchronize(s.d2h_done));    // slot free again
        size_t cb = std::max<size_t>(e1 - g1, paired ? (size_t)(e2 - g2) : 0) + 64;
        int rc = ensure_staging(ctx, s, cb, (size_t)cn + 8, out_host ? (size_t)cn * out_elem : 0);
        if (rc) return rc;
        // mate-2 qualities may stay in page-locked host memory when the lane-per-pair kernel runs (AQC_BATCH_QUAL2_IN_PLACE)
        const uint8_t *q2_in_place = nullptr;
        if (paired && (b->flags & AQC_BATCH_QUAL2_IN_PLACE) && e2 > a2 && lane_path(ctx, x0.mode, maxl))
            q2_in_place = device_view_of_host(b->qual2, a2, e2 - 1);
        const uint8_t *const srcs[4] = {b->seq1 + g1, b->qual1 + g1, paired ? b->seq2 + g2 : nullptr, paired ? b->qual2 + g2 : nullptr};
        const size_t lens[4] = {(size_t)(e1 - g1), (size_t)(e1 - g1), (size_t)(e2 - g2), (size_t)(e2 - g2)};
        CK(cudaMemcpyAsync(s.col[0], srcs[0], lens[0], cudaMemcpyHostToDevice, ctx->copy_in));
        CK(cudaMemcpyAsync(s.col[1], srcs[1], lens[1], cudaMemcpyHostToDevice, ctx->copy_in));
        CK(cudaMemcpyAsync(s.off[0], b->off1 + lo, (size_t)(cn + 1) * 4, cudaMemcpyHostToDevice, ctx->copy_in));
        if (paired) {
            CK(cudaMemcpyAsync(s.col[2], srcs[2], lens[2], cudaMemcpyHostToDevice, ctx->copy_in));
            if (!q2_in_place) CK(cudaMemcpyAsync(s.col[3], srcs[3], lens[3], cudaMemcpyHostToDevice, ctx->copy_in));
            CK(cudaMemcpyAsync(s.off[1], b->off2 + lo, (size_t)(cn + 1) * 4, cudaMemcpyHostToDevice, ctx->copy_in));
        }
        CK(cudaEventRecord(s.h2d_done, ctx->copy_in));
        CK(cudaStreamWaitEvent(ctx->compute, s.h2d_done, 0));
        DevBatch d;
        // virtual column bases so that the absolute offsets of the chunk index the staged bytes
        d.seq1 = s.col[0] - g1; d.qual1 = s.col[1] - g1;
        d.seq2 = paired ? s.col[2] - g2 : nullptr; d.qual2 = paired ? (q2_in_place ? q2_in_place : s.col[3] - g2) : nullptr;
        d.off1 = s.off[0]; d.off2 = paired ? s.off[1] : nullptr;
        d.n = cn; d.first_index = b->first_index + lo; d.max_len = maxl;
        d.qual2_in_sysmem = q2_in_place != nullptr;
        LaunchExtra x = x0;
        x.out = s.res;
        rc = launch(ctx, d, x, ctx->compute, true);
        if (rc) return rc;
        CK(cudaEventRecord(s.k_done, ctx->compute));
        CK(cudaStreamWaitEvent(ctx->copy_out, s.k_done, 0));
        if (out_host)
            CK(cudaMemcpyAsync((uint8_t *)out_host + (size_t)lo * out_elem, s.res, (size_t)cn * out_elem, cudaMemcpyDeviceToHost, ctx->copy_out));
        CK(cudaEventRecord(s.d2h_done, ctx->copy_out));
        used[slot] = true;
        slot ^= 1;
        done = hi;
    }
    CK(cudaStreamSynchronize(ctx->copy_out));
    CK(cudaStreamSynchronize(ctx->compute));
    int e = poll_error(ctx);
    return e;
}

int run_device(aqc_ctx *ctx, const aqc_batch *b, const LaunchExtra &x) {
    ctx->ev_used = 0;
    DevBatch d;
    d.seq1 = b->seq1; d.qual1 = b->qual1; d.seq2 = b->seq2; d.qual2 = b->qual2; d.off1 = b->off1; d.off2 = b->off2;
    d.n = b->n; d.first_index = b->first_index;
    int hint = (int)(b->flags & 0xFFFFu);
    if (hint > 0) d.max_len = hint;
    else {
        uint32_t m = 0;
        CK(cudaMemsetAsync(ctx->d_maxlen, 0, 4, ctx->compute));
        if (b->n) {
            const uint32_t *o1 = b->off1, *o2 = b->off2;
            uint32_t nn = b->n, *dm = ctx->d_maxlen;
            void *margs[4] = {(void *)&o1, (void *)&o2, (void *)&nn, (void *)&dm};
#ifndef AQC_EMU
            const void *mk = (const void *)maxlen_kernel;
#else
            const void *mk = (const void *)(simt::Entry)emu_maxlen_kernel;
#endif
            CK(cudaLaunchKernel(mk, dim3(std::min<uint32_t>((b->n + 255) / 256, 1024u)), dim3(256), margs, 0, ctx->compute));
            ctx->launches++;
        }
        CK(cudaMemcpyAsync(&m, ctx->d_maxlen, 4, cudaMemcpyDeviceToHost, ctx->compute));
        CK(cudaStreamSynchronize(ctx->compute));
        d.max_len = (int)m;
    }
    return launch(ctx, d, x, ctx->compute, true);
}

int check_batch(aqc_ctx *ctx, const aqc_batch *b, int mem) {
    if (!ctx || !b) return AQC_ERR_INVALID;
    if (mem != AQC_MEM_HOST && mem != AQC_MEM_DEVICE) return fail(ctx, AQC_ERR_INVALID, "bad memory space");
    if (b->n && (!b->seq1 || !b->qual1 || !b->off1)) return fail(ctx, AQC_ERR_INVALID, "null mate-1 column");
    if (b->seq2 && (!b->qual2 || !b->off2)) return fail(ctx, AQC_ERR_INVALID, "incomplete mate-2 columns");
    if (ctx->sticky_err) return ctx->sticky_err;
    return 0;
}

}  // namespace

// =============================================================================================
extern "C" {

int aqc_abi_version(void) { return AQC_ABI_VERSION; }

int aqc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char *aqc_last_error(const aqc_ctx *ctx) { return ctx ? ctx->err : g_create_err; }

int aqc_create(int device, const aqc_params *params, aqc_ctx **out) {
    if (!params || !out) { snprintf(g_create_err, sizeof g_create_err, "null argument"); return AQC_ERR_INVALID; }
    int rc = check_params(params, g_create_err, sizeof g_create_err);
    if (rc) return rc;
    aqc_ctx *ctx = new aqc_ctx();
    ctx->p = *params;
    auto bail = [&](int code) { snprintf(g_create_err, sizeof g_create_err, "%s", ctx->err); aqc_destroy(ctx); return code; };
    cudaError_t e;
    if (device < 0) { e = cudaGetDevice(&device); if (e != cudaSuccess) { snprintf(ctx->err, sizeof ctx->err, "cudaGetDevice: %s", cudaGetErrorString(e)); return bail(AQC_ERR_CUDA); } }
    ctx->device = device;
    e = cudaSetDevice(device);
    if (e != cudaSuccess) { snprintf(ctx->err, sizeof ctx->err, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e)); return bail(AQC_ERR_CUDA); }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { snprintf(ctx->err, sizeof ctx->err, "cudaGetDeviceProperties: %s", cudaGetErrorString(e)); return bail(AQC_ERR_CUDA); }
    ctx->sm_count = prop.multiProcessorCount;
    auto init = [&]() -> int {
        CK(cudaStreamCreateWithFlags(&ctx->compute, cudaStreamNonBlocking));
        ctx->own_compute = ctx->compute;
        CK(cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
        CK(cudaMalloc(&ctx->d_counters, sizeof(unsigned long long) * AQC_C_TOTAL));
        CK(cudaMalloc(&ctx->d_luts, sizeof(Luts)));
        CK(cudaMalloc(&ctx->d_error, sizeof(int)));
        CK(cudaMalloc(&ctx->d_maxlen, sizeof(uint32_t)));
        Luts L; fill_luts(L);
        CK(cudaMemcpy(ctx->d_luts, &L, sizeof L, cudaMemcpyHostToDevice));
        int optin = 0;
        CK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
        const void *kernels[8] = {kernel_for(MODE_FILTER, true), kernel_for(MODE_FILTER, false), kernel_for(MODE_STAT, true),
                                  kernel_for(MODE_STAT, false), kernel_for(MODE_OPS, true), kernel_for(MODE_OPS, false),
                                  kernel_for(MODE_LIST, true), kernel_for(MODE_LIST, false)};
        size_t max_static = 0;
        for (const void *k : kernels) {
            cudaFuncAttributes fa;
            CK(cudaFuncGetAttributes(&fa, k));
            max_static = std::max(max_static, fa.sharedSizeBytes);
        }
        ctx->max_dyn_smem = (size_t)optin - max_static - 64;
        std::vector<const void *> lanes;
        for (int nw : {4, 5, 8})
            for (bool pe : {true, false}) {
                lanes.push_back(lane_kernel_for(pe, nw));
                for (bool post : {false, true}) lanes.push_back(stat_kernel_for(pe, nw, post));
            }
        for (const void *k : lanes) {
            cudaFuncAttributes fa;
            CK(cudaFuncGetAttributes(&fa, k));
            max_static = std::max(max_static, fa.sharedSizeBytes);
        }
        ctx->max_dyn_smem = (size_t)optin - max_static - 64;
        for (const void *k : kernels) {
            CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->max_dyn_smem));
            CK(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        }
        for (const void *k : lanes) {
            CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->max_dyn_smem));
            CK(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        }
        CK(cudaMalloc(&ctx->d_fb_count, 2 * sizeof(uint32_t)));
        CK(cudaMalloc(&ctx->d_kbits, (2 * (size_t)stat_kbit_words(ctx->p.qc_kmer) + 2) * sizeof(uint32_t)));
        if (const char *cp = getenv("AQC_CHUNK_PAIRS")) {          // host-path chunk size (tests exercise the multi-chunk pipeline with small batches)
            long v = atol(cp);
            if (v >= 4) ctx->chunk_pairs = (uint32_t)std::min<long>(v & ~3L, 1L << 24);
        }
        if (const char *sh = getenv("AQC_STAT_HEAD")) {            // tests: a short head leaves k-mers for stat_kernel's own stamp path
            long v = atol(sh);
            if (v >= 1) ctx->stat_head = (uint32_t)std::min<long>(v, 1L << 24);
        }
        for (int s = 0; s < AQC_NUM_QC; s++) { int r = alloc_qc(ctx, ctx->qc[s]); if (r) return r; }
        return aqc_reset(ctx);
    };
    rc = init();
    if (rc) return bail(rc);
    *out = ctx;
    return 0;
}

void aqc_destroy(aqc_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->compute) cudaStreamSynchronize(ctx->compute);
    for (int s = 0; s < AQC_NUM_QC; s++) free_qc(ctx->qc[s]);
    for (auto &st : ctx->stg) {
        for (int k = 0; k < 4; k++) cudaFree(st.col[k]);
        for (int k = 0; k < 2; k++) cudaFree(st.off[k]);
        cudaFree(st.res);
        if (st.h2d_done) { cudaEventDestroy(st.h2d_done); cudaEventDestroy(st.k_done); cudaEventDestroy(st.d2h_done); }
    }
    for (auto &ev : ctx->ev_pool) for (auto e : ev) cudaEventDestroy(e);
    cudaFree(ctx->d_counters); cudaFree(ctx->d_luts); cudaFree(ctx->d_error); cudaFree(ctx->d_maxlen);
    cudaFree(ctx->d_fb_list); cudaFree(ctx->d_fb_count); cudaFree(ctx->d_kbits);
    for (auto &ps : ctx->parse) {
        cudaFree(ps.text); cudaFree(ps.seq); cudaFree(ps.qual); cudaFree(ps.blk); cudaFree(ps.nl); cudaFree(ps.line_start);
        cudaFree(ps.line_len); cudaFree(ps.rec); cudaFree(ps.part); cudaFree(ps.flags);
    }
    if (ctx->own_compute) cudaStreamDestroy(ctx->own_compute);
    if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
    if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
    delete ctx;
}

int aqc_set_params(aqc_ctx *ctx, const aqc_params *params) {
    if (!ctx || !params) return AQC_ERR_INVALID;
    int rc = check_params(params, ctx->err, sizeof ctx->err);
    if (rc) return rc;
    if (params->qc_kmer != ctx->p.qc_kmer) return fail(ctx, AQC_ERR_INVALID, "qc_kmer cannot change after create");
    int keep = ctx->p.kmer_side_log2;
    ctx->p = *params;
    ctx->p.kmer_side_log2 = keep;
    return 0;
}

int aqc_reset(aqc_ctx *ctx) {
    if (!ctx) return AQC_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemsetAsync(ctx->d_counters, 0, sizeof(unsigned long long) * AQC_C_TOTAL, ctx->compute));
    CK(cudaMemsetAsync(ctx->d_error, 0, sizeof(int), ctx->compute));
    for (int s = 0; s < AQC_NUM_QC; s++) { int r = zero_qc(ctx, ctx->qc[s]); if (r) return r; }
    CK(cudaStreamSynchronize(ctx->compute));
    ctx->sticky_err = 0; ctx->err[0] = 0;
    return 0;
}

int aqc_reset_filter(aqc_ctx *ctx) {
    if (!ctx) return AQC_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemsetAsync(ctx->d_counters, 0, sizeof(unsigned long long) * AQC_C_TOTAL, ctx->compute));
    for (int s = AQC_QC_R1_POST; s < AQC_NUM_QC; s++) { int r = zero_qc(ctx, ctx->qc[s]); if (r) return r; }
    CK(cudaStreamSynchronize(ctx->compute));
    return 0;
}

int aqc_host_alloc(size_t bytes, void **out) {
    if (!out) return AQC_ERR_INVALID;
    return cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault) == cudaSuccess ? 0 : AQC_ERR_NOMEM;
}
void aqc_host_free(void *p) { if (p) cudaFreeHost(p); }

int aqc_device_alloc(aqc_ctx *ctx, size_t bytes, void **out) {
    if (!ctx || !out) return AQC_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMalloc(out, bytes ? bytes : 16));
    return 0;
}
void aqc_device_free(aqc_ctx *ctx, void *p) { if (ctx && p) { cudaSetDevice(ctx->device); cudaFree(p); } }
int aqc_memcpy_h2d(aqc_ctx *ctx, void *dst, const void *src, size_t bytes) {
    if (!ctx) return AQC_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return 0;
}
int aqc_memcpy_d2h(aqc_ctx *ctx, void *dst, const void *src, size_t bytes) {
    if (!ctx) return AQC_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->compute));
    CK(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

int aqc_stat_reads(aqc_ctx *ctx, const aqc_batch *batch, int mem, int qc1, int qc2, uint64_t stat_lo, uint64_t stat_hi, uint64_t order_base) {
    int rc = check_batch(ctx, batch, mem);
    if (rc) return rc;
    if (qc1 >= AQC_NUM_QC || qc2 >= AQC_NUM_QC) return fail(ctx, AQC_ERR_INVALID, "bad QC slot");
    CK(cudaSetDevice(ctx->device));
    LaunchExtra x; x.mode = MODE_STAT; x.qc1 = qc1; x.qc2 = batch->seq2 ? qc2 : -1;
    x.stat_lo = stat_lo; x.stat_hi = stat_hi; x.order_base = order_base; x.out = nullptr;
    return mem == AQC_MEM_HOST ? run_host(ctx, batch, x, nullptr, 0) : run_device(ctx, batch, x);
}

int aqc_filter_pairs(aqc_ctx *ctx, const aqc_batch *batch, int mem, aqc_result *results) {
    int rc = check_batch(ctx, batch, mem);
    if (rc) return rc;
    if (!results && batch->n) return fail(ctx, AQC_ERR_INVALID, "null results");
    CK(cudaSetDevice(ctx->device));
    LaunchExtra x; x.mode = MODE_FILTER; x.qc1 = AQC_QC_R1_POST; x.qc2 = batch->seq2 ? AQC_QC_R2_POST : -1;
    x.stat_lo = 0; x.stat_hi = 0; x.order_base = 0; x.out = results;
    return mem == AQC_MEM_HOST ? run_host(ctx, batch, x, results, sizeof(aqc_result)) : run_device(ctx, batch, x);
}

int aqc_ops_pairs(aqc_ctx *ctx, const aqc_batch *batch, int mem, aqc_ops *out) {
    int rc = check_batch(ctx, batch, mem);
    if (rc) return rc;
    if (!out && batch->n) return fail(ctx, AQC_ERR_INVALID, "null output");
    CK(cudaSetDevice(ctx->device));
    LaunchExtra x; x.mode = MODE_OPS; x.qc1 = -1; x.qc2 = -1;
    x.stat_lo = 0; x.stat_hi = 0; x.order_base = 0; x.out = out;
    return mem == AQC_MEM_HOST ? run_host(ctx, batch, x, out, sizeof(aqc_ops)) : run_device(ctx, batch, x);
}

int aqc_set_stream(aqc_ctx *ctx, void *cuda_stream) {
    if (!ctx) return AQC_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->compute));
    ctx->compute = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_compute;
    return 0;
}

int aqc_device_ptr(aqc_ctx *ctx, int what, int slot, void **ptr_out, uint64_t *n_out) {
    if (!ctx || !ptr_out || !n_out) return AQC_ERR_INVALID;
    if (what == 0) { *ptr_out = ctx->d_counters; *n_out = AQC_C_TOTAL; return 0; }
    if (slot < 0 || slot >= AQC_NUM_QC) return AQC_ERR_INVALID;
    QcHost &q = ctx->qc[slot];
    switch (what) {
        case 1: *ptr_out = q.d.cls_cnt; *n_out = QC_CLASSES * AQC_MAX_LEN; break;
        case 2: *ptr_out = q.d.cls_qsum; *n_out = QC_CLASSES * AQC_MAX_LEN; break;
        case 3: *ptr_out = q.d.disc; *n_out = AQC_MAX_LEN; break;
        case 4: *ptr_out = q.d.gchist; *n_out = AQC_MAX_LEN + 1; break;
        case 5: *ptr_out = q.d.scal; *n_out = 2; break;
        case 6: *ptr_out = q.d.kcnt; *n_out = q.dense_n; break;
        case 7: *ptr_out = q.d.kfirst; *n_out = q.dense_n; break;
        case 8: *ptr_out = q.d.skeys; *n_out = q.side_cap; break;
        case 9: *ptr_out = q.d.scnt; *n_out = q.side_cap; break;
        case 10: *ptr_out = q.d.sfirst; *n_out = q.side_cap; break;
        case 11: *ptr_out = q.d.sseed; *n_out = q.side_cap; break;
        default: return AQC_ERR_INVALID;
    }
    return 0;
}

int aqc_sync(aqc_ctx *ctx) {
    if (!ctx) return AQC_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->compute));
    return poll_error(ctx);
}

int aqc_get_counters(aqc_ctx *ctx, int64_t *out) {
    if (!ctx || !out) return AQC_ERR_INVALID;
    int rc = aqc_sync(ctx);
    CK(cudaMemcpy(out, ctx->d_counters, sizeof(int64_t) * AQC_C_TOTAL, cudaMemcpyDeviceToHost));
    return rc;
}

int aqc_add_counters(aqc_ctx *ctx, const int64_t *in) {
    if (!ctx || !in) return AQC_ERR_INVALID;
    std::vector<int64_t> cur(AQC_C_TOTAL);
    int rc = aqc_get_counters(ctx, cur.data());
    if (rc) return rc;
    for (int i = 0; i < AQC_C_TOTAL; i++) cur[i] += in[i];
    CK(cudaMemcpy(ctx->d_counters, cur.data(), sizeof(int64_t) * AQC_C_TOTAL, cudaMemcpyHostToDevice));
    return 0;
}

int aqc_get_qc(aqc_ctx *ctx, int slot, aqc_qc_counters *out) {
    if (!ctx || !out || slot < 0 || slot >= AQC_NUM_QC) return AQC_ERR_INVALID;
    int rc = aqc_sync(ctx);
    QcHost &q = ctx->qc[slot];
    std::vector<unsigned long long> cnt(QC_CLASSES * AQC_MAX_LEN), qs(QC_CLASSES * AQC_MAX_LEN), disc(AQC_MAX_LEN), gch(AQC_MAX_LEN + 1);
    unsigned long long scal[2];
    CK(cudaMemcpy(cnt.data(), q.d.cls_cnt, cnt.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(qs.data(), q.d.cls_qsum, qs.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(disc.data(), q.d.disc, disc.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(gch.data(), q.d.gchist, gch.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(scal, q.d.scal, sizeof scal, cudaMemcpyDeviceToHost));
    memset(out, 0, sizeof *out);
    for (int i = 0; i < AQC_MAX_LEN; i++) {
        int64_t tn = 0, tq = 0;
        for (int c = 0; c < QC_CLASSES; c++) { tn += (int64_t)cnt[c * AQC_MAX_LEN + i]; tq += (int64_t)qs[c * AQC_MAX_LEN + i]; }
        out->totalNum[i] = tn;
        out->totalQual[i] = tq - 33 * tn;                       // util.qualNum: ord(q) - 33
        for (int c = 0; c < 4; c++) {
            out->baseCounts[c][i] = (int64_t)cnt[c * AQC_MAX_LEN + i];
            out->baseTotalQual[c][i] = (int64_t)qs[c * AQC_MAX_LEN + i] - 33 * (int64_t)cnt[c * AQC_MAX_LEN + i];
        }
        out->totalDiscontinuity[i] = (int64_t)disc[i];
    }
    for (int i = 0; i <= AQC_MAX_LEN; i++) out->gcHistogram[i] = (int64_t)gch[i];
    out->totalKmer = (int64_t)scal[0];
    out->reads = (int64_t)scal[1];
    return rc;
}

int aqc_get_kmer_dense(aqc_ctx *ctx, int slot, uint64_t *counts, uint64_t *first) {
    if (!ctx || !counts || !first || slot < 0 || slot >= AQC_NUM_QC) return AQC_ERR_INVALID;
    int rc = aqc_sync(ctx);
    QcHost &q = ctx->qc[slot];
    const int K = ctx->p.qc_kmer;
    std::vector<unsigned long long> c(q.dense_n), f(q.dense_n);
    CK(cudaMemcpy(c.data(), q.d.kcnt, q.dense_n * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(f.data(), q.d.kfirst, q.dense_n * 8, cudaMemcpyDeviceToHost));
    // internal index (plane1 bits << K) | plane0 bits, bit t = base t  ->  natural index, base 0 most significant.
    // The device keeps DIRECT first sightings; a k-mer is also inserted (with count 0) when its reverse complement is
    // first seen (qualitycontrol.py:118-122): first[X] = min(direct[X], direct[rc(X)] | 1).
    const uint32_t kmask = (1u << K) - 1u;
    auto brevK = [&](uint32_t w) { uint32_t r = 0; for (int t = 0; t < K; t++) r |= ((w >> t) & 1u) << (K - 1 - t); return r; };
    for (size_t in = 0; in < q.dense_n; in++) {
        uint32_t w0 = (uint32_t)in & kmask, w1 = (uint32_t)(in >> K);
        size_t nat = 0;
        for (int t = 0; t < K; t++) {
            uint32_t code = ((w0 >> t) & 1u) | (((w1 >> t) & 1u) << 1);
            nat |= (size_t)code << (2 * (K - 1 - t));
        }
        const size_t rin = ((size_t)brevK(~w1 & kmask) << K) | brevK(~w0 & kmask);   // reverse complement (both code bits inverted)
        unsigned long long fst = f[in];
        if (f[rin] != AQC_KMER_NEVER) fst = std::min(fst, f[rin] | 1ULL);
        counts[nat] = c[in];
        first[nat] = fst;
    }
    return rc;
}

int aqc_get_kmer_side(aqc_ctx *ctx, int slot, uint64_t *keys, uint64_t *counts, uint64_t *first, uint32_t cap, uint32_t *n_out) {
    if (!ctx || !n_out || slot < 0 || slot >= AQC_NUM_QC) return AQC_ERR_INVALID;
    int rc = aqc_sync(ctx);
    QcHost &q = ctx->qc[slot];
    const int K = ctx->p.qc_kmer;
    std::vector<unsigned long long> k(q.side_cap), c(q.side_cap), d(q.side_cap), sf(q.side_cap);
    CK(cudaMemcpy(k.data(), q.d.skeys, (size_t)q.side_cap * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(c.data(), q.d.scnt, (size_t)q.side_cap * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(d.data(), q.d.sfirst, (size_t)q.side_cap * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(sf.data(), q.d.sseed, (size_t)q.side_cap * 8, cudaMemcpyDeviceToHost));
    Luts L; fill_luts(L);
    auto find = [&](unsigned long long key) -> long {
        uint32_t h = (uint32_t)host_side_hash(key) & q.d.smask;
        for (uint32_t probe = 0; probe <= q.d.smask; probe++) {
            if (k[h] == key) return (long)h;
            if (k[h] == AQC_KMER_NEVER) return -1;
            h = (h + 1) & q.d.smask;
        }
        return -1;
    };
    // resolve the insertion stamp of every slot (see the comment above stat_read in aqc_device.cuh)
    std::vector<unsigned long long> stamp(q.side_cap, AQC_KMER_NEVER);
    uint32_t n = 0;
    for (uint32_t i = 0; i < q.side_cap; i++) {
        if (k[i] == AQC_KMER_NEVER) continue;
        unsigned long long key = k[i], rkey = 0;
        bool foreign = false;
        for (int j = 0; j < K; j++) {
            uint8_t bj = (uint8_t)(key >> (8 * (K - 1 - j)));
            rkey |= (unsigned long long)L.lut3[bj] << (8 * j);
            foreign |= (L.lut1[bj] & 15u) == 15u;
        }
        unsigned long long p;
        if (foreign) p = d[i];                                  // no pre-image: present from its first sighting
        else {
            p = std::min(d[i], sf[i]);
            if (rkey != key) {
                long j = find(rkey);
                if (j >= 0 && d[j] < d[i]) {                    // partner sighted first: it seeds us iff it was newly inserted
                    bool partner_new = d[j] < sf[j];
                    if (partner_new) p = std::min(p, d[j] | 1ULL);
                }
            }
        }
        stamp[i] = p;
        if (p != AQC_KMER_NEVER) n++;
    }
    *n_out = n;
    if (cap == 0 && !keys) return rc;
    if (cap < n || !keys || !counts || !first) return fail(ctx, AQC_ERR_INVALID, "side-table output too small");
    uint32_t o = 0;
    for (uint32_t i = 0; i < q.side_cap; i++)
        if (stamp[i] != AQC_KMER_NEVER) { keys[o] = k[i]; counts[o] = c[i]; first[o] = stamp[i]; o++; }
    return rc;
}

int aqc_get_kmer_side_raw(aqc_ctx *ctx, int slot, uint64_t *keys, uint64_t *counts, uint64_t *first_direct, uint64_t *first_seed,
                          uint32_t cap, uint32_t *n_out) {
    if (!ctx || !n_out || slot < 0 || slot >= AQC_NUM_QC) return AQC_ERR_INVALID;
    int rc = aqc_sync(ctx);
    QcHost &q = ctx->qc[slot];
    std::vector<unsigned long long> k(q.side_cap), c(q.side_cap), d(q.side_cap), sf(q.side_cap);
    CK(cudaMemcpy(k.data(), q.d.skeys, (size_t)q.side_cap * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(c.data(), q.d.scnt, (size_t)q.side_cap * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(d.data(), q.d.sfirst, (size_t)q.side_cap * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(sf.data(), q.d.sseed, (size_t)q.side_cap * 8, cudaMemcpyDeviceToHost));
    uint32_t n = 0;
    for (uint32_t i = 0; i < q.side_cap; i++) n += k[i] != AQC_KMER_NEVER;
    *n_out = n;
    if (cap == 0 && !keys) return rc;
    if (cap < n || !keys || !counts || !first_direct || !first_seed) return fail(ctx, AQC_ERR_INVALID, "side-table output too small");
    uint32_t o = 0;
    for (uint32_t i = 0; i < q.side_cap; i++)
        if (k[i] != AQC_KMER_NEVER) { keys[o] = k[i]; counts[o] = c[i]; first_direct[o] = d[i]; first_seed[o] = sf[i]; o++; }
    return rc;
}

int aqc_edit_distance_batch(aqc_ctx *ctx, const uint8_t *a, const uint32_t *a_off, const uint8_t *b, const uint32_t *b_off,
                            uint32_t n, int mem, int32_t *out) {
    if (!ctx || (n && (!a_off || !b_off || !out))) return AQC_ERR_INVALID;
    if (mem != AQC_MEM_HOST && mem != AQC_MEM_DEVICE) return fail(ctx, AQC_ERR_INVALID, "bad memory space");
    if (n == 0) return 0;
    CK(cudaSetDevice(ctx->device));
    const uint8_t *da = a, *db = b;
    const uint32_t *dao = a_off, *dbo = b_off;
    int32_t *dout = out;
    void *tmp[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    auto release = [&]() { for (void *p : tmp) cudaFree(p); };
    if (mem == AQC_MEM_HOST) {
        const size_t na = a_off[n], nb = b_off[n];
        auto up = [&](int k, const void *src, size_t bytes) -> int {
            CK(cudaMalloc(&tmp[k], bytes + 16));
            if (bytes) CK(cudaMemcpyAsync(tmp[k], src, bytes, cudaMemcpyHostToDevice, ctx->compute));
            return 0;
        };
        int rc = up(0, a, na); if (!rc) rc = up(1, b, nb); if (!rc) rc = up(2, a_off, (size_t)(n + 1) * 4); if (!rc) rc = up(3, b_off, (size_t)(n + 1) * 4);
        if (!rc && cudaMalloc(&tmp[4], (size_t)n * 4) != cudaSuccess) rc = AQC_ERR_NOMEM;
        if (rc) { release(); return rc; }
        da = (const uint8_t *)tmp[0]; db = (const uint8_t *)tmp[1]; dao = (const uint32_t *)tmp[2]; dbo = (const uint32_t *)tmp[3]; dout = (int32_t *)tmp[4];
    }
    void *args[6] = {(void *)&da, (void *)&dao, (void *)&db, (void *)&dbo, (void *)&n, (void *)&dout};
#ifndef AQC_EMU
    const void *kern = (const void *)edit_distance_kernel;
#else
    const void *kern = (const void *)(simt::Entry)emu_edit_distance_kernel;
#endif
    cudaError_t e = cudaLaunchKernel(kern, dim3(std::max<uint32_t>(1u, std::min<uint32_t>((n + 127) / 128, (uint32_t)ctx->sm_count * 16u))), dim3(128), args, 0, ctx->compute);
    if (e == cudaSuccess) e = cudaGetLastError();
    ctx->launches++;
    if (e == cudaSuccess && mem == AQC_MEM_HOST) {
        e = cudaMemcpyAsync(out, dout, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->compute);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->compute);
    }
    release();
    if (e != cudaSuccess) { snprintf(ctx->err, sizeof ctx->err, "edit distance launch failed: %s", cudaGetErrorString(e)); return AQC_ERR_CUDA; }
    return 0;
}

// ---- FASTQ text -> packed columns in HBM (aqc_parse.cuh) ----
extern "C++" {
namespace {
#ifndef AQC_EMU
#define PARSE_KERNEL(name) ((const void *)name)
#else
#define PARSE_KERNEL(name) ((const void *)(simt::Entry)emu_##name)
#endif
template <class T> int grow(aqc_ctx *ctx, T *&p, size_t &cap, size_t want) {
    if (want <= cap && p) return 0;
    cudaFree(p); p = nullptr; cap = 0;
    const size_t n = want + want / 4 + 64;
    CK(cudaMalloc((void **)&p, n * sizeof(T)));
    cap = n;
    return 0;
}
int launch1(aqc_ctx *ctx, const void *k, uint32_t grid, const void *arg_struct) {
    void *args[1] = {const_cast<void *>(arg_struct)};
    CK(cudaLaunchKernel(k, dim3(std::max<uint32_t>(1u, grid)), dim3(PARSE_BLOCK_THREADS), args, 0, ctx->compute));
    CK(cudaGetLastError());
    ctx->launches++;
    return 0;
}
// exclusive scan of data[0 .. n) in place, data[n] = total
int device_scan(aqc_ctx *ctx, aqc_ctx::ParseSlot &ps, uint32_t *data, uint32_t n) {
    ScanArgs S;
    S.data = data; S.n = n; S.n_part = (n + SCAN_BLOCK_ELEMS - 1) / SCAN_BLOCK_ELEMS;
    size_t pc = ps.part_cap;
    int rc = grow(ctx, ps.part, pc, (size_t)S.n_part + 1);
    ps.part_cap = pc;
    if (rc) return rc;
    S.part = ps.part;
    const uint32_t g = std::min<uint32_t>(std::max<uint32_t>(S.n_part, 1u), (uint32_t)ctx->sm_count * 8u);
    if ((rc = launch1(ctx, PARSE_KERNEL(scan_partial_kernel), g, &S))) return rc;
    if ((rc = launch1(ctx, PARSE_KERNEL(scan_single_kernel), 1, &S))) return rc;
    return launch1(ctx, PARSE_KERNEL(scan_final_kernel), g, &S);
}
}  // namespace
}  // extern "C++"

int aqc_fastq_parse_device(aqc_ctx *ctx, int slot, const uint8_t *text, uint64_t n, int mem, int final, uint64_t max_records,
                           aqc_parsed *out) {
    if (!ctx || !out || slot < 0 || slot > 1 || (n && !text)) return AQC_ERR_INVALID;
    if (mem != AQC_MEM_HOST && mem != AQC_MEM_DEVICE) return fail(ctx, AQC_ERR_INVALID, "bad memory space");
    if (n > 0xFFFFFF00ull) return fail(ctx, AQC_ERR_INVALID, "a text buffer must stay below 4 GiB (32-bit positions): split it");
    memset(out, 0, sizeof *out);
    out->hit_eof = final ? 1 : 0;
    if (n == 0) return 0;
    CK(cudaSetDevice(ctx->device));
    aqc_ctx::ParseSlot &ps = ctx->parse[slot];
    cudaStream_t st = ctx->compute;
    int rc;
    // the text: copied next to a 16-byte boundary of our own (host memory), or used where it lies (device memory) unless its
    // last line has no newline and this is the end of the file -- the kernels see a '\n' after every line
    uint8_t last = 0;
    if (mem == AQC_MEM_HOST) last = text[n - 1];
    else { CK(cudaMemcpyAsync(&last, text + n - 1, 1, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st)); }
    const bool append = final && last != (uint8_t)'\n';
    const uint8_t *dtext = text;
    if (mem == AQC_MEM_HOST || append || ((uintptr_t)text & 15u)) {
        if ((rc = grow(ctx, ps.text, ps.text_cap, (size_t)n + 32))) return rc;
        CK(cudaMemcpyAsync(ps.text, text, n, mem == AQC_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, st));
        if (append) CK(cudaMemsetAsync(ps.text + n, '\n', 1, st));
        dtext = ps.text;
    }
    ParseArgs P;
    memset(&P, 0, sizeof P);
    P.text = dtext;
    P.n = (uint32_t)(n + (append ? 1 : 0));
    P.n_blk = (P.n + PARSE_BLOCK_BYTES - 1) / PARSE_BLOCK_BYTES;
    if ((rc = grow(ctx, ps.blk, ps.blk_cap, (size_t)P.n_blk + 1))) return rc;
    if ((rc = grow(ctx, ps.flags, ps.flags_cap, 2))) return rc;
    P.blk = ps.blk; P.flags = ps.flags;
    const uint32_t gmax = (uint32_t)ctx->sm_count * 8u;
    if ((rc = launch1(ctx, PARSE_KERNEL(newline_count_kernel), std::min(P.n_blk, gmax), &P))) return rc;
    if ((rc = device_scan(ctx, ps, ps.blk, P.n_blk))) return rc;
    uint32_t n_lines = 0;
    CK(cudaMemcpyAsync(&n_lines, ps.blk + P.n_blk, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemsetAsync(ps.flags, 0xFF, 8, st));
    CK(cudaStreamSynchronize(st));
    const uint32_t n_cand = (uint32_t)std::min<uint64_t>(n_lines / 4u, max_records);
    if (n_cand == 0) {
        // fewer than four lines (or no record wanted).  The reference stops at an empty line wherever it is: look at the few
        // complete lines there are (host text only; a caller with device text sees it with the next, longer buffer)
        out->hit_eof = (final && max_records) ? 1 : 0;
        if (!out->hit_eof && max_records && mem == AQC_MEM_HOST) {
            uint64_t p = 0;
            for (uint32_t l = 0; l < n_lines && l < 4; l++) {
                const uint8_t *nl = (const uint8_t *)memchr(text + p, '\n', n - p);
                if (!nl) break;
                uint64_t e = (uint64_t)(nl - text);
                while (e > p && (text[e - 1] == ' ' || text[e - 1] == '\t' || text[e - 1] == '\r' || text[e - 1] == 0x0b || text[e - 1] == 0x0c)) e--;
                if (e == p) { out->hit_eof = 1; break; }
                p = (uint64_t)(nl - text) + 1;
            }
        }
        return 0;
    }
    if ((rc = grow(ctx, ps.nl, ps.nl_cap, (size_t)n_lines + 1))) return rc;
    if ((rc = grow(ctx, ps.line_start, ps.line_cap, (size_t)4 * n_cand))) return rc;
    if ((rc = grow(ctx, ps.line_len, ps.line2_cap, (size_t)4 * n_cand))) return rc;
    if ((rc = grow(ctx, ps.rec, ps.rec_cap, (size_t)n_cand + 1))) return rc;
    P.nl_pos = ps.nl; P.n_rec = n_cand; P.line_start = ps.line_start; P.line_len = ps.line_len; P.rec_len = ps.rec;
    if ((rc = launch1(ctx, PARSE_KERNEL(newline_write_kernel), std::min(P.n_blk, gmax), &P))) return rc;
    if ((rc = launch1(ctx, PARSE_KERNEL(record_kernel), std::min((n_cand + PARSE_BLOCK_THREADS - 1) / PARSE_BLOCK_THREADS, gmax), &P))) return rc;
    uint32_t flags[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};
    CK(cudaMemcpyAsync(flags, ps.flags, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    // the reference reads record by record: the first empty line ends the file, a bad record before it is an error
    uint32_t n_keep = n_cand;
    bool eof = false, bad = false;
    if (flags[0] < n_keep) { n_keep = flags[0]; eof = true; }
    if (flags[1] < n_keep) { n_keep = flags[1]; eof = false; bad = true; }
    P.n_keep = n_keep;
    uint32_t end_pos = 0, seq_bytes = 0;
    if (n_keep) {
        if ((rc = device_scan(ctx, ps, ps.rec, n_keep))) return rc;
        CK(cudaMemcpyAsync(&seq_bytes, ps.rec + n_keep, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&end_pos, ps.nl + (size_t)4 * n_keep - 1, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (!ps.seq || !ps.qual || ps.col_cap < (size_t)seq_bytes + 32) {
            cudaFree(ps.seq); cudaFree(ps.qual); ps.seq = ps.qual = nullptr; ps.col_cap = 0;
            const size_t cap = (size_t)seq_bytes + seq_bytes / 4 + 64;
            CK(cudaMalloc((void **)&ps.seq, cap));
            CK(cudaMalloc((void **)&ps.qual, cap));
            ps.col_cap = cap;
        }
        P.seq = ps.seq; P.qual = ps.qual;
        if ((rc = launch1(ctx, PARSE_KERNEL(gather_kernel), std::min((n_keep + 31) / 32, gmax), &P))) return rc;
        CK(cudaStreamSynchronize(st));
    }
    out->n_records = n_keep;
    out->consumed = n_keep ? std::min<uint64_t>((uint64_t)end_pos + 1, n) : 0;
    out->seq_bytes = seq_bytes;
    out->bad_record = bad ? n_keep : 0;
    // as aqc_fastq_parse: an empty line ends the file; at the end of the file whatever is left (a partial record) is dropped
    // (a call that stops at max_records has not looked further: end of file only if the text is used up)
    out->hit_eof = eof ? 1 : (bad ? 0 : ((uint64_t)n_keep == max_records ? (final && out->consumed >= n) : (final ? 1 : 0)));
    if (!out->hit_eof && !bad && !final && (uint64_t)n_keep < max_records && n_lines > 4u * n_keep && mem == AQC_MEM_HOST) {
        // the trailing partial record: the reference would still stop at an empty line in it (host text only, see above)
        uint32_t nlp[5] = {0, 0, 0, 0, 0};
        const uint32_t first = n_keep ? 4u * n_keep - 1u : 0u, cnt = n_lines - first;          // <= 4 positions
        CK(cudaMemcpyAsync(nlp, ps.nl + first, (size_t)std::min<uint32_t>(cnt, 5u) * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        uint64_t p = n_keep ? (uint64_t)nlp[0] + 1 : 0;
        for (uint32_t i = n_keep ? 1u : 0u; i < cnt && i < 5u; i++) {
            uint64_t e = nlp[i];
            while (e > p && (text[e - 1] == ' ' || text[e - 1] == '\t' || text[e - 1] == '\r' || text[e - 1] == 0x0b || text[e - 1] == 0x0c)) e--;
            if (e == p) { out->hit_eof = 1; break; }
            p = (uint64_t)nlp[i] + 1;
        }
    }
    out->seq = ps.seq; out->qual = ps.qual; out->off = ps.rec; out->line_start = ps.line_start; out->line_len = ps.line_len;
    out->text = dtext;
    if (bad) return fail(ctx, AQC_ERR_INVALID, "a quality line is not as long as its sequence line");
    return 0;
}

// ---- libed.so-compatible entry points (editdistance/_editdistance.h:16,23), so that the untouched reference can load this
// library through its own loader (util.py:15-18 -> ed_ctypes.edit_distance / ed_ctypes.seek_overlap).  One call = one pair:
// a lazily created context on the current device serves them (a batch entry exists for real work: aqc_edit_distance_batch,
// aqc_ops_pairs).  seek_overlap has util.overlap_hm's semantics (util.py:158-212: the closed form of the leaked loop variable,
// constants 3 / 30 / 50 whatever order the caller passes them in -- the reference's own call site swaps two of them), NOT the
// semantics of the C function of that name, which differs from the Python on ~12 % of pairs.
namespace {
aqc_ctx *g_shim_ctx = nullptr;
std::mutex g_shim_mutex;
aqc_ctx *shim_ctx() {
    if (!g_shim_ctx) {
        aqc_params p;
        memset(&p, 0, sizeof p);
        p.paired = 1; p.seq_len_req = 35; p.poly_size_limit = 35; p.allow_mismatch_in_poly = 2; p.qualified_quality_phred = 15;
        p.unqualified_base_limit = 60; p.n_base_limit = 5; p.qc_sample = 200000; p.qc_kmer = 8;
        const char *dv = getenv("AQC_DEVICE");
        if (aqc_create(dv ? atoi(dv) : -1, &p, &g_shim_ctx) != 0) g_shim_ctx = nullptr;
    }
    return g_shim_ctx;
}
}  // namespace

unsigned int edit_distance(const char *a, const unsigned int asize, const char *b, const unsigned int bsize) {
    std::lock_guard<std::mutex> lock(g_shim_mutex);
    aqc_ctx *ctx = shim_ctx();
    if (!ctx) return 0xFFFFFFFFu;
    const uint32_t ao[2] = {0, asize}, bo[2] = {0, bsize};
    int32_t d = -1;
    if (aqc_edit_distance_batch(ctx, (const uint8_t *)a, ao, (const uint8_t *)b, bo, 1, AQC_MEM_HOST, &d) != 0) return 0xFFFFFFFFu;
    return (unsigned int)d;
}

int seek_overlap(const char *r1, const int len1, const char *rc_r2, const int len2, const int limit_distance, const int p5, const int p6) {
    if (len1 < 0 || len2 < 0 || len1 > AQC_MAX_LEN || len2 > AQC_MAX_LEN) return 0x7FFFFFFF;
    if (limit_distance != 3 || !((p5 == 30 && p6 == 50) || (p5 == 50 && p6 == 30))) return 0x7FFFFFFF;     // overlap_hm's constants only
    std::lock_guard<std::mutex> lock(g_shim_mutex);
    aqc_ctx *ctx = shim_ctx();
    if (!ctx) return 0x7FFFFFFF;
    // the caller hands over reverseComplement(r2) (util.py:217); the engine takes r2 itself: complement back (an involution on
    // util.COMP's alphabet; bytes outside it are 'N' by now and stay 'N')
    std::vector<uint8_t> s1((size_t)len1 + 16, 0), q1((size_t)len1 + 16, 'I'), s2((size_t)len2 + 16, 0), q2((size_t)len2 + 16, 'I');
    memcpy(s1.data(), r1, (size_t)len1);
    Luts L; fill_luts(L);
    for (int i = 0; i < len2; i++) s2[(size_t)i] = L.lut3[(uint8_t)rc_r2[len2 - 1 - i]];
    const uint32_t o1[2] = {0, (uint32_t)len1}, o2[2] = {0, (uint32_t)len2};
    aqc_batch b;
    memset(&b, 0, sizeof b);
    b.n = 1; b.seq1 = s1.data(); b.qual1 = q1.data(); b.off1 = o1; b.seq2 = s2.data(); b.qual2 = q2.data(); b.off2 = o2;
    aqc_ops r;
    if (aqc_ops_pairs(ctx, &b, AQC_MEM_HOST, &r) != 0) return 0x7FFFFFFF;
    if (r.ov_len == 0) return 0x7FFFFFFF;                       // overlap_hm's (0, 0, 0): "not matched"
    return (int)((uint32_t)((int)r.ov_offset << 8) + std::min<uint32_t>(r.ov_diff, 255u));
}

uint64_t aqc_launch_count(const aqc_ctx *ctx) { return ctx ? ctx->launches : 0; }

float aqc_last_phase_ms(const aqc_ctx *ctx, int phase) {
    if (!ctx || phase < -1 || phase > 2) return 0.f;
    const int a = phase < 0 ? 0 : phase, b = phase < 0 ? 3 : phase + 1;
    float total = 0.f;
    for (size_t i = 0; i < ctx->ev_used; i++) {
        float ms = 0.f;
        if (cudaEventSynchronize(ctx->ev_pool[i][3]) == cudaSuccess &&
            cudaEventElapsedTime(&ms, ctx->ev_pool[i][a], ctx->ev_pool[i][b]) == cudaSuccess) total += ms;
    }
    return total;
}

float aqc_last_kernel_ms(const aqc_ctx *ctx) { return aqc_last_phase_ms(ctx, -1); }

}  // extern "C"
