// aqc_pack.cpp -- host side of the packed base transport (see aqc_pack.hpp).  No CUDA here.
#include "aqc_pack.hpp"

#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif
#if defined(__linux__)
#include <sched.h>
#endif

namespace aqc_pack {

struct Pool {
    std::vector<std::thread> workers;
    std::mutex m;
    std::condition_variable cv_go, cv_done;
    std::function<void(int)> job;
    uint64_t generation = 0;
    int pending = 0;
    bool stop = false;
    int n = 1;
    std::vector<std::vector<uint32_t>> exc_pos;      // per column and thread
    std::vector<std::vector<uint8_t>> exc_val;
};

static void worker(Pool *p, int tid) {
    uint64_t seen = 0;
    for (;;) {
        std::function<void(int)> job;
        {
            std::unique_lock<std::mutex> lk(p->m);
            p->cv_go.wait(lk, [&] { return p->stop || p->generation != seen; });
            if (p->stop) return;
            seen = p->generation;
            job = p->job;
        }
        job(tid);
        {
            std::lock_guard<std::mutex> lk(p->m);
            if (--p->pending == 0) p->cv_done.notify_all();
        }
    }
}

Pool *pool_create(int threads) {
    int n = threads;
    if (const char *e = getenv("AQC_PACK_THREADS")) n = atoi(e);
    if (n <= 0) {
        n = (int)std::thread::hardware_concurrency();
#if defined(__linux__)
        cpu_set_t set;                           // the CPUs this process may run on (containers, taskset, torchrun bindings)
        CPU_ZERO(&set);
        if (sched_getaffinity(0, sizeof set, &set) == 0 && CPU_COUNT(&set) > 0) n = std::min(n > 0 ? n : 1 << 20, (int)CPU_COUNT(&set));
#endif
        n = n > 4 ? n / 2 : n;                   // leave cores to the caller's own threads (readers, writers)
        n = std::max(1, std::min(n, 48));
    }
    Pool *p = new Pool;
    p->n = n;
    p->exc_pos.resize((size_t)MAX_COLUMNS * n);          // [column][thread]
    p->exc_val.resize((size_t)MAX_COLUMNS * n);
    for (int t = 1; t < n; t++) p->workers.emplace_back(worker, p, t);     // piece 0 runs on the calling thread
    return p;
}

void pool_destroy(Pool *p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(p->m);
        p->stop = true;
    }
    p->cv_go.notify_all();
    for (auto &t : p->workers) t.join();
    delete p;
}

int pool_threads(const Pool *p) { return p ? p->n : 0; }

static void run_all(Pool *p, const std::function<void(int)> &f) {
    if (p->n == 1) { f(0); return; }
    {
        std::lock_guard<std::mutex> lk(p->m);
        p->job = f;
        p->pending = p->n - 1;
        p->generation++;
    }
    p->cv_go.notify_all();
    f(0);
    std::unique_lock<std::mutex> lk(p->m);
    p->cv_done.wait(lk, [&] { return p->pending == 0; });
}

// ---- one piece: src[0..n) -> dst[0..(n+3)/4), exceptions appended with positions relative to `base` ----
static inline bool is_acgt(uint8_t b) { return b == 'A' || b == 'C' || b == 'G' || b == 'T'; }

static void pack_scalar(const uint8_t *src, size_t n, uint8_t *dst, size_t base, std::vector<uint32_t> &xp, std::vector<uint8_t> &xv, size_t cap) {
    size_t i = 0;
    for (; i + 4 <= n; i += 4) {
        uint8_t o = 0;
        for (int j = 0; j < 4; j++) {
            const uint8_t b = src[i + j];
            o |= (uint8_t)(((b >> 1) & 3u) << (2 * j));
            if (!is_acgt(b)) { xp.push_back((uint32_t)(base + i + j)); xv.push_back(b); }
        }
        dst[i >> 2] = o;
        if (xp.size() > cap) return;             // the caller gives up on this chunk anyway
    }
    if (i < n) {
        uint8_t o = 0;
        for (int j = 0; i + j < n; j++) {
            const uint8_t b = src[i + j];
            o |= (uint8_t)(((b >> 1) & 3u) << (2 * j));
            if (!is_acgt(b)) { xp.push_back((uint32_t)(base + i + j)); xv.push_back(b); }
        }
        dst[i >> 2] = o;
    }
}

#if defined(__x86_64__)
__attribute__((target("avx2")))
static void pack_avx2(const uint8_t *src, size_t n, uint8_t *dst, size_t base, std::vector<uint32_t> &xp, std::vector<uint8_t> &xv, size_t cap) {
    const __m256i three = _mm256_set1_epi8(3);
    const __m256i table = _mm256_setr_epi8('A', 'C', 'T', 'G', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 'A', 'C', 'T', 'G', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i mul1 = _mm256_set1_epi16(0x0401);          // bytes 1, 4: c0 + 4 * c1
    const __m256i mul2 = _mm256_set1_epi32(0x00100001);      // words 1, 16: t0 + 16 * t1
    const __m256i pick = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    size_t i = 0;
    for (; i + 32 <= n; i += 32) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i));
        const __m256i code = _mm256_and_si256(_mm256_srli_epi16(v, 1), three);
        const uint32_t bad = ~(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_shuffle_epi8(table, code), v));
        const __m256i t = _mm256_maddubs_epi16(code, mul1);
        const __m256i u = _mm256_shuffle_epi8(_mm256_madd_epi16(t, mul2), pick);
        const uint32_t lo = (uint32_t)_mm256_extract_epi32(u, 0), hi = (uint32_t)_mm256_extract_epi32(u, 4);
        const uint64_t out = (uint64_t)lo | ((uint64_t)hi << 32);
        memcpy(dst + (i >> 2), &out, 8);
        if (__builtin_expect(bad != 0u, 0)) {
            uint32_t m = bad;
            while (m) {
                const int j = __builtin_ctz(m);
                m &= m - 1;
                xp.push_back((uint32_t)(base + i + (size_t)j));
                xv.push_back(src[i + (size_t)j]);
            }
            if (xp.size() > cap) return;         // the caller gives up on this chunk anyway
        }
    }
    if (i < n) pack_scalar(src + i, n - i, dst + (i >> 2), base + i, xp, xv, cap);
}
#endif

// ---- qualities: 6 bits per byte, four codes -> three bytes ----
static void pack6_scalar(const uint8_t *src, size_t n, uint8_t *dst, size_t base, std::vector<uint32_t> &xp, std::vector<uint8_t> &xv, size_t cap) {
    for (size_t i = 0; i < n; i += 4) {
        uint32_t v = 0;
        for (size_t j = 0; j < 4 && i + j < n; j++) {
            const uint8_t b = src[i + j];
            const uint8_t c = (uint8_t)(b - 33);
            if (c > 63) { xp.push_back((uint32_t)(base + i + j)); xv.push_back(b); }
            v |= (uint32_t)(c & 63u) << (6 * j);
        }
        uint8_t *o = dst + 3 * (i >> 2);
        o[0] = (uint8_t)v; o[1] = (uint8_t)(v >> 8); o[2] = (uint8_t)(v >> 16);
        if (xp.size() > cap) return;
    }
}

#if defined(__x86_64__)
__attribute__((target("avx2")))
static void pack6_avx2(const uint8_t *src, size_t n, uint8_t *dst, size_t base, std::vector<uint32_t> &xp, std::vector<uint8_t> &xv, size_t cap) {
    const __m256i off = _mm256_set1_epi8(33), m63 = _mm256_set1_epi8(63), hi2 = _mm256_set1_epi8((char)0xC0), zero = _mm256_setzero_si256();
    const __m256i mul1 = _mm256_set1_epi16(0x4001);          // bytes 1, 64: a + 64 * b
    const __m256i mul2 = _mm256_set1_epi32(0x10000001);      // words 1, 4096: t0 + 4096 * t1
    const __m256i pick = _mm256_setr_epi8(0, 1, 2, 4, 5, 6, 8, 9, 10, 12, 13, 14, -1, -1, -1, -1, 0, 1, 2, 4, 5, 6, 8, 9, 10, 12, 13, 14, -1, -1, -1, -1);
    const __m256i join = _mm256_setr_epi32(0, 1, 2, 4, 5, 6, 7, 7);        // 12 + 12 bytes of the two lanes -> 24 contiguous bytes
    size_t i = 0;
    for (; i + 32 <= n; i += 32) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i));
        const __m256i c = _mm256_sub_epi8(v, off);
        const uint32_t bad = ~(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_and_si256(c, hi2), zero));
        const __m256i t = _mm256_maddubs_epi16(_mm256_and_si256(c, m63), mul1);
        const __m256i u = _mm256_permutevar8x32_epi32(_mm256_shuffle_epi8(_mm256_madd_epi16(t, mul2), pick), join);
        uint8_t *o = dst + 3 * (i >> 2);
        _mm_storeu_si128(reinterpret_cast<__m128i *>(o), _mm256_castsi256_si128(u));
        _mm_storel_epi64(reinterpret_cast<__m128i *>(o + 16), _mm256_extracti128_si256(u, 1));
        if (__builtin_expect(bad != 0u, 0)) {
            uint32_t m = bad;
            while (m) {
                const int j = __builtin_ctz(m);
                m &= m - 1;
                xp.push_back((uint32_t)(base + i + (size_t)j));
                xv.push_back(src[i + (size_t)j]);
            }
            if (xp.size() > cap) return;
        }
    }
    if (i < n) pack6_scalar(src + i, n - i, dst + 3 * (i >> 2), base + i, xp, xv, cap);
}
#endif

static void pack6_piece(const uint8_t *src, size_t n, uint8_t *dst, size_t base, std::vector<uint32_t> &xp, std::vector<uint8_t> &xv, size_t cap) {
#if defined(__x86_64__)
    if (__builtin_cpu_supports("avx2") && !getenv("AQC_PACK_SCALAR")) { pack6_avx2(src, n, dst, base, xp, xv, cap); return; }
#endif
    pack6_scalar(src, n, dst, base, xp, xv, cap);
}

static void pack_piece(const uint8_t *src, size_t n, uint8_t *dst, size_t base, std::vector<uint32_t> &xp, std::vector<uint8_t> &xv, size_t cap) {
#if defined(__x86_64__)
    if (__builtin_cpu_supports("avx2") && !getenv("AQC_PACK_SCALAR")) { pack_avx2(src, n, dst, base, xp, xv, cap); return; }   // once per piece
#endif
    pack_scalar(src, n, dst, base, xp, xv, cap);
}

void pack_columns(Pool *p, Column *cols, int n_cols) {
    const int T = p->n;
    if (n_cols > MAX_COLUMNS) n_cols = MAX_COLUMNS;
    // every thread takes one piece of every column: pieces of a multiple of 32 bytes (whole packed groups, whole vector iterations)
    size_t per[MAX_COLUMNS] = {0, 0, 0, 0};
    for (int c = 0; c < n_cols; c++) {
        per[c] = ((cols[c].n + (size_t)T - 1) / (size_t)T + 31) & ~(size_t)31;
        if (per[c] < 4096) per[c] = 4096;
        cols[c].n_exc = 0; cols[c].ok = true;
    }
    run_all(p, [&](int tid) {
        for (int c = 0; c < n_cols; c++) {
            std::vector<uint32_t> &xp = p->exc_pos[(size_t)(c * T + tid)];
            std::vector<uint8_t> &xv = p->exc_val[(size_t)(c * T + tid)];
            xp.clear(); xv.clear();
            const size_t lo = (size_t)tid * per[c];
            if (lo >= cols[c].n) continue;
            const size_t hi = std::min(cols[c].n, lo + per[c]);
            if (cols[c].kind == KIND_BASES) pack_piece(cols[c].src + lo, hi - lo, cols[c].dst + (lo >> 2), lo, xp, xv, cols[c].max_exc);
            else pack6_piece(cols[c].src + lo, hi - lo, cols[c].dst + 3 * (lo >> 2), lo, xp, xv, cols[c].max_exc);
        }
    });
    for (int c = 0; c < n_cols; c++) {
        size_t total = 0;
        for (int t = 0; t < T; t++) total += p->exc_pos[(size_t)(c * T + t)].size();
        if (total > cols[c].max_exc) { cols[c].ok = false; continue; }
        size_t w = 0;
        for (int t = 0; t < T; t++) {
            const std::vector<uint32_t> &xp = p->exc_pos[(size_t)(c * T + t)];
            const size_t k = xp.size();
            if (k) {
                memcpy(cols[c].exc_pos + w, xp.data(), k * sizeof(uint32_t));
                memcpy(cols[c].exc_val + w, p->exc_val[(size_t)(c * T + t)].data(), k);
                w += k;
            }
        }
        cols[c].n_exc = total;
    }
}

bool pack_bases(Pool *p, const uint8_t *src, size_t n, uint8_t *dst, uint32_t *exc_pos, uint8_t *exc_val, size_t max_exc, size_t *n_exc) {
    Column c{KIND_BASES, src, n, dst, exc_pos, exc_val, max_exc, 0, true};
    pack_columns(p, &c, 1);
    *n_exc = c.n_exc;
    return c.ok;
}

}  // namespace aqc_pack
