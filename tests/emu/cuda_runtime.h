// tests/emu/cuda_runtime.h -- TEST INFRASTRUCTURE ONLY.
//
// A host stand-in for <cuda_runtime.h> that lets g++ compile the engine's .cu/.cuh sources unchanged
// (macro -DAQC_EMU) and run the DEVICE code under a SIMT emulator (simt_emu.cpp): every CUDA thread is a
// cooperative fiber, one CTA at a time; warp collectives (__shfl_sync, __ballot_sync, __reduce_*_sync ...)
// and __syncthreads are rendezvous points between fibers, atomics are plain read-modify-writes, the
// mbarrier / bulk-copy (TMA) wrappers of aqc_device.cuh are modelled with their transaction counts.
// The resulting library (tests/emu/_build/libafterqc_b200_emu.so) exports the same C-ABI and exists so that
// the kernels' logic can be checked against the oracle on a machine without a GPU.  It is never built by
// afterqc_b200.build, never loaded by afterqc_b200._native and never shipped: the product has no CPU path.
#pragma once
#ifndef AQC_EMU
#error "tests/emu/cuda_runtime.h is only for the -DAQC_EMU test build"
#endif
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

// ---- qualifiers --------------------------------------------------------------------------------------
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static          // one CTA runs at a time
#define __restrict__ __restrict

struct uint2 { uint32_t x, y; };
struct uint3 { uint32_t x, y, z; };
struct __attribute__((aligned(16))) uint4 { uint32_t x, y, z, w; };
struct dim3 {
    uint32_t x, y, z;
    dim3(uint32_t a = 1, uint32_t b = 1, uint32_t c = 1) : x(a), y(b), z(c) {}
};
static inline uint4 make_uint4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { uint4 v; v.x = a; v.y = b; v.z = c; v.w = d; return v; }
static inline uint2 make_uint2(uint32_t a, uint32_t b) { uint2 v; v.x = a; v.y = b; return v; }

extern uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;

// ---- the emulator core (simt_emu.cpp) ------------------------------------------------------------------
namespace simt {
enum Op { OP_SHFL = 1, OP_SHFL_UP, OP_SHFL_DOWN, OP_SHFL_XOR, OP_BALLOT, OP_REDUX, OP_SYNCWARP, OP_MATCH };
const uint32_t *warp_xchg(uint32_t v, int op);      // rendezvous of the 32 lanes; returns every lane's value
void cta_barrier();
void wait_word_change(volatile uint32_t *w, uint32_t seen);   // suspend until *w != seen
uint8_t *dyn_smem();
int lane_id();
typedef void (*Entry)(void **args);
void launch(uint32_t grid, uint32_t block, size_t smem_bytes, Entry fn, void **args);
void fail(const char *msg);
extern uint64_t collectives;      // statistics: warp rendezvous executed (per lane)
}  // namespace simt

// ---- integer intrinsics ----------------------------------------------------------------------------
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((uint32_t)x); }
static inline uint32_t __brev(uint32_t x) {
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    return __builtin_bswap32(x);
}
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) {
    const uint64_t v = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        const uint32_t sel = (s >> (4 * i)) & 0xFu;
        uint32_t b = (uint32_t)(v >> (8 * (sel & 7u))) & 0xFFu;
        if (sel & 8u) b = (b & 0x80u) ? 0xFFu : 0x00u;
        r |= b << (8 * i);
    }
    return r;
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) {
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    return (uint32_t)(v >> (sh & 31u));
}
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t sh) {
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    return (uint32_t)((v << (sh & 31u)) >> 32);
}
static inline uint32_t __funnelshift_rc(uint32_t lo, uint32_t hi, uint32_t sh) {
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    return (uint32_t)(v >> (sh > 32u ? 32u : sh));
}
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
template <class T> static inline T __ldcg(const T *p) { return *p; }
template <class T> static inline T __ldg(const T *p) { return *p; }

static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, int b) { return a < (unsigned)b ? a : (unsigned)b; }
static inline unsigned min(int a, unsigned b) { return (unsigned)a < b ? (unsigned)a : b; }
static inline unsigned max(unsigned a, int b) { return a > (unsigned)b ? a : (unsigned)b; }
static inline unsigned max(int a, unsigned b) { return (unsigned)a > b ? (unsigned)a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
static inline unsigned long min(unsigned long a, unsigned long b) { return a < b ? a : b; }
static inline unsigned long max(unsigned long a, unsigned long b) { return a > b ? a : b; }

// ---- warp collectives (full mask only: the kernels keep collectives in warp-uniform control flow) --------
static inline void simt_check_mask(unsigned m) { if (m != 0xffffffffu) simt::fail("collective with a partial mask"); }
template <class T> static inline uint32_t simt_bits(T v) { static_assert(sizeof(T) == 4, "32-bit values only"); uint32_t u; memcpy(&u, &v, 4); return u; }
template <class T> static inline T simt_from(uint32_t u) { T v; memcpy(&v, &u, 4); return v; }

template <class T> static inline T __shfl_sync(unsigned m, T v, int src, int width = 32) {
    simt_check_mask(m);
    const uint32_t *a = simt::warp_xchg(simt_bits(v), simt::OP_SHFL);
    const int l = simt::lane_id();
    const int s = (l & ~(width - 1)) | (src & (width - 1));
    return simt_from<T>(a[s]);
}
template <class T> static inline T __shfl_up_sync(unsigned m, T v, unsigned d) {
    simt_check_mask(m);
    const uint32_t *a = simt::warp_xchg(simt_bits(v), simt::OP_SHFL_UP);
    const int l = simt::lane_id();
    return simt_from<T>(l >= (int)d ? a[l - (int)d] : a[l]);
}
template <class T> static inline T __shfl_down_sync(unsigned m, T v, unsigned d) {
    simt_check_mask(m);
    const uint32_t *a = simt::warp_xchg(simt_bits(v), simt::OP_SHFL_DOWN);
    const int l = simt::lane_id();
    return simt_from<T>(l + (int)d < 32 ? a[l + (int)d] : a[l]);
}
template <class T> static inline T __shfl_xor_sync(unsigned m, T v, int x) {
    simt_check_mask(m);
    const uint32_t *a = simt::warp_xchg(simt_bits(v), simt::OP_SHFL_XOR);
    return simt_from<T>(a[(simt::lane_id() ^ x) & 31]);
}
static inline unsigned __ballot_sync(unsigned m, int pred) {
    simt_check_mask(m);
    const uint32_t *a = simt::warp_xchg(pred ? 1u : 0u, simt::OP_BALLOT);
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= (a[i] & 1u) << i;
    return r;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0u; }
static inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }
// The call site is part of the collective's identity here: hardware would pair __syncwarp()s of different source lines, but in
// these kernels that is always a bug (lanes that took different decisions about a warp-wide step), and this check finds it.
static inline void simt_syncwarp_at(int line, unsigned m) { simt_check_mask(m); simt::warp_xchg(0, simt::OP_SYNCWARP + (line << 8)); }
static inline unsigned simt_mask_or_full() { return 0xffffffffu; }
static inline unsigned simt_mask_or_full(unsigned m) { return m; }
#define __syncwarp(...) simt_syncwarp_at(__LINE__, simt_mask_or_full(__VA_ARGS__))
static inline unsigned __reduce_add_sync(unsigned m, unsigned v) {
    simt_check_mask(m);
    const uint32_t *a = simt::warp_xchg(v, simt::OP_REDUX);
    unsigned r = 0; for (int i = 0; i < 32; i++) r += a[i];
    return r;
}
static inline int __reduce_add_sync(unsigned m, int v) { return (int)__reduce_add_sync(m, (unsigned)v); }
static inline unsigned __reduce_max_sync(unsigned m, unsigned v) {
    simt_check_mask(m);
    const uint32_t *a = simt::warp_xchg(v, simt::OP_REDUX);
    unsigned r = 0; for (int i = 0; i < 32; i++) r = a[i] > r ? a[i] : r;
    return r;
}
static inline unsigned __reduce_min_sync(unsigned m, unsigned v) {
    simt_check_mask(m);
    const uint32_t *a = simt::warp_xchg(v, simt::OP_REDUX);
    unsigned r = 0xffffffffu; for (int i = 0; i < 32; i++) r = a[i] < r ? a[i] : r;
    return r;
}
static inline unsigned __reduce_or_sync(unsigned m, unsigned v) {
    simt_check_mask(m);
    const uint32_t *a = simt::warp_xchg(v, simt::OP_REDUX);
    unsigned r = 0; for (int i = 0; i < 32; i++) r |= a[i];
    return r;
}
static inline unsigned __match_any_sync(unsigned m, unsigned v) {
    simt_check_mask(m);
    const uint32_t *a = simt::warp_xchg(v, simt::OP_MATCH);
    unsigned r = 0; for (int i = 0; i < 32; i++) if (a[i] == v) r |= 1u << i;
    return r;
}
static inline void __syncthreads() { simt::cta_barrier(); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

// ---- atomics (fibers are cooperative: no real concurrency) --------------------------------------------
template <class T, class U> static inline T atomicAdd(T *p, U v) { T o = *p; *p = (T)(o + (T)v); return o; }
template <class T, class U> static inline T atomicSub(T *p, U v) { T o = *p; *p = (T)(o - (T)v); return o; }
template <class T, class U> static inline T atomicMin(T *p, U v) { T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <class T, class U> static inline T atomicMax(T *p, U v) { T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <class T, class U> static inline T atomicExch(T *p, U v) { T o = *p; *p = (T)v; return o; }
template <class T, class U> static inline T atomicOr(T *p, U v) { T o = *p; *p = (T)(o | (T)v); return o; }
template <class T, class U> static inline T atomicAnd(T *p, U v) { T o = *p; *p = (T)(o & (T)v); return o; }
template <class T, class U, class V> static inline T atomicCAS(T *p, U cmp, V v) { T o = *p; if (o == (T)cmp) *p = (T)v; return o; }

// ---- mbarrier + bulk copy model used by the AQC_EMU branch of aqc_device.cuh ---------------------------
// 64-bit barrier word: bit 63 phase | bits 48..62 init count | bits 32..47 pending arrivals | bits 0..31 tx bytes
namespace simt {
static inline void mbar_settle(uint64_t *bar) {
    uint64_t b = *bar;
    const uint32_t pending = (uint32_t)(b >> 32) & 0xFFFFu, tx = (uint32_t)b;
    if (pending == 0 && tx == 0) {
        const uint64_t init = (b >> 48) & 0x7FFFu;
        *bar = ((b ^ (1ull << 63)) & (1ull << 63)) | (init << 48) | (init << 32);
    }
}
static inline void mbar_init(uint64_t *bar, uint32_t count) { *bar = ((uint64_t)count << 48) | ((uint64_t)count << 32); }
static inline void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    uint64_t b = *bar;
    const uint32_t pending = (uint32_t)(b >> 32) & 0xFFFFu;
    if (pending == 0) fail("mbarrier: arrive on a barrier with no pending arrivals");
    const uint32_t tx = (uint32_t)b + bytes;
    *bar = (b & 0xFFFF000000000000ull) | ((uint64_t)(pending - 1) << 32) | tx;
    mbar_settle(bar);
}
static inline void mbar_arrive(uint64_t *bar) { mbar_arrive_expect_tx(bar, 0); }
static inline void mbar_complete_tx(uint64_t *bar, uint32_t bytes) {
    uint64_t b = *bar;
    const uint32_t tx = (uint32_t)b;
    if (tx < bytes) fail("mbarrier: more bytes completed than expected (expect_tx must precede the copies)");
    *bar = (b & 0xFFFFFFFF00000000ull) | (tx - bytes);
    mbar_settle(bar);
}
static inline bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t hi = (uint32_t)(*bar >> 32);
    if ((hi >> 31) != (parity & 1u)) return true;
    wait_word_change(reinterpret_cast<volatile uint32_t *>(bar) + 1, hi);      // little endian: high word
    return false;
}
// the copy itself is performed at issue time; 16-byte granularity as the hardware demands
static inline void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    if ((reinterpret_cast<uintptr_t>(dst) & 15u) || (reinterpret_cast<uintptr_t>(src) & 15u) || (bytes & 15u) || bytes == 0)
        fail("cp.async.bulk: dst, src and size must be non-zero multiples of 16");
    memcpy(dst, src, bytes);
    mbar_complete_tx(bar, bytes);
}
}  // namespace simt

// ---- the handful of runtime calls the engine makes (all synchronous here) ------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
typedef struct simt_stream *cudaStream_t;
typedef struct simt_event *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0 };
enum cudaDeviceAttr { cudaDevAttrMaxSharedMemoryPerBlockOptin = 97 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
enum { cudaSharedmemCarveoutMaxShared = 100 };
struct cudaDeviceProp { int multiProcessorCount; char name[64]; };
struct cudaFuncAttributes { size_t sharedSizeBytes; int numRegs; };

enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void *devicePointer; void *hostPointer; };
// every host pointer is "device accessible" here unless AQC_EMU_PAGEABLE=1 (to exercise the copy fall-back)
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *p) {
    const bool pageable = getenv("AQC_EMU_PAGEABLE") != nullptr;
    a->type = pageable ? cudaMemoryTypeUnregistered : cudaMemoryTypeHost;
    a->device = 0; a->devicePointer = pageable ? nullptr : const_cast<void *>(p); a->hostPointer = const_cast<void *>(p);
    return cudaSuccess;
}
static inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline int simt_env_int(const char *name, int dflt) { const char *s = getenv(name); return s ? atoi(s) : dflt; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { memset(p, 0, sizeof *p); p->multiProcessorCount = simt_env_int("AQC_EMU_SMS", 3); return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr, int) { *v = 232448; return cudaSuccess; }
static inline cudaError_t cudaFuncGetAttributes(cudaFuncAttributes *a, const void *) { a->sharedSizeBytes = 1024; a->numRegs = 64; return cudaSuccess; }
static inline cudaError_t cudaFuncSetAttribute(const void *, cudaFuncAttribute, int) { return cudaSuccess; }
static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, const void *, int, size_t) { *n = simt_env_int("AQC_EMU_OCC", 2); return cudaSuccess; }
// device allocations end (16-byte granular) right at an inaccessible guard page: a kernel or bulk copy that reads or writes
// 16 bytes or more past the end of a buffer faults under the emulator (the GPU would read garbage or raise an illegal address)
namespace simt { void *guarded_alloc(size_t n); void guarded_free(void *p); }
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t n) { *p = (T *)simt::guarded_alloc(n); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void *p) { simt::guarded_free(p); return cudaSuccess; }
static inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { *p = aligned_alloc(256, (n + 255) & ~(size_t)255); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (cudaStream_t)malloc(8); return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (cudaEvent_t)malloc(8); return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = (cudaEvent_t)malloc(8); return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
// kernels are launched through trampolines void(void **args) (see AQC_EMU in aqc_engine.cu)
static inline cudaError_t cudaLaunchKernel(const void *fn, dim3 grid, dim3 block, void **args, size_t smem, cudaStream_t) {
    simt::launch(grid.x, block.x, smem, (simt::Entry)fn, args);
    return cudaSuccess;
}
