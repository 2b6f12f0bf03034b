// tests/emu/simt_emu.cpp -- TEST INFRASTRUCTURE ONLY: the fiber scheduler behind tests/emu/cuda_runtime.h.
//
// One CTA at a time; each CUDA thread is a cooperative fiber on its own stack (hand-rolled x86-64 context
// switch).  A fiber runs until it reaches a rendezvous (warp collective, __syncthreads, mbarrier wait) that is
// not complete yet; the scheduler then resumes the next fiber whose wait condition has changed.  A pass over
// all fibers without progress is a deadlock (on the GPU: a hang) and aborts with a diagnosis.
#include <sys/mman.h>
#include <unistd.h>
#include <map>
#include <vector>
#include "cuda_runtime.h"

#if !defined(__x86_64__)
#error "the SIMT emulator's context switch is written for x86-64"
#endif

uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;

extern "C" void simt_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl simt_switch
.type simt_switch,@function
simt_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size simt_switch,.-simt_switch
)");

namespace simt {

uint64_t collectives = 0;

namespace {

constexpr size_t STACK_BYTES = 512 * 1024;
constexpr int MAX_THREADS = 1024;

struct Fiber {
    void *sp = nullptr;
    char *stack = nullptr;
    int tid = 0;
    bool done = false;
    volatile uint32_t *wait_ptr = nullptr;
    uint32_t wait_val = 0;
    const char *where = "";
};

struct Warp {
    uint32_t in[32];
    uint32_t out[2][32];
    int op[32];
    int arrived = 0;
    uint32_t gen = 0;
};

Fiber fibers[MAX_THREADS];
Warp warps[MAX_THREADS / 32];
int n_threads = 0;
Fiber *cur = nullptr;
void *main_sp = nullptr;
Entry g_entry = nullptr;
void **g_args = nullptr;
uint8_t *g_smem = nullptr;
size_t g_smem_cap = 0;
int cta_arrived = 0;
uint32_t cta_gen = 0;

void to_main() { simt_switch(&cur->sp, main_sp); }

void block_on(volatile uint32_t *p, uint32_t v, const char *where) {
    cur->wait_ptr = p; cur->wait_val = v; cur->where = where;
    to_main();
}

void fiber_main() {
    g_entry(g_args);
    cur->done = true;
    to_main();
    fail("resumed a finished fiber");
}

void prepare(Fiber &f, int tid) {
    if (!f.stack) {
        void *m = mmap(nullptr, STACK_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (m == MAP_FAILED) fail("mmap of a fiber stack failed");
        mprotect(m, 4096, PROT_NONE);                          // guard page
        f.stack = (char *)m;
    }
    f.tid = tid; f.done = false; f.wait_ptr = nullptr; f.where = "";
    uintptr_t top = ((uintptr_t)f.stack + STACK_BYTES) & ~(uintptr_t)15;
    uint64_t *s = (uint64_t *)top;
    *--s = 0;                                   // fake return address of fiber_main (never used)
    *--s = (uint64_t)(uintptr_t)&fiber_main;    // popped by simt_switch's ret
    for (int i = 0; i < 6; i++) *--s = 0;       // r15 r14 r13 r12 rbx rbp
    f.sp = s;
}

}  // namespace

void fail(const char *msg) {
    fprintf(stderr, "[simt emulator] %s", msg);
    if (cur) fprintf(stderr, " (block %u thread %d)", blockIdx.x, cur->tid);
    fprintf(stderr, "\n");
    fflush(stderr);
    abort();
}

namespace {
std::map<void *, std::pair<void *, size_t>> g_allocs;     // user pointer -> (mapping base, mapping length)
}

void *guarded_alloc(size_t n) {
    const size_t page = (size_t)sysconf(_SC_PAGESIZE);
    const size_t body = ((n ? n : 1) + 15) & ~(size_t)15;
    const size_t len = ((body + page - 1) / page) * page + page;
    void *m = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (m == MAP_FAILED) return nullptr;
    char *guard = (char *)m + len - page;
    mprotect(guard, page, PROT_NONE);
    void *p = guard - body;
    memset(p, 0xA5, body);                                   // device memory is not zero-initialised
    g_allocs[p] = {m, len};
    return p;
}

void guarded_free(void *p) {
    if (!p) return;
    auto it = g_allocs.find(p);
    if (it == g_allocs.end()) fail("cudaFree of a pointer cudaMalloc did not return");
    munmap(it->second.first, it->second.second);
    g_allocs.erase(it);
}

int lane_id() { return cur->tid & 31; }
uint8_t *dyn_smem() { return g_smem; }

const uint32_t *warp_xchg(uint32_t v, int op) {
    Warp &W = warps[cur->tid >> 5];
    const int l = cur->tid & 31;
    collectives++;
    W.in[l] = v; W.op[l] = op;
    const uint32_t g = W.gen;
    if (++W.arrived == 32) {
        for (int i = 1; i < 32; i++)
            if (W.op[i] != W.op[0]) fail("lanes of one warp met in different collectives (divergent control flow around a *_sync)");
        memcpy(W.out[g & 1u], W.in, sizeof W.in);
        W.arrived = 0;
        W.gen = g + 1;
    } else {
        block_on(&W.gen, g, "warp collective");
    }
    return W.out[g & 1u];
}

void cta_barrier() {
    const uint32_t g = cta_gen;
    if (++cta_arrived == n_threads) { cta_arrived = 0; cta_gen = g + 1; }
    else block_on(&cta_gen, g, "__syncthreads");
}

void wait_word_change(volatile uint32_t *w, uint32_t seen) { block_on(w, seen, "mbarrier wait"); }

void launch(uint32_t grid, uint32_t block, size_t smem_bytes, Entry fn, void **args) {
    if (block == 0 || block > MAX_THREADS || (block & 31u)) fail("block size must be a multiple of 32 and <= 1024");
    if (smem_bytes > 232448) fail("dynamic shared memory beyond the 227 KB opt-in limit");
    {   // dynamic shared memory ends (128-byte granular) at a guard page: reads past the launch's allocation fault
        static void *map_base = nullptr; static size_t map_len = 0;
        const size_t page = (size_t)sysconf(_SC_PAGESIZE);
        const size_t body = (smem_bytes + 127) & ~(size_t)127;
        if (map_base) munmap(map_base, map_len);
        map_len = ((body + page - 1) / page) * page + 2 * page;
        map_base = mmap(nullptr, map_len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (map_base == MAP_FAILED) fail("mmap of shared memory failed");
        char *guard = (char *)map_base + map_len - page;
        mprotect(guard, page, PROT_NONE);
        g_smem = (uint8_t *)(guard - body);
        g_smem_cap = body;
    }
    g_entry = fn; g_args = args;
    n_threads = (int)block;
    gridDim = dim3(grid); blockDim = dim3(block);
    for (uint32_t b = 0; b < grid; b++) {
        blockIdx.x = b; blockIdx.y = blockIdx.z = 0;
        memset(g_smem, 0xCD, g_smem_cap);                      // shared memory starts uninitialised on the device
        for (int w = 0; w < (int)block / 32; w++) { warps[w].arrived = 0; warps[w].gen = 0; }
        cta_arrived = 0; cta_gen = 0;
        for (int t = 0; t < (int)block; t++) prepare(fibers[t], t);
        int remaining = (int)block;
        // fiber order of a scheduling pass: AQC_EMU_SCHED=reverse|random shakes out code that relies on lane 0 (or warp 0)
        // running first; the default is thread order
        static const char *sched = getenv("AQC_EMU_SCHED");
        static uint64_t rng = 0x9E3779B97F4A7C15ull;
        static int order[MAX_THREADS];
        for (int t = 0; t < (int)block; t++) order[t] = (sched && sched[0] == 'r' && sched[1] == 'e') ? (int)block - 1 - t : t;
        while (remaining > 0) {
            bool progressed = false;
            if (sched && sched[0] == 'r' && sched[1] == 'a') {
                for (int t = (int)block - 1; t > 0; t--) {
                    rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
                    const int j = (int)(rng % (uint64_t)(t + 1));
                    const int tmp = order[t]; order[t] = order[j]; order[j] = tmp;
                }
            }
            for (int oi = 0; oi < (int)block; oi++) {
                const int t = order[oi];
                Fiber &f = fibers[t];
                if (f.done) continue;
                if (f.wait_ptr) {
                    if (*f.wait_ptr == f.wait_val) continue;
                    f.wait_ptr = nullptr;
                }
                cur = &f;
                threadIdx.x = (uint32_t)t; threadIdx.y = threadIdx.z = 0;
                simt_switch(&main_sp, f.sp);
                progressed = true;
                if (f.done) remaining--;
            }
            if (!progressed) {
                fprintf(stderr, "[simt emulator] deadlock in block %u: ", b);
                int shown = 0;
                for (int t = 0; t < (int)block && shown < 8; t++)
                    if (!fibers[t].done) { fprintf(stderr, "thread %d waits in %s; ", t, fibers[t].where); shown++; }
                fprintf(stderr, "\n");
                abort();
            }
        }
        cur = nullptr;
    }
}

}  // namespace simt
