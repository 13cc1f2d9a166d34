// cuda_runtime.h (EMULATOR) -- TEST INFRASTRUCTURE ONLY, never part of the product.
//
// A minimal SIMT emulator that lets g++ compile osm_renderer_b200/csrc/*.cu(h) unchanged and run the kernels on the
// host, so that kernel LOGIC changes (warp-synchronous control flow, closed forms, compaction) can be checked against
// the oracle in the GPU-less build container before a GPU box is spent on them:
//   * every CUDA thread of a block is a fiber (own stack, hand-written context switch); blocks run one after another;
//   * warp collectives (__ballot_sync, __shfl*_sync, __any_sync, __syncwarp) and __syncthreads are rendezvous points:
//     a lane deposits its operand and yields until all live lanes of the warp / block have arrived.  Lanes that meet at
//     DIFFERENT call sites are reported (divergent collective: undefined behaviour / a hang on the real GPU);
//   * a watchdog aborts a block whose fibers spin without any lane making progress through a rendezvous (deadlock);
//   * __shared__ is a function-local static (blocks are sequential), atomics are plain read-modify-writes;
//   * the runtime API (cudaMalloc, cudaMemcpyAsync, streams, events) maps to malloc / memcpy / no-ops.
// Arithmetic is the host's IEEE f64 (compile with -ffp-contract=off, the analogue of nvcc -fmad=false); tan/log come
// from glibc instead of libdevice.  tests/emu/build_emu.py rewrites `k<<<g, b, s, st>>>(args)` into EMU_LAUNCH.
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <system_error>
#include <mutex>
#include <new>
#include <shared_mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <utility>
#include <vector>

#define OSMR_EMULATED 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __shared__ static
#define __constant__ static
#define __launch_bounds__(...)

// ---------------------------------------------------------------------------------------------------------
// vector types (CUDA alignments)
// ---------------------------------------------------------------------------------------------------------
struct alignas(8) int2 { int x, y; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) short4 { short x, y, z, w; };
struct alignas(4) uchar4 { unsigned char x, y, z, w; };
struct alignas(16) double2 { double x, y; };
struct alignas(16) double4 { double x, y, z, w; };
struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline short4 make_short4(short x, short y, short z, short w) { return short4{x, y, z, w}; }
static inline uchar4 make_uchar4(unsigned char x, unsigned char y, unsigned char z, unsigned char w) { return uchar4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline double4 make_double4(double x, double y, double z, double w) { return double4{x, y, z, w}; }

// CUDA's global min / max overloads
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
static inline long min(long a, long b) { return a < b ? a : b; }
static inline long max(long a, long b) { return a > b ? a : b; }
static inline unsigned long min(unsigned long a, unsigned long b) { return a < b ? a : b; }
static inline unsigned long max(unsigned long a, unsigned long b) { return a > b ? a : b; }
static inline long long min(long long a, int b) { return a < b ? a : b; }
static inline long long max(long long a, int b) { return a > b ? a : b; }
static inline long long min(int a, long long b) { return a < b ? a : b; }
static inline long long max(int a, long long b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, int b) { return a < (unsigned)b ? a : (unsigned)b; }
static inline unsigned max(unsigned a, int b) { return a > (unsigned)b ? a : (unsigned)b; }
static inline unsigned min(int a, unsigned b) { return (unsigned)a < b ? (unsigned)a : b; }
static inline unsigned max(int a, unsigned b) { return (unsigned)a > b ? (unsigned)a : b; }
static inline double min(double a, double b) { return std::fmin(a, b); }
static inline double max(double a, double b) { return std::fmax(a, b); }

using std::abs;
using std::ceil;
using std::fabs;
using std::floor;
using std::fmax;
using std::fmin;
using std::fmod;
using std::llabs;
using std::log;
using std::round;
using std::sqrt;
using std::tan;

// ---------------------------------------------------------------------------------------------------------
// scalar intrinsics
// ---------------------------------------------------------------------------------------------------------
static inline int __double2int_rz(double v) {  // cvt.rzi.s32.f64: saturating, NaN -> 0
    if (v != v) return 0;
    if (v >= 2147483647.0) return 2147483647;
    if (v <= -2147483648.0) return (int)0x80000000;
    return (int)v;
}
static inline void __threadfence_system() {}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
static inline unsigned __brev(unsigned v) {
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
    return __builtin_bswap32(v);
}
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) {
    return (unsigned long long)(((unsigned __int128)a * (unsigned __int128)b) >> 64);
}
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline double __longlong_as_double(long long v) {
    double d;
    memcpy(&d, &v, 8);
    return d;
}
static inline long long __double_as_longlong(double d) {
    long long v;
    memcpy(&v, &d, 8);
    return v;
}

// ---------------------------------------------------------------------------------------------------------
// the SIMT engine
// ---------------------------------------------------------------------------------------------------------
extern "C" void emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
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
.size emu_switch,.-emu_switch
)");

namespace emu {

constexpr int kMaxThreads = 1024;
constexpr size_t kStack = 256 * 1024;

struct Rendezvous {  // one per warp, one per block
    unsigned long long val[2][kMaxThreads];
    int site[2][kMaxThreads];
    unsigned arrived_mask[2][kMaxThreads / 32];
    int count = 0, live = 0, size = 32;
    unsigned long long gen = 0;
};

struct Fiber {
    void* sp = nullptr;
    char* stack = nullptr;
    bool done = true;
};

struct Engine {
    Fiber fib[kMaxThreads];
    Rendezvous warp[kMaxThreads / 32];
    Rendezvous block;
    void* sched_sp = nullptr;
    int cur = 0, n_threads = 0;
    const std::function<void()>* body = nullptr;
    unsigned long long progress = 0;  // completed rendezvous + exited threads
    unsigned long long divergent = 0;
    const char* kernel = "";
};
inline Engine& E() {
    static Engine e;
    return e;
}

}  // namespace emu

struct EmuIdx {
    unsigned x = 0, y = 0, z = 0;
};
inline EmuIdx threadIdx, blockIdx, blockDim, gridDim;

namespace emu {

inline void yield() {
    Engine& e = E();
    emu_switch(&e.fib[e.cur].sp, e.sched_sp);
}

inline void complete(Rendezvous& r) {
    r.count = 0;
    r.gen++;
    E().progress++;
}

// Deposit `v` at call site `site`, wait for every live thread of the group (first..first+n), return the buffer index.
inline int rendezvous(Rendezvous& r, int slot, int site, unsigned long long v, const char* what) {
    Engine& e = E();
    const unsigned long long g = r.gen;
    const int b = (int)(g & 1);
    if (r.count == 0) memset(r.arrived_mask[b], 0, sizeof r.arrived_mask[b]);
    r.val[b][slot] = v;
    r.site[b][slot] = site;
    r.arrived_mask[b][slot >> 5] |= 1u << (slot & 31);
    r.count++;
    if (r.count >= r.live) {
        int first_site = site;
        for (int i = 0; i < r.size; ++i)
            if ((r.arrived_mask[b][i >> 5] >> (i & 31)) & 1u)
                if (r.site[b][i] != first_site) {
                    if (e.divergent++ < 20)
                        fprintf(stderr, "[emu] %s: divergent %s: lanes met at source lines %d and %d (block %u)\n", e.kernel, what,
                                first_site, r.site[b][i], blockIdx.x);
                    break;
                }
        complete(r);
    } else {
        while (r.gen == g) yield();
    }
    return b;
}

inline void thread_exit_hook() {
    Engine& e = E();
    const int t = e.cur;
    Rendezvous& w = e.warp[t >> 5];
    w.live--;
    if (w.count > 0 && w.count >= w.live) complete(w);
    e.block.live--;
    if (e.block.count > 0 && e.block.count >= e.block.live) complete(e.block);
    e.progress++;
}

inline void fiber_main() {
    Engine& e = E();
    (*e.body)();
    e.fib[e.cur].done = true;
    thread_exit_hook();
    for (;;) yield();
}

inline void run_block(const std::function<void()>& body, int n_threads) {
    Engine& e = E();
    e.body = &body;
    e.n_threads = n_threads;
    for (int w = 0; w < (n_threads + 31) / 32; ++w) {
        e.warp[w].count = 0;
        e.warp[w].gen = 0;
        e.warp[w].live = std::min(32, n_threads - 32 * w);
    }
    e.block.count = 0;
    e.block.gen = 0;
    e.block.live = n_threads;
    e.block.size = n_threads;
    for (int t = 0; t < n_threads; ++t) {
        Fiber& f = e.fib[t];
        if (!f.stack) f.stack = (char*)aligned_alloc(64, kStack);
        uintptr_t top = ((uintptr_t)f.stack + kStack) & ~(uintptr_t)15;
        void** sp = (void**)top;
        *--sp = nullptr;               // fake return address of fiber_main
        *--sp = (void*)&fiber_main;    // popped by emu_switch's ret
        for (int i = 0; i < 6; ++i) *--sp = nullptr;  // callee-saved registers
        f.sp = sp;
        f.done = false;
    }
    int alive = n_threads;
    unsigned long long last_progress = e.progress;
    unsigned long long idle_rounds = 0;
    while (alive) {
        alive = 0;
        for (int t = 0; t < n_threads; ++t) {
            if (e.fib[t].done) continue;
            e.cur = t;
            threadIdx.x = (unsigned)t;
            emu_switch(&e.sched_sp, e.fib[t].sp);
            if (!e.fib[t].done) ++alive;
        }
        if (e.progress == last_progress) {
            if (++idle_rounds > 4) {
                fprintf(stderr, "[emu] %s: DEADLOCK in block %u: %d threads wait at a rendezvous that cannot complete\n", e.kernel,
                        blockIdx.x, alive);
                for (int w = 0; w < (n_threads + 31) / 32; ++w) {
                    Rendezvous& r = e.warp[w];
                    if (r.count) {
                        const int b = (int)(r.gen & 1);
                        fprintf(stderr, "  warp %d: %d of %d live lanes arrived; sites:", w, r.count, r.live);
                        for (int i = 0; i < 32; ++i)
                            if ((r.arrived_mask[b][0] >> i) & 1u) fprintf(stderr, " %d:%d", i, r.site[b][i]);
                        fprintf(stderr, "\n");
                    }
                }
                abort();
            }
        } else {
            idle_rounds = 0;
            last_progress = e.progress;
        }
    }
}

template <typename F>
inline void launch(const char* name, dim3 grid, dim3 block, F&& f) {
    Engine& e = E();
    e.kernel = name;
    const std::function<void()> body(std::forward<F>(f));
    gridDim.x = grid.x;
    blockDim.x = block.x;
    if (block.x > (unsigned)kMaxThreads || grid.y != 1 || block.y != 1) {
        fprintf(stderr, "[emu] unsupported launch shape\n");
        abort();
    }
    for (unsigned b = 0; b < grid.x; ++b) {
        blockIdx.x = b;
        run_block(body, (int)block.x);
    }
}

inline int cur_lane() { return E().cur & 31; }
inline Rendezvous& cur_warp() { return E().warp[E().cur >> 5]; }

inline unsigned ballot(int site, bool p) {
    Rendezvous& r = cur_warp();
    const int b = rendezvous(r, cur_lane(), site, p ? 1ull : 0ull, "warp collective");
    unsigned m = 0;
    for (int i = 0; i < 32; ++i)
        if (((r.arrived_mask[b][0] >> i) & 1u) && r.val[b][i]) m |= 1u << i;
    return m;
}
template <typename T>
inline T shfl(int site, T v, int src) {
    static_assert(sizeof(T) <= 8, "shfl operand");
    unsigned long long bits = 0;
    memcpy(&bits, &v, sizeof(T));
    Rendezvous& r = cur_warp();
    const int b = rendezvous(r, cur_lane(), site, bits, "warp collective");
    src &= 31;
    unsigned long long got = ((r.arrived_mask[b][0] >> src) & 1u) ? r.val[b][src] : bits;  // inactive source: undefined on the GPU
    T out;
    memcpy(&out, &got, sizeof(T));
    return out;
}
inline unsigned match_any(int site, unsigned long long v) {
    Rendezvous& r = cur_warp();
    const int b = rendezvous(r, cur_lane(), site, v, "warp collective");
    unsigned m = 0;
    for (int i = 0; i < 32; ++i)
        if (((r.arrived_mask[b][0] >> i) & 1u) && r.val[b][i] == v) m |= 1u << i;
    return m;
}
inline void syncwarp(int site) { rendezvous(cur_warp(), cur_lane(), site, 0, "__syncwarp"); }
inline void syncthreads(int site) { rendezvous(E().block, E().cur, site, 0, "__syncthreads"); }

}  // namespace emu

#define __ballot_sync(m, p) emu::ballot(__LINE__, (p))
#define __any_sync(m, p) (emu::ballot(__LINE__, (p)) != 0u)
#define __all_sync(m, p) (emu::ballot(__LINE__, !(p)) == 0u)
#define __shfl_sync(m, v, src) emu::shfl(__LINE__, (v), (int)(src))
#define __shfl_up_sync(m, v, d) emu::shfl(__LINE__, (v), emu::cur_lane() >= (int)(d) ? emu::cur_lane() - (int)(d) : emu::cur_lane())
#define __shfl_down_sync(m, v, d) emu::shfl(__LINE__, (v), emu::cur_lane() + (int)(d) < 32 ? emu::cur_lane() + (int)(d) : emu::cur_lane())
#define __shfl_xor_sync(m, v, x) emu::shfl(__LINE__, (v), emu::cur_lane() ^ (int)(x))
#define __match_any_sync(m, v) emu::match_any(__LINE__, (unsigned long long)(v))
#define __syncwarp() emu::syncwarp(__LINE__)
// REDUX (sm_80+): warp reductions of 32-bit integers
static inline unsigned emu_reduce_add(unsigned v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
static inline int emu_reduce_min(int v) {
    for (int o = 16; o > 0; o >>= 1) {
        const int y = __shfl_xor_sync(0xffffffffu, v, o);
        v = y < v ? y : v;
    }
    return v;
}
static inline int emu_reduce_max(int v) {
    for (int o = 16; o > 0; o >>= 1) {
        const int y = __shfl_xor_sync(0xffffffffu, v, o);
        v = y > v ? y : v;
    }
    return v;
}
#define __reduce_add_sync(m, v) emu_reduce_add(v)
#define __reduce_min_sync(m, v) emu_reduce_min(v)
#define __reduce_max_sync(m, v) emu_reduce_max(v)
#define __syncthreads() emu::syncthreads(__LINE__)

template <typename T>
static inline T atomicAdd(T* p, T v) {
    T old = *p;
    *p = old + v;
    return old;
}
template <typename T>
static inline T atomicOr(T* p, T v) {
    T old = *p;
    *p = old | v;
    return old;
}
template <typename T>
static inline T atomicMax(T* p, T v) {
    T old = *p;
    if (v > old) *p = v;
    return old;
}
template <typename T>
static inline T atomicMin(T* p, T v) {
    T old = *p;
    if (v < old) *p = v;
    return old;
}
static inline int atomicOr(volatile int* p, int v) {
    int old = *p;
    *p = old | v;
    return old;
}

// named event counters for divergence / work statistics (printed at exit when OSMR_EMU_STATS is set); the kernels use
// them through OSMR_COUNT(name, n), which is a no-op in the CUDA build
namespace emu {
struct Counters {
    std::vector<std::pair<std::string, unsigned long long>> c;
    ~Counters() {
        if (getenv("OSMR_EMU_STATS"))
            for (auto& kv : c) fprintf(stderr, "[emu-stat] %-40s %llu\n", kv.first.c_str(), kv.second);
    }
    unsigned long long& at(const char* name) {
        for (auto& kv : c)
            if (kv.first == name) return kv.second;
        c.emplace_back(name, 0ull);
        return c.back().second;
    }
};
inline Counters& counters() {
    static Counters k;
    return k;
}
}  // namespace emu
#define OSMR_COUNT(name, n) (emu::counters().at(name) += (unsigned long long)(n))

#define EMU_LAUNCH(kernel, grid, block, ...) emu::launch(#kernel, dim3(grid), dim3(block), [=]() { kernel(__VA_ARGS__); })

// ---------------------------------------------------------------------------------------------------------
// runtime API
// ---------------------------------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
typedef struct EmuStream* cudaStream_t;
typedef struct EmuEvent* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaDevAttrMultiProcessorCount = 16 };
struct EmuEvent {
    std::chrono::steady_clock::time_point t;
};
static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) {
    *n = 1;
    return cudaSuccess;
}
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) {
    *v = 2;  // "SMs": persistent kernels launch a handful of blocks
    return cudaSuccess;
}
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) {
    *lo = 0;
    *hi = -1;
    return cudaSuccess;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned);
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned f, int) { return cudaStreamCreateWithFlags(s, f); }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) {
    *s = (cudaStream_t)(uintptr_t)1;
    return cudaSuccess;
}
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) {
    *e = new EmuEvent();
    return cudaSuccess;
}
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) {
    delete e;
    return cudaSuccess;
}
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) {
    e->t = std::chrono::steady_clock::now();
    return cudaSuccess;
}
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
    return cudaSuccess;
}
template <typename T>
static inline cudaError_t cudaMalloc(T** p, size_t bytes) {
    *p = (T*)aligned_alloc(256, (bytes + 255) & ~(size_t)255);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
template <typename T>
static inline cudaError_t cudaMallocHost(T** p, size_t bytes) {
    return cudaMalloc(p, bytes);
}
static inline cudaError_t cudaFree(void* p) {
    free(p);
    return cudaSuccess;
}
static inline cudaError_t cudaFreeHost(void* p) {
    free(p);
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) {
    memcpy(d, s, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) {
    memcpy(d, s, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) {
    memset(d, v, n);
    return cudaSuccess;
}
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes {
    cudaMemoryType type;
    int device;
    void* devicePointer;
    void* hostPointer;
};
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) {
    // OSMR_EMU_PINNED=1: treat every host pointer as page-locked and mapped (exercises the direct-output path)
    const bool pinned = getenv("OSMR_EMU_PINNED") != nullptr;
    a->type = pinned ? cudaMemoryTypeHost : cudaMemoryTypeUnregistered;
    a->device = 0;
    a->devicePointer = pinned ? const_cast<void*>(p) : nullptr;
    a->hostPointer = const_cast<void*>(p);
    return cudaSuccess;
}
template <typename T>
static inline cudaError_t cudaMemcpyToSymbol(T& sym, const void* src, size_t n) {
    memcpy((void*)&sym, src, n);
    return cudaSuccess;
}
