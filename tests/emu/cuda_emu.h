// cuda_emu.h -- TEST-ONLY host emulation of the small CUDA subset used by csrc/*.cu.
//
// Purpose: the build container has no GPU.  Compiling the *same* kernel sources with g++ against
// this shim (-DFFB_EMU) lets `pytest -m "not gpu"` execute the tiling / halo / ring-buffer /
// reduction logic of every kernel on small inputs and compare it with the oracle before any GPU
// time is spent.  It is never loaded by the product (funscript_flow_b200 only ever opens
// libffb.so, which is built by nvcc for sm_100a and needs a real device).
//
// Model: each CUDA thread of a block is a ucontext fiber; __syncthreads() and the warp shuffles
// are cooperative barriers served by a round-robin scheduler; blocks run one after another.
// A divergent barrier (not all live threads arriving) aborts -- which is a useful check in itself.
#pragma once
#ifndef FFB_EMU
#error "cuda_emu.h is only for -DFFB_EMU builds"
#endif

#include <ucontext.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <functional>
#include <vector>

// ------------------------------------------------------------------ qualifiers
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) alignas(n)
#define __constant__ static

// ------------------------------------------------------------------ vector types
struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) double2 { double x, y; };
struct uchar4 { unsigned char x, y, z, w; };
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline int2 make_int2(int a, int b) { return int2{a, b}; }
static inline double2 make_double2(double a, double b) { return double2{a, b}; }

// ------------------------------------------------------------------ runtime state
namespace emu {
struct Fiber {
    ucontext_t ctx;
    char* stack = nullptr;
    int state = 0;   // 0 runnable, 1 waiting block barrier, 2 waiting warp barrier, 3 done, 4 waiting cluster barrier
    uint3 tid{0, 0, 0};
    uint3 bid{0, 0, 0};
    int linear = 0;  // thread index inside its CTA
    int cta = 0;     // rank of its CTA inside the cluster
    char* smem = nullptr;
};
inline uint3 g_threadIdx, g_blockIdx;
inline dim3 g_blockDim, g_gridDim;
inline ucontext_t g_sched;
inline Fiber* g_cur = nullptr;
inline std::vector<Fiber> g_fibers;
inline std::function<void()> g_body;
inline char* g_dyn_smem = nullptr;       // dynamic shared memory of the running CTA
inline char* g_smem_pool = nullptr;      // one slice per CTA of the running cluster
inline size_t g_smem_pool_cap = 0, g_smem_slice = 0;
inline unsigned g_cluster_ctas = 1;
inline unsigned long long g_warp_buf[512][32];   // shuffle exchange, per (cta, warp)
inline long long g_launches = 0;
inline unsigned long long g_spin_rounds = 0;
constexpr size_t kStack = 256 * 1024;

inline void fiber_entry() {
    g_body();
    g_cur->state = 3;
    swapcontext(&g_cur->ctx, &g_sched);
}
inline void yield_state(int st) {
    Fiber* f = g_cur;
    f->state = st;
    swapcontext(&f->ctx, &g_sched);
    // resumed: scheduler restored the thread identity
}
// Runs the CTAs of one cluster (ncta >= 1) to completion; `bids` holds their block indices.
inline void run_cluster(unsigned nthreads, unsigned ncta, const uint3* bids) {
    const unsigned total = nthreads * ncta;
    if (g_fibers.size() < total) {
        size_t old = g_fibers.size();
        g_fibers.resize(total);
        for (size_t i = old; i < total; ++i) g_fibers[i].stack = (char*)malloc(kStack);
    }
    for (unsigned k = 0; k < total; ++k) {
        Fiber& f = g_fibers[k];
        const unsigned i = k % nthreads;
        f.state = 0;
        f.linear = (int)i;
        f.cta = (int)(k / nthreads);
        f.bid = bids[f.cta];
        f.smem = g_smem_pool ? g_smem_pool + (size_t)f.cta * g_smem_slice : nullptr;
        f.tid.x = i % g_blockDim.x;
        f.tid.y = (i / g_blockDim.x) % g_blockDim.y;
        f.tid.z = i / (g_blockDim.x * g_blockDim.y);
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = kStack;
        f.ctx.uc_link = &g_sched;
        makecontext(&f.ctx, (void (*)())fiber_entry, 0);
    }
    const unsigned wpc = (nthreads + 31) / 32;
    for (;;) {
        for (unsigned k = 0; k < total; ++k) {
            Fiber& f = g_fibers[k];
            if (f.state != 0) continue;
            g_cur = &f;
            g_threadIdx = f.tid;
            g_blockIdx = f.bid;
            g_dyn_smem = f.smem;
            swapcontext(&g_sched, &f.ctx);
        }
        // every fiber is now waiting, done, or still runnable (it polled a flag and yielded: ffb_spin_pause)
        unsigned done = 0, wcluster = 0, runnable = 0;
        for (unsigned k = 0; k < total; ++k) {
            done += g_fibers[k].state == 3;
            wcluster += g_fibers[k].state == 4;
            runnable += g_fibers[k].state == 0;
        }
        if (done == total) break;
        bool released = false;
        for (unsigned c = 0; c < ncta; ++c) {               // warp barriers
            for (unsigned w = 0; w < wpc; ++w) {
                unsigned lo = c * nthreads + w * 32, hi = std::min(c * nthreads + nthreads, lo + 32), live = 0, ww = 0;
                for (unsigned k = lo; k < hi; ++k) {
                    live += g_fibers[k].state != 3;
                    ww += g_fibers[k].state == 2;
                }
                if (ww && ww == live) {
                    for (unsigned k = lo; k < hi; ++k) if (g_fibers[k].state == 2) g_fibers[k].state = 0;
                    released = true;
                }
            }
        }
        if (!released) {
            for (unsigned c = 0; c < ncta; ++c) {           // block barriers, per CTA
                unsigned live = 0, wb = 0;
                for (unsigned k = c * nthreads; k < (c + 1) * nthreads; ++k) {
                    live += g_fibers[k].state != 3;
                    wb += g_fibers[k].state == 1;
                }
                if (wb && wb == live) {
                    for (unsigned k = c * nthreads; k < (c + 1) * nthreads; ++k) if (g_fibers[k].state == 1) g_fibers[k].state = 0;
                    released = true;
                }
            }
        }
        if (!released && wcluster && wcluster + done == total) {   // cluster barrier
            for (unsigned k = 0; k < total; ++k) if (g_fibers[k].state == 4) g_fibers[k].state = 0;
            released = true;
        }
        if (!released && runnable) {          // pollers wait for one another: give them more rounds, but not forever
            if (++g_spin_rounds > 50000000ull) {
                fprintf(stderr, "cuda_emu: spin-wait never satisfied in block (%u,%u,%u)\n", g_blockIdx.x, g_blockIdx.y, g_blockIdx.z);
                abort();
            }
            continue;
        }
        g_spin_rounds = 0;
        if (!released) {
            fprintf(stderr, "cuda_emu: deadlock / divergent barrier in block (%u,%u,%u)\n", g_blockIdx.x, g_blockIdx.y, g_blockIdx.z);
            abort();
        }
    }
}
// cluster = CTAs per cluster along (x, y, z); the grid must be a multiple of it
template <class F>
inline void launch(dim3 grid, dim3 block, size_t smem, F body, dim3 cluster = dim3(1, 1, 1)) {
    ++g_launches;
    const unsigned ncta = cluster.x * cluster.y * cluster.z;
    g_smem_slice = (smem + 127) / 128 * 128;
    if (g_smem_slice * ncta > g_smem_pool_cap) {
        free(g_smem_pool);
        g_smem_pool_cap = g_smem_slice * ncta;
        g_smem_pool = (char*)aligned_alloc(128, g_smem_pool_cap);
    }
    g_body = body;
    g_gridDim = grid;
    g_blockDim = block;
    g_cluster_ctas = ncta;
    const unsigned nthreads = block.x * block.y * block.z;
    if (grid.x % cluster.x || grid.y % cluster.y || grid.z % cluster.z) {
        fprintf(stderr, "cuda_emu: grid is not a multiple of the cluster shape\n");
        abort();
    }
    uint3 bids[16];
    for (unsigned cz = 0; cz < grid.z; cz += cluster.z)
        for (unsigned cy = 0; cy < grid.y; cy += cluster.y)
            for (unsigned cx = 0; cx < grid.x; cx += cluster.x) {
                unsigned n = 0;
                for (unsigned z = 0; z < cluster.z; ++z)
                    for (unsigned y = 0; y < cluster.y; ++y)
                        for (unsigned x = 0; x < cluster.x; ++x) bids[n++] = uint3{cx + x, cy + y, cz + z};
                if (smem) memset(g_smem_pool, 0xCD, g_smem_slice * ncta);   // poison: catches reads of unwritten smem
                run_cluster(nthreads, ncta, bids);
            }
}
template <class T>
inline T shfl_generic(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    int lin = g_cur->linear;
    int warp = g_cur->cta * 64 + lin / 32, lane = lin % 32;
    unsigned long long bits = 0;
    memcpy(&bits, &v, sizeof(T));
    g_warp_buf[warp][lane] = bits;
    yield_state(2);
    unsigned long long got = g_warp_buf[warp][src_lane & 31];
    yield_state(2);
    T out;
    memcpy(&out, &got, sizeof(T));
    return out;
}
// thread-block cluster support: rank, barrier, distributed-shared-memory address of a peer CTA
inline unsigned cluster_rank() { return (unsigned)g_cur->cta; }
inline void cluster_sync() { yield_state(4); }
template <class T>
inline T* map_shared(T* p, unsigned rank) {
    const size_t off = (size_t)((char*)p - g_cur->smem);
    if (off >= g_smem_slice || rank >= g_cluster_ctas) { fprintf(stderr, "cuda_emu: bad DSMEM mapping\n"); abort(); }
    return (T*)(g_smem_pool + (size_t)rank * g_smem_slice + off);
}
}  // namespace emu

#define threadIdx emu::g_threadIdx
#define blockIdx emu::g_blockIdx
#define blockDim emu::g_blockDim
#define gridDim emu::g_gridDim
static constexpr int warpSize = 32;

static inline void __syncthreads() { emu::yield_state(1); }
static inline void __threadfence_block() {}
static inline void __nanosleep(unsigned) { emu::yield_state(0); }      // a polling thread lets the others run
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::yield_state(2); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emu::shfl_generic(v, (emu::g_cur->linear % 32) ^ m); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d) {
    int lane = emu::g_cur->linear % 32;
    return emu::shfl_generic(v, lane + d < 32 ? lane + d : lane);
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, int d) {
    int lane = emu::g_cur->linear % 32;
    return emu::shfl_generic(v, lane - d >= 0 ? lane - d : lane);
}
template <class T> static inline T __shfl_sync(unsigned, T v, int l) { return emu::shfl_generic(v, l); }

// ------------------------------------------------------------------ intrinsics
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __int2float_rn(int a) { return (float)a; }
static inline int __float2int_rd(float a) { return (int)floorf(a); }
using std::max;
using std::min;
static inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; if (v > o) *p = v; return o; }
static inline int atomicAdd(int* p, int v) { int o = *p; *p += v; return o; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p += v; return o; }
static inline float atomicAdd(float* p, float v) { float o = *p; *p += v; return o; }
static inline double atomicAdd(double* p, double v) { double o = *p; *p += v; return o; }

// ------------------------------------------------------------------ runtime API subset
typedef int cudaError_t;
typedef struct emu_stream_* cudaStream_t;
typedef struct emu_event_* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void* devicePointer; void* hostPointer; };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaEventDefault = 0, cudaHostAllocDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline const char* cudaGetErrorString(cudaError_t e) { return e == 0 ? "no error" : "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
struct cudaDeviceProp { char name[256]; };
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { snprintf(p->name, sizeof(p->name), "emulated device"); return cudaSuccess; }
static inline cudaError_t cudaDeviceGetPCIBusId(char* buf, int len, int) { snprintf(buf, len, "0000:00:00.0"); return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256 + 256); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { return cudaMalloc(p, n); }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t wbytes, size_t rows, cudaMemcpyKind, cudaStream_t = nullptr) {
    for (size_t r = 0; r < rows; ++r) memcpy((char*)d + r * dp, (const char*)s + r * sp, wbytes);
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpy2D(void* d, size_t dp, const void* s, size_t sp, size_t wbytes, size_t rows, cudaMemcpyKind k) { return cudaMemcpy2DAsync(d, dp, s, sp, wbytes, rows, k); }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
// linear-memory textures: the "object" is the device pointer itself (see ffb_make_linear_texture)
typedef unsigned long long cudaTextureObject_t;
template <class T> static inline T tex1Dfetch(cudaTextureObject_t t, int i) { return reinterpret_cast<const T*>(t)[i]; }
static inline cudaError_t cudaDestroyTextureObject(cudaTextureObject_t) { return cudaSuccess; }
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void*) { a->type = cudaMemoryTypeUnregistered; a->device = 0; return cudaSuccess; }
template <class K> static inline cudaError_t cudaFuncSetAttribute(K, int, int) { return cudaSuccess; }
