// ffb_common.h -- shared host/device declarations for the sm_100a kernels.
#pragma once

#ifdef FFB_EMU
#include "cuda_emu.h"   // tests/emu: host emulation used only by the CPU test-suite
#define FFB_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emu::launch((grid), (block), (smem), [=]() { kernel(__VA_ARGS__); })
#define FFB_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::g_dyn_smem)
#else
#include <cuda_runtime.h>
#define FFB_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define FFB_DYN_SMEM(type, name) \
    extern __shared__ __align__(16) unsigned char ffb_dyn_smem_raw[]; \
    type* name = reinterpret_cast<type*>(ffb_dyn_smem_raw)
#endif

#include <stddef.h>
#include <stdint.h>

#define FFB_MAX_LEVELS 4
#define FFB_POLY_N 5
#define FFB_WIN 15
#define FFB_WIN_R 7
#define FFB_ITERS 3

// Half of a symmetric 1-D kernel: k[0] is the centre tap, k[i] the taps at +-i.
struct FfbTaps {
    int r;
    float k[10];
};

// Polynomial-expansion constants (FarnebackPrepareGaussian, computed on the host in double).
struct FfbPolyConsts {
    float g[FFB_POLY_N + 1];
    float xg[FFB_POLY_N + 1];
    float xxg[FFB_POLY_N + 1];
    float ig11, ig03, ig33, ig55;
};

// A ring of equally sized device buffers: element j lives at base + ((first + j) % mod) * stride.
struct FfbRing {
    char* base;
    size_t stride;   // bytes
    int first;
    int mod;
};
__host__ __device__ __forceinline__ char* ffb_ring_at(const FfbRing& r, int j) {
    return r.base + (size_t)((r.first + j) % r.mod) * r.stride;
}

static __host__ __device__ __forceinline__ int ffb_round_up(int v, int m) { return (v + m - 1) / m * m; }
