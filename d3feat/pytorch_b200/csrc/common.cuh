// Shared helpers for libd3feat_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../../include/d3feat_b200.h"

void d3f_set_error(const char* fmt, ...);
int* d3f_fail_flag_device();   // gemm_tcgen05.cu: device int raised when a tensor-core kernel gives up on an mbarrier

#define D3F_CHECK_CUDA(expr)                                                                   \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            d3f_set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,                 \
                          cudaGetErrorString(_e));                                             \
            return D3F_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

extern unsigned long long g_d3f_launches;   // kernels launched by this library (diagnostic; not thread-safe)
#define D3F_CHECK_LAUNCH()                                                                     \
    do {                                                                                       \
        ++g_d3f_launches;                                                                      \
        D3F_CHECK_CUDA(cudaGetLastError());                                                    \
    } while (0)

#define D3F_REQUIRE(cond, code, msg)                                                           \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            d3f_set_error("%s (%s) at %s:%d", msg, #cond, __FILE__, __LINE__);                 \
            return code;                                                                       \
        }                                                                                      \
    } while (0)

static inline size_t d3f_align(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// bump allocator over the caller's workspace
struct WsCursor {
    char* base; size_t off; size_t cap;
    template <typename T> T* take(size_t n) {
        T* p = (T*)(base + off);
        off += d3f_align(n * sizeof(T));
        return p;
    }
    bool ok() const { return off <= cap; }
};

__host__ __device__ static inline int d3f_ceil_div(int a, int b) { return (a + b - 1) / b; }

// fill.cu: up to 6 regions of 32-bit words, each set to its own pattern, in ONE kernel launch
struct D3fFillSegs {
    void* p[6];
    size_t words[6];
    uint32_t v[6];
    int n;
    D3fFillSegs() : n(0) {}
    void add(void* ptr, size_t bytes, uint32_t pattern) {
        if (bytes == 0) return;
        p[n] = ptr; words[n] = bytes / 4; v[n] = pattern; ++n;
    }
};
int d3f_fill_segments(const D3fFillSegs& s, cudaStream_t stream);
__host__ __device__ static inline uint32_t d3f_pow2ceil(uint32_t v) {
    uint32_t p = 1; while (p < v) p <<= 1; return p;
}

// 64-bit mix (splitmix64 finaliser) for the open-addressing cell tables
__device__ __forceinline__ uint32_t d3f_hash64(uint64_t k) {
    k ^= k >> 30; k *= 0xbf58476d1ce4e5b9ULL;
    k ^= k >> 27; k *= 0x94d049bb133111ebULL;
    k ^= k >> 31;
    return (uint32_t)k;
}

#define D3F_EMPTY_KEY 0xFFFFFFFFFFFFFFFFULL

// batch element of stacked row i given per-element lengths (n_batch is tiny); -1 if i >= sum(len):
// buffers may be allocated larger than the real row count (static capacities for CUDA-graph capture)
__device__ __forceinline__ int d3f_batch_of(int i, const int32_t* __restrict__ len, int nb, int* start) {
    int s = 0;
    for (int b = 0; b < nb; ++b) {
        int l = len[b];
        if (i < s + l) { *start = s; return b; }
        s += l;
    }
    *start = s;
    return -1;   // row i lies beyond the real rows (capacity padding)
}
