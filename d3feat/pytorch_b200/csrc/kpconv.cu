// KPConv forward / backward (replaces the ATen chain of models/blocks.py:237-382): C ABI, workspace layout, the wiring
// of gather kernel -> contraction GEMM, and the first-generation (v1) gather kernels kept as a selectable baseline.
//
// Math per query i:
//   w[i,k,h]  = influence(|| (s[idx[i,h]] - q[i]) - kp[k] ||^2)           (shadow idx -> no contribution)
//   wf[i,k,:] = m[i,k] * sum_h w[i,k,h] * x[idx[i,h], :]
//   out[i,:]  = act((sum_k wf[i,k,:] @ W[k]) / max(1, #{h : sum_c x[idx[i,h],c] > 0}) + bias)
//
// Kernels
//   kp_rowpos     : per support row, (sum_c x[j,c] > 0)                      (density count, blocks.py:377)
//   gather        : kp2_correlate (kpconv2.cu, default) or v1 kp_correlate (this file): wf [Nq, K*Cin] row-major = the A
//                   operand of the contraction
//   contraction   : d3f_gemm_launch (gemm.cu / gemm_tcgen05.cu), 3xTF32 on tcgen05 with fused epilogue:
//                   out = act(diag(inv_n) wf W + bias),  dW = wf^T diag(inv_n) g,  dwf = diag(inv_n) g W^T
//   backward data : kp2t_correlate over transposed neighbour lists + GEMM with W^T (atomic-free), or dwf GEMM +
//                   kp2_scatter / v1 kp_scatter (reductions)
#include "common.cuh"
#include "gemm.cuh"
#include "kpconv.cuh"
#include <stdlib.h>

namespace {

constexpr int KP = 16;            // kernel points padded to 16 in shared memory
constexpr float SHADOW = 1e6f;    // blocks.py:277

__global__ void kp_rowpos_kernel(const float* __restrict__ x, int ns, int cin, unsigned char* __restrict__ rowpos) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= ns) return;
    float s = 0.f;
    for (int c = lane; c < cin; c += 32) s += x[(size_t)warp * cin + c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) rowpos[warp] = s > 0.0f;
}

__device__ __forceinline__ float kp_influence(float sq, float extent, int influence) {
    if (influence == D3F_INFLUENCE_LINEAR) return fmaxf(1.0f - sqrtf(sq) / extent, 0.0f);
    if (influence == D3F_INFLUENCE_GAUSSIAN) {
        const float sigma = extent * 0.3f;
        return expf(-sq / (2.0f * sigma * sigma + 1e-9f));
    }
    return 1.0f;
}

// d influence / d sq  (for the kernel-point gradient of deformable layers)
__device__ __forceinline__ float kp_influence_grad(float sq, float w, float extent, int influence) {
    if (influence == D3F_INFLUENCE_LINEAR) {
        if (!(1.0f - sqrtf(sq) / extent >= 0.0f)) return 0.0f;
        return -1.0f / (2.0f * extent * sqrtf(sq));
    }
    if (influence == D3F_INFLUENCE_GAUSSIAN) {
        const float sigma = extent * 0.3f;
        return -w / (2.0f * sigma * sigma + 1e-9f);
    }
    return 0.0f;
}

struct KpArgs {
    const float* q; const float* s; const void* inds; long long ld; const float* x;
    const float* kp; const float* mod; const unsigned char* rowpos;
    int nq, ns, H, K, cin; float extent; int influence, aggregation;
};

// Phase 1 for one query (whole warp).  Fills w_s[h][KP] (0 for dropped / shadow neighbours),
// idx_s[h] (-1 = skip), optionally rel_s[h][3]; returns the density count.
template <bool IDX64, bool DEFORMED>
__device__ __forceinline__ int kp_phase1(const KpArgs& a, int qi, int lane, const float* kp_s /*[KP*3]*/,
                                         float* w_s, int* idx_s, float* rel_s, float* mind2 /*[KP] per lane or null*/) {
    const float qx = a.q[3 * (size_t)qi], qy = a.q[3 * (size_t)qi + 1], qz = a.q[3 * (size_t)qi + 2];
    const float ext2 = a.extent * a.extent;
    int count = 0;
    for (int h0 = 0; h0 < a.H; h0 += 32) {
        const int h = h0 + lane;
        bool pos = false;
        if (h < a.H) {
            long long idx = IDX64 ? ((const long long*)a.inds)[(size_t)qi * a.ld + h]
                                  : (long long)((const int*)a.inds)[(size_t)qi * a.ld + h];
            const bool valid = idx >= 0 && idx < a.ns;
            float rx, ry, rz;
            if (valid) {
                rx = a.s[3 * idx] - qx; ry = a.s[3 * idx + 1] - qy; rz = a.s[3 * idx + 2] - qz;
            } else {
                rx = SHADOW - qx; ry = SHADOW - qy; rz = SHADOW - qz;
            }
            float wv[KP];
            bool in_range = false;
            float best = INFINITY; int best_k = 0;
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                wv[k] = 0.f;
                if (k < a.K) {
                    const float dx = rx - kp_s[3 * k], dy = ry - kp_s[3 * k + 1], dz = rz - kp_s[3 * k + 2];
                    const float sq = dx * dx + dy * dy + dz * dz;
                    if (DEFORMED) {
                        in_range |= sq < ext2;
                        if (mind2) mind2[k] = fminf(mind2[k], sq);
                    }
                    if (sq < best) { best = sq; best_k = k; }
                    wv[k] = kp_influence(sq, a.extent, a.influence);
                }
            }
            if (a.aggregation == D3F_AGGREGATION_CLOSEST) {
#pragma unroll
                for (int k = 0; k < KP; ++k) if (k != best_k) wv[k] = 0.f;
            }
            const bool keep = valid && (!DEFORMED || in_range);
            pos = keep && a.rowpos[idx];
            idx_s[h] = keep ? (int)idx : -1;
#pragma unroll
            for (int k = 0; k < KP; ++k) w_s[h * KP + k] = keep ? wv[k] : 0.f;
            if (rel_s) { rel_s[3 * h] = rx; rel_s[3 * h + 1] = ry; rel_s[3 * h + 2] = rz; }
        }
        count += __popc(__ballot_sync(0xffffffffu, pos));
    }
    __syncwarp();
    return count;
}

// per-warp shared memory: w_s[H*KP] | idx_s[H] | kp_s[KP*3] | rel_s[H*3] (backward only)
__host__ __device__ inline size_t kp_warp_smem_floats(int H, bool with_rel) {
    size_t f = (size_t)H * KP + H + KP * 3 + (with_rel ? (size_t)H * 3 : 0);
    return (f + 3) & ~(size_t)3;
}

template <bool IDX64, bool DEFORMED, int CG>
__global__ void __launch_bounds__(256)
kp_correlate_kernel(KpArgs a, float* __restrict__ wf, float* __restrict__ wf_unmod, float* __restrict__ inv_n,
                    float* __restrict__ min_d2) {
    extern __shared__ float4 smem_f4[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qi = blockIdx.x * (blockDim.x >> 5) + warp;
    float* base = (float*)smem_f4 + (size_t)warp * kp_warp_smem_floats(a.H, false);
    float* w_s = base;
    int* idx_s = (int*)(w_s + (size_t)a.H * KP);
    float* kp_s = (float*)(idx_s + a.H);
    if (qi >= a.nq) return;
    for (int t = lane; t < KP * 3; t += 32) {
        const int k = t / 3;
        kp_s[t] = k < a.K ? (DEFORMED ? a.kp[(size_t)qi * a.K * 3 + t] : a.kp[t]) : 0.f;
    }
    __syncwarp();
    float mind2[KP];
    if (DEFORMED) {
#pragma unroll
        for (int k = 0; k < KP; ++k) mind2[k] = INFINITY;
    }
    const int count = kp_phase1<IDX64, DEFORMED>(a, qi, lane, kp_s, w_s, idx_s, nullptr,
                                                 (DEFORMED && min_d2) ? mind2 : nullptr);
    if (lane == 0) inv_n[qi] = 1.0f / (float)max(count, 1);
    if (DEFORMED && min_d2) {
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            float v = mind2[k];
            for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
            if (lane == 0 && k < a.K) min_d2[(size_t)qi * a.K + k] = v;
        }
    }
    const float4* w4 = (const float4*)w_s;
    for (int c0 = 0; c0 < a.cin; c0 += 32 * CG) {
        float acc[CG][KP];
#pragma unroll
        for (int j = 0; j < CG; ++j)
#pragma unroll
            for (int k = 0; k < KP; ++k) acc[j][k] = 0.f;
        for (int h = 0; h < a.H; ++h) {
            const int idx = idx_s[h];
            if (idx < 0) continue;
            float xv[CG];
#pragma unroll
            for (int j = 0; j < CG; ++j) {
                const int c = c0 + j * 32 + lane;
                xv[j] = c < a.cin ? __ldg(&a.x[(size_t)idx * a.cin + c]) : 0.f;
            }
            float w[KP];
#pragma unroll
            for (int v = 0; v < KP / 4; ++v) {
                const float4 t = w4[h * (KP / 4) + v];
                w[4 * v] = t.x; w[4 * v + 1] = t.y; w[4 * v + 2] = t.z; w[4 * v + 3] = t.w;
            }
#pragma unroll
            for (int j = 0; j < CG; ++j)
#pragma unroll
                for (int k = 0; k < KP; ++k) acc[j][k] = fmaf(w[k], xv[j], acc[j][k]);
        }
#pragma unroll
        for (int j = 0; j < CG; ++j) {
            const int c = c0 + j * 32 + lane;
            if (c < a.cin) {
#pragma unroll
                for (int k = 0; k < KP; ++k)
                    if (k < a.K) {
                        float v = acc[j][k];
                        const size_t o = ((size_t)qi * a.K + k) * a.cin + c;
                        if (a.mod) {
                            if (wf_unmod) wf_unmod[o] = v;
                            v *= a.mod[(size_t)qi * a.K + k];
                        }
                        wf[o] = v;
                    }
            }
        }
    }
}

// --------------------------------------------------------------------------------------------
// backward scatter: dx[idx[i,h], c] += sum_k m[i,k] w[i,k,h] dwf[i,k,c]
// deformed: also dkp[i,k,:] and (modulated) dmod[i,k] = sum_c dwf[i,k,c] * wf_unmod[i,k,c]
template <bool IDX64, bool DEFORMED, int CG>
__global__ void __launch_bounds__(256)
kp_scatter_kernel(KpArgs a, const float* __restrict__ dwf, const float* __restrict__ wf_unmod,
                  float* __restrict__ grad_x, float* __restrict__ grad_kp, float* __restrict__ grad_mod) {
    extern __shared__ float4 smem_f4[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qi = blockIdx.x * (blockDim.x >> 5) + warp;
    float* base = (float*)smem_f4 + (size_t)warp * kp_warp_smem_floats(a.H, true);
    float* w_s = base;
    int* idx_s = (int*)(w_s + (size_t)a.H * KP);
    float* kp_s = (float*)(idx_s + a.H);
    float* rel_s = kp_s + KP * 3;
    if (qi >= a.nq) return;
    for (int t = lane; t < KP * 3; t += 32) {
        const int k = t / 3;
        kp_s[t] = k < a.K ? (DEFORMED ? a.kp[(size_t)qi * a.K * 3 + t] : a.kp[t]) : 0.f;
    }
    __syncwarp();
    kp_phase1<IDX64, DEFORMED>(a, qi, lane, kp_s, w_s, idx_s, rel_s, nullptr);
    const float4* w4 = (const float4*)w_s;

    float mk[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) mk[k] = (a.mod && k < a.K) ? a.mod[(size_t)qi * a.K + k] : 1.0f;

    // t_hk accumulators for the kernel-point gradient: dkp[k] over this lane's channel slice
    float gkp[KP][3];
    float gmod[KP];
    if (DEFORMED) {
#pragma unroll
        for (int k = 0; k < KP; ++k) { gkp[k][0] = gkp[k][1] = gkp[k][2] = 0.f; gmod[k] = 0.f; }
    }

    for (int c0 = 0; c0 < a.cin; c0 += 32 * CG) {
        float d[CG][KP];
#pragma unroll
        for (int j = 0; j < CG; ++j) {
            const int c = c0 + j * 32 + lane;
#pragma unroll
            for (int k = 0; k < KP; ++k)
                d[j][k] = (k < a.K && c < a.cin) ? dwf[((size_t)qi * a.K + k) * a.cin + c] : 0.f;
        }
        if (DEFORMED && grad_mod && wf_unmod) {
#pragma unroll
            for (int j = 0; j < CG; ++j) {
                const int c = c0 + j * 32 + lane;
                if (c < a.cin)
#pragma unroll
                    for (int k = 0; k < KP; ++k)
                        if (k < a.K) gmod[k] = fmaf(d[j][k], wf_unmod[((size_t)qi * a.K + k) * a.cin + c], gmod[k]);
            }
        }
        for (int h = 0; h < a.H; ++h) {
            const int idx = idx_s[h];
            if (idx < 0) continue;
            float w[KP];
#pragma unroll
            for (int v = 0; v < KP / 4; ++v) {
                const float4 t = w4[h * (KP / 4) + v];
                w[4 * v] = t.x; w[4 * v + 1] = t.y; w[4 * v + 2] = t.z; w[4 * v + 3] = t.w;
            }
            float xv[CG];
            if (DEFORMED && grad_kp) {
#pragma unroll
                for (int j = 0; j < CG; ++j) {
                    const int c = c0 + j * 32 + lane;
                    xv[j] = c < a.cin ? __ldg(&a.x[(size_t)idx * a.cin + c]) : 0.f;
                }
            }
            if (grad_x) {
#pragma unroll
                for (int j = 0; j < CG; ++j) {
                    const int c = c0 + j * 32 + lane;
                    float v = 0.f;
#pragma unroll
                    for (int k = 0; k < KP; ++k) v = fmaf(w[k] * mk[k], d[j][k], v);
                    if (c < a.cin) atomicAdd(&grad_x[(size_t)idx * a.cin + c], v);
                }
            }
            if (DEFORMED && grad_kp) {
                // dw[k] = m[k] * <dwf[k,:], x[idx,:]>   (this lane's channel slice; reduced over lanes at the end
                // because d sq/d kp is lane-independent)
                const float rx = rel_s[3 * h], ry = rel_s[3 * h + 1], rz = rel_s[3 * h + 2];
                float best = INFINITY; int best_k = 0;
                if (a.aggregation == D3F_AGGREGATION_CLOSEST) {
#pragma unroll
                    for (int k = 0; k < KP; ++k)
                        if (k < a.K) {
                            const float dx = rx - kp_s[3 * k], dy = ry - kp_s[3 * k + 1], dz = rz - kp_s[3 * k + 2];
                            const float sq = dx * dx + dy * dy + dz * dz;
                            if (sq < best) { best = sq; best_k = k; }
                        }
                }
#pragma unroll
                for (int k = 0; k < KP; ++k)
                    if (k < a.K) {
                        float t = 0.f;
#pragma unroll
                        for (int j = 0; j < CG; ++j) t = fmaf(d[j][k], xv[j], t);
                        const float dx = rx - kp_s[3 * k], dy = ry - kp_s[3 * k + 1], dz = rz - kp_s[3 * k + 2];
                        const float sq = dx * dx + dy * dy + dz * dz;
                        float gw = kp_influence_grad(sq, w[k], a.extent, a.influence);
                        if (a.aggregation == D3F_AGGREGATION_CLOSEST && k != best_k) gw = 0.f;
                        const float f = t * mk[k] * gw * (-2.0f);
                        gkp[k][0] = fmaf(f, dx, gkp[k][0]);
                        gkp[k][1] = fmaf(f, dy, gkp[k][1]);
                        gkp[k][2] = fmaf(f, dz, gkp[k][2]);
                    }
            }
        }
    }
    if (DEFORMED) {
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            if (k >= a.K) continue;
            if (grad_kp) {
#pragma unroll
                for (int ax = 0; ax < 3; ++ax) {
                    float v = gkp[k][ax];
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    if (lane == 0) grad_kp[((size_t)qi * a.K + k) * 3 + ax] = v;
                }
            }
            if (grad_mod) {
                float v = gmod[k];
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) grad_mod[(size_t)qi * a.K + k] = v;
            }
        }
    }
}

int kp_warps_per_cta(int H, bool with_rel, size_t* smem) {
    const size_t per = kp_warp_smem_floats(H, with_rel) * sizeof(float);
    int warps = 8;
    while (warps > 1 && per * warps > 160 * 1024) warps >>= 1;
    *smem = per * warps;
    return warps;
}

struct KpWs { unsigned char* rowpos; float* dwf; float* det; size_t det_bytes; };
// dwf (backward) and the deterministic split-K partials (forward) share the same region
size_t kp_layout(KpWs* w, void* base, size_t cap, int nq, int ns, int K, int cin, int cout) {
    WsCursor c{(char*)base, 0, cap};
    w->rowpos = c.take<unsigned char>((size_t)(ns > 0 ? ns : 1));
    const size_t dwf_floats = (size_t)(nq > 0 ? nq : 1) * K * cin;
    w->det_bytes = d3f_gemm_det_workspace_bytes(nq, cout, K * cin);
    size_t floats = dwf_floats > w->det_bytes / sizeof(float) ? dwf_floats : w->det_bytes / sizeof(float);
    const size_t g_floats = (size_t)(ns > 0 ? ns : 1) * K * cout;   // G of the atomic-free backward (same region)
    if (g_floats > floats) floats = g_floats;
    w->dwf = c.take<float>(floats);
    w->det = w->dwf;
    return c.off;
}

int kp_check(int nq, int ns, int H, int K, int cin, int cout) {
    D3F_REQUIRE(nq >= 0 && ns >= 0 && H >= 0 && cin >= 1 && cout >= 1, D3F_ERR_INVALID, "bad sizes");
    D3F_REQUIRE(K >= 1 && K <= KP, D3F_ERR_UNSUPPORTED, "K must be in [1,16]");
    D3F_REQUIRE(kp_warp_smem_floats(H, true) * sizeof(float) <= 160 * 1024, D3F_ERR_UNSUPPORTED,
                "too many neighbour columns for one warp's shared memory");
    return D3F_OK;
}

template <typename Kern>
int kp_set_smem(Kern kern, size_t smem) {
    if (smem > 48 * 1024)
        D3F_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return D3F_OK;
}

#define KP_DISPATCH(NAME, IDX64, DEF, CG, ...)                                            \
    do {                                                                                  \
        auto kern = NAME<IDX64, DEF, CG>;                                                 \
        int rc_ = kp_set_smem(kern, smem);                                                \
        if (rc_) return rc_;                                                              \
        kern<<<grid, warps * 32, smem, stream>>>(__VA_ARGS__);                            \
    } while (0)

#define KP_DISPATCH_ALL(NAME, ...)                                                        \
    do {                                                                                  \
        const int cg = cin <= 32 ? 1 : (cin <= 64 ? 2 : 4);                               \
        if (idx_is_64) {                                                                  \
            if (deformed) { if (cg == 1) KP_DISPATCH(NAME, true, true, 1, __VA_ARGS__);   \
                            else if (cg == 2) KP_DISPATCH(NAME, true, true, 2, __VA_ARGS__); \
                            else KP_DISPATCH(NAME, true, true, 4, __VA_ARGS__); }         \
            else { if (cg == 1) KP_DISPATCH(NAME, true, false, 1, __VA_ARGS__);           \
                   else if (cg == 2) KP_DISPATCH(NAME, true, false, 2, __VA_ARGS__);      \
                   else KP_DISPATCH(NAME, true, false, 4, __VA_ARGS__); }                 \
        } else {                                                                          \
            if (deformed) { if (cg == 1) KP_DISPATCH(NAME, false, true, 1, __VA_ARGS__);  \
                            else if (cg == 2) KP_DISPATCH(NAME, false, true, 2, __VA_ARGS__); \
                            else KP_DISPATCH(NAME, false, true, 4, __VA_ARGS__); }        \
            else { if (cg == 1) KP_DISPATCH(NAME, false, false, 1, __VA_ARGS__);          \
                   else if (cg == 2) KP_DISPATCH(NAME, false, false, 2, __VA_ARGS__);     \
                   else KP_DISPATCH(NAME, false, false, 4, __VA_ARGS__); }                \
        }                                                                                 \
    } while (0)

}  // namespace

// Gather-kernel generation: 0 = v1 (this file), 1 = v2 with FFMA accumulation, 2 = v2 with mma.sync 3xTF32
// accumulation (kpconv2.cu).  Default from D3F_KPCONV_IMPL = v1 | ffma | mma, else v2/MMA.
static int g_kp_impl = -1;
extern "C" void d3f_set_kpconv_impl(int impl) { g_kp_impl = impl < 0 ? -1 : (impl > 2 ? 2 : impl); }
static int kp_impl() {
    if (g_kp_impl < 0) {
        const char* e = getenv("D3F_KPCONV_IMPL");
        g_kp_impl = !e ? 2 : (e[0] == 'v' ? 0 : (e[0] == 'f' ? 1 : 2));
    }
    return g_kp_impl;
}
extern "C" int d3f_get_kpconv_impl(void) { return kp_impl(); }

// Optional CUDA events recorded immediately before / after the forward gather kernel (kp_correlate / kp2_correlate) on
// the caller's stream, so a benchmark can time that kernel alone inside d3f_kpconv_forward.  NULL switches it off.
static cudaEvent_t g_kp_ev0 = nullptr, g_kp_ev1 = nullptr;
extern "C" void d3f_kpconv_set_gather_events(void* start_event, void* stop_event) {
    g_kp_ev0 = (cudaEvent_t)start_event;
    g_kp_ev1 = (cudaEvent_t)stop_event;
}

extern "C" size_t d3f_kpconv_workspace_bytes(int n_queries, int n_supports, int n_neighbors, int K, int c_in,
                                             int c_out) {
    (void)n_neighbors;
    KpWs w;
    return kp_layout(&w, nullptr, 0, n_queries, n_supports, K, c_in, c_out);
}

extern "C" int d3f_kpconv_forward_ex(const float* q_pts, const float* s_pts, const void* inds, int idx_is_64,
                                     int64_t ld_inds, const float* x, const float* weights,
                                     const float* kernel_points, int deformed, const float* modulations,
                                     int nq, int ns, int H, int K, int cin, int cout, float kp_extent, int influence,
                                     int aggregation, const float* bias, int leaky_relu, float slope,
                                     float* out, float* wf, float* wf_unmod, float* inv_n,
                                     float* min_d2, void* workspace, size_t workspace_bytes, d3f_stream stream_);

extern "C" int d3f_kpconv_forward(const float* q_pts, const float* s_pts, const void* inds, int idx_is_64,
                                  int64_t ld_inds, const float* x, const float* weights,
                                  const float* kernel_points, int deformed, const float* modulations,
                                  int nq, int ns, int H, int K, int cin, int cout, float kp_extent, int influence,
                                  int aggregation, float* out, float* wf, float* wf_unmod, float* inv_n,
                                  float* min_d2, void* workspace, size_t workspace_bytes, d3f_stream stream_) {
    return d3f_kpconv_forward_ex(q_pts, s_pts, inds, idx_is_64, ld_inds, x, weights, kernel_points, deformed, modulations,
                                 nq, ns, H, K, cin, cout, kp_extent, influence, aggregation, nullptr, 0, 0.f,
                                 out, wf, wf_unmod, inv_n, min_d2, workspace, workspace_bytes, stream_);
}

// out = act(diag(1/n) wf W + bias): the learned bias that replaces batch norm and the LeakyReLU of SimpleBlock /
// ResnetBottleneckBlock (blocks.py:597, :671) ride in the contraction's epilogue.
extern "C" int d3f_kpconv_forward_ex(const float* q_pts, const float* s_pts, const void* inds, int idx_is_64,
                                     int64_t ld_inds, const float* x, const float* weights,
                                     const float* kernel_points, int deformed, const float* modulations,
                                     int nq, int ns, int H, int K, int cin, int cout, float kp_extent, int influence,
                                     int aggregation, const float* bias, int leaky_relu, float slope,
                                     float* out, float* wf, float* wf_unmod, float* inv_n,
                                     float* min_d2, void* workspace, size_t workspace_bytes, d3f_stream stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = kp_check(nq, ns, H, K, cin, cout);
    if (rc) return rc;
    if (nq == 0) return D3F_OK;
    D3F_REQUIRE(q_pts && s_pts && (inds || H == 0) && x && weights && kernel_points && out && wf && inv_n,
                D3F_ERR_INVALID, "null pointer");
    D3F_REQUIRE(!modulations || wf_unmod, D3F_ERR_INVALID, "wf_unmod is required with modulations");
    D3F_REQUIRE(influence >= 0 && influence <= 2 && aggregation >= 0 && aggregation <= 1, D3F_ERR_INVALID, "bad mode");
    KpWs w;
    const size_t need = kp_layout(&w, workspace, workspace_bytes, nq, ns, K, cin, cout);
    D3F_REQUIRE(workspace && need <= workspace_bytes, D3F_ERR_WORKSPACE, "workspace too small");
    if (ns > 0) {
        kp_rowpos_kernel<<<d3f_ceil_div(ns, 8), 256, 0, stream>>>(x, ns, cin, w.rowpos);
        D3F_CHECK_LAUNCH();
    }
    KpArgs a{q_pts, s_pts, inds, (long long)ld_inds, x, kernel_points, modulations, w.rowpos,
             nq, ns, H, K, cin, kp_extent, influence, aggregation};
    if (g_kp_ev0) D3F_CHECK_CUDA(cudaEventRecord(g_kp_ev0, stream));
    if (kp_impl() >= 1 && kp2_supported(H, ns, cin)) {
        Kp2Args a2{q_pts, s_pts, inds, (long long)ld_inds, x, kernel_points, modulations, w.rowpos,
                   nq, ns, H, K, cin, kp_extent, influence, aggregation, idx_is_64 ? 1 : 0, deformed ? 1 : 0};
        rc = kp2_correlate_launch(a2, wf, wf_unmod, inv_n, deformed ? min_d2 : nullptr, kp_impl() == 2 ? 1 : 0, stream);
        if (rc) return rc;
    } else {
        size_t smem;
        const int warps = kp_warps_per_cta(H, false, &smem);
        const int grid = d3f_ceil_div(nq, warps);
        KP_DISPATCH_ALL(kp_correlate_kernel, a, wf, wf_unmod, inv_n, deformed ? min_d2 : nullptr);
        D3F_CHECK_LAUNCH();
    }
    if (g_kp_ev1) D3F_CHECK_CUDA(cudaEventRecord(g_kp_ev1, stream));
    D3fGemm g{nq, cout, K * cin, wf, K * cin, weights, cout, out, cout, inv_n, nullptr, bias, leaky_relu, slope, 0, nullptr};
    // forward: deterministic, padding-independent split-K (a 1e-7 perturbation here can flip a LeakyReLU mask)
    float dummy;
    return d3f_gemm_launch(g, false, false, stream, w.det_bytes ? w.det : &dummy, w.det_bytes);
}

extern "C" int d3f_kpconv_backward_ex(const float* q_pts, const float* s_pts, const void* inds, int idx_is_64,
                                      int64_t ld_inds, const float* x, const float* weights,
                                      const float* kernel_points, int deformed, const float* modulations,
                                      int nq, int ns, int H, int K, int cin, int cout, float kp_extent,
                                      int influence, int aggregation, const float* wf, const float* wf_unmod,
                                      const float* inv_n, const float* grad_out, float* grad_x,
                                      float* grad_weights, float* grad_kernel_points, float* grad_modulations,
                                      const int32_t* t_offsets, const int32_t* t_src,
                                      void* workspace, size_t workspace_bytes, d3f_stream stream_);

extern "C" int d3f_kpconv_backward(const float* q_pts, const float* s_pts, const void* inds, int idx_is_64,
                                   int64_t ld_inds, const float* x, const float* weights,
                                   const float* kernel_points, int deformed, const float* modulations,
                                   int nq, int ns, int H, int K, int cin, int cout, float kp_extent,
                                   int influence, int aggregation, const float* wf, const float* wf_unmod,
                                   const float* inv_n, const float* grad_out, float* grad_x,
                                   float* grad_weights, float* grad_kernel_points, float* grad_modulations,
                                   void* workspace, size_t workspace_bytes, d3f_stream stream_) {
    return d3f_kpconv_backward_ex(q_pts, s_pts, inds, idx_is_64, ld_inds, x, weights, kernel_points, deformed, modulations,
                                  nq, ns, H, K, cin, cout, kp_extent, influence, aggregation, wf, wf_unmod, inv_n, grad_out,
                                  grad_x, grad_weights, grad_kernel_points, grad_modulations, nullptr, nullptr,
                                  workspace, workspace_bytes, stream_);
}

// With the transposed neighbour lists of d3f_neighbors_transpose (t_offsets [Ns+1], t_src), rigid layers whose Cout is
// a multiple of 32 compute grad_x WITHOUT atomics: G = gather over the lists (kp2t_correlate), grad_x = G x W^T.
extern "C" int d3f_kpconv_backward_ex(const float* q_pts, const float* s_pts, const void* inds, int idx_is_64,
                                      int64_t ld_inds, const float* x, const float* weights,
                                      const float* kernel_points, int deformed, const float* modulations,
                                      int nq, int ns, int H, int K, int cin, int cout, float kp_extent,
                                      int influence, int aggregation, const float* wf, const float* wf_unmod,
                                      const float* inv_n, const float* grad_out, float* grad_x,
                                      float* grad_weights, float* grad_kernel_points, float* grad_modulations,
                                      const int32_t* t_offsets, const int32_t* t_src,
                                      void* workspace, size_t workspace_bytes, d3f_stream stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = kp_check(nq, ns, H, K, cin, cout);
    if (rc) return rc;
    D3F_REQUIRE(influence >= 0 && influence <= 2 && aggregation >= 0 && aggregation <= 1, D3F_ERR_INVALID, "bad mode");
    const bool transposed = grad_x && t_offsets && t_src && !deformed && !modulations && kp_impl() >= 1 &&
                            kp2t_supported(nq, cout) && ns > 0 && nq > 0;
    // the scatter accumulates into grad_x with reductions; the transposed path's GEMM overwrites it
    if (grad_x && ns > 0 && !transposed)
        D3F_CHECK_CUDA(cudaMemsetAsync(grad_x, 0, sizeof(float) * (size_t)ns * cin, stream));
    if (nq == 0) {
        if (grad_weights) D3F_CHECK_CUDA(cudaMemsetAsync(grad_weights, 0, sizeof(float) * (size_t)K * cin * cout, stream));
        return D3F_OK;
    }
    D3F_REQUIRE(q_pts && s_pts && (inds || H == 0) && x && weights && kernel_points && wf && inv_n && grad_out,
                D3F_ERR_INVALID, "null pointer");
    KpWs w;
    const size_t need = kp_layout(&w, workspace, workspace_bytes, nq, ns, K, cin, cout);
    D3F_REQUIRE(workspace && need <= workspace_bytes, D3F_ERR_WORKSPACE, "workspace too small");
    const int KC = K * cin;
    // dW[kc, o] = sum_i wf[i, kc] * inv_n[i] * g[i, o]
    if (grad_weights) {
        D3fGemm g{KC, cout, nq, wf, KC, grad_out, cout, grad_weights, cout, nullptr, inv_n, nullptr, 0, 0.f, 0, nullptr};
        rc = d3f_gemm_launch(g, true, false, stream);
        if (rc) return rc;
    }
    const bool need_scatter = grad_x || (deformed && (grad_kernel_points || grad_modulations));
    if (!need_scatter) return D3F_OK;
    if (transposed) {
        // atomic-free: G[j,k,o] over the transposed lists, then grad_x[j,c] = sum_{k,o} G[j,k,o] W[k,c,o]
        float* G = w.dwf;
        Kp2tArgs ta{q_pts, s_pts, t_offsets, t_src, grad_out, inv_n, kernel_points, nq, ns, K, cout, kp_extent, influence,
                    aggregation};
        rc = kp2t_correlate_launch(ta, G, stream);
        if (rc) return rc;
        D3fGemm g{ns, cin, K * cout, G, K * cout, weights, cout, grad_x, cin, nullptr, nullptr, nullptr, 0, 0.f, 0, nullptr,
                  nullptr, nullptr, 0, cout, (long long)cin * cout};
        return d3f_gemm_launch(g, false, true, stream);
    }
    // dwf[i, kc] = inv_n[i] * sum_o g[i, o] * W[kc, o]
    {
        D3fGemm g{nq, KC, cout, grad_out, cout, weights, cout, w.dwf, KC, inv_n, nullptr, nullptr, 0, 0.f, 0, nullptr};
        rc = d3f_gemm_launch(g, false, true, stream);
    }
    if (rc) return rc;
    if (ns > 0) {
        kp_rowpos_kernel<<<d3f_ceil_div(ns, 8), 256, 0, stream>>>(x, ns, cin, w.rowpos);
        D3F_CHECK_LAUNCH();
    }
    KpArgs a{q_pts, s_pts, inds, (long long)ld_inds, x, kernel_points, modulations, w.rowpos,
             nq, ns, H, K, cin, kp_extent, influence, aggregation};
    if (kp_impl() >= 1 && kp2_supported(H, ns, cin)) {
        Kp2Args a2{q_pts, s_pts, inds, (long long)ld_inds, x, kernel_points, modulations, w.rowpos,
                   nq, ns, H, K, cin, kp_extent, influence, aggregation, idx_is_64 ? 1 : 0, deformed ? 1 : 0};
        return kp2_scatter_launch(a2, w.dwf, wf_unmod, grad_x, deformed ? grad_kernel_points : nullptr,
                                  deformed ? grad_modulations : nullptr, stream);
    }
    size_t smem;
    const int warps = kp_warps_per_cta(H, true, &smem);
    const int grid = d3f_ceil_div(nq, warps);
    KP_DISPATCH_ALL(kp_scatter_kernel, a, w.dwf, wf_unmod, grad_x, deformed ? grad_kernel_points : nullptr,
                    deformed ? grad_modulations : nullptr);
    D3F_CHECK_LAUNCH();
    return D3F_OK;
}
