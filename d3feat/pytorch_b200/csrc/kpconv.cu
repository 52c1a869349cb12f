// KPConv forward / backward (replaces the ATen chain of models/blocks.py:237-382): C ABI, workspace layout, the wiring
// of gather kernel -> contraction GEMM (the first-generation gather kernels of round 1 were deleted in round 2).
//
// Math per query i:
//   w[i,k,h]  = influence(|| (s[idx[i,h]] - q[i]) - kp[k] ||^2)           (shadow idx -> no contribution)
//   wf[i,k,:] = m[i,k] * sum_h w[i,k,h] * x[idx[i,h], :]
//   out[i,:]  = act((sum_k wf[i,k,:] @ W[k]) / max(1, #{h : sum_c x[idx[i,h],c] > 0}) + bias)
//
// Kernels
//   kp_rowpos     : per support row, (sum_c x[j,c] > 0)                      (density count, blocks.py:377)
//   gather        : kp2_correlate (kpconv2.cu): wf [Nq, K*Cin] row-major = the A operand of the contraction
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
    D3F_REQUIRE(kp2_supported(H, 1, 1), D3F_ERR_UNSUPPORTED, "too many neighbour columns for one warp's shared memory");
    return D3F_OK;
}

}  // namespace

// Forward path (debug selector, include/d3feat_b200_debug.h): 1 = gather kernel with FFMA correlation + contraction GEMM,
// 2 = gather kernel with mma.sync 3xTF32 correlation + contraction GEMM (kpconv2.cu), 3 = fused kernel (kpconv_fused.cu)
// where the layer is eligible, else 2.  Default from D3F_KPCONV_IMPL = ffma | mma | fused, else 3.
static int g_kp_impl = -1;
extern "C" void d3f_set_kpconv_impl(int impl) { g_kp_impl = impl < 0 ? -1 : (impl < 1 ? 1 : (impl > 3 ? 3 : impl)); }
static int kp_impl() {
    if (g_kp_impl < 0) {
        const char* e = getenv("D3F_KPCONV_IMPL");
        g_kp_impl = !e ? 3 : (e[0] == 'f' && e[1] == 'f' ? 1 : (e[0] == 'm' ? 2 : 3));
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
    D3F_REQUIRE(q_pts && s_pts && (inds || H == 0) && x && weights && kernel_points && out && inv_n,
                D3F_ERR_INVALID, "null pointer");
    D3F_REQUIRE(!modulations || wf_unmod, D3F_ERR_INVALID, "wf_unmod is required with modulations");
    const bool fused = kp_impl() == 3 && !deformed && !modulations && influence == D3F_INFLUENCE_LINEAR &&
                       aggregation == D3F_AGGREGATION_SUM && kpf_fused_eligible(H, K, cin, cout) && ns > 0 &&
                       (((size_t)x | (size_t)out | (size_t)wf) & 15) == 0 && (long long)ns * cin < (1LL << 31);
    D3F_REQUIRE(wf || fused, D3F_ERR_INVALID, "wf may only be NULL for layers d3f_kpconv_fused_eligible() accepts");
    D3F_REQUIRE(influence >= 0 && influence <= 2 && aggregation >= 0 && aggregation <= 1, D3F_ERR_INVALID, "bad mode");
    KpWs w;
    const size_t need = kp_layout(&w, workspace, workspace_bytes, nq, ns, K, cin, cout);
    D3F_REQUIRE(workspace && need <= workspace_bytes, D3F_ERR_WORKSPACE, "workspace too small");
    if (ns > 0) {
        kp_rowpos_kernel<<<d3f_ceil_div(ns, 8), 256, 0, stream>>>(x, ns, cin, w.rowpos);
        D3F_CHECK_LAUNCH();
    }
    if (g_kp_ev0) D3F_CHECK_CUDA(cudaEventRecord(g_kp_ev0, stream));
    if (fused) {
        // gather + correlation + contraction + 1/n + bias + LeakyReLU: one launch, wf only written on request
        Kp2Args a2{q_pts, s_pts, inds, (long long)ld_inds, x, kernel_points, nullptr, w.rowpos,
                   nq, ns, H, K, cin, kp_extent, influence, aggregation, idx_is_64 ? 1 : 0, 0};
        rc = kpf_fused_launch(a2, weights, bias, leaky_relu, slope, out, inv_n, wf, stream);
        if (rc) return rc;
        if (g_kp_ev1) D3F_CHECK_CUDA(cudaEventRecord(g_kp_ev1, stream));
        return D3F_OK;
    }
    D3F_REQUIRE((long long)ns * cin < (1LL << 31), D3F_ERR_UNSUPPORTED, "Ns * Cin must stay below 2^31");
    {
        Kp2Args a2{q_pts, s_pts, inds, (long long)ld_inds, x, kernel_points, modulations, w.rowpos,
                   nq, ns, H, K, cin, kp_extent, influence, aggregation, idx_is_64 ? 1 : 0, deformed ? 1 : 0};
        rc = kp2_correlate_launch(a2, wf, wf_unmod, inv_n, deformed ? min_d2 : nullptr, kp_impl() >= 2 ? 1 : 0, stream);
        if (rc) return rc;
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
    // (a layer whose input needs no gradient -- the first one -- and that kept wf takes the single wf^T g GEMM below: the
    // gather over the lists plus G^T x cost 195 us at the tail of the step for the 960 weights of the 1 -> 64 layer)
    const bool transposed = (grad_x || (grad_weights && !wf)) && t_offsets && t_src && !modulations &&
                            kp2t_supported(nq, cout) && ns > 0 && nq > 0 && (cout & 3) == 0;
    // the scatter accumulates into grad_x with reductions; the transposed path's GEMM overwrites it
    if (grad_x && ns > 0 && !transposed)
        D3F_CHECK_CUDA(cudaMemsetAsync(grad_x, 0, sizeof(float) * (size_t)ns * cin, stream));
    if (nq == 0) {
        if (grad_weights) D3F_CHECK_CUDA(cudaMemsetAsync(grad_weights, 0, sizeof(float) * (size_t)K * cin * cout, stream));
        return D3F_OK;
    }
    D3F_REQUIRE(q_pts && s_pts && (inds || H == 0) && x && weights && kernel_points && inv_n && grad_out,
                D3F_ERR_INVALID, "null pointer");
    KpWs w;
    const size_t need = kp_layout(&w, workspace, workspace_bytes, nq, ns, K, cin, cout);
    D3F_REQUIRE(workspace && need <= workspace_bytes, D3F_ERR_WORKSPACE, "workspace too small");
    const int KC = K * cin;
    const bool need_scatter = grad_x || (deformed && (grad_kernel_points || grad_modulations));
    float* scatter_gx = grad_x;
    if (transposed) {
        // atomic-free, and wf-free: G[j,k,o] = sum over the queries i that list support j of w * (1/n_i) * g[i,o]
        // (forward-style gather over the transposed lists), then
        //   grad_x[j,c]   = sum_{k,o} G[j,k,o] W[k,c,o]          (one GEMM, W addressed block-wise)
        //   grad_W[k,c,o] = sum_j x[j,c] G[j,k,o]               (one GEMM, C written in the [K,Cin,Cout] layout)
        // -- the weight gradient no longer needs the kernel-point-weighted features wf [Nq,K,Cin] of the forward pass,
        // so the forward (fused kernel) never writes them.
        float* G = w.dwf;
        Kp2tArgs ta{q_pts, s_pts, t_offsets, t_src, grad_out, inv_n, kernel_points, nq, ns, K, cout, kp_extent, influence,
                    aggregation, deformed ? 1 : 0};
        rc = kp2t_correlate_launch(ta, G, stream);
        if (rc) return rc;
        if (grad_weights) {
            D3fGemm g{K * cout, cin, ns, G, K * cout, x, cin, grad_weights, cout, nullptr, nullptr, nullptr, 0, 0.f, 0, nullptr,
                      nullptr, nullptr, 0, 0, 0, cout, (long long)cin * cout, 1};
            rc = d3f_gemm_launch(g, true, false, stream);
            if (rc) return rc;
        }
        if (grad_x) {
            D3fGemm g{ns, cin, K * cout, G, K * cout, weights, cout, grad_x, cin, nullptr, nullptr, nullptr, 0, 0.f, 0, nullptr,
                      nullptr, nullptr, 0, cout, (long long)cin * cout};
            rc = d3f_gemm_launch(g, false, true, stream);
            if (rc) return rc;
        }
        // deformable layer: the gradient of the deformed kernel points is query-major (every query owns its K points), so
        // it stays with the dwf GEMM + the scatter kernel below -- WITHOUT that kernel's grad_x reductions (59 M float
        // atomics onto 344 K addresses at level 3 of BASELINE config 4: 0.85 ms per layer).  G (in w.dwf) has been
        // consumed by the GEMMs queued above on this stream.
        if (!(deformed && grad_kernel_points)) return D3F_OK;
        scatter_gx = nullptr;
    } else {
        D3F_REQUIRE(wf, D3F_ERR_INVALID, "wf is required without transposed neighbour lists");
        // dW[kc, o] = sum_i wf[i, kc] * inv_n[i] * g[i, o]
        if (grad_weights) {
            D3fGemm g{KC, cout, nq, wf, KC, grad_out, cout, grad_weights, cout, nullptr, inv_n, nullptr, 0, 0.f, 0, nullptr};
            rc = d3f_gemm_launch(g, true, false, stream);
            if (rc) return rc;
        }
        if (!need_scatter) return D3F_OK;
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
    D3F_REQUIRE((long long)ns * cin < (1LL << 31), D3F_ERR_UNSUPPORTED, "Ns * Cin must stay below 2^31");
    Kp2Args a2{q_pts, s_pts, inds, (long long)ld_inds, x, kernel_points, modulations, w.rowpos,
               nq, ns, H, K, cin, kp_extent, influence, aggregation, idx_is_64 ? 1 : 0, deformed ? 1 : 0};
    return kp2_scatter_launch(a2, w.dwf, wf_unmod, scatter_gx, deformed ? grad_kernel_points : nullptr,
                              deformed ? grad_modulations : nullptr, stream);
}

// The two halves of the list-based backward as separate entry points, so that a caller can run the weight-gradient and
// the data-gradient GEMM concurrently on two streams once G is there (blocks._KPConvFunction.backward does).
extern "C" int d3f_kpconv_gather_transposed(const float* q_pts, const float* s_pts, const int32_t* t_offsets,
                                            const int32_t* t_src, const float* grad_out, const float* inv_n,
                                            const float* kernel_points, int nq, int ns, int K, int cout, float kp_extent,
                                            int influence, int aggregation, float* G, d3f_stream stream) {
    D3F_REQUIRE(nq >= 0 && ns >= 0 && K >= 1 && K <= 16 && cout >= 1, D3F_ERR_INVALID, "bad sizes");
    D3F_REQUIRE(kp2t_supported(nq, cout), D3F_ERR_UNSUPPORTED, "Cout must be a multiple of 32 and Nq * Cout < 2^31");
    D3F_REQUIRE(influence >= 0 && influence <= 2 && aggregation >= 0 && aggregation <= 1, D3F_ERR_INVALID, "bad mode");
    if (ns == 0) return D3F_OK;
    D3F_REQUIRE(q_pts && s_pts && t_offsets && t_src && grad_out && inv_n && kernel_points && G, D3F_ERR_INVALID, "null pointer");
    Kp2tArgs ta{q_pts, s_pts, t_offsets, t_src, grad_out, inv_n, kernel_points, nq, ns, K, cout, kp_extent, influence, aggregation};
    return kp2t_correlate_launch(ta, G, (cudaStream_t)stream);
}

extern "C" int d3f_kpconv_grads_from_gathered(const float* G, const float* x, const float* weights, int ns, int K, int cin,
                                              int cout, float* grad_x, float* grad_weights, int grad_weights_prezeroed,
                                              d3f_stream stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    D3F_REQUIRE(ns >= 0 && K >= 1 && cin >= 1 && cout >= 1 && (cout & 3) == 0, D3F_ERR_INVALID, "bad sizes");
    if (ns == 0) {
        if (grad_weights && !grad_weights_prezeroed)
            D3F_CHECK_CUDA(cudaMemsetAsync(grad_weights, 0, sizeof(float) * (size_t)K * cin * cout, stream));
        return D3F_OK;
    }
    D3F_REQUIRE(G && (!grad_weights || x) && (!grad_x || weights), D3F_ERR_INVALID, "null pointer");
    if (grad_weights) {     // grad_W^T = G^T x ([K*Cout, Cin]), every [Cout, Cin] block stored transposed: the [K, Cin, Cout] weight layout
        D3fGemm g{K * cout, cin, ns, G, K * cout, x, cin, grad_weights, cout, nullptr, nullptr, nullptr, 0, 0.f, 0, nullptr,
                  nullptr, nullptr, 0, 0, 0, cout, (long long)cin * cout, 1, grad_weights_prezeroed};
        const int rc = d3f_gemm_launch(g, true, false, stream);
        if (rc) return rc;
    }
    if (grad_x) {           // grad_x[j,c] = sum_{k,o} G[j,k,o] W[k,c,o]: W addressed block-wise as B^T
        D3fGemm g{ns, cin, K * cout, G, K * cout, weights, cout, grad_x, cin, nullptr, nullptr, nullptr, 0, 0.f, 0, nullptr,
                  nullptr, nullptr, 0, cout, (long long)cin * cout};
        return d3f_gemm_launch(g, false, true, stream);
    }
    return D3F_OK;
}

// 1 if d3f_kpconv_forward[_ex] runs this layer shape as ONE fused kernel (rigid, unmodulated, linear influence, sum
// aggregation; and the fused path is selected)
extern "C" int d3f_kpconv_fused_eligible(int n_neighbors, int K, int c_in, int c_out) {
    return kp_impl() == 3 && kpf_fused_eligible(n_neighbors, K, c_in, c_out) ? 1 : 0;
}
