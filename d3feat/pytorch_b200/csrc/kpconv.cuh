// Internal interface between kpconv.cu (C ABI, GEMM wiring, v1 kernels) and kpconv2.cu (v2 gather kernels).
#pragma once
#include <cuda_runtime.h>

struct Kp2Args {
    const float* q; const float* s; const void* inds; long long ld; const float* x;
    const float* kp; const float* mod; const unsigned char* rowpos;
    int nq, ns, H, K, cin; float extent; int influence, aggregation;
    int idx64, deformed;
};

// mode: 0 = FFMA accumulation, 1 = mma.sync 3xTF32 accumulation (needs cin % 4 == 0 and 16-byte aligned x / wf)
int kp2_correlate_launch(const Kp2Args& a, float* wf, float* wf_unmod, float* inv_n, float* min_d2, int mode,
                         cudaStream_t stream);
int kp2_scatter_launch(const Kp2Args& a, const float* dwf, const float* wf_unmod, float* grad_x, float* grad_kp,
                       float* grad_mod, cudaStream_t stream);
// the v2 kernels need one warp's shared-memory slab to fit and 32-bit row offsets (Ns * Cin < 2^31)
bool kp2_supported(int H, int ns, int cin);

// atomic-free backward (rigid layers): forward-style gather over the transposed neighbour lists of transpose.cu
struct Kp2tArgs {
    const float* q; const float* s; const int* t_off; const int* t_src; const float* g; const float* inv_n;
    const float* kp; int nq, ns, K, cout; float extent; int influence, aggregation;
    int deformed;   // kp is per query [nq, K, 3] (deformed kernel points) and the in-range filter of blocks.py:300-324 applies
};
bool kp2t_supported(int nq, int cout);
// G [ns, K, cout] = sum over listing queries of w * inv_n * grad_out rows
int kp2t_correlate_launch(const Kp2tArgs& a, float* G, cudaStream_t stream);

// fused forward (kpconv_fused.cu): gather + correlation + tcgen05 contraction + 1/n, bias, LeakyReLU in one kernel.
// wf may be null (not written).  Needs 16-byte aligned x / out and Ns * Cin < 2^31.
bool kpf_fused_eligible(int H, int K, int cin, int cout);
int kpf_fused_launch(const Kp2Args& g, const float* weights, const float* bias, int act, float slope, float* out,
                     float* inv_n, float* wf, cudaStream_t stream);
