// Optimiser step of the reference trainer on flat buffers, with no host round trip:
//   SGD (momentum, weight decay, dampening 0: training_3DMatch.py:62-69 -> torch.optim.SGD)
//   + the "skip the step if any gradient is not finite" guard of trainer.py:104-111 as a DEVICE predicate
//   + a learning rate read from device memory (ExponentialLR multiplies it between epochs, training_3DMatch.py:77-80).
// Two kernels: a finite scan of the gradient (grid-stride, 128-bit loads, one atomicOr per block that saw a bad value) and
// the update, which returns early when the flag is set -- so the whole step stays inside a CUDA graph.
#include "common.cuh"

namespace {

__device__ __forceinline__ bool nonfinite4(const float4 v) {
    const unsigned m = 0x7f800000u;
    return ((__float_as_uint(v.x) & m) == m) | ((__float_as_uint(v.y) & m) == m) | ((__float_as_uint(v.z) & m) == m) |
           ((__float_as_uint(v.w) & m) == m);
}

__global__ void __launch_bounds__(256)
sgd_scan_kernel(const float* __restrict__ g, size_t n, int32_t* __restrict__ flag) {
    const size_t n4 = n >> 2;
    bool bad = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
        bad |= nonfinite4(__ldg((const float4*)g + i));
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const float v = g[(n4 << 2) + threadIdx.x];
        bad |= (__float_as_uint(v) & 0x7f800000u) == 0x7f800000u;
    }
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(flag, 1);
}

// p, m updated in place;  g' = g + wd * p;  m = mu * m + g';  p -= lr * m      (torch.optim.SGD, dampening 0, no nesterov;
// a zero-initialised momentum buffer reproduces torch's "first step: buf = g'")
__global__ void __launch_bounds__(256)
sgd_apply_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, size_t n,
                 const float* __restrict__ lr_ptr, float mu, float wd, const int32_t* __restrict__ skip, int zero_grads) {
    const size_t n4 = n >> 2;
    if (skip && *skip) {       // non-finite gradient: no update (trainer.py:104-111); the gradients are still consumed
        if (zero_grads) {
            for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
                ((float4*)g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (blockIdx.x == 0 && threadIdx.x < (n & 3)) g[(n4 << 2) + threadIdx.x] = 0.f;
        }
        return;
    }
    const float lr = *lr_ptr;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 pv = ((float4*)p)[i], mv = ((float4*)m)[i];
        const float4 gv = ((const float4*)g)[i];
        if (zero_grads) ((float4*)g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        mv.x = mu * mv.x + (gv.x + wd * pv.x); mv.y = mu * mv.y + (gv.y + wd * pv.y);
        mv.z = mu * mv.z + (gv.z + wd * pv.z); mv.w = mu * mv.w + (gv.w + wd * pv.w);
        pv.x -= lr * mv.x; pv.y -= lr * mv.y; pv.z -= lr * mv.z; pv.w -= lr * mv.w;
        ((float4*)m)[i] = mv;
        ((float4*)p)[i] = pv;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const size_t i = (n4 << 2) + threadIdx.x;
        const float mv = mu * m[i] + (g[i] + wd * p[i]);
        m[i] = mv;
        p[i] -= lr * mv;
        if (zero_grads) g[i] = 0.f;
    }
}

}  // namespace

// params / grads / momentum: flat fp32 buffers of n elements (16-byte aligned); lr: device float; nonfinite_flag: device
// int32 that is OR-ed with 1 when a gradient element is inf / nan (the caller clears it; when set the update is skipped,
// exactly trainer.py:104-111).  check_finite = 0 skips the scan (flag is still honoured).  zero_grads != 0 clears the
// gradient buffer in the same pass (optimizer.zero_grad() of the next step: the backward kernels may then accumulate
// into it without their own zero fills).
extern "C" int d3f_sgd_step(float* params, float* grads, float* momentum_buf, size_t n, const float* lr,
                            float momentum, float weight_decay, int32_t* nonfinite_flag, int check_finite,
                            int zero_grads, d3f_stream stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n == 0) return D3F_OK;
    D3F_REQUIRE(params && grads && momentum_buf && lr, D3F_ERR_INVALID, "null pointer");
    D3F_REQUIRE((((size_t)params | (size_t)grads | (size_t)momentum_buf) & 15) == 0, D3F_ERR_INVALID,
                "flat buffers must be 16-byte aligned");
    const size_t n4 = (n >> 2) + 1;
    const int grid = (int)((n4 + 255) / 256 < 148 * 8 ? (n4 + 255) / 256 : 148 * 8);
    if (check_finite) {
        D3F_REQUIRE(nonfinite_flag, D3F_ERR_INVALID, "check_finite needs a flag");
        sgd_scan_kernel<<<grid, 256, 0, stream>>>(grads, n, nonfinite_flag);
        D3F_CHECK_LAUNCH();
    }
    sgd_apply_kernel<<<grid, 256, 0, stream>>>(params, grads, momentum_buf, n, lr, momentum, weight_decay, nonfinite_flag,
                                               zero_grads);
    D3F_CHECK_LAUNCH();
    return D3F_OK;
}
