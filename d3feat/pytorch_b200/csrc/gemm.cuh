// Internal interface of the 3xTF32 tensor-core GEMM (gemm.cu), shared with kpconv.cu.
#pragma once
#include <cuda_runtime.h>

struct D3fGemm {
    int M, N, K;
    const float* A; int lda;
    const float* B; int ldb;
    float* C; int ldc;
    const float* rs;     // optional row scale  [M]   (applied to C rows)
    const float* ks;     // optional k scale    [K]   (applied to B[k][:]; non-transposed B only)
    const float* bias;   // optional bias       [N]
    int act;             // 1 = LeakyReLU(slope) after the bias
    float slope;
    int k_per_split;     // filled by the launcher
};

// C[M,N] = act(rs[m] * sum_k opA(m,k) * ks[k] * opB(k,n) + bias[n]);  ta: A stored [K,M];  tb: B stored [N,K]
int d3f_gemm_launch(const D3fGemm& g, bool ta, bool tb, cudaStream_t stream);
