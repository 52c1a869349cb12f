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
    float* partial;      // filled by the launcher: deterministic split-K partials [splits][M][N], or null
    const float* bias2;  // optional second bias [N] (UnaryBlock: Linear bias + the learned bias that replaces batch norm)
    const float* res;    // optional residual [M, ldr] added before the activation (ResnetBottleneckBlock shortcut)
    int ldr;
    // B stored [N, K] (tb) in K blocks: element (n, k) lives at B[(k / bblk) * bblk_stride + n * ldb + k % bblk]
    // (bblk = 0: plain).  KPConv backward reads W [K_pts, Cin, Cout] as B^T[c][k * Cout + o] this way without a
    // transposed copy of the weights; bblk must be a multiple of 32.
    int bblk; long long bblk_stride;
    // C written in column blocks: element (m, n) lives at C[(n / cblk) * cblk_stride + m * ldc + n % cblk] (cblk = 0: plain;
    // cblk a multiple of 4).  KPConv's weight gradient from the transposed gather, dW[k][c][o] = sum_j x[j][c] G[j][k*Cout+o],
    // lands in the [K_pts, Cin, Cout] layout of the weights this way (cblk = Cout, cblk_stride = Cin*Cout, ldc = Cout).
    int cblk; long long cblk_stride;
    // ctrans (with cblk): the blocks run along M and each block is stored transposed: element (m, n) lives at
    // C[(m / cblk) * cblk_stride + n * ldc + m % cblk] -- dW^T = G^T x ([K*Cout, Cin]: 480 useful rows per 128-row tile
    // instead of 32) written straight into the [K_pts, Cin, Cout] weight layout.
    int ctrans;
    // the caller guarantees C is zero on entry (a gradient buffer cleared once per step): split-K partial sums are
    // accumulated with atomics WITHOUT the library's own zero fill (one memset node less per GEMM in the step's graph)
    int c_zeroed;
};

// C[M,N] = act(rs[m] * sum_k opA(m,k) * ks[k] * opB(k,n) + bias[n] + bias2[n] + res[m,n]);  ta: A stored [K,M];  tb: B stored [N,K]
// `det_ws` != null selects the DETERMINISTIC split-K used by the forward pass: the split size depends on K
// only (never on M), partial tiles go to det_ws and are summed in split order by a second kernel, so a row of C
// is bit-identical run to run and independent of how many (padding) rows M has.  Without it, split-K partials
// are combined with float atomics (backward GEMMs: order-dependent at the 1e-7 level, like the scatter-adds).
size_t d3f_gemm_det_workspace_bytes(int M, int N, int K);
int d3f_gemm_launch(const D3fGemm& g, bool ta, bool tb, cudaStream_t stream, float* det_ws = nullptr,
                    size_t det_ws_bytes = 0);
