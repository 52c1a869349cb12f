// fp32-accurate tensor-core GEMM (3xTF32 error-compensated: a_hi*b_hi + a_hi*b_lo + a_lo*b_hi with fp32
// accumulation) with fused epilogues.  Used for every dense contraction on the hot path:
//   KPConv   out = diag(1/n) wf W          (NN, row scale)        models/blocks.py:369-380
//            dwf = diag(1/n) g W^T         (NT, row scale)
//            dW  = wf^T diag(1/n) g        (TN, k scale, split-K)
//   Unary    y = leaky(x W^T + b)          (NT, bias + LeakyReLU)  models/blocks.py:481-515
//            dx = dz W (NN), dW = dz^T x (TN)
// plain TF32 in the K*Cin contraction alone costs 4.5e-5 of the 1e-4 budget, so the
// operands are split in registers (cvt.rna.tf32) and three MMAs are issued per product term.
//
// Tiling: CTA 128x64x32, 8 warps (4 along M x 2 along N), warp tile 32x32 = 2x4 m16n8k8 MMAs, operands
// staged K-major in shared memory with a +4 float row pad (conflict-free fragment loads), register
// prefetch of the next K tile + two shared-memory stages (one barrier per K tile).
#include "common.cuh"
#include "gemm.cuh"
#include <stdlib.h>

namespace {

constexpr int BM = 128, BN = 64, BK = 32, NT = 256;
constexpr int LDS = BK + 4;      // K-major tiles  [rows][LDS]: fragment bank = 4*row + k  -> conflict-free
constexpr int LDA_T = BM + 8;    // M-major A tile [BK][LDA_T] (TA):  fragment bank = 8*k + m -> conflict-free
constexpr int LDB_N = BN + 8;    // N-major B tile [BK][LDB_N] (!TB): fragment bank = 8*k + n -> conflict-free
constexpr int A_STAGE = BM * LDS;   // >= BK * LDA_T
constexpr int B_STAGE = BN * LDS;   // == BK * LDB_N

__device__ __forceinline__ uint32_t to_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}

__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
    hi = to_tf32(v);
    lo = to_tf32(v - __uint_as_float(hi));
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// up to 4 consecutive floats starting at p; `valid` of them are in bounds
__device__ __forceinline__ float4 ld4(const float* __restrict__ p, int valid, bool vec) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid >= 4 && vec) return __ldg((const float4*)p);
    if (valid > 0) v.x = __ldg(p);
    if (valid > 1) v.y = __ldg(p + 1);
    if (valid > 2) v.z = __ldg(p + 2);
    if (valid > 3) v.w = __ldg(p + 3);
    return v;
}

// TA: A is [K, M] (read transposed); TB: B is [N, K] (K-major already)
template <bool TA, bool TB>
__global__ void __launch_bounds__(NT)
tc_gemm_kernel(D3fGemm g) {
    extern __shared__ float smem[];
    float* As = smem;                       // [2][A_STAGE]
    float* Bs = smem + 2 * A_STAGE;         // [2][B_STAGE]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gq = lane >> 2, tq = lane & 3;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * g.k_per_split, kend = min(g.K, kbeg + g.k_per_split);
    const int wm = (warp & 3) * 32, wn = (warp >> 2) * 32;
    const bool a_vec = (g.lda & 3) == 0 && (((size_t)g.A) & 15) == 0;
    const bool b_vec = (g.ldb & 3) == 0 && (((size_t)g.B) & 15) == 0;

    float acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;

    float4 ra[4], rb[2];
    auto load_tile = [&](int k0) {
        if (!TA) {  // A[m][k], contiguous along k: 128 rows x 8 float4
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int m = m0 + (tid >> 3) + 32 * r, k = k0 + (tid & 7) * 4;
                ra[r] = (m < g.M) ? ld4(g.A + (size_t)m * g.lda + k, kend - k, a_vec) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else {    // A[k][m], contiguous along m: 32 k-rows x 32 float4
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int k = k0 + (tid >> 5) + 8 * r, m = m0 + (tid & 31) * 4;
                ra[r] = (k < kend) ? ld4(g.A + (size_t)k * g.lda + m, g.M - m, a_vec) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        if (TB) {   // B[n][k], contiguous along k: 64 rows x 8 float4
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int n = n0 + (tid >> 3) + 32 * r, k = k0 + (tid & 7) * 4;
                const size_t kb = g.bblk ? (size_t)(k / g.bblk) * g.bblk_stride + (k % g.bblk) : (size_t)k;
                rb[r] = (n < g.N) ? ld4(g.B + (size_t)n * g.ldb + kb, kend - k, b_vec) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else {    // B[k][n], contiguous along n: 32 k-rows x 16 float4
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int k = k0 + (tid >> 4) + 16 * r, n = n0 + (tid & 15) * 4;
                float4 v = (k < kend) ? ld4(g.B + (size_t)k * g.ldb + n, g.N - n, b_vec) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (g.ks && k < kend) { const float s = g.ks[k]; v.x *= s; v.y *= s; v.z *= s; v.w *= s; }
                rb[r] = v;
            }
        }
    };
    auto store_tile = [&](int buf) {
        float* as = As + buf * A_STAGE;
        float* bs = Bs + buf * B_STAGE;
        if (!TA) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
                *(float4*)&as[((tid >> 3) + 32 * r) * LDS + (tid & 7) * 4] = ra[r];
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r)
                *(float4*)&as[((tid >> 5) + 8 * r) * LDA_T + (tid & 31) * 4] = ra[r];
        }
        if (TB) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
                *(float4*)&bs[((tid >> 3) + 32 * r) * LDS + (tid & 7) * 4] = rb[r];
        } else {
#pragma unroll
            for (int r = 0; r < 2; ++r)
                *(float4*)&bs[((tid >> 4) + 16 * r) * LDB_N + (tid & 15) * 4] = rb[r];
        }
    };

    const int nk = (kend - kbeg + BK - 1) / BK;
    if (nk > 0) {
        load_tile(kbeg);
        store_tile(0);
    }
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        if (kt + 1 < nk) load_tile(kbeg + (kt + 1) * BK);
        const float* as = As + (kt & 1) * A_STAGE + (TA ? wm : wm * LDS);
        const float* bs = Bs + (kt & 1) * B_STAGE + (TB ? wn * LDS : wn);
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks) {
            uint32_t ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) {
                const float* p = TA ? as + (ks * 8 + tq) * LDA_T + mi * 16 + gq
                                    : as + (mi * 16 + gq) * LDS + ks * 8 + tq;
                constexpr int DM = TA ? 8 : 8 * LDS, DK = TA ? 4 * LDA_T : 4;   // +8 rows of m, +4 of k
                split_tf32(p[0], ah[mi][0], al[mi][0]);
                split_tf32(p[DM], ah[mi][1], al[mi][1]);
                split_tf32(p[DK], ah[mi][2], al[mi][2]);
                split_tf32(p[DM + DK], ah[mi][3], al[mi][3]);
            }
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const float* p = TB ? bs + (ni * 8 + gq) * LDS + ks * 8 + tq
                                    : bs + (ks * 8 + tq) * LDB_N + ni * 8 + gq;
                constexpr int DKB = TB ? 4 : 4 * LDB_N;
                split_tf32(p[0], bh[ni][0], bl[ni][0]);
                split_tf32(p[DKB], bh[ni][1], bl[ni][1]);
            }
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
                    mma_tf32(acc[mi][ni], al[mi], bh[ni]);
                    mma_tf32(acc[mi][ni], ah[mi], bl[ni]);
                    mma_tf32(acc[mi][ni], ah[mi], bh[ni]);
                }
        }
        if (kt + 1 < nk) store_tile((kt + 1) & 1);
        __syncthreads();
    }

    // epilogue: rows m0+wm+mi*16+gq (+8), cols n0+wn+ni*8+2*tq (+1)
    if (g.partial) {   // deterministic split-K: raw partial sums, reduced in split order by gemm_reduce_kernel
        float* part = g.partial + (size_t)blockIdx.z * g.M * g.N;
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int m = m0 + wm + mi * 16 + gq + half * 8;
                if (m >= g.M) continue;
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int n = n0 + wn + ni * 8 + 2 * tq + e;
                        if (n < g.N) part[(size_t)m * g.N + n] = acc[mi][ni][half * 2 + e];
                    }
            }
        return;
    }
    const bool atomic = gridDim.z > 1;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int m = m0 + wm + mi * 16 + gq + half * 8;
            if (m >= g.M) continue;
            const float sc = g.rs ? g.rs[m] : 1.0f;
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int n = n0 + wn + ni * 8 + 2 * tq + e;
                    if (n >= g.N) continue;
                    float v = acc[mi][ni][half * 2 + e] * sc;
                    float* dst = g.ctrans ? g.C + (size_t)(m / g.cblk) * g.cblk_stride + (size_t)n * g.ldc + (m % g.cblk)
                                 : g.cblk ? g.C + (size_t)(n / g.cblk) * g.cblk_stride + (size_t)m * g.ldc + (n % g.cblk)
                                          : g.C + (size_t)m * g.ldc + n;
                    if (atomic) { atomicAdd(dst, v); continue; }
                    if (g.bias) v += g.bias[n];
                    if (g.bias2) v += g.bias2[n];
                    if (g.res) v += g.res[(size_t)m * g.ldr + n];
                    if (g.act) v = v > 0.f ? v : v * g.slope;
                    *dst = v;
                }
        }
}

__global__ void gemm_reduce_kernel(const float* __restrict__ part, int splits, int M, int N, float* __restrict__ C,
                                   int ldc, const float* __restrict__ rs, const float* __restrict__ bias, int act,
                                   float slope, const float* __restrict__ bias2, const float* __restrict__ res, int ldr) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)M * N) return;
    const int m = (int)(t / N), n = (int)(t % N);
    float v = 0.f;
    for (int z = 0; z < splits; ++z) v += part[(size_t)z * M * N + t];   // fixed order
    if (rs) v *= rs[m];
    if (bias) v += bias[n];
    if (bias2) v += bias2[n];
    if (res) v += res[(size_t)m * ldr + n];
    if (act) v = v > 0.f ? v : v * slope;
    C[(size_t)m * ldc + n] = v;
}

// deterministic split: a function of K only (so the summation order of a row never depends on M, i.e. on how many
// capacity-padding rows the static pipeline adds).  Round 1c (ncu): the L1 contraction [13312 x 960] x [960 x 64] ran
// as 104 CTAs x 30 serial K tiles (89 us) because large-M problems were never split; every K >= 512 problem is split
// now, which also covers the small-M / large-K unary GEMMs of the deep levels (16-44 tiles on 148 SMs otherwise).
constexpr int DET_KPS = 256;
inline int det_splits(int M, int K) { (void)M; return K >= 2 * DET_KPS ? d3f_ceil_div(K, DET_KPS) : 1; }

}  // namespace

size_t d3f_gemm_det_workspace_bytes(int M, int N, int K) {
    const int s = det_splits(M, K);
    return s > 1 ? sizeof(float) * (size_t)s * M * N : 0;
}

int d3f_gemm_tcgen05_launch(const D3fGemm& g, bool ta, bool tb, int splits, cudaStream_t stream);

// 1 = tcgen05 / TMEM kernel (gemm_tcgen05.cu, default), 0 = legacy mma.sync kernel (this file)
static int g_gemm_impl = -1;
extern "C" void d3f_set_gemm_impl(int use_tcgen05) { g_gemm_impl = use_tcgen05 ? 1 : 0; }
static int gemm_impl() {
    if (g_gemm_impl < 0) {
        const char* e = getenv("D3F_GEMM_IMPL");
        g_gemm_impl = (e && e[0] == 'm') ? 0 : 1;
    }
    return g_gemm_impl;
}

// debug / tuning (include/d3feat_b200_debug.h): force the N tile width and the number of atomically combined K splits
static int g_force_bn = 0, g_force_splits = 0;
extern "C" void d3f_set_gemm_tuning(int bn, int splits) { g_force_bn = bn; g_force_splits = splits; }
int d3f_gemm_forced_bn() { return g_force_bn; }

int d3f_gemm_launch(const D3fGemm& in, bool ta, bool tb, cudaStream_t stream, float* det_ws, size_t det_ws_bytes) {
    D3fGemm g = in;
    g.partial = nullptr;
    if (g.M <= 0 || g.N <= 0) return D3F_OK;
    const int tiles = d3f_ceil_div(g.M, BM) * d3f_ceil_div(g.N, BN);
    int splits = 1, kps;
    const bool plain = !g.bias && !g.act && !g.bias2 && !g.res;   // atomically combined partials cannot take an epilogue
    D3F_REQUIRE(!g.cblk || (!det_ws && (g.cblk & 3) == 0 && (g.ctrans ? g.M : g.N) % g.cblk == 0 && g.K > 0 && !g.bias &&
                            !g.act && !g.bias2 && !g.res && !g.rs),
                D3F_ERR_UNSUPPORTED, "blocked C needs the plain path without epilogue, cblk % 4 == 0 and a whole number of blocks");
    D3F_REQUIRE(!g.ctrans || g.cblk, D3F_ERR_INVALID, "ctrans needs cblk");
    if (det_ws) {
        splits = det_splits(g.M, g.K);
        kps = splits > 1 ? DET_KPS : d3f_ceil_div(g.K > 0 ? g.K : 1, BK) * BK;
        if (splits > 1) {
            D3F_REQUIRE(det_ws_bytes >= sizeof(float) * (size_t)splits * g.M * g.N, D3F_ERR_WORKSPACE,
                        "deterministic split-K workspace too small");
            g.partial = det_ws;
        }
    } else {
        if (g.K > 0 && plain) {
            // atomically combined partials.  Fitted to the sweeps of tools/gemm_tune.py over every GEMM of the step
            // (profiles/r2_gemm_tune*.txt, re-fitted after the K loop got faster): fill the SMs first (splitting is free
            // while CTAs < SMs: [256 x 1024 x 512] 25 us unsplit, 15 us in four); beyond that a split costs ~4-7 us of
            // reductions (+ a zero fill outside engine.PairStep's arena) and pays only while a CTA keeps >= 8 K tiles
            // ([768 x 1024 x 256], 96 tiles x 8 K tiles: 17 us unsplit, 21 us in three); never more than 32 K tiles in one
            // CTA's serial loop ([256 x 768 x 13312]: 144 us in three, 58 us in twelve).
            const int kt = d3f_ceil_div(g.K, BK);
            splits = min(148 / tiles, kt / 2);
            splits = max(splits, min(296 / tiles, kt / 8));
            splits = max(splits, d3f_ceil_div(kt, 32));
            splits = min(splits, kt / 2);
            if (splits < 1) splits = 1;
        }
        if (g_force_splits > 0 && plain && g.K > 0) splits = min(g_force_splits, d3f_ceil_div(g.K, BK));
        kps = d3f_ceil_div(d3f_ceil_div(g.K > 0 ? g.K : 1, splits), BK) * BK;
        splits = d3f_ceil_div(g.K > 0 ? g.K : 1, kps);
        const size_t c_floats = g.ctrans ? (size_t)(g.M / g.cblk) * g.cblk_stride
                                : g.cblk ? (size_t)(g.N / g.cblk) * g.cblk_stride : (size_t)g.M * g.ldc;
        if (splits > 1 && !g.c_zeroed) D3F_CHECK_CUDA(cudaMemsetAsync(g.C, 0, sizeof(float) * c_floats, stream));
    }
    g.k_per_split = kps;
    if (g.K == 0) {
        D3F_CHECK_CUDA(cudaMemsetAsync(g.C, 0, sizeof(float) * (size_t)g.M * g.ldc, stream));
        return D3F_OK;
    }
    if (gemm_impl() == 1) {
        int rc5 = d3f_gemm_tcgen05_launch(g, ta, tb, splits, stream);
        if (rc5) return rc5;
    } else {
    const size_t smem = sizeof(float) * 2 * (A_STAGE + B_STAGE);
    dim3 grid(d3f_ceil_div(g.N, BN), d3f_ceil_div(g.M, BM), splits);
#define LAUNCH(TA_, TB_)                                                                                   \
    do {                                                                                                   \
        static bool attr_set = false;                                                                      \
        if (!attr_set) {                                                                                   \
            D3F_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<TA_, TB_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            attr_set = true;                                                                               \
        }                                                                                                  \
        tc_gemm_kernel<TA_, TB_><<<grid, NT, smem, stream>>>(g);                                           \
    } while (0)
    if (ta && !tb) LAUNCH(true, false);
    else if (!ta && tb) LAUNCH(false, true);
    else if (!ta && !tb) LAUNCH(false, false);
    else { d3f_set_error("gemm: TT mode is not used on the hot path"); return D3F_ERR_UNSUPPORTED; }
#undef LAUNCH
    D3F_CHECK_LAUNCH();
    }
    if (g.partial) {
        const size_t total = (size_t)g.M * g.N;
        gemm_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(g.partial, splits, g.M, g.N, g.C, g.ldc,
                                                                              g.rs, g.bias, g.act, g.slope, g.bias2, g.res, g.ldr);
        D3F_CHECK_LAUNCH();
    }
    return D3F_OK;
}

// C ABI: generic entries used by the fused UnaryBlock (blocks.py) and by tests.
extern "C" size_t d3f_gemm_workspace_bytes(int M, int N, int K) { return d3f_gemm_det_workspace_bytes(M, N, K); }

extern "C" int d3f_gemm_ex(int trans_a, int trans_b, int M, int N, int K, const float* A, int lda, const float* B,
                           int ldb, float* C, int ldc, const float* row_scale, const float* k_scale,
                           const float* bias, const float* bias2, const float* residual, int ld_residual,
                           int leaky_relu, float slope, void* workspace, size_t workspace_bytes, d3f_stream stream) {
    D3F_REQUIRE(M >= 0 && N >= 0 && K >= 0, D3F_ERR_INVALID, "bad sizes");
    if (M == 0 || N == 0) return D3F_OK;
    D3F_REQUIRE(C && (K == 0 || (A && B)), D3F_ERR_INVALID, "null pointer");
    D3F_REQUIRE(!(k_scale && trans_b), D3F_ERR_UNSUPPORTED, "k_scale is applied on B[k][n] loads only");
    const size_t need = d3f_gemm_det_workspace_bytes(M, N, K);
    D3F_REQUIRE(need == 0 || (workspace && workspace_bytes >= need), D3F_ERR_WORKSPACE, "workspace too small");
    D3fGemm g{M, N, K, A, lda, B, ldb, C, ldc, row_scale, k_scale, bias, leaky_relu, slope, 0, nullptr,
              bias2, residual, ld_residual};
    float dummy;
    return d3f_gemm_launch(g, trans_a != 0, trans_b != 0, (cudaStream_t)stream, need ? (float*)workspace : &dummy, need);
}

static int gemm_plain(int trans_a, int trans_b, int M, int N, int K, const float* A, int lda, const float* B,
                      int ldb, float* C, int ldc, const float* row_scale, const float* k_scale,
                      const float* bias, int leaky_relu, float slope, int c_zeroed, d3f_stream stream) {
    D3F_REQUIRE(M >= 0 && N >= 0 && K >= 0, D3F_ERR_INVALID, "bad sizes");
    if (M == 0 || N == 0) return D3F_OK;
    D3F_REQUIRE(C && (K == 0 || (A && B)), D3F_ERR_INVALID, "null pointer");
    D3F_REQUIRE(!(k_scale && trans_b), D3F_ERR_UNSUPPORTED, "k_scale is applied on B[k][n] loads only");
    D3fGemm g{M, N, K, A, lda, B, ldb, C, ldc, row_scale, k_scale, bias, leaky_relu, slope, 0, nullptr};
    g.c_zeroed = c_zeroed && K > 0;
    return d3f_gemm_launch(g, trans_a != 0, trans_b != 0, (cudaStream_t)stream);
}

extern "C" int d3f_gemm(int trans_a, int trans_b, int M, int N, int K, const float* A, int lda, const float* B,
                        int ldb, float* C, int ldc, const float* row_scale, const float* k_scale,
                        const float* bias, int leaky_relu, float slope, d3f_stream stream) {
    return gemm_plain(trans_a, trans_b, M, N, K, A, lda, B, ldb, C, ldc, row_scale, k_scale, bias, leaky_relu, slope, 0, stream);
}

// d3f_gemm for a C that the caller has already cleared (a slice of a gradient buffer zeroed once per step)
extern "C" int d3f_gemm_prezeroed(int trans_a, int trans_b, int M, int N, int K, const float* A, int lda, const float* B,
                                  int ldb, float* C, int ldc, const float* row_scale, const float* k_scale,
                                  const float* bias, int leaky_relu, float slope, d3f_stream stream) {
    return gemm_plain(trans_a, trans_b, M, N, K, A, lda, B, ldb, C, ldc, row_scale, k_scale, bias, leaky_relu, slope, 1, stream);
}
