// Skinny-N variant of the fp32-accurate GEMM:  C[M,N] = act(rs[m] * A[M,K] B[K,N] + bias + bias2 + res),  N <= 64, large M.
//
// The KPConv contraction ([40000 x 480] x [480 x 32] at level 0) spends its time in the tcgen05 kernel re-reading the
// 128-row A tile from shared memory for every MMA (N = 32 gives each MMA 16 cycles of math for 40 cycles of operand
// fetch) behind a per-K-tile barrier chain (profiles/r1e_micro_kpconv_gemm.txt: 42-45 us, 2.8 us per K tile per CTA).
// Here the roles are swapped: the small operand B (K x N, the weights) is split into tf32 hi / remainder ONCE per CTA
// and parked in shared memory already in mma.sync fragment order (one conflict-free LDS.128 per fragment), while A is
// streamed straight from global memory / L2 into registers as 128-bit loads -- no shared-memory staging of A, no
// barrier in the main loop.  A lane's float4 holds 4 consecutive k of one row; the K index inside each group of 16 is
// permuted consistently on both operands (position p of k-step u <-> k = 4*(p%4) + 2u + p/4) so that float4 maps onto
// the m16n8k8 A fragment of two k-steps without shuffles.  One warp owns a 16-row tile over the whole K range
// (deterministic, M-independent summation order: usable by the forward pass); the warp count per CTA is chosen so
// that the tiles fill 148 SMs in whole waves.
#include "common.cuh"
#include "gemm.cuh"
#include <stdlib.h>

namespace {

__device__ __forceinline__ void sk_split(float v, uint32_t& hi, uint32_t& lo) {
    hi = (__float_as_uint(v) + 0x1000u) & 0xffffe000u;      // round to tf32 by hand (cvt.rna is emulated on sm_100a)
    lo = __float_as_uint(v - __uint_as_float(hi));          // remainder: the tensor core drops its low 13 bits
}

__device__ __forceinline__ void sk_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// NT8 = N tile / 8 (4 or 8); KC = K chunk resident in shared memory (multiple of 16)
template <int NT8, bool TB>
__global__ void __launch_bounds__(576)
sk_gemm_kernel(D3fGemm g, int KC, int n_tiles, int total_warps) {
    extern __shared__ float4 bs[];        // [KC/16][2][NT8][32] float4 {b0_hi, b1_hi, b0_lo, b1_lo}
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int gq = lane >> 2, tq = lane & 3;
    constexpr int N8 = NT8 * 8;
    float* bsf = (float*)bs;

    for (int tile = blockIdx.x * nwarps + warp, round = 0; ; tile += total_warps, ++round) {
        // every warp of the CTA runs the same number of rounds (barriers below); surplus warps idle through them
        const int first = blockIdx.x * nwarps + round * total_warps;
        if (first >= n_tiles) break;                        // uniform across the CTA
        const bool active = tile < n_tiles;
        const int r0 = tile * 16;
        const int rowA = r0 + gq, rowB = r0 + gq + 8;
        const bool okA = active && rowA < g.M, okB = active && rowB < g.M;
        const float* pa = g.A + (size_t)(okA ? rowA : 0) * g.lda + 4 * tq;
        const float* pb = g.A + (size_t)(okB ? rowB : 0) * g.lda + 4 * tq;
        float acc[NT8][4];
#pragma unroll
        for (int nt = 0; nt < NT8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;

        for (int k0 = 0; k0 < g.K; k0 += KC) {
            const int kc = min(KC, g.K - k0);               // multiple of 16
            const bool resident = g.K <= KC && round > 0;   // a single chunk stays in shared memory across rounds
            if (!resident) __syncthreads();                 // previous chunk / round fully consumed
            // ---- B chunk -> shared memory in fragment order (zero beyond N)
            for (int t = resident ? kc * N8 : tid; t < kc * N8; t += blockDim.x) {
                int kk, n;
                if (TB) { kk = t % kc; n = t / kc; } else { n = t % N8; kk = t / N8; }   // contiguous in global memory
                float v = 0.f;
                if (n < g.N) {
                    const int k = k0 + kk;
                    if (TB) {
                        const size_t kb = g.bblk ? (size_t)(k / g.bblk) * g.bblk_stride + (k % g.bblk) : (size_t)k;
                        v = g.B[(size_t)n * g.ldb + kb];
                    } else {
                        v = g.B[(size_t)k * g.ldb + n];
                    }
                }
                uint32_t hi, lo;
                sk_split(v, hi, lo);
                const int s = kk >> 4, kl = kk & 15, q = kl >> 2, r = kl & 3, u = r >> 1, which = r & 1;
                const int ln = (n & 7) * 4 + q, nt = n >> 3;
                float* dst = bsf + ((((size_t)(s * 2 + u) * NT8 + nt) * 32 + ln) << 2);
                dst[which] = __uint_as_float(hi);
                dst[2 + which] = __uint_as_float(lo);
            }
            if (!resident) __syncthreads();
            if (active) {
                // A: a ring of PF groups (16 k each) per row pair is always in flight -- with one group the kernel kept
                // ~2.5 MB in flight chip-wide and ran at 1.9 TB/s (profiles/r1m_micro_kpconv_gemm.txt)
                const int ngroups = kc >> 4;
                constexpr int PF = 4;
                float4 ra[PF], rb[PF];
#pragma unroll
                for (int j = 0; j < PF; ++j) {
                    ra[j] = rb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (j < ngroups) {
                        if (okA) ra[j] = __ldg((const float4*)(pa + k0 + 16 * j));
                        if (okB) rb[j] = __ldg((const float4*)(pb + k0 + 16 * j));
                    }
                }
                for (int s0 = 0; s0 < ngroups; s0 += PF) {
#pragma unroll
                    for (int j = 0; j < PF; ++j) {
                        const int s = s0 + j;
                        if (s < ngroups) {
                            const float4 va = ra[j], vb = rb[j];
                            if (s + PF < ngroups) {
                                if (okA) ra[j] = __ldg((const float4*)(pa + k0 + 16 * (s + PF)));
                                if (okB) rb[j] = __ldg((const float4*)(pb + k0 + 16 * (s + PF)));
                            }
                            const float xa[4] = {va.x, va.y, va.z, va.w}, xb[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
                            for (int u = 0; u < 2; ++u) {
                                uint32_t ah[4], al[4];
                                sk_split(xa[2 * u], ah[0], al[0]);          // a0: row gq,   position tq
                                sk_split(xb[2 * u], ah[1], al[1]);          // a1: row gq+8, position tq
                                sk_split(xa[2 * u + 1], ah[2], al[2]);      // a2: row gq,   position tq+4
                                sk_split(xb[2 * u + 1], ah[3], al[3]);      // a3: row gq+8, position tq+4
                                const float4* bp = bs + ((size_t)(s * 2 + u) * NT8) * 32 + lane;
#pragma unroll
                                for (int nt = 0; nt < NT8; ++nt) {
                                    const float4 b = bp[nt * 32];
                                    sk_mma(acc[nt], al, __float_as_uint(b.x), __float_as_uint(b.y));
                                    sk_mma(acc[nt], ah, __float_as_uint(b.z), __float_as_uint(b.w));
                                    sk_mma(acc[nt], ah, __float_as_uint(b.x), __float_as_uint(b.y));
                                }
                            }
                        }
                    }
                }
            }
        }
        // ---- epilogue: c0 (gq, 2tq) c1 (gq, 2tq+1) c2 (gq+8, 2tq) c3 (gq+8, 2tq+1) of n-tile nt
        if (active) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int row = r0 + gq + 8 * half;
                if (row >= g.M) continue;
                const float sc = g.rs ? g.rs[row] : 1.0f;
#pragma unroll
                for (int nt = 0; nt < NT8; ++nt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int n = nt * 8 + 2 * tq + e;
                        if (n >= g.N) continue;
                        float v = acc[nt][2 * half + e] * sc;
                        if (g.bias) v += g.bias[n];
                        if (g.bias2) v += g.bias2[n];
                        if (g.res) v += g.res[(size_t)row * g.ldr + n];
                        if (g.act) v = v > 0.f ? v : v * g.slope;
                        g.C[(size_t)row * g.ldc + n] = v;
                    }
            }
        }
    }
}

int g_skinny = -1;
int skinny_enabled() {
    if (g_skinny < 0) {
        const char* e = getenv("D3F_GEMM_SKINNY");
        g_skinny = e ? (e[0] == '0' ? 0 : (e[0] == '2' ? 2 : 1)) : D3F_GEMM_SKINNY_DEFAULT;
    }
    return g_skinny;
}

}  // namespace

extern "C" void d3f_set_gemm_skinny(int on) { g_skinny = on < 0 ? -1 : (on > 2 ? 2 : on); }

// true if the problem is one this kernel takes (the caller has already ruled out split-K)
bool d3f_gemm_skinny_eligible(const D3fGemm& g, bool ta, bool tb) {
    const int mode = skinny_enabled();      // 1 = only where it was measured to win, 2 = every problem the kernel takes
    if (!mode || ta || g.ks || g.partial) return false;
    if (g.N < 1 || g.N > 64 || g.M < 2048 || g.K < 16 || (g.K & 15)) return false;
    if (mode == 1 && !(g.M >= 16384 && g.K <= (g.N <= 32 ? 480 : 240))) return false;   // one resident B chunk, >= 7 warps/SM
    if ((g.lda & 3) || (((size_t)g.A) & 15)) return false;
    if (g.bblk && (!tb || (g.bblk & 15))) return false;
    return true;
}

int d3f_gemm_skinny_launch(const D3fGemm& g, bool tb, cudaStream_t stream) {
    const int nt8 = g.N <= 32 ? 4 : 8;
    const int KC = nt8 == 4 ? 480 : 240;                     // 120 KB of B fragments (hi + lo) per chunk
    const int kc = g.K < KC ? g.K : KC;
    const size_t smem = (size_t)kc * nt8 * 8 * 8;            // kc * N8 * 2 floats... = kc/16 * 2 * nt8 * 32 * 16 bytes
    const int n_tiles = d3f_ceil_div(g.M, 16);
    // whole waves over 148 SMs: the fewest rounds r whose warp count per CTA fits, then the warps that cover the tiles
    int warps = 18, rounds = 1;
    for (rounds = 1; ; ++rounds) {
        warps = d3f_ceil_div(n_tiles, 148 * rounds);
        if (warps <= 18) break;
    }
    if (warps < 4) warps = 4;
    const int ctas = d3f_ceil_div(n_tiles, warps * rounds) < 148 ? d3f_ceil_div(n_tiles, warps * rounds) : 148;
    const int total_warps = ctas * warps;
#define SK_GO(NT8_, TB_)                                                                                          \
    do {                                                                                                          \
        static size_t attr = 0;                                                                                   \
        if (smem > attr) {                                                                                        \
            D3F_CHECK_CUDA(cudaFuncSetAttribute(sk_gemm_kernel<NT8_, TB_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                (int)(122880)));                                                  \
            attr = 122880;                                                                                        \
        }                                                                                                         \
        sk_gemm_kernel<NT8_, TB_><<<ctas, warps * 32, smem, stream>>>(g, KC, n_tiles, total_warps);               \
    } while (0)
    if (nt8 == 4) { if (tb) SK_GO(4, true); else SK_GO(4, false); }
    else { if (tb) SK_GO(8, true); else SK_GO(8, false); }
#undef SK_GO
    D3F_CHECK_LAUNCH();
    return D3F_OK;
}
