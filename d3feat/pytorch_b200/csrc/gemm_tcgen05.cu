// 5th-generation tensor-core version of the fp32-accurate GEMM (see gemm.cu for the contract):
// tcgen05.mma kind::tf32, 3xTF32 error compensation, accumulators in TMEM.
//
// One CTA (256 threads) computes a 128 x 64 tile.  Per 32-wide K tile all threads load A/B with coalesced
// 128-bit global loads, split every value into tf32 hi + fp32 remainder lo, and store the four operand
// tiles (A_hi, A_lo, B_hi, B_lo) into shared memory in the UMMA canonical K-MAJOR NO-SWIZZLE layout
//     off(r,k) = (k/4)*LBO + (r/8)*SBO + (r%8)*16 + (k%4)*4          (16-byte unit = 4 consecutive k of one row)
// Operands that are stored transposed in global memory ([k][rows]: the TN / NN modes) are transposed in
// registers (4x4) on the way in: for tf32 the only MN-major shared-memory layout the tensor core accepts is the
// special 128B_BASE32B swizzle (cutlass sm100_common.inl), so everything is fed K-major instead.
// SBO = 144 and LBO = (rows/8)*144 + 16 are padded so that every quarter-warp's 128-bit shared stores -- plain and
// transposed -- land in 8 different bank groups.  One elected thread then issues 4 k-steps x 3 tcgen05.mma
// (a_lo*b_hi, a_hi*b_lo, a_hi*b_hi) into a 128-lane x 64-column fp32 accumulator in tensor memory and commits
// to an mbarrier; two shared-memory stages let the loads of tile t+1 overlap the MMAs of tile t.
// Epilogue: tcgen05.ld (32 lanes x 32 columns per warp) -> row scale / bias / LeakyReLU -> global
// (plain, atomic split-K, or deterministic split-K partials), exactly as the mma.sync kernel.
#include "common.cuh"
#include "gemm.cuh"
#include <stdlib.h>

#ifndef D3F_GEMM_PIPELINE_DEFAULT
#define D3F_GEMM_PIPELINE_DEFAULT 0   // tc5 until a variant beats it on the GPU (round 1e: tc6 is slower everywhere)
#endif

namespace {

constexpr int BM = 128, BK = 32, NT = 256;
constexpr int SBO = 144;                              // stride between 8-row groups (128 + 16 pad)
constexpr int A_LBO = (BM / 8) * SBO + 16;            // stride between 16-byte k units: 2320
constexpr int A_TILE = (BK / 4) * A_LBO;              // 18560
// the N tile is a template parameter (32 / 64 / 128): skinny outputs do not pay for a half-empty MMA and wide ones
// re-read A half as often
template <int BN> struct Cfg {
    static constexpr int B_LBO = (BN / 8) * SBO + 16;
    static constexpr int B_TILE = (BK / 4) * B_LBO;
    static constexpr int STAGE_BYTES = 2 * A_TILE + 2 * B_TILE;   // hi + lo for A and B
    // ONE shared-memory stage: these GEMMs are short (K = 32..960 for most of them) and latency-bound, so the
    // latency is hidden by co-resident CTAs (4 per SM at 46-56 KB) rather than by a deep pipeline inside one CTA
    // (ncu, round 1: 2 stages -> 1-2 CTAs/SM, 12-22 % warps active, every pipe under 27 %).
    static constexpr int SMEM_BYTES = STAGE_BYTES + 64;           // one stage + barrier / tmem pointer
    static constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // cute::UMMA::SmemDescriptor: start[0,14) | LBO[16,30) | SBO[32,46) | version=1 [46,48) | layout_type=0 (no swizzle)
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ULL << 46);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* fail, long long max_spin = (1LL << 26)) {
    uint32_t done = 0;
    for (long long spin = 0; spin < max_spin; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
    }
    if (fail) atomicExch(fail, 1);   // never hang the GPU: give up and flag the launch
}

__device__ __forceinline__ float tf32_hi(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

__device__ __forceinline__ float4 ld4g(const float* __restrict__ p, int valid, bool vec) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid >= 4 && vec) return __ldg((const float4*)p);
    if (valid > 0) v.x = __ldg(p);
    if (valid > 1) v.y = __ldg(p + 1);
    if (valid > 2) v.z = __ldg(p + 2);
    if (valid > 3) v.w = __ldg(p + 3);
    return v;
}

__device__ __forceinline__ void st_split(char* hi, char* lo, uint32_t off, float4 v) {
    float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
    *(float4*)(hi + off) = h;
    *(float4*)(lo + off) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
}

__device__ int g_tc5_fail = 0;

// -DD3F_TC5_TIMING (tools/tc5_timing.py builds a separate diagnostic library): per-phase clock64() sums of the main
// loop for threads 0 (MMA issuer) and 255 of one CTA in the middle of the grid.  Never compiled into the product.
#ifdef D3F_TC5_TIMING
__device__ unsigned long long g_tc5_t[32];
#define TC5_T(i) do { if (timed) { const long long now_ = clock64(); tacc[i] += now_ - tlast; tlast = now_; } } while (0)
#else
#define TC5_T(i) do { } while (0)
#endif

template <bool TA, bool TB, int BN>
__global__ void __launch_bounds__(NT, BN >= 128 ? 3 : 4)
tc5_gemm_kernel(D3fGemm g) {
    constexpr int B_LBO = Cfg<BN>::B_LBO, B_TILE = Cfg<BN>::B_TILE, STAGE_BYTES = Cfg<BN>::STAGE_BYTES;
    constexpr uint32_t TMEM_COLS = Cfg<BN>::TMEM_COLS;
    extern __shared__ __align__(128) char smem[];
    uint64_t* bars = (uint64_t*)(smem + STAGE_BYTES);          // [0]: the MMAs of the current K tile are done
    uint32_t* tmem_ptr = (uint32_t*)(smem + STAGE_BYTES + 32);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * g.k_per_split, kend = min(g.K, kbeg + g.k_per_split);
    const int nk = (kend - kbeg + BK - 1) / BK;
    const bool a_vec = (g.lda & 3) == 0 && (((size_t)g.A) & 15) == 0;
    const bool b_vec = (g.ldb & 3) == 0 && (((size_t)g.B) & 15) == 0;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(tmem_ptr)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 32) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(&bars[0])), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_d = *tmem_ptr;

    // instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6) | a,b = TF32 (2) [7,10),[10,13) | a_major [15] |
    // b_major [16] | N>>3 [17,23) | M>>4 [24,29)
    // (both operands are fed K-major: a_major = b_major = 0)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

    float4 ra[4], rb[BN >= 128 ? BN / 32 : 4];
    auto load_tile = [&](int k0) {
        if (!TA) {      // A[m][k]: thread = (row, 16-byte k unit)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int m = m0 + (tid >> 3) + 32 * r, k = k0 + (tid & 7) * 4;
                ra[r] = (m < g.M) ? ld4g(g.A + (size_t)m * g.lda + k, kend - k, a_vec) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else {        // A[k][m]: thread = (4 k rows, 4 consecutive m), transposed in registers at store time
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = k0 + (tid >> 5) * 4 + j, m = m0 + (tid & 31) * 4;
                ra[j] = (k < kend) ? ld4g(g.A + (size_t)k * g.lda + m, g.M - m, a_vec) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        if (TB) {       // B[n][k]
#pragma unroll
            for (int r = 0; r < BN / 32; ++r) {
                const int n = n0 + (tid >> 3) + 32 * r, k = k0 + (tid & 7) * 4;
                const size_t kb = g.bblk ? (size_t)(k / g.bblk) * g.bblk_stride + (k % g.bblk) : (size_t)k;
                rb[r] = (n < g.N) ? ld4g(g.B + (size_t)n * g.ldb + kb, kend - k, b_vec) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else if (tid < 2 * BN) {   // B[k][n]: thread = (4 k rows, 4 consecutive n)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = k0 + (tid / (BN / 4)) * 4 + j, n = n0 + (tid % (BN / 4)) * 4;
                float4 v = (k < kend) ? ld4g(g.B + (size_t)k * g.ldb + n, g.N - n, b_vec) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (g.ks && k < kend) { const float sc = g.ks[k]; v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc; }
                rb[j] = v;
            }
        }
    };
    auto store_tile = [&]() {
        char* a_hi = smem;
        char* a_lo = a_hi + A_TILE;
        char* b_hi = a_lo + A_TILE;
        char* b_lo = b_hi + B_TILE;
        if (!TA) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int m = (tid >> 3) + 32 * r, k4 = tid & 7;
                st_split(a_hi, a_lo, k4 * A_LBO + (m >> 3) * SBO + (m & 7) * 16, ra[r]);
            }
        } else {
            const int k4 = tid >> 5, mb = (tid & 31) * 4;
            const float t[4][4] = {{ra[0].x, ra[1].x, ra[2].x, ra[3].x}, {ra[0].y, ra[1].y, ra[2].y, ra[3].y},
                                   {ra[0].z, ra[1].z, ra[2].z, ra[3].z}, {ra[0].w, ra[1].w, ra[2].w, ra[3].w}};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int m = mb + e;
                st_split(a_hi, a_lo, k4 * A_LBO + (m >> 3) * SBO + (m & 7) * 16, make_float4(t[e][0], t[e][1], t[e][2], t[e][3]));
            }
        }
        if (TB) {
#pragma unroll
            for (int r = 0; r < BN / 32; ++r) {
                const int n = (tid >> 3) + 32 * r, k4 = tid & 7;
                st_split(b_hi, b_lo, k4 * B_LBO + (n >> 3) * SBO + (n & 7) * 16, rb[r]);
            }
        } else if (tid < 2 * BN) {
            const int k4 = tid / (BN / 4), nb = (tid % (BN / 4)) * 4;
            const float t[4][4] = {{rb[0].x, rb[1].x, rb[2].x, rb[3].x}, {rb[0].y, rb[1].y, rb[2].y, rb[3].y},
                                   {rb[0].z, rb[1].z, rb[2].z, rb[3].z}, {rb[0].w, rb[1].w, rb[2].w, rb[3].w}};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int n = nb + e;
                st_split(b_hi, b_lo, k4 * B_LBO + (n >> 3) * SBO + (n & 7) * 16, make_float4(t[e][0], t[e][1], t[e][2], t[e][3]));
            }
        }
    };

#ifdef D3F_TC5_TIMING
    const bool timed = (tid == 0 || tid == 255) && blockIdx.x == 0 && blockIdx.y == gridDim.y / 2 && blockIdx.z == 0;
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64();
    const long long tstart = tlast;
#endif
    if (nk > 0) load_tile(kbeg);
    TC5_T(6);                                                                      // prologue: alloc, barrier init, first loads issued
    for (int kt = 0; kt < nk; ++kt) {
        if (kt >= 1) mbar_wait(smem_u32(&bars[0]), (kt - 1) & 1, &g_tc5_fail);   // MMAs of tile kt-1 have read the stage
        TC5_T(0);
        store_tile();
        TC5_T(1);
        if (kt + 1 < nk) load_tile(kbeg + (kt + 1) * BK);                          // global loads overlap the MMAs
        TC5_T(2);
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");              // generic-proxy stores -> async proxy (UMMA)
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        TC5_T(3);
        __syncthreads();
        TC5_T(4);
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint32_t a_hi = smem_u32(smem), a_lo = a_hi + A_TILE;
            const uint32_t b_hi = a_lo + A_TILE, b_lo = b_hi + B_TILE;
#pragma unroll
            for (int ks = 0; ks < BK / 8; ++ks) {
                // one MMA = 8 k values = two 16-byte k units
                const uint32_t ao = ks * 2 * A_LBO, bo = ks * 2 * B_LBO;
                const uint64_t dah = make_desc(a_hi + ao, A_LBO, SBO), dal = make_desc(a_lo + ao, A_LBO, SBO);
                const uint64_t dbh = make_desc(b_hi + bo, B_LBO, SBO), dbl = make_desc(b_lo + bo, B_LBO, SBO);
                mma_tf32(tmem_d, dal, dbh, idesc, (kt | ks) ? 1u : 0u);
                mma_tf32(tmem_d, dah, dbl, idesc, 1u);
                mma_tf32(tmem_d, dah, dbh, idesc, 1u);
            }
            // tcgen05.commit: arrive on the barrier when every MMA issued so far has completed
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
                         :: "r"(smem_u32(&bars[0])) : "memory");
        }
        TC5_T(5);
    }

    // ---- epilogue
    if (nk > 0) mbar_wait(smem_u32(&bars[0]), (nk - 1) & 1, &g_tc5_fail);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    TC5_T(0);
    // TMEM -> registers (one row per thread) -> shared C tile [128][BN+4] (row stride = 4 banks mod 32: the
    // 128-bit stores of a quarter warp are conflict-free) -> 128-bit row-contiguous global stores.
    constexpr int LDC_S = BN + 4;
    float* cs = (float*)smem;                       // the operand stages are free once bars[2] has completed
    {
        const int r_loc = (warp & 3) * 32 + lane, col0 = (warp >> 2) * (BN / 2);
#pragma unroll
        for (int part = 0; part < BN / 32; ++part) {
            uint32_t v[16];
            const uint32_t taddr = tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(col0 + part * 16);
            if (nk > 0) {
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                               "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                             : "r"(taddr) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) v[e] = 0u;
            }
#pragma unroll
            for (int e = 0; e < 16; e += 4)
                *(uint4*)&cs[r_loc * LDC_S + col0 + part * 16 + e] = make_uint4(v[e], v[e + 1], v[e + 2], v[e + 3]);
        }
    }
    __syncthreads();
    {
        const bool atomic = gridDim.z > 1 && !g.partial;
        constexpr int TPR = BN / 4, RPP = NT / TPR;      // threads per row, rows per pass
        const int c4 = (tid % TPR) * 4, n = n0 + c4;
        const bool vec_ok = g.partial ? ((g.N & 3) == 0) : ((g.ldc & 3) == 0 && (((size_t)g.C) & 15) == 0);
#pragma unroll
        for (int it = 0; it < BM / RPP; ++it) {
            const int r_loc = tid / TPR + RPP * it, row = m0 + r_loc;
            if (row >= g.M || n >= g.N) continue;
            float4 x = *(const float4*)&cs[r_loc * LDC_S + c4];
            float xs[4] = {x.x, x.y, x.z, x.w};
            if (g.partial) {
                float* dst = g.partial + (size_t)blockIdx.z * g.M * g.N + (size_t)row * g.N + n;
                if (vec_ok && n + 3 < g.N) *(float4*)dst = x;
                else for (int e = 0; e < 4; ++e) if (n + e < g.N) dst[e] = xs[e];
                continue;
            }
            const float sc = g.rs ? g.rs[row] : 1.0f;
            float* dst = g.C + (size_t)row * g.ldc + n;
            if (atomic) {
                for (int e = 0; e < 4; ++e) if (n + e < g.N) atomicAdd(dst + e, xs[e] * sc);
                continue;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float y = xs[e] * sc;
                if (n + e < g.N) {
                    if (g.bias) y += g.bias[n + e];
                    if (g.bias2) y += g.bias2[n + e];
                    if (g.res) y += g.res[(size_t)row * g.ldr + n + e];
                }
                if (g.act) y = y > 0.f ? y : y * g.slope;
                xs[e] = y;
            }
            if (vec_ok && n + 3 < g.N) *(float4*)dst = make_float4(xs[0], xs[1], xs[2], xs[3]);
            else for (int e = 0; e < 4; ++e) if (n + e < g.N) dst[e] = xs[e];
        }
    }
    TC5_T(7);                                                                      // epilogue
#ifdef D3F_TC5_TIMING
    if (timed) {
        unsigned long long* o = g_tc5_t + (tid == 0 ? 0 : 16);
        for (int i = 0; i < 8; ++i) o[i] = (unsigned long long)tacc[i];
        o[8] = (unsigned long long)(clock64() - tstart);
        o[9] = (unsigned long long)nk;
    }
#endif
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem_d), "r"(TMEM_COLS) : "memory");
}


// ------------------------------------------------------------------------------------------------------------------
// tc6: the same MMA / TMEM / epilogue path with the A operand fed by a cp.async ring.
//
// ncu (round 1d) showed the tc5 kernel latency-bound: one K tile per CTA in flight (held in registers), 12-26 % warps
// active, ~3 us per K tile when a problem has fewer CTAs than SMs.  Here every thread issues its 16-byte cp.async
// copies of A for K tile kt+S-1 into a ring of S raw fp32 stages before it converts tile kt, so S-1 A tiles (32-48 KB
// per CTA, two CTAs per SM) are always in flight and the global-memory latency leaves the per-tile dependency chain:
//     wait(A tile kt landed) -> barrier -> refill the slot freed by tile kt-1 -> wait(MMAs of kt-1 done)
//     -> raw stage -> registers -> tf32 hi / remainder -> UMMA-layout stage -> fence + barrier -> 12 tcgen05.mma
// B (the weight matrix in the NN / NT modes: small and L2-resident) keeps tc5's one-tile register prefetch.
// Needs a 16-byte aligned A (lda a multiple of 4): everything on the hot path except the K = 15 first layer.
template <int BN> struct Cfg6 {
    static constexpr int S = BN == 64 ? 3 : 4;                       // raw A stages
    static constexpr int RAW_A = BM * BK * 4;                        // bytes
    static constexpr int OP_BYTES = Cfg<BN>::STAGE_BYTES;            // A_hi | A_lo | B_hi | B_lo in UMMA layout
    static constexpr int SMEM_BYTES = OP_BYTES + S * RAW_A + 64;     // 32: 110.7 KB, 64: 103.6 KB (2 CTAs/SM); 128: 138.3 KB
};

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(smem_u32(dst_smem)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N) : "memory"); }

template <bool TA, bool TB, int BN>
__global__ void __launch_bounds__(NT, BN >= 128 ? 1 : 2)
tc6_gemm_kernel(D3fGemm g) {
    constexpr int B_LBO = Cfg<BN>::B_LBO, B_TILE = Cfg<BN>::B_TILE, OP_BYTES = Cfg6<BN>::OP_BYTES;
    constexpr int S = Cfg6<BN>::S, RAW_A = Cfg6<BN>::RAW_A, RAW_STAGE = Cfg6<BN>::RAW_A;
    constexpr uint32_t TMEM_COLS = Cfg<BN>::TMEM_COLS;
    extern __shared__ __align__(128) char smem[];
    char* raw = smem + OP_BYTES;
    uint64_t* bars = (uint64_t*)(smem + OP_BYTES + S * RAW_STAGE);
    uint32_t* tmem_ptr = (uint32_t*)(smem + OP_BYTES + S * RAW_STAGE + 32);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * g.k_per_split, kend = min(g.K, kbeg + g.k_per_split);
    const int nk = (kend - kbeg + BK - 1) / BK;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(tmem_ptr)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 32) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(&bars[0])), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }

    const bool b_vec = (g.ldb & 3) == 0 && (((size_t)g.B) & 15) == 0;
    // 16-byte copies of the A part of K tile `kt` into raw stage kt % S; out-of-range elements are zero-filled
    auto issue_tile = [&](int kt) {
        char* ra = raw + (kt % S) * RAW_STAGE;
        const int k0 = kbeg + kt * BK;
        if (!TA) {      // A[m][k] -> ra[m][32]
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int ml = (tid >> 3) + 32 * r, k4 = tid & 7;
                const int m = m0 + ml, k = k0 + k4 * 4;
                const int nb = (m < g.M) ? 4 * max(0, min(4, kend - k)) : 0;
                cp_async16(ra + (ml * 8 + k4) * 16, nb ? (const void*)(g.A + (size_t)m * g.lda + k) : (const void*)g.A, nb);
            }
        } else {        // A[k][m] -> ra[k][128]
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int kl = (tid >> 5) * 4 + j, m4 = tid & 31;
                const int k = k0 + kl, m = m0 + m4 * 4;
                const int nb = (k < kend) ? 4 * max(0, min(4, g.M - m)) : 0;
                cp_async16(ra + (kl * 32 + m4) * 16, nb ? (const void*)(g.A + (size_t)k * g.lda + m) : (const void*)g.A, nb);
            }
        }
    };
    float4 rb[BN >= 128 ? BN / 32 : 4];
    auto load_b = [&](int kt) {     // B part of K tile `kt` -> registers (as tc5)
        const int k0 = kbeg + kt * BK;
        if (TB) {       // B[n][k]
#pragma unroll
            for (int r = 0; r < BN / 32; ++r) {
                const int n = n0 + (tid >> 3) + 32 * r, k = k0 + (tid & 7) * 4;
                const size_t kb = g.bblk ? (size_t)(k / g.bblk) * g.bblk_stride + (k % g.bblk) : (size_t)k;
                rb[r] = (n < g.N) ? ld4g(g.B + (size_t)n * g.ldb + kb, kend - k, b_vec) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else if (tid < 2 * BN) {   // B[k][n]: thread = (4 k rows, 4 consecutive n)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = k0 + (tid / (BN / 4)) * 4 + j, n = n0 + (tid % (BN / 4)) * 4;
                float4 v = (k < kend) ? ld4g(g.B + (size_t)k * g.ldb + n, g.N - n, b_vec) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (g.ks && k < kend) { const float sc = g.ks[k]; v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc; }
                rb[j] = v;
            }
        }
    };
    // raw A stage / B registers -> tf32 hi + remainder operand tiles (same thread <-> element mapping as tc5)
    auto convert_tile = [&](int kt) {
        const char* ra = raw + (kt % S) * RAW_STAGE;
        char* a_hi = smem;
        char* a_lo = a_hi + A_TILE;
        char* b_hi = a_lo + A_TILE;
        char* b_lo = b_hi + B_TILE;
        if (!TA) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int m = (tid >> 3) + 32 * r, k4 = tid & 7;
                st_split(a_hi, a_lo, k4 * A_LBO + (m >> 3) * SBO + (m & 7) * 16, *(const float4*)(ra + (m * 8 + k4) * 16));
            }
        } else {
            const int k4 = tid >> 5, mb = (tid & 31) * 4;
            float4 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = *(const float4*)(ra + ((k4 * 4 + j) * 32 + (tid & 31)) * 16);
            const float t[4][4] = {{v[0].x, v[1].x, v[2].x, v[3].x}, {v[0].y, v[1].y, v[2].y, v[3].y},
                                   {v[0].z, v[1].z, v[2].z, v[3].z}, {v[0].w, v[1].w, v[2].w, v[3].w}};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int m = mb + e;
                st_split(a_hi, a_lo, k4 * A_LBO + (m >> 3) * SBO + (m & 7) * 16, make_float4(t[e][0], t[e][1], t[e][2], t[e][3]));
            }
        }
        if (TB) {
#pragma unroll
            for (int r = 0; r < BN / 32; ++r) {
                const int n = (tid >> 3) + 32 * r, k4 = tid & 7;
                st_split(b_hi, b_lo, k4 * B_LBO + (n >> 3) * SBO + (n & 7) * 16, rb[r]);
            }
        } else if (tid < 2 * BN) {
            const int k4 = tid / (BN / 4), nb = (tid % (BN / 4)) * 4;
            const float t[4][4] = {{rb[0].x, rb[1].x, rb[2].x, rb[3].x}, {rb[0].y, rb[1].y, rb[2].y, rb[3].y},
                                   {rb[0].z, rb[1].z, rb[2].z, rb[3].z}, {rb[0].w, rb[1].w, rb[2].w, rb[3].w}};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int n = nb + e;
                st_split(b_hi, b_lo, k4 * B_LBO + (n >> 3) * SBO + (n & 7) * 16, make_float4(t[e][0], t[e][1], t[e][2], t[e][3]));
            }
        }
    };

    // the ring is primed while warp 0 allocates tensor memory
#pragma unroll
    for (int s = 0; s < S - 1; ++s) {
        if (s < nk) issue_tile(s);
        cp_async_commit();
    }
    if (nk > 0) load_b(0);
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_d = *tmem_ptr;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<S - 2>();               // this thread's copies of tile kt have landed
        __syncthreads();                      // ... everyone's have; and everyone is done converting tile kt-1
        if (kt + S - 1 < nk) issue_tile(kt + S - 1);   // refill the slot tile kt-1 occupied
        cp_async_commit();
        if (kt >= 1) mbar_wait(smem_u32(&bars[0]), (kt - 1) & 1, &g_tc5_fail);   // MMAs of tile kt-1 have read the operand stage
        convert_tile(kt);
        if (kt + 1 < nk) load_b(kt + 1);      // in flight during the fence / barrier / MMA issue below
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint32_t a_hi = smem_u32(smem), a_lo = a_hi + A_TILE;
            const uint32_t b_hi = a_lo + A_TILE, b_lo = b_hi + B_TILE;
#pragma unroll
            for (int ks = 0; ks < BK / 8; ++ks) {
                const uint32_t ao = ks * 2 * A_LBO, bo = ks * 2 * B_LBO;
                const uint64_t dah = make_desc(a_hi + ao, A_LBO, SBO), dal = make_desc(a_lo + ao, A_LBO, SBO);
                const uint64_t dbh = make_desc(b_hi + bo, B_LBO, SBO), dbl = make_desc(b_lo + bo, B_LBO, SBO);
                mma_tf32(tmem_d, dal, dbh, idesc, (kt | ks) ? 1u : 0u);
                mma_tf32(tmem_d, dah, dbl, idesc, 1u);
                mma_tf32(tmem_d, dah, dbh, idesc, 1u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
                         :: "r"(smem_u32(&bars[0])) : "memory");
        }
    }
    cp_async_wait<0>();

    // ---- epilogue (identical to tc5): TMEM -> registers -> shared C tile -> coalesced global stores
    if (nk > 0) mbar_wait(smem_u32(&bars[0]), (nk - 1) & 1, &g_tc5_fail);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    constexpr int LDC_S = BN + 4;
    float* cs = (float*)smem;
    {
        const int r_loc = (warp & 3) * 32 + lane, col0 = (warp >> 2) * (BN / 2);
#pragma unroll
        for (int part = 0; part < BN / 32; ++part) {
            uint32_t v[16];
            const uint32_t taddr = tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(col0 + part * 16);
            if (nk > 0) {
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                               "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                             : "r"(taddr) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) v[e] = 0u;
            }
#pragma unroll
            for (int e = 0; e < 16; e += 4)
                *(uint4*)&cs[r_loc * LDC_S + col0 + part * 16 + e] = make_uint4(v[e], v[e + 1], v[e + 2], v[e + 3]);
        }
    }
    __syncthreads();
    {
        const bool atomic = gridDim.z > 1 && !g.partial;
        constexpr int TPR = BN / 4, RPP = NT / TPR;
        const int c4 = (tid % TPR) * 4, n = n0 + c4;
        const bool vec_ok = g.partial ? ((g.N & 3) == 0) : ((g.ldc & 3) == 0 && (((size_t)g.C) & 15) == 0);
#pragma unroll
        for (int it = 0; it < BM / RPP; ++it) {
            const int r_loc = tid / TPR + RPP * it, row = m0 + r_loc;
            if (row >= g.M || n >= g.N) continue;
            float4 x = *(const float4*)&cs[r_loc * LDC_S + c4];
            float xs[4] = {x.x, x.y, x.z, x.w};
            if (g.partial) {
                float* dst = g.partial + (size_t)blockIdx.z * g.M * g.N + (size_t)row * g.N + n;
                if (vec_ok && n + 3 < g.N) *(float4*)dst = x;
                else for (int e = 0; e < 4; ++e) if (n + e < g.N) dst[e] = xs[e];
                continue;
            }
            const float sc = g.rs ? g.rs[row] : 1.0f;
            float* dst = g.C + (size_t)row * g.ldc + n;
            if (atomic) {
                for (int e = 0; e < 4; ++e) if (n + e < g.N) atomicAdd(dst + e, xs[e] * sc);
                continue;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float y = xs[e] * sc;
                if (n + e < g.N) {
                    if (g.bias) y += g.bias[n + e];
                    if (g.bias2) y += g.bias2[n + e];
                    if (g.res) y += g.res[(size_t)row * g.ldr + n + e];
                }
                if (g.act) y = y > 0.f ? y : y * g.slope;
                xs[e] = y;
            }
            if (vec_ok && n + 3 < g.N) *(float4*)dst = make_float4(xs[0], xs[1], xs[2], xs[3]);
            else for (int e = 0; e < 4; ++e) if (n + e < g.N) dst[e] = xs[e];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem_d), "r"(TMEM_COLS) : "memory");
}


// ------------------------------------------------------------------------------------------------------------------
// tc7: warp-specialised version.  Round 1e measured 2.8 us per K tile per CTA for tc5 (L0 contraction, 2.1 CTAs/SM)
// and the cp.async ring (tc6) made it worse: the time is not global-load latency but the per-tile SERIALISATION of
//   MMAs done (mbarrier) -> all 256 threads convert -> __syncthreads -> thread 0 issues 12 MMAs + commit -> ...
// (ncu source page: 38 % of the stall samples sit at the block barrier, 13 % in the mbarrier wait).  Here
//   * warps 0-7 (converters) load global -> registers TWO K tiles ahead, split to tf32 hi / remainder and fill one of
//     two operand stages; each warp signals `full[s]` on its own (fence.proxy.async + __syncwarp + one mbarrier arrive),
//     so no warp ever waits for another converter;
//   * warp 8 (one lane) waits for `full[s]` (8 arrivals), issues the 12 tcgen05.mma of the tile and commits them to
//     `empty[s]`; converters only wait for `empty[s]` when they come back to that stage two tiles later;
//   * the epilogue (TMEM -> registers -> shared C tile -> global) is tc5's, run by the 8 converter warps.
constexpr int NT7 = 288, NS7 = 2;
template <int BN> struct Cfg7 {
    static constexpr int STAGE_BYTES = Cfg<BN>::STAGE_BYTES;
    static constexpr int SMEM_BYTES = NS7 * STAGE_BYTES + 128;    // 32: 93.3 KB, 64: 111.7 KB -> 2 CTAs/SM
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" :: "r"(bar) : "memory");
}

template <bool TA, bool TB, int BN>
__global__ void __launch_bounds__(NT7, 2)
tc7_gemm_kernel(D3fGemm g) {
    constexpr int B_LBO = Cfg<BN>::B_LBO, B_TILE = Cfg<BN>::B_TILE, STAGE_BYTES = Cfg<BN>::STAGE_BYTES;
    constexpr uint32_t TMEM_COLS = Cfg<BN>::TMEM_COLS;
    extern __shared__ __align__(128) char smem[];
    uint64_t* bars = (uint64_t*)(smem + NS7 * STAGE_BYTES);      // full[0], full[1], empty[0], empty[1]
    uint32_t* tmem_ptr = (uint32_t*)(smem + NS7 * STAGE_BYTES + 64);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * g.k_per_split, kend = min(g.K, kbeg + g.k_per_split);
    const int nk = (kend - kbeg + BK - 1) / BK;
    const bool a_vec = (g.lda & 3) == 0 && (((size_t)g.A) & 15) == 0;
    const bool b_vec = (g.ldb & 3) == 0 && (((size_t)g.B) & 15) == 0;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(tmem_ptr)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 32) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(&bars[0])), "r"(8) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(&bars[1])), "r"(8) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(&bars[2])), "r"(1) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(&bars[3])), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_d = *tmem_ptr;

    if (warp == 8) {
        // ---------------- MMA issuer: the whole warp walks the tiles (it must reach the block barriers below
        // converged), lane 0 issues
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        for (int kt = 0; kt < nk; ++kt) {
            const int s = kt & 1;
            mbar_wait(smem_u32(&bars[s]), (kt >> 1) & 1, &g_tc5_fail, 1LL << 18);          // all 8 converter warps filled stage s
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            if (lane == 0) {
                const uint32_t a_hi = smem_u32(smem) + s * STAGE_BYTES, a_lo = a_hi + A_TILE;
                const uint32_t b_hi = a_lo + A_TILE, b_lo = b_hi + B_TILE;
#pragma unroll
                for (int ks = 0; ks < BK / 8; ++ks) {
                    const uint32_t ao = ks * 2 * A_LBO, bo = ks * 2 * B_LBO;
                    const uint64_t dah = make_desc(a_hi + ao, A_LBO, SBO), dal = make_desc(a_lo + ao, A_LBO, SBO);
                    const uint64_t dbh = make_desc(b_hi + bo, B_LBO, SBO), dbl = make_desc(b_lo + bo, B_LBO, SBO);
                    mma_tf32(tmem_d, dal, dbh, idesc, (kt | ks) ? 1u : 0u);
                    mma_tf32(tmem_d, dah, dbl, idesc, 1u);
                    mma_tf32(tmem_d, dah, dbh, idesc, 1u);
                }
                // arrives on empty[s] once every MMA issued so far has completed (the stage may be refilled)
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
                             :: "r"(smem_u32(&bars[2 + s])) : "memory");
            }
            __syncwarp();
        }
    } else {
        // ---------------- converters: global -> registers (two tiles ahead) -> hi / lo operand stage
        float4 ra[2][4], rb[2][BN >= 128 ? BN / 32 : 4];
        auto load_tile = [&](int kt, float4 (&a)[4], float4 (&b)[BN >= 128 ? BN / 32 : 4]) {
            const int k0 = kbeg + kt * BK;
            if (!TA) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int m = m0 + (tid >> 3) + 32 * r, k = k0 + (tid & 7) * 4;
                    a[r] = (m < g.M) ? ld4g(g.A + (size_t)m * g.lda + k, kend - k, a_vec) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int k = k0 + (tid >> 5) * 4 + j, m = m0 + (tid & 31) * 4;
                    a[j] = (k < kend) ? ld4g(g.A + (size_t)k * g.lda + m, g.M - m, a_vec) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            if (TB) {
#pragma unroll
                for (int r = 0; r < BN / 32; ++r) {
                    const int n = n0 + (tid >> 3) + 32 * r, k = k0 + (tid & 7) * 4;
                    const size_t kb = g.bblk ? (size_t)(k / g.bblk) * g.bblk_stride + (k % g.bblk) : (size_t)k;
                    b[r] = (n < g.N) ? ld4g(g.B + (size_t)n * g.ldb + kb, kend - k, b_vec) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            } else if (tid < 2 * BN) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int k = k0 + (tid / (BN / 4)) * 4 + j, n = n0 + (tid % (BN / 4)) * 4;
                    float4 v = (k < kend) ? ld4g(g.B + (size_t)k * g.ldb + n, g.N - n, b_vec) : make_float4(0.f, 0.f, 0.f, 0.f);
                    if (g.ks && k < kend) { const float sc = g.ks[k]; v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc; }
                    b[j] = v;
                }
            }
        };
        auto store_tile = [&](int s, const float4 (&a)[4], const float4 (&b)[BN >= 128 ? BN / 32 : 4]) {
            char* a_hi = smem + s * STAGE_BYTES;
            char* a_lo = a_hi + A_TILE;
            char* b_hi = a_lo + A_TILE;
            char* b_lo = b_hi + B_TILE;
            if (!TA) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int m = (tid >> 3) + 32 * r, k4 = tid & 7;
                    st_split(a_hi, a_lo, k4 * A_LBO + (m >> 3) * SBO + (m & 7) * 16, a[r]);
                }
            } else {
                const int k4 = tid >> 5, mb = (tid & 31) * 4;
                const float t[4][4] = {{a[0].x, a[1].x, a[2].x, a[3].x}, {a[0].y, a[1].y, a[2].y, a[3].y},
                                       {a[0].z, a[1].z, a[2].z, a[3].z}, {a[0].w, a[1].w, a[2].w, a[3].w}};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int m = mb + e;
                    st_split(a_hi, a_lo, k4 * A_LBO + (m >> 3) * SBO + (m & 7) * 16, make_float4(t[e][0], t[e][1], t[e][2], t[e][3]));
                }
            }
            if (TB) {
#pragma unroll
                for (int r = 0; r < BN / 32; ++r) {
                    const int n = (tid >> 3) + 32 * r, k4 = tid & 7;
                    st_split(b_hi, b_lo, k4 * B_LBO + (n >> 3) * SBO + (n & 7) * 16, b[r]);
                }
            } else if (tid < 2 * BN) {
                const int k4 = tid / (BN / 4), nb = (tid % (BN / 4)) * 4;
                const float t[4][4] = {{b[0].x, b[1].x, b[2].x, b[3].x}, {b[0].y, b[1].y, b[2].y, b[3].y},
                                       {b[0].z, b[1].z, b[2].z, b[3].z}, {b[0].w, b[1].w, b[2].w, b[3].w}};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int n = nb + e;
                    st_split(b_hi, b_lo, k4 * B_LBO + (n >> 3) * SBO + (n & 7) * 16, make_float4(t[e][0], t[e][1], t[e][2], t[e][3]));
                }
            }
        };
        // one tile: wait until the MMAs that read stage s two tiles ago are done, fill it, prefetch tile kt + 2, signal
        auto step = [&](int kt, float4 (&a)[4], float4 (&b)[BN >= 128 ? BN / 32 : 4]) {
            const int s = kt & 1;
            if (kt >= NS7) mbar_wait(smem_u32(&bars[2 + s]), ((kt >> 1) - 1) & 1, &g_tc5_fail, 1LL << 18);
            store_tile(s, a, b);
            if (kt + 2 < nk) load_tile(kt + 2, a, b);
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");    // this thread's stores -> async proxy (UMMA)
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars[s]));
        };
        if (nk > 0) load_tile(0, ra[0], rb[0]);
        if (nk > 1) load_tile(1, ra[1], rb[1]);
        for (int kt = 0; kt < nk; kt += 2) {
            step(kt, ra[0], rb[0]);
            if (kt + 1 < nk) step(kt + 1, ra[1], rb[1]);
        }
        // every MMA has completed once the commit of the last tile has arrived
        if (nk > 0) mbar_wait(smem_u32(&bars[2 + ((nk - 1) & 1)]), ((nk - 1) >> 1) & 1, &g_tc5_fail, 1LL << 18);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    }

    // ---- epilogue (tc5's): TMEM -> registers -> shared C tile [128][BN+4] -> coalesced global stores; warp 8 only
    // takes part in the block barriers
    constexpr int LDC_S = BN + 4;
    float* cs = (float*)smem;
    if (warp < 8) {
        const int r_loc = (warp & 3) * 32 + lane, col0 = (warp >> 2) * (BN / 2);
#pragma unroll
        for (int part = 0; part < BN / 32; ++part) {
            uint32_t v[16];
            const uint32_t taddr = tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(col0 + part * 16);
            if (nk > 0) {
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                               "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                             : "r"(taddr) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) v[e] = 0u;
            }
#pragma unroll
            for (int e = 0; e < 16; e += 4)
                *(uint4*)&cs[r_loc * LDC_S + col0 + part * 16 + e] = make_uint4(v[e], v[e + 1], v[e + 2], v[e + 3]);
        }
    }
    __syncthreads();
    if (warp < 8) {
        const bool atomic = gridDim.z > 1 && !g.partial;
        constexpr int TPR = BN / 4, RPP = 256 / TPR;
        const int c4 = (tid % TPR) * 4, n = n0 + c4;
        const bool vec_ok = g.partial ? ((g.N & 3) == 0) : ((g.ldc & 3) == 0 && (((size_t)g.C) & 15) == 0);
#pragma unroll
        for (int it = 0; it < BM / RPP; ++it) {
            const int r_loc = tid / TPR + RPP * it, row = m0 + r_loc;
            if (row >= g.M || n >= g.N) continue;
            float4 x = *(const float4*)&cs[r_loc * LDC_S + c4];
            float xs[4] = {x.x, x.y, x.z, x.w};
            if (g.partial) {
                float* dst = g.partial + (size_t)blockIdx.z * g.M * g.N + (size_t)row * g.N + n;
                if (vec_ok && n + 3 < g.N) *(float4*)dst = x;
                else for (int e = 0; e < 4; ++e) if (n + e < g.N) dst[e] = xs[e];
                continue;
            }
            const float sc = g.rs ? g.rs[row] : 1.0f;
            float* dst = g.C + (size_t)row * g.ldc + n;
            if (atomic) {
                for (int e = 0; e < 4; ++e) if (n + e < g.N) atomicAdd(dst + e, xs[e] * sc);
                continue;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float y = xs[e] * sc;
                if (n + e < g.N) {
                    if (g.bias) y += g.bias[n + e];
                    if (g.bias2) y += g.bias2[n + e];
                    if (g.res) y += g.res[(size_t)row * g.ldr + n + e];
                }
                if (g.act) y = y > 0.f ? y : y * g.slope;
                xs[e] = y;
            }
            if (vec_ok && n + 3 < g.N) *(float4*)dst = make_float4(xs[0], xs[1], xs[2], xs[3]);
            else for (int e = 0; e < 4; ++e) if (n + e < g.N) dst[e] = xs[e];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem_d), "r"(TMEM_COLS) : "memory");
}


// ------------------------------------------------------------------------------------------------------------------
// tc8 (EXPERIMENTAL -- written at the end of round 1 from the clock64 breakdown in profiles/r1p_tc5_phase_timing.txt;
// it passed the 36 fp64-parity cases of tests/test_gpu_gemm.py on a B200 with the round's last GPU seconds
// (profiles/r1q_tc8_parity.log) but has NOT been timed yet; selectable with D3F_GEMM_PIPELINE=tmem only): the A operand
// lives in TENSOR MEMORY.
//
// tc5/tc6/tc7 all push A through shared memory twice over (hi + lo stores, then every MMA re-reads the 128-row tile:
// ~5x the tile's bytes through a 128 B/clk pipe, ~3700 cycles per K tile).  Here
//   * warps 0-3 own 32 rows each: coalesced 128-bit loads of the [32 x 32] fp32 slab (one K tile ahead, in registers),
//     a swizzled 4 KB shared-memory transpose so that a thread holds ITS row, hi = the raw fp32 bits (the tensor core
//     drops the low 13 bits), lo = v - trunc(v), and two tcgen05.st.32x32b.x32 into the stage's TMEM columns;
//     with A stored [K][M] (TA) a thread's row is already what coalesced loads give: no transpose;
//   * warps 4-7 split the small B tile into the usual K-major shared-memory stage;
//   * warp 8 issues tcgen05.mma with A from TMEM ([taddr]) and B from a shared-memory descriptor, two stages, the
//     same full[s] / empty[s] mbarrier protocol as tc7.
// TMEM columns (256 allocated, 2 CTAs/SM): accumulator [0, BN) | stage s: A_hi [128 + 64s, +32), A_lo [160 + 64s, +32).
constexpr int NT8 = 288;
template <int BN> struct Cfg8 {
    static constexpr int B_STAGE = 2 * Cfg<BN>::B_TILE;                 // B_hi | B_lo
    static constexpr int SLAB = 32 * 32 * 4;                            // one warp's transpose slab
    static constexpr int C_BYTES = BM * (BN + 4) * 4;
    static constexpr int WORK = 2 * B_STAGE + 4 * SLAB;
    static constexpr int SMEM_BYTES = (WORK > C_BYTES ? WORK : C_BYTES) + 128;
};

__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};\n"
        :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
           "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
           "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
           "r"(v[30]), "r"(v[31]) : "memory");
}

template <bool TA, bool TB, int BN>
__global__ void __launch_bounds__(NT8, 2)
tc8_gemm_kernel(D3fGemm g) {
    constexpr int B_LBO = Cfg<BN>::B_LBO, B_TILE = Cfg<BN>::B_TILE, B_STAGE = Cfg8<BN>::B_STAGE, SLAB = Cfg8<BN>::SLAB;
    constexpr uint32_t TMEM_COLS = 256, A_COL0 = 128;
    extern __shared__ __align__(128) char smem[];
    char* slabs = smem + 2 * B_STAGE;
    uint64_t* bars = (uint64_t*)(smem + Cfg8<BN>::SMEM_BYTES - 128);     // full[0], full[1], empty[0], empty[1]
    uint32_t* tmem_ptr = (uint32_t*)(smem + Cfg8<BN>::SMEM_BYTES - 64);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * g.k_per_split, kend = min(g.K, kbeg + g.k_per_split);
    const int nk = (kend - kbeg + BK - 1) / BK;
    const bool a_vec = (g.lda & 3) == 0 && (((size_t)g.A) & 15) == 0;
    const bool b_vec = (g.ldb & 3) == 0 && (((size_t)g.B) & 15) == 0;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(tmem_ptr)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 32) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(&bars[0])), "r"(8) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(&bars[1])), "r"(8) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(&bars[2])), "r"(1) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(&bars[3])), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_d = *tmem_ptr;

    if (warp == 8) {
        // ---------------- MMA issuer (whole warp walks the tiles, lane 0 issues)
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        for (int kt = 0; kt < nk; ++kt) {
            const int s = kt & 1;
            mbar_wait(smem_u32(&bars[s]), (kt >> 1) & 1, &g_tc5_fail, 1LL << 18);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            if (lane == 0) {
                const uint32_t b_hi = smem_u32(smem) + s * B_STAGE, b_lo = b_hi + B_TILE;
                const uint32_t a_hi = tmem_d + A_COL0 + 64 * s, a_lo = a_hi + 32;
#pragma unroll
                for (int ks = 0; ks < BK / 8; ++ks) {
                    const uint32_t bo = ks * 2 * B_LBO;
                    const uint64_t dbh = make_desc(b_hi + bo, B_LBO, SBO), dbl = make_desc(b_lo + bo, B_LBO, SBO);
                    mma_tf32_ts(tmem_d, a_lo + 8 * ks, dbh, idesc, (kt | ks) ? 1u : 0u);
                    mma_tf32_ts(tmem_d, a_hi + 8 * ks, dbl, idesc, 1u);
                    mma_tf32_ts(tmem_d, a_hi + 8 * ks, dbh, idesc, 1u);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
                             :: "r"(smem_u32(&bars[2 + s])) : "memory");
            }
            __syncwarp();
        }
    } else if (warp < 4) {
        // ---------------- A converters: rows 32*warp .. +31 of the tile -> TMEM
        char* slab = slabs + warp * SLAB;
        const int rbase = 32 * warp;
        float4 pre[8];                       // !TA: slab of the next tile as coalesced float4 (row 4j + lane/8, k4 = lane%8)
        float prt[TA ? 32 : 1];              //  TA: this thread's row of the next tile (k = 0..31), coalesced over lanes
        auto load_a = [&](int kt) {
            const int k0 = kbeg + kt * BK;
            if (!TA) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int m = m0 + rbase + 4 * j + (lane >> 3), k = k0 + (lane & 7) * 4;
                    pre[j] = (m < g.M) ? ld4g(g.A + (size_t)m * g.lda + k, kend - k, a_vec) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            } else {
                const int m = m0 + rbase + lane;
#pragma unroll
                for (int k = 0; k < 32; ++k)
                    prt[TA ? k : 0] = (m < g.M && k0 + k < kend) ? __ldg(g.A + (size_t)(k0 + k) * g.lda + m) : 0.f;
            }
        };
        if (nk > 0) load_a(0);
        for (int kt = 0; kt < nk; ++kt) {
            const int s = kt & 1;
            float row[32];
            if (!TA) {
                __syncwarp();                                   // the previous tile's row reads of the slab are done
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int r = 4 * j + (lane >> 3), q = lane & 7;
                    *(float4*)(slab + (r * 8 + (q ^ (r & 7))) * 16) = pre[j];
                }
                __syncwarp();
                if (kt + 1 < nk) load_a(kt + 1);
#pragma unroll
                for (int q = 0; q < 8; ++q) {                   // my row = lane; chunk q sits at position q ^ (lane & 7)
                    const float4 v = *(const float4*)(slab + (lane * 8 + (q ^ (lane & 7))) * 16);
                    row[4 * q] = v.x; row[4 * q + 1] = v.y; row[4 * q + 2] = v.z; row[4 * q + 3] = v.w;
                }
            } else {
#pragma unroll
                for (int k = 0; k < 32; ++k) row[k] = prt[TA ? k : 0];
                if (kt + 1 < nk) load_a(kt + 1);
            }
            uint32_t hi[32], lo[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                hi[k] = __float_as_uint(row[k]);                                        // tensor core drops the low 13 bits
                lo[k] = __float_as_uint(row[k] - __uint_as_float(hi[k] & 0xffffe000u)); // exact remainder
            }
            if (kt >= 2) mbar_wait(smem_u32(&bars[2 + s]), ((kt >> 1) - 1) & 1, &g_tc5_fail, 1LL << 18);   // stage s free
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint32_t ta = tmem_d + ((uint32_t)rbase << 16) + A_COL0 + 64 * s;
            tmem_st32(ta, hi);
            tmem_st32(ta + 32, lo);
            asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars[s]));
        }
        if (nk > 0) mbar_wait(smem_u32(&bars[2 + ((nk - 1) & 1)]), ((nk - 1) >> 1) & 1, &g_tc5_fail, 1LL << 18);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    } else {
        // ---------------- B converters (128 threads): global -> registers (one tile ahead) -> hi / lo shared-memory stage
        const int u = tid - 128;
        float4 rb[4];
        auto load_b = [&](int kt) {
            const int k0 = kbeg + kt * BK;
            if (TB) {       // B[n][k]: BN rows x 8 float4 = BN/16 per thread
#pragma unroll
                for (int r = 0; r < BN / 16; ++r) {
                    const int n = n0 + (u >> 3) + 16 * r, k = k0 + (u & 7) * 4;
                    const size_t kb = g.bblk ? (size_t)(k / g.bblk) * g.bblk_stride + (k % g.bblk) : (size_t)k;
                    rb[r] = (n < g.N) ? ld4g(g.B + (size_t)n * g.ldb + kb, kend - k, b_vec) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            } else if (u < 2 * BN) {   // B[k][n]: thread = (4 k rows, 4 consecutive n)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int k = k0 + (u / (BN / 4)) * 4 + j, n = n0 + (u % (BN / 4)) * 4;
                    float4 v = (k < kend) ? ld4g(g.B + (size_t)k * g.ldb + n, g.N - n, b_vec) : make_float4(0.f, 0.f, 0.f, 0.f);
                    if (g.ks && k < kend) { const float sc = g.ks[k]; v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc; }
                    rb[j] = v;
                }
            }
        };
        if (nk > 0) load_b(0);
        for (int kt = 0; kt < nk; ++kt) {
            const int s = kt & 1;
            if (kt >= 2) mbar_wait(smem_u32(&bars[2 + s]), ((kt >> 1) - 1) & 1, &g_tc5_fail, 1LL << 18);
            char* b_hi = smem + s * B_STAGE;
            char* b_lo = b_hi + B_TILE;
            if (TB) {
#pragma unroll
                for (int r = 0; r < BN / 16; ++r) {
                    const int n = (u >> 3) + 16 * r, k4 = u & 7;
                    st_split(b_hi, b_lo, k4 * B_LBO + (n >> 3) * SBO + (n & 7) * 16, rb[r]);
                }
            } else if (u < 2 * BN) {
                const int k4 = u / (BN / 4), nb = (u % (BN / 4)) * 4;
                const float t[4][4] = {{rb[0].x, rb[1].x, rb[2].x, rb[3].x}, {rb[0].y, rb[1].y, rb[2].y, rb[3].y},
                                       {rb[0].z, rb[1].z, rb[2].z, rb[3].z}, {rb[0].w, rb[1].w, rb[2].w, rb[3].w}};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int n = nb + e;
                    st_split(b_hi, b_lo, k4 * B_LBO + (n >> 3) * SBO + (n & 7) * 16, make_float4(t[e][0], t[e][1], t[e][2], t[e][3]));
                }
            }
            if (kt + 1 < nk) load_b(kt + 1);
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars[s]));
        }
        if (nk > 0) mbar_wait(smem_u32(&bars[2 + ((nk - 1) & 1)]), ((nk - 1) >> 1) & 1, &g_tc5_fail, 1LL << 18);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    }

    // ---- epilogue (tc5's): TMEM -> registers -> shared C tile [128][BN+4] -> coalesced global stores
    __syncthreads();                     // every warp is past its last shared-memory / TMEM use of the main loop
    constexpr int LDC_S = BN + 4;
    float* cs = (float*)smem;
    if (warp < 8) {
        const int r_loc = (warp & 3) * 32 + lane, col0 = (warp >> 2) * (BN / 2);
#pragma unroll
        for (int part = 0; part < BN / 32; ++part) {
            uint32_t v[16];
            const uint32_t taddr = tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(col0 + part * 16);
            if (nk > 0) {
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                               "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                             : "r"(taddr) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) v[e] = 0u;
            }
#pragma unroll
            for (int e = 0; e < 16; e += 4)
                *(uint4*)&cs[r_loc * LDC_S + col0 + part * 16 + e] = make_uint4(v[e], v[e + 1], v[e + 2], v[e + 3]);
        }
    }
    __syncthreads();
    if (warp < 8) {
        const bool atomic = gridDim.z > 1 && !g.partial;
        constexpr int TPR = BN / 4, RPP = 256 / TPR;
        const int c4 = (tid % TPR) * 4, n = n0 + c4;
        const bool vec_ok = g.partial ? ((g.N & 3) == 0) : ((g.ldc & 3) == 0 && (((size_t)g.C) & 15) == 0);
#pragma unroll
        for (int it = 0; it < BM / RPP; ++it) {
            const int r_loc = tid / TPR + RPP * it, row = m0 + r_loc;
            if (row >= g.M || n >= g.N) continue;
            float4 x = *(const float4*)&cs[r_loc * LDC_S + c4];
            float xs[4] = {x.x, x.y, x.z, x.w};
            if (g.partial) {
                float* dst = g.partial + (size_t)blockIdx.z * g.M * g.N + (size_t)row * g.N + n;
                if (vec_ok && n + 3 < g.N) *(float4*)dst = x;
                else for (int e = 0; e < 4; ++e) if (n + e < g.N) dst[e] = xs[e];
                continue;
            }
            const float sc = g.rs ? g.rs[row] : 1.0f;
            float* dst = g.C + (size_t)row * g.ldc + n;
            if (atomic) {
                for (int e = 0; e < 4; ++e) if (n + e < g.N) atomicAdd(dst + e, xs[e] * sc);
                continue;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float y = xs[e] * sc;
                if (n + e < g.N) {
                    if (g.bias) y += g.bias[n + e];
                    if (g.bias2) y += g.bias2[n + e];
                    if (g.res) y += g.res[(size_t)row * g.ldr + n + e];
                }
                if (g.act) y = y > 0.f ? y : y * g.slope;
                xs[e] = y;
            }
            if (vec_ok && n + 3 < g.N) *(float4*)dst = make_float4(xs[0], xs[1], xs[2], xs[3]);
            else for (int e = 0; e < 4; ++e) if (n + e < g.N) dst[e] = xs[e];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem_d), "r"(TMEM_COLS) : "memory");
}

}  // namespace

// launched by d3f_gemm_launch (gemm.cu) with the split decision already made
// tcgen05 kernel variant: 0 = tc5 (register-fed, one stage, 3-4 CTAs/SM), 1 = tc6 (A through a cp.async ring; needs a
// 16-byte aligned A), 2 = tc7 (warp-specialised, two operand stages).  Default from D3F_GEMM_PIPELINE = reg | cpasync | ws.
static int g_tc_pipeline = -1;
extern "C" void d3f_set_gemm_pipeline(int variant) { g_tc_pipeline = variant < 0 ? -1 : (variant > 3 ? 3 : variant); }
static int tc_pipeline() {
    if (g_tc_pipeline < 0) {
        const char* e = getenv("D3F_GEMM_PIPELINE");
        g_tc_pipeline = !e ? D3F_GEMM_PIPELINE_DEFAULT : (e[0] == 'r' ? 0 : (e[0] == 'c' ? 1 : (e[0] == 't' ? 3 : 2)));
    }
    return g_tc_pipeline;
}

template <bool TA, bool TB, int BN>
static int launch_bn(const D3fGemm& g, int splits, cudaStream_t stream) {
    dim3 grid(d3f_ceil_div(g.N, BN), d3f_ceil_div(g.M, BM), splits);
    const bool aligned = (g.lda & 3) == 0 && (((size_t)g.A) & 15) == 0;
    if (tc_pipeline() == 1 && aligned) {
        static bool attr6_set = false;
        if (!attr6_set) {
            D3F_CHECK_CUDA(cudaFuncSetAttribute(tc6_gemm_kernel<TA, TB, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                Cfg6<BN>::SMEM_BYTES));
            attr6_set = true;
        }
        tc6_gemm_kernel<TA, TB, BN><<<grid, NT, Cfg6<BN>::SMEM_BYTES, stream>>>(g);
        D3F_CHECK_LAUNCH();
        return D3F_OK;
    }
    if constexpr (BN <= 64) {
        if (tc_pipeline() == 3) {    // experimental: A operand in tensor memory
            static bool attr8_set = false;
            if (!attr8_set) {
                D3F_CHECK_CUDA(cudaFuncSetAttribute(tc8_gemm_kernel<TA, TB, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    Cfg8<BN>::SMEM_BYTES));
                attr8_set = true;
            }
            tc8_gemm_kernel<TA, TB, BN><<<grid, NT8, Cfg8<BN>::SMEM_BYTES, stream>>>(g);
            D3F_CHECK_LAUNCH();
            return D3F_OK;
        }
        if (tc_pipeline() == 2) {
            static bool attr7_set = false;
            if (!attr7_set) {
                D3F_CHECK_CUDA(cudaFuncSetAttribute(tc7_gemm_kernel<TA, TB, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    Cfg7<BN>::SMEM_BYTES));
                attr7_set = true;
            }
            tc7_gemm_kernel<TA, TB, BN><<<grid, NT7, Cfg7<BN>::SMEM_BYTES, stream>>>(g);
            D3F_CHECK_LAUNCH();
            return D3F_OK;
        }
    }
    static bool attr_set = false;
    if (!attr_set) {
        D3F_CHECK_CUDA(cudaFuncSetAttribute(tc5_gemm_kernel<TA, TB, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            Cfg<BN>::SMEM_BYTES));
        attr_set = true;
    }
    tc5_gemm_kernel<TA, TB, BN><<<grid, NT, Cfg<BN>::SMEM_BYTES, stream>>>(g);
    D3F_CHECK_LAUNCH();
    return D3F_OK;
}

template <bool TA, bool TB>
static int launch_mode(const D3fGemm& g, int splits, cudaStream_t stream) {
    if (g.N <= 32) return launch_bn<TA, TB, 32>(g, splits, stream);
    // wide outputs with enough row tiles to fill the chip: 128-wide tiles halve the A re-reads
    if (tc_pipeline() < 2 && g.N >= 256 && d3f_ceil_div(g.M, BM) * d3f_ceil_div(g.N, 128) * splits >= 148)
        return launch_bn<TA, TB, 128>(g, splits, stream);
    return launch_bn<TA, TB, 64>(g, splits, stream);
}

int d3f_gemm_tcgen05_launch(const D3fGemm& g, bool ta, bool tb, int splits, cudaStream_t stream) {
    if (ta && !tb) return launch_mode<true, false>(g, splits, stream);
    if (!ta && tb) return launch_mode<false, true>(g, splits, stream);
    if (!ta && !tb) return launch_mode<false, false>(g, splits, stream);
    d3f_set_error("gemm: TT mode is not used on the hot path");
    return D3F_ERR_UNSUPPORTED;
}

#ifdef D3F_TC5_TIMING
extern "C" int d3f_tc5_timing(unsigned long long* out32) {
    return cudaMemcpyFromSymbol(out32, g_tc5_t, sizeof(unsigned long long) * 32) == cudaSuccess ? 0 : -1;
}
#endif

// 1 if any tcgen05 GEMM gave up waiting on an mbarrier (diagnostic; reads a device symbol -> synchronises)
extern "C" int d3f_gemm_tcgen05_failed(void) {
    int v = 0;
    if (cudaMemcpyFromSymbol(&v, g_tc5_fail, sizeof(int)) != cudaSuccess) return -1;
    return v;
}
