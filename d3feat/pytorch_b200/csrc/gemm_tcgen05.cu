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

namespace {

constexpr int BM = 128, BK = 32, NT = 256;
constexpr int SBO = 144;                              // stride between 8-row groups (128 + 16 pad)
constexpr int A_LBO = (BM / 8) * SBO + 16;            // stride between 16-byte k units: 2320
constexpr int A_TILE = (BK / 4) * A_LBO;              // 18560
// the N tile is a template parameter (32 / 64 / 128): skinny outputs do not pay for a half-empty MMA and wide ones
// re-read A half as often
template <int BN> struct Cfg {
    // BN <= 64: the tf32-hi and remainder tiles of B are STACKED along N (rows 0..BN-1 = hi, BN..2BN-1 = lo of one
    // K-major tile), so that one MMA of width 2*BN yields a_hi*b_hi and a_hi*b_lo side by side in the accumulator and a
    // K step costs 2 tensor-core instructions instead of 3.  tools/umma_probe.cu measured >= 49 cycles per tcgen05.mma
    // whatever N <= 96: with 3-4 co-resident CTAs the skinny GEMMs were bound by the NUMBER of MMAs, not by their math.
    static constexpr bool STACK = BN <= 64;
    static constexpr int B_LBO = ((STACK ? 2 * BN : BN) / 8) * SBO + 16;
    static constexpr int B_TILE = (BK / 4) * B_LBO;               // STACK: holds hi and lo
    static constexpr int B_LO_OFF = STACK ? (BN / 8) * SBO : B_TILE;   // byte offset of the remainder rows / tile
    static constexpr int STAGE_BYTES = 2 * A_TILE + (STACK ? 1 : 2) * B_TILE;   // A hi + lo, B hi + lo
    // ONE shared-memory stage: these GEMMs are short (K = 32..960 for most of them) and latency-bound, so the
    // latency is hidden by co-resident CTAs (4 per SM at 46-56 KB) rather than by a deep pipeline inside one CTA
    // (ncu, round 1: 2 stages -> 1-2 CTAs/SM, 12-22 % warps active, every pipe under 27 %).
    static constexpr int SMEM_BYTES = STAGE_BYTES + 64;           // one stage + barrier / tmem pointer
    static constexpr uint32_t TMEM_COLS = STACK ? (2 * BN < 32 ? 32 : 2 * BN) : BN;
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

// Returns false if the barrier did not complete within max_spin probes: the caller never hangs the GPU, it flags the
// launch (g_tc5_fail, surfaced by d3f_gemm_status_snapshot) and poisons its output tile with NaN so that the failure
// cannot go unnoticed downstream (the optimiser's non-finite guard then skips the step).
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, int* fail, long long max_spin = (1LL << 26)) {
    uint32_t done = 0;
    for (long long spin = 0; spin < max_spin; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return true;
    }
    if (fail) atomicExch(fail, 1);
    return false;
}

// tf32 "hi" part.  cvt.rna.tf32 is emulated on sm_100a (FSETP + IADD + LOP3): round to nearest by hand (add half an ulp
// of the 10-bit mantissa, clear 13 bits; inputs are finite).  hi + (v - hi) == v exactly either way.
__device__ __forceinline__ float tf32_hi(float v) {
    return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
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
    constexpr int B_LBO = Cfg<BN>::B_LBO, B_LO_OFF = Cfg<BN>::B_LO_OFF, STAGE_BYTES = Cfg<BN>::STAGE_BYTES;
    constexpr bool STACK = Cfg<BN>::STACK;
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
    const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * BN) >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

    float4 ra[4], rb[BN >= 128 ? BN / 32 : 4];
    // Interior tiles (a full BK of K, 16-byte aligned operands, full M / N tile where the operand is transposed): every
    // thread's loads sit at a fixed offset from a base that advances by one tile per iteration, so the 64-bit address
    // and bounds arithmetic of the generic path below (~900 cycles per tile for the single warp that also issues the
    // MMAs) is done once.  The generic path still takes the last partial tile and unaligned / scaled operands.
    const bool fast = a_vec && b_vec && !g.ks && (!g.bblk || ((g.bblk % BK) == 0 && (g.bblk_stride & 3) == 0)) && (!TA || m0 + BM <= g.M) &&
                      (TB || n0 + BN <= g.N);
    const float* ap0;
    const float* bp0;
    size_t a_rs, b_rs, a_step, b_step;
    uint32_t amask = 0, bmask = 0;
    if (!TA) {
        ap0 = g.A + (size_t)(m0 + (tid >> 3)) * g.lda + kbeg + (tid & 7) * 4;
        a_rs = (size_t)32 * g.lda; a_step = BK;
#pragma unroll
        for (int r = 0; r < 4; ++r) amask |= (m0 + (tid >> 3) + 32 * r < g.M ? 1u : 0u) << r;
    } else {
        ap0 = g.A + (size_t)(kbeg + (tid >> 5) * 4) * g.lda + m0 + (tid & 31) * 4;
        a_rs = (size_t)g.lda; a_step = (size_t)BK * g.lda; amask = 0xFu;
    }
    if (TB) {
        bp0 = g.B + (size_t)(n0 + (tid >> 3)) * g.ldb + (tid & 7) * 4;      // + the tile's k offset (block-wise B: per tile)
        b_rs = (size_t)32 * g.ldb; b_step = BK;
#pragma unroll
        for (int r = 0; r < BN / 32; ++r) bmask |= (n0 + (tid >> 3) + 32 * r < g.N ? 1u : 0u) << r;
    } else {
        bp0 = g.B + (size_t)(kbeg + (tid / (BN / 4)) * 4) * g.ldb + n0 + (tid % (BN / 4)) * 4;
        b_rs = (size_t)g.ldb; b_step = (size_t)BK * g.ldb; bmask = tid < 2 * BN ? 0xFu : 0u;
    }
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto load_tile = [&](int k0) {
        if (fast && k0 + BK <= kend) {
            const size_t t = (size_t)((k0 - kbeg) / BK);
            const float* ap = ap0 + t * a_step;
            const float* bp;
            if (TB) bp = bp0 + (g.bblk ? (size_t)(k0 / g.bblk) * g.bblk_stride + (size_t)(k0 % g.bblk) : (size_t)k0);
            else bp = bp0 + t * b_step;
#pragma unroll
            for (int r = 0; r < 4; ++r) ra[r] = (amask >> r) & 1u ? __ldg((const float4*)(ap + r * a_rs)) : zero4;
#pragma unroll
            for (int r = 0; r < (TB ? BN / 32 : 4); ++r)
                rb[r] = (bmask >> r) & 1u ? __ldg((const float4*)(bp + r * b_rs)) : zero4;
            return;
        }
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
        char* b_lo = b_hi + B_LO_OFF;
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
    long long tacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64();
    const long long tstart = tlast;
#endif
    bool ok = true;
    if (nk > 0) load_tile(kbeg);
    TC5_T(6);                                                                      // prologue: alloc, barrier init, first loads issued
    for (int kt = 0; kt < nk; ++kt) {
        if (kt >= 1) ok &= mbar_wait(smem_u32(&bars[0]), (kt - 1) & 1, &g_tc5_fail);   // MMAs of tile kt-1 have read the stage
        TC5_T(0);
        store_tile();
        TC5_T(1);
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");              // generic-proxy stores -> async proxy (UMMA)
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        TC5_T(3);
        __syncthreads();
        TC5_T(4);
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint32_t a_hi = smem_u32(smem), a_lo = a_hi + A_TILE;
            const uint32_t b_hi = a_lo + A_TILE, b_lo = b_hi + B_LO_OFF;
#pragma unroll
            for (int ks = 0; ks < BK / 8; ++ks) {
                // one MMA = 8 k values = two 16-byte k units
                const uint32_t ao = ks * 2 * A_LBO, bo = ks * 2 * B_LBO;
                const uint64_t dah = make_desc(a_hi + ao, A_LBO, SBO), dal = make_desc(a_lo + ao, A_LBO, SBO);
                const uint64_t dbh = make_desc(b_hi + bo, B_LBO, SBO);
                if (STACK) {
                    // D[:, 0:BN] += a_hi b_hi + a_lo b_hi,  D[:, BN:2BN] += a_hi b_lo   (added in the epilogue)
                    mma_tf32(tmem_d, dah, dbh, idesc2, (kt | ks) ? 1u : 0u);   // B rows 0..2BN-1 = hi | lo
                    mma_tf32(tmem_d, dal, dbh, idesc, 1u);
                } else {
                    const uint64_t dbl = make_desc(b_lo + bo, B_LBO, SBO);
                    mma_tf32(tmem_d, dal, dbh, idesc, (kt | ks) ? 1u : 0u);
                    mma_tf32(tmem_d, dah, dbl, idesc, 1u);
                    mma_tf32(tmem_d, dah, dbh, idesc, 1u);
                }
            }
            // tcgen05.commit: arrive on the barrier when every MMA issued so far has completed
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
                         :: "r"(smem_u32(&bars[0])) : "memory");
        }
        TC5_T(5);
        // the next tile's global loads are issued AFTER the MMAs (they used to sit between the stores and the barrier,
        // delaying every MMA issue by their address arithmetic): they now overlap the tensor core's work on this tile
        if (kt + 1 < nk) load_tile(kbeg + (kt + 1) * BK);
        TC5_T(2);
    }

    // ---- epilogue
    if (nk > 0) ok &= mbar_wait(smem_u32(&bars[0]), (nk - 1) & 1, &g_tc5_fail);
    ok = __syncthreads_and(ok);          // one thread's timeout poisons the whole tile
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
                if (STACK) {        // the a_hi * b_lo half of the accumulator sits BN columns further
                    uint32_t u[16];
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                                   "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
                                 : "r"(taddr + (uint32_t)BN) : "memory");
                    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) + __uint_as_float(u[e]));
                } else {
                    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
                }
            } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) v[e] = 0u;
            }
#pragma unroll
            for (int e = 0; e < 16; e += 4)
                *(uint4*)&cs[r_loc * LDC_S + col0 + part * 16 + e] = make_uint4(v[e], v[e + 1], v[e + 2], v[e + 3]);
        }
    }
    TC5_T(10);                                                                     // TMEM -> registers -> shared C tile
    __syncthreads();
    TC5_T(11);
    {
        // Shared C tile -> global.  This code runs ONCE per CTA: ncu's source page showed the former fully unrolled,
        // every-mode-inlined version of it (2000 instructions) spending 47 % of a small GEMM's samples in
        // stall_no_inst -- instruction-cache misses on straight-line code (3.5 us of a 7.4 us CTA).  Hence: a rolled loop,
        // the column-invariant work (bias loads, bounds, address parts) hoisted, one compact body per store mode.
        const bool atomic = gridDim.z > 1 && !g.partial;
        constexpr int TPR = BN / 4, RPP = NT / TPR;      // threads per row, rows per pass
        const int c4 = (tid % TPR) * 4, n = n0 + c4;
        const int nv = min(4, g.N - n);                  // valid columns of this thread's quad (<= 0: none)
        const bool vec_ok = g.partial ? ((g.N & 3) == 0) : ((g.ldc & 3) == 0 && (((size_t)g.C) & 15) == 0);
        const float qnan = __int_as_float(0x7fc00000);
        if (nv > 0) {
            if (g.partial) {
                float* base = g.partial + (size_t)blockIdx.z * g.M * g.N + n;
#pragma unroll 1
                for (int r_loc = tid / TPR; r_loc < BM && m0 + r_loc < g.M; r_loc += RPP) {
                    float4 x = *(const float4*)&cs[r_loc * LDC_S + c4];
                    if (!ok) x = make_float4(qnan, qnan, qnan, qnan);
                    float* dst = base + (size_t)(m0 + r_loc) * g.N;
                    if (vec_ok && nv == 4) *(float4*)dst = x;
                    else { const float xs[4] = {x.x, x.y, x.z, x.w}; for (int e = 0; e < nv; ++e) dst[e] = xs[e]; }
                }
            } else if (g.ctrans) {     // transposed column blocks (see gemm.cuh): scalar, element stride ldc
#pragma unroll 1
                for (int r_loc = tid / TPR; r_loc < BM && m0 + r_loc < g.M; r_loc += RPP) {
                    const int row = m0 + r_loc;
                    float4 x = *(const float4*)&cs[r_loc * LDC_S + c4];
                    if (!ok) x = make_float4(qnan, qnan, qnan, qnan);
                    const float sc = g.rs ? g.rs[row] : 1.0f;
                    const float xs[4] = {x.x * sc, x.y * sc, x.z * sc, x.w * sc};
                    float* dt = g.C + (size_t)(row / g.cblk) * g.cblk_stride + (size_t)n * g.ldc + (row % g.cblk);
                    for (int e = 0; e < nv; ++e) {
                        if (atomic) atomicAdd(dt + (size_t)e * g.ldc, xs[e]);
                        else dt[(size_t)e * g.ldc] = xs[e];
                    }
                }
            } else {
                float* cbase = g.cblk ? g.C + (size_t)(n / g.cblk) * g.cblk_stride + (n % g.cblk) : g.C + n;
                float b1[4] = {0.f, 0.f, 0.f, 0.f}, b2[4] = {0.f, 0.f, 0.f, 0.f};
                if (!atomic) {
                    for (int e = 0; e < nv; ++e) {
                        if (g.bias) b1[e] = g.bias[n + e];
                        if (g.bias2) b2[e] = g.bias2[n + e];
                    }
                }
#pragma unroll 1
                for (int r_loc = tid / TPR; r_loc < BM && m0 + r_loc < g.M; r_loc += RPP) {
                    const int row = m0 + r_loc;
                    float4 x = *(const float4*)&cs[r_loc * LDC_S + c4];
                    if (!ok) x = make_float4(qnan, qnan, qnan, qnan);
                    const float sc = g.rs ? g.rs[row] : 1.0f;
                    float xs[4] = {x.x * sc, x.y * sc, x.z * sc, x.w * sc};
                    float* dst = cbase + (size_t)row * g.ldc;
                    if (atomic) {
                        for (int e = 0; e < nv; ++e) atomicAdd(dst + e, xs[e]);
                        continue;
                    }
                    float rr[4] = {0.f, 0.f, 0.f, 0.f};
                    if (g.res) {
                        const float* rp = g.res + (size_t)row * g.ldr + n;
                        if (nv == 4 && (g.ldr & 3) == 0 && (((size_t)g.res) & 15) == 0) {
                            const float4 t = *(const float4*)rp;
                            rr[0] = t.x; rr[1] = t.y; rr[2] = t.z; rr[3] = t.w;
                        } else {
                            for (int e = 0; e < nv; ++e) rr[e] = rp[e];
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float y = xs[e];
                        if (g.bias) y += b1[e];
                        if (g.bias2) y += b2[e];
                        if (g.res) y += rr[e];
                        if (g.act) y = y > 0.f ? y : y * g.slope;
                        xs[e] = y;
                    }
                    if (vec_ok && nv == 4) *(float4*)dst = make_float4(xs[0], xs[1], xs[2], xs[3]);
                    else for (int e = 0; e < nv; ++e) dst[e] = xs[e];
                }
            }
        }
    }
    TC5_T(7);                                                                      // epilogue
#ifdef D3F_TC5_TIMING
    if (timed) {
        unsigned long long* o = g_tc5_t + (tid == 0 ? 0 : 16);
        for (int i = 0; i < 8; ++i) o[i] = (unsigned long long)tacc[i];
        o[8] = (unsigned long long)(clock64() - tstart);
        o[9] = (unsigned long long)nk;
        o[10] = (unsigned long long)tacc[10];
        o[11] = (unsigned long long)tacc[11];
    }
#endif
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem_d), "r"(TMEM_COLS) : "memory");
}


}  // namespace

// launched by d3f_gemm_launch (gemm.cu) with the split decision already made.
// Round 2 decision (profiles/r2a_micro_kpconv_gemm.txt, B200): the register-fed one-stage kernel above beat the three
// alternatives that round 1 kept selectable -- A through a cp.async ring (42 vs 67 us on [40000x480]x[480x32]), the
// warp-specialised two-stage kernel (51 us) and the A-operand-in-TMEM kernel (59 us) -- on all 12 hot-path shapes, so
// those variants were deleted.  What replaces the largest contraction is not a faster stand-alone GEMM but the fused
// KPConv kernel (kpconv_fused.cu), which never materialises the A operand.
template <bool TA, bool TB, int BN>
static int launch_bn(const D3fGemm& g, int splits, cudaStream_t stream) {
    dim3 grid(d3f_ceil_div(g.N, BN), d3f_ceil_div(g.M, BM), splits);
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

int d3f_gemm_forced_bn();

template <bool TA, bool TB>
static int launch_mode(const D3fGemm& g, int splits, cudaStream_t stream) {
    const int forced = d3f_gemm_forced_bn();
    if (forced == 32) return launch_bn<TA, TB, 32>(g, splits, stream);
    if (forced == 64) return launch_bn<TA, TB, 64>(g, splits, stream);
    if (forced == 128) return launch_bn<TA, TB, 128>(g, splits, stream);
    if (g.N <= 32) return launch_bn<TA, TB, 32>(g, splits, stream);
    // 128-wide tiles halve the A re-reads but cost 3 MMAs per K step where the stacked 64-wide tile costs 2: they win
    // only on the largest problems (profiles/r2_gemm_tune.txt: >= 2.2 GFLOP won by 5-12 %, <= 0.54 GFLOP lost by 10-30 %)
    if (g.N >= 256 && (double)g.M * g.N * g.K >= 1.0e9 && d3f_ceil_div(g.M, BM) * d3f_ceil_div(g.N, 128) * splits >= 148)
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

// Asynchronous, capturable form: out[0] = 1 if any tcgen05 GEMM of this process gave up waiting on an mbarrier (its
// output tile was poisoned with NaN), copied device-to-device on `stream` so that a sync-free pipeline can fold it into
// its status vector (engine.PairStep).
// device address of the library-wide "a tensor-core kernel timed out on an mbarrier" flag (also raised by kpconv_fused.cu)
int* d3f_fail_flag_device() {
    static int* addr = nullptr;
    if (!addr && cudaGetSymbolAddress((void**)&addr, g_tc5_fail) != cudaSuccess) addr = nullptr;
    return addr;
}

extern "C" int d3f_gemm_status_snapshot(int32_t* out, d3f_stream stream) {
    int* addr = d3f_fail_flag_device();
    D3F_REQUIRE(addr, D3F_ERR_CUDA, "cudaGetSymbolAddress failed");
    D3F_REQUIRE(out, D3F_ERR_INVALID, "null pointer");
    D3F_CHECK_CUDA(cudaMemcpyAsync(out, addr, sizeof(int), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return D3F_OK;
}

// 1 if any tcgen05 GEMM gave up waiting on an mbarrier (diagnostic; reads a device symbol -> synchronises)
extern "C" int d3f_gemm_tcgen05_failed(void) {
    int v = 0;
    if (cudaMemcpyFromSymbol(&v, g_tc5_fail, sizeof(int)) != cudaSuccess) return -1;
    return v;
}
