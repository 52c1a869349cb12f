// Fused KPConv forward for the rigid 32 -> 32 layers of pyramid level 0 (models/blocks.py:277-380): neighbour gather +
// kernel-point correlation + the [K*Cin x Cout] contraction + density normalisation + bias + LeakyReLU in ONE kernel.
// The kernel-point-weighted features wf [Nq, K*Cin] (77 MB at level 0; written and re-read by the two-kernel path) never
// leave the SM unless the caller asks for them.
//
// One persistent CTA per SM, 19 warps, each CTA owns a contiguous range of 16-query rounds:
//
//   warps 0-15  "gather": one query per round, software-pipelined ACROSS queries so that no global load is ever waited
//     for: while query j is being multiplied, the index row of query j+2, the support points / flags of query j+1 and
//     (group by group, into the registers query j has just freed) the feature rows of query j+1 are in flight.
//     A  lanes over neighbours: index -> support point -> record (-2r, |r|^2) in shared memory (r = s - q), density
//        count from the per-support "positive row" flags (blocks.py:377).
//     B  influences on the legacy tensor path: sq[kp, h] = |r_h|^2 - 2 r_h.k + |k|^2 is ONE 16x8x8 product per 8
//        neighbours, [kp | (k, |k|^2, 1)] x [(-2r, 1, |r|^2) | h], 3xTF32; its C fragment (kernel point x neighbour) becomes,
//        after w = max(0, 1 - sqrt(sq)/extent), exactly the A fragment of the correlation product, so the influences
//        never touch shared memory.
//     C  correlation wf[16 kp x 32 ch] += w[16 x 8] * X[8 x 32] per 8 neighbours (mma.sync m16n8k8, 3xTF32); the 8
//        neighbour rows are read as two 128-bit loads per lane (a row = 128 contiguous bytes = one line).
//     D  the warp's wf row (480 floats) is split into THREE bf16 terms (b1 + b2 + b3 = the fp32 value to 2^-24) and stored
//        into the B operand tile of the contraction (UMMA K-major, no swizzle: rows 0-31 = b1 of the 32 queries of a
//        super-batch, rows 32-63 = b2, rows 64-95 = b3); a lane's 8 consecutive channels are exactly one 16-byte K unit.
//   warps 16-18 "contraction": W^T lives in TENSOR MEMORY for the whole kernel, also as three bf16 terms -- TMEM lanes
//        0-31 / 32-63 / 64-95 = b1 / b2 / b3 of W^T [Cout, K*Cin], two bf16 per column, written once with tcgen05.st from
//        coalesced loads.  Per super-batch of 32 queries ONE tcgen05.mma (kind::f16, M 128 x N 96 x K 16) per 16-wide K
//        step, 30 for K = 15: D[128 x 96] += A_tmem * B_smem holds all NINE partial products bi(W) x bj(wf) in its
//        3 x 3 blocks of 32 x 32, which the epilogue adds (fp32): full fp32 accuracy from bf16 tensor-core products.
//        tools/umma_probe.cu measured what shapes this: a tcgen05.mma costs >= 49 cycles whatever N <= 96 (then
//        128 N / 256), so the contraction wants FEW, WIDE instructions -- 30 per 32 queries here against 180 per 16
//        queries for a 3xTF32 formulation -- and two bf16 sit in one TMEM column with the even k in the low half.
//        Epilogue: 1/n, bias, LeakyReLU, [32 x 32] output tile -> global memory with one bulk (TMA) store.
//
// Synchronisation: mbarriers `full[2]` (32 arrivals: the B tile of a super-batch is written) and `done[2]`
// (tcgen05.commit: the tile may be overwritten, D may be read); the B tile is double-buffered so gather warps run up to
// one super-batch ahead of the tensor core.
#include "common.cuh"
#include "kpconv.cuh"
#include <limits.h>

namespace {

constexpr int FQ = 16;                    // gather warps = queries per round
constexpr int SBQ = 2 * FQ;               // queries per super-batch (two rounds)
constexpr int NT = (FQ + 3) * 32;         // + 3 contraction warps (TMEM lanes 0-31, 32-63, 64-95)
constexpr int CIN = 32, COUT = 32;
constexpr int BN = 3 * SBQ;               // rows of the B tile / columns of D
constexpr int B_SBO = 128;                // 8-row group stride (dense core matrices: 8 rows x 16 bytes)
constexpr int B_LBO = (BN / 8) * 128 + 16;   // stride of a 16-byte K unit (8 bf16): 96 rows + 16 bytes of bank spread
constexpr uint32_t TF32_MASK = 0xffffe000u, BF16_MASK = 0xffff0000u;
constexpr unsigned FULL = 0xffffffffu;
constexpr uint32_t TMEM_COLS = 512, D_COL = 256;

struct FusedArgs {
    const float* q; const float* s; const void* inds; long long ld; const float* x; const unsigned char* rowpos;
    const float* kp; const float* w; const float* bias; float* out; float* inv_n; float* wf;
    int nq, ns, H, K; float extent; int act; float slope; int* fail;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float fast_sqrt(float v) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

// v = hi + lo with hi exactly representable in tf32 (low 13 mantissa bits clear) and lo the exact remainder
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(v) & TF32_MASK;
    lo = __float_as_uint(v - __uint_as_float(hi));
}

// v = b1 + b2 + b3 (+ O(2^-24 v)), every term exactly representable in bf16 (returned as fp32 bit patterns, low 16 bits 0)
__device__ __forceinline__ void split_bf16x3(float v, uint32_t& b1, uint32_t& b2, uint32_t& b3) {
    b1 = __float_as_uint(v) & BF16_MASK;
    const float r1 = v - __uint_as_float(b1);
    b2 = __float_as_uint(r1) & BF16_MASK;
    const float r2 = r1 - __uint_as_float(b2);
    b3 = __float_as_uint(r2) & BF16_MASK;
}
// two bf16 (given as fp32 bit patterns) in one 32-bit word, `even` in the low half (measured: tools/umma_probe.cu)
__device__ __forceinline__ uint32_t pack_bf16(uint32_t even, uint32_t odd) { return __byte_perm(even, odd, 0x7632); }

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// Returns false if the phase never completes: the caller never hangs the GPU, it raises the library's failure flag and
// poisons its output.  (A suspend-time hint on try_wait compiles to a fixed NANOSLEEP of the hint after every failed
// probe -- measured: 20 us stalls -- so this is a plain probe loop; `backoff` adds a short sleep for the warps whose
// wait is long by design.)
template <bool BACKOFF = false>
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
#pragma unroll 1
    for (int spin = 0; spin < (1 << 24); ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return true;
        if (BACKOFF) __nanosleep(64);
    }
    return false;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" :: "r"(bar) : "memory");
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // UMMA shared-memory descriptor (K-major, no swizzle): start[0,14) | LBO[16,30) | SBO[32,46) | version 1 [46,48)
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ULL << 46);
}

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
        :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
           "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
          "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
          "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// shared memory (bytes): B tiles [2] (16 kernel-point slots each: the 16th takes the lanes that own no kernel point, so the
// stores need no predicate; the MMAs read K of them) | rec [FQ][HP] float4 | idx [2][FQ][HP] i32 | inv_n [4][SBQ] f32 |
// xch [2][32][SBQ+1] f32 | out tile [SBQ][COUT] f32 | barriers
constexpr int NG = 6, HP = NG * 8, NH = 2;      // neighbour columns are padded to 48 = 6 groups of 8 = 2 passes of 32 lanes
constexpr int RING = 3;                         // feature-row loads kept in flight per lane: 3 groups (24 registers)
struct FusedSmem {
    int b_bytes, rec_off, idx_off, invn_off, xch_off, out_off, bar_off, total;
};
__host__ __device__ inline FusedSmem fused_smem() {
    FusedSmem m;
    m.b_bytes = (16 * CIN / 8) * B_LBO;
    m.rec_off = (2 * m.b_bytes + 127) & ~127;
    m.idx_off = m.rec_off + FQ * HP * 16;
    m.invn_off = m.idx_off + 2 * FQ * HP * 4;
    m.xch_off = m.invn_off + 4 * SBQ * 4;
    m.out_off = (m.xch_off + 2 * 32 * (SBQ + 1) * 4 + 127) & ~127;
    m.bar_off = m.out_off + SBQ * COUT * 4;
    m.total = m.bar_off + 64;                 // 6 mbarriers + the tensor-memory base address
    return m;
}

template <bool IDX64>
__global__ void __launch_bounds__(NT, 1)
kpf_fused_kernel(FusedArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const FusedSmem L = fused_smem();
    float4* rec_all = (float4*)(smem + L.rec_off);
    int* idx_all = (int*)(smem + L.idx_off);
    float* invn_s = (float*)(smem + L.invn_off);
    float* xch = (float*)(smem + L.xch_off);
    float* out_s = (float*)(smem + L.out_off);
    uint64_t* bars = (uint64_t*)(smem + L.bar_off);          // [0,1] full, [2,3] done, [4] W landed, [5] W staging free
    uint32_t* tmem_ptr = (uint32_t*)(smem + L.bar_off + 48);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // W [K*Cin, Cout] fp32 is staged ONCE per CTA by one bulk (TMA) copy into the second B tile, which the gather warps
    // do not touch before their third round
    float* w_stage = (float*)(smem + L.b_bytes);
    const uint32_t w_bytes = (uint32_t)(a.K * CIN * COUT * 4);
    // this CTA's contiguous range of 16-query rounds
    const int n_rounds = (a.nq + FQ - 1) / FQ;
    const int r_base = n_rounds / (int)gridDim.x, r_rem = n_rounds % (int)gridDim.x;
    const int my_rounds = r_base + ((int)blockIdx.x < r_rem ? 1 : 0);
    const int r_start = (int)blockIdx.x * r_base + min((int)blockIdx.x, r_rem);
    const int my_sb = (my_rounds + 1) >> 1;

    if (warp == FQ) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(tmem_ptr)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(&bars[0])), "r"(SBQ) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(&bars[1])), "r"(SBQ) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(&bars[2])), "r"(1) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(&bars[3])), "r"(1) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(&bars[4])), "r"(1) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(&bars[5])), "r"(3) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        // global -> shared, completion counted in bytes on bars[4]
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(smem_u32(&bars[4])), "r"(w_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                     :: "r"(smem_u32(w_stage)), "l"(a.w), "r"(w_bytes), "r"(smem_u32(&bars[4])) : "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = *tmem_ptr;
    const uint32_t bar_full = smem_u32(&bars[0]), bar_done = smem_u32(&bars[2]);     // + 8 * buffer
    const uint32_t bar_w = smem_u32(&bars[4]), bar_wfree = smem_u32(&bars[5]);

    if (warp < FQ) {
        // =========================================================================================== gather warps
        float4* rec = rec_all + (size_t)warp * HP;
        int* idx_buf = idx_all + (size_t)warp * HP;                 // + FQ * HP for the other parity
        const int gq = lane >> 2, tq = lane & 3;
        // A fragment of the distance product: row kp = (kx, ky, kz, |k|^2, 1, 0, 0, 0); rows >= K are zero
        uint32_t kh[4], kl[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = gq + 8 * (i & 1), c = tq + 4 * (i >> 1);     // a0 (gq,tq) a1 (gq+8,tq) a2 (gq,tq+4) a3 (gq+8,tq+4)
            float v = 0.f;
            if (k < a.K) {
                const float kx = a.kp[3 * k], ky = a.kp[3 * k + 1], kz = a.kp[3 * k + 2];
                v = c == 0 ? kx : c == 1 ? ky : c == 2 ? kz : c == 3 ? (kx * kx + ky * ky + kz * kz) : c == 4 ? 1.0f : 0.f;
            }
            split_tf32(v, kh[i], kl[i]);
        }
        const float inv_ext = 1.0f / a.extent;
        const float* __restrict__ xg = a.x + 4 * gq;
        const int nq = a.nq, H = a.H, ns = a.ns;
        const long long ld = a.ld;
        const int q_first = r_start * FQ + warp;                    // this warp's queries: q_first + 16 j
        bool ok = true;

        // ---- pipeline stages (see the header).  S0: index row -> registers, RAW (validated one iteration later, when it is
        // consumed: checking it here would wait for the load that was just issued); the low word is kept for 64-bit
        // indices together with a "high word is zero" flag folded into bit 31
        auto load_idx = [&](int j, int (&id)[NH]) {
            const int qi = q_first + FQ * j;
            const bool live = j < my_rounds && qi < nq;
#pragma unroll
            for (int t = 0; t < NH; ++t) {
                const int h = 32 * t + lane;
                int v = -1;
                if (live && h < H) {
                    if (IDX64) {
                        const int2 w = __ldg((const int2*)a.inds + (size_t)qi * ld + h);
                        v = w.x | (w.y ? 0x80000000 : 0);             // any high bits -> negative -> shadow
                    } else {
                        v = __ldg((const int*)a.inds + (size_t)qi * ld + h);
                    }
                }
                id[t] = v;
            }
        };
        // S1a: support points / flags of the indexed neighbours -> registers; index row -> shared memory; group count
        auto load_pts = [&](int j, int (&id)[NH], float (&sx)[NH], float (&sy)[NH], float (&sz)[NH], int (&rp)[NH],
                            float (&qv)[3], int* idx_s) -> int {
            const int qi = q_first + FQ * j;
            int hend = 0;
#pragma unroll
            for (int t = 0; t < NH; ++t) {
                const int h = 32 * t + lane;
                if ((unsigned)id[t] >= (unsigned)ns) id[t] = -1;    // shadow index (= Ns), padding, garbage: no neighbour
                const int i = max(id[t], 0);                        // shadow rows read point 0 and are masked below
                const float* sp = a.s + 3 * (size_t)i;
                sx[t] = __ldg(sp); sy[t] = __ldg(sp + 1); sz[t] = __ldg(sp + 2);
                rp[t] = __ldg(a.rowpos + i);
                if (h < HP) idx_s[h] = i;
                const unsigned bv = __ballot_sync(FULL, id[t] >= 0);
                if (bv) hend = 32 * t + 32 - __clz(bv);
            }
            const float* qp = a.q + 3 * (size_t)min(qi, nq - 1);
            qv[0] = __ldg(qp); qv[1] = __ldg(qp + 1); qv[2] = __ldg(qp + 2);
            return (hend + 7) >> 3;
        };
        // S1b: records (-2r, |r|^2) -> shared memory, density count
        auto store_rec = [&](const int (&id)[NH], const float (&sx)[NH], const float (&sy)[NH], const float (&sz)[NH],
                             const int (&rp)[NH], const float (&qv)[3]) -> int {
            int count = 0;
#pragma unroll
            for (int t = 0; t < NH; ++t) {
                const int h = 32 * t + lane;
                const bool valid = id[t] >= 0;
                const float rx = sx[t] - qv[0], ry = sy[t] - qv[1], rz = sz[t] - qv[2];
                float4 r = make_float4(-2.0f * rx, -2.0f * ry, -2.0f * rz, rx * rx + ry * ry + rz * rz);
                if (!valid) r = make_float4(0.f, 0.f, 0.f, 1e30f);             // shadow / padding: w = 0
                if (h < HP) rec[h] = r;
                count += __popc(__ballot_sync(FULL, valid && rp[t] != 0));
            }
            return count;
        };
        auto load_x = [&](const int* idx_s, int g, float4& xa, float4& xb) {
            const int2 id = *(const int2*)&idx_s[8 * g + 2 * tq];              // neighbours 2tq, 2tq+1 of the group
            xa = __ldg((const float4*)(xg + (unsigned)id.x * (unsigned)CIN));
            xb = __ldg((const float4*)(xg + (unsigned)id.y * (unsigned)CIN));
        };

        // ---- prologue: query 0 completely staged, its first RING groups of rows and the index row of query 1 in flight
        int nid[NH], nid2[NH];
        float sx[NH], sy[NH], sz[NH], qv[3];
        int rp[NH];
        float4 xa[RING], xb[RING];
        load_idx(0, nid);
        int ng = load_pts(0, nid, sx, sy, sz, rp, qv, idx_buf);
        int count = store_rec(nid, sx, sy, sz, rp, qv);
        __syncwarp();
#pragma unroll
        for (int g = 0; g < RING; ++g)
            if (g < ng) load_x(idx_buf, g, xa[g], xb[g]);
        load_idx(1, nid);

        for (int j = 0; j < my_rounds; ++j) {
            const int qi = q_first + FQ * j;
            const bool qvalid = qi < nq;
            const int* idx_cur = idx_buf + (j & 1) * (FQ * HP);
            int* idx_nxt = idx_buf + ((j + 1) & 1) * (FQ * HP);
            // ---- stage the next queries: points of j+1 (its index row arrived during query j-1), index row of j+2
            __syncwarp();                                   // idx_nxt held query j-1's row: every lane is done with it
            const int ng_next = load_pts(j + 1, nid, sx, sy, sz, rp, qv, idx_nxt);
            load_idx(j + 2, nid2);
            __syncwarp();                                   // idx_nxt visible to the feature-row loads below

            float acc[4][4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[e][i] = 0.f;
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                if (g < ng) {
                    // ---- B: sq[kp, neighbour] for the group's 8 neighbours, then the influences
                    const float rv = ((const float*)rec)[(8 * g + gq) * 4 + tq];      // B[k = tq][n = gq], k = 3 is the constant 1
                    const float r2 = ((const float*)rec)[(8 * g + gq) * 4 + 3];       // |r|^2 of neighbour gq (broadcast read)
                    uint32_t bh[2], bl[2];
                    split_tf32(tq == 3 ? 1.0f : rv, bh[0], bl[0]);
                    split_tf32(tq == 0 ? r2 : 0.f, bh[1], bl[1]);                     // B[k = tq + 4][n = gq]
                    float sq[4] = {0.f, 0.f, 0.f, 0.f};
                    mma_tf32(sq, kl, bh);
                    mma_tf32(sq, kh, bl);
                    mma_tf32(sq, kh, bh);
                    // C fragment: c0 (kp gq, nb 2tq) c1 (gq, 2tq+1) c2 (gq+8, 2tq) c3 (gq+8, 2tq+1)
                    float w[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) w[i] = fmaxf(fmaf(-fast_sqrt(fmaxf(sq[i], 0.f)), inv_ext, 1.0f), 0.f);
                    // A fragment of the correlation: a0 (kp gq, col tq) a1 (gq+8, tq) a2 (gq, tq+4) a3 (gq+8, tq+4) with
                    // column tq <-> neighbour 2tq and column tq+4 <-> neighbour 2tq+1 of the group
                    uint32_t ah[4], al[4];
                    split_tf32(w[0], ah[0], al[0]);
                    split_tf32(w[2], ah[1], al[1]);
                    split_tf32(w[1], ah[2], al[2]);
                    split_tf32(w[3], ah[3], al[3]);
                    const float va[4] = {xa[g % RING].x, xa[g % RING].y, xa[g % RING].z, xa[g % RING].w};
                    const float vb[4] = {xb[g % RING].x, xb[g % RING].y, xb[g % RING].z, xb[g % RING].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {          // n-tile e holds channels 4n + e, n = 0..7
                        uint32_t xh[2], xl[2];
                        split_tf32(va[e], xh[0], xl[0]);   // B[k = tq][n = gq]      = x[nb 2tq][4gq + e]
                        split_tf32(vb[e], xh[1], xl[1]);   // B[k = tq + 4][n = gq]  = x[nb 2tq+1][4gq + e]
                        mma_tf32(acc[e], al, xh);
                        mma_tf32(acc[e], ah, xl);
                        mma_tf32(acc[e], ah, xh);
                    }
                }
                // the freed ring slot takes the rows RING groups ahead: of this query, or of the next one
                if (g + RING < NG) {
                    if (g + RING < ng) load_x(idx_cur, g + RING, xa[g % RING], xb[g % RING]);
                } else {
                    if (g + RING - NG < ng_next) load_x(idx_nxt, g + RING - NG, xa[g % RING], xb[g % RING]);
                }
            }
            // ---- records of query j+1 (its support points have arrived by now)
            __syncwarp();                                   // all lanes are done reading rec (query j)
            const int count_next = store_rec(nid, sx, sy, sz, rp, qv);

            // acc[e][i]: i = 0 (kp gq, col 2tq) 1 (gq, 2tq+1) 2 (gq+8, 2tq) 3 (gq+8, 2tq+1); channel = 4 col + e
            const float invn = 1.0f / (float)max(count, 1);
            if (lane == 0 && qvalid) a.inv_n[qi] = invn;
            if (a.wf && qvalid) {
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int k = gq + 8 * half;
                    if (k < a.K) {
                        float* o = a.wf + ((size_t)qi * a.K + k) * CIN + 8 * tq;
                        *(float4*)o = make_float4(acc[0][2 * half], acc[1][2 * half], acc[2][2 * half], acc[3][2 * half]);
                        *(float4*)(o + 4) = make_float4(acc[0][2 * half + 1], acc[1][2 * half + 1], acc[2][2 * half + 1], acc[3][2 * half + 1]);
                    }
                }
            }
            // ---- D: this query's three rows of the B tile
            const int sb = j >> 1, slot = (j & 1) * FQ + warp;
            if ((j & 1) == 0 && sb >= 2) ok &= mbar_wait(bar_done + 8 * (sb & 1), ((sb >> 1) - 1) & 1);   // MMAs of sb-2 have read it
            if (j == 2) ok &= mbar_wait(bar_wfree, 0);                      // the second tile was the staging area of W
            // (those MMAs were issued after the epilogue of super-batch sb-3, and sb-4 used this inv_n slot: safe to overwrite)
            if (lane == 0) invn_s[(sb & 3) * SBQ + slot] = invn;
            {
                unsigned char* tile = smem + (size_t)(sb & 1) * L.b_bytes + (slot >> 3) * B_SBO + (slot & 7) * 16;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    // channels 8tq .. 8tq+7 of kernel point k = gq + 8 half = ONE 16-byte K unit (kc = 32 k + 8 tq); v = b1 + b2 +
                    // b3 with bf16 terms: packing the HIGH halves of two fp32 words truncates them to bf16
                    float v[8], r[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) v[c] = acc[c & 3][2 * half + (c >> 2)];
                    unsigned char* p = tile + (size_t)((gq + 8 * half) * (CIN / 8) + tq) * B_LBO;
                    *(uint4*)p = make_uint4(pack_bf16(__float_as_uint(v[0]), __float_as_uint(v[1])), pack_bf16(__float_as_uint(v[2]), __float_as_uint(v[3])),
                                            pack_bf16(__float_as_uint(v[4]), __float_as_uint(v[5])), pack_bf16(__float_as_uint(v[6]), __float_as_uint(v[7])));
#pragma unroll
                    for (int c = 0; c < 8; ++c) r[c] = v[c] - __uint_as_float(__float_as_uint(v[c]) & BF16_MASK);
                    *(uint4*)(p + 4 * B_SBO) = make_uint4(pack_bf16(__float_as_uint(r[0]), __float_as_uint(r[1])), pack_bf16(__float_as_uint(r[2]), __float_as_uint(r[3])),
                                                          pack_bf16(__float_as_uint(r[4]), __float_as_uint(r[5])), pack_bf16(__float_as_uint(r[6]), __float_as_uint(r[7])));
#pragma unroll
                    for (int c = 0; c < 8; ++c) r[c] = r[c] - __uint_as_float(__float_as_uint(r[c]) & BF16_MASK);
                    *(uint4*)(p + 8 * B_SBO) = make_uint4(pack_bf16(__float_as_uint(r[0]), __float_as_uint(r[1])), pack_bf16(__float_as_uint(r[2]), __float_as_uint(r[3])),
                                                          pack_bf16(__float_as_uint(r[4]), __float_as_uint(r[5])), pack_bf16(__float_as_uint(r[6]), __float_as_uint(r[7])));
                }
            }
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");      // generic-proxy stores -> async proxy (UMMA)
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full + 8 * (sb & 1));
            // rotate the pipeline registers
#pragma unroll
            for (int t = 0; t < NH; ++t) nid[t] = nid2[t];
            ng = ng_next;
            count = count_next;
        }
        if ((my_rounds & 1) && lane == 0) mbar_arrive(bar_full + 8 * ((my_rounds >> 1) & 1));   // the missing round of the last super-batch
        if (!ok && lane == 0 && a.fail) atomicExch(a.fail, 1);
    } else {
        // =========================================================================================== contraction warps
        const int part = warp - FQ;                  // 0 / 1 / 2: TMEM lanes 0-31 / 32-63 / 64-95 = b1 / b2 / b3 of W^T
        const uint32_t lane_base = (uint32_t)(32 * part) << 16;
        // W^T -> tensor memory: thread `lane` owns output channel o = lane; column c holds kc = 2c (low half), 2c+1
        bool ok = mbar_wait<true>(bar_w, 0);                                 // the bulk copy has landed
        for (int k = 0; k < a.K; ++k) {
            uint32_t v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float w0 = w_stage[(k * CIN + 2 * i) * COUT + lane];  // conflict-free: consecutive lanes, consecutive words
                const float w1 = w_stage[(k * CIN + 2 * i + 1) * COUT + lane];
                uint32_t e[3], o[3];
                split_bf16x3(w0, e[0], e[1], e[2]);
                split_bf16x3(w1, o[0], o[1], o[2]);
                v[i] = part == 0 ? pack_bf16(e[0], o[0]) : part == 1 ? pack_bf16(e[1], o[1]) : pack_bf16(e[2], o[2]);
            }
            tmem_st16(tmem + lane_base + (uint32_t)(16 * k), v);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_wfree);                               // this warp no longer reads the staging area
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        asm volatile("bar.sync 1, 96;\n" ::: "memory");                     // all three terms of A are in place
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");

        // D[128 x 96] (fp32) += A[128 x 16] (bf16, TMEM) * B[16 x 96] (bf16, shared memory), both K-major
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const float bias = (a.bias && lane < COUT) ? a.bias[lane] : 0.f;
        const int nks = a.K * CIN / 16;                                      // 30 K steps
        for (int sb = 0; sb < my_sb; ++sb) {
            const int buf = sb & 1;
            if (part == 0) {
                ok &= mbar_wait<true>(bar_full + 8 * buf, (sb >> 1) & 1);    // all 96 rows of the B tile are written
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                if (lane == 0) {
                    const uint64_t bdesc0 = make_desc(smem_u32(smem + (size_t)buf * L.b_bytes), B_LBO, B_SBO);
                    for (int ks = 0; ks < nks; ++ks)                         // one MMA = 16 k values = two 16-byte units
                        umma_bf16_ts(tmem + D_COL, tmem + 8 * ks, bdesc0 + (uint64_t)((2 * B_LBO * ks) >> 4), idesc, ks ? 1u : 0u);
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
                                 :: "r"(bar_done + 8 * buf) : "memory");
                }
                __syncwarp();
            }
            ok &= mbar_wait<true>(bar_done + 8 * buf, (sb >> 1) & 1);        // D is complete
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            float y[SBQ];
            {
                uint32_t d[32];
                tmem_ld32(tmem + lane_base + D_COL, d);                      // x b1(wf)
#pragma unroll
                for (int s = 0; s < SBQ; ++s) y[s] = __uint_as_float(d[s]);
                tmem_ld32(tmem + lane_base + D_COL + 32, d);                 // x b2(wf)
#pragma unroll
                for (int s = 0; s < SBQ; ++s) y[s] += __uint_as_float(d[s]);
                tmem_ld32(tmem + lane_base + D_COL + 64, d);                 // x b3(wf)
#pragma unroll
                for (int s = 0; s < SBQ; ++s) y[s] += __uint_as_float(d[s]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            if (part > 0) {
                float* xo = xch + (size_t)(part - 1) * 32 * (SBQ + 1) + lane * (SBQ + 1);
#pragma unroll
                for (int s = 0; s < SBQ; ++s) xo[s] = y[s];
            }
            asm volatile("bar.sync 1, 96;\n" ::: "memory");                 // b2 / b3 parts handed over; D may be overwritten
            if (part == 0) {
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");   // previous tile store has read out_s
                __syncwarp();
                const int q0 = (r_start + 2 * sb) * FQ;
                const float* x1 = xch + lane * (SBQ + 1);
                const float* x2 = x1 + 32 * (SBQ + 1);
#pragma unroll
                for (int s = 0; s < SBQ; ++s) {
                    float v = ((y[s] + x1[s]) + x2[s]) * invn_s[(sb & 3) * SBQ + s] + bias;
                    if (a.act) v = v > 0.f ? v : v * a.slope;
                    if (!ok) v = __int_as_float(0x7fc00000);
                    out_s[s * COUT + lane] = v;
                }
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    const int rows = min(min(SBQ, a.nq - q0), (my_rounds - 2 * sb) * FQ);
                    // [rows x 32] fp32 = rows * 128 contiguous bytes in global memory: one bulk (TMA) store
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n"
                                 :: "l"(a.out + (size_t)q0 * COUT), "r"(smem_u32(out_s)), "r"(rows * COUT * 4) : "memory");
                    asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
                }
            }
            asm volatile("bar.sync 1, 96;\n" ::: "memory");                 // xch may be rewritten by the next super-batch
        }
        if (part == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
        if (!ok && lane == 0 && a.fail) atomicExch(a.fail, 1);
    }

    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == FQ)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem), "r"(TMEM_COLS) : "memory");
}

template <bool IDX64>
int fused_launch(const FusedArgs& a, int grid, cudaStream_t stream) {
    const FusedSmem L = fused_smem();
    auto kern = kpf_fused_kernel<IDX64>;
    static bool attr_set = false;
    if (!attr_set) {
        D3F_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
        attr_set = true;
    }
    kern<<<grid, NT, L.total, stream>>>(a);
    D3F_CHECK_LAUNCH();
    return D3F_OK;
}

}  // namespace

// rigid, unmodulated, linear influence, sum aggregation, Cin = Cout = 32, K <= 15 (W^T as bf16 x 3 takes 16 K of the 512
// tensor-memory columns, D 96 more), <= 48 neighbour columns (all of a query's row loads are kept in flight in registers)
bool kpf_fused_eligible(int H, int K, int cin, int cout) {
    return cin == CIN && cout == COUT && K >= 1 && K <= 15 && H >= 1 && H <= HP;
}

int kpf_fused_launch(const Kp2Args& g, const float* weights, const float* bias, int act, float slope, float* out,
                     float* inv_n, float* wf, cudaStream_t stream) {
    static int n_sm = 0;
    if (!n_sm) {
        int dev = 0;
        D3F_CHECK_CUDA(cudaGetDevice(&dev));
        D3F_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    const int n_rounds = (g.nq + FQ - 1) / FQ;
    const int grid = n_rounds < n_sm ? n_rounds : n_sm;   // persistent: one CTA per SM (it owns all 512 tensor-memory columns)
    FusedArgs a{g.q, g.s, g.inds, g.ld, g.x, g.rowpos, g.kp, weights, bias, out, inv_n, wf,
                g.nq, g.ns, g.H, g.K, g.extent, act, slope, d3f_fail_flag_device()};
    return g.idx64 ? fused_launch<true>(a, grid, stream) : fused_launch<false>(a, grid, stream);
}
