// KPConv gather kernels, second generation (models/blocks.py:237-382; see kpconv.cu for the math and the C ABI).
//
// The v1 kernels were ISSUE-bound (ncu, round 1c: 5713 executed warp instructions per query at Cin = 32, 65 % issue
// active): half of them went into the influence evaluation (lanes over neighbours, 16 IEEE sqrt + div per lane, 29
// idle lanes in the second pass for H = 35), the rest into a neighbour loop whose loads could not overlap.  Here:
//
//   phase A  lanes over neighbours: index -> (s - q, code) records in shared memory, density count, last valid row
//   phase B  lanes over (neighbour, kernel point) PAIRS: lane = (h & 1) * 16 + k, so a lane keeps ITS kernel point in
//            registers, evaluates one influence per step (sqrt.approx, reciprocal extent) and stores it conflict-free;
//            closest-point masking / in-range filtering / min_d2 are 16-lane shuffles or ballots
//   phase C  lanes over channels.  FFMA mode: 4 neighbour rows are loaded per step before any arithmetic (4-8
//            independent 128-byte row loads in flight per warp), then 16*CG FFMA per row with the 16 weights read as
//            4 broadcast LDS.128.  MMA mode: per 8 neighbours one m16n8k8 3xTF32 product per 8 channels,
//            A = w [16 kernel points x 8 neighbours] from shared memory, B = 8 neighbour rows read as LDG.128 (the
//            lane's 4 consecutive channels feed 4 n-tiles), C = wf [16 x 8 channels] in registers.
//   scatter  same phases A/B, then dx[idx[h], :] += sum_k w[k,h] * m[k] * dwf[k,:] with the modulation folded into
//            the dwf registers once; for Cin % 32 == 0 (rigid) a lane owns 4 channels and the warp issues one
//            128-bit vector reduction (red.global.add.v4.f32) for 1-4 neighbour rows at a time.
#include "common.cuh"
#include "kpconv.cuh"
#include <limits.h>

namespace {

constexpr int KP = 16;            // kernel points padded to 16
constexpr int WS = 24;            // row stride of w_s (floats): conflict-free MMA fragment reads, 16-byte aligned rows
constexpr int U = 4;              // neighbour rows loaded per step in the FFMA loops
constexpr float SHADOW = 1e6f;    // blocks.py:277
constexpr int CODE_NONE = INT_MIN;
constexpr unsigned FULL = 0xffffffffu;

__host__ __device__ inline int kp2_hpad(int H) { return H <= 8 ? 8 : (H + 7) & ~7; }
// per-warp shared memory (floats): w_s[HP*WS] | rec_s[HP] float4 | idx_s[HP] | kp_s[48]
__host__ __device__ inline size_t kp2_warp_floats(int HP) { return (size_t)HP * (WS + 5) + 48; }

__device__ __forceinline__ float fast_sqrt(float v) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

// d influence / d sq  (kernel-point gradient of deformable layers)
__device__ __forceinline__ float kp2_influence_grad(float sq, float w, float extent, int influence) {
    if (influence == D3F_INFLUENCE_LINEAR) {
        if (!(1.0f - sqrtf(sq) / extent >= 0.0f)) return 0.0f;
        return -1.0f / (2.0f * extent * sqrtf(sq));
    }
    if (influence == D3F_INFLUENCE_GAUSSIAN) {
        const float sigma = extent * 0.3f;
        return -w / (2.0f * sigma * sigma + 1e-9f);
    }
    return 0.0f;
}

struct Kp2Warp {
    float* w_s; float4* rec_s; int* idx_s; float* kp_s;
    int count;    // neighbours with a positive feature sum (density normalisation, blocks.py:377)
    int hend;     // 1 + last kept neighbour row
    float md;     // min over neighbours of |rel - kp[lane & 15]|^2  (deformed only)
};

// Phases A and B for one query (whole warp).  After the call w_s[h*WS + k] holds the influence of kernel point k on
// neighbour h (0 for dropped / shadow / padding rows and for k >= K), idx_s[h] the support row (-1 = dropped) and
// rec_s[h].xyz the neighbour position relative to the query.
template <bool IDX64, bool DEFORMED>
__device__ __forceinline__ void kp2_phases(const Kp2Args& a, int qi, int lane, int HP, Kp2Warp& wp, bool want_kp_s) {
    const float qx = a.q[3 * (size_t)qi], qy = a.q[3 * (size_t)qi + 1], qz = a.q[3 * (size_t)qi + 2];
    int count = 0, hend = 0;
    // ---- A: lanes over neighbours
    for (int h0 = 0; h0 < HP; h0 += 32) {
        const int h = h0 + lane;
        bool valid = false, pos = false;
        if (h < HP) {
            long long idx = -1;
            if (h < a.H)
                idx = IDX64 ? ((const long long*)a.inds)[(size_t)qi * a.ld + h]
                            : (long long)((const int*)a.inds)[(size_t)qi * a.ld + h];
            valid = idx >= 0 && idx < a.ns;
            float rx = SHADOW - qx, ry = SHADOW - qy, rz = SHADOW - qz;
            int code = CODE_NONE;
            if (valid) {
                rx = a.s[3 * idx] - qx; ry = a.s[3 * idx + 1] - qy; rz = a.s[3 * idx + 2] - qz;
                pos = a.rowpos[idx] != 0;
                code = pos ? (int)idx : ~(int)idx;
            }
            wp.rec_s[h] = make_float4(rx, ry, rz, __int_as_float(code));
            wp.idx_s[h] = valid ? (int)idx : -1;
        }
        const unsigned bv = __ballot_sync(FULL, valid);
        count += __popc(__ballot_sync(FULL, pos));
        if (bv) hend = h0 + 32 - __clz(bv);
    }
    // ---- B: lanes over (neighbour parity, kernel point)
    const int k = lane & 15, hh = lane >> 4;
    const bool kvalid = k < a.K;
    float kx = 0.f, ky = 0.f, kz = 0.f;
    if (kvalid) {
        const float* p = DEFORMED ? a.kp + ((size_t)qi * a.K + k) * 3 : a.kp + 3 * k;
        kx = p[0]; ky = p[1]; kz = p[2];
    }
    if (want_kp_s && lane < KP) { wp.kp_s[3 * lane] = kx; wp.kp_s[3 * lane + 1] = ky; wp.kp_s[3 * lane + 2] = kz; }
    __syncwarp();
    const float ext2 = a.extent * a.extent, inv_ext = 1.0f / a.extent;
    const float sigma = a.extent * 0.3f, gden = 2.0f * sigma * sigma + 1e-9f;
    float md = INFINITY;
    if (DEFORMED) { count = 0; hend = 0; }
    for (int t = 0; t < (HP >> 1); ++t) {
        const int h = 2 * t + hh;
        const float4 r = wp.rec_s[h];
        const int code = __float_as_int(r.w);
        const float dx = r.x - kx, dy = r.y - ky, dz = r.z - kz;
        const float sq = dx * dx + dy * dy + dz * dz;
        float w;
        if (a.influence == D3F_INFLUENCE_LINEAR) w = fmaxf(1.0f - fast_sqrt(sq) * inv_ext, 0.0f);
        else if (a.influence == D3F_INFLUENCE_GAUSSIAN) w = expf(-sq / gden);
        else w = 1.0f;
        if (a.aggregation == D3F_AGGREGATION_CLOSEST) {   // keep the nearest kernel point only (first minimum)
            float bs = kvalid ? sq : INFINITY;
            int bk = k;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                const float os = __shfl_xor_sync(FULL, bs, o);
                const int ok = __shfl_xor_sync(FULL, bk, o);
                if (os < bs || (os == bs && ok < bk)) { bs = os; bk = ok; }
            }
            if (bk != k) w = 0.f;
        }
        bool keep = code != CODE_NONE;
        if (DEFORMED) {   // blocks.py:300-324: a neighbour counts only if it is inside some kernel point's extent
            const unsigned b = __ballot_sync(FULL, kvalid && sq < ext2);
            keep = keep && ((b >> (hh * 16)) & 0xFFFFu) != 0;
            if (h < a.H) md = fminf(md, sq);
            const bool lead = k == 0;
            const unsigned bk = __ballot_sync(FULL, keep && lead);
            count += __popc(__ballot_sync(FULL, keep && lead && code >= 0));
            if (lead) wp.idx_s[h] = keep ? (code >= 0 ? code : ~code) : -1;
            if (bk & 0x10000u) hend = 2 * t + 2;
            else if (bk & 1u) hend = 2 * t + 1;
        }
        wp.w_s[h * WS + k] = (keep && kvalid) ? w : 0.f;
    }
    __syncwarp();
    wp.count = count; wp.hend = hend; wp.md = md;
}

__device__ __forceinline__ void kp2_slab(Kp2Warp& wp, float* base, int warp, int HP) {
    wp.w_s = base + (size_t)warp * kp2_warp_floats(HP);
    wp.rec_s = (float4*)(wp.w_s + (size_t)HP * WS);
    wp.idx_s = (int*)(wp.rec_s + HP);
    wp.kp_s = (float*)(wp.idx_s + HP);
}

// 3xTF32 operand split.  cvt.rna.tf32 is emulated on sm_100a (FSETP + IADD + LOP3), so round to nearest by hand
// (add half an ulp of the 10-bit mantissa, clear 13 bits: 2 instructions; inputs are finite) and leave the remainder
// unrounded: the tensor core ignores its low 13 bits, an error of 2^-21 relative to v.
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
    hi = (__float_as_uint(v) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(v - __uint_as_float(hi));
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// ------------------------------------------------------------------------------------------------ forward
// wf[qi, k, c] = m[qi,k] * sum_h w[k,h] * x[idx[h], c]      one warp per query
template <bool IDX64, bool DEFORMED, int CG, int MODE>
__global__ void __launch_bounds__(256)
kp2_correlate_kernel(Kp2Args a, int HP, float* __restrict__ wf, float* __restrict__ wf_unmod, float* __restrict__ inv_n,
                     float* __restrict__ min_d2) {
    extern __shared__ float4 smem_f4[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qi = blockIdx.x * (blockDim.x >> 5) + warp;
    if (qi >= a.nq) return;                          // warps are independent: no block-wide barrier below
    Kp2Warp wp;
    kp2_slab(wp, (float*)smem_f4, warp, HP);
    kp2_phases<IDX64, DEFORMED>(a, qi, lane, HP, wp, false);
    if (lane == 0) inv_n[qi] = 1.0f / (float)max(wp.count, 1);
    if (DEFORMED && min_d2) {
        const float v = fminf(wp.md, __shfl_xor_sync(FULL, wp.md, 16));
        if (lane < a.K) min_d2[(size_t)qi * a.K + lane] = v;
    }
    const float* __restrict__ x = a.x;
    const int cin = a.cin;
    if (cin == 1 && !a.mod) {
        // ---- one input channel (the first layer: the reference feeds a constant feature, datasets/ThreeDMatch.py):
        // wf[qi, k] = sum_h w[k,h] x[idx[h]].  The generic loops below put the lanes over CHANNELS -- one lane of 32 at
        // work, 96 us for 40000 queries at the head of the forward pass; here the lanes are (neighbour parity, kernel
        // point), as in phase B.
        const int k = lane & 15, hh = lane >> 4;
        float acc = 0.f;
        for (int h = hh; h < wp.hend; h += 2) {
            const int idx = wp.idx_s[h];
            if (idx >= 0) acc = fmaf(wp.w_s[h * WS + k], __ldg(x + idx), acc);
        }
        acc += __shfl_xor_sync(FULL, acc, 16);
        if (lane < a.K) wf[(size_t)qi * a.K + lane] = acc;
        return;
    }
    // gridDim.y > 1: one 32*CG-channel chunk per blockIdx.y (deep levels: a few hundred queries x 256-512 channels would
    // otherwise occupy 32-96 CTAs for tens of microseconds of pure latency)
    const int c_lo = gridDim.y > 1 ? (int)blockIdx.y * 32 * CG : 0;
    const int c_hi = gridDim.y > 1 ? min(cin, c_lo + 32 * CG) : cin;

    if (MODE == 0) {
        // ---- C (FFMA): lanes over channels, U rows in flight.  Row offsets are 32-bit (the launcher checks
        // Ns * Cin < 2^31); lanes beyond Cin read channel Cin-1 and drop the result, so the loop has no predicates.
        const int hr = (wp.hend + U - 1) & ~(U - 1);
        const unsigned ucin = (unsigned)cin;
        for (int c0 = c_lo; c0 < c_hi; c0 += 32 * CG) {
            float acc[CG][KP];
            const float* xc[CG];
#pragma unroll
            for (int j = 0; j < CG; ++j) {
                xc[j] = x + min(c0 + j * 32 + lane, cin - 1);
#pragma unroll
                for (int k = 0; k < KP; ++k) acc[j][k] = 0.f;
            }
            for (int h0 = 0; h0 < hr; h0 += U) {
                const int4 id4 = *(const int4*)&wp.idx_s[h0];
                const unsigned off[U] = {(unsigned)max(id4.x, 0) * ucin, (unsigned)max(id4.y, 0) * ucin,
                                         (unsigned)max(id4.z, 0) * ucin, (unsigned)max(id4.w, 0) * ucin};
                float xv[U][CG];
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int j = 0; j < CG; ++j) xv[u][j] = __ldg(xc[j] + off[u]);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const float4* w4 = (const float4*)(wp.w_s + (h0 + u) * WS);
                    float w[KP];
#pragma unroll
                    for (int v = 0; v < KP / 4; ++v) {
                        const float4 t = w4[v];
                        w[4 * v] = t.x; w[4 * v + 1] = t.y; w[4 * v + 2] = t.z; w[4 * v + 3] = t.w;
                    }
#pragma unroll
                    for (int j = 0; j < CG; ++j)
#pragma unroll
                        for (int k = 0; k < KP; ++k) acc[j][k] = fmaf(w[k], xv[u][j], acc[j][k]);
                }
            }
#pragma unroll
            for (int j = 0; j < CG; ++j) {
                const int c = c0 + j * 32 + lane;
                if (c < cin) {
                    float* o = wf + (size_t)qi * a.K * cin + c;
                    if (!a.mod) {
#pragma unroll
                        for (int k = 0; k < KP; ++k)
                            if (k < a.K) o[(size_t)k * cin] = acc[j][k];
                    } else {
                        float* ou = wf_unmod ? wf_unmod + (size_t)qi * a.K * cin + c : nullptr;
#pragma unroll
                        for (int k = 0; k < KP; ++k)
                            if (k < a.K) {
                                if (ou) ou[(size_t)k * cin] = acc[j][k];
                                o[(size_t)k * cin] = acc[j][k] * a.mod[(size_t)qi * a.K + k];
                            }
                    }
                }
            }
        }
    } else {
        // ---- C (tensor cores): wf[16 x 32*CG] = w[16 x 8] * X[8 x 32*CG] per 8 neighbours, 3xTF32
        const int gq = lane >> 2, tq = lane & 3;
        const int hr = (wp.hend + 7) & ~7;
        for (int c0 = c_lo; c0 < c_hi; c0 += 32 * CG) {
            float acc[CG][4][4];
#pragma unroll
            for (int j = 0; j < CG; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[j][e][i] = 0.f;
            for (int h0 = 0; h0 < hr; h0 += 8) {
                const unsigned ra = (unsigned)max(wp.idx_s[h0 + tq], 0) * (unsigned)cin;
                const unsigned rb = (unsigned)max(wp.idx_s[h0 + tq + 4], 0) * (unsigned)cin;
                float4 xa[CG], xb[CG];
#pragma unroll
                for (int j = 0; j < CG; ++j) {   // lanes beyond Cin re-read the last 4 channels and drop the result
                    const float* xc = x + min(c0 + j * 32 + 4 * gq, cin - 4);
                    xa[j] = __ldg((const float4*)(xc + ra));
                    xb[j] = __ldg((const float4*)(xc + rb));
                }
                // A[m = kernel point][k = neighbour]: a0 (gq, tq)  a1 (gq+8, tq)  a2 (gq, tq+4)  a3 (gq+8, tq+4)
                const float* wq = wp.w_s + (h0 + tq) * WS + gq;
                uint32_t ah[4], al[4];
                split_tf32(wq[0], ah[0], al[0]);
                split_tf32(wq[8], ah[1], al[1]);
                split_tf32(wq[4 * WS], ah[2], al[2]);
                split_tf32(wq[4 * WS + 8], ah[3], al[3]);
#pragma unroll
                for (int j = 0; j < CG; ++j) {
                    const float va[4] = {xa[j].x, xa[j].y, xa[j].z, xa[j].w};
                    const float vb[4] = {xb[j].x, xb[j].y, xb[j].z, xb[j].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {   // n-tile e holds channels 4*n + e, n = 0..7
                        uint32_t bh[2], bl[2];
                        split_tf32(va[e], bh[0], bl[0]);     // B[k = tq][n = gq]
                        split_tf32(vb[e], bh[1], bl[1]);     // B[k = tq + 4][n = gq]
                        mma_tf32(acc[j][e], al, bh);
                        mma_tf32(acc[j][e], ah, bl);
                        mma_tf32(acc[j][e], ah, bh);
                    }
                }
            }
            // C[m][n]: c0 (gq, 2tq)  c1 (gq, 2tq+1)  c2 (gq+8, 2tq)  c3 (gq+8, 2tq+1); channel = c0 + 32j + 4n + e
#pragma unroll
            for (int j = 0; j < CG; ++j)
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int k = gq + 8 * half;
                    if (k >= a.K) continue;
                    const float m = a.mod ? a.mod[(size_t)qi * a.K + k] : 1.0f;
#pragma unroll
                    for (int nn = 0; nn < 2; ++nn) {
                        const int c = c0 + j * 32 + 4 * (2 * tq + nn);
                        if (c >= cin) continue;
                        const int i = 2 * half + nn;
                        float4 v = make_float4(acc[j][0][i], acc[j][1][i], acc[j][2][i], acc[j][3][i]);
                        const size_t o = ((size_t)qi * a.K + k) * cin + c;
                        if (a.mod) {
                            if (wf_unmod) *(float4*)(wf_unmod + o) = v;
                            v.x *= m; v.y *= m; v.z *= m; v.w *= m;
                        }
                        *(float4*)(wf + o) = v;
                    }
                }
        }
    }
}

// ------------------------------------------------------------------------------------------------ backward
// dx[idx[h], c] += sum_k m[k] w[k,h] dwf[k,c];  deformed: dkp[qi,k,:], dmod[qi,k] = sum_c dwf[k,c] wf_unmod[k,c]
template <bool IDX64, bool DEFORMED, int CG>
__global__ void __launch_bounds__(256)
kp2_scatter_kernel(Kp2Args a, int HP, const float* __restrict__ dwf, const float* __restrict__ wf_unmod,
                   float* __restrict__ grad_x, float* __restrict__ grad_kp, float* __restrict__ grad_mod) {
    extern __shared__ float4 smem_f4[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qi = blockIdx.x * (blockDim.x >> 5) + warp;
    if (qi >= a.nq) return;
    Kp2Warp wp;
    kp2_slab(wp, (float*)smem_f4, warp, HP);
    kp2_phases<IDX64, DEFORMED>(a, qi, lane, HP, wp, DEFORMED);
    const int cin = a.cin;
    const int hr = (wp.hend + U - 1) & ~(U - 1);

    float mk[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) mk[k] = (a.mod && k < a.K) ? a.mod[(size_t)qi * a.K + k] : 1.0f;
    float gkp[KP][3];
    float gmod[KP];
    if (DEFORMED) {
#pragma unroll
        for (int k = 0; k < KP; ++k) { gkp[k][0] = gkp[k][1] = gkp[k][2] = 0.f; gmod[k] = 0.f; }
    }

    // gridDim.y > 1: the channel chunks of a query are spread over blockIdx.y (few queries x many channels x wide
    // neighbourhoods -- the deformable layers of levels 3-4 -- would otherwise run on a fraction of the SMs); the
    // per-query outputs are then combined with atomics into zero-filled buffers
    const bool split = gridDim.y > 1;
    const int c_lo = split ? (int)blockIdx.y * 32 * CG : 0, c_hi = split ? min(cin, c_lo + 32 * CG) : cin;
    for (int c0 = c_lo; c0 < c_hi; c0 += 32 * CG) {
        float d[CG][KP];   // m[k] * dwf[k, c]
#pragma unroll
        for (int j = 0; j < CG; ++j) {
            const int c = c0 + j * 32 + lane;
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                float v = (k < a.K && c < cin) ? dwf[((size_t)qi * a.K + k) * cin + c] : 0.f;
                if (DEFORMED && grad_mod && wf_unmod && k < a.K && c < cin)
                    gmod[k] = fmaf(v, wf_unmod[((size_t)qi * a.K + k) * cin + c], gmod[k]);
                d[j][k] = v * mk[k];
            }
        }
        for (int h0 = 0; h0 < hr; h0 += U) {
            const int4 id4 = *(const int4*)&wp.idx_s[h0];
            const int ids[U] = {id4.x, id4.y, id4.z, id4.w};
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int idx = ids[u];
                if (idx < 0) continue;                // warp-uniform
                const int h = h0 + u;
                const float4* w4 = (const float4*)(wp.w_s + h * WS);
                float w[KP];
#pragma unroll
                for (int v = 0; v < KP / 4; ++v) {
                    const float4 t = w4[v];
                    w[4 * v] = t.x; w[4 * v + 1] = t.y; w[4 * v + 2] = t.z; w[4 * v + 3] = t.w;
                }
                if (grad_x) {
#pragma unroll
                    for (int j = 0; j < CG; ++j) {
                        const int c = c0 + j * 32 + lane;
                        float v = 0.f;
#pragma unroll
                        for (int k = 0; k < KP; ++k) v = fmaf(w[k], d[j][k], v);
                        if (c < cin) atomicAdd(&grad_x[(size_t)idx * cin + c], v);
                    }
                }
                if (DEFORMED && grad_kp) {
                    float xv[CG];
#pragma unroll
                    for (int j = 0; j < CG; ++j) {
                        const int c = c0 + j * 32 + lane;
                        xv[j] = c < cin ? __ldg(&a.x[(size_t)idx * cin + c]) : 0.f;
                    }
                    const float4 r = wp.rec_s[h];
                    float best = INFINITY; int best_k = 0;
                    if (a.aggregation == D3F_AGGREGATION_CLOSEST) {
#pragma unroll
                        for (int k = 0; k < KP; ++k)
                            if (k < a.K) {
                                const float dx = r.x - wp.kp_s[3 * k], dy = r.y - wp.kp_s[3 * k + 1], dz = r.z - wp.kp_s[3 * k + 2];
                                const float sq = dx * dx + dy * dy + dz * dz;
                                if (sq < best) { best = sq; best_k = k; }
                            }
                    }
#pragma unroll
                    for (int k = 0; k < KP; ++k)
                        if (k < a.K) {
                            float t = 0.f;   // <m[k] dwf[k,:], x[idx,:]> over this lane's channels (reduced over lanes at the end)
#pragma unroll
                            for (int j = 0; j < CG; ++j) t = fmaf(d[j][k], xv[j], t);
                            const float dx = r.x - wp.kp_s[3 * k], dy = r.y - wp.kp_s[3 * k + 1], dz = r.z - wp.kp_s[3 * k + 2];
                            const float sq = dx * dx + dy * dy + dz * dz;
                            float gw = kp2_influence_grad(sq, w[k], a.extent, a.influence);
                            if (a.aggregation == D3F_AGGREGATION_CLOSEST && k != best_k) gw = 0.f;
                            const float f = t * gw * (-2.0f);
                            gkp[k][0] = fmaf(f, dx, gkp[k][0]);
                            gkp[k][1] = fmaf(f, dy, gkp[k][1]);
                            gkp[k][2] = fmaf(f, dz, gkp[k][2]);
                        }
                }
            }
        }
    }
    if (DEFORMED) {
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            if (k >= a.K) continue;
            if (grad_kp) {
#pragma unroll
                for (int ax = 0; ax < 3; ++ax) {
                    float v = gkp[k][ax];
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
                    if (lane == 0) {
                        if (split) atomicAdd(&grad_kp[((size_t)qi * a.K + k) * 3 + ax], v);
                        else grad_kp[((size_t)qi * a.K + k) * 3 + ax] = v;
                    }
                }
            }
            if (grad_mod) {
                float v = gmod[k];
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
                if (lane == 0) {
                    if (split) atomicAdd(&grad_mod[(size_t)qi * a.K + k], v);
                    else grad_mod[(size_t)qi * a.K + k] = v;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ backward, kernel points
// grad_kernel_points of a deformable (unmodulated) layer on the tensor path:
//   T[k, h] = sum_c dwf[qi, k, c] * x[idx[h], c]                     one [16 x Cin] x [Cin x 8] mma chain per 8 neighbours
//   dkp[qi, k, :] = sum_h T[k, h] * dw/dsq(h, k) * (-2) * (rel[h] - kp[k])
// The scatter kernel above computes the same sums with the lanes over CHANNELS, so all 32 lanes repeat the geometry of
// every (neighbour, kernel point) pair: 1.05 ms per layer at level 3 of BASELINE config 4 (1344 queries x 173 neighbours
// x 256 channels), three quarters of that step's backward.  Here the contraction over channels is 3xTF32 mma.sync and a
// lane evaluates the geometry of the 4 (k, h) pairs its accumulator fragment holds.  Cin % 32 == 0.
template <bool IDX64>
__global__ void __launch_bounds__(256)
kp2_gradkp_kernel(Kp2Args a, int HP, const float* __restrict__ dwf, float* __restrict__ grad_kp) {
    extern __shared__ float4 smem_f4[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qi = blockIdx.x * (blockDim.x >> 5) + warp;
    if (qi >= a.nq) return;
    Kp2Warp wp;
    kp2_slab(wp, (float*)smem_f4, warp, HP);
    kp2_phases<IDX64, true>(a, qi, lane, HP, wp, true);     // w_s, rec_s, idx_s (kept neighbours only), kp_s
    const int cin = a.cin, gq = lane >> 2, tq = lane & 3;
    const float* __restrict__ D0 = dwf + ((size_t)qi * a.K + gq) * cin + 4 * tq;                 // kernel point gq
    const bool k1 = gq + 8 < a.K;
    const float* __restrict__ D1 = dwf + ((size_t)qi * a.K + (k1 ? gq + 8 : gq)) * cin + 4 * tq;  // kernel point gq + 8
    const float* __restrict__ x = a.x + 4 * tq;
    float g0[3] = {0.f, 0.f, 0.f}, g1[3] = {0.f, 0.f, 0.f};
    const int nblk = (wp.hend + 7) >> 3;
    constexpr int GB = 6;                                   // 8-neighbour blocks per pass: 24 accumulator registers
    for (int hb0 = 0; hb0 < nblk; hb0 += GB) {
        float acc[GB][4];
        unsigned row[GB];
#pragma unroll
        for (int b = 0; b < GB; ++b) {
            acc[b][0] = acc[b][1] = acc[b][2] = acc[b][3] = 0.f;
            const int id = hb0 + b < nblk ? wp.idx_s[(hb0 + b) * 8 + gq] : -1;
            row[b] = (unsigned)max(id, 0) * (unsigned)cin;  // dropped neighbours read row 0; their T is never used
        }
        for (int c0 = 0; c0 < cin; c0 += 32) {
            // mma k index <-> channel: step s covers channels c0 + 4 tq + s (k = tq) and c0 + 16 + 4 tq + s (k = tq + 4)
            const float4 d0 = __ldg((const float4*)(D0 + c0)), d2 = __ldg((const float4*)(D0 + c0 + 16));
            float4 d1 = make_float4(0.f, 0.f, 0.f, 0.f), d3 = d1;
            if (k1) { d1 = __ldg((const float4*)(D1 + c0)); d3 = __ldg((const float4*)(D1 + c0 + 16)); }
            const float da[4][4] = {{d0.x, d1.x, d2.x, d3.x}, {d0.y, d1.y, d2.y, d3.y}, {d0.z, d1.z, d2.z, d3.z},
                                    {d0.w, d1.w, d2.w, d3.w}};
            uint32_t ah[4][4], al[4][4];
#pragma unroll
            for (int s = 0; s < 4; ++s)
#pragma unroll
                for (int e = 0; e < 4; ++e) split_tf32(da[s][e], ah[s][e], al[s][e]);
#pragma unroll
            for (int b = 0; b < GB; ++b) {
                if (hb0 + b >= nblk) break;                 // warp-uniform
                const float4 xa = __ldg((const float4*)(x + row[b] + c0));
                const float4 xb = __ldg((const float4*)(x + row[b] + c0 + 16));
                const float va[4] = {xa.x, xa.y, xa.z, xa.w}, vb[4] = {xb.x, xb.y, xb.z, xb.w};
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    uint32_t bh[2], bl[2];
                    split_tf32(va[s], bh[0], bl[0]);
                    split_tf32(vb[s], bh[1], bl[1]);
                    mma_tf32(acc[b], al[s], bh);
                    mma_tf32(acc[b], ah[s], bl);
                    mma_tf32(acc[b], ah[s], bh);
                }
            }
        }
        // accumulator fragment: acc[b][2 * half + e] = T[k = gq + 8 * half][h = 8 * (hb0 + b) + 2 * tq + e]
#pragma unroll
        for (int b = 0; b < GB; ++b) {
            if (hb0 + b >= nblk) break;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int h = (hb0 + b) * 8 + 2 * tq + e;
                if (wp.idx_s[h] < 0) continue;
                const float4 r = wp.rec_s[h];
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int k = gq + 8 * half;
                    if (k >= a.K) continue;
                    const float dx = r.x - wp.kp_s[3 * k], dy = r.y - wp.kp_s[3 * k + 1], dz = r.z - wp.kp_s[3 * k + 2];
                    const float sq = dx * dx + dy * dy + dz * dz;
                    const float w = wp.w_s[h * WS + k];
                    float gw = kp2_influence_grad(sq, w, a.extent, a.influence);
                    if (a.aggregation == D3F_AGGREGATION_CLOSEST && w == 0.f) gw = 0.f;   // w_s is zero off the closest point
                    const float f = acc[b][2 * half + e] * gw * (-2.0f);
                    float* g = half ? g1 : g0;
                    g[0] = fmaf(f, dx, g[0]); g[1] = fmaf(f, dy, g[1]); g[2] = fmaf(f, dz, g[2]);
                }
            }
        }
    }
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
        float v0 = g0[ax], v1 = g1[ax];
        v0 += __shfl_xor_sync(FULL, v0, 1); v0 += __shfl_xor_sync(FULL, v0, 2);
        v1 += __shfl_xor_sync(FULL, v1, 1); v1 += __shfl_xor_sync(FULL, v1, 2);
        if (tq == 0) {
            if (gq < a.K) grad_kp[((size_t)qi * a.K + gq) * 3 + ax] = v0;
            if (k1) grad_kp[((size_t)qi * a.K + gq + 8) * 3 + ax] = v1;
        }
    }
}

// Rigid layers with Cin a multiple of 32 (<= 128 per pass): a lane owns 4 consecutive channels, VL = 8*CV lanes cover
// a 32*CV-channel row, and the warp's 32/VL lane groups take different neighbours, so one step reads the weights of
// 32/VL neighbours with 4 LDS.128 and issues ONE vector reduction (16 bytes per lane).
template <bool IDX64, int CV>   // CV = channels of one pass / 32: 1, 2 or 4
__global__ void __launch_bounds__(256)
kp2_scatter_vec_kernel(Kp2Args a, int HP, const float* __restrict__ dwf, float* __restrict__ grad_x) {
    extern __shared__ float4 smem_f4[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qi = blockIdx.x * (blockDim.x >> 5) + warp;
    if (qi >= a.nq) return;
    Kp2Warp wp;
    kp2_slab(wp, (float*)smem_f4, warp, HP);
    kp2_phases<IDX64, false>(a, qi, lane, HP, wp, false);
    constexpr int VL = 8 * CV, NU = 32 / VL;        // lanes per row, rows per step
    const int sub = lane / VL, cl = (lane % VL) * 4;
    const int cin = a.cin;
    const int hr = (wp.hend + NU - 1) / NU * NU;    // <= HP (HP is a multiple of 8 >= NU); rows >= hend are dropped
    for (int c0 = 0; c0 < cin; c0 += 32 * CV) {
        float4 d[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            d[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < a.K) {
                d[k] = *(const float4*)(dwf + ((size_t)qi * a.K + k) * cin + c0 + cl);
                if (a.mod) { const float m = a.mod[(size_t)qi * a.K + k]; d[k].x *= m; d[k].y *= m; d[k].z *= m; d[k].w *= m; }
            }
        }
        for (int h0 = 0; h0 < hr; h0 += NU) {
            const int h = h0 + sub;
            const int idx = wp.idx_s[h];
            const float4* w4 = (const float4*)(wp.w_s + h * WS);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int q4 = 0; q4 < KP / 4; ++q4) {
                const float4 t = w4[q4];
                const float w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float4 dk = d[4 * q4 + e];
                    v.x = fmaf(w[e], dk.x, v.x); v.y = fmaf(w[e], dk.y, v.y);
                    v.z = fmaf(w[e], dk.z, v.z); v.w = fmaf(w[e], dk.w, v.w);
                }
            }
            if (idx >= 0) atomicAdd((float4*)(grad_x + (size_t)idx * cin + c0 + cl), v);
        }
    }
}

// ------------------------------------------------------------------------------------------------ backward, atomic-free
// G[j, k, o] = sum over the queries i that list support j of  w[i, k, h(i,j)] * inv_n[i] * g[i, o]
// -- the forward gather over the TRANSPOSED neighbour lists (t_off / t_src, transpose.cu), reading rows of the output
// gradient instead of rows of x.  dx = G [Ns, K*Cout] x W^T then is one GEMM (kpconv.cu).  One warp per support point,
// lists walked in chunks of HT entries with the accumulators kept in registers; Cout % 32 == 0.  Deformable layers read
// the listing query's own kernel points per entry (the gradient of those points is query-major and stays with the
// scatter kernel, which then skips its grad_x reductions).
constexpr int HT = 64;
__host__ __device__ inline size_t kp2t_warp_floats() { return (size_t)HT * (WS + 5); }   // w_s | rec_s (float4) | idx_s

template <int CG>
__global__ void __launch_bounds__(256)
kp2t_correlate_kernel(Kp2tArgs a, float* __restrict__ G) {
    extern __shared__ float4 smem_f4[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * (blockDim.x >> 5) + warp;
    if (j >= a.ns) return;
    float* w_s = (float*)smem_f4 + (size_t)warp * kp2t_warp_floats();
    float4* rec_s = (float4*)(w_s + HT * WS);
    int* idx_s = (int*)(rec_s + HT);
    const float sx = a.s[3 * (size_t)j], sy = a.s[3 * (size_t)j + 1], sz = a.s[3 * (size_t)j + 2];
    const int e0 = a.t_off[j], e1 = a.t_off[j + 1];
    const int k = lane & 15, hh = lane >> 4;
    const bool kvalid = k < a.K;
    float kx = 0.f, ky = 0.f, kz = 0.f;
    if (kvalid && !a.deformed) { kx = a.kp[3 * k]; ky = a.kp[3 * k + 1]; kz = a.kp[3 * k + 2]; }
    const float inv_ext = 1.0f / a.extent, ext2 = a.extent * a.extent;
    const float sigma = a.extent * 0.3f, gden = 2.0f * sigma * sigma + 1e-9f;
    const int gq = lane >> 2, tq = lane & 3;
    const int cout = a.cout;
    const float* __restrict__ g = a.g;

    const int c_lo = gridDim.y > 1 ? (int)blockIdx.y * 32 * CG : 0;          // one channel chunk per blockIdx.y (small grids)
    const int c_hi = gridDim.y > 1 ? min(cout, c_lo + 32 * CG) : cout;
    for (int c0 = c_lo; c0 < c_hi; c0 += 32 * CG) {
        float acc[CG][4][4];
#pragma unroll
        for (int jj = 0; jj < CG; ++jj)
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[jj][e][i] = 0.f;
        for (int base = e0; base < e1; base += HT) {
            const int cnt = min(HT, e1 - base), cnt8 = (cnt + 7) & ~7;
            // ---- A: lanes over list entries: (s_j - q_i, inv_n[i]) records
            for (int h0 = 0; h0 < cnt8; h0 += 32) {
                const int h = h0 + lane;
                if (h < cnt8) {
                    int i = -1;
                    float rx = SHADOW, ry = SHADOW, rz = SHADOW, sc = 0.f;
                    if (h < cnt) {
                        i = a.t_src[base + h];
                        rx = sx - a.q[3 * (size_t)i]; ry = sy - a.q[3 * (size_t)i + 1]; rz = sz - a.q[3 * (size_t)i + 2];
                        sc = a.inv_n[i];
                    }
                    rec_s[h] = make_float4(rx, ry, rz, sc);
                    idx_s[h] = i;
                }
            }
            __syncwarp();
            // ---- B: lanes over (entry parity, kernel point); the density normalisation rides in the weight
            for (int t = 0; t < (cnt8 >> 1); ++t) {
                const int h = 2 * t + hh;
                const float4 r = rec_s[h];
                if (a.deformed) {      // the listing query's own (deformed) kernel points, blocks.py:286-291
                    const int i = idx_s[h];
                    if (kvalid && i >= 0) {
                        const float* p = a.kp + ((size_t)i * a.K + k) * 3;
                        kx = p[0]; ky = p[1]; kz = p[2];
                    }
                }
                const float dx = r.x - kx, dy = r.y - ky, dz = r.z - kz;
                const float sq = dx * dx + dy * dy + dz * dz;
                float w;
                if (a.influence == D3F_INFLUENCE_LINEAR) w = fmaxf(1.0f - fast_sqrt(sq) * inv_ext, 0.0f);
                else if (a.influence == D3F_INFLUENCE_GAUSSIAN) w = expf(-sq / gden);
                else w = 1.0f;
                if (a.aggregation == D3F_AGGREGATION_CLOSEST) {
                    float bs = kvalid ? sq : INFINITY;
                    int bk = k;
#pragma unroll
                    for (int o = 8; o > 0; o >>= 1) {
                        const float os = __shfl_xor_sync(FULL, bs, o);
                        const int ok = __shfl_xor_sync(FULL, bk, o);
                        if (os < bs || (os == bs && ok < bk)) { bs = os; bk = ok; }
                    }
                    if (bk != k) w = 0.f;
                }
                if (a.deformed) {      // blocks.py:300-324: the pair counts only inside some kernel point's extent
                    const unsigned b = __ballot_sync(FULL, kvalid && sq < ext2);
                    if (((b >> (hh * 16)) & 0xFFFFu) == 0) w = 0.f;
                }
                w_s[h * WS + k] = kvalid ? w * r.w : 0.f;      // r.w = 0 for the padding entries
            }
            __syncwarp();
            // ---- C: G[16 x 32*CG] += w[16 x 8] * g_rows[8 x 32*CG] per 8 entries, 3xTF32
            for (int h0 = 0; h0 < cnt8; h0 += 8) {
                const unsigned ra = (unsigned)max(idx_s[h0 + tq], 0) * (unsigned)cout;
                const unsigned rb = (unsigned)max(idx_s[h0 + tq + 4], 0) * (unsigned)cout;
                float4 xa[CG], xb[CG];
#pragma unroll
                for (int jj = 0; jj < CG; ++jj) {
                    const float* gc = g + min(c0 + jj * 32 + 4 * gq, cout - 4);
                    xa[jj] = __ldg((const float4*)(gc + ra));
                    xb[jj] = __ldg((const float4*)(gc + rb));
                }
                const float* wq = w_s + (h0 + tq) * WS + gq;
                uint32_t ah[4], al[4];
                split_tf32(wq[0], ah[0], al[0]);
                split_tf32(wq[8], ah[1], al[1]);
                split_tf32(wq[4 * WS], ah[2], al[2]);
                split_tf32(wq[4 * WS + 8], ah[3], al[3]);
#pragma unroll
                for (int jj = 0; jj < CG; ++jj) {
                    const float va[4] = {xa[jj].x, xa[jj].y, xa[jj].z, xa[jj].w};
                    const float vb[4] = {xb[jj].x, xb[jj].y, xb[jj].z, xb[jj].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        uint32_t bh[2], bl[2];
                        split_tf32(va[e], bh[0], bl[0]);
                        split_tf32(vb[e], bh[1], bl[1]);
                        mma_tf32(acc[jj][e], al, bh);
                        mma_tf32(acc[jj][e], ah, bl);
                        mma_tf32(acc[jj][e], ah, bh);
                    }
                }
            }
            __syncwarp();      // the next chunk overwrites the slab
        }
#pragma unroll
        for (int jj = 0; jj < CG; ++jj)
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int kk = gq + 8 * half;
                if (kk >= a.K) continue;
#pragma unroll
                for (int nn = 0; nn < 2; ++nn) {
                    const int c = c0 + jj * 32 + 4 * (2 * tq + nn);
                    if (c >= cout) continue;
                    const int i = 2 * half + nn;
                    *(float4*)(G + ((size_t)j * a.K + kk) * cout + c) =
                        make_float4(acc[jj][0][i], acc[jj][1][i], acc[jj][2][i], acc[jj][3][i]);
                }
            }
    }
}

int kp2_warps_per_cta(int HP, size_t* smem) {
    const size_t per = kp2_warp_floats(HP) * sizeof(float);
    int warps = 8;
    while (warps > 1 && per * warps > 160 * 1024) warps >>= 1;
    *smem = per * warps;
    return warps;
}

template <typename Kern>
int kp2_set_smem(Kern kern, size_t smem) {
    if (smem > 48 * 1024)
        D3F_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return D3F_OK;
}

template <bool IDX64, bool DEF, int MODE>
int correlate_cg(const Kp2Args& a, int HP, int grid, int warps, size_t smem, float* wf, float* wf_unmod, float* inv_n,
                 float* min_d2, cudaStream_t stream) {
    const int cg = a.cin <= 32 ? 1 : (a.cin <= 64 ? 2 : 4);
    const int chunks = (a.cin + 32 * cg - 1) / (32 * cg);
    const dim3 grid2(grid, (chunks > 1 && grid < 2 * 148) ? chunks : 1);
#define KP2_GO(CG_)                                                                                       \
    do {                                                                                                  \
        auto kern = kp2_correlate_kernel<IDX64, DEF, CG_, MODE>;                                          \
        int rc_ = kp2_set_smem(kern, smem);                                                               \
        if (rc_) return rc_;                                                                              \
        kern<<<grid2, warps * 32, smem, stream>>>(a, HP, wf, wf_unmod, inv_n, min_d2);                    \
    } while (0)
    if (cg == 1) KP2_GO(1); else if (cg == 2) KP2_GO(2); else KP2_GO(4);
#undef KP2_GO
    D3F_CHECK_LAUNCH();
    return D3F_OK;
}

template <bool IDX64, bool DEF>
int scatter_cg(const Kp2Args& a, int HP, int grid, int warps, size_t smem, const float* dwf, const float* wf_unmod,
               float* grad_x, float* grad_kp, float* grad_mod, cudaStream_t stream) {
    const int cg = a.cin <= 32 ? 1 : (a.cin <= 64 ? 2 : 4);
    // few CTAs and several 128-channel chunks: one chunk per blockIdx.y (see the kernel)
    const int chunks = (a.cin + 32 * cg - 1) / (32 * cg);
    const int gy = (chunks > 1 && grid < 2 * 148) ? chunks : 1;
    if (gy > 1) {
        if (grad_kp) D3F_CHECK_CUDA(cudaMemsetAsync(grad_kp, 0, sizeof(float) * (size_t)a.nq * a.K * 3, stream));
        if (grad_mod) D3F_CHECK_CUDA(cudaMemsetAsync(grad_mod, 0, sizeof(float) * (size_t)a.nq * a.K, stream));
    }
    const dim3 grid2(grid, gy);
#define KP2_GO(CG_)                                                                                       \
    do {                                                                                                  \
        auto kern = kp2_scatter_kernel<IDX64, DEF, CG_>;                                                  \
        int rc_ = kp2_set_smem(kern, smem);                                                               \
        if (rc_) return rc_;                                                                              \
        kern<<<grid2, warps * 32, smem, stream>>>(a, HP, dwf, wf_unmod, grad_x, grad_kp, grad_mod);       \
    } while (0)
    if (cg == 1) KP2_GO(1); else if (cg == 2) KP2_GO(2); else KP2_GO(4);
#undef KP2_GO
    D3F_CHECK_LAUNCH();
    return D3F_OK;
}

template <bool IDX64>
int scatter_vec(const Kp2Args& a, int HP, int grid, int warps, size_t smem, const float* dwf, float* grad_x,
                cudaStream_t stream) {
#define KP2_GO(CV_)                                                                                       \
    do {                                                                                                  \
        auto kern = kp2_scatter_vec_kernel<IDX64, CV_>;                                                   \
        int rc_ = kp2_set_smem(kern, smem);                                                               \
        if (rc_) return rc_;                                                                              \
        kern<<<grid, warps * 32, smem, stream>>>(a, HP, dwf, grad_x);                                     \
    } while (0)
    if (a.cin == 32) KP2_GO(1); else if (a.cin == 64) KP2_GO(2); else KP2_GO(4);
#undef KP2_GO
    D3F_CHECK_LAUNCH();
    return D3F_OK;
}

}  // namespace

bool kp2_supported(int H, int ns, int cin) {
    return kp2_warp_floats(kp2_hpad(H)) * sizeof(float) <= 160 * 1024 && (long long)ns * cin < (1LL << 31);
}

int kp2_correlate_launch(const Kp2Args& a, float* wf, float* wf_unmod, float* inv_n, float* min_d2, int mode,
                         cudaStream_t stream) {
    const int HP = kp2_hpad(a.H);
    size_t smem;
    const int warps = kp2_warps_per_cta(HP, &smem);
    const int grid = d3f_ceil_div(a.nq, warps);
    const bool aligned = (a.cin & 3) == 0 && (((size_t)a.x | (size_t)wf | (size_t)wf_unmod) & 15) == 0;
    const bool mma = mode == 1 && aligned && a.cin >= 8;
    if (a.idx64) {
        if (a.deformed) return mma ? correlate_cg<true, true, 1>(a, HP, grid, warps, smem, wf, wf_unmod, inv_n, min_d2, stream)
                                   : correlate_cg<true, true, 0>(a, HP, grid, warps, smem, wf, wf_unmod, inv_n, min_d2, stream);
        return mma ? correlate_cg<true, false, 1>(a, HP, grid, warps, smem, wf, wf_unmod, inv_n, min_d2, stream)
                   : correlate_cg<true, false, 0>(a, HP, grid, warps, smem, wf, wf_unmod, inv_n, min_d2, stream);
    }
    if (a.deformed) return mma ? correlate_cg<false, true, 1>(a, HP, grid, warps, smem, wf, wf_unmod, inv_n, min_d2, stream)
                               : correlate_cg<false, true, 0>(a, HP, grid, warps, smem, wf, wf_unmod, inv_n, min_d2, stream);
    return mma ? correlate_cg<false, false, 1>(a, HP, grid, warps, smem, wf, wf_unmod, inv_n, min_d2, stream)
               : correlate_cg<false, false, 0>(a, HP, grid, warps, smem, wf, wf_unmod, inv_n, min_d2, stream);
}

int kp2_scatter_launch(const Kp2Args& a, const float* dwf, const float* wf_unmod, float* grad_x, float* grad_kp,
                       float* grad_mod, cudaStream_t stream) {
    const int HP = kp2_hpad(a.H);
    size_t smem;
    const int warps = kp2_warps_per_cta(HP, &smem);
    const int grid = d3f_ceil_div(a.nq, warps);
    // round 1e micro-benchmark: the reduction rate is per 4-byte element, so the vector form only pays where it also
    // saves shared-memory reads (Cin >= 128: 129 vs 131 us at level 2); at Cin = 32 it is slower (386 vs 300 us)
    const int vec_min_cin = 128;
    const bool vec = !a.deformed && grad_x && !grad_kp && !grad_mod && a.cin >= vec_min_cin &&
                     (a.cin == 32 || a.cin == 64 || (a.cin & 127) == 0) && (((size_t)dwf | (size_t)grad_x) & 15) == 0;
    if (vec) return a.idx64 ? scatter_vec<true>(a, HP, grid, warps, smem, dwf, grad_x, stream)
                            : scatter_vec<false>(a, HP, grid, warps, smem, dwf, grad_x, stream);
    // kernel-point gradient alone (the data gradient came from the transposed lists): tensor-path kernel
    if (a.deformed && !grad_x && grad_kp && !grad_mod && !a.mod && (a.cin & 31) == 0 &&
        (((size_t)dwf | (size_t)a.x) & 15) == 0) {
        if (a.idx64) {
            int rc_ = kp2_set_smem(kp2_gradkp_kernel<true>, smem);
            if (rc_) return rc_;
            kp2_gradkp_kernel<true><<<grid, warps * 32, smem, stream>>>(a, HP, dwf, grad_kp);
        } else {
            int rc_ = kp2_set_smem(kp2_gradkp_kernel<false>, smem);
            if (rc_) return rc_;
            kp2_gradkp_kernel<false><<<grid, warps * 32, smem, stream>>>(a, HP, dwf, grad_kp);
        }
        D3F_CHECK_LAUNCH();
        return D3F_OK;
    }
    if (a.idx64) {
        if (a.deformed) return scatter_cg<true, true>(a, HP, grid, warps, smem, dwf, wf_unmod, grad_x, grad_kp, grad_mod, stream);
        return scatter_cg<true, false>(a, HP, grid, warps, smem, dwf, wf_unmod, grad_x, grad_kp, grad_mod, stream);
    }
    if (a.deformed) return scatter_cg<false, true>(a, HP, grid, warps, smem, dwf, wf_unmod, grad_x, grad_kp, grad_mod, stream);
    return scatter_cg<false, false>(a, HP, grid, warps, smem, dwf, wf_unmod, grad_x, grad_kp, grad_mod, stream);
}

bool kp2t_supported(int nq, int cout) { return (cout & 31) == 0 && (long long)nq * cout < (1LL << 31); }

int kp2t_correlate_launch(const Kp2tArgs& a, float* G, cudaStream_t stream) {
    if (a.ns <= 0) return D3F_OK;
    const int warps = 8;
    const size_t smem = kp2t_warp_floats() * sizeof(float) * warps;
    const int grid = d3f_ceil_div(a.ns, warps);
    const int cgt = a.cout <= 32 ? 1 : (a.cout <= 64 ? 2 : 4);
    const int chunks = (a.cout + 32 * cgt - 1) / (32 * cgt);
    const dim3 grid2(grid, (chunks > 1 && grid < 2 * 148) ? chunks : 1);
#define KP2T_GO(CG_)                                                                                      \
    do {                                                                                                  \
        auto kern = kp2t_correlate_kernel<CG_>;                                                           \
        int rc_ = kp2_set_smem(kern, smem);                                                               \
        if (rc_) return rc_;                                                                              \
        kern<<<grid2, warps * 32, smem, stream>>>(a, G);                                                  \
    } while (0)
    if (a.cout <= 32) KP2T_GO(1); else if (a.cout <= 64) KP2T_GO(2); else KP2T_GO(4);
#undef KP2T_GO
    D3F_CHECK_LAUNCH();
    return D3F_OK;
}
