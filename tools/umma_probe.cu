// Stand-alone probe (nvcc -gencode arch=compute_100a,code=sm_100a tools/umma_probe.cu -o umma_probe): measures what the
// fused KPConv kernel's design rests on and that no document in the image states:
//   1. cycles per tcgen05.mma with the A operand in TENSOR MEMORY and B in shared memory, as a function of N and kind
//      (tf32 K = 8, bf16 K = 16), issued back to back into one accumulator (the contraction's K loop);
//   2. how 16-bit A elements are packed in tensor-memory columns for kind::f16 (checked against a CPU product).
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ULL << 46);
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (int s = 0; s < (1 << 22) && !done; ++s)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

// KIND 0: tf32 (A: one fp32 per column, K = 8 per MMA); KIND 1: bf16 (A: two bf16 per column, K = 16 per MMA)
template <int KIND>
__global__ void __launch_bounds__(128, 1) probe(int N, int reps, int pack_hi_first, const float* Ain /*[128][K]*/, const float* Bin /*[N][K]*/,
                                                float* Dout /*[128][N]*/, long long* cycles) {
    constexpr int K = KIND ? 16 : 8;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(&tmem_ptr)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = tmem_ptr;
    // ---- A -> tensor memory: thread = row (lane of TMEM), columns 0.. (8 for both kinds)
    {
        uint32_t v[8];
        for (int c = 0; c < 8; ++c) {
            if (KIND == 0) v[c] = __float_as_uint(Ain[tid * K + c]);
            else {
                const uint32_t e0 = __bfloat16_as_ushort(__float2bfloat16_rn(Ain[tid * K + 2 * c]));
                const uint32_t e1 = __bfloat16_as_ushort(__float2bfloat16_rn(Ain[tid * K + 2 * c + 1]));
                v[c] = pack_hi_first ? ((e0 << 16) | e1) : ((e1 << 16) | e0);
            }
        }
        const uint32_t ta = tmem + ((uint32_t)(32 * warp) << 16);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n"
                     :: "r"(ta), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    }
    // ---- B -> shared memory, K-major no swizzle: 16-byte unit = 4 tf32 / 8 bf16 along K
    const int SBO = 128, LBO = (N / 8) * 128 + 16;
    for (int i = tid; i < N * K; i += 128) {
        const int n = i / K, k = i % K;
        if (KIND == 0) {
            const int off = (k / 4) * LBO + (n / 8) * SBO + (n % 8) * 16 + (k % 4) * 4;
            *(float*)(smem + off) = Bin[n * K + k];
        } else {
            const int off = (k / 8) * LBO + (n / 8) * SBO + (n % 8) * 16 + (k % 8) * 2;
            *(__nv_bfloat16*)(smem + off) = __float2bfloat16_rn(Bin[n * K + k]);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t fmt = KIND ? 1u : 2u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t bdesc = make_desc(smem_u32(smem), LBO, SBO);
    const uint32_t dcol = 256;
    long long t0 = 0, t1 = 0, t2 = 0;
    if (warp == 0) {
        t0 = clock64();
        if (lane == 0) {
            for (int r = 0; r < reps; ++r) {
                const uint32_t acc = r ? 1u : 0u;
                if (KIND == 0)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                                 :: "r"(tmem + dcol), "r"(tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                                 :: "r"(tmem + dcol), "r"(tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
            }
            t1 = clock64();
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(smem_u32(&bar)) : "memory");
        }
        __syncwarp();
    }
    mbar_wait(smem_u32(&bar), 0);
    if (tid == 0) { t2 = clock64(); cycles[0] = t1 - t0; cycles[1] = t2 - t0; }
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    // ---- D -> global (row = tid)
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t d[8];
        const uint32_t ta = tmem + ((uint32_t)(32 * warp) << 16) + dcol + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                     : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]) : "r"(ta) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        for (int c = 0; c < 8; ++c) Dout[tid * N + c0 + c] = __uint_as_float(d[c]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem), "r"(512) : "memory");
}

static float bf16r(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

int main() {
    const int MAXN = 256;
    std::vector<float> A(128 * 16), B(MAXN * 16), D(128 * MAXN);
    srand(1);
    for (auto& v : A) v = (rand() % 2001 - 1000) / 1000.0f;
    for (auto& v : B) v = (rand() % 2001 - 1000) / 1000.0f;
    float *dA, *dB, *dD; long long* dC;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dC, 16);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int kind = 0; kind < 2; ++kind) {
        const int K = kind ? 16 : 8;
        // ---- correctness / packing
        for (int pack = 0; pack < (kind ? 2 : 1); ++pack) {
            const int N = 32;
            cudaMemset(dD, 0, D.size() * 4);
            if (kind == 0) probe<0><<<1, 128, 48 * 1024>>>(N, 1, pack, dA, dB, dD, dC);
            else probe<1><<<1, 128, 48 * 1024>>>(N, 1, pack, dA, dB, dD, dC);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("kind %d pack %d: CUDA error %s\n", kind, pack, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(D.data(), dD, 128 * N * 4, cudaMemcpyDeviceToHost);
            double worst = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < N; ++n) {
                    double ref = 0;
                    for (int k = 0; k < K; ++k) {
                        float a = A[m * K + k], b = B[n * K + k];
                        if (kind) { a = bf16r(a); b = bf16r(b); }
                        else { uint32_t ua, ub; memcpy(&ua, &a, 4); memcpy(&ub, &b, 4); ua &= 0xffffe000u; ub &= 0xffffe000u; memcpy(&a, &ua, 4); memcpy(&b, &ub, 4); }
                        ref += (double)a * b;
                    }
                    worst = fmax(worst, fabs(ref - D[m * N + n]));
                }
            printf("kind %s pack(e0 in %s half): max |D - ref| = %.3e  %s\n", kind ? "bf16" : "tf32", pack ? "high" : "low", worst,
                   worst < 1e-3 ? "MATCH" : "mismatch");
        }
        // ---- timing
        for (int N : {16, 32, 64, 96, 128, 192, 256}) {
            long long c[2];
            const int reps = 64;
            for (int rep = 0; rep < 2; ++rep) {
                if (kind == 0) probe<0><<<1, 128, 48 * 1024>>>(N, reps, 0, dA, dB, dD, dC);
                else probe<1><<<1, 128, 48 * 1024>>>(N, reps, 0, dA, dB, dD, dC);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost);
            printf("kind %s N=%3d: %d back-to-back MMAs (A in TMEM): issue %.1f cyc/MMA, issue+complete %.1f cyc/MMA (floor 128*N/256 = %d)\n",
                   kind ? "bf16" : "tf32", N, reps, (double)c[0] / reps, (double)c[1] / reps, 128 * N / 256);
        }
    }
    return 0;
}
