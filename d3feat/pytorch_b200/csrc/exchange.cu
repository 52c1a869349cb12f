// The one data-path exchange of the multi-GPU step: every rank contributes its P selected descriptor
// pairs, scores and keypoint distances; each rank then evaluates the (W*P)^2 cross-fragment loss.  The pack / unpack
// around the NCCL all-gather used to be ~30 small ATen kernels on the critical path between forward and loss (concatenations,
// dtype conversions, an 8 MB fill and W block copies: +0.25 ms on the forward at W = 8); here each side is ONE launch.
//
// chunk layout (per rank, `d3f_exchange_chunk_bytes(P, D)` bytes, 8-byte aligned sections):
//   f32 anchor [P,D] | f32 positive [P,D] | f32 anc_score [P] | f32 pos_score [P] | f64 dist_keypts [P,P]
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
xc_pack_kernel(const float* __restrict__ a, const float* __restrict__ p, const float* __restrict__ sa,
               const float* __restrict__ sp, const void* __restrict__ dk, int dk_is_f64, int P, int D,
               unsigned char* __restrict__ chunk) {
    const size_t pd = (size_t)P * D, n32 = 2 * pd + 2 * (size_t)P, pp = (size_t)P * P;
    float* o32 = (float*)chunk;
    double* o64 = (double*)(chunk + n32 * sizeof(float));
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n32 + pp; i += (size_t)gridDim.x * blockDim.x) {
        if (i < pd) o32[i] = a[i];
        else if (i < 2 * pd) o32[i] = p[i - pd];
        else if (i < 2 * pd + P) o32[i] = sa[i - 2 * pd];
        else if (i < n32) o32[i] = sp[i - 2 * pd - P];
        else o64[i - n32] = dk_is_f64 ? ((const double*)dk)[i - n32] : (double)((const float*)dk)[i - n32];
    }
}

__global__ void __launch_bounds__(256)
xc_unpack_kernel(const unsigned char* __restrict__ all, size_t chunk_bytes, int W, int P, int D, float* __restrict__ A,
                 float* __restrict__ Pos, float* __restrict__ SA, float* __restrict__ SP, double* __restrict__ DK) {
    const size_t pd = (size_t)P * D, n32 = 2 * pd + 2 * (size_t)P, WP = (size_t)W * P;
    const size_t n_small = (size_t)W * n32, n_dk = WP * WP;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_small + n_dk; i += (size_t)gridDim.x * blockDim.x) {
        if (i < n_small) {
            const size_t r = i / n32, e = i % n32;
            const float v = ((const float*)(all + r * chunk_bytes))[e];
            if (e < pd) A[r * pd + e] = v;
            else if (e < 2 * pd) Pos[r * pd + (e - pd)] = v;
            else if (e < 2 * pd + P) SA[r * P + (e - 2 * pd)] = v;
            else SP[r * P + (e - 2 * pd - P)] = v;
        } else {
            // block-diagonal keypoint distances; pairs of different fragments are always valid negatives (+inf)
            const size_t t = i - n_small, row = t / WP, col = t % WP;
            const size_t r = row / P;
            double v = INFINITY;
            if (col / P == r)
                v = ((const double*)(all + r * chunk_bytes + n32 * sizeof(float)))[(row % P) * P + (col % P)];
            DK[t] = v;
        }
    }
}

}  // namespace

extern "C" size_t d3f_exchange_chunk_bytes(int P, int D) {
    return ((size_t)2 * P * D + 2 * (size_t)P) * sizeof(float) + (size_t)P * P * sizeof(double);
}

// anchor, positive [P,D] f32; anc_score, pos_score [P] f32; dist_keypts [P,P] f64 (dk_is_f64) or f32 -> chunk
extern "C" int d3f_exchange_pack(const float* anchor, const float* positive, const float* anc_score, const float* pos_score,
                                 const void* dist_keypts, int dk_is_f64, int P, int D, void* chunk, d3f_stream stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    D3F_REQUIRE(P >= 1 && D >= 1, D3F_ERR_INVALID, "bad sizes");
    D3F_REQUIRE(anchor && positive && anc_score && pos_score && dist_keypts && chunk, D3F_ERR_INVALID, "null pointer");
    D3F_REQUIRE((((size_t)chunk) & 7) == 0, D3F_ERR_INVALID, "chunk must be 8-byte aligned");
    const size_t n = (size_t)2 * P * D + 2 * (size_t)P + (size_t)P * P;
    const size_t blocks = (n + 255) / 256;
    xc_pack_kernel<<<(unsigned)(blocks > 148 * 8 ? 148 * 8 : blocks), 256, 0, stream>>>(
        anchor, positive, anc_score, pos_score, dist_keypts, dk_is_f64, P, D, (unsigned char*)chunk);
    D3F_CHECK_LAUNCH();
    return D3F_OK;
}

// all: W chunks back to back (the all-gather's output) -> A, Pos [W*P, D] f32; SA, SP [W*P] f32; DK [W*P, W*P] f64
extern "C" int d3f_exchange_unpack(const void* all, int W, int P, int D, float* A, float* Pos, float* SA, float* SP,
                                   double* DK, d3f_stream stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    D3F_REQUIRE(W >= 1 && P >= 1 && D >= 1, D3F_ERR_INVALID, "bad sizes");
    D3F_REQUIRE(all && A && Pos && SA && SP && DK, D3F_ERR_INVALID, "null pointer");
    D3F_REQUIRE((((size_t)all) & 7) == 0, D3F_ERR_INVALID, "buffer must be 8-byte aligned");
    const size_t n = (size_t)W * (2 * (size_t)P * D + 2 * (size_t)P) + (size_t)W * P * W * P;
    const size_t blocks = (n + 255) / 256;
    xc_unpack_kernel<<<(unsigned)(blocks > 148 * 16 ? 148 * 16 : blocks), 256, 0, stream>>>(
        (const unsigned char*)all, d3f_exchange_chunk_bytes(P, D), W, P, D, A, Pos, SA, SP, DK);
    D3F_CHECK_LAUNCH();
    return D3F_OK;
}
