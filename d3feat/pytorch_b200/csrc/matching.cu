// Mutual nearest-neighbour matching of descriptors: replaces build_correspondence
// (geometric_registration/common.py:5-21), which the reference runs in numpy on the host for 250-5000 keypoints per
// fragment:
//     distance = sqrt(2 - 2 * (S @ T^T));  i <-> argmin_j distance[i, j]  kept when  argmin_i distance[i, j] == i
// The N x M matrix is never materialised: one warp per row walks the other descriptor set staged through shared memory
// and keeps its running (min, first index).  numpy's argmin returns the FIRST NaN of a row when 2 - 2<s,t> < 0
// (descriptors slightly longer than 1); that rule is kept.  Pairs are emitted in ascending source index.
#include "common.cuh"

namespace {

constexpr int MNN_TILE = 128;   // rows of the other set per shared-memory tile

// arg[i] = numpy.argmin_j sqrt(2 - 2 <a_i, b_j>)   (fp32; first NaN wins, else first minimum)
__global__ void __launch_bounds__(256)
mnn_argmin_kernel(const float* __restrict__ a, const float* __restrict__ b, int na, int nb, int d, int* __restrict__ arg) {
    extern __shared__ float smem[];                     // [MNN_TILE][d + 1] tile of b | [8][d] rows of a
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ldt = d + 1;
    float* tile = smem;
    float* arow = smem + (size_t)MNN_TILE * ldt + (size_t)warp * d;
    const int i = blockIdx.x * 8 + warp;
    if (i < na)
        for (int c = lane; c < d; c += 32) arow[c] = a[(size_t)i * d + c];
    float best = INFINITY;
    int best_j = 0x7fffffff, nan_j = 0x7fffffff;
    for (int j0 = 0; j0 < nb; j0 += MNN_TILE) {
        __syncthreads();
        const int rows = min(MNN_TILE, nb - j0);
        for (int t = threadIdx.x; t < rows * d; t += 256) tile[(t / d) * ldt + (t % d)] = b[(size_t)j0 * d + t];
        __syncthreads();
        if (i < na) {
            for (int jj = lane; jj < rows; jj += 32) {
                const float* br = tile + jj * ldt;
                float dot = 0.f;
                for (int c = 0; c < d; ++c) dot = fmaf(arow[c], br[c], dot);
                const float dist = sqrtf(2.0f - 2.0f * dot);
                const int j = j0 + jj;
                if (dist != dist) { if (j < nan_j) nan_j = j; }
                else if (dist < best) { best = dist; best_j = j; }     // j ascends per lane: first minimum kept
            }
        }
    }
    if (i >= na) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oj = __shfl_xor_sync(0xffffffffu, best_j, o);
        const int on = __shfl_xor_sync(0xffffffffu, nan_j, o);
        if (ob < best || (ob == best && oj < best_j)) { best = ob; best_j = oj; }
        nan_j = min(nan_j, on);
    }
    if (lane == 0) arg[i] = nan_j != 0x7fffffff ? nan_j : (best_j != 0x7fffffff ? best_j : 0);
}

// pairs (i, row_arg[i]) with col_arg[row_arg[i]] == i, ascending i.  One CTA.
__global__ void __launch_bounds__(1024)
mnn_emit_kernel(const int* __restrict__ row_arg, const int* __restrict__ col_arg, int na, int nb, int* __restrict__ pairs,
                int* __restrict__ n_pairs) {
    __shared__ int warp_sum[32];
    __shared__ int total_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int chunk = (na + 1023) / 1024;
    const int i0 = min(na, tid * chunk), i1 = min(na, i0 + chunk);
    int local = 0;
    for (int i = i0; i < i1; ++i) {
        const int j = row_arg[i];
        local += (j >= 0 && j < nb && col_arg[j] == i);
    }
    int incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int ws = warp_sum[lane], wi = ws;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += v;
        }
        warp_sum[lane] = wi - ws;
        if (lane == 31) total_s = wi;
    }
    __syncthreads();
    int pos = warp_sum[warp] + incl - local;
    for (int i = i0; i < i1; ++i) {
        const int j = row_arg[i];
        if (j >= 0 && j < nb && col_arg[j] == i) { pairs[2 * pos] = i; pairs[2 * pos + 1] = j; ++pos; }
    }
    if (tid == 0) *n_pairs = total_s;
}

}  // namespace

extern "C" int d3f_mutual_nn(const float* source, const float* target, int n_source, int n_target, int dim,
                             int32_t* source_arg, int32_t* target_arg, int32_t* pairs, int32_t* n_pairs, d3f_stream stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    D3F_REQUIRE(n_source >= 0 && n_target >= 0 && dim >= 1 && n_pairs, D3F_ERR_INVALID, "bad arguments");
    D3F_CHECK_CUDA(cudaMemsetAsync(n_pairs, 0, sizeof(int32_t), stream));
    if (n_source == 0 || n_target == 0) return D3F_OK;
    D3F_REQUIRE(source && target && source_arg && target_arg && pairs, D3F_ERR_INVALID, "null pointer");
    const size_t smem = sizeof(float) * ((size_t)MNN_TILE * (dim + 1) + 8 * (size_t)dim);
    D3F_REQUIRE(smem <= 200 * 1024, D3F_ERR_UNSUPPORTED, "descriptor dimension too large for one shared-memory tile");
    if (smem > 48 * 1024)
        D3F_CHECK_CUDA(cudaFuncSetAttribute(mnn_argmin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mnn_argmin_kernel<<<d3f_ceil_div(n_source, 8), 256, smem, stream>>>(source, target, n_source, n_target, dim, source_arg);
    D3F_CHECK_LAUNCH();
    mnn_argmin_kernel<<<d3f_ceil_div(n_target, 8), 256, smem, stream>>>(target, source, n_target, n_source, dim, target_arg);
    D3F_CHECK_LAUNCH();
    mnn_emit_kernel<<<1, 1024, 0, stream>>>(source_arg, target_arg, n_source, n_target, pairs, n_pairs);
    D3F_CHECK_LAUNCH();
    return D3F_OK;
}
