// One launch that initialises several scratch regions with (possibly different) 32-bit patterns.
// The pyramid build clears a hash table (0xFF), its counters (0) and an info vector before every search / subsampling:
// as separate cudaMemsetAsync nodes these were 68 of the 126 memset nodes of a captured pair step, each a few
// microseconds of dependency latency at the head of a latency-bound chain.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) fill_segments_kernel(D3fFillSegs s) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    for (int k = 0; k < s.n; ++k) {
        uint32_t* p = (uint32_t*)s.p[k];
        const size_t words = s.words[k];
        const uint32_t v = s.v[k];
        if ((((size_t)p) & 15) == 0) {
            const size_t n4 = words >> 2;
            const uint4 v4 = make_uint4(v, v, v, v);
            for (size_t i = tid; i < n4; i += nth) ((uint4*)p)[i] = v4;
            for (size_t i = (n4 << 2) + tid; i < words; i += nth) p[i] = v;
        } else {
            for (size_t i = tid; i < words; i += nth) p[i] = v;
        }
    }
}

}  // namespace

int d3f_fill_segments(const D3fFillSegs& s, cudaStream_t stream) {
    size_t most = 0;
    for (int k = 0; k < s.n; ++k) {
        D3F_REQUIRE(s.words[k] == 0 || (s.p[k] && (((size_t)s.p[k]) & 3) == 0), D3F_ERR_INVALID, "fill: unaligned segment");
        most = s.words[k] > most ? s.words[k] : most;
    }
    if (most == 0) return D3F_OK;
    const size_t blocks = (most / 4 + 255) / 256;
    fill_segments_kernel<<<(unsigned)(blocks < 1 ? 1 : (blocks > 148 * 8 ? 148 * 8 : blocks)), 256, 0, stream>>>(s);
    D3F_CHECK_LAUNCH();
    return D3F_OK;
}
