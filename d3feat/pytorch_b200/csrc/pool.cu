// Neighbourhood pooling / row gather / detection score kernels: the gathers that sit between the
// KPConv layers.  The reference issues them as ATen advanced indexing
// (models/blocks.py:79-110, models/architectures.py:322-368) whose backward (sort-based index_put)
// dominates a training step on the GPU; here every op is one warp-per-row kernel, forward and backward.
//
//   max_pool      out[i,c] = max_h xpad[inds[i,h], c]     (shadow row = 0, blocks.py:94-110)
//   gather_rows   out[m,:] = xpad[idx[m], :]               (closest_pool blocks.py:79-91, row selects)
//   det_scores    D3Feat keypoint score (architectures.py:322-368), train and eval mode
#include "common.cuh"

namespace {

template <bool IDX64>
__device__ __forceinline__ long long load_idx(const void* p, size_t off) {
    return IDX64 ? ((const long long*)p)[off] : (long long)((const int*)p)[off];
}

// ------------------------------------------------------------------------------------------- max pool
// One warp per (query, 128-channel chunk): the deep levels have few queries and many channels (256 x 1024 at level 4),
// where a warp per query walked 8 chunks x 47 neighbours of dependent index -> row loads (97 us for 1 MB in the
// round-1d profile).  The neighbour indices of a row are loaded once, 32 at a time, and broadcast by shuffle; four
// row loads are in flight per lane.  Neighbours are visited in column order with a strict '>' (first maximum wins).
template <bool IDX64>
__global__ void mp_forward_kernel(const float* __restrict__ x, const void* __restrict__ inds, long long ld, int nq,
                                  int ns, int H, int C, int nchunks, float* __restrict__ out, int* __restrict__ arg,
                                  const int* __restrict__ valid_width) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= nq * nchunks) return;
    const int qi = w / nchunks, c = (w % nchunks) * 128 + lane * 4;
    // columns >= *valid_width do not exist in the reference's matrix (its width is min(max_count, limit)): they
    // must not contribute the zero shadow row to the max
    if (valid_width) H = min(H, max(*valid_width, 0));
    const bool vec = c + 3 < C && (C & 3) == 0;
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    int who[4] = {-1, -1, -1, -1};
    for (int hb = 0; hb < H; hb += 32) {
        const int cnt = min(32, H - hb);
        long long mine = -1;
        if (lane < cnt) mine = load_idx<IDX64>(inds, (size_t)qi * ld + hb + lane);
        for (int j0 = 0; j0 < cnt; j0 += 4) {
            long long id[4];
            float v[4][4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                id[u] = __shfl_sync(0xffffffffu, mine, min(j0 + u, 31));
                const bool real = j0 + u < cnt && id[u] >= 0 && id[u] < ns;
                if (!real) id[u] = -1;
                v[u][0] = v[u][1] = v[u][2] = v[u][3] = 0.f;
                if (real && c < C) {
                    if (vec) {
                        const float4 t = *(const float4*)&x[(size_t)id[u] * C + c];
                        v[u][0] = t.x; v[u][1] = t.y; v[u][2] = t.z; v[u][3] = t.w;
                    } else {
                        for (int e = 0; e < 4; ++e) if (c + e < C) v[u][e] = x[(size_t)id[u] * C + c + e];
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (j0 + u < cnt) {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (v[u][e] > best[e]) { best[e] = v[u][e]; who[e] = (int)id[u]; }
                }
        }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e)
        if (c + e < C) {
            out[(size_t)qi * C + c + e] = H > 0 ? best[e] : 0.f;
            arg[(size_t)qi * C + c + e] = who[e];
        }
}

__global__ void mp_backward_kernel(const float* __restrict__ g, const int* __restrict__ arg, size_t total, int C,
                                   float* __restrict__ gx) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int a = arg[t];
    if (a >= 0) atomicAdd(&gx[(size_t)a * C + (t % C)], g[t]);
}

// ------------------------------------------------------------------------------------------- row gather
template <bool IDX64>
__global__ void gr_forward_kernel(const float* __restrict__ x, const void* __restrict__ idx, long long stride, int m,
                                  int ns, int C, float* __restrict__ out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= m) return;
    const long long i = load_idx<IDX64>(idx, (size_t)warp * stride);
    const bool real = i >= 0 && i < ns;
    for (int c = lane; c < C; c += 32) out[(size_t)warp * C + c] = real ? x[(size_t)i * C + c] : 0.f;
}

template <bool IDX64>
__global__ void gr_backward_kernel(const float* __restrict__ g, const void* __restrict__ idx, long long stride, int m,
                                   int ns, int C, float* __restrict__ gx) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= m) return;
    const long long i = load_idx<IDX64>(idx, (size_t)warp * stride);
    if (i < 0 || i >= ns) return;
    for (int c = lane; c < C; c += 32) atomicAdd(&gx[(size_t)i * C + c], g[(size_t)warp * C + c]);
}

// ------------------------------------------------------------------------------------------- detection scores
// global max over F and the zero shadow row; also its flat argmax (first occurrence) for the backward
__global__ void ds_gmax_kernel(const float* __restrict__ F, size_t total, unsigned long long* __restrict__ packed) {
    // packed = (ordered float bits << 32) | (0xFFFFFFFF - index): atomicMax gives max value, then smallest index
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long best = 0;
    for (; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const float v = F[t];
        unsigned int b = __float_as_uint(v);
        b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
        const unsigned long long k = ((unsigned long long)b << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned int)t);
        best = k > best ? k : best;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long ob = __shfl_xor_sync(0xffffffffu, best, o);
        best = ob > best ? ob : best;
    }
    if ((threadIdx.x & 31) == 0) atomicMax(packed, best);
}

__device__ __forceinline__ float ds_unpack_max(unsigned long long packed, long long* arg) {
    unsigned int b = (unsigned int)(packed >> 32);
    b = (b & 0x80000000u) ? (b & 0x7FFFFFFFu) : ~b;
    float g = __uint_as_float(b);
    *arg = (long long)(0xFFFFFFFFu - (unsigned int)(packed & 0xFFFFFFFFu));
    if (!(g > 0.0f)) { g = 0.0f; *arg = -1; }  // the appended zero shadow row wins (architectures.py:330-331)
    return g;
}

__device__ __forceinline__ float ds_softplus(float v) { return v > 20.0f ? v : log1pf(expf(v)); }
__device__ __forceinline__ float ds_sigmoid(float v) { return v > 20.0f ? 1.0f : 1.0f / (1.0f + expf(-v)); }

// One warp per point; lane = channel (C <= 32).  Forward writes score[i]; backward (gscore != null) writes
// the direct part of dL/dF and accumulates dL/dgmax.
template <bool IDX64, bool BACKWARD>
__global__ void ds_kernel(const float* __restrict__ F, const void* __restrict__ nb, long long ld, int n, int H, int C,
                          const unsigned long long* __restrict__ packed, int eval_mode, float* __restrict__ score,
                          const float* __restrict__ gscore, float* __restrict__ gF, float* __restrict__ gacc,
                          const int* __restrict__ valid_width) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= n) return;
    if (valid_width) H = min(H, max(*valid_width, 0));
    long long garg;
    const float inv = 1.0f / (ds_unpack_max(*packed, &garg) + 1e-6f);
    float gs = 0.f;
    if (BACKWARD) {
        gs = gscore[i];
        if (gs == 0.0f) return;  // only the selected correspondences carry gradient
    }
    const bool act = lane < C;
    const float f = act ? F[(size_t)i * C + lane] * inv : 0.f;
    float sum = 0.f, nmax = -INFINITY;
    int cnt = 0;
    for (int h = 0; h < H; ++h) {
        const long long j = load_idx<IDX64>(nb, (size_t)i * ld + h);
        const bool real = j >= 0 && j < n;
        const float v = (real && act) ? F[(size_t)j * C + lane] * inv : 0.f;
        float rs = v;
        for (int o = 16; o > 0; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
        cnt += rs != 0.0f;
        sum += v;
        nmax = fmaxf(nmax, v);
    }
    const float cn = (float)max(cnt, 1);
    const float mean = sum / cn;
    const float local = ds_softplus(f - mean);
    float rmax = act ? f : -INFINITY;
    for (int o = 16; o > 0; o >>= 1) rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
    const float den = 1e-6f + rmax;
    const float depth = f / den;
    float s = act ? local * depth : -INFINITY;
    // max over channels with first-index argmax
    float best = s; int bc = lane;
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
        if (ob > best || (ob == best && oc < bc)) { best = ob; bc = oc; }
    }
    float gate = 1.0f;
    if (eval_mode) {
        const unsigned any = __ballot_sync(0xffffffffu, act && H > 0 && f == nmax);
        gate = any ? 1.0f : 0.0f;
    }
    if (!BACKWARD) {
        if (lane == 0) score[i] = best * gate;
        return;
    }
    // ---- backward: score = local[c*] * depth[c*] (* gate, constant)
    const float g = gs * gate;
    // values at the winning channel, broadcast
    const float local_s = __shfl_sync(0xffffffffu, local, bc);
    const float depth_s = __shfl_sync(0xffffffffu, depth, bc);
    const float f_s = __shfl_sync(0xffffffffu, f, bc);
    const float mean_s = __shfl_sync(0xffffffffu, mean, bc);
    const float dlocal = g * depth_s;                 // dL/dlocal[c*]
    const float ddepth = g * local_s;                 // dL/ddepth[c*]
    const float sg = ds_sigmoid(f_s - mean_s);
    // d f[c*] : local via softplus, depth numerator;  d rmax : depth denominator;  d mean[c*]
    const float df_cs = dlocal * sg + ddepth / den;
    const float drmax = -ddepth * f_s / (den * den);
    const float dmean = -dlocal * sg;
    // row argmax of f (first index) receives drmax
    float rbest = act ? f : -INFINITY; int rc = lane;
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, rbest, o);
        const int oc = __shfl_xor_sync(0xffffffffu, rc, o);
        if (ob > rbest || (ob == rbest && oc < rc)) { rbest = ob; rc = oc; }
    }
    // dL/df for this row (normalised features); convert to dL/dF = df * inv and accumulate sum(df * f) for gmax
    float df_own = 0.f;
    if (lane == bc) df_own += df_cs;
    if (lane == rc) df_own += drmax;
    float acc = df_own * f;
    if (act && df_own != 0.0f) atomicAdd(&gF[(size_t)i * C + lane], df_own * inv);
    // neighbours: mean[c*] = sum_h nf[h, c*] / cn  -> each real neighbour gets dmean / cn at channel c*
    const float dn = dmean / cn;
    if (dn != 0.0f) {
        for (int h0 = 0; h0 < H; h0 += 32) {
            const int h = h0 + lane;
            if (h < H) {
                const long long j = load_idx<IDX64>(nb, (size_t)i * ld + h);
                if (j >= 0 && j < n) {
                    atomicAdd(&gF[(size_t)j * C + bc], dn * inv);
                    acc += dn * (F[(size_t)j * C + bc] * inv);
                }
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    // f = F / (gmax + eps): dL/dgmax = -sum(df * f) / (gmax + eps)
    if (lane == 0 && acc != 0.0f) atomicAdd(gacc, -acc * inv);
}

__global__ void ds_gmax_backward_kernel(const unsigned long long* __restrict__ packed, const float* __restrict__ gacc,
                                        float* __restrict__ gF) {
    long long garg;
    ds_unpack_max(*packed, &garg);
    if (garg >= 0) gF[garg] += *gacc;
}

// out[n] = sum_m x[m, n]   (bias gradients of the fused UnaryBlock; blocks.py:473 / nn.Linear bias)
__global__ void colsum_kernel(const float* __restrict__ x, int M, int N, int rows_per_cta, float* __restrict__ out) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int m0 = blockIdx.y * rows_per_cta, m1 = min(M, m0 + rows_per_cta);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    int m = m0;
    for (; m + 3 < m1; m += 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] += x[(size_t)(m + j) * N + n];
    }
    for (; m < m1; ++m) acc[0] += x[(size_t)m * N + n];
    atomicAdd(&out[n], (acc[0] + acc[1]) + (acc[2] + acc[3]));
}

// Vector version for N = 4 * 2^j (every channel count of the network): the CTA's 256 threads tile [R rows x N/4 float4
// columns], so a warp reads 512 contiguous bytes per load and every lane is busy even at N = 32 (the scalar kernel
// above keeps 32 of 128 threads busy there: 25 us for 5 MB in the round-1d profile); 4 independent loads in flight per
// thread, a shared-memory reduction over the R row groups, one atomicAdd per column per CTA.
// FUSE: x is the incoming gradient; dz = x * (y > 0 ? 1 : slope) is written out and its columns are summed (LeakyReLU
// backward + bias gradient of the fused blocks in one pass over the gradient).
template <bool FUSE>
__global__ void __launch_bounds__(256)
colsum_vec_kernel(const float4* __restrict__ x, int M, int nv, int nv_cta, int rows_per_cta, float* __restrict__ out,
                  const float4* __restrict__ y, float slope, float4* __restrict__ dz) {
    __shared__ float4 red[256];
    const int tid = threadIdx.x;
    const int cv = tid % nv_cta, rsub = tid / nv_cta, R = 256 / nv_cta;
    const int col = blockIdx.x * nv_cta + cv;                     // float4 column
    const int m0 = blockIdx.y * rows_per_cta, m1 = min(M, m0 + rows_per_cta);
    auto ld = [&](int m) {
        float4 v = x[(size_t)m * nv + col];
        if (FUSE) {
            const float4 r = y[(size_t)m * nv + col];
            v.x *= r.x > 0.f ? 1.0f : slope; v.y *= r.y > 0.f ? 1.0f : slope;
            v.z *= r.z > 0.f ? 1.0f : slope; v.w *= r.w > 0.f ? 1.0f : slope;
            dz[(size_t)m * nv + col] = v;
        }
        return v;
    };
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
    int m = m0 + rsub;
    for (; m + 3 * R < m1; m += 4 * R) {
        const float4 v0 = ld(m), v1 = ld(m + R), v2 = ld(m + 2 * R), v3 = ld(m + 3 * R);
        a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
        a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
        a2.x += v2.x; a2.y += v2.y; a2.z += v2.z; a2.w += v2.w;
        a3.x += v3.x; a3.y += v3.y; a3.z += v3.z; a3.w += v3.w;
    }
    for (; m < m1; m += R) {
        const float4 v0 = ld(m);
        a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
    }
    red[tid] = make_float4((a0.x + a1.x) + (a2.x + a3.x), (a0.y + a1.y) + (a2.y + a3.y),
                           (a0.z + a1.z) + (a2.z + a3.z), (a0.w + a1.w) + (a2.w + a3.w));
    __syncthreads();
    if (rsub == 0) {
        float4 t = red[cv];
        for (int r = 1; r < R; ++r) {
            const float4 v = red[r * nv_cta + cv];
            t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
        float* o = out + 4 * (size_t)col;
        atomicAdd(o, t.x); atomicAdd(o + 1, t.y); atomicAdd(o + 2, t.z); atomicAdd(o + 3, t.w);
    }
}

// shared launch geometry of the vector kernels; false if the shape needs the scalar path
static bool colsum_vec_geometry(int n_rows, int n_cols, dim3* grid, int* nv, int* nv_cta, int* rpc) {
    *nv = n_cols / 4;
    if ((n_cols & 3) != 0 || (*nv & (*nv - 1)) != 0) return false;
    *nv_cta = *nv < 256 ? *nv : 256;
    const int R = 256 / *nv_cta, col_ctas = *nv / *nv_cta;
    int row_ctas = d3f_ceil_div(444, col_ctas);              // ~3 CTAs per SM
    *rpc = d3f_ceil_div(n_rows, row_ctas);
    if (*rpc < 8 * R) *rpc = 8 * R;
    *rpc = d3f_ceil_div(*rpc, R) * R;
    row_ctas = d3f_ceil_div(n_rows, *rpc);
    *grid = dim3(col_ctas, row_ctas);
    return true;
}

}  // namespace

static int colsum_impl(const float* x, int n_rows, int n_cols, float* out, int prezeroed, d3f_stream stream_);
extern "C" int d3f_colsum(const float* x, int n_rows, int n_cols, float* out, d3f_stream stream) {
    return colsum_impl(x, n_rows, n_cols, out, 0, stream);
}
// `out` already cleared by the caller (gradient buffer zeroed once per step): no zero fill here
extern "C" int d3f_colsum_prezeroed(const float* x, int n_rows, int n_cols, float* out, d3f_stream stream) {
    return colsum_impl(x, n_rows, n_cols, out, 1, stream);
}
static int colsum_impl(const float* x, int n_rows, int n_cols, float* out, int prezeroed, d3f_stream stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    D3F_REQUIRE(n_rows >= 0 && n_cols >= 1 && out, D3F_ERR_INVALID, "bad arguments");
    if (!prezeroed) D3F_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * n_cols, stream));
    if (n_rows == 0) return D3F_OK;
    D3F_REQUIRE(x, D3F_ERR_INVALID, "null pointer");
    dim3 grid;
    int nv, nv_cta, rpc;
    if ((((size_t)x) & 15) == 0 && colsum_vec_geometry(n_rows, n_cols, &grid, &nv, &nv_cta, &rpc)) {
        colsum_vec_kernel<false><<<grid, 256, 0, stream>>>((const float4*)x, n_rows, nv, nv_cta, rpc, out, nullptr, 0.f, nullptr);
        D3F_CHECK_LAUNCH();
        return D3F_OK;
    }
    const int col_ctas = d3f_ceil_div(n_cols, 128);
    int row_ctas = d3f_ceil_div(592, col_ctas);                  // ~4 CTAs per SM
    rpc = d3f_ceil_div(n_rows, row_ctas);
    if (rpc < 32) rpc = 32;
    row_ctas = d3f_ceil_div(n_rows, rpc);
    colsum_kernel<<<dim3(col_ctas, row_ctas), 128, 0, stream>>>(x, n_rows, n_cols, rpc, out);
    D3F_CHECK_LAUNCH();
    return D3F_OK;
}

// dz = grad * (y > 0 ? 1 : slope) and colsum[n] = sum_m dz[m, n] in one pass (y = the saved LeakyReLU OUTPUT).
// Returns D3F_ERR_UNSUPPORTED for shapes the vector kernel does not take (n_cols must be 4 * 2^j, 16-byte aligned rows);
// the caller then uses the two separate ops.
static int leaky_colsum_impl(const float* grad, const float* y, float slope, int n_rows, int n_cols, float* dz,
                             float* colsum, int prezeroed, d3f_stream stream_);
extern "C" int d3f_leaky_backward_colsum(const float* grad, const float* y, float slope, int n_rows, int n_cols, float* dz,
                                         float* colsum, d3f_stream stream) {
    return leaky_colsum_impl(grad, y, slope, n_rows, n_cols, dz, colsum, 0, stream);
}
extern "C" int d3f_leaky_backward_colsum_prezeroed(const float* grad, const float* y, float slope, int n_rows, int n_cols,
                                                   float* dz, float* colsum, d3f_stream stream) {
    return leaky_colsum_impl(grad, y, slope, n_rows, n_cols, dz, colsum, 1, stream);
}
static int leaky_colsum_impl(const float* grad, const float* y, float slope, int n_rows, int n_cols, float* dz,
                             float* colsum, int prezeroed, d3f_stream stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    D3F_REQUIRE(n_rows >= 0 && n_cols >= 1 && colsum, D3F_ERR_INVALID, "bad arguments");
    dim3 grid;
    int nv, nv_cta, rpc;
    if ((((size_t)grad | (size_t)y | (size_t)dz) & 15) != 0 || !colsum_vec_geometry(n_rows, n_cols, &grid, &nv, &nv_cta, &rpc))
        return D3F_ERR_UNSUPPORTED;
    if (!prezeroed) D3F_CHECK_CUDA(cudaMemsetAsync(colsum, 0, sizeof(float) * n_cols, stream));
    if (n_rows == 0) return D3F_OK;
    D3F_REQUIRE(grad && y && dz, D3F_ERR_INVALID, "null pointer");
    colsum_vec_kernel<true><<<grid, 256, 0, stream>>>((const float4*)grad, n_rows, nv, nv_cta, rpc, colsum, (const float4*)y,
                                                     slope, (float4*)dz);
    D3F_CHECK_LAUNCH();
    return D3F_OK;
}

extern "C" int d3f_max_pool_forward(const float* x, const void* inds, int idx_is_64, int64_t ld_inds, int n_queries,
                                    int n_supports, int n_neighbors, int channels, const int32_t* valid_width,
                                    float* out, int32_t* argmax, d3f_stream stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    D3F_REQUIRE(n_queries >= 0 && n_supports >= 0 && n_neighbors >= 0 && channels >= 1, D3F_ERR_INVALID, "bad sizes");
    if (n_queries == 0) return D3F_OK;
    D3F_REQUIRE(x && (inds || n_neighbors == 0) && out && argmax, D3F_ERR_INVALID, "null pointer");
    auto kern = idx_is_64 ? mp_forward_kernel<true> : mp_forward_kernel<false>;
    const int nchunks = d3f_ceil_div(channels, 128);
    const long long warps = (long long)n_queries * nchunks;
    D3F_REQUIRE(warps < (1LL << 28), D3F_ERR_UNSUPPORTED, "max_pool: too many (query, channel chunk) pairs");
    kern<<<(unsigned)((warps + 7) / 8), 256, 0, stream>>>(x, inds, (long long)ld_inds, n_queries, n_supports, n_neighbors,
                                                         channels, nchunks, out, argmax, valid_width);
    D3F_CHECK_LAUNCH();
    return D3F_OK;
}

extern "C" int d3f_max_pool_backward(const float* grad_out, const int32_t* argmax, int n_queries, int n_supports,
                                     int channels, float* grad_x, d3f_stream stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    D3F_REQUIRE(n_queries >= 0 && n_supports >= 0 && channels >= 1, D3F_ERR_INVALID, "bad sizes");
    D3F_REQUIRE(grad_x || n_supports == 0, D3F_ERR_INVALID, "null pointer");
    if (n_supports > 0) D3F_CHECK_CUDA(cudaMemsetAsync(grad_x, 0, sizeof(float) * (size_t)n_supports * channels, stream));
    if (n_queries == 0 || n_supports == 0) return D3F_OK;
    D3F_REQUIRE(grad_out && argmax, D3F_ERR_INVALID, "null pointer");
    const size_t total = (size_t)n_queries * channels;
    mp_backward_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(grad_out, argmax, total, channels, grad_x);
    D3F_CHECK_LAUNCH();
    return D3F_OK;
}

extern "C" int d3f_gather_rows_forward(const float* x, const void* idx, int idx_is_64, int64_t idx_stride, int n_rows,
                                       int n_supports, int channels, float* out, d3f_stream stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    D3F_REQUIRE(n_rows >= 0 && n_supports >= 0 && channels >= 1, D3F_ERR_INVALID, "bad sizes");
    if (n_rows == 0) return D3F_OK;
    D3F_REQUIRE(x && idx && out, D3F_ERR_INVALID, "null pointer");
    auto kern = idx_is_64 ? gr_forward_kernel<true> : gr_forward_kernel<false>;
    kern<<<d3f_ceil_div(n_rows, 8), 256, 0, stream>>>(x, idx, (long long)idx_stride, n_rows, n_supports, channels, out);
    D3F_CHECK_LAUNCH();
    return D3F_OK;
}

extern "C" int d3f_gather_rows_backward(const float* grad_out, const void* idx, int idx_is_64, int64_t idx_stride,
                                        int n_rows, int n_supports, int channels, float* grad_x, d3f_stream stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    D3F_REQUIRE(n_rows >= 0 && n_supports >= 0 && channels >= 1, D3F_ERR_INVALID, "bad sizes");
    if (n_supports > 0) {
        D3F_REQUIRE(grad_x, D3F_ERR_INVALID, "null pointer");
        D3F_CHECK_CUDA(cudaMemsetAsync(grad_x, 0, sizeof(float) * (size_t)n_supports * channels, stream));
    }
    if (n_rows == 0 || n_supports == 0) return D3F_OK;
    D3F_REQUIRE(grad_out && idx, D3F_ERR_INVALID, "null pointer");
    auto kern = idx_is_64 ? gr_backward_kernel<true> : gr_backward_kernel<false>;
    kern<<<d3f_ceil_div(n_rows, 8), 256, 0, stream>>>(grad_out, idx, (long long)idx_stride, n_rows, n_supports, channels, grad_x);
    D3F_CHECK_LAUNCH();
    return D3F_OK;
}

extern "C" int d3f_detection_scores_forward(const float* features, const void* neighbors, int idx_is_64,
                                            int64_t ld_inds, int n_points, int n_neighbors, int channels,
                                            int eval_mode, const int32_t* valid_width, float* scores,
                                            void* gmax_state, d3f_stream stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    D3F_REQUIRE(n_points >= 0 && n_neighbors >= 0, D3F_ERR_INVALID, "bad sizes");
    D3F_REQUIRE(channels >= 1 && channels <= 32, D3F_ERR_UNSUPPORTED, "detection scores support up to 32 channels");
    D3F_REQUIRE(gmax_state, D3F_ERR_INVALID, "null gmax_state");
    D3F_CHECK_CUDA(cudaMemsetAsync(gmax_state, 0, 16, stream));
    if (n_points == 0) return D3F_OK;
    D3F_REQUIRE(features && (neighbors || n_neighbors == 0) && scores, D3F_ERR_INVALID, "null pointer");
    D3F_REQUIRE((size_t)n_points * channels < 0xFFFFFFFFull, D3F_ERR_UNSUPPORTED, "too many feature elements");
    ds_gmax_kernel<<<296, 256, 0, stream>>>(features, (size_t)n_points * channels, (unsigned long long*)gmax_state);
    D3F_CHECK_LAUNCH();
    auto kern = idx_is_64 ? ds_kernel<true, false> : ds_kernel<false, false>;
    kern<<<d3f_ceil_div(n_points, 8), 256, 0, stream>>>(features, neighbors, (long long)ld_inds, n_points, n_neighbors,
                                                       channels, (const unsigned long long*)gmax_state, eval_mode,
                                                       scores, nullptr, nullptr, nullptr, valid_width);
    D3F_CHECK_LAUNCH();
    return D3F_OK;
}

extern "C" int d3f_detection_scores_backward(const float* features, const void* neighbors, int idx_is_64,
                                             int64_t ld_inds, int n_points, int n_neighbors, int channels,
                                             int eval_mode, const int32_t* valid_width, const void* gmax_state,
                                             const float* grad_scores, float* grad_features, d3f_stream stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    D3F_REQUIRE(n_points >= 0 && n_neighbors >= 0, D3F_ERR_INVALID, "bad sizes");
    D3F_REQUIRE(channels >= 1 && channels <= 32, D3F_ERR_UNSUPPORTED, "detection scores support up to 32 channels");
    if (n_points == 0) return D3F_OK;
    D3F_REQUIRE(features && (neighbors || n_neighbors == 0) && gmax_state && grad_scores && grad_features,
                D3F_ERR_INVALID, "null pointer");
    D3F_CHECK_CUDA(cudaMemsetAsync(grad_features, 0, sizeof(float) * (size_t)n_points * channels, stream));
    float* gacc = (float*)((char*)gmax_state + 8);  // second 8 bytes of the state: dL/dgmax accumulator
    D3F_CHECK_CUDA(cudaMemsetAsync(gacc, 0, 8, stream));
    auto kern = idx_is_64 ? ds_kernel<true, true> : ds_kernel<false, true>;
    kern<<<d3f_ceil_div(n_points, 8), 256, 0, stream>>>(features, neighbors, (long long)ld_inds, n_points, n_neighbors,
                                                       channels, (const unsigned long long*)gmax_state, eval_mode,
                                                       nullptr, grad_scores, grad_features, gacc, valid_width);
    D3F_CHECK_LAUNCH();
    ds_gmax_backward_kernel<<<1, 1, 0, stream>>>((const unsigned long long*)gmax_state, gacc, grad_features);
    D3F_CHECK_LAUNCH();
    return D3F_OK;
}
