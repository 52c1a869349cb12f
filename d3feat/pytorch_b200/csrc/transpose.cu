// Transposed neighbour lists (CSR) for the atomic-free KPConv backward.
//
// KPConv's gradient with respect to the support features is
//     dx[j, c] = sum over (i, h) with idx[i, h] = j of  sum_k w[i, k, h] * dwf[i, k, c]
// i.e. a sum over the queries i that list support j.  Walking the neighbour matrix row by row (query by query) turns
// it into 45 M float reductions at level 0, and the LSU retires about one reduction lane per 1.3 cycles per SM: 200 us,
// whatever the kernel around it does (round 1e micro-benchmark).  Walking it COLUMN-wise needs, for every support j,
// the list of queries that reference it: this file builds that list once per neighbour matrix (count -> scan -> fill),
// and kp2t_correlate (kpconv2.cu) then computes the backward as a forward gather over it -- no atomics.
#include "common.cuh"

namespace {

template <bool IDX64>
__global__ void nt_count_kernel(const void* __restrict__ inds, long long ld, int nq, int H, int ns, int* __restrict__ cnt) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)nq * H) return;
    const int i = (int)(t / H), h = (int)(t % H);
    const long long j = IDX64 ? ((const long long*)inds)[(size_t)i * ld + h] : (long long)((const int*)inds)[(size_t)i * ld + h];
    if (j >= 0 && j < ns) atomicAdd(&cnt[j], 1);
}

// in place: cnt[0..n) -> exclusive offsets, cnt[n] = total; cursor[0..n) = offsets.  One CTA of 1024 threads; a thread
// owns `chunk` (multiple of 4) consecutive counters and reads them as independent 128-bit loads (the first version
// walked them one dependent load at a time: 18 us per call in the round-1f profile).
__global__ void __launch_bounds__(1024) nt_scan_kernel(int* __restrict__ cnt, int* __restrict__ cursor, int n) {
    __shared__ int warp_sum[32];
    __shared__ int carry_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int chunk = (((n + 1023) / 1024) + 3) & ~3;
    const int i0 = min(n, tid * chunk), i1 = min(n, i0 + chunk);
    const bool vec = (((size_t)cnt) & 15) == 0;
    int local = 0;
    {
        int i = i0;
        if (vec) {
            int4 acc = make_int4(0, 0, 0, 0);
#pragma unroll 4
            for (; i + 3 < i1; i += 4) {
                const int4 v = *(const int4*)(cnt + i);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            local = (acc.x + acc.y) + (acc.z + acc.w);
        }
        for (; i < i1; ++i) local += cnt[i];
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
        warp_sum[lane] = wi - ws;          // exclusive prefix of the warp totals
        if (lane == 31) carry_s = wi;      // grand total
    }
    __syncthreads();
    int run = warp_sum[warp] + incl - local;
    int i = i0;
    if (vec && (((size_t)cursor) & 15) == 0) {
        for (; i + 3 < i1; i += 4) {
            const int4 v = *(const int4*)(cnt + i);
            const int4 o = make_int4(run, run + v.x, run + v.x + v.y, run + v.x + v.y + v.z);
            run = o.w + v.w;
            *(int4*)(cnt + i) = o;
            *(int4*)(cursor + i) = o;
        }
    }
    for (; i < i1; ++i) {
        const int c = cnt[i];
        cnt[i] = run; cursor[i] = run;
        run += c;
    }
    if (tid == 0) cnt[n] = carry_s;        // index n belongs to no chunk (i1 <= n)
}

template <bool IDX64>
__global__ void nt_fill_kernel(const void* __restrict__ inds, long long ld, int nq, int H, int ns, int* __restrict__ cursor,
                               int* __restrict__ src) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)nq * H) return;
    const int i = (int)(t / H), h = (int)(t % H);
    const long long j = IDX64 ? ((const long long*)inds)[(size_t)i * ld + h] : (long long)((const int*)inds)[(size_t)i * ld + h];
    if (j >= 0 && j < ns) src[atomicAdd(&cursor[j], 1)] = i;
}

// The fill above places entries with atomic cursors, so the order INSIDE a list varies run to run; the backward gather
// sums rows in list order, i.e. grad_x would not be bit-reproducible.  One warp per list puts it into
// ascending query order: an entry's rank is the number of smaller entries (ties, i.e. a row that lists a support twice,
// broken by position) -- counted from a shared-memory copy (lists hold ~H entries; longer ones are ranked from global
// memory into the same place after a copy-out through registers in chunks).
constexpr int NT_SORT_CAP = 256;     // entries per warp in shared memory
__global__ void __launch_bounds__(256)
nt_sort_kernel(const int* __restrict__ off, int ns, int* __restrict__ src) {
    __shared__ int buf[8][NT_SORT_CAP];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * 8 + warp;
    if (j >= ns) return;
    const int b = off[j], n = off[j + 1] - b;
    if (n <= 1) return;
    if (n <= NT_SORT_CAP) {
        int* s = buf[warp];
        for (int t = lane; t < n; t += 32) s[t] = src[b + t];
        __syncwarp();
        for (int t = lane; t < n; t += 32) {
            const int v = s[t];
            int r = 0;
            for (int u = 0; u < n; ++u) r += (s[u] < v) | ((s[u] == v) & (u < t));   // (a row that lists a support twice)
            src[b + r] = v;
        }
        return;
    }
    // long list (in-degree above 256: only wide deformable-radius matrices): odd-even transposition in place
    for (int pass = 0; pass < n; ++pass) {
        for (int t = 2 * lane + (pass & 1); t + 1 < n; t += 64) {
            const int a = src[b + t], c = src[b + t + 1];
            if (a > c) { src[b + t] = c; src[b + t + 1] = a; }
        }
        __syncwarp();
    }
}

}  // namespace

extern "C" size_t d3f_neighbors_transpose_workspace_bytes(int n_supports) {
    return d3f_align(sizeof(int) * (size_t)(n_supports > 0 ? n_supports : 1));
}

extern "C" int d3f_neighbors_transpose(const void* inds, int idx_is_64, int64_t ld_inds, int nq, int ns, int H,
                                       int32_t* t_offsets, int32_t* t_src, void* workspace, size_t workspace_bytes,
                                       d3f_stream stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    D3F_REQUIRE(nq >= 0 && ns >= 0 && H >= 0 && t_offsets, D3F_ERR_INVALID, "bad arguments");
    D3F_REQUIRE(workspace && workspace_bytes >= d3f_neighbors_transpose_workspace_bytes(ns), D3F_ERR_WORKSPACE,
                "workspace too small");
    D3F_CHECK_CUDA(cudaMemsetAsync(t_offsets, 0, sizeof(int32_t) * ((size_t)ns + 1), stream));
    const long long total = (long long)nq * H;
    if (total == 0 || ns == 0) return D3F_OK;
    D3F_REQUIRE(inds && t_src, D3F_ERR_INVALID, "null pointer");
    D3F_REQUIRE(total < (1LL << 31), D3F_ERR_UNSUPPORTED, "neighbour matrix too large for 32-bit offsets");
    int* cursor = (int*)workspace;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    if (idx_is_64) nt_count_kernel<true><<<blocks, 256, 0, stream>>>(inds, (long long)ld_inds, nq, H, ns, t_offsets);
    else nt_count_kernel<false><<<blocks, 256, 0, stream>>>(inds, (long long)ld_inds, nq, H, ns, t_offsets);
    D3F_CHECK_LAUNCH();
    nt_scan_kernel<<<1, 1024, 0, stream>>>(t_offsets, cursor, ns);
    D3F_CHECK_LAUNCH();
    if (idx_is_64) nt_fill_kernel<true><<<blocks, 256, 0, stream>>>(inds, (long long)ld_inds, nq, H, ns, cursor, t_src);
    else nt_fill_kernel<false><<<blocks, 256, 0, stream>>>(inds, (long long)ld_inds, nq, H, ns, cursor, t_src);
    D3F_CHECK_LAUNCH();
    nt_sort_kernel<<<d3f_ceil_div(ns, 8), 256, 0, stream>>>(t_offsets, ns, t_src);   // deterministic list order
    D3F_CHECK_LAUNCH();
    return D3F_OK;
}
