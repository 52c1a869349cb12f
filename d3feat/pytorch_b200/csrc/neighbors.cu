// Radius neighbours on a device hash grid (replaces the nanoflann kd-tree path of
// cpp_wrappers/cpp_neighbors/neighbors/neighbors.cpp:211-332).
//
// Pipeline per call (all on `stream`, no host sync):
//   nb_insert   : every support claims its cell in an open-addressing table (key = batch|cx|cy|cz,
//                 cell edge = 1.01*radius) and takes a rank inside the cell
//   nb_alloc    : every occupied cell gets a contiguous range in the cell-sorted copy
//   nb_scatter  : supports are written cell-contiguously as float4 (x, y, z, index)
//   nb_query    : one warp per query: the 27 surrounding cells are looked up by 27 lanes, each lane
//                 streams its own cell, hits (fp32 d2 < r2, no FMA) are compacted with ballots into
//                 a per-warp shared-memory buffer of 64-bit (d2 bits | index) keys, bitonic-sorted
//                 and the first max_cols indices written (padded with n_supports).
// Set membership and order are exactly those of the reference except inside runs of equal d2,
// where the reference's unstable std::sort (nanoflann.hpp:1286) is replaced by index order.
#include "common.cuh"

namespace {

constexpr int CELL_BIAS = 1 << 17;
constexpr int CELL_MAXC = (1 << 18) - 1;

__device__ __forceinline__ int cell_coord(float v, float inv_cs) {
    float f = floorf(v * inv_cs);
    f = fminf(fmaxf(f, -(float)CELL_BIAS), (float)(CELL_BIAS - 1));
    return (int)f + CELL_BIAS;
}

__device__ __forceinline__ uint64_t cell_key(int b, int cx, int cy, int cz) {
    return ((uint64_t)b << 54) | ((uint64_t)cx << 36) | ((uint64_t)cy << 18) | (uint64_t)cz;
}

__global__ void nb_insert_kernel(const float* __restrict__ s, const int32_t* __restrict__ s_len, int nb,
                                 int ns, float inv_cs, unsigned long long* keys, uint32_t* cnt,
                                 uint32_t mask, uint32_t* slot_of, uint32_t* rank) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ns) return;
    int st;
    int b = d3f_batch_of(i, s_len, nb, &st);
    if (b < 0) return;  // capacity padding
    uint64_t key = cell_key(b, cell_coord(s[3 * i], inv_cs), cell_coord(s[3 * i + 1], inv_cs),
                            cell_coord(s[3 * i + 2], inv_cs));
    uint32_t slot = d3f_hash64(key) & mask;
    while (true) {
        unsigned long long old = atomicCAS(&keys[slot], D3F_EMPTY_KEY, (unsigned long long)key);
        if (old == D3F_EMPTY_KEY || old == key) break;
        slot = (slot + 1) & mask;
    }
    slot_of[i] = slot;
    rank[i] = atomicAdd(&cnt[slot], 1u);
}

__global__ void nb_alloc_kernel(const uint32_t* __restrict__ cnt, uint32_t* start, uint32_t table,
                                uint32_t* cursor /* [0]=points, [1]=cells */) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t c = t < table ? cnt[t] : 0u;
    // warp-aggregated range allocation
    uint32_t lane = threadIdx.x & 31, incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    uint32_t occupied = __popc(__ballot_sync(0xffffffffu, c != 0));
    uint32_t base = 0;
    if (lane == 31 && total) {
        base = atomicAdd(&cursor[0], total);
        atomicAdd(&cursor[1], occupied);
    }
    base = __shfl_sync(0xffffffffu, base, 31);
    if (t < table) start[t] = base + incl - c;
}

__global__ void nb_scatter_kernel(const float* __restrict__ s, int ns, const int32_t* __restrict__ s_len, int nb,
                                  const uint32_t* __restrict__ slot_of,
                                  const uint32_t* __restrict__ rank, const uint32_t* __restrict__ start,
                                  float4* sorted) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ns) return;
    int st;
    if (d3f_batch_of(i, s_len, nb, &st) < 0) return;
    uint32_t dst = start[slot_of[i]] + rank[i];
    sorted[dst] = make_float4(s[3 * i], s[3 * i + 1], s[3 * i + 2], __int_as_float(i));
}

// ascending bitonic sort of cand[0, m) by one warp (entries beyond m up to the next power of two are set to ~0)
__device__ __forceinline__ void nb_warp_sort(unsigned long long* cand, int m, int lane) {
    int n2 = 1;
    while (n2 < m) n2 <<= 1;
    for (int t = m + lane; t < n2; t += 32) cand[t] = ~0ULL;
    __syncwarp();
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < (n2 >> 1); t += 32) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const unsigned long long a = cand[i], c = cand[l];
                const bool up = (i & k) == 0;
                if ((a > c) == up) { cand[i] = c; cand[l] = a; }
            }
            __syncwarp();
        }
    }
}

template <bool IDX64>
__global__ void nb_query_kernel(const float* __restrict__ q, const int32_t* __restrict__ q_len, int nb, int nq,
                                int ns, float inv_cs, float r2, const unsigned long long* __restrict__ keys,
                                const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ start,
                                uint32_t mask, const float4* __restrict__ sorted, int max_cols, void* out,
                                int32_t* info, int cap, int pad_index) {
    extern __shared__ unsigned long long cand_all[];
    __shared__ int s_max, s_ovf;
    if (threadIdx.x == 0) { s_max = 0; s_ovf = 0; }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qi = blockIdx.x * (blockDim.x >> 5) + warp;
    unsigned long long* cand = cand_all + (size_t)warp * cap;
    int count = 0, nbuf = 0;      // in-range supports seen / candidates currently buffered
    const bool can_compact = out != nullptr && max_cols + 32 <= cap;
    int st;
    const int b = qi < nq ? d3f_batch_of(qi, q_len, nb, &st) : -1;
    if (qi < nq && b < 0 && out != nullptr) {
        // capacity padding row: no neighbours at all
        for (int col = lane; col < max_cols; col += 32) {
            if (IDX64) ((long long*)out)[(size_t)qi * max_cols + col] = pad_index;
            else ((int*)out)[(size_t)qi * max_cols + col] = pad_index;
        }
    }
    if (b >= 0) {
        const float qx = q[3 * qi], qy = q[3 * qi + 1], qz = q[3 * qi + 2];
        uint32_t beg = 0, n = 0;
        if (lane < 27 && ns > 0) {
            int cx = cell_coord(qx, inv_cs) + (lane % 3) - 1;
            int cy = cell_coord(qy, inv_cs) + ((lane / 3) % 3) - 1;
            int cz = cell_coord(qz, inv_cs) + (lane / 9) - 1;
            if (cx >= 0 && cy >= 0 && cz >= 0 && cx <= CELL_MAXC && cy <= CELL_MAXC && cz <= CELL_MAXC) {
                uint64_t key = cell_key(b, cx, cy, cz);
                uint32_t slot = d3f_hash64(key) & mask;
                while (true) {
                    unsigned long long k = keys[slot];
                    if (k == key) { beg = start[slot]; n = cnt[slot]; break; }
                    if (k == D3F_EMPTY_KEY) break;
                    slot = (slot + 1) & mask;
                }
            }
        }
        const uint32_t nmax = __reduce_max_sync(0xffffffffu, n);
        for (uint32_t it = 0; it < nmax; ++it) {
            bool hit = false;
            unsigned long long ck = 0;
            if (it < n) {
                const float4 p = sorted[beg + it];
                // nanoflann.hpp:431-439: ((dx*dx) + dy*dy) + dz*dz, fp32 round-to-nearest, no FMA
                const float dx = __fsub_rn(qx, p.x), dy = __fsub_rn(qy, p.y), dz = __fsub_rn(qz, p.z);
                float d2 = __fmul_rn(dx, dx);
                d2 = __fadd_rn(d2, __fmul_rn(dy, dy));
                d2 = __fadd_rn(d2, __fmul_rn(dz, dz));
                hit = d2 < r2;
                ck = ((unsigned long long)__float_as_uint(d2) << 32) | (uint32_t)__float_as_int(p.w);
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, hit);
            const int nh = __popc(bal);
            // The buffer holds `cap` candidates.  Only the max_cols nearest are ever written, so when it would overflow
            // it is compacted in place: sort, keep the max_cols smallest keys seen so far, go on.  A row with ANY number
            // of in-range supports (deformable radii: hundreds) therefore still yields its exact nearest max_cols; the
            // overflow flag is left for the case that cannot be compacted (max_cols + 32 > cap: count-only / untruncated
            // calls, which the drop-in wrapper retries with a larger buffer).
            if (nbuf + nh > cap && can_compact) {          // warp-uniform
                __syncwarp();
                nb_warp_sort(cand, nbuf, lane);
                nbuf = min(nbuf, max_cols);
            }
            if (hit) {
                int pos = nbuf + __popc(bal & ((1u << lane) - 1u));
                if (pos < cap) cand[pos] = ck;
            }
            nbuf += nh;
            count += nh;
        }
        if (lane == 0) {
            atomicMax(&s_max, count);
            if (nbuf > cap) s_ovf = 1;
        }
        if (out != nullptr) {
            const int m = min(nbuf, cap);
            __syncwarp();
            nb_warp_sort(cand, m, lane);
            for (int col = lane; col < max_cols; col += 32) {
                const int v = col < m ? (int)(uint32_t)(cand[col] & 0xffffffffULL) : pad_index;
                if (IDX64) ((long long*)out)[(size_t)qi * max_cols + col] = v;
                else ((int*)out)[(size_t)qi * max_cols + col] = v;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_max > 0) atomicMax(&info[0], s_max);
        if (s_ovf) atomicMax(&info[1], 1);
    }
}

struct NbWs {
    unsigned long long* keys; uint32_t* cnt; uint32_t* start; uint32_t* slot_of; uint32_t* rank;
    float4* sorted; uint32_t* cursor; uint32_t table;
};

uint32_t nb_table_size(int ns) {
    uint32_t t = d3f_pow2ceil((uint32_t)(ns > 0 ? ns : 1) * 2u);
    return t < 1024u ? 1024u : t;
}

size_t nb_layout(NbWs* w, void* base, size_t cap, int ns) {
    WsCursor c{(char*)base, 0, cap};
    w->table = nb_table_size(ns);
    w->keys = c.take<unsigned long long>(w->table);
    w->cnt = c.take<uint32_t>(w->table);
    w->cursor = c.take<uint32_t>(64);
    w->start = c.take<uint32_t>(w->table);
    w->slot_of = c.take<uint32_t>(ns > 0 ? ns : 1);
    w->rank = c.take<uint32_t>(ns > 0 ? ns : 1);
    w->sorted = c.take<float4>(ns > 0 ? ns : 1);
    return c.off;
}

}  // namespace

extern "C" size_t d3f_radius_neighbors_workspace_bytes(int n_queries, int n_supports, int n_batch) {
    (void)n_queries; (void)n_batch;
    NbWs w;
    return nb_layout(&w, nullptr, 0, n_supports);
}

extern "C" int d3f_radius_neighbors(const float* queries, const float* supports, const int32_t* q_lengths,
                                    const int32_t* s_lengths, int n_batch, int n_queries, int n_supports,
                                    float radius, int max_cols, void* out_idx, int idx_is_64, int pad_index,
                                    int32_t* out_info, int row_capacity, void* workspace,
                                    size_t workspace_bytes, d3f_stream stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    D3F_REQUIRE(n_queries >= 0 && n_supports >= 0 && n_batch >= 1 && n_batch < 1024, D3F_ERR_INVALID, "bad sizes");
    D3F_REQUIRE(out_info != nullptr, D3F_ERR_INVALID, "out_info is required");
    D3F_REQUIRE(radius > 0.f, D3F_ERR_INVALID, "radius must be positive");
    D3F_REQUIRE(out_idx == nullptr || max_cols > 0, D3F_ERR_INVALID, "max_cols must be positive when out_idx is given");
    D3F_REQUIRE(row_capacity >= 64 && row_capacity <= 8192 && (row_capacity & (row_capacity - 1)) == 0,
                D3F_ERR_INVALID, "row_capacity must be a power of two in [64, 8192]");
    if (n_queries == 0) {
        D3F_CHECK_CUDA(cudaMemsetAsync(out_info, 0, 4 * sizeof(int32_t), stream));
        return D3F_OK;
    }
    D3F_REQUIRE(queries && supports && q_lengths && s_lengths, D3F_ERR_INVALID, "null input");
    NbWs w;
    size_t need = nb_layout(&w, workspace, workspace_bytes, n_supports);
    D3F_REQUIRE(workspace != nullptr && need <= workspace_bytes, D3F_ERR_WORKSPACE, "workspace too small");

    const float cs = radius * 1.01f;
    const float inv_cs = 1.0f / cs;
    const float r2 = radius * radius;  // neighbors.cpp:226 (fp32 product)
    {   // info vector, hash keys (empty = all ones) and the contiguous cnt + cursor block: one launch (fill.cu)
        D3fFillSegs f;
        f.add(out_info, 4 * sizeof(int32_t), 0u);
        f.add(w.keys, (size_t)w.table * sizeof(unsigned long long), 0xFFFFFFFFu);
        f.add(w.cnt, (size_t)((char*)w.start - (char*)w.cnt), 0u);
        int rcf = d3f_fill_segments(f, stream);
        if (rcf) return rcf;
    }
    if (n_supports > 0) {
        const int T = 256;
        nb_insert_kernel<<<d3f_ceil_div(n_supports, T), T, 0, stream>>>(
            supports, s_lengths, n_batch, n_supports, inv_cs, w.keys, w.cnt, w.table - 1, w.slot_of, w.rank);
        D3F_CHECK_LAUNCH();
        nb_alloc_kernel<<<d3f_ceil_div((int)w.table, T), T, 0, stream>>>(w.cnt, w.start, w.table, w.cursor);
        D3F_CHECK_LAUNCH();
        nb_scatter_kernel<<<d3f_ceil_div(n_supports, T), T, 0, stream>>>(supports, n_supports, s_lengths, n_batch, w.slot_of,
                                                                        w.rank, w.start, w.sorted);
        D3F_CHECK_LAUNCH();
    }
    int warps = 8;
    while (warps > 1 && (size_t)warps * row_capacity * 8 > 96 * 1024) warps >>= 1;
    const size_t smem = (size_t)warps * row_capacity * 8;
    auto kern = idx_is_64 ? nb_query_kernel<true> : nb_query_kernel<false>;
    if (smem > 48 * 1024)
        D3F_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<d3f_ceil_div(n_queries, warps), warps * 32, smem, stream>>>(
        queries, q_lengths, n_batch, n_queries, n_supports, inv_cs, r2, w.keys, w.cnt, w.start, w.table - 1,
        w.sorted, max_cols, out_idx, out_info, row_capacity, pad_index < 0 ? n_supports : pad_index);
    D3F_CHECK_LAUNCH();
    // info[2] = occupied cells (diagnostic)
    D3F_CHECK_CUDA(cudaMemcpyAsync(out_info + 2, w.cursor + 1, sizeof(int32_t), cudaMemcpyDeviceToDevice, stream));
    return D3F_OK;
}
