// Grid subsampling (barycentre per occupied voxel) on a device hash grid; replaces
// cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.cpp:5-106,109-211.
//
// Bit-exactness contract:
//   * voxel index arithmetic is the reference's fp32 sequence (IEEE sub/div/floor, no FMA);
//   * the barycentre is a SEQUENTIAL fp32 sum in input order times the fp32-rounded reciprocal,
//     so every cell's members are put in index order before they are summed;
//   * output order = iteration order of libstdc++'s unordered_map<size_t,...> with identity hash:
//     reproduced with the parallel closed form -- per bucket-count n of the prime policy, order
//     the processed sequence by (first-touch time of the bucket desc, position desc) -- which is
//     one segmented-min plus one sort per round, run by a single CTA per batch element.
//
// Kernels (all on `stream`, no host sync):
//   sub_bounds   : per batch element min/max -> origin, NX, NY            (grid = n_batch)
//   sub_insert   : per point voxel key -> open-addressing table, rank in cell, first index of cell
//   sub_alloc    : contiguous member range per occupied cell
//   sub_scatter  : member lists
//   sub_reduce   : per cell: members in index order, sequential sum, barycentre
//   sub_order    : per batch element: first-insertion rank of cells, unordered_map order replay
//   sub_emit     : gather barycentres in final order
#include "common.cuh"

namespace {

constexpr unsigned long long KEY_MASK = (1ULL << 56) - 1ULL;
constexpr int ORDER_THREADS = 1024;
constexpr int SM_SORT_CAP = 16384;   // entries of the shared-memory sort buffer
constexpr int SM_FT_CAP = 20753;     // bucket counts up to this prime fit the shared first-touch array
constexpr uint32_t R20 = (1u << 20) - 1u;

// bucket of a voxel key (identity hash, key % bucket_count); 32-bit fast path (64-bit '%' is a long software loop)
__device__ __forceinline__ uint32_t bucket_of(unsigned long long key, uint32_t nbkt) {
    return key <= 0xFFFFFFFFull ? (uint32_t)key % nbkt : (uint32_t)(key % nbkt);
}

__constant__ unsigned int c_primes[24] = {13, 29, 59, 127, 257, 541, 1109, 2357, 5087, 10273, 20753,
                                          42043, 85229, 172933, 351061, 712697, 1447153, 2938679,
                                          5967347, 12117689, 24607243, 0, 0, 0};
static const unsigned int h_primes[21] = {13, 29, 59, 127, 257, 541, 1109, 2357, 5087, 10273, 20753,
                                          42043, 85229, 172933, 351061, 712697, 1447153, 2938679,
                                          5967347, 12117689, 24607243};

struct SubGrid {
    float ox, oy, oz, pad;
    unsigned long long nx, ny;
};

__global__ void sub_bounds_kernel(const float* __restrict__ p, const int32_t* __restrict__ len, int nb, float dl,
                                  SubGrid* grid) {
    const int b = blockIdx.x;
    int st = 0;
    for (int i = 0; i < b; ++i) st += len[i];
    const int n = len[b];
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = threadIdx.x; i < n; i += blockDim.x)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = p[3 * (size_t)(st + i) + a];
            mn[a] = fminf(mn[a], v);
            mx[a] = fmaxf(mx[a], v);
        }
    __shared__ float smn[3][32], smx[3][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
        if (lane == 0) { smn[a][warp] = mn[a]; smx[a][warp] = mx[a]; }
    }
    __syncthreads();
    if (threadIdx.x == 0 && n > 0) {
        const int nw = blockDim.x >> 5;
        float org[3], hi[3];
        const float inv = __fdiv_rn(1.0f, dl);  // grid_subsampling.cpp:27  (1/sampleDl) in fp32
        for (int a = 0; a < 3; ++a) {
            float lo = smn[a][0], h = smx[a][0];
            for (int w = 1; w < nw; ++w) { lo = fminf(lo, smn[a][w]); h = fmaxf(h, smx[a][w]); }
            org[a] = __fmul_rn(floorf(__fmul_rn(lo, inv)), dl);
            hi[a] = h;
        }
        SubGrid g;
        g.ox = org[0]; g.oy = org[1]; g.oz = org[2]; g.pad = 0.f;
        // grid_subsampling.cpp:30-31
        g.nx = (unsigned long long)floorf(__fdiv_rn(__fsub_rn(hi[0], org[0]), dl)) + 1ULL;
        g.ny = (unsigned long long)floorf(__fdiv_rn(__fsub_rn(hi[1], org[1]), dl)) + 1ULL;
        grid[b] = g;
    }
}

__global__ void sub_insert_kernel(const float* __restrict__ p, const int32_t* __restrict__ len, int nb, int n,
                                  float dl, const SubGrid* __restrict__ grid, unsigned long long* keys,
                                  uint32_t* cnt, uint32_t* minidx, uint32_t mask, uint32_t* slot_of,
                                  uint32_t* rank, int32_t* info) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int st;
    const int b = d3f_batch_of(i, len, nb, &st);
    if (b < 0) return;  // capacity padding
    const SubGrid g = grid[b];
    // grid_subsampling.cpp:53-56
    const unsigned long long ix = (unsigned long long)floorf(__fdiv_rn(__fsub_rn(p[3 * (size_t)i], g.ox), dl));
    const unsigned long long iy = (unsigned long long)floorf(__fdiv_rn(__fsub_rn(p[3 * (size_t)i + 1], g.oy), dl));
    const unsigned long long iz = (unsigned long long)floorf(__fdiv_rn(__fsub_rn(p[3 * (size_t)i + 2], g.oz), dl));
    const unsigned long long map_idx = ix + g.nx * iy + g.nx * g.ny * iz;
    if (map_idx > KEY_MASK) atomicMax(&info[0], 1);  // grid too large for the 56-bit table key
    const unsigned long long key = ((unsigned long long)b << 56) | (map_idx & KEY_MASK);
    uint32_t slot = d3f_hash64(key) & mask;
    while (true) {
        const unsigned long long old = atomicCAS(&keys[slot], D3F_EMPTY_KEY, key);
        if (old == D3F_EMPTY_KEY || old == key) break;
        slot = (slot + 1) & mask;
    }
    slot_of[i] = slot;
    rank[i] = atomicAdd(&cnt[slot], 1u);
    atomicMin(&minidx[slot], (uint32_t)i);
}

__global__ void sub_alloc_kernel(const uint32_t* __restrict__ cnt, uint32_t* start, uint32_t table, uint32_t* cursor) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t c = t < table ? cnt[t] : 0u;
    const uint32_t lane = threadIdx.x & 31;
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    uint32_t base = 0;
    if (lane == 31 && total) base = atomicAdd(cursor, total);
    base = __shfl_sync(0xffffffffu, base, 31);
    if (t < table) start[t] = base + incl - c;
}

__global__ void sub_scatter_kernel(int n, const int32_t* __restrict__ len, int nb, const uint32_t* __restrict__ slot_of,
                                   const uint32_t* __restrict__ rank,
                                   const uint32_t* __restrict__ start, uint32_t* member) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int st;
    if (d3f_batch_of(i, len, nb, &st) < 0) return;
    member[start[slot_of[i]] + rank[i]] = (uint32_t)i;
}

__global__ void sub_reduce_kernel(const float* __restrict__ p, const int32_t* __restrict__ len, int nb,
                                  const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ cnt,
                                  const uint32_t* __restrict__ start, const uint32_t* __restrict__ slot_of,
                                  uint32_t* member, uint32_t table, float4* bary) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= table) return;
    const uint32_t c = cnt[t];
    if (c == 0) return;
    float sx = 0.f, sy = 0.f, sz = 0.f;  // SampledData(): point = PointXYZ() = 0  (grid_subsampling.h:26-29)
    if (c <= 32) {
        uint32_t* m = member + start[t];
        for (uint32_t a = 1; a < c; ++a) {  // insertion sort: members into input (index) order
            const uint32_t v = m[a];
            int bq = (int)a - 1;
            while (bq >= 0 && m[bq] > v) { m[bq + 1] = m[bq]; --bq; }
            m[bq + 1] = v;
        }
        for (uint32_t a = 0; a < c; ++a) {  // grid_subsampling.h:74-79: point += p, sequential fp32
            const size_t i = m[a];
            sx = __fadd_rn(sx, p[3 * i]); sy = __fadd_rn(sy, p[3 * i + 1]); sz = __fadd_rn(sz, p[3 * i + 2]);
        }
    } else {
        // heavy cell: scan the batch element in input order
        const int b = (int)(keys[t] >> 56);
        int st = 0;
        for (int i = 0; i < b; ++i) st += len[i];
        const int e = st + len[b];
        for (int i = st; i < e; ++i)
            if (slot_of[i] == t) {
                sx = __fadd_rn(sx, p[3 * (size_t)i]); sy = __fadd_rn(sy, p[3 * (size_t)i + 1]);
                sz = __fadd_rn(sz, p[3 * (size_t)i + 2]);
            }
    }
    const float w = (float)(1.0 / (double)c);  // grid_subsampling.cpp:87: point * (1.0 / count)
    bary[t] = make_float4(__fmul_rn(sx, w), __fmul_rn(sy, w), __fmul_rn(sz, w), 0.f);
}

// One CTA per batch element.
__global__ void __launch_bounds__(ORDER_THREADS, 1)
sub_order_kernel(const int32_t* __restrict__ len, int nb, const unsigned long long* __restrict__ keys,
                 const uint32_t* __restrict__ minidx, const uint32_t* __restrict__ slot_of,
                 uint32_t* cell_of_rank, unsigned long long* ckey, uint32_t* seq_all,
                 unsigned long long* gbuf_all, uint32_t* gft_all, size_t gbuf_stride, size_t gft_stride,
                 int32_t* out_len, int32_t* info) {
    extern __shared__ unsigned long long sm_dyn[];
    unsigned long long* sbuf = sm_dyn;                         // SM_SORT_CAP entries
    uint32_t* sft = (uint32_t*)(sm_dyn + SM_SORT_CAP);         // SM_FT_CAP entries
    __shared__ uint32_t s_scan[ORDER_THREADS];
    const int b = blockIdx.x, tid = threadIdx.x;
    int st = 0;
    for (int i = 0; i < b; ++i) st += len[i];
    const int n = len[b];

    // ---- first-insertion rank of every cell = rank of its first point among first points
    const int chunk = (n + ORDER_THREADS - 1) / ORDER_THREADS;
    const int i0 = st + min(n, tid * chunk), i1 = st + min(n, (tid + 1) * chunk);
    uint32_t local = 0;
    for (int i = i0; i < i1; ++i) local += (minidx[slot_of[i]] == (uint32_t)i);
    s_scan[tid] = local;
    __syncthreads();
    for (int o = 1; o < ORDER_THREADS; o <<= 1) {
        uint32_t v = tid >= o ? s_scan[tid - o] : 0u;
        __syncthreads();
        s_scan[tid] += v;
        __syncthreads();
    }
    const uint32_t M = s_scan[ORDER_THREADS - 1];
    uint32_t j = s_scan[tid] - local;
    uint32_t* cor = cell_of_rank + st;
    unsigned long long* ck = ckey + st;
    for (int i = i0; i < i1; ++i) {
        const uint32_t s = slot_of[i];
        if (minidx[s] == (uint32_t)i) { cor[j] = s; ck[j] = keys[s] & KEY_MASK; ++j; }
    }
    const bool bad = M > R20 || info[0] != 0;   // > 2^20 cells in one element, or voxel keys beyond 56 bits
    if (tid == 0) out_len[b] = bad ? -1 : (int32_t)M;
    __syncthreads();
    if (M == 0 || bad) return;

    // ---- unordered_map iteration order (closed form, one round per bucket count)
    uint32_t* seq = seq_all + st;
    const bool in_smem = M <= (uint32_t)SM_SORT_CAP;
    unsigned long long* buf = in_smem ? sbuf : gbuf_all + (size_t)b * gbuf_stride;
    uint32_t* ft = in_smem ? sft : gft_all + (size_t)b * gft_stride;
    uint32_t start = 0;
    for (int r = 0; c_primes[r] != 0; ++r) {
        const uint32_t nbkt = c_primes[r];
        const uint32_t end = min(M, nbkt);
        const uint32_t L = end;
        for (uint32_t p = start + tid; p < end; p += ORDER_THREADS) seq[p] = p;  // new inserts follow the rehashed list
        for (uint32_t q = tid; q < nbkt; q += ORDER_THREADS) ft[q] = 0xFFFFFFFFu;
        __syncthreads();
        for (uint32_t p = tid; p < L; p += ORDER_THREADS)
            atomicMin(&ft[bucket_of(ck[seq[p]], nbkt)], p);
        __syncthreads();
        uint32_t n2 = 1;
        while (n2 < L) n2 <<= 1;
        for (uint32_t p = tid; p < n2; p += ORDER_THREADS) {
            unsigned long long k = ~0ULL;
            if (p < L) {
                const uint32_t e = seq[p];
                const uint32_t f = ft[bucket_of(ck[e], nbkt)];
                k = ((unsigned long long)(R20 - f) << 40) | ((unsigned long long)(R20 - p) << 20) | e;
            }
            buf[p] = k;
        }
        __syncthreads();
        for (uint32_t k = 2; k <= n2; k <<= 1)
            for (uint32_t jj = k >> 1; jj > 0; jj >>= 1) {
                for (uint32_t t = tid; t < (n2 >> 1); t += ORDER_THREADS) {
                    const uint32_t i = ((t & ~(jj - 1)) << 1) | (t & (jj - 1));
                    const uint32_t l = i | jj;
                    const unsigned long long a = buf[i], c = buf[l];
                    if ((a > c) == ((i & k) == 0)) { buf[i] = c; buf[l] = a; }
                }
                __syncthreads();
            }
        for (uint32_t p = tid; p < L; p += ORDER_THREADS) seq[p] = (uint32_t)(buf[p] & R20);
        __syncthreads();
        start = end;
        if (end >= M) break;
    }
}

__global__ void sub_emit_kernel(const int32_t* __restrict__ len, int32_t* __restrict__ out_len, int nb,
                                int n, const uint32_t* __restrict__ seq_all, const uint32_t* __restrict__ cell_of_rank,
                                const float4* __restrict__ bary, float* out, int out_capacity) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == 0) {  // the caller's output buffer is too small: report it instead of writing out of bounds
        long long tot = 0;
        for (int e = 0; e < nb; ++e) tot += out_len[e] > 0 ? out_len[e] : 0;
        if (tot > out_capacity) out_len[nb] = 1;   // out_lengths has n_batch + 1 entries: [n_batch] = overflow flag
    }
    if (i >= out_capacity) return;
    // i indexes the OUTPUT row; find its batch element from out_len
    int so = 0, si = 0, b = -1;
    for (int e = 0; e < nb; ++e) {
        const int ol = out_len[e] > 0 ? out_len[e] : 0;
        if (i < so + ol) { b = e; break; }
        so += ol;
        si += len[e];
    }
    if (b < 0) return;
    const uint32_t r = seq_all[si + (i - so)];
    const float4 v = bary[cell_of_rank[si + r]];
    out[3 * (size_t)i] = v.x; out[3 * (size_t)i + 1] = v.y; out[3 * (size_t)i + 2] = v.z;
}

// If the caller's output buffer was too small, clamp the reported lengths to the rows that were written so
// that later stages never index past out_capacity (the overflow flag stays set in out_len[nb]).
__global__ void sub_clamp_kernel(int32_t* out_len, int nb, int out_capacity) {
    long long room = out_capacity;
    for (int e = 0; e < nb; ++e) {
        const long long l = out_len[e] > 0 ? out_len[e] : 0;
        if (l > room) out_len[e] = (int32_t)room;
        room -= l < room ? l : room;
    }
}

struct SubWs {
    SubGrid* grid; unsigned long long* keys; uint32_t* cnt; uint32_t* cursor; uint32_t* start; uint32_t* minidx;
    float4* bary; uint32_t* slot_of; uint32_t* rank; uint32_t* member; uint32_t* cell_of_rank;
    unsigned long long* ckey; uint32_t* seq; unsigned long long* gbuf; uint32_t* gft; int32_t* info;
    uint32_t table; size_t gbuf_stride, gft_stride;
};

size_t sub_layout(SubWs* w, void* base, size_t cap, int n, int nb) {
    WsCursor c{(char*)base, 0, cap};
    const size_t nn = n > 0 ? n : 1;
    uint32_t t = d3f_pow2ceil((uint32_t)nn * 2u);
    w->table = t < 1024u ? 1024u : t;
    w->grid = c.take<SubGrid>(nb);
    w->keys = c.take<unsigned long long>(w->table);
    w->cnt = c.take<uint32_t>(w->table);
    w->cursor = c.take<uint32_t>(64);
    w->info = c.take<int32_t>(64);
    w->start = c.take<uint32_t>(w->table);
    w->minidx = c.take<uint32_t>(w->table);
    w->bary = c.take<float4>(w->table);
    w->slot_of = c.take<uint32_t>(nn);
    w->rank = c.take<uint32_t>(nn);
    w->member = c.take<uint32_t>(nn);
    w->cell_of_rank = c.take<uint32_t>(nn);
    w->ckey = c.take<unsigned long long>(nn);
    w->seq = c.take<uint32_t>(nn);
    // global fallbacks of the order replay (used only when one element has > SM_SORT_CAP cells)
    if ((int)nn > SM_SORT_CAP) {
        w->gbuf_stride = d3f_pow2ceil((uint32_t)nn);
        size_t pr = 0;
        for (int i = 0; i < 21; ++i) { pr = h_primes[i]; if (pr >= nn) break; }
        w->gft_stride = pr;
        w->gbuf = c.take<unsigned long long>(w->gbuf_stride * nb);
        w->gft = c.take<uint32_t>(w->gft_stride * nb);
    } else {
        w->gbuf_stride = w->gft_stride = 0;
        w->gbuf = nullptr; w->gft = nullptr;
    }
    return c.off;
}

}  // namespace

extern "C" size_t d3f_grid_subsample_workspace_bytes(int n_points, int n_batch) {
    SubWs w;
    return sub_layout(&w, nullptr, 0, n_points, n_batch > 0 ? n_batch : 1);
}

extern "C" int d3f_grid_subsample(const float* points, const int32_t* lengths, int n_batch, int n_points,
                                  float sample_dl, float* out_points, int out_capacity, int32_t* out_lengths,
                                  void* workspace, size_t workspace_bytes, d3f_stream stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    D3F_REQUIRE(n_points >= 0 && n_batch >= 1 && n_batch < 256, D3F_ERR_INVALID, "bad sizes");
    D3F_REQUIRE(lengths && out_lengths, D3F_ERR_INVALID, "null lengths");
    D3F_REQUIRE(sample_dl > 0.f, D3F_ERR_INVALID, "sample_dl must be positive");
    D3F_REQUIRE(n_points < (1 << 30), D3F_ERR_UNSUPPORTED, "too many points");
    D3F_REQUIRE(out_capacity >= 0, D3F_ERR_INVALID, "bad out_capacity");
    D3fFillSegs f;          // every region this call initialises, cleared by ONE launch (fill.cu)
    f.add(out_lengths, (size_t)(n_batch + 1) * sizeof(int32_t), 0u);
    if (out_points && out_capacity > 0) f.add(out_points, sizeof(float) * 3 * (size_t)out_capacity, 0u);
    if (n_points == 0) return d3f_fill_segments(f, stream);
    D3F_REQUIRE(points && out_points, D3F_ERR_INVALID, "null points");
    SubWs w;
    const size_t need = sub_layout(&w, workspace, workspace_bytes, n_points, n_batch);
    D3F_REQUIRE(workspace != nullptr && need <= workspace_bytes, D3F_ERR_WORKSPACE, "workspace too small");

    const int T = 256;
    f.add(w.keys, (size_t)w.table * 8, 0xFFFFFFFFu);
    f.add(w.cnt, (size_t)((char*)w.start - (char*)w.cnt), 0u);  // cnt, cursor, info
    f.add(w.minidx, (size_t)w.table * 4, 0xFFFFFFFFu);
    {
        int rcf = d3f_fill_segments(f, stream);
        if (rcf) return rcf;
    }
    sub_bounds_kernel<<<n_batch, 1024, 0, stream>>>(points, lengths, n_batch, sample_dl, w.grid);
    D3F_CHECK_LAUNCH();
    sub_insert_kernel<<<d3f_ceil_div(n_points, T), T, 0, stream>>>(points, lengths, n_batch, n_points, sample_dl,
                                                                  w.grid, w.keys, w.cnt, w.minidx, w.table - 1,
                                                                  w.slot_of, w.rank, w.info);
    D3F_CHECK_LAUNCH();
    sub_alloc_kernel<<<d3f_ceil_div((int)w.table, T), T, 0, stream>>>(w.cnt, w.start, w.table, w.cursor);
    D3F_CHECK_LAUNCH();
    sub_scatter_kernel<<<d3f_ceil_div(n_points, T), T, 0, stream>>>(n_points, lengths, n_batch, w.slot_of, w.rank, w.start, w.member);
    D3F_CHECK_LAUNCH();
    sub_reduce_kernel<<<d3f_ceil_div((int)w.table, T), T, 0, stream>>>(points, lengths, n_batch, w.keys, w.cnt,
                                                                      w.start, w.slot_of, w.member, w.table, w.bary);
    D3F_CHECK_LAUNCH();
    const size_t smem = (size_t)SM_SORT_CAP * 8 + (size_t)SM_FT_CAP * 4;
    D3F_CHECK_CUDA(cudaFuncSetAttribute(sub_order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sub_order_kernel<<<n_batch, ORDER_THREADS, smem, stream>>>(lengths, n_batch, w.keys, w.minidx, w.slot_of,
                                                              w.cell_of_rank, w.ckey, w.seq, w.gbuf, w.gft,
                                                              w.gbuf_stride, w.gft_stride, out_lengths, w.info);
    D3F_CHECK_LAUNCH();
    sub_emit_kernel<<<d3f_ceil_div(n_points, T), T, 0, stream>>>(lengths, out_lengths, n_batch, n_points, w.seq,
                                                                w.cell_of_rank, w.bary, out_points, out_capacity);
    D3F_CHECK_LAUNCH();
    sub_clamp_kernel<<<1, 1, 0, stream>>>(out_lengths, n_batch, out_capacity);
    D3F_CHECK_LAUNCH();
    return D3F_OK;
}
