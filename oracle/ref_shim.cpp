// TEST INFRASTRUCTURE ONLY -- never linked into or imported by the product path.
//
// C-ABI shim around the UNMODIFIED reference C++ (compiled where it lies under
// /root/reference by oracle/Makefile; outputs go to oracle/_ref/ only).  The
// reference's own CPython glue (cpp_wrappers/cpp_neighbors/wrapper.cpp:58-238,
// cpp_wrappers/cpp_subsampling/wrapper.cpp:62-333) does not compile against
// NumPy 2.x, so this shim does the same buffer -> std::vector<PointXYZ>
// marshalling that glue does and calls the same two entry points:
//   batch_nanoflann_neighbors  (cpp_neighbors/neighbors/neighbors.cpp:211-332)
//   batch_grid_subsampling     (cpp_subsampling/grid_subsampling/grid_subsampling.cpp:109-211)
#include "cpp_neighbors/neighbors/neighbors.h"
#include "cpp_subsampling/grid_subsampling/grid_subsampling.h"
#include <cstring>
#include <cstdlib>

static std::vector<PointXYZ> as_points(const float* p, int n) {
    std::vector<PointXYZ> v(n);
    for (int i = 0; i < n; ++i) v[i] = PointXYZ(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
    return v;
}

extern "C" {

// Returns a malloc'ed [nq, *max_count] int32 matrix (caller frees with ref_free).
int* ref_batch_neighbors(const float* q, int nq, const float* s, int ns,
                         const int* q_len, const int* s_len, int nb, float radius,
                         int* max_count) {
    std::vector<PointXYZ> queries = as_points(q, nq), supports = as_points(s, ns);
    std::vector<int> qb(q_len, q_len + nb), sb(s_len, s_len + nb), out;
    batch_nanoflann_neighbors(queries, supports, qb, sb, out, radius);
    *max_count = nq > 0 ? (int)(out.size() / (size_t)nq) : 0;
    int* buf = (int*)malloc(sizeof(int) * (out.size() ? out.size() : 1));
    if (out.size()) memcpy(buf, out.data(), sizeof(int) * out.size());
    return buf;
}

// Points-only variant (the only one D3Feat uses: datasets/dataloader.py:138).
// out_pts has room for 3*n floats, out_len for nb ints.  Returns total sampled count.
int ref_batch_grid_subsampling(const float* p, int n, const int* len, int nb,
                               float dl, int max_p, float* out_pts, int* out_len) {
    std::vector<PointXYZ> pts = as_points(p, n), sub;
    std::vector<float> f, sf;
    std::vector<int> c, sc, ob(len, len + nb), sbatch;
    batch_grid_subsampling(pts, sub, f, sf, c, sc, ob, sbatch, dl, max_p);
    for (size_t i = 0; i < sub.size(); ++i) {
        out_pts[3 * i] = sub[i].x; out_pts[3 * i + 1] = sub[i].y; out_pts[3 * i + 2] = sub[i].z;
    }
    for (int b = 0; b < nb; ++b) out_len[b] = sbatch[b];
    return (int)sub.size();
}

void ref_free(void* p) { free(p); }

}  // extern "C"
