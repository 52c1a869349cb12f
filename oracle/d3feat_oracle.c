/*
 * TEST INFRASTRUCTURE ONLY.  Plain-C CPU restatement of the index-producing
 * half of the D3Feat hot path.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product path (d3feat/pytorch_b200) never does.
 *
 * Parity status: PINNED.  The reference ships no golden vectors for this path
 * (SURVEY.md 8c), so this restatement is pinned against the reference itself:
 * tests/test_oracle_cpu.py compares it bit-for-bit with oracle/_ref
 * (the unmodified reference C++ behind ref_shim.cpp) and with the committed
 * fixtures in tests/golden/ generated from that library.
 *
 * What is restated (reference file:line):
 *   orc_radius_neighbors  <- cpp_wrappers/cpp_neighbors/neighbors/neighbors.cpp:211-332
 *       batch_nanoflann_neighbors: per batch element, every support with
 *       d2 < r*r (nanoflann.hpp:431-439 L2_Simple_Adaptor::evalMetric, fp32, no FMA;
 *       strict '<' from RadiusResultSet::worstDist nanoflann.hpp:250-253,1361),
 *       rows sorted by ascending d2 (nanoflann.hpp:1286-1287), global index =
 *       local + sum of previous support lengths, rows padded with supports.size()
 *       up to the largest row (neighbors.cpp:304-327).
 *       The kd-tree is an access structure only; this port is a brute-force scan.
 *       Ties in d2: the reference's std::sort is unstable; this port orders ties
 *       by ascending index (documented contract).
 *   orc_grid_subsampling  <- cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.cpp:5-106,109-211
 *       + grid_subsampling.h:74-79 (SampledData::update_points), cloud.h:120-143,
 *       cloud.cpp:27-67 (min_point / max_point).  Output ORDER is the iteration
 *       order of libstdc++'s std::unordered_map<size_t, SampledData>
 *       (grid_subsampling.cpp:48,85); that container is a third-party dependency
 *       (GCC 13.3 libstdc++, hashtable.h / hashtable_policy.h) restated below:
 *       identity hash, bucket = key % n, single forward list, a node goes to the
 *       front of its bucket if the bucket is non-empty and to the global list
 *       head otherwise, both on insert (_M_insert_bucket_begin) and on rehash
 *       (_M_rehash_aux); bucket counts follow _Prime_rehash_policy.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* Bucket counts produced by _Prime_rehash_policy with max_load_factor 1.0 and
 * growth factor 2 starting from an empty map; measured from this container's
 * libstdc++ (see oracle/README.md).  A rehash to PR[r+1] happens when the
 * element count would exceed PR[r]. */
static const size_t PR[] = {13, 29, 59, 127, 257, 541, 1109, 2357, 5087, 10273, 20753,
                            42043, 85229, 172933, 351061, 712697, 1447153, 2938679,
                            5967347, 12117689, 24607243};
#define NPR ((int)(sizeof(PR) / sizeof(PR[0])))

/* ------------------------------------------------------------------ */
/* radius neighbours                                                   */
/* ------------------------------------------------------------------ */
typedef struct { float d2; int idx; } cand_t;

static int cand_cmp(const void* a, const void* b) {
    const cand_t* x = (const cand_t*)a; const cand_t* y = (const cand_t*)b;
    if (x->d2 < y->d2) return -1;
    if (x->d2 > y->d2) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx);
}

/* Returns 0 on success.  *out is a malloc'ed [nq, *max_count] int32 matrix. */
int orc_radius_neighbors(const float* q, int nq, const float* s, int ns,
                         const int* q_len, const int* s_len, int nb, float radius,
                         int** out, int* max_count) {
    float r2 = radius * radius;                       /* neighbors.cpp:226 */
    int* cnt = (int*)calloc((size_t)(nq > 0 ? nq : 1), sizeof(int));
    size_t* off = (size_t*)calloc((size_t)nq + 1, sizeof(size_t));
    size_t cap = 1 << 20, used = 0;
    cand_t* all = (cand_t*)malloc(cap * sizeof(cand_t));
    int qi = 0, sum_sb = 0, mx = 0;
    for (int b = 0; b < nb; ++b) {
        for (int i = 0; i < q_len[b]; ++i, ++qi) {
            const float qx = q[3 * qi], qy = q[3 * qi + 1], qz = q[3 * qi + 2];
            off[qi] = used;
            for (int j = 0; j < s_len[b]; ++j) {
                const float* p = s + 3 * (size_t)(sum_sb + j);
                /* nanoflann.hpp:431-439: result += diff*diff, dims in order */
                float dx = qx - p[0], dy = qy - p[1], dz = qz - p[2];
                float d2 = dx * dx;
                d2 = d2 + dy * dy;
                d2 = d2 + dz * dz;
                if (d2 < r2) {
                    if (used == cap) { cap *= 2; all = (cand_t*)realloc(all, cap * sizeof(cand_t)); }
                    all[used].d2 = d2; all[used].idx = sum_sb + j; ++used;
                }
            }
            cnt[qi] = (int)(used - off[qi]);
            if (cnt[qi] > mx) mx = cnt[qi];
            qsort(all + off[qi], (size_t)cnt[qi], sizeof(cand_t), cand_cmp);
        }
        sum_sb += s_len[b];
    }
    off[nq] = used;
    int* res = (int*)malloc(sizeof(int) * ((size_t)nq * (size_t)mx + 1));
    for (int i = 0; i < nq; ++i)
        for (int j = 0; j < mx; ++j)
            res[(size_t)i * mx + j] = j < cnt[i] ? all[off[i] + j].idx : ns; /* pad = supports.size() */
    free(all); free(cnt); free(off);
    *out = res; *max_count = mx;
    return 0;
}

/* ------------------------------------------------------------------ */
/* grid subsampling with libstdc++ unordered_map iteration order        */
/* ------------------------------------------------------------------ */
typedef struct {
    size_t nbkt;     /* bucket count */
    int pr;          /* index into PR of the current bucket count, -1 = single bucket */
    int* before;     /* per bucket: node BEFORE its first node; -1 empty; 0 = list head sentinel */
    int* next;       /* per node (1-based; 0 = sentinel): next node, 0 = end */
    size_t* key;
    int size;
} umap_t;

static void umap_rehash(umap_t* m, size_t n) {
    /* hashtable.h _M_rehash_aux(__n, true_type) */
    int* nb = (int*)malloc(n * sizeof(int));
    for (size_t i = 0; i < n; ++i) nb[i] = -1;
    int p = m->next[0];
    m->next[0] = 0;
    size_t bbegin_bkt = 0;
    while (p) {
        int nx = m->next[p];
        size_t bkt = m->key[p] % n;
        if (nb[bkt] < 0) {
            m->next[p] = m->next[0];
            m->next[0] = p;
            nb[bkt] = 0;
            if (m->next[p]) nb[bbegin_bkt] = p;
            bbegin_bkt = bkt;
        } else {
            m->next[p] = m->next[nb[bkt]];
            m->next[nb[bkt]] = p;
        }
        p = nx;
    }
    free(m->before);
    m->before = nb; m->nbkt = n;
}

/* returns node id (>=1) of key, inserting it if absent (*fresh = 1) */
static int umap_get(umap_t* m, size_t k, int* fresh) {
    size_t bkt = k % m->nbkt;
    int prev = m->before[bkt];
    if (prev >= 0) {
        for (int p = m->next[prev]; p && m->key[p] % m->nbkt == bkt; p = m->next[p])
            if (m->key[p] == k) { *fresh = 0; return p; }
    }
    /* _M_insert_unique_node: _M_need_rehash(bkt_count, size, 1) */
    if (m->pr < 0 || (size_t)m->size + 1 > PR[m->pr]) {
        m->pr += 1;
        umap_rehash(m, PR[m->pr]);
        bkt = k % m->nbkt;
    }
    int node = ++m->size;
    m->key[node] = k;
    /* _M_insert_bucket_begin */
    if (m->before[bkt] >= 0) {
        m->next[node] = m->next[m->before[bkt]];
        m->next[m->before[bkt]] = node;
    } else {
        m->next[node] = m->next[0];
        m->next[0] = node;
        if (m->next[node]) m->before[m->key[m->next[node]] % m->nbkt] = node;
        m->before[bkt] = 0;
    }
    *fresh = 1;
    return node;
}

static int subsample_one(const float* p, int n, float dl, float* out) {
    if (n == 0) return 0;
    /* cloud.cpp:27-67 */
    float mn[3] = {p[0], p[1], p[2]}, mx[3] = {p[0], p[1], p[2]};
    for (int i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) {
            float v = p[3 * i + a];
            if (v < mn[a]) mn[a] = v;
            if (v > mx[a]) mx[a] = v;
        }
    /* grid_subsampling.cpp:27: floor(minCorner * (1/sampleDl)) * sampleDl  (1/float -> float) */
    float inv = 1 / dl;
    float org[3];
    for (int a = 0; a < 3; ++a) { float t = mn[a] * inv; org[a] = floorf(t) * dl; }
    /* grid_subsampling.cpp:30-31 */
    size_t NX = (size_t)floorf((mx[0] - org[0]) / dl) + 1;
    size_t NY = (size_t)floorf((mx[1] - org[1]) / dl) + 1;

    umap_t m;
    m.nbkt = 1; m.pr = -1; m.size = 0;
    m.before = (int*)malloc(sizeof(int)); m.before[0] = -1;
    m.next = (int*)calloc((size_t)n + 2, sizeof(int));
    m.key = (size_t*)calloc((size_t)n + 2, sizeof(size_t));
    float* sum = (float*)calloc(3 * ((size_t)n + 2), sizeof(float));
    int* cnt = (int*)calloc((size_t)n + 2, sizeof(int));
    for (int i = 0; i < n; ++i) {
        /* grid_subsampling.cpp:53-56 */
        size_t ix = (size_t)floorf((p[3 * i] - org[0]) / dl);
        size_t iy = (size_t)floorf((p[3 * i + 1] - org[1]) / dl);
        size_t iz = (size_t)floorf((p[3 * i + 2] - org[2]) / dl);
        size_t k = ix + NX * iy + NX * NY * iz;
        int fresh;
        int node = umap_get(&m, k, &fresh);
        /* grid_subsampling.h:74-79 update_points: count += 1; point += p (fp32, input order) */
        cnt[node] += 1;
        sum[3 * node] += p[3 * i];
        sum[3 * node + 1] += p[3 * i + 1];
        sum[3 * node + 2] += p[3 * i + 2];
    }
    /* grid_subsampling.cpp:85-87: for (auto& v : data) push_back(point * (1.0 / count)) ;
       operator*(PointXYZ, float) => the double reciprocal is rounded to float first */
    int o = 0;
    for (int nd = m.next[0]; nd; nd = m.next[nd], ++o) {
        float w = (float)(1.0 / cnt[nd]);
        out[3 * o] = sum[3 * nd] * w;
        out[3 * o + 1] = sum[3 * nd + 1] * w;
        out[3 * o + 2] = sum[3 * nd + 2] * w;
    }
    free(m.before); free(m.next); free(m.key); free(sum); free(cnt);
    return o;
}

/* grid_subsampling.cpp:109-211 (points only, max_p = 0 => no truncation).
 * out_pts has room for 3*n floats.  Returns the total number of sampled points. */
int orc_grid_subsampling(const float* p, int n, const int* len, int nb, float dl,
                         float* out_pts, int* out_len) {
    int sum_b = 0, total = 0;
    (void)n;
    for (int b = 0; b < nb; ++b) {
        int m = subsample_one(p + 3 * (size_t)sum_b, len[b], dl, out_pts + 3 * (size_t)total);
        out_len[b] = m;
        total += m;
        sum_b += len[b];
    }
    return total;
}

void orc_free(void* p) { free(p); }
