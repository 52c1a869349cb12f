// Pairwise descriptor distance + circle / hardest-contrastive descriptor loss + detector loss,
// forward and backward (replaces utils/loss.py:8-44, 55-97, 111-141, 149-158 as wired by
// trainer.py:90-98).  P is 64..128 for one pair and B*128 after the cross-fragment all-gather.
//
// forward : pl_dist (D[i,j], + the +10 bumps of ContrastiveLoss) -> pl_rowcol (one warp per row and
//           one per column: two-pass logsumexp of the circle terms, furthest positive, closest
//           negative, row sum) -> pl_final (one CTA: means, accuracy, detector loss)
// backward: pl_grad (dL/dD[i,j] scaled by the metric's 1/D factor) -> pl_grad_desc (one warp per anchor
//           row / positive row accumulates dD/d(descriptor))
#include "common.cuh"

namespace {

__device__ __forceinline__ bool pl_gt(const void* kp, int f64, size_t i, double thr) {
    return f64 ? (((const double*)kp)[i] > thr) : ((double)((const float*)kp)[i] > thr);
}
// ContrastiveLoss bump test, literally loss.py:58-61: (eye*10 + dist_keypts) < safe_radius
__device__ __forceinline__ bool pl_bump(const void* kp, int f64, size_t e, bool diag, double thr) {
    const double v = f64 ? ((const double*)kp)[e] : (double)((const float*)kp)[e];
    return ((diag ? 10.0 : 0.0) + v) < thr;
}

__device__ __forceinline__ float pl_metric(const float* __restrict__ a, const float* __restrict__ b, int D, int metric) {
    float acc = 0.f;
    if (metric == D3F_METRIC_COSINE || metric == D3F_METRIC_ARCCOSINE) {
        for (int d = 0; d < D; ++d) acc = fmaf(a[d], b[d], acc);
        return metric == D3F_METRIC_COSINE ? sqrtf(2.0f - 2.0f * acc) : acosf(acc);
    }
    if (metric == D3F_METRIC_CITYBLOCK) {
        for (int d = 0; d < D; ++d) acc += fabsf(a[d] - b[d]);
        return acc;
    }
    for (int d = 0; d < D; ++d) { const float t = a[d] - b[d]; acc = fmaf(t, t, acc); }
    return metric == D3F_METRIC_SQEUCLIDEAN ? acc : sqrtf(acc + 1e-12f);
}

// dists[i,j]; bump: ContrastiveLoss adds 10 where (dist_keypts + 10*I) < safe_radius (loss.py:58-61)
__global__ void pl_dist_kernel(const float* __restrict__ a, const float* __restrict__ b, int Pa, int Pb, int D,
                               int metric, const void* keypts, int f64, double safe_radius, int bump,
                               float* __restrict__ dists) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= Pb || i >= Pa) return;
    float d = pl_metric(a + (size_t)i * D, b + (size_t)j * D, D, metric);
    if (bump && pl_bump(keypts, f64, (size_t)i * Pb + j, i == j, safe_radius)) d += 10.0f;
    dists[(size_t)i * Pb + j] = d;
}

struct PlAux {  // views into the aux buffer, each [P]
    float *lpr, *lnr, *lpc, *lnc, *sig_row, *sig_col, *fp, *cn, *rsum; int* cn_arg; float* G;
};
__host__ __device__ inline PlAux pl_aux(float* aux, int P) {
    PlAux x;
    x.lpr = aux; x.lnr = aux + P; x.lpc = aux + 2 * P; x.lnc = aux + 3 * P; x.sig_row = aux + 4 * P;
    x.sig_col = aux + 5 * P; x.fp = aux + 6 * P; x.cn = aux + 7 * P; x.rsum = aux + 8 * P;
    x.cn_arg = (int*)(aux + 9 * P); x.G = aux + 16 * (size_t)P + 16;
    return x;
}

__device__ __forceinline__ void circle_terms(float d, bool neg_mask, float pm, float nm, float s, float* zp,
                                             float* zn, float* pw, float* nw) {
    // loss.py:125-135, literally, in fp32
    const float pos = d - 1e5f * (neg_mask ? 1.0f : 0.0f);
    *pw = fmaxf(pos - pm, 0.0f);
    *zp = s * (pos - pm) * (*pw);
    const float neg = d + 1e5f * (neg_mask ? 0.0f : 1.0f);
    *nw = fmaxf(nm - neg, 0.0f);
    *zn = s * (nm - neg) * (*nw);
}

__device__ __forceinline__ float warp_max(float v) {
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// warp w < P: row w.  warp w >= P: column w-P (circle only).
__global__ void pl_rowcol_kernel(const float* __restrict__ dists, int P, const void* keypts, int f64,
                                 double safe_radius, int loss_kind, float pm, float nm, float s, float* aux) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= 2 * P) return;
    const PlAux x = pl_aux(aux, P);
    const bool is_col = w >= P;
    const int r = is_col ? w - P : w;
    if (is_col && loss_kind != D3F_LOSS_CIRCLE) return;
    float mp = -INFINITY, mn = -INFINITY;
    if (loss_kind == D3F_LOSS_CIRCLE) {
        for (int t = lane; t < P; t += 32) {
            const size_t e = is_col ? (size_t)t * P + r : (size_t)r * P + t;
            float zp, zn, pw, nw;
            circle_terms(dists[e], pl_gt(keypts, f64, e, safe_radius), pm, nm, s, &zp, &zn, &pw, &nw);
            mp = fmaxf(mp, zp); mn = fmaxf(mn, zn);
        }
        mp = warp_max(mp); mn = warp_max(mn);
        float sp = 0.f, sn = 0.f;
        for (int t = lane; t < P; t += 32) {
            const size_t e = is_col ? (size_t)t * P + r : (size_t)r * P + t;
            float zp, zn, pw, nw;
            circle_terms(dists[e], pl_gt(keypts, f64, e, safe_radius), pm, nm, s, &zp, &zn, &pw, &nw);
            sp += expf(zp - mp); sn += expf(zn - mn);
        }
        sp = warp_sum(sp); sn = warp_sum(sn);
        if (lane == 0) {
            const float lp = mp + logf(sp), ln = mn + logf(sn);
            if (is_col) { x.lpc[r] = lp; x.lnc[r] = ln; } else { x.lpr[r] = lp; x.lnr[r] = ln; }
        }
    }
    if (!is_col) {
        // furthest positive = max_j(D * I), closest negative = min_j(D + 1e5 * I), row sum (loss.py:120-122)
        float rs = 0.f, best = INFINITY; int arg = 0x7fffffff;
        for (int t = lane; t < P; t += 32) {
            const float d = dists[(size_t)r * P + t];
            rs += d;
            const float v = d + (t == r ? 1e5f : 0.0f);
            if (v < best) { best = v; arg = t; }
        }
        rs = warp_sum(rs);
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (ob < best || (ob == best && oa < arg)) { best = ob; arg = oa; }
        }
        if (lane == 0) {
            const float dd = dists[(size_t)r * P + r];
            x.fp[r] = P > 1 ? fmaxf(dd, 0.0f) : dd;
            x.cn[r] = best; x.cn_arg[r] = arg; x.rsum[r] = rs;
        }
    }
}

__device__ __forceinline__ float softplus1(float v) { return v > 20.0f ? v : log1pf(expf(v)); }
__device__ __forceinline__ float sigmoid_sp(float v) { return v > 20.0f ? 1.0f : 1.0f / (1.0f + expf(-v)); }

__global__ void pl_final_kernel(int P, int loss_kind, float pm, float nm, float s, const float* __restrict__ sa,
                                const float* __restrict__ sp, float* aux, float* stats, float* fp_out,
                                float* an_out) {
    const PlAux x = pl_aux(aux, P);
    __shared__ float red[5][32];
    float l = 0.f, det = 0.f, acc = 0.f, mfp = 0.f, man = 0.f;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const float fp = x.fp[i], cn = x.cn[i];
        if (loss_kind == D3F_LOSS_CIRCLE) {
            const float ar = x.lpr[i] + x.lnr[i], ac = x.lpc[i] + x.lnc[i];
            l += softplus1(ar) / s + softplus1(ac) / s;
            x.sig_row[i] = sigmoid_sp(ar); x.sig_col[i] = sigmoid_sp(ac);
        } else {
            l += fmaxf(fp - pm, 0.0f) + fmaxf(nm - cn, 0.0f);
        }
        if (sa && sp) det += (fp - cn) * (sa[i] + sp[i]);
        acc += (fp - cn) < 0.0f ? 1.0f : 0.0f;
        const float an = (x.rsum[i] - fp) / (float)(P - 1);
        mfp += fp; man += an;
        if (fp_out) fp_out[i] = fp;
        if (an_out) an_out[i] = an;
    }
    float v[5] = {l, det, acc, mfp, man};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        v[q] = warp_sum(v[q]);
        if (lane == 0) red[q][warp] = v[q];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = blockDim.x >> 5;
        float t[5] = {0, 0, 0, 0, 0};
        for (int q = 0; q < 5; ++q) for (int w2 = 0; w2 < nw; ++w2) t[q] += red[q][w2];
        stats[0] = t[0] / P; stats[1] = t[1] / P; stats[2] = t[2] * 100.0f / P; stats[3] = t[3] / P;
        stats[4] = t[4] / P; stats[5] = stats[6] = stats[7] = 0.f;
    }
}

// G[i,j] = dL/dD[i,j] * (metric factor so that dD/da = G * direction)
__global__ void pl_grad_kernel(const float* __restrict__ dists, int P, const void* keypts, int f64,
                               double safe_radius, int loss_kind, int metric, float pm, float nm, float s,
                               const float* __restrict__ sa, const float* __restrict__ sp,
                               const float* __restrict__ gl, float* aux) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= P || i >= P) return;
    const PlAux x = pl_aux(aux, P);
    const size_t e = (size_t)i * P + j;
    const float d = dists[e];
    const float gdesc = gl[0] / P, gdet = gl[1] / P;
    float g = 0.f;
    const bool is_fp = (i == j) && (P == 1 || d > 0.0f);
    const bool is_cn = (j == x.cn_arg[i]);
    if (loss_kind == D3F_LOSS_CIRCLE) {
        float zp, zn, pw, nw;
        circle_terms(d, pl_gt(keypts, f64, e, safe_radius), pm, nm, s, &zp, &zn, &pw, &nw);
        const float row = x.sig_row[i] * (expf(zp - x.lpr[i]) * pw - expf(zn - x.lnr[i]) * nw);
        const float col = x.sig_col[j] * (expf(zp - x.lpc[j]) * pw - expf(zn - x.lnc[j]) * nw);
        g += gdesc * (row + col);
    } else {
        if (is_fp && x.fp[i] - pm > 0.0f) g += gdesc;
        if (is_cn && nm - x.cn[i] > 0.0f) g -= gdesc;
    }
    if (sa && sp) {
        const float sc = gdet * (sa[i] + sp[i]);
        if (is_fp) g += sc;
        if (is_cn) g -= sc;
    }
    // strip the contrastive bump to recover the metric value
    float dm = d;
    if (loss_kind == D3F_LOSS_CONTRASTIVE && pl_bump(keypts, f64, e, i == j, safe_radius)) dm = d - 10.0f;
    float f = g;
    if (metric == D3F_METRIC_EUCLIDEAN) f = g / dm;
    else if (metric == D3F_METRIC_SQEUCLIDEAN) f = 2.0f * g;
    else if (metric == D3F_METRIC_COSINE) f = -g / dm;
    else if (metric == D3F_METRIC_ARCCOSINE) { const float c = cosf(dm); f = -g / sqrtf(fmaxf(1.0f - c * c, 1e-30f)); }
    x.G[e] = f;
}

// warp w < P: grad_anchor row w;  w >= P: grad_positive row w-P.
__global__ void pl_grad_desc_kernel(const float* __restrict__ a, const float* __restrict__ p, int P, int D,
                                    int metric, const float* __restrict__ aux_c, float* __restrict__ ga,
                                    float* __restrict__ gp) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= 2 * P) return;
    const float* G = aux_c + 16 * (size_t)P + 16;
    const bool for_p = w >= P;
    const int r = for_p ? w - P : w;
    const float* self = (for_p ? p : a) + (size_t)r * D;
    const float* other = for_p ? a : p;
    float* out = (for_p ? gp : ga) + (size_t)r * D;
    const bool dot_metric = metric == D3F_METRIC_COSINE || metric == D3F_METRIC_ARCCOSINE;
    // The coefficient row (or column) G[r, :] is read 32 entries at a time by the lanes and broadcast with shuffles, so the
    // 32 row loads of a chunk are independent and all in flight (the scalar loop this replaces chained one L2 round
    // trip per t behind a data-dependent `continue`: 64 us for P = 128; entries with g == 0 contribute exactly 0 here too)
    for (int d0 = 0; d0 < D; d0 += 32) {
        const int d = d0 + lane;
        const float sv = d < D ? self[d] : 0.f;
        float acc = 0.f;
        for (int t0 = 0; t0 < P; t0 += 32) {
            const int tl = t0 + lane;
            const float gv = tl < P ? (for_p ? G[(size_t)tl * P + r] : G[(size_t)r * P + tl]) : 0.f;
            float ov[32];
#pragma unroll
            for (int tt = 0; tt < 32; ++tt) ov[tt] = (d < D && t0 + tt < P) ? other[(size_t)(t0 + tt) * D + d] : sv;
#pragma unroll
            for (int tt = 0; tt < 32; ++tt) {
                const float g = __shfl_sync(0xffffffffu, gv, tt);
                if (dot_metric) acc = fmaf(g, ov[tt], acc);
                else {
                    // anchor: +(a-p) ; positive: d/dp of f(a-p) = -(a-p) = (p-a)
                    const float diff = sv - ov[tt];
                    if (metric == D3F_METRIC_CITYBLOCK) acc += g * (diff > 0.f ? 1.0f : (diff < 0.f ? -1.0f : 0.0f));
                    else acc = fmaf(g, diff, acc);
                }
            }
        }
        if (d < D) out[d] = acc;
    }
}

__global__ void pl_grad_score_kernel(int P, const float* __restrict__ aux_c, const float* __restrict__ gl,
                                     float* __restrict__ gsa, float* __restrict__ gsp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float* fp = aux_c + 6 * (size_t)P;
    const float* cn = aux_c + 7 * (size_t)P;
    const float v = gl[1] * (fp[i] - cn[i]) / P;
    if (gsa) gsa[i] = v;
    if (gsp) gsp[i] = v;
}

int pl_check(int P, int D, int loss_kind, int metric) {
    D3F_REQUIRE(P >= 1 && D >= 1, D3F_ERR_INVALID, "bad sizes");
    D3F_REQUIRE(P <= 32768, D3F_ERR_UNSUPPORTED, "P too large");
    D3F_REQUIRE(loss_kind == D3F_LOSS_CIRCLE || loss_kind == D3F_LOSS_CONTRASTIVE, D3F_ERR_INVALID, "bad loss kind");
    D3F_REQUIRE(metric >= 0 && metric <= 4, D3F_ERR_INVALID, "bad metric");
    return D3F_OK;
}

// ---- detector loss on an ARBITRARY distance matrix (utils/loss.py:149-158), for callers that do not come through
// CircleLoss / ContrastiveLoss of this package (any [P,P] fp32 matrix, e.g. a cdist):
//   fp_i = max_j dists[i,j]·[i==j]  (= max(d_ii, 0) for P > 1),  cn_i = min_j dists[i,j] + 1e5·[i==j],
//   loss = mean_i (fp_i - cn_i)·(anc_i + pos_i).
// One warp per row; ties take the smallest column (torch leaves them unspecified); fixed-order final sum.
__global__ void __launch_bounds__(256)
det_rows_kernel(const float* __restrict__ dists, int ld, int P, float* __restrict__ rowval, int* __restrict__ arg) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= P) return;
    const float* row = dists + (size_t)i * ld;
    float best = INFINITY;
    int bj = 0x7fffffff;
    bool nan_seen = false;
    for (int j = lane; j < P; j += 32) {
        const float v = j == i ? __fadd_rn(row[j], 1e5f) : row[j];
        nan_seen |= v != v;
        if (v < best || (v == best && j < bj)) { best = v; bj = j; }
    }
    for (int o = 16; o; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
        if (ov < best || (ov == best && oj < bj)) { best = ov; bj = oj; }
    }
    nan_seen = __any_sync(0xffffffffu, nan_seen);
    if (lane == 0) {
        const float dii = row[i];
        const bool diag = P == 1 || dii > 0.f;
        const float fp = diag ? dii : (dii != dii ? dii : 0.f);
        rowval[i] = nan_seen ? NAN : fp - best;
        arg[i] = diag ? i : -1;
        arg[P + i] = bj == 0x7fffffff ? -1 : bj;
    }
}

__global__ void __launch_bounds__(256)
det_final_kernel(const float* __restrict__ rowval, const float* __restrict__ anc, const float* __restrict__ pos, int P,
                 float* __restrict__ loss) {
    __shared__ double sh[256];
    double acc = 0.0;
    for (int i = threadIdx.x; i < P; i += 256) acc += (double)(rowval[i] * (anc[i] + pos[i]));
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) loss[0] = (float)(sh[0] / (double)P);
}

__global__ void __launch_bounds__(256)
det_grad_kernel(const float* __restrict__ rowval, const int* __restrict__ arg, const float* __restrict__ anc,
                const float* __restrict__ pos, int P, const float* __restrict__ grad_loss, float* __restrict__ grad_dists,
                int ld, float* __restrict__ grad_anc, float* __restrict__ grad_pos) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= P) return;
    const float g = grad_loss[0] / (float)P;
    if (grad_anc) grad_anc[i] = g * rowval[i];
    if (grad_pos) grad_pos[i] = g * rowval[i];
    if (grad_dists) {      // zero-filled by the caller of this kernel; one thread owns row i
        const float gs = g * (anc[i] + pos[i]);
        if (arg[i] >= 0) grad_dists[(size_t)i * ld + arg[i]] += gs;
        if (arg[P + i] >= 0) grad_dists[(size_t)i * ld + arg[P + i]] -= gs;
    }
}

}  // namespace

extern "C" size_t d3f_pair_loss_aux_floats(int P) {
    return 16 * (size_t)(P > 0 ? P : 1) + 16 + (size_t)P * P;
}

extern "C" int d3f_pair_dist(const float* a, const float* b, int Pa, int Pb, int D, int metric, float* dists,
                             d3f_stream stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    D3F_REQUIRE(Pa >= 0 && Pb >= 0 && D >= 1 && metric >= 0 && metric <= 4, D3F_ERR_INVALID, "bad arguments");
    if (Pa == 0 || Pb == 0) return D3F_OK;
    D3F_REQUIRE(a && b && dists, D3F_ERR_INVALID, "null pointer");
    D3F_REQUIRE(Pa <= 65535, D3F_ERR_UNSUPPORTED, "Pa too large");
    pl_dist_kernel<<<dim3(d3f_ceil_div(Pb, 128), Pa), 128, 0, stream>>>(a, b, Pa, Pb, D, metric, nullptr, 0, 0.0, 0, dists);
    D3F_CHECK_LAUNCH();
    return D3F_OK;
}

extern "C" int d3f_pair_loss_forward(const float* anchor, const float* positive, int P, int D,
                                     const void* dist_keypts, int keypts_is_f64, const float* anc_score,
                                     const float* pos_score, int loss_kind, int metric, double safe_radius,
                                     float pos_margin, float neg_margin, float log_scale, float* dists,
                                     float* stats, float* furthest_pos, float* avg_neg, float* aux,
                                     d3f_stream stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = pl_check(P, D, loss_kind, metric);
    if (rc) return rc;
    D3F_REQUIRE(anchor && positive && dist_keypts && dists && stats && aux, D3F_ERR_INVALID, "null pointer");
    pl_dist_kernel<<<dim3(d3f_ceil_div(P, 128), P), 128, 0, stream>>>(
        anchor, positive, P, P, D, metric, dist_keypts, keypts_is_f64, safe_radius,
        loss_kind == D3F_LOSS_CONTRASTIVE, dists);
    D3F_CHECK_LAUNCH();
    pl_rowcol_kernel<<<d3f_ceil_div(2 * P, 8), 256, 0, stream>>>(dists, P, dist_keypts, keypts_is_f64, safe_radius,
                                                                loss_kind, pos_margin, neg_margin, log_scale, aux);
    D3F_CHECK_LAUNCH();
    pl_final_kernel<<<1, 256, 0, stream>>>(P, loss_kind, pos_margin, neg_margin, log_scale, anc_score, pos_score, aux,
                                          stats, furthest_pos, avg_neg);
    D3F_CHECK_LAUNCH();
    return D3F_OK;
}

extern "C" int d3f_pair_loss_backward(const float* anchor, const float* positive, int P, int D,
                                      const void* dist_keypts, int keypts_is_f64, const float* anc_score,
                                      const float* pos_score, int loss_kind, int metric, double safe_radius,
                                      float pos_margin, float neg_margin, float log_scale, const float* dists,
                                      const float* aux, const float* grad_losses, float* grad_anchor,
                                      float* grad_positive, float* grad_anc_score, float* grad_pos_score,
                                      d3f_stream stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = pl_check(P, D, loss_kind, metric);
    if (rc) return rc;
    D3F_REQUIRE(anchor && positive && dist_keypts && dists && aux && grad_losses && grad_anchor && grad_positive,
                D3F_ERR_INVALID, "null pointer");
    pl_grad_kernel<<<dim3(d3f_ceil_div(P, 128), P), 128, 0, stream>>>(
        dists, P, dist_keypts, keypts_is_f64, safe_radius, loss_kind, metric, pos_margin, neg_margin, log_scale,
        anc_score, pos_score, grad_losses, (float*)aux);
    D3F_CHECK_LAUNCH();
    pl_grad_desc_kernel<<<d3f_ceil_div(2 * P, 8), 256, 0, stream>>>(anchor, positive, P, D, metric, aux, grad_anchor,
                                                                   grad_positive);
    D3F_CHECK_LAUNCH();
    if (anc_score && pos_score && (grad_anc_score || grad_pos_score)) {
        pl_grad_score_kernel<<<d3f_ceil_div(P, 128), 128, 0, stream>>>(P, aux, grad_losses, grad_anc_score,
                                                                     grad_pos_score);
        D3F_CHECK_LAUNCH();
    }
    return D3F_OK;
}

// Detector loss of utils/loss.py:149-158 on any [P,P] fp32 distance matrix (row stride ld).
//   loss [1] out; rowval [P] (fp - cn per row) and arg [2P] (column of the furthest positive or -1, column of the closest
//   negative) are kept for the backward.
extern "C" int d3f_det_loss_forward(const float* dists, int ld, const float* anc_score, const float* pos_score, int P,
                                    float* loss, float* rowval, int32_t* arg, d3f_stream stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    D3F_REQUIRE(P >= 1 && ld >= P, D3F_ERR_INVALID, "bad arguments");
    D3F_REQUIRE(dists && anc_score && pos_score && loss && rowval && arg, D3F_ERR_INVALID, "null pointer");
    det_rows_kernel<<<d3f_ceil_div(P, 8), 256, 0, stream>>>(dists, ld, P, rowval, arg);
    D3F_CHECK_LAUNCH();
    det_final_kernel<<<1, 256, 0, stream>>>(rowval, anc_score, pos_score, P, loss);
    D3F_CHECK_LAUNCH();
    return D3F_OK;
}

// grad_loss [1] (device).  grad_dists [P,P] (row stride ld; zero-filled here) or NULL; grad_anc_score / grad_pos_score [P] or NULL.
extern "C" int d3f_det_loss_backward(const float* rowval, const int32_t* arg, const float* anc_score, const float* pos_score,
                                     int P, const float* grad_loss, float* grad_dists, int ld, float* grad_anc_score,
                                     float* grad_pos_score, d3f_stream stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    D3F_REQUIRE(P >= 1 && (!grad_dists || ld >= P), D3F_ERR_INVALID, "bad arguments");
    D3F_REQUIRE(rowval && arg && anc_score && pos_score && grad_loss, D3F_ERR_INVALID, "null pointer");
    if (grad_dists) D3F_CHECK_CUDA(cudaMemsetAsync(grad_dists, 0, sizeof(float) * (size_t)P * ld, stream));
    det_grad_kernel<<<d3f_ceil_div(P, 256), 256, 0, stream>>>(rowval, arg, anc_score, pos_score, P, grad_loss, grad_dists, ld,
                                                            grad_anc_score, grad_pos_score);
    D3F_CHECK_LAUNCH();
    return D3F_OK;
}
