/*
 * d3feat_b200_debug.h -- diagnostic entry points of libd3feat_b200.so.  NOT part of the drop-in ABI (d3feat_b200.h):
 * process-global selectors used by the parity tests to run the same case on every kernel generation, and measurement
 * hooks used by bench.py / tools/.  Not thread-safe by design; a product never needs to call them.
 */
#ifndef D3FEAT_B200_DEBUG_H
#define D3FEAT_B200_DEBUG_H

#include "d3feat_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* KPConv forward path of d3f_kpconv_forward[_ex]: 1 = gather kernel with FFMA correlation + separate contraction GEMM,
 * 2 = gather kernel with mma.sync 3xTF32 correlation + separate contraction GEMM (csrc/kpconv2.cu), 3 = the fused
 * kernel (csrc/kpconv_fused.cu: gather + correlation + tcgen05 contraction in one launch) wherever a layer is eligible,
 * else 2; -1 restores the default (environment D3F_KPCONV_IMPL = ffma | mma | fused, else 3).  All compute the same
 * function; the selector exists for A/B measurements and parity tests. */
void d3f_set_kpconv_impl(int impl);
int d3f_get_kpconv_impl(void);
/* 1 if d3f_kpconv_forward[_ex] runs a rigid, unmodulated layer of this shape as one fused kernel */
int d3f_kpconv_fused_eligible(int n_neighbors, int K, int c_in, int c_out);
/* Measurement hook: cudaEvent_t handles (or NULL, NULL) recorded on the caller's stream right before and right after
 * the forward gather / fused kernel of the next d3f_kpconv_forward calls, so that kernel can be timed alone. */
void d3f_kpconv_set_gather_events(void* start_event, void* stop_event);

/* GEMM back end: 1 = tcgen05.mma + TMEM (default), 0 = legacy mma.sync. */
void d3f_set_gemm_impl(int use_tcgen05);
/* Tuning hook of tools/gemm_tune.py: force the N tile width (32 / 64 / 128, 0 = heuristic) of the tcgen05 GEMM and the
 * number of atomically combined K splits of plain (epilogue-free) GEMMs (0 = heuristic). */
void d3f_set_gemm_tuning(int bn, int splits);
/* 1 if a tcgen05 kernel ever gave up waiting on its mbarrier (synchronises the device; the asynchronous form is
 * d3f_gemm_status_snapshot in d3feat_b200.h). */
int d3f_gemm_tcgen05_failed(void);

#ifdef __cplusplus
}
#endif
#endif /* D3FEAT_B200_DEBUG_H */
