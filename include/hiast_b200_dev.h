/*
 * hiast_b200 -- development / test interface.  NOT part of the drop-in boundary (include/hiast_b200.h): host-side test
 * hooks of the scan arithmetic, A/B toggles used by tools/ and tests/, and a device self test.  The symbols are exported by
 * the same library so that the tests can reach them; nothing under hiast_b200/*.py's product path calls them.
 */
#ifndef HIAST_B200_DEV_H_
#define HIAST_B200_DEV_H_

#include "hiast_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* 1 when the library was compiled with -DHIAST_DEV_VARIANTS (the measured-and-dropped phase-A variants, the TMA-staged
 * kernel and the fused persistent window kernel are then selectable); 0 for the product build.                        */
HIAST_API int hiast_dev_variants(void);

/* ---- host-side test hooks (no GPU needed; used by tests only) --------------------------- */
/* x^n by double-double repeated squaring, the integer-gamma power used by the scan.          */
HIAST_API double hiast_testhook_powi(double x, int n);
/* One class, one group of hiast_ias_threshold_scan on the HOST from an inclusive-prefix
 * histogram row; returns the new threshold, *temp_out = the float32 quantile.               */
HIAST_API double hiast_testhook_threshold_step(const uint32_t* prefix_row_host, int key_lo,
                                     double thr, double alpha, double beta, double gamma,
                                     float* temp_out, int* error_out);

/* ---- development hooks -------------------------------------------------------------------- */
/* Per-unit timeline of the next hiast_ias_fused_window launches: dev_buffer = u64 [n_SMs][256][6]
 * (kind << 32 | unit, begin, end, closer: wait begin, wait end, published; %globaltimer ns), zeroed by the caller;
 * NULL switches tracing off.                                                                   */
HIAST_API int hiast_debug_validate_direct(int on);   /* 1: hiast_probs_upsample_argmax always takes the direct (unstaged) kernel */
HIAST_API int hiast_debug_png_variant(int v);         /* emit kernel token loop: 0 nested (divergent), 1 (default) one token per iteration */
HIAST_API int hiast_debug_set_fused_trace(void* dev_buffer);
/* on != 0: hiast_st_loss_fwd / _bwd use the scalar vector kernels instead of the packed-pair (f32x2) ones for
 * the SoftCE consistency kind (A/B measurements and cross-checks).                               */
HIAST_API int hiast_debug_loss_scalar(int on);
/* on != 0: hiast_ias_upsample_softmax_hist uses its first kernel (4 horizontally adjacent pixels per thread).   */
HIAST_API int hiast_debug_upsample_v1(int on);

/* ---- device-side self test (needs a GPU; used by tests only) ---------------------------- */
/* Sweeps EVERY non-positive float (bit patterns 0x80000000..0xFF800000 and +0) through the packed
 * (f32x2) exponential used by phase A and compares it bit for bit with CUDA's expf();
 * *mismatches_dev (device u64) receives the number of differing inputs.                      */
HIAST_API int hiast_selftest_packed_expf(unsigned long long* mismatches_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HIAST_B200_DEV_H_ */
