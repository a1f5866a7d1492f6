/*
 * viewfusion_b200 — hardware probes (tests / profiling scripts only).
 *
 * These three entry points are NOT part of the product library: they live in libviewfusion_b200_probes.so (the product
 * objects + csrc/k_debug.cu), which view_fusion_b200/build.py links next to libviewfusion_b200.so and only the tests and
 * scripts/ load (view_fusion_b200._lib.load_probes()).  They answer questions about tcgen05 behaviour that the convolution
 * relies on (shifted SWIZZLE_128B descriptors, MMA issue rate, MN-major operands); DESIGN.md 5.1 cites their results.
 */
#ifndef VIEWFUSION_B200_PROBES_H_
#define VIEWFUSION_B200_PROBES_H_

#include "viewfusion_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Hardware probe (tests only): out[i][n] = sum_k A[shift_rows + i][k] * B[n][k] through ONE TMA-loaded, 128B-swizzled
 * smem tile and a tcgen05 descriptor whose start is shifted by `shift_rows` rows. A [rows>=256, 64] bf16, B [64, 64]. */
/* Hardware probe: cycles for n_groups x 4 back-to-back tcgen05.mma (M=128, N, K=16) issued by one thread per CTA. */
/* Special-function throughput probe (mode 0 tanh.f32, 1 ex2+rcp, 2 tanh.f16x2, 3 ex2, 4 rcp, 5 FMA only): cycles per CTA. */
int vf_debug_mufu_rate(int mode, int iters, int grid, float* scratch, long long* cycles_out, vf_stream stream);
int vf_debug_umma_rate(int N, int shift_rows, int n_groups, int commit_every, int grid, long long* cycles_out, vf_stream stream);
/* Hardware probe: MN-major operands: out[m][n] = sum_{k<64} A[k][m] * B[shift_rows + k][n] (A [.,128], B [.,64] bf16). */
int vf_debug_umma_mn(const void* A, int rowsA, const void* B, int rowsB, int shift_rows, int lbo_bytes, int sbo_bytes, float* out,
                     vf_stream stream);
int vf_debug_umma_shift(const void* A, int rows, const void* B, int shift_rows, int use_base_offset, float* out,
                        vf_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* VIEWFUSION_B200_PROBES_H_ */
