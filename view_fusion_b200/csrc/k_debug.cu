// Hardware probe (test hook, not on the product path): does a tcgen05 shared-memory descriptor whose start address
// is shifted by a whole number of 128-byte rows — i.e. not aligned to the 1024-byte swizzle atom — address the
// rows TMA wrote with SWIZZLE_128B correctly, and does it need the descriptor's base-offset field?
// The answer decides whether a 3x3 convolution can reuse ONE halo'd activation tile for all nine taps.
#include <cuda.h>

#include "tc_ptx.cuh"
#include "vf_common.cuh"

namespace vf {

int encode_bf16_map(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box);

__global__ void __launch_bounds__(128) umma_shift_probe(const __grid_constant__ CUtensorMap mapA,
                                                        const __grid_constant__ CUtensorMap mapB, int shift_rows,
                                                        int use_base_offset, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const uint32_t bar_ld = base, bar_mma = base + 8, tmem_slot = base + 16;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gbase + 16);
  const uint32_t sA = base + 1024;               // 256 rows x 128 B
  const uint32_t sB = sA + 256 * 128;            // 64 rows x 128 B
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar_ld, 1);
    ptx::mbar_init(bar_mma, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) { ptx::tmem_alloc(tmem_slot, 64); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_d = *tmem_slot_ptr;
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(bar_ld, 256 * 128 + 64 * 128);
    ptx::tma_load_2d(sA, &mapA, bar_ld, 0, 0);
    ptx::tma_load_2d(sB, &mapB, bar_ld, 0, 0);
    ptx::mbar_wait(bar_ld, 0);
    ptx::tc_fence_after();
    const uint32_t a0 = sA + (uint32_t)shift_rows * 128;
    const uint32_t idesc = ptx::make_idesc_bf16(128, 64, 0, 0);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint64_t ad = ptx::make_smem_desc(a0 + k * 32, 16, 1024);
      if (use_base_offset) ad |= (uint64_t)((a0 >> 7) & 7) << 49;
      const uint64_t bd = ptx::make_smem_desc(sB + k * 32, 16, 1024);
      ptx::umma_f16(tmem_d, ad, bd, idesc, k > 0 ? 1u : 0u);
    }
    ptx::umma_commit(bar_mma);
  }
  ptx::mbar_wait(bar_mma, 0);
  ptx::tc_fence_after();
  const int r = warp * 32 + lane;
  for (int c0 = 0; c0 < 64; c0 += 16) {
    uint32_t rr[16];
    ptx::tmem_ld16(tmem_d + ((uint32_t)(warp * 32) << 16) + c0, rr);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) out[r * 64 + c0 + j] = __uint_as_float(rr[j]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_d, 64); }
}

}  // namespace vf

// A: [rows >= 256, 64] bf16, B: [64, 64] bf16 (row n = output column, K-major), out: [128, 64] fp32
// out[i][n] = sum_k A[shift_rows + i][k] * B[n][k] if the shifted descriptor works.
extern "C" __attribute__((visibility("default"))) int vf_debug_umma_shift(const void* A, int rows, const void* B, int shift_rows,
                                                                          int use_base_offset, float* out, vf_stream stream) {
  using namespace vf;
  VF_REQUIRE(A && B && out && rows >= 256 && shift_rows >= 0 && shift_rows <= 128, "vf_debug_umma_shift: bad args");
  CUtensorMap mA, mB;
  {
    const uint64_t dims[2] = {64, (uint64_t)rows};
    const uint64_t strides[1] = {128};
    const uint32_t box[2] = {64, 256};
    int rc = encode_bf16_map(&mA, A, 2, dims, strides, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {64, 64};
    const uint64_t strides[1] = {128};
    const uint32_t box[2] = {64, 64};
    int rc = encode_bf16_map(&mB, B, 2, dims, strides, box);
    if (rc) return rc;
  }
  const size_t smem = 1024 + 1024 + 256 * 128 + 64 * 128;
  VF_CUDA(cudaFuncSetAttribute(umma_shift_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_shift_probe<<<1, 128, smem, as_stream(stream)>>>(mA, mB, shift_rows, use_base_offset, out);
  VF_LAUNCH_CHECK();
  return VF_OK;
}

// ---- probe 2: sustained tcgen05.mma issue/execute rate from one thread (no loads; smem contents are irrelevant) ------
namespace vf {
// KIND 0: plain MMA, 1: .ws without collector hints, 2: .ws keeping the weight tile in collector buffer b0 across the G row tiles
template <int KIND, int G>
__device__ __forceinline__ void ws_pattern(uint32_t tmem_d, uint64_t ad, uint64_t bd, uint32_t idesc, int N, int n_groups, uint32_t sink_bar) {
  for (int i = 0; i < n_groups; i += 4 * G) {
#pragma unroll
    for (int tap = 0; tap < 4; ++tap) {
      const uint64_t a_t = ad + (uint64_t)(tap * 67 * 8), b_t = bd + (uint64_t)(tap * N * 8);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const uint32_t d = tmem_d + (uint32_t)(g * N);
          const uint64_t ag = a_t + (uint64_t)(2 * k + g * 1024), bk = b_t + (uint64_t)(2 * k);
          if (KIND == 0) ptx::umma_f16(d, ag, bk, idesc, 1u);
          else if (KIND == 1) ptx::umma_ws_f16<0>(d, ag, bk, idesc, 1u);
          else if (g == 0) ptx::umma_ws_f16<1>(d, ag, bk, idesc, 1u);
          else if (g == G - 1) ptx::umma_ws_f16<3>(d, ag, bk, idesc, 1u);
          else ptx::umma_ws_f16<2>(d, ag, bk, idesc, 1u);
        }
      }
      ptx::umma_commit(sink_bar);
    }
  }
}

__global__ void __launch_bounds__(128) umma_rate_probe(int N, int shift_rows, int n_groups, int commit_every, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const uint32_t bar = base, tmem_slot = base + 16;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gbase + 16);
  const uint32_t sA = base + 1024, sB = sA + 512 * 128;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_barrier_init(); }
  if (warp == 0) { ptx::tmem_alloc(tmem_slot, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_d = *tmem_slot_ptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::make_idesc_bf16(128, N, 0, 0);
    const uint64_t d0 = ptx::make_smem_desc(0, 16, 1024);
    const uint64_t ad = d0 + (((sA + (uint32_t)shift_rows * 128u) & 0x3FFFF) >> 4), bd = d0 + ((sB & 0x3FFFF) >> 4);
    uint32_t phase = 0;
    const long long t0 = clock64();
    if (commit_every == -100 || commit_every == -101) {
      // realistic operand pattern: 4 "taps" with distinct weight tiles and row-shifted A, 2 accumulators, k4 per (tap, g);
      // -101 orders the MMAs k-outer / g-inner so consecutive instructions share the B operand
      const uint32_t sink_bar = base + 40;
      ptx::mbar_init(sink_bar, 1);
      ptx::fence_barrier_init();
      for (int i = 0; i < n_groups; i += 8) {
        for (int tap = 0; tap < 4; ++tap) {
          const uint64_t a_t = ad + (uint64_t)(tap * 67 * 8), b_t = bd + (uint64_t)(tap * N * 8);
          if (commit_every == -100) {
            for (int g = 0; g < 2; ++g) ptx::umma_f16_k4(tmem_d + (uint32_t)(g * N), a_t + (uint64_t)(g * 1024), b_t, idesc, 1u);
          } else {
            for (int k = 0; k < 4; ++k)
              for (int g = 0; g < 2; ++g)
                ptx::umma_f16(tmem_d + (uint32_t)(g * N), a_t + (uint64_t)(g * 1024 + 2 * k), b_t + (uint64_t)(2 * k), idesc, 1u);
          }
          ptx::umma_commit(sink_bar);
        }
      }
    } else if (commit_every <= -200 && commit_every >= -209) {
      // weight-gradient pattern: both operands MN-major, 5 accumulators (tap pairs) whose A descriptors differ in start
      // row and leading-dimension offset, 8 k-steps of 16 rows per 128-row chunk.  Order: -200 accumulator outer,
      // -201 k-step outer, -202 chains of four k-steps, -203 chains of two, -204 accumulator outer with identical A
      const uint32_t idm = ptx::make_idesc_bf16(128, N, 1, 1);
      const int shifts[5] = {0, 2, 66, 131, 132};
      uint64_t adp[5];
      for (int a = 0; a < 5; ++a)
        adp[a] = ptx::make_smem_desc(sA + (uint32_t)(commit_every == -204 ? 0 : shifts[a]) * 128u, a == 4 ? 128u : (a == 1 ? 63u * 128u : 128u), 1024);
      const uint64_t bdm = ptx::make_smem_desc(sB, 128 * 128, 1024);
      const uint32_t sink_bar = base + 40;
      ptx::mbar_init(sink_bar, 1);
      ptx::fence_barrier_init();
      const int mode = -200 - commit_every;
      const int chain = mode == 1 ? 1 : (mode == 2 ? 4 : (mode == 3 ? 2 : 8));
      for (int i = 0; i < n_groups; i += 10) {           // 10 groups of 4 = 40 MMAs = one chunk
        for (int kb = 0; kb < 8; kb += chain)
          for (int a = 0; a < 5; ++a)
            for (int k = kb; k < kb + chain; ++k)
              ptx::umma_f16(tmem_d + (uint32_t)(a * N), adp[a] + (uint64_t)(k * 128), bdm + (uint64_t)(k * 128), idm, 1u);
        ptx::umma_commit(sink_bar);
      }
    } else if (commit_every <= -300 && commit_every >= -309) {
      // weight-stationary MMAs: G row tiles against the same weight tile, k-step outer / row tile inner.
      // -300: plain MMA, G=2; -301: .ws without hints, G=2; -302: .ws fill/lastuse, G=2; -303: plain, G=4; -304: .ws no hints, G=4;
      // -305: .ws fill/use/use/lastuse, G=4
      const uint32_t sink_bar = base + 40;
      ptx::mbar_init(sink_bar, 1);
      ptx::fence_barrier_init();
      switch (-300 - commit_every) {
        case 0: ws_pattern<0, 2>(tmem_d, ad, bd, idesc, N, n_groups, sink_bar); break;
        case 1: ws_pattern<1, 2>(tmem_d, ad, bd, idesc, N, n_groups, sink_bar); break;
        case 2: ws_pattern<2, 2>(tmem_d, ad, bd, idesc, N, n_groups, sink_bar); break;
        case 3: ws_pattern<0, 4>(tmem_d, ad, bd, idesc, N, n_groups, sink_bar); break;
        case 4: ws_pattern<1, 4>(tmem_d, ad, bd, idesc, N, n_groups, sink_bar); break;
        default: ws_pattern<2, 4>(tmem_d, ad, bd, idesc, N, n_groups, sink_bar); break;
      }
    } else if (commit_every >= 0) {
      for (int i = 0; i < n_groups; ++i) {
        ptx::umma_f16_k4(tmem_d, ad, bd, idesc, 1u);
        if (commit_every > 0 && (i + 1) % commit_every == 0 && i + 1 < n_groups) { ptx::umma_commit(bar); ptx::mbar_wait(bar, phase); phase ^= 1; }
      }
    } else {
      // mimic the conv main loop: per "tap" = -commit_every groups: wait on an already-complete barrier, fence, MMAs,
      // commit to a barrier nobody waits on
      const uint32_t done_bar = base + 32, sink_bar = base + 40;
      ptx::mbar_init(done_bar, 1); ptx::mbar_init(sink_bar, 1);
      ptx::fence_barrier_init();
      ptx::mbar_arrive(done_bar);                       // phase 0 complete
      const int per = -commit_every;
      for (int i = 0; i < n_groups; i += per) {
        ptx::mbar_wait(done_bar, 0);
        ptx::tc_fence_after();
        for (int g = 0; g < per; ++g) ptx::umma_f16_k4(tmem_d + (uint32_t)(g * 0), ad + (uint64_t)(g * 1024), bd, idesc, 1u);
        ptx::umma_commit(sink_bar);
      }
    }
    ptx::umma_commit(bar);
    ptx::mbar_wait(bar, phase);
    out[blockIdx.x] = clock64() - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_d, 512); }
}
}  // namespace vf

// cycles_out[grid] = SM cycles for n_groups x 4 MMAs (M=128, N, K=16) issued by one thread per CTA; commit_every > 0 waits
// for completion every that many groups (exposes latency), 0 = fully pipelined.
extern "C" __attribute__((visibility("default"))) int vf_debug_umma_rate(int N, int shift_rows, int n_groups, int commit_every, int grid,
                                                                         long long* cycles_out, vf_stream stream) {
  using namespace vf;
  VF_REQUIRE(N % 16 == 0 && N >= 16 && N <= 256 && grid > 0 && cycles_out, "vf_debug_umma_rate: bad args");
  const size_t smem = 1024 + 1024 + 512 * 128 + 4 * 256 * 128;
  VF_CUDA(cudaFuncSetAttribute(umma_rate_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_rate_probe<<<grid, 128, smem, as_stream(stream)>>>(N, shift_rows, n_groups, commit_every, cycles_out);
  VF_LAUNCH_CHECK();
  return VF_OK;
}

// ---- probe 3: MN-major operands (the weight-gradient GEMM reduces over pixel ROWS, so both operands have the reduction
// dimension K as the slow axis: A[k][m], B[k][n] as TMA loads them from NHWC matrices) ------------------------------------
namespace vf {
__global__ void __launch_bounds__(128) umma_mn_probe(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                     int shift_rows, int lbo_bytes, int sbo_bytes, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const uint32_t bar_ld = base, bar_mma = base + 8, tmem_slot = base + 16;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gbase + 16);
  const uint32_t sA = base + 1024;               // 2 boxes of [64 k-rows x 64 m] = 2 x 8 KB  (M = 128)
  const uint32_t sB = sA + 2 * 8192;             // [128 k-rows x 64 n] (rows 0..127 loaded so that shifted reads stay inside)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { ptx::mbar_init(bar_ld, 1); ptx::mbar_init(bar_mma, 1); ptx::fence_barrier_init(); }
  if (warp == 0) { ptx::tmem_alloc(tmem_slot, 64); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_d = *tmem_slot_ptr;
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(bar_ld, 2 * 8192 + 2 * 8192);
    ptx::tma_load_2d(sA, &mapA, bar_ld, 0, 0);          // m 0..63,  k-rows 0..63
    ptx::tma_load_2d(sA + 8192, &mapA, bar_ld, 64, 0);  // m 64..127
    ptx::tma_load_2d(sB, &mapB, bar_ld, 0, 0);          // n 0..63, k-rows 0..63
    ptx::tma_load_2d(sB + 8192, &mapB, bar_ld, 0, 64);  // k-rows 64..127
    ptx::mbar_wait(bar_ld, 0);
    ptx::tc_fence_after();
    const uint32_t idesc = ptx::make_idesc_bf16(128, 64, 1, 1);     // both operands MN-major
    for (int k = 0; k < 4; ++k) {                                    // K = 64 rows = 4 x 16
      const uint64_t ad = ptx::make_smem_desc(sA + k * 2048, (uint32_t)lbo_bytes, (uint32_t)sbo_bytes);
      const uint64_t bd = ptx::make_smem_desc(sB + (uint32_t)shift_rows * 128 + k * 2048, (uint32_t)lbo_bytes, (uint32_t)sbo_bytes);
      ptx::umma_f16(tmem_d, ad, bd, idesc, k > 0 ? 1u : 0u);
    }
    ptx::umma_commit(bar_mma);
  }
  ptx::mbar_wait(bar_mma, 0);
  ptx::tc_fence_after();
  const int r = warp * 32 + lane;
  for (int c0 = 0; c0 < 64; c0 += 16) {
    uint32_t rr[16];
    ptx::tmem_ld16(tmem_d + ((uint32_t)(warp * 32) << 16) + c0, rr);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) out[r * 64 + c0 + j] = __uint_as_float(rr[j]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_d, 64); }
}
}  // namespace vf

// A: [rows >= 64, 128] bf16 (row = k, col = m), B: [rows >= 128, 64] bf16 (row = k, col = n); out [128, 64] fp32:
// out[m][n] = sum_{k<64} A[k][m] * B[shift_rows + k][n]
extern "C" __attribute__((visibility("default"))) int vf_debug_umma_mn(const void* A, int rowsA, const void* B, int rowsB, int shift_rows,
                                                                       int lbo_bytes, int sbo_bytes, float* out, vf_stream stream) {
  using namespace vf;
  VF_REQUIRE(A && B && out && rowsA >= 64 && rowsB >= 128 && shift_rows >= 0 && shift_rows <= 64, "vf_debug_umma_mn: bad args");
  CUtensorMap mA, mB;
  {
    const uint64_t dims[2] = {128, (uint64_t)rowsA};
    const uint64_t strides[1] = {256};
    const uint32_t box[2] = {64, 64};
    int rc = encode_bf16_map(&mA, A, 2, dims, strides, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {64, (uint64_t)rowsB};
    const uint64_t strides[1] = {128};
    const uint32_t box[2] = {64, 64};
    int rc = encode_bf16_map(&mB, B, 2, dims, strides, box);
    if (rc) return rc;
  }
  const size_t smem = 1024 + 1024 + 2 * 8192 + 2 * 8192;
  VF_CUDA(cudaFuncSetAttribute(umma_mn_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_mn_probe<<<1, 128, smem, as_stream(stream)>>>(mA, mB, shift_rows, lbo_bytes, sbo_bytes, out);
  VF_LAUNCH_CHECK();
  return VF_OK;
}

// ---- probe 4: special-function throughput (what bounds the SiLU of the GroupNorm pass) -------------------------------------
namespace vf {
// mode 0: tanh.approx.f32, 1: ex2.approx + rcp.approx (sigmoid), 2: tanh.approx.f16x2 (two elements per operation), 3: ex2 only,
// 4: rcp only, 5: FMA only (baseline of the loop)
template <int MODE>
__global__ void __launch_bounds__(256) mufu_probe_kernel(int iters, float seed, float* out, long long* cyc) {
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = seed + 0.001f * (float)(threadIdx.x + j);
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      float a = v[j], b = v[j + 1];
      if (MODE == 0) {
        float ta, tb;
        asm volatile("tanh.approx.f32 %0, %1;" : "=f"(ta) : "f"(a));
        asm volatile("tanh.approx.f32 %0, %1;" : "=f"(tb) : "f"(b));
        a = fmaf(a, ta, a); b = fmaf(b, tb, b);
      } else if (MODE == 1) {
        float ea, eb, ra, rb;
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(ea) : "f"(-a));
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(eb) : "f"(-b));
        asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(ra) : "f"(1.f + ea));
        asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(rb) : "f"(1.f + eb));
        a = a * ra; b = b * rb;
      } else if (MODE == 2) {
        uint32_t h, t;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(b), "f"(a));
        asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(h));
        float ta, tb;
        asm volatile("{.reg .f16 lo, hi; mov.b32 {lo, hi}, %2; cvt.f32.f16 %0, lo; cvt.f32.f16 %1, hi;}" : "=f"(ta), "=f"(tb) : "r"(t));
        a = fmaf(a, ta, a); b = fmaf(b, tb, b);
      } else if (MODE == 3) {
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(a) : "f"(a));
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(b) : "f"(b));
      } else if (MODE == 4) {
        asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(a) : "f"(a));
        asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(b) : "f"(b));
      } else {
        a = fmaf(a, 0.999f, 0.001f); b = fmaf(b, 0.999f, 0.001f);
      }
      v[j] = a * 0.9f + 0.05f; v[j + 1] = b * 0.9f + 0.05f;
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += v[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
}  // namespace vf

// cycles_out[grid] = cycles of `iters` x 8 elements per thread, 256 threads per CTA, `ctas_per_sm` x 148 CTAs
extern "C" __attribute__((visibility("default"))) int vf_debug_mufu_rate(int mode, int iters, int grid, float* scratch, long long* cycles_out, vf_stream stream) {
  using namespace vf;
  cudaStream_t st = as_stream(stream);
  switch (mode) {
    case 0: mufu_probe_kernel<0><<<grid, 256, 0, st>>>(iters, 0.3f, scratch, cycles_out); break;
    case 1: mufu_probe_kernel<1><<<grid, 256, 0, st>>>(iters, 0.3f, scratch, cycles_out); break;
    case 2: mufu_probe_kernel<2><<<grid, 256, 0, st>>>(iters, 0.3f, scratch, cycles_out); break;
    case 3: mufu_probe_kernel<3><<<grid, 256, 0, st>>>(iters, 0.3f, scratch, cycles_out); break;
    case 4: mufu_probe_kernel<4><<<grid, 256, 0, st>>>(iters, 0.3f, scratch, cycles_out); break;
    default: mufu_probe_kernel<5><<<grid, 256, 0, st>>>(iters, 0.3f, scratch, cycles_out); break;
  }
  VF_LAUNCH_CHECK();
  return VF_OK;
}
