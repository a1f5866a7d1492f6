// Shared helpers for the viewfusion_b200 CUDA sources (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/viewfusion_b200.h"

namespace vf {

// ---- error plumbing --------------------------------------------------------------------------------
void set_error(const char* fmt, ...);

#define VF_REQUIRE(cond, ...)                 \
  do {                                        \
    if (!(cond)) {                            \
      ::vf::set_error(__VA_ARGS__);           \
      return VF_ERR_ARG;                      \
    }                                         \
  } while (0)

#define VF_CUDA(expr)                                                                            \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      ::vf::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return VF_ERR_CUDA;                                                                        \
    }                                                                                            \
  } while (0)

#define VF_LAUNCH_CHECK() VF_CUDA(cudaGetLastError())

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------------
// Consecutive kernels of a step are launched with programmatic stream serialization: a kernel calls
// pdl_launch_dependents() first thing (the next kernel may then be scheduled onto SMs as they drain) and pdl_wait()
// before its first access to global memory (returns once every earlier kernel has completed and flushed).  Launch
// latency, barrier / TMEM set-up and the tail of the previous grid overlap; both calls are no-ops in a plain launch.
// EVERY thread that touches global memory must have passed pdl_wait(): the predecessor may itself still be waiting.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled();      // VF_PDL=0 in the environment turns the launch attribute off (vf_api.cu)
void pdl_set_suspended(bool s);

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Row splits per image for a (splits, images) grid of a bandwidth kernel: around `want`, at most `max_splits`, chosen so
// that splits*images CTAs fill whole waves of `resident` co-resident CTAs (a 3.03-wave grid runs as long as a 4-wave one).
inline int wave_splits(int images, int want, int max_splits, int resident) {
  if (max_splits < 1) max_splits = 1;
  if (want < 1) want = 1;
  if (want > max_splits) want = max_splits;
  int lo = want / 2 > 0 ? want / 2 : 1, hi = want * 2 < max_splits ? want * 2 : max_splits;
  int best = want;
  double best_eff = -1.0;
  for (int s = lo; s <= hi; ++s) {
    const long ctas = (long)s * images;
    const long waves = (ctas + resident - 1) / resident;
    const double eff = (double)ctas / (double)(waves * resident);
    const int d = s > want ? s - want : want - s, bd = best > want ? best - want : want - best;
    if (eff > best_eff + 0.02 || (eff > best_eff - 0.02 && d < bd && eff >= best_eff - 1e-9)) { best = s; best_eff = eff > best_eff ? eff : best_eff; }
  }
  return best;
}

// Row splits for SMALL layers of the same kind of kernel.  There a CTA's fixed cost (the statistics prologue: two
// dependent global round trips, two barriers and the per-group loops, ~3 us) rivals its streaming time, so filling the
// machine with many short CTAs multiplies the prologue by the number of waves (ncu: 25 us for the 33 MB of a 16x16 192-
// channel GroupNorm backward pass in 4 waves).  Model: time = waves * (t_pro + iterations * t_iter), t_iter = one memory
// round trip per unrolled row batch; the fewest CTAs win ties.
inline int latency_splits(int images, int rows, int rows_per_iter, int max_splits, int resident, double t_pro_us, double t_iter_us) {
  if (max_splits < 1) max_splits = 1;
  if (rows_per_iter < 1) rows_per_iter = 1;
  int best = 1;
  double best_cost = 1e30;
  for (int s = 1; s <= max_splits; ++s) {
    const long ctas = (long)s * images;
    const long waves = (ctas + resident - 1) / resident;
    const int rpc = (rows + s - 1) / s;
    const int iters = (rpc + rows_per_iter - 1) / rows_per_iter;
    const double cost = (double)waves * (t_pro_us + iters * t_iter_us);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
  }
  return best;
}
// layers up to 32x32 x 256 channels per image (<= 94 MB per tensor at 168 view-images) are scheduled by latency_splits
inline bool latency_bound_layer(int rows_per_img, int C) { return (long)rows_per_img * C <= 1089L * 256; }

inline cudaStream_t as_stream(vf_stream s) { return reinterpret_cast<cudaStream_t>(s); }
inline size_t dtype_size(int dt) { return dt == VF_BF16 ? 2 : 4; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int cdiv(int a, int b) { return (a + b - 1) / b; }
int sm_count();

// True exactly once per (call site, CUDA device): kernel attributes (opt-in shared memory) are per-device state, so a
// process that drives several GPUs must set them on each one.
struct PerDeviceOnce {
  unsigned long long mask = 0;
  bool first() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) return true;
    const unsigned long long bit = 1ull << (d & 63);
    const unsigned long long old = __atomic_fetch_or(&mask, bit, __ATOMIC_ACQ_REL);
    return !(old & bit);
  }
};
#define VF_SET_MAX_SMEM(kernel, bytes)                                                                          \
  do {                                                                                                          \
    static ::vf::PerDeviceOnce _once;                                                                           \
    if (_once.first()) VF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
  } while (0)

// ---- device-side scalar/vector helpers -------------------------------------------------------------
template <typename T> struct VecOf;            // 16-byte vector of T
template <> struct VecOf<float> { static constexpr int N = 4; };
template <> struct VecOf<__nv_bfloat16> { static constexpr int N = 8; };

__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(__nv_bfloat16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

// 16-byte load of VecOf<T>::N elements into fp32 registers
__device__ __forceinline__ void load_vec(const float* p, float (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load_vec(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 t = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {     // bf16 -> fp32 is a 16-bit shift: one SHL and one LOP per pair (the intrinsic costs PRMT + SHL for the high half)
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ void store_vec(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store_vec(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 t;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = t;
}

__device__ __forceinline__ float silu(float x) { return x / (1.f + __expf(-x)); }
// silu(2h) = h + h*tanh(h)
__device__ __forceinline__ float silu_half(float h) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
// x * sigmoid(x) with sigmoid = 0.5 * tanh(0.5 x) + 0.5: one MUFU op per element (tanh.approx, rel. error 2^-11, well
// below the bf16 rounding of the stored result) instead of ex2 + rcp.  The SiLU pass is otherwise MUFU-bound.
__device__ __forceinline__ float silu_fast(float x) {
#ifdef VF_SILU_EXACT      // A/B build: 2 MUFU (ex2 + rcp), relative error ~1e-7 instead of 2^-11
  return __fdividef(x, 1.f + __expf(-x));
#else
  // h + h*tanh(h), h = x/2: three operations per element (the convolution's fused operand transform folds the halving into its
  // coefficients and must produce the very same bits as vf_gn_apply)
  return silu_half(0.5f * x);
#endif
}
template <typename T> __device__ __forceinline__ float silu_for(float x);
template <> __device__ __forceinline__ float silu_for<float>(float x) { return silu(x); }
template <> __device__ __forceinline__ float silu_for<__nv_bfloat16>(float x) { return silu_fast(x); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- FLAT / PADDED row orders (include/viewfusion_b200.h) --------------------------------------------------
struct RowGeom {
  int H, W, HW, W1, P, images;
  int rows_total;                    // GEMM M: images*P (PADDED sources) or images*H*W (FLAT)
  int in_padded, out_padded, stride2;
  float rcp_rows, rcp_w;             // 1 / rows per image (P or HW), 1 / line length (W+1 or W): exact small-integer division
};
// Bias / embedding gradient of the convolution that produced a GroupNorm input, emitted by the GroupNorm backward itself:
// dx = cA*dz + cB*x + cC per (image, channel), so sum_pixels dx = cA*sum(dz) + cB*sum(x) + cC*HW needs no pass over dx.
struct GnColsum {
  float* db;            // [C] bias gradient (+=) or NULL
  float* demb;          // embedding-table gradient [rows, emb_ld] (+= at column col) or NULL
  const int* img_row;   // [images] embedding row of each image
  int emb_ld, col;
};
// s[img][ch] of a vf_gn_shift (emb_row = emb + img_row[img] * emb_ld, or NULL)
__device__ __forceinline__ float gn_shift_value(const vf_gn_shift& sh, const float* emb_row, int ch) {
  float v = sh.bias ? __ldg(sh.bias + ch) : 0.f;
  if (emb_row) v += __ldg(emb_row + ch);
  return v;
}
struct RowInfo {
  int img, pix;      // pix = y*Wo + x at the OUTPUT resolution
  bool valid;
  long out_row;
};
inline RowGeom make_geom(int images, int H, int W, int in_padded, int out_padded, int stride2) {
  RowGeom g;
  g.H = H; g.W = W; g.HW = H * W; g.W1 = W + 1; g.P = (H + 1) * (W + 1); g.images = images;
  g.in_padded = in_padded; g.out_padded = out_padded; g.stride2 = stride2;
  g.rows_total = images * (in_padded ? g.P : g.HW);
  g.rcp_rows = 1.0f / (float)(in_padded ? g.P : g.HW);
  g.rcp_w = 1.0f / (float)(in_padded ? g.W1 : g.W);
  return g;
}
// n / d for 0 <= n < 2^24 through the float reciprocal (|error| < 1 before truncation, one correction step): a few
// instructions instead of the ~100-cycle integer division sequence on the epilogue's critical path.
template <bool kChecked = true>
__device__ __forceinline__ int div_small(int n, int d, float rcp) {
  if (kChecked && n >= (1 << 24)) return n / d;
  int q = __float2int_rz(__int2float_rn(n) * rcp);
  const int r = n - q * d;
  q += (r >= d) - (r < 0);
  return q;
}
// kFast: the caller guarantees rows_total < 2^24 (no integer-division fallback is compiled in)
template <bool kFast = false>
__device__ __forceinline__ RowInfo decode_row(const RowGeom& p, int m) {
  RowInfo r;
  r.valid = m < p.rows_total;
  int img = 0, y = 0, x = 0;
  if (p.in_padded) {
    img = div_small<!kFast>(m, p.P, p.rcp_rows);
    const int rem = m - img * p.P;
    const int yy = div_small<!kFast>(rem, p.W1, p.rcp_w), xx = rem - yy * p.W1;
    r.valid = r.valid && yy >= 1 && xx >= 1;
    y = yy - 1; x = xx - 1;
  } else {
    img = div_small<!kFast>(m, p.HW, p.rcp_rows);
    const int rem = m - img * p.HW;
    y = div_small<!kFast>(rem, p.W, p.rcp_w); x = rem - y * p.W;
  }
  int Wo = p.W, Ho = p.H;
  if (p.stride2) {
    r.valid = r.valid && ((y | x) & 1) == 0;
    y >>= 1; x >>= 1; Wo >>= 1; Ho >>= 1;
  }
  r.img = img;
  r.pix = y * Wo + x;
  r.out_row = p.out_padded ? (long)img * (Ho + 1) * (Wo + 1) + (long)(y + 1) * (Wo + 1) + (x + 1) : (long)img * Ho * Wo + r.pix;
  return r;
}

// Column sums of a 32 x 16 register tile (one row per lane): 16 shuffles instead of 16 x 5.  Every lane returns the
// total of column warp_col16(lane); lanes l and l^1 hold the same column.
__device__ __forceinline__ float warp_colsum16(const float (&v)[16], int lane) {
  float a[8], b[4], c[2];
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float send = h16 ? v[j] : v[j + 8], keep = h16 ? v[j + 8] : v[j];
    a[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float send = h8 ? a[j] : a[j + 4], keep = h8 ? a[j + 4] : a[j];
    b[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float send = h4 ? b[j] : b[j + 2], keep = h4 ? b[j + 2] : b[j];
    c[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const float send = h2 ? c[0] : c[1], keep = h2 ? c[1] : c[0];
  float d = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  d += __shfl_xor_sync(0xffffffffu, d, 1);
  return d;
}
__device__ __forceinline__ int warp_col16(int lane) { return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1); }

}  // namespace vf
