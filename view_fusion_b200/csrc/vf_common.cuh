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

inline cudaStream_t as_stream(vf_stream s) { return reinterpret_cast<cudaStream_t>(s); }
inline size_t dtype_size(int dt) { return dt == VF_BF16 ? 2 : 4; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int cdiv(int a, int b) { return (a + b - 1) / b; }
int sm_count();

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
  uint4 t = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
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
