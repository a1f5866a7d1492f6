// GroupNorm(32, C, eps=1e-5, affine) + Swish  (reference: model/unet.py:207-218 Block, :254 attention norm).
//
// Two HBM-bound passes over NHWC activations with 128-bit accesses:
//   vf_gn_stats : per (image, channel) sum / sum-of-squares; the convolution epilogues can emit the same
//                 partials instead (vf_conv_args::stats), in which case this pass is skipped.
//   vf_gn_apply : turns the channel sums into per-group mean / rstd, folds them with gamma/beta into one
//                 per-channel FMA held in shared memory, applies x*a+b (+ Swish) and writes the tensor the next
//                 convolution's TMA reads.  It reads up to two sources so that torch.cat((x, skip), 1)
//                 (unet.py:134) is never materialised un-normalised.
// Algorithmic bytes (bf16): stats 2 B/elem, apply 4 B/elem.
#include "vf_common.cuh"

namespace vf {

constexpr int kGnThreads = 256;

// block = (CV channel-vectors) x (PY pixel lanes); every thread keeps sums for its fixed 16-byte channel vector
template <typename T>
__global__ void __launch_bounds__(kGnThreads) gn_stats_kernel(const T* __restrict__ s0, int C0, const T* __restrict__ s1,
                                                              int C1, int HW, int pix_per_cta, float* __restrict__ stats) {
  constexpr int VEC = VecOf<T>::N;
  extern __shared__ float acc[];                     // [C][2]
  const int C = C0 + C1, CV = C / VEC;
  const int img = blockIdx.y;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  const int PY = blockDim.x / CV;                    // pixel lanes (blockDim = CV*PY exactly)
  const int cv = threadIdx.x % CV, py = threadIdx.x / CV;
  const int c = cv * VEC;
  const T* src = c < C0 ? s0 + (size_t)img * HW * C0 + c : s1 + (size_t)img * HW * C1 + (c - C0);
  const int ld = c < C0 ? C0 : C1;
  const int p0 = blockIdx.x * pix_per_cta, p1 = min(HW, p0 + pix_per_cta);
  float s[VEC], q[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) s[j] = q[j] = 0.f;
  for (int p = p0 + py; p < p1; p += PY) {
    float v[VEC];
    load_vec(src + (size_t)p * ld, v);
#pragma unroll
    for (int j = 0; j < VEC; ++j) { s[j] += v[j]; q[j] += v[j] * v[j]; }
  }
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    atomicAdd(&acc[2 * (c + j)], s[j]);
    atomicAdd(&acc[2 * (c + j) + 1], q[j]);
  }
  __syncthreads();
  float* dst = stats + (size_t)img * C * 2;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(dst + i, acc[i]);
}

template <typename T, bool kSwish>
__global__ void __launch_bounds__(kGnThreads) gn_apply_kernel(const T* __restrict__ s0, int C0, const T* __restrict__ s1,
                                                              int C1, int HW, int groups, int pix_per_cta,
                                                              const float* __restrict__ st0, int ld0,
                                                              const float* __restrict__ st1, int ld1,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, T* __restrict__ dst) {
  constexpr int VEC = VecOf<T>::N;
  extern __shared__ float ab[];                      // [C][2]: y = x*a + b
  const int C = C0 + C1, CV = C / VEC;
  const int img = blockIdx.y;
  const int gs = C / groups;
  const float inv_n = 1.f / ((float)gs * (float)HW);
  const float* sa = st0 + (size_t)img * ld0 * 2;
  const float* sb = st1 ? st1 + (size_t)img * ld1 * 2 : nullptr;
  for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
    const int g0 = ch / gs * gs;
    float s = 0.f, q = 0.f;
    for (int j = 0; j < gs; ++j) {           // a group may straddle the two sources of a concatenation
      const int cc = g0 + j;
      const float* e = cc < C0 ? sa + 2 * cc : sb + 2 * (cc - C0);
      s += __ldg(e); q += __ldg(e + 1);
    }
    const float mean = s * inv_n;
    const float var = fmaxf(q * inv_n - mean * mean, 0.f);
    const float rstd = rsqrtf(var + 1e-5f);
    const float a = rstd * __ldg(gamma + ch);
    ab[2 * ch] = a;
    ab[2 * ch + 1] = __ldg(beta + ch) - mean * a;
  }
  __syncthreads();
  const int PY = blockDim.x / CV;
  const int cv = threadIdx.x % CV, py = threadIdx.x / CV;
  const int c = cv * VEC;
  const T* src = c < C0 ? s0 + (size_t)img * HW * C0 + c : s1 + (size_t)img * HW * C1 + (c - C0);
  const int ld = c < C0 ? C0 : C1;
  T* out = dst + (size_t)img * HW * C + c;
  float a[VEC], b[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) { a[j] = ab[2 * (c + j)]; b[j] = ab[2 * (c + j) + 1]; }
  const int p0 = blockIdx.x * pix_per_cta, p1 = min(HW, p0 + pix_per_cta);
  constexpr int UN = 4;                              // independent 16-byte loads in flight per thread
  int p = p0 + py;
  for (; p + (UN - 1) * PY < p1; p += UN * PY) {
    float v[UN][VEC];
#pragma unroll
    for (int u = 0; u < UN; ++u) load_vec(src + (size_t)(p + u * PY) * ld, v[u]);
#pragma unroll
    for (int u = 0; u < UN; ++u) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        float y = v[u][j] * a[j] + b[j];
        v[u][j] = kSwish ? silu(y) : y;
      }
      store_vec(out + (size_t)(p + u * PY) * C, v[u]);
    }
  }
  for (; p < p1; p += PY) {
    float v[VEC];
    load_vec(src + (size_t)p * ld, v);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      float y = v[j] * a[j] + b[j];
      v[j] = kSwish ? silu(y) : y;
    }
    store_vec(out + (size_t)p * C, v);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) upsample2x_kernel(const T* __restrict__ src, int H, int W, int C, size_t total,
                                                         T* __restrict__ dst) {
  constexpr int VEC = VecOf<T>::N;
  const int CV = C / VEC;
  size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // over output (img, y, x, cv)
  if (gid >= total) return;
  const int cv = (int)(gid % CV);
  size_t r = gid / CV;
  const int x = (int)(r % (2 * W)); r /= (2 * W);
  const int y = (int)(r % (2 * H));
  const size_t img = r / (2 * H);
  const uint4 v = *reinterpret_cast<const uint4*>(src + ((img * H + (y >> 1)) * W + (x >> 1)) * C + cv * VEC);
  *reinterpret_cast<uint4*>(dst + gid * VEC) = v;
}

struct GnGeom { int threads, pix_per_cta, splits; };
static GnGeom gn_geom(int C, int vec, int HW, int images) {
  GnGeom g;
  const int CV = C / vec;
  const int PY = kGnThreads / CV > 0 ? kGnThreads / CV : 1;
  g.threads = CV * PY;
  // enough CTAs for ~4 waves of 148 SMs x 4 resident CTAs, but at least 4 pixels per pixel-lane
  int want = cdiv(148 * 16, images > 0 ? images : 1);
  int max_splits = HW / (PY * 4) > 0 ? HW / (PY * 4) : 1;
  g.splits = want < 1 ? 1 : (want > max_splits ? max_splits : want);
  g.pix_per_cta = cdiv(HW, g.splits);
  g.splits = cdiv(HW, g.pix_per_cta);
  return g;
}

}  // namespace vf

extern "C" __attribute__((visibility("default"))) int vf_gn_stats(const void* src0, int C0, const void* src1, int C1, int dtype, int images, int HW,
                           float* stats, vf_stream stream) {
  using namespace vf;
  VF_REQUIRE(src0 && stats && images > 0 && HW > 0 && C0 > 0, "vf_gn_stats: bad args");
  if (!src1) C1 = 0;
  const int vec = dtype == VF_BF16 ? 8 : 4;
  VF_REQUIRE(C0 % vec == 0 && C1 % vec == 0, "vf_gn_stats: channels (%d,%d) not a multiple of %d", C0, C1, vec);
  const int C = C0 + C1;
  VF_REQUIRE(C / vec <= kGnThreads, "vf_gn_stats: C=%d too large", C);
  GnGeom g = gn_geom(C, vec, HW, images);
  dim3 grid(g.splits, images);
  const size_t smem = 2 * C * sizeof(float);
  if (dtype == VF_BF16)
    gn_stats_kernel<__nv_bfloat16><<<grid, g.threads, smem, as_stream(stream)>>>((const __nv_bfloat16*)src0, C0, (const __nv_bfloat16*)src1, C1, HW, g.pix_per_cta, stats);
  else
    gn_stats_kernel<float><<<grid, g.threads, smem, as_stream(stream)>>>((const float*)src0, C0, (const float*)src1, C1, HW, g.pix_per_cta, stats);
  VF_LAUNCH_CHECK();
  return VF_OK;
}

extern "C" __attribute__((visibility("default"))) int vf_gn_apply(const void* src0, int C0, const float* stats0, int stats0_ld, const void* src1, int C1,
                           const float* stats1, int stats1_ld, int dtype, int images, int HW, int groups,
                           const float* gamma, const float* beta, int swish, void* dst, vf_stream stream) {
  using namespace vf;
  VF_REQUIRE(src0 && stats0 && gamma && beta && dst && images > 0 && HW > 0 && C0 > 0, "vf_gn_apply: bad args");
  if (!src1) C1 = 0;
  VF_REQUIRE(C1 == 0 || stats1, "vf_gn_apply: second source needs statistics");
  VF_REQUIRE(stats0_ld >= C0 && (C1 == 0 || stats1_ld >= C1), "vf_gn_apply: bad statistics stride");
  const int vec = dtype == VF_BF16 ? 8 : 4;
  VF_REQUIRE(C0 % vec == 0 && C1 % vec == 0, "vf_gn_apply: channels (%d,%d) not a multiple of %d", C0, C1, vec);
  const int C = C0 + C1;
  VF_REQUIRE(groups > 0 && C % groups == 0, "vf_gn_apply: C=%d not divisible by groups=%d", C, groups);
  VF_REQUIRE(C / vec <= kGnThreads, "vf_gn_apply: C=%d too large", C);
  GnGeom g = gn_geom(C, vec, HW, images);
  dim3 grid(g.splits, images);
  const size_t smem = 2 * C * sizeof(float);
  cudaStream_t st = as_stream(stream);
#define VF_GN_LAUNCH(T, SW) \
  gn_apply_kernel<T, SW><<<grid, g.threads, smem, st>>>((const T*)src0, C0, (const T*)src1, C1, HW, groups, g.pix_per_cta, stats0, stats0_ld, C1 ? stats1 : nullptr, stats1_ld, gamma, beta, (T*)dst)
  if (dtype == VF_BF16) { if (swish) VF_GN_LAUNCH(__nv_bfloat16, true); else VF_GN_LAUNCH(__nv_bfloat16, false); }
  else { if (swish) VF_GN_LAUNCH(float, true); else VF_GN_LAUNCH(float, false); }
#undef VF_GN_LAUNCH
  VF_LAUNCH_CHECK();
  return VF_OK;
}

extern "C" __attribute__((visibility("default"))) int vf_upsample2x(const void* src, int dtype, int images, int H, int W, int C, void* dst, vf_stream stream) {
  using namespace vf;
  VF_REQUIRE(src && dst && images > 0 && H > 0 && W > 0 && C > 0, "vf_upsample2x: bad args");
  const int vec = dtype == VF_BF16 ? 8 : 4;
  VF_REQUIRE(C % vec == 0, "vf_upsample2x: C=%d not a multiple of %d", C, vec);
  const size_t total = (size_t)images * 4 * H * W * (C / vec);
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (dtype == VF_BF16) upsample2x_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)src, H, W, C, total, (__nv_bfloat16*)dst);
  else upsample2x_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((const float*)src, H, W, C, total, (float*)dst);
  VF_LAUNCH_CHECK();
  return VF_OK;
}
