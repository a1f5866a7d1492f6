// Layout kernels at the edges of the hot path.
//  - vf_pack_views: view stacking of model/view_fusion.py:95-115 / :244-263 (first V_b views, target repeated per
//    view, channel concat) fused with the im2col of the first 3x3 convolution (unet.py:42), so that layer runs as
//    a plain K=K0 GEMM.  Bandwidth-trivial: reads 12-24 B and writes K0*sizeof(T) per view-pixel.
//  - vf_pack_nchw / vf_nhwc_to_nchw: the NCHW fp32 <-> NHWC conversions of the stand-alone UNet.forward API.
//  - vf_pack_conv_weight: fp32 OIHW masters -> K-major GEMM rows.
#include "vf_common.cuh"

namespace vf {

__global__ void img_sample_kernel(const int* __restrict__ voff, int B, int images, int* __restrict__ img_sample) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= images) return;
  int b = 0;
  while (b + 1 < B && __ldg(voff + b + 1) <= i) ++b;
  img_sample[i] = b;
}

// One thread per (image, pixel), consecutive threads = consecutive x: every (tap, channel) load is coalesced along the
// image row (the 9x re-reads of a pixel hit L1), the K0 values of the pixel are staged in a conflict-free smem column
// and leave as one contiguous K0*sizeof(T) row.
// CC > 0: conditioning channels known at compile time -> the 9 * (CC + 3) loads are fully unrolled and independent.
constexpr int kPackThreads = 128;
template <typename T, int CC>
__global__ void __launch_bounds__(kPackThreads) pack_views_kernel(const float* __restrict__ y_cond, const float* __restrict__ y_t,
                                                                  const int* __restrict__ voff, const int* __restrict__ img_sample,
                                                                  int n_max, int Cc, int H, int W, int images, int K0,
                                                                  T* __restrict__ x0) {
  constexpr int VEC = VecOf<T>::N;
  extern __shared__ float stage[];                 // [K0][kPackThreads + 1]
  const size_t total = (size_t)images * H * W;
  const size_t gid = (size_t)blockIdx.x * kPackThreads + threadIdx.x;
  const bool live = gid < total;
  const size_t ip = live ? gid : total - 1;
  const int pix = (int)(ip % (size_t)(H * W));
  const int img = (int)(ip / (size_t)(H * W));
  const int y = pix / W, x = pix - y * W;
  const int b = __ldg(img_sample + img);
  const int view = img - __ldg(voff + b);
  if (CC > 0) Cc = CC;
  const int Cin = Cc + 3;
  const float* cond = y_cond + ((size_t)b * n_max + view) * Cc * H * W;
  const float* tgt = y_t + (size_t)b * 3 * H * W;
  // staging row stride of kPackThreads + 1 floats: the transposed read-out below is bank-conflict free
  constexpr int LD = kPackThreads + 1;
  float* col = stage + threadIdx.x;
  int k = 0;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
    const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
    const size_t off = (size_t)yy * W + xx;
    if (CC > 0) {
      float vals[CC + 3];
#pragma unroll
      for (int c = 0; c < CC + 3; ++c)
        vals[c] = in ? (c < CC ? __ldg(cond + (size_t)c * H * W + off) : __ldg(tgt + (size_t)(c - CC) * H * W + off)) : 0.f;
#pragma unroll
      for (int c = 0; c < CC + 3; ++c, ++k) col[k * LD] = vals[c];
    } else {
      for (int c = 0; c < Cin; ++c, ++k) {
        float val = 0.f;
        if (in) val = c < Cc ? __ldg(cond + (size_t)c * H * W + off) : __ldg(tgt + (size_t)(c - Cc) * H * W + off);
        col[k * LD] = val;
      }
    }
  }
  for (; k < K0; ++k) col[k * LD] = 0.f;
  __syncthreads();
  // write-out: consecutive threads take consecutive 16-byte chunks of consecutive pixels -> full 128-byte lines
  const int chunks = K0 / VEC;
  const size_t first = (size_t)blockIdx.x * kPackThreads;
  for (int i = threadIdx.x; i < kPackThreads * chunks; i += kPackThreads) {
    const int pl = i / chunks, ch = i - pl * chunks;
    if (first + pl >= total) break;
    float v[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) v[j] = stage[(ch * VEC + j) * LD + pl];
    store_vec(x0 + (first + pl) * (size_t)K0 + ch * VEC, v);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) pack_nchw_kernel(const float* __restrict__ x, int C, int H, int W, int R, int K0,
                                                        T* __restrict__ x0) {
  constexpr int VEC = VecOf<T>::N;
  const int chunks = K0 / VEC;
  const size_t total = (size_t)R * H * W * chunks;
  size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int ch = (int)(gid % chunks);
  const size_t ip = gid / chunks;
  const int pix = (int)(ip % (H * W));
  const int img = (int)(ip / (H * W));
  const int y = pix / W, xq = pix % W;
  const float* src = x + (size_t)img * C * H * W;
  float v[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const int k = ch * VEC + j;
    float val = 0.f;
    if (k < 9 * C) {
      const int tap = k / C, c = k % C;
      const int yy = y + tap / 3 - 1, xx = xq + tap % 3 - 1;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) val = __ldg(src + ((size_t)c * H + yy) * W + xx);
    }
    v[j] = val;
  }
  store_vec(x0 + gid * VEC, v);
}

__global__ void nhwc_to_nchw_kernel(const float* __restrict__ src, int ld, int C, int HW, size_t total,
                                    float* __restrict__ dst) {
  size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // over (r, c, pix): coalesced writes
  if (gid >= total) return;
  const int pix = (int)(gid % HW);
  const size_t rc = gid / HW;
  const int c = (int)(rc % C);
  const size_t r = rc / C;
  dst[gid] = __ldg(src + (r * HW + pix) * ld + c);
}

template <typename T>
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, int cout, int cin, int kk, T* __restrict__ dst,
                                        int cout_pad, int k_total, int k_off) {
  const int per_row = kk * cin;
  const size_t total = (size_t)cout_pad * per_row;
  size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int n = (int)(gid / per_row);
  const int r = (int)(gid % per_row);
  const int tap = r / cin, c = r % cin;
  float v = n < cout ? __ldg(w + ((size_t)n * cin + c) * kk + tap) : 0.f;   // OIHW: ((n*cin + c)*k + kh)*k + kw
  dst[(size_t)n * k_total + k_off + r] = from_f<T>(v);
}

// identity block: dst[n][k_off + c] = (c == n), used to ride an identity residual as a 1x1 K-segment of a GEMM
template <typename T>
__global__ void pack_identity_kernel(T* __restrict__ dst, int n_rows, int k_total, int k_off) {
  const size_t total = (size_t)n_rows * n_rows;
  size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int n = (int)(gid / n_rows), c = (int)(gid % n_rows);
  dst[(size_t)n * k_total + k_off + c] = from_f<T>(c == n ? 1.f : 0.f);
}

int pack_identity(void* dst, int dtype, int n_rows, int k_total, int k_off, cudaStream_t st) {
  const size_t total = (size_t)n_rows * n_rows;
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (dtype == VF_BF16) pack_identity_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((__nv_bfloat16*)dst, n_rows, k_total, k_off);
  else pack_identity_kernel<float><<<grid, 256, 0, st>>>((float*)dst, n_rows, k_total, k_off);
  VF_LAUNCH_CHECK();
  return VF_OK;
}

}  // namespace vf

extern "C" __attribute__((visibility("default"))) int vf_pack_views(const float* y_cond, const float* y_t, const int* view_offset, int B, int n_max,
                             int cond_channels, int H, int W, int images, int k0, int x0_dtype, void* x0,
                             int* img_sample, vf_stream stream) {
  using namespace vf;
  VF_REQUIRE(y_cond && y_t && view_offset && x0 && img_sample, "vf_pack_views: null tensor");
  VF_REQUIRE(B > 0 && images > 0 && H > 0 && W > 0 && n_max > 0, "vf_pack_views: bad shape");
  VF_REQUIRE(k0 % 8 == 0 && k0 >= 9 * (cond_channels + 3), "vf_pack_views: k0=%d too small for %d channels", k0,
             cond_channels + 3);
  cudaStream_t st = as_stream(stream);
  img_sample_kernel<<<cdiv(images, 128), 128, 0, st>>>(view_offset, B, images, img_sample);
  VF_LAUNCH_CHECK();
  const size_t total = (size_t)images * H * W;
  const unsigned grid = (unsigned)((total + kPackThreads - 1) / kPackThreads);
  const size_t smem = (size_t)k0 * (kPackThreads + 1) * sizeof(float);
  VF_REQUIRE(smem <= 160 * 1024, "vf_pack_views: k0=%d too large", k0);
  // K0 = 64 (in_channel 6) stages 33 KB; the relative variant (in_channel 9, K0 = 128) needs 66 KB: opt in above 48 KB
#define VF_PACK_LAUNCH(T, CC)                                                          \
  do {                                                                                 \
    if (smem > 48 * 1024) VF_SET_MAX_SMEM((pack_views_kernel<T, CC>), 160 * 1024);     \
    pack_views_kernel<T, CC><<<grid, kPackThreads, smem, st>>>(y_cond, y_t, view_offset, img_sample, n_max, cond_channels, H, W, images, k0, (T*)x0); \
  } while (0)
  if (x0_dtype == VF_BF16) {
    if (cond_channels == 3) VF_PACK_LAUNCH(__nv_bfloat16, 3); else if (cond_channels == 6) VF_PACK_LAUNCH(__nv_bfloat16, 6); else VF_PACK_LAUNCH(__nv_bfloat16, 0);
  } else {
    if (cond_channels == 3) VF_PACK_LAUNCH(float, 3); else if (cond_channels == 6) VF_PACK_LAUNCH(float, 6); else VF_PACK_LAUNCH(float, 0);
  }
#undef VF_PACK_LAUNCH
  VF_LAUNCH_CHECK();
  return VF_OK;
}

extern "C" __attribute__((visibility("default"))) int vf_pack_nchw(const float* x, int R, int C, int H, int W, int k0, int x0_dtype, void* x0, vf_stream stream) {
  using namespace vf;
  VF_REQUIRE(x && x0 && R > 0 && C > 0 && H > 0 && W > 0, "vf_pack_nchw: bad args");
  VF_REQUIRE(k0 % 8 == 0 && k0 >= 9 * C, "vf_pack_nchw: k0=%d too small for C=%d", k0, C);
  const size_t vec = x0_dtype == VF_BF16 ? 8 : 4;
  const size_t total = (size_t)R * H * W * (k0 / vec);
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (x0_dtype == VF_BF16) pack_nchw_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(x, C, H, W, R, k0, (__nv_bfloat16*)x0);
  else pack_nchw_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(x, C, H, W, R, k0, (float*)x0);
  VF_LAUNCH_CHECK();
  return VF_OK;
}

extern "C" __attribute__((visibility("default"))) int vf_nhwc_to_nchw(const float* src, int ld, int R, int C, int H, int W, float* dst, vf_stream stream) {
  using namespace vf;
  VF_REQUIRE(src && dst && R > 0 && C > 0 && C <= ld && H > 0 && W > 0, "vf_nhwc_to_nchw: bad args");
  const size_t total = (size_t)R * C * H * W;
  nhwc_to_nchw_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(src, ld, C, H * W, total, dst);
  VF_LAUNCH_CHECK();
  return VF_OK;
}

extern "C" __attribute__((visibility("default"))) int vf_pack_conv_weight(const float* w_oihw, int cout, int cin, int ksize, int dtype, void* dst, int cout_pad,
                                   int k_total, int k_off, vf_stream stream) {
  using namespace vf;
  VF_REQUIRE(w_oihw && dst && cout > 0 && cin > 0 && (ksize == 1 || ksize == 3), "vf_pack_conv_weight: bad args");
  VF_REQUIRE(cout_pad >= cout && k_off + ksize * ksize * cin <= k_total, "vf_pack_conv_weight: row overflow");
  const size_t total = (size_t)cout_pad * ksize * ksize * cin;
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (dtype == VF_BF16)
    pack_conv_weight_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(w_oihw, cout, cin, ksize * ksize, (__nv_bfloat16*)dst, cout_pad, k_total, k_off);
  else
    pack_conv_weight_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(w_oihw, cout, cin, ksize * ksize, (float*)dst, cout_pad, k_total, k_off);
  VF_LAUNCH_CHECK();
  return VF_OK;
}
