// Evaluation metrics on the device: PSNR and SSIM of generated vs ground-truth views,
//   utils/metrics.py:6-8   psnr = 20 log10(1 / sqrt(mean_{c,h,w} (g - t)^2))                      -> (B,)
//   utils/metrics.py:11-12 ssim = pytorch_msssim.ssim(g, t, data_range=1.0, size_average=False)   -> (B,)
// (pytorch-msssim 1.0.0: 11-tap Gaussian, sigma 1.5, separable VALID depth-wise filter along H then W of x, y, x*x, y*y,
// x*y; C1 = 0.01^2, C2 = 0.03^2; mean of the SSIM map per channel, then over channels; oracle: vf_oracle.ssim).
// One CTA per (image, channel) plane: the two planes sit in shared memory, the H pass writes five filtered planes next to
// them, the W pass folds them into the SSIM map and its sum; both results leave as one atomic per plane.  Inputs are the
// reference's NCHW fp32 tensors, so `eval()` needs no host round trip once sampling runs on the device.
#include <mutex>

#include "vf_common.cuh"

namespace vf {

constexpr int MT_THREADS = 256;
constexpr int MT_WIN = 11;

struct MetricParams {
  const float* x;
  const float* y;
  int C, H, W;
  float g[MT_WIN];
  float c1, c2;
  float* mse_sum;   // [B] (+=)
  float* ssim;      // [B] (+=)
};

__device__ __forceinline__ float block_sum(float v, float* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float t = 0.f;
  if (warp == 0) {
    t = lane < (int)(blockDim.x >> 5) ? scratch[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;       // valid in thread 0
}

__global__ void __launch_bounds__(MT_THREADS) metrics_kernel(const MetricParams p) {
  extern __shared__ float sm[];
  __shared__ float red[MT_THREADS / 32];
  const int HW = p.H * p.W, Ho = p.H - MT_WIN + 1, Wo = p.W - MT_WIN + 1;
  float* sx = sm;
  float* sy = sm + HW;
  float* tmp = sm + 2 * HW;                  // [5][Ho][W]: x, y, x*x, y*y, x*y filtered along H
  const int plane = blockIdx.x, b = plane / p.C;
  const float* x = p.x + (size_t)plane * HW;
  const float* y = p.y + (size_t)plane * HW;
  float se = 0.f;
  for (int i = threadIdx.x; i < HW; i += MT_THREADS) {
    const float xv = __ldg(x + i), yv = __ldg(y + i), d = xv - yv;
    sx[i] = xv; sy[i] = yv;
    se = fmaf(d, d, se);
  }
  se = block_sum(se, red);                   // also orders the smem writes above before the reads below
  if (threadIdx.x == 0) atomicAdd(p.mse_sum + b, se);
  const int n1 = Ho * p.W;
  for (int idx = threadIdx.x; idx < n1; idx += MT_THREADS) {
    const int i = idx / p.W, j = idx - i * p.W;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
#pragma unroll
    for (int k = 0; k < MT_WIN; ++k) {
      const float xv = sx[(i + k) * p.W + j], yv = sy[(i + k) * p.W + j], gk = p.g[k];
      a0 = fmaf(gk, xv, a0); a1 = fmaf(gk, yv, a1);
      a2 = fmaf(gk, xv * xv, a2); a3 = fmaf(gk, yv * yv, a3); a4 = fmaf(gk, xv * yv, a4);
    }
    tmp[idx] = a0; tmp[n1 + idx] = a1; tmp[2 * n1 + idx] = a2; tmp[3 * n1 + idx] = a3; tmp[4 * n1 + idx] = a4;
  }
  __syncthreads();
  float acc = 0.f;
  const int n2 = Ho * Wo;
  for (int idx = threadIdx.x; idx < n2; idx += MT_THREADS) {
    const int i = idx / Wo, j = idx - i * Wo;
    const float* t = tmp + i * p.W + j;
    float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
    for (int k = 0; k < MT_WIN; ++k) {
      const float gk = p.g[k];
      m1 = fmaf(gk, t[k], m1); m2 = fmaf(gk, t[n1 + k], m2);
      e11 = fmaf(gk, t[2 * n1 + k], e11); e22 = fmaf(gk, t[3 * n1 + k], e22); e12 = fmaf(gk, t[4 * n1 + k], e12);
    }
    const float s11 = e11 - m1 * m1, s22 = e22 - m2 * m2, s12 = e12 - m1 * m2;
    const float cs = (2.f * s12 + p.c2) / (s11 + s22 + p.c2);
    acc += (2.f * m1 * m2 + p.c1) / (m1 * m1 + m2 * m2 + p.c1) * cs;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(p.ssim + b, acc / ((float)n2 * (float)p.C));
}

__global__ void metrics_finalize_kernel(float* psnr, int B, float inv_n) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float mse = psnr[b] * inv_n;
  psnr[b] = 20.f * log10f(1.f / sqrtf(mse));
}

}  // namespace vf

extern "C" __attribute__((visibility("default"))) int vf_eval_metrics(const float* generated, const float* target, int B, int C, int H, int W,
                                                                      float* psnr, float* ssim, vf_stream stream) {
  using namespace vf;
  VF_REQUIRE(generated && target && psnr && ssim && B > 0 && C > 0, "vf_eval_metrics: bad args");
  VF_REQUIRE(H >= MT_WIN && W >= MT_WIN, "vf_eval_metrics: %dx%d images are smaller than the %d-tap window", H, W, MT_WIN);
  const size_t smem = ((size_t)2 * H * W + (size_t)5 * (H - MT_WIN + 1) * W) * sizeof(float);
  constexpr size_t kMaxDyn = 226 * 1024;      // 227 KB per CTA minus the kernel's static reduction scratch
  VF_REQUIRE(smem <= kMaxDyn, "vf_eval_metrics: %dx%d planes need %zu B of shared memory", H, W, smem);
  VF_SET_MAX_SMEM(metrics_kernel, 226 * 1024);
  cudaStream_t st = as_stream(stream);
  MetricParams p{};
  p.x = generated; p.y = target; p.C = C; p.H = H; p.W = W; p.mse_sum = psnr; p.ssim = ssim;
  {
    // the window of pytorch_msssim._fspecial_gauss_1d(11, 1.5) in fp32
    float g[MT_WIN], s = 0.f;
    for (int k = 0; k < MT_WIN; ++k) { const float c = (float)(k - MT_WIN / 2); g[k] = expf(-(c * c) / (2.f * 1.5f * 1.5f)); s += g[k]; }
    for (int k = 0; k < MT_WIN; ++k) p.g[k] = g[k] / s;
  }
  p.c1 = 0.01f * 0.01f; p.c2 = 0.03f * 0.03f;
  VF_CUDA(cudaMemsetAsync(psnr, 0, (size_t)B * sizeof(float), st));
  VF_CUDA(cudaMemsetAsync(ssim, 0, (size_t)B * sizeof(float), st));
  metrics_kernel<<<B * C, MT_THREADS, smem, st>>>(p);
  metrics_finalize_kernel<<<(B + 127) / 128, 128, 0, st>>>(psnr, B, 1.f / ((float)C * H * W));
  VF_LAUNCH_CHECK();
  return VF_OK;
}
