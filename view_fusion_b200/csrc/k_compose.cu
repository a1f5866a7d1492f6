// Fused view composition + DDPM update  (reference: model/view_fusion.py:116-160, :70-84, :166-177)
// and the training tail composition + MSE + output gradient (view_fusion.py:265-298).
//
// HBM-bound: per (sample, pixel) thread the kernel streams the V_b view outputs once (32 B per view-pixel,
// perfectly coalesced across the warp), keeps the softmax over views online in registers, and writes the
// 12 B of y_{t-1}.  Algorithmic bytes per sample-step: V*H*W*32 + H*W*3*4*2 (+ H*W*12 when z is injected).
#include "vf_common.cuh"

namespace vf {

// ---- Philox4x32-10 + Box-Muller (in-kernel noise draw for the perf path) ----------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
__device__ __forceinline__ void normal3(uint64_t seed, uint64_t offset, uint64_t idx, float (&z)[3]) {
  uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const float k = 2.3283064365386963e-10f;  // 2^-32
  float u0 = (c[0] + 0.5f) * k, u1 = (c[1] + 0.5f) * k, u2 = (c[2] + 0.5f) * k, u3 = (c[3] + 0.5f) * k;
  float r0 = sqrtf(-2.f * __logf(u0)), r1 = sqrtf(-2.f * __logf(u2));
  float s, co;
  __sincosf(6.283185307179586f * u1, &s, &co);
  z[0] = r0 * co; z[1] = r0 * s;
  z[2] = r1 * __cosf(6.283185307179586f * u3);
}

struct ComposeParams {
  vf_compose_args a;
  vf_schedule s;
};

// online softmax-weighted sum over the views of one (sample, pixel)
template <bool kWeighting>
__device__ __forceinline__ void compose_pixel(const float* __restrict__ out, int v0, int v1, int HW, int pix,
                                              float (&eps)[3], float (&mx)[3], float (&sum)[3]) {
  float acc[3] = {0.f, 0.f, 0.f};
  mx[0] = mx[1] = mx[2] = -INFINITY;
  sum[0] = sum[1] = sum[2] = 0.f;
  const float4* base = reinterpret_cast<const float4*>(out) + ((size_t)v0 * HW + pix) * 2;
  const size_t vstride = (size_t)HW * 2;
  for (int v = v0; v < v1; v += 4) {
    float4 lo[4], hi[4];
    const int n = min(4, v1 - v);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < n) {
        lo[j] = __ldg(base + (size_t)(v - v0 + j) * vstride);
        hi[j] = __ldg(base + (size_t)(v - v0 + j) * vstride + 1);
      }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < n) {
        const float e[3] = {lo[j].x, lo[j].y, lo[j].z};
        if (kWeighting) {
          const float l[3] = {lo[j].w, hi[j].x, hi[j].y};
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float m2 = fmaxf(mx[c], l[c]);
            float sc = expf(mx[c] - m2);      // exp(-inf) = 0 on the first view
            float p = expf(l[c] - m2);
            sum[c] = sum[c] * sc + p;
            acc[c] = acc[c] * sc + p * e[c];
            mx[c] = m2;
          }
        } else {
#pragma unroll
          for (int c = 0; c < 3; ++c) { acc[c] += e[c]; sum[c] += 1.f; }
        }
      }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) eps[c] = acc[c] / sum[c];
}

template <bool kWeighting>
__global__ void __launch_bounds__(256) compose_ddpm_kernel(const ComposeParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const vf_compose_args& a = p.a;
  const int HW = a.H * a.W;
  const int b = blockIdx.y;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= HW) return;
  const int v0 = __ldg(a.view_offset + b), v1 = __ldg(a.view_offset + b + 1);
  float eps[3], mx[3], sum[3];
  compose_pixel<kWeighting>(a.unet_out, v0, v1, HW, pix, eps, mx, sum);

  const int t = __ldg(a.t + b);
  const float A = __ldg(p.s.sqrt_recip_gammas + t), Bm = __ldg(p.s.sqrt_recipm1_gammas + t);
  const float c1 = __ldg(p.s.posterior_mean_coef1 + t), c2 = __ldg(p.s.posterior_mean_coef2 + t);
  // the `any(t > 0)` decision and the Philox offset come from the launch arguments or, in a replayed CUDA graph, from the
  // device-resident step record that vf_step_prepare wrote
  const bool add_noise = a.add_noise == 2 ? (a.step != nullptr && a.step->any_t_positive != 0) : a.add_noise != 0;
  const float sigma = add_noise ? expf(0.5f * __ldg(p.s.posterior_log_variance_clipped + t)) : 0.f;
  const size_t o = (size_t)b * 3 * HW + pix;
  float z[3] = {0.f, 0.f, 0.f};
  if (add_noise) {
    if (a.z) {
#pragma unroll
      for (int c = 0; c < 3; ++c) z[c] = __ldg(a.z + o + (size_t)c * HW);
    } else {
      normal3((a.step && a.seed == 0) ? a.step->seed : a.seed, a.step ? a.step->noise_offset : a.offset, (uint64_t)b * HW + pix, z);
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float yt = a.y_t[o + (size_t)c * HW];            // plain load: y_prev may alias y_t (in-place reverse loop)
    float y0 = A * yt - Bm * eps[c];
    if (a.clip_denoised) y0 = fminf(fmaxf(y0, -1.f), 1.f);
    const float mean = c1 * y0 + c2 * yt;
    a.y_prev[o + (size_t)c * HW] = mean + z[c] * sigma;
    if (a.eps_out) a.eps_out[o + (size_t)c * HW] = eps[c];
  }
  if (kWeighting && (a.weights_out || a.logits_out)) {   // snapshot steps only (8 of T): second pass from L1/L2
    const float4* base = reinterpret_cast<const float4*>(a.unet_out);
    for (int v = v0; v < v1; ++v) {
      float4 lo = __ldg(base + ((size_t)v * HW + pix) * 2), hi = __ldg(base + ((size_t)v * HW + pix) * 2 + 1);
      const float l[3] = {lo.w, hi.x, hi.y};
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (a.logits_out) a.logits_out[((size_t)v * 3 + c) * HW + pix] = l[c];
        if (a.weights_out)
          a.weights_out[(((size_t)b * a.max_v + (v - v0)) * 3 + c) * HW + pix] = expf(l[c] - mx[c]) / sum[c];
      }
    }
    if (a.weights_out)
      for (int v = v1 - v0; v < a.max_v; ++v)
#pragma unroll
        for (int c = 0; c < 3; ++c) a.weights_out[(((size_t)b * a.max_v + v) * 3 + c) * HW + pix] = 0.f;
  }
}

template <bool kWeighting>
__global__ void __launch_bounds__(256) compose_mse_kernel(const float* __restrict__ out, const int* __restrict__ voff,
                                                          const float* __restrict__ noise, int B, int HW,
                                                          float* __restrict__ loss_acc, float* __restrict__ eps_out,
                                                          float* __restrict__ grad, float gscale) {
  const int b = blockIdx.y;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  float part = 0.f;
  if (pix < HW) {
    const int v0 = __ldg(voff + b), v1 = __ldg(voff + b + 1);
    float eps[3], mx[3], sum[3], g[3];
    compose_pixel<kWeighting>(out, v0, v1, HW, pix, eps, mx, sum);
    const size_t o = (size_t)b * 3 * HW + pix;
    const float inv = 1.f / (3.f * (float)B * (float)HW);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float d = eps[c] - __ldg(noise + o + (size_t)c * HW);
      part += d * d;
      g[c] = 2.f * d * inv * gscale;
      if (eps_out) eps_out[o + (size_t)c * HW] = eps[c];
    }
    if (grad) {
      const float4* base = reinterpret_cast<const float4*>(out);
      float4* gb = reinterpret_cast<float4*>(grad);
      for (int v = v0; v < v1; ++v) {
        const size_t idx = ((size_t)v * HW + pix) * 2;
        float4 lo = __ldg(base + idx), hi = __ldg(base + idx + 1);
        const float e[3] = {lo.x, lo.y, lo.z}, l[3] = {lo.w, hi.x, hi.y};
        float de[3], dl[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          if (kWeighting) {
            const float w = expf(l[c] - mx[c]) / sum[c];
            de[c] = w * g[c];
            dl[c] = w * (e[c] - eps[c]) * g[c];
          } else {
            de[c] = g[c] / sum[c];
            dl[c] = 0.f;
          }
        }
        gb[idx] = make_float4(de[0], de[1], de[2], dl[0]);
        gb[idx + 1] = make_float4(dl[1], dl[2], 0.f, 0.f);
      }
    }
  }
  part = warp_sum(part);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(loss_acc, v / (3.f * (float)B * (float)HW));
  }
}

__global__ void q_sample_kernel(const float* __restrict__ y0, const float* __restrict__ noise,
                                const float* __restrict__ gammas, int chw, size_t total, float* __restrict__ y) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float g = __ldg(gammas + i / chw);
  y[i] = sqrtf(g) * y0[i] + sqrtf(1.f - g) * noise[i];
}

}  // namespace vf

extern "C" __attribute__((visibility("default"))) int vf_compose_ddpm_step(const vf_compose_args* a, const vf_schedule* s, vf_stream stream) {
  using namespace vf;
  VF_REQUIRE(a && s, "vf_compose_ddpm_step: null args");
  VF_REQUIRE(a->unet_out && a->view_offset && a->t && a->y_t && a->y_prev, "vf_compose_ddpm_step: null tensor");
  VF_REQUIRE(a->B > 0 && a->H > 0 && a->W > 0, "vf_compose_ddpm_step: bad shape B=%d H=%d W=%d", a->B, a->H, a->W);
  VF_REQUIRE(!a->weights_out || a->max_v > 0, "vf_compose_ddpm_step: weights_out needs max_v");
  VF_REQUIRE(a->weighting || (!a->weights_out && !a->logits_out), "vf_compose_ddpm_step: no weights without weighting");
  ComposeParams p{*a, *s};
  const int HW = a->H * a->W;
  dim3 grid(cdiv(HW, 256), a->B);
  if (a->weighting) VF_CUDA(launch_pdl(compose_ddpm_kernel<true>, grid, dim3(256), 0, as_stream(stream), p));
  else VF_CUDA(launch_pdl(compose_ddpm_kernel<false>, grid, dim3(256), 0, as_stream(stream), p));
  VF_LAUNCH_CHECK();
  return VF_OK;
}

extern "C" __attribute__((visibility("default"))) int vf_compose_mse(const float* unet_out, const int* view_offset, const float* noise, int B, int H, int W,
                              int weighting, float* loss_acc, float* eps_out, float* grad_out, float grad_scale,
                              vf_stream stream) {
  using namespace vf;
  VF_REQUIRE(unet_out && view_offset && noise && loss_acc, "vf_compose_mse: null tensor");
  VF_REQUIRE(B > 0 && H > 0 && W > 0, "vf_compose_mse: bad shape");
  const int HW = H * W;
  dim3 grid(cdiv(HW, 256), B);
  if (weighting)
    compose_mse_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(unet_out, view_offset, noise, B, HW, loss_acc, eps_out, grad_out, grad_scale);
  else
    compose_mse_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(unet_out, view_offset, noise, B, HW, loss_acc, eps_out, grad_out, grad_scale);
  VF_LAUNCH_CHECK();
  return VF_OK;
}

extern "C" __attribute__((visibility("default"))) int vf_q_sample(const float* y0, const float* noise, const float* gammas, int B, int chw, float* y,
                           vf_stream stream) {
  using namespace vf;
  VF_REQUIRE(y0 && noise && gammas && y && B > 0 && chw > 0, "vf_q_sample: bad args");
  size_t total = (size_t)B * chw;
  q_sample_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(y0, noise, gammas, chw, total, y);
  VF_LAUNCH_CHECK();
  return VF_OK;
}
