// Fused multi-tensor Adam (reference: torch.optim.Adam as constructed in experiment.py:115-120, stepped at :293).
//
// One launch updates every parameter tensor of the model: a device table holds (param, grad, exp_avg, exp_avg_sq,
// numel) per tensor and each CTA takes one fixed-size chunk of one tensor.  HBM-bound: 16 B read + 12 B written per
// element (fp32 master weights, gradient, two moments), 128-bit accesses when the four pointers are 16-byte aligned.
// Arithmetic is torch's (no amsgrad, no maximize):
//   g += wd * p;  m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g*g;  p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
#include "vf_common.cuh"

namespace vf {

constexpr int kAdamChunk = 4096;      // elements per CTA
constexpr int kAdamThreads = 256;

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float lr_c, float b1, float omb1, float b2, float omb2,
                                         float eps, float wd, float inv_sqrt_c2) {
  g = fmaf(wd, p, g);
  m = fmaf(b1, m, omb1 * g);
  v = fmaf(b2, v, omb2 * g * g);
  const float denom = fmaf(sqrtf(v), inv_sqrt_c2, eps);
  p -= lr_c * (m / denom);
}

__global__ void __launch_bounds__(kAdamThreads) adam_kernel(const vf_adam_entry* __restrict__ table, const int* __restrict__ chunk_entry,
                                                            const int* __restrict__ chunk_start, float lr_c, float b1, float omb1, float b2,
                                                            float omb2, float eps, float wd, float inv_sqrt_c2) {
  const vf_adam_entry e = table[chunk_entry[blockIdx.x]];
  const int start = chunk_start[blockIdx.x];
  const int end = min(e.numel, start + kAdamChunk);
  float* p = e.param + start;
  const float* g = e.grad + start;
  float* m = e.exp_avg + start;
  float* v = e.exp_avg_sq + start;
  const int n = end - start;
  const bool aligned = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                         reinterpret_cast<uintptr_t>(v)) & 15u) == 0;
  if (aligned) {
    const int n4 = n >> 2;
    for (int i = threadIdx.x; i < n4; i += kAdamThreads) {
      float4 pv = reinterpret_cast<float4*>(p)[i], mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
      const float4 gv = reinterpret_cast<const float4*>(g)[i];
      adam_one(pv.x, gv.x, mv.x, vv.x, lr_c, b1, omb1, b2, omb2, eps, wd, inv_sqrt_c2);
      adam_one(pv.y, gv.y, mv.y, vv.y, lr_c, b1, omb1, b2, omb2, eps, wd, inv_sqrt_c2);
      adam_one(pv.z, gv.z, mv.z, vv.z, lr_c, b1, omb1, b2, omb2, eps, wd, inv_sqrt_c2);
      adam_one(pv.w, gv.w, mv.w, vv.w, lr_c, b1, omb1, b2, omb2, eps, wd, inv_sqrt_c2);
      reinterpret_cast<float4*>(p)[i] = pv; reinterpret_cast<float4*>(m)[i] = mv; reinterpret_cast<float4*>(v)[i] = vv;
    }
    for (int i = (n4 << 2) + threadIdx.x; i < n; i += kAdamThreads) adam_one(p[i], g[i], m[i], v[i], lr_c, b1, omb1, b2, omb2, eps, wd, inv_sqrt_c2);
  } else {
    for (int i = threadIdx.x; i < n; i += kAdamThreads) adam_one(p[i], g[i], m[i], v[i], lr_c, b1, omb1, b2, omb2, eps, wd, inv_sqrt_c2);
  }
}

}  // namespace vf

extern "C" __attribute__((visibility("default"))) int vf_adam_chunk_elems(void) { return vf::kAdamChunk; }

extern "C" __attribute__((visibility("default"))) int vf_adam_step(const vf_adam_entry* table_dev, const int* chunk_entry_dev,
                                                                    const int* chunk_start_dev, int n_chunks, double lr, double beta1,
                                                                    double beta2, double eps, double weight_decay, int step,
                                                                    vf_stream stream) {
  using namespace vf;
  VF_REQUIRE(table_dev && chunk_entry_dev && chunk_start_dev && n_chunks > 0 && step >= 1, "vf_adam_step: bad args");
  // hyper-parameters arrive as doubles (Python floats): 1 - beta is formed in double like torch does
  const double c1 = 1.0 - pow(beta1, (double)step), c2 = 1.0 - pow(beta2, (double)step);
  const float lr_c = (float)(lr / c1), inv_sqrt_c2 = (float)(1.0 / sqrt(c2));
  adam_kernel<<<n_chunks, kAdamThreads, 0, as_stream(stream)>>>(table_dev, chunk_entry_dev, chunk_start_dev, lr_c, (float)beta1,
                                                                (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps,
                                                                (float)weight_decay, inv_sqrt_c2);
  VF_LAUNCH_CHECK();
  return VF_OK;
}
