// CUDA-core single-head self-attention core (reference: model/unet.py:267-274), fp32 accumulation.
// Used by the fp32 reference-precision mode and as the on-device cross-check of the tcgen05 kernel.
//   S = Q K^T / sqrt(C);  P = softmax_keys(S);  O = P V        per image, L = H*W tokens
#include "vf_common.cuh"

namespace vf {

constexpr int AQ = 8;   // queries per CTA

template <typename T>
__global__ void __launch_bounds__(256) attn_simt_kernel(const T* __restrict__ qk, int ld, const T* __restrict__ vt, int L,
                                                        int C, T* __restrict__ out) {
  extern __shared__ float sm[];
  float* q = sm;               // [AQ][C]
  float* s = sm + AQ * C;      // [AQ][L]
  const int img = blockIdx.y, q0 = blockIdx.x * AQ;
  const T* base = qk + (size_t)img * L * ld;
  for (int i = threadIdx.x; i < AQ * C; i += blockDim.x) {
    const int qi = i / C, c = i % C;
    q[i] = (q0 + qi < L) ? to_f(base[(size_t)(q0 + qi) * ld + c]) : 0.f;
  }
  __syncthreads();
  const float scale = rsqrtf((float)C);
  for (int i = threadIdx.x; i < AQ * L; i += blockDim.x) {
    const int qi = i / L, kj = i % L;
    const T* kr = base + (size_t)kj * ld + C;
    const float* qr = q + qi * C;
    float acc = 0.f;
    for (int c = 0; c < C; ++c) acc += qr[c] * to_f(kr[c]);
    s[i] = acc * scale;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int qi = warp; qi < AQ; qi += blockDim.x >> 5) {
    float* row = s + qi * L;
    float m = -INFINITY;
    for (int j = lane; j < L; j += 32) m = fmaxf(m, row[j]);
    m = warp_max(m);
    float sum = 0.f;
    for (int j = lane; j < L; j += 32) { float e = expf(row[j] - m); row[j] = e; sum += e; }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int j = lane; j < L; j += 32) row[j] *= inv;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < AQ * C; i += blockDim.x) {
    const int qi = i / C, c = i % C;
    if (q0 + qi >= L) continue;
    const float* row = s + qi * L;
    float acc = 0.f;
    if (vt) {
      const T* vr = vt + ((size_t)img * C + c) * L;
      for (int j = 0; j < L; ++j) acc += row[j] * to_f(vr[j]);
    } else {
      const T* vc = base + 2 * C + c;
      for (int j = 0; j < L; ++j) acc += row[j] * to_f(vc[(size_t)j * ld]);
    }
    out[((size_t)img * L + q0 + qi) * C + c] = from_f<T>(acc);
  }
}

int attention_simt(const void* qk, const void* vt, int dtype, int images, int L, int C, void* out, cudaStream_t st) {
  dim3 grid(cdiv(L, AQ), images);
  const size_t smem = (size_t)AQ * (C + L) * sizeof(float);
  VF_REQUIRE(smem <= 48 * 1024, "vf_attention(simt): L=%d C=%d too large", L, C);
  if (dtype == VF_BF16)
    attn_simt_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>((const __nv_bfloat16*)qk, 3 * C, (const __nv_bfloat16*)vt, L, C, (__nv_bfloat16*)out);
  else
    attn_simt_kernel<float><<<grid, 256, smem, st>>>((const float*)qk, 3 * C, (const float*)vt, L, C, (float*)out);
  VF_LAUNCH_CHECK();
  return VF_OK;
}

}  // namespace vf
