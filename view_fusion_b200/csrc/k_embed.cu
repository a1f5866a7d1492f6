// Noise-level / angle embedding path (reference: model/unet.py:115-116, :27-32, :142-157, :160-177).
//
// The reference evaluates PositionalEncoding + noise_level_mlp once per VIEW and then one Linear(ic -> Cout)
// inside each of the 30 ResnetBlocks.  Level and angle are per SAMPLE, so here one launch evaluates the MLP once
// per embedding row and produces all blocks' additive biases as a single [rows, E] fp32 table (E = sum of Cout,
// 5568 for the small config) that the convolution epilogues consume.
#include "vf_common.cuh"

namespace vf {

// grid (chunks, rows); every CTA recomputes the tiny MLP (ic*4*ic*2 MACs) and then its slice of the E outputs.
__global__ void __launch_bounds__(256) embed_kernel(const float* __restrict__ level, const float* __restrict__ angle,
                                                    int ic, const float* __restrict__ w0, const float* __restrict__ b0,
                                                    const float* __restrict__ w2, const float* __restrict__ b2,
                                                    const float* __restrict__ ew, const float* __restrict__ eb, int E,
                                                    float* __restrict__ out) {
  extern __shared__ float sm[];
  float* pe = sm;            // [ic]
  float* hid = sm + ic;      // [4*ic]
  float* tv = hid + 4 * ic;  // [ic]
  const int row = blockIdx.y;
  const int half = ic / 2, cnt = ic / 4;   // PositionalEncoding(dim = ic/2): count = dim/2 frequencies
  for (int i = threadIdx.x; i < ic; i += blockDim.x) {
    // t_angle = cat(PE(time), PE(angle)); PE(x) = [sin(x f_0..f_{cnt-1}), cos(x f_0..)], f_k = exp(-ln(1e4) k / cnt)
    const float x = i < half ? __ldg(level + row) : __ldg(angle + row);
    const int j = i % half;
    const int k = j % cnt;
    const float f = expf(-9.210340371976184f * ((float)k / (float)cnt));
    pe[i] = j < cnt ? sinf(x * f) : cosf(x * f);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < 4 * ic; o += blockDim.x) {
    float acc = __ldg(b0 + o);
    for (int i = 0; i < ic; ++i) acc += __ldg(w0 + (size_t)o * ic + i) * pe[i];
    hid[o] = acc / (1.f + expf(-acc));      // Swish
  }
  __syncthreads();
  for (int o = threadIdx.x; o < ic; o += blockDim.x) {
    float acc = __ldg(b2 + o);
    for (int i = 0; i < 4 * ic; ++i) acc += __ldg(w2 + (size_t)o * 4 * ic + i) * hid[i];
    tv[o] = acc;
  }
  __syncthreads();
  const int per = (E + gridDim.x - 1) / gridDim.x;
  const int e0 = blockIdx.x * per, e1 = min(E, e0 + per);
  for (int e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
    float acc = __ldg(eb + e);
    const float* w = ew + (size_t)e * ic;
    for (int i = 0; i < ic; ++i) acc += __ldg(w + i) * tv[i];
    out[(size_t)row * E + e] = acc;
  }
}

}  // namespace vf

extern "C" __attribute__((visibility("default"))) int vf_embed(const float* level, const float* angle, int rows, int inner_channel, const float* w0,
                        const float* b0, const float* w2, const float* b2, const float* emb_w, const float* emb_b, int E,
                        float* out, vf_stream stream) {
  using namespace vf;
  VF_REQUIRE(level && angle && w0 && b0 && w2 && b2 && emb_w && emb_b && out, "vf_embed: null tensor");
  VF_REQUIRE(rows > 0 && E > 0 && inner_channel >= 4 && inner_channel % 4 == 0, "vf_embed: bad shape");
  const int chunks = E >= 2048 ? 8 : 1;
  dim3 grid(chunks, rows);
  embed_kernel<<<grid, 256, 6 * inner_channel * sizeof(float), as_stream(stream)>>>(level, angle, inner_channel, w0, b0, w2, b2, emb_w, emb_b, E, out);
  VF_LAUNCH_CHECK();
  return VF_OK;
}
