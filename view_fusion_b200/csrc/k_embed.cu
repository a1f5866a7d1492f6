// Noise-level / angle embedding path (reference: model/unet.py:115-116, :27-32, :142-157, :160-177).
//
// The reference evaluates PositionalEncoding + noise_level_mlp once per VIEW and then one Linear(ic -> Cout)
// inside each of the 30 ResnetBlocks.  Level and angle are per SAMPLE, so here one launch evaluates the MLP once
// per embedding row and produces all blocks' additive biases as a single [rows, E] fp32 table (E = sum of Cout,
// 5568 for the small config) that the convolution epilogues consume.
#include "vf_common.cuh"

namespace vf {

// grid (chunks, rows); every CTA recomputes the tiny MLP (ic*4*ic*2 MACs) and then its slice of the E outputs.
// All three matrix-vector products are warp-per-output: the lanes stride the inner dimension (coalesced weight rows)
// and reduce with shuffles.
// U outputs per call: the U weight rows are fetched with independent loads (the weights are cold in L2 every step, so a
// serial loop over outputs would pay one DRAM round trip per output)
template <int U>
__device__ __forceinline__ void warp_dots(const float* __restrict__ w, size_t row_stride, int rows_left, const float* x, int n, int lane,
                                          float (&acc)[U]) {
#pragma unroll
  for (int u = 0; u < U; ++u) acc[u] = 0.f;
  for (int i = lane; i < n; i += 32) {
    float wv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) wv[u] = u < rows_left ? __ldg(w + (size_t)u * row_stride + i) : 0.f;
    const float xv = x[i];
#pragma unroll
    for (int u = 0; u < U; ++u) acc[u] += wv[u] * xv;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int u = 0; u < U; ++u) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
}

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
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
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
  constexpr int U = 8;
  float acc[U];
  for (int o = warp * U; o < 4 * ic; o += nwarps * U) {
    warp_dots<U>(w0 + (size_t)o * ic, (size_t)ic, 4 * ic - o, pe, ic, lane, acc);
    if (lane < U && o + lane < 4 * ic) {
      float a = 0.f;
#pragma unroll
      for (int u = 0; u < U; ++u) a = lane == u ? acc[u] : a;
      a += __ldg(b0 + o + lane);
      hid[o + lane] = a / (1.f + expf(-a));      // Swish
    }
  }
  __syncthreads();
  for (int o = warp * U; o < ic; o += nwarps * U) {
    warp_dots<U>(w2 + (size_t)o * 4 * ic, (size_t)4 * ic, ic - o, hid, 4 * ic, lane, acc);
    if (lane < U && o + lane < ic) {
      float a = 0.f;
#pragma unroll
      for (int u = 0; u < U; ++u) a = lane == u ? acc[u] : a;
      tv[o + lane] = a + __ldg(b2 + o + lane);
    }
  }
  __syncthreads();
  const int per = (E + gridDim.x - 1) / gridDim.x;
  const int e0 = blockIdx.x * per, e1 = min(E, e0 + per);
  for (int e = e0 + warp * U; e < e1; e += nwarps * U) {
    warp_dots<U>(ew + (size_t)e * ic, (size_t)ic, e1 - e, tv, ic, lane, acc);
    if (lane < U && e + lane < e1) {
      float a = 0.f;
#pragma unroll
      for (int u = 0; u < U; ++u) a = lane == u ? acc[u] : a;
      out[(size_t)row * E + e + lane] = a + __ldg(eb + e + lane);
    }
  }
}

}  // namespace vf

extern "C" __attribute__((visibility("default"))) int vf_embed(const float* level, const float* angle, int rows, int inner_channel, const float* w0,
                        const float* b0, const float* w2, const float* b2, const float* emb_w, const float* emb_b, int E,
                        float* out, vf_stream stream) {
  using namespace vf;
  VF_REQUIRE(level && angle && w0 && b0 && w2 && b2 && emb_w && emb_b && out, "vf_embed: null tensor");
  VF_REQUIRE(rows > 0 && E > 0 && inner_channel >= 4 && inner_channel % 4 == 0, "vf_embed: bad shape");
  const int chunks = E >= 2048 ? 16 : 1;
  dim3 grid(chunks, rows);
  embed_kernel<<<grid, 256, 6 * inner_channel * sizeof(float), as_stream(stream)>>>(level, angle, inner_channel, w0, b0, w2, b2, emb_w, emb_b, E, out);
  VF_LAUNCH_CHECK();
  return VF_OK;
}
