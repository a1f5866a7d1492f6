// Input path on the device: what data/nmr_dataset.py:10-52 (`process_sample`) does per object on the
// host, for a whole batch in one launch.  The decoded views arrive as uint8 HWC images (webdataset's "rgb" decoder yields
// the same values as float32 / 255); given the per-object view permutation the reference draws with np.random.shuffle,
//   target[b]      = views[b, perm[b, 0]] / 255            (C, H, W) fp32               nmr_dataset.py:17-19
//   cond[b, j]     = views[b, perm[b, j + 1]] / 255        (V - 1, C, H, W) fp32        nmr_dataset.py:43 ("cond")
//   angle[b]       = 2 pi / V * perm[b, 0]                                              nmr_dataset.py:20-24
// i.e. a gather + HWC -> CHW transpose + uint8 -> float conversion: 1 byte read and 4 bytes written per element, no
// float image ever crosses PCIe.  One thread per (object, view slot, pixel): the C channel bytes of a pixel are
// adjacent in the source, the C stores go to C planes with consecutive threads on consecutive addresses.
#include "vf_common.cuh"

namespace vf {

__global__ void __launch_bounds__(256) prepare_batch_u8_kernel(const uint8_t* __restrict__ views, const int* __restrict__ perm, int V, int C, int HW,
                                                               size_t total, float* __restrict__ target, float* __restrict__ cond,
                                                               float* __restrict__ angle) {
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // over (b, slot, pixel)
  if (gid >= total) return;
  const int pix = (int)(gid % HW);
  const size_t bs = gid / HW;
  const int slot = (int)(bs % V);
  const size_t b = bs / V;
  const int v = __ldg(perm + b * V + slot);
  const uint8_t* src = views + ((b * V + v) * (size_t)HW + pix) * C;
  float* dst = slot == 0 ? target + b * (size_t)C * HW : cond + ((b * (V - 1) + (slot - 1)) * (size_t)C) * HW;
  for (int c = 0; c < C; ++c) dst[(size_t)c * HW + pix] = (float)src[c] / 255.0f;     // IEEE division: bit-exact with numpy
  if (slot == 0 && pix == 0) angle[b] = (float)(2.0 * 3.14159265358979323846 / (double)V * (double)v);
}

}  // namespace vf

extern "C" __attribute__((visibility("default"))) int vf_prepare_batch_u8(const uint8_t* views, const int* perm, int B, int V, int C, int H, int W,
                                                                          float* target, float* cond, float* angle, vf_stream stream) {
  using namespace vf;
  VF_REQUIRE(views && perm && target && cond && angle && B > 0 && V > 1 && C > 0 && H > 0 && W > 0, "vf_prepare_batch_u8: bad args");
  const size_t total = (size_t)B * V * H * W;
  const unsigned grid = (unsigned)((total + 255) / 256);
  prepare_batch_u8_kernel<<<grid, 256, 0, as_stream(stream)>>>(views, perm, V, C, H * W, total, target, cond, angle);
  VF_LAUNCH_CHECK();
  return VF_OK;
}
