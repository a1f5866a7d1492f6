// GroupNorm(32, C, eps=1e-5, affine) + Swish  (reference: model/unet.py:207-218 Block, :254 attention norm).
//
// HBM-bound passes over PADDED NHWC activations with 128-bit accesses:
//   vf_gn_apply : turns per-(image, channel) sums (emitted by the producing convolution's epilogue) into per-group
//                 mean / rstd, folds them with gamma/beta into one per-channel FMA held in shared memory, applies
//                 x*a+b (+ Swish) and writes the PADDED tensor the next convolution's TMA reads (exact zeros in the
//                 padding rows = the convolution's zero padding).  It reads up to two sources so that
//                 torch.cat((x, skip), 1) (unet.py:134) is never materialised un-normalised.
//   vf_gn_stats : stand-alone statistics pass (API completeness / tests; the plan uses the fused sums).
// Algorithmic bytes (bf16): apply 4 B/elem, stats 2 B/elem.
#include <cstdlib>

#include "vf_common.cuh"

namespace vf {

constexpr int kGnThreads = 256;

// block = (CV channel-vectors) x (PY row lanes); every thread keeps sums for its fixed 16-byte channel vector
template <typename T>
__global__ void __launch_bounds__(kGnThreads) gn_stats_kernel(const T* __restrict__ s0, int C0, const T* __restrict__ s1,
                                                              int C1, int W1, int P, int rows_per_cta, float* __restrict__ stats) {
  constexpr int VEC = VecOf<T>::N;
  extern __shared__ float acc[];                     // [C][2]
  const int C = C0 + C1, CV = C / VEC;
  const int img = blockIdx.y;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  const int PY = blockDim.x / CV;                    // row lanes (blockDim = CV*PY exactly)
  const int cv = threadIdx.x % CV, py = threadIdx.x / CV;
  const int c = cv * VEC;
  const T* src = c < C0 ? s0 + (size_t)img * P * C0 + c : s1 + (size_t)img * P * C1 + (c - C0);
  const int ld = c < C0 ? C0 : C1;
  const int p0 = blockIdx.x * rows_per_cta, p1 = min(P, p0 + rows_per_cta);
  float s[VEC], q[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) s[j] = q[j] = 0.f;
  for (int p = p0 + py; p < p1; p += PY) {
    const int yy = p / W1, xx = p - yy * W1;
    if (yy == 0 || xx == 0) continue;                // padding row
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
__global__ void __launch_bounds__(kGnThreads, 4) gn_apply_kernel(const T* __restrict__ s0, int C0, const T* __restrict__ s1,
                                                              int C1, int HW, int W1, int P, int groups, int rows_per_cta,
                                                              const float* __restrict__ st0, int ld0,
                                                              const float* __restrict__ st1, int ld1,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, T* __restrict__ dst, const vf_gn_shift sh) {
  constexpr int VEC = VecOf<T>::N;
  constexpr bool kHalved = kSwish && sizeof(T) == 2;
  pdl_launch_dependents();
  pdl_wait();                                        // the statistics and the sources come from the previous kernels
  extern __shared__ float ab[];                      // [C][2]: y = x*a + b, then [C][2] staging of the raw statistics
  const int C = C0 + C1, CV = C / VEC;
  float* raw = ab + 2 * C;
  const int img = blockIdx.y;
  const int gs = C / groups;
  const float inv_n = 1.f / ((float)gs * (float)HW);
  const float* sa = st0 + (size_t)img * ld0 * 2;
  const float* sb = st1 ? st1 + (size_t)img * ld1 * 2 : nullptr;
  const int PY = blockDim.x / CV;
  const int cv = threadIdx.x % CV, py = threadIdx.x / CV;
  const int c = cv * VEC;
  const T* src = c < C0 ? s0 + (size_t)img * P * C0 + c : s1 + (size_t)img * P * C1 + (c - C0);
  const int ld = c < C0 ? C0 : C1;
  T* out = dst + (size_t)img * P * C + c;
  float a[VEC], b[VEC];        // y = x*a + b, filled after the statistics prologue below
  const int p0 = blockIdx.x * rows_per_cta, p1 = min(P, p0 + rows_per_cta);
  // Software pipeline: two half-batches of UN rows; while one is normalised and stored, the loads of the next are in
  // flight (a warp never sits in a pure wait-then-compute cycle).  (yy, xx) of the thread's row advance incrementally.
  constexpr int UN = 2;
  int yy = (p0 + py) / W1, xx = (p0 + py) - yy * W1;
  const int dy = PY / W1, dx = PY - dy * W1;
  // Rows are visited in order by both lambdas (every call takes the next UN rows of the thread), so the source / destination
  // addresses are running pointers advanced by a constant stride instead of 64-bit products per row.
  const T* fp = src + (size_t)(p0 + py) * ld;
  T* ep = out + (size_t)(p0 + py) * C;
  const size_t fstep = (size_t)PY * ld, estep = (size_t)PY * C;
  auto fetch = [&](int pb, uint4 (&raw)[UN], uint32_t& inm, uint32_t& padm) {
    inm = 0; padm = 0;
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int p = pb + u * PY;
      const bool in = p < p1, pad = yy == 0 || xx == 0;
      inm |= (uint32_t)in << u; padm |= (uint32_t)pad << u;
      if (in && !pad) raw[u] = *reinterpret_cast<const uint4*>(fp);
      fp += fstep;
      yy += dy; xx += dx;
      if (xx >= W1) { xx -= W1; ++yy; }
    }
  };
  auto emit = [&](int pb, const uint4 (&raw)[UN], uint32_t inm, uint32_t padm) {
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      T* dstp = ep;
      ep += estep;
      if (!((inm >> u) & 1u)) continue;
      float v[VEC];
      if ((padm >> u) & 1u) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) v[j] = 0.f;
      } else {
        load_vec(reinterpret_cast<const T*>(&raw[u]), v);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const float y = fmaf(v[j], a[j], b[j]);
          v[j] = kSwish ? (kHalved ? silu_half(y) : silu_for<T>(y)) : y;
        }
      }
      store_vec(dstp, v);
    }
  };
  uint4 ra[UN], rb2[UN];
  uint32_t ia, pa, ib, pb2;
  int pb = p0 + py;
  fetch(pb, ra, ia, pa);
  // The first rows are requested BEFORE the statistics are turned into coefficients, so the prologue (a global round trip,
  // two barriers, the per-group loops) overlaps that load instead of preceding it (measured neutral on B200 at 168
  // view-images: 1.64 ms per step either way; kept because it removes a dependency, not for a number).
  // one coalesced load of the image's (sum, sumsq) pairs; the per-group loops then run out of shared memory
  const bool shifted = sh.bias != nullptr || sh.emb != nullptr;
  const float* sh_emb = sh.emb ? sh.emb + (size_t)__ldg(sh.img_row + img) * sh.emb_ld : nullptr;
  if (!shifted) {
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) raw[i] = i < 2 * C0 ? __ldg(sa + i) : __ldg(sb + (i - 2 * C0));
  } else {
    // source 0 is stored without its per-(image, channel) constant s: shift its raw sums in closed form
    for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
      float S1, S2;
      if (ch < C0) {
        S1 = __ldg(sa + 2 * ch); S2 = __ldg(sa + 2 * ch + 1);
        const float sv = gn_shift_value(sh, sh_emb, ch);
        S2 = fmaf(2.f * sv, S1, S2) + (float)HW * sv * sv;
        S1 = fmaf((float)HW, sv, S1);
      } else {
        S1 = __ldg(sb + 2 * (ch - C0)); S2 = __ldg(sb + 2 * (ch - C0) + 1);
      }
      raw[2 * ch] = S1; raw[2 * ch + 1] = S2;
    }
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
    const int g0 = ch / gs * gs;
    float s = 0.f, q = 0.f;
    for (int j = 0; j < gs; ++j) { s += raw[2 * (g0 + j)]; q += raw[2 * (g0 + j) + 1]; }   // a group may straddle the two sources
    float mean = s * inv_n;
    const float var = fmaxf(q * inv_n - mean * mean, 0.f);
    const float rstd = rsqrtf(var + 1e-5f);
    const float a = rstd * __ldg(gamma + ch);
    if (shifted && ch < C0) mean -= gn_shift_value(sh, sh_emb, ch);     // (x + s - mean) = x - (mean - s)
    ab[2 * ch] = a;
    ab[2 * ch + 1] = __ldg(beta + ch) - mean * a;
  }
  __syncthreads();
  // bf16 + Swish: silu(z) = h + h*tanh(h) with h = z/2 — the halving is folded into the coefficients (scaling by 0.5 commutes with
  // the rounding of the fma, so the bits are those of silu_fast(fmaf(x, a, b))) and the loop saves one multiply per element
#pragma unroll
  for (int j = 0; j < VEC; ++j) { a[j] = (kHalved ? 0.5f : 1.f) * ab[2 * (c + j)]; b[j] = (kHalved ? 0.5f : 1.f) * ab[2 * (c + j) + 1]; }
  for (; pb < p1; pb += 2 * UN * PY) {
    fetch(pb + UN * PY, rb2, ib, pb2);
    emit(pb, ra, ia, pa);
    fetch(pb + 2 * UN * PY, ra, ia, pa);
    emit(pb + UN * PY, rb2, ib, pb2);
  }
}

// PADDED (H, W) -> PADDED (2H, 2W); one thread per output (row, 16-byte channel vector)
template <typename T>
__global__ void __launch_bounds__(256) upsample2x_kernel(const T* __restrict__ src, int H, int W, int C, size_t total,
                                                         T* __restrict__ dst) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int VEC = VecOf<T>::N;
  const int CV = C / VEC;
  size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int cv = (int)(gid % CV);
  size_t r = gid / CV;
  const int Wo1 = 2 * W + 1, Po = (2 * H + 1) * Wo1;
  const int rem = (int)(r % Po);
  const size_t img = r / Po;
  const int yy = rem / Wo1, xx = rem - yy * Wo1;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (yy > 0 && xx > 0) {
    const int y = (yy - 1) >> 1, x = (xx - 1) >> 1;
    v = *reinterpret_cast<const uint4*>(src + ((img * (H + 1) + (y + 1)) * (W + 1) + (x + 1)) * C + cv * VEC);
  }
  *reinterpret_cast<uint4*>(dst + gid * VEC) = v;
}

// zeroes the padding rows (top row of every image, left column of every line) of a PADDED tensor
template <typename T>
__global__ void __launch_bounds__(256) zero_padding_kernel(T* __restrict__ dst, int H, int W, int C, size_t total) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int VEC = VecOf<T>::N;
  const int CV = C / VEC;
  size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // over (img, padding row index, cv)
  if (gid >= total) return;
  const int cv = (int)(gid % CV);
  size_t r = gid / CV;
  const int npad = H + W + 1;                                      // (W+1) top-row entries + H left-column entries
  const int k = (int)(r % npad);
  const size_t img = r / npad;
  const int W1 = W + 1;
  const size_t row = img * (size_t)(H + 1) * W1 + (k <= W ? k : (size_t)(k - W) * W1);
  *reinterpret_cast<uint4*>(dst + row * C + cv * VEC) = make_uint4(0, 0, 0, 0);
}

// FLAT <-> PADDED copies (tests / taps); one thread per PADDED (row, 16-byte vector)
template <typename T, bool kToPadded>
__global__ void __launch_bounds__(256) repad_kernel(const T* __restrict__ src, int H, int W, int C, size_t total, T* __restrict__ dst) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int VEC = VecOf<T>::N;
  const int CV = C / VEC;
  size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int cv = (int)(gid % CV);
  size_t r = gid / CV;
  const int W1 = W + 1, P = (H + 1) * W1;
  const int rem = (int)(r % P);
  const size_t img = r / P;
  const int yy = rem / W1, xx = rem - yy * W1;
  const bool pad = yy == 0 || xx == 0;
  const size_t flat = ((img * H + (yy - 1)) * W + (xx - 1)) * C + cv * VEC;
  if (kToPadded) {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (!pad) v = *reinterpret_cast<const uint4*>(src + flat);
    *reinterpret_cast<uint4*>(dst + gid * VEC) = v;
  } else if (!pad) {
    *reinterpret_cast<uint4*>(dst + flat) = *reinterpret_cast<const uint4*>(src + gid * VEC);
  }
}

struct GnGeom { int threads, pix_per_cta, splits; };
static GnGeom gn_geom(int C, int vec, int HW, int images) {   // HW = rows per image (P)
  GnGeom g;
  const int CV = C / vec;
  const int PY = kGnThreads / CV > 0 ? kGnThreads / CV : 1;
  g.threads = CV * PY;
  // enough CTAs for ~4 waves of 148 SMs x 4 resident CTAs, but at least 4 pixels per pixel-lane
  int want = cdiv(148 * 16, images > 0 ? images : 1);
  int max_splits = HW / (PY * 4) > 0 ? HW / (PY * 4) : 1;
  g.splits = wave_splits(images > 0 ? images : 1, want, max_splits, 148 * 4);
  if (latency_bound_layer(HW, C) && !getenv("VF_GN_OLD_SPLITS"))
    g.splits = latency_splits(images > 0 ? images : 1, HW, 2 * PY, max_splits, 148 * 4, 3.0, 0.8);   // UN = 2 rows per half-batch
  g.pix_per_cta = cdiv(HW, g.splits);
  g.splits = cdiv(HW, g.pix_per_cta);
  return g;
}

}  // namespace vf

// Test hook (host only, no device needed): the (row splits, threads) the GroupNorm forward would launch with.
extern "C" __attribute__((visibility("default"))) int vf_debug_gn_splits(int images, int H, int W, int C, int dtype, int* threads_out) {
  const int vec = dtype == VF_BF16 ? 8 : 4;
  if (images <= 0 || H <= 0 || W <= 0 || C <= 0 || C % vec || C / vec > vf::kGnThreads) return -1;
  const vf::GnGeom g = vf::gn_geom(C, vec, (H + 1) * (W + 1), images);
  if (threads_out) *threads_out = g.threads;
  return g.splits;
}

extern "C" __attribute__((visibility("default"))) int vf_gn_stats(const void* src0, int C0, const void* src1, int C1, int dtype, int images, int H, int W,
                           float* stats, vf_stream stream) {
  using namespace vf;
  VF_REQUIRE(src0 && stats && images > 0 && H > 0 && W > 0 && C0 > 0, "vf_gn_stats: bad args");
  if (!src1) C1 = 0;
  const int vec = dtype == VF_BF16 ? 8 : 4;
  VF_REQUIRE(C0 % vec == 0 && C1 % vec == 0, "vf_gn_stats: channels (%d,%d) not a multiple of %d", C0, C1, vec);
  const int C = C0 + C1, P = (H + 1) * (W + 1);
  VF_REQUIRE(C / vec <= kGnThreads, "vf_gn_stats: C=%d too large", C);
  GnGeom g = gn_geom(C, vec, P, images);
  dim3 grid(g.splits, images);
  const size_t smem = 2 * C * sizeof(float);
  if (dtype == VF_BF16)
    gn_stats_kernel<__nv_bfloat16><<<grid, g.threads, smem, as_stream(stream)>>>((const __nv_bfloat16*)src0, C0, (const __nv_bfloat16*)src1, C1, W + 1, P, g.pix_per_cta, stats);
  else
    gn_stats_kernel<float><<<grid, g.threads, smem, as_stream(stream)>>>((const float*)src0, C0, (const float*)src1, C1, W + 1, P, g.pix_per_cta, stats);
  VF_LAUNCH_CHECK();
  return VF_OK;
}

extern "C" __attribute__((visibility("default"))) int vf_gn_apply(const void* src0, int C0, const float* stats0, int stats0_ld, const void* src1, int C1,
                           const float* stats1, int stats1_ld, int dtype, int images, int H, int W, int groups,
                           const float* gamma, const float* beta, int swish, void* dst, const vf_gn_shift* shift, vf_stream stream) {
  using namespace vf;
  vf_gn_shift sh{};
  if (shift) sh = *shift;
  VF_REQUIRE(!sh.emb || sh.img_row, "vf_gn_apply: shift.emb needs img_row");
  VF_REQUIRE(src0 && stats0 && gamma && beta && dst && images > 0 && H > 0 && W > 0 && C0 > 0, "vf_gn_apply: bad args");
  if (!src1) C1 = 0;
  VF_REQUIRE(C1 == 0 || stats1, "vf_gn_apply: second source needs statistics");
  VF_REQUIRE(stats0_ld >= C0 && (C1 == 0 || stats1_ld >= C1), "vf_gn_apply: bad statistics stride");
  const int vec = dtype == VF_BF16 ? 8 : 4;
  VF_REQUIRE(C0 % vec == 0 && C1 % vec == 0, "vf_gn_apply: channels (%d,%d) not a multiple of %d", C0, C1, vec);
  const int C = C0 + C1, P = (H + 1) * (W + 1);
  VF_REQUIRE(groups > 0 && C % groups == 0, "vf_gn_apply: C=%d not divisible by groups=%d", C, groups);
  VF_REQUIRE(C / vec <= kGnThreads, "vf_gn_apply: C=%d too large", C);
  cudaStream_t st = as_stream(stream);
  GnGeom g = gn_geom(C, vec, P, images);
  dim3 grid(g.splits, images);
  const size_t smem = 4 * C * sizeof(float);
#define VF_GN_LAUNCH(T, SW) \
  VF_CUDA(launch_pdl(gn_apply_kernel<T, SW>, grid, dim3(g.threads), smem, st, (const T*)src0, C0, (const T*)src1, C1, H * W, W + 1, P, groups, g.pix_per_cta, stats0, stats0_ld, C1 ? stats1 : nullptr, stats1_ld, gamma, beta, (T*)dst, sh))
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
  const size_t total = (size_t)images * (2 * H + 1) * (2 * W + 1) * (C / vec);
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (dtype == VF_BF16) VF_CUDA(launch_pdl(upsample2x_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, as_stream(stream), (const __nv_bfloat16*)src, H, W, C, total, (__nv_bfloat16*)dst));
  else VF_CUDA(launch_pdl(upsample2x_kernel<float>, dim3(grid), dim3(256), 0, as_stream(stream), (const float*)src, H, W, C, total, (float*)dst));
  VF_LAUNCH_CHECK();
  return VF_OK;
}

extern "C" __attribute__((visibility("default"))) int vf_zero_padding(void* dst, int dtype, int images, int H, int W, int C, vf_stream stream) {
  using namespace vf;
  VF_REQUIRE(dst && images > 0 && H > 0 && W > 0 && C > 0, "vf_zero_padding: bad args");
  const int vec = dtype == VF_BF16 ? 8 : 4;
  VF_REQUIRE(C % vec == 0, "vf_zero_padding: C=%d not a multiple of %d", C, vec);
  const size_t total = (size_t)images * (H + W + 1) * (C / vec);
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (dtype == VF_BF16) VF_CUDA(launch_pdl(zero_padding_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, as_stream(stream), (__nv_bfloat16*)dst, H, W, C, total));
  else VF_CUDA(launch_pdl(zero_padding_kernel<float>, dim3(grid), dim3(256), 0, as_stream(stream), (float*)dst, H, W, C, total));
  VF_LAUNCH_CHECK();
  return VF_OK;
}

static int repad(const void* src, int dtype, int images, int H, int W, int C, void* dst, vf_stream stream, bool to_padded) {
  using namespace vf;
  VF_REQUIRE(src && dst && images > 0 && H > 0 && W > 0 && C > 0, "vf_flat_to_padded/vf_padded_to_flat: bad args");
  const int vec = dtype == VF_BF16 ? 8 : 4;
  VF_REQUIRE(C % vec == 0, "repad: C=%d not a multiple of %d", C, vec);
  const size_t total = (size_t)images * (H + 1) * (W + 1) * (C / vec);
  const unsigned grid = (unsigned)((total + 255) / 256);
  cudaStream_t st = as_stream(stream);
  if (dtype == VF_BF16) {
    if (to_padded) VF_CUDA(launch_pdl(repad_kernel<__nv_bfloat16, true>, dim3(grid), dim3(256), 0, st, (const __nv_bfloat16*)src, H, W, C, total, (__nv_bfloat16*)dst));
    else VF_CUDA(launch_pdl(repad_kernel<__nv_bfloat16, false>, dim3(grid), dim3(256), 0, st, (const __nv_bfloat16*)src, H, W, C, total, (__nv_bfloat16*)dst));
  } else {
    if (to_padded) VF_CUDA(launch_pdl(repad_kernel<float, true>, dim3(grid), dim3(256), 0, st, (const float*)src, H, W, C, total, (float*)dst));
    else VF_CUDA(launch_pdl(repad_kernel<float, false>, dim3(grid), dim3(256), 0, st, (const float*)src, H, W, C, total, (float*)dst));
  }
  VF_LAUNCH_CHECK();
  return VF_OK;
}
extern "C" __attribute__((visibility("default"))) int vf_flat_to_padded(const void* src, int dtype, int images, int H, int W, int C, void* dst, vf_stream stream) {
  return repad(src, dtype, images, H, W, C, dst, stream, true);
}
extern "C" __attribute__((visibility("default"))) int vf_padded_to_flat(const void* src, int dtype, int images, int H, int W, int C, void* dst, vf_stream stream) {
  return repad(src, dtype, images, H, W, C, dst, stream, false);
}
