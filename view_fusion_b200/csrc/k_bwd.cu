// Backward operators of the UNet (reference: autograd through model/unet.py, driven by experiment.py:292).
//
//   vf_conv2d_wgrad        dW += dY^T X (per tap / segment) into a packed fp32 buffer, CUDA-core version (all layouts)
//   vf_pack_conv_weight_t  transposed + tap-flipped weight pack, so that the data gradient is a forward vf_conv2d
//   vf_unpack_conv_wgrad   packed fp32 weight gradient -> OIHW parameter gradient
//   vf_gn_backward         GroupNorm (+Swish) backward: two HBM passes (reduce, apply), two-source aware
//   vf_attention_backward  softmax(QK^T/sqrt(C))V backward (CUDA cores)
//   vf_upsample2x_backward / vf_zero_insert2x / vf_add_inplace / vf_grad8_to_act   small layout kernels
#include <mutex>

#include <cstdlib>

#include "vf_common.cuh"

namespace vf {

// ---------------------------------------------------------------------------------------------------------------
// weight gradient, CUDA cores: block = 64 output channels x 32 K-columns, rows split over blockIdx.z
// ---------------------------------------------------------------------------------------------------------------
struct WgradParams {
  RowGeom geo;
  const void* src[3];
  int src_c[3];
  int ksize[3];
  int n_seg;
  const void* dy;
  int dy_ld;
  int cout, cout_pad, k_total;
  float* dwp;
  int rows_per_split;
};

template <typename T>
__global__ void __launch_bounds__(256) conv_wgrad_simt_kernel(const WgradParams p) {
  __shared__ float Ys[16][64 + 4];
  __shared__ float Xs[16][32 + 4];
  const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 64;
  // locate the segment / tap / channel of this K tile (a tile never straddles a tap: C % 32 == 0)
  int seg = 0, koff = 0;
  while (seg + 1 < p.n_seg && k0 >= koff + p.ksize[seg] * p.ksize[seg] * p.src_c[seg]) { koff += p.ksize[seg] * p.ksize[seg] * p.src_c[seg]; ++seg; }
  const int C = p.src_c[seg], ks = p.ksize[seg];
  const int tap = (k0 - koff) / C, c0 = (k0 - koff) - tap * C;
  const int shift = ks == 3 ? (tap / 3 - 1) * p.geo.W1 + (tap % 3 - 1) : 0;
  const T* X = reinterpret_cast<const T*>(p.src[seg]);
  const T* dY = reinterpret_cast<const T*>(p.dy);
  const int M = p.geo.rows_total;
  const int r0 = blockIdx.z * p.rows_per_split, r1 = min(M, r0 + p.rows_per_split);
  const int tid = threadIdx.x;
  const int tn = tid % 16, tk = tid / 16;          // thread computes n = tn*4..+3, k = tk*2..+1
  float acc[4][2] = {};
  for (int rb = r0; rb < r1; rb += 16) {
    // stage 16 rows of dY (64 n) and of the shifted X (32 c)
    for (int i = tid; i < 16 * 64; i += 256) {
      const int rr = i / 64, n = i % 64;
      const int m = rb + rr;
      float v = 0.f;
      if (m < r1 && n0 + n < p.cout) {
        const RowInfo ri = decode_row(p.geo, m);
        if (ri.valid) v = to_f(dY[(size_t)ri.out_row * p.dy_ld + n0 + n]);
      }
      Ys[rr][n] = v;
    }
    for (int i = tid; i < 16 * 32; i += 256) {
      const int rr = i / 32, c = i % 32;
      const long m = (long)rb + rr + shift;
      float v = 0.f;
      if (rb + rr < r1 && m >= 0 && m < M) v = to_f(X[(size_t)m * C + c0 + c]);
      Xs[rr][c] = v;
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < 16; ++rr) {
      float y[4], x[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) y[i] = Ys[rr][tn * 4 + i];
      x[0] = Xs[rr][tk * 2]; x[1] = Xs[rr][tk * 2 + 1];
#pragma unroll
      for (int i = 0; i < 4; ++i) { acc[i][0] += y[i] * x[0]; acc[i][1] += y[i] * x[1]; }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + tn * 4 + i;
    if (n >= p.cout) continue;
#pragma unroll
    for (int j = 0; j < 2; ++j) atomicAdd(p.dwp + (size_t)n * p.k_total + k0 + tk * 2 + j, acc[i][j]);
  }
}

template <typename T>
__global__ void pack_conv_weight_t_kernel(const float* __restrict__ w, int cout, int cin, int kk, T* __restrict__ dst, int cin_pad,
                                          int k_total, int k_off, int n_stride) {
  // dst[c][k_off + tap' * n_stride + n] = w[n][c][tap], tap' = kk-1-tap (the data gradient correlates with flipped taps)
  const size_t total = (size_t)cin_pad * kk * n_stride;
  size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int c = (int)(gid / ((size_t)kk * n_stride));
  const int r = (int)(gid % ((size_t)kk * n_stride));
  const int tapf = r / n_stride, n = r % n_stride;
  const int tap = kk - 1 - tapf;
  const float v = (c < cin && n < cout) ? __ldg(w + ((size_t)n * cin + c) * kk + tap) : 0.f;
  dst[(size_t)c * k_total + k_off + r] = from_f<T>(v);
}

__global__ void unpack_conv_wgrad_kernel(const float* __restrict__ dwp, int cout, int cin, int kk, int k_total, int k_off,
                                         float* __restrict__ dw, int cin_total, int c_off) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t total = (size_t)cout * cin * kk;
  size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // over the segment's [cout][cin][kk] slice
  if (gid >= total) return;
  const int tap = (int)(gid % kk);
  const int c = (int)((gid / kk) % cin);
  const int n = (int)(gid / ((size_t)kk * cin));
  dw[((size_t)n * cin_total + c_off + c) * kk + tap] += dwp[(size_t)n * k_total + k_off + tap * cin + c];
}

// ---------------------------------------------------------------------------------------------------------------
// GroupNorm (+Swish) backward
//   z = xh*gamma + beta, xh = (x-mu)*rstd, y = swish(z);  dz = dy * swish'(z)
//   pass 1: per (image, channel)  A = sum dz, B = sum dz*xh         (+ dgamma += B, dbeta += A)
//   pass 2: dx = rstd * (gamma*dz - S1/n - xh*S2/n),  S1 = sum_group gamma*A, S2 = sum_group gamma*B
// ---------------------------------------------------------------------------------------------------------------
constexpr int kGbThreads = 256;

struct GnBwdParams {
  const void* s0; int C0; const float* st0; int ld0;
  const void* s1; int C1; const float* st1; int ld1;
  int HW, W1, P, groups, rows_per_cta, swish, img0;
  const float* gamma; const float* beta;
  const void* dy;
  float* red;            // [images][C][2]
  float* dgamma; float* dbeta;
  void* dx0; int acc0; void* dx1; int acc1;
  GnColsum cs;           // optional (cs.db / cs.demb non-null): closed-form column sums of dx0 (single-source case)
  vf_gn_shift sh;        // optional: source 0 is stored without a per-(image, channel) constant (vf_gn_apply)
};

// d/dz [z * sigmoid(z)]; the bf16 path takes sigmoid from one tanh.approx like the forward kernel
template <typename T> __device__ __forceinline__ float dswish_for(float z);
template <> __device__ __forceinline__ float dswish_for<float>(float z) {
  const float s = 1.f / (1.f + __expf(-z));
  return s * (1.f + z * (1.f - s));
}
// g * swish'(z) given h = z/2:  swish'(z) = 0.5 * (1 + t + h*(1 - t*t)),  t = tanh(h)
template <typename T> __device__ __forceinline__ float dz_swish(float g, float h);
template <> __device__ __forceinline__ float dz_swish<float>(float g, float h) {
  const float t = tanhf(h);
  const float q = fmaf(h, fmaf(-t, t, 1.f), t);
  const float gh = 0.5f * g;
  return fmaf(gh, q, gh);
}
template <> __device__ __forceinline__ float dz_swish<__nv_bfloat16>(float g, float h) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  const float q = fmaf(h, fmaf(-t, t, 1.f), t);
  const float gh = 0.5f * g;
  return fmaf(gh, q, gh);
}
template <> __device__ __forceinline__ float dswish_for<__nv_bfloat16>(float z) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * z));
  const float s = 0.5f * t + 0.5f;
  return s * (1.f + z * (1.f - s));
}

// shared prologue: per channel mu / rstd of its group -> smem mr[C][2].  The raw (sum, sumsq) pairs of the image are
// staged into smem with one coalesced load first, so the per-group loops run out of shared memory (one global round
// trip instead of one per group member).  `raw` is scratch of 2C floats; ends with a __syncthreads().
__device__ __forceinline__ void gn_group_stats(const GnBwdParams& p, int img, float* mr, float* raw) {
  const int C = p.C0 + p.C1, gs = C / p.groups;
  const float inv_n = 1.f / ((float)gs * (float)p.HW);
  const float* sa = p.st0 + (size_t)img * p.ld0 * 2;
  const float* sb = p.st1 ? p.st1 + (size_t)img * p.ld1 * 2 : nullptr;
  const bool shifted = p.sh.bias != nullptr || p.sh.emb != nullptr;
  const float* sh_emb = p.sh.emb ? p.sh.emb + (size_t)__ldg(p.sh.img_row + img) * p.sh.emb_ld : nullptr;
  if (!shifted) {
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) raw[i] = i < 2 * p.C0 ? __ldg(sa + i) : __ldg(sb + (i - 2 * p.C0));
  } else {
    // source 0 was normalised as x + s (vf_gn_shift): shift its raw sums the way the forward did
    for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
      float S1, S2;
      if (ch < p.C0) {
        S1 = __ldg(sa + 2 * ch); S2 = __ldg(sa + 2 * ch + 1);
        const float sv = gn_shift_value(p.sh, sh_emb, ch);
        S2 = fmaf(2.f * sv, S1, S2) + (float)p.HW * sv * sv;
        S1 = fmaf((float)p.HW, sv, S1);
      } else {
        S1 = __ldg(sb + 2 * (ch - p.C0)); S2 = __ldg(sb + 2 * (ch - p.C0) + 1);
      }
      raw[2 * ch] = S1; raw[2 * ch + 1] = S2;
    }
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
    const int g0 = ch / gs * gs;
    float s = 0.f, q = 0.f;
    for (int j = 0; j < gs; ++j) { s += raw[2 * (g0 + j)]; q += raw[2 * (g0 + j) + 1]; }
    float mean = s * inv_n;
    const float var = fmaxf(q * inv_n - mean * mean, 0.f);
    // every later use of the mean is (x - mean) with the STORED x: with a shift that is x - (mean - s)
    if (shifted && ch < p.C0) mean -= gn_shift_value(p.sh, sh_emb, ch);
    mr[2 * ch] = mean;
    mr[2 * ch + 1] = rsqrtf(var + 1e-5f);
  }
  __syncthreads();
}

constexpr int kGbUnroll = 4;     // rows in flight per thread: independent 16-byte loads issued before any use

// (yy, xx) of a thread's PADDED row advance incrementally by a constant stride: no division in the streaming loops
struct RowWalk {
  int yy, xx, dy, dx, W1;
  __device__ __forceinline__ RowWalk(int row, int stride, int w1) : W1(w1) {
    yy = row / w1; xx = row - yy * w1;
    dy = stride / w1; dx = stride - dy * w1;
  }
  __device__ __forceinline__ bool pad() const { return yy == 0 || xx == 0; }
  __device__ __forceinline__ void step() {
    yy += dy; xx += dx;
    if (xx >= W1) { xx -= W1; ++yy; }
  }
};

template <typename T>
__global__ void __launch_bounds__(kGbThreads, 2) gn_bwd_reduce_kernel(const GnBwdParams p) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int VEC = VecOf<T>::N;
  constexpr int UN = kGbUnroll;    // (two or three rows in flight at 3 CTAs per SM measured 5-20 % slower)
  extern __shared__ float sm[];
  const int C = p.C0 + p.C1, CV = C / VEC;
  float* mr = sm;              // [C][2]
  float* part = sm + 2 * C;    // [PY][C][2] per-row-lane partial sums (also the staging scratch of the prologue)
  const int img = p.img0 + blockIdx.y;
  gn_group_stats(p, img, mr, part);
  const int PY = blockDim.x / CV;
  const int cv = threadIdx.x % CV, py = threadIdx.x / CV;
  const int c = cv * VEC;
  const T* src = c < p.C0 ? (const T*)p.s0 + (size_t)img * p.P * p.C0 + c : (const T*)p.s1 + (size_t)img * p.P * p.C1 + (c - p.C0);
  const int ld = c < p.C0 ? p.C0 : p.C1;
  const T* dy = (const T*)p.dy + (size_t)img * p.P * C + c;
  // per-channel constants: h = z/2 = x*cA + cD with z the GroupNorm output.  The sums are taken over dz and dz*x;
  // sum(dz * xhat) = rs * sum(dz*x) - mu*rs * sum(dz) is formed once at the end.
  float cA[VEC], cD[VEC], sA[VEC], sB[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const float mu = mr[2 * (c + j)], rs = mr[2 * (c + j) + 1], ga = __ldg(p.gamma + c + j), be = __ldg(p.beta + c + j);
    cA[j] = 0.5f * rs * ga; cD[j] = 0.5f * (be - mu * rs * ga);
    sA[j] = sB[j] = 0.f;
  }
  const int p0 = blockIdx.x * p.rows_per_cta, p1 = min(p.P, p0 + p.rows_per_cta);
  RowWalk rw(p0 + py, PY, p.W1);
  for (int rb = p0 + py; rb < p1; rb += UN * PY) {
    uint4 xr[UN], gr[UN];
    uint32_t live = 0;
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int r = rb + u * PY;
      const bool ok = r < p1 && !rw.pad();
      live |= (uint32_t)ok << u;
      if (ok) {
        xr[u] = *reinterpret_cast<const uint4*>(src + (size_t)r * ld);
        gr[u] = *reinterpret_cast<const uint4*>(dy + (size_t)r * C);
      }
      rw.step();
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      if (!((live >> u) & 1u)) continue;
      float x[VEC], g[VEC];
      load_vec(reinterpret_cast<const T*>(&xr[u]), x);
      load_vec(reinterpret_cast<const T*>(&gr[u]), g);
      if (p.swish) {
        // dz = dy * swish'(z) replaces dy in place (stored in the activation type): the apply pass then needs no
        // transcendental at all.  (Recomputing dz in the apply pass instead of this 16-byte store was measured: one tensor
        // pass less, but 131 -> 166 us per 64x64 x 64 layer, the tanh + FMAs cost the apply pass more than the store costs here.)
#pragma unroll
        for (int j = 0; j < VEC; ++j) g[j] = dz_swish<T>(g[j], fmaf(x[j], cA[j], cD[j]));
        store_vec(const_cast<T*>(dy) + (size_t)(rb + u * PY) * C, g);
      }
#pragma unroll
      for (int j = 0; j < VEC; ++j) { sA[j] += g[j]; sB[j] = fmaf(g[j], x[j], sB[j]); }
    }
  }
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const float mu = mr[2 * (c + j)], rs = mr[2 * (c + j) + 1];
    part[(py * C + c + j) * 2] = sA[j];
    part[(py * C + c + j) * 2 + 1] = rs * (sB[j] - mu * sA[j]);       // sum(dz * xhat)
  }
  __syncthreads();
  // per-image sums; gn_bwd_apply folds them into dgamma / dbeta (one CTA per image)
  float* red = p.red + (size_t)img * C * 2;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float t = 0.f;
    for (int q = 0; q < PY; ++q) t += part[q * 2 * C + i];
    atomicAdd(red + i, t);
  }
}

template <typename T>
__global__ void __launch_bounds__(kGbThreads, 3) gn_bwd_apply_kernel(const GnBwdParams p) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int VEC = VecOf<T>::N;
  constexpr int UN = 2;            // up to three 16-byte loads per row: two rows in flight keep the kernel at 3 CTAs per SM
  extern __shared__ float sm[];
  const int C = p.C0 + p.C1, CV = C / VEC;
  float* mr = sm;              // [C][2] mean, rstd
  float* tt = sm + 2 * C;      // [C][2] t1 = rstd*S1/n, t2 = rstd*S2/n
  float* raw = sm + 4 * C;     // [C][2] staging: forward statistics, then gamma-weighted reduce sums
  const int img = p.img0 + blockIdx.y;
  gn_group_stats(p, img, mr, raw);
  {
    const int gs = C / p.groups;
    const float inv_n = 1.f / ((float)gs * (float)p.HW);
    const float* red = p.red + (size_t)img * C * 2;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
      const float r = __ldg(red + i);
      raw[i] = r * __ldg(p.gamma + (i >> 1));
      // parameter gradients: dbeta = sum_img red[.][0], dgamma = sum_img red[.][1] (one CTA per image adds them)
      if (blockIdx.x == 0) atomicAdd(((i & 1) ? p.dgamma : p.dbeta) + (i >> 1), r);
    }
    __syncthreads();
    for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
      const int g0 = ch / gs * gs;
      float S1 = 0.f, S2 = 0.f;
      for (int j = 0; j < gs; ++j) { S1 += raw[2 * (g0 + j)]; S2 += raw[2 * (g0 + j) + 1]; }
      tt[2 * ch] = mr[2 * ch + 1] * S1 * inv_n;
      tt[2 * ch + 1] = mr[2 * ch + 1] * S2 * inv_n;
      if (blockIdx.x == 0 && (p.cs.db || p.cs.demb)) {
        // sum over the image's pixels of dx = cA*dz + cB*x + cC (constants as below): the bias / embedding gradient of the
        // convolution that produced x, without reading dx back
        const float mu = mr[2 * ch], rs = mr[2 * ch + 1], ga = __ldg(p.gamma + ch);
        const float t1 = tt[2 * ch], t2 = tt[2 * ch + 1];
        const float sum_dz = __ldg(red + 2 * ch), sum_x = __ldg(p.st0 + ((size_t)img * p.ld0 + ch) * 2);
        const float v = rs * ga * sum_dz - rs * t2 * sum_x + (mu * rs * t2 - t1) * (float)p.HW;
        if (p.cs.db) atomicAdd(p.cs.db + ch, v);
        if (p.cs.demb) atomicAdd(p.cs.demb + (size_t)__ldg(p.cs.img_row + img) * p.cs.emb_ld + p.cs.col + ch, v);
      }
    }
  }
  __syncthreads();
  const int PY = blockDim.x / CV;
  const int cv = threadIdx.x % CV, py = threadIdx.x / CV;
  const int c = cv * VEC;
  const bool first = c < p.C0;
  const T* src = first ? (const T*)p.s0 + (size_t)img * p.P * p.C0 + c : (const T*)p.s1 + (size_t)img * p.P * p.C1 + (c - p.C0);
  const int ld = first ? p.C0 : p.C1;
  T* dx = first ? (T*)p.dx0 + (size_t)img * p.P * p.C0 + c : (T*)p.dx1 + (size_t)img * p.P * p.C1 + (c - p.C0);
  const bool accum = first ? p.acc0 : p.acc1;
  const T* dy = (const T*)p.dy + (size_t)img * p.P * C + c;
  // per-channel constants: dx = cA*dz + x*cB + cC  with cA = rs*gamma, cB = -rs*t2, cC = mu*rs*t2 - t1
  float cA[VEC], cB[VEC], cC[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const float mu = mr[2 * (c + j)], rs = mr[2 * (c + j) + 1], ga = __ldg(p.gamma + c + j);
    const float t1 = tt[2 * (c + j)], t2 = tt[2 * (c + j) + 1];
    cA[j] = rs * ga; cB[j] = -rs * t2; cC[j] = mu * rs * t2 - t1;
  }
  const int p0 = blockIdx.x * p.rows_per_cta, p1 = min(p.P, p0 + p.rows_per_cta);
  RowWalk rw(p0 + py, PY, p.W1);
  for (int rb = p0 + py; rb < p1; rb += UN * PY) {
    uint4 xr[UN], gr[UN], orr[UN];
    uint32_t live = 0, padm = 0;
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int r = rb + u * PY;
      const bool in = r < p1, pad = rw.pad();
      live |= (uint32_t)in << u; padm |= (uint32_t)pad << u;
      if (in && !pad) {
        xr[u] = *reinterpret_cast<const uint4*>(src + (size_t)r * ld);
        gr[u] = *reinterpret_cast<const uint4*>(dy + (size_t)r * C);
        if (accum) orr[u] = *reinterpret_cast<const uint4*>(dx + (size_t)r * ld);
      }
      rw.step();
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      if (!((live >> u) & 1u)) continue;
      const int r = rb + u * PY;
      float o[VEC];
      if ((padm >> u) & 1u) {
        if (accum) continue;
#pragma unroll
        for (int j = 0; j < VEC; ++j) o[j] = 0.f;          // gradients of padding rows are exact zeros
      } else {
        float x[VEC], g[VEC];
        load_vec(reinterpret_cast<const T*>(&xr[u]), x);
        load_vec(reinterpret_cast<const T*>(&gr[u]), g);
#pragma unroll
        for (int j = 0; j < VEC; ++j) o[j] = fmaf(cA[j], g[j], fmaf(x[j], cB[j], cC[j]));     // g = dz (written by the reduce pass)
        if (accum) {
          float old[VEC];
          load_vec(reinterpret_cast<const T*>(&orr[u]), old);
#pragma unroll
          for (int j = 0; j < VEC; ++j) o[j] += old[j];
        }
      }
      store_vec(dx + (size_t)r * ld, o);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// One-pass GroupNorm (+Swish) backward: a CTA owns (image, slab of whole groups) and keeps the slab's x and dz in shared
// memory between the reduction and the apply phase, so x and dy are read from HBM ONCE and nothing but dx is written
// (3 tensor passes instead of the 6-7 of gn_bwd_reduce + gn_bwd_apply: dy is no longer rewritten with dz, and the second
// read of x / dz never leaves the SM).  No cross-CTA reduction: groups are contiguous channel ranges, the slab holds whole
// groups.  Used whenever the slab fits (gn_fused_slab); the two-pass kernels remain for the rest (64x64 x 192 channels).
// Same arithmetic as the two-pass kernels (dz rounded to the activation type before the apply phase, sums taken in fp32
// from the unrounded values); the partial sums are combined in a fixed order instead of by atomics.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kGfUnroll = 4;
struct FlatWalk {          // (y, x) of pixel index i advancing by a constant stride, -> PADDED row (y+1)*W1 + x+1
  int yy, xx, dy, dx, W;
  __device__ __forceinline__ FlatWalk(int i, int stride, int w) : W(w) { yy = i / w; xx = i - yy * w; dy = stride / w; dx = stride - dy * w; }
  __device__ __forceinline__ int prow(int W1) const { return (yy + 1) * W1 + xx + 1; }
  __device__ __forceinline__ void step() { yy += dy; xx += dx; if (xx >= W) { xx -= W; ++yy; } }
};

template <typename T>
__global__ void __launch_bounds__(512) gn_bwd_fused_kernel(const GnBwdParams p, int SC) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int VEC = VecOf<T>::N;
  constexpr int UN = kGfUnroll;
  extern __shared__ __align__(16) uint8_t gf_smem[];
  const int C = p.C0 + p.C1, gs = C / p.groups, W = p.W1 - 1, HW = p.HW;
  const int img = p.img0 + blockIdx.y, c_lo = blockIdx.x * SC;
  const int SCV = SC / VEC, PY = blockDim.x / SCV;           // blockDim = SCV * PY exactly
  float* mr = reinterpret_cast<float*>(gf_smem);              // [SC][2] mean (of the stored x), rstd
  float* tt = mr + 2 * SC;                                     // [SC][2] t1, t2
  float* raw = tt + 2 * SC;                                    // [SC][2] staging
  float* part2 = raw + 2 * SC;                                 // [blockDim + 2 SC] second reduction stage
  float* part = part2 + blockDim.x + 2 * SC;                   // [PY][SC][2]
  T* xs = reinterpret_cast<T*>((reinterpret_cast<uintptr_t>(part + (size_t)PY * SC * 2) + 15) & ~(uintptr_t)15);    // [HW][SC]
  T* ds = xs + (size_t)HW * SC;                                // [HW][SC] dz in the activation type
  const float inv_n = 1.f / ((float)gs * (float)HW);
  const bool shifted = p.sh.bias != nullptr || p.sh.emb != nullptr;
  const float* sh_emb = p.sh.emb ? p.sh.emb + (size_t)__ldg(p.sh.img_row + img) * p.sh.emb_ld : nullptr;
  // ---- forward statistics of the slab's groups -> mean / rstd per channel (as gn_group_stats) ----
  for (int i = threadIdx.x; i < SC; i += blockDim.x) {
    const int ch = c_lo + i;
    float S1, S2;
    if (ch < p.C0) {
      const float* sa = p.st0 + ((size_t)img * p.ld0 + ch) * 2;
      S1 = __ldg(sa); S2 = __ldg(sa + 1);
      if (shifted) {
        const float sv = gn_shift_value(p.sh, sh_emb, ch);
        S2 = fmaf(2.f * sv, S1, S2) + (float)HW * sv * sv;
        S1 = fmaf((float)HW, sv, S1);
      }
    } else {
      const float* sb = p.st1 + ((size_t)img * p.ld1 + (ch - p.C0)) * 2;
      S1 = __ldg(sb); S2 = __ldg(sb + 1);
    }
    raw[2 * i] = S1; raw[2 * i + 1] = S2;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < SC; i += blockDim.x) {
    const int g0 = i / gs * gs;
    float s = 0.f, q = 0.f;
    for (int j = 0; j < gs; ++j) { s += raw[2 * (g0 + j)]; q += raw[2 * (g0 + j) + 1]; }
    float mean = s * inv_n;
    const float var = fmaxf(q * inv_n - mean * mean, 0.f);
    if (shifted && c_lo + i < p.C0) mean -= gn_shift_value(p.sh, sh_emb, c_lo + i);
    mr[2 * i] = mean;
    mr[2 * i + 1] = rsqrtf(var + 1e-5f);
  }
  __syncthreads();
  // ---- phase 1: stream x, dy in; dz and the per-channel sums ----
  const int cvl = threadIdx.x % SCV, py = threadIdx.x / SCV;
  const int cl = cvl * VEC, c = c_lo + cl;                    // local / global first channel of the thread's vector
  const bool first = c < p.C0;
  const T* src = first ? (const T*)p.s0 + (size_t)img * p.P * p.C0 + c : (const T*)p.s1 + (size_t)img * p.P * p.C1 + (c - p.C0);
  const int ld = first ? p.C0 : p.C1;
  const T* dy = (const T*)p.dy + (size_t)img * p.P * C + c;
  float cA[VEC], cD[VEC], sA[VEC], sB[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const float mu = mr[2 * (cl + j)], rs = mr[2 * (cl + j) + 1], ga = __ldg(p.gamma + c + j), be = __ldg(p.beta + c + j);
    cA[j] = 0.5f * rs * ga; cD[j] = 0.5f * (be - mu * rs * ga);
    sA[j] = sB[j] = 0.f;
  }
  {
    FlatWalk fw(py, PY, W);
    for (int ib = py; ib < HW; ib += UN * PY) {
      uint4 xr[UN], gr[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        if (ib + u * PY < HW) {
          const size_t r = (size_t)fw.prow(p.W1);
          xr[u] = *reinterpret_cast<const uint4*>(src + r * ld);
          gr[u] = *reinterpret_cast<const uint4*>(dy + r * C);
        }
        fw.step();
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int i = ib + u * PY;
        if (i >= HW) continue;
        float x[VEC], g[VEC];
        load_vec(reinterpret_cast<const T*>(&xr[u]), x);
        load_vec(reinterpret_cast<const T*>(&gr[u]), g);
        if (p.swish) {
#pragma unroll
          for (int j = 0; j < VEC; ++j) g[j] = dz_swish<T>(g[j], fmaf(x[j], cA[j], cD[j]));
        }
        *reinterpret_cast<uint4*>(xs + (size_t)i * SC + cl) = xr[u];
        store_vec(ds + (size_t)i * SC + cl, g);
#pragma unroll
        for (int j = 0; j < VEC; ++j) { sA[j] += g[j]; sB[j] = fmaf(g[j], x[j], sB[j]); }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    part[(py * SC + cl + j) * 2] = sA[j];
    part[(py * SC + cl + j) * 2 + 1] = sB[j];
  }
  __syncthreads();
  // two-stage sum over the PY row lanes: (value i, chunk q) then chunks
  const int nv = 2 * SC, nch = blockDim.x / nv > 0 ? blockDim.x / nv : 1;
  for (int j = threadIdx.x; j < nv * nch; j += blockDim.x) {
    const int i = j % nv, q = j / nv;
    float t = 0.f;
    for (int r = q; r < PY; r += nch) t += part[r * nv + i];
    part2[q * nv + i] = t;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < SC; i += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int q = 0; q < nch; ++q) { a += part2[q * nv + 2 * i]; b += part2[q * nv + 2 * i + 1]; }
    const int ch = c_lo + i;
    const float mu = mr[2 * i], rs = mr[2 * i + 1], ga = __ldg(p.gamma + ch);
    const float bx = rs * (b - mu * a);                        // sum(dz * xhat)
    atomicAdd(p.dbeta + ch, a);
    atomicAdd(p.dgamma + ch, bx);
    raw[2 * i] = a * ga; raw[2 * i + 1] = bx * ga;
    tt[2 * i] = a;                                             // kept for the closed-form column sums below
  }
  __syncthreads();
  for (int i = threadIdx.x; i < SC; i += blockDim.x) {
    const int g0 = i / gs * gs;
    float S1 = 0.f, S2 = 0.f;
    for (int j = 0; j < gs; ++j) { S1 += raw[2 * (g0 + j)]; S2 += raw[2 * (g0 + j) + 1]; }
    const float sum_dz = tt[2 * i];
    const float rs = mr[2 * i + 1];
    const float t1 = rs * S1 * inv_n, t2 = rs * S2 * inv_n;
    tt[2 * i] = t1; tt[2 * i + 1] = t2;                        // (entry i is read and written by this thread only)
    if (p.cs.db || p.cs.demb) {
      const int ch = c_lo + i;
      const float mu = mr[2 * i], ga = __ldg(p.gamma + ch);
      const float sum_x = __ldg(p.st0 + ((size_t)img * p.ld0 + ch) * 2);
      const float v = rs * ga * sum_dz - rs * t2 * sum_x + (mu * rs * t2 - t1) * (float)HW;
      if (p.cs.db) atomicAdd(p.cs.db + ch, v);
      if (p.cs.demb) atomicAdd(p.cs.demb + (size_t)__ldg(p.cs.img_row + img) * p.cs.emb_ld + p.cs.col + ch, v);
    }
  }
  __syncthreads();
  // ---- phase 2: dx = cA*dz + cB*x + cC out of shared memory ----
  T* dx = first ? (T*)p.dx0 + (size_t)img * p.P * p.C0 + c : (T*)p.dx1 + (size_t)img * p.P * p.C1 + (c - p.C0);
  const bool accum = first ? p.acc0 : p.acc1;
  float kA[VEC], kB[VEC], kC[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const float mu = mr[2 * (cl + j)], rs = mr[2 * (cl + j) + 1], ga = __ldg(p.gamma + c + j);
    const float t1 = tt[2 * (cl + j)], t2 = tt[2 * (cl + j) + 1];
    kA[j] = rs * ga; kB[j] = -rs * t2; kC[j] = mu * rs * t2 - t1;
  }
  {
    FlatWalk fw(py, PY, W);
    for (int ib = py; ib < HW; ib += UN * PY) {
      uint4 orr[UN];
      size_t rr[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        rr[u] = (size_t)fw.prow(p.W1);
        if (accum && ib + u * PY < HW) orr[u] = *reinterpret_cast<const uint4*>(dx + rr[u] * ld);
        fw.step();
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int i = ib + u * PY;
        if (i >= HW) continue;
        float x[VEC], g[VEC], o[VEC];
        load_vec(xs + (size_t)i * SC + cl, x);
        load_vec(ds + (size_t)i * SC + cl, g);
#pragma unroll
        for (int j = 0; j < VEC; ++j) o[j] = fmaf(kA[j], g[j], fmaf(x[j], kB[j], kC[j]));
        if (accum) {
          float old[VEC];
          load_vec(reinterpret_cast<const T*>(&orr[u]), old);
#pragma unroll
          for (int j = 0; j < VEC; ++j) o[j] += old[j];
        }
        store_vec(dx + rr[u] * ld, o);
      }
    }
  }
  // gradients of padding rows are exact zeros (row 0 and column 0 of the PADDED image)
  if (!accum) {
    float z[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) z[j] = 0.f;
    const int H = HW / W, npad = p.W1 + H;
    for (int k = py; k < npad; k += PY) {
      const int r = k < p.W1 ? k : (k - p.W1 + 1) * p.W1;
      store_vec(dx + (size_t)r * ld, z);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// attention backward (CUDA cores): per (image, 8 queries)
// ---------------------------------------------------------------------------------------------------------------
constexpr int BQ = 8;
template <typename T>
__global__ void __launch_bounds__(256) attn_bwd_simt_kernel(const T* __restrict__ qk, int ld, const T* __restrict__ vt, const T* __restrict__ dO,
                                                            int L, int C, T* __restrict__ dq_out, int dq_ld, float* __restrict__ dkv) {
  extern __shared__ float sm[];
  float* q = sm;                 // [BQ][C]
  float* go = q + BQ * C;        // [BQ][C]  dO rows
  float* pr = go + BQ * C;       // [BQ][L]  P then dS
  float* dp = pr + BQ * L;       // [BQ][L]
  const int img = blockIdx.y, q0 = blockIdx.x * BQ;
  const T* base = qk + (size_t)img * L * ld;
  for (int i = threadIdx.x; i < BQ * C; i += blockDim.x) {
    const int qi = i / C, c = i % C;
    const bool in = q0 + qi < L;
    q[i] = in ? to_f(base[(size_t)(q0 + qi) * ld + c]) : 0.f;
    go[i] = in ? to_f(dO[((size_t)img * L + q0 + qi) * C + c]) : 0.f;
  }
  __syncthreads();
  const float scale = rsqrtf((float)C);
  for (int i = threadIdx.x; i < BQ * L; i += blockDim.x) {
    const int qi = i / L, kj = i % L;
    const T* kr = base + (size_t)kj * ld + C;
    float s = 0.f, d = 0.f;
    for (int c = 0; c < C; ++c) {
      s += q[qi * C + c] * to_f(kr[c]);
      const float v = vt ? to_f(vt[((size_t)img * C + c) * L + kj]) : to_f(base[(size_t)kj * ld + 2 * C + c]);
      d += go[qi * C + c] * v;
    }
    pr[i] = s * scale;
    dp[i] = d;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int qi = warp; qi < BQ; qi += blockDim.x >> 5) {
    float* row = pr + qi * L;
    float* drow = dp + qi * L;
    float m = -INFINITY;
    for (int j = lane; j < L; j += 32) m = fmaxf(m, row[j]);
    m = warp_max(m);
    float sum = 0.f;
    for (int j = lane; j < L; j += 32) { const float e = expf(row[j] - m); row[j] = e; sum += e; }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    float delta = 0.f;
    for (int j = lane; j < L; j += 32) { row[j] *= inv; delta += row[j] * drow[j]; }
    delta = warp_sum(delta);
    // dV[j][c] += P[qi][j] * dO[qi][c] is accumulated below; here dS = P * (dP - delta)
    for (int j = lane; j < L; j += 32) { const float pj = row[j]; drow[j] = pj * (drow[j] - delta); }
  }
  __syncthreads();
  // dQ[qi][c] = scale * sum_j dS[qi][j] K[j][c]
  for (int i = threadIdx.x; i < BQ * C; i += blockDim.x) {
    const int qi = i / C, c = i % C;
    if (q0 + qi >= L) continue;
    float acc = 0.f;
    for (int j = 0; j < L; ++j) acc += dp[qi * L + j] * to_f(base[(size_t)j * ld + C + c]);
    dq_out[((size_t)img * L + q0 + qi) * dq_ld + c] = from_f<T>(acc * scale);
  }
  // dK[j][c] += scale * sum_qi dS[qi][j] Q[qi][c];  dV[j][c] += sum_qi P[qi][j] dO[qi][c]   (fp32 atomics)
  float* dk = dkv + (size_t)img * L * 2 * C;
  for (int i = threadIdx.x; i < L * C; i += blockDim.x) {
    const int j = i / C, c = i % C;
    float ak = 0.f, av = 0.f;
#pragma unroll
    for (int qi = 0; qi < BQ; ++qi) { ak += dp[qi * L + j] * q[qi * C + c]; av += pr[qi * L + j] * go[qi * C + c]; }
    atomicAdd(dk + (size_t)j * 2 * C + c, ak * scale);
    atomicAdd(dk + (size_t)j * 2 * C + C + c, av);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// attention backward for short sequences (L = 64: the 8x8 mid block, C = 320): one CTA per image, every operand of the
// image resident in shared memory (bf16), the five products on CUDA cores with register tiles.  No atomics, no
// scratch: dq | dk | dv rows are written straight into dqkv.
// ---------------------------------------------------------------------------------------------------------------
constexpr int AL = 64;                 // tokens per image
struct AttnL64Smem {
  static __host__ __device__ int ldh(int C) { return C + 4; }          // bf16 row stride: (C+4)/2 words == 2 (mod 32) for C % 64 == 0
  static __host__ __device__ size_t bytes(int C) {
    return (size_t)3 * AL * ldh(C) * 2 + (size_t)C * (AL + 2) * 2 + (size_t)2 * AL * (AL + 1) * 4;
  }
};

__global__ void __launch_bounds__(256) attn_bwd_l64_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ vt,
                                                           const __nv_bfloat16* __restrict__ dO, int C, __nv_bfloat16* __restrict__ dqkv) {
  extern __shared__ __align__(16) uint8_t smraw[];
  const int LDH = AttnL64Smem::ldh(C), LDV = AL + 2, LDP = AL + 1;
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(smraw);
  __nv_bfloat16* Ks = Qs + AL * LDH;
  __nv_bfloat16* Gs = Ks + AL * LDH;                 // dO
  __nv_bfloat16* Vts = Gs + AL * LDH;                // [C][LDV]
  float* Ps = reinterpret_cast<float*>(Vts + (size_t)C * LDV);
  float* Ds = Ps + AL * LDP;
  const int img = blockIdx.x, t = threadIdx.x;
  const int ld = 3 * C;
  const __nv_bfloat16* base = qkv + (size_t)img * AL * ld;
  // ---- stage Q, K, dO rows (8-byte pieces: the padded rows are 8-byte aligned) and V^T
  const int c4 = C / 4;
  for (int i = t; i < AL * c4; i += 256) {
    const int r = i / c4, c = (i - r * c4) * 4;
    *reinterpret_cast<uint2*>(Qs + r * LDH + c) = *reinterpret_cast<const uint2*>(base + (size_t)r * ld + c);
    *reinterpret_cast<uint2*>(Ks + r * LDH + c) = *reinterpret_cast<const uint2*>(base + (size_t)r * ld + C + c);
    *reinterpret_cast<uint2*>(Gs + r * LDH + c) = *reinterpret_cast<const uint2*>(dO + ((size_t)img * AL + r) * C + c);
  }
  const __nv_bfloat16* vsrc = vt + (size_t)img * C * AL;
  for (int i = t; i < C * (AL / 2); i += 256) {
    const int c = i / (AL / 2), k = (i - c * (AL / 2)) * 2;
    *reinterpret_cast<uint32_t*>(Vts + c * LDV + k) = *reinterpret_cast<const uint32_t*>(vsrc + (size_t)c * AL + k);
  }
  __syncthreads();
  // ---- S = Q K^T / sqrt(C), dP = dO V^T : thread tile q = tq + 16 i, k = tk + 16 j (conflict-free smem rows)
  const float scale = rsqrtf((float)C);
  {
    const int tq = t >> 4, tk = t & 15;
    float sacc[4][4], pacc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) sacc[i][j] = pacc[i][j] = 0.f;
    for (int c = 0; c < C; c += 2) {
      float2 q[4], g[4], kk[4], v0[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        q[i] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(Qs + (tq + 16 * i) * LDH + c));
        g[i] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(Gs + (tq + 16 * i) * LDH + c));
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        kk[j] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(Ks + (tk + 16 * j) * LDH + c));
        v0[j].x = __bfloat162float(Vts[c * LDV + tk + 16 * j]);
        v0[j].y = __bfloat162float(Vts[(c + 1) * LDV + tk + 16 * j]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          sacc[i][j] = fmaf(q[i].x, kk[j].x, fmaf(q[i].y, kk[j].y, sacc[i][j]));
          pacc[i][j] = fmaf(g[i].x, v0[j].x, fmaf(g[i].y, v0[j].y, pacc[i][j]));
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        Ps[(tq + 16 * i) * LDP + tk + 16 * j] = sacc[i][j] * scale;
        Ds[(tq + 16 * i) * LDP + tk + 16 * j] = pacc[i][j];
      }
  }
  __syncthreads();
  // ---- softmax rows: P, then dS = P * (dP - sum_k P dP) / sqrt(C)
  {
    const int warp = t >> 5, lane = t & 31;
    for (int r = warp; r < AL; r += 8) {
      float* prow = Ps + r * LDP;
      float* drow = Ds + r * LDP;
      const float s0 = prow[lane], s1 = prow[lane + 32];
      const float m = warp_max(fmaxf(s0, s1));
      const float e0 = __expf(s0 - m), e1 = __expf(s1 - m);
      const float inv = 1.f / warp_sum(e0 + e1);
      const float p0 = e0 * inv, p1 = e1 * inv;
      const float d0 = drow[lane], d1 = drow[lane + 32];
      const float delta = warp_sum(p0 * d0 + p1 * d1);
      prow[lane] = p0; prow[lane + 32] = p1;
      drow[lane] = p0 * (d0 - delta) * scale; drow[lane + 32] = p1 * (d1 - delta) * scale;
    }
  }
  __syncthreads();
  // ---- dQ = dS K, dK = dS^T Q, dV = P^T dO : work unit = (block of 8 rows, channel pair)
  const int cp_n = C / 2;
  __nv_bfloat16* orow = dqkv + (size_t)img * AL * ld;
  for (int idx = t; idx < 3 * (AL / 8) * cp_n; idx += 256) {
    const int which = idx / ((AL / 8) * cp_n);
    const int rem = idx - which * (AL / 8) * cp_n;
    const int rb = rem / cp_n, c = (rem - rb * cp_n) * 2;
    float2 acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_float2(0.f, 0.f);
    if (which == 0) {            // dQ[q][c] = sum_k dS[q][k] K[k][c]
      for (int k = 0; k < AL; ++k) {
        const float2 kv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(Ks + k * LDH + c));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float w = Ds[(rb * 8 + i) * LDP + k];
          acc[i].x = fmaf(w, kv.x, acc[i].x); acc[i].y = fmaf(w, kv.y, acc[i].y);
        }
      }
    } else {                     // dK[k][c] = sum_q dS[q][k] Q[q][c];  dV[k][c] = sum_q P[q][k] dO[q][c]
      const float* W = which == 1 ? Ds : Ps;
      const __nv_bfloat16* X = which == 1 ? Qs : Gs;
      for (int q = 0; q < AL; ++q) {
        const float2 xv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(X + q * LDH + c));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float w = W[q * LDP + rb * 8 + i];
          acc[i].x = fmaf(w, xv.x, acc[i].x); acc[i].y = fmaf(w, xv.y, acc[i].y);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
      *reinterpret_cast<__nv_bfloat162*>(orow + (size_t)(rb * 8 + i) * ld + which * C + c) = __floats2bfloat162_rn(acc[i].x, acc[i].y);
  }
}

template <typename T>
__global__ void dkv_to_act_kernel(const float* __restrict__ dkv, int C, size_t rows, T* __restrict__ dqkv) {
  const size_t total = rows * 2 * C;
  size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const size_t r = gid / (2 * C);
  const int c = (int)(gid % (2 * C));
  dqkv[r * 3 * C + C + c] = from_f<T>(dkv[gid]);
}

// ---------------------------------------------------------------------------------------------------------------
// small layout kernels (PADDED tensors, 16-byte vectors)
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) upsample2x_bwd_kernel(const T* __restrict__ dy, int H, int W, int C, size_t total, T* __restrict__ dx, int accumulate) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int VEC = VecOf<T>::N;
  const int CV = C / VEC;
  size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;       // over PADDED low-res (row, vec)
  if (gid >= total) return;
  const int cv = (int)(gid % CV);
  size_t r = gid / CV;
  const int W1 = W + 1, P = (H + 1) * W1;
  const int rem = (int)(r % P);
  const size_t img = r / P;
  const int yy = rem / W1, xx = rem - yy * W1;
  float o[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) o[j] = 0.f;
  if (yy > 0 && xx > 0) {
    const int Wo1 = 2 * W + 1;
    const T* b = dy + (img * (size_t)(2 * H + 1) * Wo1) * C + cv * VEC;
#pragma unroll
    for (int dyy = 0; dyy < 2; ++dyy)
#pragma unroll
      for (int dxx = 0; dxx < 2; ++dxx) {
        float v[VEC];
        load_vec(b + ((size_t)(2 * (yy - 1) + dyy + 1) * Wo1 + (2 * (xx - 1) + dxx + 1)) * C, v);
#pragma unroll
        for (int j = 0; j < VEC; ++j) o[j] += v[j];
      }
    if (accumulate) {
      float old[VEC];
      load_vec(dx + gid * VEC, old);
#pragma unroll
      for (int j = 0; j < VEC; ++j) o[j] += old[j];
    }
  } else if (accumulate) {
    return;
  }
  store_vec(dx + gid * VEC, o);
}

// dst (PADDED 2H x 2W) = dy (PADDED H x W) at even pixels, zero elsewhere
template <typename T>
__global__ void __launch_bounds__(256) zero_insert2x_kernel(const T* __restrict__ dy, int H, int W, int C, size_t total, T* __restrict__ dst) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int VEC = VecOf<T>::N;
  const int CV = C / VEC;
  size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;       // over PADDED full-res (row, vec)
  if (gid >= total) return;
  const int cv = (int)(gid % CV);
  size_t r = gid / CV;
  const int Wo1 = 2 * W + 1, Po = (2 * H + 1) * Wo1;
  const int rem = (int)(r % Po);
  const size_t img = r / Po;
  const int yy = rem / Wo1, xx = rem - yy * Wo1;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (yy > 0 && xx > 0 && ((yy - 1) & 1) == 0 && ((xx - 1) & 1) == 0)
    v = *reinterpret_cast<const uint4*>(dy + ((img * (H + 1) + ((yy - 1) / 2 + 1)) * (size_t)(W + 1) + ((xx - 1) / 2 + 1)) * C + cv * VEC);
  *reinterpret_cast<uint4*>(dst + gid * VEC) = v;
}

template <typename T>
__global__ void __launch_bounds__(256) add_inplace_kernel(T* __restrict__ dst, const T* __restrict__ src, size_t nvec) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int VEC = VecOf<T>::N;
  size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= nvec) return;
  float a[VEC], b[VEC];
  load_vec(dst + gid * VEC, a);
  load_vec(src + gid * VEC, b);
#pragma unroll
  for (int j = 0; j < VEC; ++j) a[j] += b[j];
  store_vec(dst + gid * VEC, a);
}

template <typename T>
__global__ void grad8_to_act_kernel(const float* __restrict__ g8, size_t rows, int ld, T* __restrict__ dst) {
  const size_t total = rows * ld;
  size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const size_t r = gid / ld;
  const int c = (int)(gid % ld);
  dst[gid] = from_f<T>(c < 8 ? g8[r * 8 + c] : 0.f);
}

}  // namespace vf

using namespace vf;
#define VF_API extern "C" __attribute__((visibility("default")))

namespace vf {
bool wgrad_tc_supported(const vf_conv_args* a, int dy_ld);
int conv2d_wgrad_tc(const vf_conv_args* a, const void* dy, int dy_ld, float* dwp, cudaStream_t st);
extern int g_force_simt_flag;
}  // namespace vf

VF_API int vf_conv2d_wgrad(const vf_conv_args* a, const void* dy, int dy_ld, float* dwp, vf_stream stream) {
  VF_REQUIRE(a && dy && dwp, "vf_conv2d_wgrad: null args");
  if (!g_force_simt_flag && wgrad_tc_supported(a, dy_ld)) return conv2d_wgrad_tc(a, dy, dy_ld, dwp, as_stream(stream));
  WgradParams p{};
  p.geo = make_geom(a->images, a->H, a->W, a->in_padded, a->out_padded, a->stride == 2);
  int k_total = 0;
  for (int s = 0; s < a->n_seg; ++s) {
    p.src[s] = a->src[s]; p.src_c[s] = a->src_c[s]; p.ksize[s] = a->ksize[s];
    VF_REQUIRE(a->src_c[s] % 32 == 0, "vf_conv2d_wgrad: segment channels %d not a multiple of 32", a->src_c[s]);
    k_total += a->ksize[s] * a->ksize[s] * a->src_c[s];
  }
  p.n_seg = a->n_seg; p.dy = dy; p.dy_ld = dy_ld; p.cout = a->cout; p.cout_pad = a->cout_pad; p.k_total = k_total; p.dwp = dwp;
  const int M = p.geo.rows_total;
  int splits = cdiv(M, 4096);
  if (splits > 512) splits = 512;
  p.rows_per_split = (int)align_up((size_t)cdiv(M, splits), 16);
  splits = cdiv(M, p.rows_per_split);
  dim3 grid(k_total / 32, cdiv(a->cout, 64), splits);
  if (a->dtype == VF_BF16) conv_wgrad_simt_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(p);
  else conv_wgrad_simt_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(p);
  VF_LAUNCH_CHECK();
  return VF_OK;
}

VF_API int vf_pack_conv_weight_t(const float* w_oihw, int cout, int cin, int ksize, int dtype, void* dst, int cin_pad, int k_total,
                                 int k_off, int n_stride, vf_stream stream) {
  VF_REQUIRE(w_oihw && dst && cout > 0 && cin > 0 && (ksize == 1 || ksize == 3) && n_stride >= cout && cin_pad >= cin, "vf_pack_conv_weight_t: bad args");
  VF_REQUIRE(k_off + ksize * ksize * n_stride <= k_total, "vf_pack_conv_weight_t: row overflow");
  const size_t total = (size_t)cin_pad * ksize * ksize * n_stride;
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (dtype == VF_BF16) pack_conv_weight_t_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(w_oihw, cout, cin, ksize * ksize, (__nv_bfloat16*)dst, cin_pad, k_total, k_off, n_stride);
  else pack_conv_weight_t_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(w_oihw, cout, cin, ksize * ksize, (float*)dst, cin_pad, k_total, k_off, n_stride);
  VF_LAUNCH_CHECK();
  return VF_OK;
}

VF_API int vf_unpack_conv_wgrad(const float* dwp, int cout, int cin, int ksize, int k_total, int k_off, float* dw_oihw, int cin_total,
                                int c_off, vf_stream stream) {
  VF_REQUIRE(dwp && dw_oihw && cout > 0 && cin > 0 && c_off >= 0 && c_off + cin <= cin_total, "vf_unpack_conv_wgrad: bad args");
  const size_t total = (size_t)cout * cin * ksize * ksize;
  VF_CUDA(launch_pdl(unpack_conv_wgrad_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, as_stream(stream), dwp, cout, cin, ksize * ksize, k_total, k_off, dw_oihw, cin_total, c_off));
  VF_LAUNCH_CHECK();
  return VF_OK;
}

namespace vf {
// row splits per image of the two GroupNorm backward passes (one grid for both)
static int gn_bwd_splits(int images, int P, int C, int PY) {
  // the reduce pass runs 2 CTAs per SM: whole waves of 296 CTAs, ~6 CTAs per SM in total.  (64x64 x 64 channels, 168 view-images:
  // 7 row splits 147 us, 13 splits 164 us, 26 splits 207 us — every CTA repeats the statistics prologue and the partial-sum epilogue)
  const int want = cdiv(148 * 6, images);
  const int max_splits = P / (PY * 4) > 0 ? P / (PY * 4) : 1;
  int splits = wave_splits(images, want, max_splits, 148 * 2);
  // small layers by the latency model (the ones that still come here: most of them run the one-pass kernel)
  if ((long)P * C <= 1089L * 128 && !getenv("VF_GN_OLD_SPLITS"))
    splits = latency_splits(images, P, kGbUnroll * PY, max_splits, 148 * 2, 3.0, 0.8);   // the reduce pass runs 2 CTAs per SM
  return splits;
}

// Slab (channels per CTA) of the one-pass kernel, 0 if it does not apply: a multiple of the group size and of the vector width
// that divides C; the smallest one with >= 32-byte rows and >= 16 K elements per CTA that fits in shared memory (else the
// largest that fits).  threads = (slab / vec) * row lanes; 512-thread CTAs when only one fits per SM.
int tc_debug_flags();   // k_gemm_tc.cu (vf_debug_flags)
static int gn_fused_slab(int C, int groups, int HW, int dtype, int* threads_out, size_t* smem_out) {
  static const bool off = [] { const char* e = getenv("VF_GN_BWD_FUSED"); return e && e[0] == '0'; }();      // A/B knob
  if (off || (tc_debug_flags() & 0x2000)) return 0;                                                        // test hook: two-pass kernels
  const int vec = dtype == VF_BF16 ? 8 : 4, es = dtype == VF_BF16 ? 2 : 4;
  const int gs = C / groups;
  int unit = gs;
  while (unit % vec) unit += gs;                 // lcm(gs, vec)
  if (C % unit) return 0;
  static const long min_elems = [] { const char* e = getenv("VF_GNF_MIN_ELEMS"); return e ? atol(e) : 16384L; }();      // A/B knobs
  static const size_t limit = [] { const char* e = getenv("VF_GNF_SMEM_KB"); return (size_t)(e ? atol(e) : 90L) * 1024; }();
  auto need = [&](int SC, int threads) { return (size_t)(8 * SC + threads) * 4 + (size_t)threads * vec * 8 + (size_t)2 * HW * SC * es + 16; };
  int best = 0;
  for (int SC = unit; SC <= C; SC += unit) {
    if (C % SC || SC / vec > 512) continue;
    if (need(SC, 256) > limit) break;
    best = SC;
    if (SC * es >= 32 && (long)HW * SC >= min_elems) break;
  }
  if (!best) return 0;
  const int SCV = best / vec;
  // (512-thread CTAs for wide slabs were measured slower: 8x8 x 192 channels 14.6 -> 22.7 us)
  const int want = need(best, 256) > 100 * 1024 ? 512 : 256;
  const int PY = want / SCV > 0 ? want / SCV : 1;
  *threads_out = SCV * PY;
  *smem_out = need(best, SCV * PY);
  return best;
}

int gn_backward_impl(const void* src0, int C0, const float* stats0, int stats0_ld, const void* src1, int C1, const float* stats1,
                     int stats1_ld, int dtype, int images, int H, int W, int groups, const float* gamma, const float* beta, int swish,
                     const void* dy, float* scratch, bool scratch_zeroed, float* dgamma, float* dbeta, void* dx0, int acc0, void* dx1,
                     int acc1, cudaStream_t st, const GnColsum* colsum, const vf_gn_shift* shift) {
  VF_REQUIRE(src0 && stats0 && gamma && beta && dy && scratch && dgamma && dbeta && dx0, "vf_gn_backward: null args");
  if (!src1) C1 = 0;
  VF_REQUIRE(C1 == 0 || (stats1 && dx1), "vf_gn_backward: second source needs stats and dx");
  const int vec = dtype == VF_BF16 ? 8 : 4;
  const int C = C0 + C1;
  VF_REQUIRE(C0 % vec == 0 && C1 % vec == 0 && C % groups == 0 && C / vec <= kGbThreads, "vf_gn_backward: bad channel counts (%d, %d)", C0, C1);
  GnBwdParams p{};
  p.s0 = src0; p.C0 = C0; p.st0 = stats0; p.ld0 = stats0_ld; p.s1 = src1; p.C1 = C1; p.st1 = C1 ? stats1 : nullptr; p.ld1 = stats1_ld;
  p.HW = H * W; p.W1 = W + 1; p.P = (H + 1) * (W + 1); p.groups = groups; p.swish = swish;
  p.gamma = gamma; p.beta = beta; p.dy = dy; p.red = scratch; p.dgamma = dgamma; p.dbeta = dbeta;
  p.dx0 = dx0; p.acc0 = acc0; p.dx1 = dx1; p.acc1 = acc1;
  if (colsum) {
    VF_REQUIRE(C1 == 0 && !acc0, "vf_gn_backward: closed-form column sums need a single source and a fresh dx");
    p.cs = *colsum;
  }
  if (shift) {
    VF_REQUIRE(!shift->emb || shift->img_row, "vf_gn_backward: shift.emb needs img_row");
    p.sh = *shift;
  }
  {
    // one-pass kernel when an (image, slab of whole groups) fits in shared memory
    int threads = 0;
    size_t smem = 0;
    const int SC = gn_fused_slab(C, groups, H * W, dtype, &threads, &smem);
    if (SC > 0) {
      p.img0 = 0;
      dim3 grid(C / SC, images);
      if (dtype == VF_BF16) {
        VF_SET_MAX_SMEM(gn_bwd_fused_kernel<__nv_bfloat16>, 227 * 1024);
        VF_CUDA(launch_pdl(gn_bwd_fused_kernel<__nv_bfloat16>, grid, dim3(threads), smem, st, p, SC));
      } else {
        VF_SET_MAX_SMEM(gn_bwd_fused_kernel<float>, 227 * 1024);
        VF_CUDA(launch_pdl(gn_bwd_fused_kernel<float>, grid, dim3(threads), smem, st, p, SC));
      }
      VF_LAUNCH_CHECK();
      return VF_OK;
    }
  }
  if (!scratch_zeroed) VF_CUDA(cudaMemsetAsync(scratch, 0, (size_t)images * C * 2 * sizeof(float), st));
  const int CV = C / vec, PY = kGbThreads / CV > 0 ? kGbThreads / CV : 1;
  const int threads = CV * PY;
  // (Processing large layers in L2-sized image groups so that the apply pass re-reads x / dy from L2 was measured
  // 25 % slower than whole-batch launches: the smaller grids lose more than the L2 hits win.)
  const int group = images;
  int splits = gn_bwd_splits(group, p.P, C, PY);
  {
    static const int force = [] { const char* e = getenv("VF_GN_BWD_SPLITS"); return e ? atoi(e) : 0; }();      // A/B knob
    if (force > 0) splits = force;
  }
  p.rows_per_cta = cdiv(p.P, splits);
  splits = cdiv(p.P, p.rows_per_cta);
  const size_t smem_r = (size_t)(2 * C + 2 * C * PY) * sizeof(float), smem_a = (size_t)6 * C * sizeof(float);
  VF_REQUIRE(smem_r <= 48 * 1024, "vf_gn_backward: C=%d needs %zu B of shared memory", C, smem_r);
  for (int i0 = 0; i0 < images; i0 += group) {
    const int ni = images - i0 < group ? images - i0 : group;
    p.img0 = i0;
    dim3 grid(splits, ni);
    if (dtype == VF_BF16) {
      VF_CUDA(launch_pdl(gn_bwd_reduce_kernel<__nv_bfloat16>, grid, dim3(threads), smem_r, st, p));
      VF_CUDA(launch_pdl(gn_bwd_apply_kernel<__nv_bfloat16>, grid, dim3(threads), smem_a, st, p));
    } else {
      VF_CUDA(launch_pdl(gn_bwd_reduce_kernel<float>, grid, dim3(threads), smem_r, st, p));
      VF_CUDA(launch_pdl(gn_bwd_apply_kernel<float>, grid, dim3(threads), smem_a, st, p));
    }
  }
  VF_LAUNCH_CHECK();
  return VF_OK;
}
}  // namespace vf

// Test hook (host only): the one-pass GroupNorm backward's plan for C channels in `groups` groups over H x W pixels: channels per
// CTA (0 = the layer takes the two-pass kernels), threads and dynamic shared memory through the out parameters.
VF_API int vf_debug_gn_bwd_slab(int C, int groups, int H, int W, int dtype, int* threads_out, int* smem_out) {
  const int vec = dtype == VF_BF16 ? 8 : 4;
  if (C <= 0 || groups <= 0 || C % groups || C % vec || H <= 0 || W <= 0) return -1;
  int threads = 0;
  size_t smem = 0;
  const int sc = vf::gn_fused_slab(C, groups, H * W, dtype, &threads, &smem);
  if (threads_out) *threads_out = threads;
  if (smem_out) *smem_out = (int)smem;
  return sc;
}

// Test hook (host only): row splits the GroupNorm backward would launch with.
VF_API int vf_debug_gn_bwd_splits(int images, int H, int W, int C, int dtype) {
  const int vec = dtype == VF_BF16 ? 8 : 4;
  if (images <= 0 || H <= 0 || W <= 0 || C <= 0 || C % vec || C / vec > vf::kGbThreads) return -1;
  const int CV = C / vec, PY = vf::kGbThreads / CV > 0 ? vf::kGbThreads / CV : 1;
  return vf::gn_bwd_splits(images, (H + 1) * (W + 1), C, PY);
}

VF_API int vf_gn_backward(const void* src0, int C0, const float* stats0, int stats0_ld, const void* src1, int C1, const float* stats1,
                          int stats1_ld, int dtype, int images, int H, int W, int groups, const float* gamma, const float* beta, int swish,
                          const void* dy, float* scratch, float* dgamma, float* dbeta, void* dx0, int acc0, void* dx1, int acc1,
                          const vf_gn_shift* shift, vf_stream stream) {
  return vf::gn_backward_impl(src0, C0, stats0, stats0_ld, src1, C1, stats1, stats1_ld, dtype, images, H, W, groups, gamma, beta, swish, dy,
                              scratch, false, dgamma, dbeta, dx0, acc0, dx1, acc1, as_stream(stream), nullptr, shift);
}

namespace vf {
bool attention_bwd_tc_supported(int L, int C);
int attention_bwd_tc(const void* qkv, const void* vt, const void* out, const float* lse, const void* d_out, int images, int L, int C,
                     float* scratch, void* dqkv, cudaStream_t st);
}  // namespace vf

VF_API int vf_attention_backward(const void* qk, const void* vt, const void* out, const float* lse, const void* d_out, int dtype, int images,
                                 int L, int C, float* scratch, void* dqkv, vf_stream stream) {
  VF_REQUIRE(qk && d_out && scratch && dqkv && images > 0 && L > 0 && C > 0, "vf_attention_backward: bad args");
  if (dtype == VF_BF16 && !g_force_simt_flag && vt && out && lse && attention_bwd_tc_supported(L, C))
    return attention_bwd_tc(qk, vt, out, lse, d_out, images, L, C, scratch, dqkv, as_stream(stream));
  cudaStream_t st = as_stream(stream);
  if (dtype == VF_BF16 && !g_force_simt_flag && vt && L == AL && C % 64 == 0 && AttnL64Smem::bytes(C) <= 227 * 1024) {
    VF_SET_MAX_SMEM(attn_bwd_l64_kernel, 227 * 1024);
    attn_bwd_l64_kernel<<<images, 256, AttnL64Smem::bytes(C), st>>>((const __nv_bfloat16*)qk, (const __nv_bfloat16*)vt, (const __nv_bfloat16*)d_out,
                                                                    C, (__nv_bfloat16*)dqkv);
    VF_LAUNCH_CHECK();
    return VF_OK;
  }
  const size_t smem = (size_t)BQ * (2 * C + 2 * L) * sizeof(float);
  VF_REQUIRE(smem <= 48 * 1024, "vf_attention_backward: L=%d C=%d too large", L, C);
  VF_CUDA(cudaMemsetAsync(scratch, 0, (size_t)images * L * 2 * C * sizeof(float), st));
  dim3 grid(cdiv(L, BQ), images);
  const size_t rows = (size_t)images * L;
  const unsigned g2 = (unsigned)((rows * 2 * C + 255) / 256);
  if (dtype == VF_BF16) {
    attn_bwd_simt_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>((const __nv_bfloat16*)qk, 3 * C, (const __nv_bfloat16*)vt, (const __nv_bfloat16*)d_out, L, C, (__nv_bfloat16*)dqkv, 3 * C, scratch);
    dkv_to_act_kernel<__nv_bfloat16><<<g2, 256, 0, st>>>(scratch, C, rows, (__nv_bfloat16*)dqkv);
  } else {
    attn_bwd_simt_kernel<float><<<grid, 256, smem, st>>>((const float*)qk, 3 * C, (const float*)vt, (const float*)d_out, L, C, (float*)dqkv, 3 * C, scratch);
    dkv_to_act_kernel<float><<<g2, 256, 0, st>>>(scratch, C, rows, (float*)dqkv);
  }
  VF_LAUNCH_CHECK();
  return VF_OK;
}

VF_API int vf_upsample2x_backward(const void* dy, int dtype, int images, int H, int W, int C, void* dx, int accumulate, vf_stream stream) {
  VF_REQUIRE(dy && dx && images > 0 && H > 0 && W > 0 && C > 0, "vf_upsample2x_backward: bad args");
  const int vec = dtype == VF_BF16 ? 8 : 4;
  VF_REQUIRE(C % vec == 0, "vf_upsample2x_backward: C=%d", C);
  const size_t total = (size_t)images * (H + 1) * (W + 1) * (C / vec);
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (dtype == VF_BF16) VF_CUDA(launch_pdl(upsample2x_bwd_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, as_stream(stream), (const __nv_bfloat16*)dy, H, W, C, total, (__nv_bfloat16*)dx, accumulate));
  else VF_CUDA(launch_pdl(upsample2x_bwd_kernel<float>, dim3(grid), dim3(256), 0, as_stream(stream), (const float*)dy, H, W, C, total, (float*)dx, accumulate));
  VF_LAUNCH_CHECK();
  return VF_OK;
}

VF_API int vf_zero_insert2x(const void* dy, int dtype, int images, int H, int W, int C, void* dst, vf_stream stream) {
  VF_REQUIRE(dy && dst && images > 0 && H > 0 && W > 0 && C > 0, "vf_zero_insert2x: bad args");
  const int vec = dtype == VF_BF16 ? 8 : 4;
  VF_REQUIRE(C % vec == 0, "vf_zero_insert2x: C=%d", C);
  const size_t total = (size_t)images * (2 * H + 1) * (2 * W + 1) * (C / vec);
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (dtype == VF_BF16) VF_CUDA(launch_pdl(zero_insert2x_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, as_stream(stream), (const __nv_bfloat16*)dy, H, W, C, total, (__nv_bfloat16*)dst));
  else VF_CUDA(launch_pdl(zero_insert2x_kernel<float>, dim3(grid), dim3(256), 0, as_stream(stream), (const float*)dy, H, W, C, total, (float*)dst));
  VF_LAUNCH_CHECK();
  return VF_OK;
}

VF_API int vf_add_inplace(void* dst, const void* src, int dtype, size_t n_elems, vf_stream stream) {
  VF_REQUIRE(dst && src, "vf_add_inplace: null args");
  const size_t vec = dtype == VF_BF16 ? 8 : 4;
  VF_REQUIRE(n_elems % vec == 0, "vf_add_inplace: n_elems not a multiple of %zu", vec);
  const size_t nvec = n_elems / vec;
  const unsigned grid = (unsigned)((nvec + 255) / 256);
  if (dtype == VF_BF16) VF_CUDA(launch_pdl(add_inplace_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, as_stream(stream), (__nv_bfloat16*)dst, (const __nv_bfloat16*)src, nvec));
  else VF_CUDA(launch_pdl(add_inplace_kernel<float>, dim3(grid), dim3(256), 0, as_stream(stream), (float*)dst, (const float*)src, nvec));
  VF_LAUNCH_CHECK();
  return VF_OK;
}

VF_API int vf_grad8_to_act(const float* g8, size_t rows, int dtype, int ld, void* dst, vf_stream stream) {
  VF_REQUIRE(g8 && dst && ld >= 8, "vf_grad8_to_act: bad args");
  const size_t total = rows * ld;
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (dtype == VF_BF16) grad8_to_act_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(g8, rows, ld, (__nv_bfloat16*)dst);
  else grad8_to_act_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(g8, rows, ld, (float*)dst);
  VF_LAUNCH_CHECK();
  return VF_OK;
}
