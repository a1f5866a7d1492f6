// CUDA-core implicit-GEMM convolution: the fp32 "reference-precision" mode of vf_conv2d (1e-4 parity bar, no
// TF32/bf16 rounding anywhere) and the on-device cross-check for the tcgen05 kernel.  Same contract as the
// tensor-core path: FLAT / PADDED row orders, K-major weights, up to three K-segments (3x3 main conv + 1x1 res_conv
// over one or two sources) accumulated into one tile, fused bias + embedding + residual epilogue, GroupNorm sums.
// Reference: nn.Conv2d call sites model/unet.py:42,189,198,214,238,255,256.
#include "vf_common.cuh"

namespace vf {

constexpr int SBM = 64, SBN = 64, SBK = 16;

struct SimtConvParams {
  RowGeom geo;
  const void* src[3];
  int src_c[3];
  int ksize[3];
  int n_seg;
  const void* weight;
  int k_total;
  int cout, cout_pad;
  const float* bias;
  const float* emb;
  const int* img_row;
  int emb_ld;
  const void* residual;
  void* out;
  int out_ld;
  int qkv_split;
  void* out_vt;
  float* stats;
};

__device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
  float4 t = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load4(const __nv_bfloat16* p, float (&v)[4]) {
  uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
  float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

template <typename TA, typename TO>
__global__ void __launch_bounds__(256) conv_simt_kernel(const SimtConvParams p) {
  __shared__ float As[SBK][SBM + 4];
  __shared__ float Bs[SBK][SBN + 4];
  const int M = p.geo.rows_total;
  const int m0 = blockIdx.x * SBM, n0 = blockIdx.y * SBN;
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;          // 16x16 threads, each 4 rows x 4 cols
  const int lrow = tid / 4, lk = (tid % 4) * 4;    // loader roles
  const int lm = m0 + lrow;
  const int ln = n0 + lrow;
  const TA* wt = reinterpret_cast<const TA*>(p.weight);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  int koff = 0;
  for (int s = 0; s < p.n_seg; ++s) {
    const int C = p.src_c[s];
    const int ks = p.ksize[s];
    const TA* src = reinterpret_cast<const TA*>(p.src[s]);
    for (int tap = 0; tap < ks * ks; ++tap) {
      // PADDED rows: the 3x3 neighbourhood is a constant row offset; padding rows hold zeros
      const int off = ks == 3 ? (tap / 3 - 1) * p.geo.W1 + (tap % 3 - 1) : 0;
      const long row = (long)lm + off;
      const bool inb = lm < M && row >= 0 && row < M;
      const TA* arow = src + (inb ? row : 0) * C;
      for (int c0 = 0; c0 < C; c0 += SBK) {
        float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
        if (inb) load4(arow + c0 + lk, a);
        if (ln < p.cout_pad) load4(wt + (size_t)ln * p.k_total + koff + tap * C + c0 + lk, b);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) { As[lk + j][lrow] = a[j]; Bs[lk + j][lrow] = b[j]; }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < SBK; ++k) {
          float av[4], bv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) { av[i] = As[k][ty * 4 + i]; bv[i] = Bs[k][tx * 4 + i]; }
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] += av[i] * bv[j];
        }
      }
    }
    koff += ks * ks * C;
  }

  // epilogue
  TO* out = reinterpret_cast<TO*>(p.out);
  const TA* res = reinterpret_cast<const TA*>(p.residual);
  TA* vt = reinterpret_cast<TA*>(p.out_vt);
  const int HWo = p.geo.stride2 ? p.geo.HW / 4 : p.geo.HW;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    const RowInfo ri = decode_row(p.geo, m);
    if (!ri.valid) continue;
    const float* erow = p.emb ? p.emb + (size_t)__ldg(p.img_row + ri.img) * p.emb_ld : nullptr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.cout) continue;
      float v = acc[i][j];
      if (p.bias) v += __ldg(p.bias + n);
      if (erow) v += __ldg(erow + n);
      if (res) v += to_f(res[(size_t)ri.out_row * p.cout + n]);
      if (p.qkv_split > 0 && n >= 2 * p.qkv_split)
        vt[((size_t)ri.img * p.qkv_split + (n - 2 * p.qkv_split)) * HWo + ri.pix] = from_f<TA>(v);
      else
        out[(size_t)ri.out_row * p.out_ld + n] = from_f<TO>(v);
      if (p.stats) {
        const float x = to_f(from_f<TO>(v));
        atomicAdd(p.stats + ((size_t)ri.img * p.cout + n) * 2, x);
        atomicAdd(p.stats + ((size_t)ri.img * p.cout + n) * 2 + 1, x * x);
      }
    }
  }
}

int conv2d_simt(const vf_conv_args* a, cudaStream_t st) {
  SimtConvParams p{};
  p.geo = make_geom(a->images, a->H, a->W, a->in_padded, a->out_padded, a->stride == 2);
  int k_total = 0;
  for (int s = 0; s < a->n_seg; ++s) {
    p.src[s] = a->src[s]; p.src_c[s] = a->src_c[s]; p.ksize[s] = a->ksize[s];
    VF_REQUIRE(a->src_c[s] % SBK == 0, "vf_conv2d(simt): segment channels %d not a multiple of %d", a->src_c[s], SBK);
    VF_REQUIRE(a->ksize[s] == 1 || a->in_padded, "vf_conv2d(simt): 3x3 needs PADDED sources");
    k_total += a->ksize[s] * a->ksize[s] * a->src_c[s];
  }
  p.n_seg = a->n_seg;
  p.weight = a->weight; p.k_total = k_total; p.cout = a->cout; p.cout_pad = a->cout_pad;
  p.bias = a->bias; p.emb = a->emb; p.img_row = a->img_row; p.emb_ld = a->emb_ld; p.residual = a->residual;
  p.out = a->out; p.out_ld = a->out_ld; p.qkv_split = a->qkv_split; p.out_vt = a->out_vt; p.stats = a->stats;
  dim3 grid(cdiv(p.geo.rows_total, SBM), cdiv(a->cout, SBN));
  if (a->dtype == VF_F32) {
    VF_REQUIRE(a->out_dtype == VF_F32, "vf_conv2d(simt): fp32 activations need fp32 output");
    conv_simt_kernel<float, float><<<grid, 256, 0, st>>>(p);
  } else if (a->out_dtype == VF_F32) {
    conv_simt_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>(p);
  } else {
    conv_simt_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>(p);
  }
  VF_LAUNCH_CHECK();
  return VF_OK;
}

}  // namespace vf
