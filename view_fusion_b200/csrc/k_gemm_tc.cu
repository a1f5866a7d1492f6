// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (bf16 operands, fp32 accumulation in TMEM).
// Reference call sites: every nn.Conv2d of model/unet.py (:42, :189, :198, :214, :238, :255, :256).
//
//   D[m, n] = sum_seg sum_tap sum_c A_seg[pix(m, tap), c] * Wt[n, koff(seg) + tap*C_seg + c]      (fp32, in TMEM)
//   out     = D + bias[n] + emb[img_row[img(m)], n] + residual[m, n]                              (bf16 or fp32)
//
// * M tile = 128 output pixels = one UMMA_M=128 accumulator (TMEM lane == pixel).  A tile is one TMA box
//   (64 channels x box_w x box_h x box_n) of the NHWC activation: the 3x3 halo and the zero padding come from the
//   box start coordinate (x0+kw-1, y0+kh-1) and TMA out-of-bounds zero fill, so there is no im2col buffer and no
//   predication in the loader.  Stride-2 convolutions use a 5-D view (2C, W/2, 2, H/2, N) of the same tensor in
//   which the row/column parity is a coordinate.
// * N tile = block_n (<= 256) output channels = one UMMA_N; weights are K-major [Cout][K] rows.
// * K is walked in 64-channel steps (one 128-byte swizzle row); up to three K segments are accumulated into the
//   same TMEM tile (3x3 conv over h  +  1x1 res_conv over x and the skip tensor), which fuses the ResnetBlock's
//   residual projection (unet.py:245) and the decoder's torch.cat (unet.py:134) into the GEMM.
// * Warp roles: warps 0-3 epilogue (TMEM -> registers -> global), warp 4 TMA producer, warp 5 MMA issuer + TMEM
//   allocator.  smem ring of `stages` x (A 16 KB + B block_n*128 B); ~100 KB so that two CTAs share an SM and one
//   CTA's epilogue overlaps the other's main loop.
#include <cuda.h>

#include <mutex>

#include "tc_ptx.cuh"
#include "vf_common.cuh"

namespace vf {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;
constexpr int TC_THREADS = 192;
constexpr int TC_MAX_BOXN = 16;

struct TcSeg {
  int ntaps;     // 1 or 9
  int nchunks;   // C / 64
  int C;
  int mode;      // 0: stride-1 (4-D map), 2: stride-2 3x3 (5-D map)
  int koff;      // first weight column of the segment
};

struct TcParams {
  int M, HW, H, W;
  int box_w, box_h, box_n;
  int block_n, stages, tmem_cols;
  int n_seg;
  TcSeg seg[3];
  int num_k_iters;
  uint32_t idesc;
  const float* bias;
  const float* emb;
  const int* img_row;
  int emb_ld;
  const __nv_bfloat16* residual;
  void* out;
  int out_f32;
  int out_ld;
  int cout;
  int qkv_split;
  __nv_bfloat16* out_vt;
  int images;
  float* stats;   // [images, cout, 2] GroupNorm partial sums of the stored output, or null
};

__global__ void __launch_bounds__(TC_THREADS) conv_tc_kernel(const __grid_constant__ CUtensorMap mapA0,
                                                             const __grid_constant__ CUtensorMap mapA1,
                                                             const __grid_constant__ CUtensorMap mapA2,
                                                             const __grid_constant__ CUtensorMap mapB,
                                                             const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;          // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* gbase = smem_raw + (base - raw);
  // header (first 1 KB): barriers, tmem pointer; then bias/emb table; then the ring
  const uint32_t bar_full = base;                         // stages x 8 B
  const uint32_t bar_empty = base + 64;                   // stages x 8 B
  const uint32_t bar_tmem = base + 128;                   // 8 B
  const uint32_t tmem_slot = base + 136;                  // 4 B
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gbase + 136);
  float* sm_bias = reinterpret_cast<float*>(gbase + 1024);                 // [box_n][block_n]
  const uint32_t bias_bytes = ((uint32_t)(p.box_n * p.block_n * 4) + 1023u) & ~1023u;
  const uint32_t ring = base + 1024 + bias_bytes;
  const uint32_t a_bytes = TC_BM * TC_BK * 2;
  const uint32_t b_bytes = (uint32_t)p.block_n * TC_BK * 2;
  const uint32_t stage_bytes = a_bytes + b_bytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tile_m = blockIdx.x;
  const int n0 = blockIdx.y * p.block_n;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(bar_full + 8 * s, 1);
      ptx::mbar_init(bar_empty + 8 * s, 1);
    }
    ptx::mbar_init(bar_tmem, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 4 && lane == 0) {
    ptx::prefetch_tmap(&mapA0);
    ptx::prefetch_tmap(&mapB);
    if (p.n_seg > 1) ptx::prefetch_tmap(&mapA1);
    if (p.n_seg > 2) ptx::prefetch_tmap(&mapA2);
  }
  if (warp == 5) {
    ptx::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_d = *tmem_slot_ptr;

  // tile origin in (image, row) space; box_w == W always
  int img0, y0;
  if (p.box_n == 1) {
    const int tiles_per_img = p.HW / TC_BM;
    img0 = tile_m / tiles_per_img;
    y0 = (tile_m % tiles_per_img) * p.box_h;
  } else {
    img0 = tile_m * p.box_n;
    y0 = 0;
  }

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int it = 0;
      for (int s = 0; s < p.n_seg; ++s) {
        const TcSeg sg = p.seg[s];
        const CUtensorMap* mA = s == 0 ? &mapA0 : (s == 1 ? &mapA1 : &mapA2);
        for (int tap = 0; tap < sg.ntaps; ++tap) {
          const int kh = sg.ntaps == 9 ? tap / 3 : 1, kw = sg.ntaps == 9 ? tap % 3 : 1;
          for (int ch = 0; ch < sg.nchunks; ++ch, ++it) {
            const int stage = it % p.stages;
            const uint32_t phase = (uint32_t)(it / p.stages) & 1u;
            ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
            const uint32_t fb = bar_full + 8 * stage;
            ptx::mbar_arrive_expect_tx(fb, stage_bytes);
            const uint32_t sa = ring + stage * stage_bytes;
            const uint32_t sb = sa + a_bytes;
            if (sg.mode == 0) {
              ptx::tma_load_4d(sa, mA, fb, ch * TC_BK, kw - 1, y0 + kh - 1, img0);
            } else {
              // input pixel (2y+kh-1, 2x+kw-1): parity = (k != 1), half-index offset = (k == 0 ? -1 : 0)
              const int wp = kw != 1, hp = kh != 1;
              ptx::tma_load_5d(sa, mA, fb, wp * sg.C + ch * TC_BK, kw == 0 ? -1 : 0, hp, y0 + (kh == 0 ? -1 : 0), img0);
            }
            ptx::tma_load_2d(sb, &mapB, fb, sg.koff + tap * sg.C + ch * TC_BK, n0);
          }
        }
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      for (int it = 0; it < p.num_k_iters; ++it) {
        const int stage = it % p.stages;
        const uint32_t phase = (uint32_t)(it / p.stages) & 1u;
        ptx::mbar_wait(bar_full + 8 * stage, phase);
        ptx::tc_fence_after();
        const uint32_t sa = ring + stage * stage_bytes;
        const uint32_t sb = sa + a_bytes;
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k) {
          const uint64_t ad = ptx::make_smem_desc(sa + k * 32, 16, 1024);
          const uint64_t bd = ptx::make_smem_desc(sb + k * 32, 16, 1024);
          ptx::umma_f16(tmem_d, ad, bd, p.idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        ptx::umma_commit(bar_empty + 8 * stage);   // smem slot is free once these MMAs have read it
      }
      ptx::umma_commit(bar_tmem);                  // accumulator complete
    }
  } else {
    // ===================== epilogue (warps 0..3 == TMEM lane quarters 0..3) =====================
    // bias + embedding table for this tile, built while the main loop runs
    for (int i = threadIdx.x; i < p.box_n * p.block_n; i += 128) {
      const int li = i / p.block_n, n = n0 + i % p.block_n;
      const int img = img0 + li;
      float v = 0.f;
      if (n < p.cout) {
        if (p.bias) v += __ldg(p.bias + n);
        if (p.emb && img < p.images) v += __ldg(p.emb + (size_t)__ldg(p.img_row + img) * p.emb_ld + n);
      }
      sm_bias[i] = v;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");

    const int r = warp * 32 + lane;                 // row in tile == TMEM lane
    const int m = tile_m * TC_BM + r;
    const bool valid = m < p.M;
    const int img = valid ? m / p.HW : 0;
    const int pix = valid ? m % p.HW : 0;
    const float* brow = sm_bias + (p.box_n == 1 ? 0 : (r / p.HW)) * p.block_n;

    ptx::mbar_wait(bar_tmem, 0);
    ptx::tc_fence_after();
    const uint32_t trow = tmem_d + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < p.block_n; c0 += 16) {
      uint32_t rr[16];
      ptx::tmem_ld16(trow + (uint32_t)c0, rr);
      ptx::tmem_ld_wait();
      const int n = n0 + c0;
      const bool act = valid && (n < p.cout || p.out_f32);   // structured: the warp reconverges before the next tcgen05.ld
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = act ? __uint_as_float(rr[j]) + brow[c0 + j] : 0.f;
      if (act && p.residual) {
        const __nv_bfloat16* rp = p.residual + (size_t)m * p.cout + n;
        float r0[8], r1[8];
        load_vec(rp, r0);
        load_vec(rp + 8, r1);
#pragma unroll
        for (int j = 0; j < 8; ++j) { v[j] += r0[j]; v[8 + j] += r1[j]; }
      }
      if (!p.out_f32) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __bfloat162float(__float2bfloat16_rn(v[j]));   // what is stored (and normalised later)
      }
      if (act) {
        if (p.out_f32) {
          float* op = reinterpret_cast<float*>(p.out) + (size_t)m * p.out_ld + n;
          const int cnt = min(16, p.out_ld - n);     // out_ld is the padded channel count of the fp32 output
          for (int j = 0; j + 4 <= cnt; j += 4) *reinterpret_cast<float4*>(op + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else if (p.qkv_split > 0 && n >= 2 * p.qkv_split) {
          __nv_bfloat16* vp = p.out_vt + ((size_t)img * p.qkv_split + (n - 2 * p.qkv_split)) * p.HW + pix;
#pragma unroll
          for (int j = 0; j < 16; ++j) vp[(size_t)j * p.HW] = __float2bfloat16_rn(v[j]);
        } else {
          __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)m * p.out_ld + n;
          float lo[8], hi[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) { lo[j] = v[j]; hi[j] = v[8 + j]; }
          store_vec(op, lo);
          store_vec(op + 8, hi);
        }
      }
      if (p.stats && n < p.cout) {
        // GroupNorm partial sums of the stored values, column-reduced over the warp's 32 pixels (16 + 16 shuffles).
        // HW % 32 == 0 (checked on the host), so a warp never straddles two images; invalid rows contribute zeros.
        float sq[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) sq[j] = v[j] * v[j];
        const float cs = warp_colsum16(v, lane), cq = warp_colsum16(sq, lane);
        const int wimg = __shfl_sync(0xffffffffu, img, 0);
        const int wvalid = __shfl_sync(0xffffffffu, (int)valid, 0);
        if (wvalid && (lane & 1) == 0) {
          float* sp = p.stats + ((size_t)wimg * p.cout + n + warp_col16(lane)) * 2;
          atomicAdd(sp, cs);
          atomicAdd(sp + 1, cq);
        }
      }
    }
    ptx::tc_fence_before();
  }
  __syncthreads();
  if (warp == 5) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_d, (uint32_t)p.tmem_cols);
  }
}

// ---- host side --------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

int encode_bf16_map(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return VF_ERR_CUDA; }
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu,%llu,%llu] box [%u,%u,%u,%u,%u]", (int)r,
              rank, (unsigned long long)gd[0], (unsigned long long)(rank > 1 ? gd[1] : 0),
              (unsigned long long)(rank > 2 ? gd[2] : 0), (unsigned long long)(rank > 3 ? gd[3] : 0),
              (unsigned long long)(rank > 4 ? gd[4] : 0), bx[0], rank > 1 ? bx[1] : 0, rank > 2 ? bx[2] : 0,
              rank > 3 ? bx[3] : 0, rank > 4 ? bx[4] : 0);
    return VF_ERR_CUDA;
  }
  return VF_OK;
}

static int pick_block_n(int cout_pad) {
  for (int t = 1; t <= 16; ++t)
    if (cout_pad % t == 0 && cout_pad / t <= 256 && (cout_pad / t) % 16 == 0) return cout_pad / t;
  return 0;
}

int conv2d_tc(const vf_conv_args* a, cudaStream_t st) {
  VF_REQUIRE(a->dtype == VF_BF16, "vf_conv2d(tc): bf16 activations only");
  const int H = a->H, W = a->W;
  VF_REQUIRE(W <= TC_BM && (W & (W - 1)) == 0 && (H & (H - 1)) == 0, "vf_conv2d(tc): H=%d W=%d must be powers of two <= 128", H, W);
  TcParams p{};
  p.H = H; p.W = W; p.HW = H * W; p.images = a->images; p.M = a->images * H * W;
  p.box_w = W;
  p.box_h = (TC_BM / W) < H ? (TC_BM / W) : H;
  p.box_n = TC_BM / (p.box_w * p.box_h);
  VF_REQUIRE(p.box_n <= TC_MAX_BOXN, "vf_conv2d(tc): feature map %dx%d too small", H, W);
  VF_REQUIRE(a->cout_pad % 16 == 0, "vf_conv2d(tc): cout_pad=%d not a multiple of 16", a->cout_pad);
  p.block_n = pick_block_n(a->cout_pad);
  VF_REQUIRE(p.block_n > 0, "vf_conv2d(tc): no N tiling for cout_pad=%d", a->cout_pad);
  p.tmem_cols = 32;
  while (p.tmem_cols < p.block_n) p.tmem_cols *= 2;
  p.idesc = ptx::make_idesc_bf16(TC_BM, p.block_n, 0, 0);

  CUtensorMap maps[3];
  int k_total = 0;
  p.n_seg = a->n_seg;
  p.num_k_iters = 0;
  for (int s = 0; s < a->n_seg; ++s) {
    const int C = a->src_c[s];
    VF_REQUIRE(C % TC_BK == 0, "vf_conv2d(tc): segment %d channels %d not a multiple of 64", s, C);
    VF_REQUIRE(a->ksize[s] == 1 || a->ksize[s] == 3, "vf_conv2d(tc): ksize must be 1 or 3");
    const int stride = s == 0 ? a->stride : 1;
    VF_REQUIRE(stride == 1 || (stride == 2 && a->ksize[s] == 3), "vf_conv2d(tc): stride 2 needs ksize 3");
    TcSeg& sg = p.seg[s];
    sg.ntaps = a->ksize[s] * a->ksize[s];
    sg.nchunks = C / TC_BK;
    sg.C = C;
    sg.mode = stride == 2 ? 2 : 0;
    sg.koff = k_total;
    k_total += sg.ntaps * C;
    p.num_k_iters += sg.ntaps * sg.nchunks;
    if (stride == 1) {
      const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)a->images};
      const uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
      const uint32_t box[4] = {TC_BK, (uint32_t)p.box_w, (uint32_t)p.box_h, (uint32_t)p.box_n};
      int rc = encode_bf16_map(&maps[s], a->src[s], 4, dims, strides, box);
      if (rc) return rc;
    } else {
      const int Win = 2 * W, Hin = 2 * H;
      const uint64_t dims[5] = {(uint64_t)2 * C, (uint64_t)W, 2, (uint64_t)H, (uint64_t)a->images};
      const uint64_t strides[4] = {(uint64_t)2 * C * 2, (uint64_t)Win * C * 2, (uint64_t)2 * Win * C * 2,
                                   (uint64_t)Hin * Win * C * 2};
      const uint32_t box[5] = {TC_BK, (uint32_t)p.box_w, 1, (uint32_t)p.box_h, (uint32_t)p.box_n};
      int rc = encode_bf16_map(&maps[s], a->src[s], 5, dims, strides, box);
      if (rc) return rc;
    }
  }
  for (int s = a->n_seg; s < 3; ++s) maps[s] = maps[0];
  CUtensorMap mapB;
  {
    const uint64_t dims[2] = {(uint64_t)k_total, (uint64_t)a->cout_pad};
    const uint64_t strides[1] = {(uint64_t)k_total * 2};
    const uint32_t box[2] = {TC_BK, (uint32_t)p.block_n};
    int rc = encode_bf16_map(&mapB, a->weight, 2, dims, strides, box);
    if (rc) return rc;
  }
  p.bias = a->bias; p.emb = a->emb; p.img_row = a->img_row; p.emb_ld = a->emb_ld;
  VF_REQUIRE(!a->emb || a->img_row, "vf_conv2d(tc): emb needs img_row");
  p.residual = reinterpret_cast<const __nv_bfloat16*>(a->residual);
  p.out = a->out; p.out_f32 = a->out_dtype == VF_F32; p.out_ld = a->out_ld; p.cout = a->cout;
  p.qkv_split = a->qkv_split; p.out_vt = reinterpret_cast<__nv_bfloat16*>(a->out_vt);
  p.stats = a->stats;
  VF_REQUIRE(!a->stats || ((H * W) % 32 == 0 && !p.out_f32 && !a->qkv_split), "vf_conv2d(tc): fused statistics need H*W %% 32 == 0 and a plain bf16 output");
  VF_REQUIRE(!p.out_f32 || (a->out_ld % 4 == 0 && a->out_ld <= a->cout_pad && !a->residual), "vf_conv2d(tc): bad fp32 output layout");
  VF_REQUIRE(p.out_f32 || (a->cout % 16 == 0 && a->out_ld % 8 == 0), "vf_conv2d(tc): bf16 output needs cout %% 16 == 0");
  VF_REQUIRE(!a->qkv_split || (a->qkv_split % 16 == 0 && a->out_vt), "vf_conv2d(tc): bad qkv split");

  const uint32_t stage_bytes = TC_BM * TC_BK * 2 + p.block_n * TC_BK * 2;
  const uint32_t bias_bytes = ((uint32_t)(p.box_n * p.block_n * 4) + 1023u) & ~1023u;
  const uint32_t fixed = 1024 /*align slack*/ + 1024 /*header*/ + bias_bytes;
  const uint32_t budget = 110 * 1024;               // two CTAs per SM
  int stages = (int)((budget - fixed) / stage_bytes);
  if (stages < 2) stages = 2;
  if (stages > 8) stages = 8;
  if (stages > p.num_k_iters) stages = p.num_k_iters < 1 ? 1 : p.num_k_iters;
  p.stages = stages;
  const size_t smem = fixed + (size_t)stages * stage_bytes;
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    attr_err = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  });
  VF_CUDA(attr_err);
  dim3 grid(cdiv(p.M, TC_BM), a->cout_pad / p.block_n);
  conv_tc_kernel<<<grid, TC_THREADS, smem, st>>>(maps[0], maps[1], maps[2], mapB, p);
  VF_LAUNCH_CHECK();
  return VF_OK;
}

}  // namespace vf
