// Fused single-head self-attention on tcgen05 (reference: model/unet.py:267-274).
//   per CTA: 128 query rows (one half of a 256-token image, or 128/L whole images when L < 128)
//   S = Q K^T  -> TMEM (fp32, 128 x nkeys)          tcgen05.mma, A = Q (K-major), B = K (K-major), TMA-fed
//   P = exp((S - rowmax) / sqrt(C)) -> bf16, written by the softmax warps straight into a 128B-swizzled smem tile
//   O = P V    -> TMEM (fp32, 128 x C)              A = P (K-major), B = V^T (K-major; the qkv GEMM epilogue writes
//                                                   V transposed so no MN-major operand is needed)
//   out = O / rowsum -> bf16
// The L x L score matrix never touches HBM (the reference materialises it in fp32, unet.py:267-272).  When a tile
// holds several images (L = 64: two) the cross-image blocks of S are masked to -inf, i.e. P is block-diagonal.
// smem: phase A [Q | K], phase B [P | V^T] share the same 160 KB; TMEM: nkeys + C <= 512 columns.
#include <cuda.h>

#include <mutex>

#include "tc_ptx.cuh"
#include "vf_common.cuh"

namespace vf {

int encode_bf16_map(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box);

struct AttnTcParams {
  int L, C, images, M;
  int nkeys;          // keys per tile: L if L >= 128 else 128
  int imgs_per_tile;  // 1 or 128/L
  int tiles_per_img;  // L/128 or 1
  int kchunks;        // C / 64
  int pchunks;        // nkeys / 64
  int n_half;         // UMMA_N of the PV product (C or C/2)
  int n_parts;        // 1 or 2
  uint32_t off_k;     // byte offset of the K / V^T region
  uint32_t idesc_s, idesc_o;
  int v_mn;           // V is read row-major from the qkv tensor (MN-major B operand) instead of from a transposed copy
  int part_n[2], part_c0[2];   // v_mn: UMMA_N and first channel of each PV part (multiples of 64)
  uint32_t idesc_o_part[2];
  int tmem_cols;
  float scale_log2;   // log2(e) / sqrt(C)
  __nv_bfloat16* out;
  float* lse;          // optional [images*L]: log2 of the softmax denominator in the scaled-log2 domain (training)
};

constexpr int ATT_SM_WARPS = 8;      // softmax / epilogue warps: two per TMEM lane quarter, each owns half of the columns
constexpr int ATT_THREADS = 32 * (ATT_SM_WARPS + 1);
constexpr int ATT_HDR = 3072;        // barriers + the row max / row sum exchange between the two column halves

__global__ void __launch_bounds__(ATT_THREADS, 1) attn_tc_kernel(const __grid_constant__ CUtensorMap mapQK,
                                                                 const __grid_constant__ CUtensorMap mapVT,
                                                                 const AttnTcParams p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const uint32_t bar_qk = base, bar_s = base + 8, bar_v = base + 16, bar_p = base + 24, bar_o = base + 32;
  const uint32_t tmem_slot = base + 40;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gbase + 40);
  const uint32_t regQ = base + ATT_HDR;         // Q (phase A) / P (phase B)
  const uint32_t regK = regQ + p.off_k;         // K (phase A) / V^T (phase B)
  uint8_t* gP = gbase + ATT_HDR;
  float* x_max = reinterpret_cast<float*>(gbase + 1024);       // [2][128]
  float* x_sum = x_max + 256;                                  // [2][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  const int row0 = tile * 128;                                      // first query row in the [images*L] token axis
  const int img0 = p.imgs_per_tile > 1 ? tile * p.imgs_per_tile : tile / p.tiles_per_img;
  const int key_row0 = p.imgs_per_tile > 1 ? row0 : img0 * p.L;     // first key row

  if (threadIdx.x == 0) {
    ptx::mbar_init(bar_qk, 1);
    ptx::mbar_init(bar_s, 1);
    ptx::mbar_init(bar_v, 1);
    ptx::mbar_init(bar_p, 32 * ATT_SM_WARPS);
    ptx::mbar_init(bar_o, 1);
    ptx::fence_barrier_init();
  }
  if (warp == ATT_SM_WARPS) {
    if (lane == 0) { ptx::prefetch_tmap(&mapQK); ptx::prefetch_tmap(&mapVT); }
    ptx::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_s = *tmem_slot_ptr;
  const uint32_t tmem_o = tmem_s + (uint32_t)p.nkeys;
  pdl_wait();

  const uint32_t q_chunk = 128 * 128;                       // 128 rows x 128 B
  const uint32_t k_chunk = (uint32_t)p.nkeys * 128;
  const uint32_t v_chunk = (uint32_t)p.C * 128;             // C rows (channels) x 64 keys

  if (warp == ATT_SM_WARPS) {
    if (ptx::elect_one()) {
      // ---- phase A: Q, K -> smem; S = Q K^T
      ptx::mbar_arrive_expect_tx(bar_qk, (uint32_t)p.kchunks * (q_chunk + k_chunk));
      for (int c = 0; c < p.kchunks; ++c) {
        ptx::tma_load_2d(regQ + c * q_chunk, &mapQK, bar_qk, c * 64, row0);
        for (int h = 0; h < p.nkeys / 128; ++h)   // K tile rows are loaded as 128-row boxes, contiguous in smem
          ptx::tma_load_2d(regK + c * k_chunk + h * q_chunk, &mapQK, bar_qk, p.C + c * 64, key_row0 + h * 128);
      }
      ptx::mbar_wait(bar_qk, 0);
      ptx::tc_fence_after();
      for (int c = 0; c < p.kchunks; ++c)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t ad = ptx::make_smem_desc(regQ + c * q_chunk + k * 32, 16, 1024);
          const uint64_t bd = ptx::make_smem_desc(regK + c * k_chunk + k * 32, 16, 1024);
          ptx::umma_f16(tmem_s, ad, bd, p.idesc_s, (c > 0 || k > 0) ? 1u : 0u);
        }
      ptx::umma_commit(bar_s);
      // ---- phase B: V^T -> the K region (free once S is complete)
      ptx::mbar_wait(bar_s, 0);
      if (p.v_mn) {
        // V rows (keys) x 64-channel atoms straight from the qkv tensor: atom a = [nkeys rows][128 B], 128B-swizzled, i.e. the
        // canonical MN-major layout with SBO = 1024 (8 keys) and LBO = atom stride
        ptx::mbar_arrive_expect_tx(bar_v, (uint32_t)p.kchunks * k_chunk);
        for (int a = 0; a < p.kchunks; ++a)
          ptx::tma_load_2d(regK + a * k_chunk, &mapVT, bar_v, 2 * p.C + a * 64, key_row0);
        ptx::mbar_wait(bar_v, 0);
        ptx::mbar_wait(bar_p, 0);
        ptx::tc_fence_after();
        for (int part = 0; part < p.n_parts; ++part)
          for (int c = 0; c < p.pchunks; ++c)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t ad = ptx::make_smem_desc(regQ + c * q_chunk + k * 32, 16, 1024);
              const uint64_t bd = ptx::make_smem_desc(regK + (uint32_t)(p.part_c0[part] / 64) * k_chunk + (uint32_t)(c * 64 + k * 16) * 128u,
                                                      k_chunk, 1024);
              ptx::umma_f16(tmem_o + (uint32_t)p.part_c0[part], ad, bd, p.idesc_o_part[part], (c > 0 || k > 0) ? 1u : 0u);
            }
        ptx::umma_commit(bar_o);
      } else {
      ptx::mbar_arrive_expect_tx(bar_v, (uint32_t)p.pchunks * v_chunk);
      for (int c = 0; c < p.pchunks; ++c) {
        // chunk c = 64 keys: keys [64c, 64c+64) of image img0 (L >= 128) or all 64 keys of image img0 + c (L == 64)
        const int kcol = p.imgs_per_tile > 1 ? 0 : c * 64;
        const int vrow = (p.imgs_per_tile > 1 ? img0 + c : img0) * p.C;
        for (int part = 0; part < p.n_parts; ++part)
          ptx::tma_load_2d(regK + c * v_chunk + part * p.n_half * 128, &mapVT, bar_v, kcol, vrow + part * p.n_half);
      }
      ptx::mbar_wait(bar_v, 0);
      ptx::mbar_wait(bar_p, 0);
      ptx::tc_fence_after();
      for (int part = 0; part < p.n_parts; ++part)
        for (int c = 0; c < p.pchunks; ++c)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = ptx::make_smem_desc(regQ + c * q_chunk + k * 32, 16, 1024);
            const uint64_t bd = ptx::make_smem_desc(regK + c * v_chunk + part * p.n_half * 128 + k * 32, 16, 1024);
            ptx::umma_f16(tmem_o + (uint32_t)(part * p.n_half), ad, bd, p.idesc_o, (c > 0 || k > 0) ? 1u : 0u);
          }
      ptx::umma_commit(bar_o);
      }
    }
  } else {
    // ---- softmax + epilogue: lane quarter q = warp & 3 (TMEM lanes 32q..32q+31 = query rows), column half h = warp >> 2.
    // Two warps per scheduler hide each other's TMEM / MUFU latencies; TMEM is read 64 columns per round trip.
    const int q = warp & 3, h = warp >> 2;
    const int r = q * 32 + lane;
    const int m = row0 + r;
    const bool valid = m < p.M;
    const int my_img_local = p.imgs_per_tile > 1 ? r / p.L : 0;
    const uint32_t trow = (uint32_t)(q * 32) << 16;
    const int kh = p.nkeys / 2, k_lo = h * kh;                       // this warp's key columns (a multiple of 64)
    ptx::mbar_wait(bar_s, 0);
    ptx::tc_fence_after();
    float mx = -INFINITY;
    for (int c0 = k_lo; c0 < k_lo + kh; c0 += 64) {
      uint32_t rr[4][16];
#pragma unroll
      for (int b = 0; b < 4; ++b) ptx::tmem_ld16(tmem_s + trow + (uint32_t)(c0 + 16 * b), rr[b]);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const bool same = p.imgs_per_tile == 1 || ((c0 + 16 * b) / p.L) == my_img_local;   // L is a multiple of 16
        if (same) {
#pragma unroll
          for (int j = 0; j < 16; ++j) mx = fmaxf(mx, __uint_as_float(rr[b][j]));
        }
      }
    }
    x_max[h * 128 + r] = mx;
    asm volatile("bar.sync 1, %0;" ::"n"(32 * ATT_SM_WARPS) : "memory");
    mx = fmaxf(x_max[r], x_max[128 + r]);
    float sum = 0.f;
    for (int c0 = k_lo; c0 < k_lo + kh; c0 += 64) {
      uint32_t rr[4][16];
#pragma unroll
      for (int b = 0; b < 4; ++b) ptx::tmem_ld16(tmem_s + trow + (uint32_t)(c0 + 16 * b), rr[b]);
      ptx::tmem_ld_wait();
      uint8_t* tile_p = gP + (c0 / 64) * q_chunk + r * 128;          // P tile chunk (c0/64), row r; 128B swizzle: unit ^= (r & 7)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const bool same = p.imgs_per_tile == 1 || ((c0 + 16 * b) / p.L) == my_img_local;
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float e0 = 0.f, e1 = 0.f;
          if (same) {
            e0 = exp2f((__uint_as_float(rr[b][2 * j]) - mx) * p.scale_log2);
            e1 = exp2f((__uint_as_float(rr[b][2 * j + 1]) - mx) * p.scale_log2);
          }
          __nv_bfloat162 hh = __floats2bfloat162_rn(e0, e1);
          // the row sum uses the bf16-rounded values the tensor core will see
          sum += __bfloat162float(hh.x) + __bfloat162float(hh.y);
          pk[j] = *reinterpret_cast<uint32_t*>(&hh);
        }
        const int u = 2 * b;                                         // 16-byte units u, u+1 of the 128-byte row
        *reinterpret_cast<uint4*>(tile_p + ((u ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(tile_p + (((u + 1) ^ (r & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
    }
    x_sum[h * 128 + r] = sum;
    ptx::fence_proxy_async();       // generic-proxy smem writes -> visible to the tensor core (async proxy)
    ptx::mbar_arrive(bar_p);
    asm volatile("bar.sync 1, %0;" ::"n"(32 * ATT_SM_WARPS) : "memory");
    sum = x_sum[r] + x_sum[128 + r];
    const float inv = 1.f / sum;
    if (p.lse && valid && h == 0) p.lse[m] = mx * p.scale_log2 + log2f(sum);   // P = exp2(s * scale_log2 - lse)
    ptx::mbar_wait(bar_o, 0);
    ptx::tc_fence_after();
    const int ch = p.C / 2, c_lo = h * ch;                           // this warp's output channels (a multiple of 32)
    for (int c0 = c_lo; c0 < c_lo + ch; c0 += 32) {
      uint32_t rr[2][16];
      ptx::tmem_ld16(tmem_o + trow + (uint32_t)c0, rr[0]);
      ptx::tmem_ld16(tmem_o + trow + (uint32_t)(c0 + 16), rr[1]);
      ptx::tmem_ld_wait();
      if (valid) {
        __nv_bfloat16* op = p.out + (size_t)m * p.C + c0;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          float lo[8], hi[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) { lo[j] = __uint_as_float(rr[b][j]) * inv; hi[j] = __uint_as_float(rr[b][8 + j]) * inv; }
          store_vec(op + 16 * b, lo);
          store_vec(op + 16 * b + 8, hi);
        }
      }
    }
    ptx::tc_fence_before();
  }
  __syncthreads();
  if (warp == ATT_SM_WARPS) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_s, (uint32_t)p.tmem_cols);
  }
}

int attention_tc(const void* qk, const void* vt, int images, int L, int C, void* out, float* lse, cudaStream_t st) {
  VF_REQUIRE(C % 64 == 0, "vf_attention(tc): C=%d not a multiple of 64", C);
  VF_REQUIRE(L == 64 || (L >= 128 && L <= 256 && L % 128 == 0), "vf_attention(tc): L=%d unsupported (64, 128, 256)", L);
  AttnTcParams p{};
  p.L = L; p.C = C; p.images = images; p.M = images * L;
  p.imgs_per_tile = L < 128 ? 128 / L : 1;
  p.tiles_per_img = L < 128 ? 1 : L / 128;
  p.nkeys = L < 128 ? 128 : L;
  p.kchunks = C / 64;
  p.pchunks = p.nkeys / 64;
  p.n_parts = C > 256 ? 2 : 1;
  p.n_half = C / p.n_parts;
  p.v_mn = vt == nullptr;
  if (p.v_mn) {
    // parts start on 64-channel atoms: C <= 256 -> one part, otherwise 192 + (C - 192)
    p.part_c0[0] = 0; p.part_n[0] = C > 256 ? 192 : C;
    p.part_c0[1] = p.part_n[0]; p.part_n[1] = C - p.part_n[0];
    VF_REQUIRE(p.part_n[1] <= 256, "vf_attention(tc): C=%d cannot be tiled", C);
    for (int i = 0; i < 2; ++i) p.idesc_o_part[i] = ptx::make_idesc_bf16(128, p.part_n[i] > 0 ? p.part_n[i] : 64, 0, 1);
  }
  VF_REQUIRE(p.n_half % 16 == 0 && p.n_half <= 256, "vf_attention(tc): C=%d cannot be tiled", C);
  VF_REQUIRE(p.nkeys + C <= 512, "vf_attention(tc): L=%d C=%d exceed TMEM", L, C);
  p.tmem_cols = 32;
  while (p.tmem_cols < p.nkeys + C) p.tmem_cols *= 2;
  p.idesc_s = ptx::make_idesc_bf16(128, p.nkeys, 0, 0);
  p.idesc_o = ptx::make_idesc_bf16(128, p.n_half, 0, 0);
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)C);
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  const uint32_t q_bytes = (uint32_t)p.kchunks * 128 * 128, pp_bytes = (uint32_t)p.pchunks * 128 * 128;
  const uint32_t k_bytes = (uint32_t)p.kchunks * p.nkeys * 128, v_bytes = (uint32_t)p.pchunks * C * 128;
  p.off_k = q_bytes > pp_bytes ? q_bytes : pp_bytes;
  const size_t smem = 1024 + ATT_HDR + p.off_k + (k_bytes > v_bytes ? k_bytes : v_bytes);
  VF_REQUIRE(smem <= 227 * 1024, "vf_attention(tc): L=%d C=%d need %zu B of shared memory", L, C, smem);

  CUtensorMap mapQK, mapVT;
  {
    const uint64_t dims[2] = {(uint64_t)3 * C, (uint64_t)images * L};
    const uint64_t strides[1] = {(uint64_t)3 * C * 2};
    const uint32_t box_q[2] = {64, 128};
    // Q rows (128) and K rows (nkeys = 128 or 256) come from the same tensor through 128-row boxes
    int rc = encode_bf16_map(&mapQK, qk, 2, dims, strides, box_q);
    if (rc) return rc;
  }
  if (p.v_mn) {
    const uint64_t dims[2] = {(uint64_t)3 * C, (uint64_t)images * L};
    const uint64_t strides[1] = {(uint64_t)3 * C * 2};
    const uint32_t box[2] = {64, (uint32_t)p.nkeys};
    int rc = encode_bf16_map(&mapVT, qk, 2, dims, strides, box);
    if (rc) return rc;
  } else {
    const uint64_t dims[2] = {(uint64_t)L, (uint64_t)images * C};
    const uint64_t strides[1] = {(uint64_t)L * 2};
    const uint32_t box[2] = {64, (uint32_t)(C > 256 ? C / 2 : C)};
    int rc = encode_bf16_map(&mapVT, vt, 2, dims, strides, box);
    if (rc) return rc;
  }
  VF_SET_MAX_SMEM(attn_tc_kernel, 227 * 1024);
  VF_CUDA(launch_pdl(attn_tc_kernel, dim3(cdiv(p.M, 128)), dim3(ATT_THREADS), smem, st, mapQK, mapVT, p));
  VF_LAUNCH_CHECK();
  return VF_OK;
}

}  // namespace vf
