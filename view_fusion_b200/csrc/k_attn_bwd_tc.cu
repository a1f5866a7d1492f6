// Self-attention backward on tcgen05 (reference: autograd of model/unet.py:267-274, experiment.py:292).
//
// One CTA per (image, 128-query block i); loop over the 128-key blocks j of the image:
//   S  = Q_i K_j^T            A = Q (K-major),   B = K (K-major)                     -> TMEM [0,128)
//   dP = dO_i V_j^T           A = dO (K-major),  B = V^T tile as loaded (MN-major)   -> TMEM [128,256)
//   P  = exp2(S * log2e/sqrt(C) - lse),  dS = P * (dP - delta) / sqrt(C)             (thread == query row; bf16 tiles
//        in 128B-swizzled smem; lse comes from the forward kernel, delta = rowsum(dO * O) is computed here)
//   dQ_i += dS K_j            A = dS (K-major),  B = K_j tile as loaded (MN-major)   -> TMEM [256,256+C), kept over j
//   dV_j  = P^T dO_i          A = P tile read MN-major, B = dO tile read MN-major    -> TMEM [0,C) -> fp32 partial slot i
//   dK_j  = dS^T Q_i          A = dS tile read MN-major, B = Q tile read MN-major    -> TMEM [0,C) -> fp32 partial slot i
// No operand is ever transposed in memory: a [rows x 64] swizzled tile is K-major for one product and MN-major for
// the next (descriptor conventions probed in k_debug.cu and shared with k_wgrad_tc.cu).
// The per-query-block partial dK/dV go to slot i of the scratch (plain stores, deterministic); attn_dkv_finish sums
// the slots into the bf16 dqkv matrix.
// smem (C=192): Q, dO, K, V^T tiles 48 KB each + dS 32 KB; P reuses the V^T region once dP is complete.
#include <cuda.h>

#include <mutex>

#include "tc_ptx.cuh"
#include "vf_common.cuh"

namespace vf {

int encode_bf16_map(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box);

struct AttnBwdParams {
  int L, C, images;
  int nblk;             // L / 128
  int kchunks;          // C / 64
  uint32_t tile_bytes;  // kchunks * 16 KB
  uint32_t idesc_s, idesc_dp, idesc_dq, idesc_kv;
  float scale_log2, scale;
  const __nv_bfloat16* o;
  const __nv_bfloat16* d_out;
  const float* lse;
  __nv_bfloat16* dqkv;
  float* part;          // [nblk][images*L][2C]
};

constexpr int AB_THREADS = 160;

__global__ void __launch_bounds__(AB_THREADS, 1) attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap mapQKV,
                                                                    const __grid_constant__ CUtensorMap mapDO,
                                                                    const __grid_constant__ CUtensorMap mapVT,
                                                                    const AttnBwdParams p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const uint32_t bar_q = base, bar_ld = base + 8, bar_s = base + 16, bar_p = base + 24, bar_v = base + 32, bar_d = base + 40,
                 bar_k = base + 48, bar_e = base + 56;
  const uint32_t tmem_slot = base + 64;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gbase + 64);
  const uint32_t regQ = base + 1024;
  const uint32_t regDO = regQ + p.tile_bytes;
  const uint32_t regK = regDO + p.tile_bytes;
  const uint32_t regV = regK + p.tile_bytes;     // V^T tile, then P
  const uint32_t regDS = regV + p.tile_bytes;
  uint8_t* gP = gbase + 1024 + 3 * (size_t)p.tile_bytes;
  uint8_t* gDS = gP + p.tile_bytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int img = blockIdx.x / p.nblk, qb = blockIdx.x % p.nblk;
  const int row0 = img * p.L + qb * 128;        // first query row of this CTA
  const int key_row0 = img * p.L;
  const int C = p.C;
  constexpr uint32_t CH = 128 * 128;            // one 64-column chunk of a 128-row tile

  if (threadIdx.x == 0) {
    ptx::mbar_init(bar_q, 1);
    ptx::mbar_init(bar_ld, 1);
    ptx::mbar_init(bar_s, 1);
    ptx::mbar_init(bar_p, 128);
    ptx::mbar_init(bar_v, 1);
    ptx::mbar_init(bar_d, 128);
    ptx::mbar_init(bar_k, 1);
    ptx::mbar_init(bar_e, 128);
    ptx::fence_barrier_init();
  }
  if (warp == 4) {
    if (lane == 0) { ptx::prefetch_tmap(&mapQKV); ptx::prefetch_tmap(&mapDO); ptx::prefetch_tmap(&mapVT); }
    ptx::tmem_alloc(tmem_slot, 512u);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = *tmem_slot_ptr;
  const uint32_t tm_s = tm, tm_dp = tm + 128, tm_dq = tm + 256, tm_kv = tm;
  pdl_wait();                                    // barrier / TMEM set-up above overlapped the previous kernel's tail

  if (warp == 4) {
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(bar_q, 2u * p.tile_bytes);
      for (int c = 0; c < p.kchunks; ++c) {
        ptx::tma_load_2d(regQ + c * CH, &mapQKV, bar_q, c * 64, row0);
        ptx::tma_load_2d(regDO + c * CH, &mapDO, bar_q, c * 64, row0);
      }
      for (int j = 0; j < p.nblk; ++j) {
        const uint32_t ph = (uint32_t)j & 1u;
        // K_j rows and the V^T tile (C channel rows x 128 keys as two 64-key boxes); the previous block's readers of
        // both regions have completed (bar_d / bar_e were waited below)
        ptx::mbar_arrive_expect_tx(bar_ld, 2u * p.tile_bytes);
        for (int c = 0; c < p.kchunks; ++c) ptx::tma_load_2d(regK + c * CH, &mapQKV, bar_ld, C + c * 64, key_row0 + j * 128);
        for (int h = 0; h < 2; ++h) ptx::tma_load_2d(regV + h * (uint32_t)C * 128u, &mapVT, bar_ld, j * 128 + h * 64, img * C);
        if (j == 0) ptx::mbar_wait(bar_q, 0);
        ptx::mbar_wait(bar_ld, ph);
        ptx::tc_fence_after();
        // S = Q K^T (both K-major)
        for (int c = 0; c < p.kchunks; ++c)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = ptx::make_smem_desc(regQ + c * CH + k * 32, 16, 1024);
            const uint64_t bd = ptx::make_smem_desc(regK + c * CH + k * 32, 16, 1024);
            ptx::umma_f16(tm_s, ad, bd, p.idesc_s, (c > 0 || k > 0) ? 1u : 0u);
          }
        // dP = dO V^T: B is the V^T tile [channel rows][keys] = MN-major, N atoms (64 keys) C*128 bytes apart
        for (int c = 0; c < p.kchunks; ++c)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = ptx::make_smem_desc(regDO + c * CH + k * 32, 16, 1024);
            const uint64_t bd = ptx::make_smem_desc(regV + (uint32_t)(c * 64 + k * 16) * 128u, (uint32_t)C * 128u, 1024);
            ptx::umma_f16(tm_dp, ad, bd, p.idesc_dp, (c > 0 || k > 0) ? 1u : 0u);
          }
        ptx::umma_commit(bar_s);
        ptx::mbar_wait(bar_p, ph);
        ptx::tc_fence_after();
        // dQ += dS K_j (A K-major over keys; B = K tile read MN-major: rows = keys, atoms = 64-channel chunks)
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t ad = ptx::make_smem_desc(regDS + (kk >> 2) * CH + (kk & 3) * 32, 16, 1024);
          const uint64_t bd = ptx::make_smem_desc(regK + (uint32_t)(16 * kk) * 128u, CH, 1024);
          ptx::umma_f16(tm_dq, ad, bd, p.idesc_dq, (j > 0 || kk > 0) ? 1u : 0u);
        }
        // dV_j = P^T dO (both MN-major over the 128 query rows)
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t ad = ptx::make_smem_desc(regV + (uint32_t)(16 * kk) * 128u, CH, 1024);
          const uint64_t bd = ptx::make_smem_desc(regDO + (uint32_t)(16 * kk) * 128u, CH, 1024);
          ptx::umma_f16(tm_kv, ad, bd, p.idesc_kv, kk > 0 ? 1u : 0u);
        }
        ptx::umma_commit(bar_v);
        ptx::mbar_wait(bar_d, ph);
        ptx::tc_fence_after();
        // dK_j = dS^T Q
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t ad = ptx::make_smem_desc(regDS + (uint32_t)(16 * kk) * 128u, CH, 1024);
          const uint64_t bd = ptx::make_smem_desc(regQ + (uint32_t)(16 * kk) * 128u, CH, 1024);
          ptx::umma_f16(tm_kv, ad, bd, p.idesc_kv, kk > 0 ? 1u : 0u);
        }
        ptx::umma_commit(bar_k);
        ptx::mbar_wait(bar_e, ph);      // dK drained: TMEM [0,C), the K / V^T / dS regions are free again
        ptx::tc_fence_after();
      }
    }
  } else {
    const int r = warp * 32 + lane;
    const int m = row0 + r;
    const uint32_t trow = (uint32_t)(warp * 32) << 16;
    // delta = sum_c dO[m][c] * O[m][c]
    float delta = 0.f;
    {
      const __nv_bfloat16* a = p.d_out + (size_t)m * C;
      const __nv_bfloat16* b = p.o + (size_t)m * C;
      for (int c = 0; c < C; c += 8) {
        float x[8], y[8];
        load_vec(a + c, x);
        load_vec(b + c, y);
#pragma unroll
        for (int t = 0; t < 8; ++t) delta += x[t] * y[t];
      }
    }
    const float lse = __ldg(p.lse + m);
    for (int j = 0; j < p.nblk; ++j) {
      const uint32_t ph = (uint32_t)j & 1u;
      ptx::mbar_wait(bar_s, ph);
      ptx::tc_fence_after();
      for (int c0 = 0; c0 < 128; c0 += 16) {
        uint32_t rs[16], rd[16];
        ptx::tmem_ld16(tm_s + trow + (uint32_t)c0, rs);
        ptx::tmem_ld16(tm_dp + trow + (uint32_t)c0, rd);
        ptx::tmem_ld_wait();
        uint32_t pp[8], ds[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const float p0 = exp2f(__uint_as_float(rs[2 * t]) * p.scale_log2 - lse);
          const float p1 = exp2f(__uint_as_float(rs[2 * t + 1]) * p.scale_log2 - lse);
          const float d0 = p0 * (__uint_as_float(rd[2 * t]) - delta) * p.scale;
          const float d1 = p1 * (__uint_as_float(rd[2 * t + 1]) - delta) * p.scale;
          __nv_bfloat162 hp = __floats2bfloat162_rn(p0, p1), hd = __floats2bfloat162_rn(d0, d1);
          pp[t] = *reinterpret_cast<uint32_t*>(&hp);
          ds[t] = *reinterpret_cast<uint32_t*>(&hd);
        }
        const int u = (c0 % 64) / 8;
        const uint32_t off = (uint32_t)(c0 / 64) * CH + (uint32_t)r * 128u;
        const uint32_t u0 = (uint32_t)((u ^ (r & 7)) << 4), u1 = (uint32_t)(((u + 1) ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(gP + off + u0) = make_uint4(pp[0], pp[1], pp[2], pp[3]);
        *reinterpret_cast<uint4*>(gP + off + u1) = make_uint4(pp[4], pp[5], pp[6], pp[7]);
        *reinterpret_cast<uint4*>(gDS + off + u0) = make_uint4(ds[0], ds[1], ds[2], ds[3]);
        *reinterpret_cast<uint4*>(gDS + off + u1) = make_uint4(ds[4], ds[5], ds[6], ds[7]);
      }
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar_p);
      // partial dV then dK of key row (key_row0 + 128 j + r): slot qb, columns [C,2C) and [0,C)
      float* prow = p.part + ((size_t)qb * p.images * p.L + key_row0 + j * 128 + r) * (size_t)(2 * C);
      for (int pass = 0; pass < 2; ++pass) {
        ptx::mbar_wait(pass == 0 ? bar_v : bar_k, ph);
        ptx::tc_fence_after();
        float* dst = prow + (pass == 0 ? C : 0);
        for (int c0 = 0; c0 < C; c0 += 16) {
          uint32_t rr[16];
          ptx::tmem_ld16(tm_kv + trow + (uint32_t)c0, rr);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int t = 0; t < 16; t += 4)
            *reinterpret_cast<float4*>(dst + c0 + t) = make_float4(__uint_as_float(rr[t]), __uint_as_float(rr[t + 1]),
                                                                    __uint_as_float(rr[t + 2]), __uint_as_float(rr[t + 3]));
        }
        ptx::tc_fence_before();
        ptx::mbar_arrive(pass == 0 ? bar_d : bar_e);
      }
    }
    // dQ rows (complete: the last bar_k commit covers every earlier MMA)
    ptx::tc_fence_after();
    __nv_bfloat16* dq = p.dqkv + (size_t)m * 3 * C;
    for (int c0 = 0; c0 < C; c0 += 16) {
      uint32_t rr[16];
      ptx::tmem_ld16(tm_dq + trow + (uint32_t)c0, rr);
      ptx::tmem_ld_wait();
      float lo[8], hi[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) { lo[t] = __uint_as_float(rr[t]); hi[t] = __uint_as_float(rr[8 + t]); }
      store_vec(dq + c0, lo);
      store_vec(dq + c0 + 8, hi);
    }
    ptx::tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tm, 512u);
  }
}

// dqkv[:, C:3C] = sum over the query-block slots of the fp32 partials
__global__ void __launch_bounds__(256) attn_dkv_finish_kernel(const float* __restrict__ part, int nslots, size_t rows, int C,
                                                              __nv_bfloat16* __restrict__ dqkv) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t total = rows * (size_t)(2 * C / 8);
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const size_t r = gid / (2 * C / 8);
  const int c = (int)(gid % (2 * C / 8)) * 8;
  float acc[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) acc[t] = 0.f;
  for (int s = 0; s < nslots; ++s) {
    const float* src = part + ((size_t)s * rows + r) * (size_t)(2 * C) + c;
    const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
    acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
    acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
  }
  store_vec(dqkv + r * (size_t)(3 * C) + C + c, acc);
}

bool attention_bwd_tc_supported(int L, int C) { return (L == 128 || L == 256) && (C == 128 || C == 192); }

int attention_bwd_tc(const void* qkv, const void* vt, const void* out, const float* lse, const void* d_out, int images, int L, int C,
                     float* scratch, void* dqkv, cudaStream_t st) {
  AttnBwdParams p{};
  p.L = L; p.C = C; p.images = images;
  p.nblk = L / 128;
  p.kchunks = C / 64;
  p.tile_bytes = (uint32_t)p.kchunks * 128u * 128u;
  p.idesc_s = ptx::make_idesc_bf16(128, 128, 0, 0);
  p.idesc_dp = ptx::make_idesc_bf16(128, 128, 0, 1);
  p.idesc_dq = ptx::make_idesc_bf16(128, C, 0, 1);
  p.idesc_kv = ptx::make_idesc_bf16(128, C, 1, 1);
  p.scale = 1.f / sqrtf((float)C);
  p.scale_log2 = 1.4426950408889634f * p.scale;
  p.o = reinterpret_cast<const __nv_bfloat16*>(out);
  p.d_out = reinterpret_cast<const __nv_bfloat16*>(d_out);
  p.lse = lse;
  p.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv);
  p.part = scratch;
  const size_t rows = (size_t)images * L;
  CUtensorMap mapQKV, mapDO, mapVT;
  {
    const uint64_t dims[2] = {(uint64_t)3 * C, rows};
    const uint64_t strides[1] = {(uint64_t)3 * C * 2};
    const uint32_t box[2] = {64, 128};
    int rc = encode_bf16_map(&mapQKV, qkv, 2, dims, strides, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)C, rows};
    const uint64_t strides[1] = {(uint64_t)C * 2};
    const uint32_t box[2] = {64, 128};
    int rc = encode_bf16_map(&mapDO, d_out, 2, dims, strides, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)L, (uint64_t)images * C};
    const uint64_t strides[1] = {(uint64_t)L * 2};
    const uint32_t box[2] = {64, (uint32_t)C};
    int rc = encode_bf16_map(&mapVT, vt, 2, dims, strides, box);
    if (rc) return rc;
  }
  const size_t smem = 1024 + 1024 + 4 * (size_t)p.tile_bytes + 2 * 128 * 128;
  VF_REQUIRE(smem <= 227 * 1024, "vf_attention_backward(tc): L=%d C=%d need %zu B of shared memory", L, C, smem);
  VF_SET_MAX_SMEM(attn_bwd_tc_kernel, 227 * 1024);
  VF_CUDA(launch_pdl(attn_bwd_tc_kernel, dim3(images * p.nblk), dim3(AB_THREADS), smem, st, mapQKV, mapDO, mapVT, p));
  VF_LAUNCH_CHECK();
  const size_t total = rows * (size_t)(2 * C / 8);
  VF_CUDA(launch_pdl(attn_dkv_finish_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, (const float*)scratch, p.nblk, rows, C, p.dqkv));
  VF_LAUNCH_CHECK();
  return VF_OK;
}

}  // namespace vf
