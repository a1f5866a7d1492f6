// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (bf16 operands, fp32 accumulation in TMEM).
// Reference call sites: every nn.Conv2d of model/unet.py (:42, :189, :198, :214, :238, :255, :256).
//
//   D[m, n] = sum_seg sum_tap sum_c A_seg[row(m) + shift(tap), c] * Wt[n, koff(seg) + tap*C_seg + c]   (fp32, in TMEM)
//   out     = D + bias[n] + emb[img_row[img(m)], n] + residual[m, n]                                  (bf16 or fp32)
//
// What bounds this kernel on B200 is the L2 -> SM ingest (~42 B/clk/SM measured, B300_MICROARCH.md "LTS cap"), not
// the tensor pipe: a plain 128 x N tile needs (128+N)*128 B per 64-wide K step, i.e. 2-4x more than an SM can pull
// in while the MMAs of that step run.  The design therefore maximises operand reuse out of shared memory:
//   * PADDED row order (include/viewfusion_b200.h): the nine taps of a 3x3 convolution are nine constant ROW
//     OFFSETS of one matrix.  A K step loads ONE halo'd A slab ((128*G + 2W + 4) rows x 64 channels) and issues the
//     MMAs of all nine taps from it through smem descriptors whose start address is shifted by whole 128-byte rows
//     (verified on hardware: the 128B swizzle is a function of the absolute smem address, tests/test_gpu_ops.py::
//     test_probe_shifted_umma_descriptor).  Activation traffic drops ~9x.
//   * G accumulators of 128 rows share every weight tile (B traffic / G); TMEM holds 2 sets x G x block_n columns so
//     the epilogue of work item i overlaps the main loop of item i+1 (persistent CTAs, one per SM).
//   * (block_n, G) are chosen per layer by a small cost model of ingest bytes vs MMA cycles vs wave quantisation.
// Up to three K segments accumulate into the same tile (3x3 conv over h + 1x1 res_conv over x and the skip tensor),
// fusing the ResnetBlock's residual projection (unet.py:245) and the decoder's torch.cat (unet.py:134).
// Warp roles: warps 0..TC_EPI_WARPS-1 (8) epilogue (TMEM -> registers -> smem -> TMA store, + GroupNorm partial sums), then
// one TMA producer warp and one MMA issuer + TMEM allocator warp.
// 1x1 layers that change the row order and the stride-2 Downsample run on the same kernel through other tensor maps
// (TcParams::a_lines / epi_lines / s2_cchunks).
#include <cuda.h>

#include <cstdlib>
#include <mutex>

#include "tc_ptx.cuh"
#include "vf_common.cuh"

namespace vf {

constexpr int TC_BK = 64;
#ifndef VF_TC_EPI_WARPS
#define VF_TC_EPI_WARPS 8      // 16 was measured slower (conv class 5.48 vs 4.99 ms per step): the epilogue is not latency-bound per warp
#endif
constexpr int TC_EPI_WARPS = VF_TC_EPI_WARPS;          // epilogue warps: TC_EPI_WARPS/4 per TMEM lane quarter
constexpr int TC_EPI_THREADS = 32 * TC_EPI_WARPS;
constexpr int TC_NSUB = TC_EPI_WARPS / 4;
constexpr int TC_THREADS = TC_EPI_THREADS + 64;   // + TMA producer + MMA issuer
constexpr int TC_ABOX = 64;        // rows per A TMA box
constexpr int TC_MAX_STAGES = 16;      // weight ring
constexpr int TC_MAX_A_STAGES = 8;     // activation ring

struct TcSeg {
  int C;
  int nchunks;   // C / 64
  int ntaps;     // 1 or 9
  int koff;      // first weight column of the segment
  int halo;      // rows loaded before/after the block: (W+1)+1 for 3x3, 0 for 1x1
  int tail_rows; // >0: the last A box of a slab is this many rows (a second tensor map) instead of a full TC_ABOX: the slab is
                 // 128*G + 2*halo rows rounded up to 8, not to 64 (12-15 % fewer activation bytes per work item)
};

struct TcParams {
  RowGeom geo;
  int G, block_n, n_tiles_n, n_mblocks, n_items;
  int a_stage_bytes, a_stages, b_stages;
  int b_resident;                 // the CTA's whole weight tile stays in smem (b_stages == taps * chunks)
  int n_fixed;                    // >0 (= n_tiles_n, with b_resident): every CTA keeps ONE N tile (blockIdx % n_fixed) resident and
                                  // walks M blocks only, so layers with several N tiles also load their weights once per CTA
  int tmem_cols, max_imgs;
  int n_seg;
  TcSeg seg[3];
  uint32_t idesc;
  const float* bias;
  const float* emb;
  const int* img_row;
  int emb_ld;
  const __nv_bfloat16* residual;
  void* out;
  int out_f32, out_ld, cout;
  int qkv_split;
  __nv_bfloat16* out_vt;
  float* stats;
  long long* dbg_out;   // test hook: per-CTA cycle counters of the MMA thread [grid][4] = total, wait A, wait B, wait acc
  int dbg;        // test hook (vf_debug_flags): bit0 no stats, bit1 no store, bit2 no unit work, bit3 no bias table
  int epi_tma;    // staged TMA epilogue enabled (bf16 row-contiguous output, block_n % 64 == 0)
  // 1x1 layers that change the row order (the attention block's qkv / out projections and their data gradients):
  int a_lines;    // >0: the source is stored PADDED but read as FLAT rows: A boxes are (64 ch, W, 64/W lines) of a 3D map
  int epi_lines;  // >0: FLAT rows written to a PADDED output / residual: 32-row tiles are (64 ch, W, 32/W lines) of 3D maps
  // stride-2 3x3 (Downsample): the GEMM runs over the OUTPUT pixels only.  K chunk = (tap, 64 channels); its A tile is
  // gathered by TMA from the pixel phase (row parity, column parity) of the PADDED input the tap reads: four 4D maps
  // (channels, x/2, y/2, image), tap shifts of -1 become out-of-bounds coordinates that TMA fills with zeros.
  int s2_cchunks; // >0: Cin / 64 in that mode
};

struct MmaCtx {
  uint32_t bar_afull, bar_aempty, bar_bfull, bar_bempty, bar_accfull, bar_accempty, ringA, ringB, b_bytes, tmem_base;
  int it_first, it_stride, it_end;      // this CTA's work items: it_first, it_first + it_stride, ... < it_end
};

// The single MMA-issuing thread.  tcgen05.mma is asynchronous, but the tensor pipe only stays busy if the scalar work
// between two instructions is shorter than one MMA (57-64 cycles for N <= 128, scripts/probe_rate.py): ring indices,
// phases and descriptors therefore advance incrementally (no division / modulo), the nine tap offsets sit in
// registers, taps and row tiles are unrolled, and the cycle counters exist only in the PROF instantiation.
template <int G, bool RESIDENT, bool PROF>
__device__ __forceinline__ void mma_issue_loop(const TcParams& p, const MmaCtx& c) {
  const uint64_t desc0 = ptx::make_smem_desc(0, 16, 1024);     // K-major SWIZZLE_128B, start address 0
  const uint32_t block_n = (uint32_t)p.block_n, idesc = p.idesc;
  const uint32_t b_step = c.b_bytes >> 4;                       // descriptors hold (byte address >> 4)
  const uint64_t bd_base = desc0 + (uint64_t)(c.ringB >> 4);
  const uint32_t a_step = (uint32_t)p.a_stage_bytes >> 4;
  const uint64_t ad_ring = desc0 + (uint64_t)(c.ringA >> 4);
  const int a_stages = p.a_stages, b_stages = p.b_stages;
  uint32_t sh[9];                                               // tap (kh, kw) starts kh*(W+1) + kw rows into the slab
#pragma unroll
  for (int t = 0; t < 9; ++t) sh[t] = (uint32_t)((t / 3) * p.geo.W1 + (t % 3)) * 8u;
  int as = 0, bs = 0;
  uint32_t a_par = 0, b_par = 0, set = 0, acc_par = 1;
  uint64_t bd = bd_base, ad_stage = ad_ring;
  long long t_a = 0, t_b = 0, t_acc = 0, t0 = 0;
  const long long t_start = PROF ? clock64() : 0;
  if (RESIDENT) ptx::mbar_wait(c.bar_bfull, 0);

  auto tap_mma = [&](uint32_t acc0, uint64_t ad, uint32_t accum) {
    if (!RESIDENT) {
      if (PROF) t0 = clock64();
      ptx::mbar_wait(c.bar_bfull + 8 * bs, b_par);
      if (PROF) t_b += clock64() - t0;
      ptx::tc_fence_after();
    }
#pragma unroll
    for (int g = 0; g < G; ++g) ptx::umma_f16_k4(acc0 + (uint32_t)g * block_n, ad + (uint64_t)(g * 1024), bd, idesc, accum);
    if (!RESIDENT) {
      ptx::umma_commit(c.bar_bempty + 8 * bs);
      if (++bs == b_stages) { bs = 0; b_par ^= 1u; bd = bd_base; } else bd += b_step;
    } else {
      bd += b_step;
    }
  };

  for (int item = c.it_first; item < c.it_end; item += c.it_stride) {
    if (RESIDENT) bd = bd_base;
    if (PROF) t0 = clock64();
    ptx::mbar_wait(c.bar_accempty + 8 * set, acc_par);
    if (PROF) t_acc += clock64() - t0;
    ptx::tc_fence_after();
    const uint32_t acc0 = c.tmem_base + set * (uint32_t)G * block_n;
    uint32_t accum = 0;
    for (int s = 0; s < p.n_seg; ++s) {
      const int nchunks = p.seg[s].nchunks;
      const bool nine = p.seg[s].ntaps == 9;
      for (int ch = 0; ch < nchunks; ++ch) {
        if (PROF) t0 = clock64();
        ptx::mbar_wait(c.bar_afull + 8 * as, a_par);
        if (PROF) t_a += clock64() - t0;
        ptx::tc_fence_after();
        if (nine) {
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            tap_mma(acc0, ad_stage + sh[tap], accum);
            accum = 1;
          }
        } else {
          tap_mma(acc0, ad_stage, accum);
          accum = 1;
        }
        ptx::umma_commit(c.bar_aempty + 8 * as);
        if (++as == a_stages) { as = 0; a_par ^= 1u; ad_stage = ad_ring; } else ad_stage += a_step;
      }
    }
    ptx::umma_commit(c.bar_accfull + 8 * set);
    set ^= 1u;
    if (set == 0) acc_par ^= 1u;
  }
  if (PROF) {
    long long* o = p.dbg_out + blockIdx.x * 4;
    o[0] = clock64() - t_start; o[1] = t_a; o[2] = t_b; o[3] = t_acc;
  }
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {   // one F2FP: {hi, lo} -> bf16x2, round to nearest even
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// one 16-byte vector reduction instead of four scalar REDs (sum, sum of squares of two adjacent channels)
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// kDbg = false is the PRODUCTION instantiation: cycle counters, wall-clock stamps and the ablation bits of vf_debug_flags are
// compiled out; kDbg = true is launched only while a test hook (vf_debug_flags / vf_debug_counters) is armed.
template <bool kDbg>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const __grid_constant__ CUtensorMap mapA0,
                                                                const __grid_constant__ CUtensorMap mapA1,
                                                                const __grid_constant__ CUtensorMap mapA2,
                                                                const __grid_constant__ CUtensorMap mapA3,
                                                                const __grid_constant__ CUtensorMap mapB,
                                                                const __grid_constant__ CUtensorMap mapOut,
                                                                const __grid_constant__ CUtensorMap mapRes,
                                                                const __grid_constant__ CUtensorMap mapT0,
                                                                const __grid_constant__ CUtensorMap mapT1,
                                                                const __grid_constant__ CUtensorMap mapT2,
                                                                const TcParams p) {
  // test hook: wall-clock stamps of CTA 0 (entry, after set-up + dependency wait, after the role loops, exit) per launch,
  // appended at dbg_out[grid*12 + 1 + 4*launch ..]; dbg_out[grid*12] counts the launches
  long long* stamps = nullptr;
  const int dbg = kDbg ? p.dbg : 0;
  if (kDbg && p.dbg_out && blockIdx.x == 0 && threadIdx.x == 0) {
    long long* cnt = p.dbg_out + (size_t)gridDim.x * 12;
    const long long li = atomicAdd(reinterpret_cast<unsigned long long*>(cnt), 1ull);
    if (li < 64) { stamps = cnt + 1 + 4 * li; stamps[0] = (long long)global_timer_ns(); }
  }
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;          // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* gbase = smem_raw + (base - raw);
  // header (1 KB): barriers + tmem pointer; bias/emb table; A ring; B ring
  // header: A full / empty (8 stages each), B full / empty (16 each), accumulator full / empty, TMEM slot, per-warp residual barriers
  const uint32_t bar_afull = base, bar_aempty = base + 64, bar_bfull = base + 128, bar_bempty = base + 256;
  const uint32_t bar_accfull = base + 384, bar_accempty = base + 400;
  const uint32_t tmem_slot = base + 416;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gbase + 416);
  float* sm_bias_all = reinterpret_cast<float*>(gbase + 1024);             // [max_imgs][block_n] bias + embedding
  const uint32_t bias_bytes = ((uint32_t)(p.max_imgs * p.block_n * 4) + 1023u) & ~1023u;
  const uint32_t ringA = base + 1024 + 2 * bias_bytes;
  const uint32_t ringB = ringA + (uint32_t)p.a_stages * (uint32_t)p.a_stage_bytes;
  const uint32_t b_bytes = (uint32_t)p.block_n * TC_BK * 2;
  const int BM = 128 * p.G;
  // work items of this CTA.  Generic: item = blockIdx, blockIdx + grid, ... over (N tile major, M block minor).  Fixed-N-tile mode:
  // the CTA owns N tile blockIdx % n_fixed (its weights stay resident) and strides over the M blocks with the CTAs of that tile.
  int it_first = blockIdx.x, it_stride = gridDim.x, it_end = p.n_items;
  if (p.n_fixed > 0) {
    const int nt = (int)blockIdx.x % p.n_fixed, rank = (int)blockIdx.x / p.n_fixed;
    it_stride = ((int)gridDim.x - nt + p.n_fixed - 1) / p.n_fixed;
    it_first = nt * p.n_mblocks + rank;
    it_end = (nt + 1) * p.n_mblocks;
    if (rank >= p.n_mblocks) it_first = it_end;
  }

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.a_stages; ++s) { ptx::mbar_init(bar_afull + 8 * s, 1); ptx::mbar_init(bar_aempty + 8 * s, 1); }
    for (int s = 0; s < (p.b_resident ? 1 : p.b_stages); ++s) { ptx::mbar_init(bar_bfull + 8 * s, 1); ptx::mbar_init(bar_bempty + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(bar_accfull + 8 * s, 1); ptx::mbar_init(bar_accempty + 8 * s, TC_EPI_THREADS); }
    ptx::fence_barrier_init();
  }
  if (warp == TC_EPI_WARPS && lane == 0) {
    ptx::prefetch_tmap(&mapA0);
    ptx::prefetch_tmap(&mapB);
    if (p.n_seg > 1) ptx::prefetch_tmap(&mapA1);
    if (p.n_seg > 2 || p.s2_cchunks) { ptx::prefetch_tmap(&mapA1); ptx::prefetch_tmap(&mapA2); }
    for (int s = 0; s < p.n_seg; ++s)
      if (p.seg[s].tail_rows) ptx::prefetch_tmap(s == 0 ? &mapT0 : (s == 1 ? &mapT1 : &mapT2));
    if (p.s2_cchunks) ptx::prefetch_tmap(&mapA3);
  }
  if (warp == TC_EPI_WARPS + 1) {
    ptx::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();                                            // set-up above overlapped the previous kernel's tail
  if (stamps) stamps[1] = (long long)global_timer_ns();

  if (warp == TC_EPI_WARPS) {
    // ===================== TMA producer =====================
    if (ptx::elect_one()) {
      int as = 0, bs = 0;
      uint32_t a_par = 1, b_par = 1;                 // "empty" barriers: the first pass over each ring does not wait
      if (p.b_resident) {                        // weights-stationary: every tap's tile is loaded once per CTA
        ptx::mbar_arrive_expect_tx(bar_bfull, (uint32_t)p.b_stages * b_bytes);
        int bt = 0;
        for (int s = 0; s < p.n_seg; ++s)
          for (int ch = 0; ch < p.seg[s].nchunks; ++ch)
            for (int tap = 0; tap < p.seg[s].ntaps; ++tap, ++bt)
              ptx::tma_load_2d(ringB + bt * b_bytes, &mapB, bar_bfull, p.seg[s].koff + tap * p.seg[s].C + ch * TC_BK,
                               p.n_fixed > 0 ? ((int)blockIdx.x % p.n_fixed) * p.block_n : 0);
      }
      const int a_stages = p.a_stages, b_stages = p.b_stages;
      const bool stream_b = !p.b_resident;
      for (int item = it_first; item < it_end; item += it_stride) {
        const int n_tile = item / p.n_mblocks, m_blk = item - n_tile * p.n_mblocks;   // neighbours share the weight tile
        const int m0 = m_blk * BM, n0 = n_tile * p.block_n;
        for (int s = 0; s < p.n_seg; ++s) {
          const TcSeg sg = p.seg[s];
          const CUtensorMap* mA = s == 0 ? &mapA0 : (s == 1 ? &mapA1 : &mapA2);
          const int nbox = (BM + 2 * sg.halo + TC_ABOX - 1) / TC_ABOX;
          for (int ch = 0; ch < sg.nchunks; ++ch) {
            ptx::mbar_wait(bar_aempty + 8 * as, a_par);
            const uint32_t fa = bar_afull + 8 * as;
            ptx::mbar_arrive_expect_tx(fa, (uint32_t)((nbox - 1) * TC_ABOX + (sg.tail_rows ? sg.tail_rows : TC_ABOX)) * 128);
            const uint32_t sa = ringA + (uint32_t)as * (uint32_t)p.a_stage_bytes;
            if (p.s2_cchunks) {
              // output pixel (y, x), tap (kh, kw) reads input pixel (2y + kh - 1, 2x + kw - 1): phase (kh != 1, kw != 1),
              // shifted by -1 in the half-resolution grid when kh == 0 / kw == 0
              const int tap = ch / p.s2_cchunks, cc = ch - tap * p.s2_cchunks;
              const int kh = tap / 3, kw = tap - 3 * kh;
              const int ph = (kh != 1 ? 2 : 0) + (kw != 1 ? 1 : 0);
              const CUtensorMap* mP = ph == 0 ? &mapA0 : (ph == 1 ? &mapA1 : (ph == 2 ? &mapA2 : &mapA3));
              for (int b = 0; b < nbox; ++b) {
                const int f = m0 + b * TC_ABOX, img = f / p.geo.HW, y = (f - img * p.geo.HW) / p.geo.W;
                ptx::tma_load_4d(sa + b * (TC_ABOX * 128), mP, fa, cc * TC_BK, kw == 0 ? -1 : 0, y - (kh == 0 ? 1 : 0), img);
              }
            } else if (p.a_lines) {
              // FLAT row f of the GEMM = pixel (img, y, x) of a PADDED tensor: a 64-row box is 64/W whole image lines
              for (int b = 0; b < nbox; ++b) {
                const int f = m0 + b * TC_ABOX, img = f / p.geo.HW, y = (f - img * p.geo.HW) / p.geo.W;
                ptx::tma_load_3d(sa + b * (TC_ABOX * 128), mA, fa, ch * TC_BK, 0, img * (p.geo.H + 1) + y);
              }
            } else {
              const int nfull = sg.tail_rows ? nbox - 1 : nbox;
              for (int b = 0; b < nfull; ++b)
                ptx::tma_load_2d(sa + b * (TC_ABOX * 128), mA, fa, ch * TC_BK, m0 - sg.halo + b * TC_ABOX);
              if (sg.tail_rows)
                ptx::tma_load_2d(sa + nfull * (TC_ABOX * 128), s == 0 ? &mapT0 : (s == 1 ? &mapT1 : &mapT2), fa, ch * TC_BK, m0 - sg.halo + nfull * TC_ABOX);
            }
            if (++as == a_stages) { as = 0; a_par ^= 1u; }
            if (stream_b) {
              int kcol = sg.koff + ch * TC_BK;
              for (int tap = 0; tap < sg.ntaps; ++tap, kcol += sg.C) {
                ptx::mbar_wait(bar_bempty + 8 * bs, b_par);
                const uint32_t fb = bar_bfull + 8 * bs;
                ptx::mbar_arrive_expect_tx(fb, b_bytes);
                ptx::tma_load_2d(ringB + bs * b_bytes, &mapB, fb, kcol, n0);
                if (++bs == b_stages) { bs = 0; b_par ^= 1u; }
              }
            }
          }
        }
      }
    }
  } else if (warp == TC_EPI_WARPS + 1) {
    // ===================== MMA issuer =====================
    if (ptx::elect_one()) {
      MmaCtx mc{bar_afull, bar_aempty, bar_bfull, bar_bempty, bar_accfull, bar_accempty, ringA, ringB, b_bytes, tmem_base, it_first, it_stride, it_end};
      const bool prof = kDbg && p.dbg_out != nullptr;
      // one instantiation per (row-tile count, weights resident?, cycle counters?): the loop body is a few scalar
      // instructions per tap so the tensor pipe, not this thread, sets the pace
#define VF_MMA_CASE(GG)                                                                       \
  case GG:                                                                                    \
    if (p.b_resident) { if (kDbg && prof) mma_issue_loop<GG, true, kDbg>(p, mc); else mma_issue_loop<GG, true, false>(p, mc); }    \
    else { if (kDbg && prof) mma_issue_loop<GG, false, kDbg>(p, mc); else mma_issue_loop<GG, false, false>(p, mc); }               \
    break;
      switch (p.G) {
        VF_MMA_CASE(1)
        VF_MMA_CASE(2)
        VF_MMA_CASE(3)
        default:
        VF_MMA_CASE(4)
      }
#undef VF_MMA_CASE
    }
  } else {
    // ===================== epilogue: 8 warps = 2 x 4 TMEM lane quarters =====================
    // A unit is (accumulator g, 64-column panel): G * ceil(block_n/64) <= 4 per work item; the two warps of a lane
    // quarter take alternate units of the same item.
    // Fast path (bf16 output whose 32 rows per warp are contiguous in memory): the thread adds accumulator + bias
    // (+ a TMA-fetched residual panel) in a 128B-swizzled staging tile, the tile leaves by TMA store, and the GroupNorm
    // column sums are read back from the tile: no scattered LSU traffic.
    // Fallback (fp32 / qkv / stride-2 / non-contiguous rows): row-per-thread global accesses.
    const int q = warp & 3, sub = warp >> 2;                     // sub in [0, TC_NSUB)
    const int et = threadIdx.x;                                  // 0..255
    const int npanel = (p.block_n + 63) / 64;
    const int U = p.G * npanel;                                  // <= 4
    float* sm_bias = sm_bias_all;
    const uint32_t stg = ringB + (uint32_t)p.b_stages * b_bytes + (uint32_t)warp * 4096u;   // one 4 KB tile per warp
    uint8_t* stg_g = gbase + (stg - base);
    const uint32_t rbar = base + 448 + (uint32_t)warp * 8;
    if (lane == 0) { ptx::mbar_init(rbar, 1); ptx::fence_barrier_init(); }
    __syncwarp();
    uint32_t res_phase = 0;
    const bool has_res = p.residual != nullptr;
    // false: nothing is added in the epilogue (bias / embedding deferred to the consumer) and the table is never built; the
    // residual path reads the table unconditionally (zeros when there is no bias)
    const bool has_tab = p.bias != nullptr || p.emb != nullptr || p.residual != nullptr;
    int k_idx = 0;
    // test hook: cycle counters of epilogue warp 0 -> dbg_out[grid*4 + cta*8 + ...] = total, table, wait acc, ld+pack, store wait, stats, prep
    const bool eprof = kDbg && p.dbg_out != nullptr && et == 0;
    long long ec[7] = {0, 0, 0, 0, 0, 0, 0}, et0 = 0;
    const long long e_start = eprof ? clock64() : 0;
#define VF_EP_BEGIN() do { if (eprof) et0 = clock64(); } while (0)
#define VF_EP_END(i) do { if (eprof) ec[i] += clock64() - et0; } while (0)
    // bias + embedding table of an item, one value per (image of the block, channel): max_imgs * block_n <= 1024 entries
    constexpr int TAB_PER_THREAD = (1024 + TC_EPI_THREADS - 1) / TC_EPI_THREADS;
    float tab_next[TAB_PER_THREAD];
    auto load_table = [&](int it) {
      const int nt = it / p.n_mblocks, mb = it - nt * p.n_mblocks;
      const int img0 = div_small<false>(mb * BM, p.geo.in_padded ? p.geo.P : p.geo.HW, p.geo.rcp_rows), nn0 = nt * p.block_n;
      const float rcp_bn = 1.0f / (float)p.block_n;
#pragma unroll
      for (int j = 0; j < TAB_PER_THREAD; ++j) {
        const int i = et + j * TC_EPI_THREADS;
        float v = 0.f;
        if (i < p.max_imgs * p.block_n) {
          const int li = div_small<false>(i, p.block_n, rcp_bn), n = nn0 + i - li * p.block_n;
          const int img = img0 + li;
          if (n < p.cout) {
            if (p.bias) v += __ldg(p.bias + n);
            if (p.emb && img < p.geo.images) v += __ldg(p.emb + (size_t)__ldg(p.img_row + img) * p.emb_ld + n);
          }
        }
        tab_next[j] = v;
      }
    };
    if (p.emb != nullptr && it_first < it_end && !(dbg & 8)) load_table(it_first);
    int tab_n_tile = -1;
    // (combining the GroupNorm sums per CTA in shared memory first was tried: shared-memory float atomics cost more than
    // the global REDs they save, 12 -> 31 kclk per CTA in the statistics section)
    for (int item = it_first; item < it_end; item += it_stride, ++k_idx) {
      const int set = k_idx & 1;
      const int n_tile = item / p.n_mblocks, m_blk = item - n_tile * p.n_mblocks;
      const int m0 = m_blk * BM, n0 = n_tile * p.block_n;
      const int rows_img = p.geo.in_padded ? p.geo.P : p.geo.HW;
      const int img_first = div_small<false>(m0, rows_img, p.geo.rcp_rows);
      // bias + embedding rows of the images this block touches.  With an embedding the table changes with every item: its
      // values were fetched into registers during the previous item (tab_next), so only the two barriers around the smem
      // write are on this item's critical path.  Without one it is the bias of the N tile: rebuilt only when the CTA moves
      // to another N tile, and the epilogue warps run from item to item without meeting at a barrier.
      VF_EP_BEGIN();
      if (has_tab && (p.emb != nullptr || n_tile != tab_n_tile)) {
        if (p.emb == nullptr && !(dbg & 8)) load_table(item);
        asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_THREADS) : "memory");     // previous item's readers are done
        if (!(dbg & 8)) {
#pragma unroll
          for (int j = 0; j < TAB_PER_THREAD; ++j) {
            const int i = et + j * TC_EPI_THREADS;
            if (i < p.max_imgs * p.block_n) sm_bias[i] = tab_next[j];
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_THREADS) : "memory");
        tab_n_tile = n_tile;
        if (p.emb != nullptr && !(dbg & 8)) {
          const int item_next = item + it_stride;
          if (item_next < it_end) load_table(item_next);
        }
      }
      VF_EP_END(1);

      // the first unit's residual is fetched under the main loop
      RowInfo ri{};
      long row_base = 0;
      bool contig = false;
      auto unit_gc = [&](int u, int& g, int& c0) {       // u = g * npanel + panel, U <= 4: no integer division
        g = 0;
        int pu = u;
        while (pu >= npanel) { pu -= npanel; ++g; }
        c0 = pu * 64;
      };
      auto prep_unit = [&](int u, bool issue) {
        int g, c0;
        unit_gc(u, g, c0);
        const int f0 = m0 + g * 128 + q * 32;
        ri = decode_row<true>(p.geo, f0 + lane);
        if (p.epi_lines) {
          // FLAT rows -> PADDED output: the warp's 32 rows are 32/W whole lines of one image (3D maps, see conv2d_tc)
          row_base = (long)ri.img * (p.geo.H + 1) + __shfl_sync(0xffffffffu, ri.pix, 0) / p.geo.W;
          row_base = __shfl_sync(0xffffffffu, (int)row_base, 0);
          contig = p.epi_tma && f0 < p.geo.rows_total;
        } else {
          row_base = __shfl_sync(0xffffffffu, ri.out_row, 0);
          contig = p.epi_tma && __all_sync(0xffffffffu, ri.out_row == row_base + lane);
        }
        if (p.qkv_split > 0 && n0 + c0 >= 2 * p.qkv_split) contig = false;   // V panels leave transposed (fallback path)
        if (issue && contig) {
          if (lane == 0) {
            if (has_res) {
              ptx::tma_store_wait_read<0>();               // the staging tile's previous store has been read out
              ptx::mbar_arrive_expect_tx(rbar, 4096);
              if (p.epi_lines) ptx::tma_load_3d(stg, &mapRes, rbar, n0 + c0, 0, (int)row_base);
              else ptx::tma_load_2d(stg, &mapRes, rbar, n0 + c0, (int)row_base);
            }
          }
          __syncwarp();
        }
      };
      VF_EP_BEGIN();
      if (sub < U) prep_unit(sub, true);
      VF_EP_END(6);

      VF_EP_BEGIN();
      ptx::mbar_wait(bar_accfull + 8 * set, (uint32_t)(k_idx >> 1) & 1u);
      ptx::tc_fence_after();
      VF_EP_END(2);
      for (int u = sub; u < U && !(dbg & 4); u += TC_NSUB) {
        int g, c0;
        unit_gc(u, g, c0);
        const int width = min(64, p.block_n - c0);
        if (u != sub) prep_unit(u, true);
        const bool valid = ri.valid;
        const int li = ri.img - img_first;
        const float* brow = sm_bias + (li < p.max_imgs ? li : 0) * p.block_n + c0;
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((set * p.G + g) * p.block_n + c0);
        const int img_lo = __shfl_sync(0xffffffffu, ri.img, 0), img_hi = __shfl_sync(0xffffffffu, ri.img, 31);
        if (contig) {
          // ---------------- fast path: staged through shared memory, TMA in / TMA out ----------------
          uint8_t* tile = stg_g + lane * 128;                                // this thread's 128-byte row
          if (has_res) { ptx::mbar_wait(rbar, res_phase); res_phase ^= 1u; }
          VF_EP_BEGIN();
          // TMEM round trips per unit: with two warps per scheduler all 64 columns are requested at once; with four the
          // other warps hide the latency and two 32-column halves keep the register count under the 640-thread limit
          constexpr int NHALF = TC_EPI_WARPS > 8 ? 2 : 1, LDS_PER = 4 / NHALF, SLOTS = 8 / NHALF;
          const uint32_t sw = (uint32_t)(lane & 7);
#pragma unroll
          for (int hh = 0; hh < NHALF; ++hh) {
            uint32_t rr[LDS_PER][16];
#pragma unroll
            for (int h4 = 0; h4 < LDS_PER; ++h4) ptx::tmem_ld16(trow + (uint32_t)(hh * (64 / NHALF) + 16 * h4), rr[h4]);
            if (hh == 0 && !has_res) {                                       // the previous store must have read the tile
              const long long ts = eprof ? clock64() : 0;
              if (lane == 0) ptx::tma_store_wait_read<0>();                  // out; overlaps with the TMEM load latency
              __syncwarp();
              if (eprof) ec[4] += clock64() - ts;
            }
            ptx::tmem_ld_wait();
            // one 16-byte slot = 8 channels; the residual / padding-row cases are separate straight-line loops so the
            // common path is (2 LDS + 8 FADD + 4 F2FP + 1 STS) per slot without selects or branches
            if (!valid) {
#pragma unroll
              for (int jj = 0; jj < SLOTS; ++jj)                             // padding rows are stored as zeros
                *reinterpret_cast<uint4*>(tile + (((uint32_t)(hh * SLOTS + jj) ^ sw) << 4)) = make_uint4(0u, 0u, 0u, 0u);
            } else if (has_res) {
#pragma unroll
              for (int jj = 0; jj < SLOTS; ++jj) {
                const int j = hh * SLOTS + jj;
                uint4* slot = reinterpret_cast<uint4*>(tile + (((uint32_t)j ^ sw) << 4));
                const float4 b0 = *reinterpret_cast<const float4*>(brow + j * 8), b1 = *reinterpret_cast<const float4*>(brow + j * 8 + 4);
                const uint4 rv = *slot;
                const uint32_t* acc = &rr[jj >> 1][(jj & 1) * 8];
                uint4 o;
                o.x = pack_bf16x2(__uint_as_float(acc[0]) + b0.x + bf16_lo(rv.x), __uint_as_float(acc[1]) + b0.y + bf16_hi(rv.x));
                o.y = pack_bf16x2(__uint_as_float(acc[2]) + b0.z + bf16_lo(rv.y), __uint_as_float(acc[3]) + b0.w + bf16_hi(rv.y));
                o.z = pack_bf16x2(__uint_as_float(acc[4]) + b1.x + bf16_lo(rv.z), __uint_as_float(acc[5]) + b1.y + bf16_hi(rv.z));
                o.w = pack_bf16x2(__uint_as_float(acc[6]) + b1.z + bf16_lo(rv.w), __uint_as_float(acc[7]) + b1.w + bf16_hi(rv.w));
                *slot = o;
              }
            } else if (has_tab) {
#pragma unroll
              for (int jj = 0; jj < SLOTS; ++jj) {
                const int j = hh * SLOTS + jj;
                // (fetching the table entries before the wait on the TMEM load was measured: no effect on the 64x64 layers)
                const float4 b0 = *reinterpret_cast<const float4*>(brow + j * 8), b1 = *reinterpret_cast<const float4*>(brow + j * 8 + 4);
                const uint32_t* acc = &rr[jj >> 1][(jj & 1) * 8];
                uint4 o;
                o.x = pack_bf16x2(__uint_as_float(acc[0]) + b0.x, __uint_as_float(acc[1]) + b0.y);
                o.y = pack_bf16x2(__uint_as_float(acc[2]) + b0.z, __uint_as_float(acc[3]) + b0.w);
                o.z = pack_bf16x2(__uint_as_float(acc[4]) + b1.x, __uint_as_float(acc[5]) + b1.y);
                o.w = pack_bf16x2(__uint_as_float(acc[6]) + b1.z, __uint_as_float(acc[7]) + b1.w);
                *reinterpret_cast<uint4*>(tile + (((uint32_t)j ^ sw) << 4)) = o;
              }
            } else {
              // nothing to add (bias / embedding deferred into the consumer's GroupNorm): 4 F2FP + 1 STS per slot
#pragma unroll
              for (int jj = 0; jj < SLOTS; ++jj) {
                const int j = hh * SLOTS + jj;
                const uint32_t* acc = &rr[jj >> 1][(jj & 1) * 8];
                uint4 o;
                o.x = pack_bf16x2(__uint_as_float(acc[0]), __uint_as_float(acc[1]));
                o.y = pack_bf16x2(__uint_as_float(acc[2]), __uint_as_float(acc[3]));
                o.z = pack_bf16x2(__uint_as_float(acc[4]), __uint_as_float(acc[5]));
                o.w = pack_bf16x2(__uint_as_float(acc[6]), __uint_as_float(acc[7]));
                *reinterpret_cast<uint4*>(tile + (((uint32_t)j ^ sw) << 4)) = o;
              }
            }
          }
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0 && !(dbg & 2)) {
            if (p.epi_lines) ptx::tma_store_3d(&mapOut, stg, n0 + c0, 0, (int)row_base);
            else ptx::tma_store_2d(&mapOut, stg, n0 + c0, (int)row_base);
            ptx::tma_store_commit();
          }
          VF_EP_END(3);
          VF_EP_BEGIN();
          if (p.stats && !(dbg & 1)) {
            // column sums straight from the staging tile: lane owns channels (2*lane, 2*lane+1); conflict-free reads.
            // Padding rows were stored as zeros, so a tile inside one image is summed without any per-row test.
            const uint8_t* col = stg_g + (lane & 3) * 4;
            const uint32_t sw = (uint32_t)(lane >> 2);
            if (img_lo == img_hi) {
              float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
              for (int r = 0; r < 32; ++r) {
                const uint32_t w = *reinterpret_cast<const uint32_t*>(col + r * 128 + ((sw ^ (uint32_t)(r & 7)) << 4));
                const float a = __uint_as_float(w << 16), b = __uint_as_float(w & 0xFFFF0000u);
                s0 += a; q0 = fmaf(a, a, q0); s1 += b; q1 = fmaf(b, b, q1);
              }
              if (img_lo < p.geo.images) {
                float* sp = p.stats + ((size_t)img_lo * p.cout + n0 + c0 + 2 * lane) * 2;
                red_add_v4(sp, s0, q0, s1, q1);
              }
            } else {
              for (int im = img_lo; im <= img_hi; ++im) {
                const uint32_t rows = __ballot_sync(0xffffffffu, ri.img == im);
                float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll 8
                for (int r = 0; r < 32; ++r) {
                  if (!((rows >> r) & 1u)) continue;
                  const uint32_t w = *reinterpret_cast<const uint32_t*>(col + r * 128 + ((sw ^ (uint32_t)(r & 7)) << 4));
                  const float a = __uint_as_float(w << 16), b = __uint_as_float(w & 0xFFFF0000u);
                  s0 += a; q0 = fmaf(a, a, q0); s1 += b; q1 = fmaf(b, b, q1);
                }
                if (im < p.geo.images) {
                  float* sp = p.stats + ((size_t)im * p.cout + n0 + c0 + 2 * lane) * 2;
                  red_add_v4(sp, s0, q0, s1, q1);
                }
              }
            }
          }
          VF_EP_END(5);
        } else if (__any_sync(0xffffffffu, valid)) {
          // ---------------- fallback: row-per-thread global accesses, 16 columns at a time ----------------
          for (int cc = 0; cc < width; cc += 16) {
            uint32_t rr[16];
            ptx::tmem_ld16(trow + (uint32_t)cc, rr);
            ptx::tmem_ld_wait();
            const int n = n0 + c0 + cc;
            const bool act = valid && (n < p.cout || p.out_f32);  // structured: the warp reconverges before the next tcgen05.ld
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = act ? __uint_as_float(rr[j]) + (has_tab ? brow[cc + j] : 0.f) : 0.f;
            if (act && p.residual) {
              const __nv_bfloat16* rp = p.residual + (size_t)ri.out_row * p.cout + n;
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
                float* op = reinterpret_cast<float*>(p.out) + (size_t)ri.out_row * p.out_ld + n;
                const int cnt = min(16, p.out_ld - n);     // out_ld is the padded channel count of the fp32 output
                for (int j = 0; j + 4 <= cnt; j += 4) *reinterpret_cast<float4*>(op + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              } else if (p.qkv_split > 0 && n >= 2 * p.qkv_split) {
                __nv_bfloat16* vp = p.out_vt + ((size_t)ri.img * p.qkv_split + (n - 2 * p.qkv_split)) * p.geo.HW + ri.pix;
#pragma unroll
                for (int j = 0; j < 16; ++j) vp[(size_t)j * p.geo.HW] = __float2bfloat16_rn(v[j]);
              } else {
                __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)ri.out_row * p.out_ld + n;
                float lo[8], hi[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) { lo[j] = v[j]; hi[j] = v[8 + j]; }
                store_vec(op, lo);
                store_vec(op + 8, hi);
              }
            }
            if (p.stats && n < p.cout) {
              for (int im = img_lo; im <= img_hi; ++im) {
                float s1[16], s2[16];
                const bool own = valid && ri.img == im;
#pragma unroll
                for (int j = 0; j < 16; ++j) { s1[j] = own ? v[j] : 0.f; s2[j] = s1[j] * s1[j]; }
                const float cs = warp_colsum16(s1, lane), cq = warp_colsum16(s2, lane);
                if ((lane & 1) == 0 && im < p.geo.images) {
                  float* sp = p.stats + ((size_t)im * p.cout + n + warp_col16(lane)) * 2;
                  atomicAdd(sp, cs);
                  atomicAdd(sp + 1, cq);
                }
              }
            }
          }
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar_accempty + 8 * set);
    }
    if (eprof) {
      ec[0] = clock64() - e_start;
      long long* o = p.dbg_out + (size_t)gridDim.x * 4 + (size_t)blockIdx.x * 8;
      for (int i = 0; i < 7; ++i) o[i] = ec[i];
    }
#undef VF_EP_BEGIN
#undef VF_EP_END
    if (stamps) stamps[2] = (long long)global_timer_ns();
    if (lane == 0) ptx::tma_store_wait_all<0>();     // staged stores must land before the CTA exits
  }
  __syncthreads();
  if (warp == TC_EPI_WARPS + 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
  if (stamps) stamps[3] = (long long)global_timer_ns();
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

static int g_tc_dbg = 0;     // test hook: see vf_debug_flags
static long long* g_tc_dbg_out = nullptr;
void set_tc_debug(int f) { g_tc_dbg = f; }
int tc_debug_flags() { return g_tc_dbg; }
void set_tc_debug_out(long long* p) { g_tc_dbg_out = p; }

// ---- tiling choice -------------------------------------------------------------------------------------
struct TcTiling {
  int block_n = 0, G = 0, a_stages = 0, b_stages = 0, a_stage_bytes = 0, max_imgs = 0, b_resident = 0, n_fixed = 0;
  size_t smem = 0;
  double cost = 0;
};

static const double kIngestBytesPerClk = 42.5;   // L2 -> SM, per SM (B300_MICROARCH.md: ~6300 B/clk chip-wide / 148)
static const size_t kSmemBudget = 227 * 1024;

// the last box of an activation slab may be shorter than TC_ABOX rows: plain 2D maps only (PADDED -> PADDED and FLAT -> FLAT)
static bool a_tail_boxes(const TcParams& p) {
  static const bool off = [] { const char* e = getenv("VF_TC_TAIL"); return e && e[0] == '0'; }();      // A/B knob
  return !off && !p.a_lines && !p.s2_cchunks && !(g_tc_dbg & 0x4000);
}

static bool pick_tiling(const TcParams& p, int cout_pad, int sms, bool epi_tma, bool light_epilogue, TcTiling* best) {
  bool found = false;
  const int rows_per_img = p.geo.in_padded ? p.geo.P : p.geo.HW;
  int halo_max = 0;
  long chunks = 0, taps = 0;
  for (int s = 0; s < p.n_seg; ++s) {
    halo_max = p.seg[s].halo > halo_max ? p.seg[s].halo : halo_max;
    chunks += p.seg[s].nchunks;
    taps += (long)p.seg[s].nchunks * p.seg[s].ntaps;
  }
  const size_t a_unit = a_tail_boxes(p) ? 8 : TC_ABOX;      // granularity of an activation slab (TcSeg::tail_rows)
  for (int t = 1; t <= 32; ++t) {
    if (cout_pad % t) continue;
    const int bn = cout_pad / t;
    if (bn > 256 || bn % 16) continue;
    if (epi_tma && bn % 64) continue;
    const int force_bn = ((g_tc_dbg >> 20) & 0xFF) * 16, force_g = (g_tc_dbg >> 16) & 0xF;     // test hook: tiling sweep
    if (force_bn && bn != force_bn) continue;
    for (int G = 1; G <= 4; ++G) {
      if (force_g && G != force_g) continue;
      if (2 * G * bn > 512) continue;
      if (G * ((bn + 63) / 64) > 4) continue;      // one epilogue warp per (accumulator, 64-column panel) and lane quarter
      const int BM = 128 * G;
      TcTiling c;
      c.block_n = bn; c.G = G;
      c.a_stage_bytes = (int)align_up((size_t)BM + 2 * halo_max, a_unit) * 128;
      c.max_imgs = (BM + rows_per_img - 1) / rows_per_img + 1;
      const size_t bias_bytes = align_up((size_t)c.max_imgs * bn * 4, 1024);
      const size_t fixed = 1024 + 1024 + 2 * bias_bytes + (epi_tma ? TC_EPI_WARPS * 4096 : 0);   // + staging tiles of the TMA epilogue
      const size_t bstage = (size_t)bn * 128;
      // pipeline depth is a hard requirement (both rings run across work items): >= 2 A slabs so the next slab
      // loads under the current one's MMAs, >= 4 weight tiles so a tap never waits for its own load
      c.a_stages = 2;
      // (a deeper activation ring — up to 8 slabs for the 1x1-only layers — was measured: no effect, scripts/sweep_1x1.py)
      bool only_1x1 = true;
      for (int s2 = 0; s2 < p.n_seg; ++s2) only_1x1 = only_1x1 && p.seg[s2].ntaps == 1;
      const int a_want = 3;
      const long mblocks = (p.geo.rows_total + BM - 1) / BM;
      const bool fits_resident = taps <= 64 && fixed + 2 * (size_t)c.a_stage_bytes + (size_t)taps * bstage <= kSmemBudget;
      const bool pin_tile = !(g_tc_dbg & 512) && !(g_tc_dbg & 4096) && t > 1 && t <= 8 && fits_resident && mblocks * t >= 4L * sms;
      if ((!(g_tc_dbg & 512) && t == 1 && fits_resident) || pin_tile) {
        // weights-stationary: the CTA's whole [block_n x K] tile lives in smem.  With several N tiles (pin_tile) every CTA keeps
        // N tile (blockIdx % t) and strides over the M blocks only: the weights are fetched once per CTA instead of once per
        // work item (1x1 projections: qkv)
        c.b_resident = 1;
        c.n_fixed = pin_tile ? t : 0;
        c.b_stages = (int)taps;
        const size_t left = kSmemBudget - fixed - (size_t)taps * bstage;
        c.a_stages = (int)(left / c.a_stage_bytes);
        if (c.a_stages > a_want) c.a_stages = a_want;
      } else {
        const int b_min = taps < 4 ? (int)taps : 4;
        if (fixed + (size_t)c.a_stages * c.a_stage_bytes + (size_t)b_min * bstage > kSmemBudget) continue;
        const int b_want = taps < 8 ? (int)taps : 8;                      // weight tiles in flight before A gets more than two slabs
        if (fixed + (size_t)b_want * bstage + 2 * (size_t)c.a_stage_bytes <= kSmemBudget) {
          c.a_stages = (int)((kSmemBudget - fixed - (size_t)b_want * bstage) / c.a_stage_bytes);
          if (c.a_stages > a_want) c.a_stages = a_want;
        }
        size_t left = kSmemBudget - fixed - (size_t)c.a_stages * c.a_stage_bytes;
        c.b_stages = (int)(left / bstage);
        if (c.b_stages > TC_MAX_STAGES) c.b_stages = TC_MAX_STAGES;
      }
      c.smem = fixed + (size_t)c.a_stages * c.a_stage_bytes + (size_t)c.b_stages * bstage;
      // cost model per work item
      double bytes = 0, cyc = 0;
      for (int s = 0; s < p.n_seg; ++s) {
        const TcSeg& sg = p.seg[s];
        const double arows = (double)align_up((size_t)BM + 2 * sg.halo, a_unit);
        bytes += sg.nchunks * (arows * 128.0 + (c.b_resident ? 0.0 : sg.ntaps * bn * 128.0));
        // measured (scripts/probe_rate.py, prof_conv.py): with two or more accumulators in rotation an MMA completes every
        // max(N/2, ~58) clk (the floor is the smem operand fetch); a single accumulator chains at ~91 clk; plus ~120 clk
        // of barrier / descriptor work per tap
        double base = bn <= 64 ? 60.0 : (bn <= 128 ? 66.0 : bn / 2.0 + 4.0);
        if (G == 1 && base < 91.0) base = 91.0;
        const double per_mma = base + 120.0 / (4.0 * G);
        cyc += (double)sg.nchunks * sg.ntaps * 4 * G * per_mma;
      }
      // the epilogue of an item overlaps the next item's main loop; the two warps of a TMEM lane quarter take alternate
      // (row tile, 64-column panel) units, ~2600 clk each with the GroupNorm sums
      double epi = 1500.0 + 2600.0 * ((G * ((bn + 63) / 64) + TC_NSUB - 1) / TC_NSUB);
      if (only_1x1 && light_epilogue) {
        // 1x1-only layers without bias / residual / statistics (qkv) have so few MMAs per item that the item time IS the epilogue's latency chain; measured per item at
        // 168 view-images (scripts/sweep_1x1.py, qkv 192 -> 576): 1 / 2 / 3 / 4 units of (128 rows x 64 columns) = 5.2 / 6.2 /
        // 6.8 / 10.8 kclk — three units of ONE 192-column accumulator cost barely more than two
        static const double kEpi1x1[5] = {0.0, 5200.0, 6200.0, 6800.0, 10800.0};
        epi = kEpi1x1[G * ((bn + 63) / 64)];
      }
      double item = cyc > bytes / kIngestBytesPerClk ? cyc : bytes / kIngestBytesPerClk;
      if (epi > item) item = epi;
      const long items = (long)((p.geo.rows_total + BM - 1) / BM) * t;
      long rounds = (items + sms - 1) / sms;
      if (c.n_fixed) {                            // CTAs are split evenly over the N tiles; the one-off weight fetch is amortised over a CTA's items
        const long per_tile = sms / t > 0 ? sms / t : 1;
        rounds = (mblocks + per_tile - 1) / per_tile;
        item += (double)taps * bstage / kIngestBytesPerClk / (double)rounds;
      }
      // larger row groups shrink the weight ring and lengthen the drain of the last item: measured ~8 % per extra tile
      c.cost = rounds * item * (G > 2 ? 1.0 + 0.08 * (G - 2) : 1.0) + 3000.0;
      {
        static const double fixed_bias = [] { const char* e = getenv("VF_TC_FIXED_BIAS"); return e ? atof(e) : 1.0; }();   // A/B knob
        if (c.n_fixed) c.cost *= fixed_bias;
      }
      if (!found || c.cost < best->cost * 0.999 || (c.cost < best->cost * 1.001 && bn > best->block_n)) { *best = c; found = true; }
    }
  }
  return found;
}

// Whether the stride-2 3x3 convolution of a PADDED [images*(H+1)*(W+1), C] source runs in the phase-gather mode (which
// never reads the source's padding rows) or in the full-resolution fallback (which needs zeros there).  The plan asks
// this before deciding to skip vf_zero_padding in inference.
bool conv2d_tc_stride2_gathers(int H, int W, int C) {
  const int Wo2 = W / 2, Ho2 = H / 2;
  return H % 2 == 0 && W % 2 == 0 && Wo2 >= 1 && Wo2 <= TC_ABOX && TC_ABOX % Wo2 == 0 && Ho2 % (TC_ABOX / Wo2) == 0 && C % TC_BK == 0 &&
         !(g_tc_dbg & 1024);
}

// Test hook (vf_debug_conv_tiling): when set, conv2d_tc stops after the host-side planning (geometry, mode selection,
// tiling) and reports it here instead of encoding tensor maps and launching.
static thread_local int* g_tc_plan_out = nullptr;

int conv2d_tc(const vf_conv_args* a, cudaStream_t st) {
  VF_REQUIRE(a->dtype == VF_BF16, "vf_conv2d(tc): bf16 activations only");
  TcParams p{};
  // Downsample (3x3, stride 2): GEMM rows are the output pixels (FLAT order at half resolution), see TcParams::s2_cchunks
  const int Wo2 = a->W / 2, Ho2 = a->H / 2;
  const bool s2 = a->stride == 2 && a->n_seg == 1 && a->ksize[0] == 3 && a->in_padded && a->out_padded &&
                  conv2d_tc_stride2_gathers(a->H, a->W, a->src_c[0]);
  const int H = s2 ? Ho2 : a->H, W = s2 ? Wo2 : a->W;      // resolution of the GEMM rows
  // A 1x1 layer from a PADDED source to a FLAT output runs as a FLAT -> FLAT GEMM: TMA gathers the valid pixels (whole
  // image lines) out of the padded tensor, so the rows map 1:1 onto the output and the staged TMA epilogue applies.
  const bool gather = a->n_seg == 1 && a->ksize[0] == 1 && a->in_padded && !a->out_padded && a->stride != 2 && W <= TC_ABOX &&
                      TC_ABOX % W == 0 && H % (TC_ABOX / W) == 0 && !(g_tc_dbg & 1024);
  const int in_padded = (gather || s2) ? 0 : a->in_padded;
  p.a_lines = gather ? W : 0;
  p.s2_cchunks = s2 ? a->src_c[0] / TC_BK : 0;
  p.geo = make_geom(a->images, H, W, in_padded, a->out_padded, a->stride == 2 && !s2);
  VF_REQUIRE(p.geo.rows_total < (1 << 24), "vf_conv2d(tc): %d rows exceed the 2^24 limit of the epilogue's row arithmetic", p.geo.rows_total);
  VF_REQUIRE(!p.geo.stride2 || (H % 2 == 0 && W % 2 == 0), "vf_conv2d(tc): stride 2 needs even H, W");
  VF_REQUIRE(a->cout_pad % 16 == 0, "vf_conv2d(tc): cout_pad=%d not a multiple of 16", a->cout_pad);

  int k_total = 0;
  p.n_seg = a->n_seg;
  for (int s = 0; s < a->n_seg; ++s) {
    const int C = a->src_c[s];
    VF_REQUIRE(C % TC_BK == 0, "vf_conv2d(tc): segment %d channels %d not a multiple of 64", s, C);
    VF_REQUIRE(a->ksize[s] == 1 || (a->ksize[s] == 3 && a->in_padded), "vf_conv2d(tc): 3x3 needs PADDED sources");
    TcSeg& sg = p.seg[s];
    sg.C = C; sg.nchunks = C / TC_BK; sg.ntaps = a->ksize[s] * a->ksize[s]; sg.koff = k_total;
    sg.halo = a->ksize[s] == 3 ? W + 2 : 0;
    k_total += sg.ntaps * C;
    if (s2) {   // nine taps x C/64 chunks, each with its own gathered A tile: a 1x1-like segment over K = 9C ([tap][cin] weight order)
      sg.C = 9 * C; sg.nchunks = 9 * C / TC_BK; sg.ntaps = 1; sg.halo = 0;
    }
  }
  const bool out_f32 = a->out_dtype == VF_F32;
  // staged TMA epilogue: the 32 output rows of an epilogue warp must be one box of the output tensor map
  //   PADDED -> PADDED, FLAT -> FLAT: rows map 1:1 (2D maps);   FLAT -> PADDED: W % 32 == 0 (part of one line, 2D maps) or
  //   32 % W == 0 (32/W whole lines of one image, 3D maps: p.epi_lines)
  const bool lines = !in_padded && a->out_padded && W % 32 != 0 && 32 % W == 0 && (H * W) % 32 == 0 && !(g_tc_dbg & 1024);
  p.epi_lines = lines ? W : 0;
  const bool layout_ok = (in_padded && a->out_padded) || (!in_padded && !a->out_padded && !(g_tc_dbg & 1024)) ||
                         (!in_padded && a->out_padded && (W % 32 == 0 || lines));
  p.epi_tma = !out_f32 && (!a->qkv_split || ((2 * a->qkv_split) % 64 == 0 && !(g_tc_dbg & 1024))) && !p.geo.stride2 && layout_ok &&
              a->cout_pad % 64 == 0 && a->cout == a->cout_pad;
  TcTiling tl;
  const bool light_epilogue = !a->stats && !a->bias && !a->emb && !a->residual;
  VF_REQUIRE(pick_tiling(p, a->cout_pad, sm_count(), p.epi_tma != 0, light_epilogue, &tl), "vf_conv2d(tc): no tiling for cout_pad=%d W=%d", a->cout_pad, W);
  p.block_n = tl.block_n; p.G = tl.G; p.a_stages = tl.a_stages; p.b_stages = tl.b_stages; p.a_stage_bytes = tl.a_stage_bytes;
  p.b_resident = tl.b_resident;
  p.n_fixed = tl.n_fixed;
  p.max_imgs = tl.max_imgs;
  VF_REQUIRE(p.max_imgs * p.block_n <= 1024, "vf_conv2d(tc): bias table of %d x %d entries exceeds the epilogue's registers", p.max_imgs, p.block_n);
  p.n_tiles_n = a->cout_pad / p.block_n;
  p.n_mblocks = cdiv(p.geo.rows_total, 128 * p.G);
  p.n_items = p.n_mblocks * p.n_tiles_n;
  p.tmem_cols = 32;
  while (p.tmem_cols < 2 * p.G * p.block_n) p.tmem_cols *= 2;
  p.idesc = ptx::make_idesc_bf16(128, p.block_n, 0, 0);
  if (g_tc_plan_out) {
    int* o = g_tc_plan_out;
    o[0] = p.block_n; o[1] = p.G; o[2] = p.a_stages; o[3] = p.b_stages; o[4] = p.b_resident; o[5] = (int)tl.smem; o[6] = p.tmem_cols;
    o[7] = p.n_items; o[8] = p.n_items < sm_count() ? p.n_items : sm_count(); o[9] = p.epi_tma; o[10] = p.a_lines; o[11] = p.epi_lines;
    o[12] = p.s2_cchunks; o[13] = p.geo.rows_total; o[14] = p.max_imgs; o[15] = k_total;
    return VF_OK;
  }

  CUtensorMap maps[3], maps4[4];
  // (channels, x, line) view of a PADDED tensor [images*P, ld] restricted to its valid pixels: pixel (img, y, x) is row
  // (img*(H+1) + y + 1)*(W+1) + x + 1, i.e. line img*(H+1) + y of a view that starts W+2 rows into the tensor
  auto encode_lines = [&](CUtensorMap* m, const void* ptr, int channels, int ld, int rows_per_box) {
    const uint64_t dims[3] = {(uint64_t)channels, (uint64_t)W, (uint64_t)a->images * (H + 1) - 1};
    const uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)(W + 1) * ld * 2};
    const uint32_t box[3] = {64, (uint32_t)W, (uint32_t)(rows_per_box / W)};
    return encode_bf16_map(m, reinterpret_cast<const __nv_bfloat16*>(ptr) + (size_t)(W + 2) * ld, 3, dims, strides, box);
  };
  if (s2) {
    // phase (py, px): input pixels (2*yh + py, 2*xh + px) = PADDED rows base + img*P + yh*2(Win+1) + xh*2
    const int Win = a->W, Hin = a->H, C = a->src_c[0];
    for (int ph = 0; ph < 4; ++ph) {
      const int py = ph >> 1, px = ph & 1;
      const uint64_t dims[4] = {(uint64_t)C, (uint64_t)Wo2, (uint64_t)Ho2, (uint64_t)a->images};
      const uint64_t strides[3] = {(uint64_t)2 * C * 2, (uint64_t)2 * (Win + 1) * C * 2, (uint64_t)(Hin + 1) * (Win + 1) * C * 2};
      const uint32_t box[4] = {TC_BK, (uint32_t)Wo2, (uint32_t)(TC_ABOX / Wo2), 1};
      const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(a->src[0]) + ((size_t)(py + 1) * (Win + 1) + px + 1) * C;
      int rc = encode_bf16_map(&maps4[ph], base, 4, dims, strides, box);
      if (rc) return rc;
    }
  }
  for (int s = 0; s < a->n_seg && !s2; ++s) {
    int rc;
    if (gather) {
      rc = encode_lines(&maps[s], a->src[s], a->src_c[s], a->src_c[s], TC_ABOX);
    } else {
      const uint64_t dims[2] = {(uint64_t)a->src_c[s], (uint64_t)p.geo.rows_total};
      const uint64_t strides[1] = {(uint64_t)a->src_c[s] * 2};
      const uint32_t box[2] = {TC_BK, TC_ABOX};
      rc = encode_bf16_map(&maps[s], a->src[s], 2, dims, strides, box);
    }
    if (rc) return rc;
  }
  CUtensorMap tails[3];
  for (int s = 0; s < 3; ++s) tails[s] = maps[0];
  if (a_tail_boxes(p)) {
    for (int s = 0; s < a->n_seg; ++s) {
      const int need = 128 * p.G + 2 * p.seg[s].halo, nbox = (need + TC_ABOX - 1) / TC_ABOX;
      const int tail = (int)align_up((size_t)(need - (nbox - 1) * TC_ABOX), 8);
      if (tail >= TC_ABOX) continue;
      const uint64_t dims[2] = {(uint64_t)a->src_c[s], (uint64_t)p.geo.rows_total};
      const uint64_t strides[1] = {(uint64_t)a->src_c[s] * 2};
      const uint32_t box[2] = {TC_BK, (uint32_t)tail};
      int rc = encode_bf16_map(&tails[s], a->src[s], 2, dims, strides, box);
      if (rc) return rc;
      p.seg[s].tail_rows = tail;
    }
  }
  for (int s = a->n_seg; s < 3 && !s2; ++s) maps[s] = maps[0];
  if (s2) { maps[0] = maps4[0]; maps[1] = maps4[1]; maps[2] = maps4[2]; } else maps4[3] = maps[0];
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
  p.dbg = g_tc_dbg;
  p.dbg_out = g_tc_dbg_out;
  VF_REQUIRE(!p.out_f32 || (a->out_ld % 4 == 0 && a->out_ld <= a->cout_pad && !a->residual), "vf_conv2d(tc): bad fp32 output layout");
  VF_REQUIRE(p.out_f32 || (a->cout % 16 == 0 && a->out_ld % 8 == 0), "vf_conv2d(tc): bf16 output needs cout %% 16 == 0");
  VF_REQUIRE(!a->qkv_split || (a->qkv_split % 16 == 0 && a->out_vt && !a->out_padded && !p.geo.stride2), "vf_conv2d(tc): bad qkv split");
  VF_REQUIRE(!a->stats || (!p.out_f32 && !a->qkv_split), "vf_conv2d(tc): fused statistics need a plain bf16 output");

  CUtensorMap mapOut = mapB, mapRes = mapB;
  if (p.epi_tma && lines) {
    int rc = encode_lines(&mapOut, a->out, a->cout, a->out_ld, 32);
    if (rc) return rc;
    if (a->residual) {
      rc = encode_lines(&mapRes, a->residual, a->cout, a->cout, 32);
      if (rc) return rc;
    }
  } else if (p.epi_tma) {
    const uint64_t out_rows = a->out_padded ? (uint64_t)a->images * (H + 1) * (W + 1) : (uint64_t)a->images * H * W;
    const uint32_t box[2] = {64, 32};
    {
      const uint64_t dims[2] = {(uint64_t)a->cout, out_rows};
      const uint64_t strides[1] = {(uint64_t)a->out_ld * 2};
      int rc = encode_bf16_map(&mapOut, a->out, 2, dims, strides, box);
      if (rc) return rc;
    }
    if (a->residual) {
      const uint64_t dims[2] = {(uint64_t)a->cout, out_rows};
      const uint64_t strides[1] = {(uint64_t)a->cout * 2};
      int rc = encode_bf16_map(&mapRes, a->residual, 2, dims, strides, box);
      if (rc) return rc;
    }
  }
  int grid = p.n_items < sm_count() ? p.n_items : sm_count();
  if (p.n_fixed > 0 && grid < p.n_fixed) grid = p.n_fixed;
  if (g_tc_dbg & 256)
    fprintf(stderr, "[vf tc] rows %d W %d cout %d segs %d ktot %d | bn %d G %d a_stages %d (%d B) b_stages %d resident %d smem %zu items %d grid %d epi_tma %d a_lines %d epi_lines %d\n",
            p.geo.rows_total, W, a->cout, a->n_seg, k_total, p.block_n, p.G, p.a_stages, p.a_stage_bytes, p.b_stages, p.b_resident, tl.smem,
            p.n_items, grid, p.epi_tma, p.a_lines, p.epi_lines);
  if (g_tc_dbg != 0 || g_tc_dbg_out != nullptr) {          // a test hook is armed: the instrumented instantiation
    VF_SET_MAX_SMEM(conv_tc_kernel<true>, kSmemBudget);
    VF_CUDA(launch_pdl(conv_tc_kernel<true>, dim3(grid), dim3(TC_THREADS), tl.smem, st, maps[0], maps[1], maps[2], maps4[3], mapB, mapOut, mapRes, tails[0], tails[1], tails[2], p));
  } else {
    VF_SET_MAX_SMEM(conv_tc_kernel<false>, kSmemBudget);
    VF_CUDA(launch_pdl(conv_tc_kernel<false>, dim3(grid), dim3(TC_THREADS), tl.smem, st, maps[0], maps[1], maps[2], maps4[3], mapB, mapOut, mapRes, tails[0], tails[1], tails[2], p));
  }
  VF_LAUNCH_CHECK();
  return VF_OK;
}

int conv2d_tc_plan(const vf_conv_args* a, int* out16) {
  g_tc_plan_out = out16;
  const int rc = conv2d_tc(a, nullptr);
  g_tc_plan_out = nullptr;
  return rc;
}

}  // namespace vf

// Test hook (host arithmetic only, usable without a device): the plan vf_conv2d's tcgen05 path would use for `a` (pointers
// are not dereferenced).  out[16] = block_n, G, a_stages, b_stages, weights resident, smem bytes, TMEM columns, work items,
// grid, staged epilogue, gathered-source lines, line-map epilogue, stride-2 gather chunks, GEMM rows, bias-table images, K.
extern "C" __attribute__((visibility("default"))) int vf_debug_conv_tiling(const vf_conv_args* a, int* out16) {
  using namespace vf;
  VF_REQUIRE(a && out16, "vf_debug_conv_tiling: null args");
  return conv2d_tc_plan(a, out16);
}
