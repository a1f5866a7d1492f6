// tcgen05 weight-gradient GEMM (reference: the dW half of conv autograd, experiment.py:292).
//
//   dWp[n][koff + tap*C + c] += sum_rows  X[row + shift(tap)][c] * dY[row][n]
//
// The reduction runs over pixel ROWS, so both operands are "MN-major" for the tensor core exactly as TMA loads them
// from the NHWC matrices (row = K, 64 channels = one 128-byte swizzle atom): no transposes anywhere.
//   A (M = 128) = two taps of the same 64-channel slab: atom 0 / atom 1 are the slab shifted by the taps' row offsets,
//                 i.e. the descriptor's leading-dimension offset is the difference of the two tap shifts
//   B (N = block_n <= 256) = dY tile, atoms of 64 output channels, one TMA box each
//   D = fp32 in TMEM, one accumulator per tap pair; after the CTA's row range it is added to the packed gradient
//       with coalesced fp32 reductions (lane == channel c, column == output channel n)
// Work item (blockIdx.y) = (segment, 64-channel chunk, group of tap pairs, N tile); blockIdx.x splits the rows.
// Both hardware facts used here were probed first (k_debug.cu): MN-major SWIZZLE_128B descriptors with LBO = atom
// stride / SBO = 1024, and start addresses shifted by whole rows.
#include <cuda.h>

#include <cstdlib>
#include <mutex>

#include "tc_ptx.cuh"
#include "vf_common.cuh"

namespace vf {

int encode_bf16_map(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box);

constexpr int WG_R = 128;          // rows per K chunk
constexpr int WG_THREADS = 192;    // warps 0-3 epilogue, 4 TMA, 5 MMA
constexpr int WG_MAX_ACC = 8;

struct WgSeg {
  int C, nchunks, ntaps, koff, halo;
  int npairs;        // accumulators needed per chunk: ceil(ntaps / 2)
  int ngroups;       // ceil(npairs / acc_max)
};

struct WgParams {
  int rows_total, W1;
  int n_seg;
  WgSeg seg[3];
  int block_n, n_tiles_n, acc_max, stages;
  int x_rows;              // slab rows per stage (multiple of 64)
  int stage_bytes, x_bytes;
  int k_total, cout;
  int rows_per_split;
  int tmem_cols;
  int order;               // MMA issue order, see the kernel
  float* dwp;
};

__global__ void __launch_bounds__(WG_THREADS, 1) conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap mapX0,
                                                                      const __grid_constant__ CUtensorMap mapX1,
                                                                      const __grid_constant__ CUtensorMap mapX2,
                                                                      const __grid_constant__ CUtensorMap mapDY, const WgParams p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const uint32_t bar_full = base, bar_empty = base + 64, bar_done = base + 128, tmem_slot = base + 136;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gbase + 136);
  const uint32_t ring = base + 1024;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // decode the work item
  // (row splits are the fast grid axis; putting the jobs that share a row range next to each other for L2 reuse was
  // measured 13 % slower)
  int job = blockIdx.y, seg = 0;
  for (; seg < p.n_seg; ++seg) {
    const int per = p.seg[seg].nchunks * p.seg[seg].ngroups * p.n_tiles_n;
    if (job < per) break;
    job -= per;
  }
  const WgSeg sg = p.seg[seg];
  const int n_tile = job % p.n_tiles_n;
  const int grp = (job / p.n_tiles_n) % sg.ngroups;
  const int chunk = job / (p.n_tiles_n * sg.ngroups);
  const int pair0 = grp * p.acc_max;
  const int npair = min(p.acc_max, sg.npairs - pair0);
  const int n0 = n_tile * p.block_n;
  const int r_begin = blockIdx.x * p.rows_per_split;
  const int r_end = min(p.rows_total, r_begin + p.rows_per_split);
  const int nk = r_begin < r_end ? (r_end - r_begin + WG_R - 1) / WG_R : 0;
  const int nbox_x = (WG_R + 2 * sg.halo + 63) / 64;
  const int natom_n = (p.block_n + 63) / 64;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { ptx::mbar_init(bar_full + 8 * s, 1); ptx::mbar_init(bar_empty + 8 * s, 1); }
    ptx::mbar_init(bar_done, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 5) { ptx::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();

  if (warp == 4) {
    if (lane == 0 && nk > 0) {
      const CUtensorMap* mX = seg == 0 ? &mapX0 : (seg == 1 ? &mapX1 : &mapX2);
      int st = 0;
      uint32_t par = 1;
      for (int it = 0; it < nk; ++it) {
        ptx::mbar_wait(bar_empty + 8 * st, par);
        const uint32_t fb = bar_full + 8 * st;
        ptx::mbar_arrive_expect_tx(fb, (uint32_t)nbox_x * 64 * 128 + (uint32_t)natom_n * WG_R * 128);
        const uint32_t sx = ring + (uint32_t)st * (uint32_t)p.stage_bytes;
        const uint32_t sy = sx + (uint32_t)p.x_bytes;
        const int row0 = r_begin + it * WG_R;
        for (int b = 0; b < nbox_x; ++b) ptx::tma_load_2d(sx + b * 8192, mX, fb, chunk * 64, row0 - sg.halo + b * 64);
        for (int a = 0; a < natom_n; ++a) ptx::tma_load_2d(sy + a * (WG_R * 128), &mapDY, fb, n0 + a * 64, row0);
        if (++st == p.stages) { st = 0; par ^= 1u; }
      }
    }
  } else if (warp == 5) {
    if (nk > 0 && ptx::elect_one()) {
      // Everything that does not depend on the pipeline stage is computed once: the per-pair A descriptors (tap shift
      // in the start address, shift difference as the leading-dimension offset) and the dY descriptor.  Per chunk the
      // thread then only adds the stage offset and the k-step (16 rows = 2048 bytes) to the start-address fields, so the
      // tensor pipe is not left waiting for scalar code between MMAs.
      const uint32_t idesc = ptx::make_idesc_bf16(128, p.block_n, 1, 1);       // A and B MN-major
      uint64_t adp[WG_MAX_ACC];
#pragma unroll
      for (int a = 0; a < WG_MAX_ACC; ++a) {
        const int t0 = min(2 * (pair0 + a), sg.ntaps - 1), t1 = min(t0 + 1, sg.ntaps - 1);
        const int sh0 = sg.ntaps == 9 ? (t0 / 3) * p.W1 + (t0 % 3) : 0;         // slab starts `halo` rows before the chunk
        const int sh1 = sg.ntaps == 9 ? (t1 / 3) * p.W1 + (t1 % 3) : 0;
        const uint32_t lbo = t1 > t0 ? (uint32_t)(sh1 - sh0) * 128u : 128u;     // single tap: atom 1 is a harmless neighbour
        adp[a] = ptx::make_smem_desc(ring + (uint32_t)sh0 * 128u, lbo, 1024);
      }
      const uint64_t bd0 = ptx::make_smem_desc(ring + (uint32_t)p.x_bytes, WG_R * 128, 1024);
      const uint32_t st_step = (uint32_t)p.stage_bytes >> 4;
      const uint32_t block_n = (uint32_t)p.block_n;
      int st = 0;
      uint32_t par = 0, st_off = 0, acc = 0;
      // splits are multiples of WG_R rows; rows past the end of the tensors are zero-filled by TMA
      for (int it = 0; it < nk; ++it) {
        ptx::mbar_wait(bar_full + 8 * st, par);
        ptx::tc_fence_after();
        // MMA order (p.order, test hook): 0 = accumulator outer / 8 k-steps inner, 1 = k-step outer / accumulator inner,
        // 2 = chains of four k-steps per accumulator
        const uint64_t bd = bd0 + st_off;
        if (p.order == 1) {
#pragma unroll
          for (int k = 0; k < WG_R / 16; ++k) {
#pragma unroll
            for (int a = 0; a < WG_MAX_ACC; ++a) {
              if (a < npair)
                ptx::umma_f16(tmem_base + (uint32_t)a * block_n, adp[a] + st_off + (uint64_t)(k * 128), bd + (uint64_t)(k * 128), idesc,
                              k > 0 ? 1u : acc);
            }
          }
        } else if (p.order == 2) {
#pragma unroll
          for (int kb = 0; kb < WG_R / 16; kb += 4) {
#pragma unroll
            for (int a = 0; a < WG_MAX_ACC; ++a) {
              if (a < npair) {
#pragma unroll
                for (int k = kb; k < kb + 4; ++k)
                  ptx::umma_f16(tmem_base + (uint32_t)a * block_n, adp[a] + st_off + (uint64_t)(k * 128), bd + (uint64_t)(k * 128), idesc,
                                k > 0 ? 1u : acc);
              }
            }
          }
        } else {
#pragma unroll
          for (int a = 0; a < WG_MAX_ACC; ++a) {
            if (a < npair) {
#pragma unroll
              for (int k = 0; k < WG_R / 16; ++k)
                ptx::umma_f16(tmem_base + (uint32_t)a * block_n, adp[a] + st_off + (uint64_t)(k * 128), bd + (uint64_t)(k * 128), idesc,
                              k > 0 ? 1u : acc);
            }
          }
        }
        ptx::umma_commit(bar_empty + 8 * st);
        acc = 1;
        if (++st == p.stages) { st = 0; par ^= 1u; st_off = 0; } else st_off += st_step;
      }
      ptx::umma_commit(bar_done);
    }
  } else if (nk > 0) {
    // epilogue: lane quarter q holds accumulator rows 32q..32q+31 = (atom, channel)
    ptx::mbar_wait(bar_done, 0);
    ptx::tc_fence_after();
    const int m = warp * 32 + lane;
    const int atom = m >> 6, c = m & 63;
    for (int a = 0; a < npair; ++a) {
      const int tap = 2 * (pair0 + a) + atom;
      const bool live = tap < sg.ntaps;
      float* dst = p.dwp + (size_t)sg.koff + (size_t)tap * sg.C + chunk * 64 + c;
      for (int c0 = 0; c0 < p.block_n; c0 += 16) {
        uint32_t rr[16];
        ptx::tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * p.block_n + c0), rr);
        ptx::tmem_ld_wait();
        if (live) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int n = n0 + c0 + j;
            if (n < p.cout) atomicAdd(dst + (size_t)n * p.k_total, __uint_as_float(rr[j]));
          }
        }
      }
    }
    ptx::tc_fence_before();
  }
  __syncthreads();
  if (warp == 5) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols); }
}

// ---------------------------------------------------------------------------------------------------------------------
// Cout <= 64 variant.  With N = 64 the MN-major MMA above costs ~81 clk for 32 clk of math (the 64x64-resolution layers:
// 35 % of the model's FLOPs ran at 36 % of the tensor peak).  Here the three KERNEL ROWS ride in N instead:
//   dW[n][kh][kw][c] = sum_r  X[r + kw - 1][c] * dY[r - (kh - 1) * W1][n]          (r = padded row index)
//   A (M = 128) = X slab at two column taps (kw, kw + 1): atom 1 is the slab one row further (LBO = 128 B)
//   B (N = 192) = dY slab at the three row taps kh = 2, 1, 0: atoms W1 rows apart (LBO = W1 * 128 B), 64 output channels each
// Two MMAs per 16-row k-step (kw = 0,1 and kw = 2 + an idle half) cover all nine taps: 192 clk instead of 5 x 81, and every
// (chunk of 64 input channels) is ONE job holding its nine taps in two 192-column accumulators.
// ---------------------------------------------------------------------------------------------------------------------
struct Wg64Params {
  int rows_total, W1;
  int C, koff;             // segment channels, first weight column of the segment
  int stages, x_rows, y_rows, stage_bytes, x_bytes;
  int k_total, cout;
  int rows_per_split;
  float* dwp;
};

__global__ void __launch_bounds__(WG_THREADS, 1) conv_wgrad64_tc_kernel(const __grid_constant__ CUtensorMap mapX,
                                                                        const __grid_constant__ CUtensorMap mapDY, const Wg64Params p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const uint32_t bar_full = base, bar_empty = base + 64, bar_done = base + 128, tmem_slot = base + 136;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gbase + 136);
  const uint32_t ring = base + 1024;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = blockIdx.y;
  const int r_begin = blockIdx.x * p.rows_per_split;
  const int r_end = min(p.rows_total, r_begin + p.rows_per_split);
  const int nk = r_begin < r_end ? (r_end - r_begin + WG_R - 1) / WG_R : 0;
  constexpr uint32_t kTmemCols = 512;          // two accumulators of 192 columns

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { ptx::mbar_init(bar_full + 8 * s, 1); ptx::mbar_init(bar_empty + 8 * s, 1); }
    ptx::mbar_init(bar_done, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 5) { ptx::tmem_alloc(tmem_slot, kTmemCols); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();

  if (warp == 4) {
    if (lane == 0 && nk > 0) {
      int st = 0;
      uint32_t par = 1;
      const int nbx = p.x_rows / 64, nby = p.y_rows / 64;
      for (int it = 0; it < nk; ++it) {
        ptx::mbar_wait(bar_empty + 8 * st, par);
        const uint32_t fb = bar_full + 8 * st;
        ptx::mbar_arrive_expect_tx(fb, (uint32_t)(nbx + nby) * 64 * 128);
        const uint32_t sx = ring + (uint32_t)st * (uint32_t)p.stage_bytes;
        const uint32_t sy = sx + (uint32_t)p.x_bytes;
        const int row0 = r_begin + it * WG_R;
        for (int b = 0; b < nbx; ++b) ptx::tma_load_2d(sx + b * 8192, &mapX, fb, chunk * 64, row0 - 1 + b * 64);          // X rows r0-1 ..
        for (int b = 0; b < nby; ++b) ptx::tma_load_2d(sy + b * 8192, &mapDY, fb, 0, row0 - p.W1 + b * 64);               // dY rows r0-W1 ..
        if (++st == p.stages) { st = 0; par ^= 1u; }
      }
    }
  } else if (warp == 5) {
    if (nk > 0 && ptx::elect_one()) {
      const uint32_t idesc = ptx::make_idesc_bf16(128, 192, 1, 1);                      // A and B MN-major
      // A: atoms = column taps (kw, kw + 1), one row (128 B) apart.  accumulator 0: kw = 0, 1; accumulator 1: kw = 2 (+ idle atom)
      const uint64_t ad0 = ptx::make_smem_desc(ring, 128, 1024);
      const uint64_t ad1 = ptx::make_smem_desc(ring + 2u * 128u, 128, 1024);
      // B: atoms = row taps kh = 2, 1, 0 of the dY slab (it starts W1 rows before the chunk), W1 rows apart
      const uint64_t bd0 = ptx::make_smem_desc(ring + (uint32_t)p.x_bytes, (uint32_t)p.W1 * 128u, 1024);
      const uint32_t st_step = (uint32_t)p.stage_bytes >> 4;
      int st = 0;
      uint32_t par = 0, st_off = 0, acc = 0;
      for (int it = 0; it < nk; ++it) {
        ptx::mbar_wait(bar_full + 8 * st, par);
        ptx::tc_fence_after();
#pragma unroll
        for (int k = 0; k < WG_R / 16; ++k) {
          ptx::umma_f16(tmem_base, ad0 + st_off + (uint64_t)(k * 128), bd0 + st_off + (uint64_t)(k * 128), idesc, k > 0 ? 1u : acc);
          ptx::umma_f16(tmem_base + 192u, ad1 + st_off + (uint64_t)(k * 128), bd0 + st_off + (uint64_t)(k * 128), idesc, k > 0 ? 1u : acc);
        }
        ptx::umma_commit(bar_empty + 8 * st);
        acc = 1;
        if (++st == p.stages) { st = 0; par ^= 1u; st_off = 0; } else st_off += st_step;
      }
      ptx::umma_commit(bar_done);
    }
  } else if (nk > 0) {
    // epilogue: lane = (kw half, channel c); column = (2 - kh) * 64 + n
    ptx::mbar_wait(bar_done, 0);
    ptx::tc_fence_after();
    const int m = warp * 32 + lane;
    const int half = m >> 6, c = m & 63;
    for (int a = 0; a < 2; ++a) {
      const int kw = 2 * a + half;
      const bool live = kw < 3;
      for (int j = 0; j < 3; ++j) {
        const int kh = 2 - j;
        float* dst = p.dwp + (size_t)p.koff + (size_t)(kh * 3 + kw) * p.C + chunk * 64 + c;
        for (int c0 = 0; c0 < 64; c0 += 16) {
          uint32_t rr[16];
          ptx::tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * 192 + j * 64 + c0), rr);
          ptx::tmem_ld_wait();
          if (live) {
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              const int n = c0 + q;
              if (n < p.cout) atomicAdd(dst + (size_t)n * p.k_total, __uint_as_float(rr[q]));
            }
          }
        }
      }
    }
    ptx::tc_fence_before();
  }
  __syncthreads();
  if (warp == 5) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, kTmemCols); }
}

// One launch for one 3x3 segment of a Cout <= 64 layer (dy_ld == 64): grid (row splits, 64-channel chunks).
static int wgrad64_segment(const void* x, int C, int koff, int rows_total, int W, const void* dy, int cout, int k_total, float* dwp,
                           cudaStream_t st) {
  Wg64Params p{};
  p.rows_total = rows_total; p.W1 = W + 1; p.C = C; p.koff = koff; p.k_total = k_total; p.cout = cout; p.dwp = dwp;
  p.x_rows = (int)align_up((size_t)WG_R + 2, 64);
  p.y_rows = (int)align_up((size_t)WG_R + 2 * p.W1, 64);
  p.x_bytes = p.x_rows * 128;
  p.stage_bytes = p.x_bytes + p.y_rows * 128;
  p.stages = (int)((227 * 1024 - 2048) / p.stage_bytes);
  if (p.stages > 4) p.stages = 4;
  VF_REQUIRE(p.stages >= 2, "vf_conv2d_wgrad(tc64): stage of %d B does not fit twice", p.stage_bytes);
  const int jobs = C / 64;
  int splits = cdiv(2 * sm_count(), jobs);
  const int max_splits = cdiv(rows_total, 8 * WG_R);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  p.rows_per_split = (int)align_up((size_t)cdiv(rows_total, splits), WG_R);
  splits = cdiv(rows_total, p.rows_per_split);
  CUtensorMap mapX, mapDY;
  {
    const uint64_t dims[2] = {(uint64_t)C, (uint64_t)rows_total};
    const uint64_t strides[1] = {(uint64_t)C * 2};
    const uint32_t box[2] = {64, 64};
    int rc = encode_bf16_map(&mapX, x, 2, dims, strides, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {64, (uint64_t)rows_total};
    const uint64_t strides[1] = {64 * 2};
    const uint32_t box[2] = {64, 64};
    int rc = encode_bf16_map(&mapDY, dy, 2, dims, strides, box);
    if (rc) return rc;
  }
  VF_SET_MAX_SMEM(conv_wgrad64_tc_kernel, 227 * 1024);
  const size_t smem = 2048 + (size_t)p.stages * p.stage_bytes;
  VF_CUDA(launch_pdl(conv_wgrad64_tc_kernel, dim3(splits, jobs), dim3(WG_THREADS), smem, st, mapX, mapDY, p));
  VF_LAUNCH_CHECK();
  return VF_OK;
}

// Supported: bf16, stride 1, X and dY in the same row order (3x3 taps need PADDED), every segment a multiple of 64
// channels, dY leading dimension a multiple of 64 (columns >= cout are skipped by the epilogue).
bool wgrad_tc_supported(const vf_conv_args* a, int dy_ld) {
  if (a->dtype != VF_BF16 || a->stride != 1 || (a->in_padded != 0) != (a->out_padded != 0)) return false;
  if (dy_ld % 64) return false;
  for (int s = 0; s < a->n_seg; ++s) {
    if (a->src_c[s] % 64) return false;
    if (a->ksize[s] == 3 && !a->in_padded) return false;
  }
  return true;
}

int tc_debug_flags();

static int conv2d_wgrad_tc_generic(const vf_conv_args* a, const void* dy, int dy_ld, float* dwp, cudaStream_t st, int seg_mask, int k_total_all);

int conv2d_wgrad_tc(const vf_conv_args* a, const void* dy, int dy_ld, float* dwp, cudaStream_t st) {
  static const bool use64 = [] { const char* e = getenv("VF_WG64"); return !(e && e[0] == '0'); }();
  int k_total = 0;
  for (int s = 0; s < a->n_seg; ++s) k_total += a->ksize[s] * a->ksize[s] * a->src_c[s];
  if (use64 && dy_ld == 64 && a->in_padded && a->out_padded) {
    // Cout <= 64: the 3x3 segments take the N = 192 kernel (kernel rows in N), the 1x1 segments the generic one
    const int rows_total = a->images * (a->H + 1) * (a->W + 1);
    int koff = 0, rest = 0;
    for (int s = 0; s < a->n_seg; ++s) {
      if (a->ksize[s] == 3) {
        const int rc = wgrad64_segment(a->src[s], a->src_c[s], koff, rows_total, a->W, dy, a->cout, k_total, dwp, st);
        if (rc != VF_OK) return rc;
      } else {
        rest |= 1 << s;
      }
      koff += a->ksize[s] * a->ksize[s] * a->src_c[s];
    }
    if (!rest) return VF_OK;
    return conv2d_wgrad_tc_generic(a, dy, dy_ld, dwp, st, rest, k_total);
  }
  return conv2d_wgrad_tc_generic(a, dy, dy_ld, dwp, st, (1 << a->n_seg) - 1, k_total);
}

// seg_mask: the segments this launch covers (the others keep their place in the K layout but produce no jobs)
static int conv2d_wgrad_tc_generic(const vf_conv_args* a, const void* dy, int dy_ld, float* dwp, cudaStream_t st, int seg_mask, int k_total_all) {
  WgParams p{};
  p.order = (tc_debug_flags() >> 28) & 3;
  const int H = a->H, W = a->W;
  p.rows_total = a->in_padded ? a->images * (H + 1) * (W + 1) : a->images * H * W;
  p.W1 = W + 1;
  p.cout = a->cout;
  p.dwp = dwp;
  // N tile: MN-major MMAs cost ~81 clk up to N = 160 and N/2 beyond (scripts/probe_rate.py), so the widest tile that
  // divides the (64-padded) output channels wins: n_cols / t, a multiple of 16, at most 256 (320 -> 160, 576 -> 192)
  const int n_cols = (int)align_up((size_t)a->cout, 64);
  p.block_n = 64;
  for (int t = 1; t <= n_cols / 16; ++t)
    if (n_cols % t == 0 && (n_cols / t) % 16 == 0 && n_cols / t <= 256) { p.block_n = n_cols / t; break; }
  p.n_tiles_n = n_cols / p.block_n;
  p.acc_max = 512 / p.block_n;
  if (p.acc_max > WG_MAX_ACC) p.acc_max = WG_MAX_ACC;
  // Few rows (the 16x16 and 8x8 levels): the reduction is short and the output large, so parallelism should come from
  // output tiles (one tap pair per job) rather than from row splits, whose partial sums all go through fp32 REDs
  if (p.rows_total < 65536) {
    static const int env_acc = [] { const char* e = getenv("VF_WG_ACC_SMALL"); return e ? atoi(e) : 0; }();     // A/B knob
    const int small_acc = env_acc > 0 ? env_acc : 1;
    if (p.acc_max > small_acc) p.acc_max = small_acc;
  }
  p.n_seg = a->n_seg;
  int k_total = 0, halo_max = 0, jobs = 0, max_pairs = 0;
  for (int s = 0; s < a->n_seg; ++s) {
    WgSeg& sg = p.seg[s];
    sg.C = a->src_c[s]; sg.nchunks = sg.C / 64; sg.ntaps = a->ksize[s] * a->ksize[s]; sg.koff = k_total;
    sg.halo = a->ksize[s] == 3 ? W + 2 : 0;
    sg.npairs = (sg.ntaps + 1) / 2;
    sg.ngroups = (sg.npairs + p.acc_max - 1) / p.acc_max;
    k_total += sg.ntaps * sg.C;
    if (!((seg_mask >> s) & 1)) { sg.nchunks = 0; continue; }     // covered by another launch: no jobs, but its K columns stay in place
    halo_max = sg.halo > halo_max ? sg.halo : halo_max;
    jobs += sg.nchunks * sg.ngroups * p.n_tiles_n;
    const int mp = sg.npairs < p.acc_max ? sg.npairs : p.acc_max;
    max_pairs = mp > max_pairs ? mp : max_pairs;
  }
  p.k_total = k_total;
  p.x_rows = (int)align_up((size_t)WG_R + 2 * halo_max, 64);
  p.x_bytes = p.x_rows * 128;
  p.stage_bytes = p.x_bytes + (p.block_n + 63) / 64 * WG_R * 128;
  p.stages = (int)((227 * 1024 - 2048) / p.stage_bytes);
  if (p.stages > 4) p.stages = 4;
  VF_REQUIRE(p.stages >= 2, "vf_conv2d_wgrad(tc): stage of %d B does not fit twice", p.stage_bytes);
  p.tmem_cols = 32;
  while (p.tmem_cols < max_pairs * p.block_n) p.tmem_cols *= 2;
  // row splits: enough CTAs for ~2 waves, each at least 8 chunks deep
  int splits = cdiv(2 * sm_count(), jobs);
  const int max_splits = cdiv(p.rows_total, 8 * WG_R);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  p.rows_per_split = (int)align_up((size_t)cdiv(p.rows_total, splits), WG_R);
  splits = cdiv(p.rows_total, p.rows_per_split);

  CUtensorMap maps[3], mapDY;
  for (int s = 0; s < a->n_seg; ++s) {
    const uint64_t dims[2] = {(uint64_t)a->src_c[s], (uint64_t)p.rows_total};
    const uint64_t strides[1] = {(uint64_t)a->src_c[s] * 2};
    const uint32_t box[2] = {64, 64};
    int rc = encode_bf16_map(&maps[s], a->src[s], 2, dims, strides, box);
    if (rc) return rc;
  }
  for (int s = a->n_seg; s < 3; ++s) maps[s] = maps[0];
  {
    const uint64_t dims[2] = {(uint64_t)dy_ld, (uint64_t)p.rows_total};
    const uint64_t strides[1] = {(uint64_t)dy_ld * 2};
    const uint32_t box[2] = {64, WG_R};
    int rc = encode_bf16_map(&mapDY, dy, 2, dims, strides, box);
    if (rc) return rc;
  }
  VF_SET_MAX_SMEM(conv_wgrad_tc_kernel, 227 * 1024);
  const size_t smem = 2048 + (size_t)p.stages * p.stage_bytes;
  dim3 grid(splits, jobs);
  VF_CUDA(launch_pdl(conv_wgrad_tc_kernel, grid, dim3(WG_THREADS), smem, st, maps[0], maps[1], maps[2], mapDY, p));
  VF_LAUNCH_CHECK();
  return VF_OK;
}

}  // namespace vf
