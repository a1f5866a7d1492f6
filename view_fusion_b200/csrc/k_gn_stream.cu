// GroupNorm(+Swish) forward / backward as PERSISTENT, TMA-fed streaming kernels (bulk copies into a shared-memory ring).
// Reference: nn.GroupNorm + Swish of model/unet.py:207-218, :254 and their autograd (experiment.py:292).
//
// Why: the per-(image, channel) coefficients need a statistics prologue (a global round trip, two barriers, the per-group
// loops: ~3 us).  The first version launched one short-lived CTA per (image, row split) — 2000+ CTAs per layer, each paying the
// prologue with nothing in flight and each holding only a few 16-byte loads per thread: 57 % of the HBM copy rate on the
// large layers, 20-40 % on the small ones.  Here a CTA is resident for the whole launch (grid = SMs x 2), walks a contiguous
// range of (image, row tile) work items, and ONE elected thread of a producer warp keeps a ring of bulk copies
// (cp.async.bulk, 24-32 KB per stage) in flight — memory-level parallelism no longer depends on registers or on the number
// of CTAs, and the prologue of an image runs while its tiles are already landing.  A tile is a run of whole PADDED rows of
// one image, which is one contiguous chunk of the NHWC tensor, so no tensor map is needed.
//
//   mode FWD : y = swish?(x*a + b)                      reads x (two sources = the never-materialised torch.cat), writes y
//   mode RED : per (image, channel)  A = sum dz, B = sum dz*xhat, dz = dy * swish'(z)      reads x, dy; writes nothing
//   mode APP : dx = rstd*(gamma*dz - S1/n - xhat*S2/n)  (+= old)                            reads x, dy (+ old dx); writes dx
// dz is RECOMPUTED in APP instead of being written by RED and read back (10 instead of 12 bytes per element, and dy is
// no longer consumed).  Algorithmic bytes (bf16): FWD 4, RED 4, APP 6 (+2 when accumulating) per element.
#include <cstdlib>

#include "tc_ptx.cuh"
#include "vf_common.cuh"

namespace vf {

constexpr int GS_CONSUMERS = 256;
constexpr int GS_THREADS = GS_CONSUMERS + 32;      // + one producer warp
constexpr int GS_MAX_STAGES = 4;
#ifndef VF_GS_MIN_CTAS
#define VF_GS_MIN_CTAS 2
#endif
#ifndef VF_GS_UN
#define VF_GS_UN 4
#endif
constexpr int GS_MIN_CTAS = VF_GS_MIN_CTAS;       // resident CTAs per SM the kernel is compiled for (register cap)
enum { GS_FWD = 0, GS_RED = 1, GS_APP = 2 };

struct GsParams {
  const uint8_t* s0; const uint8_t* s1; int C0, C1;
  const uint8_t* dy;
  uint8_t* dst;
  uint8_t* dx0; uint8_t* dx1; int acc0, acc1;
  const float* st0; int ld0; const float* st1; int ld1;
  const float* gamma; const float* beta;
  float* red; float* dgamma; float* dbeta;
  vf_gn_shift sh;
  GnColsum cs;
  int images, HW, W1, P, groups, swish;
  int R, tiles_per_img, n_items, items_per_cta, CV, PY, stages;
  uint32_t tile0_bytes, tile1_bytes, tiled_bytes, stage_bytes, table_bytes;
};

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(bar)
               : "memory");
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(GS_CONSUMERS) : "memory"); }

template <typename T> __device__ __forceinline__ float gs_dz(float g, float h);      // g * swish'(2h)
template <> __device__ __forceinline__ float gs_dz<float>(float g, float h) {
  const float t = tanhf(h);
  const float q = fmaf(h, fmaf(-t, t, 1.f), t);
  const float gh = 0.5f * g;
  return fmaf(gh, q, gh);
}
template <> __device__ __forceinline__ float gs_dz<__nv_bfloat16>(float g, float h) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  const float q = fmaf(h, fmaf(-t, t, 1.f), t);
  const float gh = 0.5f * g;
  return fmaf(gh, q, gh);
}

// per-channel (mean - shift, rstd) of the image into mr[2C]; raw[2C] is scratch.  Consumer threads only (t < GS_CONSUMERS).
__device__ __forceinline__ void gs_group_stats(const GsParams& p, int img, int t, float* mr, float* raw) {
  const int C = p.C0 + p.C1, gs = C / p.groups;
  const float inv_n = 1.f / ((float)gs * (float)p.HW);
  const float* sa = p.st0 + (size_t)img * p.ld0 * 2;
  const float* sb = p.st1 ? p.st1 + (size_t)img * p.ld1 * 2 : nullptr;
  const bool shifted = p.sh.bias != nullptr || p.sh.emb != nullptr;
  const float* sh_emb = p.sh.emb ? p.sh.emb + (size_t)__ldg(p.sh.img_row + img) * p.sh.emb_ld : nullptr;
  for (int ch = t; ch < C; ch += GS_CONSUMERS) {
    float S1, S2;
    if (ch < p.C0) {
      S1 = __ldg(sa + 2 * ch); S2 = __ldg(sa + 2 * ch + 1);
      if (shifted) {        // the source is stored without s = bias + emb: shift its raw sums in closed form (vf_gn_shift)
        const float sv = gn_shift_value(p.sh, sh_emb, ch);
        S2 = fmaf(2.f * sv, S1, S2) + (float)p.HW * sv * sv;
        S1 = fmaf((float)p.HW, sv, S1);
      }
    } else {
      S1 = __ldg(sb + 2 * (ch - p.C0)); S2 = __ldg(sb + 2 * (ch - p.C0) + 1);
    }
    raw[2 * ch] = S1; raw[2 * ch + 1] = S2;
  }
  consumer_sync();
  for (int ch = t; ch < C; ch += GS_CONSUMERS) {
    const int g0 = ch / gs * gs;
    float s = 0.f, q = 0.f;
    for (int j = 0; j < gs; ++j) { s += raw[2 * (g0 + j)]; q += raw[2 * (g0 + j) + 1]; }     // a group may straddle the two sources
    float mean = s * inv_n;
    const float var = fmaxf(q * inv_n - mean * mean, 0.f);
    if (shifted && ch < p.C0) mean -= gn_shift_value(p.sh, sh_emb, ch);                     // (x + s - mean) = x - (mean - s)
    mr[2 * ch] = mean;
    mr[2 * ch + 1] = rsqrtf(var + 1e-5f);
  }
  consumer_sync();
}

template <typename T, int MODE, bool kSwish>
__global__ void __launch_bounds__(GS_THREADS, GS_MIN_CTAS) gn_stream_kernel(const GsParams p) {
  constexpr int VEC = VecOf<T>::N;
  constexpr int ES = (int)sizeof(T);
  constexpr int UN = MODE == GS_APP ? 2 : VF_GS_UN;    // rows per thread and batch: independent 16-byte shared-memory loads (APP holds 3 tensors)
  pdl_launch_dependents();
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sbase = ptx::smem_u32(smem);
  const uint32_t bar_full = sbase, bar_empty = sbase + 64;
  float* tab = reinterpret_cast<float*>(smem + 128);
  uint8_t* ring = smem + 128 + p.table_bytes;
  const uint32_t ring_u32 = sbase + 128 + p.table_bytes;
  const int C = p.C0 + p.C1;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  // balanced contiguous ranges: every resident CTA slot gets floor or ceil of n_items / grid work items
  const int it0 = (int)((long long)blockIdx.x * p.n_items / gridDim.x);
  const int it1 = (int)((long long)(blockIdx.x + 1) * p.n_items / gridDim.x);
  const int S = p.stages;

  if (t == 0) {
    for (int s = 0; s < S; ++s) { ptx::mbar_init(bar_full + 8 * s, 1); ptx::mbar_init(bar_empty + 8 * s, GS_CONSUMERS / 32); }
    ptx::fence_barrier_init();
  }
  __syncthreads();
  pdl_wait();                                   // sources, statistics and gradients come from the previous kernels

  if (warp == GS_CONSUMERS / 32) {
    // ===================== producer: one thread keeps the ring of bulk copies full =====================
    if (lane == 0) {
      int s = 0;
      uint32_t par = 1;                         // "empty" barriers: the first pass over the ring does not wait
      for (int it = it0; it < it1; ++it) {
        const int img = it / p.tiles_per_img, tile = it - img * p.tiles_per_img;
        const int r0 = tile * p.R, nrows = min(p.R, p.P - r0);
        const size_t row0 = (size_t)img * p.P + r0;
        ptx::mbar_wait(bar_empty + 8 * s, par);
        const uint32_t b0 = (uint32_t)nrows * p.C0 * ES, b1 = (uint32_t)nrows * p.C1 * ES, bd = (uint32_t)nrows * C * ES;
        const bool old0 = MODE == GS_APP && p.acc0, old1 = MODE == GS_APP && p.acc1 && p.C1;
        const uint32_t total = b0 + b1 + (MODE != GS_FWD ? bd : 0u) + (old0 ? b0 : 0u) + (old1 ? b1 : 0u);
        const uint32_t fb = bar_full + 8 * s;
        ptx::mbar_arrive_expect_tx(fb, total);
        uint32_t dst = ring_u32 + (uint32_t)s * p.stage_bytes;
        bulk_g2s(dst, p.s0 + row0 * p.C0 * ES, b0, fb);
        dst += p.tile0_bytes;
        if (p.C1) bulk_g2s(dst, p.s1 + row0 * p.C1 * ES, b1, fb);
        dst += p.tile1_bytes;
        if (MODE != GS_FWD) {
          bulk_g2s(dst, p.dy + row0 * C * ES, bd, fb);
          dst += p.tiled_bytes;
          if (old0) bulk_g2s(dst, p.dx0 + row0 * p.C0 * ES, b0, fb);
          dst += p.tile0_bytes;
          if (old1) bulk_g2s(dst, p.dx1 + row0 * p.C1 * ES, b1, fb);
        }
        if (++s == S) { s = 0; par ^= 1u; }
      }
    }
    return;
  }

  // ===================== consumers: 8 warps; thread = (16-byte channel vector cv, row lane py) =====================
  const int NT = p.CV * p.PY;
  const bool active = t < NT;
  const int cv = active ? t % p.CV : 0, py = active ? t / p.CV : 0;
  const int c = cv * VEC;
  const bool first = c < p.C0;
  const int ld = first ? p.C0 : p.C1;                 // row length of this thread's source
  const int cl = first ? c : c - p.C0;                // channel inside it
  float* mr = tab;                                     // [C][2]
  float* tb2 = tab + 2 * C;                            // FWD: ab [C][2];  RED: part [PY][C][2];  APP: tt [C][2]
  float* raw = MODE == GS_RED ? tab + 2 * C + 2 * C * p.PY : tab + 4 * C;     // [C][2] scratch
  float k0[VEC], k1[VEC], k2[VEC], k3[VEC], k4[VEC];   // per-thread per-channel constants (meaning depends on MODE)
  float gam[VEC], bet[VEC];                            // the thread's gamma / beta: constant over the launch, fetched once
#pragma unroll
  for (int j = 0; j < VEC; ++j) { gam[j] = __ldg(p.gamma + c + j); bet[j] = __ldg(p.beta + c + j); }
  float sA[VEC], sB[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) { k0[j] = k1[j] = k2[j] = k3[j] = k4[j] = 0.f; sA[j] = sB[j] = 0.f; }

  auto flush_red = [&](int img) {              // RED: block-reduce the row lanes' sums of image `img` and add them to p.red
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      if (active) {
        const float mu = mr[2 * (c + j)], rs = mr[2 * (c + j) + 1];
        tb2[(py * C + c + j) * 2] = sA[j];
        tb2[(py * C + c + j) * 2 + 1] = rs * (sB[j] - mu * sA[j]);       // sum(dz * xhat)
      }
      sA[j] = sB[j] = 0.f;
    }
    consumer_sync();
    float* red = p.red + (size_t)img * C * 2;
    for (int i = t; i < 2 * C; i += GS_CONSUMERS) {
      float v = 0.f;
      for (int q = 0; q < p.PY; ++q) v += tb2[q * 2 * C + i];
      atomicAdd(red + i, v);
    }
    consumer_sync();
  };

  int cur_img = -1;
  int s = 0;
  uint32_t par = 0;
  for (int it = it0; it < it1; ++it) {
    const int img = it / p.tiles_per_img, tile = it - img * p.tiles_per_img;
    const int r0 = tile * p.R, nrows = min(p.R, p.P - r0);
    if (img != cur_img) {
      // ---------------- per-image prologue (its tiles are already in flight) ----------------
      if (MODE == GS_RED && cur_img >= 0) flush_red(cur_img);
      cur_img = img;
      gs_group_stats(p, img, t, mr, raw);
      if (MODE == GS_FWD) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const float a = mr[2 * (c + j) + 1] * gam[j];
          k0[j] = a;
          k1[j] = bet[j] - mr[2 * (c + j)] * a;          // y = x*a + b
        }
      } else {
        if (MODE == GS_APP) {
          // gamma-weighted group sums of the reduce pass -> t1 = rstd*S1/n, t2 = rstd*S2/n; the CTA that owns the image's
          // first tile also folds the per-image sums into dgamma / dbeta (and the closed-form column sums of dx)
          const int gs = C / p.groups;
          const float inv_n = 1.f / ((float)gs * (float)p.HW);
          const float* red = p.red + (size_t)img * C * 2;
          const bool owner = tile == 0;
          for (int i = t; i < 2 * C; i += GS_CONSUMERS) {
            const float r = __ldg(red + i);
            raw[i] = r * __ldg(p.gamma + (i >> 1));
            if (owner) atomicAdd(((i & 1) ? p.dgamma : p.dbeta) + (i >> 1), r);
          }
          consumer_sync();
          for (int ch = t; ch < C; ch += GS_CONSUMERS) {
            const int g0 = ch / gs * gs;
            float S1 = 0.f, S2 = 0.f;
            for (int j = 0; j < gs; ++j) { S1 += raw[2 * (g0 + j)]; S2 += raw[2 * (g0 + j) + 1]; }
            const float t1 = mr[2 * ch + 1] * S1 * inv_n, t2 = mr[2 * ch + 1] * S2 * inv_n;
            tb2[2 * ch] = t1;
            tb2[2 * ch + 1] = t2;
            if (owner && (p.cs.db || p.cs.demb)) {
              // sum over the image's pixels of dx = kA*dz + kB*x + kC: the bias / embedding gradient of the convolution that
              // produced x, without reading dx back
              const float mu = mr[2 * ch], rs = mr[2 * ch + 1], ga = __ldg(p.gamma + ch);
              const float sum_dz = __ldg(red + 2 * ch), sum_x = __ldg(p.st0 + ((size_t)img * p.ld0 + ch) * 2);
              const float v = rs * ga * sum_dz - rs * t2 * sum_x + (mu * rs * t2 - t1) * (float)p.HW;
              if (p.cs.db) atomicAdd(p.cs.db + ch, v);
              if (p.cs.demb) atomicAdd(p.cs.demb + (size_t)__ldg(p.cs.img_row + img) * p.cs.emb_ld + p.cs.col + ch, v);
            }
          }
          consumer_sync();
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const float mu = mr[2 * (c + j)], rs = mr[2 * (c + j) + 1], ga = gam[j], be = bet[j];
          k0[j] = 0.5f * rs * ga;                      // h = z/2 = x*k0 + k1
          k1[j] = 0.5f * (be - mu * rs * ga);
          if (MODE == GS_APP) {
            const float t1 = tb2[2 * (c + j)], t2 = tb2[2 * (c + j) + 1];
            k2[j] = rs * ga; k3[j] = -rs * t2; k4[j] = mu * rs * t2 - t1;     // dx = k2*dz + x*k3 + k4
          }
        }
      }
    }
    // ---------------- one tile ----------------
    ptx::mbar_wait(bar_full + 8 * s, par);
    if (active) {
      const uint8_t* stage = ring + (size_t)s * p.stage_bytes;
      const T* xs = reinterpret_cast<const T*>(stage + (first ? 0u : p.tile0_bytes)) + cl;
      const T* gsm = reinterpret_cast<const T*>(stage + p.tile0_bytes + p.tile1_bytes) + c;
      const T* os = reinterpret_cast<const T*>(stage + p.tile0_bytes + p.tile1_bytes + p.tiled_bytes + (first ? 0u : p.tile0_bytes)) + cl;
      const bool accum = MODE == GS_APP && (first ? p.acc0 : p.acc1);
      T* outp = nullptr;
      int out_ld = 0;
      if (MODE == GS_FWD) { outp = reinterpret_cast<T*>(p.dst) + ((size_t)img * p.P + r0) * C + c; out_ld = C; }
      if (MODE == GS_APP) { outp = reinterpret_cast<T*>(first ? p.dx0 : p.dx1) + ((size_t)img * p.P + r0) * ld + cl; out_ld = ld; }
      int yy = (r0 + py) / p.W1, xx = (r0 + py) - yy * p.W1;
      const int dyy = p.PY / p.W1, dxx = p.PY - dyy * p.W1;
      for (int rb = py; rb < nrows; rb += UN * p.PY) {
        uint4 xr[UN], gr[UN], orr[UN];
        uint32_t live = 0, padm = 0;
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          const int r = rb + u * p.PY;
          const bool in = r < nrows, pad = yy == 0 || xx == 0;
          live |= (uint32_t)in << u; padm |= (uint32_t)pad << u;
          if (in && !pad) {
            xr[u] = *reinterpret_cast<const uint4*>(xs + (size_t)r * ld);
            if (MODE != GS_FWD) gr[u] = *reinterpret_cast<const uint4*>(gsm + (size_t)r * C);
            if (accum) orr[u] = *reinterpret_cast<const uint4*>(os + (size_t)r * ld);
          }
          yy += dyy; xx += dxx;
          if (xx >= p.W1) { xx -= p.W1; ++yy; }
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          if (!((live >> u) & 1u)) continue;
          const int r = rb + u * p.PY;
          const bool pad = (padm >> u) & 1u;
          float x[VEC], o[VEC];
          if (!pad) {
            load_vec(reinterpret_cast<const T*>(&xr[u]), x);
          } else {
#pragma unroll
            for (int j = 0; j < VEC; ++j) x[j] = 0.f;
          }
          if (MODE == GS_FWD) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
              const float y = fmaf(x[j], k0[j], k1[j]);
              o[j] = pad ? 0.f : (kSwish ? silu_for<T>(y) : y);          // padding rows are written as exact zeros
            }
            store_vec(outp + (size_t)r * out_ld, o);
          } else {
            if (pad) {
              if (MODE == GS_APP && !accum) {                            // gradients of padding rows are exact zeros
#pragma unroll
                for (int j = 0; j < VEC; ++j) o[j] = 0.f;
                store_vec(outp + (size_t)r * out_ld, o);
              }
              continue;
            }
            float g[VEC];
            load_vec(reinterpret_cast<const T*>(&gr[u]), g);
            if (kSwish) {
#pragma unroll
              for (int j = 0; j < VEC; ++j) g[j] = gs_dz<T>(g[j], fmaf(x[j], k0[j], k1[j]));
            }
            if (MODE == GS_RED) {
#pragma unroll
              for (int j = 0; j < VEC; ++j) { sA[j] += g[j]; sB[j] = fmaf(g[j], x[j], sB[j]); }
            } else {
#pragma unroll
              for (int j = 0; j < VEC; ++j) o[j] = fmaf(k2[j], g[j], fmaf(x[j], k3[j], k4[j]));
              if (accum) {
                float old[VEC];
                load_vec(reinterpret_cast<const T*>(&orr[u]), old);
#pragma unroll
                for (int j = 0; j < VEC; ++j) o[j] += old[j];
              }
              store_vec(outp + (size_t)r * out_ld, o);
            }
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(bar_empty + 8 * s);      // this warp is done reading the stage
    if (++s == S) { s = 0; par ^= 1u; }
  }
  if (MODE == GS_RED && cur_img >= 0) flush_red(cur_img);
}

// ---- host side -----------------------------------------------------------------------------------------------------
struct GsPlan { int R, tiles_per_img, n_items, grid, items_per_cta, CV, PY, stages; uint32_t tile0, tile1, tiled, stage_bytes, table_bytes; size_t smem; };

static bool gs_plan(int mode, int C0, int C1, int es, int images, int P, bool any_acc, GsPlan* g) {
  const int vec = 16 / es;
  const int C = C0 + C1;
  if (C0 % vec || C1 % vec || C / vec > GS_CONSUMERS || C / vec < 1) return false;
  g->CV = C / vec;
  g->PY = GS_CONSUMERS / g->CV;
  // bytes per row and stage: FWD x; RED x + dy; APP x + dy (+ old dx)
  const int row_bytes = C * es * (mode == GS_FWD ? 1 : (mode == GS_RED ? 2 : (any_acc ? 3 : 2)));
  static const int env_kb = [] { const char* e = getenv("VF_GS_TILE_KB"); return e ? atoi(e) : 0; }();
  static const int env_stages = [] { const char* e = getenv("VF_GS_STAGES"); return e ? atoi(e) : 0; }();
  static const int env_cta = [] { const char* e = getenv("VF_GS_CTAS"); return e ? atoi(e) : 0; }();
  const int target = env_kb > 0 ? env_kb * 1024 : (mode == GS_FWD ? 24 * 1024 : 32 * 1024);
  int R = target / row_bytes / g->PY * g->PY;
  if (R < g->PY) R = g->PY;
  const int Pup = (P + g->PY - 1) / g->PY * g->PY;
  if (R > Pup) R = Pup;
  g->R = R;
  g->tiles_per_img = (P + R - 1) / R;
  g->n_items = g->tiles_per_img * images;
  auto up128 = [](size_t x) { return (uint32_t)((x + 127) / 128 * 128); };
  g->tile0 = up128((size_t)R * C0 * es);
  g->tile1 = C1 ? up128((size_t)R * C1 * es) : 0u;
  g->tiled = mode == GS_FWD ? 0u : up128((size_t)R * C * es);
  g->stage_bytes = g->tile0 + g->tile1 + g->tiled + (mode == GS_APP && any_acc ? g->tile0 + g->tile1 : 0u);
  const size_t table_floats = mode == GS_RED ? (size_t)2 * C + (size_t)2 * C * g->PY + 2 * C : (size_t)6 * C;
  g->table_bytes = up128(table_floats * 4);
  const size_t budget = env_cta == 1 ? 220 * 1024 : (size_t)(226 * 1024 / GS_MIN_CTAS - 1024);     // GS_MIN_CTAS CTAs per SM
  g->stages = env_stages >= 2 && env_stages <= GS_MAX_STAGES ? env_stages : GS_MAX_STAGES;
  while (g->stages > 2 && 128 + g->table_bytes + (size_t)g->stages * g->stage_bytes > budget) --g->stages;
  g->smem = 128 + g->table_bytes + (size_t)g->stages * g->stage_bytes;
  if (g->smem > 227 * 1024) return false;
  const int resident = env_cta == 1 ? sm_count() : (int)(226 * 1024 / (g->smem + 1024)) * sm_count();
  g->grid = g->n_items < resident ? g->n_items : resident;
  g->items_per_cta = (g->n_items + g->grid - 1) / g->grid;
  return true;
}

template <typename T, int MODE>
static int gs_launch(const GsParams& p, const GsPlan& g, cudaStream_t st) {
  if (p.swish) {
    VF_SET_MAX_SMEM((gn_stream_kernel<T, MODE, true>), 227 * 1024);
    VF_CUDA(launch_pdl(gn_stream_kernel<T, MODE, true>, dim3(g.grid), dim3(GS_THREADS), g.smem, st, p));
  } else {
    VF_SET_MAX_SMEM((gn_stream_kernel<T, MODE, false>), 227 * 1024);
    VF_CUDA(launch_pdl(gn_stream_kernel<T, MODE, false>, dim3(g.grid), dim3(GS_THREADS), g.smem, st, p));
  }
  VF_LAUNCH_CHECK();
  return VF_OK;
}

static void gs_fill(GsParams& p, const GsPlan& g) {
  p.R = g.R; p.tiles_per_img = g.tiles_per_img; p.n_items = g.n_items; p.items_per_cta = g.items_per_cta; p.CV = g.CV; p.PY = g.PY;
  p.stages = g.stages; p.tile0_bytes = g.tile0; p.tile1_bytes = g.tile1; p.tiled_bytes = g.tiled; p.stage_bytes = g.stage_bytes;
  p.table_bytes = g.table_bytes;
}

bool gn_stream_enabled() {
  static const bool on = [] { const char* e = getenv("VF_GN_STREAM"); return !(e && e[0] == '0'); }();
  return on;
}

// Forward: returns VF_OK, or 1 when the shape is not supported by the streaming kernel (the caller falls back).
int gn_apply_stream(const void* src0, int C0, const float* stats0, int ld0, const void* src1, int C1, const float* stats1, int ld1, int dtype,
                    int images, int H, int W, int groups, const float* gamma, const float* beta, int swish, void* dst, const vf_gn_shift& sh,
                    cudaStream_t st) {
  const int es = dtype == VF_BF16 ? 2 : 4;
  GsPlan g;
  if (!gs_plan(GS_FWD, C0, C1, es, images, (H + 1) * (W + 1), false, &g)) return 1;
  GsParams p{};
  p.s0 = (const uint8_t*)src0; p.s1 = (const uint8_t*)src1; p.C0 = C0; p.C1 = C1; p.dst = (uint8_t*)dst;
  p.st0 = stats0; p.ld0 = ld0; p.st1 = C1 ? stats1 : nullptr; p.ld1 = ld1; p.gamma = gamma; p.beta = beta; p.sh = sh;
  p.images = images; p.HW = H * W; p.W1 = W + 1; p.P = (H + 1) * (W + 1); p.groups = groups; p.swish = swish;
  gs_fill(p, g);
  return dtype == VF_BF16 ? gs_launch<__nv_bfloat16, GS_FWD>(p, g, st) : gs_launch<float, GS_FWD>(p, g, st);
}

// Backward: reduce + apply.  Same return convention.  `scratch` ([images, C, 2]) must be zero on entry.
int gn_backward_stream(const void* src0, int C0, const float* stats0, int ld0, const void* src1, int C1, const float* stats1, int ld1, int dtype,
                       int images, int H, int W, int groups, const float* gamma, const float* beta, int swish, const void* dy, float* scratch,
                       float* dgamma, float* dbeta, void* dx0, int acc0, void* dx1, int acc1, const GnColsum* colsum, const vf_gn_shift& sh,
                       cudaStream_t st) {
  const int es = dtype == VF_BF16 ? 2 : 4;
  const int P = (H + 1) * (W + 1);
  GsPlan gr, ga;
  const bool any_acc = acc0 || (C1 && acc1);
  if (!gs_plan(GS_RED, C0, C1, es, images, P, false, &gr) || !gs_plan(GS_APP, C0, C1, es, images, P, any_acc, &ga)) return 1;
  GsParams p{};
  p.s0 = (const uint8_t*)src0; p.s1 = (const uint8_t*)src1; p.C0 = C0; p.C1 = C1; p.dy = (const uint8_t*)dy;
  p.dx0 = (uint8_t*)dx0; p.dx1 = (uint8_t*)dx1; p.acc0 = acc0; p.acc1 = C1 ? acc1 : 0;
  p.st0 = stats0; p.ld0 = ld0; p.st1 = C1 ? stats1 : nullptr; p.ld1 = ld1; p.gamma = gamma; p.beta = beta;
  p.red = scratch; p.dgamma = dgamma; p.dbeta = dbeta; p.sh = sh;
  if (colsum) p.cs = *colsum;
  p.images = images; p.HW = H * W; p.W1 = W + 1; p.P = P; p.groups = groups; p.swish = swish;
  gs_fill(p, gr);
  int rc = dtype == VF_BF16 ? gs_launch<__nv_bfloat16, GS_RED>(p, gr, st) : gs_launch<float, GS_RED>(p, gr, st);
  if (rc != VF_OK) return rc;
  gs_fill(p, ga);
  return dtype == VF_BF16 ? gs_launch<__nv_bfloat16, GS_APP>(p, ga, st) : gs_launch<float, GS_APP>(p, ga, st);
}

}  // namespace vf
