// Host-side UNet plan: the layer table of model/unet.py:38-112 and the launch sequence of UNet.forward
// (unet.py:114-138) expressed over the stage-level operators of this library.  Owns no device memory: the
// caller passes master parameter pointers, the packed-weight buffer and the workspace.
//
// Fusions relative to the reference module graph:
//   * torch.cat((x, skip)) (unet.py:134) is never materialised: GroupNorm reads both sources, res_conv is two
//     extra K-segments of the block's second convolution;
//   * res_conv (1x1, unet.py:238) + "h + res_conv(x)" (:245) ride in conv2's accumulator (K-concatenation);
//   * FeatureWiseAffine (:160-177): all 30 Linear layers are one table computed per sample by vf_embed and added
//     in conv1's epilogue; conv bias, identity residual and the attention residual (:277) are epilogue adds;
//   * the first 3x3 conv runs as a K=64 GEMM over the im2col rows written by vf_pack_views.
#include <array>
#include <map>
#include <string>
#include <vector>

#include "vf_common.cuh"

namespace vf {
extern int g_force_simt_flag;     // vf_api.cu: CUDA-core cross-check mode
int tc_debug_flags();             // k_gemm_tc.cu: test hooks of the tcgen05 convolution
bool conv2d_tc_stride2_gathers(int H, int W, int C);
}

namespace vf {

int pack_identity(void* dst, int dtype, int n_rows, int k_total, int k_off, cudaStream_t st);

struct ParamInfo {
  std::string name;
  int64_t shape[4];
  int ndim;
};

struct ResBlock {
  std::string name;
  int c0 = 0, c1 = 0;   // channels of x and of the skip tensor (0 when none)
  int cout = 0;
  bool attn = false;
  // parameter indices
  int nf_w, nf_b, g1_w, g1_b, c1_w, c1_b, g2_w, g2_b, c2_w, c2_b, rs_w = -1, rs_b = -1;
  int an_w = -1, an_b = -1, qkv_w = -1, ao_w = -1, ao_b = -1;
  // packed offsets (bytes from the start of the packed buffer)
  size_t w1 = 0, w2 = 0, wqkv = 0, wout = 0, bias2 = 0;
  // transposed (data-gradient) packs in packed_t
  size_t wt1 = 0, wt2 = 0, wtr0 = 0, wtr1 = 0, wtqkv = 0, wtout = 0;
  int emb_col = 0;
};

struct Layer {
  int kind;  // 0 conv0, 1 resblock, 2 down, 3 up
  int rb = -1;
  std::string name;
  int c = 0;
  int w_idx = -1, b_idx = -1;
  size_t w = 0, wt = 0;
};

}  // namespace vf

struct vf_unet {
  vf_unet_config cfg;
  int dtype;
  std::vector<vf::ParamInfo> params;
  std::vector<vf::ResBlock> blocks;
  std::vector<vf::Layer> downs, mid, ups;
  int final_c = 0, fin_gw = -1, fin_gb = -1, fin_w = -1, fin_b = -1;
  size_t final_w = 0, final_wt = 0;
  int final_npad = 16;
  int mlp_w0, mlp_b0, mlp_w2, mlp_b2;
  int E = 0;
  int k0 = 0;
  size_t packed_bytes = 0, emb_w_off = 0, emb_b_off = 0, conv0_w = 0;
  std::vector<const float*> master;   // device pointers of the fp32 masters (valid after pack_weights)
  bool packed = false;
  int launches = 0;                   // kernels enqueued by the last forward
  struct Tap { size_t off; int C, H, W; int dtype; int ld; };
  std::map<std::string, Tap> taps;
  int last_images = 0;
  int capacity = 0;               // arena layout is computed for max(capacity, images) view-images (vf_unet_set_capacity)
  int last_layout = 0;            // layout image count of the last forward (the backward lays its arena out the same way)
  unsigned long long fwd_gen = 0; // forward generation: bumped by every vf_unet_forward (vf_unet_forward_generation)
  void* bw_state = nullptr;       // vf::BwdCtx of a phased backward in flight (vf_unet_backward_phase)
  std::vector<int> bw_phase_stop; // its phase boundaries
  // optional per-kernel-class timing of one forward (CUDA events around every launch; perturbs overlap, so it is
  // only enabled for the roofline breakdown, never for the throughput measurement)
  // training tape: what the last forward did, in order (replayed in reverse by vf_unet_backward)
  struct TapeOp {
    int kind;                          // 0 conv, 1 gn, 2 attention, 3 upsample
    vf_conv_args conv;
    int w_idx[3], c_off[3], cin_total[3];   // OIHW parameter of each K segment (-1: identity / none), channel offset in it
    int b_idx[2];                      // bias parameters summed into this conv's bias (-1: none)
    int emb_col, nf_w, nf_b;           // embedding columns added in the epilogue (-1: none)
    size_t wt_off[3];                  // transposed weight pack per segment in packed_t (SIZE_MAX: no data gradient)
    const void* gsrc0; const void* gsrc1; int gC0, gC1; const float* gst0; const float* gst1; int gld0, gld1; int gw, gb, swish;
    vf_gn_shift gshift;               // deferred bias + embedding of source 0 (all-null: none)
    void* gdst; int gH, gW;
    const void* qkv; const void* vt; void* o; float* lse; int aC, aL;
    const void* usrc; void* udst; int uH, uW, uC;
  };
  std::vector<TapeOp> tape;
  std::map<const void*, size_t> act_bytes;
  const float* last_emb = nullptr;     // [rows, E] of the last forward
  const float* emb_w_dev = nullptr;    // concatenated embedding matrix inside the packed buffer of the last forward
  const float* last_level = nullptr; const float* last_angle = nullptr; const int* last_img_row = nullptr;
  int last_rows = 0;
  const void* last_x0 = nullptr; float* last_out = nullptr;
  size_t packed_t_bytes = 0, pack_tab_off = 0, pack_t_tab_off = 0;
  std::vector<uint8_t> pack_cache, pack_t_cache, unpack_cache;            // job tables as last uploaded (re-uploaded only when they change)
  bool packed_t = false;
  bool profiling = false;
  bool stash = true;             // the forward also produces what only the backward reads (V^T of the attention blocks)
  bool last_stash = true;
  std::vector<cudaEvent_t> ev;
  std::vector<int> ev_kind;      // kind of the launch between ev[i] and ev[i+1]
  std::vector<std::array<int, 6>> ev_desc;   // images, H, C_in (K total for conv), C_out, ksize, stride of that launch
  int ev_used = 0;
};

namespace vf {

static int add_param(vf_unet* u, const std::string& name, std::initializer_list<int64_t> shape) {
  ParamInfo p;
  p.name = name;
  p.ndim = (int)shape.size();
  int i = 0;
  for (auto s : shape) p.shape[i++] = s;
  for (; i < 4; ++i) p.shape[i] = 1;
  u->params.push_back(p);
  return (int)u->params.size() - 1;
}

static void add_resblock(vf_unet* u, std::vector<Layer>& sec, const std::string& name, int c0, int c1, int cout, bool attn) {
  const int ic = u->cfg.inner_channel;
  const int cin = c0 + c1;
  ResBlock b;
  b.name = name; b.c0 = c0; b.c1 = c1; b.cout = cout; b.attn = attn;
  const std::string r = name + ".res_block";
  // registration order of ResnetBlock.__init__ (unet.py:232-238): noise_func, block1, block2, res_conv
  b.nf_w = add_param(u, r + ".noise_func.noise_func.0.weight", {cout, ic});
  b.nf_b = add_param(u, r + ".noise_func.noise_func.0.bias", {cout});
  b.g1_w = add_param(u, r + ".block1.block.0.weight", {cin});
  b.g1_b = add_param(u, r + ".block1.block.0.bias", {cin});
  b.c1_w = add_param(u, r + ".block1.block.3.weight", {cout, cin, 3, 3});
  b.c1_b = add_param(u, r + ".block1.block.3.bias", {cout});
  b.g2_w = add_param(u, r + ".block2.block.0.weight", {cout});
  b.g2_b = add_param(u, r + ".block2.block.0.bias", {cout});
  b.c2_w = add_param(u, r + ".block2.block.3.weight", {cout, cout, 3, 3});
  b.c2_b = add_param(u, r + ".block2.block.3.bias", {cout});
  if (cin != cout) {
    b.rs_w = add_param(u, r + ".res_conv.weight", {cout, cin, 1, 1});
    b.rs_b = add_param(u, r + ".res_conv.bias", {cout});
  }
  if (attn) {
    b.an_w = add_param(u, name + ".attn.norm.weight", {cout});
    b.an_b = add_param(u, name + ".attn.norm.bias", {cout});
    b.qkv_w = add_param(u, name + ".attn.qkv.weight", {3 * cout, cout, 1, 1});
    b.ao_w = add_param(u, name + ".attn.out.weight", {cout, cout, 1, 1});
    b.ao_b = add_param(u, name + ".attn.out.bias", {cout});
  }
  b.emb_col = u->E;
  u->E += cout;
  u->blocks.push_back(b);
  Layer l;
  l.kind = 1; l.rb = (int)u->blocks.size() - 1; l.name = name; l.c = cout;
  sec.push_back(l);
}

static void add_resample(vf_unet* u, std::vector<Layer>& sec, const std::string& name, int kind, int c) {
  Layer l;
  l.kind = kind; l.name = name; l.c = c;
  l.w_idx = add_param(u, name + ".conv.weight", {c, c, 3, 3});
  l.b_idx = add_param(u, name + ".conv.bias", {c});
  sec.push_back(l);
}

// ---- execution context: the same walk either sizes the workspace (dry) or enqueues the kernels -----------
struct Exec {
  bool dry;
  uint8_t* base;
  size_t off = 0;
  size_t cap = 0;
  cudaStream_t st;
  int rc = VF_OK;
  int launches = 0;
  int alloc_images = 0;          // image count the buffers are SIZED for (>= the images processed): keeps the arena layout fixed
  vf_unet* u = nullptr;
  float* stats_base = nullptr;   // GroupNorm statistics arena (zeroed once per forward)
  size_t stats_used = 0;         // floats
  float* alloc_stats(size_t n) {
    float* p = dry ? nullptr : stats_base + stats_used;
    stats_used += n;
    return p;
  }
  void* alloc(size_t bytes) {
    off = align_up(off, 256);
    void* p = dry ? nullptr : base + off;
    off += bytes;
    if (!dry && u) u->act_bytes[p] = bytes;
    return p;
  }
};

enum { K_CONV = 0, K_GN_STATS, K_GN_APPLY, K_ATTN, K_UPSAMPLE, K_EMBED, K_NUM };

static void prof_mark(vf_unet* u, cudaStream_t st, int kind) {
  if ((size_t)u->ev_used >= u->ev.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    u->ev.push_back(e);
    u->ev_kind.push_back(-1);
    u->ev_desc.push_back({0, 0, 0, 0, 0, 0});
  }
  cudaEventRecord(u->ev[u->ev_used], st);
  u->ev_kind[u->ev_used] = kind;
  u->ev_desc[u->ev_used] = {0, 0, 0, 0, 0, 0};
  ++u->ev_used;
}

// shape of the launch that the last prof_mark opened (per-launch table of vf_unet_profile_launches)
static void prof_describe(vf_unet* u, int images, int H, int cin, int cout, int ksize, int stride) {
  if (u && u->profiling && u->ev_used > 0) u->ev_desc[u->ev_used - 1] = {images, H, cin, cout, ksize, stride};
}

#define VF_RUN(ex, kind, call)                                       \
  do {                                                               \
    if (!(ex).dry && (ex).rc == VF_OK) {                             \
      if ((ex).u && (ex).u->profiling) prof_mark((ex).u, (ex).st, kind); \
      (ex).rc = (call);                                              \
      (ex).launches += 1;                                            \
    }                                                                \
  } while (0)

template <typename T>
__global__ void tap_to_nchw_kernel(const T* __restrict__ src, int ld, int C, int H, int W, size_t total, float* __restrict__ dst) {
  size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // over (r, c, y, x) of the NCHW output
  if (gid >= total) return;
  const int HW = H * W;
  const int pix = (int)(gid % HW);
  const size_t rc = gid / HW;
  const int c = (int)(rc % C);
  const size_t r = rc / C;
  const int y = pix / W, x = pix - y * W;
  dst[gid] = to_f(src[((r * (H + 1) + (y + 1)) * (W + 1) + (x + 1)) * ld + c]);   // PADDED row order
}

__global__ void add_bias_kernel(const float* a, const float* b, int n, float* dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = a[i] + (b ? b[i] : 0.f);
}

// ---- table-driven weight packing: ONE launch derives every GEMM-ready weight tensor of a pack from the fp32 masters ----
struct PackJob {
  const float* src;
  const float* src2;
  void* dst;
  int type;       // 0 conv K-major, 1 conv transposed + tap-flipped, 2 transposed 1x1 slice, 3 identity block, 4 bias sum, 5 fp32 copy
  int cout, cin, kk, rows_pad, k_total, k_off, n_stride, c_off;
  int total;      // destination elements of the job
  int pad_;
};
constexpr int kPackChunk = 4096;

template <typename T>
__global__ void __launch_bounds__(256) multi_pack_kernel(const PackJob* __restrict__ jobs, const int* __restrict__ chunk_job,
                                                         const int* __restrict__ chunk_start) {
  const PackJob j = jobs[chunk_job[blockIdx.x]];
  const int start = chunk_start[blockIdx.x];
  const int end = min(j.total, start + kPackChunk);
  T* dst = reinterpret_cast<T*>(j.dst);
  for (int gid = start + threadIdx.x; gid < end; gid += 256) {
    switch (j.type) {
      case 0: {   // dst[n][k_off + tap*cin + c] = w[n][c][tap]            (vf_pack_conv_weight)
        const int per_row = j.kk * j.cin;
        const int n = gid / per_row, r = gid - n * per_row;
        const int tap = r / j.cin, c = r - tap * j.cin;
        const float v = n < j.cout ? __ldg(j.src + ((size_t)n * j.cin + c) * j.kk + tap) : 0.f;
        dst[(size_t)n * j.k_total + j.k_off + r] = from_f<T>(v);
        break;
      }
      case 1: {   // dst[c][k_off + (kk-1-tap)*n_stride + n] = w[n][c][tap]  (vf_pack_conv_weight_t)
        const int per_row = j.kk * j.n_stride;
        const int c = gid / per_row, r = gid - c * per_row;
        const int tapf = r / j.n_stride, n = r - tapf * j.n_stride;
        const int tap = j.kk - 1 - tapf;
        const float v = (c < j.cin && n < j.cout) ? __ldg(j.src + ((size_t)n * j.cin + c) * j.kk + tap) : 0.f;
        dst[(size_t)c * j.k_total + j.k_off + r] = from_f<T>(v);
        break;
      }
      case 2: {   // dst[c][n] = w[n][c_off + c]                          (transposed slice of a 1x1 weight)
        const int c = gid / j.cout, n = gid - c * j.cout;
        dst[gid] = from_f<T>(__ldg(j.src + (size_t)n * j.cin + j.c_off + c));
        break;
      }
      case 3: {   // identity block riding as a 1x1 K segment
        const int n = gid / j.cout, c = gid - n * j.cout;
        dst[(size_t)n * j.k_total + j.k_off + c] = from_f<T>(c == n ? 1.f : 0.f);
        break;
      }
      case 4:     // fused bias b + b_res (fp32)
        reinterpret_cast<float*>(j.dst)[gid] = __ldg(j.src + gid) + (j.src2 ? __ldg(j.src2 + gid) : 0.f);
        break;
      default:    // plain fp32 copy (embedding Linear rows)
        reinterpret_cast<float*>(j.dst)[gid] = __ldg(j.src + gid);
        break;
    }
  }
}

struct PackList {
  std::vector<PackJob> jobs;
  std::vector<int> chunk_job, chunk_start;
  void add(PackJob j) {
    if (j.total <= 0) return;
    const int id = (int)jobs.size();
    jobs.push_back(j);
    for (int s0 = 0; s0 < j.total; s0 += kPackChunk) { chunk_job.push_back(id); chunk_start.push_back(s0); }
  }
  void conv(const float* w, int cout, int cin, int ksize, void* dst, int cout_pad, int k_total, int k_off) {
    PackJob j{};
    j.src = w; j.dst = dst; j.type = 0; j.cout = cout; j.cin = cin; j.kk = ksize * ksize; j.rows_pad = cout_pad; j.k_total = k_total;
    j.k_off = k_off; j.total = cout_pad * j.kk * cin;
    add(j);
  }
  void conv_t(const float* w, int cout, int cin, int ksize, void* dst, int cin_pad, int k_total, int k_off, int n_stride) {
    PackJob j{};
    j.src = w; j.dst = dst; j.type = 1; j.cout = cout; j.cin = cin; j.kk = ksize * ksize; j.rows_pad = cin_pad; j.k_total = k_total;
    j.k_off = k_off; j.n_stride = n_stride; j.total = cin_pad * j.kk * n_stride;
    add(j);
  }
  void slice_t(const float* w, int cout, int cin, int c_off, int c_cnt, void* dst) {
    PackJob j{};
    j.src = w; j.dst = dst; j.type = 2; j.cout = cout; j.cin = cin; j.c_off = c_off; j.total = c_cnt * cout;
    add(j);
  }
  void identity(void* dst, int n_rows, int k_total, int k_off) {
    PackJob j{};
    j.dst = dst; j.type = 3; j.cout = n_rows; j.k_total = k_total; j.k_off = k_off; j.total = n_rows * n_rows;
    add(j);
  }
  void bias(const float* a, const float* b, int n, void* dst) {
    PackJob j{};
    j.src = a; j.src2 = b; j.dst = dst; j.type = 4; j.total = n;
    add(j);
  }
  void copy(const float* src, int n, void* dst) {
    PackJob j{};
    j.src = src; j.dst = dst; j.type = 5; j.total = n;
    add(j);
  }
  size_t table_bytes() const { return align_up(jobs.size() * sizeof(PackJob), 256) + 2 * align_up(chunk_job.size() * sizeof(int), 256); }
  // Launches the one kernel over the job table living in `region` (inside the caller's pack buffer).  The table is
  // uploaded only when it differs from the last upload (first call, or the caller moved a parameter / the buffer):
  // in steady state a re-pack is one memset and one launch, with no host-to-device traffic.
  int run(std::vector<uint8_t>& cache, uint8_t* region, size_t region_bytes, int dtype, cudaStream_t st) const {
    VF_REQUIRE(table_bytes() <= region_bytes, "weight pack: job table of %zu B exceeds the reserved %zu B", table_bytes(), region_bytes);
    const size_t o_cj = align_up(jobs.size() * sizeof(PackJob), 256), o_cs = o_cj + align_up(chunk_job.size() * sizeof(int), 256);
    std::vector<uint8_t> blob(table_bytes() + sizeof(void*), 0);
    memcpy(blob.data(), jobs.data(), jobs.size() * sizeof(PackJob));
    memcpy(blob.data() + o_cj, chunk_job.data(), chunk_job.size() * sizeof(int));
    memcpy(blob.data() + o_cs, chunk_start.data(), chunk_start.size() * sizeof(int));
    memcpy(blob.data() + table_bytes(), &region, sizeof(void*));            // the table's own location is part of the key
    if (blob != cache) {
      VF_CUDA(cudaMemcpyAsync(region, blob.data(), table_bytes(), cudaMemcpyHostToDevice, st));   // pageable: staged before returning
      cache = blob;
    }
    const unsigned grid = (unsigned)chunk_job.size();
    if (dtype == VF_BF16)
      multi_pack_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const PackJob*)region, (const int*)(region + o_cj), (const int*)(region + o_cs));
    else
      multi_pack_kernel<float><<<grid, 256, 0, st>>>((const PackJob*)region, (const int*)(region + o_cj), (const int*)(region + o_cs));
    VF_LAUNCH_CHECK();
    return VF_OK;
  }
};
constexpr size_t kPackTableBytes = 768 * 1024;      // reserved at the end of each pack buffer for the job table

struct Act {
  void* p;
  int C, H, W;
  float* stats;   // [images, C, 2] sum / sum-of-squares written by the producing conv's epilogue, or null
  // deferred additive constant: the tensor is stored WITHOUT bias[c] + emb[img_row[img]][c]; its one consumer (a GroupNorm)
  // folds it into the normalisation (vf_gn_shift)
  const float* sh_bias = nullptr;
  const float* sh_emb = nullptr;
};

static Act new_act(Exec& ex, const vf_unet* u, int images, int C, int H, int W, bool want_stats) {
  // spatial activations live in the PADDED row order: images * (H+1) * (W+1) rows
  const int li = ex.alloc_images > images ? ex.alloc_images : images;
  Act a{ex.alloc((size_t)li * (H + 1) * (W + 1) * C * (u->dtype == VF_BF16 ? 2 : 4)), C, H, W, nullptr};
  if (want_stats) a.stats = ex.alloc_stats((size_t)li * C * 2);
  return a;
}

static int k_elems(const vf_unet* u) { return u->dtype == VF_BF16 ? 2 : 4; }

}  // namespace vf

using namespace vf;

extern "C" __attribute__((visibility("default"))) int vf_unet_create(const vf_unet_config* cfg, int act_dtype, vf_unet** out) {
  VF_REQUIRE(cfg && out, "vf_unet_create: null args");
  VF_REQUIRE(act_dtype == VF_F32 || act_dtype == VF_BF16, "vf_unet_create: act_dtype=%d", act_dtype);
  VF_REQUIRE(cfg->n_mults >= 1 && cfg->n_mults <= VF_MAX_LEVELS && cfg->n_attn_res >= 0 && cfg->n_attn_res <= VF_MAX_LEVELS,
             "vf_unet_create: bad level counts");
  VF_REQUIRE(cfg->inner_channel > 0 && cfg->inner_channel % 4 == 0 && cfg->res_blocks >= 1 && cfg->in_channel >= 1,
             "vf_unet_create: bad channel config");
  VF_REQUIRE(cfg->norm_groups > 0 && cfg->inner_channel % cfg->norm_groups == 0,
             "vf_unet_create: inner_channel %d not divisible by norm_groups %d", cfg->inner_channel, cfg->norm_groups);
  const int outc = cfg->out_channel > 0 ? cfg->out_channel : cfg->in_channel;
  VF_REQUIRE(outc <= 8, "vf_unet_create: out_channel=%d > 8 unsupported", outc);
  const int S = cfg->image_size;
  VF_REQUIRE(S > 0 && (S & (S - 1)) == 0 && (S >> (cfg->n_mults - 1)) >= 1, "vf_unet_create: image_size=%d must be a power of two", S);
  const int quantum = act_dtype == VF_BF16 ? 64 : 16;
  for (int i = 0; i < cfg->n_mults; ++i)
    VF_REQUIRE((cfg->inner_channel * cfg->channel_mults[i]) % quantum == 0,
               "vf_unet_create: level %d has %d channels; this precision mode needs multiples of %d", i,
               cfg->inner_channel * cfg->channel_mults[i], quantum);
  vf_unet* u = new vf_unet();
  u->cfg = *cfg;
  u->cfg.out_channel = outc;
  u->dtype = act_dtype;
  const int ic = cfg->inner_channel;
  u->mlp_w0 = add_param(u, "noise_level_mlp.0.weight", {4 * ic, ic});
  u->mlp_b0 = add_param(u, "noise_level_mlp.0.bias", {4 * ic});
  u->mlp_w2 = add_param(u, "noise_level_mlp.2.weight", {ic, 4 * ic});
  u->mlp_b2 = add_param(u, "noise_level_mlp.2.bias", {ic});
  auto is_attn = [&](int res) {
    for (int i = 0; i < cfg->n_attn_res; ++i)
      if (cfg->attn_res[i] == res) return true;
    return false;
  };
  // ---- unet.py:38-64 (downs)
  int pre = ic, res = S;
  std::vector<int> feat{pre};
  {
    Layer l;
    l.kind = 0; l.name = "downs.0"; l.c = ic;
    l.w_idx = add_param(u, "downs.0.weight", {ic, cfg->in_channel, 3, 3});
    l.b_idx = add_param(u, "downs.0.bias", {ic});
    u->downs.push_back(l);
  }
  for (int i = 0; i < cfg->n_mults; ++i) {
    const bool last = i == cfg->n_mults - 1;
    const int cm = ic * cfg->channel_mults[i];
    for (int r = 0; r < cfg->res_blocks; ++r) {
      add_resblock(u, u->downs, "downs." + std::to_string(u->downs.size()), pre, 0, cm, is_attn(res));
      feat.push_back(cm);
      pre = cm;
    }
    if (!last) {
      add_resample(u, u->downs, "downs." + std::to_string(u->downs.size()), 2, pre);
      feat.push_back(pre);
      res /= 2;
    }
  }
  // ---- :66-85 (mid)
  add_resblock(u, u->mid, "mid.0", pre, 0, pre, true);
  add_resblock(u, u->mid, "mid.1", pre, 0, pre, false);
  // ---- :87-108 (ups)
  for (int i = cfg->n_mults - 1; i >= 0; --i) {
    const bool last = i < 1;
    const int cm = ic * cfg->channel_mults[i];
    for (int r = 0; r < cfg->res_blocks + 1; ++r) {
      const int skip = feat.back();
      feat.pop_back();
      add_resblock(u, u->ups, "ups." + std::to_string(u->ups.size()), pre, skip, cm, is_attn(res));
      pre = cm;
    }
    if (!last) {
      add_resample(u, u->ups, "ups." + std::to_string(u->ups.size()), 3, pre);
      res *= 2;
    }
  }
  // ---- :110-112 (final_conv)
  u->final_c = pre;
  u->fin_gw = add_param(u, "final_conv.block.0.weight", {pre});
  u->fin_gb = add_param(u, "final_conv.block.0.bias", {pre});
  u->fin_w = add_param(u, "final_conv.block.3.weight", {outc, pre, 3, 3});
  u->fin_b = add_param(u, "final_conv.block.3.bias", {outc});

  // ---- packed-weight layout
  const size_t es = dtype_size(act_dtype);
  size_t off = 0;
  auto take = [&](size_t bytes) { off = align_up(off, 256); size_t o = off; off += bytes; return o; };
  u->k0 = (int)align_up((size_t)9 * cfg->in_channel, 64);
  u->conv0_w = take((size_t)ic * u->k0 * es);
  for (auto& b : u->blocks) {
    const int cin = b.c0 + b.c1;
    b.w1 = take((size_t)b.cout * 9 * cin * es);
    const int k2 = 9 * b.cout + (b.rs_w >= 0 ? cin : 0);   // conv2 taps + res_conv columns (the projection rides in the GEMM)
    b.w2 = take((size_t)b.cout * k2 * es);
    b.bias2 = take((size_t)b.cout * 4);
    if (b.attn) {
      b.wqkv = take((size_t)3 * b.cout * b.cout * es);
      b.wout = take((size_t)b.cout * b.cout * es);
    }
  }
  for (auto* sec : {&u->downs, &u->mid, &u->ups})
    for (auto& l : *sec)
      if (l.kind == 2 || l.kind == 3) l.w = take((size_t)l.c * 9 * l.c * es);
  u->final_w = take((size_t)16 * 9 * u->final_c * es);
  u->emb_w_off = take((size_t)u->E * ic * 4);
  u->emb_b_off = take((size_t)u->E * 4);
  u->pack_tab_off = align_up(off, 256);
  u->packed_bytes = u->pack_tab_off + kPackTableBytes;
  // transposed packs (data gradients): rows = source channels, K = taps * (gradient channels)
  off = 0;
  for (auto& b : u->blocks) {
    const int cin = b.c0 + b.c1;
    b.wt1 = take((size_t)cin * 9 * b.cout * es);
    b.wt2 = take((size_t)b.cout * 9 * b.cout * es);
    if (b.rs_w >= 0) {
      b.wtr0 = take((size_t)b.c0 * b.cout * es);
      if (b.c1) b.wtr1 = take((size_t)b.c1 * b.cout * es);
    }
    if (b.attn) {
      b.wtqkv = take((size_t)b.cout * 3 * b.cout * es);
      b.wtout = take((size_t)b.cout * b.cout * es);
    }
  }
  for (auto* sec : {&u->downs, &u->mid, &u->ups})
    for (auto& l : *sec)
      if (l.kind == 2 || l.kind == 3) l.wt = take((size_t)l.c * 9 * l.c * es);
  u->final_npad = act_dtype == VF_BF16 ? 64 : 16;
  u->final_wt = take((size_t)u->final_c * 9 * u->final_npad * es);
  u->pack_t_tab_off = align_up(off, 256);
  u->packed_t_bytes = u->pack_t_tab_off + kPackTableBytes;
  *out = u;
  return VF_OK;
}

namespace vf { void bw_state_free(vf_unet* u); }
extern "C" __attribute__((visibility("default"))) void vf_unet_destroy(vf_unet* u) {
  if (!u) return;
  vf::bw_state_free(u);
  for (auto e : u->ev) cudaEventDestroy(e);
  delete u;
}
extern "C" __attribute__((visibility("default"))) int vf_unet_num_params(const vf_unet* u) { return u ? (int)u->params.size() : 0; }
extern "C" __attribute__((visibility("default"))) int vf_unet_emb_channels(const vf_unet* u) { return u ? u->E : 0; }
extern "C" __attribute__((visibility("default"))) int vf_unet_k0(const vf_unet* u) { return u ? u->k0 : 0; }
extern "C" __attribute__((visibility("default"))) int vf_unet_act_dtype(const vf_unet* u) { return u ? u->dtype : -1; }
extern "C" __attribute__((visibility("default"))) size_t vf_unet_packed_bytes(const vf_unet* u) { return u ? u->packed_bytes : 0; }
extern "C" __attribute__((visibility("default"))) int vf_unet_last_launches(const vf_unet* u) { return u ? u->launches : 0; }

extern "C" __attribute__((visibility("default"))) int vf_unet_param_info(const vf_unet* u, int index, char* name_buf, int name_cap, int64_t shape[4], int* ndim) {
  VF_REQUIRE(u && index >= 0 && index < (int)u->params.size(), "vf_unet_param_info: index %d out of range", index);
  const ParamInfo& p = u->params[index];
  if (name_buf && name_cap > 0) snprintf(name_buf, (size_t)name_cap, "%s", p.name.c_str());
  if (shape) for (int i = 0; i < 4; ++i) shape[i] = p.shape[i];
  if (ndim) *ndim = p.ndim;
  return VF_OK;
}

extern "C" __attribute__((visibility("default"))) int vf_unet_pack_weights(vf_unet* u, const float* const* params_host, void* packed, vf_stream stream) {
  VF_REQUIRE(u && params_host && packed, "vf_unet_pack_weights: null args");
  cudaStream_t st = as_stream(stream);
  u->master.assign(params_host, params_host + u->params.size());
  for (size_t i = 0; i < u->master.size(); ++i) VF_REQUIRE(u->master[i], "vf_unet_pack_weights: parameter %zu (%s) is null", i, u->params[i].name.c_str());
  uint8_t* pk = reinterpret_cast<uint8_t*>(packed);
  const int dt = u->dtype;
  const vf_unet_config& c = u->cfg;
  VF_CUDA(cudaMemsetAsync(pk, 0, u->pack_tab_off, st));          // zero padding rows / columns of every pack
  PackList pl;
  pl.conv(u->master[u->downs[0].w_idx], c.inner_channel, c.in_channel, 3, pk + u->conv0_w, c.inner_channel, u->k0, 0);
  for (auto& b : u->blocks) {
    const int cin = b.c0 + b.c1;
    pl.conv(u->master[b.c1_w], b.cout, cin, 3, pk + b.w1, b.cout, 9 * cin, 0);
    const int k2 = 9 * b.cout + (b.rs_w >= 0 ? cin : 0);
    pl.conv(u->master[b.c2_w], b.cout, b.cout, 3, pk + b.w2, b.cout, k2, 0);
    if (b.rs_w >= 0) pl.conv(u->master[b.rs_w], b.cout, cin, 1, pk + b.w2, b.cout, k2, 9 * b.cout);
    pl.bias(u->master[b.c2_b], b.rs_b >= 0 ? u->master[b.rs_b] : nullptr, b.cout, pk + b.bias2);
    if (b.attn) {
      pl.conv(u->master[b.qkv_w], 3 * b.cout, b.cout, 1, pk + b.wqkv, 3 * b.cout, b.cout, 0);
      pl.conv(u->master[b.ao_w], b.cout, b.cout, 1, pk + b.wout, b.cout, b.cout, 0);
    }
    // embedding Linear of this block -> rows [emb_col, emb_col + cout) of the concatenated matrix
    pl.copy(u->master[b.nf_w], b.cout * c.inner_channel, pk + u->emb_w_off + (size_t)b.emb_col * c.inner_channel * 4);
    pl.copy(u->master[b.nf_b], b.cout, pk + u->emb_b_off + (size_t)b.emb_col * 4);
  }
  for (auto* sec : {&u->downs, &u->mid, &u->ups})
    for (auto& l : *sec)
      if (l.kind == 2 || l.kind == 3) pl.conv(u->master[l.w_idx], l.c, l.c, 3, pk + l.w, l.c, 9 * l.c, 0);
  pl.conv(u->master[u->fin_w], c.out_channel, u->final_c, 3, pk + u->final_w, 16, 9 * u->final_c, 0);
  {
    const int rc = pl.run(u->pack_cache, pk + u->pack_tab_off, kPackTableBytes, dt, st);
    if (rc != VF_OK) return rc;
  }
  u->packed = true;
  return VF_OK;
}

namespace vf {

struct ConvMeta {
  int w_idx[3] = {-1, -1, -1}, c_off[3] = {0, 0, 0}, cin_total[3] = {0, 0, 0};
  int b_idx[2] = {-1, -1};
  int emb_col = -1, nf_w = -1, nf_b = -1;
  size_t wt_off[3] = {SIZE_MAX, SIZE_MAX, SIZE_MAX};
};

static void conv_call(Exec& ex, vf_unet* u, vf_conv_args& a, const ConvMeta& m) {
  a.dtype = u->dtype;
  if (a.out_dtype < 0) a.out_dtype = u->dtype;
  VF_RUN(ex, K_CONV, vf_conv2d(&a, (vf_stream)ex.st));
  if (!ex.dry) {
    int ktot = 0;
    for (int i = 0; i < a.n_seg; ++i) ktot += a.ksize[i] * a.ksize[i] * a.src_c[i];
    prof_describe(ex.u, a.images, a.H, ktot, a.cout, a.ksize[0], a.stride);
    vf_unet::TapeOp t{};
    t.kind = 0; t.conv = a;
    for (int i = 0; i < 3; ++i) { t.w_idx[i] = m.w_idx[i]; t.c_off[i] = m.c_off[i]; t.cin_total[i] = m.cin_total[i]; t.wt_off[i] = m.wt_off[i]; }
    t.b_idx[0] = m.b_idx[0]; t.b_idx[1] = m.b_idx[1];
    t.emb_col = m.emb_col; t.nf_w = m.nf_w; t.nf_b = m.nf_b;
    u->tape.push_back(t);
  }
}

static vf_conv_args conv_args_init() {
  vf_conv_args a{};
  a.stride = 1;
  a.out_dtype = -1;
  a.in_padded = 1;
  a.out_padded = 1;
  return a;
}

// GroupNorm (+Swish) of cat(x, skip) -> new activation.  Statistics come from the producers' epilogues when
// available, otherwise from a separate vf_gn_stats pass.
static Act gn_block(Exec& ex, vf_unet* u, int images, const Act& x, const Act* skip, int gw, int gb, bool swish, const int* img_row = nullptr) {
  const int C1 = skip ? skip->C : 0;
  const int C = x.C + C1;
  const float *s0 = x.stats, *s1 = skip ? skip->stats : nullptr;
  int ld0 = x.C, ld1 = C1;
  if (!x.stats || (skip && !skip->stats)) {
    float* st = ex.alloc_stats((size_t)(ex.alloc_images > images ? ex.alloc_images : images) * C * 2);
    VF_RUN(ex, K_GN_STATS, vf_gn_stats(x.p, x.C, skip ? skip->p : nullptr, C1, u->dtype, images, x.H, x.W, st, (vf_stream)ex.st));
    s0 = st; s1 = st + 2 * x.C; ld0 = ld1 = C;
  }
  Act y = new_act(ex, u, images, C, x.H, x.W, false);
  vf_gn_shift shift{};
  if (x.sh_bias || x.sh_emb) { shift.bias = x.sh_bias; shift.emb = x.sh_emb; shift.img_row = img_row; shift.emb_ld = u->E; }
  VF_RUN(ex, K_GN_APPLY, vf_gn_apply(x.p, x.C, s0, ld0, skip ? skip->p : nullptr, C1, s1, ld1, u->dtype, images, x.H, x.W, u->cfg.norm_groups,
                                     u->master[gw], u->master[gb], swish ? 1 : 0, y.p, &shift, (vf_stream)ex.st));
  if (!ex.dry) {
    prof_describe(ex.u, images, x.H, C, C, 0, 0);
    vf_unet::TapeOp t{};
    t.kind = 1;
    t.gsrc0 = x.p; t.gC0 = x.C; t.gst0 = s0; t.gld0 = ld0;
    t.gsrc1 = skip ? skip->p : nullptr; t.gC1 = C1; t.gst1 = s1; t.gld1 = ld1;
    t.gw = gw; t.gb = gb; t.swish = swish ? 1 : 0; t.gdst = y.p; t.gH = x.H; t.gW = x.W;
    t.gshift = shift;
    u->tape.push_back(t);
  }
  return y;
}

static Act run_resblock(Exec& ex, vf_unet* u, const uint8_t* pk, int images, const ResBlock& b, const Act& x, const Act* skip,
                        const float* emb, const int* img_row) {
  const int HW = x.H * x.W;
  const size_t es = k_elems(u);
  const int cin = b.c0 + b.c1;
  const size_t li = (size_t)(ex.alloc_images > images ? ex.alloc_images : images);   // buffers are sized for the layout image count
  // block1: GN -> Swish -> conv3x3 (+bias +embedding)                                   unet.py:242-243
  Act a1 = gn_block(ex, u, images, x, skip, b.g1_w, b.g1_b, true);
  Act h1 = new_act(ex, u, images, b.cout, x.H, x.W, true);
  {
    vf_conv_args a = conv_args_init();
    a.images = images; a.H = x.H; a.W = x.W; a.n_seg = 1;
    a.src[0] = a1.p; a.src_c[0] = cin; a.ksize[0] = 3;
    a.weight = pk + b.w1; a.cout = b.cout; a.cout_pad = b.cout;
    // bias + FeatureWiseAffine add (unet.py:243, :176) are NOT applied here: h1 feeds block2's GroupNorm and nothing else, so
    // the per-(image, channel) constant folds into that normalisation (vf_gn_shift) and this epilogue adds nothing
    if (!ex.dry) { h1.sh_bias = u->master[b.c1_b]; h1.sh_emb = emb ? emb + b.emb_col : nullptr; }
    a.out = h1.p; a.out_ld = b.cout; a.stats = h1.stats;
    ConvMeta m;
    m.w_idx[0] = b.c1_w; m.cin_total[0] = cin; m.b_idx[0] = b.c1_b; m.wt_off[0] = b.wt1;
    m.emb_col = b.emb_col; m.nf_w = b.nf_w; m.nf_b = b.nf_b;
    conv_call(ex, u, a, m);
  }
  // block2: GN -> Swish -> conv3x3, + res_conv(x) or + x                                 unet.py:244-245
  Act a2 = gn_block(ex, u, images, h1, nullptr, b.g2_w, b.g2_b, true, img_row);
  Act out = new_act(ex, u, images, b.cout, x.H, x.W, true);
  {
    vf_conv_args a = conv_args_init();
    a.images = images; a.H = x.H; a.W = x.W;
    a.src[0] = a2.p; a.src_c[0] = b.cout; a.ksize[0] = 3; a.n_seg = 1;
    // res_conv(x) accumulates in the same TMEM tile (two extra K segments over x and the skip tensor).  An identity residual
    // ("h + x", Cin == Cout) goes through the epilogue's TMA-fetched residual port instead: as a 1x1 K segment it cost a second
    // activation slab per work item (64x64 64->64: 80 us against 58 us for the same layer without it)
    if (b.rs_w >= 0) {
      a.src[1] = x.p; a.src_c[1] = x.C; a.ksize[1] = 1; a.n_seg = 2;
      if (skip) { a.src[2] = skip->p; a.src_c[2] = skip->C; a.ksize[2] = 1; a.n_seg = 3; }
    } else {
      a.residual = x.p;
    }
    a.weight = pk + b.w2; a.cout = b.cout; a.cout_pad = b.cout;
    a.bias = reinterpret_cast<const float*>(pk + b.bias2);
    a.out = out.p; a.out_ld = b.cout; a.stats = out.stats;
    ConvMeta m;
    m.w_idx[0] = b.c2_w; m.cin_total[0] = b.cout; m.b_idx[0] = b.c2_b; m.b_idx[1] = b.rs_b; m.wt_off[0] = b.wt2;
    // segments 1 (x) and 2 (skip): res_conv weight slices
    if (b.rs_w >= 0) {
      m.w_idx[1] = b.rs_w; m.c_off[1] = 0; m.cin_total[1] = cin; m.wt_off[1] = b.wtr0;
      if (skip) { m.w_idx[2] = b.rs_w; m.c_off[2] = b.c0; m.cin_total[2] = cin; m.wt_off[2] = b.wtr1; }
    }
    conv_call(ex, u, a, m);
  }
  if (!b.attn) return out;
  // SelfAttention: GN -> qkv 1x1 -> softmax(QK^T/sqrt(C)) V -> out 1x1 (+bias) + input     unet.py:258-277
  const int C = b.cout;
  Act n = gn_block(ex, u, images, out, nullptr, b.an_w, b.an_b, false);
  void* qkv = ex.alloc(li * HW * 3 * C * es);
  // V^T feeds the attention backward only; without it (inference) the forward kernel reads V row-major from qkv and the
  // whole qkv projection leaves through the staged TMA epilogue.  The workspace is always sized for the stash.
  void* vt = u->dtype == VF_BF16 ? ex.alloc(li * HW * C * es) : nullptr;
  if (!ex.dry && !u->stash) vt = nullptr;
  {
    vf_conv_args a = conv_args_init();
    a.images = images; a.H = x.H; a.W = x.W; a.n_seg = 1;
    a.src[0] = n.p; a.src_c[0] = C; a.ksize[0] = 1;
    a.weight = pk + b.wqkv; a.cout = 3 * C; a.cout_pad = 3 * C;
    a.out = qkv; a.out_ld = 3 * C; a.out_padded = 0;          // attention works on FLAT token rows
    if (vt) { a.qkv_split = C; a.out_vt = vt; }
    ConvMeta m;
    m.w_idx[0] = b.qkv_w; m.cin_total[0] = C; m.wt_off[0] = b.wtqkv;
    conv_call(ex, u, a, m);
  }
  Act o{ex.alloc(li * HW * C * es), C, x.H, x.W, nullptr};   // FLAT
  float* lse = reinterpret_cast<float*>(ex.alloc(li * HW * 4));
  VF_RUN(ex, K_ATTN, vf_attention(qkv, vt, u->dtype, images, HW, C, o.p, lse, (vf_stream)ex.st));
  if (!ex.dry) {
    vf_unet::TapeOp t{};
    t.kind = 2; t.qkv = qkv; t.vt = vt; t.o = o.p; t.lse = lse; t.aC = C; t.aL = HW;
    u->tape.push_back(t);
  }
  Act out2 = new_act(ex, u, images, C, x.H, x.W, true);
  {
    vf_conv_args a = conv_args_init();
    a.images = images; a.H = x.H; a.W = x.W; a.n_seg = 1;
    a.src[0] = o.p; a.src_c[0] = C; a.ksize[0] = 1; a.in_padded = 0;
    a.weight = pk + b.wout; a.cout = C; a.cout_pad = C;
    a.bias = ex.dry ? nullptr : u->master[b.ao_b];
    a.residual = out.p;
    a.out = out2.p; a.out_ld = C; a.stats = out2.stats;
    ConvMeta m;
    m.w_idx[0] = b.ao_w; m.cin_total[0] = C; m.b_idx[0] = b.ao_b; m.wt_off[0] = b.wtout;
    conv_call(ex, u, a, m);
  }
  return out2;
}

static int walk(vf_unet* u, Exec& ex, const uint8_t* pk, int images, const void* x0, const float* level, const float* angle,
                int rows, const int* img_row, float* out) {
  const vf_unet_config& c = u->cfg;
  const int S = c.image_size;
  const size_t es = k_elems(u);
  auto tap = [&](const std::string& name, const Act& a) {
    if (!ex.dry) u->taps[name] = vf_unet::Tap{(size_t)((uint8_t*)a.p - ex.base), a.C, a.H, a.W, u->dtype, a.C};
  };
  // embedding table [rows, E]
  float* emb = reinterpret_cast<float*>(ex.alloc((size_t)(ex.alloc_images > rows ? ex.alloc_images : rows) * u->E * 4));
  if (!ex.dry) { u->last_emb = emb; u->emb_w_dev = reinterpret_cast<const float*>(pk + u->emb_w_off); }
  VF_RUN(ex, K_EMBED, vf_embed(level, angle, rows, c.inner_channel, u->master[u->mlp_w0], u->master[u->mlp_b0], u->master[u->mlp_w2],
                         u->master[u->mlp_b2], reinterpret_cast<const float*>(pk + u->emb_w_off),
                         reinterpret_cast<const float*>(pk + u->emb_b_off), u->E, emb, (vf_stream)ex.st));
  std::vector<Act> feats;
  // downs[0]: 3x3 conv as a K0 GEMM over the packed im2col rows                            unet.py:42
  Act x = new_act(ex, u, images, c.inner_channel, S, S, true);
  {
    vf_conv_args a = conv_args_init();
    a.images = images; a.H = S; a.W = S; a.n_seg = 1;
    a.src[0] = x0; a.src_c[0] = u->k0; a.ksize[0] = 1; a.in_padded = 0;
    a.weight = pk + u->conv0_w; a.cout = c.inner_channel; a.cout_pad = c.inner_channel;
    a.bias = ex.dry ? nullptr : u->master[u->downs[0].b_idx];
    a.out = x.p; a.out_ld = c.inner_channel; a.stats = x.stats;
    ConvMeta m;                       // first layer: weight gradient only (no gradient w.r.t. the images)
    m.w_idx[0] = u->downs[0].w_idx; m.cin_total[0] = c.in_channel; m.b_idx[0] = u->downs[0].b_idx;
    conv_call(ex, u, a, m);
  }
  feats.push_back(x);
  tap("downs.0", x);
  for (size_t i = 1; i < u->downs.size(); ++i) {
    const Layer& l = u->downs[i];
    if (l.kind == 1) {
      x = run_resblock(ex, u, pk, images, u->blocks[l.rb], x, nullptr, emb, img_row);
    } else {   // Downsample: conv3x3 stride 2                                              unet.py:195-201
      Act y = new_act(ex, u, images, l.c, x.H / 2, x.W / 2, true);
      // x is a raw convolution output whose padding rows were never written.  The tcgen05 stride-2 path gathers pixel
      // phases by TMA and gets its zero halo from out-of-bounds fill, so only the CUDA-core path and the training
      // backward (the weight gradient reads x with its halo) need real zeros there.
      if (u->dtype != VF_BF16 || u->stash || vf::tc_debug_flags() != 0 || vf::g_force_simt_flag || !vf::conv2d_tc_stride2_gathers(x.H, x.W, x.C))
        VF_RUN(ex, K_UPSAMPLE, vf_zero_padding(x.p, u->dtype, images, x.H, x.W, x.C, (vf_stream)ex.st));
      vf_conv_args a = conv_args_init();
      a.images = images; a.H = x.H; a.W = x.W; a.n_seg = 1; a.stride = 2;
      a.src[0] = x.p; a.src_c[0] = l.c; a.ksize[0] = 3;
      a.weight = pk + l.w; a.cout = l.c; a.cout_pad = l.c;
      a.bias = ex.dry ? nullptr : u->master[l.b_idx];
      a.out = y.p; a.out_ld = l.c; a.stats = y.stats;
      ConvMeta m;
      m.w_idx[0] = l.w_idx; m.cin_total[0] = l.c; m.b_idx[0] = l.b_idx; m.wt_off[0] = l.wt;
      conv_call(ex, u, a, m);
      x = y;
    }
    feats.push_back(x);
    tap(l.name, x);
  }
  for (auto& l : u->mid) {
    x = run_resblock(ex, u, pk, images, u->blocks[l.rb], x, nullptr, emb, img_row);
    tap(l.name, x);
  }
  for (auto& l : u->ups) {
    if (l.kind == 1) {
      Act skip = feats.back();
      feats.pop_back();
      x = run_resblock(ex, u, pk, images, u->blocks[l.rb], x, &skip, emb, img_row);
    } else {   // Upsample: nearest x2 then conv3x3                                         unet.py:185-192
      Act up = new_act(ex, u, images, l.c, 2 * x.H, 2 * x.W, false);
      VF_RUN(ex, K_UPSAMPLE, vf_upsample2x(x.p, u->dtype, images, x.H, x.W, l.c, up.p, (vf_stream)ex.st));
      if (!ex.dry) {
        vf_unet::TapeOp t{};
        t.kind = 3; t.usrc = x.p; t.udst = up.p; t.uH = x.H; t.uW = x.W; t.uC = l.c;
        u->tape.push_back(t);
      }
      Act y = new_act(ex, u, images, l.c, up.H, up.W, true);
      vf_conv_args a = conv_args_init();
      a.images = images; a.H = y.H; a.W = y.W; a.n_seg = 1;
      a.src[0] = up.p; a.src_c[0] = l.c; a.ksize[0] = 3;
      a.weight = pk + l.w; a.cout = l.c; a.cout_pad = l.c;
      a.bias = ex.dry ? nullptr : u->master[l.b_idx];
      a.out = y.p; a.out_ld = l.c; a.stats = y.stats;
      ConvMeta m;
      m.w_idx[0] = l.w_idx; m.cin_total[0] = l.c; m.b_idx[0] = l.b_idx; m.wt_off[0] = l.wt;
      conv_call(ex, u, a, m);
      x = y;
    }
    tap(l.name, x);
  }
  // final_conv: GN -> Swish -> conv3x3 -> fp32 [., 8]                                       unet.py:110-112, :138
  Act f = gn_block(ex, u, images, x, nullptr, u->fin_gw, u->fin_gb, true);
  {
    vf_conv_args a = conv_args_init();
    a.images = images; a.H = S; a.W = S; a.n_seg = 1;
    a.src[0] = f.p; a.src_c[0] = u->final_c; a.ksize[0] = 3;
    a.weight = pk + u->final_w; a.cout = c.out_channel; a.cout_pad = 16;
    a.bias = ex.dry ? nullptr : u->master[u->fin_b];
    a.out = out; a.out_dtype = VF_F32; a.out_ld = 8; a.out_padded = 0;
    ConvMeta m;
    m.w_idx[0] = u->fin_w; m.cin_total[0] = u->final_c; m.b_idx[0] = u->fin_b; m.wt_off[0] = u->final_wt;
    conv_call(ex, u, a, m);
  }
  return ex.rc;
}

}  // namespace vf

extern "C" __attribute__((visibility("default"))) size_t vf_unet_workspace_bytes(const vf_unet* u, int max_images) {
  if (!u || max_images <= 0) return 0;
  Exec ex{true, nullptr};
  ex.st = nullptr;
  ex.alloc_images = max_images;
  walk(const_cast<vf_unet*>(u), ex, nullptr, max_images, nullptr, nullptr, nullptr, max_images, nullptr, nullptr);
  return align_up(ex.off, 256) + ex.stats_used * 4 + 256;
}

extern "C" __attribute__((visibility("default"))) int vf_unet_forward(vf_unet* u, const void* packed, void* workspace, size_t workspace_bytes, int images,
                               const void* x0, const float* level, const float* angle, int rows, const int* img_row,
                               float* out, vf_stream stream) {
  VF_REQUIRE(u && packed && workspace && x0 && level && angle && img_row && out, "vf_unet_forward: null tensor");
  VF_REQUIRE(images > 0 && rows > 0 && rows <= images, "vf_unet_forward: images=%d rows=%d", images, rows);
  if (!u->packed) { set_error("vf_unet_forward: weights were not packed (call vf_unet_pack_weights)"); return VF_ERR_STATE; }
  Exec ex{false, reinterpret_cast<uint8_t*>(workspace)};
  ex.st = as_stream(stream);
  ex.cap = workspace_bytes;
  // every buffer is sized for max(capacity, images) view-images, so the arena layout (hence the zero padding rows the
  // caller established once) does not move when the batch / view counts change from call to call
  ex.alloc_images = u->capacity > images ? u->capacity : images;
  u->last_layout = ex.alloc_images;
  ++u->fwd_gen;
  {
    Exec dry{true, nullptr};
    dry.st = nullptr;
    dry.alloc_images = ex.alloc_images;
    walk(u, dry, nullptr, images, nullptr, nullptr, nullptr, rows, nullptr, nullptr);
    const size_t act = align_up(dry.off, 256), need = act + dry.stats_used * 4;
    VF_REQUIRE(need <= workspace_bytes, "vf_unet_forward: workspace too small (%zu < %zu)", workspace_bytes, need);
    ex.stats_base = reinterpret_cast<float*>(ex.base + act);
    VF_CUDA(cudaMemsetAsync(ex.stats_base, 0, dry.stats_used * 4, ex.st));
  }
  u->taps.clear();
  u->tape.clear();
  u->act_bytes.clear();
  u->last_images = images;
  u->last_level = level; u->last_angle = angle; u->last_img_row = img_row; u->last_rows = rows; u->last_x0 = x0; u->last_out = out;
  ex.u = u;
  u->ev_used = 0;
  u->last_stash = u->stash;
  int rc = walk(u, ex, reinterpret_cast<const uint8_t*>(packed), images, x0, level, angle, rows, img_row, out);
  if (u->profiling) prof_mark(u, ex.st, -1);
  u->launches = ex.launches;
  return rc;
}

extern "C" __attribute__((visibility("default"))) int vf_unet_set_capacity(vf_unet* u, int max_images) {
  VF_REQUIRE(u && max_images >= 0, "vf_unet_set_capacity: bad args");
  u->capacity = max_images;
  return VF_OK;
}
extern "C" __attribute__((visibility("default"))) unsigned long long vf_unet_forward_generation(const vf_unet* u) { return u ? u->fwd_gen : 0; }

extern "C" __attribute__((visibility("default"))) int vf_unet_set_stash(vf_unet* u, int on) {
  VF_REQUIRE(u, "vf_unet_set_stash: null plan");
  u->stash = on != 0;
  return VF_OK;
}

extern "C" __attribute__((visibility("default"))) int vf_unet_set_profiling(vf_unet* u, int on) {
  VF_REQUIRE(u, "vf_unet_set_profiling: null plan");
  u->profiling = on != 0;
  u->ev_used = 0;
  vf::pdl_set_suspended(on != 0);      // per-launch event times are only meaningful when the launches do not overlap
  return VF_OK;
}

// Synchronises the last recorded event and sums the elapsed ms per kernel class of the last profiled forward:
// ms[0..5] = conv, gn_stats, gn_apply, attention, upsample, embed; counts[] = launches of each class.
extern "C" __attribute__((visibility("default"))) int vf_unet_profile_read(vf_unet* u, float* ms_host, int* counts_host) {
  VF_REQUIRE(u && ms_host && counts_host, "vf_unet_profile_read: null args");
  for (int k = 0; k < K_NUM; ++k) { ms_host[k] = 0.f; counts_host[k] = 0; }
  if (u->ev_used < 2) return VF_OK;
  VF_CUDA(cudaEventSynchronize(u->ev[u->ev_used - 1]));
  for (int i = 0; i + 1 < u->ev_used; ++i) {
    const int k = u->ev_kind[i];
    if (k < 0 || k >= K_NUM) continue;
    float ms = 0.f;
    VF_CUDA(cudaEventElapsedTime(&ms, u->ev[i], u->ev[i + 1]));
    ms_host[k] += ms;
    counts_host[k] += 1;
  }
  return VF_OK;
}

// Per-launch table of the last profiled forward: ms[i], kind[i] (kernel class) and desc[6*i..] = images, H, C_in (conv:
// K total), C_out, ksize, stride.  Returns the number of launches (<= cap written), negative on error.
extern "C" __attribute__((visibility("default"))) int vf_unet_profile_launches(vf_unet* u, float* ms_host, int* kind_host, int* desc_host, int cap) {
  VF_REQUIRE(u && ms_host && kind_host && desc_host, "vf_unet_profile_launches: null args");
  if (u->ev_used < 2) return 0;
  VF_CUDA(cudaEventSynchronize(u->ev[u->ev_used - 1]));
  int n = 0;
  for (int i = 0; i + 1 < u->ev_used && n < cap; ++i, ++n) {
    VF_CUDA(cudaEventElapsedTime(&ms_host[n], u->ev[i], u->ev[i + 1]));
    kind_host[n] = u->ev_kind[i];
    for (int j = 0; j < 6; ++j) desc_host[6 * n + j] = u->ev_desc[i][j];
  }
  return n;
}

extern "C" __attribute__((visibility("default"))) int vf_unet_read_tap(vf_unet* u, const void* workspace, const char* name, float* dst, int64_t* chw, vf_stream stream) {
  VF_REQUIRE(u && workspace && name && dst, "vf_unet_read_tap: null args");
  auto it = u->taps.find(name);
  VF_REQUIRE(it != u->taps.end(), "vf_unet_read_tap: no module '%s' in the last forward", name);
  const vf_unet::Tap& t = it->second;
  const size_t total = (size_t)u->last_images * t.C * t.H * t.W;
  const uint8_t* src = reinterpret_cast<const uint8_t*>(workspace) + t.off;
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (t.dtype == VF_BF16)
    tap_to_nchw_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)src, t.ld, t.C, t.H, t.W, total, dst);
  else
    tap_to_nchw_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((const float*)src, t.ld, t.C, t.H, t.W, total, dst);
  VF_LAUNCH_CHECK();
  if (chw) { chw[0] = t.C; chw[1] = t.H; chw[2] = t.W; }
  return VF_OK;
}

// =====================================================================================================================
// Training: backward of the last vf_unet_forward (reference: loss.backward(), experiment.py:292)
// =====================================================================================================================
namespace vf {

// Bias / embedding gradients of one convolution: per-image column sums of dY [rows, ld] (padding rows hold zeros),
//   db0 / db1 [n] += sum_rows dY[row][n];   demb[img_row[img]][col + n] += sum_rows(img) dY[row][n]
// thread = (16-byte column vector, row lane); four independent loads in flight; block-level reduction in smem,
// then one atomic per (CTA, column) and destination.
template <typename T>
__global__ void __launch_bounds__(256) colsum_bias_kernel(const T* __restrict__ dy, int ld, int cout, int rows_per_img, int rows_per_cta,
                                                          float* db0, float* db1, float* demb, const int* __restrict__ img_row, int emb_ld,
                                                          int col) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int VEC = VecOf<T>::N;
  constexpr int UN = 4;
  extern __shared__ float red[];                 // [PY][CV * VEC] per-row-lane partial sums
  const int CV = (cout + VEC - 1) / VEC, PY = blockDim.x / CV;
  const int img = blockIdx.y;
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(rows_per_img, r0 + rows_per_cta);
  const int cv = threadIdx.x % CV, py = threadIdx.x / CV;
  {
    const T* base = dy + (size_t)img * rows_per_img * ld + cv * VEC;
    float acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
    for (int rb = r0 + py; rb < r1; rb += UN * PY) {
      uint4 raw[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u)
        if (rb + u * PY < r1) raw[u] = *reinterpret_cast<const uint4*>(base + (size_t)(rb + u * PY) * ld);
#pragma unroll
      for (int u = 0; u < UN; ++u)
        if (rb + u * PY < r1) {
          float v[VEC];
          load_vec(reinterpret_cast<const T*>(&raw[u]), v);
#pragma unroll
          for (int j = 0; j < VEC; ++j) acc[j] += v[j];
        }
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) red[(py * CV + cv) * VEC + j] = acc[j];
  }
  __syncthreads();
  const int row = demb ? __ldg(img_row + img) : 0;
  for (int n = threadIdx.x; n < cout; n += blockDim.x) {
    float v = 0.f;
    for (int q = 0; q < PY; ++q) v += red[q * CV * VEC + n];
    if (db0) atomicAdd(db0 + n, v);
    if (db1) atomicAdd(db1 + n, v);
    if (demb) atomicAdd(demb + (size_t)row * emb_ld + col + n, v);
  }
}

// [M, 8] fp32 FLAT gradient of the UNet output -> [rows_p, ld] activation dtype, PADDED, zero padding rows / channels.
// One thread per (padded row, group of 8 channels): 32-byte read, 16/32-byte write.
template <typename T>
__global__ void __launch_bounds__(256) grad8_to_padded_kernel(const float* __restrict__ g8, int H, int W, int ld, size_t total, T* __restrict__ dst) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // over (padded row, channel octet)
  if (gid >= total) return;
  const int oct = ld / 8;
  const int o = (int)(gid % oct);
  const size_t r = gid / oct;
  const int W1 = W + 1, P = (H + 1) * W1;
  const int rem = (int)(r % P);
  const size_t img = r / P;
  const int yy = rem / W1, xx = rem - yy * W1;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = 0.f;
  if (yy > 0 && xx > 0 && o == 0) {
    const float4* src = reinterpret_cast<const float4*>(g8 + ((img * H + (yy - 1)) * W + (xx - 1)) * 8);
    const float4 a = __ldg(src), b = __ldg(src + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  T* out = dst + r * (size_t)ld + o * 8;
  if constexpr (sizeof(T) == 2) {
    store_vec(out, v);
  } else {
    float lo[4] = {v[0], v[1], v[2], v[3]}, hi[4] = {v[4], v[5], v[6], v[7]};
    store_vec(out, lo);
    store_vec(out + 4, hi);
  }
}

// dst_j[i] += src_j[i] for up to 64 small (destination, source, count) jobs in ONE launch (grid.y = job): the per-block slices of the
// embedding-table gradient go to the 30 FeatureWiseAffine weight / bias gradients (was 60 launches of axpy_f32_kernel per step)
struct AxpyJobs { float* dst[64]; const float* src[64]; int n[64]; };
__global__ void __launch_bounds__(256) multi_axpy_kernel(const AxpyJobs jobs) {
  const int j = blockIdx.y;
  float* __restrict__ d = jobs.dst[j];
  const float* __restrict__ s = jobs.src[j];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < jobs.n[j]; i += gridDim.x * blockDim.x) d[i] += s[i];
}

// Backward of vf_embed in two launches (no atomics):
//   embed_bwd_rows_kernel    one CTA per embedding row: recompute pe / a / h / t, then dt = Ew^T demb, da = (W2^T dt) * swish'(a);
//                            writes the per-row vectors [pe(ic) | h(4ic) | t(ic) | dt(ic) | da(4ic)] to scratch
//   embed_bwd_params_kernel  one thread per parameter-gradient element: sums the rank-1 contributions over the rows
//                            dEw[e][i] = sum demb[e] t[i]; dEb[e] = sum demb[e]; dW2[o][i] = sum dt[o] h[i]; db2 = sum dt;
//                            dW0[o][i] = sum da[o] pe[i]; db0 = sum da
constexpr int kEmbBwdThreads = 1024;   // the E-long reduction below is a chain of dependent L2 loads per thread: 16 thread groups instead of 4
constexpr int kEmbBwdGroups = 16;
__global__ void __launch_bounds__(kEmbBwdThreads) embed_bwd_rows_kernel(const float* __restrict__ level, const float* __restrict__ angle, int ic,
                                                             const float* __restrict__ w0, const float* __restrict__ b0,
                                                             const float* __restrict__ w2, const float* __restrict__ b2,
                                                             const float* __restrict__ ew, int E, const float* __restrict__ demb,
                                                             float* __restrict__ rowbuf) {
  extern __shared__ float sm[];
  float* pe = sm;              // [ic]
  float* av = pe + ic;         // [4ic] pre-activation
  float* hid = av + 4 * ic;    // [4ic]
  float* dt = hid + 4 * ic;    // [ic]
  float* part = dt + ic;       // [kEmbBwdGroups][ic] partial dt
  const int row = blockIdx.x;
  const int half = ic / 2, cnt = ic / 4;
  float* out = rowbuf + (size_t)row * 11 * ic;
  for (int i = threadIdx.x; i < ic; i += blockDim.x) {
    const float x = i < half ? __ldg(level + row) : __ldg(angle + row);
    const int j = i % half, k = j % cnt;
    const float f = expf(-9.210340371976184f * ((float)k / (float)cnt));
    pe[i] = j < cnt ? sinf(x * f) : cosf(x * f);
    out[i] = pe[i];
  }
  __syncthreads();
  for (int o = threadIdx.x; o < 4 * ic; o += blockDim.x) {
    float acc = __ldg(b0 + o);
    for (int i = 0; i < ic; ++i) acc += __ldg(w0 + (size_t)o * ic + i) * pe[i];
    av[o] = acc;
    hid[o] = acc / (1.f + expf(-acc));
    out[ic + o] = hid[o];
  }
  __syncthreads();
  for (int o = threadIdx.x; o < ic; o += blockDim.x) {
    float acc = __ldg(b2 + o);
    for (int i = 0; i < 4 * ic; ++i) acc += __ldg(w2 + (size_t)o * 4 * ic + i) * hid[i];
    out[5 * ic + o] = acc;
  }
  // dt[i] = sum_e demb[e] * Ew[e][i]: the E range is split over blockDim/ic thread groups (coalesced over i)
  const float* g = demb + (size_t)row * E;
  {
    const int ngrp = blockDim.x / ic > 0 ? (int)(blockDim.x / ic) : 1;
    const int grp = threadIdx.x / ic, i = threadIdx.x % ic;
    if (grp < ngrp && grp < kEmbBwdGroups) {
      const int gcount = ngrp < kEmbBwdGroups ? ngrp : kEmbBwdGroups;
      float acc = 0.f;
#pragma unroll 4
      for (int e = grp; e < E; e += gcount) acc += __ldg(g + e) * __ldg(ew + (size_t)e * ic + i);
      part[grp * ic + i] = acc;
    }
    __syncthreads();
    const int gcount = ngrp < kEmbBwdGroups ? ngrp : kEmbBwdGroups;
    for (int k = threadIdx.x; k < ic; k += blockDim.x) {
      float acc = 0.f;
      for (int q = 0; q < gcount; ++q) acc += part[q * ic + k];
      dt[k] = acc;
      out[6 * ic + k] = acc;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 4 * ic; i += blockDim.x) {
    float acc = 0.f;
    for (int o = 0; o < ic; ++o) acc += dt[o] * __ldg(w2 + (size_t)o * 4 * ic + i);
    const float a = av[i], sg = 1.f / (1.f + expf(-a));
    out[7 * ic + i] = acc * sg * (1.f + a * (1.f - sg));
  }
}

__global__ void __launch_bounds__(256) embed_bwd_params_kernel(const float* __restrict__ rowbuf, const float* __restrict__ demb, int rows,
                                                               int ic, int E, float* __restrict__ dew, float* __restrict__ deb,
                                                               float* dw0, float* db0, float* dw2, float* db2) {
  const size_t n_ew = (size_t)E * ic, n_w2 = (size_t)ic * 4 * ic, n_w0 = (size_t)4 * ic * ic;
  const size_t total = n_ew + E + n_w2 + ic + n_w0 + 4 * ic;
  size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int R = 11 * ic;
  float acc = 0.f;
  if (gid < n_ew) {
    const int e = (int)(gid / ic), i = (int)(gid % ic);
    for (int r = 0; r < rows; ++r) acc += __ldg(demb + (size_t)r * E + e) * __ldg(rowbuf + (size_t)r * R + 5 * ic + i);
    dew[gid] = acc;
    return;
  }
  gid -= n_ew;
  if (gid < (size_t)E) {
    for (int r = 0; r < rows; ++r) acc += __ldg(demb + (size_t)r * E + gid);
    deb[gid] = acc;
    return;
  }
  gid -= E;
  if (gid < n_w2) {
    const int o = (int)(gid / (4 * ic)), i = (int)(gid % (4 * ic));
    for (int r = 0; r < rows; ++r) acc += __ldg(rowbuf + (size_t)r * R + 6 * ic + o) * __ldg(rowbuf + (size_t)r * R + ic + i);
    dw2[gid] += acc;
    return;
  }
  gid -= n_w2;
  if (gid < (size_t)ic) {
    for (int r = 0; r < rows; ++r) acc += __ldg(rowbuf + (size_t)r * R + 6 * ic + gid);
    db2[gid] += acc;
    return;
  }
  gid -= ic;
  if (gid < n_w0) {
    const int o = (int)(gid / ic), i = (int)(gid % ic);
    for (int r = 0; r < rows; ++r) acc += __ldg(rowbuf + (size_t)r * R + 7 * ic + o) * __ldg(rowbuf + (size_t)r * R + i);
    dw0[gid] += acc;
    return;
  }
  gid -= n_w0;
  for (int r = 0; r < rows; ++r) acc += __ldg(rowbuf + (size_t)r * R + 7 * ic + gid);
  db0[gid] += acc;
}

int gn_backward_impl(const void* src0, int C0, const float* stats0, int stats0_ld, const void* src1, int C1, const float* stats1,
                     int stats1_ld, int dtype, int images, int H, int W, int groups, const float* gamma, const float* beta, int swish,
                     const void* dy, float* scratch, bool scratch_zeroed, float* dgamma, float* dbeta, void* dx0, int acc0, void* dx1,
                     int acc1, cudaStream_t st, const GnColsum* colsum, const vf_gn_shift* shift);

// ---- all packed weight gradients -> OIHW parameter gradients in ONE launch at the end of the backward -----------------
// (every convolution owns a slice of the packed-gradient arena, so nothing has to be unpacked layer by layer)
struct UnpackJob {
  const float* dwp;      // packed gradient [cout][k_total] of one convolution
  long long dw_off;      // destination = grad_base + dw_off (floats): offsets are stable across steps, the base is not
  int cout, cin, kk, k_total, k_off, cin_total, c_off;
  int total;
};
__global__ void __launch_bounds__(256) multi_unpack_kernel(const UnpackJob* __restrict__ jobs, const int* __restrict__ chunk_job,
                                                           const int* __restrict__ chunk_start, float* __restrict__ grad_base) {
  pdl_launch_dependents();
  pdl_wait();
  const UnpackJob j = jobs[chunk_job[blockIdx.x]];
  const int start = chunk_start[blockIdx.x];
  const int end = min(j.total, start + kPackChunk);
  float* dw = grad_base + j.dw_off;
  for (int gid = start + threadIdx.x; gid < end; gid += 256) {       // over the segment's [cout][cin][kk] slice
    const int tap = gid % j.kk;
    const int c = (gid / j.kk) % j.cin;
    const int n = gid / (j.kk * j.cin);
    dw[((size_t)n * j.cin_total + j.c_off + c) * j.kk + tap] += j.dwp[(size_t)n * j.k_total + j.k_off + tap * j.cin + c];
  }
}

static inline long long float_offset(const float* p, const float* base) {
  return (long long)((reinterpret_cast<intptr_t>(p) - reinterpret_cast<intptr_t>(base)) / (intptr_t)sizeof(float));
}

struct UnpackList {
  std::vector<UnpackJob> jobs;
  std::vector<int> chunk_job, chunk_start;
  void add(const float* dwp, int cout, int cin, int ksize, int k_total, int k_off, long long dw_off, int cin_total, int c_off) {
    UnpackJob j{dwp, dw_off, cout, cin, ksize * ksize, k_total, k_off, cin_total, c_off, cout * cin * ksize * ksize};
    const int id = (int)jobs.size();
    jobs.push_back(j);
    for (int s0 = 0; s0 < j.total; s0 += kPackChunk) { chunk_job.push_back(id); chunk_start.push_back(s0); }
  }
  size_t table_bytes() const { return align_up(jobs.size() * sizeof(UnpackJob), 256) + 2 * align_up(chunk_job.size() * sizeof(int), 256); }
};

// Backward of vf_embed (two launches, no atomics): demb [rows, E] -> dew [E, ic] / deb [E] (overwritten) and += into the
// noise_level_mlp gradients.  rowbuf: rows * 11 * ic floats of scratch.
static int embed_backward_launch(const float* level, const float* angle, int rows, int ic, const float* w0, const float* b0, const float* w2,
                                 const float* b2, const float* emb_w, int E, const float* demb, float* rowbuf, float* dew, float* deb, float* dw0,
                                 float* db0, float* dw2, float* db2, cudaStream_t st) {
  embed_bwd_rows_kernel<<<rows, kEmbBwdThreads, (size_t)(10 + kEmbBwdGroups) * ic * sizeof(float), st>>>(level, angle, ic, w0, b0, w2, b2, emb_w, E, demb, rowbuf);
  const size_t total = (size_t)E * ic + E + (size_t)8 * ic * ic + 5 * ic;
  embed_bwd_params_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(rowbuf, demb, rows, ic, E, dew, deb, dw0, db0, dw2, db2);
  VF_LAUNCH_CHECK();
  return VF_OK;
}

struct BwdCtx {
  vf_unet* u;
  cudaStream_t st;
  uint8_t* gbase;
  UnpackList unpack;
  size_t goff = 0, gcap = 0;
  size_t li = 0;             // layout image count (u->last_layout): buffers are sized for it, kernels run over u->last_images
  bool dry = false;
  int rc = VF_OK;
  std::map<const void*, std::pair<void*, bool>> grads;     // forward tensor -> (gradient buffer, written?)
  // resumable walk (vf_unet_backward_phase): scratch pointers and the position in the reversed tape
  float *dwp = nullptr, *cs = nullptr, *demb = nullptr, *dew = nullptr, *deb = nullptr, *emb_rows = nullptr, *gn_scratch = nullptr, *att_scratch = nullptr;
  uint8_t* unpack_tab = nullptr;
  void* g_out = nullptr;
  float* dwp_cur = nullptr;
  float* gn_cur = nullptr;
  int colsum_done = -1;          // tape index of the convolution whose bias / embedding gradients the GroupNorm backward produced
  int next_op = -1;              // next tape index to differentiate (counts down to 0)
  size_t chunks_done = 0;        // unpack chunks already launched
  void* galloc(size_t bytes) {
    goff = align_up(goff, 256);
    void* p = dry ? nullptr : gbase + goff;
    goff += bytes;
    return p;
  }
  std::pair<void*, bool>& grad_of(const void* fwd) {
    auto it = grads.find(fwd);
    if (it != grads.end()) return it->second;
    size_t bytes = 0;
    auto ab = u->act_bytes.find(fwd);
    if (ab != u->act_bytes.end()) bytes = ab->second;
    auto& e = grads[fwd];
    e.first = galloc(bytes);
    e.second = false;
    return e;
  }
};

#define VF_B(call)                                  \
  do {                                              \
    if (!cx.dry && cx.rc == VF_OK) cx.rc = (call);  \
  } while (0)

// One launch: per-image column sums of dY [images*rows_per_img, ld] (first `cout` columns) added into db0 / db1 [cout] and
// into demb[img_row[img]][col + n].
static int colsum_bias_launch(const void* dy, int dt, int images, int rows_per_img, int ld, int cout, float* db0, float* db1, float* demb,
                              const int* img_row, int emb_ld, int col, cudaStream_t st) {
  const int vec = dt == VF_BF16 ? 8 : 4;
  const int cv = (cout + vec - 1) / vec;
  VF_REQUIRE(cv >= 1 && cv <= 256 && ld % vec == 0, "colsum_bias: cout=%d ld=%d", cout, ld);
  const int py = 256 / cv;
  int per = py * 4 * 8;                                   // >= 8 unrolled iterations per thread
  if (per < 512) per = 512;
  dim3 grid(cdiv(rows_per_img, per), images);
  const size_t smem = (size_t)py * cv * vec * sizeof(float);
  cudaError_t le;
  if (dt == VF_BF16)
    le = launch_pdl(colsum_bias_kernel<__nv_bfloat16>, grid, dim3(cv * py), smem, st, (const __nv_bfloat16*)dy, ld, cout, rows_per_img, per, db0, db1,
                    demb, img_row, emb_ld, col);
  else
    le = launch_pdl(colsum_bias_kernel<float>, grid, dim3(cv * py), smem, st, (const float*)dy, ld, cout, rows_per_img, per, db0, db1, demb, img_row,
                    emb_ld, col);
  if (le != cudaSuccess) { set_error("colsum_bias launch: %s", cudaGetErrorString(le)); return VF_ERR_CUDA; }
  return VF_OK;
}

static void conv_backward(BwdCtx& cx, const vf_unet::TapeOp& t, const uint8_t* pkt, const void* dY, int dy_ld, float* dwp, float* cs,
                          float* const* pg, float* demb, bool skip_colsum = false) {
  vf_unet* u = cx.u;
  const vf_conv_args& f = t.conv;
  const int dt = u->dtype;
  const size_t es = dtype_size(dt);
  vf_conv_args a = f;                              // geometry for the gradient problems
  const void* dYs = dY;
  int H = f.H, W = f.W;
  const int images = f.images;
  if (f.stride == 2) {
    // Downsample: scatter dY (H/2 x W/2) onto the even pixels of a zeroed H x W grid -> stride-1 problems at source resolution
    void* z = cx.galloc(cx.li * (H + 1) * (W + 1) * dy_ld * es);
    VF_B(vf_zero_insert2x(dY, dt, images, H / 2, W / 2, dy_ld, z, (vf_stream)cx.st));
    dYs = z;
    a.stride = 1;
  }
  const int out_rows_per_img = a.out_padded ? (H + 1) * (W + 1) : H * W;
  // ---- bias / embedding gradients: per-image column sums of dY
  if ((t.b_idx[0] >= 0 || t.emb_col >= 0) && !skip_colsum) {
    if (!cx.dry && cx.rc == VF_OK) {
      float* db0 = t.b_idx[0] >= 0 ? pg[t.b_idx[0]] : nullptr;
      float* db1 = t.b_idx[1] >= 0 ? pg[t.b_idx[1]] : nullptr;
      float* de = t.emb_col >= 0 ? demb : nullptr;
      const int col = t.emb_col >= 0 ? t.emb_col : 0;
      cx.rc = colsum_bias_launch(dYs, dt, images, out_rows_per_img, dy_ld, f.cout, db0, db1, de, (const int*)u->last_img_row, u->E, col, cx.st);
    }
  }
  // ---- weight gradient into the packed scratch, then scatter to the OIHW parameter gradients
  int k_total = 0;
  for (int s = 0; s < f.n_seg; ++s) k_total += f.ksize[s] * f.ksize[s] * f.src_c[s];
  {
    vf_conv_args aw = a;
    const void* dYw = dYs;
    bool all_1x1 = true;
    for (int s = 0; s < f.n_seg; ++s) all_1x1 = all_1x1 && f.ksize[s] == 1;
    if (dt == VF_BF16 && all_1x1 && (a.in_padded != 0) != (a.out_padded != 0)) {
      // 1x1 convolution between the two row orders (qkv, attention out-projection, first layer): bring dY into the
      // row order of X so the tensor-core weight-gradient GEMM sees one row index on both operands
      if (a.in_padded) {
        void* z = cx.galloc(cx.li * (H + 1) * (W + 1) * dy_ld * es);
        VF_B(vf_flat_to_padded(dYs, dt, images, H, W, dy_ld, z, (vf_stream)cx.st));
        dYw = z;
      } else {
        void* z = cx.galloc(cx.li * H * W * dy_ld * es);
        VF_B(vf_padded_to_flat(dYs, dt, images, H, W, dy_ld, z, (vf_stream)cx.st));
        dYw = z;
      }
      aw.out_padded = a.in_padded;
    }
    VF_B(vf_conv2d_wgrad(&aw, dYw, dy_ld, dwp, (vf_stream)cx.st));
  }
  int koff = 0;
  for (int s = 0; s < f.n_seg; ++s) {
    const int kk = f.ksize[s] * f.ksize[s];
    if (t.w_idx[s] >= 0) {
      if (s == 0 && f.src[0] == u->last_x0) {
        // first layer: the K0 columns are (tap, channel) of the im2col'd input with cin_total real channels
        if (!cx.dry) cx.unpack.add(dwp, f.cout, t.cin_total[0], 3, k_total, 0, float_offset(pg[t.w_idx[0]], pg[0]), t.cin_total[0], 0);
      } else {
        if (!cx.dry) cx.unpack.add(dwp, f.cout, f.src_c[s], f.ksize[s], k_total, koff, float_offset(pg[t.w_idx[s]], pg[0]), t.cin_total[s], t.c_off[s]);
      }
    }
    koff += kk * f.src_c[s];
  }
  // ---- data gradients
  koff = 0;
  for (int s = 0; s < f.n_seg; ++s) {
    const bool first_layer = s == 0 && f.src[0] == u->last_x0;
    if (!first_layer) {
      auto& g = cx.grad_of(f.src[s]);
      const size_t n_elems = (size_t)images * (a.in_padded ? (H + 1) * (W + 1) : H * W) * f.src_c[s];
      if (t.w_idx[s] < 0) {
        // identity segment: grad(x) (+)= dY
        if (g.second) VF_B(vf_add_inplace(g.first, dYs, dt, n_elems, (vf_stream)cx.st));
        else if (!cx.dry && cx.rc == VF_OK && cudaMemcpyAsync(g.first, dYs, n_elems * es, cudaMemcpyDeviceToDevice, cx.st) != cudaSuccess) cx.rc = VF_ERR_CUDA;
      } else {
        vf_conv_args d{};
        d.dtype = dt; d.images = images; d.H = H; d.W = W;
        d.in_padded = a.out_padded; d.out_padded = a.in_padded;
        d.n_seg = 1; d.src[0] = dYs; d.src_c[0] = dy_ld; d.ksize[0] = f.ksize[s]; d.stride = 1;
        d.weight = pkt + t.wt_off[s]; d.cout = f.src_c[s]; d.cout_pad = f.src_c[s];
        d.residual = g.second ? g.first : nullptr;
        d.out = g.first; d.out_dtype = dt; d.out_ld = f.src_c[s];
        VF_B(vf_conv2d(&d, (vf_stream)cx.st));
      }
      g.second = true;
    }
    koff += f.ksize[s] * f.ksize[s] * f.src_c[s];
  }
  // ---- epilogue residual (attention: out + input): gradient flows through unchanged
  if (f.residual) {
    auto& g = cx.grad_of(f.residual);
    const size_t n_elems = (size_t)images * out_rows_per_img * f.cout;
    if (g.second) VF_B(vf_add_inplace(g.first, dYs, dt, n_elems, (vf_stream)cx.st));
    else if (!cx.dry && cx.rc == VF_OK && cudaMemcpyAsync(g.first, dYs, n_elems * es, cudaMemcpyDeviceToDevice, cx.st) != cudaSuccess) cx.rc = VF_ERR_CUDA;
    g.second = true;
  }
}

}  // namespace vf

namespace vf {

// fixed layout of the unpack job table inside its reserved region: jobs | chunk -> job | chunk -> first element
constexpr size_t kUnpackJobsBytes = 96 * 1024, kUnpackChunkBytes = 320 * 1024;
static_assert(kUnpackJobsBytes + 2 * kUnpackChunkBytes <= kPackTableBytes, "unpack table regions");

// ---- the backward as a resumable walk: begin -> ops (reversed tape, possibly in several calls) -> unpack -> finish ------
static void bw_begin(BwdCtx& cx, const float* g8) {
  vf_unet* u = cx.u;
  const int dt = u->dtype;
  const size_t es = dtype_size(dt);
  const int images = u->last_images, S = u->cfg.image_size;
  const vf_unet_config& c = u->cfg;
  float *&dwp = cx.dwp, *&cs = cx.cs, *&demb = cx.demb, *&dew = cx.dew, *&deb = cx.deb, *&emb_rows = cx.emb_rows, *&gn_scratch = cx.gn_scratch,
        *&att_scratch = cx.att_scratch;
  uint8_t*& unpack_tab = cx.unpack_tab;
  void*& g_out = cx.g_out;
  // scratch: packed weight gradient (largest conv), per-image column sums, embedding-table gradient
  // Every convolution gets its own slice of ONE zero-filled packed-gradient arena and every GroupNorm its own slice of
  // one zero-filled reduction arena: two memsets per backward instead of one per layer, and no memset nodes between
  // the kernels of the chain (they would break the programmatic dependent launches).
  size_t dwp_floats = 0, cs_floats = 0;
  for (auto& t : u->tape)
    if (t.kind == 0) {
      size_t k = 0;
      for (int s = 0; s < t.conv.n_seg; ++s) k += (size_t)t.conv.ksize[s] * t.conv.ksize[s] * t.conv.src_c[s];
      dwp_floats += align_up((size_t)t.conv.cout_pad * k, 64);
      cs_floats = std::max(cs_floats, cx.li * t.conv.cout);
    }
  dwp = (float*)cx.galloc(dwp_floats * 4);
  cs = (float*)cx.galloc(cs_floats * 4);
  demb = (float*)cx.galloc(cx.li * u->E * 4);
  dew = (float*)cx.galloc((size_t)u->E * c.inner_channel * 4);
  deb = (float*)cx.galloc((size_t)u->E * 4);
  emb_rows = (float*)cx.galloc(cx.li * 11 * c.inner_channel * 4);
  size_t gn_floats = 0, att_floats = 0;
  for (auto& t : u->tape) {
    if (t.kind == 1) gn_floats += align_up(cx.li * (t.gC0 + t.gC1) * 2, 64);
    if (t.kind == 2) att_floats = std::max(att_floats, cx.li * t.aL * 2 * t.aC * std::max(1, t.aL / 128));
  }
  gn_scratch = (float*)cx.galloc(gn_floats * 4);
  att_scratch = (float*)cx.galloc(att_floats * 4);
  unpack_tab = (uint8_t*)cx.galloc(kPackTableBytes);     // job table of the multi-tensor unpack launches
  if (!cx.dry) {
    cudaMemsetAsync(dwp, 0, dwp_floats * 4, cx.st);
    cudaMemsetAsync(gn_scratch, 0, gn_floats * 4, cx.st);
    cudaMemsetAsync(demb, 0, (size_t)u->last_rows * u->E * 4, cx.st);
    cudaMemsetAsync(dew, 0, (size_t)u->E * c.inner_channel * 4, cx.st);
    cudaMemsetAsync(deb, 0, (size_t)u->E * 4, cx.st);
  }
  // gradient of the UNet output -> PADDED activation-dtype matrix with final_npad channels
  const int np = u->final_npad;
  g_out = cx.galloc(cx.li * (S + 1) * (S + 1) * np * es);
  if (!cx.dry) {
    const size_t total = (size_t)images * (S + 1) * (S + 1) * (np / 8);
    const unsigned grid = (unsigned)((total + 255) / 256);
    if (dt == VF_BF16) launch_pdl(grad8_to_padded_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, cx.st, g8, S, S, np, total, (__nv_bfloat16*)g_out);
    else launch_pdl(grad8_to_padded_kernel<float>, dim3(grid), dim3(256), 0, cx.st, g8, S, S, np, total, (float*)g_out);
  }
  cx.dwp_cur = dwp;
  cx.gn_cur = gn_scratch;
  cx.colsum_done = -1;
  cx.next_op = (int)u->tape.size() - 1;
  cx.chunks_done = 0;
}

// differentiate tape ops next_op, next_op - 1, ..., stop_at
static void bw_run_ops(BwdCtx& cx, const uint8_t* pkt, float* const* pg, int stop_at) {
  vf_unet* u = cx.u;
  const int dt = u->dtype;
  const int images = u->last_images;
  const vf_unet_config& c = u->cfg;
  const int np = u->final_npad;
  float *&dwp_cur = cx.dwp_cur, *&gn_cur = cx.gn_cur;
  int& colsum_done = cx.colsum_done;
  float* const cs = cx.cs;
  float* const demb = cx.demb;
  float* const att_scratch = cx.att_scratch;
  void* const g_out = cx.g_out;
  for (int& i = cx.next_op; i >= stop_at && cx.rc == VF_OK; --i) {
    const vf_unet::TapeOp& t = u->tape[i];
    if (t.kind == 0) {
      float* dwp_l = dwp_cur;
      {
        size_t k = 0;
        for (int s = 0; s < t.conv.n_seg; ++s) k += (size_t)t.conv.ksize[s] * t.conv.ksize[s] * t.conv.src_c[s];
        dwp_cur += align_up((size_t)t.conv.cout_pad * k, 64);
      }
      if (t.conv.out == (void*)u->last_out) {
        vf_unet::TapeOp tf = t;               // final conv: its gradient arrives PADDED with np channels
        tf.conv.out_padded = 1;
        tf.conv.cout = np; tf.conv.cout_pad = np;
        // only the real output channels have parameters: unpack / bias use the true cout below
        vf_unet::TapeOp tt = tf;
        tt.conv.cout = c.out_channel; tt.conv.cout_pad = np;
        conv_backward(cx, tt, pkt, g_out, np, dwp_l, cs, pg, demb);
      } else {
        auto& g = cx.grad_of(t.conv.out);
        const int ld = t.conv.qkv_split ? 3 * t.conv.qkv_split : t.conv.out_ld;
        conv_backward(cx, t, pkt, g.first, ld, dwp_l, cs, pg, demb, colsum_done == i);
      }
    } else if (t.kind == 1) {
      auto& gy = cx.grad_of(t.gdst);
      auto& g0 = cx.grad_of(t.gsrc0);
      std::pair<void*, bool>* g1 = t.gsrc1 ? &cx.grad_of(t.gsrc1) : nullptr;
      // The first convolution of a ResnetBlock (the one with the embedding add) feeds this GroupNorm and nothing else, so
      // its bias / embedding gradients (per-image column sums of dx) come out of the GroupNorm backward in closed form and
      // conv_backward skips its pass over dY.
      GnColsum csum{};
      bool fused_cs = false;
      if (i > 0 && !t.gsrc1 && !g0.second) {
        const vf_unet::TapeOp& pc = u->tape[i - 1];
        if (pc.kind == 0 && pc.conv.out == t.gsrc0 && pc.emb_col >= 0 && pc.b_idx[1] < 0 && pc.conv.stride == 1 && pc.conv.cout == t.gC0 &&
            !(vf::tc_debug_flags() & 2048)) {
          csum.db = (!cx.dry && pc.b_idx[0] >= 0) ? pg[pc.b_idx[0]] : nullptr;      // the sizing walk has no gradient table
          csum.demb = demb; csum.img_row = (const int*)u->last_img_row; csum.emb_ld = u->E; csum.col = pc.emb_col;
          fused_cs = true;
          colsum_done = i - 1;
        }
      }
      VF_B(gn_backward_impl(t.gsrc0, t.gC0, t.gst0, t.gld0, t.gsrc1, t.gC1, t.gst1, t.gld1, dt, images, t.gH, t.gW, c.norm_groups,
                            u->master[t.gw], u->master[t.gb], t.swish, gy.first, gn_cur, true, pg[t.gw], pg[t.gb], g0.first, g0.second ? 1 : 0,
                            g1 ? g1->first : nullptr, g1 && g1->second ? 1 : 0, cx.st, fused_cs ? &csum : nullptr, &t.gshift));
      gn_cur += align_up(cx.li * (t.gC0 + t.gC1) * 2, 64);
      g0.second = true;
      if (g1) g1->second = true;
    } else if (t.kind == 2) {
      auto& go = cx.grad_of(t.o);
      auto& gq = cx.grad_of(t.qkv);
      VF_B(vf_attention_backward(t.qkv, t.vt, t.o, t.lse, go.first, dt, images, t.aL, t.aC, att_scratch, gq.first, (vf_stream)cx.st));
      gq.second = true;
    } else if (t.kind == 3) {
      auto& gd = cx.grad_of(t.udst);
      auto& gs = cx.grad_of(t.usrc);
      VF_B(vf_upsample2x_backward(gd.first, dt, images, t.uH, t.uW, t.uC, gs.first, gs.second ? 1 : 0, (vf_stream)cx.st));
      gs.second = true;
    }
  }
}

// packed weight gradients produced since the last call -> OIHW parameter gradients (one launch).  The job table has a fixed
// layout in its region and is identical from step to step, so in steady state nothing is uploaded: only slices that differ
// from what the device already holds are copied.
static void bw_flush_unpack(BwdCtx& cx, float* const* pg) {
  vf_unet* u = cx.u;
  if (cx.dry || cx.rc != VF_OK) return;
  const UnpackList& ul = cx.unpack;
  const size_t n_chunks = ul.chunk_job.size();
  if (n_chunks <= cx.chunks_done) return;
  if (ul.jobs.size() * sizeof(UnpackJob) > kUnpackJobsBytes || n_chunks * sizeof(int) > kUnpackChunkBytes) {
    set_error("vf_unet_backward: unpack table (%zu jobs, %zu chunks) exceeds its reserved region", ul.jobs.size(), n_chunks);
    cx.rc = VF_ERR_ARG;
    return;
  }
  uint8_t* tab = cx.unpack_tab;
  std::vector<uint8_t>& cache = u->unpack_cache;           // host image of the device table + the table's address
  if (cache.size() != kPackTableBytes + sizeof(void*) || memcmp(cache.data() + kPackTableBytes, &tab, sizeof(void*)) != 0) {
    cache.assign(kPackTableBytes + sizeof(void*), 0xFF);   // other location (or first use): nothing on the device is valid
    memcpy(cache.data() + kPackTableBytes, &tab, sizeof(void*));
  }
  auto sync_slice = [&](size_t off, const void* src, size_t bytes) {
    if (bytes == 0 || cx.rc != VF_OK) return;
    if (memcmp(cache.data() + off, src, bytes) == 0) return;
    if (cudaMemcpyAsync(tab + off, src, bytes, cudaMemcpyHostToDevice, cx.st) != cudaSuccess) { cx.rc = VF_ERR_CUDA; return; }   // pageable: staged before returning
    memcpy(cache.data() + off, src, bytes);
  };
  sync_slice(0, ul.jobs.data(), ul.jobs.size() * sizeof(UnpackJob));
  sync_slice(kUnpackJobsBytes + cx.chunks_done * sizeof(int), ul.chunk_job.data() + cx.chunks_done, (n_chunks - cx.chunks_done) * sizeof(int));
  sync_slice(kUnpackJobsBytes + kUnpackChunkBytes + cx.chunks_done * sizeof(int), ul.chunk_start.data() + cx.chunks_done,
             (n_chunks - cx.chunks_done) * sizeof(int));
  if (cx.rc != VF_OK) return;
  cudaError_t le = launch_pdl(multi_unpack_kernel, dim3((unsigned)(n_chunks - cx.chunks_done)), dim3(256), 0, cx.st, (const UnpackJob*)tab,
                              (const int*)(tab + kUnpackJobsBytes) + cx.chunks_done,
                              (const int*)(tab + kUnpackJobsBytes + kUnpackChunkBytes) + cx.chunks_done, pg[0]);
  if (le != cudaSuccess) { set_error("multi_unpack launch: %s", cudaGetErrorString(le)); cx.rc = VF_ERR_CUDA; }
  cx.chunks_done = n_chunks;
}

// embedding path (always last: every ResnetBlock feeds it)
static void bw_finish(BwdCtx& cx, float* const* pg) {
  vf_unet* u = cx.u;
  const vf_unet_config& c = u->cfg;
  float* const demb = cx.demb;
  float* const dew = cx.dew;
  float* const deb = cx.deb;
  float* const emb_rows = cx.emb_rows;
  // embedding path: demb [rows, E] -> noise_level_mlp and the per-block Linear(ic -> Cout) parameters
  if (!cx.dry && cx.rc == VF_OK) {
    const int ic = c.inner_channel;
    const uint8_t* pk = nullptr; (void)pk;
    const int rows = u->last_rows;
    cx.rc = embed_backward_launch(u->last_level, u->last_angle, rows, ic, u->master[u->mlp_w0], u->master[u->mlp_b0], u->master[u->mlp_w2],
                                  u->master[u->mlp_b2], u->emb_w_dev, u->E, demb, emb_rows, dew, deb, pg[u->mlp_w0], pg[u->mlp_b0], pg[u->mlp_w2],
                                  pg[u->mlp_b2], cx.st);
    AxpyJobs jobs{};
    int nj = 0, nmax = 0;
    auto flush = [&]() {
      if (nj > 0) multi_axpy_kernel<<<dim3((unsigned)((nmax + 1023) / 1024 > 0 ? (nmax + 1023) / 1024 : 1), (unsigned)nj), 256, 0, cx.st>>>(jobs);
      nj = 0; nmax = 0;
    };
    auto add = [&](float* d, const float* s2, int n) {
      if (nj == 64) flush();
      jobs.dst[nj] = d; jobs.src[nj] = s2; jobs.n[nj] = n;
      nmax = n > nmax ? n : nmax;
      ++nj;
    };
    for (auto& b : u->blocks) {
      add(pg[b.nf_w], dew + (size_t)b.emb_col * ic, b.cout * ic);
      add(pg[b.nf_b], deb + b.emb_col, b.cout);
    }
    flush();
    if (cudaGetLastError() != cudaSuccess) { set_error("vf_unet_backward: embedding backward launch failed"); cx.rc = VF_ERR_CUDA; }
  }
}

static int backward_walk(BwdCtx& cx, const uint8_t* pkt, const float* g8, float* const* pg) {
  bw_begin(cx, g8);
  bw_run_ops(cx, pkt, pg, 0);
  bw_flush_unpack(cx, pg);
  bw_finish(cx, pg);
  return cx.rc;
}

// ---- phases for overlapping the data-parallel gradient exchange with the rest of the backward ---------------------------
// The reversed tape is cut into n_phases runs of (roughly) equal parameter bytes; a parameter belongs to the phase whose ops
// produce its gradient, the embedding MLP and the per-block Linear(ic -> Cout) parameters to the extra last phase n_phases
// (their gradients are complete only after every block was differentiated).  phase_stop[k] = lowest tape index of phase k.
static void plan_phases(const vf_unet* u, int n_phases, std::vector<int>& phase_stop, std::vector<int>& param_phase) {
  const int np = (int)u->params.size();
  auto numel = [&](int idx) { size_t n = 1; for (int d = 0; d < u->params[idx].ndim; ++d) n *= (size_t)u->params[idx].shape[d]; return n; };
  param_phase.assign(np, n_phases);
  std::vector<size_t> op_bytes(u->tape.size(), 0);
  std::vector<char> seen(np, 0);
  size_t total = 0;
  auto touch = [&](int idx, size_t& acc) { if (idx >= 0 && !seen[idx]) { seen[idx] = 1; acc += numel(idx); } };
  for (int i = (int)u->tape.size() - 1; i >= 0; --i) {
    const vf_unet::TapeOp& t = u->tape[i];
    size_t b = 0;
    if (t.kind == 0) { for (int sgi = 0; sgi < 3; ++sgi) touch(t.w_idx[sgi], b); touch(t.b_idx[0], b); touch(t.b_idx[1], b); }
    if (t.kind == 1) { touch(t.gw, b); touch(t.gb, b); }
    op_bytes[i] = b;
    total += b;
  }
  phase_stop.assign(n_phases, 0);
  std::fill(seen.begin(), seen.end(), 0);
  size_t acc = 0;
  int ph = 0;
  for (int i = (int)u->tape.size() - 1; i >= 0; --i) {
    const vf_unet::TapeOp& t = u->tape[i];
    auto mark = [&](int idx) { if (idx >= 0 && !seen[idx]) { seen[idx] = 1; param_phase[idx] = ph; } };
    if (t.kind == 0) { for (int sgi = 0; sgi < 3; ++sgi) mark(t.w_idx[sgi]); mark(t.b_idx[0]); mark(t.b_idx[1]); }
    if (t.kind == 1) { mark(t.gw); mark(t.gb); }
    acc += op_bytes[i];
    // a GroupNorm that emits the bias gradient of the convolution below it (closed-form column sums) must stay in that
    // convolution's phase: cut only after a convolution
    const bool cut_ok = t.kind == 0;
    if (ph < n_phases - 1 && cut_ok && acc * (size_t)n_phases >= total * (size_t)(ph + 1)) {
      phase_stop[ph] = i;
      ++ph;
    }
  }
  for (int k = ph; k < n_phases; ++k) phase_stop[k] = 0;
}

}  // namespace vf

namespace vf {
template <typename T>
__global__ void pack_t_slice_kernel(const float* __restrict__ w, int cout, int cin, int c_off, int c_cnt, T* __restrict__ dst) {
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (size_t)c_cnt * cout) return;
  const int c = (int)(gid / cout), n = (int)(gid % cout);
  dst[gid] = from_f<T>(__ldg(w + (size_t)n * cin + c_off + c));
}
static int pack_t_slice(const float* w, int cout, int cin, int c_off, int c_cnt, int dtype, void* dst, cudaStream_t st) {
  const size_t total = (size_t)c_cnt * cout;
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (dtype == VF_BF16) pack_t_slice_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(w, cout, cin, c_off, c_cnt, (__nv_bfloat16*)dst);
  else pack_t_slice_kernel<float><<<grid, 256, 0, st>>>(w, cout, cin, c_off, c_cnt, (float*)dst);
  VF_LAUNCH_CHECK();
  return VF_OK;
}
}  // namespace vf

extern "C" __attribute__((visibility("default"))) size_t vf_unet_packed_t_bytes(const vf_unet* u) { return u ? u->packed_t_bytes : 0; }

extern "C" __attribute__((visibility("default"))) int vf_unet_pack_weights_t(vf_unet* u, void* packed_t, vf_stream stream) {
  VF_REQUIRE(u && packed_t, "vf_unet_pack_weights_t: null args");
  if (!u->packed) { set_error("vf_unet_pack_weights_t: call vf_unet_pack_weights first"); return VF_ERR_STATE; }
  uint8_t* pk = reinterpret_cast<uint8_t*>(packed_t);
  const int dt = u->dtype;
  VF_CUDA(cudaMemsetAsync(pk, 0, u->pack_t_tab_off, as_stream(stream)));
  PackList pl;
  for (auto& b : u->blocks) {
    const int cin = b.c0 + b.c1;
    pl.conv_t(u->master[b.c1_w], b.cout, cin, 3, pk + b.wt1, cin, 9 * b.cout, 0, b.cout);
    pl.conv_t(u->master[b.c2_w], b.cout, b.cout, 3, pk + b.wt2, b.cout, 9 * b.cout, 0, b.cout);
    if (b.rs_w >= 0) {
      // res_conv [cout][cin][1][1]: the x slice (channels 0..c0) and the skip slice (c0..cin) get their own transposed packs
      pl.slice_t(u->master[b.rs_w], b.cout, cin, 0, b.c0, pk + b.wtr0);
      if (b.c1) pl.slice_t(u->master[b.rs_w], b.cout, cin, b.c0, b.c1, pk + b.wtr1);
    }
    if (b.attn) {
      pl.conv_t(u->master[b.qkv_w], 3 * b.cout, b.cout, 1, pk + b.wtqkv, b.cout, 3 * b.cout, 0, 3 * b.cout);
      pl.conv_t(u->master[b.ao_w], b.cout, b.cout, 1, pk + b.wtout, b.cout, b.cout, 0, b.cout);
    }
  }
  for (auto* sec : {&u->downs, &u->mid, &u->ups})
    for (auto& l : *sec)
      if (l.kind == 2 || l.kind == 3) pl.conv_t(u->master[l.w_idx], l.c, l.c, 3, pk + l.wt, l.c, 9 * l.c, 0, l.c);
  pl.conv_t(u->master[u->fin_w], u->cfg.out_channel, u->final_c, 3, pk + u->final_wt, u->final_c, 9 * u->final_npad, 0, u->final_npad);
  {
    const int rc = pl.run(u->pack_t_cache, pk + u->pack_t_tab_off, kPackTableBytes, dt, as_stream(stream));
    if (rc != VF_OK) return rc;
  }
  u->packed_t = true;
  return VF_OK;
}

extern "C" __attribute__((visibility("default"))) size_t vf_unet_backward_workspace_bytes(vf_unet* u) {
  if (!u || u->tape.empty()) return 0;
  BwdCtx cx{u, nullptr, nullptr};
  cx.dry = true;
  cx.li = (size_t)(u->last_layout > u->last_images ? u->last_layout : u->last_images);
  backward_walk(cx, nullptr, nullptr, nullptr);
  return align_up(cx.goff, 256) + 256;
}

extern "C" __attribute__((visibility("default"))) int vf_unet_backward(vf_unet* u, const void* packed_t, void* grad_workspace, size_t grad_workspace_bytes,
                                                                     const float* grad_out8, float* const* param_grads_host, vf_stream stream) {
  VF_REQUIRE(u && packed_t && grad_workspace && grad_out8 && param_grads_host, "vf_unet_backward: null args");
  if (u->tape.empty() || !u->packed_t) { set_error("vf_unet_backward: needs a forward and vf_unet_pack_weights_t first"); return VF_ERR_STATE; }
  if (!u->last_stash) { set_error("vf_unet_backward: the last forward ran with vf_unet_set_stash(0) (inference mode)"); return VF_ERR_STATE; }
  for (size_t i = 0; i < u->params.size(); ++i) VF_REQUIRE(param_grads_host[i], "vf_unet_backward: gradient %zu (%s) is null", i, u->params[i].name.c_str());
  {
    BwdCtx dry{u, nullptr, nullptr};
    dry.dry = true;
    dry.li = (size_t)(u->last_layout > u->last_images ? u->last_layout : u->last_images);
    backward_walk(dry, nullptr, nullptr, nullptr);
    VF_REQUIRE(dry.goff <= grad_workspace_bytes, "vf_unet_backward: workspace too small (%zu < %zu)", grad_workspace_bytes, dry.goff);
  }
  BwdCtx cx{u, as_stream(stream), reinterpret_cast<uint8_t*>(grad_workspace)};
  cx.gcap = grad_workspace_bytes;
  cx.li = (size_t)(u->last_layout > u->last_images ? u->last_layout : u->last_images);
  int rc = backward_walk(cx, reinterpret_cast<const uint8_t*>(packed_t), grad_out8, param_grads_host);
  if (rc == VF_OK) VF_LAUNCH_CHECK();
  return rc;
}

// ---- stage-level exports of two backward helpers (unit parity; the plan calls the same launchers) ----------------------
extern "C" __attribute__((visibility("default"))) int vf_colsum_bias(const void* dy, int dtype, int images, int rows_per_img, int ld, int cout,
                                                                   float* db0, float* db1, float* demb, const int* img_row, int emb_ld, int col,
                                                                   vf_stream stream) {
  VF_REQUIRE(dy && images > 0 && rows_per_img > 0 && cout > 0 && (db0 || db1 || demb), "vf_colsum_bias: bad args");
  VF_REQUIRE(!demb || img_row, "vf_colsum_bias: demb needs img_row");
  return vf::colsum_bias_launch(dy, dtype, images, rows_per_img, ld, cout, db0, db1, demb, img_row, emb_ld, col, as_stream(stream));
}

extern "C" __attribute__((visibility("default"))) int vf_embed_backward(const float* level, const float* angle, int rows, int inner_channel,
                                                                      const float* w0, const float* b0, const float* w2, const float* b2,
                                                                      const float* emb_w, int E, const float* demb, float* rowbuf, float* dew,
                                                                      float* deb, float* dw0, float* db0, float* dw2, float* db2, vf_stream stream) {
  VF_REQUIRE(level && angle && w0 && b0 && w2 && b2 && emb_w && demb && rowbuf && dew && deb && dw0 && db0 && dw2 && db2 && rows > 0 && E > 0,
             "vf_embed_backward: null args");
  VF_REQUIRE(inner_channel > 0 && inner_channel % 4 == 0 && inner_channel <= 256, "vf_embed_backward: inner_channel=%d", inner_channel);
  return vf::embed_backward_launch(level, angle, rows, inner_channel, w0, b0, w2, b2, emb_w, E, demb, rowbuf, dew, deb, dw0, db0, dw2, db2,
                                   as_stream(stream));
}

// ---- phased backward: the same walk in n_phases + 1 calls, so that the caller can start the gradient exchange of a phase's
// parameters (NCCL all-reduce over NVLink) while the next phases are still running --------------------------------------
namespace vf {
void bw_state_free(vf_unet* u) {
  if (u && u->bw_state) { delete reinterpret_cast<BwdCtx*>(u->bw_state); u->bw_state = nullptr; }
}
}  // namespace vf

extern "C" __attribute__((visibility("default"))) int vf_unet_backward_plan(vf_unet* u, int n_phases, int* param_phase_out) {
  VF_REQUIRE(u && param_phase_out && n_phases >= 1 && n_phases <= 64, "vf_unet_backward_plan: bad args");
  if (u->tape.empty()) { set_error("vf_unet_backward_plan: needs a forward first (the plan follows the recorded tape)"); return VF_ERR_STATE; }
  std::vector<int> stop, phase;
  vf::plan_phases(u, n_phases, stop, phase);
  for (size_t i = 0; i < phase.size(); ++i) param_phase_out[i] = phase[i];
  return VF_OK;
}

extern "C" __attribute__((visibility("default"))) int vf_unet_backward_phase(vf_unet* u, const void* packed_t, void* grad_workspace,
                                                                           size_t grad_workspace_bytes, const float* grad_out8,
                                                                           float* const* param_grads_host, int phase, int n_phases,
                                                                           vf_stream stream) {
  VF_REQUIRE(u && packed_t && grad_workspace && grad_out8 && param_grads_host, "vf_unet_backward_phase: null args");
  VF_REQUIRE(n_phases >= 1 && n_phases <= 64 && phase >= 0 && phase <= n_phases, "vf_unet_backward_phase: phase %d of %d", phase, n_phases);
  if (phase == 0) {
    if (u->tape.empty() || !u->packed_t) { set_error("vf_unet_backward_phase: needs a forward and vf_unet_pack_weights_t first"); return VF_ERR_STATE; }
    if (!u->last_stash) { set_error("vf_unet_backward_phase: the last forward ran with vf_unet_set_stash(0) (inference mode)"); return VF_ERR_STATE; }
    for (size_t i = 0; i < u->params.size(); ++i) VF_REQUIRE(param_grads_host[i], "vf_unet_backward_phase: gradient %zu (%s) is null", i, u->params[i].name.c_str());
    {
      BwdCtx dry{u, nullptr, nullptr};
      dry.dry = true;
      dry.li = (size_t)(u->last_layout > u->last_images ? u->last_layout : u->last_images);
      backward_walk(dry, nullptr, nullptr, nullptr);
      VF_REQUIRE(dry.goff <= grad_workspace_bytes, "vf_unet_backward_phase: workspace too small (%zu < %zu)", grad_workspace_bytes, dry.goff);
    }
    vf::bw_state_free(u);
    BwdCtx* cx = new BwdCtx{u, as_stream(stream), reinterpret_cast<uint8_t*>(grad_workspace)};
    cx->gcap = grad_workspace_bytes;
    cx->li = (size_t)(u->last_layout > u->last_images ? u->last_layout : u->last_images);
    u->bw_state = cx;
    std::vector<int> phase_of;
    vf::plan_phases(u, n_phases, u->bw_phase_stop, phase_of);
    vf::bw_begin(*cx, grad_out8);
  }
  BwdCtx* cx = reinterpret_cast<BwdCtx*>(u->bw_state);
  if (!cx || (int)u->bw_phase_stop.size() != n_phases) { set_error("vf_unet_backward_phase: phases must be called in order 0..n_phases"); return VF_ERR_STATE; }
  cx->st = as_stream(stream);
  if (phase < n_phases) {
    vf::bw_run_ops(*cx, reinterpret_cast<const uint8_t*>(packed_t), param_grads_host, u->bw_phase_stop[phase]);
    vf::bw_flush_unpack(*cx, param_grads_host);
  } else {
    if (cx->next_op >= 0) { vf::bw_run_ops(*cx, reinterpret_cast<const uint8_t*>(packed_t), param_grads_host, 0); vf::bw_flush_unpack(*cx, param_grads_host); }
    vf::bw_finish(*cx, param_grads_host);
  }
  int rc = cx->rc;
  if (phase == n_phases || rc != VF_OK) vf::bw_state_free(u);
  if (rc == VF_OK) VF_LAUNCH_CHECK();
  return rc;
}
