// C-ABI plumbing: error string, device check, and the dtype dispatch of the stage-level operators.
#include <cstring>

#include "vf_common.cuh"

namespace vf {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  int& n = cached[dev & 63];
  if (n == 0 && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
  return n;
}

int conv2d_simt(const vf_conv_args* a, cudaStream_t st);
int conv2d_tc(const vf_conv_args* a, cudaStream_t st);
int attention_simt(const void* qk, const void* vt, int dtype, int images, int L, int C, void* out, cudaStream_t st);
int attention_tc(const void* qk, const void* vt, int images, int L, int C, void* out, float* lse, cudaStream_t st);
void set_tc_debug(int f);
void set_tc_debug_out(long long* p);

}  // namespace vf

extern "C" __attribute__((visibility("default"))) const char* vf_last_error(void) { return vf::g_err; }
extern "C" __attribute__((visibility("default"))) int vf_abi_version(void) { return VF_ABI_VERSION; }

extern "C" __attribute__((visibility("default"))) int vf_device_check(void) {
  using namespace vf;
  int dev = 0, major = 0, minor = 0;
  VF_CUDA(cudaGetDevice(&dev));
  VF_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  VF_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) {
    set_error("viewfusion_b200 needs an sm_100 device (B200); found sm_%d%d", major, minor);
    return VF_ERR_ARCH;
  }
  return VF_OK;
}

// force_simt: debugging / cross-check switch (environment VF_FORCE_SIMT=1 is read by the Python tests only)
namespace vf {
static bool g_pdl_suspended = false;
void pdl_set_suspended(bool s) { g_pdl_suspended = s; }     // per-kernel event timing needs serialised launches
bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("VF_PDL"); return !(e && e[0] == '0'); }();
  return on && !g_pdl_suspended;
}
}  // namespace vf
static int g_force_simt = 0;
namespace vf { int g_force_simt_flag = 0; }
extern "C" __attribute__((visibility("default"))) void vf_debug_force_simt(int on) { g_force_simt = on; vf::g_force_simt_flag = on; }
extern "C" __attribute__((visibility("default"))) void vf_debug_flags(int flags) { vf::set_tc_debug(flags); }
extern "C" __attribute__((visibility("default"))) void vf_debug_counters(long long* dev_buf) { vf::set_tc_debug_out(dev_buf); }

extern "C" __attribute__((visibility("default"))) int vf_conv2d(const vf_conv_args* a, vf_stream stream) {
  using namespace vf;
  VF_REQUIRE(a, "vf_conv2d: null args");
  VF_REQUIRE(a->n_seg >= 1 && a->n_seg <= 3, "vf_conv2d: n_seg=%d", a->n_seg);
  VF_REQUIRE(a->images > 0 && a->H > 0 && a->W > 0 && a->cout > 0 && a->cout_pad >= a->cout, "vf_conv2d: bad shape");
  VF_REQUIRE(a->weight && a->out, "vf_conv2d: null tensor");
  VF_REQUIRE(a->stride == 1 || a->stride == 2, "vf_conv2d: stride=%d", a->stride);
  for (int s = 0; s < a->n_seg; ++s) {
    VF_REQUIRE(a->src[s] && a->src_c[s] > 0, "vf_conv2d: segment %d empty", s);
    VF_REQUIRE(a->ksize[s] == 1 || a->ksize[s] == 3, "vf_conv2d: ksize=%d", a->ksize[s]);
    VF_REQUIRE(s == 0 || a->ksize[s] == 1, "vf_conv2d: only segment 0 may be 3x3");
  }
  if (a->dtype == VF_BF16 && !g_force_simt) return conv2d_tc(a, as_stream(stream));
  return conv2d_simt(a, as_stream(stream));
}

extern "C" __attribute__((visibility("default"))) int vf_attention(const void* qk, const void* vt, int dtype, int images, int L, int C, void* out, float* lse, vf_stream stream) {
  using namespace vf;
  VF_REQUIRE(qk && out && images > 0 && L > 0 && C > 0, "vf_attention: bad args");
  if (dtype == VF_BF16 && !g_force_simt) return attention_tc(qk, vt, images, L, C, out, lse, as_stream(stream));
  return attention_simt(qk, vt, dtype, images, L, C, out, as_stream(stream));
}
