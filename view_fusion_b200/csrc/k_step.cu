// One reverse-diffusion step as a single enqueue (reference: ViewFusion.p_sample / the loop body of generate,
// model/view_fusion.py:166-177, :196-206), with the loop state on the DEVICE so that a captured CUDA graph of the step can
// be replayed T times: nothing that changes from step to step (time-step, noise level, the `any(t > 0)` decision, the
// Philox offset) is a launch argument.
#include "vf_common.cuh"

namespace vf {

__global__ void __launch_bounds__(256) step_prepare_kernel(int* __restrict__ t_state, int B, const float* __restrict__ gammas, int T, int advance,
                                                           unsigned long long* noise_ctr, float* __restrict__ level, int* __restrict__ t_cur,
                                                           vf_step_record* __restrict__ rec) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ int any_pos;
  if (threadIdx.x == 0) any_pos = 0;
  __syncthreads();
  int mine = 0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    int t = t_state[b];
    t = t < 0 ? 0 : (t >= T ? T - 1 : t);
    level[b] = __ldg(gammas + t);
    t_cur[b] = t;
    mine |= t > 0;
    if (advance) t_state[b] = t > 0 ? t - 1 : 0;
  }
  if (mine) atomicOr(&any_pos, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    rec->any_t_positive = any_pos;
    rec->reserved = 0;
    unsigned long long off = 0, seed = 0;
    if (noise_ctr) {
      off = noise_ctr[0];
      seed = noise_ctr[1];
      if (advance) noise_ctr[0] = off + 1;
    }
    rec->noise_offset = off;
    rec->seed = seed;
  }
}

}  // namespace vf

using namespace vf;

extern "C" __attribute__((visibility("default"))) int vf_step_prepare(int* t_state, int B, const float* gammas, int num_timesteps, int advance,
                                                                    unsigned long long* noise_ctr, float* level, int* t_cur,
                                                                    vf_step_record* rec, vf_stream stream) {
  VF_REQUIRE(t_state && gammas && level && t_cur && rec && B > 0 && num_timesteps > 0, "vf_step_prepare: bad args");
  VF_CUDA(launch_pdl(step_prepare_kernel, dim3(1), dim3(256), 0, as_stream(stream), t_state, B, gammas, num_timesteps, advance, noise_ctr, level,
                     t_cur, rec));
  VF_LAUNCH_CHECK();
  return VF_OK;
}

extern "C" __attribute__((visibility("default"))) int vf_p_sample_step(vf_unet* u, const vf_sample_step_args* a, vf_stream stream) {
  VF_REQUIRE(u && a, "vf_p_sample_step: null args");
  VF_REQUIRE(a->y_cond && a->view_offset && a->angle && a->y_t && a->y_prev && a->t_state && a->x0 && a->img_sample && a->unet_out && a->level &&
                 a->t_cur && a->rec,
             "vf_p_sample_step: null tensor");
  VF_REQUIRE(a->z || a->noise_ctr || a->add_noise == 0, "vf_p_sample_step: noise needs z or a device counter");
  int rc = vf_step_prepare(a->t_state, a->B, a->sched.gammas, a->sched.num_timesteps, a->advance, a->noise_ctr, a->level, a->t_cur, a->rec, stream);
  if (rc != VF_OK) return rc;
  const int k0 = vf_unet_k0(u);
  rc = vf_pack_views(a->y_cond, a->y_t, a->view_offset, a->B, a->n_max, a->cond_channels, a->H, a->W, a->images, k0, vf_unet_act_dtype(u), a->x0,
                     a->img_sample, stream);
  if (rc != VF_OK) return rc;
  rc = vf_unet_forward(u, a->packed, a->workspace, a->workspace_bytes, a->images, a->x0, a->level, a->angle, a->B, a->img_sample, a->unet_out, stream);
  if (rc != VF_OK) return rc;
  vf_compose_args c{};
  c.unet_out = a->unet_out; c.view_offset = a->view_offset; c.t = a->t_cur; c.y_t = a->y_t; c.y_prev = a->y_prev; c.z = a->z;
  c.seed = a->seed; c.offset = 0; c.add_noise = a->add_noise; c.clip_denoised = a->clip_denoised; c.weighting = a->weighting;
  c.B = a->B; c.H = a->H; c.W = a->W;
  c.eps_out = a->eps_out; c.weights_out = a->weights_out; c.max_v = a->max_v; c.logits_out = a->logits_out;
  c.step = a->rec;
  return vf_compose_ddpm_step(&c, &a->sched, stream);
}
