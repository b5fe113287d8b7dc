// ppo_loss.cuh -- the per-row arithmetic of one PPO mini-batch loss (S/rsl_rl/ext/algorithms/ppo.py:118-171), shared by ppo_loss_grad_kernel
// (ppo_update.cu) and the fused forward + loss + weight-gradient kernel (actor_backward.cu).  Everything is returned UN-normalised (no 1/rows,
// no value_loss_coef): the callers apply their own factors, in the order the stand-alone kernel always used.
#pragma once
#include "../../include/gracing.h"

namespace gr {

struct PpoActorRow {
  float g_logp;            // d(clipped surrogate of this row) / d(log-prob)
  float d[4], inv_s[4];    // action - mean, 1 / std
  float surrogate, kl;
};
// log-prob of the stored action under the current policy (Normal.log_prob summed over the action dims), KL(old || new) (ppo.py:126-129),
// ratio and clipped surrogate (:143-149)
__device__ __forceinline__ PpoActorRow ppo_actor_row(const float4 mu, const float4 sg, const float4 a, const float4 omu, const float4 osg, const float adv,
                                                     const float old_logp, const float clip_param) {
  const float mus[4] = {mu.x, mu.y, mu.z, mu.w}, as[4] = {a.x, a.y, a.z, a.w}, sgs[4] = {sg.x, sg.y, sg.z, sg.w};
  const float omus[4] = {omu.x, omu.y, omu.z, omu.w}, osgs[4] = {osg.x, osg.y, osg.z, osg.w};
  PpoActorRow r;
  float logp = 0.0f, kl = 0.0f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float d = as[k] - mus[k];
    logp += -(d * d) / (2.0f * sgs[k] * sgs[k]) - logf(sgs[k]) - 0.91893853320467274178f;
    const float dm = omus[k] - mus[k];
    kl += logf(sgs[k] / osgs[k] + 1.0e-5f) + (osgs[k] * osgs[k] + dm * dm) / (2.0f * sgs[k] * sgs[k]) - 0.5f;
    r.d[k] = d;
    r.inv_s[k] = 1.0f / sgs[k];
  }
  const float ratio = expf(logp - old_logp);
  const float s1 = -adv * ratio;
  const float s2 = -adv * fminf(fmaxf(ratio, 1.0f - clip_param), 1.0f + clip_param);
  r.surrogate = fmaxf(s1, s2);
  // d max(s1, s2) / d logp: s1 carries -adv * ratio; s2 carries it only inside the clip range (where s1 == s2: the tie's two halves add up)
  r.g_logp = s1 >= s2 ? -adv * ratio : 0.0f;
  r.kl = kl;
  return r;
}

struct PpoCriticRow { float vloss, g_v; };
// value loss (ppo.py:153-160) and its derivative w.r.t. the value
__device__ __forceinline__ PpoCriticRow ppo_critic_row(const float v, const float ret, const float ov, const bool clipped, const float clip_param) {
  PpoCriticRow r;
  if (clipped) {
    const float dv = v - ov;
    const bool inside = dv >= -clip_param && dv <= clip_param;     // torch.clamp passes the gradient at the bounds
    const float vc = ov + fminf(fmaxf(dv, -clip_param), clip_param);
    const float l1 = (v - ret) * (v - ret), l2 = (vc - ret) * (vc - ret);
    r.vloss = fmaxf(l1, l2);
    r.g_v = l1 >= l2 ? 2.0f * (v - ret) : (inside ? 2.0f * (vc - ret) : 0.0f);
    if (l1 == l2 && !inside) r.g_v *= 0.5f;                        // exact tie outside the range: only the l1 half has a gradient
  } else {
    r.vloss = (ret - v) * (ret - v);
    r.g_v = 2.0f * (v - ret);
  }
  return r;
}

}  // namespace gr
