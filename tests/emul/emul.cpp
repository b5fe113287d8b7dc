// TEST INFRASTRUCTURE -- see cuda_shim.h.  Exposes emul_* twins of the step/reset/startup/bwd/fill entry
// points of include/gracing.h that execute the kernel sources on the CPU (1 thread per block, sequential).
#define GR_CPU_EMUL 1
#include "cuda_shim.h"
#include "../../generalizableracing_b200/csrc/racing_step.cu"
#include "../../generalizableracing_b200/csrc/racing_bwd.cu"
#include "../../generalizableracing_b200/csrc/reach_step.cu"
#include "../../generalizableracing_b200/csrc/reach_bwd.cu"
// rollout storage: the block reductions of the GAE moments merge (count, mean, M2) triples pulled from other lanes; with one thread per block
// there is no other lane, i.e. an empty triple, which merge() skips.  (Valid for this file's merge-style reductions only.)
static inline double __shfl_down_sync(unsigned, double, int) { return 0.0; }
#include "../../generalizableracing_b200/csrc/rollout.cu"
#include "../../generalizableracing_b200/csrc/mesh_collision.cu"
#include <vector>

using namespace gr;

template <typename F>
static void run_grid(int64_t n, F&& f) {
  blockDim_.x = 1; gridDim_.x = (unsigned)n; threadIdx_.x = 0;
  for (int64_t b = 0; b < n; ++b) { blockIdx_.x = (unsigned)b; f(); }
}

extern "C" {

int emul_step_fwd(const GrConfig* cfg, const GrTrack* tr, const GrState* st, const GrRandom* rng, const GrStepIO* io) {
  const bool diff = io->loss || io->tape || io->loss_terms, stats = st->num_planes == GR_NUM_PLANES_WITH_STATS, philox = rng->rnd == nullptr,
             noise = cfg->add_cmd_noise != 0;
#define GO(a, b, c, d) if (noise == a && diff == b && philox == c && stats == d) { run_grid(st->num_envs, [&] { racing_step_fwd_kernel<a, b, c, d>(*cfg, *tr, *st, *rng, *io); }); return 0; }
  GO(false, false, false, false) GO(false, false, false, true) GO(false, false, true, false) GO(false, false, true, true)
  GO(false, true, false, false) GO(false, true, false, true) GO(false, true, true, false) GO(false, true, true, true)
  GO(true, false, false, false) GO(true, false, false, true) GO(true, false, true, false) GO(true, false, true, true)
  GO(true, true, false, false) GO(true, true, false, true) GO(true, true, true, false) GO(true, true, true, true)
#undef GO
  return -100;
}

int emul_rollout_fwd(const GrConfig* cfg, const GrTrack* tr, const GrState* st, const GrRandom* rng, const GrRolloutIO* io) {
  const bool diff = io->loss || io->tape || io->loss_terms, stats = st->num_planes == GR_NUM_PLANES_WITH_STATS, philox = rng->rnd == nullptr,
             noise = cfg->add_cmd_noise != 0;
#define GO(a, b, c, d) if (noise == a && diff == b && philox == c && stats == d) { \
    if (io->obs_seq) run_grid(st->num_envs, [&] { racing_rollout_fwd_kernel<a, b, c, d, true>(*cfg, *tr, *st, *rng, *io); }); \
    else run_grid(st->num_envs, [&] { racing_rollout_fwd_kernel<a, b, c, d, false>(*cfg, *tr, *st, *rng, *io); }); \
    return 0; }
  GO(false, false, false, false) GO(false, false, false, true) GO(false, false, true, false) GO(false, false, true, true)
  GO(false, true, false, false) GO(false, true, false, true) GO(false, true, true, false) GO(false, true, true, true)
  GO(true, false, false, false) GO(true, false, false, true) GO(true, false, true, false) GO(true, false, true, true)
  GO(true, true, false, false) GO(true, true, false, true) GO(true, true, true, false) GO(true, true, true, true)
#undef GO
  return -100;
}

int emul_env_reset(const GrConfig* cfg, const GrTrack* tr, const GrState* st, const GrRandom* rng, const uint8_t* mask, int mode, float* obs,
                   float* critic, float* aux) {
  const bool stats = st->num_planes == GR_NUM_PLANES_WITH_STATS, philox = rng->rnd == nullptr, noise = cfg->add_cmd_noise != 0;
#define GO(a, c, d) if (noise == a && philox == c && stats == d) { run_grid(st->num_envs, [&] { racing_reset_kernel<a, c, d>(*cfg, *tr, *st, *rng, mask, mode, obs, critic, aux); }); return 0; }
  GO(false, false, false) GO(false, false, true) GO(false, true, false) GO(false, true, true)
  GO(true, false, false) GO(true, false, true) GO(true, true, false) GO(true, true, true)
#undef GO
  return -100;
}

int emul_env_startup(const GrConfig* cfg, const GrTrack* tr, const GrState* st, const int32_t* types, int32_t* chunk_types, const float* srnd,
                     uint64_t seed) {
  run_grid(st->num_envs, [&] { racing_startup_kernel(*cfg, *tr, *st, types, chunk_types, srnd, seed); });
  return 0;
}

int emul_step_bwd(const GrConfig* cfg, const GrState* st, const GrBwdIO* io) {
  run_grid(st->num_envs, [&] { racing_step_bwd_kernel(*cfg, *st, *io); });
  return 0;
}

int emul_fill_rand(float* rnd, int32_t num_envs, int32_t env_id_offset, uint64_t seed, uint32_t step) {
  run_grid(num_envs * (GR_RND_STRIDE / 4), [&] { fill_rand_kernel(reinterpret_cast<float4*>(rnd), num_envs, env_id_offset, seed, step); });
  return 0;
}

int emul_fill_startup_rand(float* srnd, int32_t num_envs, int32_t env_id_offset, uint64_t seed) {
  run_grid(num_envs, [&] { fill_startup_rand_kernel(srnd, num_envs, env_id_offset, seed); });
  return 0;
}

int emul_reach_step_fwd(const GrReachConfig* cfg, const GrReachState* st, const GrRandom* rng, const GrReachStepIO* io) {
  const bool diff = io->loss || io->tape || io->loss_terms, philox = rng->rnd == nullptr;
#define GO(a, b) if (diff == a && philox == b) { run_grid(st->num_envs, [&] { reach_step_fwd_kernel<a, b>(*cfg, *st, *rng, *io); }); return 0; }
  GO(false, false) GO(false, true) GO(true, false) GO(true, true)
#undef GO
  return -100;
}

int emul_reach_rollout_fwd(const GrReachConfig* cfg, const GrReachState* st, const GrRandom* rng, const GrReachRolloutIO* io) {
  const bool diff = io->loss || io->tape || io->loss_terms, philox = rng->rnd == nullptr;
#define GO(a, b) if (diff == a && philox == b) { run_grid(st->num_envs, [&] { reach_rollout_fwd_kernel<a, b>(*cfg, *st, *rng, *io); }); return 0; }
  GO(false, false) GO(false, true) GO(true, false) GO(true, true)
#undef GO
  return -100;
}

int emul_reach_reset(const GrReachConfig* cfg, const GrReachState* st, const GrRandom* rng, const uint8_t* mask, int mode, float* obs) {
  const GrRandom none{nullptr, 0, 0};
  const GrRandom& r = rng ? *rng : none;
  if (r.rnd == nullptr) run_grid(st->num_envs, [&] { reach_reset_kernel<true>(*cfg, *st, r, mask, mode, obs); });
  else run_grid(st->num_envs, [&] { reach_reset_kernel<false>(*cfg, *st, r, mask, mode, obs); });
  return 0;
}

int emul_reach_step_bwd(const GrReachConfig* cfg, const GrReachState* st, const GrBwdIO* io) {
  run_grid(st->num_envs, [&] { reach_step_bwd_kernel(*cfg, *st, *io); });
  return 0;
}

int emul_reach_fill_rand(float* rnd, int32_t num_envs, int32_t env_id_offset, uint64_t seed, uint32_t step) {
  run_grid(num_envs, [&] { reach_fill_rand_kernel(rnd, num_envs, env_id_offset, seed, step); });
  return 0;
}

// ---- rollout storage (csrc/rollout.cu): the launch arithmetic of gr_storage_add / gr_compute_returns / gr_advantage_normalize / gr_storage_gather
static inline bool rows_vec4_(const GrStorage* s) { return !((s->obs_dim & 3) || (s->act_dim & 3) || (s->critic_obs && (s->critic_dim & 3))); }

int emul_storage_add(const GrStorage* s, const GrTransition* tr, int32_t step) {
  if (step < 0 || step >= s->T) return GR_ERR_SIZE;
  const int64_t N = s->N;
  const int V = rows_vec4_(s) ? 4 : 1;
  const int64_t total = N * s->obs_dim / V + (s->critic_obs ? N * s->critic_dim / V : 0) + 3 * (N * s->act_dim / V) + N;
  if (V == 4) run_grid(total, [&] { storage_add_kernel<4>(*s, *tr, step); });
  else run_grid(total, [&] { storage_add_kernel<1>(*s, *tr, step); });
  return 0;
}

int emul_advantage_normalize(const GrStorage* s, const double* moments) {
  const int64_t total = (int64_t)s->T * s->N;
  run_grid((total + 3) / 4, [&] { adv_normalize_kernel(s->advantages, total, moments); });
  return 0;
}

int emul_compute_returns(const GrStorage* s, const float* last_values, float gamma, float lam, void* scratch, double* moments, int32_t normalize) {
  // one partial per (one-thread) block: more partials than the scratch of gr_gae_scratch_bytes holds, so they live here
  std::vector<double> partials(3 * (size_t)s->N + 3);
  const int64_t blocks128 = ((int64_t)s->N + 127) / 128;
  double* mom = moments ? moments : reinterpret_cast<double*>(scratch) + 3 * blocks128;
  run_grid(s->N, [&] { gae_kernel(*s, last_values, gamma, lam, partials.data()); });
  run_grid(1, [&] { gae_moments_kernel(partials.data(), s->N, mom); });
  if (normalize) return emul_advantage_normalize(s, mom);
  return 0;
}

int emul_storage_gather(const GrStorage* s, const int64_t* indices, int32_t B, const GrMiniBatch* out) {
  const int V = rows_vec4_(s) ? 4 : 1;
  const int per_row = s->obs_dim / V + (s->critic_obs ? s->critic_dim / V : 0) + 3 * (s->act_dim / V) + 1;
  const int64_t total = (int64_t)B * per_row;
  if (V == 4) run_grid(total, [&] { storage_gather_kernel<4>(*s, indices, B, *out); });
  else run_grid(total, [&] { storage_gather_kernel<1>(*s, indices, B, *out); });
  return 0;
}

int emul_storage_pack_records(const GrStorage* s, const int64_t* perm, int64_t num, float* records) {
  const int64_t rows = perm ? num : (int64_t)s->T * s->N;
  if (rows < 1 || rows > (int64_t)s->T * s->N) return GR_ERR_SIZE;
  run_grid(rows * 12, [&] { storage_pack_records_kernel(*s, perm, reinterpret_cast<float4*>(records), rows); });
  return 0;
}

int emul_uav_collision_ray(const GrMesh* mesh, const float* pos, const float* quat, int32_t n, const float* lattices, int32_t num_lattices, float max_dist,
                           float arm_length, float height, int32_t* out) {
  std::memset(out, 0, sizeof(int32_t) * (size_t)n);
  run_grid((int64_t)n * (num_lattices > 0 ? num_lattices : 1), [&] { uav_collision_ray_kernel(*mesh, pos, quat, lattices, num_lattices, n, max_dist, arm_length, height, out); });
  return 0;
}

int emul_mesh_query_rays(const GrMesh* mesh, const float* origins, const float* dirs, int64_t num_rays, float max_t, float* t_out, float* sign_out) {
  run_grid(num_rays, [&] { mesh_query_rays_kernel(*mesh, origins, dirs, num_rays, max_t, t_out, sign_out); });
  return 0;
}
}
