// bptt_collect.cu -- the forward half of a BPTT window in ONE sm_100a kernel: T x [actor MLP on the tensor cores -> rsample ->
// differentiable env.step with its tape and losses].
//
// Reference path replaced (S = standalone): the rollout loop of AlgoRunner.learn (S/diff_rl/algorithms/runner.py:110-126):
// BaseModel.act (S/diff_rl/algorithms/model.py:63-99: ActorCritic.update_distribution + rsample; actor 16 -> 256 -> 128 -> 4,
// lrelu, QD/agents/diff_rl_naive_cfg.py:26-32) and env.step with the LossManager terms (QD/mdp/losses.py:72-117).  The env
// side is racing_step_body<diff> -- the tape planes, losses, rewards, resets of gr_step_fwd -- with the state in registers
// for the whole window; the reverse sweep stays gr_step_bwd.  What the policy's backward needs is recorded per step (the
// observation and the standard-normal draw behind each action), so that the host side replaces T small autograd graphs by
// ONE batched actor forward/backward over [T*N] rows (generalizableracing_b200/collect.py::FusedBpttCollector).
// Structure, operand layouts and numerics as in ppo_collect.cu / mlp_tc.cuh (one net, three MMA stages per step).
#include "racing_step_core.cuh"
#include "mlp_tc.cuh"

namespace gr {

// observations of step t feed step t+1: fp32 row -> obs_seq[t+1] (or the "next observation" buffer after the last step),
// fp16 row -> registers until the activation tile is free; the critic row is only needed after the last step
struct BpttObsSink {
  float4* obs_row; float4* critic_row; float* aux_ptr;
  uint4* policy_pk;
  __device__ __forceinline__ void policy(int, float4 o0, float4 o1, float4 o2, float4 o3) const {
    __stcs(obs_row + 0, o0); __stcs(obs_row + 1, o1); __stcs(obs_row + 2, o2); __stcs(obs_row + 3, o3);
    policy_pk[0] = pack8(o0, o1); policy_pk[1] = pack8(o2, o3);
  }
  __device__ __forceinline__ bool wants_critic() const { return critic_row != nullptr; }
  __device__ __forceinline__ void critic(int, float4 c0, float4 c1, float4 c2, float4 c3) const {
    __stcs(critic_row + 0, c0); __stcs(critic_row + 1, c1); __stcs(critic_row + 2, c2); __stcs(critic_row + 3, c3);
  }
  __device__ __forceinline__ void aux(int, float v) const { if (aux_ptr) *aux_ptr = v; }
  __device__ __forceinline__ constexpr bool wants_policy() const { return true; }
  template <bool kNoise, bool kDiff, bool kStats>          // (the state stays in registers over the rollout)
  __device__ __forceinline__ bool state_final(EnvRegs&, float4&, const float4&, const float (&)[GR_NUM_REWARD_TERMS], float, bool, bool) const { return false; }
};

template <class NL, int G, bool kNoise, bool kStats>
__global__ void __launch_bounds__(G * kTileEnvs, 1) bptt_collect_kernel(const GrConfig cfg, const GrTrack track, const GrState st, const GrRandom rng,
                                                                       const GrPolicy pol, const GrBpttCollectIO cio, const int track_in_smem,
                                                                       const int coop_off, const int coop_k) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* w_smem = smem;                                        // actor
  uint8_t* h_smem = smem + NL::kNetBytes;                        // G activation tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(h_smem + G * NL::kHBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + G);
  float4* track_rows = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(bars) + 128);

  const int tid = threadIdx.x, grp = tid / kTileEnvs, row = tid % kTileEnvs;
  const int i = blockIdx.x * (G * kTileEnvs) + tid;
  const bool active = i < st.num_envs;
  const int li = active ? i : st.num_envs - 1;
  const int N = st.num_envs, T = cio.T;

  {
    const uint4* src = reinterpret_cast<const uint4*>(pol.packed);
    uint4* dst = reinterpret_cast<uint4*>(w_smem);
    for (int k = tid; k < NL::kNetBytes / 16; k += G * kTileEnvs) dst[k] = __ldg(src + k);
  }
  if (tid < G) mbar_init(&bars[tid], 1);
  __syncwarp();
  if (tid < 32) tmem_alloc(tmem_slot, G * NL::kCols);
  TrackSmem tr{reinterpret_cast<const float4*>(track.rows), 0, track.levels, track.gates};
  if (track_in_smem) tr = stage_track(track, reinterpret_cast<const int2*>(st.chunk_types), N, track_rows);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();

  GroupCtx g = make_group_ctx(h_smem, NL::kHBytes, bars, *tmem_slot, NL::kCols, grp, row, pol.negative_slope);
  const uint32_t w_addr = smem_u32(w_smem);

  float4* __restrict__ tile = tile_ptr(reinterpret_cast<float4*>(st.planes), li);
  EnvRegs e;
  load_env<kNoise>(e, tile);
  float4 eps0 = make_float4(0.f, 0.f, 0.f, 0.f), lsum = eps0;
  if (kStats) { eps0 = ld_plane(tile, PL_EPSUM0); lsum = ld_plane(tile, PL_LOSSSUM); }
  const float4 sigma = *reinterpret_cast<const float4*>(pol.sigma);
  bool any_reset = false, any_noise_dirty = false, last_noise_dirty = false;
  const int64_t tape_step = (int64_t)(cio.tape_stride / kTile) * GR_TAPE_PLANES * kTile * 4;       // floats per tape step

  uint4 policy_pk[2];
  {
    const float4* o = reinterpret_cast<const float4*>(cio.obs0) + (int64_t)li * 4;
    const float4 o0 = __ldg(o), o1 = __ldg(o + 1), o2 = __ldg(o + 2), o3 = __ldg(o + 3);
    if (active) {
      float4* so = reinterpret_cast<float4*>(cio.obs_seq) + (int64_t)i * 4;
      __stcs(so, o0); __stcs(so + 1, o1); __stcs(so + 2, o2); __stcs(so + 3, o3);
    }
    policy_pk[0] = pack8(o0, o1); policy_pk[1] = pack8(o2, o3);
    write_x_row(g.hrow, policy_pk[0], policy_pk[1]);
  }

#pragma unroll 1
  for (int t = 0; t < T; ++t) {
    const int64_t tn = (int64_t)t * N + i;
    // ---- BaseModel.act: mean from the actor, reparameterised sample (model.py:96-99)
    stage_issue<NL>(g, w_addr, kL1);
    GrRandom rt = rng;
    rt.step = rng.step + (uint32_t)t;
    const RandSrc<true> rs(rt, li, st.env_id_offset + li);
    float4 n01, n23;
    rs.normals8(n01, n23);
    float4 eps;
    {
      const uint4 x = rs.ph(GR_PHILOX_CALL_ACTION);
      const float2 a0 = box_muller(x.x, x.y), a1 = box_muller(x.z, x.w);
      eps = make_float4(a0.x, a0.y, a1.x, a1.y);
    }
    if (active) __stcs(reinterpret_cast<float4*>(cio.eps_seq) + tn, eps);
    stage_wait(g);
    epilogue1<NL>(g);
    stage_issue<NL>(g, w_addr, kL2); stage_wait(g);
    epilogue2<NL>(g, w_smem);
    stage_issue<NL>(g, w_addr, kL3); stage_wait(g);
    const float4 mu = read_head<NL>(g, w_smem);
    tc_fence_before_sync();
    const float4 a_t = make_float4(mu.x + sigma.x * eps.x, mu.y + sigma.y * eps.y, mu.z + sigma.z * eps.z, mu.w + sigma.w * eps.w);
    if (active && cio.actions) __stcs(reinterpret_cast<float4*>(cio.actions) + tn, a_t);

    // ---- differentiable env.step: tape planes + losses of this step, observations for the next one
    GrStepIO io = {};
    io.log_accum = cio.log_accum;
    io.tape = cio.tape + (int64_t)t * tape_step;
    io.tape_stride = cio.tape_stride;
    io.loss = cio.loss + (int64_t)t * N;
    io.loss_terms = cio.loss_terms ? cio.loss_terms + (int64_t)t * N * 3 : nullptr;
    const CompactDraws draws{rs, coop_k > 0 ? track_rows + coop_off + (tid >> 5) * (7 * coop_k) : nullptr, coop_k};      // (ppo_collect.cu)
    const bool last = t == T - 1;
    BpttObsSink sink;
    sink.obs_row = (last ? reinterpret_cast<float4*>(cio.obs_out) : reinterpret_cast<float4*>(cio.obs_seq) + (int64_t)(t + 1) * N * 4) + (int64_t)i * 4;
    sink.critic_row = last ? reinterpret_cast<float4*>(cio.critic_obs_out) + (int64_t)i * 4 : nullptr;
    sink.aux_ptr = (last && cio.aux_out) ? cio.aux_out + i : nullptr;
    sink.policy_pk = policy_pk;
    StepOut so;
    const bool alive = racing_step_body<kNoise, true, true, kStats>(cfg, tr, e, a_t, n01, n23, draws, eps0, lsum, io, i, active, sink, so);
    write_x_row(g.hrow, policy_pk[0], policy_pk[1]);              // layer 3 has been read: the tile is free for the next step's operand
    if (alive) {
      if (kStats && !so.reset) add_episode_sums(eps0, e, so.terms, cfg.dt);
      any_reset |= so.reset;
      any_noise_dirty |= so.noise_dirty;
      last_noise_dirty = so.noise_dirty;
      if (cio.reward) cio.reward[tn] = so.reward;
      if (cio.dones) cio.dones[tn] = so.reset ? 1 : 0;
    }
  }

  if (active) {
    store_env<kNoise>(e, tile, any_reset, any_noise_dirty);
    if (kNoise && any_noise_dirty && !last_noise_dirty)      // the flag means "rewritten by the LAST step"
      st_plane(tile, PL_LINVEL, pack(e.v, __uint_as_float(eplen_word(e.eplen, e.aux != 0.0f, false, e.arate, e.metrics_zero))));
    if (kStats) { st_plane(tile, PL_EPSUM0, eps0); st_plane(tile, PL_LOSSSUM, lsum); }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (tid < 32) tmem_dealloc(*tmem_slot, G * NL::kCols);
}

}  // namespace gr

// =============================================================================================
// C ABI
// =============================================================================================
using namespace gr;

static inline bool misaligned(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) != 0; }

template <class NL, int G, bool kNoise, bool kStats>
static int launch_bptt(const GrConfig* cfg, const GrTrack* tr, const GrState* st, const GrRandom* rng, const GrPolicy* pol, const GrBpttCollectIO* io,
                       cudaStream_t s) {
  int types = ((G * kTileEnvs + 255) / 256) * st->max_types_per_block;
  if (types > tr->types) types = tr->types;
  size_t track_bytes = (size_t)types * tr->levels * (tr->gates + 1) * sizeof(float4);
  const size_t fixed = (size_t)NL::kNetBytes + (size_t)G * NL::kHBytes + 128;
  const int track_in_smem = fixed + track_bytes <= 227 * 1024;
  if (!track_in_smem) track_bytes = 0;
  if (fixed + track_bytes > 227 * 1024) return GR_ERR_SMEM;
  const size_t per_k = (size_t)(G * kTileEnvs / 32) * 7 * sizeof(float4);      // staging columns of the cooperative reset draws (gr_common.cuh)
  int coop_k = (int)((227 * 1024 - fixed - track_bytes) / per_k);
  coop_k = coop_k > 4 ? 4 : coop_k;
  const size_t bytes = fixed + track_bytes + (size_t)coop_k * per_k;
  auto kernel = bptt_collect_kernel<NL, G, kNoise, kStats>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return (int)e;
  const int grid = (st->num_envs + G * kTileEnvs - 1) / (G * kTileEnvs);
  kernel<<<grid, G * kTileEnvs, bytes, s>>>(*cfg, *tr, *st, *rng, *pol, *io, track_in_smem, (int)(track_bytes / sizeof(float4)), coop_k);
  return (int)cudaGetLastError();
}

template <class NL, int G>
static int dispatch_bptt(const GrConfig* cfg, const GrTrack* tr, const GrState* st, const GrRandom* rng, const GrPolicy* pol, const GrBpttCollectIO* io,
                         cudaStream_t s) {
  const bool stats = st->num_planes == GR_NUM_PLANES_WITH_STATS;
  if (cfg->add_cmd_noise) return stats ? launch_bptt<NL, G, true, true>(cfg, tr, st, rng, pol, io, s) : launch_bptt<NL, G, true, false>(cfg, tr, st, rng, pol, io, s);
  return stats ? launch_bptt<NL, G, false, true>(cfg, tr, st, rng, pol, io, s) : launch_bptt<NL, G, false, false>(cfg, tr, st, rng, pol, io, s);
}

extern "C" int gr_bptt_collect(const GrConfig* cfg, const GrTrack* track, const GrState* st, const GrRandom* rng, const GrPolicy* policy,
                               int32_t hidden, int32_t hidden2, const GrBpttCollectIO* io, void* stream) {
  if (!cfg || !track || !st || !rng || !policy || !io) return GR_ERR_NULL;
  if (!st->planes || !track->rows || !st->chunk_types || !policy->packed || !policy->sigma) return GR_ERR_NULL;
  if (!io->obs0 || !io->obs_out || !io->critic_obs_out || !io->obs_seq || !io->eps_seq || !io->loss || !io->tape) return GR_ERR_NULL;
  if (rng->rnd) return GR_ERR_CONFIG;
  if (st->num_envs <= 0 || io->T < 1 || io->tape_stride < ((st->num_envs + 31) & ~31) || (io->tape_stride & 31)) return GR_ERR_SIZE;
  if (st->num_planes != GR_NUM_PLANES && st->num_planes != GR_NUM_PLANES_WITH_STATS) return GR_ERR_SIZE;
  if (st->plane_stride < ((st->num_envs + 31) & ~31)) return GR_ERR_SIZE;
  if (track->types < 1 || track->types > 32 || track->levels < 1 || track->levels > 64 || track->gates < 1 || track->gates > GR_MAX_GATES) return GR_ERR_SIZE;
  if (st->max_types_per_block < 1 || st->max_types_per_block > track->types) return GR_ERR_SIZE;
  if (policy->negative_slope < 0.0f || policy->negative_slope > 1.0f) return GR_ERR_CONFIG;
  if (!((hidden == 128 || hidden == 256) && hidden2 == 128)) return GR_ERR_SIZE;
  if (misaligned(st->planes) || misaligned(track->rows) || misaligned(policy->packed) || misaligned(policy->sigma) || misaligned(io->obs0) ||
      misaligned(io->obs_out) || misaligned(io->critic_obs_out) || misaligned(io->obs_seq) || misaligned(io->eps_seq) || misaligned(io->tape) ||
      (io->actions && misaligned(io->actions)))
    return GR_ERR_ALIGN;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int G = io->groups_per_cta;
  const int tiles = (st->num_envs + kTileEnvs - 1) / kTileEnvs;
  if (G == 0) G = tiles <= 148 ? 1 : ((hidden == 256 || tiles <= 2 * 148) ? 2 : 4);       // fewest tiles per block that fit one wave (width 256: at most 2)
  if (hidden == 256) {
    if (G == 1) return dispatch_bptt<NetLayout<256, 128>, 1>(cfg, track, st, rng, policy, io, s);
    if (G == 2) return dispatch_bptt<NetLayout<256, 128>, 2>(cfg, track, st, rng, policy, io, s);
    return GR_ERR_SIZE;
  }
  if (G == 1) return dispatch_bptt<NetLayout<128, 128>, 1>(cfg, track, st, rng, policy, io, s);
  if (G == 2) return dispatch_bptt<NetLayout<128, 128>, 2>(cfg, track, st, rng, policy, io, s);
  if (G == 4) return dispatch_bptt<NetLayout<128, 128>, 4>(cfg, track, st, rng, policy, io, s);
  return GR_ERR_SIZE;
}
