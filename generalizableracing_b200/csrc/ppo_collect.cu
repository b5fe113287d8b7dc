// ppo_collect.cu -- the whole PPO collection phase in ONE sm_100a kernel: T x [actor MLP -> Normal sample -> log-prob ->
// critic MLP -> env.step -> time-out bootstrap -> add_transitions], then V(last obs) for the GAE bootstrap.
//
// Reference path replaced (S = standalone): the rollout loop of OnPolicyRunner.learn (S/rsl_rl/ext/runners/
// on_policy_runner.py:141-175): PPO.act (S/rsl_rl/ext/algorithms/ppo.py:71-83; rsl_rl ActorCritic 16->128->128->4 /
// ->1, lrelu, QD/agents/rsl_rl_ppo_cfg.py:22-27), env.step (the ~700 aten launches racing_step.cu fuses),
// PPO.process_env_step (ppo.py:85-97), RolloutStorage.add_transitions (S/rsl_rl/ext/storage/rollout_storage.py:71-88),
// the episode book keeping (on_policy_runner.py:160-173) and PPO.compute_returns' last_values (ppo.py:99-100).
// Measured on a B200 (profiles/r1_rollout_cost_torch_policy_vs_env.json): per collection step at 65,536 envs the torch
// policy costs 432 us eager / 296 us under a CUDA graph against 9 us for the env step -- the MLP is the dense
// contraction that bounds collection, so it goes to the tensor cores here (north star: tensor cores only for the MLP).
//
// Design
//  * one thread per env, 128 envs (4 warps) = one UMMA tile (M = 128); a CTA runs G tiles ("groups") that share one
//    copy of the packed weights in shared memory and overlap each other's MMA latency; env state lives in REGISTERS for
//    the whole rollout (loaded once, stored once): per env-step HBM sees only the rollout-storage rows.
//  * per group and net: X[128x32] . W1 -> D (TMEM) -> lrelu -> H1[128x128] fp16 (smem) . W2 -> D -> +b2, lrelu -> H2 . W3
//    (N = 16) -> D -> 4 means / 1 value.  tcgen05.mma kind::f16 (fp16 operands, fp32 accumulate), issued by one thread per
//    group, completion through an mbarrier; accumulator row = TMEM lane = the env's own thread, so the epilogue is
//    thread-local (tcgen05.ld 32x32b) and each thread writes its own activation row back as the next A operand.
//    Layer-1 bias rides in the GEMM (two constant-one K columns against fp16 hi/lo bias rows), layer-2 bias is a packed
//    half2 add in the epilogue, layer-3 bias is fp32.
//  * operands are K-major, non-swizzled: [K/8 chunks][rows][8 halfs]; a thread's 16-byte chunk stores are conflict-free.
//  * numerics: policy inference at fp16-operand / fp32-accumulate precision (the error against the fp32 torch MLP is
//    measured in tests/test_ppo_collect.py and DESIGN.md); everything the env does (state, rewards, dones, observations)
//    is the same code as gr_step_fwd: bit for bit given the same actions in the default variant <noise, stats>; within ~1 ulp
//    per step in the others (FMA contraction of the inlined body differs between the two kernels).
#include "racing_step_core.cuh"
#include "mlp_tc.cuh"

namespace gr {

using NL = NetLayout<128, 128>;          // QD/agents/rsl_rl_ppo_cfg.py:22-27: actor / critic hidden dims [128, 128]
constexpr int kHid = 128;
constexpr int kNetBytes = NL::kNetBytes, kHBytes = NL::kHBytes, kB2Off = NL::kB2Off;

// observation sink of the fused kernel: fp32 rows go to the rollout storage (or to the "next observation" buffers after
// the last step); both rows also wait in 8 registers each, as fp16, to become layer-1 operands once the group's
// activation tile is free (the env step runs while the critic's layer-2 MMAs still read it)
struct FusedObsSink {
  float4* obs_row; float4* critic_row; float* aux_ptr;
  uint4* policy_pk; uint4* critic_pk;     // -> two uint4 each in the caller's registers
  __device__ __forceinline__ void policy(int, float4 o0, float4 o1, float4 o2, float4 o3) const {
    __stcs(obs_row + 0, o0); __stcs(obs_row + 1, o1); __stcs(obs_row + 2, o2); __stcs(obs_row + 3, o3);
    policy_pk[0] = pack8(o0, o1); policy_pk[1] = pack8(o2, o3);
  }
  __device__ __forceinline__ bool wants_critic() const { return true; }
  __device__ __forceinline__ void critic(int, float4 c0, float4 c1, float4 c2, float4 c3) const {
    __stcs(critic_row + 0, c0); __stcs(critic_row + 1, c1); __stcs(critic_row + 2, c2); __stcs(critic_row + 3, c3);
    critic_pk[0] = pack8(c0, c1); critic_pk[1] = pack8(c2, c3);
  }
  __device__ __forceinline__ void aux(int, float v) const { if (aux_ptr) *aux_ptr = v; }
  __device__ __forceinline__ constexpr bool wants_policy() const { return true; }
  template <bool kNoise, bool kDiff, bool kStats>          // (the state stays in registers over the rollout)
  __device__ __forceinline__ bool state_final(EnvRegs&, float4&, const float4&, const float (&)[GR_NUM_REWARD_TERMS], float, bool, bool) const { return false; }
};

// ---------------------------------------------------------------------------------------------
// the collection kernel
// ---------------------------------------------------------------------------------------------
template <int G, bool kNoise, bool kStats>
__global__ void __launch_bounds__(G * kTileEnvs, 1) ppo_collect_kernel(const GrConfig cfg, const GrTrack track, const GrState st, const GrRandom rng,
                                                                      const GrPolicy pol, const GrStorage sto, const GrCollectIO cio,
                                                                      const int track_in_smem, const int coop_off, const int coop_k) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* w_smem = smem;                                        // actor net | critic net
  uint8_t* h_smem = smem + 2 * kNetBytes;                        // G activation tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(h_smem + G * kHBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + G);
  float4* track_rows = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(bars) + 128);

  constexpr bool kWarpIssue = G < 4;        // (mlp_tc.cuh issue_layer: the converged issue form needs registers the 128-register build does not have)
  const int tid = threadIdx.x, grp = tid / kTileEnvs, row = tid % kTileEnvs;
  const int i = blockIdx.x * (G * kTileEnvs) + tid;
  const bool active = i < st.num_envs;
  const int li = active ? i : st.num_envs - 1;
  const int N = st.num_envs, T = sto.T;

  // ---- one-time setup: weights -> smem, gate-table slice, barriers, tensor memory
  {
    const uint4* src = reinterpret_cast<const uint4*>(pol.packed);
    uint4* dst = reinterpret_cast<uint4*>(w_smem);
    for (int k = tid; k < 2 * kNetBytes / 16; k += G * kTileEnvs) dst[k] = __ldg(src + k);
  }
  if (tid < G) mbar_init(&bars[tid], 1);
  __syncwarp();
  if (tid < 32) tmem_alloc(tmem_slot, G * kHid < 32 ? 32 : G * kHid);
  // gate table: the slice of the terrain types this CTA spans goes to shared memory when it fits next to the weights and
  // activation tiles, otherwise the (L1-resident, <= 29 KB) table is read in place
  TrackSmem tr{reinterpret_cast<const float4*>(track.rows), 0, track.levels, track.gates};
  if (track_in_smem) tr = stage_track(track, reinterpret_cast<const int2*>(st.chunk_types), N, track_rows);     // ends with __syncthreads()
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();

  GroupCtx g = make_group_ctx(h_smem, kHBytes, bars, *tmem_slot, kHid, grp, row, pol.negative_slope);
  const uint32_t w_addr = smem_u32(w_smem);

  // ---- env state -> registers (once per rollout)
  float4* __restrict__ tile = tile_ptr(reinterpret_cast<float4*>(st.planes), li);
  EnvRegs e;
  load_env<kNoise>(e, tile);
  float4 eps0 = make_float4(0.f, 0.f, 0.f, 0.f), lsum = eps0;
  if (kStats) { eps0 = ld_plane(tile, PL_EPSUM0); }
  float2 epacc = cio.episode_acc ? reinterpret_cast<const float2*>(cio.episode_acc)[li] : make_float2(0.f, 0.f);
  const float4 sigma = *reinterpret_cast<const float4*>(pol.sigma);
  const float4 log_sigma = make_float4(logf(sigma.x), logf(sigma.y), logf(sigma.z), logf(sigma.w));
  bool any_reset = false, any_noise_dirty = false, last_noise_dirty = false;
  GrStepIO io = {};
  io.log_accum = cio.log_accum;

  // ---- observations the rollout starts from -> storage slot 0 + operand rows
  uint4 critic_pk[2];
  {
    const float4* o = reinterpret_cast<const float4*>(cio.obs0) + (int64_t)li * 4;
    const float4* c = reinterpret_cast<const float4*>(cio.critic_obs0) + (int64_t)li * 4;
    const float4 o0 = __ldg(o), o1 = __ldg(o + 1), o2 = __ldg(o + 2), o3 = __ldg(o + 3);
    const float4 c0 = __ldg(c), c1 = __ldg(c + 1), c2 = __ldg(c + 2), c3 = __ldg(c + 3);
    if (active) {
      float4* so = reinterpret_cast<float4*>(sto.obs) + (int64_t)i * 4;
      float4* sc = reinterpret_cast<float4*>(sto.critic_obs) + (int64_t)i * 4;
      __stcs(so, o0); __stcs(so + 1, o1); __stcs(so + 2, o2); __stcs(so + 3, o3);
      __stcs(sc, c0); __stcs(sc + 1, c1); __stcs(sc + 2, c2); __stcs(sc + 3, c3);
    }
    write_x_row(g.hrow, pack8(o0, o1), pack8(o2, o3));
    critic_pk[0] = pack8(c0, c1); critic_pk[1] = pack8(c2, c3);
  }

  const uint8_t* critic_smem = w_smem + kNetBytes;
  const uint32_t critic_addr = w_addr + kNetBytes;
  uint4 policy_pk[2] = {critic_pk[0], critic_pk[1]};       // (lanes past the last env never refresh it)

  // Per step: actor L1 L2 L3 -> sample -> critic L1 L2 L3, env.step.  Work that does not depend on the MMA in flight is
  // placed between its issue and its wait: the Philox draws under actor L1, the storage rows of the action under critic L1,
  // the whole env step under critic L2 (the longest MMA), the episode sums under critic L3.
#pragma unroll 1
  for (int t = 0; t < T; ++t) {
    const int64_t tn = (int64_t)t * N + i;
    // ---- PPO.act: actor mean, sample, log-prob (ppo.py:71-83; Normal(mean, std).sample() / .log_prob().sum(-1))
    stage_issue<NL, kWarpIssue>(g, w_addr, kL1);
    GrRandom rt = rng;
    rt.step = rng.step + (uint32_t)t;
    const RandSrc<true> rs(rt, li, st.env_id_offset + li);
    float4 n01, n23;
    rs.normals8(n01, n23);
    float2 an0, an1;
    {
      const uint4 x = rs.ph(GR_PHILOX_CALL_ACTION);
      an0 = box_muller(x.x, x.y); an1 = box_muller(x.z, x.w);
    }
    stage_wait(g);
    epilogue1<NL>(g);
    stage_issue<NL, kWarpIssue>(g, w_addr, kL2); stage_wait(g);
    epilogue2<NL>(g, w_smem);
    stage_issue<NL, kWarpIssue>(g, w_addr, kL3); stage_wait(g);
    const float4 mu = read_head<NL>(g, w_smem);
    const float4 a_t = make_float4(mu.x + sigma.x * an0.x, mu.y + sigma.y * an0.y, mu.z + sigma.z * an1.x, mu.w + sigma.w * an1.y);
    // ---- critic value of the same state
    write_x_row(g.hrow, critic_pk[0], critic_pk[1]);
    stage_issue<NL, kWarpIssue>(g, critic_addr, kL1);
    {
      const float dx = a_t.x - mu.x, dy = a_t.y - mu.y, dz = a_t.z - mu.z, dw = a_t.w - mu.w;
      const float kLogSqrt2Pi = 0.91893853320467274178f;
      const float logp = (-(dx * dx) / (2.0f * sigma.x * sigma.x) - log_sigma.x - kLogSqrt2Pi) + (-(dy * dy) / (2.0f * sigma.y * sigma.y) - log_sigma.y - kLogSqrt2Pi) +
                         (-(dz * dz) / (2.0f * sigma.z * sigma.z) - log_sigma.z - kLogSqrt2Pi) + (-(dw * dw) / (2.0f * sigma.w * sigma.w) - log_sigma.w - kLogSqrt2Pi);
      if (active) {
        __stcs(reinterpret_cast<float4*>(sto.actions) + tn, a_t);
        __stcs(reinterpret_cast<float4*>(sto.mu) + tn, mu);
        __stcs(reinterpret_cast<float4*>(sto.sigma) + tn, sigma);
        sto.log_prob[tn] = logp;
      }
    }
    stage_wait(g);
    epilogue1<NL>(g);
    stage_issue<NL, kWarpIssue>(g, critic_addr, kL2);

    // ---- env.step (same body as gr_step_fwd) while the critic's layer 2 runs; its observations are the next step's
    //      operands (kept packed in registers until the activation tile is free) and storage rows
    // reset draws: the warp generates the seven Philox calls of its first coop_k resetting lanes together (CompactDraws, gr_common.cuh) --
    // left to the resetting lane alone they are ~770 issue slots of the warp in a kernel that is bound by issue slots
    const CompactDraws draws{rs, coop_k > 0 ? track_rows + coop_off + (tid >> 5) * (7 * coop_k) : nullptr, coop_k};
    const bool last = t == T - 1;
    FusedObsSink sink;
    sink.obs_row = (last ? reinterpret_cast<float4*>(cio.obs_out) : reinterpret_cast<float4*>(sto.obs) + (int64_t)(t + 1) * N * 4) + (int64_t)i * 4;
    sink.critic_row = (last ? reinterpret_cast<float4*>(cio.critic_obs_out) : reinterpret_cast<float4*>(sto.critic_obs) + (int64_t)(t + 1) * N * 4) + (int64_t)i * 4;
    sink.aux_ptr = (last && cio.aux_out) ? cio.aux_out + i : nullptr;
    sink.policy_pk = policy_pk;
    sink.critic_pk = critic_pk;
    StepOut so;
    const bool alive = racing_step_body<kNoise, false, true, kStats>(cfg, tr, e, a_t, n01, n23, draws, eps0, lsum, io, i, active, sink, so);

    stage_wait(g);
    epilogue2<NL>(g, critic_smem);
    stage_issue<NL, kWarpIssue>(g, critic_addr, kL3);
    if (alive) {
      if (kStats && !so.reset) add_episode_sums(eps0, e, so.terms, cfg.dt);
      any_reset |= so.reset;
      any_noise_dirty |= so.noise_dirty;
      last_noise_dirty = so.noise_dirty;
      sto.dones[tn] = so.reset ? 1 : 0;
    }
    stage_wait(g);
    const float value = read_head<NL>(g, critic_smem).x;
    write_x_row(g.hrow, policy_pk[0], policy_pk[1]);          // the tile is free again: next step's actor operand
    if (alive) {
      // ---- PPO.process_env_step (ppo.py:85-97) + add_transitions: bootstrap on time-outs with V(s_t)
      sto.rewards[tn] = so.reward + cio.gamma * (value * (so.time_out ? 1.0f : 0.0f));
      sto.values[tn] = value;
      // ---- episode book keeping of the runner (on_policy_runner.py:160-173)
      epacc.x += so.reward;
      epacc.y += 1.0f;
      if (so.reset) {
        if (cio.episode_log) {
          float* acc_row = cio.episode_log + (size_t)((i >> 5) & (GR_LOG_SHARDS - 1)) * 4;
          atomicAdd(acc_row + 0, epacc.x);
          atomicAdd(acc_row + 1, epacc.y);
          atomicAdd(acc_row + 2, 1.0f);
        }
        epacc = make_float2(0.f, 0.f);
      }
    }
  }

  // ---- V(observation after the last step) for the GAE bootstrap (ppo.py:99-100)
  write_x_row(g.hrow, critic_pk[0], critic_pk[1]);
  const float last_value = run_net<NL, kWarpIssue>(g, critic_smem, critic_addr).x;
  if (active) {
    cio.last_values[i] = last_value;
    // ---- env state -> HBM (once per rollout).  Every env that reset at ANY step rewrote its read-mostly planes: the
    // caller clears GR_LAUNCH_PREFETCH for the next single-step launch (env.py).
    store_env<kNoise>(e, tile, any_reset, any_noise_dirty);
    if (kNoise && any_noise_dirty && !last_noise_dirty)      // the flag means "rewritten by the LAST step"
      st_plane(tile, PL_LINVEL, pack(e.v, __uint_as_float(eplen_word(e.eplen, e.aux != 0.0f, false, e.arate, e.metrics_zero))));
    if (kStats) { st_plane(tile, PL_EPSUM0, eps0); }
    if (cio.episode_acc) reinterpret_cast<float2*>(cio.episode_acc)[i] = epacc;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (tid < 32) tmem_dealloc(*tmem_slot, G * kHid < 32 ? 32 : G * kHid);
}

}  // namespace gr

// =============================================================================================
// C ABI
// =============================================================================================
using namespace gr;

static inline bool bad16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) != 0; }

extern "C" int64_t gr_policy_packed_bytes(int32_t hidden, int32_t hidden2, int32_t nets) {
  if (nets < 1 || nets > 2) return GR_ERR_SIZE;
  if (hidden == 128 && hidden2 == 128) return (int64_t)nets * NetLayout<128, 128>::kNetBytes;
  if (hidden == 256 && hidden2 == 128) return (int64_t)nets * NetLayout<256, 128>::kNetBytes;
  return GR_ERR_SIZE;
}

extern "C" int gr_policy_pack(const GrMlp* actor, const GrMlp* critic, void* packed, void* stream) {
  if (!actor || !packed) return GR_ERR_NULL;
  const int nets = critic ? 2 : 1;
  const GrMlp* ms[2] = {actor, critic};
  for (int k = 0; k < nets; ++k) {
    const GrMlp* m = ms[k];
    if (!m->w1 || !m->b1 || !m->w2 || !m->b2 || !m->w3 || !m->b3) return GR_ERR_NULL;
    if (m->in_dim != kObsDim || m->hidden != actor->hidden || m->hidden2 != actor->hidden2 || m->out_dim < 1 || m->out_dim > 4) return GR_ERR_SIZE;
  }
  if (gr_policy_packed_bytes(actor->hidden, actor->hidden2, nets) < 0) return GR_ERR_SIZE;
  if (bad16(packed)) return GR_ERR_ALIGN;
  const GrMlp second = critic ? *critic : *actor;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (actor->hidden == 128) policy_pack_kernel<NetLayout<128, 128>><<<dim3((128 * 128 + 255) / 256, nets), 256, 0, s>>>(*actor, second, static_cast<uint8_t*>(packed));
  else policy_pack_kernel<NetLayout<256, 128>><<<dim3((256 * 128 + 255) / 256, nets), 256, 0, s>>>(*actor, second, static_cast<uint8_t*>(packed));
  return (int)cudaGetLastError();
}

template <int G, bool kNoise, bool kStats>
static int launch_collect(const GrConfig* cfg, const GrTrack* tr, const GrState* st, const GrRandom* rng, const GrPolicy* pol, const GrStorage* sto,
                          const GrCollectIO* io, cudaStream_t s) {
  int types = ((G * kTileEnvs + 255) / 256) * st->max_types_per_block;
  if (types > tr->types) types = tr->types;
  size_t track_bytes = (size_t)types * tr->levels * (tr->gates + 1) * sizeof(float4);
  const size_t fixed = 2 * (size_t)kNetBytes + (size_t)G * kHBytes + 128;
  const int track_in_smem = fixed + track_bytes <= 227 * 1024;
  if (!track_in_smem) track_bytes = 0;
  // what is left goes to the warps' staging columns of the cooperative reset draws: 7 calls x K columns x 16 B per warp, K <= 4
  const size_t per_k = (size_t)(G * kTileEnvs / 32) * 7 * sizeof(float4);
  int coop_k = (int)((227 * 1024 - fixed - track_bytes) / per_k);
  coop_k = coop_k > 4 ? 4 : coop_k;
  if (io->coop_reset_columns < 0) coop_k = 0;
  else if (io->coop_reset_columns > 0 && io->coop_reset_columns < coop_k) coop_k = io->coop_reset_columns;
  const size_t bytes = fixed + track_bytes + (size_t)coop_k * per_k;
  auto kernel = ppo_collect_kernel<G, kNoise, kStats>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return (int)e;
  const int grid = (st->num_envs + G * kTileEnvs - 1) / (G * kTileEnvs);
  kernel<<<grid, G * kTileEnvs, bytes, s>>>(*cfg, *tr, *st, *rng, *pol, *sto, *io, track_in_smem, (int)(track_bytes / sizeof(float4)), coop_k);
  return (int)cudaGetLastError();
}

template <int G>
static int dispatch_collect(const GrConfig* cfg, const GrTrack* tr, const GrState* st, const GrRandom* rng, const GrPolicy* pol, const GrStorage* sto,
                            const GrCollectIO* io, cudaStream_t s) {
  const bool stats = st->num_planes == GR_NUM_PLANES_WITH_STATS;
  if (cfg->add_cmd_noise) return stats ? launch_collect<G, true, true>(cfg, tr, st, rng, pol, sto, io, s) : launch_collect<G, true, false>(cfg, tr, st, rng, pol, sto, io, s);
  return stats ? launch_collect<G, false, true>(cfg, tr, st, rng, pol, sto, io, s) : launch_collect<G, false, false>(cfg, tr, st, rng, pol, sto, io, s);
}

extern "C" int gr_ppo_collect(const GrConfig* cfg, const GrTrack* track, const GrState* st, const GrRandom* rng, const GrPolicy* policy,
                              const GrStorage* storage, const GrCollectIO* io, void* stream) {
  if (!cfg || !track || !st || !rng || !policy || !storage || !io) return GR_ERR_NULL;
  if (!st->planes || !track->rows || !st->chunk_types || !policy->packed || !policy->sigma) return GR_ERR_NULL;
  if (!io->obs0 || !io->critic_obs0 || !io->obs_out || !io->critic_obs_out || !io->last_values) return GR_ERR_NULL;
  if (!storage->obs || !storage->critic_obs || !storage->actions || !storage->rewards || !storage->dones || !storage->values || !storage->log_prob ||
      !storage->mu || !storage->sigma)
    return GR_ERR_NULL;
  if (rng->rnd) return GR_ERR_CONFIG;                     // the fused path draws in-kernel (Philox) only
  if (st->num_envs <= 0 || storage->N != st->num_envs || storage->T < 1) return GR_ERR_SIZE;
  if (storage->obs_dim != kObsDim || storage->critic_dim != kObsDim || storage->act_dim != GR_NUM_ACTIONS) return GR_ERR_SIZE;
  if (st->num_planes != GR_NUM_PLANES && st->num_planes != GR_NUM_PLANES_WITH_STATS) return GR_ERR_SIZE;
  if (st->plane_stride < ((st->num_envs + 31) & ~31)) return GR_ERR_SIZE;
  if (track->types < 1 || track->types > 32 || track->levels < 1 || track->levels > 64 || track->gates < 1 || track->gates > GR_MAX_GATES) return GR_ERR_SIZE;
  if (st->max_types_per_block < 1 || st->max_types_per_block > track->types) return GR_ERR_SIZE;
  if (policy->negative_slope < 0.0f || policy->negative_slope > 1.0f) return GR_ERR_CONFIG;
  if (bad16(st->planes) || bad16(track->rows) || bad16(policy->packed) || bad16(policy->sigma) || bad16(io->obs0) || bad16(io->critic_obs0) ||
      bad16(io->obs_out) || bad16(io->critic_obs_out) || bad16(storage->obs) || bad16(storage->critic_obs) || bad16(storage->actions) ||
      bad16(storage->mu) || bad16(storage->sigma) || (io->episode_acc && (reinterpret_cast<uintptr_t>(io->episode_acc) & 7u)))
    return GR_ERR_ALIGN;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int G = io->groups_per_cta;
  if (G == 0) {            // smallest CTA that still fits the rollout in one wave of 148 SMs
    const int tiles = (st->num_envs + kTileEnvs - 1) / kTileEnvs;
    G = tiles <= 148 ? 1 : (tiles <= 2 * 148 ? 2 : 4);
  }
  switch (G) {
    case 1: return dispatch_collect<1>(cfg, track, st, rng, policy, storage, io, s);
    case 2: return dispatch_collect<2>(cfg, track, st, rng, policy, storage, io, s);
    case 4: return dispatch_collect<4>(cfg, track, st, rng, policy, storage, io, s);
    default: return GR_ERR_SIZE;
  }
}
