// racing_step.cu -- one fused sm_100a kernel per env.step() of the racing task (one thread per env).
//
// Reference path replaced (relative to the reference root; QD = extensions/diff.lab_tasks/diff/
// lab_tasks/tasks/quadcopter_diff, L = extensions/diff.lab/diff/lab): the ~700 eager aten launches of
// ManagerBasedDiffRLEnv.step (L/envs/manager_based_diff_rl_env.py:160-267).  Section comments carry the
// file:line of what each block computes.  Data layout: SoA of float4 "planes" (see gr_common.cuh);
// every persistent column is read once and written once per step with 128-bit coalesced accesses; the
// gate table slice of the block's terrain types is staged in shared memory.
#include "racing_step_core.cuh"

namespace gr {

template <bool kNoise, bool kDiff, bool kPhilox, bool kStats>
__global__ void __launch_bounds__(256) racing_step_fwd_kernel(const GrConfig cfg, const GrTrack track, const GrState st,
                                                              const GrRandom rng, const GrStepIO io) {
  GR_DYN_SMEM(float4, smem_rows);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = i < st.num_envs;
  const int li = active ? i : st.num_envs - 1;        // inactive threads shadow the last env (loads only)
  float4* __restrict__ tile = tile_ptr(reinterpret_cast<float4*>(st.planes), li);

  // Two orderings of the same prologue.  Launched with programmatic dependent launch (st.launch_flags & GR_LAUNCH_PDL)
  // the part that does not depend on earlier kernels -- Philox draws, gate-table staging -- runs BEFORE the grid
  // dependency wait and overlaps the tail of the previous kernel; otherwise the state loads are issued first so that
  // the same work hides under their latency.
  const bool pdl = (st.launch_flags & GR_LAUNCH_PDL) != 0;
  GR_STAMP(0);
  EnvRegs e;
  float4 a_t, eps0 = make_float4(0.f, 0.f, 0.f, 0.f), lsum = eps0;
  if (!pdl) {
    load_env<kNoise>(e, tile);
    a_t = __ldcs(reinterpret_cast<const float4*>(io.action) + li);
    if (kStats) { eps0 = ld_plane(tile, PL_EPSUM0); if (kDiff) lsum = ld_plane(tile, PL_LOSSSUM); }
  }
  const RandSrc<kPhilox> rs(rng, li, st.env_id_offset + li);
  float4 n01, n23;                                     // obs normals (slots 0..5), thr_est_error normal (slot 6)
  if (kPhilox || !pdl) rs.normals8(n01, n23);
  // PDL: the read-mostly planes (drag, gains, filter constants, command noise: 112 of the 272 B an env reads) are
  // fetched BEFORE the grid dependency and overlap the previous kernel's tail.  They are rewritten only by the reset /
  // gate-switch tail of a step kernel, which flags it in the hot planes (fresh bit, ANGACC.w): a flagged env re-reads
  // them through L2 after the wait.  (Host-side edits of those planes must clear GR_LAUNCH_PREFETCH for the next step.)
  // GR_LAUNCH_PREFETCH_L2 instead: the same planes are only pulled into L2 here (no registers, nothing to go stale) and loaded
  // with everything else after the wait, as L2 hits -- no second, data-dependent round trip for the envs that reset.
  const bool prefetch_l2 = pdl && (st.launch_flags & GR_LAUNCH_PREFETCH_L2) != 0;
  const bool prefetch = pdl && !prefetch_l2 && (st.launch_flags & GR_LAUNCH_PREFETCH) != 0;
  if (prefetch) { load_cold<false>(e, tile); load_noise<kNoise, false>(e, tile); }
  if (prefetch_l2) {
#pragma unroll
    for (int p = PL_DRAG2; p <= (kNoise ? PL_NOISE1 : PL_ETAU); ++p) prefetch_l2_line(tile + p * kTile);
  }
  const TrackSmem tr = stage_track(track, reinterpret_cast<const int2*>(st.chunk_types), st.num_envs, smem_rows);
  if (pdl) {
    pdl_wait();
    GR_STAMP(1);
    load_hot(e, tile);
    a_t = __ldcs(reinterpret_cast<const float4*>(io.action) + li);
    if (kStats) { eps0 = ld_plane(tile, PL_EPSUM0); if (kDiff) lsum = ld_plane(tile, PL_LOSSSUM); }
    if (!kPhilox) rs.normals8(n01, n23);
    if (prefetch_l2) {
      load_cold<true>(e, tile);
      load_noise<kNoise, true>(e, tile);
    } else if (prefetch) {
      const bool stale = pk_fresh(e.pk) != 0u;
      if (stale) load_cold<true>(e, tile);
      if (kNoise && (stale || e.noise_dirty_prev)) load_noise<kNoise, true>(e, tile);
    } else {
      load_cold<false>(e, tile);
      load_noise<kNoise, false>(e, tile);
    }
  }
  // (measured: generating the rare-path draws speculatively for every env while the loads are in flight costs more
  //  than it saves -- +1.2 us median compute, stragglers unchanged -- so the reset / pass tails draw on demand)
  __shared__ float4 obs_stage[8 * 128];                 // 2 KB per warp, up to 8 warps per block (GlobalObsSink)
  // stage_reset_draws ([7 calls][<= 64 threads], GR_LAUNCH_COOP_RESET with blocks of <= 64 threads) borrows the part of the staging buffer
  // that warps 2..7 would use: static shared memory stays at 16 KB (more costs a resident block per SM, i.e. the single wave)
  const bool coop = kPhilox && (st.launch_flags & GR_LAUNCH_COOP_RESET) != 0 && blockDim.x <= 64;
  const Draws<kPhilox> draws{rs, coop ? obs_stage + 2 * 128 : nullptr};
  pdl_launch_dependents();
  // a warp that is entirely past the last env leaves; in the (single) ragged warp the inactive lanes keep shadowing
  // the last env so the warp collectives below stay full-width, and skip every store
  const unsigned live = __ballot_sync(0xffffffffu, active);
  if (live == 0u) return;

  StepOut so;
  GlobalObsSink sink{io, live, obs_stage + (threadIdx.x >> 5) * 128};
  if (st.launch_flags & GR_LAUNCH_EARLY_STORE) sink.tile = tile;
  if (!racing_step_body<kNoise, kDiff, kPhilox, kStats>(cfg, tr, e, a_t, n01, n23, draws, eps0, lsum, io, i, active, sink, so)) return;

  GR_STAMP(3);
  // ---- 12. outputs + state write-back ----
  if (!so.stored) {
    if (kStats && !so.reset) add_episode_sums(eps0, e, so.terms, cfg.dt);        // deferred to here: the episode-sum plane is the last load to arrive
    store_env<kNoise>(e, tile, so.reset, so.noise_dirty);
    if (kStats) {
      st_plane(tile, PL_EPSUM0, eps0);
      if (kDiff) st_plane(tile, PL_LOSSSUM, lsum);
    }
  }
  io.reward[i] = so.reward;
  io.terminated[i] = so.terminated ? 1 : 0;
  io.time_out[i] = so.time_out ? 1 : 0;
  if (io.dones) io.dones[i] = so.reset ? 1 : 0;
  if (io.dones_u8) io.dones_u8[i] = so.reset ? 1 : 0;
  if (io.gate_passed) io.gate_passed[i] = so.passed ? 1 : 0;
  if (io.reward_terms) {
#pragma unroll
    for (int k = 0; k < GR_NUM_REWARD_TERMS; ++k) io.reward_terms[i * GR_NUM_REWARD_TERMS + k] = so.terms[k];
  }
  GR_STAMP(4);
}

// =============================================================================================
// T steps in one launch for actions known in advance (gr_rollout_fwd): the same body, the env state in registers over
// the window.  One thread per env; per step it reads its action (16 B) and writes what the caller asked to record.
// =============================================================================================
// kObs: does this step record / hand out an observation?  A COMPILE-TIME property of the loop section the step belongs to (measured: as a
// run-time test inside the body the observation section became a scheduling barrier -- 4.06 -> 4.35 us per step at 65,536 envs).
template <bool kObs>
struct RolloutObsSink {
  GlobalObsSink g;                                     // rows(): the coalesced warp store through shared memory
  float4* seq_rows; float4* out_rows; float4* critic_rows; float* aux_ptr;        // destinations of THIS step (nullptr: skip)
  __device__ __forceinline__ void policy(int i, float4 o0, float4 o1, float4 o2, float4 o3) const {
    if (seq_rows) g.rows(seq_rows, i, o0, o1, o2, o3);
    if (out_rows) g.rows(out_rows, i, o0, o1, o2, o3);
  }
  template <bool kNoise, bool kDiff, bool kStats>
  __device__ __forceinline__ bool state_final(EnvRegs&, float4&, const float4&, const float (&)[GR_NUM_REWARD_TERMS], float, bool, bool) const { return false; }
  // a step whose observation nobody records skips the whole observation section (and its eight normals)
  __device__ __forceinline__ constexpr bool wants_policy() const { return kObs; }
  __device__ __forceinline__ bool wants_critic() const { return critic_rows != nullptr; }
  __device__ __forceinline__ void critic(int i, float4 c0, float4 c1, float4 c2, float4 c3) const { g.rows(critic_rows, i, c0, c1, c2, c3); }
  __device__ __forceinline__ void aux(int i, float v) const { if (aux_ptr) aux_ptr[i] = v; }
};

// loop-carried bookkeeping of a window (what T single steps would leave behind in the planes)
struct RolloutCarry { bool any_reset = false, any_noise_dirty = false, last_noise_dirty = false; };

// one step of the window: action a_t is in hand, the next one is fetched one step ahead
template <bool kNoise, bool kDiff, bool kPhilox, bool kStats, bool kObs>
__device__ __forceinline__ void rollout_one_step(const GrConfig& cfg, const TrackSmem& tr, const GrState& st, const GrRandom& rng, const GrRolloutIO& rio,
                                                 const GlobalObsSink& gsink, float4* coop_draws, const int t, const int i, const int li,
                                                 const bool active, EnvRegs& e, float4& eps0, float4& lsum, float4& a_next, RolloutCarry& c) {
  const int N = st.num_envs, T = rio.T;
  const int64_t tn = (int64_t)t * N + i;
  const float4 a_t = a_next;
  if (t + 1 < T) a_next = __ldcs(reinterpret_cast<const float4*>(rio.actions) + (int64_t)(t + 1) * N + li);      // one step ahead
  GrRandom rt = rng;
  rt.step = rng.step + (uint32_t)t;
  if (!kPhilox) rt.rnd = rng.rnd + (int64_t)t * N * GR_RND_STRIDE;
  const RandSrc<kPhilox> rs(rt, li, st.env_id_offset + li);
  const bool last = t == T - 1;
  float4 n01 = make_float4(0.f, 0.f, 0.f, 0.f), n23 = n01;
  if (kObs) rs.normals8(n01, n23);
  const Draws<kPhilox> draws{rs, coop_draws, kObs};
  GrStepIO io = {};
  io.log_accum = rio.log_accum;
  if (kDiff) {
    const int64_t tape_step = (int64_t)(rio.tape_stride / kTile) * GR_TAPE_PLANES * kTile * 4;       // floats per tape step
    io.tape = rio.tape ? rio.tape + (int64_t)t * tape_step : nullptr;
    io.tape_stride = rio.tape_stride;
    io.loss = rio.loss ? rio.loss + (int64_t)t * N : nullptr;
    io.loss_terms = rio.loss_terms ? rio.loss_terms + (int64_t)t * N * 3 : nullptr;
  }
  RolloutObsSink<kObs> sink{gsink, nullptr, nullptr, nullptr, nullptr};
  if (kObs) {
    if (rio.obs_seq) sink.seq_rows = reinterpret_cast<float4*>(rio.obs_seq) + (int64_t)t * N * 4;
    if (last) {
      sink.out_rows = reinterpret_cast<float4*>(rio.obs_out);
      sink.critic_rows = reinterpret_cast<float4*>(rio.critic_obs_out);
      sink.aux_ptr = rio.aux_out;
    }
  }
  StepOut so;
  const bool alive = racing_step_body<kNoise, kDiff, kPhilox, kStats>(cfg, tr, e, a_t, n01, n23, draws, eps0, lsum, io, i, active, sink, so);
  if (alive) {
    if (kStats && !so.reset) add_episode_sums(eps0, e, so.terms, cfg.dt);
    c.any_reset |= so.reset;
    c.any_noise_dirty |= so.noise_dirty;
    c.last_noise_dirty = so.noise_dirty;
    if (rio.reward) rio.reward[tn] = so.reward;
    if (rio.dones) rio.dones[tn] = so.reset ? 1 : 0;
    if (rio.terminated) rio.terminated[tn] = so.terminated ? 1 : 0;
    if (rio.time_out) rio.time_out[tn] = so.time_out ? 1 : 0;
  }
}

// 216-225 registers, no spills: 4 blocks of 64 threads per SM, i.e. two waves at 65,536 envs.  (Measured: capping the kernel at 128
// registers -- a handful of spilled words, one wave of 8 blocks per SM -- is SLOWER at every size tried: 4.22 vs 3.90 us per step at
// 65,536 envs, 2.66 vs 2.00 us at 16,384 with tape; tools/rollout_bench.py.)
// kRecord: the observation of EVERY step is recorded (rio.obs_seq); otherwise only the last step computes one -- two loop sections, each
// with its observation section decided at compile time.
template <bool kNoise, bool kDiff, bool kPhilox, bool kStats, bool kRecord>
__global__ void __launch_bounds__(256) racing_rollout_fwd_kernel(const GrConfig cfg, const GrTrack track, const GrState st,
                                                                 const GrRandom rng, const GrRolloutIO rio) {
  GR_DYN_SMEM(float4, smem_rows);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int N = st.num_envs, T = rio.T;
  const bool active = i < N;
  const int li = active ? i : N - 1;
  float4* __restrict__ tile = tile_ptr(reinterpret_cast<float4*>(st.planes), li);
  EnvRegs e;
  load_env<kNoise>(e, tile);
  float4 eps0 = make_float4(0.f, 0.f, 0.f, 0.f), lsum = eps0;
  if (kStats) { eps0 = ld_plane(tile, PL_EPSUM0); if (kDiff) lsum = ld_plane(tile, PL_LOSSSUM); }
  const TrackSmem tr = stage_track(track, reinterpret_cast<const int2*>(st.chunk_types), N, smem_rows);
  const unsigned live = __ballot_sync(0xffffffffu, active);
  if (live == 0u) return;
  __shared__ float4 obs_stage[8 * 128];
  const GrStepIO no_io = {};
  const GlobalObsSink gsink{no_io, live, obs_stage + (threadIdx.x >> 5) * 128};
  RolloutCarry c;
  const bool coop = kPhilox && (st.launch_flags & GR_LAUNCH_COOP_RESET) != 0 && blockDim.x <= 64;
  float4* const coop_draws = coop ? obs_stage + 2 * 128 : nullptr;       // (the staging rows of warps 2..7: unused with blocks of <= 64 threads)

  float4 a_next = __ldcs(reinterpret_cast<const float4*>(rio.actions) + li);
  if (kRecord) {
#pragma unroll 1
    for (int t = 0; t < T; ++t)
      rollout_one_step<kNoise, kDiff, kPhilox, kStats, true>(cfg, tr, st, rng, rio, gsink, coop_draws, t, i, li, active, e, eps0, lsum, a_next, c);
  } else {
#pragma unroll 1
    for (int t = 0; t < T - 1; ++t)
      rollout_one_step<kNoise, kDiff, kPhilox, kStats, false>(cfg, tr, st, rng, rio, gsink, coop_draws, t, i, li, active, e, eps0, lsum, a_next, c);
    rollout_one_step<kNoise, kDiff, kPhilox, kStats, true>(cfg, tr, st, rng, rio, gsink, coop_draws, T - 1, i, li, active, e, eps0, lsum, a_next, c);
  }
  if (active) {      // what T single steps leave behind: cold planes rewritten if any step reset, the noise-dirty flag = the LAST step's
    store_env<kNoise>(e, tile, c.any_reset, c.any_noise_dirty);
    if (kNoise && c.any_noise_dirty && !c.last_noise_dirty)
      st_plane(tile, PL_LINVEL, pack(e.v, __uint_as_float(eplen_word(e.eplen, e.aux != 0.0f, false, e.arate, e.metrics_zero))));
    if (kStats) { st_plane(tile, PL_EPSUM0, eps0); if (kDiff) st_plane(tile, PL_LOSSSUM, lsum); }
  }
}

// =============================================================================================
// reset / observe: ManagerBasedRLEnv.reset() = _reset_idx(ids) + observation_manager.compute();
// with an all-zero mask this is observation_manager.compute() alone (get_observations).
// The "last action" observation uses the FIFO content (a_t) -- see DESIGN.md (get_observations caveat).
// =============================================================================================
template <bool kNoise, bool kPhilox, bool kStats>
__global__ void __launch_bounds__(256) racing_reset_kernel(const GrConfig cfg, const GrTrack track, const GrState st, const GrRandom rng,
                                                           const uint8_t* __restrict__ mask, const int mode /*0 mask,1 all,2 none*/,
                                                           float* __restrict__ obs, float* __restrict__ critic, float* __restrict__ aux_out) {
  GR_DYN_SMEM(float4, smem_rows);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = i < st.num_envs;
  const int li = active ? i : st.num_envs - 1;
  float4* __restrict__ tile = tile_ptr(reinterpret_cast<float4*>(st.planes), li);
  EnvRegs e;
  load_env<kNoise>(e, tile);
  const RandSrc<kPhilox> rs(rng, li, st.env_id_offset + li);
  float4 n01, n23;
  rs.normals8(n01, n23);
  const TrackSmem tr = stage_track(track, reinterpret_cast<const int2*>(st.chunk_types), st.num_envs, smem_rows);
  if (!active) return;
  const bool do_reset = mode == 1 || (mode == 0 && mask[i] != 0);
  const int type = (int)pk_type(e.pk);
  V3 origin;
  if (do_reset) {
    origin = reset_env<kNoise, kPhilox>(cfg, tr, e, Draws<kPhilox>{rs, nullptr}, n23.z);
    e.arate = 0.f; e.metrics_zero = true;                 // CommandTerm.reset: metrics logged and zeroed, no command update follows
    store_env<kNoise>(e, tile, true, true);
    if (kStats) { st_plane(tile, PL_EPSUM0, make_float4(0.f, 0.f, 0.f, 0.f)); st_plane(tile, PL_LOSSSUM, make_float4(0.f, 0.f, 0.f, 0.f)); }
  } else {
    origin = xyz(tr.origin_row(type, (int)pk_level(e.pk)));
  }
  const int level = (int)pk_level(e.pk), gate_id = (int)pk_gate(e.pk);
  if (obs) {
    const V3 gate_rel = tr.gate(type, level, gate_id), next_rel = tr.gate(type, level, (gate_id + 1) % tr.gates);
    write_observations<kNoise>(cfg, e, origin, gate_rel, next_rel, e.fifo, n01, n23, e.aux, i, obs, critic, aux_out);
  }
}

// =============================================================================================
// startup: construction-time state + domain randomisation (QD/mdp/events.py:105-137,
// QD/mdp/diff_action.py:86, QD/mdp/dynamics/droneDynamics.py:23-34, TerrainImporter env-origin assignment)
// =============================================================================================
__device__ __forceinline__ void startup_draws(const float* __restrict__ srnd, uint64_t seed, int i, int env_id, float (&s)[GR_SRND_STRIDE]) {
  if (srnd) {
    const float4* row = reinterpret_cast<const float4*>(srnd) + (int64_t)i * (GR_SRND_STRIDE / 4);
#pragma unroll
    for (int c = 0; c < GR_SRND_STRIDE / 4; ++c) { const float4 v = __ldg(row + c); s[4 * c] = v.x; s[4 * c + 1] = v.y; s[4 * c + 2] = v.z; s[4 * c + 3] = v.w; }
  } else {
    const Philox ph(seed, (uint32_t)env_id, 0xFFFFFFFFu);
#pragma unroll
    for (int c = 0; c < 3; ++c) { const uint4 x = ph((uint32_t)c); s[4 * c] = u01(x.x); s[4 * c + 1] = u01(x.y); s[4 * c + 2] = u01(x.z); s[4 * c + 3] = u01(x.w); }
    const uint4 x = ph(3u);
    const float2 a = box_muller(x.x, x.y), b = box_muller(x.z, x.w);
    s[12] = a.x; s[13] = a.y; s[14] = b.x; s[15] = b.y;
  }
}

__global__ void racing_startup_kernel(const GrConfig cfg, const GrTrack track, const GrState st, const int32_t* __restrict__ terrain_types,
                                      int32_t* __restrict__ chunk_types, const float* __restrict__ srnd, const uint64_t seed) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= st.num_envs) return;
  float4* __restrict__ P = reinterpret_cast<float4*>(st.planes);
  float s[GR_SRND_STRIDE];
  startup_draws(srnd, seed, i, st.env_id_offset + i, s);
  const int type = terrain_types[i];
  if ((i & 63) == 0) {
    const int j = min(i + 63, st.num_envs - 1);
    chunk_types[2 * (i >> 6)] = type;
    chunk_types[2 * (i >> 6) + 1] = terrain_types[j];
  }
  int max_init = cfg.max_init_level < track.levels - 1 ? cfg.max_init_level : track.levels - 1;
  int level = (int)floorf(s[10] * (float)(max_init + 1));
  if (level > max_init) level = max_init;
  const float ps = cfg.pid_scale_span, ds = cfg.delay_scale_span;
  const V3 kp = v3(cfg.kp[0] * (s[0] * ps + cfg.pid_scale_lo), cfg.kp[1] * (s[1] * ps + cfg.pid_scale_lo), cfg.kp[2] * (s[2] * ps + cfg.pid_scale_lo));
  const V3 kd = v3(cfg.kd[0] * (s[3] * ps + cfg.pid_scale_lo), cfg.kd[1] * (s[4] * ps + cfg.pid_scale_lo), cfg.kd[2] * (s[5] * ps + cfg.pid_scale_lo));
  const float thrust_delay = cfg.thrust_delay * (s[6] * ds + cfg.delay_scale_lo);
  const V3 torque_delay = v3(cfg.torque_delay[0] * (s[7] * ds + cfg.delay_scale_lo), cfg.torque_delay[1] * (s[8] * ds + cfg.delay_scale_lo),
                             cfg.torque_delay[2] * (s[9] * ds + cfg.delay_scale_lo));
  const float m = cfg.mass;
  const float b2 = cfg.drag2 * m, b1 = cfg.drag1 * m;
  const float zero = 0.0f;
  P[pidx(PL_QUAT, i)] = make_float4(1.f, 0.f, 0.f, 0.f);
  P[pidx(PL_POS, i)] = make_float4(zero, zero, zero, zero);
  P[pidx(PL_LINVEL, i)] = make_float4(zero, zero, zero, __uint_as_float(eplen_word(0, false, false, 0.f, true)));
  P[pidx(PL_ANGVEL, i)] = make_float4(zero, zero, zero, __uint_as_float(pk_make(0u, 0u, (uint32_t)level, (uint32_t)type, 1u)));
  P[pidx(PL_TORQUE, i)] = make_float4(zero, zero, zero, zero);
  P[pidx(PL_ANGACC, i)] = make_float4(zero, zero, zero, zero);
  P[pidx(PL_FIFO, i)] = make_float4(zero, zero, zero, zero);
  P[pidx(PL_DRAG2, i)] = make_float4(b2, b2, b2 * cfg.z_drag, m);
  P[pidx(PL_DRAG1, i)] = make_float4(b1, b1, b1 * cfg.z_drag, expf(-cfg.dt / thrust_delay));
  P[pidx(PL_KP, i)] = pack(kp, 1.0f + s[12] * cfg.thr_err_init_std);
  P[pidx(PL_KD, i)] = pack(kd, 0.f);
  P[pidx(PL_ETAU, i)] = make_float4(expf(-cfg.dt / torque_delay.x), expf(-cfg.dt / torque_delay.y), expf(-cfg.dt / torque_delay.z), 0.f);
  P[pidx(PL_NOISE0, i)] = make_float4(zero, zero, zero, zero);
  P[pidx(PL_NOISE1, i)] = make_float4(zero, zero, cfg.cmd_noise_pos, 1.0f);
  if (st.num_planes >= GR_NUM_PLANES_WITH_STATS) {
    P[pidx(PL_EPSUM0, i)] = make_float4(zero, zero, zero, zero);
    P[pidx(PL_LOSSSUM, i)] = make_float4(zero, zero, zero, zero);
  }
}

__global__ void fill_rand_kernel(float4* __restrict__ out, const int num_envs, const int env_id_offset, const uint64_t seed, const uint32_t step) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int calls = GR_RND_STRIDE / 4;
  if (idx >= num_envs * calls) return;
  const int i = idx / calls, c = idx - i * calls;
  GrRandom r; r.rnd = nullptr; r.seed = seed; r.step = step;
  const RandSrc<true> rs(r, i, env_id_offset + i);
  if (c < 2) {
    float4 a, b;
    rs.normals8(a, b);
    out[idx] = c == 0 ? a : b;
  } else {
    out[idx] = rs.get4(c);
  }
}

// the square root of the gate predicate `|gate - pos| < update_threshold`, exposed for a test: it must stay correctly rounded whatever
// the library-wide -prec-sqrt / -prec-div flags are (a 1-ulp error flips gate passes against the reference)
__global__ void sqrt_rn_kernel(const float* __restrict__ x, float* __restrict__ y, const int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = sqrt_rn(x[i]);
}

__global__ void fill_startup_rand_kernel(float* __restrict__ out, const int num_envs, const int env_id_offset, const uint64_t seed) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= num_envs) return;
  float s[GR_SRND_STRIDE];
  startup_draws(nullptr, seed, i, env_id_offset + i, s);
#pragma unroll
  for (int k = 0; k < GR_SRND_STRIDE; ++k) out[(int64_t)i * GR_SRND_STRIDE + k] = s[k];
}

}  // namespace gr

#ifndef GR_CPU_EMUL   // host API (the CPU emulation harness in tests/emul includes only the device code)
// =============================================================================================
// C ABI
// =============================================================================================
using namespace gr;

static inline bool misaligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) != 0; }

static int check_common(const GrConfig* cfg, const GrTrack* tr, const GrState* st) {
  if (!cfg || !tr || !st || !st->planes || !tr->rows || !st->chunk_types) return GR_ERR_NULL;
  if (st->num_envs <= 0 || st->plane_stride < ((st->num_envs + 31) & ~31)) return GR_ERR_SIZE;
  if (st->num_planes != GR_NUM_PLANES && st->num_planes != GR_NUM_PLANES_WITH_STATS) return GR_ERR_SIZE;
  if (tr->types < 1 || tr->types > 32 || tr->levels < 1 || tr->levels > 64 || tr->gates < 1 || tr->gates > GR_MAX_GATES) return GR_ERR_SIZE;
  if (misaligned16(st->planes) || misaligned16(tr->rows)) return GR_ERR_ALIGN;
  if (st->max_types_per_block < 1 || st->max_types_per_block > tr->types) return GR_ERR_SIZE;
  if (st->block_threads < 0 || st->block_threads > 256 || (st->block_threads & 31)) return GR_ERR_SIZE;
  return GR_OK;
}

static inline int block_threads(const GrState* st) { return st->block_threads > 0 ? st->block_threads : 64; }

static inline size_t track_smem_bytes(const GrTrack* tr, const GrState* st) {
  return (size_t)st->max_types_per_block * tr->levels * (tr->gates + 1) * sizeof(float4);
}

template <typename K>
static int prepare_smem(K kernel, size_t bytes) {
  if (bytes > 200 * 1024) return GR_ERR_SMEM;
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
  }
  return GR_OK;
}

template <bool kNoise, bool kDiff, bool kPhilox, bool kStats>
static int launch_step(const GrConfig* cfg, const GrTrack* tr, const GrState* st, const GrRandom* rng, const GrStepIO* io, cudaStream_t s) {
  auto kernel = racing_step_fwd_kernel<kNoise, kDiff, kPhilox, kStats>;
  const size_t bytes = track_smem_bytes(tr, st);
  int rc = prepare_smem(kernel, bytes);
  if (rc != GR_OK) return rc;
  const int kBlock = block_threads(st);
  const int grid = (st->num_envs + kBlock - 1) / kBlock;
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3((unsigned)grid);
  lc.blockDim = dim3((unsigned)kBlock);
  lc.dynamicSmemBytes = bytes;
  lc.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr;
  lc.numAttrs = (st->launch_flags & GR_LAUNCH_PDL) ? 1 : 0;
  return (int)cudaLaunchKernelEx(&lc, kernel, *cfg, *tr, *st, *rng, *io);
}

template <bool kNoise, bool kDiff, bool kPhilox>
static int dispatch_stats(bool stats, const GrConfig* cfg, const GrTrack* tr, const GrState* st, const GrRandom* rng, const GrStepIO* io, cudaStream_t s) {
  return stats ? launch_step<kNoise, kDiff, kPhilox, true>(cfg, tr, st, rng, io, s) : launch_step<kNoise, kDiff, kPhilox, false>(cfg, tr, st, rng, io, s);
}
template <bool kNoise, bool kDiff>
static int dispatch_philox(bool philox, bool stats, const GrConfig* cfg, const GrTrack* tr, const GrState* st, const GrRandom* rng, const GrStepIO* io, cudaStream_t s) {
  return philox ? dispatch_stats<kNoise, kDiff, true>(stats, cfg, tr, st, rng, io, s) : dispatch_stats<kNoise, kDiff, false>(stats, cfg, tr, st, rng, io, s);
}
template <bool kNoise>
static int dispatch_diff(bool diff, bool philox, bool stats, const GrConfig* cfg, const GrTrack* tr, const GrState* st, const GrRandom* rng, const GrStepIO* io, cudaStream_t s) {
  return diff ? dispatch_philox<kNoise, true>(philox, stats, cfg, tr, st, rng, io, s) : dispatch_philox<kNoise, false>(philox, stats, cfg, tr, st, rng, io, s);
}

extern "C" int gr_abi_version(void) { return GR_ABI_VERSION; }

extern "C" int gr_step_fwd(const GrConfig* cfg, const GrTrack* track, const GrState* st, const GrRandom* rng, const GrStepIO* io, void* stream) {
  int rc = check_common(cfg, track, st);
  if (rc != GR_OK) return rc;
  if (!rng || !io || !io->action || !io->obs || !io->reward || !io->terminated || !io->time_out) return GR_ERR_NULL;
  if (misaligned16(io->action) || misaligned16(io->obs) || (io->critic_obs && misaligned16(io->critic_obs)) ||
      (rng->rnd && misaligned16(rng->rnd)) || (io->tape && misaligned16(io->tape)))
    return GR_ERR_ALIGN;
  if (io->tape && io->tape_stride < ((st->num_envs + 31) & ~31)) return GR_ERR_SIZE;
  const bool diff = io->loss != nullptr || io->tape != nullptr || io->loss_terms != nullptr;
  const bool stats = st->num_planes == GR_NUM_PLANES_WITH_STATS;
  const bool philox = rng->rnd == nullptr;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return cfg->add_cmd_noise ? dispatch_diff<true>(diff, philox, stats, cfg, track, st, rng, io, s)
                            : dispatch_diff<false>(diff, philox, stats, cfg, track, st, rng, io, s);
}

template <bool kNoise, bool kDiff, bool kPhilox, bool kStats, bool kRecord>
static int launch_rollout_k(const GrConfig* cfg, const GrTrack* tr, const GrState* st, const GrRandom* rng, const GrRolloutIO* io, cudaStream_t s) {
  const size_t bytes = track_smem_bytes(tr, st);
  const int kBlock = block_threads(st);
  const int grid = (st->num_envs + kBlock - 1) / kBlock;
  auto kernel = racing_rollout_fwd_kernel<kNoise, kDiff, kPhilox, kStats, kRecord>;
  int rc = prepare_smem(kernel, bytes);
  if (rc != GR_OK) return rc;
  kernel<<<grid, kBlock, bytes, s>>>(*cfg, *tr, *st, *rng, *io);
  return (int)cudaGetLastError();
}

template <bool kNoise, bool kDiff, bool kPhilox>
static int launch_rollout(bool stats, const GrConfig* cfg, const GrTrack* tr, const GrState* st, const GrRandom* rng, const GrRolloutIO* io, cudaStream_t s) {
  const bool record = io->obs_seq != nullptr;
  if (stats) return record ? launch_rollout_k<kNoise, kDiff, kPhilox, true, true>(cfg, tr, st, rng, io, s) : launch_rollout_k<kNoise, kDiff, kPhilox, true, false>(cfg, tr, st, rng, io, s);
  return record ? launch_rollout_k<kNoise, kDiff, kPhilox, false, true>(cfg, tr, st, rng, io, s) : launch_rollout_k<kNoise, kDiff, kPhilox, false, false>(cfg, tr, st, rng, io, s);
}

extern "C" int gr_rollout_fwd(const GrConfig* cfg, const GrTrack* track, const GrState* st, const GrRandom* rng, const GrRolloutIO* io, void* stream) {
  int rc = check_common(cfg, track, st);
  if (rc != GR_OK) return rc;
  if (!rng || !io || !io->actions || !io->obs_out) return GR_ERR_NULL;
  if (io->T < 1) return GR_ERR_SIZE;
  if (misaligned16(io->actions) || misaligned16(io->obs_out) || (io->critic_obs_out && misaligned16(io->critic_obs_out)) ||
      (io->obs_seq && misaligned16(io->obs_seq)) || (rng->rnd && misaligned16(rng->rnd)) || (io->tape && misaligned16(io->tape)))
    return GR_ERR_ALIGN;
  if (io->tape && (io->tape_stride < ((st->num_envs + 31) & ~31) || (io->tape_stride & 31))) return GR_ERR_SIZE;
  const bool diff = io->loss != nullptr || io->tape != nullptr || io->loss_terms != nullptr;
  const bool stats = st->num_planes == GR_NUM_PLANES_WITH_STATS;
  const bool philox = rng->rnd == nullptr;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
#define GR_GO(a, b, c) if ((cfg->add_cmd_noise != 0) == a && diff == b && philox == c) return launch_rollout<a, b, c>(stats, cfg, track, st, rng, io, s);
  GR_GO(false, false, false) GR_GO(false, false, true) GR_GO(false, true, false) GR_GO(false, true, true)
  GR_GO(true, false, false) GR_GO(true, false, true) GR_GO(true, true, false) GR_GO(true, true, true)
#undef GR_GO
  return GR_ERR_CONFIG;
}

template <bool kNoise, bool kPhilox>
static int launch_reset(bool stats, const GrConfig* cfg, const GrTrack* tr, const GrState* st, const GrRandom* rng, const uint8_t* mask, int mode,
                        float* obs, float* critic, float* aux, cudaStream_t s) {
  const size_t bytes = track_smem_bytes(tr, st);
  const int kBlock = block_threads(st);
  const int grid = (st->num_envs + kBlock - 1) / kBlock;
  if (stats) {
    auto kernel = racing_reset_kernel<kNoise, kPhilox, true>;
    int rc = prepare_smem(kernel, bytes);
    if (rc != GR_OK) return rc;
    kernel<<<grid, kBlock, bytes, s>>>(*cfg, *tr, *st, *rng, mask, mode, obs, critic, aux);
  } else {
    auto kernel = racing_reset_kernel<kNoise, kPhilox, false>;
    int rc = prepare_smem(kernel, bytes);
    if (rc != GR_OK) return rc;
    kernel<<<grid, kBlock, bytes, s>>>(*cfg, *tr, *st, *rng, mask, mode, obs, critic, aux);
  }
  return (int)cudaGetLastError();
}

static int reset_impl(const GrConfig* cfg, const GrTrack* track, const GrState* st, const GrRandom* rng, const uint8_t* mask, int mode,
                      float* obs, float* critic, float* aux, void* stream) {
  int rc = check_common(cfg, track, st);
  if (rc != GR_OK) return rc;
  if (!rng) return GR_ERR_NULL;
  if ((obs && misaligned16(obs)) || (critic && misaligned16(critic)) || (rng->rnd && misaligned16(rng->rnd))) return GR_ERR_ALIGN;
  if (critic && !obs) return GR_ERR_NULL;
  const bool stats = st->num_planes == GR_NUM_PLANES_WITH_STATS;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const bool philox = rng->rnd == nullptr;
  if (cfg->add_cmd_noise)
    return philox ? launch_reset<true, true>(stats, cfg, track, st, rng, mask, mode, obs, critic, aux, s)
                  : launch_reset<true, false>(stats, cfg, track, st, rng, mask, mode, obs, critic, aux, s);
  return philox ? launch_reset<false, true>(stats, cfg, track, st, rng, mask, mode, obs, critic, aux, s)
                : launch_reset<false, false>(stats, cfg, track, st, rng, mask, mode, obs, critic, aux, s);
}

extern "C" int gr_env_reset(const GrConfig* cfg, const GrTrack* track, const GrState* st, const GrRandom* rng, const uint8_t* reset_mask,
                            float* obs, float* critic_obs, float* aux_obs, void* stream) {
  return reset_impl(cfg, track, st, rng, reset_mask, reset_mask ? 0 : 1, obs, critic_obs, aux_obs, stream);
}

extern "C" int gr_env_observe(const GrConfig* cfg, const GrTrack* track, const GrState* st, const GrRandom* rng, float* obs, float* critic_obs,
                              float* aux_obs, void* stream) {
  if (!obs) return GR_ERR_NULL;
  return reset_impl(cfg, track, st, rng, nullptr, 2, obs, critic_obs, aux_obs, stream);
}

extern "C" int gr_env_startup(const GrConfig* cfg, const GrTrack* track, const GrState* st, const int32_t* terrain_types, int32_t* chunk_types_out,
                              const float* srnd, uint64_t seed, void* stream) {
  if (!cfg || !track || !st || !st->planes || !terrain_types || !chunk_types_out) return GR_ERR_NULL;
  if (st->num_envs <= 0 || st->plane_stride < ((st->num_envs + 31) & ~31)) return GR_ERR_SIZE;
  if (st->num_planes != GR_NUM_PLANES && st->num_planes != GR_NUM_PLANES_WITH_STATS) return GR_ERR_SIZE;
  if (misaligned16(st->planes) || (srnd && misaligned16(srnd))) return GR_ERR_ALIGN;
  if (track->types < 1 || track->types > 32 || track->levels < 1 || track->levels > 64) return GR_ERR_SIZE;
  const int grid = (st->num_envs + 127) / 128;
  racing_startup_kernel<<<grid, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*cfg, *track, *st, terrain_types, chunk_types_out, srnd, seed);
  return (int)cudaGetLastError();
}

extern "C" int gr_fill_rand(float* rnd, int32_t num_envs, int32_t env_id_offset, uint64_t seed, uint32_t step, void* stream) {
  if (!rnd) return GR_ERR_NULL;
  if (num_envs <= 0) return GR_ERR_SIZE;
  if (misaligned16(rnd)) return GR_ERR_ALIGN;
  const int total = num_envs * (GR_RND_STRIDE / 4);
  fill_rand_kernel<<<(total + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<float4*>(rnd), num_envs, env_id_offset, seed, step);
  return (int)cudaGetLastError();
}

extern "C" int gr_selftest_sqrt_rn(const float* x, float* y, int64_t n, void* stream) {
  if (!x || !y) return GR_ERR_NULL;
  if (n <= 0) return GR_ERR_SIZE;
  sqrt_rn_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, y, n);
  return (int)cudaGetLastError();
}

extern "C" int gr_fill_startup_rand(float* srnd, int32_t num_envs, int32_t env_id_offset, uint64_t seed, void* stream) {
  if (!srnd) return GR_ERR_NULL;
  if (num_envs <= 0) return GR_ERR_SIZE;
  fill_startup_rand_kernel<<<(num_envs + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(srnd, num_envs, env_id_offset, seed);
  return (int)cudaGetLastError();
}
#endif  // GR_CPU_EMUL
