"""TEST INFRASTRUCTURE: CPU emulation of the step/reset/startup/bwd kernels and of the rollout-storage kernels (same .cu sources, g++).

Lets the kernel logic be debugged against the oracle without a GPU.  Only tests import this; the product
(`RacingVecEnv` without the private ``_lib`` argument) loads libgracing.so and refuses non-CUDA devices.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

from generalizableracing_b200 import _lib as B

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgracing_emul.so")
_CSRC = os.path.join(_HERE, "..", "..", "generalizableracing_b200", "csrc")


def _stale():
    if not os.path.isfile(_SO):
        return True
    t = os.path.getmtime(_SO)
    deps = [os.path.join(_HERE, f) for f in ("emul.cpp", "cuda_shim.h")] + \
           [os.path.join(_CSRC, f) for f in ("racing_step.cu", "racing_bwd.cu", "gr_math.cuh", "gr_common.cuh", "reach_step.cu", "reach_bwd.cu", "reach_core.cuh", "racing_step_core.cuh", "rollout.cu", "mesh_collision.cu")] + \
           [os.path.join(_HERE, "..", "..", "include", "gracing.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build():
    if _stale():
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", _SO,
                        os.path.join(_HERE, "emul.cpp")], check=True)
    return _SO


class EmulLib:
    """Duck-types the subset of libgracing.so that RacingVecEnv / BpttWindow call (stream argument ignored)."""

    def __init__(self):
        self._l = C.CDLL(build())
        P = C.POINTER
        self._l.emul_step_fwd.argtypes = [P(B.GrConfig), P(B.GrTrack), P(B.GrState), P(B.GrRandom), P(B.GrStepIO)]
        self._l.emul_rollout_fwd.argtypes = [P(B.GrConfig), P(B.GrTrack), P(B.GrState), P(B.GrRandom), P(B.GrRolloutIO)]
        self._l.emul_env_reset.argtypes = [P(B.GrConfig), P(B.GrTrack), P(B.GrState), P(B.GrRandom), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        self._l.emul_env_startup.argtypes = [P(B.GrConfig), P(B.GrTrack), P(B.GrState), C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        self._l.emul_step_bwd.argtypes = [P(B.GrConfig), P(B.GrState), P(B.GrBwdIO)]
        self._l.emul_fill_rand.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_uint64, C.c_uint32]
        self._l.emul_fill_startup_rand.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_uint64]
        self._l.emul_reach_step_fwd.argtypes = [P(B.GrReachConfig), P(B.GrReachState), P(B.GrRandom), P(B.GrReachStepIO)]
        self._l.emul_reach_rollout_fwd.argtypes = [P(B.GrReachConfig), P(B.GrReachState), P(B.GrRandom), P(B.GrReachRolloutIO)]
        self._l.emul_reach_reset.argtypes = [P(B.GrReachConfig), P(B.GrReachState), P(B.GrRandom), C.c_void_p, C.c_int, C.c_void_p]
        self._l.emul_reach_step_bwd.argtypes = [P(B.GrReachConfig), P(B.GrReachState), P(B.GrBwdIO)]
        self._l.emul_reach_fill_rand.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_uint64, C.c_uint32]
        self._l.emul_storage_add.argtypes = [P(B.GrStorage), P(B.GrTransition), C.c_int32]
        self._l.emul_compute_returns.argtypes = [P(B.GrStorage), C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_int32]
        self._l.emul_advantage_normalize.argtypes = [P(B.GrStorage), C.c_void_p]
        self._l.emul_storage_gather.argtypes = [P(B.GrStorage), C.c_void_p, C.c_int32, P(B.GrMiniBatch)]
        self._l.emul_storage_pack_records.argtypes = [P(B.GrStorage), C.c_void_p, C.c_int64, C.c_void_p]
        self._l.emul_uav_collision_ray.argtypes = [P(B.GrMesh), C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_void_p]
        self._l.emul_mesh_query_rays.argtypes = [P(B.GrMesh), C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_void_p]

    def gr_step_fwd(self, cfg, tr, st, rng, io, stream):
        return self._l.emul_step_fwd(cfg, tr, st, rng, io)

    def gr_rollout_fwd(self, cfg, tr, st, rng, io, stream):
        return self._l.emul_rollout_fwd(cfg, tr, st, rng, io)

    def gr_env_reset(self, cfg, tr, st, rng, mask, obs, critic, aux, stream):
        return self._l.emul_env_reset(cfg, tr, st, rng, mask, 0 if mask else 1, obs, critic, aux)

    def gr_env_observe(self, cfg, tr, st, rng, obs, critic, aux, stream):
        return self._l.emul_env_reset(cfg, tr, st, rng, None, 2, obs, critic, aux)

    def gr_env_startup(self, cfg, tr, st, types, chunk, srnd, seed, stream):
        return self._l.emul_env_startup(cfg, tr, st, types, chunk, srnd, seed)

    def gr_step_bwd(self, cfg, st, io, stream):
        return self._l.emul_step_bwd(cfg, st, io)

    def gr_fill_rand(self, rnd, n, off, seed, step, stream):
        return self._l.emul_fill_rand(rnd, n, off, seed, step)

    def gr_fill_startup_rand(self, srnd, n, off, seed, stream):
        return self._l.emul_fill_startup_rand(srnd, n, off, seed)

    def gr_reach_step_fwd(self, cfg, st, rng, io, stream):
        return self._l.emul_reach_step_fwd(cfg, st, rng, io)

    def gr_reach_rollout_fwd(self, cfg, st, rng, io, stream):
        return self._l.emul_reach_rollout_fwd(cfg, st, rng, io)

    def gr_reach_reset(self, cfg, st, rng, mask, obs, stream):
        return self._l.emul_reach_reset(cfg, st, rng, mask, 0 if mask else 1, obs)

    def gr_reach_observe(self, cfg, st, obs, stream):
        return self._l.emul_reach_reset(cfg, st, None, None, 2, obs)

    def gr_reach_step_bwd(self, cfg, st, io, stream):
        return self._l.emul_reach_step_bwd(cfg, st, io)

    def gr_reach_fill_rand(self, rnd, n, off, seed, step, stream):
        return self._l.emul_reach_fill_rand(rnd, n, off, seed, step)

    # ---- rollout storage (csrc/rollout.cu)
    def gr_gae_scratch_bytes(self, n):
        return ((int(n) + 127) // 128 * 3 + 3) * 8

    def gr_storage_add(self, s, tr, step, stream):
        return self._l.emul_storage_add(s, tr, step)

    def gr_compute_returns(self, s, last_values, gamma, lam, scratch, moments, normalize, stream):
        return self._l.emul_compute_returns(s, last_values, gamma, lam, scratch, moments, normalize)

    def gr_advantage_normalize(self, s, moments, stream):
        return self._l.emul_advantage_normalize(s, moments)

    def gr_storage_gather(self, s, idx, b, out, stream):
        return self._l.emul_storage_gather(s, idx, b, out)

    def gr_storage_pack_records(self, s, records, stream):
        return self._l.emul_storage_pack_records(s, None, 0, records)

    def gr_storage_pack_records_permuted(self, s, perm, num, records, stream):
        return self._l.emul_storage_pack_records(s, perm, num, records)

    def gr_uav_collision_ray(self, mesh, pos, quat, n, lattices, num_lattices, max_dist, arm, height, out, stream):
        return self._l.emul_uav_collision_ray(mesh, pos, quat, n, lattices, num_lattices, max_dist, arm, height, out)

    def gr_mesh_query_rays(self, mesh, origins, dirs, n, max_t, t_out, sign_out, stream):
        return self._l.emul_mesh_query_rays(mesh, origins, dirs, n, max_t, t_out, sign_out)
