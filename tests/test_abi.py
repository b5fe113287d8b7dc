"""The C-ABI library loads here (no GPU) and exports every symbol include/gracing.h declares; ctypes struct layouts
match the header; argument errors are reported without launching (no compute calls without a GPU)."""
import ctypes as C
import os
import re

import pytest

from generalizableracing_b200 import _lib as B
from generalizableracing_b200 import build as BLD
from generalizableracing_b200 import layout as L_

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "gracing.h")).read()


@pytest.fixture(scope="module")
def lib():
    BLD.build()
    return B.load()


def test_every_declared_symbol_is_exported_and_bound(lib):
    declared = set(re.findall(r"^\s*(?:int|int64_t)\s+(gr_\w+)\s*\(", HEADER, flags=re.M))
    assert len(declared) >= 13
    assert declared == set(B.PROTOTYPES), declared ^ set(B.PROTOTYPES)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.gr_abi_version() == int(re.search(r"#define GR_ABI_VERSION (\d+)", HEADER).group(1))


def test_layout_constants_match_header():
    for macro, val in [("GR_OBS_DIM", L_.OBS_DIM), ("GR_NUM_ACTIONS", L_.NUM_ACTIONS), ("GR_NUM_REWARD_TERMS", L_.NUM_REWARD_TERMS),
                       ("GR_RND_STRIDE", L_.RND_STRIDE), ("GR_SRND_STRIDE", L_.SRND_STRIDE), ("GR_NUM_PLANES", L_.NUM_PLANES),
                       ("GR_NUM_PLANES_WITH_STATS", L_.NUM_PLANES_WITH_STATS), ("GR_TAPE_PLANES", L_.TAPE_PLANES),
                       ("GR_TILE_PLANES", L_.TILE_PLANES), ("GR_LOG_SLOTS", B.GR_LOG_SLOTS), ("GR_LOG_SHARDS", B.GR_LOG_SHARDS),
                       ("GR_LAUNCH_PDL", B.GR_LAUNCH_PDL)]:
        assert int(re.search(rf"#define {macro} (\d+)", HEADER).group(1)) == val, macro


def _c_struct_fields(name):
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), HEADER, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(",")[0:1] + [p for p in decl.split(",")[1:]]:
            m = re.search(r"(\w+)\s*(?:\[[^\]]*\])?\s*$", part.strip())
            fields.append(m.group(1))
    return fields


@pytest.mark.parametrize("name", ["GrConfig", "GrTrack", "GrState", "GrRandom", "GrStepIO", "GrBwdIO", "GrTransition", "GrStorage", "GrMiniBatch", "GrHostStep", "GrMlp", "GrPolicy", "GrCollectIO", "GrBpttCollectIO", "GrMlpGrad", "GrPpoBatch", "GrAdamStep", "GrBackwardJob", "GrPeerReduce", "GrReachConfig", "GrReachState", "GrReachStepIO"])
def test_ctypes_structs_follow_the_header(name):
    assert [f for f, _ in getattr(B, name)._fields_] == _c_struct_fields(name)


def test_argument_errors_are_reported_without_launch(lib):
    assert lib.gr_step_fwd(None, None, None, None, None, None) == -1                      # GR_ERR_NULL
    assert lib.gr_fill_rand(None, 4, 0, 0, 0, None) == -1
    assert lib.gr_fill_rand(C.c_void_p(16), 0, 0, 0, 0, None) == -2                        # GR_ERR_SIZE
    assert lib.gr_fill_rand(C.c_void_p(8), 4, 0, 0, 0, None) == -3                         # GR_ERR_ALIGN
    assert lib.gr_gae_scratch_bytes(4096) >= 3 * 8 * (4096 // 128)
    st = B.GrStorage()
    assert lib.gr_compute_returns(C.byref(st), None, 0.99, 0.95, None, None, 1, None) == -1
    # fused collection / update entry points: argument errors come back before anything is launched
    assert lib.gr_ppo_collect(None, None, None, None, None, None, None, None) == -1
    assert lib.gr_bptt_collect(None, None, None, None, None, 256, 128, None, None) == -1
    assert lib.gr_actor_backward(None, 256, 128, None, None, None, 128, None, None) == -1
    assert lib.gr_policy_forward(None, None, None, None, None, 128, None) == -1
    assert lib.gr_ppo_loss_grad(None, 128, None, None, None, None) == -1
    assert lib.gr_policy_pack(None, None, None, None) == -1
    assert lib.gr_adam_clip_step(None, None) == -1
    assert lib.gr_actor_backward_jobs(None, 2, 128, 128, 128, None) == -1
    assert lib.gr_policy_packed_bytes(128, 128, 2) == 2 * 45440 and lib.gr_policy_packed_bytes(256, 128, 1) == 86400
    assert lib.gr_policy_packed_bytes(64, 64, 1) < 0 and lib.gr_policy_packed_bytes(128, 128, 3) < 0
    pol, grads = B.GrPolicy(16, 16, 0.01), B.GrMlpGrad(16, 16, 16, 16, 16, 16, 4, 0)
    assert lib.gr_actor_backward(C.byref(pol), 64, 128, C.c_void_p(16), C.c_void_p(16), C.c_void_p(16), 128, C.byref(grads), None) == -2      # widths
    assert lib.gr_actor_backward(C.byref(pol), 256, 128, C.c_void_p(16), C.c_void_p(16), C.c_void_p(16), 0, C.byref(grads), None) == -2       # rows
    assert lib.gr_actor_backward(C.byref(pol), 256, 128, C.c_void_p(8), C.c_void_p(16), C.c_void_p(16), 128, C.byref(grads), None) == -3      # alignment
    # reach-target entry points
    assert lib.gr_reach_step_fwd(None, None, None, None, None) == -1
    assert lib.gr_reach_reset(None, None, None, None, None, None) == -1
    assert lib.gr_reach_observe(None, None, None, None) == -1
    assert lib.gr_reach_step_bwd(None, None, None, None) == -1
    assert lib.gr_reach_fill_rand(None, 4, 0, 0, 0, None) == -1
    assert lib.gr_reach_fill_rand(C.c_void_p(16), 0, 0, 0, 0, None) == -2
    assert lib.gr_reach_fill_rand(C.c_void_p(8), 4, 0, 0, 0, None) == -3
    from generalizableracing_b200.config import ReachTargetCfg
    from generalizableracing_b200.reach_env import make_gr_reach_config
    rcfg, rst, rng = make_gr_reach_config(ReachTargetCfg.lv()), B.GrReachState(16, 32, 8, 0), B.GrRandom(None, 0, 0)
    assert lib.gr_reach_reset(C.byref(rcfg), C.byref(B.GrReachState(16, 0, 8, 0)), C.byref(rng), None, None, None) == -2     # stride < envs
    assert lib.gr_reach_reset(C.byref(rcfg), C.byref(B.GrReachState(8, 32, 8, 0)), C.byref(rng), None, None, None) == -3     # alignment
    rcfg.controller = 7
    assert lib.gr_reach_reset(C.byref(rcfg), C.byref(rst), C.byref(rng), None, None, None) == -4                               # GR_ERR_CONFIG
    # multi-step window entry points
    assert lib.gr_rollout_fwd(None, None, None, None, None, None) == -1
    assert lib.gr_reach_rollout_fwd(None, None, None, None, None) == -1
    rcfg = make_gr_reach_config(ReachTargetCfg.lv())
    rio = B.GrReachRolloutIO()
    assert lib.gr_reach_rollout_fwd(C.byref(rcfg), C.byref(rst), C.byref(rng), C.byref(rio), None) == -1                     # no actions / obs_out
    rio.actions, rio.obs_out, rio.T = 16, 16, 0
    assert lib.gr_reach_rollout_fwd(C.byref(rcfg), C.byref(rst), C.byref(rng), C.byref(rio), None) == -2                     # T < 1
    rio.T, rio.actions = 4, 8
    assert lib.gr_reach_rollout_fwd(C.byref(rcfg), C.byref(rst), C.byref(rng), C.byref(rio), None) == -3                     # alignment
    rio.actions, rio.tape, rio.tape_stride = 16, 16, 8
    assert lib.gr_reach_rollout_fwd(C.byref(rcfg), C.byref(rst), C.byref(rng), C.byref(rio), None) == -2                     # tape stride < envs
    from generalizableracing_b200.config import RacingCfg
    from generalizableracing_b200.env import make_gr_config
    gcfg, trk = make_gr_config(RacingCfg.for_stage(1)), B.GrTrack(16, 2, 3, 4)
    gst = B.GrState(16, 32, 8, L_.NUM_PLANES, 0, 1, 0, 0, 16)
    io = B.GrRolloutIO()
    assert lib.gr_rollout_fwd(C.byref(gcfg), C.byref(trk), C.byref(gst), C.byref(rng), C.byref(io), None) == -1              # no actions / obs_out
    io.actions, io.obs_out, io.T = 16, 16, 0
    assert lib.gr_rollout_fwd(C.byref(gcfg), C.byref(trk), C.byref(gst), C.byref(rng), C.byref(io), None) == -2              # T < 1
    io.T, io.obs_out = 3, 8
    assert lib.gr_rollout_fwd(C.byref(gcfg), C.byref(trk), C.byref(gst), C.byref(rng), C.byref(io), None) == -3              # alignment
    io.obs_out, io.tape, io.tape_stride = 16, 16, 40
    assert lib.gr_rollout_fwd(C.byref(gcfg), C.byref(trk), C.byref(gst), C.byref(rng), C.byref(io), None) == -2              # tape stride not a tile multiple
    pipe = C.c_void_p()
    assert lib.gr_host_pipe_create(64, 2, None, None) == -1
    assert lib.gr_host_pipe_create(0, 2, None, C.byref(pipe)) == -2
    assert lib.gr_host_pipe_create(64, B.GR_HOST_PIPE_MAX_DEPTH + 1, None, C.byref(pipe)) == -2
    assert lib.gr_host_pipe_step(None, None, None, None, None, None, None, None) == -1
    assert lib.gr_host_pipe_wait(None, 0) == -1
    with pytest.raises(B.GracingError):
        B.check(-5, "x")


def test_product_refuses_cpu():
    from generalizableracing_b200.config import RacingCfg
    from generalizableracing_b200.env import RacingVecEnv
    from generalizableracing_b200.storage import RolloutStorage
    from generalizableracing_b200.tracks import figure_eight_track
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        RacingVecEnv(RacingCfg.for_stage(0), figure_eight_track(), 8, device="cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        RolloutStorage("rl", 8, 4, [16], [16], [4], device="cpu")
    from generalizableracing_b200.config import ReachTargetCfg
    from generalizableracing_b200.reach_env import ReachTargetVecEnv
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ReachTargetVecEnv(ReachTargetCfg.lv(), 8, device="cpu")
