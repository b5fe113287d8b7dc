"""B200-native racing hot path of GeneralizableRacing (DiffLab): fused sm_100a env-step / BPTT / rollout kernels behind
the reference's env, storage and runner interfaces.  See DESIGN.md."""
from .config import RacingCfg, ReachTargetCfg
from .tracks import GateTable, figure_eight_track, synthetic_track_table

__all__ = ["RacingCfg", "GateTable", "figure_eight_track", "synthetic_track_table", "RacingVecEnv", "RolloutStorage", "make_env", "ReachTargetCfg", "ReachTargetVecEnv", "make_reach_env", "TerrainMesh", "get_uav_collision_num_ray"]


def __getattr__(name):
    if name == "RacingVecEnv":
        from .env import RacingVecEnv
        return RacingVecEnv
    if name == "RolloutStorage":
        from .storage import RolloutStorage
        return RolloutStorage
    if name == "make_env":
        from .env import make_env
        return make_env
    if name in ("ReachTargetVecEnv", "make_reach_env"):
        from . import reach_env
        return getattr(reach_env, name)
    if name in ("TerrainMesh", "get_uav_collision_num_ray", "collision_penalty_custom", "LATTICE_TENSOR"):
        from . import mesh
        return getattr(mesh, name)
    raise AttributeError(name)
