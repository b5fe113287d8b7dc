"""B200-native racing hot path of GeneralizableRacing (DiffLab): fused sm_100a env-step / BPTT / rollout kernels behind
the reference's env, storage and runner interfaces.  See DESIGN.md."""
from .config import RacingCfg, ReachTargetCfg
from .tracks import GateTable, figure_eight_track, synthetic_track_table

__all__ = ["RacingCfg", "GateTable", "figure_eight_track", "synthetic_track_table", "RacingVecEnv", "RolloutStorage", "make_env", "ReachTargetCfg", "ReachTargetVecEnv", "make_reach_env"]


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
    raise AttributeError(name)
