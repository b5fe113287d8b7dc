from .ppo import PPO
from .bptt import BPTT

__all__ = ["PPO", "BPTT"]
