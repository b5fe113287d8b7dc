from .on_policy_runner import OnPolicyRunner
from .algo_runner import AlgoRunner

__all__ = ["OnPolicyRunner", "AlgoRunner"]
