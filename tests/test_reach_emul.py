"""Reach-target tasks, kernel logic on the CPU: the very .cu sources of reach_step.cu / reach_bwd.cu compiled by g++ (tests/emul)
against oracle/reach_oracle.py on identical draws -- observations, rewards, masks, state, losses, episode log and the analytic
BPTT gradient against torch.autograd through the oracle, for the three command modes.  The GPU twin: tests/test_reach_parity.py."""
import pytest
import torch

from tests import reach_cases as RC

CASES = ["lv", "ps", "ctbr", "ctbr_sim2real", "lv_literal"]


@pytest.fixture(scope="module")
def emul():
    from tests.emul import EmulLib
    return EmulLib()


@pytest.mark.parametrize("case", CASES)
def test_forward_rollout_matches_oracle(emul, case):
    RC.check_forward(case, num_envs=67, steps=120, device="cpu", lib=emul)


@pytest.mark.parametrize("case", ["lv", "ps", "ctbr"])
def test_bptt_gradient_matches_autograd(emul, case):
    RC.check_bptt(case, num_envs=37, horizon=24, device="cpu", lib=emul)


def test_philox_fill_matches_in_kernel_draws(emul):
    RC.check_philox(num_envs=45, steps=40, device="cpu", lib=emul)


def test_masked_reset_and_observe(emul):
    RC.check_masked_reset(num_envs=40, device="cpu", lib=emul)
