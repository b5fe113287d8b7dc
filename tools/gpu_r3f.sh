#!/usr/bin/env bash
# PPO update kernels after a change: their tests, then whole-iteration timing at 65,536 / 4,096 envs (twice each)
set -u
out=gpurun_out; mkdir -p "$out"; tag=${1:-r4f}
timeout 900 python -m pytest tests/test_ppo_update_kernels.py tests/test_ppo_graphed_update.py tests/test_actor_backward.py tests/test_runners_gpu.py tests/test_training_parity.py -m gpu -q -x > "$out/${tag}_pytest.log" 2>&1
echo "pytest: exit $?" | tee "$out/${tag}_status.txt"
tail -n 3 "$out/${tag}_pytest.log"
for rep in 1 2; do
for n in 65536 4096; do
timeout 300 python tools/train.py ppo --num_envs $n --iters 10 --fused --kernel_update 2>&1 | tail -n 1 | sed "s/^/n=$n /" | tee -a "$out/${tag}_status.txt"
done; done
