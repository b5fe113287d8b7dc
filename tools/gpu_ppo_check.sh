#!/usr/bin/env bash
# PPO update kernels: their GPU tests, then whole-iteration timings at 65,536 and 4,096 envs with the one-launch step forced on / off
set -u
out=gpurun_out; mkdir -p "$out"; tag=${1:-ppochk}
timeout 900 python -m pytest tests/test_actor_backward.py tests/test_ppo_update_kernels.py tests/test_ppo_graphed_update.py tests/test_runners_gpu.py tests/test_training_parity.py -m gpu -q -x > "$out/${tag}_pytest.log" 2>&1
echo "pytest: exit $?" | tee "$out/${tag}_status.txt"
tail -n 3 "$out/${tag}_pytest.log"
for f in 1 0; do
for n in 65536 4096; do
GRACING_PPO_FUSED_STEP=$f timeout 300 python tools/train.py ppo --num_envs $n --iters 6 --fused --kernel_update 2>&1 | tail -n 1 | sed "s/^/n=$n one_launch_step=$f /" | tee -a "$out/${tag}_status.txt"
done; done
