#!/usr/bin/env bash
set -u
out=gpurun_out; mkdir -p "$out"; tag=${1:-r2q}
timeout 900 python -m pytest tests/test_actor_backward.py tests/test_ppo_update_kernels.py tests/test_ppo_graphed_update.py tests/test_runners_gpu.py tests/test_training_parity.py -m gpu -q -x > "$out/${tag}_pytest.log" 2>&1
echo "pytest: exit $?" | tee "$out/${tag}_status.txt"
tail -n 6 "$out/${tag}_pytest.log"
for f in 1 0; do
GRACING_PPO_FUSED_STEP=$f timeout 300 python tools/train.py ppo --num_envs 65536 --iters 6 --fused --kernel_update > "$out/${tag}_ppo_f$f.log" 2>&1
echo "ppo fused=$f: exit $?" | tee -a "$out/${tag}_status.txt"; tail -n 2 "$out/${tag}_ppo_f$f.log"
GRACING_PPO_FUSED_STEP=$f timeout 300 python tools/train.py ppo --num_envs 4096 --iters 6 --fused --kernel_update > "$out/${tag}_ppo4k_f$f.log" 2>&1
echo "ppo4k fused=$f: exit $?" | tee -a "$out/${tag}_status.txt"; tail -n 1 "$out/${tag}_ppo4k_f$f.log"
done
