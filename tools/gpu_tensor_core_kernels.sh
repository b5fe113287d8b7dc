#!/usr/bin/env bash
# tcgen05 kernels after a change of the issue path: their tests, collection timing, PPO / BPTT iteration timing
set -u
out=gpurun_out; mkdir -p "$out"; tag=${1:-tc}
timeout 900 python -m pytest tests/test_ppo_collect.py tests/test_bptt_collect.py tests/test_ppo_update_kernels.py tests/test_ppo_graphed_update.py tests/test_actor_backward.py tests/test_runners_gpu.py -m gpu -q -x > "$out/${tag}_pytest.log" 2>&1
echo "pytest: exit $?" | tee "$out/${tag}_status.txt"
tail -n 3 "$out/${tag}_pytest.log"
for rep in 1 2; do
REPS=5 SKIP_EAGER=1 ENVS=65536,16384,4096 TILE_GROUPS=0 timeout 300 python tools/collect_bench.py 2>&1 | grep us_per_step | tr '\n' ' ' | sed "s/^/collect 65536 16384 4096: /" | tee -a "$out/${tag}_status.txt"; echo | tee -a "$out/${tag}_status.txt"
done
timeout 300 python tools/train.py ppo --num_envs 65536 --iters 10 --fused --kernel_update 2>&1 | tail -n 1 | sed "s/^/ppo n=65536 /" | tee -a "$out/${tag}_status.txt"
timeout 300 python tools/train.py ppo --num_envs 4096 --iters 12 --fused --kernel_update 2>&1 | tail -n 1 | sed "s/^/ppo n=4096 /" | tee -a "$out/${tag}_status.txt"
timeout 300 python tools/train.py bptt --num_envs 16384 --iters 40 --fused --fused_backward 2>&1 | tail -n 1 | sed "s/^/bptt n=16384 /" | tee -a "$out/${tag}_status.txt"
