#!/usr/bin/env bash
# update kernels after a change: tests, whole-iteration timing (twice), ncu launch list
set -u
out=gpurun_out; mkdir -p "$out"; tag=${1:-upd}
timeout 900 python -m pytest tests/test_ppo_update_kernels.py tests/test_ppo_graphed_update.py tests/test_actor_backward.py tests/test_bptt_collect.py tests/test_runners_gpu.py -m gpu -q -x > "$out/${tag}_pytest.log" 2>&1
echo "pytest: exit $?" | tee "$out/${tag}_status.txt"
tail -n 3 "$out/${tag}_pytest.log"
for rep in 1 2; do
timeout 300 python tools/train.py ppo --num_envs 65536 --iters 10 --fused --kernel_update 2>&1 | tail -n 1 | sed "s/^/n=65536 /" | tee -a "$out/${tag}_status.txt"
done
timeout 300 python tools/train.py ppo --num_envs 4096 --iters 12 --fused --kernel_update 2>&1 | tail -n 1 | sed "s/^/n=4096 /" | tee -a "$out/${tag}_status.txt"
ppo="python tools/train.py ppo --num_envs 65536 --iters 4 --fused --kernel_update"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:actor_backward|adam_|policy_|ppo_|storage_|gae" -c 500 --csv --log-file "$out/${tag}_launches.csv" $ppo > "$out/${tag}_ncu_list.log" 2>&1
echo "ncu list: exit $?" | tee -a "$out/${tag}_status.txt"
