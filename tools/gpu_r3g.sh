#!/usr/bin/env bash
# forward kernel: interleaved actor / critic chains A/B -- tests with both, whole-iteration timing, ncu launch lists
set -u
out=gpurun_out; mkdir -p "$out"; tag=${1:-r4g}
GRACING_FWD_INTERLEAVE=1 timeout 600 python -m pytest tests/test_ppo_update_kernels.py tests/test_ppo_graphed_update.py -m gpu -q -x > "$out/${tag}_pytest_inter.log" 2>&1
echo "pytest (interleaved): exit $?" | tee "$out/${tag}_status.txt"
timeout 600 python -m pytest tests/test_ppo_update_kernels.py tests/test_ppo_graphed_update.py -m gpu -q -x > "$out/${tag}_pytest.log" 2>&1
echo "pytest: exit $?" | tee -a "$out/${tag}_status.txt"
for rep in 1 2; do for f in 0 1; do
GRACING_FWD_INTERLEAVE=$f timeout 300 python tools/train.py ppo --num_envs 65536 --iters 10 --fused --kernel_update 2>&1 | tail -n 1 | sed "s/^/interleave=$f /" | tee -a "$out/${tag}_status.txt"
done; done
ppo="python tools/train.py ppo --num_envs 65536 --iters 4 --fused --kernel_update"
for f in 0 1; do
GRACING_FWD_INTERLEAVE=$f timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:actor_backward|adam_|policy_|ppo_|storage_|gae" -c 500 --csv --log-file "$out/${tag}_launches_inter$f.csv" $ppo > "$out/${tag}_ncu_list$f.log" 2>&1
echo "ncu list $f: exit $?" | tee -a "$out/${tag}_status.txt"
done
