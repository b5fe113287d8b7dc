#!/usr/bin/env bash
set -u
out=gpurun_out; mkdir -p "$out"; tag=${1:-r2m}
ppo="python tools/train.py ppo --num_envs 65536 --iters 5 --fused --kernel_update"
if timeout 300 $ppo > "$out/${tag}_ppo_plain.log" 2>&1; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:actor_backward|adam_|policy_|ppo_|storage_|gae|adv_|racing_|moments" -c 900 --csv --log-file "$out/${tag}_ppo_launches.csv" $ppo > "$out/${tag}_ppo_ncu.log" 2>&1
  echo "ncu ppo: exit $?" | tee -a "$out/${tag}_status.txt"
else echo "ppo plain failed" | tee -a "$out/${tag}_status.txt"; fi
tail -n 3 "$out/${tag}_ppo_plain.log"
