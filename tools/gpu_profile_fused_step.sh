#!/usr/bin/env bash
# the one-launch PPO step with dense records at 65,536 envs: plain run, ncu launch list, one --set full capture with source
set -u
out=gpurun_out; mkdir -p "$out"; tag=${1:-fusedprof}
ppo="python tools/train.py ppo --num_envs 65536 --iters 4 --fused --kernel_update"
if timeout 300 $ppo > "$out/${tag}_plain.log" 2>&1; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:actor_backward|adam_|policy_|ppo_|storage_|gae" -c 400 --csv --log-file "$out/${tag}_launches.csv" $ppo > "$out/${tag}_ncu_list.log" 2>&1
  echo "ncu list: exit $?" | tee -a "$out/${tag}_status.txt"
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:actor_backward" -s 30 -c 1 -f -o "$out/${tag}_fused" $ppo > "$out/${tag}_ncu_full.log" 2>&1
  echo "ncu full: exit $?" | tee -a "$out/${tag}_status.txt"
fi
tail -n 1 "$out/${tag}_plain.log"
