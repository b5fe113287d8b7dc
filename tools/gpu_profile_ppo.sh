#!/usr/bin/env bash
# ncu evidence for the PPO update kernels at BASELINE C5 scale (65,536 envs: 393,216-row mini-batches): launch list of our kernels and one
# --set full capture of the forward+loss kernel and of the weight-gradient kernel (plain run first, same command)
set -u
out=gpurun_out; mkdir -p "$out"; tag=${1:-profppo}
ppo="python tools/train.py ppo --num_envs 65536 --iters 5 --fused --kernel_update"
if timeout 300 $ppo > "$out/${tag}_plain.log" 2>&1; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:actor_backward|adam_|policy_|ppo_|storage_|gae|adv_|racing_|moments" -c 900 --csv --log-file "$out/${tag}_launches.csv" $ppo > "$out/${tag}_ncu_list.log" 2>&1
  echo "ncu list: exit $?" | tee -a "$out/${tag}_status.txt"
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:actor_backward|policy_forward" -s 50 -c 2 -f -o "$out/${tag}_update" $ppo > "$out/${tag}_ncu_full.log" 2>&1
  echo "ncu full: exit $?" | tee -a "$out/${tag}_status.txt"
fi
tail -n 2 "$out/${tag}_plain.log"
