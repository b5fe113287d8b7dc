#!/usr/bin/env bash
# launch lists of whole training iterations at scale: PPO at 65,536 envs (fused collection + kernel update), BPTT at C3 (fused window + kernel backward)
set -u
out=gpurun_out; mkdir -p "$out"; tag=${1:-r2l}
ppo="python tools/train.py ppo --num_envs 65536 --iters 4 --fused --kernel_update"
bptt="python tools/train.py bptt --num_envs 16384 --iters 6 --fused --fused_backward"
if timeout 300 $ppo > "$out/${tag}_ppo_plain.log" 2>&1; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 700 --csv --log-file "$out/${tag}_ppo_launches.csv" $ppo > "$out/${tag}_ppo_ncu.log" 2>&1
  echo "ncu ppo: exit $?" | tee -a "$out/${tag}_status.txt"
else echo "ppo plain failed" | tee -a "$out/${tag}_status.txt"; tail -5 "$out/${tag}_ppo_plain.log"; fi
if timeout 300 $bptt > "$out/${tag}_bptt_plain.log" 2>&1; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 300 --csv --log-file "$out/${tag}_bptt_launches.csv" $bptt > "$out/${tag}_bptt_ncu.log" 2>&1
  echo "ncu bptt: exit $?" | tee -a "$out/${tag}_status.txt"
else echo "bptt plain failed" | tee -a "$out/${tag}_status.txt"; tail -5 "$out/${tag}_bptt_plain.log"; fi
tail -3 "$out/${tag}_ppo_plain.log" "$out/${tag}_bptt_plain.log"
