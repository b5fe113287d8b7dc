#!/usr/bin/env bash
set -u
out=gpurun_out; mkdir -p "$out"; tag=${1:-r2s}
ppo="python tools/train.py ppo --num_envs 65536 --iters 4 --fused --kernel_update"
for f in 1 0; do
export GRACING_PPO_FUSED_STEP=$f
if timeout 300 $ppo > "$out/${tag}_plain$f.log" 2>&1; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:actor_backward -s 25 -c 1 -f -o "$out/${tag}_ab$f" $ppo > "$out/${tag}_ncu$f.log" 2>&1
  echo "ncu fused=$f: exit $?" | tee -a "$out/${tag}_status.txt"
fi
done
