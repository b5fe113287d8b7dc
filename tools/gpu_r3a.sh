#!/usr/bin/env bash
# round-2 late session, call A: dense (pre-permuted) transition records against gathered ones; step kernel with byte dones / 32-thread blocks
set -u
out=gpurun_out; mkdir -p "$out"; tag=${1:-r4a}
timeout 600 python -m pytest tests/test_ppo_update_kernels.py tests/test_ppo_graphed_update.py tests/test_runners_gpu.py -m gpu -q -x > "$out/${tag}_pytest.log" 2>&1
echo "pytest: exit $?" | tee "$out/${tag}_status.txt"
tail -n 3 "$out/${tag}_pytest.log"
for d in 1 0 1 0; do
for n in 65536; do
GRACING_PPO_DENSE_RECORDS=$d timeout 300 python tools/train.py ppo --num_envs $n --iters 8 --fused --kernel_update 2>&1 | tail -n 1 | sed "s/^/n=$n dense_records=$d /" | tee -a "$out/${tag}_status.txt"
done; done
for d in 1 0; do
GRACING_PPO_DENSE_RECORDS=$d timeout 300 python tools/train.py ppo --num_envs 4096 --iters 12 --fused --kernel_update 2>&1 | tail -n 1 | sed "s/^/n=4096 dense_records=$d /" | tee -a "$out/${tag}_status.txt"
done
for dn in i64 u8 none i64 u8; do
timeout 300 python tools/step_timing.py --blocks 64 --complex 1 --dones $dn 2>&1 | tail -n 1 | tee -a "$out/${tag}_step.txt"
done
timeout 300 python tools/step_timing.py --blocks 32,96 --complex 1 --dones u8 2>&1 | tail -n 2 | tee -a "$out/${tag}_step.txt"
