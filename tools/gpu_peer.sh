#!/usr/bin/env bash
# N GPUs: peer all-reduce check, then PPO training with NCCL / with the peer kernel (same seeds: same curves up to summation order)
set -u
out=gpurun_out; mkdir -p "$out"; tag=${1:-peer}; n=${2:-2}
run="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
timeout 240 $run --master-port 29521 tools/peer_reduce_check.py 2>&1 | tail -n 1 | tee "$out/${tag}_check_n${n}.json"
for f in 0 1 0 1; do
GRACING_PEER_ALLREDUCE=$f timeout 300 $run --master-port 2953$f tools/train.py ppo --num_envs 65536 --iters 10 --fused --kernel_update 2>&1 | tail -n 1 | sed "s/^/n=$n peer=$f /" | tee -a "$out/${tag}_status.txt"
done
