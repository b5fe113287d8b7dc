#!/usr/bin/env bash
# N GPUs (default 8): peer all-reduce check, the bench line (collective block: PPO iteration with the peer kernel, both exchange kernels alone), the
# same PPO iteration with NCCL for comparison
set -u
out=gpurun_out; mkdir -p "$out"; tag=${1:-peer8}; n=${2:-8}
run="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
timeout 240 $run --master-port 29521 tools/peer_reduce_check.py 2>&1 | tail -n 1 | tee "$out/${tag}_check_n${n}.json"
timeout 900 $run --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 > "$out/${tag}_bench_n${n}.json" 2> "$out/${tag}_bench_n${n}.err"
echo "bench n=$n: exit $?" | tee -a "$out/${tag}_status.txt"
for f in 0 1; do
GRACING_PEER_ALLREDUCE=$f timeout 300 $run --master-port 2953$f tools/train.py ppo --num_envs 65536 --iters 10 --fused --kernel_update 2>&1 | tail -n 1 | sed "s/^/n=$n peer=$f /" | tee -a "$out/${tag}_status.txt"
done
tail -c 300 "$out/${tag}_bench_n${n}.err"
