#!/usr/bin/env bash
# the driver's scaling leg at N GPUs: bench.py under torchrun (our arm, then the reference arm), one run each
set -u
out=gpurun_out; mkdir -p "$out"; tag=${1:-scale}; n=${2:-8}
nvidia-smi topo -m > "$out/${tag}_topo.txt" 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 > "$out/${tag}_bench_n${n}.json" 2> "$out/${tag}_bench_n${n}.err"
echo "bench n=$n: exit $?" | tee -a "$out/${tag}_status.txt"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $n --steps 20 --warmup 5 > "$out/${tag}_bench_ref_n${n}.json" 2> "$out/${tag}_bench_ref_n${n}.err"
echo "bench ref n=$n: exit $?" | tee -a "$out/${tag}_status.txt"
tail -c 400 "$out/${tag}_bench_n${n}.err"
