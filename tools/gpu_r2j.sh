#!/usr/bin/env bash
# multi-GPU leg: bench.py under torchrun exactly as the driver launches it (N ranks), then the reference arm under torchrun
set -u
out=gpurun_out; mkdir -p "$out"; tag=${1:-r2j}; n=${2:-2}
nvidia-smi topo -m > "$out/${tag}_topo.txt" 2>&1
for rep in 1 2; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 > "$out/${tag}_bench_n${n}_$rep.json" 2> "$out/${tag}_bench_n${n}_$rep.err"
echo "bench n=$n rep $rep: exit $?" | tee -a "$out/${tag}_status.txt"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $n --steps 20 --warmup 5 > "$out/${tag}_bench_ref_n${n}.json" 2> "$out/${tag}_bench_ref_n${n}.err"
echo "bench ref n=$n: exit $?" | tee -a "$out/${tag}_status.txt"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extra --no-cpu > "$out/${tag}_bench_n1.json" 2> "$out/${tag}_bench_n1.err"
echo "bench n=1: exit $?" | tee -a "$out/${tag}_status.txt"
tail -c 600 "$out/${tag}_bench_n${n}_2.err"
