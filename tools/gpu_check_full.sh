#!/usr/bin/env bash
# what the driver runs at round end, on one GPU: the -m gpu suite, smoke(), both bench arms
set -u
out=gpurun_out; mkdir -p "$out"; tag=${1:-chk}
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > "$out/${tag}_pytest_gpu.log" 2>&1
echo "pytest -m gpu: exit $?" | tee "$out/${tag}_status.txt"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$out/${tag}_smoke.log" 2>&1
echo "smoke: exit $?" | tee -a "$out/${tag}_status.txt"
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > "$out/${tag}_bench_reference.json" 2> "$out/${tag}_bench_reference.err"
echo "bench ref: exit $?" | tee -a "$out/${tag}_status.txt"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > "$out/${tag}_bench_k20.json" 2> "$out/${tag}_bench_k20.err"
echo "bench k20: exit $?" | tee -a "$out/${tag}_status.txt"
tail -n 12 "$out/${tag}_pytest_gpu.log"; tail -n 2 "$out/${tag}_smoke.log"; tail -c 300 "$out/${tag}_bench_k20.err"
