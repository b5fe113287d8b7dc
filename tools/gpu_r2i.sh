#!/usr/bin/env bash
set -u
out=gpurun_out; mkdir -p "$out"; tag=${1:-r2i}
timeout 1200 python -m pytest tests -m gpu -q -x --durations=10 > "$out/${tag}_pytest_gpu.log" 2>&1
echo "pytest -m gpu: exit $?" | tee "$out/${tag}_status.txt"
for k in 1 2 3 4 5 6; do
  timeout 300 python -m pytest tests/test_baseline_sizes.py -m gpu -q -k "dense" > "$out/${tag}_flaky_$k.log" 2>&1
  echo "flaky run $k: exit $?" | tee -a "$out/${tag}_status.txt"
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$out/${tag}_smoke.log" 2>&1
echo "smoke: exit $?" | tee -a "$out/${tag}_status.txt"
for coop in 0 1; do
GRACING_COOP_RESET=$coop timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > "$out/${tag}_bench_k20_coop$coop.json" 2> "$out/${tag}_bench_k20_coop$coop.err"
echo "bench k20 coop$coop: exit $?" | tee -a "$out/${tag}_status.txt"
done
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > "$out/${tag}_bench_reference.json" 2> "$out/${tag}_bench_reference.err"
echo "bench ref: exit $?" | tee -a "$out/${tag}_status.txt"
tail -4 "$out/${tag}_pytest_gpu.log"; tail -2 "$out/${tag}_smoke.log"
