#!/usr/bin/env bash
# One-GPU check-up of the tree, one box:
#   gpurun --timeout 1500 -- 'bash tools/first_gpu_call.sh r2'
# 1. the whole -m gpu suite, 2. smoke(), 3. the bench line as the driver runs it (--steps 20 --warmup 5) and the reference arm,
# 4. the default bench line, 5. the ncu launch list of the driver's bench command (only after the bench exited 0 without ncu).
# Everything lands in gpurun_out/; copy what is to be judged into profiles/.
set -u
out=gpurun_out
mkdir -p "$out"
tag=${1:-r2}

timeout 1200 python -m pytest tests -m gpu -q -rA --durations=15 > "$out/${tag}_pytest_gpu.log" 2>&1
echo "pytest -m gpu: exit $?" | tee "$out/${tag}_first_call_status.txt"

timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$out/${tag}_smoke.log" 2>&1
echo "smoke: exit $?" | tee -a "$out/${tag}_first_call_status.txt"

timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > "$out/${tag}_bench_k20.json" 2> "$out/${tag}_bench_k20.err"
rc=$?
echo "bench --steps 20 --warmup 5: exit $rc" | tee -a "$out/${tag}_first_call_status.txt"

timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > "$out/${tag}_bench_reference.json" 2> "$out/${tag}_bench_reference.err"
echo "bench --impl reference: exit $?" | tee -a "$out/${tag}_first_call_status.txt"

timeout 600 python bench.py --no-cpu > "$out/${tag}_bench_default.json" 2> "$out/${tag}_bench_default.err"
echo "bench (default K): exit $?" | tee -a "$out/${tag}_first_call_status.txt"

if [ $rc -eq 0 ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
      --log-file "$out/${tag}_launches.csv" python bench.py --steps 20 --warmup 5 --repeats 3 --no-cpu --no-extra > "$out/${tag}_ncu.log" 2>&1
  echo "ncu launch list: exit $?" | tee -a "$out/${tag}_first_call_status.txt"
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv \
      --log-file "$out/${tag}_smoke_launches.csv" python -c "import __graft_entry__ as g; g.smoke()" > "$out/${tag}_ncu_smoke.log" 2>&1
  echo "ncu smoke launch list: exit $?" | tee -a "$out/${tag}_first_call_status.txt"
fi
tail -3 "$out/${tag}_pytest_gpu.log"
tail -2 "$out/${tag}_smoke.log"
head -c 3000 "$out/${tag}_bench_k20.json"
