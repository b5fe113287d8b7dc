#!/usr/bin/env bash
# First GPU call of a round, one box, one GPU:
#   gpurun --timeout 1500 -- 'bash tools/first_gpu_call.sh'
# 1. the whole -m gpu suite (5 cuda variants written after round 1's GPU budget was spent have not run on a B200 yet),
# 2. smoke(), 3. the default bench line and the reference arm, 4. the ncu launch list of the same bench command
# (only after the bench exited 0 without ncu).  Everything lands in gpurun_out/; copy what is to be judged into profiles/.
set -u
out=gpurun_out
mkdir -p "$out"
tag=${1:-r2}

timeout 900 python -m pytest tests -m gpu -q -rA --durations=15 > "$out/${tag}_pytest_gpu.log" 2>&1
echo "pytest -m gpu: exit $?" | tee "$out/${tag}_first_call_status.txt"

timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$out/${tag}_smoke.log" 2>&1
echo "smoke: exit $?" | tee -a "$out/${tag}_first_call_status.txt"

timeout 600 python bench.py > "$out/${tag}_bench.json" 2> "$out/${tag}_bench.err"
rc=$?
echo "bench: exit $rc" | tee -a "$out/${tag}_first_call_status.txt"

timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > "$out/${tag}_bench_reference.json" 2> "$out/${tag}_bench_reference.err"
echo "bench --impl reference: exit $?" | tee -a "$out/${tag}_first_call_status.txt"

if [ $rc -eq 0 ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file "$out/${tag}_launches.csv" python bench.py --steps 22 --warmup 3 --no-cpu --no-extra > "$out/${tag}_ncu.log" 2>&1
  echo "ncu launch list: exit $?" | tee -a "$out/${tag}_first_call_status.txt"
fi
tail -3 "$out/${tag}_pytest_gpu.log"
cat "$out/${tag}_bench.json"
