#!/usr/bin/env bash
# round-2 check-up + profile: suite, smoke, bench as the driver runs it, then (plain run first, same command) the ncu launch list and
# ONE --set full capture of the step kernel and of the window / sweep kernels from the same bench command
set -u
out=gpurun_out; mkdir -p "$out"; tag=${1:-prof}
timeout 1200 python -m pytest tests -m gpu -q -x --durations=8 > "$out/${tag}_pytest_gpu.log" 2>&1
echo "pytest -m gpu: exit $?" | tee "$out/${tag}_status.txt"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$out/${tag}_smoke.log" 2>&1
echo "smoke: exit $?" | tee -a "$out/${tag}_status.txt"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > "$out/${tag}_bench_k20.json" 2> "$out/${tag}_bench_k20.err"
echo "bench k20: exit $?" | tee -a "$out/${tag}_status.txt"
cmd="python bench.py --steps 20 --warmup 5 --repeats 3 --no-cpu --no-extra --no-collective"
if timeout 300 $cmd > "$out/${tag}_plain.log" 2>&1; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file "$out/${tag}_launches.csv" $cmd > "$out/${tag}_ncu_list.log" 2>&1
  echo "ncu launch list: exit $?" | tee -a "$out/${tag}_status.txt"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:racing_step_fwd -s 60 -c 3 -f -o "$out/${tag}_step" $cmd > "$out/${tag}_ncu_full.log" 2>&1
  echo "ncu full step: exit $?" | tee -a "$out/${tag}_status.txt"
  if ENVS=65536 timeout 300 python tools/collect_bench.py > "$out/${tag}_collect_plain.json" 2>&1; then
    ENVS=65536 timeout 600 ncu --set full --clock-control none --import-source on -k regex:ppo_collect -s 2 -c 1 -f -o "$out/${tag}_collect" python tools/collect_bench.py > "$out/${tag}_ncu_collect.log" 2>&1
    echo "ncu full collect: exit $?" | tee -a "$out/${tag}_status.txt"
  fi
else
  echo "plain profile command failed" | tee -a "$out/${tag}_status.txt"
fi
tail -4 "$out/${tag}_pytest_gpu.log"; tail -2 "$out/${tag}_smoke.log"
