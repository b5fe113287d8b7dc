#!/usr/bin/env bash
# call C: dense-records test with its run-to-run yardstick; compact cooperative reset draws A/B with longer timing
set -u
out=gpurun_out; mkdir -p "$out"; tag=${1:-r4c}
timeout 900 python -m pytest tests/test_ppo_graphed_update.py tests/test_ppo_collect.py tests/test_bptt_collect.py -m gpu -q -x -s > "$out/${tag}_pytest.log" 2>&1
echo "pytest: exit $?" | tee "$out/${tag}_status.txt"
grep "^it " "$out/${tag}_pytest.log" | tee -a "$out/${tag}_status.txt"
tail -n 3 "$out/${tag}_pytest.log"
for c in -1 0 2 -1 0 2; do
GRACING_COLLECT_COOP_COLUMNS=$c REPS=5 SKIP_EAGER=1 ENVS=65536 TILE_GROUPS=0 timeout 300 python tools/collect_bench.py 2>&1 | grep us_per_step | tr '\n' ' ' | sed "s/^/coop_columns=$c /" | tee -a "$out/${tag}_status.txt"; echo | tee -a "$out/${tag}_status.txt"
done
