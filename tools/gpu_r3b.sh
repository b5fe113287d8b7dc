#!/usr/bin/env bash
# call B: dense-records tests again; compact cooperative reset draws in the fused collection kernels (tests + A/B timing)
set -u
out=gpurun_out; mkdir -p "$out"; tag=${1:-r4b}
timeout 900 python -m pytest tests/test_ppo_update_kernels.py tests/test_ppo_graphed_update.py tests/test_ppo_collect.py tests/test_bptt_collect.py tests/test_runners_gpu.py tests/test_training_parity.py -m gpu -q -x > "$out/${tag}_pytest.log" 2>&1
echo "pytest: exit $?" | tee "$out/${tag}_status.txt"
tail -n 3 "$out/${tag}_pytest.log"
for c in -1 0 2 -1 0; do
GRACING_COLLECT_COOP_COLUMNS=$c SKIP_EAGER=1 ENVS=65536,4096 TILE_GROUPS=0 timeout 300 python tools/collect_bench.py 2>&1 | grep us_per_step | tr '\n' ' ' | sed "s/^/coop_columns=$c /" | tee -a "$out/${tag}_status.txt"; echo | tee -a "$out/${tag}_status.txt"
done
