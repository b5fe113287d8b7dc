#!/usr/bin/env bash
set -u
out=gpurun_out; mkdir -p "$out"; tag=${1:-ab}
timeout 300 python -m pytest tests/test_env_parity.py tests/test_baseline_sizes.py tests/test_philox_chain.py -m gpu -q -x > "$out/${tag}_pytest.log" 2>&1
echo "pytest: exit $?"; tail -n 2 "$out/${tag}_pytest.log"
for k in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-collective --no-cpu > "$out/${tag}_bench_$k.json" 2> "$out/${tag}_bench_$k.err"
python - "$out/${tag}_bench_$k.json" <<'PY'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith('{'):
        d=json.loads(line); t=d['timing']; print('step us', round(d['ms_per_step']*1e3,3), 'min', t['kernel_us_per_rank_min'], 'iso', round(t['isolated_kernel_us'],3), 'frac', round(d['roofline']['frac'],4))
PY
done
