#!/usr/bin/env bash
# A/B of the step kernel's launch variants on one GPU: gpurun --timeout 900 -- 'bash tools/ab_step.sh tag'
set -u
out=gpurun_out; mkdir -p "$out"; tag=${1:-ab}
timeout 900 python -m pytest tests -m gpu -q > "$out/${tag}_pytest_gpu.log" 2>&1
echo "pytest -m gpu: exit $?" | tee "$out/${tag}_status.txt"
for es in 0 1; do
for pf in 1 l2; do
  for k in 20 2000; do
    GRACING_EARLY_STORE=$es timeout 300 python bench.py --steps $k --warmup 5 --prefetch $pf --no-cpu --no-collective > "$out/${tag}_bench_es${es}_pf${pf}_k${k}.json" 2> "$out/${tag}_bench_es${es}_pf${pf}_k${k}.err"
    echo "bench es=$es pf=$pf k=$k: exit $?" | tee -a "$out/${tag}_status.txt"
  done
done
done
tail -5 "$out/${tag}_pytest_gpu.log"
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_bench_es*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        ex=d["extra"]; t=d["timing"]
        print(f, "kernel_us %.3f"%d["roofline"]["kernel_us"], "isolated %.3f"%t["isolated_kernel_us"], t["events"], "frac %.3f"%d["roofline"]["frac"], "low %.3f"%ex["fwd_low_reset"]["kernel_us"],
              "1pct %.3f"%ex["fwd_forced_reset_rate"]["1pct"]["kernel_us"], "10pct %.3f"%ex["fwd_forced_reset_rate"]["10pct"]["kernel_us"], "gae %.2f"%ex["gae_24x4096_us"], "e2e %.1fM"%(d["e2e"]["value"]/1e6))
    except Exception as e:
        print(f, "ERR", e)
PY
