#!/usr/bin/env bash
set -u
out=gpurun_out; mkdir -p "$out"; tag=${1:-q}
timeout 600 python -m pytest tests/test_philox_chain.py tests/test_baseline_sizes.py tests/test_rollout_window.py -m gpu -q > "$out/${tag}_pytest.log" 2>&1
echo "pytest: exit $?"; tail -3 "$out/${tag}_pytest.log"
for coop in 0 1; do
  GRACING_COOP_RESET=$coop timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-collective > "$out/${tag}_bench_coop${coop}.json" 2> "$out/${tag}_bench_coop${coop}.err"
done
python - <<PY
import json
for f in ("gpurun_out/${tag}_bench_coop0.json","gpurun_out/${tag}_bench_coop1.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        ex=d["extra"]
        print(f, "kernel_us %.3f"%d["roofline"]["kernel_us"], "frac %.3f"%d["roofline"]["frac"], "low %.3f"%ex["fwd_low_reset"]["kernel_us"], "1pct %.3f"%ex["fwd_forced_reset_rate"]["1pct"]["kernel_us"], "10pct %.3f"%ex["fwd_forced_reset_rate"]["10pct"]["kernel_us"])
        w=ex["bptt_fwd_bwd_c3_one_launch_window"]
        print("  bptt c3 one-launch:", w["ms_per_window_graph"], "fwd", w["ms_forward_window_graph"], "sweep", ex["bptt_bwd_sweep"]["ms"], " fwd window 65536x24:", ex["fwd_window_65536x24"]["us_per_step"])
    except Exception as e:
        print(f, "ERR", repr(e))
PY
