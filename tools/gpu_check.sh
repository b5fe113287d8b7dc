#!/usr/bin/env bash
# gpurun --timeout 1200 -- 'bash tools/gpu_check.sh tag'   : gpu tests + the driver's bench command + default bench
set -u
out=gpurun_out; mkdir -p "$out"; tag=${1:-chk}
timeout 900 python -m pytest tests -m gpu -q > "$out/${tag}_pytest_gpu.log" 2>&1
echo "pytest -m gpu: exit $?" | tee "$out/${tag}_status.txt"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$out/${tag}_smoke.log" 2>&1
echo "smoke: exit $?" | tee -a "$out/${tag}_status.txt"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > "$out/${tag}_bench_k20.json" 2> "$out/${tag}_bench_k20.err"
echo "bench k20: exit $?" | tee -a "$out/${tag}_status.txt"
timeout 600 python bench.py --no-cpu --no-collective > "$out/${tag}_bench_default.json" 2> "$out/${tag}_bench_default.err"
echo "bench default: exit $?" | tee -a "$out/${tag}_status.txt"
tail -4 "$out/${tag}_pytest_gpu.log"; tail -2 "$out/${tag}_smoke.log"; tail -3 "$out/${tag}_bench_k20.err"
python - <<PY
import json
for f in ("gpurun_out/${tag}_bench_k20.json","gpurun_out/${tag}_bench_default.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        ex=d["extra"]; t=d["timing"]; e=d["e2e"]
        print(f, "kernel_us %.3f"%d["roofline"]["kernel_us"], "isolated %.3f"%t["isolated_kernel_us"], t["events"], "frac %.3f"%d["roofline"]["frac"], "low %.3f"%ex["fwd_low_reset"]["kernel_us"])
        print("  e2e %.1fM (%.1f us) byte-dones %.1fM probe %.1f us %.1f GB/s"%(e["value"]/1e6, e["ms_per_step"]*1e3, e["byte_dones"]["value"]/1e6, e["copy_only_probe"]["us_per_step"], e["copy_only_probe"]["per_gpu_GBps"]))
        print("  gae %.2f graph %.2f add %.2f"%(ex["gae_24x4096_us"], ex["gae_24x4096_graph_us"], ex["add_transitions_4096_us"]))
        print("  bptt c3:", ex["bptt_fwd_bwd_c3"]["ms_per_window_graph"], "one-launch:", ex["bptt_fwd_bwd_c3_one_launch_window"]["ms_per_window_graph"], ex["bptt_fwd_bwd_c3_one_launch_window"]["ms_forward_window_python_driven"], "sweep", ex["bptt_bwd_sweep"])
        print("  fwd window 65536x24:", ex["fwd_window_65536x24"]["us_per_step"])
        print("  collection:", {k:v["fused"]["us_per_step"] for k,v in ex["ppo_collection"].items() if k!="note"})
        print("  collective:", d.get("collective"))
    except Exception as ex_:
        print(f, "ERR", repr(ex_))
PY
