#!/usr/bin/env bash
set -u
out=gpurun_out; mkdir -p "$out"; tag=${1:-r2z}
timeout 900 python -m pytest tests/test_mesh_collision.py tests/test_misc.py -m gpu -q -x > "$out/${tag}_pytest.log" 2>&1
echo "pytest: exit $?" | tee "$out/${tag}_status.txt"; tail -n 4 "$out/${tag}_pytest.log"
timeout 600 python tools/mesh_bench.py > "$out/${tag}_mesh_bench.json" 2> "$out/${tag}_mesh_bench.err"
echo "mesh bench: exit $?" | tee -a "$out/${tag}_status.txt"; cat "$out/${tag}_mesh_bench.json" | grep -E "us_per_call|env_step_with|colliding|bundle|per_ray"
