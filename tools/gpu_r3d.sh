#!/usr/bin/env bash
set -u
out=gpurun_out; mkdir -p "$out"; tag=${1:-r3d}
timeout 600 python -m pytest tests/test_host_pipe.py tests/test_abi.py -m gpu -q -x > "$out/${tag}_pytest.log" 2>&1
echo "pytest: exit $?" | tee "$out/${tag}_status.txt"; tail -n 4 "$out/${tag}_pytest.log"
timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-collective > "$out/${tag}_bench.json" 2> "$out/${tag}_bench.err"
echo "bench: exit $?" | tee -a "$out/${tag}_status.txt"; tail -c 300 "$out/${tag}_bench.err"
