#!/usr/bin/env bash
set -u
out=gpurun_out; mkdir -p "$out"; tag=${1:-r2n}
timeout 900 python -m pytest tests/test_actor_backward.py tests/test_ppo_update_kernels.py tests/test_ppo_graphed_update.py tests/test_bptt_collect.py tests/test_runners_gpu.py tests/test_mesh_collision.py -m gpu -q -x > "$out/${tag}_pytest.log" 2>&1
echo "pytest: exit $?" | tee "$out/${tag}_status.txt"
tail -n 5 "$out/${tag}_pytest.log"
timeout 300 python tools/train.py ppo --num_envs 65536 --iters 6 --fused --kernel_update > "$out/${tag}_ppo.log" 2>&1
echo "ppo: exit $?" | tee -a "$out/${tag}_status.txt"; tail -n 2 "$out/${tag}_ppo.log"
timeout 300 python tools/train.py ppo --num_envs 4096 --iters 6 --fused --kernel_update > "$out/${tag}_ppo4k.log" 2>&1
echo "ppo4k: exit $?" | tee -a "$out/${tag}_status.txt"; tail -n 2 "$out/${tag}_ppo4k.log"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$out/${tag}_smoke.log" 2>&1
echo "smoke: exit $?" | tee -a "$out/${tag}_status.txt"; tail -n 2 "$out/${tag}_smoke.log"
