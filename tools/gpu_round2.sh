#!/bin/bash
# One GPU-box visit (round 2): parity tests, smoke, bench, quick kernel timing.  Logs land in gpurun_out/.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
STAGE="${1:-all}"
if [ "$STAGE" = "all" ] || [ "$STAGE" = "quick" ]; then
  timeout 300 python tools/quick_bench.py 20 > gpurun_out/qb.txt 2>&1; echo "quick exit $?"; cat gpurun_out/qb.txt
fi
if [ "$STAGE" = "all" ] || [ "$STAGE" = "tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
  tail -60 gpurun_out/pytest_gpu.log
fi
if [ "$STAGE" = "all" ] || [ "$STAGE" = "smoke" ]; then
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
  echo "smoke exit $?" >> gpurun_out/smoke.log
  tail -5 gpurun_out/smoke.log
fi
if [ "$STAGE" = "all" ] || [ "$STAGE" = "bench" ]; then
  timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench exit $?"
  tail -c 6000 gpurun_out/bench.json
  tail -5 gpurun_out/bench.err
fi
