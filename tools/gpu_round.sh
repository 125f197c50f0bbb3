#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list.  Everything under timeouts;
# logs land in gpurun_out/ (merged back by gpurun).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
STAGE="${1:-all}"
if [ "$STAGE" = "all" ] || [ "$STAGE" = "tests" ]; then
  timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -x --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
  tail -40 gpurun_out/pytest_gpu.log
fi
if [ "$STAGE" = "all" ] || [ "$STAGE" = "smoke" ]; then
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
  echo "smoke exit $?" >> gpurun_out/smoke.log
  tail -5 gpurun_out/smoke.log
fi
if [ "$STAGE" = "all" ] || [ "$STAGE" = "bench" ]; then
  timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench exit $?"
  tail -c 3000 gpurun_out/bench.json
  tail -5 gpurun_out/bench.err
  timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
  tail -c 1500 gpurun_out/bench_ref.json
fi
if [ "$STAGE" = "all" ] || [ "$STAGE" = "ncu" ]; then
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
  echo "ncu exit $?"
  tail -5 gpurun_out/ncu.log
fi
