#!/bin/bash
# ncu --set full capture of the dominant kernel (one GPU, short command), after a plain run exits 0.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
KREGEX="${1:-mlp_tc_kernel}"
TAG="${2:-fused}"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-staged --no-e2e --frames 4194304"
if [ "$TAG" = "staged" ]; then CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --frames 4194304"; fi
timeout 300 $CMD > gpurun_out/ncu_full_plain_$TAG.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 3 -c 2 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full exit $?"
tail -3 gpurun_out/ncu_full_$TAG.log
ls -la gpurun_out/*.ncu-rep
