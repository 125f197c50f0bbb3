#!/bin/bash
# ncu --set full of the fused FNN kernel alone (one GPU); the command first runs plain and must exit 0
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-staged --no-e2e --no-legs --frames 4194304"
timeout 200 $CMD > gpurun_out/ncu_full_plain_fused.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel -s 3 -c 2 -f -o gpurun_out/prof_fused $CMD > gpurun_out/ncu_full_fused.log 2>&1
echo "fused full exit $?"
ls -la gpurun_out/prof_fused.ncu-rep
