#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 200 python tools/ks_bench.py 3 > gpurun_out/ks_bench.txt 2>&1; cat gpurun_out/ks_bench.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_kstream -c 4 -o gpurun_out/prof_ks -f python tools/ks_bench.py 1 > gpurun_out/ncu_ks.log 2>&1
tail -3 gpurun_out/ncu_ks.log
