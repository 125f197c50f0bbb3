#!/bin/bash
# ncu --set full of the TMA-staged K-stream kernel and of the streaming host step (application replay)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 100 python tools/ks_bench.py 1 > gpurun_out/ks_bench.txt 2>&1 && cat gpurun_out/ks_bench.txt &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_kstream_tma -c 2 -f -o gpurun_out/prof_ks_tma python tools/ks_bench.py 1 > gpurun_out/ncu_ks_tma.log 2>&1
echo "ks tma full exit $?"
NVD_HOST_STEP=streaming timeout 100 python tools/host_step_once.py > gpurun_out/host_step_plain.log 2>&1 &&
NVD_HOST_STEP=streaming timeout 900 ncu --set full --replay-mode application --clock-control none -k regex:mlp_tc_kernel -c 1 -s 2 -f -o gpurun_out/prof_host_step python tools/host_step_once.py > gpurun_out/ncu_host_step.log 2>&1
echo "host step full exit $?"; tail -3 gpurun_out/ncu_host_step.log
