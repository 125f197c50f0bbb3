#!/bin/bash
# ncu --set full of the streaming host step (MODE_FUSED_IO, one launch with copies arriving beside it): application replay
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
NVD_HOST_STEP=streaming timeout 100 python tools/host_step_once.py > gpurun_out/host_step_plain.log 2>&1 &&
NVD_HOST_STEP=streaming timeout 900 ncu --set full --replay-mode application --clock-control none -k regex:mlp_tc_kernel -c 1 -s 2 -f -o gpurun_out/prof_host_step python tools/host_step_once.py > gpurun_out/ncu_host_step.log 2>&1
echo "host step full exit $?"; tail -3 gpurun_out/ncu_host_step.log
