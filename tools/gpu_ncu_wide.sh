#!/bin/bash
# ncu --set full capture of the wide-layer tensor-core MLP kernel (one GPU, short command), after a plain run exits 0.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CMD="python tools/seq_bench.py 262144"
timeout 300 $CMD > gpurun_out/ncu_full_plain_wide.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_wide_layer_kernel -s 1 -c 2 -f -o gpurun_out/prof_wide $CMD > gpurun_out/ncu_full_wide.log 2>&1
echo "ncu full exit $?"
tail -3 gpurun_out/ncu_full_wide.log
