#!/bin/bash
# ncu --set full capture of the tensor-core LSTM recurrence (one GPU, short command), after a plain run exits 0.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CMD="python tools/lstm_bench.py 94720"
timeout 300 $CMD > gpurun_out/ncu_full_plain_lstm.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_tc_kernel -s 2 -c 2 -f -o gpurun_out/prof_lstm_tc $CMD > gpurun_out/ncu_full_lstm.log 2>&1
echo "ncu full exit $?"
tail -3 gpurun_out/ncu_full_lstm.log
ls -la gpurun_out/*.ncu-rep
