#!/bin/bash
# Round-2 ncu evidence (one GPU).  Every profiled command first runs plain and must exit 0.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,launch__grid_size"
# (1) staged HBM kernels K0 / K1 / K2 / K5 at 2^24, 2^26, 2^28 frames
timeout 200 python tools/hbm_kernels_bench.py > gpurun_out/hbm_kernels.txt 2>&1 && cat gpurun_out/hbm_kernels.txt &&
timeout 600 ncu --metrics $M --clock-control none -k regex:"source_|encode_|channel_|count_" --csv --log-file gpurun_out/ncu_hbm_kernels.csv python tools/hbm_kernels_bench.py > gpurun_out/ncu_hbm.log 2>&1
echo "hbm ncu exit $?"
# (2) launch list of the bench command
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "launch list exit $?"
# (3) --set full of the fused kernel, the K-streamed kernel (staged + fused) and the channel kernel
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-staged --no-e2e --no-legs --frames 4194304"
timeout 200 $CMD > gpurun_out/ncu_full_plain_fused.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel -s 3 -c 2 -f -o gpurun_out/prof_fused $CMD > gpurun_out/ncu_full_fused.log 2>&1
echo "fused full exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_kstream -c 4 -f -o gpurun_out/prof_ks python tools/ks_bench.py 1 > gpurun_out/ncu_ks.log 2>&1
echo "ks full exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:channel_ -c 3 -f -o gpurun_out/prof_channel python tools/hbm_kernels_bench.py > gpurun_out/ncu_channel.log 2>&1
echo "channel full exit $?"
# (4) the streaming host step (MODE_FUSED_IO, one launch with copies arriving beside it): application replay
NVD_HOST_STEP=streaming timeout 100 python tools/host_step_once.py > gpurun_out/host_step_plain.log 2>&1 &&
NVD_HOST_STEP=streaming timeout 900 ncu --set full --replay-mode application --clock-control none -k regex:mlp_tc_kernel -c 1 -s 2 -f -o gpurun_out/prof_host_step python tools/host_step_once.py > gpurun_out/ncu_host_step.log 2>&1
echo "host step full exit $?"; tail -3 gpurun_out/ncu_host_step.log
# (5) compute-sanitizer, if the pool allows it
timeout 300 compute-sanitizer --tool memcheck python tools/quick_bench.py 1 > gpurun_out/sanitizer_memcheck.log 2>&1; echo "sanitizer exit $?"; tail -3 gpurun_out/sanitizer_memcheck.log
ls -la gpurun_out/*.ncu-rep
