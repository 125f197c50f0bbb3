#!/bin/bash
# 8-GPU visit: topology, full bench line, per-rank host-step trace of the end-to-end call.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N="${1:-8}"
{ nvidia-smi topo -m; echo; lscpu | grep -i "numa\|socket\|model name\|^CPU(s)"; echo; for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q "^0x0302" $d/class; then echo "$d numa_node=$(cat $d/numa_node) local_cpulist=$(cat $d/local_cpulist)"; fi; done; echo; cat /sys/devices/system/node/node*/cpulist; nproc; taskset -p $$; } > gpurun_out/topo_n$N.txt 2>&1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
NCCL_DEBUG=WARN timeout 300 $RUN bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench exit $?"
tail -c 1500 gpurun_out/bench_n$N.json | head -c 600; echo
[ "${SKIP_TRACE:-0}" = "1" ] && exit 0
NVD_HOST_STEP_TRACE=1 timeout 200 $RUN bench.py --gpus $N --steps 5 --warmup 3 --no-legs --no-staged --no-cpu-baseline > gpurun_out/trace_n$N.json 2> gpurun_out/trace_n$N.err; echo "trace exit $?"
grep -c "host step trace" gpurun_out/trace_n$N.err
