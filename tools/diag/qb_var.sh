#!/bin/bash
# time the fused kernel of every variant library given on the command line (tools/build_variants.sh)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
echo "== default"; timeout 40 python tools/quick_bench.py 20 2>&1 | grep "^fused"
for v in "$@"; do echo "== $v"; NVD_LIB=$PWD/neural-network-based-viterbi-decoding_b200/build/var/libnvd_$v.so timeout 40 python tools/quick_bench.py 20 2>&1 | grep "^fused"; done
