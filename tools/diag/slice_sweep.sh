#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for d in 2 3 4 6; do for f in 2 4; do echo "== div $d first $f"; NVD_HOST_SLICE_DIV=$d NVD_HOST_SLICE_FIRST=$f timeout 60 python tools/host_step_bench.py 2>&1 | grep "^streaming" | tail -2; done; done
