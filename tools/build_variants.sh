#!/bin/bash
# build experiment variants of libnvd (only k_mlp_tc.cu and its two-CTA twin recompiled): build/var/libnvd_<VARIANT>.so
# VARIANT = tokens joined by '+': NAME -> -DNVD_EXP_NAME, NAME=VALUE -> -DNVD_NAME=VALUE
set -e
cd /root/repo/neural-network-based-viterbi-decoding_b200
mkdir -p build/var
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
for V in "$@"; do
  DEFS=""
  for d in $(echo $V | tr '+' ' '); do
    case "$d" in *=*) DEFS="$DEFS -DNVD_$d";; *) DEFS="$DEFS -DNVD_EXP_$d";; esac
  done
  nvcc $FLAGS $DEFS -c csrc/k_mlp_tc.cu -o build/var/k_mlp_tc_$V.o &
  nvcc $FLAGS $DEFS -c csrc/k_mlp_tc_g2.cu -o build/var/k_mlp_tc_g2_$V.o &
done
wait
for V in "$@"; do
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o build/var/libnvd_$V.so build/api.o build/comm.o build/k_mlp_kstream.o build/k_source_encode.o build/k_channel.o build/k_mlp_fp32.o build/k_mlp_wide_tc.o build/var/k_mlp_tc_$V.o build/var/k_mlp_tc_g2_$V.o build/k_lstm.o build/k_lstm_tc.o build/k_count.o build/k_train.o -ldl
done
ls build/var/*.so
