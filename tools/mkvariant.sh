#!/bin/bash
# build the current tree into build/lib_<name>.so (extra nvcc flags after the name)
name=$1; shift
mkdir -p /root/repo/build
cd /root/repo/gym_flock_b200/csrc && nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared "$@" -o /root/repo/build/lib_$name.so flock_abi.cu
