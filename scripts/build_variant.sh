#!/bin/bash
# build_variant.sh NAME [-D...]: an alternative libwfagpu (kernels compiled with extra defines) next to the
# objects of the normal build, for A/B measurements: WFAGPU_LIB=pywfa_b200/csrc/build/libwfagpu_NAME.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
B=pywfa_b200/csrc/build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Wno-deprecated-gpu-targets "$@" \
  -c pywfa_b200/csrc/wfa_kernels.cu -o $B/wfa_kernels_$name.o
nvcc -shared -Wno-deprecated-gpu-targets -o $B/libwfagpu_$name.so $B/wfa_kernels_$name.o $B/wfa_pack.o $B/wfa_reg_bytes.o $B/wfa_vec_bytes.o $B/wfagpu_api.o $B/pack.o -lpthread
echo $B/libwfagpu_$name.so
