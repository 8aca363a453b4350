#!/bin/bash
# usage: scripts/build_k3_variant.sh NAME "<extra nvcc flags>"   ->  pathfinder_b200/libpfb200_NAME.so
# A/B builds of the KP = 12 instantiation of K3 (the bench's kernel); everything else is shared.
set -e
cd "$(dirname "$0")/../pathfinder_b200/csrc"
NAME=$1; FLAGS=$2
mkdir -p _build/var_$NAME
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -Xptxas -v \
  -DPFB_K3_KP=12 -DPFB_K3_ENTRY=pfb_launch_k3_kp12 $FLAGS -c k3_elbo_sample_mma.cu -o _build/var_$NAME/k3_kp12.o \
  2> _build/var_$NAME/k3_kp12.ptxas.log || (cat _build/var_$NAME/k3_kp12.ptxas.log; exit 1)
OBJS=$(ls _build/*.o | grep -v k3_kp12.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libpfb200_$NAME.so $OBJS _build/var_$NAME/k3_kp12.o -ldl -Xlinker -rpath=/usr/local/cuda/lib64
grep -A2 "pfb_k3_elbo_sampleILi12ELi1ELi2ELb0" _build/var_$NAME/k3_kp12.ptxas.log | tail -2
