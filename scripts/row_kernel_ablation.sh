#!/bin/bash
# Builds timing-ablation variants of the row kernel into gpurun_ab/ (git-ignored; they travel to the GPU box) next to a copy
# of the shipped library.  Run here, then on the box: python scripts/ab_lookahead.py gpurun_ab/base.so gpurun_ab/<variant>.so ...
# Usage: bash scripts/row_kernel_ablation.sh "CN_ABLATE_EPI" "CN_ABLATE_EPI -DCN_ABLATE_MMA" ...
set -e
cd "$(dirname "$0")/../modelcrowdnav_b200/csrc"
make -j8 libcrowdnav_b200.so > /dev/null
mkdir -p ../../gpurun_ab
cp libcrowdnav_b200.so ../../gpurun_ab/base.so
for v in "$@"; do
  n=$(echo "$v" | sed 's/ -D/_/g; s/CN_//g')
  nvcc -D$v -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -c lookahead_tc.cu -o /tmp/lt_$n.o 2> /dev/null
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../gpurun_ab/$n.so capi.o orca_step.o lookahead_f32.o /tmp/lt_$n.o trainer.o world_model.o scenes_host.o -lcudart_static -lrt -lpthread -ldl
  echo built gpurun_ab/$n.so
done
