#!/bin/bash
# Builds A/B variants of the CUDA library into build/ab/ (git-ignored; they travel to the GPU box with gpurun).
# usage: tools/ab_build.sh name "-DFLAG ..." [name "-D..."]...     (name "v6" builds the pre-v7 sources from git)
set -e
cd "$(dirname "$0")/.."
mkdir -p build/ab
NVF="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Iinclude -shared -lcudart"
while [ $# -gt 0 ]; do
  name=$1; flags=$2; shift 2
  if [ "$name" = "v6" ]; then
    rm -rf build/ab_v6 && mkdir -p build/ab_v6/include
    for f in rsq_kernels.cu rsq_covstats.cu rsq_capi.cu rsq_device.h rsq_legacy.cuh; do git show 5cc1c35:rnaseqc_b200/csrc/$f > build/ab_v6/$f; done
    git show 5cc1c35:include/rsq_b200.h > build/ab_v6/include/rsq_b200.h
    /usr/local/cuda/bin/nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Ibuild/ab_v6/include -shared -o build/ab/v6.so build/ab_v6/rsq_kernels.cu build/ab_v6/rsq_covstats.cu build/ab_v6/rsq_capi.cu -lcudart &
  else
    /usr/local/cuda/bin/nvcc $NVF $flags -o build/ab/$name.so rnaseqc_b200/csrc/rsq_kernels.cu rnaseqc_b200/csrc/rsq_covstats.cu rnaseqc_b200/csrc/rsq_capi.cu &
  fi
done
wait
ls -la build/ab/
