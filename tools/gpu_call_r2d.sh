#!/bin/bash
# round 2, call D (2 GPUs): whole -m gpu suite incl. the multi-GPU tests (library NCCL, product CLI --gpus), ncu source capture
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests: exit $?" | tee -a gpurun_out/gpu_tests.log
tail -n 6 gpurun_out/gpu_tests.log
CUDA_VISIBLE_DEVICES=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:hot_path -s 2 -c 1 -o gpurun_out/prof_hot \
   python tools/ab_hot.py --records 20000000 --parity 0 --reps 1 --max-records-factor 2 rnaseqc_b200/lib/librsq_b200.so > gpurun_out/ncu_hot.log 2>&1
echo "ncu: exit $?"; ls -la gpurun_out/*.ncu-rep; tail -n 3 gpurun_out/ncu_hot.log
