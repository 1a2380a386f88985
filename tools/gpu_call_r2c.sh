#!/bin/bash
# round 2, call C (1 GPU): full -m gpu suite, the default bench line (200 M records, compact e2e, e2e_cli leg), ncu source capture
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests: exit $?" | tee -a gpurun_out/gpu_tests.log
tail -n 4 gpurun_out/gpu_tests.log
timeout 900 python bench.py --e2e-also packed > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
echo "bench: exit $?"; tail -n 4 gpurun_out/bench_1gpu.err | cut -c1-300
cut -c1-3000 gpurun_out/bench_1gpu.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hot_path -s 2 -c 1 -o gpurun_out/prof_hot \
   python tools/ab_hot.py --records 20000000 --parity 0 --reps 1 --max-records-factor 2 rnaseqc_b200/lib/librsq_b200.so > gpurun_out/ncu_hot.log 2>&1
echo "ncu: exit $?"; ls -la gpurun_out/*.ncu-rep
