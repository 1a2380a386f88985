#!/bin/bash
# round 2, call A: state check + the prepared kernel switches (A/B on one box) + box facts for the multi-GPU work
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
nvidia-smi --query-gpu=index,name,pci.bus_id,memory.total --format=csv >> gpurun_out/topo.txt 2>&1
lscpu | head -30 >> gpurun_out/topo.txt; numactl -H >> gpurun_out/topo.txt 2>&1; free -g >> gpurun_out/topo.txt
ls /usr/lib/x86_64-linux-gnu/libnccl* >> gpurun_out/topo.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests: exit $?" | tee -a gpurun_out/gpu_tests.log
variants=$(ls build/ab/*.so 2>/dev/null)
timeout 900 python tools/ab_hot.py --records 200000000 --parity 300000 --reps 5 rnaseqc_b200/lib/librsq_b200.so $variants > gpurun_out/ab_hot_stdout.log 2> gpurun_out/ab_hot.err
echo "ab_hot: exit $?"
cat gpurun_out/ab_hot_stdout.log | cut -c1-400
tail -n 3 gpurun_out/gpu_tests.log
