#!/bin/bash
# round 2, call B (2 GPUs): N>1 parity on the device (NCCL inside the library) + the reworked bench at N=2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo2.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/gpu_multi_tests.log 2>&1
echo "gpu multi tests: exit $?" | tee -a gpurun_out/gpu_multi_tests.log
tail -n 5 gpurun_out/gpu_multi_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 --records 40000000 --bed > gpurun_out/bench_2gpu_bed_40m.json 2> gpurun_out/bench_2gpu_bed.err
echo "bench N=2 bed weak: exit $?"; tail -n 3 gpurun_out/bench_2gpu_bed.err | cut -c1-300
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 --records 40000000 --scaling strong > gpurun_out/bench_2gpu_strong_40m.json 2> gpurun_out/bench_2gpu_strong.err
echo "bench N=2 strong: exit $?"; tail -n 3 gpurun_out/bench_2gpu_strong.err | cut -c1-300
cut -c1-1500 gpurun_out/bench_2gpu_bed_40m.json
