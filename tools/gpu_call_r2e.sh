#!/bin/bash
# round 2, call E (8 GPUs): box facts, H2D ceiling, the N=8 bench lines (weak, strong = configs[3], --bed = configs[4])
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{ nvidia-smi topo -m; nvidia-smi --query-gpu=index,pci.bus_id --format=csv; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)"; free -g | head -2; } > gpurun_out/topo8.txt 2>&1
timeout 300 python tools/h2d_probe.py > gpurun_out/h2d_probe.log 2>&1; echo "h2d probe: exit $?"; cat gpurun_out/h2d_probe.log | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621"
timeout 900 $TR bench.py --gpus 8 --steps 100 --warmup 3 > gpurun_out/bench_8gpu_weak.json 2> gpurun_out/bench_8gpu_weak.err
echo "bench N=8 weak: exit $?"; grep -E "pre-flight|Error|error" gpurun_out/bench_8gpu_weak.err | cut -c1-400 | tail -n 4
timeout 600 $TR bench.py --gpus 8 --steps 100 --warmup 3 --scaling strong > gpurun_out/bench_8gpu_strong.json 2> gpurun_out/bench_8gpu_strong.err
echo "bench N=8 strong: exit $?"; tail -n 2 gpurun_out/bench_8gpu_strong.err | cut -c1-300
timeout 900 $TR bench.py --gpus 8 --steps 50 --warmup 3 --bed > gpurun_out/bench_8gpu_bed.json 2> gpurun_out/bench_8gpu_bed.err
echo "bench N=8 bed: exit $?"; tail -n 2 gpurun_out/bench_8gpu_bed.err | cut -c1-300
timeout 300 python -m pytest tests/test_gpu_multi.py tests/test_cli_multi.py -m gpu -q > gpurun_out/gpu_multi8_tests.log 2>&1; echo "multi tests: exit $?"; tail -n 3 gpurun_out/gpu_multi8_tests.log
for f in weak strong bed; do python - "$f" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/bench_8gpu_{f}.json"))
    print(f, "value %.2f G/s" % (d["value"] / 1e9), "ms %.3f" % d["ms_per_step"], "e2e %.2f G/s" % (d["e2e"]["value"] / 1e9), "e2e ms %.1f" % d["e2e"]["ms_per_step"],
          "hot", [round(r["hot_kernel_ms"], 2) for r in d["per_rank"]], "h2d", [round(r.get("h2d_gb_per_s", 0), 1) for r in d["per_rank"]])
except Exception as e:
    print(f, "no result", e)
PY
done
