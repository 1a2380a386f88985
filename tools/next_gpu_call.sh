#!/bin/bash
# One gpurun call that spends ~12 GPU-minutes on everything the next round needs first:
#   1. the -m gpu suite (new this round: tests/test_zz_random_records.py — odd records and odd pairs on the device),
#   2. a bench line of the product build,
#   3. the prepared kernel experiments (built here by tools/ab_build.sh) timed and parity-checked on the same box.
# usage (from the build container):
#   tools/ab_build.sh stage_desc "-DRSQ_STAGE_DESC" dyn_tiles "-DRSQ_DYNAMIC_TILES" dyn_desc "-DRSQ_DYNAMIC_TILES -DRSQ_STAGE_DESC" \
#                     frag4x "-DRSQ_FRAG_SLOTS_PER_RECORD_X2=8" dyn_desc_frag4x "-DRSQ_DYNAMIC_TILES -DRSQ_STAGE_DESC -DRSQ_FRAG_SLOTS_PER_RECORD_X2=8"
#   gpurun --timeout 1500 -- 'bash tools/next_gpu_call.sh'
#   4. tools/e2e_cli_bench.py: the whole CLI (multi-threaded ingest + CUDA path) on one BAM,
# results: gpurun_out/gpu_tests.log, gpurun_out/bench.json, gpurun_out/ab_hot.jsonl, gpurun_out/e2e_cli.json
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests: exit $?" | tee -a gpurun_out/gpu_tests.log
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench: exit $?"
variants=$(ls build/ab/*.so 2>/dev/null)
timeout 500 python tools/ab_hot.py --records 200000000 --parity 300000 --reps 5 rnaseqc_b200/lib/librsq_b200.so $variants > gpurun_out/ab_hot_stdout.log 2> gpurun_out/ab_hot.err
echo "ab_hot: exit $?"
# 4. the whole CLI on a 30 M-record BAM after the host-side ingest changes (speculated framing, packer thread)
RSQ_INGEST_STATS=1 timeout 600 python tools/e2e_cli_bench.py --records 30000000 --ref-records 0 --threads 0 32 64 > gpurun_out/e2e_cli.json 2> gpurun_out/e2e_cli.err
echo "e2e cli: exit $?"
tail -n 12 gpurun_out/ab_hot_stdout.log
tail -n 3 gpurun_out/gpu_tests.log
