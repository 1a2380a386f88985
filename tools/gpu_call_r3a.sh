#!/bin/bash
# First device call of the next round (1 GPU, ~12 min): everything the last session of round 2 wrote without a device.
#   1. the -m gpu suite (new: six-reset swap of the fragment-tracker tables, wide bias windows in global scratch, per-contig index build)
#   2. bench.py default line, then the same with RSQ_NO_SPARE_TABLE=1: the step should lose ~1 ms at 200 M records, `roofline` should not move
#   3. the CLI leg's three switches (own DEFLATE decoder, early ingest start, quick exit) on one 60 M-record BAM, stage times on stderr
#   4. launch list + one ncu --set full capture of the hot kernel and of coverage_stats_kernel<32> (never timed since its rewrite)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests: exit $?" | tee -a gpurun_out/gpu_tests.log
tail -n 4 gpurun_out/gpu_tests.log
timeout 900 python bench.py --steps 150 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench: exit $?"; tail -c 1200 gpurun_out/bench_default.json
RSQ_NO_SPARE_TABLE=1 timeout 600 python bench.py --steps 150 --warmup 3 --no-e2e --cli-records 0 > gpurun_out/bench_no_spare.json 2> gpurun_out/bench_no_spare.err
echo "bench without the spare table: exit $?"; python - <<'PY'
import json
for f in ("bench_default", "bench_no_spare"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "ms_per_step", round(d["ms_per_step"], 3), "hot kernel", round(d["roofline"]["hot_kernel_ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 4))
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
python tools/e2e_cli_bench.py --records 60000000 --ref-records 0 --threads 16 > gpurun_out/cli_default.json 2> gpurun_out/cli_default.err
for sw in RSQ_INGEST_ZLIB RSQ_INGEST_LATE_START RSQ_FULL_TEARDOWN; do
    env $sw=1 python tools/e2e_cli_bench.py --records 60000000 --ref-records 0 --threads 16 > gpurun_out/cli_$sw.json 2> gpurun_out/cli_$sw.err
done
grep -h ours_decode gpurun_out/cli_*.json | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --cli-records 0 > gpurun_out/b_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"hot_path|coverage_stats" -s 2 -c 3 -o gpurun_out/prof_r3a \
   python tools/ab_hot.py --records 20000000 --parity 0 --reps 1 --max-records-factor 2 rnaseqc_b200/lib/librsq_b200.so > gpurun_out/ncu_r3a.log 2>&1
echo "ncu: exit $?"; ls -la gpurun_out/*.ncu-rep
