#!/usr/bin/env python
"""bench.py — reads/s of the RNA-SeQC per-read hot path on B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W                   # this repo's CUDA path
    python bench.py --impl reference --gpus 1 --steps K --warmup W  # the reference's own CPU implementation
    torchrun --nproc-per-node N ... bench.py --gpus N ...            # one rank per GPU, coordinate-partitioned records
         [--scaling strong]   BASELINE.json configs[3]: the 200 M records of the workload split over the N GPUs
         [--bed]              BASELINE.json configs[4]: fragment-size sampler + per-base coverage on top

A "step" is one whole pass of the hot path over the synthetic records of the workload: accumulator reset, the fused
filter/intersect/accumulate kernel over every record, the end-of-run exchange over NCCL when N > 1 (inside the library:
rsq_exchange — count vectors, ordered first-N fragment-size cut), integration of the coverage difference array, the per-gene
coverage / 3' bias statistics of the genes the rank owns (gathered over NCCL), and the read-back of every result table.
`value` = records processed by all ranks / step time with the records already resident in HBM; `e2e` = the same through
rsq_submit_compact (16-byte wire format) with pinned HOST buffers (H2D inside the timed region).  One JSON line on stdout (rank 0).

At N > 1 a PRE-FLIGHT runs first: one 2 M-record sorted stream is split at gene-free cuts, every rank runs its slice, the
exchange runs, and every rank's results must equal bit for bit the single-GPU run of the whole stream on rank 0 and the CPU
oracle's (counters, gene and exon tables, fragment-size histogram, coverage statistics).  A failed pre-flight aborts the run.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[2]: the configuration the metric is quoted on (GENCODE-scale GTF, 200 M 150 bp PE records)
    "gencode200m": dict(n_genes=56202, full_genome=True, records=200_000_000,
                        label="synthetic 200M 150bp PE records, GENCODE-scale collapsed GTF (56202 genes, 25 contigs)"),
    # BASELINE.json configs[1]
    "chr1_10m": dict(n_genes=5234, full_genome=False, records=10_000_000,
                     label="synthetic 10M 150bp PE records, chr1-only collapsed GTF (5234 genes)"),
}
METRIC = "reads/sec (150bp PE, GENCODE-scale GTF)"
SEED = 20260925


# stdout carries exactly ONE JSON line: anything a library prints to fd 1 (NCCL's version banner, for one) is sent to stderr
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def emit(obj):
    _JSON_OUT.write(json.dumps(obj) + "\n")
    _JSON_OUT.flush()


def make_synth_annotation(wl):
    from rnaseqc_b200 import synth
    if wl["full_genome"]:
        return synth.make_annotation(seed=SEED, n_genes=wl["n_genes"], contig_names=synth.GRCH38_NAMES, contig_lens=synth.GRCH38)
    return synth.make_annotation(seed=SEED, n_genes=wl["n_genes"])


def build_annotation(wl, tmpdir, with_bed=False):
    from rnaseqc_b200 import api, synth
    sa = make_synth_annotation(wl)
    gtf = os.path.join(tmpdir, "bench.gtf")
    synth.write_gtf(sa, gtf)
    bed = None
    if with_bed:    # BASELINE.json configs[4]: the --bed fragment-size path on top (long, isolated exons as intervals)
        bed = os.path.join(tmpdir, "bench.bed")
        synth.write_bed(sa, bed, min_len=1000)
    ann = api.Annotation(gtf, bed)
    return sa, ann, gtf


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the GPU is under load (B200_PROFILING.md).  Started BEFORE the warm-up
    steps (nvidia-smi needs about a second to deliver its first row) and read after the timed region: warm-up and timed steps
    are one uninterrupted stretch of the same work."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.device)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def n_rows(self):
        return len(self.rows)

    def stop(self, since=0.0):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for t, r in self.rows:
            if t < since:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": "warm-up + timed steps (continuous load)"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def bind_to_gpu_cores(local_rank, world):
    """Pins this rank to the host cores NVML reports as local to its GPU — BEFORE any pinned buffer is allocated, so that
    first-touch places the staging memory on the GPU's NUMA node.  When every GPU reports the same core set (a VM with one
    NUMA node), the set is split evenly among the ranks so that the copy threads do not share cores.  Returns a description."""
    info = {"bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cores = [64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1 and 64 * w + b < n_cpu]
        info["gpu_local_cores"] = f"{cores[0]}-{cores[-1]}" if cores else ""
        try:
            info["numa_node"] = pynvml.nvmlDeviceGetNumaNodeId(h)
        except Exception:
            info["numa_node"] = None
        allowed = sorted(os.sched_getaffinity(0))
        cores = [c for c in cores if c in allowed] or allowed
        if world > 1:
            same = True
            for r in range(world):
                w2 = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(r), (n_cpu + 63) // 64)
                same = same and list(w2) == list(words)
            if same and len(cores) >= world:       # one shared core set: give every rank its own slice of it
                per = len(cores) // world
                cores = cores[local_rank * per:(local_rank + 1) * per]
        os.sched_setaffinity(0, cores)
        info.update(bound=True, cores=f"{cores[0]}-{cores[-1]}", n_cores=len(cores))
    except Exception as ex:
        info["error"] = repr(ex)[:200]
    return info


def generate_on_device(sa, ann, n_records, device, region, rank, seed=SEED):
    """Coordinate-sorted packed records on `device`, drawn from `region` of the genome (None = all of it)."""
    import torch
    from rnaseqc_b200 import synth
    chunk = 20_000_000
    parts = []
    done = 0
    k = 0
    while done < n_records:
        n = min(chunk, n_records - done)
        sr = synth.simulate_reads(sa, n, seed=seed + 1000 * rank + k, device=device, sort=False, region=region,
                                  name_base=(rank << 40) + done)
        parts.append(sr)
        done += sr.n
        k += 1
    sr = synth.concat(parts)
    del parts
    sr = synth.sort_reads(sr)
    idx = [ann.contig_index(n) for n in sa.contig_names]
    rb, bb = synth.pack(sr, idx)
    del sr
    torch.cuda.empty_cache()
    return rb, bb


RESULT_FIELDS = ("counters", "gene_reads", "gene_unique", "gene_fragments", "exon_whole", "exon_frac", "fragment_hist", "gc_hist")
STAT_FIELDS = ("gene_mean", "gene_std", "gene_cv", "five_end", "three_end", "exon_cv")
DIAG_COUNTER = 53      # RSQ_C_FRAGMENT_CANDIDATES: a diagnostic, counted only while the sample quota lasts (tests/util.py)


def snapshot(res, cs):
    out = {k: np.array(getattr(res, k), copy=True) for k in RESULT_FIELDS}
    out["counters"][DIAG_COUNTER] = 0
    out["taken"] = np.array([res.fragment_samples_taken], np.uint64)
    for k in STAT_FIELDS:
        out[k] = np.array(cs[k], copy=True).view(np.uint64)      # bit patterns: NaN == NaN
    return out


def preflight(be, ann, sa, device_index, rank, world, with_bed, n_records=2_000_000, quota=20000, log_fn=log):
    """N>1 parity on the hardware the bench runs on (see the module docstring).  Returns a summary dict (rank 0) or raises."""
    import torch
    import torch.distributed as dist
    from rnaseqc_b200 import api, shard, synth
    device = f"cuda:{device_index}"
    sr = synth.simulate_reads(sa, n_records, seed=SEED + 77, device=device)           # the same stream on every rank
    contig, pos = sr.contig.cpu().numpy(), sr.pos.cpu().numpy()
    ranges, cuts = shard.split_sorted_records(sa, contig, pos, world, return_cuts=True)
    shard.assert_cuts_legal(ann, sa, cuts)
    rb, bb = synth.pack(sr, [ann.contig_index(n) for n in sa.contig_names])
    reads, blocks = synth.to_numpy(rb, bb)
    lo, hi = ranges[rank]
    r = reads[lo:hi].copy()
    b0 = int(r["block_off"][0]) if len(r) else 0
    b1 = int(r["block_off"][-1]) + int(r["n_blocks"][-1]) if len(r) else 0
    r["block_off"] -= np.uint32(b0)
    fs = quota if with_bed else 0
    ctx = be.create(api.make_config(device=device_index, max_records=max(len(r), 1024), fragment_samples=fs))
    ctx.set_annotation(ann)
    ids = [be.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.comm_init(ids[0], rank, world)
    ctx.submit(r, blocks[b0:b1].copy())
    ctx.exchange()
    res = ctx.finalize(want_coverage=False)
    cs = ctx.coverage_statistics(ann, owned=shard.owned_gene_lines(ann, sa, cuts, rank))
    mine = snapshot(res, cs)
    ctx.close()
    want = [None]
    summary = None
    if rank == 0:
        one = be.create(api.make_config(device=device_index, max_records=len(reads), fragment_samples=fs))
        one.set_annotation(ann)
        one.submit(reads, blocks)
        ref = one.finalize(want_coverage=False)
        want[0] = snapshot(ref, one.coverage_statistics(ann))
        one.close()
        summary = {"records": int(len(reads)), "partition_sizes": [int(b - a) for a, b in ranges], "fragment_sample_quota": fs,
                   "fragment_samples_taken": int(want[0]["taken"][0]), "gene_reads_sum": int(want[0]["gene_reads"].sum()),
                   "compared": list(RESULT_FIELDS) + ["fragment_samples_taken"] + list(STAT_FIELDS), "vs_single_gpu": "bit-identical on every rank"}
        oracle_lib = os.path.join(ROOT, "oracle", "_build", "librsq_oracle.so")
        if os.path.exists(oracle_lib):            # the checker (test infrastructure), never the thing measured
            octx = api.Backend(oracle_lib).create(api.make_config(fragment_samples=fs))
            octx.set_annotation(ann)
            octx.submit(reads, blocks)
            ores = octx.finalize(want_coverage=False)
            osnap = snapshot(ores, octx.coverage_statistics(ann))
            octx.close()
            for k in list(RESULT_FIELDS) + ["taken"]:
                assert np.array_equal(osnap[k], want[0][k]), f"pre-flight: single-GPU run differs from the CPU oracle in {k}"
            # block-reduced fp64 from exact integer moments on the device vs the reference's sequential fp64: same rule as
            # tests/test_gpu_parity.py::test_coverage_statistics_match_oracle (1e-9 relative, a flat profile's std is exactly 0 on
            # the device and rounding noise in the reference); the u64 bias ends and the NaN pattern exactly
            f64 = lambda snap, k: snap[k].view(np.float64)
            for k in ("five_end", "three_end"):
                assert np.array_equal(osnap[k], want[0][k]), f"pre-flight: {k} differs from the CPU oracle"
            g_mean, w_mean = f64(osnap, "gene_mean"), f64(want[0], "gene_mean")
            assert np.array_equal(g_mean == -1.0, w_mean == -1.0) and np.allclose(w_mean, g_mean, rtol=1e-9, atol=0), "pre-flight: gene_mean differs from the CPU oracle"
            assert np.allclose(f64(want[0], "gene_std"), f64(osnap, "gene_std"), rtol=1e-9, atol=1e-9 * np.maximum(np.abs(w_mean), 1e-300)), "pre-flight: gene_std differs from the CPU oracle"
            a, b = f64(want[0], "gene_cv"), f64(osnap, "gene_cv")
            assert np.array_equal(np.isnan(a), np.isnan(b)) and np.allclose(a[~np.isnan(a)], b[~np.isnan(b)], rtol=1e-9, atol=1e-9), "pre-flight: gene_cv differs from the CPU oracle"
            a, b = f64(want[0], "exon_cv"), f64(osnap, "exon_cv")
            assert np.array_equal(np.isnan(a), np.isnan(b)) and np.allclose(a[~np.isnan(a)], b[~np.isnan(b)], rtol=1e-9, atol=1e-9), "pre-flight: exon_cv differs from the CPU oracle"
            summary["vs_cpu_oracle"] = "integers bit-identical, coverage statistics within 1e-9"
    dist.broadcast_object_list(want, src=0)
    bad = [k for k in want[0] if not np.array_equal(want[0][k], mine[k])]
    flags = [None] * world
    dist.all_gather_object(flags, bad)
    if any(flags):
        raise SystemExit(f"pre-flight FAILED: sharded != single-GPU run: {flags}")
    if rank == 0:
        log_fn(f"[pre-flight] {world} partitions of {len(reads)} records == single-GPU run == oracle: {summary}")
    return summary


def run_ours(args):
    import torch
    import torch.distributed as dist
    from rnaseqc_b200 import api, shard

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    binding = bind_to_gpu_cores(local_rank, world) if not args.no_bind else {"bound": False}
    torch.cuda.set_device(local_rank)
    device = f"cuda:{local_rank}"
    if world > 1:
        import datetime
        # a rank that falls out of step with the others must end the run within minutes, not hold the box until the job's limit
        dist.init_process_group("nccl", device_id=torch.device(device), timeout=datetime.timedelta(seconds=300))
    wl = WORKLOADS[args.workload]
    total_target = args.records or wl["records"]
    strong = args.scaling == "strong" and world > 1
    n_records = (total_target + world - 1) // world if strong else total_target
    tmpdir = tempfile.mkdtemp(prefix="rsq_bench_")
    t0 = time.time()
    sa, ann, gtf = build_annotation(wl, tmpdir, with_bed=args.bed)
    log(f"[rank {rank}] annotation: {ann.n_genes} genes, {ann.n_exons} exons, {ann.cov_len} exon bases ({time.time() - t0:.1f}s); cpu binding {binding}")
    be = api.default_backend()

    pre = None
    if world > 1 and not args.no_preflight:
        # the pre-flight always runs with the --bed sampler on (the ordered first-N cut across the partitions is part of the exchange)
        ann_pf = ann if args.bed else build_annotation(wl, tempfile.mkdtemp(prefix="rsq_pf_"), with_bed=True)[1]
        pre = preflight(be, ann_pf, sa, local_rank, rank, world, with_bed=True)
        del ann_pf

    # this rank's coordinate slice of the genome (cut where no gene lies): it draws its records there and owns those genes
    region, owned = None, None
    if world > 1:
        regions = [shard.rank_region(sa, r, world) for r in range(world)]
        region = regions[rank]
        cut_lins = [rg["lin_lo"] for rg in regions[1:]]
        shard.assert_cuts_legal(ann, sa, cut_lins)
        owned = shard.owned_gene_lines(ann, sa, cut_lins, rank)
    rb, bb = generate_on_device(sa, ann, n_records, device, region, rank)
    n_reads, n_blocks = rb.shape[0], bb.shape[0]
    torch.cuda.synchronize()
    log(f"[rank {rank}] records: {n_reads} reads, {n_blocks} blocks resident in HBM ({time.time() - t0:.1f}s)")

    stream = torch.cuda.Stream()          # the library runs on this stream, so torch's CUDA events bracket its kernels
    torch.cuda.set_stream(stream)
    ctx = be.create(api.make_config(device=local_rank, max_records=n_reads, user_stream=stream.cuda_stream,
                                    fragment_samples=1000000 if args.bed else 0,
                                    collect_coverage=os.environ.get("RSQ_BENCH_COVERAGE", "1") != "0"))   # env: ablation runs only
    ctx.set_annotation(ann)
    if world > 1:
        ids = [be.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(ids[0], rank, world)            # NCCL communicator inside the library (rsq_exchange / rsq_coverage_statistics)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    hist_len = (1 << 16) if args.bed else 8      # the fragment-size histogram is only filled by the --bed sampler

    def finish():
        if world > 1:
            ctx.exchange()                                                     # collective, on the ctx stream, no host round trip
        res = ctx.finalize(want_coverage=False, hist_len=hist_len, pinned=True)   # result tables into page-locked host buffers
        res.cov_stats = ctx.coverage_statistics(ann, pinned=True, owned=owned)    # own genes only; gathered over NCCL when N > 1
        return res

    def step_resident():
        ctx.reset()
        ctx.submit_device(rb.data_ptr(), n_reads, bb.data_ptr(), n_blocks)
        return finish()

    # ---------------- device-resident timing -----------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_load = time.perf_counter()
    res = None
    for _ in range(args.warmup):
        res = step_resident()
    # warm-up goes on until min_warmup_s have passed (nvidia-smi needs about a second of load before its first row).  Every step is
    # a collective at N > 1, so the NUMBER of extra steps is fixed on rank 0 and broadcast: ranks that disagreed on it would wait
    # for each other forever (that was the hang of the first 8-GPU run of round 2)
    per_step = max((time.perf_counter() - t_load) / max(args.warmup, 1), 1e-4)
    extra = [int(min(1000, max(0.0, args.min_warmup_s - (time.perf_counter() - t_load)) / per_step))]
    if world > 1:
        dist.broadcast_object_list(extra, src=0)
    for _ in range(extra[0]):
        res = step_resident()
    warm_done = args.warmup + extra[0]
    algo_bytes = ctx.algorithmic_bytes()
    barrier()
    l0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    t_wall = time.perf_counter()
    for _ in range(args.steps):
        res = step_resident()
    ev1.record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.stop(since=t_load + 0.2)
    ms_dev = ev0.elapsed_time(ev1)
    launches = ctx.launch_count() - l0
    ref_snap = snapshot(res, res.cov_stats)       # the pinned result buffers are reused by the next finalize
    cs = res.cov_stats
    n_cov_genes = len(cs["gene_mean"])
    sanity = {"mapped_reads": res.counter("Mapped Reads"), "exonic_reads": res.counter("Exonic Reads"),
              "gene_reads_sum": int(res.gene_reads.sum()), "device_ms_events": ms_dev / args.steps,
              "genes_with_coverage_stats": int((cs["gene_mean"] > 0).sum()),
              "fragment_samples_taken": int(res.fragment_samples_taken),
              "genes_in_3prime_bias": int(((cs["five_end"] + cs["three_end"]) > 0).sum())}
    # the dominant kernel's own time (CUDA events recorded by the library around the hot-path launches of one step)
    ctx.reset()
    torch.cuda.synchronize()      # the kernel ALONE: also the side stream, where the other fragment-tracker table is being zeroed
    ctx.submit_device(rb.data_ptr(), n_reads, bb.data_ptr(), n_blocks); ctx.sync()
    hot_ms = ctx.kernel_time_ms()
    finish()
    ms_step_local = max(ms_dev, t_wall * 1e3) / args.steps
    ms_step = ms_step_local
    if world > 1:
        t = torch.tensor([ms_step], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step = float(t.item())
    total_records = n_reads if world == 1 else int(_allsum(n_reads, device))
    value = total_records / (ms_step * 1e-3)

    # ---------------- the same step over a region of >= --sustain-s seconds (the K steps above are 14 ms each) -----------------
    sustained = None
    if ms_step * args.steps < args.sustain_s * 1e3:
        m_steps = int(min(5000, args.sustain_s * 1e3 / ms_step + 1))      # from the max-over-ranks time: the same count on every rank
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        t_s = time.perf_counter()
        for _ in range(m_steps):
            step_resident()
        s1.record(stream)
        barrier()
        t_s = time.perf_counter() - t_s
        ms_s = max(s0.elapsed_time(s1), t_s * 1e3) / m_steps
        if world > 1:
            t = torch.tensor([ms_s], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_s = float(t.item())
        sustained = {"steps": m_steps, "seconds": ms_s * m_steps / 1e3, "ms_per_step": ms_s, "value": total_records / (ms_s * 1e-3), "unit": "reads/s"}

    # ---------------- BASELINE configs[3] beside the weak line: the workload's records SPLIT over the GPUs -------------------
    # (the driver launches `bench.py --gpus N` without flags; `--scaling strong` makes this the headline instead)
    strong_leg = None
    if world > 1 and not strong and not args.no_strong_leg:
        n_s = (total_target + world - 1) // world
        rb2, bb2 = generate_on_device(sa, ann, n_s, device, region, rank, seed=SEED + 7)
        n2, nb2 = rb2.shape[0], bb2.shape[0]

        def step_strong():
            ctx.reset()
            ctx.submit_device(rb2.data_ptr(), n2, bb2.data_ptr(), nb2)
            return finish()
        for _ in range(3):
            step_strong()
        barrier()
        t_s = time.perf_counter()
        for _ in range(args.steps):
            step_strong()
        barrier()
        dt_s = (time.perf_counter() - t_s) / args.steps
        ctx.reset()
        torch.cuda.synchronize()
        ctx.submit_device(rb2.data_ptr(), n2, bb2.data_ptr(), nb2); ctx.sync()
        hot_s = ctx.kernel_time_ms()
        finish()
        t = torch.tensor([dt_s, hot_s], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_s = int(_allsum(n2, device))
        strong_leg = {"scaling": "strong", "workload": "BASELINE configs[3]: the workload's records split over the GPUs by coordinate",
                      "records_total": total_s, "records_per_gpu": int(n2), "steps": args.steps, "ms_per_step": float(t[0].item()) * 1e3,
                      "value": total_s / float(t[0].item()), "unit": "reads/s", "slowest_hot_kernel_ms": float(t[1].item())}
        del rb2, bb2
        torch.cuda.empty_cache()

    # ---------------- end to end from pinned host buffers ------------------------------------------------------------
    e2e = None
    e2e_rank = {}
    have_pinned = not args.no_e2e
    if have_pinned:
        try:
            h_reads = torch.empty(rb.shape, dtype=rb.dtype, pin_memory=True)
            h_blocks = torch.empty(bb.shape, dtype=bb.dtype, pin_memory=True)
        except (MemoryError, RuntimeError) as ex:   # pinned allocation can fail on a small host
            log(f"[rank {rank}] no page-locked memory for the e2e leg: {ex!r}")
            have_pinned = False
        if world > 1:                               # every step is a collective: the ranks go through the e2e leg together or not at all
            t = torch.tensor([1 if have_pinned else 0], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            have_pinned = bool(t.item())
        if not have_pinned:
            e2e = {"value": None, "unit": "reads/s", "error": "page-locked host buffers could not be allocated on every rank"}
    if have_pinned:
        try:
            h_reads.copy_(rb); h_blocks.copy_(bb)
            torch.cuda.synchronize()
            off = np.frombuffer(h_reads.numpy().reshape(-1).data, dtype=api.READ_DTYPE)
            chunk = args.e2e_chunk
            bounds = list(range(0, n_reads, chunk)) + [n_reads]
            block_lo = [int(off["block_off"][b]) if b < n_reads else n_blocks for b in bounds]
            # block offsets are per submit call: rebase each chunk once, outside the timed region
            for a, b, lo in zip(bounds[:-1], bounds[1:], block_lo[:-1]):
                off["block_off"][a:b] -= np.uint32(lo)

            hb_np = np.frombuffer(h_blocks.numpy().reshape(-1).data, dtype=api.BLOCK_DTYPE)
            # the wire formats: every chunk is packed ONCE, outside the timed region, into page-locked memory — exactly as a
            # decoder hands its batches over (the CLI's ingest workers pack while the previous batch is on the wire)
            compact, packed = None, None
            if args.e2e_format == "compact":        # rsq_batch16: 16-byte records + tile headers + dictionary + 32-bit blocks
                try:
                    cap = sum(api.compact_bound(b - a, hi - lo) for a, b, lo, hi in zip(bounds[:-1], bounds[1:], block_lo[:-1], block_lo[1:]))
                    h_cp = torch.empty((cap + 64,), dtype=torch.uint8, pin_memory=True)
                    cp_np = h_cp.numpy()
                    compact, w = [], (-cp_np.ctypes.data) % 16
                    for a, b, lo, hi in zip(bounds[:-1], bounds[1:], block_lo[:-1], block_lo[1:]):
                        room = api.compact_bound(b - a, hi - lo)
                        got = api.pack_compact(off[a:b], hb_np[lo:hi], cp_np[w:w + room])
                        if got is None:
                            compact = None
                            break
                        compact.append((cp_np.ctypes.data + w, got.nbytes))
                        w += (got.nbytes + 15) & ~15          # the chunks lie back to back: what is not used of `room` is reused
                except Exception as ex:
                    log(f"[rank {rank}] compact wire format unavailable: {ex!r}")
                    compact = None
            if compact is None and args.e2e_format in ("compact", "packed"):   # rsq_read_packed: 32 B per record, implicit whole-read blocks
                h_pk = torch.empty((n_reads, 32), dtype=torch.uint8, pin_memory=True)
                h_pb = torch.empty((max(n_blocks, 1), 8), dtype=torch.uint8, pin_memory=True)
                pk_np = np.frombuffer(h_pk.numpy().reshape(-1).data, dtype=api.PACKED_DTYPE)
                pb_np = np.frombuffer(h_pb.numpy().reshape(-1).data, dtype=api.BLOCK_DTYPE)
                packed, w = [], 0
                for a, b, lo, hi in zip(bounds[:-1], bounds[1:], block_lo[:-1], block_lo[1:]):
                    got = api.pack_records(off[a:b], hb_np[lo:hi], pk_np[a:b], pb_np[w:w + max(hi - lo, 1)])
                    if got is None:
                        packed = None
                        break
                    packed.append((a, b, w, len(got[1])))
                    w += len(got[1])

            def step_e2e(fmt):
                ctx.reset()
                if fmt == "compact":
                    for ptr, nbytes in compact:
                        ctx.submit_compact_ptr(ptr, nbytes)
                elif fmt == "packed":
                    for a, b, w0, nw in packed:
                        ctx.submit_packed_ptr(h_pk.data_ptr() + 32 * a, b - a, h_pb.data_ptr() + 8 * w0, nw)
                else:
                    for a, b, lo, hi in zip(bounds[:-1], bounds[1:], block_lo[:-1], block_lo[1:]):
                        ctx.submit_ptr(h_reads.data_ptr() + 48 * a, b - a, h_blocks.data_ptr() + 8 * lo, hi - lo)
                return finish()

            def time_e2e(fmt):
                for _ in range(max(1, min(args.warmup, 3) - 1)):
                    r = step_e2e(fmt)
                barrier()
                t0 = time.perf_counter()
                for _ in range(args.e2e_steps):
                    r = step_e2e(fmt)
                local = (time.perf_counter() - t0) / args.e2e_steps       # before the barrier: this rank's own time
                barrier()
                dt = (time.perf_counter() - t0) / args.e2e_steps
                if world > 1:
                    t = torch.tensor([dt], device=device, dtype=torch.float64)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    dt = float(t.item())
                got = snapshot(r, r.cov_stats)
                bad = [k for k in ref_snap if not np.array_equal(ref_snap[k], got[k])]
                assert not bad, f"e2e ({fmt}) != resident in {bad}"
                return dt, local

            d2h = 8 * (api.RSQ_N_COUNTERS + 3 * ann.n_genes + 2 * ann.n_exons) + 8 * (5 * n_cov_genes + ann.n_exons)
            fmt = "compact" if compact else ("packed" if packed else "plain")
            wire = {"compact": lambda: int(sum(nb_ for _, nb_ in compact)),
                    "packed": lambda: int(32 * n_reads + 8 * sum(p[3] for p in packed)),
                    "plain": lambda: int(48 * n_reads + 8 * n_blocks)}
            names = {"compact": "rsq_batch16 (16-byte records, tile headers, per-batch dictionary, 32-bit blocks, exception list; one copy per batch)",
                     "packed": "rsq_read_packed (32 B + 8 B per block of the records that are not one whole-read run)",
                     "plain": "rsq_read (48 B + 8 B per block)"}
            t_e, t_local = time_e2e(fmt)
            h2d = wire[fmt]()
            e2e = {"value": total_records / t_e, "unit": "reads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(d2h),
                   "ms_per_step": t_e * 1e3, "steps": args.e2e_steps, "wire_format": names[fmt], "wire_bytes_per_record": h2d / n_reads,
                   "same_results_as_resident": True}
            e2e_rank = {"e2e_ms": t_local * 1e3, "h2d_gb_per_s": h2d / t_local / 1e9}
            for other in args.e2e_also:
                if other != fmt and (other != "packed" or packed) and (other != "compact" or compact):
                    t_o, _ = time_e2e(other)
                    e2e[other] = {"value": total_records / t_o, "ms_per_step": t_o * 1e3, "h2d_bytes_per_step": wire[other]()}
            del h_reads, h_blocks
        except MemoryError as ex:  # pinned allocation can fail on a small host
            e2e = {"value": None, "unit": "reads/s", "error": repr(ex)}

    # ---------------- per-rank facts (every rank, gathered on rank 0) --------------------------------------------------------
    mine = {"rank": rank, "records": int(n_reads), "hot_kernel_ms": hot_ms, "step_ms": ms_step_local, "cpu_binding": binding, **e2e_rank}
    per_rank = [mine]
    if world > 1:
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)

    # ---------------- CPU baseline (oracle port) on a bounded sample, rank 0 at N=1 only ---------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:     # N=1 only: the other ranks of a multi-GPU run would wait ~10 s for it
        oracle_lib = os.path.join(ROOT, "oracle", "_build", "librsq_oracle.so")
        if os.path.exists(oracle_lib):
            n_s = min(n_reads, args.port_sample)
            sr = np.frombuffer(rb[:n_s].cpu().numpy().reshape(-1).data, dtype=api.READ_DTYPE).copy()
            nb_s = int(sr["block_off"][-1]) + int(sr["n_blocks"][-1])
            sb = bb[:nb_s].cpu().numpy().copy()
            octx = api.Backend(oracle_lib).create(api.make_config())
            octx.set_annotation(ann)
            t_c = time.perf_counter()
            octx.submit(sr, sb)
            octx.finalize(want_coverage=False)
            t_c = time.perf_counter() - t_c
            cpu_baseline = {"value": n_s / t_c, "unit": "reads/s", "cores": 1, "kind": "port",
                            "sample": f"first {n_s} records of the same workload, oracle/rsq_oracle.cpp single thread"}
            octx.close()

    # ---------------- BAM -> report files through the product CLI (ingest included), rank 0 at N=1 only --------------------------
    e2e_cli = None
    if rank == 0 and world == 1 and args.cli_records > 0:
        try:
            e2e_cli = run_cli_leg(sa, gtf, tmpdir, args.cli_records, device)
        except Exception as ex:
            e2e_cli = {"error": repr(ex)[:300]}

    peak, peak_src = measured_peak()
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):   # DRAM bytes per record of the hot kernel from the committed ncu --set full capture, scaled to this launch
        tj = json.load(open(tp))
        traffic, traffic_src = tj["dram_bytes_per_record"] * n_reads, tj["source"]
    achieved = algo_bytes / (hot_ms * 1e-3) / 1e9 if hot_ms > 0 else None
    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm_done, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": wl["label"] if not args.records else f"{wl['label']} [records overridden: {total_target}]",
                       "records_total": int(total_records), "records_per_gpu": n_reads, "blocks_per_gpu": n_blocks, "genes": ann.n_genes, "exons": ann.n_exons,
                       "exon_bases": ann.cov_len, "bed_fragment_sizes": bool(args.bed),
                       "partitioning": ("coordinate ranges cut at gene-free positions; " + ("the workload's records split over the GPUs (BASELINE configs[3])" if strong else "every GPU draws the workload's record count from its own range")) if world > 1 else "none",
                       "step": "reset (the fragment-tracker table of the next pass is zeroed on a side stream during this one) + hot-path kernel" + (" + rsq_exchange (NCCL inside the library: count vectors, ordered first-N fragment-size cut)" if world > 1 else "") +
                               " + coverage integration + per-gene coverage/bias statistics kernel" + (" of the rank's own genes + NCCL gather" if world > 1 else "") + " + count/statistics read-back",
                       "l2_policy": f"inputs ({(48 * n_reads + 8 * n_blocks) / 1e9:.1f} GB per GPU) larger than the 126 MB L2",
                       "timed_seconds": ms_step * args.steps / 1e3,
                       "not_in_step": "text formatting of the report files and the medians over genes (host, once per sample)"},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_step": int(algo_bytes),
                         "hot_kernel_ms_per_step": hot_ms, "kernel": "hot_path_kernel (rank 0; every rank's time is in per_rank)"},
            "cpu_baseline": cpu_baseline, "e2e_cli": e2e_cli, "per_rank": per_rank, "preflight": pre, "strong_scaling": strong_leg,
            "sanity": sanity, "sustained": sustained,
        }
        emit(out)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def run_cli_leg(sa, gtf, tmpdir, n_records, device):
    """The number a drop-in user sees: BAM file -> the nine report files through rnaseqc_b200/bin/rnaseqc (multi-threaded
    BGZF/BAM ingest + CUDA path + report writers), wall clock of the whole process, against the reference binary on a bounded
    sample of the same BAM's records."""
    from rnaseqc_b200 import synth
    cli = os.path.join(ROOT, "rnaseqc_b200", "bin", "rnaseqc")
    if not os.path.exists(cli):
        return {"error": "rnaseqc_b200/bin/rnaseqc is not built"}
    sr = synth.simulate_reads(sa, n_records, seed=SEED + 5, device=device, dup_frac=0.0)
    bam = os.path.join(tmpdir, "cli.bam")
    t = time.perf_counter()
    synth.write_bam(sa, sr, bam, level=1)
    t_write = time.perf_counter() - t
    out = {"records": int(sr.n), "bam_bytes": os.path.getsize(bam), "bam_write_seconds": t_write, "host_cores": os.cpu_count(), "flags": "--coverage"}
    env = dict(os.environ, RSQ_INGEST_STATS="1", RSQ_SETUP_STATS="1")      # stage times of the ingest and of the start-up on stderr
    best = None
    for rep in range(2):
        t = time.perf_counter()
        p = subprocess.run([cli, gtf, bam, os.path.join(tmpdir, f"cli_out{rep}"), "--coverage", "-v"], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        dt = time.perf_counter() - t
        if p.returncode != 0:
            return dict(out, error=f"exit {p.returncode}: {p.stderr[-300:]}")
        if best is None or dt < best[0]:
            best = (dt, p.stdout, p.stderr)
    stages = {}
    for line in best[1].splitlines():
        for key, name in (("Finished processing GTF in", "gtf_parse_s"), ("Device context + annotation upload:", "device_setup_s"),
                          ("Time Elapsed:", "bam_loop_s"), ("Finalize + coverage statistics:", "finalize_s"), ("Report files:", "report_s")):
            if line.startswith(key):
                try:
                    stages[name] = float(line[len(key):].split()[0].rstrip(";"))
                except ValueError:
                    pass
    out.update(seconds=best[0], records_per_s=sr.n / best[0], stages=stages,
               bam_loop_records_per_s=(sr.n / stages["bam_loop_s"]) if stages.get("bam_loop_s") else None,
               ingest_stats=[l for l in best[2].splitlines() if l.startswith(("ingest busy seconds", "load_gtf:", "rsq_set_annotation:"))][-16:])
    return out


def _allsum(v, device):
    import torch
    import torch.distributed as dist
    t = torch.tensor([v], device=device, dtype=torch.int64)
    dist.all_reduce(t)
    return t.item()


def run_reference(args):
    """The reference's own CPU implementation of the path, timed on this box's host cores (single-threaded by
    construction, src/BamReader.cpp:14-15), on a bounded sample of the same workload.  Nothing of this repo's product is
    loaded here: the GTF and the BAM are written by the synthetic generator (rnaseqc_b200.synth; the BAM through htslib-free
    Python/zlib code), the timed part is a subprocess of oracle/_ref/rnaseqc."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from rnaseqc_b200 import synth
    wl = WORKLOADS[args.workload]
    tmpdir = tempfile.mkdtemp(prefix="rsq_ref_")
    sa = make_synth_annotation(wl)
    gtf = os.path.join(tmpdir, "bench.gtf")
    synth.write_gtf(sa, gtf)
    flags = ["--coverage"]
    if args.bed:
        bed = os.path.join(tmpdir, "bench.bed")
        synth.write_bed(sa, bed, min_len=1000)
        flags += ["--bed", bed]
    # every step is one run of the reference binary: ~6 s of fixed cost (its GTF parse, the report files of 54 k genes) + the
    # sample at ~0.3 M records/s.  The sample shrinks with the number of steps asked for, so that the whole arm stays within
    # --ref-budget-s whatever --steps / --warmup the caller chose.
    asked = (args.steps, args.warmup)
    max_runs = max(2, int(args.ref_budget_s / 7.0))          # even the smallest sample costs ~7 s per run
    if args.warmup + args.steps > max_runs:
        args.warmup = min(args.warmup, 1)
        args.steps = max(1, min(args.steps, max_runs - args.warmup))
    runs = max(args.warmup + args.steps, 1)
    n_s = int(min(args.cpu_sample, max(300_000, (args.ref_budget_s / runs - 6.0) * 0.3e6)))
    sr = synth.simulate_reads(sa, n_s, seed=SEED, device="cpu", dup_frac=0.0)   # no duplicates: keeps the 1e9-iteration
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "rnaseqc")                  # library-complexity scan (outside the path) out
    times = []
    if os.path.exists(ref_bin):
        bam = os.path.join(tmpdir, "sample.bam")
        synth.write_bam_python(sa, sr, bam)
        kind, sample = "reference", f"{sr.n} records of the same workload as BAM, unmodified reference binary ({' '.join(flags)}), 1 thread"
        for k in range(args.warmup + args.steps):
            t = time.perf_counter()
            subprocess.run([ref_bin, gtf, bam, os.path.join(tmpdir, f"out{k}")] + flags, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            times.append(time.perf_counter() - t)
        # GTF parse time is outside the path: measure it with an empty-ish BAM and subtract
        empty = synth.simulate_reads(sa, 2, seed=1, device="cpu")
        ebam = os.path.join(tmpdir, "empty.bam")
        synth.write_bam_python(sa, empty, ebam)
        t = time.perf_counter()
        subprocess.run([ref_bin, gtf, ebam, os.path.join(tmpdir, "oute")] + flags, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        t_gtf = time.perf_counter() - t
    else:
        # the reference did not build here: the oracle port stands in (the one other place bench.py may execute oracle/)
        from rnaseqc_b200 import api
        ann = api.Annotation(gtf)
        kind, sample = "port", f"{sr.n} records of the same workload, oracle/rsq_oracle.cpp, 1 thread"
        rb, bb = synth.pack(sr, [ann.contig_index(n) for n in sa.contig_names])
        reads, blocks = synth.to_numpy(rb, bb)
        be = api.Backend(os.path.join(ROOT, "oracle", "_build", "librsq_oracle.so"))
        t_gtf = 0.0
        for k in range(args.warmup + args.steps):
            ctx = be.create(api.make_config())
            ctx.set_annotation(ann)
            t = time.perf_counter()
            ctx.submit(reads, blocks); ctx.finalize(want_coverage=True)
            times.append(time.perf_counter() - t)
            ctx.close()
    timed = times[args.warmup:]
    per_step = max(sum(timed) / len(timed) - t_gtf, 1e-9)
    value = sr.n / per_step
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": "reads/s",
           "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
           "config": {"workload": wl["label"], "sample_records": sr.n, "steps_asked": asked[0], "warmup_asked": asked[1], "wall_budget_seconds": args.ref_budget_s,
                      "gtf_parse_seconds_subtracted": t_gtf, "flags": flags[:1] + (["--bed"] if args.bed else []),
                      "includes": "BGZF inflate + BAM record parse (about 10 % of the reference's time, SURVEY.md §6) — the GPU arm's e2e starts from decoded, packed records; bench.py's e2e_cli leg is the BAM-to-files comparison"},
           "cpu_baseline": {"value": value, "unit": "reads/s", "cores": 1, "kind": kind, "sample": sample},
           "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=150, help="timed steps (150 x ~14 ms: more than 2 s of timed work at the default workload)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--min-warmup-s", type=float, default=1.5, help="warm-up goes on until this many seconds have passed (clock sampling needs load)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="gencode200m", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = every GPU gets the workload's record count; strong = the workload's records are split (configs[3])")
    ap.add_argument("--records", type=int, default=0, help="override the workload's record count (smoke runs)")
    ap.add_argument("--cpu-sample", type=int, default=3_000_000,
                    help="records per step of `--impl reference` (the reference binary does ~0.2-0.4 M records/s: ~10 s)")
    ap.add_argument("--ref-budget-s", type=float, default=300.0, help="`--impl reference`: wall-clock budget of all its runs together (sizes the sample)")
    ap.add_argument("--sustain-s", type=float, default=2.0,
                    help="when the K timed steps take less than this, a second region of as many steps as fill it is timed the same way (reported as `sustained`)")
    ap.add_argument("--watchdog-s", type=float, default=720.0, help="dump all stacks and exit 1 when the run takes longer than this (0 = off)")
    ap.add_argument("--port-sample", type=int, default=30_000_000,
                    help="records of the cpu_baseline leg (the oracle port does ~3 M records/s on one core: ~10 s)")
    ap.add_argument("--e2e-chunk", type=int, default=8_000_000)
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--e2e-format", default="compact", choices=["compact", "packed", "plain"], help="wire format of the e2e leg")
    ap.add_argument("--e2e-also", nargs="*", default=[], choices=["packed", "plain"], help="also time these formats (reported beside the headline one)")
    ap.add_argument("--cli-records", type=int, default=60_000_000,
                    help="records of the BAM the e2e_cli leg runs the product CLI on (N=1 only; 0 = skip)")
    ap.add_argument("--bed", action="store_true", help="also run the --bed fragment-size sampler (BASELINE.json configs[4])")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-preflight", action="store_true")
    ap.add_argument("--no-strong-leg", action="store_true", help="N > 1, weak scaling: skip the extra configs[3] measurement")
    ap.add_argument("--no-bind", action="store_true", help="do not pin the rank to its GPU's host cores")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 0)
    if args.watchdog_s > 0:
        # the library's own NCCL calls have no timeout: a rank that falls out of step would hold the box until the job's limit
        import faulthandler
        faulthandler.dump_traceback_later(args.watchdog_s, exit=True)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
