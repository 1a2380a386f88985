#!/usr/bin/env python
"""bench.py — reads/s of the RNA-SeQC per-read hot path on B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W                 # this repo's CUDA path
    python bench.py --impl reference --gpus 1 --steps K --warmup W  # the reference's own CPU implementation
    torchrun --nproc-per-node N ... bench.py --gpus N ...          # one rank per GPU, coordinate-partitioned records

A "step" is one whole pass of the hot path over the synthetic records of the workload: accumulator reset, the fused
filter/intersect/accumulate kernel over every record, integration of the coverage difference array, and the read-back
of the count vectors (after the NCCL all-reduce when N > 1).  `value` = records processed by all ranks / step time with
the records already resident in HBM; `e2e` = the same through rsq_submit with pinned HOST buffers (H2D inside the timed
region).  One JSON line on stdout (rank 0).
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


# stdout carries exactly ONE JSON line: anything a library prints to fd 1 (NCCL's version banner, for one) is sent to stderr
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def emit(obj):
    _JSON_OUT.write(json.dumps(obj) + "\n")
    _JSON_OUT.flush()


def build_annotation(wl, tmpdir, with_bed=False):
    from rnaseqc_b200 import api, synth
    if wl["full_genome"]:
        sa = synth.make_annotation(seed=20260925, n_genes=wl["n_genes"], contig_names=synth.GRCH38_NAMES, contig_lens=synth.GRCH38)
    else:
        sa = synth.make_annotation(seed=20260925, n_genes=wl["n_genes"])
    gtf = os.path.join(tmpdir, "bench.gtf")
    synth.write_gtf(sa, gtf)
    bed = None
    if with_bed:    # BASELINE.json configs[4]: the --bed fragment-size path on top (long, isolated exons as intervals)
        bed = os.path.join(tmpdir, "bench.bed")
        synth.write_bed(sa, bed, min_len=1000)
    ann = api.Annotation(gtf, bed)
    return sa, ann, gtf


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.device)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def generate_on_device(sa, ann, n_records, device, rank, world, seed=20260925):
    """Coordinate-sorted packed records on `device`.  With world > 1 every rank draws its own n_records from its own
    coordinate slice of the genome (slices cut at gene-free positions): weak scaling."""
    import torch
    from rnaseqc_b200 import shard, synth
    region = shard.rank_region(sa, rank, world) if world > 1 else None
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


def run_ours(args):
    import torch
    import torch.distributed as dist
    from rnaseqc_b200 import api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = f"cuda:{local_rank}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(device))
    wl = WORKLOADS[args.workload]
    n_records = args.records or wl["records"]
    tmpdir = tempfile.mkdtemp(prefix="rsq_bench_")
    t0 = time.time()
    sa, ann, gtf = build_annotation(wl, tmpdir, with_bed=args.bed)
    log(f"[rank {rank}] annotation: {ann.n_genes} genes, {ann.n_exons} exons, {ann.cov_len} exon bases ({time.time() - t0:.1f}s)")
    rb, bb = generate_on_device(sa, ann, n_records, device, rank, world)
    n_reads, n_blocks = rb.shape[0], bb.shape[0]
    torch.cuda.synchronize()
    log(f"[rank {rank}] records: {n_reads} reads, {n_blocks} blocks resident in HBM ({time.time() - t0:.1f}s)")

    be = api.default_backend()
    stream = torch.cuda.Stream()          # the library runs on this stream, so torch's CUDA events bracket its kernels
    torch.cuda.set_stream(stream)
    ctx = be.create(api.make_config(device=local_rank, max_records=n_reads, user_stream=stream.cuda_stream,
                                    fragment_samples=1000000 if args.bed else 0,
                                    collect_coverage=os.environ.get("RSQ_BENCH_COVERAGE", "1") != "0"))   # env: ablation runs only
    ctx.set_annotation(ann)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allreduce_counts():
        ptr, n_words = ctx.export_counts()
        if world > 1:
            # alias the library's export buffer as a torch tensor (no copy) and sum it over NVLink
            t = _alias_int64(ptr, n_words, local_rank)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            torch.cuda.current_stream().synchronize()
        ctx.import_counts()

    hist_len = (1 << 16) if args.bed else 8      # the fragment-size histogram is only filled by the --bed sampler

    def step_resident():
        ctx.reset()
        ctx.submit_device(rb.data_ptr(), n_reads, bb.data_ptr(), n_blocks)
        if world > 1:
            allreduce_counts()
        res = ctx.finalize(want_coverage=False, hist_len=hist_len, pinned=True)      # result tables into page-locked host buffers
        res.cov_stats = ctx.coverage_statistics(ann, pinned=True)   # coverage integration + per-gene statistics kernel + 5 MB read-back
        return res

    # ---------------- device-resident timing -----------------------------------------------------------------------
    for _ in range(args.warmup):
        res = step_resident()
    algo_bytes = ctx.algorithmic_bytes()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
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
    clocks = sampler.stop()
    ms_dev = ev0.elapsed_time(ev1)
    launches = ctx.launch_count() - l0
    ref_counters, ref_gene_reads = res.counters.copy(), res.gene_reads.copy()   # the pinned result buffers are reused by the next finalize
    # the dominant kernel's own time (CUDA events recorded by the library around the hot-path launches of one step)
    ctx.reset(); ctx.submit_device(rb.data_ptr(), n_reads, bb.data_ptr(), n_blocks); ctx.sync()
    hot_ms = ctx.kernel_time_ms()
    ms_step = max(ms_dev, t_wall * 1e3) / args.steps
    if world > 1:
        t = torch.tensor([ms_step], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step = float(t.item())
    total_records = n_reads * world if world == 1 else int(_allsum(n_reads, device))
    value = total_records / (ms_step * 1e-3)

    # ---------------- end to end from pinned host buffers ------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        try:
            h_reads = torch.empty(rb.shape, dtype=rb.dtype, pin_memory=True)
            h_blocks = torch.empty(bb.shape, dtype=bb.dtype, pin_memory=True)
            h_reads.copy_(rb); h_blocks.copy_(bb)
            torch.cuda.synchronize()
            off = np.frombuffer(h_reads.numpy().reshape(-1).data, dtype=api.READ_DTYPE)
            chunk = args.e2e_chunk
            bounds = list(range(0, n_reads, chunk)) + [n_reads]
            block_lo = [int(off["block_off"][b]) if b < n_reads else n_blocks for b in bounds]
            # block offsets are per submit call: rebase each chunk once, outside the timed region
            for a, b, lo in zip(bounds[:-1], bounds[1:], block_lo[:-1]):
                off["block_off"][a:b] -= np.uint32(lo)

            # compact wire format (rsq_read_packed: 32 B per record, no entry for whole-read blocks): packed once per chunk,
            # outside the timed region, exactly as a decoder would hand the batches over
            packed = []
            try:
                hb_np = np.frombuffer(h_blocks.numpy().reshape(-1).data, dtype=api.BLOCK_DTYPE)
                h_pk = torch.empty((n_reads, 32), dtype=torch.uint8, pin_memory=True)
                h_pb = torch.empty((max(n_blocks, 1), 8), dtype=torch.uint8, pin_memory=True)
                pk_np = np.frombuffer(h_pk.numpy().reshape(-1).data, dtype=api.PACKED_DTYPE)
                pb_np = np.frombuffer(h_pb.numpy().reshape(-1).data, dtype=api.BLOCK_DTYPE)
                w = 0
                for a, b, lo, hi in zip(bounds[:-1], bounds[1:], block_lo[:-1], block_lo[1:]):
                    got = api.pack_records(off[a:b], hb_np[lo:hi], pk_np[a:b], pb_np[w:w + max(hi - lo, 1)])
                    if got is None:
                        packed = None
                        break
                    packed.append((a, b, w, len(got[1])))
                    w += len(got[1])
            except Exception as ex:
                log(f"[rank {rank}] packed wire format unavailable: {ex!r}")
                packed = None

            def step_e2e(use_packed):
                ctx.reset()
                if use_packed:
                    for a, b, w0, nw in packed:
                        ctx.submit_packed_ptr(h_pk.data_ptr() + 32 * a, b - a, h_pb.data_ptr() + 8 * w0, nw)
                else:
                    for a, b, lo, hi in zip(bounds[:-1], bounds[1:], block_lo[:-1], block_lo[1:]):
                        ctx.submit_ptr(h_reads.data_ptr() + 48 * a, b - a, h_blocks.data_ptr() + 8 * lo, hi - lo)
                if world > 1:
                    ctx.sync(); allreduce_counts()
                res = ctx.finalize(want_coverage=False, hist_len=hist_len, pinned=True)
                res.cov_stats = ctx.coverage_statistics(ann, pinned=True)
                return res

            def time_e2e(use_packed):
                for _ in range(max(1, args.warmup - 1)):
                    r = step_e2e(use_packed)
                barrier()
                t0 = time.perf_counter()
                for _ in range(args.steps):
                    r = step_e2e(use_packed)
                barrier()
                dt = (time.perf_counter() - t0) / args.steps
                if world > 1:
                    t = torch.tensor([dt], device=device, dtype=torch.float64)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    dt = float(t.item())
                assert np.array_equal(r.counters, ref_counters) and np.array_equal(r.gene_reads, ref_gene_reads), "e2e != resident"
                return dt

            t_plain = time_e2e(False)
            d2h = 8 * (api.RSQ_N_COUNTERS + 3 * ann.n_genes + 2 * ann.n_exons) + 8 * (5 * len(res.cov_stats["gene_mean"]) + ann.n_exons)
            e2e = {"value": total_records / t_plain, "unit": "reads/s", "h2d_bytes_per_step": int(48 * n_reads + 8 * n_blocks),
                   "d2h_bytes_per_step": int(d2h), "ms_per_step": t_plain * 1e3, "wire_format": "rsq_read (48 B + 8 B per block)"}
            if packed:
                t_packed = time_e2e(True)
                n_wire_blocks = sum(p[3] for p in packed)
                e2e = {"value": total_records / t_packed, "unit": "reads/s", "h2d_bytes_per_step": int(32 * n_reads + 8 * n_wire_blocks),
                       "d2h_bytes_per_step": int(d2h), "ms_per_step": t_packed * 1e3,
                       "wire_format": "rsq_read_packed (32 B + 8 B per block of the records that are not one whole-read run)",
                       "unpacked": {"value": total_records / t_plain, "ms_per_step": t_plain * 1e3,
                                    "h2d_bytes_per_step": int(48 * n_reads + 8 * n_blocks)}}
                del h_pk, h_pb
            del h_reads, h_blocks
        except Exception as ex:  # pinned allocation can fail on a small host
            e2e = {"value": None, "unit": "reads/s", "error": repr(ex)}

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
            ores = octx.finalize(want_coverage=False)
            t_c = time.perf_counter() - t_c
            cpu_baseline = {"value": n_s / t_c, "unit": "reads/s", "cores": 1, "kind": "port",
                            "sample": f"first {n_s} records of the same workload, oracle/rsq_oracle.cpp single thread"}
            octx.close()

    peak, peak_src = measured_peak()
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):   # DRAM bytes per record of the hot kernel from the committed ncu --set full capture, scaled to this launch
        tj = json.load(open(tp))
        traffic, traffic_src = tj["dram_bytes_per_record"] * n_reads, tj["source"]
    hot_ms_per_launch = hot_ms  # one step == one logical launch of the hot path (kSubBatch chunks are back to back)
    achieved = algo_bytes / (hot_ms_per_launch * 1e-3) / 1e9 if hot_ms_per_launch > 0 else None
    if rank == 0:
        out = {
            "metric": "reads/sec (150bp PE, GENCODE-scale GTF)", "value": value, "unit": "reads/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": wl["label"] if not args.records else f"{wl['label']} [records overridden: {n_records}]",
                       "records_per_gpu": n_reads, "blocks_per_gpu": n_blocks, "genes": ann.n_genes, "exons": ann.n_exons,
                       "exon_bases": ann.cov_len, "bed_fragment_sizes": bool(args.bed), "partitioning": "coordinate ranges cut at gene-free positions" if world > 1 else "none",
                       "step": "reset + hot-path kernel + coverage integration + per-gene coverage/bias statistics kernel + count/statistics read-back" + (" + NCCL all-reduce" if world > 1 else ""),
                       "l2_policy": f"inputs ({(48 * n_reads + 8 * n_blocks) / 1e9:.1f} GB per GPU) larger than the 126 MB L2",
                       "not_in_step": "text formatting of the report files and the medians over genes (host, once per sample)"},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_step": int(algo_bytes),
                         "hot_kernel_ms_per_step": hot_ms_per_launch},
            "cpu_baseline": cpu_baseline,
            "sanity": {"mapped_reads": res.counter("Mapped Reads"), "exonic_reads": res.counter("Exonic Reads"),
                       "gene_reads_sum": int(res.gene_reads.sum()), "device_ms_events": ms_dev / args.steps,
                       "genes_with_coverage_stats": int((res.cov_stats["gene_mean"] > 0).sum()),
                       "fragment_samples_taken": int(res.fragment_samples_taken),
                       "genes_in_3prime_bias": int(((res.cov_stats["five_end"] + res.cov_stats["three_end"]) > 0).sum())},
        }
        emit(out)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def _alias_int64(ptr, n, device):
    import torch

    class _CAI:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3}
    return torch.as_tensor(_CAI(), device=f"cuda:{device}")


def _allsum(v, device):
    import torch
    import torch.distributed as dist
    t = torch.tensor([v], device=device, dtype=torch.int64)
    dist.all_reduce(t)
    return t.item()


def run_reference(args):
    """The reference's own CPU implementation of the path, timed on this box's host cores (single-threaded by
    construction, src/BamReader.cpp:14-15), on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from rnaseqc_b200 import api, synth
    wl = WORKLOADS[args.workload]
    tmpdir = tempfile.mkdtemp(prefix="rsq_ref_")
    sa, ann, gtf = build_annotation(wl, tmpdir)
    n_s = args.cpu_sample
    sr = synth.simulate_reads(sa, n_s, seed=20260925, device="cpu", dup_frac=0.0)   # no duplicates: keeps the 1e9-iteration
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "rnaseqc")                      # library-complexity scan (outside the path) out
    times = []
    if os.path.exists(ref_bin):
        bam = os.path.join(tmpdir, "sample.bam")
        synth.write_bam(sa, sr, bam)
        kind, sample = "reference", f"{sr.n} records of the same workload as BAM, unmodified reference binary, 1 thread"
        for k in range(args.warmup + args.steps):
            t = time.perf_counter()
            subprocess.run([ref_bin, gtf, bam, os.path.join(tmpdir, f"out{k}")], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            times.append(time.perf_counter() - t)
        # GTF parse time is outside the path: measure it with an empty-ish BAM and subtract
        empty = synth.simulate_reads(sa, 2, seed=1, device="cpu")
        ebam = os.path.join(tmpdir, "empty.bam")
        synth.write_bam(sa, empty, ebam)
        t = time.perf_counter()
        subprocess.run([ref_bin, gtf, ebam, os.path.join(tmpdir, "oute")], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        t_gtf = time.perf_counter() - t
    else:
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
    out = {"impl": "reference", "metric": "reads/sec (150bp PE, GENCODE-scale GTF)", "value": value, "unit": "reads/s",
           "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
           "config": {"workload": wl["label"], "sample_records": sr.n, "gtf_parse_seconds_subtracted": t_gtf},
           "cpu_baseline": {"value": value, "unit": "reads/s", "cores": 1, "kind": kind, "sample": sample},
           "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="gencode200m", choices=sorted(WORKLOADS))
    ap.add_argument("--records", type=int, default=0, help="override the record count per GPU (smoke runs)")
    ap.add_argument("--cpu-sample", type=int, default=3_000_000,
                    help="records per step of `--impl reference` (the reference binary does ~0.2-0.4 M records/s: ~10 s)")
    ap.add_argument("--port-sample", type=int, default=30_000_000,
                    help="records of the cpu_baseline leg (the oracle port does ~3 M records/s on one core: ~10 s)")
    ap.add_argument("--e2e-chunk", type=int, default=8_000_000)
    ap.add_argument("--bed", action="store_true", help="also run the --bed fragment-size sampler (BASELINE.json configs[4])")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
