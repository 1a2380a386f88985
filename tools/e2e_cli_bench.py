#!/usr/bin/env python
"""End-to-end CLI comparison on one BAM file: this repo's `rnaseqc` (multi-threaded ingest + CUDA hot path) vs the
unmodified reference binary, same GTF, same BAM, same flags.  Wall clock of the whole process, GTF parse included.

    python tools/e2e_cli_bench.py --records 20000000 [--ref-records 3000000]
"""
import argparse, json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rnaseqc_b200 import synth  # noqa: E402


def run(cmd):
    t = time.perf_counter()
    # RSQ_INGEST_STATS=1: the ingest's per-stage busy times (bamparallel.cpp) go to this process's stderr
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=None if os.environ.get("RSQ_INGEST_STATS") else subprocess.DEVNULL)
    return time.perf_counter() - t


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=20_000_000)
    ap.add_argument("--ref-records", type=int, default=3_000_000)
    ap.add_argument("--threads", type=int, nargs="*", default=[0, 8, 32])
    a = ap.parse_args()
    d = tempfile.mkdtemp(prefix="rsq_e2e_")
    sa = synth.make_annotation(seed=20260925, n_genes=56202, contig_names=synth.GRCH38_NAMES, contig_lens=synth.GRCH38)
    gtf = os.path.join(d, "a.gtf")
    synth.write_gtf(sa, gtf)
    import torch
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    out = {"records": a.records, "host_cores": os.cpu_count()}
    big = synth.simulate_reads(sa, a.records, seed=1, device=dev, dup_frac=0.0)
    bam = os.path.join(d, "big.bam")
    t = time.perf_counter(); synth.write_bam(sa, big, bam, level=1); out["bam_write_s"] = time.perf_counter() - t
    out["bam_bytes"] = os.path.getsize(bam)
    cli = os.path.join(ROOT, "rnaseqc_b200", "bin", "rnaseqc")
    for th in a.threads:
        dt = run([cli, gtf, bam, os.path.join(d, f"out_{th}"), "--coverage", "--decode-threads", str(th)])
        out[f"ours_decode_threads_{th}"] = {"seconds": dt, "records_per_s": big.n / dt}
    ref = os.path.join(ROOT, "oracle", "_ref", "rnaseqc")
    if os.path.exists(ref) and a.ref_records:
        small = synth.simulate_reads(sa, a.ref_records, seed=1, device=dev, dup_frac=0.0)
        sbam = os.path.join(d, "small.bam")
        synth.write_bam(sa, small, sbam, level=1)
        dt = run([ref, gtf, sbam, os.path.join(d, "out_ref"), "--coverage"])
        out["reference_1_thread"] = {"records": small.n, "seconds": dt, "records_per_s": small.n / dt}
        dt = run([cli, gtf, sbam, os.path.join(d, "out_small"), "--coverage"])
        out["ours_same_small_file"] = {"records": small.n, "seconds": dt, "records_per_s": small.n / dt}
        same = subprocess.run(["cmp", "-s", os.path.join(d, "out_ref", "small.bam.gene_reads.gct"), os.path.join(d, "out_small", "small.bam.gene_reads.gct")]).returncode == 0
        out["gene_reads_identical_to_reference"] = same
    print(json.dumps(out))


if __name__ == "__main__":
    main()
