#!/usr/bin/env python
"""A/B timing of CUDA-library variants (build/ab/*.so, tools/ab_build.sh) on ONE box, same records, same annotation.

    python tools/ab_hot.py [--records 20000000] [--workload gencode200m] [--parity 300000] [--bed] variant.so ...

Per variant: bit-exact parity of the first `--parity` records against the CPU oracle (test infrastructure), then the hot
kernel's own time (library CUDA events, median of `--reps`) with the records resident in HBM.  One JSON line per variant
on stdout and in gpurun_out/ab_hot.jsonl."""
import argparse
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("libs", nargs="+")
    ap.add_argument("--records", type=int, default=20_000_000)
    ap.add_argument("--workload", default="gencode200m")
    ap.add_argument("--parity", type=int, default=300_000)
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--bed", action="store_true")
    ap.add_argument("--no-coverage", action="store_true")
    ap.add_argument("--max-records-factor", type=float, default=1.0, help="rsq_config.max_records = factor x records (2: the whole batch is ONE launch)")
    ap.add_argument("--cache", default="", help="directory to keep the generated records in between invocations (e.g. /dev/shm/ab)")
    ap.add_argument("--device", default="cuda:0", help="cpu only to smoke-test this script against the oracle library")
    args = ap.parse_args()
    import torch
    import bench
    from rnaseqc_b200 import api
    from util import assert_results_equal

    wl = bench.WORKLOADS[args.workload]
    tmp = tempfile.mkdtemp(prefix="rsq_ab_")
    sa, ann, gtf = bench.build_annotation(wl, tmp, with_bed=args.bed)
    cache = os.path.join(args.cache, f"records_{args.workload}_{args.records}.pt") if args.cache else None
    if cache and os.path.exists(cache):
        rb, bb = [t.to(args.device) for t in torch.load(cache)]
    else:
        rb, bb = bench.generate_on_device(sa, ann, args.records, args.device, None, 0)
        if cache:
            os.makedirs(args.cache, exist_ok=True)
            torch.save([rb.cpu(), bb.cpu()], cache)
    n_reads, n_blocks = rb.shape[0], bb.shape[0]
    if args.device != "cpu":
        torch.cuda.synchronize()
    print(f"records: {n_reads}, blocks: {n_blocks}, genes {ann.n_genes}, exons {ann.n_exons}", file=sys.stderr, flush=True)

    # parity sample: a slice from the middle of the stream (dense region), host copies
    want = None
    if args.parity:
        a = n_reads // 3
        b = min(n_reads, a + args.parity)
        pr = np.frombuffer(rb[a:b].cpu().numpy().reshape(-1).data, dtype=api.READ_DTYPE).copy()
        b0 = int(pr["block_off"][0]); b1 = int(pr["block_off"][-1]) + int(pr["n_blocks"][-1])
        pr["block_off"] -= np.uint32(b0)
        pb = np.frombuffer(bb[b0:b1].cpu().numpy().reshape(-1).data, dtype=api.BLOCK_DTYPE).copy()
        mk = lambda: api.make_config(max_records=len(pr), fragment_samples=1000000 if args.bed else 0)
        octx = api.Backend(os.path.join(ROOT, "oracle", "_build", "librsq_oracle.so")).create(mk())
        octx.set_annotation(ann)
        octx.submit(pr, pb)
        want = octx.finalize()
        want_bytes = octx.algorithmic_bytes()
        octx.close()

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    outf = open(os.path.join(ROOT, "gpurun_out", "ab_hot.jsonl"), "a")
    for lib in args.libs:
        row = {"lib": os.path.basename(lib), "records": n_reads}
        try:
            be = api.Backend(lib)
            if want is not None:
                ctx = be.create(mk())
                ctx.set_annotation(ann)
                ctx.submit(pr, pb)
                got = ctx.finalize()
                got_bytes = ctx.algorithmic_bytes()
                ctx.close()
                try:
                    assert_results_equal(got, want, lib)
                    assert got_bytes == want_bytes, (got_bytes, want_bytes)
                    row["parity"] = "bit-exact"
                except AssertionError as ex:
                    row["parity"] = "MISMATCH: " + str(ex)[:300]
            ctx = be.create(api.make_config(max_records=int(n_reads * args.max_records_factor), fragment_samples=1000000 if args.bed else 0,
                                            collect_coverage=not args.no_coverage))
            ctx.set_annotation(ann)
            times = []
            for k in range(args.reps + 2):
                ctx.reset()
                ctx.submit_device(rb.data_ptr(), n_reads, bb.data_ptr(), n_blocks)
                ctx.sync()
                if k >= 2:
                    times.append(ctx.kernel_time_ms())
            res = ctx.finalize(want_coverage=False)
            row.update(hot_ms_median=float(np.median(times)), hot_ms_min=float(min(times)), hot_ms_all=[round(t, 4) for t in times],
                       grec_per_s=n_reads / (float(np.median(times)) * 1e-3) / 1e9, algo_bytes=int(ctx.algorithmic_bytes()),
                       gene_reads_sum=int(res.gene_reads.sum()), exonic=res.counter("Exonic Reads"))
            ctx.close()
        except Exception as ex:  # keep going: one broken variant must not cost the whole GPU call
            row["error"] = repr(ex)[:500]
        print(json.dumps(row), flush=True)
        outf.write(json.dumps(row) + "\n"); outf.flush()


if __name__ == "__main__":
    main()
