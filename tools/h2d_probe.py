#!/usr/bin/env python
"""Aggregate host->device copy ceiling of the box: N processes (one per GPU) copy a pinned 1 GiB buffer to their GPU in a loop,
all at once; prints per-GPU and aggregate GB/s for N = 1, 2, 4, 8 with the ranks unbound and bound to NVML's GPU-local cores.
    python tools/h2d_probe.py [--gib 1] [--reps 12]
"""
import argparse, json, os, sys, time
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def worker(rank, world, gib, reps, bind, q, barrier):
    sys.path.insert(0, ROOT)
    info = {}
    if bind:
        import bench
        info = bench.bind_to_gpu_cores(rank, world)
    torch.cuda.set_device(rank)
    n = int(gib * (1 << 30))
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h.fill_(1)
    d = torch.empty(n, dtype=torch.uint8, device=f"cuda:{rank}")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    barrier.wait()
    t = time.perf_counter()
    with torch.cuda.stream(s):
        for _ in range(reps):
            d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    q.put((rank, n * reps / dt / 1e9, info.get("cores")))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=12)
    a = ap.parse_args()
    ngpu = torch.cuda.device_count()
    out = []
    ctx = mp.get_context("spawn")
    for world in [w for w in (1, 2, 4, 8) if w <= ngpu]:
        for bind in (False, True):
            q, barrier = ctx.Queue(), ctx.Barrier(world)
            ps = [ctx.Process(target=worker, args=(r, world, a.gib, a.reps, bind, q, barrier)) for r in range(world)]
            [p.start() for p in ps]
            rows = sorted(q.get() for _ in range(world))
            [p.join() for p in ps]
            row = {"gpus": world, "bound": bind, "aggregate_gb_s": sum(r[1] for r in rows), "per_gpu_gb_s": [round(r[1], 1) for r in rows],
                   "cores": [r[2] for r in rows] if bind else None}
            print(json.dumps(row), flush=True)
            out.append(row)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "h2d_probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
