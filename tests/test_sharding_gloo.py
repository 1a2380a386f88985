"""CPU test of the N>1 path: world size 2 over gloo.  Each rank runs its coordinate slice of one sorted record stream
through the C-ABI (oracle backend here; the GPU box runs the same flow with the CUDA library and NCCL in bench.py), the
exported count vectors are all-reduced, and the result must equal the single-process run bit for bit."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, gtf, out_dir):
    import ctypes

    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from rnaseqc_b200 import api, shard, synth

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sa = synth.make_annotation(seed=21, n_genes=300, contig_names=["chr1", "chr2", "chr3"], contig_lens=[30_000_000, 20_000_000, 10_000_000])
    ann = api.Annotation(gtf)
    sr = synth.simulate_reads(sa, 120000, seed=22)
    ranges = shard.split_sorted_records(sa, sr.contig.numpy(), sr.pos.numpy(), world)
    rb, bb = synth.pack(sr, [ann.contig_index(n) for n in sa.contig_names])
    reads, blocks = synth.to_numpy(rb, bb)
    lo, hi = ranges[rank]
    r = reads[lo:hi].copy()
    b0 = int(r["block_off"][0]) if len(r) else 0
    b1 = int(r["block_off"][-1]) + int(r["n_blocks"][-1]) if len(r) else 0
    r["block_off"] -= b0
    be = api.Backend(os.path.join(ROOT, "oracle", "_build", "librsq_oracle.so"))
    ctx = be.create(api.make_config())
    ctx.set_annotation(ann)
    ctx.submit(r, blocks[b0:b1].copy())
    ptr, n = ctx.export_counts()                      # host pointer in the oracle, device pointer in the product
    buf = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_int64)), shape=(n,))
    t = torch.from_numpy(buf)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)          # in place: the export buffer now holds the global sums
    ctx.import_counts()
    res = ctx.finalize(want_coverage=False)
    if rank == 0:
        np.savez(os.path.join(out_dir, "sharded.npz"), counters=res.counters, gene_reads=res.gene_reads, gene_unique=res.gene_unique,
                 gene_fragments=res.gene_fragments, exon_whole=res.exon_whole, exon_frac=res.exon_frac, ranges=np.array(ranges))
        one = be.create(api.make_config())
        one.set_annotation(ann)
        one.submit(reads, blocks)
        ref = one.finalize(want_coverage=False)
        np.savez(os.path.join(out_dir, "single.npz"), counters=ref.counters, gene_reads=ref.gene_reads, gene_unique=ref.gene_unique,
                 gene_fragments=ref.gene_fragments, exon_whole=ref.exon_whole, exon_frac=ref.exon_frac)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_run(libs, tmp_path):
    import torch.multiprocessing as mp
    from rnaseqc_b200 import synth
    sa = synth.make_annotation(seed=21, n_genes=300, contig_names=["chr1", "chr2", "chr3"], contig_lens=[30_000_000, 20_000_000, 10_000_000])
    gtf = str(tmp_path / "a.gtf")
    synth.write_gtf(sa, gtf)
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, gtf, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "sharded.npz"), np.load(tmp_path / "single.npz")
    lo, hi = a["ranges"][0]
    assert 0 < hi < 120000, "degenerate split"
    for k in ("counters", "gene_reads", "gene_unique", "gene_fragments", "exon_whole", "exon_frac"):
        assert np.array_equal(a[k], b[k]), k
    assert b["gene_reads"].sum() > 10000


def test_cut_points_are_gene_free():
    sys.path.insert(0, ROOT)
    from rnaseqc_b200 import shard, synth
    sa = synth.make_annotation(seed=5, n_genes=2000, contig_names=["c1", "c2"], contig_lens=[200_000_000, 100_000_000])
    coff = np.concatenate([[0], np.cumsum(sa.contig_lens)])
    for world in (2, 4, 8):
        regs = [shard.rank_region(sa, r, world) for r in range(world)]
        assert regs[0]["gene_lo"] == 0 and regs[-1]["gene_hi"] == sa.n_genes
        for r0, r1 in zip(regs[:-1], regs[1:]):
            assert r0["gene_hi"] == r1["gene_lo"] and r0["lin_hi"] == r1["lin_lo"]
            cut = r0["lin_hi"]
            # no gene spans the cut
            lin_s = coff[sa.gene_contig] + sa.gene_start
            lin_e = coff[sa.gene_contig] + sa.gene_end
            assert not np.any((lin_s <= cut) & (lin_e >= cut)), (world, cut)
