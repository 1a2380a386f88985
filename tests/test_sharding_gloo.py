"""CPU test of the N>1 path: world sizes 2 and 3 over gloo.  Each rank runs its coordinate slice of one sorted record stream
through the C-ABI (oracle backend here; the GPU box runs the same protocol with the CUDA library and NCCL inside the library,
bench.py's pre-flight and tests/test_gpu_multi.py), the exchange of include/rsq_b200.h runs with gloo as the transport, and
every rank's result must equal the single-process run bit for bit: counters, gene and exon tables, the fragment-size histogram
under the reference's ordered first-N cut (src/RNASeQC.cpp:372), GC bins, and the per-gene coverage / 3' bias statistics."""
import ctypes
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIELDS = ("counters", "gene_reads", "gene_unique", "gene_fragments", "exon_whole", "exon_frac", "fragment_hist", "gc_hist")
DIAG = 53        # RSQ_C_FRAGMENT_CANDIDATES (see tests/util.py assert_results_equal)
STATS = ("gene_mean", "gene_std", "gene_cv", "five_end", "three_end", "exon_cv")


def _allreduce_host(ptr, n):
    import torch
    import torch.distributed as dist
    if n == 0:
        return
    buf = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_int64)), shape=(n,))
    dist.all_reduce(torch.from_numpy(buf), op=dist.ReduceOp.SUM)     # in place: the export buffer now holds the global sums


def _worker(rank, world, port, gtf, bed, out_dir, quota, lib):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from rnaseqc_b200 import api, shard, synth

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sa = synth.make_annotation(seed=21, n_genes=300, contig_names=["chr1", "chr2", "chr3"], contig_lens=[30_000_000, 20_000_000, 10_000_000])
    ann = api.Annotation(gtf, bed)
    sr = synth.simulate_reads(sa, 120000 if "oracle" in lib else 15000, seed=22)
    ranges, cuts = shard.split_sorted_records(sa, sr.contig.numpy(), sr.pos.numpy(), world, return_cuts=True)
    shard.assert_cuts_legal(ann, sa, cuts)
    rb, bb = synth.pack(sr, [ann.contig_index(n) for n in sa.contig_names])
    reads, blocks = synth.to_numpy(rb, bb)
    lo, hi = ranges[rank]
    r = reads[lo:hi].copy()
    b0 = int(r["block_off"][0]) if len(r) else 0
    b1 = int(r["block_off"][-1]) + int(r["n_blocks"][-1]) if len(r) else 0
    r["block_off"] -= b0
    be = api.Backend(os.path.join(ROOT, "oracle", "_build", lib))
    ctx = be.create(api.make_config(fragment_samples=quota))
    ctx.set_annotation(ann)
    ctx.set_partition(rank, world)
    ctx.submit(r, blocks[b0:b1].copy())
    _allreduce_host(*ctx.export_sample_counts())       # phase 1: per-partition sample counts
    _allreduce_host(*ctx.export_counts())              # phase 2: count vector + the kept samples at their global ranks
    ctx.import_counts()
    res = ctx.finalize(want_coverage=False)
    owned = shard.owned_gene_lines(ann, sa, cuts, rank)
    cs = ctx.coverage_statistics(ann, owned=owned)
    for k in STATS:                                    # owner-gather: bitwise sum, one non-zero contributor per value
        dist.all_reduce(torch.from_numpy(cs[k].view(np.int64)), op=dist.ReduceOp.SUM)
    np.savez(os.path.join(out_dir, f"sharded{rank}.npz"), ranges=np.array(ranges), taken=res.fragment_samples_taken, owned=owned,
             **{k: getattr(res, k) for k in FIELDS}, **{k: cs[k].view(np.uint64) for k in STATS})
    if rank == 0:
        one = be.create(api.make_config(fragment_samples=quota))
        one.set_annotation(ann)
        one.submit(reads, blocks)
        ref = one.finalize(want_coverage=False)
        rcs = one.coverage_statistics(ann)
        np.savez(os.path.join(out_dir, "single.npz"), taken=ref.fragment_samples_taken, **{k: getattr(ref, k) for k in FIELDS},
                 **{k: rcs[k].view(np.uint64) for k in STATS})
    dist.barrier()
    dist.destroy_process_group()


# librsq_oracle.so: the CPU restatement.  librsq_emu.so: the PRODUCT's exchange code — pack / unpack kernels with the kept samples at
# their global ranks, ownership-masked statistics kernel, rsq_capi.cu — compiled for the host (tests/cudaemu); its "device" pointers are
# host pointers, so gloo can reduce them in place.
@pytest.mark.parametrize("world,quota,lib", [(2, 1000000, "librsq_oracle.so"), (2, 700, "librsq_oracle.so"), (3, 2500, "librsq_oracle.so"),
                                             (3, 60, "librsq_emu.so")])
def test_sharded_run_matches_single_run(libs, tmp_path, world, quota, lib):
    import torch.multiprocessing as mp
    from rnaseqc_b200 import synth
    sa = synth.make_annotation(seed=21, n_genes=300, contig_names=["chr1", "chr2", "chr3"], contig_lens=[30_000_000, 20_000_000, 10_000_000])
    gtf, bed = str(tmp_path / "a.gtf"), str(tmp_path / "a.bed")
    synth.write_gtf(sa, gtf)
    assert synth.write_bed(sa, bed, min_len=400) > 20
    port = 29500 + (os.getpid() * 7 + world * 13 + quota) % 2000
    mp.spawn(_worker, args=(world, port, gtf, bed, str(tmp_path), quota, lib), nprocs=world, join=True)
    b = np.load(tmp_path / "single.npz")
    small = "emu" in lib
    assert b["gene_reads"].sum() > (2000 if small else 10000) and b["fragment_hist"].sum() > (50 if small else 500)
    assert int(b["taken"]) == min(quota, int(b["fragment_hist"].sum())) or quota > int(b["fragment_hist"].sum())
    for rank in range(world):
        a = np.load(tmp_path / f"sharded{rank}.npz")
        for lo, hi in a["ranges"]:
            assert hi > lo, "degenerate split"
        assert 0 < a["owned"].sum() < len(a["owned"])
        for k in FIELDS + STATS:
            x, y = a[k].copy(), b[k].copy()
            if k == "counters":       # a diagnostic, not a reference output: candidates are only counted while the quota lasts
                x[DIAG] = y[DIAG] = 0
            assert np.array_equal(x, y), (rank, k)
        assert int(a["taken"]) == int(b["taken"])
    if small:
        assert int(b["taken"]) == quota
    if quota == 700 and not small:      # the quota is reached inside the stream: the cut is exercised (the single run stopped sampling early)
        assert int(b["taken"]) == 700


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


def test_extra_intervals_veto_cuts(libs, tmp_path):
    """a BED interval that is not an exon (it bridges two genes) removes the cut candidates it covers; the parsed annotation's
    islands (C++ host model) agree"""
    sys.path.insert(0, ROOT)
    from rnaseqc_b200 import api, shard, synth
    sa = synth.make_annotation(seed=9, n_genes=60, contig_names=["c1"], contig_lens=[12_000_000])
    _, cuts = shard.gene_free_cuts(sa)
    victim = int(cuts[len(cuts) // 2])
    extra = (np.array([0]), np.array([victim - 50]), np.array([victim + 50]))
    _, cuts2 = shard.gene_free_cuts(sa, extra)
    assert victim in cuts and victim not in cuts2 and len(cuts2) == len(cuts) - 1
    gtf, bed = str(tmp_path / "a.gtf"), str(tmp_path / "a.bed")
    synth.write_gtf(sa, gtf)
    with open(bed, "w") as f:
        f.write(f"c1\t{victim - 51}\t{victim + 50}\n")
    ann = api.Annotation(gtf, bed)
    shard.assert_cuts_legal(ann, sa, [int(c) for c in cuts2[1:]])
    with pytest.raises(AssertionError):
        shard.assert_cuts_legal(ann, sa, [victim])
