"""-m gpu, needs >= 2 GPUs on the box (skipped otherwise; `gpurun --gpus 2`): the coordinate-partitioned run of the CUDA library
with NCCL INSIDE the library (rsq_comm_init / rsq_exchange / collective rsq_coverage_statistics) equals the single-GPU run and
the CPU oracle bit for bit — counters, gene and exon tables, fragment-size histogram under the ordered first-N cut, coverage
statistics.  The process-group plumbing of the test is gloo on purpose: the library brings its own communicator."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir, quota):
    import json
    import tempfile

    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import bench
    from rnaseqc_b200 import api
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(rank)
    wl = dict(n_genes=1500, full_genome=False)
    d = tempfile.mkdtemp(prefix="rsq_multi_")
    sa, ann, _ = bench.build_annotation(wl, d, with_bed=True)
    summary = bench.preflight(api.default_backend(), ann, sa, rank, rank, world, with_bed=True, n_records=600_000, quota=quota,
                              log_fn=lambda *a: None)
    if rank == 0:
        json.dump(summary, open(os.path.join(out_dir, "summary.json"), "w"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("quota", [1000000, 3000])
def test_partitioned_cuda_run_matches_single_gpu_and_oracle(libs, tmp_path, quota):
    import json

    import torch
    import torch.multiprocessing as mp
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = min(n, 4)
    port = 29500 + (os.getpid() * 3 + quota) % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path), quota), nprocs=world, join=True)
    s = json.load(open(tmp_path / "summary.json"))
    assert s["gene_reads_sum"] > 50000 and s["fragment_samples_taken"] > 0
    assert all(x > 0 for x in s["partition_sizes"]) and len(s["partition_sizes"]) == world
    if quota == 3000:
        assert s["fragment_samples_taken"] == 3000          # the quota ends inside the stream: the cross-rank cut is exercised
    assert "vs_cpu_oracle" in s
