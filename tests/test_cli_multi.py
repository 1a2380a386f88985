"""The CLI's multi-GPU mode (--gpus N: one context per GPU, the coordinate-sorted stream routed lane by lane at island-free
cuts, end-of-run exchange, owner-gathered coverage statistics).  On CPU the test harness binary (the same host code over the
oracle library) runs the lanes with the exchange buffers travelling through host memory; under -m gpu the product binary runs
them on real GPUs with NCCL inside the library.  Either way all report files must equal the single-lane run's byte for byte."""
import os
import subprocess

import pytest

from util import compare_outputs, run_cli, unpack_case


def _run_and_compare(cli, d, flags, multi, n_files):
    one, many = os.path.join(d, "out_one"), os.path.join(d, "out_many")
    run_cli(cli, os.path.join(d, "a.gtf"), os.path.join(d, "a.sam"), one, *flags)
    p = run_cli(cli, os.path.join(d, "a.gtf"), os.path.join(d, "a.sam"), many, "-v", *flags, *multi)
    lanes = [l for l in p.stdout.splitlines() if l.startswith("GPU ")]
    assert len(os.listdir(one)) == n_files
    assert sorted(compare_outputs(one, many)) == sorted(os.listdir(one)), "multi-lane run differs from the single-lane run"
    return lanes


@pytest.mark.parametrize("n,quota", [(2, None), (3, 300), (4, None)])
def test_multi_lane_cli_equals_single_lane_on_cpu(libs, tmp_path, n, quota):
    d = unpack_case("synth_small", tmp_path)
    flags = ["--coverage", "--bed", os.path.join(d, "a.bed")] + (["--fragment-samples", str(quota)] if quota else [])
    lanes = _run_and_compare(libs["oracle_cli"], d, flags, ["--gpus", str(n), "--records-per-gpu", str(20000 // n), "--batch-records", "1500"], 8)
    assert len(lanes) == n and sum(int(l.split()[2]) > 0 for l in lanes) >= min(n, 3), lanes      # the stream really was spread out
    if quota is None and n == 2:      # and the reference's golden too
        assert len(compare_outputs(os.path.join(d, "expected", "coverage_bed"), os.path.join(d, "out_many"))) == 8


def test_multi_lane_cli_fasta_and_stranded(libs, tmp_path):
    d = unpack_case("synth_fasta", tmp_path)
    flags = ["--coverage", "--stranded", "rf", "--bed", os.path.join(d, "a.bed"), "--fasta", os.path.join(d, "g.fa")]
    _run_and_compare(libs["oracle_cli"], d, flags, ["--gpus", "3", "--records-per-gpu", "2500", "--batch-records", "700"], 9)


def test_multi_lane_product_cli_in_emulation(libs, tmp_path, monkeypatch):
    """the PRODUCT's side of a partitioned run — rsq_set_partition, the exchange packers with the kept samples at their global ranks,
    the ownership-masked statistics kernel — under the same CLI, through the CPU emulation of the CUDA library (tests/cudaemu; its
    exchange buffers are host memory, so they travel like the oracle's)"""
    emu_cli = os.path.join(os.path.dirname(libs["oracle_cli"]), "rnaseqc_emu")
    if not os.path.exists(emu_cli):
        pytest.skip("rnaseqc_emu is not built")
    monkeypatch.setenv("RSQ_HOST_EXCHANGE", "1")
    d = unpack_case("kat_odd", tmp_path)
    flags = ["--coverage", "--coverage-mask", "0", "--gene-length", "50", "--bed", os.path.join(d, "a.bed"), "--fragment-samples", "60"]
    lanes = _run_and_compare(emu_cli, d, flags, ["--gpus", "2", "--records-per-gpu", "150", "--batch-records", "64"], 8)
    assert sum(int(l.split()[2]) > 0 for l in lanes) >= 2, lanes


def test_multi_lane_cli_refuses_unsorted_input(libs, tmp_path):
    d = unpack_case("synth_small", tmp_path)
    lines = open(os.path.join(d, "a.sam")).read().splitlines()
    head = [l for l in lines if l.startswith("@")]
    body = [l for l in lines if not l.startswith("@")]
    body[2000], body[9000] = body[9000], body[2000]
    with open(os.path.join(d, "u.sam"), "w") as f:
        f.write("\n".join(head + body) + "\n")
    p = subprocess.run([libs["oracle_cli"], os.path.join(d, "a.gtf"), os.path.join(d, "u.sam"), os.path.join(d, "o"), "--gpus", "2",
                        "--records-per-gpu", "5000", "--batch-records", "1000"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 6 and "coordinate-sorted" in p.stderr


@pytest.mark.gpu
def test_multi_gpu_product_cli_equals_single_gpu(libs, tmp_path):
    import torch
    n = min(torch.cuda.device_count(), 4)
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    d = unpack_case("synth_small", tmp_path)
    flags = ["--coverage", "--bed", os.path.join(d, "a.bed"), "--fragment-samples", "300"]
    lanes = _run_and_compare(libs["cli"], d, flags, ["--gpus", str(n), "--records-per-gpu", str(20000 // n), "--batch-records", "1500"], 8)
    assert len(lanes) == n and sum(int(l.split()[2]) > 0 for l in lanes) >= min(n, 3), lanes
    for f in ("a.sam.gene_reads.gct", "a.sam.gene_fragments.gct"):      # and the reference's golden (the tables the quota does not touch)
        assert open(os.path.join(d, "expected", "coverage_bed", f)).read() == open(os.path.join(d, "out_many", f)).read(), f
