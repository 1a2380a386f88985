"""CRAM input: this repo's own CRAM reader (rnaseqc_b200/csrc/host/cramreader.cpp) against the unmodified
reference binary, which reads CRAM through htslib 1.5 and REGENERATES NM from the read features and the reference sequence.

* golden (runs anywhere): tests/golden/cram_small — a CRAM written by the reference's own htslib (multi-slice containers, records
  whose sequences follow a genome with N runs and lower-case bases, soft clips, insertions, deletions, mates on other contigs),
  with the outputs of the reference binary on it (tests/golden/make_cram_golden.py);
* live (build container only, where oracle/_ref/ exists): a fresh seed, other slice sizes and flag sets."""
import gzip
import os
import shutil
import subprocess
import sys

import pytest

from util import compare_outputs, run_cli, unpack_case

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_cram_golden as mk  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _golden(tmp_path):
    d = unpack_case("cram_small", tmp_path)
    return d


def test_oracle_cli_reads_cram_like_the_reference(libs, tmp_path):
    d = _golden(tmp_path)
    out = os.path.join(d, "out")
    run_cli(libs["oracle_cli"], os.path.join(d, "c.gtf"), os.path.join(d, "c.cram"), out, "--fasta", os.path.join(d, "c.fa"),
            "--coverage", "--bed", os.path.join(d, "c.bed"))
    same = compare_outputs(out, os.path.join(d, "expected"))
    assert "c.cram.gene_reads.gct" in same and "c.cram.gene_fragments.gct" in same and "c.cram.fragmentSizes.txt" in same
    metrics = dict(l.rstrip("\n").split("\t") for l in open(os.path.join(out, "c.cram.metrics.tsv")))
    assert int(metrics["High Quality Reads"]) > 1000 and int(metrics["Mismatched Bases"]) > 1000      # NM was regenerated


@pytest.mark.parametrize("threads", [0, 1, 3, 16])
def test_container_decode_threads_do_not_change_the_output(libs, tmp_path, threads):
    # containers are decoded concurrently and handed back in file order: same bytes out for any thread count
    d = _golden(tmp_path)
    out = os.path.join(d, "out")
    run_cli(libs["oracle_cli"], os.path.join(d, "c.gtf"), os.path.join(d, "c.cram"), out, "--fasta", os.path.join(d, "c.fa"),
            "--coverage", "--bed", os.path.join(d, "c.bed"), "--decode-threads", str(threads))
    same = compare_outputs(out, os.path.join(d, "expected"))
    assert "c.cram.gene_reads.gct" in same and "c.cram.fragmentSizes.txt" in same


@pytest.mark.parametrize("threads", [0, 4])
def test_truncated_cram_is_an_error(libs, tmp_path, threads):
    d = _golden(tmp_path)
    raw = open(os.path.join(d, "c.cram"), "rb").read()
    open(os.path.join(d, "t.cram"), "wb").write(raw[: len(raw) * 2 // 3])
    p = run_cli(libs["oracle_cli"], os.path.join(d, "c.gtf"), os.path.join(d, "t.cram"), os.path.join(d, "out"), "--fasta",
                os.path.join(d, "c.fa"), "--decode-threads", str(threads), check=False)
    assert p.returncode != 0


def test_cram_without_a_reference_fails_like_the_reference(libs, tmp_path):
    d = _golden(tmp_path)
    p = run_cli(libs["oracle_cli"], os.path.join(d, "c.gtf"), os.path.join(d, "c.cram"), os.path.join(d, "out"), check=False)
    assert p.returncode != 0 and "reference" in p.stderr.lower()


# sam2cram arguments: records per slice, slices per container, version, lossy names, multi-reference slices, no rANS
LIVE = [
    ("default_slices", [], ["--coverage"]),
    ("small_slices_stranded", ["333", "3"], ["--stranded", "rf", "--coverage", "--bed", "c.bed"]),
    ("one_slice_unpaired", ["100000", "1"], ["-u", "--exclude-chimeric"]),
    ("cram_2_1", ["1000", "2", "2.1"], ["--coverage", "--bed", "c.bed"]),
    ("lossy_names_multi_ref", ["700", "1", "3.0", "1", "1"], ["--bed", "c.bed"]),
    ("gzip_only", ["1500", "1", "3.0", "0", "0", "1"], []),
]


@pytest.mark.parametrize("name,slicing,flags", LIVE, ids=[c[0] for c in LIVE])
def test_live_reference_on_fresh_cram(libs, tmp_path, name, slicing, flags):
    ref_cli, sam2cram = libs["ref_cli"], os.path.join(ROOT, "oracle", "_ref", "sam2cram")
    if not (os.path.exists(ref_cli) and os.path.exists(sam2cram)):
        pytest.skip("the reference binary / its htslib are only built in the build container")
    from rnaseqc_b200 import synth
    d = str(tmp_path)
    sa = synth.make_annotation(seed=900 + len(name), n_genes=80, contig_names=["chr1", "chr2", "chrM"], contig_lens=[200000, 120000, 16569])
    synth.write_gtf(sa, os.path.join(d, "c.gtf"))
    synth.write_bed(sa, os.path.join(d, "c.bed"), min_len=300)
    synth.write_fasta(sa, os.path.join(d, "c.fa"), seed=6)
    sr = synth.simulate_reads(sa, 9000, seed=901 + len(name))
    synth.write_sam(sa, sr, os.path.join(d, "plain.sam"))
    mk.realistic_sam(os.path.join(d, "plain.sam"), os.path.join(d, "c.sam"), mk.read_fasta(os.path.join(d, "c.fa")), seed=len(name))
    subprocess.run([sam2cram, os.path.join(d, "c.sam"), os.path.join(d, "c.fa"), os.path.join(d, "c.cram")] + slicing, check=True,
                   stderr=subprocess.DEVNULL)
    for cli, out in ((ref_cli, "ref"), (libs["oracle_cli"], "mine")):
        subprocess.run([cli, "c.gtf", "c.cram", out, "--fasta", "c.fa"] + flags, cwd=d, check=True, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL)
    same = compare_outputs(os.path.join(d, "mine"), os.path.join(d, "ref"))
    assert "c.cram.gene_reads.gct" in same
