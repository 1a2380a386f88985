"""CPU tests against the LIVE reference binary (oracle/_ref/rnaseqc, built by `make ref` where /root/reference exists):
fresh seeded synthetic inputs, several flag sets, oracle-backed CLI vs reference files.  Skipped on boxes without the
binary (the committed goldens under tests/golden cover those)."""
import os

import pytest

from util import compare_outputs, run_cli

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "rnaseqc")
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="reference binary not built (make ref)")

CASES = [
    ("plain", 61, []),
    ("stranded_rf_bed", 62, ["--stranded", "rf", "--bed", "BED", "--fragment-samples", "500"]),
    ("unpaired_excl", 63, ["-u", "--exclude-chimeric", "--coverage", "--coverage-mask", "100"]),
    ("legacy_coverage", 65, ["--legacy", "--coverage"]),
    ("legacy_rf_bed_excl", 66, ["--legacy", "--stranded", "rf", "--bed", "BED", "--exclude-chimeric"]),
    ("fasta_coverage", 67, ["--fasta", "FASTA", "--coverage"]),
    ("fasta_fr_bed_legacy", 68, ["--fasta", "FASTA", "--stranded", "fr", "--bed", "BED", "--legacy"]),
    ("loose", 64, ["-q", "3", "--base-mismatch", "3", "--chimeric-distance", "5000", "--detection-threshold", "2", "--gene-length", "150", "--window-size", "60"]),
]


@pytest.mark.parametrize("name,seed,flags", CASES, ids=[c[0] for c in CASES])
def test_oracle_cli_vs_live_reference(libs, tmp_path, name, seed, flags):
    from rnaseqc_b200 import synth
    d = str(tmp_path)
    sa = synth.make_annotation(seed=seed, n_genes=120, contig_names=["chr1", "chr2", "chrM"], contig_lens=[15_000_000, 9_000_000, 16569])
    gtf, bed, sam = os.path.join(d, "a.gtf"), os.path.join(d, "a.bed"), os.path.join(d, "a.sam")
    synth.write_gtf(sa, gtf)
    synth.write_bed(sa, bed, min_len=400)
    sr = synth.simulate_reads(sa, 40000, seed=seed + 100, dup_frac=0.0)   # no duplicates: skips the reference's 1e9-step complexity scan
    synth.write_sam(sa, sr, sam)
    fasta = os.path.join(d, "g.fa")
    if "FASTA" in flags:
        synth.write_fasta(sa, fasta, seed=seed, contigs=[2, 0], extra=[("chrUn", 4000)])      # chr2 absent from the FASTA
    flags = [bed if f == "BED" else fasta if f == "FASTA" else f for f in flags]
    run_cli(REF, gtf, sam, os.path.join(d, "ref"), *flags)
    run_cli(libs["oracle_cli"], gtf, sam, os.path.join(d, "ours"), *flags)
    identical = set(compare_outputs(os.path.join(d, "ref"), os.path.join(d, "ours")))
    # metrics.tsv: every integer row exact and every float row within 1e-6 is already enforced by compare_outputs; byte identity
    # is not required because of the reference's out-of-bounds percentile rows (see util.compare_outputs)
    assert {"a.sam.gene_reads.gct", "a.sam.gene_fragments.gct", "a.sam.exon_reads.gct"} <= identical, identical
    if "FASTA" in CASES[[c[0] for c in CASES].index(name)][2]:
        assert "a.sam.gc_content.tsv" in identical, identical
