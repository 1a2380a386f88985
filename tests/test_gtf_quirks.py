"""GTF parsing against the LIVE reference binary: the reference's Feature parser (src/GTF.cpp:40-161) re-uses one Feature object, so
several attributes carry over from line to line; ids are checked for uniqueness; attribute text is split on every ';' with quotes
ignored.  `csrc/host/annotation.cpp` reproduces those rules; here hand-written GTFs exercise them one by one — same output files,
same exit codes.  Skipped where the reference binary is not built (the goldens under tests/golden cover well-formed GTFs)."""
import os

import pytest

from util import compare_outputs, run_cli

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "rnaseqc")
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="reference binary not built (make ref)")


def L(chrom, kind, start, end, strand, attrs, track="t", score=".", frame="."):
    return f"{chrom}\t{track}\t{kind}\t{start}\t{end}\t{score}\t{strand}\t{frame}\t{attrs}"


def A(**kw):
    return " ".join(f'{k} "{v}";' for k, v in kw.items())


BASE = [
    L("chr1", "gene", 1000, 5000, "+", A(gene_id="G1", transcript_id="G1", gene_type="protein_coding", gene_name="ALPHA", transcript_type="protein_coding")),
    L("chr1", "transcript", 1000, 5000, "+", A(gene_id="G1", transcript_id="G1", gene_name="ALPHA", transcript_type="protein_coding")),
    L("chr1", "exon", 1000, 1400, "+", A(gene_id="G1", transcript_id="G1", gene_name="ALPHA", transcript_type="protein_coding", exon_id="G1_1")),
    L("chr1", "exon", 2000, 2600, "+", A(gene_id="G1", transcript_id="G1", gene_name="ALPHA", transcript_type="protein_coding", exon_id="G1_2")),
    L("chr1", "exon", 4000, 5000, "+", A(gene_id="G1", transcript_id="G1", gene_name="ALPHA", transcript_type="protein_coding", exon_id="G1_3")),
    L("chr1", "gene", 8000, 9500, "-", A(gene_id="G2", transcript_id="G2", gene_name="BETA", transcript_type="lincRNA")),
    L("chr1", "exon", 8000, 8700, "-", A(gene_id="G2", transcript_id="G2", gene_name="BETA", transcript_type="lincRNA", exon_id="G2_1")),
    L("chr1", "exon", 9000, 9500, "-", A(gene_id="G2", transcript_id="G2", gene_name="BETA", transcript_type="lincRNA", exon_id="G2_2")),
]

CASES = {
    "baseline": BASE,
    # exon lines without exon_id get "<gene_id>_<n>" names (src/GTF.cpp:92-100)
    "unnamed_exons": BASE[:2] + [
        L("chr1", "exon", 1000, 1400, "+", A(gene_id="G1", transcript_id="G1", gene_name="ALPHA", transcript_type="protein_coding")),
        L("chr1", "exon", 2000, 2600, "+", A(gene_id="G1", transcript_id="G1", gene_name="ALPHA", transcript_type="protein_coding")),
    ] + BASE[5:],
    # an exon line without gene_id / transcript_type inherits both from the line before it
    "carry_over_gene_and_type": [
        L("chr1", "gene", 1000, 5000, "+", A(gene_id="G1", gene_name="ALPHA", transcript_type="rRNA")),
        L("chr1", "exon", 1000, 1400, "+", A(exon_id="E1")),
        L("chr1", "exon", 2000, 2600, "+", A(exon_id="E2", gene_name="OTHER")),
        L("chr1", "gene", 8000, 9500, "-", A(gene_id="G2", gene_name="BETA")),          # no transcript_type: still "rRNA"
        L("chr1", "exon", 8000, 8700, "-", A(gene_id="G2", exon_id="E3", transcript_type="lincRNA")),
        L("chr1", "exon", 9000, 9500, "-", A(gene_id="G2", exon_id="E4")),
    ],
    # ';' inside quotes still splits; a piece without quotes; blanks and tabs before a key; duplicate keys (the last one wins)
    "attribute_splitting": [
        L("chr1", "gene", 1000, 5000, "+", 'gene_id "G1"; gene_name "AL;PHA"; level 2; transcript_type "protein_coding"; gene_name "ALPHA2";'),
        L("chr1", "exon", 1000, 1400, "+", 'gene_id "G1";\t  exon_id "G1_1"; tag "a;b;c"; gene_name "X"'),
        L("chr1", "exon", 2000, 2600, "+", 'gene_id "G1"; exon_id "G1_2"; gene_name ALPHA3;'),
        L("chr1", "gene", 8000, 9500, "-", 'gene_id "G2"; gene_name "HBB"; transcript_type "protein_coding";'),
        L("chr1", "exon", 8000, 8700, "-", 'gene_id "G2"; exon_id "G2_1"; gene_name "HBB";'),
    ],
    # other feature kinds are parsed (their attributes reach the carried-over state and the name table) but not kept
    "other_feature_kinds": BASE[:3] + [
        L("chr1", "CDS", 1100, 1300, "+", A(gene_id="G1", transcript_id="G1", gene_name="RENAMED", transcript_type="rRNA_pseudogene")),
        L("chr1", "UTR", 1000, 1099, "+", A(gene_id="G9", gene_name="GHOST")),
    ] + BASE[3:],
    # comments in the middle, unknown strand, extra tabs inside the attributes
    "odd_lines": ["#!genome-build test"] + BASE[:4] + ["# a comment"] + [
        L("chr1", "gene", 20000, 21000, ".", A(gene_id="G3", gene_name="GAMMA", transcript_type="protein_coding")),
        L("chr1", "exon", 20000, 21000, ".", 'gene_id "G3";\texon_id "G3_1";\tgene_name "GAMMA";'),
    ] + BASE[5:],
    # the same id on a gene and on a transcript / exon of ANOTHER kind is fine; contigs the BAM does not have
    "shared_ids_and_extra_contig": BASE + [
        L("chrZ", "gene", 100, 900, "+", A(gene_id="GZ", gene_name="ZETA", transcript_type="protein_coding")),
        L("chrZ", "exon", 100, 900, "+", A(gene_id="GZ", exon_id="GZ", gene_name="ZETA_EXON")),
    ],
}

ERRORS = {
    "duplicate_gene_id": BASE + [BASE[0]],
    "duplicate_exon_id": BASE + [BASE[2]],
    "too_few_fields": BASE + ["chr1\tt\texon\t100"],
    "no_attributes": BASE + ["chr1\tt\texon\t100\t200\t.\t+\t."],
    "start_not_a_number": BASE + [L("chr1", "exon", "abc", 200, "+", A(gene_id="G1", exon_id="Q"))],
    "exon_without_ids": [L("chr1", "exon", 100, 200, "+", A(gene_name="N"))] + BASE,
    "empty_line": BASE[:3] + [""] + BASE[3:],
    "no_exons": [BASE[0], BASE[5]],
    "exon_end_before_start": BASE + [L("chr1", "exon", 4000, 3990, "+", A(gene_id="G1", gene_name="ALPHA", exon_id="G1_bad"))],
    "gene_end_before_start": BASE + [L("chr1", "gene", 30000, 29000, "+", A(gene_id="G7", gene_name="ETA"))],
}


def _sam(path):
    rows = ["@HD\tVN:1.0\tSO:coordinate", "@SQ\tSN:chr1\tLN:100000", "@SQ\tSN:chr2\tLN:50000"]
    n = 0
    for pos in list(range(900, 5200, 37)) + list(range(7900, 9700, 41)) + list(range(19900, 21100, 53)):
        for flag, mpos_d in ((99, 150), (147, -150)):
            n += 1
            rows.append(f"r{n // 2}\t{flag}\tchr1\t{pos if flag == 99 else pos + 150}\t255\t76M\t=\t{pos + 150 if flag == 99 else pos}\t{226 if flag == 99 else -226}\t*\t*\tNM:i:0")
    rows_sorted = rows[:3] + sorted(rows[3:], key=lambda r: int(r.split("\t")[3]))
    open(path, "w").write("\n".join(rows_sorted) + "\n")


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("legacy", [False, True], ids=["current", "legacy"])
def test_gtf_quirks_match_the_reference(libs, tmp_path, name, legacy):
    d = str(tmp_path)
    gtf, sam = os.path.join(d, "q.gtf"), os.path.join(d, "q.sam")
    open(gtf, "w").write("\n".join(CASES[name]) + "\n")
    _sam(sam)
    flags = ["--coverage"] + (["--legacy"] if legacy else [])
    pr = run_cli(REF, gtf, sam, os.path.join(d, "ref"), *flags)
    po = run_cli(libs["oracle_cli"], gtf, sam, os.path.join(d, "ours"), *flags)
    identical = set(compare_outputs(os.path.join(d, "ref"), os.path.join(d, "ours")))
    assert {"q.sam.gene_reads.gct", "q.sam.exon_reads.gct", "q.sam.gene_tpm.gct", "q.sam.coverage.tsv"} <= identical, identical
    # the parser's diagnostics (unnamed exons, bad ranges) are the same text
    pick = lambda p: [l for l in p.stderr.splitlines() if l.startswith(("Unnamed exon", "Bad feature range"))]
    assert pick(pr) == pick(po)


@pytest.mark.parametrize("name", sorted(ERRORS))
def test_gtf_errors_exit_like_the_reference(libs, tmp_path, name):
    d = str(tmp_path)
    gtf, sam = os.path.join(d, "q.gtf"), os.path.join(d, "q.sam")
    open(gtf, "w").write("\n".join(ERRORS[name]) + "\n")
    _sam(sam)
    pr = run_cli(REF, gtf, sam, os.path.join(d, "ref"), check=False)
    po = run_cli(libs["oracle_cli"], gtf, sam, os.path.join(d, "ours"), check=False)
    assert pr.returncode == po.returncode, (pr.returncode, po.returncode, pr.stderr[-300:], po.stderr[-300:])
    if pr.returncode == 0:
        compare_outputs(os.path.join(d, "ref"), os.path.join(d, "ours"))


def test_gtf_without_trailing_newline_and_with_crlf(libs, tmp_path):
    d = str(tmp_path)
    _sam(os.path.join(d, "q.sam"))
    for k, text in enumerate(("\n".join(BASE), "\r\n".join(BASE) + "\r\n")):
        gtf = os.path.join(d, f"q{k}.gtf")
        open(gtf, "w", newline="").write(text)
        run_cli(REF, gtf, os.path.join(d, "q.sam"), os.path.join(d, f"ref{k}"))
        run_cli(libs["oracle_cli"], gtf, os.path.join(d, "q.sam"), os.path.join(d, f"ours{k}"))
        identical = set(compare_outputs(os.path.join(d, f"ref{k}"), os.path.join(d, f"ours{k}")))
        assert {"q.sam.gene_reads.gct", "q.sam.exon_reads.gct"} <= identical, identical


BEDS = {
    "baseline": ["chr1\t900\t5300", "chr1\t7900\t9800"],
    "named_and_extra_columns": ["chr1\t900\t5300\tfirst\t0\t+", "chr1\t7900\t9800\tsecond"],
    "overlapping": ["chr1\t900\t3000", "chr1\t2500\t5300", "chr1\t7900\t9800", "chr1\t7900\t9800"],
    "other_contigs": ["chrQ\t10\t20000", "chr1\t900\t5300", "chr2\t1\t40000"],
    "abutting_and_tiny": ["chr1\t900\t2000", "chr1\t2000\t2001", "chr1\t2001\t5300"],
    "end_before_start": ["chr1\t5300\t900", "chr1\t7900\t9800"],
    "trailing_tab_and_spaces": ["chr1\t900\t5300\t", "chr1\t7900\t9800\t \t"],
}
BED_ERRORS = {
    "start_not_a_number": ["chr1\tabc\t5300"],
    "two_columns": ["chr1\t900", "chr1\t7900\t9800"],
    "numeric_contig_alone": ["7", "chr1\t900\t5300"],
    "one_column": ["chr1"],
    "empty_line": ["chr1\t900\t5300", "", "chr1\t7900\t9800"],
    "track_line": ["track name=x", "chr1\t900\t5300"],
    "comment_line": ["# c", "chr1\t900\t5300"],
    "negative_start": ["chr1\t-5\t5300"],
    "space_separated": ["chr1 900 5300"],
}


@pytest.mark.parametrize("name", sorted(BEDS) + sorted(BED_ERRORS))
def test_bed_files_match_the_reference(libs, tmp_path, name):
    d = str(tmp_path)
    gtf, sam, bed = os.path.join(d, "q.gtf"), os.path.join(d, "q.sam"), os.path.join(d, "q.bed")
    open(gtf, "w").write("\n".join(BASE) + "\n")
    open(bed, "w").write("\n".join(BEDS.get(name) or BED_ERRORS[name]) + "\n")
    _sam(sam)
    pr = run_cli(REF, gtf, sam, os.path.join(d, "ref"), "--bed", bed, check=False)
    po = run_cli(libs["oracle_cli"], gtf, sam, os.path.join(d, "ours"), "--bed", bed, check=False)
    assert pr.returncode == po.returncode, (pr.returncode, po.returncode, pr.stderr[-300:], po.stderr[-300:])
    if name in BEDS:
        assert pr.returncode == 0
    if pr.returncode == 0:
        identical = set(compare_outputs(os.path.join(d, "ref"), os.path.join(d, "ours")))
        assert {"q.sam.gene_reads.gct", "q.sam.fragmentSizes.txt"} <= identical, identical


def test_alignment_file_errors_keep_their_place_behind_the_annotation(libs, tmp_path):
    """The CLI starts the BAM ingest before it parses the GTF (csrc/host/pipeline.cpp); what goes wrong with the alignment file is still
    reported where the reference opens it — after the GTF and the BED (src/RNASeQC.cpp:105-219) — with the reference's exit codes."""
    import gzip
    import shutil
    from rnaseqc_b200 import synth
    d = str(tmp_path)
    gtf = os.path.join(d, "a.gtf")
    with open(gtf, "w") as f:
        f.write("\n".join(BASE) + "\n")
    bad_gtf = os.path.join(d, "bad.gtf")
    with open(bad_gtf, "w") as f:
        f.write("\n".join(BASE + [BASE[0]]) + "\n")                      # gene id twice: exit 11
    sa = synth.make_annotation(seed=5, n_genes=20, contig_names=["chr1"], contig_lens=[1_000_000])
    good = os.path.join(d, "good.bam")
    synth.write_bam(sa, synth.simulate_reads(sa, 500, seed=6), good, level=1)
    not_bam = os.path.join(d, "text.bam")                                 # gzip magic, no BAM inside
    with gzip.open(not_bam, "wb") as f:
        f.write(b"hello, this is not a BAM file\n" * 100)
    truncated = os.path.join(d, "cut.bam")
    shutil.copy(good, truncated)
    with open(truncated, "r+b") as f:
        f.truncate(os.path.getsize(good) // 2)
    missing = os.path.join(d, "nothing_here.bam")
    for k, (g, b) in enumerate([(gtf, good), (gtf, missing), (bad_gtf, missing), (bad_gtf, not_bam), (gtf, not_bam), (bad_gtf, good), (gtf, truncated)]):
        pr = run_cli(REF, g, b, os.path.join(d, f"ref{k}"), check=False)
        for threads in ("0", "3"):
            po = run_cli(libs["oracle_cli"], g, b, os.path.join(d, f"ours{k}_{threads}"), "--decode-threads", threads, check=False)
            if os.path.basename(b) in ("text.bam", "cut.bam") and os.path.basename(g) == "a.gtf":
                assert (pr.returncode != 0) == (po.returncode != 0), (g, b, threads, pr.returncode, po.returncode)   # htslib words its own failures
            else:
                assert pr.returncode == po.returncode, (g, b, threads, pr.returncode, po.returncode, pr.stderr[-300:], po.stderr[-300:])
        if pr.returncode == 0:
            compare_outputs(os.path.join(d, f"ref{k}"), os.path.join(d, f"ours{k}_3"))
