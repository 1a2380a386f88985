"""Random SAM records against the LIVE reference binary: flags in any combination, CIGARs over all nine operations, NM / chimeric /
filter tags present, absent or oddly typed, mates on other contigs, unmapped records with and without coordinates, contigs the GTF
does not know.  Both binaries read the same file; every output file must agree (util.compare_outputs).  This pins the decode side
(csrc/host/alnreader.cpp: what the loop sees of a record) and the oracle's filter cascade on inputs no aligner would write.
Skipped where the reference binary is not built."""
import os
import random

import pytest

from util import compare_outputs, run_cli

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "rnaseqc")
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="reference binary not built (make ref)")


def _gtf(path, rng):
    rows = []
    g = 0
    for chrom, span in (("chr1", 60000), ("chr2", 30000), ("chrM", 8000)):
        p = 500
        while p < span - 3000:
            g += 1
            strand = rng.choice("+-")
            ttype = rng.choice(["protein_coding"] * 6 + ["rRNA", "Mt_rRNA", "lincRNA"])
            gs = p
            n_ex = rng.randint(1, 5)
            exons, q = [], gs
            for _ in range(n_ex):
                ln = rng.randint(40, 600)
                exons.append((q, q + ln))
                q += ln + rng.choice([1, 2, 80, 300, 1200])
            ge = exons[-1][1]
            name = rng.choice(["HBB", "HBA1"]) if rng.random() < 0.05 else f"N{g}"
            common = f'gene_id "G{g}"; transcript_id "G{g}"; gene_type "{ttype}"; gene_name "{name}"; transcript_type "{ttype}";'
            rows.append(f"{chrom}\tf\tgene\t{gs}\t{ge}\t.\t{strand}\t.\t{common}")
            for k, (s, e) in enumerate(exons):
                rows.append(f'{chrom}\tf\texon\t{s}\t{e}\t.\t{strand}\t.\t{common} exon_id "G{g}_{k}";')
            p = gs + rng.choice([30, 200, ge - gs + 1, ge - gs + 500, ge - gs + 4000])      # overlapping genes are common
    open(path, "w").write("\n".join(rows) + "\n")
    return rows


def _cigar(rng):
    """A CIGAR htslib accepts: hard clips outermost, soft clips next, at least one base-consuming operation inside."""
    mid = []
    for _ in range(rng.choice([1, 1, 1, 2, 3, 3, 5])):
        op = rng.choice("MMMMMM=XIDNNP")
        if mid and mid[-1][1] == op and op in "NDIP":
            op = "M"
        mid.append((rng.choice([1, 2, 10, 25, 50, 76, 101, 400, 2000]) if op in "ND" else rng.choice([1, 2, 5, 20, 38, 76]), op))
    if mid[0][1] in "NDP":
        mid.insert(0, (rng.choice([1, 30]), "M"))
    if mid[-1][1] in "NDP":
        mid.append((rng.choice([1, 30]), "M"))
    parts = list(mid)
    if rng.random() < 0.3:
        parts.insert(0, (rng.randint(1, 12), "S"))
    if rng.random() < 0.3:
        parts.append((rng.randint(1, 12), "S"))
    if rng.random() < 0.1:
        parts.insert(0, (rng.randint(1, 9), "H"))
    if rng.random() < 0.1:
        parts.append((rng.randint(1, 9), "H"))
    qlen = sum(n for n, op in parts if op in "MIS=X")
    return "".join(f"{n}{op}" for n, op in parts), qlen


def _sam(path, rng, n, allow_dup):
    contigs = [("chr1", 60000), ("chr2", 30000), ("chrM", 8000), ("chrUn", 5000)]
    rows = []
    bits = [1, 2, 8, 16, 32, 64, 128, 256, 512, 2048] + ([1024] if allow_dup else [])
    for i in range(n):
        ci = rng.choice([0, 0, 0, 1, 1, 2, 3])
        chrom, span = contigs[ci]
        pos = rng.randint(1, span)
        kind = rng.random()
        if kind < 0.5:
            flag = rng.choice([99, 147, 83, 163, 0, 16, 65, 129, 97, 145, 73, 133, 89])
        else:
            flag = 0
            for b in bits:
                if rng.random() < 0.25:
                    flag |= b
        cigar, qlen = _cigar(rng)
        if rng.random() < 0.06:
            flag |= 4
            if rng.random() < 0.5:
                cigar, qlen = "*", rng.choice([0, 50])
        rnext = rng.choice(["=", "=", "=", "=", "*", "chr2", "chrUn", "chr1"])
        pnext = 0 if rnext == "*" else max(1, pos + rng.choice([0, 150, -150, 300, 5000, 2_500_000, -40]))
        tlen = rng.choice([0, 226, -226, 150, -150, 999, 1001, 100, 101, 5000])
        mapq = rng.choice([255, 255, 255, 60, 3, 0, 254])
        tags = []
        t = rng.random()
        if t < 0.6:
            tags.append(f"NM:i:{rng.choice([0, 0, 1, 5, 6, 7, 30])}")
        elif t < 0.7:
            tags.append("NM:Z:3")
        elif t < 0.75:
            tags.append("NM:f:2.0")
        elif t < 0.8:
            tags.append(f"NM:A:{rng.choice('07')}")
        if rng.random() < 0.05:
            tags.append(rng.choice(["ch:Z:chr2", "ch:i:1", "ch:A:x"]))
        if rng.random() < 0.1:
            tags.append(rng.choice(["XF:Z:bad", "XF:i:1"]))
        if rng.random() < 0.1:
            tags.append("XG:Z:y")
        if rng.random() < 0.2:
            tags.append("RG:Z:rg1")
        seq = "*" if qlen == 0 else "".join(rng.choice("ACGT") for _ in range(qlen))
        name = f"q{rng.randint(0, n // 2)}"                                     # mates, repeats and strangers share names
        if flag & 4 and rng.random() < 0.3:
            chrom_out, pos_out = "*", 0
        else:
            chrom_out, pos_out = chrom, pos
        rows.append((ci if chrom_out != "*" else 99, pos_out, "\t".join([name, str(flag), chrom_out, str(pos_out), str(mapq), cigar, rnext, str(pnext), str(tlen), seq, "*"] + tags)))
    rows.sort(key=lambda r: (r[0], r[1]))
    hdr = ["@HD\tVN:1.0\tSO:coordinate"] + [f"@SQ\tSN:{c}\tLN:{max(l, 3_000_000)}" for c, l in contigs] + ["@RG\tID:rg1\tSM:s"]
    open(path, "w").write("\n".join(hdr + [r[2] for r in rows]) + "\n")


FLAGSETS = [
    [],
    ["--coverage", "--stranded", "rf"],
    ["-u", "--exclude-chimeric", "--stranded", "fr"],
    ["--legacy"],
    ["--legacy", "--exclude-chimeric", "-u"],
    ["--tag", "XF", "--tag", "XG", "--chimeric-tag", "ch", "-q", "3", "--base-mismatch", "5", "--chimeric-distance", "1000"],
    ["--bed", "BED", "--fragment-samples", "50", "--coverage"],
]


# 5021 / 5038 / 5048: a mapped record whose CIGAR spans no reference base, lying one past the end of a gene, under --legacy (the
# whole-read interval is inverted there; found by a 60-seed sweep of this generator)
@pytest.mark.parametrize("seed", [4100, 4101, 4102, 4103, 4104, 4105, 5021, 5038, 5048])
def test_random_sam_records_match_the_reference(libs, tmp_path, seed):
    rng = random.Random(seed)
    d = str(tmp_path)
    gtf, sam, bed = os.path.join(d, "f.gtf"), os.path.join(d, "f.sam"), os.path.join(d, "f.bed")
    rows = _gtf(gtf, rng)
    _sam(sam, rng, 2500, allow_dup=False)
    ex = [r.split("\t") for r in rows if "\texon\t" in r]
    open(bed, "w").write("".join(f"{e[0]}\t{int(e[3]) - 1}\t{e[4]}\n" for e in ex[:: 3]))
    for k, flags in enumerate(FLAGSETS):
        flags = [bed if f == "BED" else f for f in flags]
        pr = run_cli(REF, gtf, sam, os.path.join(d, f"ref{k}"), *flags, check=False)
        po = run_cli(libs["oracle_cli"], gtf, sam, os.path.join(d, f"ours{k}"), *flags, check=False)
        assert pr.returncode == po.returncode, (flags, pr.returncode, po.returncode, pr.stderr[-400:], po.stderr[-400:])
        assert pr.returncode == 0, pr.stderr[-400:]
        identical = set(compare_outputs(os.path.join(d, f"ref{k}"), os.path.join(d, f"ours{k}")))
        assert {"f.sam.gene_reads.gct", "f.sam.gene_fragments.gct", "f.sam.exon_reads.gct"} <= identical, (flags, identical)


SAM2CRAM = os.path.join(ROOT, "oracle", "_ref", "sam2cram")


def _fasta(path, contigs):
    """A genome for the fuzz contigs: a repeated 4-mer pattern with runs of N and lower-case stretches (+ .fai)."""
    import numpy as np
    with open(path, "w") as f, open(path + ".fai", "w") as fai:
        for k, (name, ln) in enumerate(contigs):
            rng = np.random.default_rng(77 + k)
            seq = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, ln)].copy()
            for s in rng.integers(0, max(ln - 400, 1), 40):
                seq[s:s + int(rng.integers(1, 300))] = ord("N")
            for s in rng.integers(0, max(ln - 400, 1), 40):
                seq[s:s + 200] |= 0x20                                            # lower case
            hdr = f">{name}\n"
            f.write(hdr)
            off = f.tell()
            body = seq.tobytes().decode()
            f.write("\n".join(body[i:i + 60] for i in range(0, ln, 60)) + "\n")
            fai.write(f"{name}\t{ln}\t{off}\t60\t61\n")


def _strip_nm_z(sam_in, sam_out):
    # htslib's CRAM writer refuses NM tags of string type ("Unhandled type code for NM tag", cram_encode.c:2113-2125)
    with open(sam_in) as f, open(sam_out, "w") as g:
        for line in f:
            g.write(line.replace("\tNM:Z:3", ""))


@pytest.mark.skipif(not os.path.exists(SAM2CRAM), reason="reference htslib tool not built (make ref)")
@pytest.mark.parametrize("seed", [4200, 4201, 5021])
@pytest.mark.parametrize("fmt", ["bam", "cram", "cram_small_slices"])
def test_random_records_as_bam_and_cram_match_the_reference(libs, tmp_path, seed, fmt):
    """The same odd records written by the reference's htslib as BAM (sequential and multi-threaded readers) and CRAM (own reader;
    NM regenerated from the read features against a genome with N runs)."""
    import subprocess
    rng = random.Random(seed)
    d = str(tmp_path)
    gtf, sam0, sam, fa = os.path.join(d, "f.gtf"), os.path.join(d, "f0.sam"), os.path.join(d, "f.sam"), os.path.join(d, "f.fa")
    _gtf(gtf, rng)
    _sam(sam0, rng, 2500, allow_dup=False)
    _strip_nm_z(sam0, sam)
    ext = "bam" if fmt == "bam" else "cram"
    aln = os.path.join(d, "f." + ext)
    extra = []
    if fmt != "bam":
        _fasta(fa, [("chr1", 3_000_000), ("chr2", 3_000_000), ("chrM", 3_000_000), ("chrUn", 3_000_000)])
        extra = ["--fasta", fa]
    slicing = ["97", "2"] if fmt == "cram_small_slices" else []
    subprocess.run([SAM2CRAM, sam, fa, aln] + slicing, check=True, stderr=subprocess.DEVNULL)
    for k, flags in enumerate([[], ["--legacy", "-u"], ["--stranded", "rf", "--coverage", "--exclude-chimeric"]]):
        pr = run_cli(REF, gtf, aln, os.path.join(d, f"ref{k}"), *flags, *extra, check=False)
        outs = []
        for threads in ("0", "3"):
            po = run_cli(libs["oracle_cli"], gtf, aln, os.path.join(d, f"ours{k}_{threads}"), *flags, *extra, "--decode-threads", threads, check=False)
            assert pr.returncode == po.returncode == 0, (flags, pr.returncode, po.returncode, pr.stderr[-400:], po.stderr[-400:])
            identical = set(compare_outputs(os.path.join(d, f"ref{k}"), os.path.join(d, f"ours{k}_{threads}")))
            assert {f"f.{ext}.gene_reads.gct", f"f.{ext}.gene_fragments.gct", f"f.{ext}.exon_reads.gct"} <= identical, (flags, identical)


def _pairs_sam(path, rng, rows_gtf, n_pairs, shared_names=True):
    """High-quality pairs placed inside exons, then broken in every way the fragment-size and GC-content samplers care about: three
    and four records under one name, both mates forward, same start, mates in different exons or genes, a mate missing or unmapped,
    TLEN with the wrong sign or zero, spliced mates, insert sizes on the sampler's 100 / 1000 limits."""
    ex = [r.split("\t") for r in rows_gtf if "\texon\t" in r]
    ex = [(e[0], int(e[3]), int(e[4])) for e in ex if int(e[4]) - int(e[3]) > 120]
    recs = []
    for i in range(n_pairs):
        chrom, s, e = ex[rng.randrange(len(ex))]
        rl = rng.choice([36, 50, 76])
        ins = rng.choice([rl, rl + 1, 99, 100, 101, 150, 226, 400, 999, 1000, 1001])
        p1 = rng.randint(max(1, s - 20), max(s, e - rl))
        p2 = p1 + max(0, ins - rl)
        name = f"p{i}"
        f1, f2 = rng.choice([(99, 147), (163, 83), (97, 145), (65, 129), (99, 147), (99, 147)])
        t1, t2 = ins, -ins
        odd = rng.random()
        c1 = c2 = f"{rl}M"
        if odd < 0.05:
            f2 = f2 & ~16 | 32                      # both forward
        elif odd < 0.10:
            p2 = p1                                 # same start
        elif odd < 0.15:
            t1, t2 = -ins, ins                      # wrong signs
        elif odd < 0.20:
            t1 = t2 = 0
        elif odd < 0.25:
            c2 = f"{rl // 2}M{rng.choice([50, 300])}N{rl - rl // 2}M"   # spliced upper mate
        elif odd < 0.30 and shared_names:
            name = f"p{rng.randrange(max(i, 1))}"   # a name used before: three or four records share it
        elif odd < 0.33:
            f2 |= 4                                 # upper mate unmapped (keeps its coordinates)
        elif odd < 0.36:
            f1 |= 512                               # (no duplicate flags: they start the reference's 1e9-step complexity scan)
        elif odd < 0.40:
            chrom2, s2, e2 = ex[rng.randrange(len(ex))]
            if chrom2 == chrom:
                p2 = rng.randint(s2, max(s2, e2 - rl))                  # mate in another exon
        nm = rng.choice([0, 0, 0, 1, 6, 7])
        mq = rng.choice([255, 255, 255, 255, 60])
        seq = "".join(rng.choice("ACGT") for _ in range(rl))
        recs.append((chrom, p1, "\t".join([name, str(f1), chrom, str(p1), str(mq), c1, "=", str(p2), str(t1), seq, "*", f"NM:i:{nm}"])))
        if odd < 0.43 or odd >= 0.46:               # else: the upper mate is missing
            recs.append((chrom, p2, "\t".join([name, str(f2), chrom, str(p2), str(mq), c2, "=", str(p1), str(t2), seq, "*", f"NM:i:{nm}"])))
    order = {"chr1": 0, "chr2": 1, "chrM": 2}
    recs.sort(key=lambda r: (order[r[0]], r[1]))
    hdr = ["@HD\tVN:1.0\tSO:coordinate", "@SQ\tSN:chr1\tLN:100000", "@SQ\tSN:chr2\tLN:100000", "@SQ\tSN:chrM\tLN:100000"]
    open(path, "w").write("\n".join(hdr + [r[2] for r in recs]) + "\n")


@pytest.mark.parametrize("seed", range(4))
def test_random_pairs_fragment_sizes_and_gc_match_the_reference(libs, tmp_path, seed):
    rng = random.Random(4300 + seed)
    d = str(tmp_path)
    gtf, sam, bed, fa = (os.path.join(d, n) for n in ("f.gtf", "f.sam", "f.bed", "f.fa"))
    rows = _gtf(gtf, rng)
    _pairs_sam(sam, rng, rows, 1500)
    ex = sorted({(e[0], int(e[3]), int(e[4])) for e in (r.split("\t") for r in rows if "\texon\t" in r)})
    open(bed, "w").write("".join(f"{c}\t{s - 1}\t{e}\n" for c, s, e in ex))
    _fasta(fa, [("chr1", 100000), ("chr2", 100000), ("chrM", 100000)])
    for k, flags in enumerate([["--bed", bed], ["--bed", bed, "--fragment-samples", "25", "--fasta", fa], ["--fasta", fa, "--stranded", "fr", "--coverage"],
                               ["--bed", bed, "--legacy", "--fasta", fa], ["--bed", bed, "-u", "-q", "60", "--base-mismatch", "0"]]):
        pr = run_cli(REF, gtf, sam, os.path.join(d, f"ref{k}"), *flags, check=False)
        po = run_cli(libs["oracle_cli"], gtf, sam, os.path.join(d, f"ours{k}"), *flags, check=False)
        assert pr.returncode == po.returncode == 0, (flags, pr.returncode, po.returncode, pr.stderr[-400:], po.stderr[-400:])
        identical = set(compare_outputs(os.path.join(d, f"ref{k}"), os.path.join(d, f"ours{k}")))
        need = {"f.sam.gene_reads.gct", "f.sam.gene_fragments.gct", "f.sam.exon_reads.gct"}
        if "--bed" in flags:
            need.add("f.sam.fragmentSizes.txt")
            assert sum(int(l.split("\t")[1]) for l in open(os.path.join(d, f"ref{k}", "f.sam.fragmentSizes.txt")).readlines()[1:]) >= 20
        if "--fasta" in flags and "--legacy" not in flags:
            need.add("f.sam.gc_content.tsv")
        assert need <= identical, (flags, identical)


def _wild_gtf(path, rng):
    """Annotations no collapser would write: gene lines in any file order and on interleaved contigs, exons before their gene line,
    overlapping / nested / duplicate-coordinate exons inside one gene, genes without exons, exons
    whose gene has no gene line, transcript and CDS lines in between, one-base exons and genes, strands '.', rRNA by transcript_type."""
    rows = []
    g = 0
    for chrom, span in (("chr1", 40000), ("chr2", 20000), ("chrM", 6000)):
        for _ in range(rng.randint(6, 22)):
            g += 1
            strand = rng.choice("++--.")
            ttype = rng.choice(["protein_coding"] * 5 + ["rRNA", "lincRNA", "Mt_rRNA"])
            gs = rng.randint(1, span - 2000)
            ge = gs + rng.choice([0, 1, 50, 400, 800, 1500, 1500, 3000, 6000])
            common = f'gene_id "G{g}"; transcript_id "T{g}"; gene_type "{ttype}"; gene_name "N{g % 7}"; transcript_type "{ttype}";'
            lines = []
            if rng.random() < 0.9:
                lines.append(f"{chrom}\tw\tgene\t{gs}\t{ge}\t.\t{strand}\t.\t{common}")
            if rng.random() < 0.5:
                lines.append(f"{chrom}\tw\ttranscript\t{gs}\t{ge}\t.\t{strand}\t.\t{common}")
            for k in range(rng.choice([0, 1, 1, 2, 3, 5, 8])):
                # mostly inside the gene's span; one exon in six reaches past the gene line's end: the reference integrates a gene's
                # coverage and erases its fragment set when the gene line leaves the window, so reads on the far side come too late
                # for the coverage and count their fragments anew (rsq_eit.h, gene_eviction_bounds)
                es = rng.randint(gs, ge)
                ee = es + rng.choice([0, 1, 30, 120, 120, 300, 500, 1200, 2500])
                if rng.random() < 0.84:
                    ee = min(ge, ee)
                if k and rng.random() < 0.2:
                    es, ee = prev                                                   # same coordinates as the previous exon
                prev = (es, ee)
                estrand = strand if rng.random() < 0.9 else rng.choice("+-")
                lines.append(f'{chrom}\tw\texon\t{es}\t{ee}\t.\t{estrand}\t.\t{common} exon_id "G{g}_e{k}";')
                if rng.random() < 0.2:
                    lines.append(f'{chrom}\tw\tCDS\t{es}\t{ee}\t.\t{estrand}\t0\t{common}')
            if rng.random() < 0.3:
                rng.shuffle(lines)                                                    # exons before their gene line
            rows.append(lines)
    rng.shuffle(rows)                                                                 # file order unrelated to coordinates
    flat = [l for ls in rows for l in ls]
    open(path, "w").write("\n".join(flat) + "\n")
    return flat


@pytest.mark.parametrize("seed", range(8))
def test_wild_annotations_match_the_reference(libs, tmp_path, seed):
    rng = random.Random(4400 + seed)
    d = str(tmp_path)
    gtf, sam = os.path.join(d, "f.gtf"), os.path.join(d, "f.sam")
    rows = _wild_gtf(gtf, rng)
    if not any("\texon\t" in r for r in rows) or not any("\tgene\t" in r for r in rows):
        pytest.skip("no genes or no exons drawn")
    # well-formed high-quality pairs over the exons (and 20 bases around them), plus the random records
    ex = [r.split("\t") for r in rows if "\texon\t" in r]
    recs = []
    order = {"chr1": 0, "chr2": 1, "chrM": 2}
    for i in range(2500):
        e = ex[rng.randrange(len(ex))]
        s, t = int(e[3]), int(e[4])
        rl = rng.choice([25, 50, 76])
        p1 = max(1, rng.randint(s - 20, max(s - 20, t - rl + 20)))
        ins = rng.choice([rl, 120, 200, 350])
        p2 = p1 + max(0, ins - rl)
        cig = f"{rl}M" if rng.random() < 0.8 else f"{rl // 2}M{rng.choice([40, 150, 700])}N{rl - rl // 2}M"
        seq = "A" * rl
        for flag, pos, mpos, tl in ((99, p1, p2, ins), (147, p2, p1, -ins)):
            recs.append((order[e[0]], pos, "\t".join([f"w{i}", str(flag), e[0], str(pos), "255", cig, "=", str(mpos), str(tl), seq, "*", f"NM:i:{rng.choice([0, 0, 2])}"])))
    recs.sort(key=lambda r: (r[0], r[1]))
    hdr = ["@HD\tVN:1.0\tSO:coordinate", "@SQ\tSN:chr1\tLN:100000", "@SQ\tSN:chr2\tLN:100000", "@SQ\tSN:chrM\tLN:100000"]
    open(sam, "w").write("\n".join(hdr + [r[2] for r in recs]) + "\n")
    for k, flags in enumerate([["--coverage"], ["--coverage", "--stranded", "rf", "--coverage-mask", "20", "--gene-length", "100", "--window-size", "30", "--detection-threshold", "1"],
                               ["--legacy", "--coverage"], ["--coverage", "-u", "--offset", "10", "--window-size", "50", "--gene-length", "200"]]):
        pr = run_cli(REF, gtf, sam, os.path.join(d, f"ref{k}"), *flags, check=False)
        po = run_cli(libs["oracle_cli"], gtf, sam, os.path.join(d, f"ours{k}"), *flags, check=False)
        assert pr.returncode == po.returncode, (flags, pr.returncode, po.returncode, pr.stderr[-400:], po.stderr[-400:])
        if pr.returncode == 0:
            identical = set(compare_outputs(os.path.join(d, f"ref{k}"), os.path.join(d, f"ours{k}")))
            assert {"f.sam.gene_reads.gct", "f.sam.gene_fragments.gct", "f.sam.exon_reads.gct"} <= identical, (flags, identical)


@pytest.mark.parametrize("seed", range(4))
def test_contig_order_and_returning_contigs_match_the_reference(libs, tmp_path, seed):
    """@SQ order unrelated to the GTF's contig order, and a stream that comes BACK to contigs it has left: the reference has dropped
    their features by then (dropFeatures, src/Expression.cpp:96-103), so the returning records find nothing."""
    rng = random.Random(4500 + seed)
    d = str(tmp_path)
    gtf, sam0, sam = os.path.join(d, "f.gtf"), os.path.join(d, "f0.sam"), os.path.join(d, "f.sam")
    _gtf(gtf, rng)
    _sam(sam0, rng, 2500, allow_dup=False)
    lines = open(sam0).read().split("\n")
    hdr = [l for l in lines if l.startswith("@")]
    recs = [l for l in lines if l and not l.startswith("@")]
    sq = [l for l in hdr if l.startswith("@SQ")]
    rng.shuffle(sq)
    hdr = [hdr[0]] + sq + [l for l in hdr[1:] if not l.startswith("@SQ")]
    by = {}
    for r in recs:
        by.setdefault(r.split("\t")[2], []).append(r)
    out, tail = [], []
    for c in [l.split("\t")[1][3:] for l in sq] + ["*"]:
        rs = by.get(c, [])
        if rs and rng.random() < 0.6:
            k = rng.randrange(len(rs))
            tail.append(rs[k:])
            rs = rs[:k]
        out += rs
    rng.shuffle(tail)
    for t in tail:
        out += t
    open(sam, "w").write("\n".join(hdr + out) + "\n")
    for k, flags in enumerate([["--coverage"], ["--legacy"], ["--stranded", "rf", "-u"]]):
        pr = run_cli(REF, gtf, sam, os.path.join(d, f"ref{k}"), *flags, check=False)
        po = run_cli(libs["oracle_cli"], gtf, sam, os.path.join(d, f"ours{k}"), *flags, check=False)
        assert pr.returncode == po.returncode == 0, (flags, pr.returncode, po.returncode, pr.stderr[-400:], po.stderr[-400:])
        identical = set(compare_outputs(os.path.join(d, f"ref{k}"), os.path.join(d, f"ours{k}")))
        assert {"f.sam.gene_reads.gct", "f.sam.gene_fragments.gct", "f.sam.exon_reads.gct"} <= identical, (flags, identical)


@pytest.mark.skipif(not os.path.exists(SAM2CRAM), reason="reference htslib tool not built (make ref)")
def test_long_reads_spanning_many_bgzf_blocks(libs, tmp_path):
    """Records of up to 1.2 M bases and hundreds of CIGAR operations (one record = up to ~19 BGZF blocks) among short ones: the
    multi-threaded BAM ingest has to carry partial records across its jobs; all readers against the reference."""
    import subprocess
    rng = random.Random(3)
    d = str(tmp_path)
    gtf, sam, bam = os.path.join(d, "f.gtf"), os.path.join(d, "f.sam"), os.path.join(d, "f.bam")
    _gtf(gtf, rng)
    hdr = ["@HD\tVN:1.0\tSO:coordinate", "@SQ\tSN:chr1\tLN:3000000", "@SQ\tSN:chr2\tLN:3000000", "@SQ\tSN:chrM\tLN:3000000"]
    recs = []
    for i in range(400):
        pos = rng.randint(1, 50000)
        if rng.random() < 0.1:
            left, q, ops = rng.choice([5000, 70000, 300000, 1200000]), 0, []
            while left > 0:
                m = min(left, rng.randint(50, 4000))
                ops.append(f"{m}M")
                q += m
                left -= m
                if left > 0:
                    t, k = rng.choice("IDN"), rng.randint(1, 300)
                    ops.append(f"{k}{t}")
                    q += k if t == "I" else 0
            cigar, seq = "".join(ops), "A" * q
        else:
            cigar, seq = "76M", "C" * 76
        recs.append((pos, "\t".join([f"L{i}", str(rng.choice([0, 16])), "chr1", str(pos), "255", cigar, "*", "0", "0", seq, "*", "NM:i:0"])))
    recs.sort()
    open(sam, "w").write("\n".join(hdr + [r[1] for r in recs]) + "\n")
    subprocess.run([SAM2CRAM, sam, "-", bam], check=True, stderr=subprocess.DEVNULL)
    run_cli(REF, gtf, bam, os.path.join(d, "ref"), "-u")
    for threads in ("0", "2", "8"):
        run_cli(libs["oracle_cli"], gtf, bam, os.path.join(d, "ours" + threads), "-u", "--decode-threads", threads)
        identical = set(compare_outputs(os.path.join(d, "ref"), os.path.join(d, "ours" + threads)))
        assert {"f.bam.gene_reads.gct", "f.bam.exon_reads.gct", "f.bam.metrics.tsv"} <= identical, (threads, identical)
