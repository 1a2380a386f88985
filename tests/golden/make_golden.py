#!/usr/bin/env python
"""Generates the committed golden fixtures from the UNMODIFIED reference binary (oracle/_ref/rnaseqc, built by
`make ref` from /root/reference).  Run in the build container only; the GPU box uses the committed files.

    python tests/golden/make_golden.py

Cases
  synth_small : seeded synthetic annotation (90 genes / 3 contigs incl. a tiny chrM) + 30 k 150 bp PE records from
                rnaseqc_b200.synth; reference run with --coverage --bed.
  kat_odd     : the two shapes round 2 closed: exons that reach past their gene line (the gene leaves the reference's
                sliding window early: coverage dropped, fragments counted anew) and three, four and five records under one read
                name inside one BED interval / exon (the mate maps' erase-and-restart rule), --bed + --fasta.
  kat         : hand-written known-answer GTF + SAM, one read (pair) per quirk listed in SURVEY.md §8c; reference run
                under several flag sets (default, --stranded rf/fr, -u, --exclude-chimeric, -t XS, --rpkm -s).
"""
import gzip
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref", "rnaseqc")

KAT_FLAGSETS = {
    "default": ["--coverage"],
    "stranded_rf": ["--stranded", "rf"],
    "stranded_fr": ["--stranded", "fr"],
    "unpaired": ["-u"],
    "exclude_chimeric": ["--exclude-chimeric"],
    "tag_filter": ["-t", "XS", "-t", "XT"],
    "rpkm_sample": ["--rpkm", "-s", "mysample", "-q", "3", "--base-mismatch", "2", "--chimeric-distance", "1000"],
    "legacy": ["--legacy", "--coverage"],
}
KAT_LEGACY_FLAGSETS = {
    "legacy": ["--legacy", "--coverage"],
    "legacy_rf": ["--legacy", "--stranded", "rf"],
    "legacy_fr_excl_u": ["--legacy", "--stranded", "fr", "--exclude-chimeric", "-u"],
    "legacy_q255": ["--legacy", "-q", "255", "--base-mismatch", "2", "--coverage"],
}
SYNTH_FASTA_FLAGSETS = {"fasta": ["--fasta", "g.fa", "--coverage"], "fasta_rf_bed": ["--fasta", "g.fa", "--stranded", "rf", "--bed", "a.bed"],
                        "fasta_legacy": ["--fasta", "g.fa", "--legacy"]}
KAT_ODD_FLAGSETS = {"odd": ["--coverage", "--coverage-mask", "0", "--gene-length", "50", "--bed", "a.bed", "--fasta", "g.fa"],
                    "odd_few": ["--bed", "a.bed", "--fragment-samples", "7", "--stranded", "fr"]}
SYNTH_FLAGSETS = {"coverage_bed": ["--coverage", "--bed", "a.bed"], "legacy": ["--legacy", "--coverage", "--bed", "a.bed"]}


def run_ref(gtf, sam, out, flags, cwd):
    shutil.rmtree(out, ignore_errors=True)
    p = subprocess.run([REF, gtf, sam, out] + flags, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if p.returncode != 0:
        raise SystemExit(f"reference failed ({p.returncode}): {p.stderr}")


def make_synth_small():
    from rnaseqc_b200 import synth
    d = os.path.join(HERE, "synth_small")
    os.makedirs(d, exist_ok=True)
    ann = synth.make_annotation(seed=31, n_genes=90, contig_names=["chr1", "chr2", "chrM"], contig_lens=[12_000_000, 8_000_000, 16569])
    synth.write_gtf(ann, os.path.join(d, "a.gtf"))
    synth.write_bed(ann, os.path.join(d, "a.bed"), min_len=400)
    reads = synth.simulate_reads(ann, 30000, seed=32)
    sam = os.path.join(d, "a.sam")
    synth.write_sam(ann, reads, sam)
    for name, flags in SYNTH_FLAGSETS.items():
        run_ref("a.gtf", "a.sam", os.path.join("expected", name), flags, d)
    with open(sam, "rb") as fi, gzip.open(sam + ".gz", "wb", compresslevel=9) as fo:
        shutil.copyfileobj(fi, fo)
    os.remove(sam)


def make_bam_small():
    """A BAM of several ingest jobs (the multi-threaded reader cuts the decompressed stream into ~4 MB jobs, most of which start inside
    a record): 60 k seeded records, --coverage --bed; the reference reads the same file."""
    from rnaseqc_b200 import synth
    d = os.path.join(HERE, "bam_small")
    os.makedirs(d, exist_ok=True)
    ann = synth.make_annotation(seed=61, n_genes=120, contig_names=["chr1", "chr2", "chrM"], contig_lens=[14_000_000, 9_000_000, 16569])
    synth.write_gtf(ann, os.path.join(d, "a.gtf"))
    synth.write_bed(ann, os.path.join(d, "a.bed"), min_len=400)
    reads = synth.simulate_reads(ann, 60000, seed=62, dup_frac=0.0)
    synth.write_bam(ann, reads, os.path.join(d, "a.bam"), level=9)
    run_ref("a.gtf", "a.bam", os.path.join("expected", "coverage_bed"), ["--coverage", "--bed", "a.bed"], d)


def make_synth_fasta():
    """--fasta: fragment GC-content sampler, exon GC column, gc_content.tsv (src/Expression.cpp:459-476, src/RNASeQC.cpp:632-674).
    The FASTA lists its contigs in another order than the GTF (that order seeds the chromosome ids, src/Fasta.cpp:91-93), lacks
    chrM and has a contig no feature uses."""
    from rnaseqc_b200 import synth
    d = os.path.join(HERE, "synth_fasta")
    os.makedirs(d, exist_ok=True)
    ann = synth.make_annotation(seed=51, n_genes=40, contig_names=["chr1", "chr2", "chrM"], contig_lens=[400_000, 250_000, 16569])
    synth.write_gtf(ann, os.path.join(d, "a.gtf"))
    synth.write_bed(ann, os.path.join(d, "a.bed"), min_len=400)
    synth.write_fasta(ann, os.path.join(d, "g.fa"), seed=52, contigs=[1, 0], extra=[("chrUn", 3000)])
    reads = synth.simulate_reads(ann, 20000, seed=53, dup_frac=0.0)
    sam = os.path.join(d, "a.sam")
    synth.write_sam(ann, reads, sam)
    for name, flags in SYNTH_FASTA_FLAGSETS.items():
        run_ref("a.gtf", "a.sam", os.path.join("expected", name), flags, d)
    for f in (sam, os.path.join(d, "g.fa")):
        with open(f, "rb") as fi, gzip.open(f + ".gz", "wb", compresslevel=9) as fo:
            shutil.copyfileobj(fi, fo)
        os.remove(f)


GTF_ATTR = 'gene_id "{g}"; transcript_id "{g}"; gene_type "{t}"; gene_name "{n}"; transcript_type "{t}";'


def kat_gtf():
    L = ["##kat annotation"]

    def gene(chrom, gid, name, strand, exons, ttype="protein_coding", exon_ids=True, omit_ttype=False):
        s, e = min(x[0] for x in exons), max(x[1] for x in exons)
        attr = GTF_ATTR.format(g=gid, t=ttype, n=name)
        if omit_ttype:
            attr = f'gene_id "{gid}"; transcript_id "{gid}"; gene_name "{name}";'
        L.append(f"{chrom}\tkat\tgene\t{s}\t{e}\t.\t{strand}\t.\t{attr}")
        L.append(f"{chrom}\tkat\ttranscript\t{s}\t{e}\t.\t{strand}\t.\t{attr}")
        for k, (a, b) in enumerate(exons):
            ex = f' exon_id "{gid}_{k + 1}";' if exon_ids else ""
            L.append(f"{chrom}\tkat\texon\t{a}\t{b}\t.\t{strand}\t.\t{attr}{ex}")

    gene("chrA", "GA", "GeneA", "+", [(1000, 1199), (2000, 2199), (3000, 3499)])
    gene("chrA", "GB", "GeneB", "-", [(2100, 2600)])
    gene("chrA", "GC", "GeneC", "+", [(10000, 10999), (12000, 14000)])
    gene("chrA", "GD", "HBB", "-", [(20000, 20500)])
    gene("chrA", "GE", "Ribo", "+", [(30000, 30300)], ttype="rRNA")
    gene("chrA", "GF", "CarryOver", "+", [(31000, 31400)], omit_ttype=True)   # inherits transcript_type "rRNA" (GTF.cpp:110)
    gene("chrA", "GH", "Unnamed", "-", [(40000, 40200), (40500, 40800)], exon_ids=False)
    gene("chrA", "GI", "Dot", ".", [(50000, 50400)])
    gene("chrA", "GJ", "Short", "+", [(60000, 60150)])
    gene("chrA", "GK", "FarAway", "+", [(3000000, 3000500)])
    gene("chrB", "GL", "OnB", "+", [(500, 900), (1500, 2500)])
    gene("chrC", "GM", "NoReads", "-", [(100, 2000)])
    return "\n".join(L) + "\n"


def kat_sam():
    H = ["@HD\tVN:1.6\tSO:coordinate", "@SQ\tSN:chrA\tLN:5000000", "@SQ\tSN:chrB\tLN:100000", "@SQ\tSN:chrZ\tLN:100000"]
    R = []
    n = [0]

    def rd(chrom, pos, cigar, flag=99, mapq=255, nm=0, lq=150, rnext="=", pnext=None, tlen=300, tags=(), name=None):
        n[0] += 1
        name = name or f"k{n[0]}"
        pnext = pos + 150 if pnext is None else pnext
        seq = "*" if lq == 0 else "A" * lq
        qual = "*" if lq == 0 else "F" * lq
        t = list(tags)
        if nm is not None:
            t.insert(0, f"NM:i:{nm}")
        R.append((chrom, pos, "\t".join([name, str(flag), chrom, str(pos), str(mapq), cigar, rnext, str(pnext), str(tlen), seq, qual] + t)))

    # --- chrA, GeneA/GeneB --------------------------------------------------------------------------------------
    rd("chrA", 850, "150M")                       # ends at 999: overhang-by-one vs exon@1000 -> Ambiguous
    rd("chrA", 1010, "150M")                      # fully inside exon 1
    rd("chrA", 1100, "150M")                      # partial overlap -> Ambiguous
    rd("chrA", 1100, "100M800N50M")               # spliced exon1 -> exon2, fractions 100/150 + 50/150
    rd("chrA", 1100, "100M800N50M", flag=147, name="k4")   # its mate: same qname -> one fragment
    rd("chrA", 1150, "50M1800N100M")              # exon1 -> exon3, skipping exon2
    rd("chrA", 1300, "150M")                      # intron of GeneA
    rd("chrA", 2120, "50M", lq=50)                # inside GeneA exon2 AND GeneB exon: both genes counted
    rd("chrA", 2120, "50M", lq=50, flag=147)      # same, reverse strand second mate
    rd("chrA", 2300, "100M", lq=100)              # GeneB exon only (GeneA intron)
    rd("chrA", 2150, "100M", lq=100)              # spans GeneA exon2 end: partial for A, full for B
    rd("chrA", 3010, "70M2I78M")                  # insertion: two abutting blocks in one exon
    rd("chrA", 3010, "70M2D80M", lq=150)          # deletion
    rd("chrA", 3020, "10S140M")                   # soft clip
    rd("chrA", 3030, "50=1X99=")                  # = and X ops
    rd("chrA", 3040, "5H145M", lq=145)            # hard clip
    rd("chrA", 3050, "20S30I", lq=50)             # no aligned block at all
    rd("chrA", 3060, "150M", nm=None)             # NM absent
    rd("chrA", 3070, "150M", nm=7)                # too many mismatches -> low quality
    rd("chrA", 3080, "150M", mapq=3)              # low MAPQ
    rd("chrA", 3090, "150M", flag=99 | 0x100, mapq=0)     # secondary + low MAPQ (both counted)
    rd("chrA", 3100, "150M", flag=99 | 0x800)             # supplementary without ch -> Chimeric auto
    rd("chrA", 3110, "150M", flag=99 | 0x800, tags=["ch:A:1"])   # supplementary with ch
    rd("chrA", 3120, "150M", flag=99 | 0x200)             # QC fail
    rd("chrA", 3130, "150M", flag=99 | 0x400)             # duplicate
    rd("chrA", 3130, "150M", flag=147 | 0x400, name="k25")
    rd("chrA", 3140, "150M", flag=0)                      # unpaired
    rd("chrA", 3150, "150M", flag=73, pnext=3150, tlen=0)  # mate unmapped (0x8)
    rd("chrA", 3160, "150M", flag=97, rnext="chrB", pnext=700, tlen=0)   # mate on another contig
    rd("chrA", 3170, "150M", flag=97, pnext=4000000, tlen=0)             # mate too far
    rd("chrA", 3180, "150M", tags=["ch:Z:chimera"])        # STAR chimeric tag, first mate
    rd("chrA", 3190, "150M", flag=147, tags=["ch:A:1"])    # chimeric tag on second mate
    rd("chrA", 3200, "150M", tags=["XS:i:1"])              # user filter tag (int)
    rd("chrA", 3210, "150M", tags=["XT:Z:foo"])            # user filter tag (string)
    rd("chrA", 3220, "150M", tags=["XS:A:+"])              # type A is not seen by GetTag
    rd("chrA", 3230, "150M", tags=["XS:f:1.5"])            # float is
    rd("chrA", 3240, "150M", flag=83)                      # reverse first mate
    rd("chrA", 3250, "150M", flag=163)                     # forward second mate
    rd("chrA", 3300, "100M", lq=100)                       # shorter read
    rd("chrA", 3300, "60M6600N40M", lq=100)                # long span with short length: Read Length quirk; 2nd block in GeneC
    # --- GeneC: deep coverage for the bias / mask paths --------------------------------------------------------------
    for k in range(1400):
        p = 10000 + (k * 7) % 840
        rd("chrA", p, "150M", flag=99 if k % 2 == 0 else 147, name=f"c{k // 2}")
    for k in range(1100):
        p = 12000 + (k * 13) % 1840
        rd("chrA", p, "150M", flag=83 if k % 2 == 0 else 163, name=f"d{k // 2}")
    rd("chrA", 10900, "100M1000N50M")             # junction of GeneC
    # --- other genes ----------------------------------------------------------------------------------------------------------
    rd("chrA", 20100, "150M")                     # globin
    rd("chrA", 20100, "150M", flag=147 | 0x400)   # globin duplicate
    rd("chrA", 30050, "150M")                     # rRNA
    rd("chrA", 31100, "150M")                     # rRNA by carry-over
    rd("chrA", 40010, "150M", flag=83)            # unnamed exons
    rd("chrA", 40100, "100M300N50M", flag=163)
    rd("chrA", 50100, "150M")                     # gene with strand '.'
    rd("chrA", 60000, "150M")                     # gene shorter than 200 bp (bias skip), exactly exon-sized
    rd("chrA", 70000, "150M")                     # intergenic
    rd("chrA", 3000100, "150M")                   # far away gene (chimeric distance tests)
    # --- chrB, then chrZ (not in the GTF), then back to chrA (its features were dropped) ------------------------------------
    rd("chrB", 600, "150M")
    rd("chrB", 800, "100M600N50M")
    rd("chrB", 1600, "150M", flag=147)
    rd("chrZ", 100, "150M")
    # unmapped, no contig
    R.sort(key=lambda x: ({"chrA": 0, "chrB": 1, "chrZ": 2}[x[0]], x[1]))
    lines = H + [x[2] for x in R]
    lines.append("\t".join(["back1", "99", "chrA", "1010", "255", "150M", "=", "1160", "300", "A" * 150, "F" * 150, "NM:i:0"]))
    lines.append("\t".join(["u1", "77", "*", "0", "0", "*", "*", "0", "0", "A" * 150, "F" * 150]))
    lines.append("\t".join(["u1", "141", "*", "0", "0", "*", "*", "0", "0", "A" * 150, "F" * 150]))
    return "\n".join(lines) + "\n"


def kat_legacy_gtf():
    """Annotation for the --legacy known-answer cases (src/Expression.cpp:129-304, src/RNASeQC.cpp:129-135)."""
    L = ["##kat legacy annotation"]

    def line(chrom, kind, s, e, strand, attr):
        L.append(f"{chrom}\tkat\t{kind}\t{s}\t{e}\t.\t{strand}\t.\t{attr}")

    def gene(chrom, gid, name, strand, exons, ttype="protein_coding", span=None):
        s, e = span if span else (min(x[0] for x in exons), max(x[1] for x in exons))
        attr = GTF_ATTR.format(g=gid, t=ttype, n=name)
        line(chrom, "gene", s, e, strand, attr)
        line(chrom, "transcript", s, e, strand, attr)
        for k, (a, b) in enumerate(exons):
            line(chrom, "exon", a, b, strand, attr + f' exon_id "{gid}_{k + 1}";')

    gene("chrA", "GA", "GeneA", "+", [(1000, 1199), (2000, 2199), (3000, 3499)])
    gene("chrA", "GB", "GeneB", "-", [(2100, 2600)])
    gene("chrA", "GS", "SingleBaseExon", "+", [(5000, 5000), (5100, 5300)])       # 1-base exon: dropped from the list, stays in exonList
    gene("chrA", "GT", "SingleBaseGene", "-", [(6000, 6000)])                     # 1-base gene line and exon
    gene("chrA", "GX", "Beyond", "+", [(8000, 8300)])
    gene("chrA", "GO", "Outer", "+", [(8900, 9300)])
    gene("chrA", "GN", "LateLine", "+", [(9600, 9900)], span=(9100, 9900))        # gene line starts after GO's exon: last list hit of a read at 8900
    gene("chrA", "GR", "Ribo", "-", [(11000, 11400)], ttype="rRNA")
    gene("chrA", "GD", "Dot", ".", [(12000, 12400), (12700, 12900)])
    gene("chrA", "GW", "Wide", "+", [(20000, 20200), (130000, 130200)])           # intron > LEGACY_MAX_READ_LENGTH
    # two exons of one gene overlapping each other (uncollapsed model): list order decides which one counts
    gene("chrA", "GU", "Overlap", "+", [(14000, 14400), (14000, 14200), (14100, 14600)])
    gene("c129", "GZ", "HighTid", "+", [(1000, 1500)])
    return "\n".join(L) + "\n"


def kat_legacy_sam():
    H = ["@HD\tVN:1.6\tSO:coordinate", "@SQ\tSN:chrA\tLN:5000000"] + [f"@SQ\tSN:c{k}\tLN:100000" for k in range(1, 130)]
    R = []
    n = [0]

    def rd(chrom, pos, cigar, flag=99, mapq=255, nm=0, lq=150, rnext="=", pnext=None, tlen=300, tags=(), name=None):
        n[0] += 1
        name = name or f"L{n[0]}"
        pnext = pos + 150 if pnext is None else pnext
        t = list(tags)
        if nm is not None:
            t.insert(0, f"NM:i:{nm}")
        R.append((chrom, pos, "\t".join([name, str(flag), chrom, str(pos), str(mapq), cigar, rnext, str(pnext), str(tlen), "A" * lq, "F" * lq] + t)))

    rd("chrA", 1010, "150M")                               # one block in exon 1: counted whole
    rd("chrA", 1010, "150M", flag=147, name="L1")          # mate: one fragment
    rd("chrA", 1100, "100M800N50M")                        # split (gap 800 > 99): dosage 100/150 and 50/150 as floats
    rd("chrA", 1150, "50M50N50M", lq=100)                  # gap 50: not split; second block intronic -> last block finds nothing: not counted
    rd("chrA", 1100, "150M")                               # partial overlap with exon 1 only -> intron+no exon -> Intronic
    rd("chrA", 1150, "50M800N100M")                        # first block inside, second inside exon 2: split
    rd("chrA", 1150, "60M790N100M", lq=160)                # first block overruns exon 1 (partial), second full: junction -> Intronic
    rd("chrA", 1010, "50M100N50M1790N50M")                 # blocks: exon1, intron, exon3 -> legacyNotSplit: exon 3 counted 1.0
    rd("chrA", 3010, "40M2D40M1000N70M", flag=163)         # deletion: two blocks in exon 3 (floats summed) ... third block outside the gene's exons
    rd("chrA", 3010, "40M2D40M100N70M")                    # split by the 100-base gap; three blocks all inside exon 3
    rd("chrA", 2120, "50M", lq=50)                         # inside GA exon 2 and GB exon: both gene lines counted
    rd("chrA", 2120, "50M", lq=50, flag=147 | 0x400)       # duplicate
    rd("chrA", 2300, "100M", lq=100)                       # GB only; GA intragenic without exon
    rd("chrA", 3400, "50M4550N100M")                       # second block starts beyond GA's end -> legacyNotExonic -> Intronic (block 2 in GX exon)
    rd("chrA", 5000, "1M99N140M", lq=141)                  # first block on the dropped 1-base exon
    rd("chrA", 5150, "100M", lq=100)                       # GS second exon
    rd("chrA", 6000, "1M", lq=1)                           # the 1-base gene does not exist in legacy mode -> Intergenic
    rd("chrA", 8900, "50M200N50M", lq=100)                 # GO split read; the last list hit is GN's gene line -> legacyNotSplit set: no 'Split Reads'
    rd("chrA", 8900, "100M", lq=100)                       # GO only (GN line starts at 9100)
    rd("chrA", 9650, "150M")                               # GN exon
    rd("chrA", 11050, "150M")                              # rRNA
    rd("chrA", 11050, "150M", flag=83)
    rd("chrA", 12050, "150M")                              # '.' strand
    rd("chrA", 12300, "100M300N50M", flag=163)
    rd("chrA", 14000, "150M")                              # overlapping exons of GU: first in list order that contains the block
    rd("chrA", 14120, "60M", lq=60)                        # contained by all three
    rd("chrA", 14250, "100M", lq=100)                      # contained by exons 1 and 3 only
    rd("chrA", 20050, "100M109850N50M")                    # PositionEnd - Position > 100000: skipped after 'Mapped Reads'
    rd("chrA", 20050, "100M", lq=100)
    rd("chrA", 3060, "150M", nm=None)                      # NM absent
    rd("chrA", 3070, "150M", nm=7)                         # low quality: classified, not counted
    rd("chrA", 3080, "150M", mapq=3)                       # below the legacy default of 4
    rd("chrA", 3085, "150M", mapq=4)                       # at the legacy default
    rd("chrA", 3090, "150M", flag=99 | 0x100)              # secondary
    rd("chrA", 3100, "150M", flag=99 | 0x800)              # supplementary: never 'Chimeric Fragments_auto' in legacy
    rd("chrA", 3110, "150M", flag=99 | 0x800, tags=["ch:A:1"])
    rd("chrA", 3180, "150M", tags=["ch:Z:chimera"])        # chimeric tag ignored in legacy
    rd("chrA", 3190, "150M", flag=97, pnext=4000000, tlen=0)   # mate too far
    rd("chrA", 3200, "150M", flag=0)                       # unpaired
    rd("chrA", 3210, "10S140M")                            # soft clip: Length() counts it
    rd("chrA", 3215, "10S40M200N100M")                     # soft clip + split: dosage over Length()=150
    rd("chrA", 70000, "150M")                              # intergenic
    rd("c129", 1100, "150M")                               # ChrID 129 > 127: paired with mapped mate -> chimeric in legacy
    rd("c129", 1100, "150M", flag=147, name="hz")
    rd("c129", 1200, "150M", flag=0)                       # unpaired: unaffected
    order = {"chrA": 0}
    order.update({f"c{k}": k for k in range(1, 130)})
    R.sort(key=lambda x: (order[x[0]], x[1]))
    return "\n".join(H + [x[2] for x in R]) + "\n"


def make_kat_legacy():
    d = os.path.join(HERE, "kat_legacy")
    os.makedirs(d, exist_ok=True)
    open(os.path.join(d, "kat.gtf"), "w").write(kat_legacy_gtf())
    sam = os.path.join(d, "kat.sam")
    open(sam, "w").write(kat_legacy_sam())
    for name, flags in KAT_LEGACY_FLAGSETS.items():
        run_ref("kat.gtf", "kat.sam", os.path.join("expected", name), flags, d)
    with open(sam, "rb") as fi, gzip.open(sam + ".gz", "wb", compresslevel=9) as fo:
        shutil.copyfileobj(fi, fo)
    os.remove(sam)


def kat_odd_files():
    import random
    rng = random.Random(777)
    A = 'gene_id "{g}"; transcript_id "{g}"; gene_type "protein_coding"; gene_name "{g}n"; transcript_type "protein_coding";'
    L = []

    def feat(kind, g, s, e, strand="+", eid=None, chrom="chr1"):
        L.append(f"{chrom}\tx\t{kind}\t{s}\t{e}\t.\t{strand}\t.\t" + A.format(g=g) + (f' exon_id "{eid}";' if eid else ""))
    # a gene line that ends before its second exon does
    feat("gene", "G1", 1000, 2000); feat("exon", "G1", 1000, 1200, eid="G1_1"); feat("exon", "G1", 1800, 3000, eid="G1_2")
    # a long gene that starts earlier keeps G2's line in the window until ITS end (prefix maximum of the list)
    feat("gene", "GL", 5000, 9000, "-"); feat("exon", "GL", 5000, 5200, "-", eid="GL_1"); feat("exon", "GL", 8800, 9000, "-", eid="GL_2")
    feat("gene", "G2", 5500, 6000); feat("exon", "G2", 5500, 5700, eid="G2_1"); feat("exon", "G2", 5900, 9500, eid="G2_2")
    feat("gene", "G3", 20000, 24000); feat("exon", "G3", 20000, 24000, eid="G3_1")
    feat("gene", "G4", 30000, 33000, "-"); feat("exon", "G4", 30000, 31500, "-", eid="G4_1"); feat("exon", "G4", 32000, 33000, "-", eid="G4_2")
    feat("gene", "G5", 1000, 5000, chrom="chr2"); feat("exon", "G5", 1000, 5000, eid="G5_1", chrom="chr2")
    H = ["@HD\tVN:1.6\tSO:coordinate", "@SQ\tSN:chr1\tLN:100000", "@SQ\tSN:chr2\tLN:100000"]
    R = []

    def rd(name, pos, flag, cigar="100M", lq=100, pnext=0, tlen=0, chrom="chr1"):
        R.append((chrom, pos, "\t".join([name, str(flag), chrom, str(pos), "255", cigar, "=", str(pnext), str(tlen), "A" * lq, "F" * lq, "NM:i:0"])))

    def pair(name, p1, p2, chrom="chr1", f=(99, 147)):
        rd(name, p1, f[0], pnext=p2, tlen=p2 + 100 - p1, chrom=chrom); rd(name, p2, f[1], pnext=p1, tlen=-(p2 + 100 - p1), chrom=chrom)
    pair("p1", 1010, 1850)            # both mates while the gene line is in the window
    pair("p2", 1900, 2500)            # the second mate finds the gene evicted: a second fragment, no coverage
    pair("p3", 2100, 2600)            # both beyond
    for k in range(30):
        rd(f"q{k}", 2200 + 10 * k, 99 if k % 2 == 0 else 147, pnext=2200)
    pair("r1", 5950, 6500); pair("r2", 8000, 9100); pair("r3", 9150, 9300)
    # G3 / G4 / G5 are BED intervals too: names shared by three, four and five records
    for i in range(60):
        base = 20000 + 40 * i
        n = rng.choice([2, 2, 3, 3, 4, 5])
        for k in range(n):
            fwd = k % 2 == 0 if rng.random() < 0.8 else rng.random() < 0.5
            p = base + 110 * k + rng.choice([0, 0, 3])
            rd(f"m{i}", p, 99 if fwd else 147, pnext=base if not fwd else p + 150, tlen=(300 if fwd else -300) + k)
    for i in range(40):
        base = 1000 + 60 * i
        n = rng.choice([3, 4])
        for k in range(n):
            fwd = k % 2 == 0
            rd(f"n{i}", base + 150 * k, 99 if fwd else 147, pnext=base, tlen=(250 if fwd else -250) - k, chrom="chr2")
    order = {"chr1": 0, "chr2": 1}
    R.sort(key=lambda x: (order[x[0]], x[1]))
    bed = "chr1\t19999\t24000\nchr1\t29999\t31500\nchr2\t999\t5000\n"
    seq = {c: "".join(rng.choice("ACGT") for _ in range(100000)) for c in ("chr1", "chr2")}
    fa = "".join(f">{c}\n" + "\n".join(seq[c][i:i + 60] for i in range(0, 100000, 60)) + "\n" for c in ("chr1", "chr2"))
    fai, off = "", 0
    for c in ("chr1", "chr2"):
        off += len(c) + 2
        fai += f"{c}\t100000\t{off}\t60\t61\n"
        off += 100000 + (100000 + 59) // 60
    return "\n".join(L) + "\n", "\n".join(H + [x[2] for x in R]) + "\n", bed, fa, fai


def make_kat_odd():
    d = os.path.join(HERE, "kat_odd")
    os.makedirs(d, exist_ok=True)
    gtf, sam, bed, fa, fai = kat_odd_files()
    for name, text in (("a.gtf", gtf), ("a.sam", sam), ("a.bed", bed), ("g.fa", fa), ("g.fa.fai", fai)):
        open(os.path.join(d, name), "w").write(text)
    for name, flags in KAT_ODD_FLAGSETS.items():
        run_ref("a.gtf", "a.sam", os.path.join("expected", name), flags, d)
    for f in ("a.sam", "g.fa"):
        with open(os.path.join(d, f), "rb") as fi, gzip.open(os.path.join(d, f + ".gz"), "wb", compresslevel=9) as fo:
            shutil.copyfileobj(fi, fo)
        os.remove(os.path.join(d, f))


def make_kat():
    d = os.path.join(HERE, "kat")
    os.makedirs(d, exist_ok=True)
    open(os.path.join(d, "kat.gtf"), "w").write(kat_gtf())
    sam = os.path.join(d, "kat.sam")
    open(sam, "w").write(kat_sam())
    for name, flags in KAT_FLAGSETS.items():
        run_ref("kat.gtf", "kat.sam", os.path.join("expected", name), flags, d)
    with open(sam, "rb") as fi, gzip.open(sam + ".gz", "wb", compresslevel=9) as fo:
        shutil.copyfileobj(fi, fo)
    os.remove(sam)


if __name__ == "__main__":
    if not os.path.exists(REF):
        raise SystemExit("build the reference first: make ref")
    make_synth_small()
    make_kat()
    make_kat_legacy()
    make_synth_fasta()
    make_kat_odd()
    make_bam_small()
    print("golden fixtures written under", HERE)
