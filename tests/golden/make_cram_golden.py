#!/usr/bin/env python
"""Golden vectors for CRAM input (TEST INFRASTRUCTURE; run in the build container, where /root/reference is present).

Generates a seeded annotation + genome + records whose SEQUENCES follow the genome (so that htslib's CRAM writer stores few
features and the NM the reference regenerates while reading is small), writes the SAM, converts it to CRAM with the reference's
own htslib (oracle/_ref/sam2cram, 900 records per slice, 2 slices per container: multi-slice containers, multi-reference
slices at contig boundaries), runs the UNMODIFIED reference binary on the CRAM and commits its outputs:
    tests/golden/cram_small/{c.cram, c.fa.gz, c.fa.fai, c.gtf, c.bed, expected/*}
usage: python tests/golden/make_cram_golden.py [out_dir]      (default tests/golden/cram_small)"""
import gzip
import os
import shutil
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
FLAGS = ["--coverage", "--bed", "c.bed"]


def read_fasta(path):
    seqs, name, buf = {}, None, []
    for line in open(path):
        if line.startswith(">"):
            if name is not None:
                seqs[name] = "".join(buf)
            name, buf = line[1:].split()[0], []
        else:
            buf.append(line.strip())
    if name is not None:
        seqs[name] = "".join(buf)
    return seqs


def realistic_sam(sam_in, sam_out, genome, seed=11):
    """Rewrites SEQ/QUAL of every record: aligned bases copied from the genome (upper-cased; genome N stays N) with a few random
    mismatches, random inserted / soft-clipped bases, an occasional N in the read; NM/MD tags dropped (CRAM regenerates them)."""
    rng = np.random.default_rng(seed)
    acgt = "ACGT"
    import re
    with open(sam_in) as fi, open(sam_out, "w") as fo:
        for line in fi:
            if line.startswith("@"):
                fo.write(line)
                continue
            f = line.rstrip("\n").split("\t")
            flag, contig, pos, cigar = int(f[1]), f[2], int(f[3]), f[5]
            if flag & 4 or cigar == "*" or contig not in genome:
                f[9], f[10] = "ACGTN" * 30, "F" * 150
            else:
                ref, p, out = genome[contig], pos - 1, []
                for n, op in re.findall(r"(\d+)([MIDNSHP=X])", cigar):
                    n = int(n)
                    if op in "M=X":
                        seg = list(ref[p:p + n].upper().ljust(n, "A"))
                        k = int(rng.choice([0, 0, 0, 1, 1, 2, 3, 9]))
                        for j in rng.choice(n, size=min(k, n), replace=False):
                            seg[j] = acgt[(acgt.find(seg[j]) + 1 + int(rng.integers(3))) % 4] if seg[j] in acgt else "C"
                        if rng.random() < 0.02:
                            seg[int(rng.integers(n))] = "N"
                        out.append("".join(seg)); p += n
                    elif op in "IS":
                        out.append("".join(acgt[int(x)] for x in rng.integers(4, size=n)))
                    elif op in "DN":
                        p += n
                seq = "".join(out)
                f[9], f[10] = seq, "F" * len(seq)
            f = f[:11] + [t for t in f[11:] if not t.startswith(("NM:", "MD:"))]
            fo.write("\t".join(f) + "\n")


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "cram_small")
    from rnaseqc_b200 import synth
    os.makedirs(out, exist_ok=True)
    tmp = os.path.join(out, "_tmp")
    os.makedirs(tmp, exist_ok=True)
    sa = synth.make_annotation(seed=177, n_genes=60, contig_names=["chr1", "chr2", "chrM"], contig_lens=[150000, 90000, 16569])
    synth.write_gtf(sa, os.path.join(out, "c.gtf"))
    synth.write_bed(sa, os.path.join(out, "c.bed"), min_len=300)
    synth.write_fasta(sa, os.path.join(tmp, "c.fa"), seed=5)
    sr = synth.simulate_reads(sa, 6000, seed=178)
    synth.write_sam(sa, sr, os.path.join(tmp, "plain.sam"))
    realistic_sam(os.path.join(tmp, "plain.sam"), os.path.join(tmp, "c.sam"), read_fasta(os.path.join(tmp, "c.fa")))
    subprocess.run([os.path.join(ROOT, "oracle", "_ref", "sam2cram"), os.path.join(tmp, "c.sam"), os.path.join(tmp, "c.fa"),
                    os.path.join(out, "c.cram"), "900", "2"], check=True)
    shutil.copy(os.path.join(tmp, "c.fa.fai"), os.path.join(out, "c.fa.fai"))
    with open(os.path.join(tmp, "c.fa"), "rb") as fi, gzip.open(os.path.join(out, "c.fa.gz"), "wb") as fo:
        shutil.copyfileobj(fi, fo)
    exp = os.path.join(out, "expected")
    shutil.rmtree(exp, ignore_errors=True)
    shutil.copy(os.path.join(tmp, "c.fa"), os.path.join(out, "c.fa"))
    subprocess.run([os.path.join(ROOT, "oracle", "_ref", "rnaseqc"), "c.gtf", "c.cram", "expected", "--fasta", "c.fa"] + FLAGS, cwd=out, check=True,
                   stdout=subprocess.DEVNULL)
    os.remove(os.path.join(out, "c.fa"))
    shutil.rmtree(tmp)
    print("wrote", out, sorted(os.listdir(exp)))


if __name__ == "__main__":
    main()
