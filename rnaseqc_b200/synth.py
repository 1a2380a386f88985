"""Deterministic synthetic workloads of the shape BASELINE.json names.

Two products from one seed:

* an annotation (collapsed-model GTF: one transcript per gene, disjoint exons inside a gene, ~10 % of genes
  overlapping a neighbour on the other strand, ~1 % on the same strand, 12 globin names, ~0.1 % rRNA) that can be
  written as GTF/BED text for the reference binary and for this repo's own parser;
* coordinate-sorted 150 bp paired-end alignment records drawn from it (85 % from transcripts with log-normal
  expression, 15 % uniform genomic; secondary/supplementary/QC-fail/duplicate/mate-unmapped/mate-elsewhere flags,
  MAPQ and NM distributions, soft clips and small indels as in SURVEY.md §8d), produced with torch ops so the same
  code runs on the CPU (tests, fixtures) and on a B200 (bench.py, 200 M records resident in HBM).

The record layout is include/rsq_b200.h's rsq_read (48 B) / rsq_block (8 B).
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

GLOBINS = ["HBA1", "HBA2", "HBB", "HBD", "HBG1", "HBG2", "HBE1", "HBM", "HBQ1", "HBZ", "HBBP1", "HBZP1"]

# GRCh38 primary contig lengths (chr1..22, X, Y, M)
GRCH38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717,
          133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285, 58617616,
          64444167, 46709983, 50818468, 156040895, 57227415, 16569]
GRCH38_NAMES = [f"chr{i}" for i in range(1, 23)] + ["chrX", "chrY", "chrM"]

READ_DTYPE = np.dtype([
    ("pos", "<i4"), ("mpos", "<i4"), ("endpos", "<i4"), ("isize", "<i4"), ("l_qseq", "<i4"), ("nm", "<i4"),
    ("flag", "<u2"), ("mapq", "u1"), ("aux", "u1"), ("contig", "<i2"), ("filter_mask", "u1"), ("reserved0", "u1"),
    ("name_hash", "<u8"), ("block_off", "<u4"), ("n_blocks", "<u2"), ("reserved1", "<u2")])
BLOCK_DTYPE = np.dtype([("start", "<i4"), ("len", "<i4")])
assert READ_DTYPE.itemsize == 48 and BLOCK_DTYPE.itemsize == 8

AUX_HAS_NM, AUX_CHIMERIC_TAG, AUX_MATE_OTHER_CONTIG = 0x01, 0x02, 0x04


@dataclasses.dataclass
class SynthAnnotation:
    contig_names: List[str]
    contig_lens: List[int]
    # per gene
    gene_contig: np.ndarray      # int32
    gene_start: np.ndarray       # int64, 1-based closed
    gene_end: np.ndarray
    gene_strand: np.ndarray      # 0 '+', 1 '-'
    gene_name: List[str]
    gene_rrna: np.ndarray        # bool
    gene_exon_off: np.ndarray    # int64 [n_genes+1] into exon arrays (genomic order inside a gene)
    exon_start: np.ndarray       # int64, 1-based closed
    exon_end: np.ndarray

    @property
    def n_genes(self) -> int:
        return len(self.gene_start)

    @property
    def n_exons(self) -> int:
        return len(self.exon_start)

    def gene_id(self, g: int) -> str:
        return f"ENSG{g + 1:011d}.1"


def make_annotation(seed: int = 20260925, n_genes: int = 5234, contig_names: Optional[List[str]] = None,
                    contig_lens: Optional[List[int]] = None, mean_exons: float = 5.9) -> SynthAnnotation:
    """Collapsed-model annotation.  Defaults give the chr1-scale config (5,234 genes / ~31 k exons);
    n_genes=56202 with the 25 GRCh38 contigs gives the GENCODE-scale one."""
    rng = np.random.default_rng(seed)
    if contig_names is None:
        contig_names, contig_lens = ["chr1"], [GRCH38[0]]
    total_len = float(sum(contig_lens))
    # genes per contig proportional to length (chrM gets a handful)
    per = np.maximum(1, np.floor(np.array(contig_lens) / total_len * n_genes)).astype(int)
    per[0] += n_genes - per.sum()
    g_contig, g_start, g_end, g_strand = [], [], [], []
    ex_start, ex_end, ex_off = [], [], [0]
    for c, (clen, ng) in enumerate(zip(contig_lens, per)):
        # draw exon structure for every gene of the contig
        n_ex = np.clip(rng.geometric(1.0 / mean_exons, size=ng), 1, 30)
        spans = []
        structs = []
        for k in range(ng):
            el = np.clip(np.exp(rng.normal(np.log(130.0), 0.9, size=n_ex[k])), 20, 6000).astype(np.int64)
            il = np.clip(np.exp(rng.normal(np.log(1500.0), 1.1, size=max(n_ex[k] - 1, 0))), 60, 200000).astype(np.int64)
            span = int(el.sum() + il.sum())
            structs.append((el, il))
            spans.append(span)
        spans = np.array(spans, dtype=np.int64)
        # ~11 % of genes start inside the previous gene (overlap), the rest after a gap
        overlap = rng.random(ng) < 0.11
        overlap[0] = False
        free = clen - 2000 - int(spans[~overlap].sum())
        n_gaps = int((~overlap).sum())
        if free < n_gaps * 200:
            # contig too small for the drawn structures (chrM): shrink introns by placing genes back to back
            scale = max(clen - 2000, 1) / float(spans.sum() + n_gaps * 200)
            structs = [(np.maximum((el * scale).astype(np.int64), 20), np.maximum((il * scale).astype(np.int64), 30)) for el, il in structs]
            spans = np.array([int(el.sum() + il.sum()) for el, il in structs], dtype=np.int64)
            free = max(clen - 2000 - int(spans[~overlap].sum()), n_gaps * 10)
        gaps = rng.dirichlet(np.ones(n_gaps)) * free if n_gaps else np.zeros(0)
        cursor = 1000
        gi = 0
        prev_start, prev_end, prev_strand = 0, 0, 0
        for k in range(ng):
            el, il = structs[k]
            if overlap[k] and prev_end - prev_start > 400:
                start = int(prev_start + rng.integers(50, max(51, (prev_end - prev_start) // 2)))
                same = rng.random() < 0.09   # ~1 % of all genes overlap on the same strand
                strand = prev_strand if same else 1 - prev_strand
            else:
                start = int(cursor + gaps[min(gi, n_gaps - 1)]) if n_gaps else cursor
                gi += 1
                strand = int(rng.integers(0, 2))
            s = start
            for j in range(len(el)):
                ex_start.append(s)
                ex_end.append(s + int(el[j]) - 1)
                s += int(el[j])
                if j < len(il):
                    s += int(il[j])
            end = ex_end[-1]
            if end >= clen - 100:   # ran off the contig: drop the exons just added and stop placing genes here
                del ex_start[ex_off[-1]:], ex_end[ex_off[-1]:]
                break
            ex_off.append(len(ex_start))
            g_contig.append(c); g_start.append(start); g_end.append(end); g_strand.append(strand)
            cursor = max(cursor, end + 1)
            prev_start, prev_end, prev_strand = start, end, strand
    n = len(g_start)
    names = [f"GENE{g + 1}" for g in range(n)]
    glob_idx = rng.choice(n, size=min(12, n), replace=False)
    for i, g in enumerate(glob_idx):
        names[int(g)] = GLOBINS[i]
    rrna = rng.random(n) < 0.001
    if n > 50:
        rrna[int(rng.integers(0, n))] = True
    return SynthAnnotation(list(contig_names), list(contig_lens), np.array(g_contig, np.int32), np.array(g_start, np.int64),
                           np.array(g_end, np.int64), np.array(g_strand, np.int8), names, rrna,
                           np.array(ex_off, np.int64), np.array(ex_start, np.int64), np.array(ex_end, np.int64))


def write_gtf(ann: SynthAnnotation, path: str) -> None:
    """GTEx-collapsed-style GTF: gene, transcript and exon lines with gene_id, transcript_id, gene_type, gene_name,
    transcript_type on every line and exon_id on exons."""
    with open(path, "w") as f:
        f.write("##description: synthetic collapsed annotation (rnaseqc_b200.synth)\n")
        for g in range(ann.n_genes):
            chrom = ann.contig_names[ann.gene_contig[g]]
            strand = "+-"[ann.gene_strand[g]]
            gid = ann.gene_id(g)
            ttype = "rRNA" if ann.gene_rrna[g] else "protein_coding"
            common = f'gene_id "{gid}"; transcript_id "{gid}"; gene_type "{ttype}"; gene_name "{ann.gene_name[g]}"; transcript_type "{ttype}"; transcript_name "{ann.gene_name[g]}";'
            f.write(f"{chrom}\tsynth\tgene\t{ann.gene_start[g]}\t{ann.gene_end[g]}\t.\t{strand}\t.\t{common}\n")
            f.write(f"{chrom}\tsynth\ttranscript\t{ann.gene_start[g]}\t{ann.gene_end[g]}\t.\t{strand}\t.\t{common}\n")
            lo, hi = ann.gene_exon_off[g], ann.gene_exon_off[g + 1]
            for j in range(lo, hi):
                f.write(f"{chrom}\tsynth\texon\t{ann.exon_start[j]}\t{ann.exon_end[j]}\t.\t{strand}\t.\t{common} exon_id \"{gid}_{j - lo + 1}\"; exon_number {j - lo + 1};\n")


def write_bed(ann: SynthAnnotation, path: str, min_len: int = 1000) -> int:
    """BED of the long exons that overlap no other exon (the fragment-size intervals), sorted; returns the count."""
    order = np.lexsort((ann.exon_start, np.repeat(ann.gene_contig, np.diff(ann.gene_exon_off))))
    contig = np.repeat(ann.gene_contig, np.diff(ann.gene_exon_off))[order]
    s, e = ann.exon_start[order], ann.exon_end[order]
    n = 0
    with open(path, "w") as f:
        for i in range(len(s)):
            if e[i] - s[i] + 1 < min_len:
                continue
            if i > 0 and contig[i - 1] == contig[i] and e[i - 1] >= s[i] - 2:
                continue
            if i + 1 < len(s) and contig[i + 1] == contig[i] and s[i + 1] <= e[i] + 2:
                continue
            f.write(f"{ann.contig_names[contig[i]]}\t{s[i] - 1}\t{e[i]}\n")
            n += 1
    return n


def write_fasta(ann: SynthAnnotation, path: str, seed: int = 0, line_len: int = 60, contigs: Optional[Sequence[int]] = None,
                extra: Sequence[Tuple[str, int]] = ()) -> None:
    """Random genome for the annotation's contigs (+ `extra` (name, length) contigs that no feature uses) and its .fai index.
    GC content varies along each contig in 2 kb steps; a sprinkle of N and lower-case bases exercises gc()'s character
    classes.  `contigs` selects which annotation contigs are written (default all), in the given order."""
    rng = np.random.default_rng(seed)
    todo = [(ann.contig_names[c], int(ann.contig_lens[c])) for c in (range(len(ann.contig_names)) if contigs is None else contigs)]
    todo += [(n, int(l)) for n, l in extra]
    with open(path, "wb") as f, open(path + ".fai", "w") as fai:
        off = 0
        for name, length in todo:
            hdr = f">{name} synthetic\n".encode()
            f.write(hdr)
            off += len(hdr)
            steps = (length + 1999) // 2000
            gc_frac = np.repeat(rng.uniform(0.25, 0.75, steps), 2000)[:length]
            u = rng.random(length)
            is_gc = u < gc_frac
            pick = rng.random(length) < 0.5
            seq = np.where(is_gc, np.where(pick, ord("G"), ord("C")), np.where(pick, ord("A"), ord("T"))).astype(np.uint8)
            seq[rng.random(length) < 0.002] = ord("N")
            lower = rng.random(length) < 0.05
            seq[lower] |= 0x20
            full, rest = divmod(length, line_len)
            body = bytearray()
            if full:
                lines = np.empty((full, line_len + 1), np.uint8)
                lines[:, :line_len] = seq[:full * line_len].reshape(full, line_len)
                lines[:, line_len] = 10
                body += lines.tobytes()
            if rest:
                body += seq[full * line_len:].tobytes() + b"\n"
            f.write(body)
            fai.write(f"{name}\t{length}\t{off}\t{line_len}\t{line_len + 1}\n")
            off += len(body)


# ---------------------------------------------------------------------------------------------------------------
# read simulation (torch; device-agnostic)
# ---------------------------------------------------------------------------------------------------------------
MAXB = 8  # blocks kept per read before the indel/soft-clip edits; a 150 bp read over >=20 bp exons needs at most 8


@dataclasses.dataclass
class SynthReads:
    """Coordinate-sorted records as torch tensors on `device` (SoA); pack() gives the rsq_read/rsq_block bytes."""
    n: int
    contig: torch.Tensor; pos: torch.Tensor; mpos: torch.Tensor; endpos: torch.Tensor; isize: torch.Tensor
    l_qseq: torch.Tensor; nm: torch.Tensor; flag: torch.Tensor; mapq: torch.Tensor; aux: torch.Tensor
    name_id: torch.Tensor        # int64 pair id (qname "r<id>")
    n_blocks: torch.Tensor       # int32
    blk_start: torch.Tensor      # [n, MAXB+1] int32 (1-based), valid up to n_blocks
    blk_len: torch.Tensor        # [n, MAXB+1] int32
    lead_soft: torch.Tensor      # int32: leading S length
    ins_len: torch.Tensor        # int32: inserted bases between block 0 and 1 when they abut
    mate_contig: torch.Tensor    # int32


def _splitmix64(x: torch.Tensor) -> torch.Tensor:
    """64-bit mix in int64 two's complement arithmetic (wrapping), bit-identical on CPU and CUDA."""
    def srl(v, k):
        return (v >> k) & ((1 << (64 - k)) - 1)
    x = x + (-7046029254386353131)            # 0x9E3779B97F4A7C15
    x = (x ^ srl(x, 30)) * (-4658895280553007687)   # 0xBF58476D1CE4E5B9
    x = (x ^ srl(x, 27)) * (-7723592293110705685)   # 0x94D049BB133111EB
    return x ^ srl(x, 31)


def gene_weights(ann: SynthAnnotation, seed: int = 20260925) -> np.ndarray:
    """Sampling weight of every gene: log-normal expression (sigma 2) x usable transcript length."""
    ex_len = ann.exon_end - ann.exon_start + 1
    ex_cum = np.concatenate([[0], np.cumsum(ex_len)])
    t_len = ex_cum[ann.gene_exon_off[1:]] - ex_cum[ann.gene_exon_off[:-1]]
    expr = np.exp(np.random.default_rng(seed + 1).normal(0.0, 2.0, size=ann.n_genes))
    return expr * np.maximum(t_len - 150, 0)


def simulate_reads(ann: SynthAnnotation, n_records: int, seed: int = 20260925, device: str = "cpu",
                   genomic_frac: float = 0.15, sort: bool = True, dup_frac: float = 0.05,
                   region: Optional[dict] = None, name_base: int = 0, weight_seed: int = 20260925) -> SynthReads:
    """region (multi-GPU, rnaseqc_b200.shard.rank_region): {"gene_lo", "gene_hi", "lin_lo", "lin_hi"} restricts the
    draw to one coordinate slice of the genome; name_base offsets the pair ids so read names stay unique."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    n_pairs = (n_records + 1) // 2
    rand = lambda *shape: torch.rand(*shape, generator=g, device=dev)
    randn = lambda *shape: torch.randn(*shape, generator=g, device=dev)

    # ---- transcript tables --------------------------------------------------------------------------------------
    ex_len_np = (ann.exon_end - ann.exon_start + 1)
    ex_cum_np = np.concatenate([[0], np.cumsum(ex_len_np)])            # global stitched coordinate of every exon
    t_len_np = ex_cum_np[ann.gene_exon_off[1:]] - ex_cum_np[ann.gene_exon_off[:-1]]
    w = gene_weights(ann, weight_seed)                                  # genes shorter than a read get no fragments
    if region is not None:
        mask = np.zeros(ann.n_genes, bool)
        mask[region["gene_lo"]:region["gene_hi"]] = True
        w = np.where(mask, w, 0.0)
    if w.sum() <= 0:
        w = np.ones(ann.n_genes)
        genomic_frac = 1.0
    w = w / w.sum()
    ex_start = torch.from_numpy(ann.exon_start).to(dev)
    ex_cum = torch.from_numpy(ex_cum_np).to(dev)
    g_exoff = torch.from_numpy(ann.gene_exon_off).to(dev)
    t_len = torch.from_numpy(t_len_np).to(dev)
    g_contig = torch.from_numpy(ann.gene_contig.astype(np.int64)).to(dev)
    g_strand = torch.from_numpy(ann.gene_strand.astype(np.int64)).to(dev)
    wcum = torch.from_numpy(np.cumsum(w)).to(dev)

    is_genomic = rand(n_pairs) < genomic_frac
    # ---- transcript-derived fragments ---------------------------------------------------------------------------
    gene = torch.searchsorted(wcum, rand(n_pairs).double()).clamp_(max=ann.n_genes - 1)
    T = t_len[gene]
    flen = (250.0 + 60.0 * randn(n_pairs)).round().long().clamp_(min=150)
    flen = torch.minimum(flen, T)
    fstart = (rand(n_pairs).double() * (T - flen + 1).double()).long().clamp_(min=0)
    base = ex_cum[g_exoff[gene]]                                        # global stitched coord of the transcript start

    def project(t0: torch.Tensor, length: int):
        """stitched [t0, t0+length) of each pair's gene -> up to MAXB genomic blocks."""
        a = base + t0
        b = a + length
        e0 = torch.searchsorted(ex_cum, a, right=True) - 1              # exon containing a
        starts = torch.zeros(n_pairs, MAXB + 1, dtype=torch.int64, device=dev)
        lens = torch.zeros(n_pairs, MAXB + 1, dtype=torch.int64, device=dev)
        nb = torch.zeros(n_pairs, dtype=torch.int64, device=dev)
        last_exon = g_exoff[gene + 1] - 1
        for j in range(MAXB):
            e = torch.minimum(e0 + j, last_exon)
            lo = torch.maximum(a, ex_cum[e])
            hi = torch.minimum(b, ex_cum[e + 1])
            ok = (hi > lo) & (e0 + j <= last_exon)
            starts[:, j] = torch.where(ok, ex_start[e] + (lo - ex_cum[e]), starts[:, j])
            lens[:, j] = torch.where(ok, hi - lo, lens[:, j])
            nb += ok.long()
        # merge abutting exons?  collapsed exons of one gene never abut (introns >= 30 bp), nothing to merge
        return starts, lens, nb

    ls, ll, lnb = project(fstart, 150)
    rs, rl, rnb = project(fstart + flen - 150, 150)
    # ---- genomic fragments ----------------------------------------------------------------------------------------
    clen = torch.tensor(ann.contig_lens, dtype=torch.float64, device=dev)
    coff = torch.cumsum(clen, 0) - clen                                  # linear genome coordinate of every contig start
    lin_lo = float(region["lin_lo"]) if region is not None else 0.0
    lin_hi = float(region["lin_hi"]) if region is not None else float(clen.sum().item())
    lin = lin_lo + rand(n_pairs).double() * max(lin_hi - lin_lo - 1.0, 1.0)
    gc = (torch.searchsorted(coff, lin, right=True) - 1).clamp_(min=0, max=len(ann.contig_lens) - 1)
    gflen = (250.0 + 60.0 * randn(n_pairs)).round().long().clamp_(min=150)
    gpos = (lin - coff[gc]).long()
    gpos = torch.minimum(gpos, (clen[gc] - gflen.double() - 1000).long()).clamp_(min=1000)
    if region is not None:   # keep the fragment inside the slice (slices are cut where no gene lies)
        gpos = torch.minimum(gpos, (lin_hi - coff[gc] - gflen.double() - 10).long()).clamp_(min=1)
        gpos = torch.maximum(gpos, (lin_lo - coff[gc]).long().clamp_(min=1))
    gi = is_genomic
    ls[gi] = 0; ll[gi] = 0; rs[gi] = 0; rl[gi] = 0
    ls[gi, 0] = gpos[gi]; ll[gi, 0] = 150
    rs[gi, 0] = (gpos + gflen - 150)[gi]; rl[gi, 0] = 150
    lnb = torch.where(gi, torch.ones_like(lnb), lnb)
    rnb = torch.where(gi, torch.ones_like(rnb), rnb)
    contig_p = torch.where(gi, gc, g_contig[gene])
    flen = torch.where(gi, gflen, flen)

    # ---- pair-level attributes ------------------------------------------------------------------------------------
    read1_left = rand(n_pairs) < 0.5
    dup = rand(n_pairs) < dup_frac
    mate_unmapped = rand(n_pairs) < 0.01
    mate_elsewhere = (rand(n_pairs) < 0.003) & ~mate_unmapped
    pair_id = torch.arange(n_pairs, device=dev, dtype=torch.int64) + name_base

    def side(starts, lens, nb, is_left):
        n = n_pairs
        pos0 = starts[:, 0] - 1
        lastidx = (nb - 1).clamp_(min=0)
        end = (starts.gather(1, lastidx[:, None]) + lens.gather(1, lastidx[:, None]))[:, 0] - 1   # 0-based exclusive
        is_r1 = read1_left if is_left else ~read1_left
        flag = torch.full((n,), 0x1, dtype=torch.int64, device=dev)
        flag |= torch.where(is_r1, 0x40, 0x80)
        flag |= 0x20 if is_left else 0x10                                # left read forward / mate reverse, right read reverse
        proper = ~mate_unmapped & ~mate_elsewhere
        flag |= torch.where(proper, 0x2, 0)
        flag |= torch.where(dup, 0x400, 0)
        flag |= torch.where(mate_unmapped, 0x8, 0)
        if is_left:
            flag = torch.where(mate_unmapped, flag & ~0x20, flag)
        return pos0, end, flag

    lpos, lend, lflag = side(ls, ll, lnb, True)
    rpos, rend, rflag = side(rs, rl, rnb, False)
    # when the mate is unmapped only the left read is emitted as mapped; the right record is dropped below
    keep_right = ~mate_unmapped
    other = ((contig_p + 1) % max(len(ann.contig_lens), 1)).long()
    elsewhere_pos = (rand(n_pairs) * 1e6).long() + 1000

    def cat(a, b):
        return torch.cat([a, b[keep_right]])

    contig = cat(contig_p, contig_p)
    pos = cat(lpos, rpos)
    endp = cat(lend, rend)
    flag = cat(lflag, rflag)
    mate_contig = cat(torch.where(mate_elsewhere, other, contig_p), torch.where(mate_elsewhere, other, contig_p))
    mpos = cat(torch.where(mate_elsewhere, elsewhere_pos, torch.where(mate_unmapped, lpos, rpos)),
               torch.where(mate_elsewhere, elsewhere_pos, lpos))
    isize = cat(torch.where(mate_elsewhere | mate_unmapped, torch.zeros_like(flen), rend - lpos),
                torch.where(mate_elsewhere, torch.zeros_like(flen), -(rend - lpos)))
    name_id = cat(pair_id, pair_id)
    bs = torch.cat([ls, rs[keep_right]])
    bl = torch.cat([ll, rl[keep_right]])
    nb = cat(lnb, rnb)
    n = contig.shape[0]

    # ---- per-record attributes ---------------------------------------------------------------------------------------
    u = rand(n)
    flag = flag | torch.where(u < 0.02, 0x100, 0)                                        # secondary
    flag = flag | torch.where((u >= 0.02) & (u < 0.025), 0x800, 0)                       # supplementary
    flag = flag | torch.where((u >= 0.025) & (u < 0.035), 0x200, 0)                      # QC fail
    um = rand(n)
    mapq = torch.where(um < 0.90, 255, torch.where(um < 0.94, 3, torch.where(um < 0.97, 1, 0)))
    un = rand(n)
    nm = torch.where(un < 0.60, 0, torch.where(un < 0.80, 1, torch.where(un < 0.90, 2, torch.where(un < 0.94, 3,
         torch.where(un < 0.97, (4 + rand(n) * 3).long(), (7 + rand(n) * 4).long())))))
    has_nm = rand(n) >= 0.005
    chim = (rand(n) < 0.001)
    aux = torch.where(has_nm, AUX_HAS_NM, 0) | torch.where(chim, AUX_CHIMERIC_TAG, 0) | \
        torch.where(mate_contig != contig, AUX_MATE_OTHER_CONTIG, 0)
    nm = torch.where(has_nm, nm, torch.zeros_like(nm))
    # edits on single-block reads: 2 % soft clip (10S140M), 1 % small insertion or deletion
    ue = rand(n)
    single = nb == 1
    soft = single & (ue < 0.02)
    ins = single & (ue >= 0.02) & (ue < 0.025)
    dele = single & (ue >= 0.025) & (ue < 0.03)
    k = (1 + rand(n) * 3).long().clamp_(max=3)
    lead_soft = torch.where(soft, 10, 0)
    bs[:, 0] = torch.where(soft, bs[:, 0] + 10, bs[:, 0])
    bl[:, 0] = torch.where(soft, bl[:, 0] - 10, bl[:, 0])
    pos = torch.where(soft, pos + 10, pos)
    # insertion 70M kI (80-k)M : second block abuts the first
    bs[:, 1] = torch.where(ins, bs[:, 0] + 70, bs[:, 1])
    bl[:, 1] = torch.where(ins, 80 - k, bl[:, 1])
    bl[:, 0] = torch.where(ins, 70, bl[:, 0])
    endp = torch.where(ins, endp - k, endp)
    ins_len = torch.where(ins, k, 0)
    # deletion 70M kD 80M
    bs[:, 1] = torch.where(dele, bs[:, 0] + 70 + k, bs[:, 1])
    bl[:, 1] = torch.where(dele, 80, bl[:, 1])
    bl[:, 0] = torch.where(dele, 70, bl[:, 0])
    endp = torch.where(dele, endp + k, endp)
    nb = torch.where(ins | dele, 2, nb)
    l_qseq = torch.full((n,), 150, dtype=torch.int64, device=dev)

    i32 = lambda t: t.to(torch.int32)
    out = SynthReads(n, i32(contig), i32(pos), i32(mpos), i32(endp), i32(isize), i32(l_qseq), i32(nm), i32(flag), i32(mapq),
                     i32(aux), name_id, i32(nb), i32(bs), i32(bl), i32(lead_soft), i32(ins_len), i32(mate_contig))
    return sort_reads(out) if sort else out


_FIELDS = ["contig", "pos", "mpos", "endpos", "isize", "l_qseq", "nm", "flag", "mapq", "aux", "name_id", "n_blocks",
           "blk_start", "blk_len", "lead_soft", "ins_len", "mate_contig"]


def sort_reads(r: SynthReads) -> SynthReads:
    """stable sort by (contig, pos): coordinate order, as a BAM would be"""
    key = r.contig.long() * (1 << 32) + r.pos.long()
    order = torch.argsort(key, stable=True)
    del key
    return SynthReads(r.n, **{f: getattr(r, f)[order] for f in _FIELDS})


def concat(parts: List[SynthReads]) -> SynthReads:
    return SynthReads(sum(p.n for p in parts), **{f: torch.cat([getattr(p, f) for p in parts]) for f in _FIELDS})


def pack(reads: SynthReads, contig_index: Optional[List[int]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> (uint8 [n,48] rsq_read bytes, int32 [n_blocks,2] rsq_block) on the tensors' device.
    contig_index maps the generator's contig number to the annotation contig index of the consumer."""
    dev = reads.pos.device
    n = reads.n
    nb = reads.n_blocks.long()
    off = torch.cumsum(nb, 0) - nb
    total = int(nb.sum().item())
    j = torch.arange(reads.blk_start.shape[1], device=dev)[None, :]
    valid = j < nb[:, None]
    blocks = torch.stack([reads.blk_start[valid], reads.blk_len[valid]], dim=1).to(torch.int32).contiguous()
    assert blocks.shape[0] == total
    contig = reads.contig
    if contig_index is not None:
        contig = torch.tensor(contig_index, dtype=torch.int32, device=dev)[contig.long()]
    rec = torch.zeros(n, 12, dtype=torch.int32, device=dev)
    rec[:, 0] = reads.pos; rec[:, 1] = reads.mpos; rec[:, 2] = reads.endpos; rec[:, 3] = reads.isize
    rec[:, 4] = reads.l_qseq; rec[:, 5] = reads.nm
    rec[:, 6] = (reads.flag & 0xffff) | ((reads.mapq & 0xff) << 16) | ((reads.aux & 0x7f) << 24)
    rec[:, 7] = contig & 0xffff
    h = _splitmix64(reads.name_id)
    rec[:, 8] = (h & 0xffffffff).to(torch.int64).where((h & 0xffffffff) < (1 << 31), (h & 0xffffffff) - (1 << 32)).to(torch.int32)
    hi = (h >> 32) & 0xffffffff
    rec[:, 9] = hi.where(hi < (1 << 31), hi - (1 << 32)).to(torch.int32)
    o32 = off.where(off < (1 << 31), off - (1 << 32)).to(torch.int32)
    rec[:, 10] = o32
    rec[:, 11] = (nb & 0xffff).to(torch.int32)
    return rec.view(torch.uint8).reshape(n, 48), blocks


def to_numpy(reads_bytes: torch.Tensor, blocks: torch.Tensor) -> Tuple[np.ndarray, np.ndarray]:
    r = reads_bytes.cpu().numpy().reshape(-1).view(READ_DTYPE)
    b = blocks.cpu().numpy().reshape(-1).view(BLOCK_DTYPE)
    return r, b


def cigar_of(reads: SynthReads, i: int) -> str:
    nb = int(reads.n_blocks[i])
    s = reads.blk_start[i].tolist(); l = reads.blk_len[i].tolist()
    out = []
    if int(reads.lead_soft[i]):
        out.append(f"{int(reads.lead_soft[i])}S")
    for j in range(nb):
        out.append(f"{l[j]}M")
        if j + 1 < nb:
            gap = s[j + 1] - (s[j] + l[j])
            if gap == 0:
                out.append(f"{int(reads.ins_len[i])}I")
            elif gap <= 3:
                out.append(f"{gap}D")
            else:
                out.append(f"{gap}N")
    return "".join(out)


def write_sam(ann: SynthAnnotation, reads: SynthReads, path: str, chimeric_tag: str = "ch") -> None:
    """SAM text of the records (sequence/quality are constant: the hot path never looks at them)."""
    seq, qual = "A" * 150, "F" * 150
    contig = reads.contig.tolist(); pos = reads.pos.tolist(); mpos = reads.mpos.tolist(); isize = reads.isize.tolist()
    flag = reads.flag.tolist(); mapq = reads.mapq.tolist(); nm = reads.nm.tolist(); aux = reads.aux.tolist()
    name = reads.name_id.tolist(); mc = reads.mate_contig.tolist()
    nbl = reads.n_blocks.tolist()
    bs = reads.blk_start.tolist(); bl = reads.blk_len.tolist()
    lead = reads.lead_soft.tolist(); insl = reads.ins_len.tolist()
    with open(path, "w") as f:
        f.write("@HD\tVN:1.6\tSO:coordinate\n")
        for nme, ln in zip(ann.contig_names, ann.contig_lens):
            f.write(f"@SQ\tSN:{nme}\tLN:{ln}\n")
        for i in range(reads.n):
            parts = []
            if lead[i]:
                parts.append(f"{lead[i]}S")
            s, l = bs[i], bl[i]
            for j in range(nbl[i]):
                parts.append(f"{l[j]}M")
                if j + 1 < nbl[i]:
                    gap = s[j + 1] - (s[j] + l[j])
                    parts.append(f"{insl[i]}I" if gap == 0 else (f"{gap}D" if gap <= 3 else f"{gap}N"))
            cigar = "".join(parts)
            rnext = "=" if mc[i] == contig[i] else ann.contig_names[mc[i]]
            tags = []
            if aux[i] & AUX_HAS_NM:
                tags.append(f"NM:i:{nm[i]}")
            if aux[i] & AUX_CHIMERIC_TAG:
                tags.append(f"{chimeric_tag}:A:1")
            f.write(f"r{name[i]}\t{flag[i]}\t{ann.contig_names[contig[i]]}\t{pos[i] + 1}\t{mapq[i]}\t{cigar}\t{rnext}\t{mpos[i] + 1}\t{isize[i]}\t{seq}\t{qual}"
                    + ("\t" + "\t".join(tags) if tags else "") + "\n")


def write_bam(ann: SynthAnnotation, reads: SynthReads, path: str, level: int = 1) -> None:
    """BAM of the records through the C++ writer in librsq_host.so (fast enough for multi-million-record samples)."""
    import ctypes as C
    from . import api
    H = api.host_lib()
    c = lambda t: np.ascontiguousarray(t.cpu().numpy().astype(np.int32))
    names = (C.c_char_p * len(ann.contig_names))(*[n.encode() for n in ann.contig_names])
    lens = np.asarray(ann.contig_lens, np.int64)
    arrs = dict(tid=c(reads.contig), pos=c(reads.pos), mtid=c(reads.mate_contig), mpos=c(reads.mpos), isize=c(reads.isize),
                flag=c(reads.flag), mapq=c(reads.mapq), nm=c(reads.nm), aux=c(reads.aux), nb=c(reads.n_blocks),
                bs=c(reads.blk_start), bl=c(reads.blk_len), ls=c(reads.lead_soft), il=c(reads.ins_len))
    name_id = np.ascontiguousarray(reads.name_id.cpu().numpy().astype(np.int64))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    H.rsqh_write_bam.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.c_void_p, C.c_size_t] + [C.c_void_p] * 13 + \
        [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    rc = H.rsqh_write_bam(path.encode(), len(ann.contig_names), names, p(lens), reads.n, p(arrs["tid"]), p(arrs["pos"]),
                          p(arrs["mtid"]), p(arrs["mpos"]), p(arrs["isize"]), p(arrs["flag"]), p(arrs["mapq"]), p(arrs["nm"]),
                          p(arrs["aux"]), p(name_id), p(arrs["nb"]), p(arrs["bs"]), p(arrs["bl"]), reads.blk_start.shape[1],
                          p(arrs["ls"]), p(arrs["il"]), 150, level)
    if rc != 0:
        raise RuntimeError(f"rsqh_write_bam failed: {rc}")


def write_bam_python(ann: SynthAnnotation, reads: SynthReads, path: str, level: int = 1, l_qseq: int = 150, chunk: int = 100_000) -> None:
    """The same BAM as write_bam, byte for byte before compression, written with numpy + zlib only: `bench.py --impl reference`
    uses it so that the reference arm loads none of this repo's native libraries."""
    import struct
    import zlib
    c = lambda t: np.ascontiguousarray(t.cpu().numpy().astype(np.int64))
    tid, pos, mtid, mpos, isize, flag, mapq = c(reads.contig), c(reads.pos), c(reads.mate_contig), c(reads.mpos), c(reads.isize), c(reads.flag), c(reads.mapq)
    nm, aux, name_id, nb, ls, il = c(reads.nm), c(reads.aux), c(reads.name_id), c(reads.n_blocks), c(reads.lead_soft), c(reads.ins_len)
    bs_all, bl_all = c(reads.blk_start), c(reads.blk_len)
    stride = bs_all.shape[1]

    out = open(path, "wb")
    pending = bytearray()

    def bgzf_block(data: bytes) -> bytes:
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        comp = co.compress(data) + co.flush()
        return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(comp) + 25) + comp +
                struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data)))

    def put(data, flush=False):
        pending.extend(data)
        lim = 0xff00
        k = 0
        while len(pending) - k >= lim or (flush and k < len(pending)):
            out.write(bgzf_block(bytes(pending[k:k + lim])))
            k += lim
        del pending[:min(k, len(pending))]

    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join(f"@SQ\tSN:{n}\tLN:{l}\n" for n, l in zip(ann.contig_names, ann.contig_lens))
    hdr = b"BAM\x01" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(ann.contig_names))
    for n, l in zip(ann.contig_names, ann.contig_lens):
        hdr += struct.pack("<i", len(n) + 1) + n.encode() + b"\x00" + struct.pack("<i", int(l))
    put(hdr)

    seq_qual = np.concatenate([np.full((l_qseq + 1) // 2, 0x11, np.uint8), np.full(l_qseq, ord("F") - 33, np.uint8)])
    for a in range(0, reads.n, chunk):
        b = min(reads.n, a + chunk)
        n = b - a
        sl = slice(a, b)
        bs, bl, k_nb, k_ls = bs_all[sl], bl_all[sl], nb[sl], ls[sl]
        # CIGAR: [S] M (I|D|N M)*
        max_ops = 2 * stride
        ops = np.zeros((n, max_ops), np.uint32)
        has_s = (k_ls > 0).astype(np.int64)
        ops[np.arange(n)[has_s > 0], 0] = ((k_ls[has_s > 0] << 4) | 4).astype(np.uint32)
        ref_end = pos[sl].copy()
        for j in range(stride):
            m = k_nb > j
            if not m.any():
                break
            idx = np.nonzero(m)[0]
            ops[idx, has_s[idx] + 2 * j] = ((bl[idx, j] << 4) | 0).astype(np.uint32)
            ref_end[idx] = bs[idx, j] - 1 + bl[idx, j]
            if j + 1 < stride:
                m2 = k_nb > j + 1
                idx2 = np.nonzero(m2)[0]
                if len(idx2):
                    gap = bs[idx2, j + 1] - (bs[idx2, j] + bl[idx2, j])
                    op = np.where(gap == 0, (il[sl][idx2] << 4) | 1, np.where(gap <= 3, (gap << 4) | 2, (gap << 4) | 3))
                    ops[idx2, has_s[idx2] + 2 * j + 1] = op.astype(np.uint32)
        n_cig = has_s + np.where(k_nb > 0, 2 * k_nb - 1, 0)
        names = np.char.add("r", name_id[sl].astype(str)).astype("S")
        l_name = np.char.str_len(names) + 1
        wname = names.dtype.itemsize + 1
        name_bytes = np.zeros((n, wname), np.uint8)
        name_bytes[:, :wname - 1] = np.frombuffer(names.tobytes(), np.uint8).reshape(n, wname - 1)
        has_nm = (aux[sl] & AUX_HAS_NM) != 0
        has_ch = (aux[sl] & AUX_CHIMERIC_TAG) != 0
        size = 32 + l_name + 4 * n_cig + len(seq_qual) + 4 * has_nm + 4 * has_ch
        off = np.concatenate([[0], np.cumsum(size + 4)])
        buf = np.zeros(int(off[-1]) + 256, np.uint8)
        # bin (reg2bin of [pos, max(ref_end, pos + 1)))
        beg, end = pos[sl], np.maximum(ref_end, pos[sl] + 1) - 1
        bin_ = np.zeros(n, np.int64)
        done = np.zeros(n, bool)
        for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
            hit = ~done & ((beg >> shift) == (end >> shift))
            bin_[hit] = base + (beg[hit] >> shift)
            done |= hit
        fixed = np.zeros(n, np.dtype([("block_size", "<i4"), ("tid", "<i4"), ("pos", "<i4"), ("l_name", "u1"), ("mapq", "u1"), ("bin", "<u2"),
                                      ("n_cig", "<u2"), ("flag", "<u2"), ("l_seq", "<i4"), ("mtid", "<i4"), ("mpos", "<i4"), ("isize", "<i4")]))
        fixed["block_size"] = size; fixed["tid"] = tid[sl]; fixed["pos"] = pos[sl]; fixed["l_name"] = l_name; fixed["mapq"] = mapq[sl]
        fixed["bin"] = bin_; fixed["n_cig"] = n_cig; fixed["flag"] = flag[sl]; fixed["l_seq"] = l_qseq; fixed["mtid"] = mtid[sl]
        fixed["mpos"] = mpos[sl]; fixed["isize"] = isize[sl]

        def scatter(at, rows):
            w = rows.shape[1]
            buf[(at[:, None] + np.arange(w)[None, :]).ravel()] = rows.ravel()

        # written in record order of the fields, so that the zero padding of a variable-width field is overwritten by the next one
        scatter(off[:-1], np.frombuffer(fixed.tobytes(), np.uint8).reshape(n, 36))
        p = off[:-1] + 36
        scatter(p, name_bytes)
        p = p + l_name
        scatter(p, np.frombuffer(ops.astype("<u4").tobytes(), np.uint8).reshape(n, 4 * max_ops))
        p = p + 4 * n_cig
        scatter(p, np.broadcast_to(seq_qual, (n, len(seq_qual))))
        p = p + len(seq_qual)
        tail = np.zeros((n, 8), np.uint8)                       # aux fields; the spill of the last record lands in the slack of `buf`
        nm_rows = np.stack([np.full(n, ord("N")), np.full(n, ord("M")), np.full(n, ord("C")), nm[sl] & 0xff], axis=1).astype(np.uint8)
        ch_rows = np.broadcast_to(np.frombuffer(b"chA1", np.uint8), (n, 4))
        tail[has_nm, 0:4] = nm_rows[has_nm]
        only_ch = has_ch & ~has_nm
        tail[only_ch, 0:4] = ch_rows[only_ch]
        both = has_ch & has_nm
        tail[both, 4:8] = ch_rows[both]
        # a record's aux bytes may be 0, 4 or 8 long: write 8 and let the NEXT record's fixed part overwrite the rest — so write the
        # tails first in a second pass over a copy of the starts
        keep = buf.copy()
        scatter(p, tail)
        # restore everything that belongs to following records
        w_aux = 4 * has_nm.astype(np.int64) + 4 * has_ch.astype(np.int64)
        for extra in (0, 4):
            sel = w_aux <= extra
            if sel.any():
                at = p[sel] + extra
                idx = (at[:, None] + np.arange(4)[None, :]).ravel()
                buf[idx] = keep[idx]
        put(buf[:int(off[-1])].tobytes())
    put(b"", flush=True)
    out.write(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))   # BGZF EOF marker
    out.close()
