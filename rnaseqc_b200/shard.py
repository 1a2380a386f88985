"""Coordinate partitioning of the sorted record stream across GPUs.

Records are range-partitioned by coordinate, one contiguous range per rank, cut ONLY where no gene (and no BED
interval) lies, so that every read that can count toward a gene — and both mates of a fragment — land on the same
rank: the per-gene fragment tracker, the fragment-size mate map and the per-base coverage of a gene then have exactly
one owner, and the only exchange is one integer all-reduce of the count vectors at the end of the run.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np


def gene_free_cuts(ann, extra_intervals=None) -> Tuple[np.ndarray, np.ndarray]:
    """Candidate cut points of a SYNTHETIC annotation (rnaseqc_b200.synth): for every gene g that starts after everything
    before it on its contig has ended (or that opens a new contig), the linear genome coordinate just before it.  Returns
    (gene index, linear coordinate).  The intervals that must not be cut are the gene lines, every exon (checked to lie inside
    its gene line) and `extra_intervals` — (contig, start, end) arrays, e.g. BED intervals that are not exons.  For parsed
    annotations the C++ host model has the general rule (csrc/host/annotation.h, Islands)."""
    coff = np.concatenate([[0], np.cumsum(np.asarray(ann.contig_lens, np.int64))])
    gc = np.asarray(ann.gene_contig, np.int64)
    gs, ge = np.asarray(ann.gene_start, np.int64), np.asarray(ann.gene_end, np.int64)
    assert np.all(np.lexsort((gs, gc)) == np.arange(len(gs))) or np.all(np.diff(gc * (1 << 40) + gs) >= 0), \
        "synthetic genes are generated sorted by (contig, start)"
    eg = np.repeat(np.arange(ann.n_genes), np.diff(ann.gene_exon_off))
    assert np.all(ann.exon_start >= gs[eg]) and np.all(ann.exon_end <= ge[eg]), "an exon lies outside its gene line"
    # blocked[i] = linear intervals no cut may fall into, besides the gene lines
    xs = np.zeros(0, np.int64); xe = np.zeros(0, np.int64)
    if extra_intervals is not None:
        xc, s_, e_ = [np.asarray(a, np.int64) for a in extra_intervals]
        xs, xe = coff[xc] + s_, coff[xc] + e_
        o = np.argsort(xs); xs, xe = xs[o], xe[o]
    cuts_g, cuts_lin = [0], [0]
    run_end, run_contig = -1, -1
    for g in range(ann.n_genes):
        c = int(gc[g])
        cand = None
        if c != run_contig:
            if g > 0:
                cand = int(coff[c])
            run_contig, run_end = c, -1
        elif gs[g] > run_end + 1 and g > 0:
            cand = int(coff[c] + (run_end + gs[g]) // 2)
        if cand is not None:
            k = int(np.searchsorted(xs, cand, side="right")) - 1     # an extra interval over the candidate vetoes it
            if not (len(xs) and k >= 0 and np.max(xe[:k + 1]) >= cand - 1):
                cuts_g.append(g); cuts_lin.append(cand)
        run_end = max(run_end, int(ge[g]))
    return np.asarray(cuts_g), np.asarray(cuts_lin)


def rank_region(ann, rank: int, world: int, genomic_frac: float = 0.15, weight_seed: int = 20260925) -> Dict[str, int]:
    """The slice of the genome owned by `rank`: balances the expected record mass (transcript-derived + uniform
    genomic) over ranks, cutting only at gene-free positions."""
    from . import synth
    w = synth.gene_weights(ann, weight_seed)
    w = w / w.sum()
    cuts_g, cuts_lin = gene_free_cuts(ann)
    total_len = float(sum(ann.contig_lens))
    cum_w = np.concatenate([[0.0], np.cumsum(w)])
    mass = (1 - genomic_frac) * cum_w[cuts_g] + genomic_frac * cuts_lin / total_len     # mass left of every candidate cut
    bounds_g, bounds_lin = [0], [0]
    for r in range(1, world):
        k = int(np.argmin(np.abs(mass - r / world)))
        k = max(k, 1)
        bounds_g.append(int(cuts_g[k])); bounds_lin.append(int(cuts_lin[k]))
    bounds_g.append(ann.n_genes); bounds_lin.append(int(total_len))
    # enforce monotone bounds (degenerate annotations)
    for i in range(1, len(bounds_g)):
        bounds_g[i] = max(bounds_g[i], bounds_g[i - 1]); bounds_lin[i] = max(bounds_lin[i], bounds_lin[i - 1])
    return {"gene_lo": bounds_g[rank], "gene_hi": bounds_g[rank + 1], "lin_lo": bounds_lin[rank], "lin_hi": bounds_lin[rank + 1]}


def split_sorted_records(ann, contig, pos, world: int, return_cuts: bool = False, extra_intervals=None):
    """[lo, hi) record ranges of an EXISTING coordinate-sorted stream (numpy arrays of generator contig number and
    0-based pos), one per rank, balanced by record count and cut only at gene-free coordinates.  return_cuts: also the
    world-1 interior cuts as linear coordinates (for owned_gene_lines)."""
    coff = np.concatenate([[0], np.cumsum(np.asarray(ann.contig_lens, np.int64))])
    lin = coff[np.asarray(contig, np.int64)] + np.asarray(pos, np.int64)
    n = len(lin)
    _, cuts_lin = gene_free_cuts(ann, extra_intervals)
    # a record belongs left of cut x iff lin + 1 (1-based start) < x; reads that START before a cut and reach past it cannot
    # exist unless they span a gene-free gap, where they count toward nothing
    idx = np.searchsorted(lin + 1, cuts_lin, side="left")
    out, lo, chosen = [], 0, []
    for r in range(1, world):
        k = int(np.argmin(np.abs(idx - r * n / world)))
        hi = max(int(idx[k]), lo)
        if hi == lo and chosen:
            chosen.append(chosen[-1])            # an empty partition: repeat the previous cut
        else:
            chosen.append(int(cuts_lin[k]))
        out.append((lo, hi)); lo = hi
    out.append((lo, n))
    return (out, chosen) if return_cuts else out


def stream_offsets(api_ann, synth_ann) -> np.ndarray:
    """linear offset of every ANNOTATION contig id (api.Annotation, first-seen order of the GTF/BED) in the stream order of
    the synthetic genome; -1 for contigs the stream does not have"""
    coff = np.concatenate([[0], np.cumsum(np.asarray(synth_ann.contig_lens, np.int64))])
    where = {n: i for i, n in enumerate(synth_ann.contig_names)}
    return np.asarray([coff[where[n]] if n in where else -1 for n in api_ann.contig_names()], np.int64)


def owned_gene_lines(api_ann, synth_ann, cut_lins, rank: int) -> np.ndarray:
    """rsq_transcripts.gene_owned of partition `rank`: gene line i belongs to the partition its start falls into
    (cut_lins: the world-1 interior cuts as linear coordinates, ascending)."""
    c, s, _ = api_ann.gene_lines()
    off = stream_offsets(api_ann, synth_ann)
    lin = off[c] + s
    owner = np.searchsorted(np.asarray(cut_lins, np.int64), lin, side="right")
    owner[off[c] < 0] = 0
    return (owner == rank).astype(np.uint8)


def assert_cuts_legal(api_ann, synth_ann, cut_lins) -> None:
    """no island of the PARSED annotation (gene hulls incl. exons outside gene lines, BED intervals) contains a cut"""
    c, s, e = api_ann.islands()
    off = stream_offsets(api_ann, synth_ann)
    lo, hi = off[c] + s, off[c] + e
    for x in cut_lins:
        bad = (lo <= x) & (hi >= x - 1)
        assert not np.any(bad), f"cut {x} falls into an island"
