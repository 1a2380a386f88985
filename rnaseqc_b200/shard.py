"""Coordinate partitioning of the sorted record stream across GPUs.

Records are range-partitioned by coordinate, one contiguous range per rank, cut ONLY where no gene (and no BED
interval) lies, so that every read that can count toward a gene — and both mates of a fragment — land on the same
rank: the per-gene fragment tracker, the fragment-size mate map and the per-base coverage of a gene then have exactly
one owner, and the only exchange is one integer all-reduce of the count vectors at the end of the run.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np


def gene_free_cuts(ann) -> Tuple[np.ndarray, np.ndarray]:
    """Candidate cut points: for every gene g that starts after everything before it on its contig has ended (or that
    opens a new contig), the linear genome coordinate just before it.  Returns (gene index, linear coordinate)."""
    coff = np.concatenate([[0], np.cumsum(np.asarray(ann.contig_lens, np.int64))])
    cuts_g, cuts_lin = [0], [0]
    run_end, run_contig = -1, -1
    for g in range(ann.n_genes):
        c = int(ann.gene_contig[g])
        if c != run_contig:
            if g > 0:
                cuts_g.append(g); cuts_lin.append(int(coff[c]))
            run_contig, run_end = c, -1
        elif ann.gene_start[g] > run_end + 1 and g > 0:
            cuts_g.append(g); cuts_lin.append(int(coff[c] + (run_end + ann.gene_start[g]) // 2))
        run_end = max(run_end, int(ann.gene_end[g]))
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


def split_sorted_records(ann, contig, pos, world: int) -> List[Tuple[int, int]]:
    """[lo, hi) record ranges of an EXISTING coordinate-sorted stream (numpy arrays of generator contig number and
    0-based pos), one per rank, balanced by record count and cut only at gene-free coordinates."""
    coff = np.concatenate([[0], np.cumsum(np.asarray(ann.contig_lens, np.int64))])
    lin = coff[np.asarray(contig, np.int64)] + np.asarray(pos, np.int64)
    n = len(lin)
    _, cuts_lin = gene_free_cuts(ann)
    # a record belongs left of cut x iff lin + 1 (1-based start) < x; reads that START before a cut and reach past it cannot
    # exist unless they span a gene-free gap, where they count toward nothing
    idx = np.searchsorted(lin + 1, cuts_lin, side="left")
    out, lo = [], 0
    for r in range(1, world):
        k = int(np.argmin(np.abs(idx - r * n / world)))
        hi = max(int(idx[k]), lo)
        out.append((lo, hi)); lo = hi
    out.append((lo, n))
    return out
