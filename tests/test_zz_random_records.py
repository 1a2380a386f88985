"""Odd records through the C-ABI: the random SAM of tests/test_sam_fuzz.py (any flag combination, CIGARs without a single aligned
base, mates on other contigs, unmapped records with coordinates, contigs the GTF lacks) decoded by the host library and run
through a backend, against the CPU oracle on the same records — which tests/test_sam_fuzz.py pins to the reference binary on
these very inputs.  The CPU variant checks the harness with the oracle on both sides and the wire-format packer; the `-m gpu`
variant is the CUDA path under the current and the --legacy rules.  The pair cases below also give three and four records one read
name: the mate maps replay every name's candidates in stream order (rsq_kernels.cu, mate_* kernels)."""
import os
import random

import numpy as np
import pytest

from test_sam_fuzz import _fasta, _gtf, _pairs_sam, _sam
from util import assert_results_equal

CONFIGS = {
    "default": dict(),
    "rf_unpaired": dict(strand_mode=2, unpaired=True),
    "excl_loose": dict(exclude_chimeric=True, mapq=3, mismatches=5, chimeric_distance=1000),
    "legacy": dict(legacy=True, mapq=4),
    "legacy_fr_unpaired_excl": dict(legacy=True, strand_mode=1, unpaired=True, exclude_chimeric=True),
}


def _case(api, d, seed, legacy):
    rng = random.Random(seed)
    gtf, sam, bed = os.path.join(d, "f.gtf"), os.path.join(d, "f.sam"), os.path.join(d, "f.bed")
    rows = _gtf(gtf, rng)
    _sam(sam, rng, 2500, allow_dup=True)
    ex = [r.split("\t") for r in rows if "\texon\t" in r]
    open(bed, "w").write("".join(f"{e[0]}\t{int(e[3]) - 1}\t{e[4]}\n" for e in ex[::3]))
    ann = api.Annotation(gtf, bed, legacy=legacy)
    return ann, sam


def _run(api, lib, ann, reads, blocks, cfg, packed=False):
    ctx = api.Backend(lib).create(cfg)
    ctx.set_annotation(ann)
    if packed:
        ctx.submit_packed(*api.pack_records(reads, blocks))
    else:
        ctx.submit(reads, blocks)
    ctx.sync()
    res = ctx.finalize()
    nbytes = ctx.algorithmic_bytes()
    ctx.close()
    return res, nbytes


def _check(libs, tmp_path, lib, seed, name):
    api = libs["api"]
    kw = dict(CONFIGS[name])
    legacy = kw.pop("legacy", False)
    ann, sam = _case(api, str(tmp_path), seed, legacy)
    reads, blocks, _ = ann.decode(sam, exclude_chimeric=kw.get("exclude_chimeric", False), chimeric_distance=kw.get("chimeric_distance", 2000000))
    assert len(reads) == 2500 and int((reads["n_blocks"] == 0).sum()) > 20          # records without any aligned block are in
    cfg = lambda: api.make_config(max_records=len(reads), legacy=legacy, **kw)
    want, want_bytes = _run(api, libs["oracle"], ann, reads, blocks, cfg())
    got, got_bytes = _run(api, lib, ann, reads, blocks, cfg())
    assert_results_equal(got, want, f"{name} seed {seed}")
    assert got_bytes == want_bytes
    assert want.counter("Exonic Reads") > 50 and want.counter("Intergenic Reads") > 50
    return api, ann, reads, blocks, cfg, want


@pytest.mark.parametrize("name", ["default", "legacy"])
def test_harness_and_packer_on_the_oracle(libs, tmp_path, name):
    api, ann, reads, blocks, cfg, want = _check(libs, tmp_path, libs["oracle"], 5021, name)
    # the wire format refuses batches it cannot express (a mapped record without reference span ends AT its position under
    # htslib 1.5, not one past it); a batch without such records packs and gives the same tables
    assert api.pack_records(reads, blocks) is None
    last = np.maximum(reads["block_off"].astype(np.int64) + reads["n_blocks"] - 1, 0)
    last_end = blocks["start"][last].astype(np.int64) - 1 + blocks["len"][last] if len(blocks) else np.zeros(len(reads), np.int64)
    keep = np.where(reads["n_blocks"] == 0, reads["endpos"] == reads["pos"] + 1, reads["endpos"] == last_end)
    assert 1500 < int(keep.sum()) < len(reads)
    sub = reads[keep]
    w1, _ = _run(api, libs["oracle"], ann, sub, blocks, cfg())
    w2, _ = _run(api, libs["oracle"], ann, sub, blocks, cfg(), packed=True)
    assert_results_equal(w2, w1, "packed")


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [5021, 5038, 4100])
@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_cuda_matches_oracle_on_random_sam_records(libs, tmp_path, name, seed):
    _check(libs, tmp_path, libs["product"], seed, name)


PAIR_CONFIGS = {
    "bed": dict(fragment_samples=1000000),
    "bed_quota": dict(fragment_samples=25),
    "bed_fasta_fr": dict(fragment_samples=300, strand_mode=1, fasta=True),
    "fasta_unpaired_loose": dict(unpaired=True, mapq=60, mismatches=0, fasta=True),
    "legacy_bed_fasta": dict(legacy=True, mapq=4, fragment_samples=1000000, fasta=True),
}


def _check_pairs(libs, tmp_path, lib, seed, name):
    """Odd PAIRS (tests/test_sam_fuzz.py::_pairs_sam, incl. names shared by three and four records) through the fragment-size and
    GC-content samplers."""
    api = libs["api"]
    kw = dict(PAIR_CONFIGS[name])
    legacy, fasta = kw.pop("legacy", False), kw.pop("fasta", False)
    rng = random.Random(seed)
    d = str(tmp_path)
    gtf, sam, bed, fa = (os.path.join(d, n) for n in ("f.gtf", "f.sam", "f.bed", "f.fa"))
    rows = _gtf(gtf, rng)
    _pairs_sam(sam, rng, rows, 1500, shared_names=True)
    ex = sorted({(e[0], int(e[3]), int(e[4])) for e in (r.split("\t") for r in rows if "\texon\t" in r)})
    open(bed, "w").write("".join(f"{c}\t{s - 1}\t{e}\n" for c, s, e in ex))
    if fasta:
        _fasta(fa, [("chr1", 100000), ("chr2", 100000), ("chrM", 100000)])
    ann = api.Annotation(gtf, bed, legacy=legacy, fasta=fa if fasta else None)
    reads, blocks, _ = ann.decode(sam)
    cfg = lambda: api.make_config(max_records=len(reads), legacy=legacy, **kw)
    out = []
    for which in (libs["oracle"], lib):
        ctx = api.Backend(which).create(cfg())
        ctx.set_annotation(ann)
        ctx.submit(reads, blocks)
        ctx.sync()
        out.append(ctx.finalize())
        ctx.close()
    want, got = out
    assert_results_equal(got, want, f"{name} seed {seed}")
    assert np.array_equal(got.gc_hist, want.gc_hist), (name, seed, got.gc_hist, want.gc_hist)
    if kw.get("fragment_samples"):
        assert 20 <= int(want.fragment_hist.sum()) <= kw["fragment_samples"]
    if fasta and not legacy:
        assert int(want.gc_hist.sum()) > 20


@pytest.mark.parametrize("name", ["bed_fasta_fr", "legacy_bed_fasta"])
def test_pairs_harness_on_the_oracle(libs, tmp_path, name):
    _check_pairs(libs, tmp_path, libs["oracle"], 4301, name)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [4300, 4301])
@pytest.mark.parametrize("name", sorted(PAIR_CONFIGS))
def test_cuda_samplers_match_oracle_on_odd_pairs(libs, tmp_path, name, seed):
    _check_pairs(libs, tmp_path, libs["product"], seed, name)


@pytest.mark.parametrize("seed", [5021, 5038, 4100])
@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_host_emulation_of_the_device_logic_on_the_same_records(libs, tmp_path, name, seed):
    """What the -m gpu variant above feeds the CUDA kernels, through the kernels' own per-record sources compiled for the CPU
    (tests/hostcore): whatever the GPU run may still find can only be in the warp plumbing."""
    import test_hostcore as T
    api = libs["api"]
    kw = dict(CONFIGS[name])
    legacy = kw.pop("legacy", False)
    ann, sam = _case(api, str(tmp_path), seed, legacy)
    reads, blocks, _ = ann.decode(sam, exclude_chimeric=kw.get("exclude_chimeric", False), chimeric_distance=kw.get("chimeric_distance", 2000000))
    cfg = lambda: api.make_config(max_records=len(reads), legacy=legacy, **kw)
    want, want_bytes = T.run_oracle(api, libs["oracle"], ann, reads, blocks, cfg())
    if legacy:
        got, got_bytes = T.run_hostlegacy(api, ann, reads, blocks, cfg())
    else:
        got, got_bytes, _ = T.run_hostcore(api, ann, reads, blocks, cfg())
    T.assert_same(api, got, want, f"{name} seed {seed}")
    assert got_bytes == want_bytes
