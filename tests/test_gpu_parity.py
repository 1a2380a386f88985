"""GPU tests (`-m gpu`): the CUDA hot path, called through the C-ABI, against the CPU oracle on the same seeded
inputs (bit-exact: counters, gene tables, exact Q64.64 exon sums, per-base coverage, fragment-size histogram) and
against the golden outputs of the unmodified reference binary."""
import os
import sys

import numpy as np
import pytest

from util import assert_results_equal, compare_outputs, run_cli, unpack_case

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from conftest import SCALE, needs_device_pointers, scaled  # noqa: E402
from make_golden import KAT_FLAGSETS, KAT_LEGACY_FLAGSETS, KAT_ODD_FLAGSETS, SYNTH_FASTA_FLAGSETS, SYNTH_FLAGSETS  # noqa: E402

pytestmark = pytest.mark.gpu


def _run(api, lib, ann, reads, blocks, cfg, chunks=1):
    ctx = api.Backend(lib).create(cfg)
    ctx.set_annotation(ann)
    n = len(reads)
    bounds = [n * k // chunks for k in range(chunks + 1)]
    for a, b in zip(bounds[:-1], bounds[1:]):
        if a == b:
            continue
        r = reads[a:b].copy()
        b0 = int(r["block_off"][0])
        b1 = int(r["block_off"][-1]) + int(r["n_blocks"][-1])
        r["block_off"] -= b0
        ctx.submit(r, blocks[b0:b1].copy())
    ctx.sync()
    res = ctx.finalize()
    nbytes = ctx.algorithmic_bytes()
    ctx.close()
    return res, nbytes


@pytest.fixture(scope="module")
def synth_case(libs, tmp_path_factory):
    from rnaseqc_b200 import synth
    api = libs["api"]
    d = str(tmp_path_factory.mktemp("synth"))
    sa = synth.make_annotation(seed=101, n_genes=400, contig_names=["chr1", "chr2", "chr3", "chrM"],
                               contig_lens=[40_000_000, 30_000_000, 20_000_000, 16569])
    synth.write_gtf(sa, os.path.join(d, "a.gtf"))
    synth.write_bed(sa, os.path.join(d, "a.bed"), min_len=400)
    ann = api.Annotation(os.path.join(d, "a.gtf"), os.path.join(d, "a.bed"))
    sr = synth.simulate_reads(sa, scaled(400000), seed=102)
    rb, bb = synth.pack(sr, [ann.contig_index(n) for n in sa.contig_names])
    reads, blocks = synth.to_numpy(rb, bb)
    return ann, reads, blocks


CONFIGS = {
    "default": dict(),
    "bed": dict(fragment_samples=1000000),
    "bed_few_samples": dict(fragment_samples=777),
    "stranded_fr": dict(strand_mode=1),
    "stranded_rf": dict(strand_mode=2),
    "unpaired": dict(unpaired=True),
    "exclude_chimeric": dict(exclude_chimeric=True),
    "loose": dict(mapq=3, mismatches=2, chimeric_distance=1000),
}


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_cuda_matches_oracle_on_synthetic_records(libs, synth_case, name):
    api = libs["api"]
    ann, reads, blocks = synth_case
    cfg = lambda: api.make_config(max_records=len(reads), **CONFIGS[name])
    want, want_bytes = _run(api, libs["oracle"], ann, reads, blocks, cfg())
    got, got_bytes = _run(api, libs["product"], ann, reads, blocks, cfg())
    assert_results_equal(got, want, name)
    assert got_bytes == want_bytes
    assert want.counter("Exonic Reads") > 100000 * SCALE and want.gene_reads.sum() > 100000 * SCALE
    if "bed" in name:
        assert want.fragment_samples_taken > 0


def test_batching_does_not_change_results(libs, synth_case):
    api = libs["api"]
    ann, reads, blocks = synth_case
    one, _ = _run(api, libs["product"], ann, reads, blocks, api.make_config(fragment_samples=5000, max_records=len(reads)))
    many, _ = _run(api, libs["product"], ann, reads, blocks, api.make_config(fragment_samples=5000, max_records=1000), chunks=7)
    assert_results_equal(many, one, "7 submits vs 1")


def test_filter_tags_and_aux_bits(libs, synth_case):
    api = libs["api"]
    ann, reads, blocks = synth_case
    r = reads.copy()
    rng = np.random.default_rng(5)
    r["filter_mask"] = (rng.random(len(r)) < 0.05) * rng.integers(1, 4, len(r))
    r["aux"] |= ((rng.random(len(r)) < 0.01) * 0x08).astype(np.uint8)     # tid out of range
    r["aux"] |= ((rng.random(len(r)) < 0.02) * 0x10).astype(np.uint8)     # contig already dropped
    r["contig"][rng.random(len(r)) < 0.01] = -1
    cfg = lambda: api.make_config(n_filter_tags=2, max_records=len(r))
    want, _ = _run(api, libs["oracle"], ann, r, blocks, cfg())
    got, _ = _run(api, libs["product"], ann, r, blocks, cfg())
    assert_results_equal(got, want, "filter tags")
    assert want.counters[api.COUNTER_INDEX["Filtered by tag 0"]] > 0


def test_empty_and_tiny_batches(libs, synth_case):
    api = libs["api"]
    ann, reads, blocks = synth_case
    ctx = api.Backend(libs["product"]).create(api.make_config())
    ctx.set_annotation(ann)
    ctx.submit(reads[:0].copy(), blocks[:0].copy())
    res = ctx.finalize()
    assert res.counters.sum() == 0 and res.gene_reads.sum() == 0
    want, _ = _run(api, libs["oracle"], ann, reads[:33], blocks, api.make_config())
    got, _ = _run(api, libs["product"], ann, reads[:33], blocks, api.make_config())
    assert_results_equal(got, want, "33 reads")
    # reset gives a clean slate
    ctx.submit(reads[:1000].copy(), blocks)
    ctx.reset()
    ctx.submit(reads[:33].copy(), blocks)
    assert_results_equal(ctx.finalize(), want, "after reset")
    # a context that is reset again and again swaps between two fragment-tracker tables, the idle one zeroed on a side stream
    # (rsq_capi.cu: fresh_frag_table): every run must start from an empty set, with and without a finalize in between
    want_1000, _ = _run(api, libs["oracle"], ann, reads[:1000], blocks, api.make_config())
    for rep in range(6):
        ctx.reset()
        if rep % 2:
            ctx.submit(reads[:1000].copy(), blocks)
            assert_results_equal(ctx.finalize(), want_1000, f"reset {rep}: 1000 reads")
        else:
            ctx.submit(reads[:500].copy(), blocks)       # left without a finalize: the next reset follows queued kernels
            ctx.reset()
            ctx.submit(reads[:33].copy(), blocks)
            assert_results_equal(ctx.finalize(), want, f"reset {rep}: 33 reads")


@pytest.mark.parametrize("flagset", sorted(KAT_FLAGSETS))
def test_kat_product_cli_matches_reference(libs, tmp_path, flagset):
    d = unpack_case("kat", tmp_path)
    out = os.path.join(d, "out_" + flagset)
    run_cli(libs["cli"], os.path.join(d, "kat.gtf"), os.path.join(d, "kat.sam"), out, *KAT_FLAGSETS[flagset])
    names = set(compare_outputs(os.path.join(d, "expected", flagset), out))   # integers exact, floats 1e-6 (approx_diff semantics)
    assert any(n.endswith("gene_reads.gct") for n in names) and any(n.endswith("gene_fragments.gct") for n in names), names


@pytest.mark.parametrize("flagset", sorted(SYNTH_FLAGSETS))
def test_synth_small_product_cli_matches_reference(libs, tmp_path, flagset):
    d = unpack_case("synth_small", tmp_path)
    out = os.path.join(d, "out_" + flagset)
    flags = [f if f != "a.bed" else os.path.join(d, "a.bed") for f in SYNTH_FLAGSETS[flagset]]
    run_cli(libs["cli"], os.path.join(d, "a.gtf"), os.path.join(d, "a.sam"), out, *flags)
    identical = set(compare_outputs(os.path.join(d, "expected", flagset), out))
    assert {"a.sam.gene_reads.gct", "a.sam.gene_fragments.gct", "a.sam.fragmentSizes.txt"} <= identical, identical


@needs_device_pointers()
def test_device_resident_submit_and_export_roundtrip(libs, synth_case):
    import torch
    api = libs["api"]
    ann, reads, blocks = synth_case
    quota = 5000
    want, _ = _run(api, libs["product"], ann, reads, blocks, api.make_config(max_records=len(reads), fragment_samples=quota))
    ctx = api.Backend(libs["product"]).create(api.make_config(max_records=len(reads), fragment_samples=quota))
    ctx.set_annotation(ann)
    dr = torch.from_numpy(reads.view(np.uint8).reshape(-1, 48)).cuda()
    db = torch.from_numpy(blocks.view(np.int32).reshape(-1, 2)).cuda()
    ctx.submit_device(dr.data_ptr(), len(reads), db.data_ptr(), len(blocks))
    ptr, n = ctx.export_counts()
    # counters | 3 gene tables | 3 exon words | GC bins | samples found | the kept fragment-size samples, two per word
    assert n == api.RSQ_N_COUNTERS + 3 * ann.n_genes + 3 * ann.n_exons + 100 + 1 + (quota + 1) // 2
    ctx.import_counts()           # identity round trip through the all-reduce layout (one partition: every sample is kept)
    got = ctx.finalize()
    assert_results_equal(got, want, "device submit + export/import")
    assert ctx.launch_count() > 0 and ctx.kernel_time_ms() > 0


def _stats(api, lib, ann, reads, blocks, **kw):
    ctx = api.Backend(lib).create(api.make_config(max_records=len(reads)))
    ctx.set_annotation(ann)
    ctx.submit(reads, blocks)
    ctx.sync()
    out = ctx.coverage_statistics(ann, **kw)
    ctx.close()
    return out


@pytest.mark.parametrize("kw", [dict(), dict(mask=100, bias_window=40, bias_gene_length=120), dict(mask=0, bias_offset=20),
                                dict(bias_window=300, mask=200), dict(bias_window=150, smem_window=64), dict(bias_window=9000, bias_gene_length=150)])
def test_device_coverage_statistics_match_oracle(libs, synth_case, kw, monkeypatch):
    """K5 (rsq_covstats.cu) against the sequential restatement of Metrics.cpp:132-337: bias ends exact, fp64 within 1e-9.  Windows up
    to 128: warp-per-gene and CTA kernels; wider: CTA kernel alone; above 8192 (or the lowered test limit): windows in global scratch."""
    api = libs["api"]
    ann, reads, blocks = synth_case
    kw = dict(kw)
    if "smem_window" in kw:
        monkeypatch.setenv("RSQ_COVSTATS_SMEM_WINDOW", str(kw.pop("smem_window")))
    wide = kw.get("bias_window", 100) > 8192
    # deepen the coverage of the synthetic sample so that the >=100 peak / percentile-trimming path is exercised
    r = np.repeat(reads, 6)          # stays coordinate-sorted
    want = _stats(api, libs["oracle"], ann, r, blocks, **kw)
    got = _stats(api, libs["product"], ann, r, blocks, **kw)
    assert np.array_equal(got["five_end"], want["five_end"]) and np.array_equal(got["three_end"], want["three_end"])
    assert wide or (want["five_end"] > 0).sum() > (5 if SCALE == 1 else 0), "bias path not exercised"
    for k in ("gene_mean", "gene_std", "gene_cv", "exon_cv"):
        a, b = got[k], want[k]
        assert np.array_equal(np.isnan(a), np.isnan(b)), k
        m = ~np.isnan(b)
        assert np.allclose(a[m], b[m], rtol=1e-9, atol=1e-12), (k, np.abs(a[m] - b[m]).max())
    assert (want["gene_mean"] > 0).sum() > 50 * min(SCALE * 4, 1)


def test_dense_annotation_forces_the_unstaged_path(libs, tmp_path):
    """More exons / gene segments inside one tile window than the per-warp shared-memory stage holds (64 / 32): every tile
    must fall back to the bucket-index path in L2 and still agree with the oracle bit for bit."""
    from rnaseqc_b200 import synth
    api = libs["api"]
    n_genes, per_gene = 48, 6
    g_start, g_end, ex_s, ex_e, ex_off, strand = [], [], [], [], [0], []
    for g in range(n_genes):                       # 48 nested genes, exons of different genes interleaved every 7 bases
        s0 = 5000 + 7 * g
        for k in range(per_gene):
            ex_s.append(s0 + 420 * k); ex_e.append(s0 + 420 * k + 260)
        ex_off.append(len(ex_s)); g_start.append(s0); g_end.append(ex_e[-1]); strand.append(g % 2)
    sa = synth.SynthAnnotation(["chrD"], [200000], np.zeros(n_genes, np.int32), np.array(g_start, np.int64), np.array(g_end, np.int64),
                               np.array(strand, np.int8), [f"DENSE{g}" for g in range(n_genes)], np.zeros(n_genes, bool),
                               np.array(ex_off, np.int64), np.array(ex_s, np.int64), np.array(ex_e, np.int64))
    gtf = str(tmp_path / "dense.gtf")
    synth.write_gtf(sa, gtf)
    ann = api.Annotation(gtf)
    sr = synth.simulate_reads(sa, scaled(60000), seed=9, genomic_frac=0.3)
    rb, bb = synth.pack(sr, [ann.contig_index("chrD")])
    reads, blocks = synth.to_numpy(rb, bb)
    for kw in (dict(), dict(strand_mode=1)):
        want, _ = _run(api, libs["oracle"], ann, reads, blocks, api.make_config(max_records=len(reads), **kw))
        got, _ = _run(api, libs["product"], ann, reads, blocks, api.make_config(max_records=len(reads), **kw))
        assert_results_equal(got, want, f"dense {kw}")
    assert want.counter("Ambiguous Reads") > 0 and want.gene_reads.sum() > 1000


# ---- --legacy counting rules (src/Expression.cpp:129-304) ------------------------------------------------------------------

@pytest.fixture(scope="module")
def legacy_case(libs, tmp_path_factory):
    """Synthetic sample for the legacy rules: uncollapsed annotation quirks are covered by the KAT; this one brings volume
    (split reads, multi-gene reads, duplicates, chimeras, tid > 127 through the aux bit)."""
    from rnaseqc_b200 import synth
    api = libs["api"]
    d = str(tmp_path_factory.mktemp("legacy"))
    sa = synth.make_annotation(seed=201, n_genes=400, contig_names=["chr1", "chr2", "chr3", "chrM"],
                               contig_lens=[40_000_000, 30_000_000, 20_000_000, 16569])
    synth.write_gtf(sa, os.path.join(d, "a.gtf"))
    synth.write_bed(sa, os.path.join(d, "a.bed"), min_len=400)
    ann = api.Annotation(os.path.join(d, "a.gtf"), os.path.join(d, "a.bed"), legacy=True)
    sr = synth.simulate_reads(sa, scaled(300000), seed=202)
    rb, bb = synth.pack(sr, [ann.contig_index(n) for n in sa.contig_names])
    reads, blocks = synth.to_numpy(rb, bb)
    rng = np.random.default_rng(7)
    reads["aux"] |= ((rng.random(len(reads)) < 0.01) * 0x20).astype(np.uint8)      # RSQ_AUX_TID_GT_127
    return ann, reads, blocks


LEGACY_CONFIGS = {
    "legacy": dict(mapq=4),
    "legacy_bed": dict(mapq=4, fragment_samples=100000),
    "legacy_rf_excl": dict(mapq=4, strand_mode=2, exclude_chimeric=True),
    "legacy_fr_unpaired_q255": dict(strand_mode=1, unpaired=True),
}


@pytest.mark.parametrize("name", sorted(LEGACY_CONFIGS))
def test_cuda_legacy_rules_match_oracle(libs, legacy_case, name):
    api = libs["api"]
    ann, reads, blocks = legacy_case
    cfg = lambda: api.make_config(max_records=len(reads), legacy=True, **LEGACY_CONFIGS[name])
    want, want_bytes = _run(api, libs["oracle"], ann, reads, blocks, cfg())
    got, got_bytes = _run(api, libs["product"], ann, reads, blocks, cfg(), chunks=3)
    assert_results_equal(got, want, name)
    assert got_bytes == want_bytes
    assert want.counter("Split Reads") > 10000 * SCALE and want.gene_reads.sum() > 100000 * SCALE
    assert want.counter("Chimeric Fragments_auto") > 0


@pytest.mark.parametrize("flagset", sorted(KAT_ODD_FLAGSETS))
@pytest.mark.parametrize("batch", [None, "37"])
def test_kat_odd_product_cli_matches_reference(libs, tmp_path, flagset, batch):
    """The two shapes round 2 closed, on the device: a gene that has left the reference's window while one of its exons goes on
    (coverage dropped, fragments counted in a second epoch), and three to five records under one read name in the mate maps
    (stream-order automaton per name: rounds 1 and 2 in parallel, the rest replayed) — also with launches of 37 records, so that
    the records of one name are spread over several launches."""
    d = unpack_case("kat_odd", tmp_path)
    out = os.path.join(d, "out_" + flagset)
    flags = [os.path.join(d, f) if f in ("a.bed", "g.fa") else f for f in KAT_ODD_FLAGSETS[flagset]]
    if batch:
        flags += ["--batch-records", batch]
    run_cli(libs["cli"], os.path.join(d, "a.gtf"), os.path.join(d, "a.sam"), out, *flags)
    identical = set(compare_outputs(os.path.join(d, "expected", flagset), out))
    assert {"a.sam.gene_reads.gct", "a.sam.gene_fragments.gct", "a.sam.fragmentSizes.txt"} <= identical, identical
    if "--fasta" in flags:
        assert "a.sam.gc_content.tsv" in identical


@pytest.mark.parametrize("flagset", sorted(KAT_LEGACY_FLAGSETS))
def test_kat_legacy_product_cli_matches_reference(libs, tmp_path, flagset):
    d = unpack_case("kat_legacy", tmp_path)
    out = os.path.join(d, "out_" + flagset)
    run_cli(libs["cli"], os.path.join(d, "kat.gtf"), os.path.join(d, "kat.sam"), out, *KAT_LEGACY_FLAGSETS[flagset])
    identical = set(compare_outputs(os.path.join(d, "expected", flagset), out))
    assert {"kat.sam.gene_reads.gct", "kat.sam.gene_fragments.gct", "kat.sam.exon_reads.gct"} <= identical, identical


# ---- --fasta: fragment GC-content sampler (src/Expression.cpp:459-476) -----------------------------------------------------

@pytest.fixture(scope="module")
def fasta_case(libs, tmp_path_factory):
    from rnaseqc_b200 import synth
    api = libs["api"]
    d = str(tmp_path_factory.mktemp("fasta"))
    sa = synth.make_annotation(seed=301, n_genes=300, contig_names=["chr1", "chr2", "chr3", "chrM"],
                               contig_lens=[6_000_000, 4_000_000, 3_000_000, 16569])
    synth.write_gtf(sa, os.path.join(d, "a.gtf"))
    synth.write_bed(sa, os.path.join(d, "a.bed"), min_len=400)
    synth.write_fasta(sa, os.path.join(d, "g.fa"), seed=302, contigs=[2, 0, 1])        # chrM has no sequence
    ann = api.Annotation(os.path.join(d, "a.gtf"), os.path.join(d, "a.bed"), fasta=os.path.join(d, "g.fa"))
    sr = synth.simulate_reads(sa, scaled(300000), seed=303)
    rb, bb = synth.pack(sr, [ann.contig_index(n) for n in sa.contig_names])
    reads, blocks = synth.to_numpy(rb, bb)
    return ann, reads, blocks


@pytest.mark.parametrize("name,kw,chunks", [("default", dict(), 1), ("bed_rf", dict(fragment_samples=50000, strand_mode=2), 4),
                                            ("unpaired_small_tables", dict(unpaired=True), 9)])
def test_cuda_fragment_gc_content_matches_oracle(libs, fasta_case, name, kw, chunks):
    api = libs["api"]
    ann, reads, blocks = fasta_case
    max_records = 1000 if "small_tables" in name else len(reads)          # forces the mate map to grow between batches
    want, _ = _run(api, libs["oracle"], ann, reads, blocks, api.make_config(max_records=len(reads), **kw))
    got, _ = _run(api, libs["product"], ann, reads, blocks, api.make_config(max_records=max_records, **kw), chunks=chunks)
    assert_results_equal(got, want, name)
    assert np.array_equal(got.gc_hist, want.gc_hist), (got.gc_hist, want.gc_hist)
    assert want.gc_hist.sum() > 10000 * SCALE and (want.gc_hist > 0).sum() > (30 if SCALE == 1 else 10)


@pytest.mark.parametrize("flagset", sorted(SYNTH_FASTA_FLAGSETS))
def test_synth_fasta_product_cli_matches_reference(libs, tmp_path, flagset):
    d = unpack_case("synth_fasta", tmp_path)
    out = os.path.join(d, "out_" + flagset)
    flags = [os.path.join(d, f) if f in ("a.bed", "g.fa") else f for f in SYNTH_FASTA_FLAGSETS[flagset]]
    run_cli(libs["cli"], os.path.join(d, "a.gtf"), os.path.join(d, "a.sam"), out, *flags)
    identical = set(compare_outputs(os.path.join(d, "expected", flagset), out))
    assert {"a.sam.gene_reads.gct", "a.sam.gene_fragments.gct", "a.sam.gc_content.tsv"} <= identical, identical


def test_packed_wire_format_matches_unpacked_submit(libs, synth_case):
    """rsq_submit_packed: the device expands 32-byte records (implicit whole-read blocks, derived end positions and block
    offsets) into the hot kernel's layout; results equal the rsq_submit ones bit for bit, also over several submits."""
    api = libs["api"]
    ann, reads, blocks = synth_case
    want, want_bytes = _run(api, libs["product"], ann, reads, blocks, api.make_config(max_records=len(reads), fragment_samples=20000))
    ctx = api.Backend(libs["product"]).create(api.make_config(max_records=len(reads), fragment_samples=20000))
    ctx.set_annotation(ann)
    n, total_wire = len(reads), 0
    for k in range(5):
        lo, hi = n * k // 5, n * (k + 1) // 5
        r = reads[lo:hi].copy()
        b0 = int(r["block_off"][0]); b1 = int(r["block_off"][-1]) + int(r["n_blocks"][-1])
        r["block_off"] -= b0
        p = api.pack_records(r, blocks[b0:b1].copy())
        assert p is not None
        ctx.submit_packed(np.ascontiguousarray(p[0]), np.ascontiguousarray(p[1]))
        total_wire += 32 * len(p[0]) + 8 * len(p[1])
    got = ctx.finalize()
    assert_results_equal(got, want, "packed submit")
    assert ctx.algorithmic_bytes() == want_bytes
    assert total_wire < 0.75 * (48 * len(reads) + 8 * len(blocks))
    ctx.close()


def test_compact_wire_format_matches_unpacked_submit(libs, synth_case, tmp_path):
    """rsq_submit_compact: the device expands the 16-byte wire format (tile headers, dictionary, 32-bit blocks, exceptions) into
    the hot kernel's layout; results equal the rsq_submit ones bit for bit — plain sample over several submits, a sample with
    records the format cannot express, and a corrupted batch is reported instead of processed."""
    from test_abi import _compact_cases
    api = libs["api"]
    ann, reads, blocks = synth_case
    cfg = lambda: api.make_config(max_records=len(reads), fragment_samples=20000)
    want, want_bytes = _run(api, libs["product"], ann, reads, blocks, cfg())
    ctx = api.Backend(libs["product"]).create(cfg())
    ctx.set_annotation(ann)
    n, total_wire, keep = len(reads), 0, []
    for k in range(5):
        lo, hi = n * k // 5, n * (k + 1) // 5
        r = reads[lo:hi].copy()
        b0 = int(r["block_off"][0]); b1 = int(r["block_off"][-1]) + int(r["n_blocks"][-1])
        r["block_off"] -= b0
        bl = blocks[b0:b1].copy()
        p = api.pack_compact(r, bl)
        if p is None:                       # only the scaled-down sample of an emulated run is sparse enough for the packer to refuse
            assert SCALE != 1               # (a tile of 32 records then spans more than 65 kb): the caller's fall-back, rsq_submit
            keep.append((r, bl))
            ctx.submit(r, bl)
            total_wire += r.nbytes + bl.nbytes
            continue
        keep.append(p)                      # the buffers belong to the caller until rsq_sync
        ctx.submit_compact(p)
        total_wire += p.nbytes
    got = ctx.finalize()
    assert_results_equal(got, want, "compact submit")
    assert ctx.algorithmic_bytes() == want_bytes
    assert SCALE != 1 or total_wire < 0.45 * (48 * len(reads) + 8 * len(blocks)), total_wire / len(reads)
    ctx.close()
    ann2, reads2, blocks2, odd = _compact_cases(api, tmp_path)
    # (the odd sample is not coordinate-sorted any more — contigs flipped inside tiles —, so the reference's sliding window, which
    # the oracle restates, sees another stream than the stateless device look-ups: the same library's plain submit is the yardstick)
    want2, _ = _run(api, libs["product"], ann2, odd, blocks2, api.make_config(max_records=len(odd)))
    c2 = api.Backend(libs["product"]).create(api.make_config(max_records=len(odd)))
    c2.set_annotation(ann2)
    whole = api.pack_compact(odd, blocks2)
    c2.submit_compact(whole)
    assert_results_equal(c2.finalize(), want2, "compact submit with exceptions vs plain submit")
    off_reads = int(np.frombuffer(whole[48:56].tobytes(), np.uint64)[0])
    n_combos = int(np.frombuffer(whole[:96].tobytes(), np.uint32)[4])
    bad = whole.copy()
    bad[off_reads + 14:off_reads + 16] = np.frombuffer(np.uint16(n_combos + 5).tobytes(), np.uint8)
    c2.reset()
    c2.submit_compact(bad)
    with pytest.raises(api.RsqError):
        c2.finalize()
    c2.close()


@needs_device_pointers()
def test_full_size_properties_200m_records(libs):
    """BASELINE.json configs[2] at full size (200 M synthetic records, GENCODE-scale annotation), where the oracle cannot follow:
    size-independent properties instead — batching invariance (one submit vs three uneven ones: every table and every base of
    the coverage identical), idempotence after reset, the conservation laws of the counters, exon sums vs gene counts, and a
    300 k-record slices from four places of the stream against the oracle."""
    import torch
    from rnaseqc_b200 import synth
    api = libs["api"]
    if torch.cuda.mem_get_info()[0] < 60 << 30:
        pytest.skip("needs ~60 GB of free HBM")
    import tempfile
    n_records = 200_000_000
    d = tempfile.mkdtemp(prefix="rsq_full_")
    sa = synth.make_annotation(seed=20260925, n_genes=56202, contig_names=synth.GRCH38_NAMES, contig_lens=synth.GRCH38)
    synth.write_gtf(sa, os.path.join(d, "a.gtf"))
    ann = api.Annotation(os.path.join(d, "a.gtf"))
    parts, done, k = [], 0, 0
    while done < n_records:
        sr = synth.simulate_reads(sa, min(20_000_000, n_records - done), seed=20260925 + k, device="cuda:0", sort=False, name_base=done)
        parts.append(sr); done += sr.n; k += 1
    sr = synth.sort_reads(synth.concat(parts))
    del parts
    rb, bb = synth.pack(sr, [ann.contig_index(n) for n in sa.contig_names])
    del sr
    torch.cuda.empty_cache()
    n, nb = rb.shape[0], bb.shape[0]

    def run(bounds):
        ctx = api.Backend(libs["product"]).create(api.make_config(max_records=n))
        ctx.set_annotation(ann)
        out = []
        for rep in range(2 if len(bounds) == 2 else 1):          # the single-submit context also checks idempotence after reset
            ctx.reset()
            for a, b in zip(bounds[:-1], bounds[1:]):
                ctx.submit_device(rb.data_ptr() + 48 * a, b - a, bb.data_ptr(), nb)
            out.append(ctx.finalize())
        ctx.close()
        return out

    whole, again = run([0, n])
    (parts3,) = run([0, 61_234_567, 61_234_568, 150_000_001, n])
    assert_results_equal(again, whole, "second pass after reset")
    assert_results_equal(parts3, whole, "four uneven submits vs one")
    assert np.array_equal(parts3.coverage, whole.coverage)

    c = whole.counter
    assert c("Total Alignments") == n
    assert c("Mapped Reads") == c("Mapped Duplicate Reads") + c("Mapped Unique Reads")
    assert c("Reads used for Intron/Exon counts") == c("High Quality Reads") + c("Low Quality Reads")
    assert c("Exonic Reads") + c("Intronic Reads") + c("Intergenic Reads") + c("Ambiguous Reads") == c("Reads used for Intron/Exon counts")
    assert c("Intragenic Reads") == c("Exonic Reads") + c("Intronic Reads")
    assert c("HQ Exonic Reads") + c("HQ Intronic Reads") + c("HQ Intergenic Reads") + c("HQ Ambiguous Reads") == c("High Quality Reads")
    assert c("End 1 Mapped Reads") + c("End 2 Mapped Reads") <= c("Mapped Reads")
    assert int(whole.gene_fragments.sum()) == c("Fragment Set Entries")
    assert (whole.gene_unique <= whole.gene_reads).all() and (whole.gene_fragments <= whole.gene_reads).all()
    assert int(whole.gene_reads.sum()) >= c("HQ Exonic Reads") > 100_000_000
    # every (read, gene) assignment spreads exactly 1 over the exons of the gene (each floor loses < 2^-64)
    assert abs(float(whole.exon_counts().sum()) - float(whole.gene_reads.sum())) < 1e-3
    # slices from four places of the stream (its head, one and two thirds in, its tail) against the oracle
    for a0 in (0, n // 3, 2 * (n // 3), n - 300_000):
        pr = np.frombuffer(rb[a0:a0 + 300_000].cpu().numpy().reshape(-1).data, dtype=api.READ_DTYPE).copy()
        b0, b1 = int(pr["block_off"][0]), int(pr["block_off"][-1]) + int(pr["n_blocks"][-1])
        pr["block_off"] -= np.uint32(b0)
        pb = np.frombuffer(bb[b0:b1].cpu().numpy().reshape(-1).data, dtype=api.BLOCK_DTYPE).copy()
        want, _ = _run(api, libs["oracle"], ann, pr, pb, api.make_config(max_records=len(pr)))
        got, _ = _run(api, libs["product"], ann, pr, pb, api.make_config(max_records=len(pr)))
        assert_results_equal(got, want, f"300 k-record slice at record {a0} vs oracle")


@pytest.mark.parametrize("mode", ["sequential", "threads", "threads_no_speculation", "threads_no_digest"])
def test_product_cli_reads_a_multi_job_bam_like_the_reference(libs, tmp_path, mode, monkeypatch):
    """BAM -> report files through the PRODUCT CLI (multi-threaded ingest with and without speculated framing / job digests, wire-format
    packers, CUDA path) against the reference binary's outputs for the same BAM (tests/golden/bam_small: several ~4 MB ingest jobs)"""
    from test_oracle_golden import run_bam_small
    identical = run_bam_small(libs["cli"], tmp_path, mode, monkeypatch)
    assert {"a.bam.gene_reads.gct", "a.bam.gene_fragments.gct", "a.bam.fragmentSizes.txt", "a.bam.exon_reads.gct"} <= identical, identical


def test_product_cli_reads_cram_like_the_reference(libs, tmp_path):
    """CRAM input through the product CLI (host-side reader, tests/test_cram.py has its CPU twins) against the reference's outputs."""
    d = unpack_case("cram_small", tmp_path)
    out = os.path.join(d, "out")
    run_cli(libs["cli"], os.path.join(d, "c.gtf"), os.path.join(d, "c.cram"), out, "--fasta", os.path.join(d, "c.fa"),
            "--coverage", "--bed", os.path.join(d, "c.bed"))
    compare_outputs(out, os.path.join(d, "expected"))
