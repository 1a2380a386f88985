"""CPU tests of the hot kernel's per-record logic and of the elementary-interval index, WITHOUT a GPU.

`oracle/_build/libhostcore.so` (tests/hostcore/hostcore.cpp, test infrastructure) compiles the very sources the CUDA kernel
is made of — rnaseqc_b200/csrc/rsq_hot_core.cuh (filter cascade, interval look-up, exonAlignmentMetrics) and
rnaseqc_b200/csrc/rsq_eit.h (index builder) — for the host and drives them tile by tile as a warp does (32 records, a
128-entry staged slice, the 1 kb table for blocks outside it).  Its tables must equal the CPU oracle's bit for bit:
counters, gene tables, exact Q64.64 exon sums, every base of the coverage array, the algorithmic byte count."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOSTCORE = os.path.join(ROOT, "oracle", "_build", "libhostcore.so")


VARIANTS = {"product": HOSTCORE, "stage_desc": os.path.join(ROOT, "oracle", "_build", "libhostcore_sd.so")}


def _hostcore(path=HOSTCORE):
    lib = C.CDLL(path)
    lib.hostcore_run.restype = C.c_int
    return lib


def run_hostcore(api, ann, reads, blocks, cfg, variant="product"):
    lib = _hostcore(VARIANTS[variant])
    G, E, L = ann.n_genes, ann.n_exons, ann.cov_len
    out = {k: np.zeros(n, np.uint64) for k, n in (("counters", 64), ("gene_reads", G), ("gene_unique", G), ("gene_fragments", G),
                                                   ("exon_whole", E), ("exon_fhi", E), ("exon_flo", E), ("bytes", 1), ("stats", 8))}
    cov = np.zeros(L + 1, np.uint32)
    reads = np.ascontiguousarray(reads)
    blocks = np.ascontiguousarray(blocks)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.hostcore_run(C.byref(ann.view), C.byref(cfg), p(reads), C.c_size_t(len(reads)), p(blocks), C.c_size_t(len(blocks)),
                          p(out["counters"]), p(out["gene_reads"]), p(out["gene_unique"]), p(out["gene_fragments"]), p(out["exon_whole"]),
                          p(out["exon_fhi"]), p(out["exon_flo"]), p(cov), p(out["bytes"]), p(out["stats"]))
    assert rc == 0
    # combine_exon_kernel's carry-free fold and the coverage scan, in numpy
    hi, lo = out["exon_fhi"], out["exon_flo"]
    mid = hi + (lo >> np.uint64(32))
    res = api.Results(G, E, L)
    res.counters = out["counters"]
    res.gene_reads, res.gene_unique, res.gene_fragments = out["gene_reads"], out["gene_unique"], out["gene_fragments"]
    res.exon_frac = (mid << np.uint64(32)) | (lo & np.uint64(0xffffffff))
    res.exon_whole = out["exon_whole"] + (mid >> np.uint64(32))
    res.coverage = np.cumsum(cov, dtype=np.uint32)
    return res, int(out["bytes"][0]), out["stats"]


def run_oracle(api, lib, ann, reads, blocks, cfg):
    ctx = api.Backend(lib).create(cfg)
    ctx.set_annotation(ann)
    ctx.submit(reads, blocks)
    res = ctx.finalize()
    nbytes = ctx.algorithmic_bytes()
    ctx.close()
    return res, nbytes


def assert_same(api, got, want, what):
    skip = {api.COUNTER_INDEX["Fragment Candidates"]}
    bad = [(api.COUNTER_NAMES[i], int(got.counters[i]), int(want.counters[i])) for i in range(len(api.COUNTER_NAMES))
           if got.counters[i] != want.counters[i] and i not in skip]
    assert not bad, f"{what}: counters differ: {bad}"
    for name in ("gene_reads", "gene_unique", "gene_fragments", "exon_whole", "exon_frac"):
        x, y = getattr(got, name), getattr(want, name)
        diff = np.nonzero(x != y)[0]
        assert diff.size == 0, f"{what}: {name} differs at {diff[:10]}: {x[diff[:10]]} vs {y[diff[:10]]}"
    n = min(len(got.coverage), len(want.coverage)) - 1
    diff = np.nonzero(got.coverage[:n] != want.coverage[:n])[0]
    assert diff.size == 0, f"{what}: coverage differs at {diff[:10]}"


@pytest.fixture(scope="module")
def synth_case(libs, tmp_path_factory):
    from rnaseqc_b200 import synth
    api = libs["api"]
    d = str(tmp_path_factory.mktemp("hostcore"))
    sa = synth.make_annotation(seed=101, n_genes=400, contig_names=["chr1", "chr2", "chr3", "chrM"],
                               contig_lens=[40_000_000, 30_000_000, 20_000_000, 16569])
    synth.write_gtf(sa, os.path.join(d, "a.gtf"))
    ann = api.Annotation(os.path.join(d, "a.gtf"))
    sr = synth.simulate_reads(sa, 200000, seed=102)
    rb, bb = synth.pack(sr, [ann.contig_index(n) for n in sa.contig_names])
    reads, blocks = synth.to_numpy(rb, bb)
    return ann, reads, blocks


CONFIGS = {
    "default": dict(),
    "stranded_fr": dict(strand_mode=1),
    "stranded_rf": dict(strand_mode=2),
    "unpaired": dict(unpaired=True),
    "exclude_chimeric": dict(exclude_chimeric=True),
    "loose": dict(mapq=3, mismatches=2, chimeric_distance=1000),
}


@pytest.mark.parametrize("variant", sorted(VARIANTS))
@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_record_logic_matches_oracle_on_synthetic_records(libs, synth_case, name, variant):
    api = libs["api"]
    ann, reads, blocks = synth_case
    cfg = lambda: api.make_config(max_records=len(reads), **CONFIGS[name])
    want, want_bytes = run_oracle(api, libs["oracle"], ann, reads, blocks, cfg())
    got, got_bytes, stats = run_hostcore(api, ann, reads, blocks, cfg(), variant)
    assert_same(api, got, want, name + " " + variant)
    assert got_bytes == want_bytes
    assert want.counter("Exonic Reads") > 50000 and want.gene_reads.sum() > 50000
    assert stats[0] > 0, "the staged-slice look-up must be exercised"


def test_record_logic_filter_tags_aux_bits_and_unsorted_input(libs, synth_case):
    api = libs["api"]
    ann, reads, blocks = synth_case
    r = reads.copy()
    rng = np.random.default_rng(5)
    r["filter_mask"] = (rng.random(len(r)) < 0.05) * rng.integers(1, 4, len(r))
    r["aux"] |= ((rng.random(len(r)) < 0.01) * 0x08).astype(np.uint8)     # tid out of range
    r["aux"] |= ((rng.random(len(r)) < 0.02) * 0x10).astype(np.uint8)     # contig already dropped
    r["contig"][rng.random(len(r)) < 0.01] = -1
    r = r[rng.permutation(len(r))]                                        # tiles mix contigs and far-apart coordinates
    cfg = lambda: api.make_config(n_filter_tags=2, max_records=len(r))
    got, _, stats = run_hostcore(api, ann, r, blocks, cfg())
    assert stats[1] > 0, "the 1 kb table look-up must be exercised"
    # the oracle's sliding window drops features behind the stream position: only order-free tables compare on shuffled input
    want_sorted, _ = run_oracle(api, libs["oracle"], ann, r[np.argsort(r, order=["contig", "pos"], kind="stable")], blocks, cfg())
    got_sorted, got_bytes, _ = run_hostcore(api, ann, r[np.argsort(r, order=["contig", "pos"], kind="stable")], blocks, cfg())
    assert_same(api, got_sorted, want_sorted, "filter tags, sorted")
    # the data-parallel path is order-free: shuffled input gives the sorted answer
    assert_same(api, got, got_sorted, "shuffled vs sorted")


def test_dense_overlapping_annotation_and_edge_blocks(libs, tmp_path):
    """Hand-made annotation with nested / abutting / identical exons of several genes on both strands, a ribosomal gene and
    an exon ending at a bucket boundary; reads whose blocks start or end exactly on every cut, zero-length and overhanging
    blocks, blocks beyond the last feature and before the first."""
    api = libs["api"]
    gtf = tmp_path / "dense.gtf"
    rows = []

    def gene(gid, start, end, strand, exons, ttype="protein_coding"):
        rows.append(f'chr1\tt\tgene\t{start}\t{end}\t.\t{strand}\t.\tgene_id "{gid}"; transcript_id "{gid}"; gene_type "{ttype}"; gene_name "{gid}"; transcript_type "{ttype}";')
        for k, (s, e) in enumerate(exons):
            rows.append(f'chr1\tt\texon\t{s}\t{e}\t.\t{strand}\t.\tgene_id "{gid}"; transcript_id "{gid}"; gene_type "{ttype}"; gene_name "{gid}"; transcript_type "{ttype}"; exon_id "{gid}_{k}";')

    gene("A", 1000, 5000, "+", [(1000, 1200), (2000, 2100), (4000, 5000)])
    gene("B", 1100, 2050, "-", [(1100, 1200), (2000, 2050)])
    gene("C", 1000, 1200, "+", [(1000, 1200)])                   # identical to A's first exon
    gene("D", 2047, 2049, "+", [(2047, 2049)])                   # tiny exon across the 2048 bucket boundary
    gene("R", 3000, 3500, "+", [(3000, 3500)], ttype="rRNA")
    gene("E", 4500, 4800, "-", [(4500, 4600), (4601, 4800)])      # abutting exons
    gene("F", 900000, 900500, ".", [(900000, 900500)])
    gtf.write_text("\n".join(rows) + "\n")
    ann = api.Annotation(str(gtf))
    cuts = sorted({c for g in [(1000, 1200), (2000, 2100), (4000, 5000), (1100, 1200), (2000, 2050), (2047, 2049), (3000, 3500),
                               (4500, 4600), (4601, 4800), (900000, 900500), (1000, 5000), (1100, 2050)] for c in (g[0], g[1], g[1] + 1)})
    recs, blks = [], []
    name = 1

    def add(block_list, flag=99, nm=0):
        nonlocal name
        off = len(blks)
        blks.extend(block_list)
        pos = block_list[0][0] - 1 if block_list else 10
        end = (block_list[-1][0] - 1 + block_list[-1][1]) if block_list else pos + 1
        recs.append((pos, pos + 100, end, 200, 100, nm, flag, 255, 0x01, 0, 0, 0, name, off, len(block_list), 0))
        name += 1

    for c in cuts:
        for d in (-2, -1, 0, 1):
            for ln in (0, 1, 2, 50, 101, 1000):
                if c + d >= 1:
                    add([(c + d, ln)])
                    add([(c + d, ln)], flag=147)
            if c + d >= 1:
                add([(c + d, 30), (c + d + 900, 30)])                                  # second block usually beyond the first's window
                add([(max(1, c + d - 60), 40), (c + d, 40), (c + d + 2000, 20)], nm=7)  # three blocks, low quality
    add([(1, 10)]); add([(5000000, 100)]); add([(899990, 20)]); add([(900490, 20)])
    reads = np.array(recs, dtype=api.READ_DTYPE)
    blocks = np.array(blks, dtype=api.BLOCK_DTYPE)
    order = np.argsort(reads["pos"], kind="stable")
    reads = reads[order]
    for mode in (0, 1, 2):
        cfg = lambda: api.make_config(strand_mode=mode, max_records=len(reads))
        want, want_bytes = run_oracle(api, libs["oracle"], ann, reads, blocks, cfg())
        got, got_bytes, _ = run_hostcore(api, ann, reads, blocks, cfg())
        assert_same(api, got, want, f"dense annotation, strand mode {mode}")
        assert got_bytes == want_bytes
        assert want.counter("Ambiguous Reads") > 0 and (mode == 2 or want.counter("rRNA Reads") > 0) and want.counter("Exonic Reads") > 0


def _fuzz_case(api, tmp_path, seed):
    """Random small annotation dense enough that most reads see several overlapping genes on both strands (nested, abutting and
    identical exons, rRNA genes, a globin name, exons on 1 kb bucket boundaries), and random reads with 0-6 blocks whose ends fall
    on / next to the cuts."""
    rng = np.random.default_rng(seed)
    contigs = ["chr1", "chr2", "chrM"][: int(rng.integers(1, 4))]
    span = int(rng.choice([6000, 40000, 300000]))
    rows, cuts = [], {c: [] for c in contigs}
    n_genes = int(rng.integers(3, 60))
    for g in range(n_genes):
        chrom = contigs[int(rng.integers(0, len(contigs)))]
        gs = int(rng.integers(1, span))
        if rng.random() < 0.2:
            gs = max(1, (gs >> 10 << 10) + int(rng.integers(-2, 3)))                 # near a 1 kb bucket boundary
        ge = gs + int(rng.integers(1, max(2, span // 4)))
        strand = "+-."[int(rng.choice([0, 0, 1, 1, 2]))]
        ttype = "rRNA" if rng.random() < 0.1 else ("Mt_rRNA" if rng.random() < 0.05 else "protein_coding")
        gname = "HBB" if rng.random() < 0.05 else f"N{g}"
        common = f'gene_id "G{g}"; transcript_id "G{g}"; gene_type "{ttype}"; gene_name "{gname}"; transcript_type "{ttype}";'
        rows.append(f"{chrom}\tf\tgene\t{gs}\t{ge}\t.\t{strand}\t.\t{common}")
        k, p = 0, gs
        while p <= ge and k < 12:
            es = p
            ee = min(ge, es + int(rng.integers(0, max(1, (ge - gs) // 3 + 1))))
            if rng.random() < 0.12:                                                   # an exon that reaches past its gene line: reads on the
                ee = es + int(rng.integers(0, max(2, span // 6)))                     # far side find the gene evicted (rsq_eit.h, gene_eviction_bounds)
            rows.append(f'{chrom}\tf\texon\t{es}\t{ee}\t.\t{strand}\t.\t{common} exon_id "G{g}_{k}";')
            cuts[chrom] += [es, ee, ee + 1]
            k += 1
            p = ee + 1 + int(rng.choice([0, 0, 1, 5, 100, 1000]))                     # abutting exons are common
        cuts[chrom] += [gs, ge, ge + 1]
    gtf = tmp_path / f"fuzz{seed}.gtf"
    gtf.write_text("\n".join(rows) + "\n")
    ann = api.Annotation(str(gtf))
    recs, blks = [], []
    n_reads = 1500
    for i in range(n_reads):
        ci = int(rng.integers(0, len(contigs)))
        cs = cuts[contigs[ci]]
        nb = int(rng.choice([0, 1, 1, 1, 1, 2, 2, 3, 6]))
        if cs and rng.random() < 0.8:
            p = max(1, int(cs[int(rng.integers(0, len(cs)))]) + int(rng.integers(-3, 4)) - int(rng.choice([0, 0, 10, 50, 100])))
        else:
            p = int(rng.integers(1, span + 3000))
        off, first, bl = len(blks), p, []
        for b in range(nb):
            ln = int(rng.choice([0, 1, 2, 10, 25, 50, 76, 101, 400]))
            bl.append((p, ln))
            p += ln + int(rng.choice([0, 1, 30, 500, 2000, 20000]))
        blks.extend(bl)
        pos = first - 1
        end = (bl[-1][0] - 1 + bl[-1][1]) if bl else pos + 1
        flag = int(rng.choice([99, 147, 83, 163, 0, 16, 1024 + 99, 256 + 99, 512 + 147, 4, 2048 + 83, 65, 129, 73]))
        mapq = int(rng.choice([255, 255, 255, 60, 3, 0]))
        nm = int(rng.choice([0, 0, 1, 6, 7, 12]))
        aux = int(rng.choice([1, 1, 1, 0, 3, 5, 9, 0x11, 0x21]))
        contig = ann.contig_index(contigs[ci]) if rng.random() > 0.03 else -1
        mpos = pos + int(rng.choice([150, -150, 0, 3_000_000]))
        name = int(rng.integers(1, n_reads // 2 + 2)) * 0x9E3779B97F4A7C15 % (1 << 64)   # mates and repeats share names
        recs.append((pos, mpos, end, 200, 101, nm, flag, mapq, aux, contig, int(rng.choice([0, 0, 0, 1, 2])), 0, name, off, nb, 0))
    reads = np.array(recs, dtype=api.READ_DTYPE)
    blocks = np.array(blks, dtype=api.BLOCK_DTYPE) if blks else np.zeros(0, api.BLOCK_DTYPE)
    order = np.lexsort((reads["pos"], reads["contig"].astype(np.int64) % 70000))       # contig -1 last, then stream order
    return ann, reads[order], blocks, rng


@pytest.mark.parametrize("variant", sorted(VARIANTS))
@pytest.mark.parametrize("seed", range(24))
def test_fuzz_random_dense_annotations(libs, tmp_path, seed, variant):
    api = libs["api"]
    ann, reads, blocks, rng = _fuzz_case(api, tmp_path, 7000 + seed)
    kw = dict(strand_mode=int(rng.integers(0, 3)), unpaired=bool(rng.random() < 0.3), exclude_chimeric=bool(rng.random() < 0.3),
              n_filter_tags=int(rng.choice([0, 0, 1, 2])), mapq=int(rng.choice([255, 255, 60, 3])), mismatches=int(rng.choice([6, 6, 0, 12])))
    cfg = lambda: api.make_config(max_records=len(reads), **kw)
    want, want_bytes = run_oracle(api, libs["oracle"], ann, reads, blocks, cfg())
    got, got_bytes, _ = run_hostcore(api, ann, reads, blocks, cfg(), variant)
    assert_same(api, got, want, f"fuzz seed {seed} {kw} {variant}")
    assert got_bytes == want_bytes


# ---- the --legacy device rules (rsq_legacy.cuh) on the host -------------------------------------------------------------------
HOSTLEGACY = os.path.join(ROOT, "oracle", "_build", "libhostlegacy.so")


def run_hostlegacy(api, ann, reads, blocks, cfg):
    lib = C.CDLL(HOSTLEGACY)
    lib.hostlegacy_run.restype = C.c_int
    G, E, L = ann.n_genes, ann.n_exons, ann.cov_len
    out = {k: np.zeros(n, np.uint64) for k, n in (("counters", 64), ("gene_reads", G), ("gene_unique", G), ("gene_fragments", G),
                                                   ("exon_whole", E), ("exon_fhi", E), ("exon_flo", E), ("bytes", 1))}
    cov = np.zeros(L + 1, np.uint32)
    reads = np.ascontiguousarray(reads)
    blocks = np.ascontiguousarray(blocks)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.hostlegacy_run(C.byref(ann.view), C.byref(cfg), p(reads), C.c_size_t(len(reads)), p(blocks), C.c_size_t(len(blocks)),
                            p(out["counters"]), p(out["gene_reads"]), p(out["gene_unique"]), p(out["gene_fragments"]), p(out["exon_whole"]),
                            p(out["exon_fhi"]), p(out["exon_flo"]), p(cov), p(out["bytes"]))
    assert rc == 0
    hi, lo = out["exon_fhi"], out["exon_flo"]
    mid = hi + (lo >> np.uint64(32))
    res = api.Results(G, E, L)
    res.counters = out["counters"]
    res.gene_reads, res.gene_unique, res.gene_fragments = out["gene_reads"], out["gene_unique"], out["gene_fragments"]
    res.exon_frac = (mid << np.uint64(32)) | (lo & np.uint64(0xffffffff))
    res.exon_whole = out["exon_whole"] + (mid >> np.uint64(32))
    res.coverage = np.cumsum(cov, dtype=np.uint32)
    return res, int(out["bytes"][0])


LEGACY_CONFIGS = {
    "legacy": dict(mapq=4),
    "legacy_rf_excl": dict(mapq=4, strand_mode=2, exclude_chimeric=True),
    "legacy_fr_unpaired_q255": dict(strand_mode=1, unpaired=True),
}


@pytest.mark.parametrize("name", sorted(LEGACY_CONFIGS))
def test_legacy_device_rules_match_oracle_on_synthetic_records(libs, tmp_path_factory, name):
    from rnaseqc_b200 import synth
    api = libs["api"]
    d = str(tmp_path_factory.mktemp("hostlegacy"))
    sa = synth.make_annotation(seed=201, n_genes=400, contig_names=["chr1", "chr2", "chr3", "chrM"], contig_lens=[40_000_000, 30_000_000, 20_000_000, 16569])
    synth.write_gtf(sa, os.path.join(d, "a.gtf"))
    ann = api.Annotation(os.path.join(d, "a.gtf"), legacy=True)
    sr = synth.simulate_reads(sa, 120000, seed=202)
    rb, bb = synth.pack(sr, [ann.contig_index(n) for n in sa.contig_names])
    reads, blocks = synth.to_numpy(rb, bb)
    rng = np.random.default_rng(7)
    reads["aux"] |= ((rng.random(len(reads)) < 0.01) * 0x20).astype(np.uint8)      # RSQ_AUX_TID_GT_127
    cfg = lambda: api.make_config(max_records=len(reads), legacy=True, **LEGACY_CONFIGS[name])
    want, want_bytes = run_oracle(api, libs["oracle"], ann, reads, blocks, cfg())
    got, got_bytes = run_hostlegacy(api, ann, reads, blocks, cfg())
    assert_same(api, got, want, name)
    assert got_bytes == want_bytes
    assert want.counter("Split Reads") > 3000 and want.gene_reads.sum() > 30000


@pytest.mark.parametrize("seed", [4100, 4101, 4102, 5021, 5038, 5048])
def test_legacy_device_rules_match_oracle_on_random_sam_records(libs, tmp_path, seed):
    """The odd records of tests/test_sam_fuzz.py (pinned to the reference there) through the device's --legacy code on the host."""
    import random
    from test_sam_fuzz import _gtf, _sam
    api = libs["api"]
    rng = random.Random(seed)
    d = str(tmp_path)
    _gtf(os.path.join(d, "f.gtf"), rng)
    _sam(os.path.join(d, "f.sam"), rng, 2500, allow_dup=True)
    ann = api.Annotation(os.path.join(d, "f.gtf"), legacy=True)
    for kw in (dict(mapq=4), dict(strand_mode=1, unpaired=True, exclude_chimeric=True), dict(mapq=0, mismatches=30, strand_mode=2)):
        reads, blocks, _ = ann.decode(os.path.join(d, "f.sam"), exclude_chimeric=kw.get("exclude_chimeric", False))
        cfg = lambda: api.make_config(max_records=len(reads), legacy=True, **kw)
        want, want_bytes = run_oracle(api, libs["oracle"], ann, reads, blocks, cfg())
        got, got_bytes = run_hostlegacy(api, ann, reads, blocks, cfg())
        assert_same(api, got, want, f"seed {seed} {kw}")
        assert got_bytes == want_bytes


@pytest.mark.parametrize("seed", range(12))
def test_legacy_device_rules_fuzz_on_dense_annotations(libs, tmp_path, seed):
    api = libs["api"]
    _, reads, blocks, rng = _fuzz_case(api, tmp_path, 7100 + seed)
    ann = api.Annotation(str(tmp_path / f"fuzz{7100 + seed}.gtf"), legacy=True)
    # the ABI asks for MIXED name hashes (include/rsq_b200.h); the fuzz generator's are multiples of one constant, and the legacy key
    # (name ^ gene * constant) collides structurally on those — splitmix64 finaliser, as hash_qname ends with
    h = reads["name_hash"].astype(np.uint64)
    h = (h ^ (h >> np.uint64(30))) * np.uint64(0xbf58476d1ce4e5b9)
    h = (h ^ (h >> np.uint64(27))) * np.uint64(0x94d049bb133111eb)
    reads["name_hash"] = h ^ (h >> np.uint64(31))
    kw = dict(strand_mode=int(rng.integers(0, 3)), unpaired=bool(rng.random() < 0.3), exclude_chimeric=bool(rng.random() < 0.3),
              mapq=int(rng.choice([255, 4, 60, 3])), mismatches=int(rng.choice([6, 6, 0, 12])))
    cfg = lambda: api.make_config(max_records=len(reads), legacy=True, **kw)
    want, want_bytes = run_oracle(api, libs["oracle"], ann, reads, blocks, cfg())
    got, got_bytes = run_hostlegacy(api, ann, reads, blocks, cfg())
    assert_same(api, got, want, f"legacy fuzz seed {seed} {kw}")
    assert got_bytes == want_bytes
