"""CPU tests: both libraries load and export every symbol include/rsq_b200.h declares; the product refuses to run
without a GPU instead of falling back."""
import ctypes
import os
import re

import pytest

from conftest import has_gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "rsq_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"static inline .*?\n}\n", "", text, flags=re.S)      # header-only helpers (rsq_pack_records) are not exports
    return sorted(set(re.findall(r"\b(rsq_[a-z_]+)\s*\(", text)))


def test_header_symbols_known(libs):
    assert declared_symbols() == sorted(libs["api"].ABI_SYMBOLS)


@pytest.mark.parametrize("which", ["product", "oracle"])
def test_library_exports_every_declared_symbol(libs, which):
    lib = ctypes.CDLL(libs[which])
    for sym in declared_symbols():
        assert hasattr(lib, sym), f"{libs[which]} does not export {sym}"
    lib.rsq_backend.restype = ctypes.c_char_p
    lib.rsq_abi_version.restype = ctypes.c_uint32
    assert lib.rsq_abi_version() == 3
    assert lib.rsq_backend().decode() == {"product": "cuda-sm_100a", "oracle": "oracle-cpu"}[which]


def test_struct_sizes(libs):
    api = libs["api"]
    assert api.READ_DTYPE.itemsize == 48 and api.BLOCK_DTYPE.itemsize == 8
    assert ctypes.sizeof(api.RsqConfig) == 80


@pytest.mark.skipif(has_gpu(), reason="only meaningful on a box without a GPU")
def test_product_fails_loudly_without_gpu(libs):
    api = libs["api"]
    with pytest.raises(api.RsqError) as e:
        api.default_backend().create(api.make_config())
    assert e.value.code == api.RSQ_ERR_NO_DEVICE


def test_product_does_not_link_the_oracle(libs):
    import subprocess
    for f in (libs["product"], libs["cli"], libs["host"]):
        out = subprocess.run(["ldd", f], stdout=subprocess.PIPE, text=True).stdout
        assert "oracle" not in out, out


def test_packed_wire_format_roundtrip_through_the_oracle(libs, tmp_path):
    """rsq_pack_records + rsq_submit_packed (32-byte records, implicit whole-read blocks) give the same results as rsq_submit;
    batches that break the format's rules are refused."""
    import numpy as np
    from rnaseqc_b200 import synth
    api = libs["api"]
    assert api.PACKED_DTYPE.itemsize == 32
    sa = synth.make_annotation(seed=5, n_genes=60, contig_names=["chr1", "chr2"], contig_lens=[4_000_000, 2_000_000])
    gtf = str(tmp_path / "a.gtf")
    synth.write_gtf(sa, gtf)
    ann = api.Annotation(gtf)
    sr = synth.simulate_reads(sa, 30000, seed=6)
    rb, bb = synth.pack(sr, [ann.contig_index(n) for n in sa.contig_names])
    reads, blocks = synth.to_numpy(rb, bb)
    packed = api.pack_records(reads, blocks)
    assert packed is not None
    pr, pb = packed
    implicit = (pr["aux"] & api.RSQ_AUX_IMPLICIT_BLOCK) != 0
    assert implicit.sum() > len(reads) // 3 and len(pb) == int(reads["n_blocks"][~implicit].sum())

    def run(submit):
        ctx = api.Backend(libs["oracle"]).create(api.make_config())
        ctx.set_annotation(ann)
        submit(ctx)
        return ctx.finalize()

    want = run(lambda c: c.submit(reads, blocks))
    half = len(reads) // 2

    def packed_two_calls(c):
        for lo, hi in ((0, half), (half, len(reads))):
            r = reads[lo:hi].copy()
            b0 = int(r["block_off"][0]); b1 = int(r["block_off"][-1]) + int(r["n_blocks"][-1])
            r["block_off"] -= b0
            p = api.pack_records(r, blocks[b0:b1].copy())
            c.submit_packed(np.ascontiguousarray(p[0]), np.ascontiguousarray(p[1]))
    got = run(packed_two_calls)
    from util import assert_results_equal
    assert_results_equal(got, want, "packed vs unpacked")
    bad = reads.copy()
    bad["endpos"][7] += 3                      # end position that is not the end of the last block
    assert api.pack_records(bad, blocks) is None
    bad = reads.copy()
    bad["nm"][11] = 70000
    assert api.pack_records(bad, blocks) is None


def _compact_cases(api, tmp_path):
    """(annotation, reads, blocks) of a synthetic sample, plus a copy with records the 16-byte format cannot express"""
    import numpy as np
    from rnaseqc_b200 import synth
    sa = synth.make_annotation(seed=5, n_genes=60, contig_names=["chr1", "chr2"], contig_lens=[4_000_000, 2_000_000])
    gtf = str(tmp_path / "a.gtf")
    synth.write_gtf(sa, gtf)
    ann = api.Annotation(gtf)
    sr = synth.simulate_reads(sa, 30000, seed=6)
    rb, bb = synth.pack(sr, [ann.contig_index(n) for n in sa.contig_names])
    reads, blocks = synth.to_numpy(rb, bb)
    odd = reads.copy()
    rng = np.random.default_rng(1)
    k = rng.choice(len(odd), 600, replace=False)
    odd["mpos"][k[:100]] += 1_000_000                  # mate far away
    odd["isize"][k[100:200]] = 400_000                 # insert size beyond 16 bits
    odd["nm"][k[200:250]] = 70_000                     # NM beyond 16 bits
    odd["contig"][k[250:300]] = 1 - odd["contig"][k[250:300]]   # another contig inside a tile
    odd["l_qseq"][k[300:350]] = 100_000
    odd["flag"][k[350:600]] ^= rng.integers(0, 1 << 12, 250).astype(np.uint16)   # many distinct combinations
    return ann, reads, blocks, odd


def test_compact_wire_format_roundtrip_through_the_oracle(libs, tmp_path):
    """rsq_pack_compact + rsq_submit_compact (16-byte records, tile headers, dictionary, 32-bit blocks, exception list) give the
    same results as rsq_submit — also for records the format cannot express (they travel as exceptions) and over several
    submits; batches with too many exceptions and corrupted batches are refused."""
    import numpy as np
    from util import assert_results_equal
    api = libs["api"]
    ann, reads, blocks, odd = _compact_cases(api, tmp_path)

    def run(submit):
        ctx = api.Backend(libs["oracle"]).create(api.make_config())
        ctx.set_annotation(ann)
        submit(ctx)
        return ctx.finalize()

    for name, rr in (("plain", reads), ("odd", odd)):
        want = run(lambda c: c.submit(rr, blocks))
        whole = api.pack_compact(rr, blocks)
        assert whole is not None
        hdr = np.frombuffer(whole[:96].tobytes(), np.uint32)
        n_exc = int(hdr[6])
        if name == "plain":
            assert n_exc < len(rr) // 50 and whole.nbytes < 24 * len(rr), (n_exc, whole.nbytes / len(rr))
        else:
            assert 250 <= n_exc <= 700
        assert_results_equal(run(lambda c: c.submit_compact(whole)), want, f"compact vs unpacked ({name})")

        def three_calls(c):
            n = len(rr)
            for lo, hi in ((0, n // 3), (n // 3, n // 3 + 33), (n // 3 + 33, n)):
                r = rr[lo:hi].copy()
                b0 = int(r["block_off"][0]); b1 = int(r["block_off"][-1]) + int(r["n_blocks"][-1])
                r["block_off"] -= b0
                c.submit_compact(api.pack_compact(r, blocks[b0:b1].copy()))
        assert_results_equal(run(three_calls), want, f"compact in three calls ({name})")
    # too many exceptions: every record's mate is far away
    far = reads.copy()
    far["mpos"] += 5_000_000
    assert api.pack_compact(far, blocks) is None
    # a corrupted header or a record that points outside the batch is an error, not a crash
    whole = api.pack_compact(reads, blocks).copy()
    bad = whole.copy(); bad[0] ^= 0xff
    ctx = api.Backend(libs["oracle"]).create(api.make_config()); ctx.set_annotation(ann)
    with pytest.raises(api.RsqError):
        ctx.submit_compact(bad)
    hdr = np.frombuffer(whole[:96].tobytes(), np.uint32).copy()
    off_reads = int(np.frombuffer(whole[48:56].tobytes(), np.uint64)[0])
    bad = whole.copy()
    bad[off_reads + 14:off_reads + 16] = np.frombuffer(np.uint16(hdr[4] + 5).tobytes(), np.uint8)     # combo index past the dictionary
    with pytest.raises(api.RsqError):
        ctx.submit_compact(bad)


def test_pinned_result_buffers_match_plain_ones(libs, tmp_path):
    """Context.finalize(pinned=True) / coverage_statistics(pinned=True): page-locked buffers from rsq_host_alloc, reused by
    the next call, same content as the default numpy buffers (CPU twin of what bench.py does on the GPU)."""
    import numpy as np
    from rnaseqc_b200 import synth
    api = libs["api"]
    sa = synth.make_annotation(seed=3, n_genes=120, contig_names=["chr1", "chr2"], contig_lens=[20_000_000, 10_000_000])
    synth.write_gtf(sa, str(tmp_path / "a.gtf"))
    ann = api.Annotation(str(tmp_path / "a.gtf"))
    sr = synth.simulate_reads(sa, 20000, seed=4)
    rb, bb = synth.pack(sr, [ann.contig_index(n) for n in sa.contig_names])
    reads, blocks = synth.to_numpy(rb, bb)
    ctx = api.Backend(libs["oracle"]).create(api.make_config())
    ctx.set_annotation(ann)
    ctx.submit(reads, blocks)
    plain = ctx.finalize(want_coverage=False)
    pinned = ctx.finalize(want_coverage=False, hist_len=8, pinned=True)
    for f in ("counters", "gene_reads", "gene_unique", "gene_fragments", "exon_whole", "exon_frac"):
        assert np.array_equal(getattr(plain, f), getattr(pinned, f)), f
    c1, c2 = ctx.coverage_statistics(ann), ctx.coverage_statistics(ann, pinned=True)
    for k in c1:
        assert np.array_equal(c1[k], c2[k], equal_nan=True), k
    again = ctx.finalize(want_coverage=False, hist_len=8, pinned=True)
    assert again.gene_reads.ctypes.data == pinned.gene_reads.ctypes.data      # the same page-locked memory
    assert plain.gene_reads.sum() > 0
    ctx.close()


def test_segmented_packer_writes_the_same_bytes_however_the_batch_is_cut(libs):
    """rsq_pack_compact_segments (include/rsq_b200.h): the ingest hands a batch over as pieces lying in its workers' buffers, each
    with its own block array and the contig still a BAM reference id; the compact batch must not depend on the cuts (rsq_pack_compact
    is the one-piece case), including empty pieces, pieces of one record, cuts inside a 32-record tile and records that travel as
    exceptions."""
    import ctypes as C
    import numpy as np
    from rnaseqc_b200 import synth
    api = libs["api"]
    H = api.host_lib()
    H.rsqh_pack_compact_split.restype = C.c_int
    H.rsqh_pack_compact_split.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    sa = synth.make_annotation(seed=31, n_genes=120, contig_names=["chr1", "chr2"], contig_lens=[2_000_000, 1_000_000])
    sr = synth.simulate_reads(sa, 40000, seed=32)
    rb, bb = synth.pack(sr, [0, 1])
    reads, blocks = synth.to_numpy(rb, bb)
    reads = reads.copy()
    rng = np.random.default_rng(9)
    odd = rng.choice(len(reads), 300, replace=False)              # records the 16-byte slot cannot express: exceptions
    reads["isize"][odd[:100]] = 70000
    reads["mpos"][odd[100:200]] = reads["pos"][odd[100:200]] + 50000
    reads["contig"][odd[200:]] = -1
    want = api.pack_compact(reads, blocks)
    assert want is not None
    cap = api.compact_bound(len(reads), len(blocks))
    for trial in range(12):
        k = int(rng.integers(0, 40))
        cuts = np.sort(rng.integers(0, len(reads) + 1, k)).astype(np.uint64)
        if trial == 0:
            cuts = np.array([0, 0, 1, 2, 31, 32, 33, 33, len(reads), len(reads)], np.uint64)
        raw = np.zeros(cap + 16, np.uint8)
        a = (-raw.ctypes.data) % 16
        out = raw[a:a + cap]
        w = C.c_size_t(0)
        rc = H.rsqh_pack_compact_split(reads.ctypes.data, len(reads), blocks.ctypes.data, len(blocks), cuts.ctypes.data, len(cuts), trial % 2,
                                       out.ctypes.data, cap, C.byref(w))
        assert rc == 0, (trial, rc)
        assert w.value == want.nbytes and np.array_equal(out[:w.value], want), trial
