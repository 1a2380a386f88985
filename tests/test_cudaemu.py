"""The CUDA kernels that have no CPU twin, run on the CPU: their own sources compiled for the host through tests/cudaemu/cuda_emu.h
(one OS thread per CUDA thread, block barriers, warp shuffles and atomics emulated) — the per-gene coverage statistics
(rsq_covstats.cu: warp-per-gene and CTA-per-gene paths, ownership masks), the mate-map stages (per-name automaton in stream
order, rounds 1 / 2 / replay), the coverage scan and the 16-byte wire-format expansion — against the CPU oracle or a literal
sequential restatement.  The `-m gpu` suite runs the same kernels on the device; this file is what keeps them honest between
GPU runs."""
import ctypes as C
import os
import random

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "oracle", "_build", "librsq_emu.so")


@pytest.fixture(scope="module")
def emu(libs):
    if not os.path.exists(EMU):
        import subprocess
        subprocess.run(["make", "-C", ROOT, "oracle"], check=True, stdout=subprocess.DEVNULL)
    return C.CDLL(EMU)


def p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


@pytest.fixture(scope="module")
def cov_case(libs, tmp_path_factory):
    """annotation + deep coverage from the oracle (so that the >= 100 peak / percentile trimming path of the bias code is taken)"""
    from rnaseqc_b200 import synth
    api = libs["api"]
    d = tmp_path_factory.mktemp("emu")
    sa = synth.make_annotation(seed=77, n_genes=70, contig_names=["chr1", "chr2"], contig_lens=[3_000_000, 1_500_000])
    gtf = str(d / "a.gtf")
    synth.write_gtf(sa, gtf)
    ann = api.Annotation(gtf)
    sr = synth.simulate_reads(sa, 60000, seed=78)
    rb, bb = synth.pack(sr, [ann.contig_index(n) for n in sa.contig_names])
    reads, blocks = synth.to_numpy(rb, bb)
    reads = np.repeat(reads, 5)
    ctx = api.Backend(libs["oracle"]).create(api.make_config())
    ctx.set_annotation(ann)
    ctx.submit(reads, blocks)
    res = ctx.finalize()
    return api, ann, ctx, res.coverage.copy()


def _emu_stats(emu, api, ann, cov, n_exons, owned=None, nan_prefill=True, **kw):
    q = api.RsqTranscripts()
    api.host_lib().rsqh_transcripts(ann._h, kw.get("mask", 500), kw.get("bias_offset", 0), kw.get("bias_window", 100), kw.get("bias_gene_length", 200), C.byref(q))
    n = q.n_genes
    out = {k: np.zeros(n, np.float64) for k in ("gene_mean", "gene_std", "gene_cv")}
    out.update(five_end=np.zeros(n, np.uint64), three_end=np.zeros(n, np.uint64), exon_cv=np.zeros(n_exons, np.float64))
    emu.emu_coverage_stats(p(cov), C.c_uint32(n), q.gene_cov_off, q.gene_exon_off, q.gene_strand, q.exon_len, q.exon_slot, C.c_uint32(q.coverage_mask),
                           C.c_int(q.bias_offset), C.c_int(q.bias_window), C.c_uint32(q.bias_gene_length), p(out["gene_mean"]), p(out["gene_std"]),
                           p(out["gene_cv"]), p(out["five_end"]), p(out["three_end"]), p(out["exon_cv"]), C.c_uint32(n_exons), p(owned), C.c_int(int(nan_prefill)))
    return out


def _close(got, want):
    assert np.array_equal(got["five_end"], want["five_end"]) and np.array_equal(got["three_end"], want["three_end"])
    for k in ("gene_mean", "gene_std", "gene_cv", "exon_cv"):
        a, b = got[k], want[k]
        assert np.array_equal(np.isnan(a), np.isnan(b)), k
        m = ~np.isnan(b)
        assert np.allclose(a[m], b[m], rtol=1e-9, atol=1e-12), (k, np.abs(a[m] - b[m]).max())


@pytest.mark.parametrize("kw", [dict(), dict(mask=100, bias_window=40, bias_gene_length=120), dict(mask=0, bias_offset=20),
                                dict(bias_window=300, mask=200), dict(bias_window=129, bias_gene_length=50, mask=0),
                                dict(bias_window=9000, bias_gene_length=150, mask=0), dict(bias_window=8193, bias_offset=7, bias_gene_length=100)])
def test_coverage_statistics_kernels_match_oracle(emu, cov_case, kw):
    """windows up to 128 take the warp-per-gene kernel for short transcripts and the CTA kernel for long ones; wider windows send
    every gene through the CTA kernel with its dynamic shared memory, windows above 8192 through the kernel with global scratch"""
    api, ann, ctx, cov = cov_case
    want = ctx.coverage_statistics(ann, **kw)
    got = _emu_stats(emu, api, ann, cov, ann.n_exons, **kw)
    _close(got, want)
    if kw.get("bias_window", 100) <= 8192:            # (no transcript of this fixture is as long as the two widest windows)
        assert (want["five_end"] > 0).sum() > 3, "bias path not exercised"
    assert (want["gene_mean"] > 0).sum() > 15


@pytest.mark.parametrize("kw", [dict(), dict(bias_window=300, mask=200), dict(mask=0, bias_offset=20, bias_window=129)])
def test_wide_window_kernel_sorts_in_global_scratch(emu, cov_case, kw, monkeypatch):
    """--window-size above 8192 (the reference takes any): a few CTAs walk over all genes with their two windows in global memory.
    The limit is lowered for the test, so that this fixture's transcripts fill and sort such windows."""
    api, ann, ctx, cov = cov_case
    want = ctx.coverage_statistics(ann, **kw)
    monkeypatch.setenv("RSQ_COVSTATS_SMEM_WINDOW", "64")
    got = _emu_stats(emu, api, ann, cov, ann.n_exons, **kw)
    _close(got, want)
    assert (want["five_end"] > 0).sum() > 3, "bias path not exercised"


def test_coverage_statistics_ownership_masks_sum_to_the_whole(emu, cov_case):
    """two partitions, each evaluating its own genes: non-owned outputs are zero bits, unreferenced exon slots NaN on partition 0
    only, and the u64 sum of the two equals the single evaluation bit for bit"""
    api, ann, ctx, cov = cov_case
    whole = _emu_stats(emu, api, ann, cov, ann.n_exons)
    n = len(whole["gene_mean"])
    own0 = (np.arange(n) % 3 != 0).astype(np.uint8)
    a = _emu_stats(emu, api, ann, cov, ann.n_exons, owned=own0, nan_prefill=True)
    b = _emu_stats(emu, api, ann, cov, ann.n_exons, owned=(1 - own0).astype(np.uint8), nan_prefill=False)
    for k in whole:
        s = a[k].view(np.uint64) + b[k].view(np.uint64)
        assert np.array_equal(s, whole[k].view(np.uint64)), k


def test_coverage_scan_kernels(emu):
    rng = np.random.default_rng(5)
    for n in (1, 7, 4096, 4097, 70001):
        diff = rng.integers(0, 3, n).astype(np.uint32)
        diff[rng.integers(0, n, max(1, n // 50))] -= np.uint32(2)          # wrapping negative steps, like the -1 ends of reads
        cov = np.zeros(n, np.uint32)
        emu.emu_coverage_scan(p(diff), p(cov), C.c_ulonglong(n))
        assert np.array_equal(cov, np.cumsum(diff, dtype=np.uint32)), n


def _automaton(stream, quota):
    """the reference's map per read name (src/Expression.cpp:508-533) applied in stream order"""
    state, samples = {}, []
    for name, name_id, endpos, isz, ok in stream:
        if name not in state:
            state[name] = (name_id, endpos)
        elif state[name][0] == name_id:
            if ok and endpos > state[name][1]:
                samples.append(isz)
                del state[name]
    return samples, len(samples)


@pytest.mark.parametrize("seed,quota", [(1, 10 ** 6), (2, 10 ** 6), (3, 40), (4, 10 ** 6), (5, 7)])
def test_mate_map_stages_replay_every_name_in_stream_order(emu, seed, quota):
    """names shared by one to seven records, spread over launches of uneven size, candidates handed over in scrambled order (as
    the device emits them): round 1 / round 2 / replay must reproduce the sequential map, erase-and-restart included"""
    rng = random.Random(seed)
    n_records = 6000
    names = [rng.randrange(1, 900) * 0x9E3779B97F4A7C15 % (1 << 64) for _ in range(n_records)]      # ~6.7 records per name
    if seed == 4:
        names = [names[0]] * 300 + names[300:]                                                         # one name 300 times in a row
    stream = []          # (record index, name, interval, endpos, isize, ok) for the candidate records only
    for i in range(n_records):
        if rng.random() < 0.7:
            stream.append((i, names[i], rng.randrange(0, 3), 1000 + i + rng.randrange(-40, 40), rng.randrange(100, 600), rng.random() < 0.8))
    want, _ = _automaton([(s[1], s[2], s[3], s[4], s[5]) for s in stream], quota)
    cap = 1 << 14
    keys, first, second = (np.zeros(cap, np.uint64) for _ in range(3))
    payload = np.full(cap, np.uint64(0xFFFFFFFFFFFFFFFF))
    vals = np.zeros(max(quota, 1) if quota < 10 ** 6 else 20000, np.uint32)
    q = len(vals)
    count, err, totals = np.zeros(1, np.uint64), np.zeros(2, np.uint32), np.zeros(2, np.uint64)
    base, launch = 0, 0
    while base < n_records:
        n = min(n_records - base, rng.choice([37, 200, 1500, 3000]))
        launch += 1
        c = [s for s in stream if base <= s[0] < base + n]
        rng.shuffle(c)
        arr = lambda k, dt: np.array([x[k] for x in c], dt) if c else np.zeros(1, dt)
        emu.emu_bed_launch(C.c_uint32(n), C.c_uint32(len(c)), p(arr(1, np.uint64)), p(np.array([x[0] - base for x in c], np.uint32) if c else np.zeros(1, np.uint32)),
                           p(arr(2, np.uint32)), p(arr(3, np.int32)), p(arr(4, np.uint32)), p(np.array([int(x[5]) for x in c], np.uint32) if c else np.zeros(1, np.uint32)),
                           C.c_ulonglong(cap), p(keys), p(first), p(second), p(payload), C.c_ulonglong(launch), C.c_uint32(q), p(vals), p(count),
                           C.c_ulonglong(base), p(err), p(totals))
        base += n
    assert err[0] == 0
    assert int(count[0]) == len(want)
    take = min(len(want), q)
    assert list(vals[:take]) == want[:take]
    assert int(totals[0]) == len(stream)
    if seed == 4:
        assert len(want) > 20


def test_compact_expansion_kernel_matches_its_input(emu, libs, tmp_path):
    from test_abi import _compact_cases
    api = libs["api"]
    ann, reads, blocks, odd = _compact_cases(api, tmp_path)
    for rr in (reads, odd):
        buf = api.pack_compact(rr, blocks)
        out_r = np.zeros(len(rr), api.READ_DTYPE)
        out_b = np.zeros(int(rr["n_blocks"].sum()) + 1, api.BLOCK_DTYPE)
        err = np.zeros(2, np.uint32)
        emu.emu_expand_compact(p(buf), p(out_r), p(out_b), p(err))
        assert err[0] == 0
        for f in ("pos", "mpos", "endpos", "isize", "l_qseq", "nm", "flag", "mapq", "aux", "contig", "filter_mask", "name_hash", "n_blocks"):
            assert np.array_equal(out_r[f], rr[f]), f
        # blocks: the same per record, offsets are the running sum
        assert np.array_equal(out_r["block_off"], np.concatenate([[0], np.cumsum(rr["n_blocks"][:-1])]).astype(np.uint32))
        for i in np.random.default_rng(3).integers(0, len(rr), 400):
            a, b = int(rr["block_off"][i]), int(out_r["block_off"][i])
            assert np.array_equal(blocks[a:a + int(rr["n_blocks"][i])], out_b[b:b + int(rr["n_blocks"][i])]), i
