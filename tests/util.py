"""Shared helpers for the parity tests."""
import gzip
import math
import os
import shutil
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def run_cli(binary, gtf, aln, out_dir, *flags, check=True):
    p = subprocess.run([binary, gtf, aln, out_dir, *flags], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if check and p.returncode != 0:
        raise AssertionError(f"{binary} exited {p.returncode}: {p.stderr[-2000:]}")
    return p


def _cells(line):
    return line.rstrip("\n").split("\t")


def _is_int(s):
    try:
        int(s)
        return True
    except ValueError:
        return False


def compare_outputs(dir_a, dir_b, rel_tol=1e-6):
    """The reference's own comparator semantics (test_data/approx_diff.py): same rows, identical text and integers,
    floats within 1e-6 (relative, with an absolute floor of 1e-6 as approx_diff uses), same NaN pattern.
    Returns the list of files that were byte-identical."""
    files_a, files_b = sorted(os.listdir(dir_a)), sorted(os.listdir(dir_b))
    assert files_a == files_b, (files_a, files_b)
    identical = []
    for f in files_a:
        la, lb = open(os.path.join(dir_a, f)).readlines(), open(os.path.join(dir_b, f)).readlines()
        if la == lb:
            identical.append(f)
            continue
        assert len(la) == len(lb), f"{f}: {len(la)} vs {len(lb)} lines"
        for i, (x, y) in enumerate(zip(la, lb)):
            if x == y:
                continue
            if x.startswith("3' Bias, ") and "Percentile" in x:
                # the reference indexes ratios[ceil(.75 n) (+1)] without a bounds check (src/RNASeQC.cpp:486-507): with 2-4
                # bias genes it prints whatever lies past the vector.  A percentile of ratios outside [0, 1] is that garbage.
                try:
                    if not (0.0 <= float(_cells(x)[1]) <= 1.0):
                        continue
                except ValueError:
                    pass
            cx, cy = _cells(x.replace("-nan", "nan")), _cells(y.replace("-nan", "nan"))
            assert len(cx) == len(cy), f"{f}:{i + 1}: {x!r} vs {y!r}"
            for u, v in zip(cx, cy):
                if u == v:
                    continue
                assert not (_is_int(u) and _is_int(v)), f"{f}:{i + 1}: integer field differs: {x!r} vs {y!r}"
                fu, fv = float(u), float(v)
                assert math.isnan(fu) == math.isnan(fv), f"{f}:{i + 1}: {x!r} vs {y!r}"
                if not math.isnan(fu):
                    assert abs(fu - fv) <= max(1e-6, rel_tol * max(abs(fu), abs(fv))), f"{f}:{i + 1}: {x!r} vs {y!r}"
    return identical


def assert_results_equal(a, b, what=""):
    """bit-exact comparison of two api.Results"""
    from rnaseqc_b200 import api
    # "Fragment Candidates" is a diagnostic, not a reference output: the reference (and the oracle) stop evaluating
    # candidates once the sample quota is reached (src/RNASeQC.cpp:372), the data-parallel path cannot know that yet
    skip = {api.COUNTER_INDEX["Fragment Candidates"]}
    bad = [(api.COUNTER_NAMES[i], int(a.counters[i]), int(b.counters[i])) for i in range(len(api.COUNTER_NAMES))
           if a.counters[i] != b.counters[i] and i not in skip]
    assert not bad, f"{what} counters differ: {bad}"
    for name in ("gene_reads", "gene_unique", "gene_fragments", "exon_whole", "exon_frac"):
        x, y = getattr(a, name), getattr(b, name)
        diff = np.nonzero(x != y)[0]
        assert diff.size == 0, f"{what} {name} differs at {diff[:10]}: {x[diff[:10]]} vs {y[diff[:10]]}"
    if a.coverage is not None and b.coverage is not None:
        n = min(len(a.coverage), len(b.coverage)) - 1
        diff = np.nonzero(a.coverage[:n] != b.coverage[:n])[0]
        assert diff.size == 0, f"{what} coverage differs at {diff[:10]}"
    assert a.fragment_samples_taken == b.fragment_samples_taken, (what, a.fragment_samples_taken, b.fragment_samples_taken)
    assert np.array_equal(a.fragment_hist, b.fragment_hist), f"{what} fragment histogram differs"


def unpack_case(name, tmp):
    """Copies a golden case to tmp (gunzipping *.gz) and returns its directory."""
    src = os.path.join(GOLDEN, name)
    dst = os.path.join(str(tmp), name)
    os.makedirs(dst, exist_ok=True)
    for f in os.listdir(src):
        p = os.path.join(src, f)
        if os.path.isdir(p):
            shutil.copytree(p, os.path.join(dst, f), dirs_exist_ok=True)
        elif f.endswith(".gz"):
            with gzip.open(p, "rb") as fi, open(os.path.join(dst, f[:-3]), "wb") as fo:
                shutil.copyfileobj(fi, fo)
        else:
            shutil.copy(p, dst)
    return dst
