#!/usr/bin/env python
"""Byte-mutation fuzz of the multi-threaded BAM ingest (csrc/host/bamparallel.cpp): every mutant of a several-job BAM must give the
same exit code and the same report files whether its jobs are framed on the speculated record chain and observed through their
digests (default) or framed by the sequential hop and observed record by record (RSQ_INGEST_NO_SPECULATION=1
RSQ_INGEST_NO_SUMMARY=1), whether its BGZF blocks go through the ingest's own DEFLATE decoder (fastinflate.cpp, default) or through
zlib alone (RSQ_INGEST_ZLIB=1), and, when the file is still readable, as the sequential reader (--decode-threads 0).
CPU only (oracle-backed CLI, test infrastructure).   python tools/ingest_mutation_fuzz.py [--mutants 150] [--seed 1]"""
import argparse, filecmp, os, random, shutil, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rnaseqc_b200 import synth  # noqa: E402


def run(cli, gtf, bam, out, threads, env):
    p = subprocess.run([cli, gtf, bam, out, "--decode-threads", str(threads)], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, env=env, timeout=120)
    return p.returncode


def same_dir(a, b):
    fa, fb = sorted(os.listdir(a)) if os.path.isdir(a) else [], sorted(os.listdir(b)) if os.path.isdir(b) else []
    return fa == fb and all(filecmp.cmp(os.path.join(a, f), os.path.join(b, f), shallow=False) for f in fa)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mutants", type=int, default=150)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--records", type=int, default=60000)
    a = ap.parse_args()
    cli = os.path.join(ROOT, "oracle", "_build", "rnaseqc_oracle")
    d = tempfile.mkdtemp(prefix="rsq_mut_")
    sa = synth.make_annotation(seed=7, n_genes=150, contig_names=["chr1", "chr2"], contig_lens=[20_000_000, 10_000_000])
    gtf = os.path.join(d, "a.gtf")
    synth.write_gtf(sa, gtf)
    sr = synth.simulate_reads(sa, a.records, seed=8, dup_frac=0.0)
    rng = random.Random(a.seed)
    env_on = dict(os.environ); env_on.pop("RSQ_INGEST_NO_SPECULATION", None)
    env_on.pop("RSQ_INGEST_NO_SUMMARY", None)
    env_on.pop("RSQ_INGEST_ZLIB", None)
    env_off = dict(env_on, RSQ_INGEST_NO_SPECULATION="1", RSQ_INGEST_NO_SUMMARY="1")
    env_zlib = dict(env_on, RSQ_INGEST_ZLIB="1")
    bad = 0
    codes = {}
    for level in (0, 6):                       # stored blocks: mutations land in record bytes; deflated: mostly in the framing
        base = os.path.join(d, f"base{level}.bam")
        synth.write_bam(sa, sr, base, level=level)
        data = bytearray(open(base, "rb").read())
        for m in range(a.mutants // 2):
            mut = bytearray(data)
            kind = rng.choice(["byte", "byte", "byte", "burst", "truncate", "zero32"])
            at = rng.randrange(len(mut))
            if kind == "byte": mut[at] = rng.randrange(256)
            elif kind == "burst":
                for k in range(at, min(len(mut), at + rng.choice([2, 4, 16, 64]))): mut[k] = rng.randrange(256)
            elif kind == "zero32": mut[at:at + 4] = b"\0" * min(4, len(mut) - at)
            else: del mut[at:]
            bam = os.path.join(d, "m.bam")
            open(bam, "wb").write(mut)
            outs = [os.path.join(d, n) for n in ("on", "off", "seq", "zlib")]
            for o in outs: shutil.rmtree(o, ignore_errors=True)
            rc_on = run(cli, gtf, bam, outs[0], 3, env_on)
            rc_off = run(cli, gtf, bam, outs[1], 3, env_off)
            rc_seq = run(cli, gtf, bam, outs[2], 0, env_on)
            rc_zlib = run(cli, gtf, bam, outs[3], 3, env_zlib)
            codes[rc_on] = codes.get(rc_on, 0) + 1
            ok = rc_on == rc_off and (rc_on != 0 or same_dir(outs[0], outs[1]))
            if rc_on == 0 and rc_seq == 0: ok = ok and same_dir(outs[0], outs[2])
            ok = ok and rc_zlib == rc_on and (rc_on != 0 or same_dir(outs[0], outs[3]))
            if not ok:
                bad += 1
                keep = os.path.join(d, f"bad_{level}_{m}.bam")
                shutil.copy(bam, keep)
                print(f"DIFFERENT: level {level} mutant {m} ({kind} at {at}): exit {rc_on} / {rc_off} / sequential {rc_seq} / zlib {rc_zlib}; kept {keep}")
    print(f"{a.mutants} mutants, exit codes {codes}, {bad} differences")
    if not bad: shutil.rmtree(d, ignore_errors=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
