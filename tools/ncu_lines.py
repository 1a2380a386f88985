#!/usr/bin/env python
"""Per CUDA source line (file:line) totals of an .ncu-rep captured with --import-source on: executed warp instructions, stall
samples, long-scoreboard samples.   python tools/ncu_lines.py rep [--top N]"""
import argparse, csv, io, subprocess, collections
ap = argparse.ArgumentParser(); ap.add_argument("rep"); ap.add_argument("--top", type=int, default=60); a = ap.parse_args()
raw = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur_file, hdr, lines = None, None, []
tot_e = tot_s = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = {c: i for i, c in enumerate(r)}
        ci = [i for i, c in enumerate(r) if c == "Source"]
        continue
    if hdr is None or not r or r[0] == "Function Name" or not r[0]:
        continue      # SASS rows have an empty line number
    try:
        ex = int(r[hdr["Instructions Executed"]] or 0); sm = int(r[hdr["Warp Stall Sampling (All Samples)"]] or 0)
        lsb = int(r[hdr["stall_long_sb"]] or 0); ssb = int(r[hdr["stall_short_sb"]] or 0); wt = int(r[hdr["stall_wait"]] or 0)
    except (ValueError, KeyError):
        continue
    lines.append((cur_file, r[0], r[ci[0]].strip()[:100], ex, sm, lsb, ssb, wt))
    tot_e += ex; tot_s += sm
print(f"total executed {tot_e}  samples {tot_s}")
byfile = collections.Counter(); byfile_s = collections.Counter()
for f, ln, src, ex, sm, *_ in lines:
    byfile[f] += ex; byfile_s[f] += sm
for f, ex in byfile.most_common():
    print(f"  {f:28s} exec {100 * ex / tot_e:6.2f}%  samples {100 * byfile_s[f] / tot_s:6.2f}%")
print("\n  exec%  samp%  lsb%  ssb%  wait%  file:line  source")
for f, ln, src, ex, sm, lsb, ssb, wt in sorted(lines, key=lambda t: -(t[3] / tot_e + t[4] / tot_s))[:a.top]:
    print(f"  {100 * ex / tot_e:5.2f}  {100 * sm / tot_s:5.2f} {100 * lsb / tot_s:5.2f} {100 * ssb / tot_s:5.2f} {100 * wt / tot_s:5.2f}  {f}:{ln}  {src}")
