#!/usr/bin/env python
"""Per-source-line view of an .ncu-rep captured with --import-source on (needs -lineinfo): executed warp instructions and
stall samples aggregated by source line, and by named section of the hot kernel.

    python tools/ncu_source.py gpurun_out/prof.ncu-rep [--top 40] [--kernel hot_path]
"""
import argparse, csv, io, re, subprocess, sys, collections


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--kernel", default="hot_path")
    ap.add_argument("--sass", action="store_true", help="list the hottest SASS instructions instead of source lines")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", "--print-source", "sass,cuda" if False else "sass"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    # the CSV has one header per kernel; find the header row
    hdr_i = next(i for i, r in enumerate(rows) if "Source" in r and any("Instructions Executed" in c for c in r))
    hdr = rows[hdr_i]
    col = {c: i for i, c in enumerate(hdr)}
    c_src = col["Source"]
    c_exec = next(i for c, i in col.items() if c == "Instructions Executed" or c.startswith("Instructions Executed"))
    c_samp = next((i for c, i in col.items() if c.startswith("Warp Stall Sampling (All")), None)
    c_addr = col.get("Address")
    tot_exec = tot_samp = 0
    items = []
    for r in rows[hdr_i + 1:]:
        if len(r) <= c_exec:
            continue
        try:
            ex = int(r[c_exec] or 0)
            sm = int(r[c_samp] or 0) if c_samp is not None else 0
        except ValueError:
            continue
        items.append((ex, sm, r[c_src], r[c_addr] if c_addr is not None else ""))
        tot_exec += ex; tot_samp += sm
    print(f"instructions executed {tot_exec}, stall samples {tot_samp}, SASS lines {len(items)}")
    by_op = collections.Counter(); by_op_s = collections.Counter()
    for ex, sm, src, _ in items:
        op = src.strip().split()[0] if src.strip() else "?"
        if op.startswith("@"):
            op = src.strip().split()[1] if len(src.strip().split()) > 1 else op
        op = op.split(".")[0]
        by_op[op] += ex; by_op_s[op] += sm
    print("\nby opcode (executed %, samples %):")
    for op, ex in by_op.most_common(25):
        print(f"  {op:12s} {100.0 * ex / max(tot_exec, 1):6.2f}  {100.0 * by_op_s[op] / max(tot_samp, 1):6.2f}")
    print("\nhottest SASS by stall samples:")
    for ex, sm, src, addr in sorted(items, key=lambda t: -t[1])[:a.top]:
        print(f"  {100.0 * sm / max(tot_samp, 1):6.2f}% samp {100.0 * ex / max(tot_exec, 1):6.2f}% exec  {addr[-6:]:6s} {src.strip()[:110]}")


if __name__ == "__main__":
    main()
