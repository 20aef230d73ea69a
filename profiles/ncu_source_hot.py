#!/usr/bin/env python
"""Hot source lines of one ncu capture: stall samples per CUDA source line.

usage: python profiles/ncu_source_hot.py <file.ncu-rep> [top_n]
Needs a capture made with `--import-source on` of a library built with -lineinfo.  Sums the `Warp Stall Sampling (All
Samples)` column of the SASS instructions under every source line (ncu --page source --print-source cuda,sass --csv).
"""
import csv, subprocess, sys

def main():
    rep = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    cur_file, col, lines = "?", None, {}
    cur = None
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            col = r.index("Warp Stall Sampling (All Samples)")
            continue
        if col is None or len(r) <= col:
            continue
        if r[0]:  # a source line
            cur = (cur_file, int(r[0]), r[1].strip())
            lines.setdefault(cur, 0.0)
        elif cur is not None:
            try:
                lines[cur] += float(r[col])
            except ValueError:
                pass
    tot = sum(lines.values()) or 1.0
    print(f"total stall samples {tot:.0f}")
    top = sorted(lines.items(), key=lambda kv: -kv[1])[:top_n]
    for (f, ln, src), v in sorted(top, key=lambda kv: (kv[0][0], kv[0][1])):
        print(f"{v:8.0f} {100 * v / tot:5.1f}%  {f}:{ln}: {src[:120]}")

if __name__ == "__main__":
    main()
