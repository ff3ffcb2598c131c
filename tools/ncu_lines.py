#!/usr/bin/env python3
"""Stall samples and executed instructions per CUDA source line of a kernel in an .ncu-rep captured with
--import-source on (kernel compiled with -lineinfo):  python tools/ncu_lines.py report.ncu-rep [top_n]"""
import collections
import csv
import subprocess
import sys


def main(rep, top=40):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    cur, hdr, out = None, None, []
    for r in csv.reader(txt.splitlines()):
        if not r:
            continue
        if r[0] in ("File Path", "File Name"):
            cur, hdr = r[1], None
        elif r[0] == "Line No":
            hdr = r
        elif hdr and cur and r[0].isdigit():
            try:
                out.append((int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")]),
                            cur.split("/")[-1], int(r[0]), r[1].strip()[:100]))
            except (ValueError, IndexError):
                pass
    tot, totx = sum(o[0] for o in out) or 1, sum(o[1] for o in out) or 1
    print(f"total samples {tot}  warp instructions {totx}")
    by = collections.Counter()
    for o in out:
        by[o[2]] += o[0]
    print("samples by file:", ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in by.most_common()))
    for o in sorted(out, reverse=True)[:top]:
        print(f"{100 * o[0] / tot:5.1f}%s {100 * o[1] / totx:5.1f}%i {o[2]}:{o[3]}  {o[4]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
