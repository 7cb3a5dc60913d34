"""Aggregates an ncu report's per-SASS-instruction counters by CUDA source line (needs -lineinfo + --import-source).
usage: python tools/ncu_lines.py report.ncu-rep [top_n] [kernel-name substring]"""
import collections
import csv
import subprocess
import sys


def main(rep, top=40, kernel=None):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    agg = collections.OrderedDict()
    cur_file = cur = None
    skip = False
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            skip = kernel is not None and kernel not in r[1]
            continue
        if r[0] == "Line No" or skip:
            continue
        if r[0] != "":
            cur = (cur_file, r[0], r[1][:100])
            continue
        try:
            inst, tinst = int(r[7]), int(r[8])
            samp = int(r[6]) if r[6].isdigit() else 0
        except (ValueError, IndexError):
            continue
        a = agg.setdefault(cur, [0, 0, 0])
        a[0] += inst
        a[1] += tinst
        a[2] += samp
    tot = sum(a[0] for a in agg.values()) or 1
    tots = sum(a[2] for a in agg.values()) or 1
    print("total warp instructions %d, thread instructions %d (%.1f lanes/instr), samples %d" % (
        tot, sum(a[1] for a in agg.values()), sum(a[1] for a in agg.values()) / tot, tots))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
        print("%-14s %4s inst %5.1f%% lanes %4.1f stall-samples %5.1f%% | %s" % (
            k[0], k[1], 100 * a[0] / tot, a[1] / max(a[0], 1), 100 * a[2] / tots, k[2]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40, sys.argv[3] if len(sys.argv) > 3 else None)
