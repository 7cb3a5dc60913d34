"""Stall-reason totals of an ncu report's SASS page, grouped by opcode (needs --import-source on / --set full).
usage: python tools/ncu_stalls.py report.ncu-rep [top_n]"""
import collections, csv, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 14
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
agg = collections.defaultdict(lambda: collections.Counter())
inst = collections.Counter()
for r in rows:
    if len(r) > 3 and r[0] == "Address":
        hdr = r
        cols = [i for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
        i_inst = hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].startswith("0x"):
        continue
    toks = r[1].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0]
    inst[op] += int(r[i_inst] or 0)
    for i in cols:
        v = int(r[i] or 0)
        if v:
            agg[op][hdr[i]] += v
tot = sum(sum(c.values()) for c in agg.values())
print("total stall samples", tot, "total warp instructions", sum(inst.values()))
for op, c in sorted(agg.items(), key=lambda kv: -sum(kv[1].values()))[:top]:
    s = sum(c.values())
    print("%-10s inst %5.1f%%  samples %5.1f%%  " % (op, 100.0 * inst[op] / max(1, sum(inst.values())), 100.0 * s / tot) +
          " ".join("%s %.0f%%" % (k[6:], 100.0 * v / s) for k, v in c.most_common(5)))
