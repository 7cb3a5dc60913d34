"""Opcode histogram of every kernel in copo_b200/libcopo_b200.so (cuobjdump -sass): the Blackwell-specific mnemonics
(UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UBLKCP = bulk copy, FFMA2/FMUL2/FADD2 = packed
fp32, STG.*.256, REDUX, ACQBULK/PREEXIT = programmatic dependent launch) plus the ten most frequent opcodes.
usage: python tools/sass_histogram.py > profiles/rNN_sass_histogram.md"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "copo_b200", "libcopo_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEY = ["UTCHMMA", "LDTM", "UTMALDG", "UBLKCP", "UTMASTG", "SYNCS", "FFMA2", "FMUL2", "FADD2", "MUFU", "REDUX", "STG.256",
       "LDS", "STS", "ACQBULK", "PREEXIT", "NANOSLEEP", "ATOM", "RED", "BAR"]
cur, hist = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur:
        hist[cur][m.group(2)] += 1
print("# SASS opcode histogram of libcopo_b200.so (sm_100a), per kernel\n")
print("`cuobjdump -sass copo_b200/libcopo_b200.so`, counted by tools/sass_histogram.py.  Columns: total instructions, the "
      "Blackwell / design-relevant mnemonics (prefix match), then the ten most frequent opcodes.\n")
print("| kernel | instr | " + " | ".join(KEY) + " | top opcodes |")
print("|---|---|" + "---|" * len(KEY) + "---|")
for k, c in hist.items():
    tot = sum(c.values())
    def cnt(p):
        if p == "STG.256":
            return sum(v for o, v in c.items() if o.startswith("STG") and "256" in o)
        return sum(v for o, v in c.items() if o.split(".")[0] == p)
    base = collections.Counter()
    for o, v in c.items():
        base[o.split(".")[0]] += v
    top = ", ".join("%s %d" % (o, v) for o, v in base.most_common(10))
    print("| `%s` | %d | " % (k[:80], tot) + " | ".join(str(cnt(p)) if cnt(p) else "" for p in KEY) + " | %s |" % top)
