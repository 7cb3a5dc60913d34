"""Per source line instruction / stall-sample totals of one kernel in an ncu report (source page, cuda+sass view).
usage: ncu_lines2.py report.ncu-rep kernel-regex [top]"""
import collections, csv, subprocess, sys


def main(rep, kernel, top=45):
    out = subprocess.run(["ncu", "-i", rep, "-k", "regex:" + kernel, "--page", "source", "--csv", "--print-source",
                          "cuda,sass"], capture_output=True, text=True).stdout
    inst, thr, samp, text = collections.Counter(), collections.Counter(), collections.Counter(), {}
    cur_file, key = None, None
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] in ("Function Name", "Line No"):
            continue
        if r[0] != "":
            key = (cur_file, int(r[0]) if r[0].isdigit() else 0)
            text[key] = r[1].strip()
            continue
        try:
            i, t = int(r[7]), int(r[8])
            s = int(r[6]) if r[6].isdigit() else 0
        except (ValueError, IndexError):
            continue
        inst[key] += i; thr[key] += t; samp[key] += s
    tot, ts = sum(inst.values()), sum(samp.values())
    print("total warp instructions %d, samples %d" % (tot, ts))
    byfile = collections.Counter()
    for k, v in inst.items():
        byfile[k[0]] += v
    print({k: "%.1f%%" % (100 * v / tot) for k, v in byfile.most_common(8)})
    for k, v in inst.most_common(int(top)):
        print("%-16s %4d inst %5.2f%% lanes %4.1f samp %5.2f%%  %s" % (k[0][:16], k[1], 100 * v / tot, thr[k] / max(v, 1),
                                                                   100 * samp[k] / max(ts, 1), text.get(k, "")[:90]))


if __name__ == "__main__":
    main(*sys.argv[1:])
