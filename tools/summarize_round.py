"""Turns the files a `tools/gpu_round.sh` visit left in gpurun_out/ into tracked summaries under profiles/.
usage: python tools/summarize_round.py <tag>      (e.g. r01_e)"""
import collections, csv, json, os, re, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def main(tag):
    for name in ("bench.json", "bench_ref.json", "launches.csv"):
        src = os.path.join(G, name)
        if os.path.exists(src):
            shutil.copy(src, os.path.join(P, "%s_%s" % (tag, name.replace("bench_ref", "bench_reference"))))
    # launch list -> share table
    rows = [r for r in csv.reader(open(os.path.join(G, "launches.csv"))) if r and not r[0].startswith("==")]
    h = rows[0]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    d = collections.defaultdict(list)
    for r in rows[1:]:
        if len(r) > vi:
            d[r[ki][:80]].append(float(r[vi].replace(",", "")))
    mine = {k: v for k, v in d.items() if not k.startswith("void at::")}
    tot = sum(sum(v) for v in mine.values())
    lines = ["# ncu launch list of bench.py rollout steps (%s)" % tag, "",
             "`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare shares).",
             "torch's own kernels in the capture (the L2-flush fill between steps) are left out of the shares.", "",
             "kernel | launches | mean us | share of this repo's kernels", "---|---|---|---"]
    for k, v in sorted(mine.items(), key=lambda kv: -sum(kv[1])):
        lines.append("%s | %d | %.1f | %.1f%%" % (k, len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
    open(os.path.join(P, "%s_launch_summary.md" % tag), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
    # full captures
    traffic = {}
    for rep, label in (("env_step.ncu-rep", "env"), ("tc_linear.ncu-rep", "tc_linear")):
        path = os.path.join(G, rep)
        if not os.path.exists(path):
            continue
        h, u, rs = raw_metrics(path)
        open(os.path.join(P, "%s_%s_raw.csv" % (tag, label)), "w").write(
            subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)
        out = ["# %s kernels, ncu --set full (%s)" % (label, tag), ""]
        pat = (r"gpu__time_duration.sum|dram__bytes_(read|write).sum$|sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active|"
               r"smsp__issue_active.avg.pct|sm__warps_active.avg.pct_of_peak_sustained_active|smsp__inst_executed.sum$|"
               r"launch__registers_per_thread$|launch__grid_size|launch__block_size|launch__occupancy_limit_(registers|shared_mem)|"
               r"smsp__thread_inst_executed_per_inst_executed.ratio|l1tex__m_xbar2l1tex_read_bytes.sum$|"
               r"smsp__average_warps_issue_stalled_(barrier|wait|short_scoreboard|long_scoreboard)_per_issue_active")
        for r in rs:
            name = r[h.index("Kernel Name")]
            out.append("## %s" % name[:90])
            out.append("```")
            for i, m in enumerate(h):
                if re.search(pat, m):
                    out.append("%s %s %s" % (m, r[i], u[i]))
            out.append("```")
            if label == "env":
                mul = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
                rd = float(r[h.index("dram__bytes_read.sum")]) * mul[u[h.index("dram__bytes_read.sum")]]
                wr = float(r[h.index("dram__bytes_write.sum")]) * mul[u[h.index("dram__bytes_write.sum")]]
                traffic[name[:40]] = rd + wr
        open(os.path.join(P, "%s_%s_summary.md" % (tag, label)), "w").write("\n".join(out) + "\n")
    # learner step launch list
    lp = os.path.join(G, "learn_launches.csv")
    if os.path.exists(lp):
        rows = [r for r in csv.reader(open(lp)) if r and not r[0].startswith("==")]
        h = rows[0]
        ki, vi = h.index("Kernel Name"), h.index("Metric Value")
        d = collections.defaultdict(list)
        for r in rows[1:]:
            if len(r) > vi:
                d[r[ki][:80]].append(float(r[vi].replace(",", "")))
        tot = sum(sum(v) for v in d.values())
        lines = ["# ncu launch list of three CoPOPolicy.learn_on_batch calls, 65 536 rows, four networks (%s)" % tag, "",
                 "kernel | launches | mean us | share", "---|---|---|---"]
        for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
            lines.append("%s | %d | %.1f | %.1f%%" % (k, len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
        lines.append("")
        lines.append("total per call: %.2f ms (serialised under ncu)" % (tot / 3 / 1e6))
        open(os.path.join(P, "%s_learner_launches.md" % tag), "w").write("\n".join(lines) + "\n")
    for name in ("mlp_perf.json", "bookkeeping_perf.json"):
        if os.path.exists(os.path.join(G, name)):
            shutil.copy(os.path.join(G, name), os.path.join(P, "%s_%s" % (tag, name)))
    if traffic:
        json.dump({"kernel": "scene step (state + lidar kernels)", "dram_bytes_per_launch": sum(traffic.values()),
                   "per_kernel": traffic,
                   "source": "profiles/%s_env_raw.csv (ncu --set full, one launch each; outputs stay in the 126 MB L2, "
                             "so DRAM traffic is below the algorithmic bytes)" % tag},
                  open(os.path.join(P, "env_step_traffic.json"), "w"))
        print(traffic)


if __name__ == "__main__":
    main(sys.argv[1])
