"""Round-2 evidence: turns gpurun_out/{r02_step,r02_learner,r02_fuse}.ncu-rep and launches.csv into tracked summaries
under profiles/ (key raw metrics per kernel, stall reasons by opcode, hottest source lines).
usage: python tools/summarize_r02.py"""
import collections, csv, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
PAT = (r"gpu__time_duration.sum|dram__bytes_(read|write).sum$|sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active|"
       r"smsp__issue_active.avg.pct|sm__warps_active.avg.pct_of_peak_sustained_active|smsp__inst_executed.sum$|"
       r"launch__registers_per_thread$|launch__grid_size|launch__block_size|launch__occupancy_limit_(registers|shared_mem)|"
       r"smsp__thread_inst_executed_per_inst_executed.ratio|l1tex__m_xbar2l1tex_read_bytes.sum$|lts__t_bytes.sum$|"
       r"sm__inst_executed_pipe_(fma|fmaheavy|alu|xu|lsu|tensor).*sum$|dram__throughput.avg.pct_of_peak_sustained_elapsed|"
       r"smsp__average_warps_issue_stalled_(barrier|wait|short_scoreboard|long_scoreboard|math_pipe_throttle|not_selected)_per_issue_active")


def run(*a):
    return subprocess.run(list(a), capture_output=True, text=True).stdout


def summarize(rep, title, out_name):
    path = os.path.join(G, rep)
    if not os.path.exists(path):
        return
    raw = run("ncu", "-i", path, "--page", "raw", "--csv")
    rows = list(csv.reader(raw.splitlines()))
    h, u, rs = rows[0], rows[1], rows[2:]
    out = ["# %s" % title, "", "`ncu --set full --clock-control none --import-source on` under gpurun (one B200); per-launch values, "
           "cold caches, serialised - compare shares and ratios, not absolute times.", ""]
    for r in rs:
        out.append("## %s" % r[h.index("Kernel Name")][:100])
        out.append("```")
        for i, m in enumerate(h):
            if re.search(PAT, m) and r[i] not in ("", "n/a"):
                out.append("%-86s %s %s" % (m, r[i], u[i]))
        out.append("```")
    out += ["", "## stall reasons by opcode (all kernels of the capture)", "```",
            run(sys.executable, os.path.join(ROOT, "tools", "ncu_stalls.py"), path, "12").rstrip(), "```", "",
            "## hottest source lines", "```", run(sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), path, "30").rstrip()[:9000], "```"]
    open(os.path.join(P, out_name), "w").write("\n".join(out) + "\n")
    print("wrote", out_name)


def launches():
    src = os.path.join(G, "launches.csv")
    if not os.path.exists(src):
        return
    rows = [r for r in csv.reader(open(src)) if r and not r[0].startswith("==")]
    h = rows[0]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    d = collections.defaultdict(list)
    for r in rows[1:]:
        if len(r) > vi:
            d[r[ki][:90]].append(float(r[vi].replace(",", "")))
    mine = {k: v for k, v in d.items() if not k.startswith("void at::")}
    tot = sum(sum(v) for v in mine.values())
    lines = ["# ncu launch list of `bench.py --steps 4 --warmup 3 --train-iters 0` (r02)", "",
             "`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare shares).", "",
             "kernel | launches | mean us | share of this repo's kernels", "---|---|---|---"]
    for k, v in sorted(mine.items(), key=lambda kv: -sum(kv[1])):
        lines.append("%s | %d | %.1f | %.1f%%" % (k, len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
    open(os.path.join(P, "r02_launch_summary.md"), "w").write("\n".join(lines) + "\n")
    open(os.path.join(P, "r02_launches.csv"), "w").write(open(src).read())
    print("\n".join(lines[:16]))


launches()
summarize("r02_step.ncu-rep", "Rollout step kernels (C2: 4096 x 40 Intersection): scene step + one-kernel policy network (r02)", "r02_step_summary.md")
summarize("r02_learner.ncu-rep", "Learner kernels at 65 536 rows (CoPO learn_on_batch): tc_linear, tc_wgrad, head_backward, wgrad_reduce (r02)", "r02_learner_summary.md")
summarize("r02_fuse.ncu-rep", "cc_obs_fuse_kernel, mean-field (C3: 4096 x 40 Roundabout) (r02)", "r02_fuse_summary.md")
