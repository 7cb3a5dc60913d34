"""Round-2 evidence: turns gpurun_out/{r02_step,r02_learner,r02_fuse}.ncu-rep and launches.csv into tracked summaries
under profiles/ (key raw metrics per kernel, stall reasons by opcode, hottest source lines).
usage: python tools/summarize_r02.py [tag]      (tag: file prefix of the visit, default r02; e.g. r02_d)"""
import collections, csv, hashlib, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
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
    src = os.path.join(G, "launches.csv" if TAG == "r02" else TAG + "_launches.csv")
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
    lines = ["# ncu launch list of `bench.py --steps 4 --warmup 3 --train-iters 0` (%s)" % TAG, "",
             "`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare shares).", "",
             "kernel | launches | mean us | share of this repo's kernels", "---|---|---|---"]
    for k, v in sorted(mine.items(), key=lambda kv: -sum(kv[1])):
        lines.append("%s | %d | %.1f | %.1f%%" % (k, len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
    open(os.path.join(P, TAG + "_launch_summary.md"), "w").write("\n".join(lines) + "\n")
    open(os.path.join(P, TAG + "_launches.csv"), "w").write(open(src).read())
    print("\n".join(lines[:16]))


def traffic():
    """DRAM bytes of one scene step (state + lidar kernel) from the step capture -> profiles/env_step_traffic.json, keyed by
    the bench configuration and the hash of the kernel source it was taken on (bench.py reports it as roofline.traffic only
    while that hash matches)."""
    path = os.path.join(G, TAG + "_step.ncu-rep")
    if not os.path.exists(path):
        return
    rows = list(csv.reader(run("ncu", "-i", path, "--page", "raw", "--csv", "--metrics",
                               "dram__bytes_read.sum,dram__bytes_write.sum").splitlines()))
    h, u = rows[0], rows[1]
    ki, ri, wi = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per = {}
    for r in rows[2:]:
        if "env_" in r[ki] and r[ki] not in per:          # first launch of each scene-step kernel
            per[r[ki][:60]] = float(r[ri]) * scale[u[ri]] + float(r[wi]) * scale[u[wi]]
    sys.path.insert(0, ROOT)
    import bench                                          # the same hash bench.py checks (sources without comments)
    out = {"c2": {"kernel": "scene step (state + lidar kernels, policy operand included)", "dram_bytes_per_launch": sum(per.values()),
                  "per_kernel": per, "kernel_source_hash": bench._kernel_source_hash(),
                  "source": "profiles/%s_step_summary.md (ncu --set full of the bench command, one launch of each kernel)" % TAG}}
    json.dump(out, open(os.path.join(P, "env_step_traffic.json"), "w"), indent=1)
    print("traffic", out["c2"]["dram_bytes_per_launch"], per)


launches()
traffic()
summarize(TAG + "_step.ncu-rep", "Rollout step kernels (C2: 4096 x 40 Intersection): scene step + one-kernel policy network (%s)" % TAG, TAG + "_step_summary.md")
summarize(TAG + "_learner.ncu-rep", "Learner kernels at 65 536 rows (CoPO learn_on_batch): one-kernel forward, tc_linear (dgrad), tc_wgrad, head_backward, wgrad_reduce (%s)" % TAG, TAG + "_learner_summary.md")
summarize(TAG + "_fuse.ncu-rep", "cc_obs_fuse_kernel, mean-field (C3: 4096 x 40 Roundabout) (%s)" % TAG, TAG + "_fuse_summary.md")
