"""Per-function instruction / stall-sample shares of an ncu report (functions of sim_core.cuh; other files by name)."""
import collections, csv, re, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def main(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    src = open(os.path.join(ROOT, 'copo_b200/csrc/sim_core.cuh')).read().split('\n')
    func_at, cur = {}, None
    for n, line in enumerate(src, 1):
        m = re.match(r'B2C_HD\s+[\w:<> ]+?\s+(\w+)\(', line)
        if m: cur = m.group(1)
        func_at[n] = cur
    agg, lanes, samp = collections.Counter(), collections.Counter(), collections.Counter()
    cur_file = key = None
    for r in rows:
        if not r: continue
        if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
        if r[0] in ("Function Name", "Line No"): continue
        if r[0] != "":
            ln = int(r[0]) if r[0].isdigit() else 0
            key = func_at.get(ln, '?') if cur_file == 'sim_core.cuh' else cur_file
            continue
        try: inst, tinst = int(r[7]), int(r[8]); s = int(r[6]) if r[6].isdigit() else 0
        except (ValueError, IndexError): continue
        agg[key] += inst; lanes[key] += tinst; samp[key] += s
    tot, ts = sum(agg.values()), sum(samp.values())
    for k, v in agg.most_common(22):
        print("%-28s inst %5.1f%% lanes %4.1f samples %5.1f%%" % (k, 100*v/tot, lanes[k]/max(v,1), 100*samp[k]/max(ts,1)))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines())); h = rr[0]
    for name in ("gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "dram__bytes_read.sum", "dram__bytes_write.sum",
                 "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"):
        if name in h: print(name, rr[2][h.index(name)], rr[1][h.index(name)])

if __name__ == "__main__":
    main(sys.argv[1])
