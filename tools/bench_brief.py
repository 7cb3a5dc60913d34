"""Prints the few numbers of a bench.py line worth reading at a glance.  usage: python tools/bench_brief.py file.json ..."""
import json
import sys

for p in sys.argv[1:]:
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e:
        print(p, "unreadable:", e)
        continue
    if d.get("impl") == "reference":
        print("%s: reference %.3f M/s on %s cores" % (p, d["value"] / 1e6, d["cpu_baseline"]["cores"]))
        continue
    r = d["roofline"]
    k = {a: (round(b * 1e3, 1) if b else None) for a, b in d["kernel_ms"].items()}
    print("%s: %s value %.1f M/s (%.4f ms/step) e2e %.1f M/s (%.3f ms) | scene step %.1f us frac %.3f (with operand %.3f) | us: %s | train %s" % (
        p, d["config"].get("name"), d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, d["e2e"]["ms_per_step"],
        r["kernel_ms"] * 1e3, r["frac"], r["with_policy_operand"]["frac"], k,
        ("%.2f M/s/gpu learn %s ms ar %s" % (d["train"]["agent_env_steps_per_s_per_gpu"] / 1e6,
                                            [round(x, 1) for x in d["train"]["learn_ms"]],
                                            d["train"]["allreduce_ms_per_iteration"])) if d.get("train") else None))
