"""`from ray import tune` for the reference's training scripts (torch_copo/train_copo.py:3, utils/train.py:6-7, 181-192):
`grid_search`, `run`, `CLIReporter`.  Trials run one after the other in this process on the current GPU (under torchrun
every rank runs the same trial list; scenes shard across ranks inside the trainer), each as
`trainer(config=variant)` + `.train()` until a `stop` criterion holds; every trial gets a directory with
`params.json`, `progress.csv`, `result.json` and a final checkpoint, as Tune lays them out."""
import copy
import csv
import itertools
import json
import os
import time

from .registry import register_env  # noqa: F401


def grid_search(values):
    return {"grid_search": list(values)}


class CLIReporter:
    DEFAULT_COLUMNS = {"training_iteration": "iter", "time_total_s": "total time (s)", "timesteps_total": "ts",
                       "episode_reward_mean": "reward"}

    def __init__(self, metric_columns=None, **kwargs):
        self.metric_columns = dict(metric_columns or self.DEFAULT_COLUMNS)

    def add_metric_column(self, metric, representation=None):
        self.metric_columns[metric] = representation or metric

    def report(self, trial_name, result):
        cells = []
        for k, label in self.metric_columns.items():
            v = result.get(k)
            if isinstance(v, float):
                v = "%.4g" % v
            cells.append("%s=%s" % (label, v))
        print("[%s] %s" % (trial_name, " ".join(cells)), flush=True)


def _grid_paths(node, prefix=()):
    """Paths of every {"grid_search": [...]} leaf of a nested config."""
    out = []
    if isinstance(node, dict):
        if set(node.keys()) == {"grid_search"}:
            return [(prefix, node["grid_search"])]
        for k, v in node.items():
            out += _grid_paths(v, prefix + (k,))
    return out


def generate_variants(config):
    """Cartesian product of the grid-search leaves (Tune's BasicVariantGenerator for grid_search)."""
    paths = _grid_paths(config)
    if not paths:
        return [({}, copy.deepcopy(config))]
    variants = []
    for combo in itertools.product(*[vals for _, vals in paths]):
        cfg = copy.deepcopy(config)
        tag = {}
        for (path, _), val in zip(paths, combo):
            node = cfg
            for k in path[:-1]:
                node = node[k]
            node[path[-1]] = val
            tag["/".join(path)] = val
        variants.append((tag, cfg))
    return variants


def _flatten(d, prefix=""):
    out = {}
    for k, v in d.items():
        key = "%s/%s" % (prefix, k) if prefix else str(k)
        if isinstance(v, dict):
            out.update(_flatten(v, key))
        elif isinstance(v, (int, float, str, bool)) or v is None:
            out[key] = v
    return out


def _jsonable(o):
    if isinstance(o, dict):
        return {str(k): _jsonable(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_jsonable(v) for v in o]
    if isinstance(o, (int, float, str, bool)) or o is None:
        return o
    return getattr(o, "__name__", repr(o))


class Trial:
    def __init__(self, name, config, tag, logdir):
        self.trial_name, self.config, self.evaluated_params, self.logdir = name, config, tag, logdir
        self.results, self.last_result, self.checkpoint, self.status = [], {}, None, "PENDING"


class ExperimentAnalysis:
    def __init__(self, trials):
        self.trials = trials

    def fetch_trial_dataframes(self):
        import pandas as pd
        return {t.logdir: pd.DataFrame([_flatten(r) for r in t.results]) for t in self.trials}

    @property
    def results(self):
        return {t.trial_name: t.last_result for t in self.trials}

    def get_best_trial(self, metric="episode_reward_mean", mode="max"):
        key = lambda t: t.last_result.get(metric, float("-inf") if mode == "max" else float("inf"))
        return (max if mode == "max" else min)(self.trials, key=key)


def _should_stop(stop, result):
    if stop is None:
        return False
    if callable(stop):
        return bool(stop(result.get("trial_id"), result))
    return any(k in result and result[k] >= v for k, v in stop.items())


def run(run_or_experiment, name=None, stop=None, config=None, checkpoint_freq=0, checkpoint_at_end=False,
        local_dir=".", max_failures=0, verbose=1, progress_reporter=None, callbacks=None, **kwargs):
    """Runs every grid variant of `config` with the trainer class `run_or_experiment`; returns an ExperimentAnalysis."""
    trainer_cls = run_or_experiment
    cls_name = getattr(trainer_cls, "_name", None) or getattr(trainer_cls, "__name__", str(trainer_cls))
    exp_dir = os.path.join(os.path.abspath(os.path.expanduser(local_dir)), name or cls_name)
    rank0 = int(os.environ.get("RANK", "0")) == 0
    reporter = progress_reporter or CLIReporter()
    max_iters = int(os.environ.get("B2C_COMPAT_MAX_ITERS", "0"))
    dry = os.environ.get("B2C_COMPAT_DRY_RUN") == "1"
    if rank0 and not dry:
        os.makedirs(exp_dir, exist_ok=True)
    trials = []
    for n, (tag, cfg) in enumerate(generate_variants(config or {})):
        env_name = cfg.get("env") if isinstance(cfg.get("env"), str) else getattr(cfg.get("env"), "__name__", "env")
        trial_name = "%s_%s_%05d" % (cls_name, env_name, n)
        trial = Trial(trial_name, cfg, tag, os.path.join(exp_dir, trial_name))
        trials.append(trial)
        if dry:
            trial.status = "DRY_RUN"
            continue
        if rank0:
            os.makedirs(trial.logdir, exist_ok=True)
            json.dump(_jsonable(cfg), open(os.path.join(trial.logdir, "params.json"), "w"), indent=2)
        algo = trainer_cls(config=cfg)
        user_cb = getattr(algo, "callbacks", None)
        trial.status = "RUNNING"
        t0 = time.time()
        writer = fh = None
        try:
            while True:
                result = algo.train()
                result["time_total_s"] = time.time() - t0
                result["trial_id"] = trial_name
                if user_cb is not None and hasattr(user_cb, "on_train_result"):
                    user_cb.on_train_result(algorithm=algo, result=result)
                trial.results.append(result)
                trial.last_result = result
                if rank0:
                    flat = _flatten(result)
                    if writer is None:
                        fh = open(os.path.join(trial.logdir, "progress.csv"), "w", newline="")
                        writer = csv.DictWriter(fh, fieldnames=list(flat.keys()), extrasaction="ignore")
                        writer.writeheader()
                    writer.writerow(flat)
                    fh.flush()
                    with open(os.path.join(trial.logdir, "result.json"), "a") as jf:
                        jf.write(json.dumps(_jsonable(result)) + "\n")
                    if verbose:
                        reporter.report(trial_name, result)
                    it = result.get("training_iteration", len(trial.results))
                    if checkpoint_freq and it % checkpoint_freq == 0 and hasattr(algo, "save"):
                        trial.checkpoint = algo.save(trial.logdir)
                if _should_stop(stop, result) or (max_iters and len(trial.results) >= max_iters):
                    break
            if checkpoint_at_end and rank0 and hasattr(algo, "save"):
                trial.checkpoint = algo.save(trial.logdir)
            trial.status = "TERMINATED"
        finally:
            if fh is not None:
                fh.close()
            if hasattr(algo, "stop"):
                algo.stop()
    return ExperimentAnalysis(trials)
