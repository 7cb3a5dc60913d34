"""Generates tests/golden/ref_golden.npz by EXECUTING the reference's own functions (build container only: needs
/root/reference; nothing is copied - the sources are read, the named functions / methods are extracted with `ast` and
exec'd here against small stand-ins for the RLlib names they touch):

    python tests/golden/make_ref_golden.py

The reference imports ray / rllib / metadrive at module level, which are not installed, so its modules cannot be
imported; but the bodies this path depends on are plain numpy / torch code:

  torch_copo/utils/env_wrappers.py   CCEnv._update_distance_map, CCEnv._find_in_range            (:125-158)
                                     CCEnv.step, LCFEnv.step, LCFEnv._add_lcf over a scripted base env (:89-123, 307-418)
  torch_copo/algo_copo.py            compute_nei_advantage, compute_global_advantage             (:189-204)
                                     CoPOModel.compute_coordinated / lcf_dist / lcf_mean / lcf_std (:155-177)
                                     CoPOPolicy.loss, CoPOPolicy.meta_update                     (:228-424)
  torch_copo/algo_ccppo.py           concat_ccppo_process, mean_field_ccppo_process, CCPPOPolicy.loss (:225-311, 376-472)
  torch_copo/algo_ippo.py            IPPOPolicy.loss                                             (:79-172)

Stand-ins: `SampleBatch` / `Postprocessing` column-name constants (rllib's strings), `discount_cumsum` (rllib 2.2.0:
scipy.signal.lfilter([1], [1, -gamma], x[::-1])[::-1]), a model adapter around oracle/models.py's torch networks (their
forward is pinned separately by tests/golden/mlp_golden.npz) and a DiagGaussian adapter (pinned against
torch.distributions in tests/test_bookkeeping_cpu.py).  Stored: every input (weights, batches, positions, seeds) and
what the reference's code returned, so the tests need neither /root/reference nor this script.
"""
import ast
import json
import os
import sys
import textwrap
from collections import defaultdict

import numpy as np
import scipy.signal
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import models as om  # noqa: E402

REF = "/root/reference/copo_code/copo/torch_copo"


# ---- source extraction ------------------------------------------------------------------------------------------------
def _tree(path):
    src = open(path).read()
    return src, ast.parse(src)


def extract(path, qualname):
    """Source text of a module-level function `name` or a method `Class.name`, dedented, decorators dropped."""
    src, tree = _tree(path)
    parts = qualname.split(".")
    nodes = tree.body
    for p in parts[:-1]:
        nodes = [n for n in nodes if isinstance(n, ast.ClassDef) and n.name == p][0].body
    fn = [n for n in nodes if isinstance(n, ast.FunctionDef) and n.name == parts[-1]][0]
    is_prop = any(isinstance(d, ast.Name) and d.id == "property" for d in fn.decorator_list)
    return textwrap.dedent(ast.get_source_segment(src, fn, padded=True)), is_prop


def string_constants(path):
    """Module-level NAME = "string" assignments (the column names the functions index batches with)."""
    src, tree = _tree(path)
    out = {}
    for n in tree.body:
        if isinstance(n, ast.Assign) and isinstance(n.value, ast.Constant) and isinstance(n.value.value, str):
            for t in n.targets:
                if isinstance(t, ast.Name):
                    out[t.id] = n.value.value
    return out


class SampleBatch(dict):
    OBS = CUR_OBS = "obs"
    NEXT_OBS = "new_obs"
    ACTIONS = "actions"
    REWARDS = "rewards"
    DONES = "dones"
    INFOS = "infos"
    ACTION_LOGP = "action_logp"
    ACTION_DIST_INPUTS = "action_dist_inputs"
    VF_PREDS = "vf_preds"
    SEQ_LENS = "seq_lens"

    @property
    def count(self):
        return len(self["t"])


class Postprocessing:
    ADVANTAGES = "advantages"
    VALUE_TARGETS = "value_targets"


def discount_cumsum(x, gamma):                      # ray 2.2.0 rllib/evaluation/postprocessing.py
    return scipy.signal.lfilter([1], [1, float(-gamma)], x[::-1], axis=0)[::-1]


def namespace(*paths):
    ns = dict(np=np, torch=torch, SampleBatch=SampleBatch, Postprocessing=Postprocessing, discount_cumsum=discount_cumsum,
              sequence_mask=None, explained_variance=lambda a, b: torch.zeros(()),
              warn_if_infinite_kl_divergence=lambda *a, **k: None, defaultdict=defaultdict)
    for p in paths:
        ns.update(string_constants(p))
    return ns


def load(ns, path, qualname):
    text, is_prop = extract(path, qualname)
    loc = {}
    exec(compile(text, path + ":" + qualname, "exec"), ns, loc)
    fn = loc[qualname.split(".")[-1]]
    return property(fn) if is_prop else fn


# ---- stand-ins for the objects the methods are called on -------------------------------------------------------------
class RefModel(torch.nn.Module):
    """What the reference's loss code needs of a ModelV2: model(batch) -> (logits, state), the value heads, tower_stats,
    and - taken from the reference's own CoPOModel source - the LCF properties and compute_coordinated."""

    def __init__(self, inner, ns, copo):
        super().__init__()
        self.inner = inner
        self.tower_stats = {}
        self._obs = None
        if copo:
            self.lcf_parameters = inner.lcf_parameters
            self.model_config = {"custom_model_config": {ns["USE_DISTRIBUTIONAL_LCF"]: True, "initial_lcf_std": 0.1}}

    def forward(self, batch):
        self._obs = batch["obs"]
        return self.inner.forward(batch["obs"]), []

    def value_function(self):                        # IPPO: rllib's fully connected net, separate value branch
        return self.inner.central_value_function(self._obs)

    def central_value_function(self, cobs):
        return self.inner.central_value_function(cobs)

    def get_nei_value(self, cobs):
        return self.inner.get_nei_value(cobs)

    def get_global_value(self, cobs):
        return self.inner.get_global_value(cobs)

    def is_time_major(self):
        return False


def dist_class(inputs, model):
    return om.DiagGaussian(inputs)


class Harness:
    device = torch.device("cpu")
    dist_class = staticmethod(dist_class)

    def _lazy_tensor_dict(self, b, device=None):
        return b


def flat_grad(model):
    return torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in model.parameters()])


def state_arrays(prefix, model, out):
    for k, v in model.state_dict().items():
        out["%s/%s" % (prefix, k)] = v.detach().numpy().copy()


def make_batch(gen, B, odim, cdim, copo):
    r = lambda *s: torch.randn(*s, generator=gen)
    b = {"obs": torch.rand(B, odim, generator=gen), "actions": 0.5 * r(B, 2), "action_logp": -1.5 + 0.3 * r(B),
         "action_dist_inputs": 0.3 * r(B, 4), "advantages": r(B), "vf_preds": r(B), "value_targets": r(B)}
    b["centralized_critic_obs"] = b["obs"] if cdim == odim else torch.rand(B, cdim, generator=gen)
    if copo:
        for k in ("normalized_advantages", "nei_values", "nei_target", "global_values", "global_target",
                  "nei_advantage", "global_advantages"):
            b[k] = r(B)
    return b


def main():
    out, meta = {}, {}
    A_COPO, A_CC, A_IPPO = (os.path.join(REF, f) for f in ("algo_copo.py", "algo_ccppo.py", "algo_ippo.py"))
    WRAP = os.path.join(REF, "utils", "env_wrappers.py")

    # ---- A. CCEnv distance map / neighbour lists ---------------------------------------------------------------
    ns = namespace(WRAP)
    env = type("Env", (), {})()
    env.distance_map = defaultdict(lambda: defaultdict(lambda: float("inf")))
    rng = np.random.default_rng(7)
    pos = rng.uniform(-60, 60, (14, 2))
    pos[3] = pos[2] + [10.0, 0.0]                    # exact ties at the mean-field radius and equal distances
    pos[5] = pos[2] + [0.0, 10.0]
    pos[6] = pos[2] - [10.0, 0.0]
    names = ["agent%d" % i for i in range(14)]
    veh = {n: type("V", (), {"position": pos[i].copy()})() for i, n in enumerate(names)}
    veh["agent9"] = None                             # a vehicle slot that is None is skipped (:147)
    env.vehicles_including_just_terminated = veh
    load(ns, WRAP, "CCEnv._update_distance_map")(env)
    find = load(ns, WRAP, "CCEnv._find_in_range")
    nb = {}
    for n in names:
        if veh[n] is None:
            continue
        for dist in (40, 10, 0):
            ids, ds = find(env, n, dist)
            nb["%s@%d" % (n, dist)] = [list(ids), [float(x) for x in ds]]
    out["wrappers/pos"] = pos
    meta["wrappers"] = dict(names=names, none=["agent9"], neighbours=nb)

    # ---- A2. CCEnv.step + LCFEnv.step + _add_lcf over a stand-in base env ------------------------------------------
    # (the reference's methods with their own super() chain: TMP(LCFEnv, Base) as get_lcf_env builds it, :463-471)
    from math import cos, sin
    ns = namespace(WRAP)
    draws = np.random.RandomState(123)
    ns.update(cos=cos, sin=sin, clip=lambda a, lo, hi: min(max(a, lo), hi), get_np_random=lambda *a, **k: draws)

    class Base:
        def step(self, actions):
            o, r, d, i, pos = self._script.pop(0)
            self.vehicles_including_just_terminated = {k: type("V", (), {"position": np.asarray(p, np.float64)})()
                                                       for k, p in pos.items()}
            return o, r, d, i

    cc = type("CCEnv", (), {k: load(ns, WRAP, "CCEnv." + k) for k in ("step", "_find_in_range", "_update_distance_map")})
    lcf_cls = type("LCFEnv", (cc,), {k: load(ns, WRAP, "LCFEnv." + k) for k in ("step", "_add_lcf", "enable_copo")})
    ns["CCEnv"], ns["LCFEnv"] = cc, lcf_cls
    tmp = type("TMP", (lcf_cls, Base), {})
    steps_meta = []
    for scen, over in (("angle_native", {}), ("linear_coord", dict(lcf_mode="linear", return_native_reward=False)),
                       ("forced_normal", dict(force_lcf=0.4, return_native_reward=False)),
                       ("uniform", dict(lcf_dist="uniform"))):
        env = tmp()
        env.config = dict(neighbours_distance=40, lcf_mode="angle", lcf_dist="normal", lcf_normal_std=0.1,
                          return_native_reward=True, force_lcf=-100, enable_copo=True, add_traffic_light=False,
                          communication={ns["COMM_METHOD"]: "none"})
        env.config.update(over)
        env.distance_map = defaultdict(lambda: defaultdict(lambda: float("inf")))
        env.lcf_map, env.force_lcf = {}, env.config["force_lcf"]
        lcf_mean = 0.5 if env.config["lcf_mode"] == "linear" else 0.2        # linear mode asserts 0 <= lcf <= 1 (:347)
        env.current_lcf_mean, env.current_lcf_std = lcf_mean, 0.1
        env._last_obs, env._traffic_light_counter = None, 0
        alive = ["agent%d" % k for k in range(6)]
        script, rec = [], []
        for t in range(7):
            if t == 3:
                alive = [a for a in alive if a != "agent2"] + ["agent6"]        # one leaves, a fresh one joins
            pos = {a: rng.uniform(-30, 30, 2) for a in alive}
            o = {a: rng.uniform(0, 1, 5).astype(np.float32) for a in alive}
            r = {a: float(rng.normal()) for a in alive}
            d = {a: False for a in alive}
            script.append((o, dict(r), d, {a: {} for a in alive}, pos))
            rec.append(dict(pos={a: [float(x) for x in pos[a]] for a in alive}, reward=r,
                            obs={a: [float(x) for x in o[a]] for a in alive}))
        env._script = script
        draws.seed(123)
        for t in range(7):
            no, nr, nd, ni = env.step({})
            rec[t]["out_reward"] = {k: float(v) for k, v in nr.items()}
            rec[t]["out_obs_last"] = {k: float(v[-1]) for k, v in no.items()}
            rec[t]["info"] = {k: {q: (list(v[q]) if isinstance(v[q], list) else float(v[q]))
                                   for q in ("neighbours", "neighbours_distance", "nei_rewards", "global_rewards", "lcf",
                                             "coordinated_rewards", "native_rewards")} for k, v in ni.items()}
        steps_meta.append(dict(name=scen, config={k: v for k, v in env.config.items() if k != "communication"},
                               lcf_mean=lcf_mean, lcf_std=0.1, rng_seed=123, steps=rec))
    meta["lcf_env"] = steps_meta

    # ---- B. neighbourhood / global advantages ----------------------------------------------------------------
    ns = namespace(A_COPO)
    T = 60
    ro = SampleBatch({ns["NEI_VALUES"]: rng.normal(size=T).astype(np.float32),
                      ns["NEI_REWARDS"]: rng.normal(size=T).astype(np.float32),
                      ns["GLOBAL_VALUES"]: rng.normal(size=T).astype(np.float32),
                      ns["GLOBAL_REWARDS"]: rng.normal(size=T).astype(np.float32)})
    for k in list(ro):
        out["adv/" + k] = ro[k].copy()
    nei = load(ns, A_COPO, "compute_nei_advantage")
    glo = load(ns, A_COPO, "compute_global_advantage")
    for tag, last_r in (("cut", None), ("done", 0.0)):
        r1 = nei(SampleBatch(ro), float(ro[ns["NEI_VALUES"]][-1]) if last_r is None else last_r, 0.99, 0.95)
        r2 = glo(SampleBatch(ro), float(ro[ns["GLOBAL_VALUES"]][-1]) if last_r is None else last_r, 1.0, 0.95)
        out["adv/%s/nei_advantage" % tag], out["adv/%s/nei_target" % tag] = r1[ns["NEI_ADVANTAGE"]], r1[ns["NEI_TARGET"]]
        out["adv/%s/global_advantages" % tag] = r2[ns["GLOBAL_ADVANTAGES"]]
        out["adv/%s/global_target" % tag] = r2[ns["GLOBAL_TARGET"]]

    # ---- C. centralized-critic observation fusion ---------------------------------------------------------------
    ns = namespace(A_CC, A_COPO)
    odim, adim, nn_ = 6, 2, 4
    spans = {"a0": (0, 12), "a1": (0, 9), "a2": (3, 12), "a3": (5, 11), "a4": (0, 12), "a5": (7, 12)}
    batches = {}
    for n, (t0, t1) in spans.items():
        tt = np.arange(t0, t1)
        infos = []
        for t in tt:
            alive = [m for m, (s0, s1) in spans.items() if m != n and s0 <= t + 1 <= s1]      # the env lists successors too
            order = list(rng.permutation(alive))
            d = np.sort(rng.uniform(1, 25, len(order)))
            if len(d) > 1:
                d[1] = 10.0                                   # a tie with the radius: `>` skips, so it is kept
            infos.append({"neighbours": order, "neighbours_distance": [float(x) for x in d]})
        batches[n] = SampleBatch({"t": tt, "obs": rng.uniform(0, 1, (len(tt), odim)).astype(np.float32),
                                  "actions": rng.uniform(-1, 1, (len(tt), adim)).astype(np.float32), "infos": infos})
    for mode, fn_name in (("concat", "concat_ccppo_process"), ("mf", "mean_field_ccppo_process")):
        fn = load(ns, A_CC, fn_name)
        for cf in (True, False):
            other = odim + (adim if cf else 0)
            cdim = odim + (nn_ if mode == "concat" else 1) * other
            pol = type("P", (), {"config": {"num_neighbours": nn_, ns["COUNTERFACTUAL"]: cf, "mf_nei_distance": 10}})()
            for n, b in batches.items():
                sb = SampleBatch(b)
                cobs = np.zeros((len(b["t"]), cdim), np.float32)
                cobs[:, :odim] = b["obs"]
                sb[ns["CENTRALIZED_CRITIC_OBS"]] = cobs
                others = {m: (None, batches[m]) for m in batches if m != n}
                res = fn(pol, sb, others, odim=odim, adim=adim, other_info_dim=other if mode == "concat" else cdim - odim)
                out["fuse/%s/cf%d/%s" % (mode, int(cf), n)] = res[ns["CENTRALIZED_CRITIC_OBS"]].copy()
    for n, b in batches.items():
        out["fuse/in/%s/t" % n], out["fuse/in/%s/obs" % n], out["fuse/in/%s/actions" % n] = b["t"], b["obs"], b["actions"]
    meta["fuse"] = dict(odim=odim, adim=adim, num_neighbours=nn_, mf_nei_distance=10,
                        infos={n: b["infos"] for n, b in batches.items()})

    # ---- D. the three losses -------------------------------------------------------------------------------------
    cfgs = {"default": dict(clip_param=0.2, vf_clip_param=100.0, vf_loss_coeff=1.0, entropy_coeff=0.0, kl_coeff=0.2,
                            old_value_loss=True),
            "plain_vf": dict(clip_param=0.3, vf_clip_param=1.5, vf_loss_coeff=0.5, entropy_coeff=0.01, kl_coeff=0.0,
                             old_value_loss=False)}
    B, odim = 96, 20
    for algo, path, cls in (("copo", A_COPO, "CoPOPolicy"), ("ccppo", A_CC, "CCPPOPolicy"), ("ippo", A_IPPO, "IPPOPolicy")):
        ns = namespace(A_COPO, A_CC, A_IPPO)
        loss_fn = load(ns, path, cls + ".loss")
        for cname, c in cfgs.items():
            torch.manual_seed({"copo": 1, "ccppo": 2, "ippo": 3}[algo])
            cdim = odim if algo != "ccppo" else 2 * odim + 2
            inner = om.CoPOModel(odim, hiddens=(32, 32), cdim=cdim) if algo == "copo" else om.CCModel(odim, hiddens=(32, 32), cdim=cdim)
            for name, p in inner.named_parameters():          # O(1) heads instead of the 0.01-scaled initialisation
                if p.dim() == 2 and p.shape[0] in (1, 4):
                    p.data.mul_(30.0)
            model = RefModel(inner, ns, algo == "copo")
            if algo == "copo":
                for nm in ("lcf_mean", "lcf_std", "lcf_dist", "compute_coordinated"):
                    setattr(RefModel, nm, load(ns, A_COPO, "CoPOModel." + nm))
            gen = torch.Generator().manual_seed(11)
            batch = make_batch(gen, B, odim, cdim, algo == "copo")
            h = Harness()
            h.config = dict(c, use_critic=True, **{ns["USE_DISTRIBUTIONAL_LCF"]: True})
            h.entropy_coeff, h.kl_coeff, h.model = c["entropy_coeff"], c["kl_coeff"], model
            inner.zero_grad()
            total = loss_fn(h, model, dist_class, batch)
            total.backward()
            key = "loss/%s/%s" % (algo, cname)
            state_arrays(key + "/w", inner, out)
            for k, v in batch.items():
                out["%s/batch/%s" % (key, k)] = v.numpy().copy()
            out[key + "/grad"] = flat_grad(inner).numpy().copy()
            stats = {k: float(v) for k, v in model.tower_stats.items() if torch.is_tensor(v) and v.numel() == 1}
            meta[key] = dict(cfg=c, stats=stats, odim=odim, cdim=cdim, hiddens=[32, 32])
            print(key, {k: round(v, 6) for k, v in stats.items()})

    # ---- E. the meta update --------------------------------------------------------------------------------------
    ns = namespace(A_COPO, A_CC, A_IPPO)
    for nm in ("lcf_mean", "lcf_std", "lcf_dist", "compute_coordinated"):
        setattr(RefModel, nm, load(ns, A_COPO, "CoPOModel." + nm))
    meta_fn = load(ns, A_COPO, "CoPOPolicy.meta_update")
    torch.manual_seed(5)
    inner, target = om.CoPOModel(odim, hiddens=(32, 32)), om.CoPOModel(odim, hiddens=(32, 32))
    for m in (inner, target):
        for name, p in m.named_parameters():
            if p.dim() == 2 and p.shape[0] == 4:
                p.data.mul_(30.0)
    with torch.no_grad():
        inner.lcf_parameters.copy_(torch.tensor([0.3, np.log(0.15)], dtype=torch.float32))
    h = Harness()
    h.config = dict(clip_param=0.2, **{ns["USE_DISTRIBUTIONAL_LCF"]: True, ns["LCF_LR"]: 1e-2})
    h.model, h.target_model = RefModel(inner, ns, True), RefModel(target, ns, True)
    h._raw_lcf_adv_mean, h._raw_lcf_adv_std = 0.05, 1.3
    h._lcf_optimizer = torch.optim.Adam([inner.lcf_parameters], lr=h.config[ns["LCF_LR"]])
    gen = torch.Generator().manual_seed(13)
    batch = make_batch(gen, B, odim, odim, True)
    state_arrays("meta/w", inner, out)
    state_arrays("meta/w_old", target, out)
    for k, v in batch.items():
        out["meta/batch/" + k] = v.numpy().copy()
    torch.manual_seed(99)                              # the draws of lcf_dist.rsample(ego.size()) inside compute_coordinated
    eps = torch.distributions.normal.Normal(torch.tensor(0.0), torch.tensor(1.0)).rsample(batch["advantages"].size())
    out["meta/eps"] = eps.numpy().copy()
    torch.manual_seed(99)
    stats = meta_fn(h, batch)
    meta["meta"] = dict(stats={k: float(v) for k, v in stats.items()}, raw_mean=0.05, raw_std=1.3, lcf_lr=1e-2, clip_param=0.2,
                        odim=odim, hiddens=[32, 32])
    out["meta/lcf_parameters_after"] = inner.lcf_parameters.detach().numpy().copy()
    print("meta", {k: round(float(v), 6) for k, v in stats.items()})

    out["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "ref_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "ref_golden.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
