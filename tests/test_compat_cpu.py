"""The import shim (copo_b200/compat) on the CPU: every name the reference's training scripts import resolves, the
reference's own scripts run unchanged through it up to the launch of the first trial (dry run - the trainers need the
GPU), `tune.grid_search` expands like Tune's grid, `train()` builds the trial configs the reference would."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, "copo_b200", "compat")
REF = "/root/reference/copo_code/copo/torch_copo"

# what torch_copo/train_{copo,ippo,ccppo}.py import (module -> names); the scripts themselves are not copied
SCRIPT_IMPORTS = {
    "metadrive.envs.marl_envs": ["MultiAgentParkingLotEnv", "MultiAgentRoundaboutEnv", "MultiAgentBottleneckEnv",
                                 "MultiAgentMetaDrive", "MultiAgentTollgateEnv", "MultiAgentIntersectionEnv"],
    "ray": ["tune"],
    "ray.tune": ["grid_search", "run", "CLIReporter"],
    "copo.torch_copo.algo_copo": ["CoPOTrainer", "USE_CENTRALIZED_CRITIC", "USE_DISTRIBUTIONAL_LCF", "COUNTERFACTUAL",
                                  "CoPOConfig", "CoPOPolicy", "CoPOModel"],
    "copo.torch_copo.algo_ippo": ["IPPOTrainer", "IPPOConfig", "IPPOPolicy"],
    "copo.torch_copo.algo_ccppo": ["CCPPOTrainer", "get_ccppo_env", "CCPPOConfig", "CCPPOPolicy", "CCModel",
                                   "get_centralized_critic_obs_dim"],
    "copo.torch_copo.utils.callbacks": ["MultiAgentDrivingCallbacks"],
    "copo.torch_copo.utils.env_wrappers": ["get_lcf_env", "get_rllib_compatible_env", "get_ccenv", "get_change_n_env"],
    "copo.torch_copo.utils.train": ["train"],
    "copo.torch_copo.utils.utils": ["get_train_parser"],
    "copo.train.utils": ["initialize_ray", "get_train_parser"],
}


def _run(code_or_path, *args, env_extra=None, is_path=False, cwd=None):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, COMPAT]), B2C_COMPAT_DRY_RUN="1")
    env.update(env_extra or {})
    cmd = [sys.executable] + ([code_or_path] if is_path else ["-c", code_or_path]) + list(args)
    return subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300, cwd=cwd)


def test_every_name_the_scripts_import_resolves():
    code = "import importlib, json, sys\nspec = json.loads(sys.argv[1])\n" \
           "for mod, names in spec.items():\n" \
           "    m = importlib.import_module(mod)\n" \
           "    assert 'compat' in m.__file__, (mod, m.__file__)\n" \
           "    for n in names:\n        assert hasattr(m, n), (mod, n)\nprint('ok')\n"
    import json
    r = _run(code, json.dumps(SCRIPT_IMPORTS))
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]


def test_grid_search_and_trial_configs(tmp_path):
    code = """
import json
from ray import tune
from copo.torch_copo.algo_copo import CoPOTrainer, USE_CENTRALIZED_CRITIC
from copo.torch_copo.utils.callbacks import MultiAgentDrivingCallbacks
from copo.torch_copo.utils.env_wrappers import get_lcf_env, get_rllib_compatible_env
from copo.torch_copo.utils.train import train
from metadrive.envs.marl_envs import MultiAgentIntersectionEnv, MultiAgentRoundaboutEnv
envs = [get_rllib_compatible_env(get_lcf_env(c)) for c in (MultiAgentIntersectionEnv, MultiAgentRoundaboutEnv)]
assert envs == ["LCFMultiAgentIntersectionEnv", "LCFMultiAgentRoundaboutEnv"]
name, cls = get_rllib_compatible_env(get_lcf_env(MultiAgentIntersectionEnv), return_class=True)
assert name == envs[0] and cls.__name__ == name
cfg = dict(env=tune.grid_search(envs), env_config=dict(neighbours_distance=40, start_seed=tune.grid_search([5000, 6000])),
           num_gpus=0, **{USE_CENTRALIZED_CRITIC: tune.grid_search([False])})
an = train(CoPOTrainer, exp_name="t", stop=1000, config=cfg, num_seeds=3, custom_callback=MultiAgentDrivingCallbacks)
out = [(t.config["env"], t.config["env_config"]["start_seed"], t.config["seed"], t.status) for t in an.trials]
print(json.dumps(out))
"""
    r = _run(code, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert len(out) == 2 * 2 * 3 and all(s == "DRY_RUN" for *_, s in out)
    assert sorted(set(o[2] for o in out)) == [0, 100, 200]                 # seeds i * 100 + start_seed (train.py:72)
    assert sorted(set(o[1] for o in out)) == [5000, 6000]
    assert not os.listdir(str(tmp_path))                                  # a dry run leaves nothing behind


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is only mounted in the build container")
@pytest.mark.parametrize("script", ["train_copo.py", "train_ippo.py", "train_ccppo.py"])
def test_reference_scripts_run_unchanged_up_to_the_first_trial(script, tmp_path):
    r = _run(os.path.join(REF, script), "--exp-name", "t", "--test", is_path=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Successfully initialize Ray" in r.stdout
