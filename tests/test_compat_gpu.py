"""The import shim on the GPU: the calls the reference's `torch_copo/train_copo.py` makes (its imports, `get_lcf_env` +
`get_rllib_compatible_env`, a config with `tune.grid_search` leaves, `train(CoPOTrainer, ...)` with the driving
callbacks) run one real training iteration per trial on the CUDA path and leave Tune's trial directories behind.  Where
the reference tree is mounted (the build container) the reference's own script is executed by path, unchanged."""
import glob
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, "copo_b200", "compat")
REF = "/root/reference/copo_code/copo/torch_copo"

DRIVER = """
import json
from metadrive.envs.marl_envs import MultiAgentIntersectionEnv, MultiAgentRoundaboutEnv
from ray import tune
from copo.torch_copo.algo_copo import CoPOTrainer, USE_CENTRALIZED_CRITIC, USE_DISTRIBUTIONAL_LCF
from copo.torch_copo.utils.callbacks import MultiAgentDrivingCallbacks
from copo.torch_copo.utils.env_wrappers import get_lcf_env, get_rllib_compatible_env
from copo.torch_copo.utils.train import train
from copo.torch_copo.utils.utils import get_train_parser

args = get_train_parser().parse_args(["--exp-name", "shim", "--test"])
config = dict(
    env=tune.grid_search([get_rllib_compatible_env(get_lcf_env(c)) for c in (MultiAgentIntersectionEnv, MultiAgentRoundaboutEnv)]),
    env_config=dict(start_seed=tune.grid_search([5000]), neighbours_distance=40),
    num_gpus=0.5, counterfactual=True, fuse_mode="none", mf_nei_distance=10,
    train_batch_size=2000, rollout_fragment_length=25, sgd_minibatch_size=512, num_sgd_iter=2, lcf_num_iters=2,
    **{USE_CENTRALIZED_CRITIC: False, USE_DISTRIBUTIONAL_LCF: True},
)
an = train(CoPOTrainer, exp_name=args.exp_name, stop=int(1e6), config=config, num_seeds=1, test_mode=args.test,
           custom_callback=MultiAgentDrivingCallbacks, num_gpus=args.num_gpus if hasattr(args, "num_gpus") else 0)
out = [dict(env=t.config["env"], status=t.status, iters=t.last_result.get("training_iteration"),
            steps=t.last_result.get("timesteps_total"), success=t.last_result.get("success"),
            lcf=t.last_result.get("custom_metrics", {}).get("meta_update", {}).get("lcf")) for t in an.trials]
print("SHIM_RESULT " + json.dumps(out))
"""


def _env(extra=None):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, COMPAT]), B2C_NUM_SCENES="8", B2C_COMPAT_MAX_ITERS="1")
    env.pop("B2C_COMPAT_DRY_RUN", None)
    env.update(extra or {})
    return env


def test_reference_call_sequence_trains_on_the_gpu(tmp_path):
    r = subprocess.run([sys.executable, "-c", DRIVER], env=_env(), capture_output=True, text=True, timeout=600,
                       cwd=str(tmp_path))
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    line = [l for l in r.stdout.splitlines() if l.startswith("SHIM_RESULT ")][-1]
    out = json.loads(line[len("SHIM_RESULT "):])
    assert [o["env"] for o in out] == ["LCFMultiAgentIntersectionEnv", "LCFMultiAgentRoundaboutEnv"]
    for o in out:
        assert o["status"] == "TERMINATED" and o["iters"] == 1 and o["steps"] > 0
        assert o["lcf"] is not None and -1.0 <= o["lcf"] <= 1.0
    # Tune's layout: one directory per trial with params.json / progress.csv / result.json, the pickled progress beside
    trials = glob.glob(os.path.join(str(tmp_path), "shim", "*"))
    assert len([d for d in trials if os.path.isdir(d)]) == 2
    for d in trials:
        if os.path.isdir(d):
            assert {"params.json", "progress.csv", "result.json"} <= set(os.listdir(d))
    assert glob.glob(os.path.join(str(tmp_path), "shim-*.pkl"))


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is only mounted in the build container")
def test_reference_script_by_path(tmp_path):
    r = subprocess.run([sys.executable, os.path.join(REF, "train_copo.py"), "--exp-name", "t", "--test"],
                       env=_env(), capture_output=True, text=True, timeout=900, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-3000:]
