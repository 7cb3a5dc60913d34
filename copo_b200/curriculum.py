"""Curriculum on the number of agents: the reference's `ChangeNCallback`
(copo/algo_ippo/ippo_cl.py:41-78) grows the population to 1/4, 2/4, 3/4 and 4/4 of the target when training crosses
the quarter marks of `total_time_step`, by re-creating every env (`close_and_reset_num_agents`,
torch_copo/utils/env_wrappers.py:444-454).  On the batched simulator a population change is slot masking
(`b2c_env_set_num_agents`): slots beyond the current population stay disabled after the next reset."""


def curriculum_num_agents(last_steps, current_steps, total_time_step, target_num_agents):
    """The population the callback would set at this point, or None for "no change" (same branch order as
    ippo_cl.py:58-71)."""
    q = total_time_step / 4
    new = None
    if last_steps <= q * 1 < current_steps:
        new = int(target_num_agents / 4 * 2)
    elif last_steps <= q * 2 < current_steps:
        new = int(target_num_agents / 4 * 3)
    elif last_steps <= q * 3 < current_steps:
        new = int(target_num_agents / 4 * 4)
    if current_steps <= q * 1 and last_steps == 0:
        new = int(target_num_agents / 4 * 1)
    return new


class ChangeNCallback:
    """Attach to a trainer: `cb = ChangeNCallback(total, target); cb.on_train_result(trainer, result)` after every
    `trainer.train()`.  Changes take effect through `BatchedDrivingEnv.set_num_agents` + a reset of the scenes."""

    def __init__(self, total_time_step, target_num_agents):
        self.total_time_step, self.target_num_agents = total_time_step, target_num_agents
        self.last_steps = 0
        self.history = []

    def on_train_result(self, trainer, result):
        current = result["timesteps_total"]
        n = curriculum_num_agents(self.last_steps, current, self.total_time_step, self.target_num_agents)
        if n is not None:
            n = max(1, min(n, trainer.env.A))
            trainer.env.set_num_agents(n)
            trainer.reset_scenes()
            self.history.append((current, n))
        self.last_steps = current
        return n
