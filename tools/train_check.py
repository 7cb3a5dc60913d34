"""Multi-rank CoPO training check: parameters stay identical on every rank (same all-reduced gradients, same Adam),
LCF parameters too; prints per-iteration stats from rank 0."""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, ".")
from copo_b200.trainer import CoPOTrainer

def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tr = CoPOTrainer(dict(env="MultiAgentIntersectionEnv", num_scenes=256, rollout_fragment_length=32,
                          sgd_minibatch_size=16384, num_sgd_iter=3, lcf_num_iters=2, env_config={"num_agents": 40}, seed=0))
    for it in range(3):
        res = tr.train()
        if tr.rank == 0:
            st = res["info"]["learner"]["default"]["learner_stats"]
            print("iter %d total_loss %.4f kl %.5f lcf %.5f success %.3f agent_steps %d sample_ms %.1f learn_ms %.1f" % (
                it, st["total_loss"], st["kl"], res["custom_metrics"]["meta_update"]["lcf"], res["custom_metrics"]["success_rate"],
                res["custom_metrics"]["agent_steps"], res["timers"]["sample_time_ms"], res["timers"]["learn_time_ms"]), flush=True)
    if world > 1:
        flat = tr.policy.model.flat
        ref = flat.clone()
        dist.broadcast(ref, 0)
        lcf = tr.policy.model.lcf_parameters.clone()
        lref = lcf.clone()
        dist.broadcast(lref, 0)
        same = torch.tensor([float(torch.equal(ref, flat)), float(torch.equal(lref, lcf))], device=flat.device)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        if tr.rank == 0:
            print("ranks hold identical parameters:", bool(same[0]), "identical LCF:", bool(same[1]), flush=True)
        assert bool(same[0]) and bool(same[1])
        dist.destroy_process_group()
    tr.stop()

if __name__ == "__main__":
    main()
