"""Curriculum schedule (ippo_cl.py:41-78): quarter milestones of the training budget."""
from copo_b200.curriculum import curriculum_num_agents


def test_schedule_follows_the_reference_callback():
    total, target = 2_000_000, 40
    seen, last = [], 0
    for cur in range(0, total + 1, 20_000):
        if cur == 0:
            continue
        n = curriculum_num_agents(last, cur, total, target)
        if n is not None:
            seen.append((cur, n))
        last = cur
    assert seen[0] == (20_000, 10)                       # first call: a quarter of the target
    assert [n for _, n in seen] == [10, 20, 30, 40]
    assert seen[1][0] == 520_000 and seen[2][0] == 1_020_000 and seen[3][0] == 1_520_000
    assert curriculum_num_agents(600_000, 620_000, total, target) is None
    assert curriculum_num_agents(0, 10, total, 30) == 7   # int(30 / 4)
