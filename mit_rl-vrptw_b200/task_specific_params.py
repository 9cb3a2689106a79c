"""Task sizes (reference: task_specific_params.py:22-115) + the VRPTW-100 extrapolation.

``vrptw100`` is NOT defined by the reference (its table stops at vrptw50): capacity is vrp100's,
decode_len = 1.8 * n_cust like vrptw50.  Every number reported on it says so.
"""
from collections import namedtuple

TaskVRP = namedtuple("TaskVRP", ["task_name", "input_dim", "n_nodes", "n_cust", "decode_len", "capacity", "demand_max"])


class TaskTable(dict):
    pass


task_lst = TaskTable({
    "vrp10": TaskVRP("vrp", 3, 11, 10, 20, 20, 9),
    "vrp10_rl": TaskVRP("vrp", 3, 11, 10, 20, 30, 9),
    "vrp20": TaskVRP("vrp", 3, 21, 20, 30, 30, 9),
    "vrp20_rl": TaskVRP("vrp", 3, 21, 20, 45, 30, 9),
    "vrp50": TaskVRP("vrp", 3, 51, 50, 70, 40, 9),
    "vrp100": TaskVRP("vrp", 3, 101, 100, 140, 50, 9),
    "vrptw10": TaskVRP("vrptw", 5, 11, 10, 20, 20, 9),
    "vrptw20": TaskVRP("vrptw", 5, 21, 20, 40, 25, 9),
    "vrptw50": TaskVRP("vrptw", 5, 51, 50, 90, 40, 9),
    "vrptw100": TaskVRP("vrptw", 5, 101, 100, 180, 50, 9),   # extrapolated, see module docstring
})


def default_args(task: str, **over) -> dict:
    """The flags of configs.py:28-111 that the rollout path reads, merged with the task tuple
    (configs.py:9-26), with embedding_graph=0 (LinearEmbedding)."""
    t = task_lst[task]
    a = dict(task=task, embedding_dim=30, hidden_dim=128, rnn_layers=1, n_glimpses=0, use_tanh=False,
             tanh_exploration=10.0, mask_glimpses=True, mask_pointer=True, forget_bias=1.0, beam_width=5,
             embedding_graph=0, min_trucks=False, ups=False, random_seed=24601, test_size=1000, batch_size=128,
             dropout=0.1, decode_len=t.decode_len, n_process_blocks=3, actor_net_lr=1e-4, critic_net_lr=1e-4, max_grad_norm=2.0)
    a.update(t._asdict())
    a.update(over)
    return a
