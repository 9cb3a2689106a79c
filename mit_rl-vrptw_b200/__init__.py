"""mit_rl-vrptw_b200: B200-native (sm_100a) batched rollout of the attention policy of
jpoulletXaccount/MIT_rl-vrptw over its VRP / VRPTW environment.

The directory name contains a hyphen; import it with
``importlib.import_module("mit_rl-vrptw_b200")`` or through the alias module ``mit_rl_vrptw_b200``.
"""
from . import _lib, data, distributed, host, task_specific_params
from ._lib import build_library, VrpB200Error, LIB_PATH
from .engine import Engine, RolloutOutput, config_from_args

__all__ = ["Engine", "RolloutOutput", "config_from_args", "build_library", "VrpB200Error", "LIB_PATH"]
