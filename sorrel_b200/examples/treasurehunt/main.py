"""Mirrors sorrel/examples/treasurehunt/main.py: same config dict, world/env classes swapped.

    python -m sorrel_b200.examples.treasurehunt.main [num_envs]
"""
import sys

from sorrel_b200.examples.treasurehunt.entities import EmptyEntity
from sorrel_b200.examples.treasurehunt.env import TreasurehuntEnv
from sorrel_b200.examples.treasurehunt.world import TreasurehuntWorld

if __name__ == "__main__":
    config = {
        "experiment": {"epochs": 5, "max_turns": 100, "record_period": 50},
        "model": {"agent_vision_radius": 2, "epsilon_decay": 0.0005, "observation_spec": "onehot",
                  "num_agents": 1, "memory_size": 128, "n_frames": 5},
        "world": {"height": 21, "width": 21, "gem_value": 10, "food_value": 5, "bone_value": -10,
                  "spawn_prob": 0.005},
    }
    num_envs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    world = TreasurehuntWorld(config=config, default_entity=EmptyEntity(), num_envs=num_envs)
    env = TreasurehuntEnv(world, config)
    env.run_experiment()
