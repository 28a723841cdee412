"""Mirrors sorrel/examples/iowa/main.py: same config dict, world/env classes swapped.

    python -m sorrel_b200.examples.iowa.main [num_envs]
"""
import sys

from sorrel_b200.examples.iowa.entities import EmptyEntity
from sorrel_b200.examples.iowa.env import GamblingEnv
from sorrel_b200.examples.iowa.world import GamblingWorld

if __name__ == "__main__":
    config = {
        "experiment": {"epochs": 5, "max_turns": 100, "record_period": 50},
        "model": {"agent_vision_radius": 2, "epsilon_decay": 0.0005,
                  "parameters": {"layer_size": 250, "epsilon": 0.5, "device": "cpu", "n_frames": 5, "memory_size": 128}},
        "world": {"height": 20, "width": 20, "spawn_prob": 0.01},
    }
    num_envs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    world = GamblingWorld(config=config, default_entity=EmptyEntity(), num_envs=num_envs)
    env = GamblingEnv(world, config)
    env.run_experiment()
