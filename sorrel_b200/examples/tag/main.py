"""Mirrors sorrel/examples/tag/main.py: same config dict, base Gridworld, env class swapped.

    python -m sorrel_b200.examples.tag.main [num_envs]
"""
import sys

from sorrel_b200.entities import EmptyEntity
from sorrel_b200.examples.tag.env import TagEnv
from sorrel_b200.worlds import Gridworld

if __name__ == "__main__":
    config = {
        "experiment": {"epochs": 5, "max_turns": 20, "record_period": 50},
        "model": {"epsilon_decay": 0.0001, "memory_size": 64, "n_frames": 5},
        "agent": {"num_agents": 5, "vision_radius": 4},
        "world": {"height": 11, "width": 11, "layers": 1},
    }
    num_envs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    world = Gridworld(**config["world"], default_entity=EmptyEntity(), num_envs=num_envs)
    env = TagEnv(world, config)
    env.run_experiment()
