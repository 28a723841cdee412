"""Entry point with the config of sorrel/examples/tag/main.py (base Gridworld, a short run):

    python -m sorrel_b200.examples.tag.main [num_envs]
"""
from sorrel_b200.entities import EmptyEntity
from sorrel_b200.examples._common import run_cli
from sorrel_b200.examples.tag.env import TagEnv
from sorrel_b200.worlds import Gridworld

CONFIG = {
    "world": dict(height=11, width=11, layers=1),
    "agent": dict(num_agents=5, vision_radius=4),
    "model": dict(epsilon_decay=0.0001, memory_size=64, n_frames=5),
    "experiment": dict(epochs=5, max_turns=20, record_period=50),
}

if __name__ == "__main__":
    run_cli(lambda cfg, n: Gridworld(**cfg["world"], default_entity=EmptyEntity(), num_envs=n), TagEnv, CONFIG, 4096)
