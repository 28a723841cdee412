"""Entry point with the config of sorrel/examples/iowa/main.py (a short run):

    python -m sorrel_b200.examples.iowa.main [num_envs]
"""
from sorrel_b200.examples._common import run_cli
from sorrel_b200.examples.iowa.entities import EmptyEntity
from sorrel_b200.examples.iowa.env import GamblingEnv
from sorrel_b200.examples.iowa.world import GamblingWorld

CONFIG = {
    "world": dict(height=20, width=20, spawn_prob=0.01),
    "model": dict(agent_vision_radius=2, epsilon_decay=0.0005,
                  parameters=dict(layer_size=250, epsilon=0.5, device="cpu", n_frames=5, memory_size=128)),
    "experiment": dict(epochs=5, max_turns=100, record_period=50),
}

if __name__ == "__main__":
    run_cli(lambda cfg, n: GamblingWorld(cfg, EmptyEntity(), num_envs=n), GamblingEnv, CONFIG, 4096)
