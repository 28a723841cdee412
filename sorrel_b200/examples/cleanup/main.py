"""Mirrors sorrel/examples/cleanup/main.py with configs/config.yaml inlined: world/env classes swapped.

    python -m sorrel_b200.examples.cleanup.main [num_envs]
"""
from sorrel_b200.examples._common import run_cli
from sorrel_b200.examples.cleanup.entities import EmptyEntity
from sorrel_b200.examples.cleanup.env import CleanupEnv
from sorrel_b200.examples.cleanup.world import CleanupWorld

CONFIG = {
    "experiment": {"name": "cleanup", "epochs": 5, "max_turns": 100, "epsilon_decay": 0.0001, "record_period": 50},
    "env": {"height": 21, "width": 31, "layers": 3, "default_object": "EmptyEntity", "full_mdp": False,
            "pollution_threshold": 0.5, "initial_apples": 20, "apple_spawn_chance": 0.002,
            "pollution_spawn_chance": 0.009, "mode": "DEFAULT"},
    "model": {"iqn": {"type": "iRainbowModel", "parameters": {"memory_size": 128}}, "epsilon_decay": 0.0001},
    "agent": {"agent": {"num": 10, "model": "iqn", "beam_radius": 3, "rotation": False,
                        "obs": {"channels": 8, "vision": 5, "n_frames": 5, "embeddings": 3}}},
}

if __name__ == "__main__":
    run_cli(lambda cfg, n: CleanupWorld(cfg, EmptyEntity(), num_envs=n), CleanupEnv, CONFIG, 2048)
