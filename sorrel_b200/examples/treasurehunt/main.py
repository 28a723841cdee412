"""Entry point with the config of sorrel/examples/treasurehunt/main.py (one-hot observations, a short run):

    python -m sorrel_b200.examples.treasurehunt.main [num_envs]
"""
from sorrel_b200.examples._common import run_cli
from sorrel_b200.examples.treasurehunt.entities import EmptyEntity
from sorrel_b200.examples.treasurehunt.env import TreasurehuntEnv
from sorrel_b200.examples.treasurehunt.world import TreasurehuntWorld

WORLD = dict(height=21, width=21, spawn_prob=0.005, gem_value=10, food_value=5, bone_value=-10)
MODEL = dict(agent_vision_radius=2, observation_spec="onehot", num_agents=1, epsilon_decay=0.0005,
             memory_size=128, n_frames=5)
CONFIG = {"world": WORLD, "model": MODEL, "experiment": dict(epochs=5, max_turns=100, record_period=50)}

if __name__ == "__main__":
    run_cli(lambda cfg, n: TreasurehuntWorld(cfg, EmptyEntity(), num_envs=n), TreasurehuntEnv, CONFIG, 4096)
