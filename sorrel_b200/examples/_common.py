"""Shared plumbing of the example packages (nothing here exists in the reference)."""
from __future__ import annotations

import sys

from sorrel_b200 import config as _config


def batch_args(num_envs, device, seed, env_id_base) -> dict:
    """The four constructor arguments a batched world adds to the reference signature."""
    return {"num_envs": num_envs, "device": device, "seed": seed, "env_id_base": env_id_base}


def section(config, name: str):
    """`config.<name>` of a plain dict or an attribute config."""
    return getattr(_config.wrap(config), name)


def run_cli(make_world, make_env, config: dict, default_envs: int) -> None:
    """`python -m sorrel_b200.examples.<name>.main [num_envs]`: build, run the experiment, print per-epoch rewards."""
    num_envs = int(sys.argv[1]) if len(sys.argv) > 1 else default_envs
    world = make_world(config, num_envs)
    make_env(world, config).run_experiment()
