"""Minimal attribute-dict config (omegaconf is not a dependency of the engine).

`Environment` accepts the same nested `dict` the reference does (sorrel/environment.py:40-46) and
wraps it so `config.world.height`, `config.model.get("num_agents", 2)` and `hasattr(config.model, ...)`
behave like with an OmegaConf DictConfig.
"""
from __future__ import annotations


class Config(dict):
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as exc:
            raise AttributeError(name) from exc

    def __setattr__(self, name, value):
        self[name] = value


def wrap(obj):
    if isinstance(obj, Config):
        return obj
    if isinstance(obj, dict):
        return Config({k: wrap(v) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)):
        return type(obj)(wrap(v) for v in obj)
    return obj


def from_dotlist(items) -> Config:
    out: dict = {}
    for item in items:
        key, _, val = item.partition("=")
        node = out
        parts = key.split(".")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = val
    return wrap(out)
