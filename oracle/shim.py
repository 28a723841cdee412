"""Stub modules that let /root/reference import UNMODIFIED in the build container.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference needs `omegaconf`, `hydra`,
`matplotlib` and `IPython`, none of which are installed here; none of them touches the
step / observation / buffer arithmetic.  Import this module *before* any `sorrel` import.

Only used by `oracle/gen_golden.py` and by tests that are skipped when /root/reference is
absent (i.e. on the GPU box).
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SORREL_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "sorrel"))


class DictConfig(dict):
    """dict with attribute access, enough of omegaconf.DictConfig for the reference."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as exc:  # hasattr() relies on AttributeError
            raise AttributeError(name) from exc

    def __setattr__(self, name, value):
        self[name] = value


def _wrap(obj):
    if isinstance(obj, dict):
        return DictConfig({k: _wrap(v) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)):
        return type(obj)(_wrap(v) for v in obj)
    return obj


class OmegaConf:
    @staticmethod
    def create(obj=None):
        return _wrap(obj if obj is not None else {})

    @staticmethod
    def from_dotlist(items):
        out: dict = {}
        for item in items:
            key, _, val = item.partition("=")
            node = out
            parts = key.split(".")
            for p in parts[:-1]:
                node = node.setdefault(p, {})
            node[parts[-1]] = val
        return _wrap(out)

    @staticmethod
    def to_container(cfg, **_kw):
        return dict(cfg)


def install() -> None:
    """Register the stubs and put the reference on sys.path (idempotent)."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")

    def mod(name, **attrs):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            sys.modules[name] = m
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    mod("omegaconf", DictConfig=DictConfig, OmegaConf=OmegaConf)

    def hydra_main(*_a, **_kw):
        return lambda fn: fn

    mod("hydra", main=hydra_main)
    ipy = mod("IPython")
    disp = mod("IPython.display", clear_output=lambda *a, **k: None, display=lambda *a, **k: None)
    ipy.display = disp
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        mpl = mod("matplotlib")
        plt = mod("matplotlib.pyplot")
        mpl.pyplot = plt
        mod("matplotlib.animation")
        mod("matplotlib.colors")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
