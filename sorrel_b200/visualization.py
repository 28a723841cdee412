"""Readback hook for rendering (sorrel/utils/visualization.py:227-273): `ImageRenderer` with the reference's
interface, fed from ONE device instance.

The reference composes PNG sprites per entity; assets are not part of this engine, so every kind is painted as a
flat tile (hues spread like RGBObservationSpec's, EmptyEntity black), layers drawn bottom-up with empty kinds
transparent.  Host-side tooling: it reads back a single instance (`Engine.grid_instance`), never the whole batch.
"""
from __future__ import annotations

import colorsys
import os
from pathlib import Path

import numpy as np


def kind_palette(kind_names: list[str], empty: set[int]) -> np.ndarray:
    """uint8 [K, 3]: black for the empty kinds, evenly spread hues for the others."""
    vis = [k for k in range(len(kind_names)) if k not in empty]
    pal = np.zeros((len(kind_names), 3), dtype=np.uint8)
    for i, k in enumerate(vis):
        pal[k] = [int(c * 255) for c in colorsys.hsv_to_rgb(i / max(1, len(vis)), 0.85, 1.0)]
    return pal


def render_instance(world, env: int = 0, tile: int = 16) -> np.ndarray:
    """uint8 [H*tile, W*tile, 3] picture of device instance `env` (all layers, bottom-up)."""
    eng = world.engine
    kinds = eng.grid_instance(env)  # [H, W, L]
    spec = eng.spec
    empty = {k for k in range(len(spec.kind_names)) if not np.any(spec.appearance[k])}
    pal = kind_palette(spec.kind_names, empty)
    H, W, L = kinds.shape
    img = np.zeros((H, W, 3), dtype=np.uint8)
    for z in range(L):
        layer = kinds[:, :, z]
        show = ~np.isin(layer, list(empty))
        img[show] = pal[layer[show]]
    return np.kron(img, np.ones((tile, tile, 1), dtype=np.uint8))


class ImageRenderer:
    """visualization.py:227-273: `add_image(world)` per turn, `save_gif(epoch, folder)` per recorded epoch."""

    def __init__(self, experiment_name: str, record_period: int, num_turns: int, env: int = 0):
        self.experiment_name = experiment_name
        self.record_period = record_period
        self.num_turns = num_turns
        self.env = env
        self.frames: list = []

    def clear(self):
        del self.frames[:]

    def add_image(self, world) -> None:
        from PIL import Image

        self.frames.append(Image.fromarray(render_instance(world, self.env)))

    def save_gif(self, epoch: int, folder: os.PathLike) -> Path:
        Path(folder).mkdir(parents=True, exist_ok=True)
        path = Path(folder) / f"{self.experiment_name}_epoch{epoch}.gif"
        self.frames[0].save(path, format="GIF", append_images=self.frames[1:], save_all=True, duration=100,
                            disposal=2, loop=0)
        self.clear()
        return path
