"""Observation specs with the reference's interface (sorrel/observation/observation_spec.py).

`OneHotObservationSpec` keeps the constructor, `entity_map`, `input_size`, `override_*` and
`observe()` of the reference (observation_spec.py:116-205); the crop itself is computed by the CUDA
kernels (closed form of sorrel/observation/visual_field.py:9-101) for every instance at once.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np


def one_hot_encode(value: int, num_classes: int) -> np.ndarray:
    """sorrel/utils/helpers.py:130-150"""
    assert value <= (num_classes - 1), f"The maximum value of `value` is {num_classes - 1}."
    one_hot = np.zeros(num_classes)
    one_hot[value] = 1
    return one_hot


class ObservationSpec:
    #: transform the kernels apply to the per-cell sum over layers (_abi.SB_OBS_*)
    obs_transform = 0
    entity_map: dict
    vision_radius: int
    full_view: bool
    input_size: Sequence[int]
    fill_entity_kind: str

    def __init__(self, entity_list: list[str], full_view: bool, vision_radius: int | None = None,
                 env_dims: Sequence[int] | None = None, fill_entity_kind: str = "Wall"):
        if full_view and not isinstance(env_dims, Sequence):  # observation_spec.py:49-52
            raise TypeError("env_dims must be provided when full_view is true.")
        elif not full_view and not isinstance(vision_radius, int):
            raise TypeError("vision_radius must be provided when full_view is false.")
        self.full_view = full_view
        self.vision_radius = vision_radius if vision_radius else 0
        self.entity_list = list(entity_list)
        self.entity_map = self.generate_map(entity_list)
        self.input_size = (1,)
        self.fill_entity_kind = fill_entity_kind

    def generate_map(self, entity_list: list[str]) -> dict:
        raise NotImplementedError

    def override_entity_map(self, entity_map: dict) -> None:
        self.entity_map = entity_map

    def override_input_size(self, input_size: Sequence[int]) -> None:
        self.input_size = input_size

    @property
    def num_channels(self) -> int:
        return len(next(iter(self.entity_map.values())))

    def appearance_row(self, kind: str) -> np.ndarray:
        """entity_map[kind]; KeyError for an unknown kind like visual_field.py:49."""
        return np.asarray(self.entity_map[kind], dtype=np.float64)


class OneHotObservationSpec(ObservationSpec):
    def __init__(self, entity_list: list[str], full_view: bool, vision_radius: int | None = None,
                 env_dims: Sequence[int] | None = None, fill_entity_kind: str = "Wall"):
        super().__init__(entity_list, full_view, vision_radius, env_dims, fill_entity_kind)
        if self.full_view:
            raise NotImplementedError(
                "full_view observations are not implemented by the B200 engine yet (egocentric crops only)")
        self.input_size = (len(entity_list), 2 * self.vision_radius + 1, 2 * self.vision_radius + 1)

    def generate_map(self, entity_list: list[str]) -> dict[str, np.ndarray]:
        """observation_spec.py:150-173: one-hot of the list index, all-zero for "EmptyEntity"."""
        entity_map: dict[str, np.ndarray] = {}
        num_classes = len(entity_list)
        for i, x in enumerate(entity_list):
            if x == "EmptyEntity":
                entity_map[x] = np.zeros(num_classes)
            else:
                entity_map[x] = one_hot_encode(value=i, num_classes=num_classes)
        return entity_map

    def observe(self, world, location=None):
        """observation_spec.py:175-205 for every instance of a batched `Gridworld`.

        location None  -> float32 device tensor [N, A, C, S, S]: the view of every agent.
        location given -> (y, x[, z]) tuple or an int tensor [N, 2]: the view from that cell in
                          every instance, float32 [N, C, S, S]."""
        if location is None:
            return world.engine.observe()
        return world.observe_from(location)


class RGBObservationSpec(ObservationSpec):
    """observation_spec.py:386-483: RGB appearance per kind (EmptyEntity black, the others spread over
    the hue circle), layers summed in 0..255 space, `clip(sum, 0, 255) / 255` (kernel transform SB_OBS_RGB)."""

    obs_transform = 1

    def __init__(self, entity_list: list[str], full_view: bool, vision_radius: int | None = None,
                 env_dims: Sequence[int] | None = None, fill_entity_kind: str = "Wall"):
        super().__init__(entity_list, full_view, vision_radius, env_dims, fill_entity_kind)
        if self.full_view:
            raise NotImplementedError(
                "full_view observations are not implemented by the B200 engine yet (egocentric crops only)")
        self.input_size = (3, 2 * self.vision_radius + 1, 2 * self.vision_radius + 1)

    def generate_map(self, entity_list: list[str]) -> dict[str, np.ndarray]:
        """observation_spec.py:420-448"""
        import colorsys

        entity_map: dict[str, np.ndarray] = {}
        num_colors = len([e for e in entity_list if e != "EmptyEntity"])
        color_idx = 0
        for x in entity_list:
            if x == "EmptyEntity":
                entity_map[x] = np.zeros(3, dtype=np.uint8)
            else:
                rgb = colorsys.hsv_to_rgb(color_idx / max(1, num_colors), 1.0, 1.0)
                entity_map[x] = np.array([int(c * 255) for c in rgb], dtype=np.uint8)
                color_idx += 1
        return entity_map

    def observe(self, world, location=None):
        """observation_spec.py:450-483 for every instance (see OneHotObservationSpec.observe)."""
        if location is None:
            return world.engine.observe()
        return world.observe_from(location)


def positional_embedding(location, world, scale: tuple[int, int]) -> np.ndarray:
    """sorrel/observation/embedding.py:8-44: sin/cos code of a (row, column) location at `scale[0]` octaves
    per axis (the reference uses scale[0] for both loops).  float64, like the reference."""
    x, y = location[0:2]
    grid_size = (world.height, world.width)
    embedding = []
    for i in range(scale[0]):
        freq_x = 2 * np.pi * (2**i) / grid_size[0]
        embedding.append(np.sin(freq_x * x))
        embedding.append(np.cos(freq_x * x))
    for j in range(scale[0]):
        freq_y = 2 * np.pi * (2**j) / grid_size[1]
        embedding.append(np.sin(freq_y * y))
        embedding.append(np.cos(freq_y * y))
    return np.array(embedding)
