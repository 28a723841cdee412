"""`Gridworld` with the reference's interface (sorrel/worlds/gridworld.py:10-196), batched.

The reference world is one `numpy.ndarray(dtype=object)` of `Entity` instances.  Here the object
array `map` is only the *template* the user edits with `add()/remove()/move()` while the world is
being described (`Environment.populate_environment`); `compile()` turns it into a kind table plus a
uint8 template and creates `num_envs` device-resident instances stepped by the CUDA kernels.
"""
from __future__ import annotations

import copy

import numpy as np
import torch

from . import _abi
from .engine import MOVE_DELTAS, Engine, WorldSpec
from .entities import Entity
from .rules import (MOVE_FACING, BecomeIfLayer, BernoulliBecome, BernoulliSpawn, ExpireAfter, SpawnBeam,
                    ValueRedraw, beam_pattern)


class Gridworld:
    def __init__(self, height: int, width: int, layers: int, default_entity: Entity, num_envs: int = 1,
                 device: torch.device | str = "cuda", seed: int = 0, env_id_base: int = 0):
        self.height = height
        self.width = width
        self.layers = layers
        self.default_entity = default_entity
        self.num_envs = int(num_envs)
        self.device = torch.device(device)
        self.seed = int(seed)
        self.env_id_base = int(env_id_base)
        self.engine: Engine | None = None
        self.tuning: dict | None = None  # optional launch tuning handed to the Engine (see Engine.__init__)
        self.spec: WorldSpec | None = None
        self._kinds: list[Entity] = []
        self.create_world()
        self.turn = 0
        self.max_turns = 0
        self.is_done = False

    # ------------------------------------------------------------------------------------------
    # template editing (same calls as the reference)
    # ------------------------------------------------------------------------------------------
    def create_world(self) -> None:
        """gridworld.py:47-65.  The template shares ONE default-entity prototype per cell kind
        instead of H*W*L deep copies (the copies only exist to give every Python object its own
        location, which kind ids do not need)."""
        self.map = np.full((self.height, self.width, self.layers), self.default_entity, dtype=object)
        self._template_total_reward = 0.0

    def add(self, target_location: tuple[int, ...], entity: Entity) -> None:
        """gridworld.py:67-76"""
        entity.location = target_location
        self.map[target_location] = entity

    def remove(self, target_location: tuple[int, ...]) -> Entity:
        """gridworld.py:78-93"""
        entity = self.map[target_location]
        self.map[target_location] = self.default_entity
        return entity

    def move(self, entity: Entity, new_location: tuple[int, ...]) -> bool:
        """gridworld.py:95-122 on the template."""
        if self.map[new_location].passable:
            previous_location = entity.location
            entity.location = new_location
            self.map[new_location] = entity
            self.map[previous_location] = self.default_entity
            return True
        return False

    def observe(self, target_location: tuple[int, ...], env: int | None = None) -> Entity:
        """gridworld.py:124-133.  With `env`, reads the cell of that device instance back and
        returns the prototype entity of its kind."""
        if env is None or self.engine is None:
            return self.map[target_location]
        y, x, z = target_location
        kinds = self.engine.grid_numpy()[env]
        return self._kinds[int(kinds[y, x, z])]

    def observe_all_layers(self, target_location: tuple[int, ...], env: int | None = None) -> list[Entity]:
        """gridworld.py:135-147"""
        return [self.observe((*target_location[:-1], i), env) for i in range(self.layers)]

    def valid_location(self, index: tuple[int, ...]) -> bool:
        """gridworld.py:153-179"""
        if hasattr(index, "to_tuple"):
            index = index.to_tuple()
        shape = self.map.shape
        if len(index) != len(shape):
            raise IndexError(f"Index {index} and world shape {shape} must be the same length.")
        if min(index) < 0:
            return False
        return all(y < x for x, y in zip(shape, index))

    def get_entities_of_kind(self, kind: str) -> list[Entity]:
        """gridworld.py:181-196 (template)."""
        return [x for _, x in np.ndenumerate(self.map) if x.kind == kind]

    # ------------------------------------------------------------------------------------------
    # compilation: object template -> kind table + uint8 template -> device instances
    # ------------------------------------------------------------------------------------------
    def build_spec(self, agent_proto: Entity, observation_spec, action_spec, num_agents: int,
                   spawnable, item_groups=(), initial_flags=None) -> WorldSpec:
        """Pure host logic (no CUDA): derive the kind table from the entity prototypes.

        A kind-table row is (class, value, passable, age); `age` > 0 only for classes whose rule is
        `ExpireAfter` (the per-entity turn counter of the reference becomes part of the kind id)."""
        entity_list = list(getattr(observation_spec, "entity_list", []))
        protos: dict[tuple, Entity] = {}
        pending: list[Entity] = []

        def note(e: Entity):
            key = e.kind_key() + (0,)
            if key not in protos:
                protos[key] = e
                pending.append(e)
            return key

        note(self.default_entity)
        for _, e in np.ndenumerate(self.map):
            note(e)
        agent_alt = None
        if getattr(agent_proto, "state_kinds", None):  # two kinds for one agent class (tag: "NotIt" / "It")
            plain, flagged = agent_proto.state_kinds
            agent_proto, agent_alt = copy.copy(agent_proto), copy.copy(agent_proto)
            agent_proto.kind, agent_alt.kind = plain, flagged
            note(agent_alt)
        note(agent_proto)
        for ent, _, _ in item_groups:
            note(ent)
        effects = dict(agent_proto.action_effects(self)) if hasattr(agent_proto, "action_effects") else {}
        for eff in effects.values():
            if not isinstance(eff, SpawnBeam):
                raise ValueError(f"unsupported action effect {eff!r}")
            note(eff.entity)
        # transitive closure over everything a rule can produce
        rules: dict[tuple, object] = {}
        while pending:
            e = pending.pop()
            if e is agent_proto or e is agent_alt:
                continue  # `not isinstance(x, Agent)`: agents are not swept (environment.py:90)
            rule = e.transition_rule(self)
            if rule is None:
                continue
            key = e.kind_key() + (0,)
            rules[key] = rule
            if isinstance(rule, BernoulliSpawn):
                for c in rule.choices:
                    note(c)
            elif isinstance(rule, (BernoulliBecome, BecomeIfLayer, ExpireAfter)):
                note(rule.to)
            elif isinstance(rule, ValueRedraw):
                pass  # its value states are added below
            else:
                raise ValueError(f"{e.kind}: unsupported transition rule {rule!r}")
        spawn_rules = {(r.prob, tuple(c.kind_key() for c in r.choices)) for r in rules.values()
                       if isinstance(r, BernoulliSpawn)}
        if len(spawn_rules) > 1:
            raise ValueError("only one distinct BernoulliSpawn rule per world is supported")
        for key, r in list(rules.items()):  # aged copies of expiring entities, value states of redrawn ones
            if isinstance(r, ExpireAfter):
                for age in range(1, r.sweeps + 1):
                    protos[key[:-1] + (age,)] = protos[key]
            elif isinstance(r, ValueRedraw):
                for state, val in ((1, r.value_else), (2, r.value_if)):
                    twin = copy.copy(protos[key])
                    twin.value = val
                    protos[key[:-1] + (state,)] = twin

        default_kind_str = self.default_entity.kind

        def order(item):
            (name, kind, _, _, age) = item[0]
            if name in entity_list:
                idx = entity_list.index(name)
            elif kind in entity_list and kind != default_kind_str:  # iowa Deck("a") has kind "DeckA"
                idx = entity_list.index(kind)
            else:  # e.g. treasurehunt Sand: looks like the default entity, is its own class
                idx = len(entity_list)
            return (age > 0, idx, name, kind, age)

        keys = [k for k, _ in sorted(protos.items(), key=order)]
        if len(keys) > _abi.SB_MAX_KINDS:
            raise ValueError(f"{len(keys)} entity kinds exceed the engine limit of {_abi.SB_MAX_KINDS}")
        kid = {k: i for i, k in enumerate(keys)}
        kind_of = lambda e, age=0: kid[e.kind_key() + (age,)]  # noqa: E731
        self._kinds = [protos[k] for k in keys]
        C = observation_spec.num_channels
        appearance = np.zeros((len(keys), C), dtype=np.float64)
        for k, i in kid.items():
            appearance[i] = observation_spec.appearance_row(protos[k].kind)  # KeyError like visual_field.py:49
        template = np.zeros((self.height, self.width, self.layers), dtype=np.uint8)
        for idx, e in np.ndenumerate(self.map):
            template[idx] = kind_of(e)
        fill = [i for k, i in kid.items() if protos[k].kind == observation_spec.fill_entity_kind]
        if not fill:
            # the fill kind only matters through its appearance row: add a table-only kind
            raise KeyError(observation_spec.fill_entity_kind)
        spawnable = np.asarray(spawnable, dtype=np.int64).reshape(-1, 3)
        agent_layers = set(spawnable[:, 2].tolist())
        if len(agent_layers) != 1:
            raise ValueError("all spawn locations must be on one layer")
        agent_layer = int(spawnable[0, 2])

        # ---- rule columns of the kind table ----
        K = len(keys)
        kind_rule = [_abi.SB_RULE_NONE] * K
        prob, to, layer, mask = [0.0] * K, [_abi.SB_NO_KIND] * K, [-1] * K, [0] * K
        els = [_abi.SB_NO_KIND] * K
        spawn_rule = None
        for key, r in rules.items():
            i = kid[key]
            if isinstance(r, BernoulliSpawn):
                kind_rule[i], spawn_rule = _abi.SB_RULE_SPAWN, r
            elif isinstance(r, BernoulliBecome):
                kind_rule[i], prob[i], to[i] = _abi.SB_RULE_SPAWN_TO, float(r.prob), kind_of(r.to)
            elif isinstance(r, BecomeIfLayer):
                kind_rule[i], layer[i], to[i] = _abi.SB_RULE_BECOME_IF, int(r.layer), kind_of(r.to)
                mask[i] = sum(1 << j for k2, j in kid.items() if protos[k2].kind in r.kinds)
            elif isinstance(r, ExpireAfter):
                for age in range(r.sweeps + 1):
                    j = kid[key[:-1] + (age,)]
                    kind_rule[j] = _abi.SB_RULE_BECOME
                    to[j] = kid[key[:-1] + (age + 1,)] if age < r.sweeps else kind_of(r.to)
            elif isinstance(r, ValueRedraw):  # every value state redraws the same way
                for state in range(3):
                    j = kid[key[:-1] + (state,)]
                    kind_rule[j], prob[j] = _abi.SB_RULE_SPAWN_TO, float(r.prob)
                    to[j], els[j] = kid[key[:-1] + (2,)], kid[key[:-1] + (1,)]
        extended = any(kr not in (_abi.SB_RULE_NONE, _abi.SB_RULE_SPAWN) for kr in kind_rule)

        # ---- actions: movement decode, facings, beams ----
        deltas, facing, beam_kind = [], [], []
        beam = None
        for i in range(action_spec.n_actions):
            name = action_spec.actions[i]
            if name in MOVE_DELTAS:
                deltas.append(MOVE_DELTAS[name])
                facing.append(MOVE_FACING[name])
                beam_kind.append(_abi.SB_NO_KIND)
            elif name in effects:
                eff = effects[name]
                if beam is not None and (eff.radius, eff.blocked_by) != (beam.radius, beam.blocked_by):
                    raise ValueError("all beam actions of a world must share radius and blocked_by")
                beam = eff
                deltas.append((0, 0))
                facing.append(-1)
                beam_kind.append(kind_of(eff.entity))
            else:
                raise ValueError(f"action {name!r} has no movement decode or effect in the B200 engine")
        reward_all = bool(getattr(agent_proto, "reward_all_layers", False))
        reward_count = int(getattr(agent_proto, "reward_count", 1))
        extra = observation_spec.extras(self) if hasattr(observation_spec, "extras") else None
        groups = []
        for ent, locs, count in item_groups:
            locs = np.asarray([tuple(l) for l in locs], dtype=np.int64).reshape(-1, 3)
            if len(set(locs[:, 2].tolist())) != 1:
                raise ValueError("all locations of an item group must be on one layer")
            groups.append((kind_of(ent), int(locs[0, 2]), int(count), locs[:, :2].copy()))
        spec = WorldSpec(
            height=self.height, width=self.width, layers=self.layers,
            kind_names=[(k[1] if k[1] in entity_list and k[0] not in entity_list and k[1] != default_kind_str
                         else k[0]) + (f"#{k[4]}" if k[4] else "") for k in keys],
            kind_value=[float(protos[k].value) for k in keys],
            kind_passable=[bool(protos[k].passable) for k in keys],
            kind_rule=kind_rule,
            appearance=appearance, template=template, spawnable=spawnable[:, :2].copy(),
            agent_kind=kind_of(agent_proto), agent_layer=agent_layer,
            default_kind=kind_of(self.default_entity), fill_kind=fill[0],
            vision=int(observation_spec.vision_radius), num_agents=int(num_agents),
            spawn_prob=float(spawn_rule.prob) if spawn_rule else 0.0,
            spawn_choices=[kind_of(c) for c in spawn_rule.choices] if spawn_rule else [],
            action_deltas=deltas, obs_transform=int(getattr(observation_spec, "obs_transform", 0)),
            kind_rule_prob=prob if extended else None, kind_rule_to=to if extended else None,
            kind_rule_else=els if extended else None,
            kind_rule_layer=layer if extended else None, kind_rule_mask=mask if extended else None,
            action_facing=facing if beam is not None else None,
            action_beam_kind=beam_kind if beam is not None else None,
            beam_layer=agent_layer + 1 if beam is not None else -1,  # `tile_above`, cleanup/agents.py:107-109
            beam_offsets=np.array(beam_pattern(beam.radius), dtype=np.int64) if beam is not None else None,
            beam_block_mask=sum(1 << j for k2, j in kid.items() if protos[k2].kind in beam.blocked_by)
            if beam is not None else 0,
            reward_all_layers=reward_all, reward_count=reward_count,
            initial_facing=int(getattr(agent_proto, "direction", 2)),
            obs_extra=extra, item_groups=groups,
            agent_kind_alt=kind_of(agent_alt) if agent_alt is not None else -1,
            initial_flags=None if agent_alt is None else np.asarray(
                initial_flags if initial_flags is not None else [0] * int(num_agents), dtype=np.int64),
            touch_swap=bool(agent_alt is not None and getattr(agent_proto, "touch_swap", False)),
            reward_own=None if not getattr(agent_proto, "reward_by_state", None) else
            [float(agent_proto.reward_by_state.get(protos[k].kind, 0.0)) for k in keys],
            obs_flag=bool(agent_alt is not None and getattr(agent_proto, "observe_flag", False)),
        )
        return spec

    def compile(self, agent_proto: Entity, observation_spec, action_spec, num_agents: int, spawnable,
                item_groups=(), initial_flags=None) -> Engine:
        self.spec = self.build_spec(agent_proto, observation_spec, action_spec, num_agents, spawnable, item_groups,
                                    initial_flags)
        self.engine = Engine(self.spec, self.num_envs, device=self.device, env_id_base=self.env_id_base,
                             seed=self.seed, tuning=self.tuning)
        return self.engine

    # ------------------------------------------------------------------------------------------
    # device-side views
    # ------------------------------------------------------------------------------------------
    @property
    def total_reward(self):
        """float64 device tensor [num_envs] once compiled (world.total_reward of every instance)."""
        return self.engine.total_reward if self.engine is not None else self._template_total_reward

    @total_reward.setter
    def total_reward(self, value):
        if self.engine is not None and not isinstance(value, torch.Tensor):
            self.engine.total_reward.fill_(float(value))
        elif self.engine is None:
            self._template_total_reward = value

    def observe_from(self, location) -> torch.Tensor:
        """One-hot crop centred on `location` in every instance: float32 [N, C, S, S]."""
        eng = self.engine
        if isinstance(location, torch.Tensor):
            cells = (location[:, 0] * self.width + location[:, 1]).to(torch.int32)
        else:
            cells = torch.full((eng.N,), int(location[0]) * self.width + int(location[1]), dtype=torch.int32)
        saved = eng.agent_pos
        probe = saved.clone()
        probe[:, 0] = cells.to(eng.device)
        eng.agent_pos = probe
        try:
            out = eng.observe()[:, 0].clone()
        finally:
            eng.agent_pos = saved
        return out

    def instance_map(self, env: int = 0) -> np.ndarray:
        """Object array [H, W, L] of prototype entities for one device instance (debug / rendering)."""
        kinds = self.engine.grid_numpy()[env]
        out = np.empty(kinds.shape, dtype=object)
        for idx, k in np.ndenumerate(kinds):
            out[idx] = self._kinds[int(k)]
        return out
