"""ctypes binding of libsorrel_b200.so (include/sorrel_b200.h).

The library is the product: if it is missing, cannot be loaded, or a call fails, this module
raises -- there is no Python / CPU fallback for any entry point.
"""
from __future__ import annotations

import ctypes as C
import os

SB_ABI_VERSION = 5
SB_MAX_KINDS = 32
SB_MAX_CHANNELS = 32
SB_MAX_LAYERS = 8
SB_MAX_CHOICES = 8
SB_MAX_ACTIONS = 8
SB_MAX_AGENTS = 64
SB_PAD_KIND = 0xFF
SB_NO_KIND = 0xFF
SB_MAX_BEAM_CELLS = 16
SB_MAX_ITEM_GROUPS = 4

SB_OK, SB_EINVAL, SB_EUNSUPPORTED, SB_EALIGN, SB_ECUDA, SB_ENODEVICE, SB_EVISION = 0, -1, -2, -3, -4, -5, -6
SB_ERR_OFFGRID, SB_ERR_TAPE_ALIGN, SB_ERR_BAD_ACTION, SB_ERR_BAD_KIND, SB_ERR_BAD_INDEX = 1, 2, 4, 8, 16
SB_RULE_NONE, SB_RULE_SPAWN, SB_RULE_SPAWN_TO, SB_RULE_BECOME_IF, SB_RULE_BECOME = 0, 1, 2, 3, 4
SB_RNG_TAPE, SB_RNG_PHILOX, SB_RNG_SPARSE = 0, 1, 2
SB_SPARSE_KMAX = 64
SB_POLICY_TAPE, SB_POLICY_UNIFORM = 0, 1
SB_MODE_FUSED, SB_MODE_SPLIT = 0, 1
SB_OBS_SUM, SB_OBS_RGB = 0, 1

LIB_PATH = os.environ.get("SB_LIB_PATH") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib",
                                                         "libsorrel_b200.so")


class SbWorldDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("num_envs", C.c_int32),
        ("num_agents", C.c_int32),
        ("height", C.c_int32),
        ("width", C.c_int32),
        ("num_layers", C.c_int32),
        ("num_dyn_layers", C.c_int32),
        ("dyn_layers", C.c_int32 * SB_MAX_LAYERS),
        ("agent_layer", C.c_int32),
        ("num_kinds", C.c_int32),
        ("num_channels", C.c_int32),
        ("vision", C.c_int32),
        ("fill_kind", C.c_int32),
        ("agent_kind", C.c_int32),
        ("default_kind", C.c_int32),
        ("num_actions", C.c_int32),
        ("action_dy", C.c_int8 * SB_MAX_ACTIONS),
        ("action_dx", C.c_int8 * SB_MAX_ACTIONS),
        ("num_spawn_choices", C.c_int32),
        ("spawn_choice_kinds", C.c_uint8 * SB_MAX_CHOICES),
        ("spawn_prob", C.c_double),
        ("env_id_base", C.c_uint64),
        ("kind_value", C.c_float * SB_MAX_KINDS),
        ("kind_passable", C.c_uint8 * SB_MAX_KINDS),
        ("kind_rule", C.c_uint8 * SB_MAX_KINDS),
        ("d_appearance", C.c_void_p),
        ("h_appearance", C.c_void_p),
        ("d_template", C.c_void_p),
        ("static_layers_invisible", C.c_int32),
        ("spawn_cdf_len", C.c_int32),
        ("spawn_cdf", C.c_uint64 * 64),
        ("obs_transform", C.c_int32),
        ("reserved1", C.c_int32),
        # ---- ABI v2 ----
        ("kind_rule_prob", C.c_double * SB_MAX_KINDS),
        ("kind_rule_to", C.c_uint8 * SB_MAX_KINDS),
        ("kind_rule_layer", C.c_int8 * SB_MAX_KINDS),
        ("kind_rule_mask", C.c_uint32 * SB_MAX_KINDS),
        ("action_facing", C.c_int8 * SB_MAX_ACTIONS),
        ("action_beam_kind", C.c_uint8 * SB_MAX_ACTIONS),
        ("uses_facing", C.c_int32),
        ("beam_layer", C.c_int32),
        ("num_beam_cells", C.c_int32),
        ("beam_dy", (C.c_int8 * SB_MAX_BEAM_CELLS) * 4),
        ("beam_dx", (C.c_int8 * SB_MAX_BEAM_CELLS) * 4),
        ("beam_block_mask", C.c_uint32),
        ("reward_all_layers", C.c_int32),
        ("reward_count", C.c_int32),
        ("obs_extra", C.c_int32),
        ("d_obs_extra", C.c_void_p),
        # ---- ABI v3 ----
        ("kind_rule_else", C.c_uint8 * SB_MAX_KINDS),
        ("use_value64", C.c_int32),
        ("reserved2", C.c_int32),
        ("kind_value64", C.c_double * SB_MAX_KINDS),
        # ---- ABI v4 ----
        ("agent_kind_alt", C.c_int32),
        ("touch_swap", C.c_int32),
        ("reward_own", C.c_int32),
        ("obs_flag", C.c_int32),
        ("toroidal", C.c_int32),
        ("reserved3", C.c_int32),
        ("d_agent_flag", C.c_void_p),
    ]


class SbItemGroup(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("layer", C.c_int32),
        ("count", C.c_int32),
        ("n_cells", C.c_int32),
        ("d_cells", C.c_void_p),
        ("d_placement", C.c_void_p),
    ]


class SbBufferDesc(C.Structure):
    _fields_ = [
        ("capacity", C.c_int32),
        ("obs_size", C.c_int32),
        ("n_frames", C.c_int32),
        ("reserved", C.c_int32),
        ("rows", C.c_int64),
        ("d_states", C.c_void_p),
        ("d_actions", C.c_void_p),
        ("d_rewards", C.c_void_p),
        ("d_dones", C.c_void_p),
        ("d_positions", C.c_void_p),
        # ---- ABI v5 ----
        ("d_error", C.c_void_p),
        ("states_row_stride", C.c_int64),
        ("states_slot_stride", C.c_int64),
    ]


class SbStepArgs(C.Structure):
    _fields_ = [
        ("mode", C.c_int32),
        ("num_turns", C.c_int32),
        ("tick0", C.c_uint32),
        ("slot0", C.c_int32),
        ("rng", C.c_int32),
        ("policy", C.c_int32),
        ("seed", C.c_uint64),
        ("do_sweep", C.c_int32),
        ("act_agent", C.c_int32),
        ("obs_agent", C.c_int32),
        ("reserved", C.c_int32),
        ("d_actions", C.c_void_p),
        ("d_spawn_u", C.c_void_p),
        ("d_spawn_c", C.c_void_p),
        ("d_obs_out", C.c_void_p),
        ("d_reward_out", C.c_void_p),
        ("d_action_out", C.c_void_p),
        ("d_drew_out", C.c_void_p),
        ("d_error", C.c_void_p),
        ("d_clock", C.c_void_p),
        ("d_agent_facing", C.c_void_p),
        # ---- ABI v5: launch tuning (0 = the library chooses) ----
        ("tune_lpe", C.c_int32),
        ("tune_staged", C.c_int32),
        ("tune_par", C.c_int32),
        ("actions_u8", C.c_int32),
    ]


class SbStepPlan(C.Structure):
    _fields_ = [
        ("lpe", C.c_int32),
        ("kernel_rng", C.c_int32),
        ("fast_onehot", C.c_int32),
        ("staged", C.c_int32),
        ("par", C.c_int32),
        ("threads", C.c_int32),
        ("smem_bytes", C.c_int32),
        ("blocks", C.c_int64),
    ]

    KERNEL_RNG = {0: "replay", 1: "philox byte-SIMD", 2: "philox generic", 3: "sparse"}

    def as_dict(self) -> dict:
        return {"kernel": f"sb::step_kernel<{self.lpe},{self.kernel_rng},{self.fast_onehot},{self.staged}>",
                "lanes_per_instance": self.lpe, "sweep": self.KERNEL_RNG.get(self.kernel_rng, "?"),
                "fast_onehot": bool(self.fast_onehot), "planes_staged": bool(self.staged),
                "agent_parallel": bool(self.par), "threads": self.threads, "smem_bytes": self.smem_bytes,
                "blocks": int(self.blocks)}


# name -> (restype, argtypes); every symbol include/sorrel_b200.h declares
SIGNATURES = {
    "sb_abi_version": (C.c_int, []),
    "sb_strerror": (C.c_char_p, [C.c_int]),
    "sb_last_cuda_error": (C.c_char_p, []),
    "sb_launch_count": (C.c_uint64, []),
    "sb_plane_stride": (C.c_int32, [C.c_int32, C.c_int32]),
    "sb_grid_bytes": (C.c_size_t, [C.POINTER(SbWorldDesc)]),
    "sb_validate": (C.c_int, [C.POINTER(SbWorldDesc)]),
    "sb_reset": (C.c_int, [C.POINTER(SbWorldDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                           C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p]),
    "sb_reset_ex": (C.c_int, [C.POINTER(SbWorldDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                              C.c_void_p, C.POINTER(SbItemGroup), C.c_int32, C.c_uint64, C.c_uint32, C.c_void_p]),
    "sb_step": (C.c_int, [C.POINTER(SbWorldDesc), C.POINTER(SbStepArgs), C.c_void_p, C.c_void_p, C.c_void_p,
                          C.POINTER(SbBufferDesc), C.c_void_p]),
    "sb_step_plan": (C.c_int, [C.POINTER(SbWorldDesc), C.POINTER(SbStepArgs), C.POINTER(SbStepPlan)]),
    "sb_clock_advance": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]),
    "sb_observe": (C.c_int, [C.POINTER(SbWorldDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sb_buffer_add": (C.c_int, [C.POINTER(SbBufferDesc), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p]),
    "sb_buffer_sample": (C.c_int, [C.POINTER(SbBufferDesc), C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sb_buffer_sample_transformer": (C.c_int, [C.POINTER(SbBufferDesc), C.c_void_p, C.c_void_p, C.c_int32,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_void_p]),
    "sb_buffer_current_state": (C.c_int, [C.POINTER(SbBufferDesc), C.c_int32, C.c_void_p, C.c_void_p]),
    "sb_buffer_current_state_dev": (C.c_int, [C.POINTER(SbBufferDesc), C.c_void_p, C.c_void_p, C.c_void_p]),
    "sb_stats_scratch_bytes": (C.c_size_t, []),
    "sb_reward_stats": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sb_stats_combine": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
}

_lib = None


class SorrelB200Error(RuntimeError):
    """The CUDA library is missing or a call into it failed."""


def lib() -> C.CDLL:
    """Load libsorrel_b200.so (once).  Raises if it has not been built -- never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SorrelB200Error(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  sorrel_b200 has no CPU fallback."
            )
        handle = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the .so does not export it
            fn.restype = restype
            fn.argtypes = argtypes
        if handle.sb_abi_version() != SB_ABI_VERSION:
            raise SorrelB200Error("libsorrel_b200.so ABI version mismatch; rebuild")
        _lib = handle
    return _lib


def check(status: int, what: str = "") -> None:
    """Map an SbStatus to the exception type the reference raises in the same situation
    (SURVEY.md §8b 'Error conventions')."""
    if status == SB_OK:
        return
    msg = lib().sb_strerror(status).decode()
    if status == SB_ECUDA:
        msg += ": " + lib().sb_last_cuda_error().decode()
    if what:
        msg = f"{what}: {msg}"
    if status in (SB_EINVAL, SB_EUNSUPPORTED, SB_EALIGN, SB_EVISION):
        raise ValueError(msg)
    raise SorrelB200Error(msg)


def error_bits_to_exception(bits: int) -> Exception | None:
    """Device error word -> the reference's exception for that condition."""
    if bits == 0:
        return None
    names = []
    if bits & SB_ERR_OFFGRID:
        names.append("agent moved off the grid (IndexError in the reference, gridworld.py:95-122)")
    if bits & SB_ERR_TAPE_ALIGN:
        names.append("replay tape does not line up with the cells that drew")
    if bits & SB_ERR_BAD_ACTION:
        names.append("action outside the ActionSpec (KeyError in the reference, action_spec.py:27-36)")
    if bits & SB_ERR_BAD_KIND:
        names.append("grid holds a kind id >= num_kinds")
    if bits & SB_ERR_BAD_INDEX:
        names.append("Buffer.sample index outside the ring (IndexError in the reference, buffers.py:112-119)")
    msg = "; ".join(names)
    if bits & (SB_ERR_OFFGRID | SB_ERR_BAD_INDEX):
        return IndexError(msg)
    if bits & SB_ERR_BAD_ACTION:
        return KeyError(msg)
    return SorrelB200Error(msg)
