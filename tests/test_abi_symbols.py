"""CPU: the C-ABI library loads without a GPU and exports every symbol the header declares."""
from __future__ import annotations

import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sorrel_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def built_lib():
    import __graft_entry__ as ge

    ge.build()
    from sorrel_b200 import _abi

    assert os.path.exists(_abi.LIB_PATH)
    return _abi


def test_header_and_binding_agree(built_lib):
    assert declared_functions() == sorted(built_lib.SIGNATURES)


def test_library_exports_every_declared_symbol(built_lib):
    handle = ctypes.CDLL(built_lib.LIB_PATH)
    for name in declared_functions():
        assert hasattr(handle, name), f"{name} declared in include/sorrel_b200.h but not exported"


def test_struct_sizes_match_header(built_lib):
    """Compile a tiny C program against the header and compare sizeof() with the ctypes mirrors."""
    import subprocess
    import tempfile

    src = '#include <stdio.h>\n#include "sorrel_b200.h"\nint main(){printf("%zu %zu %zu\\n",' \
          "sizeof(SbWorldDesc),sizeof(SbBufferDesc),sizeof(SbStepArgs));return 0;}\n"
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes == [ctypes.sizeof(built_lib.SbWorldDesc), ctypes.sizeof(built_lib.SbBufferDesc),
                     ctypes.sizeof(built_lib.SbStepArgs)]


def test_pure_host_entry_points(built_lib):
    L = built_lib.lib()
    assert L.sb_abi_version() == built_lib.SB_ABI_VERSION
    assert L.sb_plane_stride(21, 21) == 448 and L.sb_plane_stride(128, 128) == 16384
    assert L.sb_strerror(built_lib.SB_EVISION).decode().startswith("vision")
    assert L.sb_stats_scratch_bytes() > 0
    d = built_lib.SbWorldDesc()
    assert L.sb_validate(ctypes.byref(d)) == built_lib.SB_EINVAL  # zeroed descriptor is rejected


def test_no_cpu_fallback():
    """The product refuses to run without CUDA instead of silently computing on the host."""
    import numpy as np
    import torch

    from sorrel_b200 import SorrelB200Error
    from sorrel_b200.buffers import Buffer

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(SorrelB200Error):
        Buffer(8, (4,), device="cpu")
    import sorrel_b200

    pkg = os.path.dirname(sorrel_b200.__file__)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def _host_desc(spec):
    """SbWorldDesc for an oracle spec without touching a GPU (device pointers are opaque non-NULL integers:
    sb_validate only reads the HOST copy of the appearance table)."""
    import numpy as np

    import sb_testutil as util
    from sorrel_b200 import _abi
    from sorrel_b200.engine import build_desc

    es = util.engine_spec(spec)
    dyn = es.resolved_dynamic_layers()
    stride = (es.height * es.width + 15) // 16 * 16
    happ = np.ascontiguousarray(es.appearance, dtype=np.float64)
    width = 0 if es.obs_extra is None else int(es.obs_extra.shape[2])
    d, _ = build_desc(es, 8, 0, dyn, len(dyn) * stride, happ, 0x1000, 0x2000, True, width, 0x3000 if width else 0)
    return d, happ, _abi


def test_descriptor_validation_without_gpu(built_lib):
    """sb_validate on real world descriptions and on every way of breaking them (status codes of SURVEY §8b)."""
    from oracle import restatement as R

    L = built_lib.lib()
    ok, inval, unsup, evision = built_lib.SB_OK, built_lib.SB_EINVAL, built_lib.SB_EUNSUPPORTED, built_lib.SB_EVISION
    for spec in (R.treasurehunt_spec(), R.treasurehunt_spec(9, 15, 4, 3, 0.03, obs="rgb"), R.cleanup_spec()):
        d, keep, _ = _host_desc(spec)
        assert L.sb_validate(ctypes.byref(d)) == ok

    def broken(spec, mutate):
        d, keep, _ = _host_desc(spec)
        mutate(d)
        return L.sb_validate(ctypes.byref(d))

    th, cl = R.treasurehunt_spec(), R.cleanup_spec()
    assert broken(th, lambda d: setattr(d, "abi_version", 1)) == inval
    assert broken(th, lambda d: setattr(d, "vision", 11)) == evision  # > (H-1)//2: visual_field.py:78-87
    assert broken(th, lambda d: setattr(d, "num_kinds", 33)) == unsup
    assert broken(th, lambda d: setattr(d, "spawn_prob", 1.5)) == inval
    assert broken(th, lambda d: setattr(d, "fill_kind", 9)) == inval
    assert broken(th, lambda d: setattr(d, "d_template", None)) == inval
    assert broken(th, lambda d: d.kind_rule.__setitem__(0, 9)) == unsup  # unknown rule id
    assert broken(th, lambda d: setattr(d, "agent_layer", 0)) == inval  # agent layer must be dynamic
    # column rules
    assert broken(cl, lambda d: d.kind_rule_prob.__setitem__(2, 2.0)) == inval  # River: probability > 1
    assert broken(cl, lambda d: d.kind_rule_to.__setitem__(4, 40)) == inval  # AppleTree -> unknown kind
    assert broken(cl, lambda d: d.kind_rule_layer.__setitem__(3, 5)) == inval  # Pollution tests a missing layer
    # beams / facings / positional code
    assert broken(cl, lambda d: setattr(d, "beam_layer", 7)) == inval
    assert broken(cl, lambda d: setattr(d, "num_beam_cells", 17)) == unsup
    assert broken(cl, lambda d: d.action_facing.__setitem__(0, 4)) == inval
    assert broken(cl, lambda d: d.action_beam_kind.__setitem__(4, 99)) == inval
    assert broken(cl, lambda d: setattr(d, "d_obs_extra", None)) == inval


def test_item_group_struct_matches_header(built_lib):
    import subprocess
    import tempfile

    src = '#include <stdio.h>\n#include "sorrel_b200.h"\nint main(){printf("%zu\\n",sizeof(SbItemGroup));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        assert int(subprocess.check_output([exe]).split()[0]) == ctypes.sizeof(built_lib.SbItemGroup)


def test_step_plan_without_gpu(built_lib):
    """sb_step_plan is pure host code: the launch variant per BASELINE shape, the tuning fields of ABI v5 (which
    replaced the v4 environment variables) and their validation."""
    from oracle import restatement as R

    L, abi = built_lib.lib(), built_lib

    def plan(spec, n, **tune):
        d, keep, _ = _host_desc(spec)
        d.num_envs = n
        a = abi.SbStepArgs()
        a.mode, a.num_turns, a.rng, a.policy = abi.SB_MODE_FUSED, 25, tune.pop("rng", abi.SB_RNG_PHILOX), abi.SB_POLICY_UNIFORM
        a.act_agent = a.obs_agent = -1
        a.tune_lpe, a.tune_staged, a.tune_par = tune.get("lpe", 0), tune.get("staged", 0), tune.get("par", 0)
        out = abi.SbStepPlan()
        return L.sb_step_plan(ctypes.byref(d), ctypes.byref(a), ctypes.byref(out)), out

    th1 = R.treasurehunt_spec(21, 21, 1, 2, 0.005)
    rc, p = plan(th1, 1 << 20)  # config 5: the headline variant
    assert rc == abi.SB_OK and (p.lpe, p.kernel_rng, p.fast_onehot, p.staged) == (4, 1, 1, 1)
    assert p.blocks == (1 << 20) // 32 and p.threads == 128 and 0 < p.smem_bytes <= 227 * 1024 // 8
    rc, p = plan(th1, 4096)  # config 2: small batches widen the groups
    assert rc == abi.SB_OK and p.lpe > 4
    rc, p = plan(R.treasurehunt_spec(21, 21, 8, 2, 0.005), 16384)  # config 3
    assert rc == abi.SB_OK and p.lpe == 8 and p.staged == 1
    rc, p = plan(R.treasurehunt_spec(128, 128, 64, 4, 0.005), 2048)  # config 4: block per instance, plane staged
    assert rc == abi.SB_OK and (p.lpe, p.staged, p.par, p.blocks) == (128, 1, 1, 2048)
    assert p.smem_bytes <= 227 * 1024 // 6  # six blocks per SM
    rc, p = plan(R.treasurehunt_spec(128, 128, 64, 4, 0.005), 2048, lpe=32)  # ... or, pinned, warp per instance
    assert rc == abi.SB_OK and (p.lpe, p.staged, p.par) == (32, 0, 1)
    rc, p = plan(R.treasurehunt_spec(128, 128, 64, 4, 0.005), 2048, par=-1)
    assert rc == abi.SB_OK and (p.lpe, p.par) == (32, 0)
    assert plan(th1, 64, lpe=128)[0] == abi.SB_EUNSUPPORTED  # one agent: the block kernel does not apply
    rc, p = plan(th1, 64, lpe=4)  # what the tests and smoke() pin
    assert rc == abi.SB_OK and p.lpe == 4
    rc, p = plan(th1, 64, staged=-1)
    assert rc == abi.SB_OK and (p.lpe, p.staged) == (32, 0)
    rc, p = plan(th1, 1 << 20, rng=abi.SB_RNG_SPARSE)
    assert rc == abi.SB_OK and p.kernel_rng == 3
    assert plan(th1, 64, lpe=5)[0] == abi.SB_EINVAL
    assert plan(th1, 64, lpe=8, staged=-1)[0] == abi.SB_EINVAL  # planes in place need one warp per instance
    assert plan(th1, 64, staged=2)[0] == abi.SB_EINVAL
    assert plan(R.treasurehunt_spec(128, 128, 64, 4, 0.005), 64, lpe=4, staged=1)[0] == abi.SB_EUNSUPPORTED  # smem


def test_library_reads_no_environment_variables():
    """ABI v5: kernel variants are chosen from SbStepArgs, never from the process environment."""
    src = os.path.join(ROOT, "sorrel_b200", "csrc")
    for f in os.listdir(src):
        if f.endswith((".cu", ".cuh", ".h")):
            assert "getenv" not in open(os.path.join(src, f)).read(), f


def test_step_plan_struct_matches_header(built_lib):
    import subprocess
    import tempfile

    src = '#include <stdio.h>\n#include "sorrel_b200.h"\nint main(){printf("%zu\\n",sizeof(SbStepPlan));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        assert int(subprocess.check_output([exe]).split()[0]) == ctypes.sizeof(built_lib.SbStepPlan)
