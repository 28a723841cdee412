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
