"""CPU-only checks of the boundary: the C-ABI library builds, loads, and exports every
symbol include/pyphy_b200.h declares; calls that need a GPU fail loudly instead of falling
back to anything on the CPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "pyphy_b200.h")).read()
    return sorted(set(re.findall(r"PPB_API\s+[\w\s\*]*?\b(ppb_\w+)\s*\(", txt)))


def test_header_declares_the_expected_entry_points():
    syms = _declared_symbols()
    for s in ("ppb_create", "ppb_destroy", "ppb_set_walls", "ppb_set_balls", "ppb_step_async",
              "ppb_render_async", "ppb_simulate_host", "ppb_read_events", "ppb_get_status", "ppb_last_error"):
        assert s in syms
    assert len(syms) >= 20


def test_library_builds_and_exports_every_declared_symbol():
    from pyphy_engine_b200 import build, _capi
    path = build.build_library()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    for s in _declared_symbols():
        assert hasattr(lib, s), "missing export: " + s
    # the ctypes binding covers exactly the declared surface
    assert sorted(_capi.SIGNATURES) == _declared_symbols()
    assert _capi.load().ppb_version() == 3


def test_no_cpu_fallback():
    """Without a CUDA device ppb_create must fail with an error, never emulate."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pyphy_engine_b200.engine import WorldBatch
    from pyphy_engine_b200._capi import PpbError
    with pytest.raises(PpbError) as ei:
        WorldBatch(4, 2, 4)
    assert "CUDA" in str(ei.value) or "cuda" in str(ei.value)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pyphy_engine_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src, fn
                assert "ref_shim" not in src, fn


def test_argument_validation_happens_before_cuda():
    from pyphy_engine_b200 import _capi
    lib = _capi.load()
    cfg = _capi.Config()
    cfg.n_worlds, cfg.n_balls, cfg.n_walls, cfg.dt = 1, 99, 4, 0.01
    h = ctypes.c_void_p()
    rc = lib.ppb_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc == -1 and b"n_balls" in lib.ppb_last_error()
    cfg.n_balls, cfg.dt = 4, 0.0
    assert lib.ppb_create(ctypes.byref(cfg), ctypes.byref(h)) == -1


def test_register_budget_of_overlapping_kernels():
    """The render kernel (one 384-thread CTA per SM, 120 registers) leaves room for an fp32 advance CTA while the
    physics of the next chunk of steps runs beside the last renders of the current one, and the advance kernel keeps
    at least 5 CTAs (20 warps) per SM when it runs alone: at most 96 registers -- this test is the tripwire."""
    import re
    import shutil
    import subprocess
    from pyphy_engine_b200 import _capi
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump) or not os.path.exists(_capi.LIB_PATH):
        pytest.skip("cuobjdump or the built library is not available")
    out = subprocess.run([cuobjdump, "-res-usage", _capi.LIB_PATH], capture_output=True, text=True).stdout
    regs = {}
    name = None
    for line in out.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
        m = re.search(r"REG:(\d+)", line)
        if m and name:
            regs[name] = int(m.group(1))
    collide = {k: v for k, v in regs.items() if "advance_kernelIfLi" in k and "ELi1E" in k}   # fp32, one warp per world
    render = {k: v for k, v in regs.items() if "render_tile_kernelIf" in k}
    assert collide and render, sorted(regs)
    assert max(collide.values()) <= 96, collide
    assert max(render.values()) <= 120, render
    assert 384 * max(render.values()) + 128 * max(collide.values()) <= 65536


def test_header_is_plain_c():
    """include/pyphy_b200.h is the drop-in boundary: it must compile as C99 and as C++17 on its own (no torch /
    CUDA types in the signatures), so that a cgo / JNI / ctypes binding can include it."""
    import shutil
    import subprocess
    hdr = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "pyphy_b200.h")
    for cc, args in (("gcc", ["-std=c99", "-x", "c"]), ("g++", ["-std=c++17", "-x", "c++"])):
        exe = shutil.which(cc)
        if not exe:
            pytest.skip("%s is not available" % cc)
        r = subprocess.run([exe, "-Wall", "-Wextra", "-Werror", "-fsyntax-only"] + args + [hdr], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
