"""The C-ABI library loads on a machine without a GPU and exports every symbol the header declares."""
import ctypes
import os
import re

import pytest

from path_tracer_python_b200 import _lib, build

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", params=["strict", "fast", "debug", "wide"])
def L(request):
    """All builds of the library: the strict (parity) one, the fast-numerics variant and the debug-bounds variant
    of the same sources."""
    {"strict": build.build_native, "fast": build.build_fast, "debug": build.build_debug, "wide": build.build_wide}[request.param]()
    return _lib.lib(request.param)


def _declared():
    text = open(os.path.join(REPO, "include", "b200pt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pt_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(L):
    names = _declared()
    assert len(names) >= 19
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/b200pt.h but not exported"
    assert sorted(_lib.EXPORTS) == names


def test_struct_sizes_match_header(L):
    # PtCamera: 6*3 floats + angle + 3 bg + max_depth + jitter; PtCounters: 7 u64
    assert ctypes.sizeof(_lib.PtCamera) == 4 * (18 + 1 + 3 + 1 + 1)
    assert ctypes.sizeof(_lib.PtCounters) == 56
    assert ctypes.sizeof(_lib.PtHostMaterials) == 13 * 8


def test_errors_are_reported_without_a_gpu(L):
    info = _lib.PtPackedInfo()
    rc = L.pt_pack_scene_info(None, ctypes.byref(info))
    assert rc == -1 and b"NULL" in L.pt_last_error()
    assert L.pt_wavefront_workspace_bytes(0, 10) == 0
    assert L.pt_wavefront_workspace_bytes(1920, 1080) > 1920 * 1080 * (6 * 16 + 16 + 12)
    # per in-flight sample and pixel: two queues (2 x 48 B), hit record 16 B, 8 bin entries 32 B, two slabs 24 B
    n = 1920 * 1080
    one, many = L.pt_wavefront_workspace_bytes_for(1920, 1080, 1), L.pt_wavefront_workspace_bytes_for(1920, 1080, 32)
    assert one == L.pt_wavefront_workspace_bytes(1920, 1080)
    assert many - one == 31 * n * 168 and 0 < one - n * 168 <= 64 * 1024
    assert b"sm_100a" in L.pt_version() and (b"strict" in L.pt_version()) != (b"fast" in L.pt_version())
    code = ctypes.c_uint32(7)
    assert L.pt_get_debug_error(None, ctypes.byref(code)) == -1


def test_no_cpu_fallback_when_library_missing(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_libs", {})
    monkeypatch.setattr(_lib, "LIB_PATHS", {"strict": str(tmp_path / "nope.so"), "fast": str(tmp_path / "nope_fast.so")})   # (and no "debug")
    for numerics in ("strict", "fast"):
        with pytest.raises(_lib.B200PTError, match="not built"):
            _lib.lib(numerics)
    with pytest.raises(_lib.B200PTError, match="unknown numerics"):
        _lib.lib("sloppy")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(REPO, "path_tracer_python_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "liboracle" not in src and "oracle/" not in src.replace(
                    "oracle/oracle_math.h", "").replace("oracle/gen_golden.py", ""), f


def test_product_never_reads_the_test_fixtures():
    """The package builds its scenes from its own classes + package data (path_tracer_python_b200/data); the
    golden fixtures under tests/golden are test inputs, loaded only by tests/golden_fixtures.py."""
    pkg = os.path.join(REPO, "path_tracer_python_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert "import golden_fixtures" not in src and "GOLDEN_DIR" not in src and "load_fixture" not in src, f
    assert os.path.exists(os.path.join(pkg, "data", "earthmap_u8.npz"))
