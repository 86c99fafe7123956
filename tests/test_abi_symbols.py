"""The C-ABI library loads on a machine without a GPU, exports every symbol include/*.h declares, and
its argument validation (which runs before any CUDA call) behaves.  No compute here."""
import ctypes as C
import glob
import os
import re

import pytest

from hgsfusion_b200 import _lib
from hgsfusion_b200.geometry import make_geometry

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names += re.findall(r"HGSF_API\s+[\w\s\*]+?\b(hgsf_\w+)\s*\(", src)
    return sorted(set(names))


def test_header_declares_the_expected_surface():
    assert declared_symbols() == sorted(_lib.EXPORTS)


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    raw = C.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(raw, name), name
    assert lib.hgsf_abi_version() == 1


def test_no_torch_types_in_the_abi():
    src = open(os.path.join(ROOT, "include", "hgsfusion_b200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)          # comments cite the reference's at::Tensor surface
    assert "at::" not in code and "torch" not in code.lower() and "#include <cuda" not in code


def test_status_strings():
    assert _lib.status_string(0) == "ok"
    for code in (_lib.ERR_INVALID_ARG, _lib.ERR_UNSUPPORTED, _lib.ERR_WORKSPACE, _lib.ERR_DRIVER):
        assert _lib.status_string(code) not in ("ok", "unknown status")


def test_capacity_and_workspace_queries():
    lib = _lib.load()
    g = make_geometry([0, -25.6, -3, 51.2, 25.6, 2], [0.16, 0.16, 5])
    assert list(g.grid) == [320, 320, 1]
    # min(n, B * min(max_voxels, cells))
    assert lib.hgsf_pillar_capacity(C.byref(g), 480000, 16, 40000) == 480000
    assert lib.hgsf_pillar_capacity(C.byref(g), 480000, 16, 1000) == 16000
    assert lib.hgsf_pillar_capacity(C.byref(g), 10 ** 7, 2, 10 ** 9) == 2 * 320 * 320
    assert lib.hgsf_pillar_capacity(C.byref(g), -1, 2, 10) == -1
    need = C.c_size_t(0)
    assert lib.hgsf_workspace_size(C.byref(g), 480000, 16, 7, C.byref(need)) == 0
    table = 16 * 320 * 320 * 12                       # tag, cnt, start per cell
    # + tile records (16 B per 32 cells); per point: key, arrival, row, pillar-list entry
    assert table < need.value < table + 16 * 320 * 10 * 16 + 480000 * (4 + 4 + 32 + 16) + (1 << 20)
    assert lib.hgsf_workspace_size(C.byref(g), 480000, 0, 7, C.byref(need)) == _lib.ERR_INVALID_ARG
    assert lib.hgsf_workspace_size(C.byref(g), 480000, 16, 7, None) == _lib.ERR_INVALID_ARG
    # B * cells must fit int32 keys
    assert lib.hgsf_workspace_size(C.byref(g), 1000, 30000, 7, C.byref(need)) == _lib.ERR_UNSUPPORTED
    assert lib.hgsf_scatter_workspace_size(C.byref(g), 16, C.byref(need)) == 0
    assert need.value >= 16 * 320 * 320 * 4


def test_invalid_arguments_are_rejected_before_any_cuda_call():
    lib = _lib.load()
    g = make_geometry([0, -25.6, -3, 51.2, 25.6, 2], [0.16, 0.16, 5])
    pts, out, pfn = _lib.Points(), _lib.PillarOutputs(), _lib.Pfn()
    assert lib.hgsf_pillarize(None, C.byref(pts), 32, 100, None, 0, C.byref(out), None) == _lib.ERR_INVALID_ARG
    assert lib.hgsf_pillarize(C.byref(g), C.byref(pts), 32, 100, None, 0, C.byref(out), None) == _lib.ERR_INVALID_ARG
    assert lib.hgsf_points_to_bev(C.byref(g), C.byref(pts), None, 32, 100, None, 0, C.byref(out), None) == _lib.ERR_INVALID_ARG
    assert lib.hgsf_pillar_vfe(C.byref(g), None, None, None, None, 1, 1, 0, 32, 7, None, None) == _lib.ERR_INVALID_ARG
    assert lib.hgsf_pointpillar_scatter(C.byref(g), None, None, 1, 0, 64, 0, None, 0, None, None) == _lib.ERR_INVALID_ARG
    bad = make_geometry([0, -25.6, -3, 51.2, 25.6, 2], [0.16, 0.16, 5])
    bad.grid[0] = 0
    assert lib.hgsf_pillar_capacity(C.byref(bad), 10, 1, 10) == -1


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libhgsfusion_b200.so")
    with pytest.raises(ImportError, match="no CPU or PyTorch fallback"):
        _lib.load()
