"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/gsd.h declares (no compute calls)."""
import os
import re

from gs_dynamics_b200 import _lib
from gs_dynamics_b200.csrc import build as cuda_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    cuda_build.build_lib()
    header = open(os.path.join(ROOT, "include", "gsd.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(gsd_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    lib = _lib.lib()
    for sym in sorted(declared):
        assert hasattr(lib, sym), "libgsd_b200.so does not export %s" % sym
    assert declared == set(_lib.EXPORTS), (declared ^ set(_lib.EXPORTS))
    assert lib.gsd_version() >= 100


def test_invalid_arguments_are_reported_not_crashed():
    lib = _lib.lib()
    import ctypes as C
    out = (C.c_size_t * 4)()
    assert lib.gsd_raster_workspace_bytes(-1, 640, 480, 1, 10, out) != 0
    assert b"invalid" in lib.gsd_last_error()
    assert lib.gsd_raster_forward(None, None) != 0
    assert lib.gsd_adam_step(None, None) != 0
    nbytes = C.c_size_t()
    assert lib.gsd_photometric_workspace_bytes(3, 480, 640, C.byref(nbytes)) == 0 and nbytes.value > 3 * 480 * 640 * 4 * 3
    assert lib.gsd_track_losses_workspace_bytes(1000, 10, C.byref(nbytes)) == 0 and nbytes.value > 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    import pytest
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.GsdError):
        _lib.lib()
