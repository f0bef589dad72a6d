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


def test_ctypes_structs_match_the_c_header(tmp_path):
    """Every struct of include/gsd.h compiled by gcc (plain C) has the size and field offsets of its ctypes mirror in _lib.py:
    the header is the contract, the Python binding must not drift from it."""
    import ctypes as C
    import subprocess
    header = open(os.path.join(ROOT, "include", "gsd.h")).read()
    clean = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    structs = {}
    for body, name in re.findall(r"typedef struct\s*\{(.*?)\}\s*(Gsd\w+)\s*;", clean, flags=re.S):
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            # "const float *a, *b" / "float w[2]" / "int32_t G, W" / "GsdRasterFwd fwd"
            first, *rest = [d.strip() for d in decl.split(",")]
            fields.append(re.sub(r"\[.*\]", "", first.split()[-1]).lstrip("*"))
            fields += [re.sub(r"\[.*\]", "", r).lstrip("*").strip() for r in rest]
        structs[name] = fields
    mirrored = {n: getattr(_lib, n) for n in structs if hasattr(_lib, n)}
    assert {"GsdRasterFwd", "GsdRasterBwd", "GsdPhotometric", "GsdTrackLosses", "GsdTrackUpdate", "GsdAdam", "GsdGnnEdges", "GsdDensifyPlan", "GsdDensifyApply"} <= set(mirrored)
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "gsd.h"', 'int main(void) {']
    for n, cls in mirrored.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (n, n))
        for f in structs[n]:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (n, f, n, f))
    lines += ['return 0;', '}']
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for n, cls in mirrored.items():
        assert int(out[n]) == C.sizeof(cls), (n, out[n], C.sizeof(cls))
        py_fields = [f[0] for f in cls._fields_]
        assert py_fields == structs[n], (n, py_fields, structs[n])
        for f in py_fields:
            assert int(out["%s.%s" % (n, f)]) == getattr(cls, f).offset, (n, f)
