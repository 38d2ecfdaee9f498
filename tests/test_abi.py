"""The C-ABI library: loads, exports every symbol include/pcr.h declares, struct layouts match
the ctypes mirror, host-only helpers agree with the oracle, and without a GPU it fails loudly
(no CPU fallback).  No device compute here."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pcr.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pcr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    from pointcloud_render_b200 import _native
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in pcr.h but not exported"
    assert sorted(_native.SYMBOLS) == declared
    assert lib.pcr_abi_version() == 4


def test_struct_layouts_match_header(lib, tmp_path):
    from pointcloud_render_b200 import _native
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "pcr.h"\n'
        'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(pcr_camera), sizeof(pcr_style), sizeof(pcr_frame),'
        ' offsetof(pcr_camera, width), offsetof(pcr_style, floor_min), offsetof(pcr_style, xform), offsetof(pcr_frame, W));return 0;}\n')
    exe = tmp_path / "layout"
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.run([cc, "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    C, S, F = _native.Camera, _native.Style, _native.Frame
    assert got == [ctypes.sizeof(C), ctypes.sizeof(S), ctypes.sizeof(F), C.width.offset, S.floor_min.offset,
                   S.xform.offset, F.W.offset]


def test_camera_frame_matches_oracle_bit_for_bit(lib, orc):
    from pointcloud_render_b200 import _native
    from pointcloud_render_b200.presets import PRESETS
    for name, cfg in PRESETS.items():
        for frame in (0, 57, 199, 211):
            for (W, H) in ((1920, 1080), (1024, 1024), (801, 601)):
                cam = cfg.camera(frame, 220, W, H)
                f = _native.camera_frame(cam)
                o = orc.camera_frame(cfg.camera_position(frame, 220), cfg.target, cfg.up, cfg.fov, 0.1, 100.0, W, H)
                for fld in ("L", "U", "D", "O"):
                    assert list(getattr(f, fld)) == list(getattr(o, fld)), (name, frame, fld)
                assert (f.T, f.Th, f.TW, f.W, f.H) == (o.T, o.Th, o.TW, o.W, o.H)
    # Mitsuba look_at convention: left = up x dir, newup = dir x left; example camera looks down at the origin
    f = _native.camera_frame(PRESETS["example"].camera())
    d = -np.array([2.2, 2.2, 4.2]) / np.linalg.norm([2.2, 2.2, 4.2])
    np.testing.assert_allclose(list(f.D), d, atol=1e-7)
    assert abs(np.dot(list(f.L), list(f.D))) < 1e-7 and abs(list(f.L)[2]) < 1e-7 and list(f.U)[2] > 0


def test_degenerate_camera_is_rejected(lib):
    from pointcloud_render_b200 import _native
    with pytest.raises(ValueError):
        _native.camera_frame(_native.make_camera((1, 1, 1), (1, 1, 1)))
    with pytest.raises(ValueError):
        _native.camera_frame(_native.make_camera((0, 0, 1), (0, 0, 0), up=(0, 0, 1)))   # up parallel to dir


def test_no_cpu_fallback(lib):
    """Without a CUDA device every compute entry fails loudly instead of falling back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from pointcloud_render_b200 import _native, renderers
    h = ctypes.c_void_p()
    assert lib.pcr_create(ctypes.byref(h), 0, 1024, 64, 64, 1, 0) == -2 and not h.value     # PCR_ERR_CUDA
    with pytest.raises(RuntimeError):
        _native.Context()
    with pytest.raises(RuntimeError):
        renderers.PointCloudRenderer.init_mitsuba_variant()
    with pytest.raises((RuntimeError, AssertionError)):
        renderers.PointCloudRenderer.standardize_point_cloud(np.zeros((8, 3), np.float32))


def test_invalid_arguments_return_status_not_crash(lib):
    assert lib.pcr_create(None, 0, 1024, 64, 64, 1, 0) == -1
    h = ctypes.c_void_p()
    assert lib.pcr_create(ctypes.byref(h), 0, 0, 64, 64, 1, 0) == -1
    assert lib.pcr_create(ctypes.byref(h), 0, 1024, 70000, 64, 1, 0) == -1
    assert lib.pcr_last_error(None) == b"null context"
    lib.pcr_destroy(None)


def test_product_package_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under pointcloud_render_b200/ may import it."""
    pkg = os.path.join(ROOT, "pointcloud_render_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), fn
                assert "liboracle" not in text and "import pcr_oracle" not in text, fn
    code = "import sys; import pointcloud_render_b200; assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
