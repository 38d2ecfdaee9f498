"""Import the UNMODIFIED reference scripts from /root/reference with their two missing
third-party modules stubbed — TEST INFRASTRUCTURE.

`mitsuba` and `plyfile` are not installable here (no network).  Everything the reference
does *before* mi.load_file — standardize_point_cloud, transform_coordinates,
compute_camera_position, compute_color, generate_xml_content — is plain numpy and runs
once the two names resolve.  Used only to pin oracle/pcr_oracle.py and to
generate tests/golden/ (oracle/gen_golden.py).  /root/reference does not exist on the GPU
box: nothing on a `-m gpu` / smoke / bench path may call this.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PCR_REFERENCE_ROOT", "/root/reference")

_MODULES = ("example_renderer", "traj_ball_renderer", "traj_original", "traj_b0", "traj_b1",
            "traj_renderer", "traj_vel_renderer")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "example_renderer.py"))


def load():
    """Return {module_name: module} for the 7 reference scripts."""
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    if "mitsuba" not in sys.modules:
        sys.modules["mitsuba"] = types.ModuleType("mitsuba")
    if "plyfile" not in sys.modules:
        ply = types.ModuleType("plyfile")

        class PlyData:  # only the name has to exist; .ply loading is not exercised
            @staticmethod
            def read(path):
                raise RuntimeError("plyfile is stubbed in the oracle harness")

        ply.PlyData = PlyData
        sys.modules["plyfile"] = ply
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    sys.dont_write_bytecode = True  # /root/reference is read-only
    return {name: importlib.import_module(name) for name in _MODULES}
