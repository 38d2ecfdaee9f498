"""File I/O at the two ends of the hot path: load_point_cloud and the PNG half of save_scene.

Mirrors load_point_cloud of example_renderer.py:101-111 / traj_ball_renderer.py:223-279
(.npy | .npz['pred'] | .ply with x,y,z [+ vx,vy,vz | nx,ny,nz]).  `plyfile` is a third-party
dependency of the reference that is absent here, so a minimal PLY vertex reader is included
(ascii and binary_little_endian / binary_big_endian, scalar properties only).
"""
import os

import numpy as np

_PLY_TYPES = {
    "char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2",
    "ushort": "u2", "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4",
    "float": "f4", "float32": "f4", "double": "f8", "float64": "f8",
}


def read_ply_vertices(path):
    """Return the 'vertex' element of a PLY file as a numpy structured array."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError("not a PLY file")
        fmt = None
        elements = []  # (name, count, [(prop, dtype)])
        while True:
            line = f.readline()
            if not line:
                raise ValueError("PLY header not terminated")
            tok = line.decode("ascii", "replace").split()
            if not tok or tok[0] == "comment" or tok[0] == "obj_info":
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                elements.append((tok[1], int(tok[2]), []))
            elif tok[0] == "property":
                if tok[1] == "list":
                    elements[-1][2].append((tok[4], ("list", _PLY_TYPES[tok[2]], _PLY_TYPES[tok[3]])))
                else:
                    elements[-1][2].append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        for name, count, props in elements:
            scalar = all(not isinstance(t, tuple) for _, t in props)
            if name != "vertex":
                if fmt == "ascii":
                    for _ in range(count):
                        f.readline()
                    continue
                if not scalar:
                    raise ValueError("non-vertex list element before 'vertex' is not supported")
                f.seek(count * sum(np.dtype(t).itemsize for _, t in props), os.SEEK_CUR)
                continue
            if not scalar:
                raise ValueError("list properties on 'vertex' are not supported")
            if fmt == "ascii":
                rows = np.loadtxt(f, max_rows=count, ndmin=2)
                out = np.empty(count, dtype=[(p, t) for p, t in props])
                for k, (p, _) in enumerate(props):
                    out[p] = rows[:, k]
                return out
            order = "<" if fmt == "binary_little_endian" else ">"
            dt = np.dtype([(p, order + t) for p, t in props])
            return np.frombuffer(f.read(count * dt.itemsize), dtype=dt, count=count)
    raise ValueError("PLY file has no 'vertex' element")


def load_point_cloud(file_path, with_velocity=True, verbose=False):
    """.npy / .npz['pred'] / .ply -> (N,3) or (N,6) (or (F,N,C) for stacked .npy), dtype preserved.

    with_velocity=False reproduces example_renderer.py:101-111 (x,y,z only);
    True reproduces traj_ball_renderer.py:223-279 (vx,vy,vz, else nx,ny,nz as velocity)."""
    ext = os.path.splitext(file_path)[1]
    if ext == ".npy":
        return np.load(file_path, allow_pickle=True)
    if ext == ".npz":
        return np.load(file_path)["pred"]
    if ext == ".ply":
        v = read_ply_vertices(file_path)
        names = v.dtype.names
        cols = [v["x"], v["y"], v["z"]]
        if with_velocity:
            if all(k in names for k in ("vx", "vy", "vz")):
                cols += [v["vx"], v["vy"], v["vz"]]
            elif all(k in names for k in ("nx", "ny", "nz")):
                cols += [v["nx"], v["ny"], v["nz"]]
        data = np.column_stack(cols)
        if verbose:
            print(f"  Loaded PLY: shape={data.shape}")
        return data
    raise ValueError("Unsupported file format.")


def write_png(path, rgba):
    """The file half of save_scene (mi.util.write_bitmap, example_renderer.py:159-161).
    rgba: (H,W,4) or (H,W,3) uint8, already sRGB-encoded by K4."""
    from PIL import Image
    a = np.asarray(rgba)
    Image.fromarray(a[..., :3] if a.shape[-1] == 4 else a).save(path)
