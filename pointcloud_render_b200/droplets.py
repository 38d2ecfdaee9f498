"""Host half of the droplet scene (traj_renderer.py / traj_vel_renderer.py).

Only what cannot or need not run on the GPU lives here:
  * the droplet mesh itself (_create_droplet_mesh, traj_renderer.py:102-153) — 340 vertices, built once;
    the reference writes it to `temp_meshes/droplet.obj` with 6 decimals, so the vertices a loader sees are
    those decimals read back as float32.  Nothing is written to disk here.
  * the rotations of points WITHOUT velocity (generate_random_rotation_matrix, :398-418), which come from
    numpy's legacy Mersenne-Twister stream seeded with the point index: a per-index constant table.
Rotations from velocities, Catmull-Rom history trails, binning, raster and shading are CUDA kernels
(csrc/pcr_droplets.cuh) behind pcr_render_droplet_frames.
"""
import functools

import numpy as np

N_SEGMENTS = 20       # traj_renderer.py:111-114
N_RINGS = 16
BASE_RADIUS = 0.008
LENGTH = 0.035


@functools.lru_cache(maxsize=None)
def droplet_vertices():
    """((N_RINGS+1)*N_SEGMENTS, 3) float32, ring-major: a sphere cap of radius BASE_RADIUS down to
    theta = pi/3, then a tail whose radius shrinks as (1-t)^2 while it is pulled back by 0.8*LENGTH*t."""
    rows = []
    for i in range(N_RINGS + 1):
        theta = np.pi * i / N_RINGS
        if theta <= np.pi / 3:
            r, z_off = BASE_RADIUS, 0
        else:
            t = (theta - np.pi / 3) / (2 * np.pi / 3)
            r, z_off = BASE_RADIUS * (1 - t) ** 2, -LENGTH * t * 0.8
        for j in range(N_SEGMENTS):
            phi = 2 * np.pi * j / N_SEGMENTS
            xyz = (r * np.sin(theta) * np.cos(phi), r * np.sin(theta) * np.sin(phi), r * np.cos(theta) + z_off)
            rows.append([float(f"{c:.6f}") for c in xyz])           # the OBJ's text round trip
    out = np.asarray(rows, np.float64).astype(np.float32)
    out.setflags(write=False)
    return out


def droplet_faces():
    """(2*N_RINGS*N_SEGMENTS, 3) int32 zero-based faces in the OBJ's order (the kernels derive them
    from the ring structure; this is for callers that want the mesh)."""
    f = []
    for i in range(N_RINGS):
        for j in range(N_SEGMENTS):
            v0, v1 = i * N_SEGMENTS + j, i * N_SEGMENTS + (j + 1) % N_SEGMENTS
            v2, v3 = v0 + N_SEGMENTS, v1 + N_SEGMENTS
            f += [[v0, v2, v1], [v1, v2, v3]]
    return np.asarray(f, np.int32)


def load_ring_mesh_obj(path, tol=2e-6):
    """A caller-supplied droplet OBJ (`droplet_mesh_path`, traj_renderer.py:93-99) -> (verts float32 ring-major,
    n_rings, n_segments) for pcr_set_droplet_mesh.  The device raster derives the faces from the ring structure and
    the shading normals from the ring profile, so the file must be what _create_droplet_mesh writes for SOME
    (n_rings, n_segments, profile): (n_rings+1) rings of n_segments vertices on circles about the z axis, vertex 0 of
    every ring in the xz half-plane, ring z strictly decreasing (pole rings may coincide with the poles), and the
    reference's face list (v0,v2,v1),(v1,v2,v3) per quad.  Anything else raises ValueError: there is no silent
    fallback to the built-in mesh."""
    verts, faces = [], []
    with open(path) as f:
        for line in f:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "v":
                verts.append([float(tok[1]), float(tok[2]), float(tok[3])])
            elif tok[0] == "f":
                idx = [int(t.split("/")[0]) - 1 for t in tok[1:]]
                if len(idx) != 3:
                    raise ValueError(f"{path}: only triangle faces are supported")
                faces.append(idx)
    if not verts or not faces:
        raise ValueError(f"{path}: no vertices / faces")
    n_seg = faces[0][1] - faces[0][0]                      # first face is (0, n_segments, 1)
    if n_seg < 3 or len(verts) % n_seg or len(verts) // n_seg < 2:
        raise ValueError(f"{path}: not a ring-structured mesh (cannot infer the ring size from the first face)")
    n_rings = len(verts) // n_seg - 1
    want = []
    for i in range(n_rings):
        for j in range(n_seg):
            v0, v1 = i * n_seg + j, i * n_seg + (j + 1) % n_seg
            want += [[v0, v0 + n_seg, v1], [v1, v0 + n_seg, v1 + n_seg]]
    if faces != want:
        raise ValueError(f"{path}: faces are not the ring topology (v0,v2,v1),(v1,v2,v3) of _create_droplet_mesh")
    v = np.asarray(verts, np.float64).reshape(n_rings + 1, n_seg, 3)
    r = np.hypot(v[:, 0, 0], v[:, 0, 1])
    phi = 2 * np.pi * np.arange(n_seg) / n_seg
    ideal = np.stack([r[:, None] * np.cos(phi), r[:, None] * np.sin(phi), np.repeat(v[:, :1, 2], n_seg, 1)], axis=-1)
    if np.abs(v[:, 0, 1]).max() > tol or np.abs(v - ideal).max() > tol:
        raise ValueError(f"{path}: rings are not circles about the z axis starting in the xz half-plane (surface of revolution)")
    if not np.all(np.diff(v[:, 0, 2]) < 0):
        raise ValueError(f"{path}: ring z must decrease strictly from the first ring to the last")
    out = np.ascontiguousarray(v.reshape(-1, 3).astype(np.float32))
    return out, int(n_rings), int(n_seg)


def write_obj(path, verts, faces):
    """The OBJ text _create_droplet_mesh writes (6 decimals, 1-based faces) — for callers that want the file."""
    with open(path, "w") as f:
        for x, y, z in np.asarray(verts, np.float64):
            f.write(f"v {x:.6f} {y:.6f} {z:.6f}\n")
        for a, b, c in np.asarray(faces):
            f.write(f"f {a + 1} {b + 1} {c + 1}\n")


_ROT_CACHE = {}


def random_rotations(n):
    """(n, 9) float32: rotation the reference gives point idx when the input has no velocity columns —
    Rodrigues about normalised randn(3) by uniform(0, 2 pi), both drawn after np.random.seed(idx)."""
    have = _ROT_CACHE.get("table")
    if have is None or have.shape[0] < n:
        start = 0 if have is None else have.shape[0]
        new = np.empty((n - start, 9), np.float32)
        for m, idx in enumerate(range(start, n)):
            rs = np.random.RandomState(idx)
            axis = rs.randn(3)
            axis = axis / np.linalg.norm(axis)
            angle = rs.uniform(0, 2 * np.pi)
            K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
            new[m] = (np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * np.dot(K, K)).reshape(9)
        have = new if have is None else np.concatenate([have, new])
        _ROT_CACHE["table"] = have
    return have[:n]
