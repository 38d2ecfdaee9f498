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
