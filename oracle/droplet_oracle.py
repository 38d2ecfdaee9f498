"""CPU restatement of the droplet / history-trail scene of traj_renderer.py and traj_vel_renderer.py
 — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates, in vectorised numpy, what the reference emits per point before Mitsuba sees it:

  droplet_mesh              _create_droplet_mesh                      traj_renderer.py:102-153
  rotation_from_velocity    generate_rotation_matrix_from_velocity    traj_renderer.py:159-202
  random_rotation           generate_random_rotation_matrix           traj_renderer.py:398-418
  history_trails            _add_trail_lines (Catmull-Rom polyline)    traj_renderer.py:204-396

Pinned against the UNMODIFIED reference: tests/test_oracle_vs_reference.py runs the reference's
own functions (stub-imported) on the same inputs and asserts bit equality with these — the mesh
against the OBJ file it writes, the trails against the curve files it writes — and
tests/golden/droplets.npz carries the reference's outputs to the GPU box.

The pixel half (which triangle a ray hits) is Mitsuba's and is PARITY UNPINNED; oracle/raycast.c
states it as arithmetic contract "VA-3" (DESIGN.md §3).
"""
import numpy as np

from .pcr_oracle import round6

N_SEGMENTS = 20          # traj_renderer.py:111
N_RINGS = 16             # :112
BASE_RADIUS = 0.008      # :113
LENGTH = 0.035           # :114
HISTORY_FRAMES = 20      # trail_length_frames, traj_renderer.py:218
N_SAMPLES = 20           # :272
MIN_DISTANCE = 1e-5      # :360
MAX_CTRL = N_SAMPLES + 1


def droplet_profile():
    """(r, z) of the 17 rings in float64, before the OBJ text round trip (traj_renderer.py:119-130)."""
    theta = np.pi * np.arange(N_RINGS + 1) / N_RINGS
    cap = theta <= np.pi / 3
    t = (theta - np.pi / 3) / (2 * np.pi / 3)
    r = np.where(cap, BASE_RADIUS, BASE_RADIUS * (1 - t) ** 2)
    z_off = np.where(cap, 0.0, -LENGTH * t * 0.8)
    return theta, r, z_off


def droplet_mesh():
    """Vertices (340,3) float32 as a loader parses the OBJ's 6-decimal text, faces (640,3) int32
    zero-based, in file order (ring-major vertices; per quad (v0,v2,v1), (v1,v2,v3))."""
    theta, r, z_off = droplet_profile()
    phi = 2 * np.pi * np.arange(N_SEGMENTS) / N_SEGMENTS
    x = (r * np.sin(theta))[:, None] * np.cos(phi)[None, :]
    y = (r * np.sin(theta))[:, None] * np.sin(phi)[None, :]
    z = np.broadcast_to((r * np.cos(theta) + z_off)[:, None], x.shape)
    v64 = np.stack([x, y, z], axis=-1).reshape(-1, 3)
    verts = np.array([[float(f"{c:.6f}") for c in row] for row in v64], np.float64).astype(np.float32)
    i, j = np.meshgrid(np.arange(N_RINGS), np.arange(N_SEGMENTS), indexing="ij")
    v0 = i * N_SEGMENTS + j
    v1 = i * N_SEGMENTS + (j + 1) % N_SEGMENTS
    v2 = (i + 1) * N_SEGMENTS + j
    v3 = (i + 1) * N_SEGMENTS + (j + 1) % N_SEGMENTS
    faces = np.stack([np.stack([v0, v2, v1], -1), np.stack([v1, v2, v3], -1)], axis=2).reshape(-1, 3)
    return verts, faces.astype(np.int32)


def rotation_from_velocity(vel):
    """(N,3) velocities -> (N,3,3) float64 rotations taking the mesh's tip axis (0,0,-1) onto the
    velocity direction (Rodrigues), identity for |v| < 1e-6.  Same operations as the reference, with
    the constant default direction folded in: dot = -t_z, axis = (t_y, -t_x, 0)."""
    v = np.asarray(vel, np.float64).reshape(-1, 3)
    n = v.shape[0]
    R = np.broadcast_to(np.eye(3), (n, 3, 3)).copy()
    vn = np.sqrt((v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]) + v[:, 2] * v[:, 2])
    for k in np.nonzero(vn >= 1e-6)[0]:
        t = v[k] / vn[k]
        d = min(max(-t[2], -1.0), 1.0)
        ax = np.array([t[1], -t[0], 0.0])
        an = np.sqrt((ax[0] * ax[0] + ax[1] * ax[1]) + ax[2] * ax[2])
        if an < 1e-8:
            if d > 0.999:
                continue
            tmp = np.array([1.0, 0.0, 0.0]) if abs(t[0]) < 0.9 else np.array([0.0, 1.0, 0.0])
            ax = np.array([t[1] * tmp[2] - t[2] * tmp[1], t[2] * tmp[0] - t[0] * tmp[2], t[0] * tmp[1] - t[1] * tmp[0]])
            an = np.sqrt((ax[0] * ax[0] + ax[1] * ax[1]) + ax[2] * ax[2])
            ax = ax / an if an > 1e-8 else np.array([0.0, 1.0, 0.0])
            ang = np.pi
        else:
            ax = ax / an
            ang = np.arccos(d)
        c, s = np.cos(ang), np.sin(ang)
        K = np.array([[0.0, -ax[2], ax[1]], [ax[2], 0.0, -ax[0]], [-ax[1], ax[0], 0.0]])
        KK = np.empty((3, 3))
        for i in range(3):
            for j in range(3):
                KK[i, j] = (K[i, 0] * K[0, j] + K[i, 1] * K[1, j]) + K[i, 2] * K[2, j]
        R[k] = (np.eye(3) + s * K) + (1 - c) * KK
    return R


def random_rotation(indices):
    """Rotations the reference gives points WITHOUT velocity: axis = normalised randn(3), angle =
    uniform(0, 2 pi) from numpy's legacy global generator seeded with the point index."""
    out = np.empty((len(indices), 3, 3))
    for m, idx in enumerate(indices):
        rs = np.random.RandomState(int(idx))
        ax = rs.randn(3)
        ax = ax / np.sqrt((ax[0] * ax[0] + ax[1] * ax[1]) + ax[2] * ax[2])
        ang = rs.uniform(0, 2 * np.pi)
        c, s = np.cos(ang), np.sin(ang)
        K = np.array([[0.0, -ax[2], ax[1]], [ax[2], 0.0, -ax[0]], [-ax[1], ax[0], 0.0]])
        out[m] = (np.eye(3) + s * K) + (1 - c) * (K @ K)
    return out


def to_world_f32(R, translation):
    """(N,3,3) f64 rotation + (N,3) f32 translation -> (N,12) float32 rows [R | t] as a loader
    reads the `{}`-formatted matrix of DROPLET_SEGMENT (f64 repr -> f32)."""
    R = np.asarray(R, np.float64)
    t = np.asarray(translation, np.float32).astype(np.float64)
    return np.concatenate([R, t[:, :, None]], axis=2).reshape(-1, 12).astype(np.float32)


def sample_plan(h):
    """For a history of h >= 3 frames: which (segment, t) pairs the reference keeps, in order.
    Returns (seg[20], t[20]) with t as python floats (traj_renderer.py:283-316)."""
    n_seg = h - 1
    sps = max(2, N_SAMPLES // n_seg)
    pairs = [(s, (i / (sps - 1)) if sps > 1 else 0) for s in range(n_seg) for i in range(sps)]
    if len(pairs) > N_SAMPLES:
        keep = np.linspace(0, len(pairs) - 1, N_SAMPLES).astype(int)
        pairs = [pairs[i] for i in keep]
    while len(pairs) < N_SAMPLES:
        pairs.append(pairs[-1])
    return [p[0] for p in pairs], [p[1] for p in pairs]


def catmull_rom_samples(hist):
    """hist: (h, N, 3) float32 history positions, oldest first, 2 <= h <= 20.  Returns the 20 samples
    per point, (N, 20, 3) float32 — every operation is a float32 numpy ufunc with the python-float
    parameters t, t^2, t^3 cast to float32 (weak scalars), in the reference's order."""
    pa = np.asarray(hist, np.float32)
    h, n, _ = pa.shape
    out = np.empty((n, N_SAMPLES, 3), np.float32)
    if h == 2:
        for i in range(N_SAMPLES):
            t = i / (N_SAMPLES - 1)
            out[:, i] = np.float32(1 - t) * pa[0] + np.float32(t) * pa[1]
        return out
    segs, ts = sample_plan(h)
    n_seg = h - 1
    two, three, four, five, half = (np.float32(c) for c in (2, 3, 4, 5, 0.5))
    for k, (s, t) in enumerate(zip(segs, ts)):
        if s == 0:
            p0 = pa[0] - (pa[1] - pa[0])
            p1, p2, p3 = pa[0], pa[1], pa[min(2, h - 1)]
        elif s == n_seg - 1:
            p0, p1, p2 = pa[s - 1], pa[s], pa[s + 1]
            p3 = pa[s + 1] + (pa[s + 1] - pa[s])
        else:
            p0, p1, p2, p3 = pa[s - 1], pa[s], pa[s + 1], pa[min(s + 2, h - 1)]
        t1, t2, t3 = np.float32(t), np.float32(t * t), np.float32(t * t * t)
        a = (-p0 + p2) * t1
        b = (((two * p0 - five * p1) + four * p2) - p3) * t2
        c = (((-p0 + three * p1) - three * p2) + p3) * t3
        out[:, k] = half * (((two * p1 + a) + b) + c)
    return out


def history_trails(hist, position):
    """_add_trail_lines for every point.  hist: (h, N, 3) float32 transformed history positions
    (oldest first; the reference keeps the last 20), position: (N,3) float32 current positions.
    Returns ctrl (N, 21, 3) float32 — the control points of the curve file after its 6-decimal text
    round trip, first `count[i]` rows valid — and count (N,) int32 (0 = no trail drawn)."""
    hist = np.asarray(hist, np.float32)[-HISTORY_FRAMES:]
    position = np.asarray(position, np.float32)
    n = position.shape[0]
    ctrl = np.zeros((n, MAX_CTRL, 3), np.float32)
    count = np.zeros(n, np.int32)
    if hist.shape[0] < 2:
        return ctrl, count
    pts = np.concatenate([catmull_rom_samples(hist).astype(np.float64), position.astype(np.float64)[:, None, :]], axis=1)
    for i in range(n):
        p = pts[i][np.all(np.isfinite(pts[i]), axis=1)]
        if len(p) < 2:
            continue
        kept = [p[0]]
        for q in p[1:]:
            d = q - kept[-1]
            if np.sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]) > MIN_DISTANCE:
                kept.append(q)
        if len(kept) >= 2:
            d = kept[0] - kept[-1]
            if np.sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]) < MIN_DISTANCE:
                kept = kept[:-1]
        if len(kept) < 2:
            continue
        count[i] = len(kept)
        ctrl[i, :len(kept)] = round6(np.array(kept)).astype(np.float32)
    return ctrl, count


# ------------------------------------------------------------------- pixel half (raycast.c)
def ring_profile(verts, n_rings=N_RINGS, n_segments=N_SEGMENTS):
    """(n_rings+1, 4) float32 rows (ring radius, ring z, normal radial, normal z): the profile curve
    the mesh samples (vertex j = 0 of every ring lies in the xz half-plane) and its smooth normals —
    the normalised sum of the two adjacent band normals, the poles pointing along the axis."""
    v = np.asarray(verts, np.float64).reshape(n_rings + 1, n_segments, 3)
    pr, pz = v[:, 0, 0], v[:, 0, 2]
    dr, dz = np.diff(pr), np.diff(pz)
    bn = np.stack([-dz, dr], axis=1)
    ln = np.sqrt(bn[:, 0] ** 2 + bn[:, 1] ** 2)
    bn = np.where(ln[:, None] > 0, bn / np.where(ln > 0, ln, 1.0)[:, None], 0.0)
    rn = np.zeros((n_rings + 1, 2))
    rn[1:-1] = bn[:-1] + bn[1:]
    rn[0], rn[-1] = (0.0, 1.0), (0.0, -1.0)
    ln = np.sqrt(rn[:, 0] ** 2 + rn[:, 1] ** 2)
    rn = rn / np.where(ln > 0, ln, 1.0)[:, None]
    return np.ascontiguousarray(np.concatenate([pr[:, None], pz[:, None], rn], axis=1), np.float32)


def _lib():
    import ctypes
    from . import pcr_oracle
    L = pcr_oracle.lib()
    if not getattr(L, "_droplets_bound", False):
        vp, i32, i64, u32, f32 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32, ctypes.c_float
        L.orc_visibility_mesh.argtypes = [vp, i32, vp, i32, vp, i64, u32, vp, vp, ctypes.c_int]
        L.orc_visibility_mesh.restype = None
        L.orc_visibility_polylines.argtypes = [vp, vp, i64, i32, f32, u32, vp, vp, ctypes.c_int]
        L.orc_visibility_polylines.restype = None
        L.orc_shade_polylines.argtypes = [vp, vp, vp, i64, i32, u32, vp, vp, vp, vp]
        L.orc_shade_polylines.restype = None
        L.orc_shade_droplets.argtypes = [vp, vp, i64, u32, vp, i32, vp, vp, vp, vp]
        L.orc_shade_droplets.restype = None
        L._droplets_bound = True
    return L


def add_droplets(vis, xf, frame, id_base=0, brute_force=False, mesh=None):
    """Merge the droplet instances xf (N,12) f32 into the (H,W) uint64 key buffer; returns a new array."""
    import ctypes
    verts, faces = mesh if mesh is not None else droplet_mesh()
    verts = np.ascontiguousarray(verts, np.float32)
    faces = np.ascontiguousarray(faces, np.int32)
    xf = np.ascontiguousarray(xf, np.float32).reshape(-1, 12)
    out = np.ascontiguousarray(vis, np.uint64).copy()
    _lib().orc_visibility_mesh(verts.ctypes.data, verts.shape[0], faces.ctypes.data, faces.shape[0], xf.ctypes.data, xf.shape[0],
                               int(id_base), ctypes.addressof(frame), out.ctypes.data, 0 if brute_force else 1)
    return out


def add_polylines(vis, ctrl, count, frame, cap_id_base, radius=0.0007, brute_force=False):
    import ctypes
    ctrl = np.ascontiguousarray(ctrl, np.float32)
    count = np.ascontiguousarray(count, np.int32)
    out = np.ascontiguousarray(vis, np.uint64).copy()
    _lib().orc_visibility_polylines(ctrl.ctypes.data, count.ctypes.data, ctrl.shape[0], ctrl.shape[1], float(radius),
                                    int(cap_id_base), ctypes.addressof(frame), out.ctypes.data, 0 if brute_force else 1)
    return out


def shade_droplet_scene(vis, xf, ctrl, count, frame, scene, rgb=(0.3, 0.3, 0.3), trail_rgb=(0.2, 1.0, 0.4), mesh=None):
    """sRGB8 image of a droplet scene: floor / miss from orc_shade, then the polylines (ids N + i) and
    the droplets (ids i) on top."""
    import ctypes
    from . import pcr_oracle
    verts, _ = mesh if mesh is not None else droplet_mesh()
    xf = np.ascontiguousarray(xf, np.float32).reshape(-1, 12)
    n = xf.shape[0]
    vis = np.ascontiguousarray(vis, np.uint64)
    out = pcr_oracle.shade(vis, np.zeros((0, 4), np.float32), np.zeros((0, 4), np.float32), frame, scene)
    L = _lib()
    if ctrl is not None:
        ctrl = np.ascontiguousarray(ctrl, np.float32)
        count = np.ascontiguousarray(count, np.int32)
        c = (ctypes.c_float * 3)(*trail_rgb)
        L.orc_shade_polylines(vis.ctypes.data, ctrl.ctypes.data, count.ctypes.data, n, ctrl.shape[1], n, c,
                              ctypes.addressof(frame), ctypes.addressof(scene), out.ctypes.data)
    prof = ring_profile(verts)
    c = (ctypes.c_float * 3)(*rgb)
    L.orc_shade_droplets(vis.ctypes.data, xf.ctypes.data, n, 0, prof.ctypes.data, prof.shape[0] - 1, c,
                         ctypes.addressof(frame), ctypes.addressof(scene), out.ctypes.data)
    return out
