"""numpy / ctypes front of the CPU oracle — TEST INFRASTRUCTURE, not product code.

Host-side restatement (numpy, same operations in the same dtype as the reference) of
everything the reference does to a frame before it is handed to Mitsuba, plus ctypes
bindings to oracle/raycast.c for visibility and shading.  Each function cites the
reference lines it follows.  Pinned against the reference itself by
tests/test_oracle_vs_reference.py (runs where /root/reference exists) and against the
fixtures oracle/gen_golden.py wrote to tests/golden/ (runs everywhere).

The image half (raycast.c) is PARITY UNPINNED: Mitsuba is absent, the reference has no
golden images (see raycast.c header).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

ID_FLOOR = 0xFFFFFFFE
ID_MISS = 0xFFFFFFFF
KEY_MISS = 0x7F800000FFFFFFFF


# --------------------------------------------------------------------------- a1
def standardize_point_cloud(pcl, exact_mean=False):
    """example_renderer.py:94-98 (3 cols) / traj_ball_renderer.py:190-202 (3|6 cols).

    center = mean, scale = largest axis extent (one scalar), arithmetic in the input
    dtype, cast to f32 last; velocity columns are cast to f32 and passed through unscaled.

    exact_mean=False is the reference verbatim: np.mean(axis=0) of an (N,3) array is a plain
    SEQUENTIAL sum in the input dtype (verified: it equals a python loop `acc = acc + row`), so
    for float32 input its error grows with N and with |mean|/extent (2e-5 at N=1e5, mean 5).
    exact_mean=True is the order-independent definition the CUDA path implements: the mean is
    summed in float64 and rounded once to the input dtype; everything else is unchanged.
    """
    pcl = np.asarray(pcl)
    positions = pcl[:, :3]
    if exact_mean:
        center = (np.sum(positions, axis=0, dtype=np.float64) / positions.shape[0]).astype(positions.dtype)
    else:
        center = np.mean(positions, axis=0)
    scale = np.amax(positions - np.amin(positions, axis=0))
    normalized = ((positions - center) / scale).astype(np.float32)
    if pcl.shape[1] == 6:
        return np.column_stack([normalized, pcl[:, 3:6].astype(np.float32)])
    return normalized


# --------------------------------------------------------------------------- a2
def transform_coordinates(pcl, flip_x=True):
    """traj_ball_renderer.py:204-221 (flip_x) / traj_b0.py:62-82, traj_original.py:40-60
    (no flip) / inline example_renderer.py:171-173.  pos' = (-+z, x, y + 0.0125)."""
    pcl = np.asarray(pcl, dtype=np.float32)
    pos = pcl[:, [2, 0, 1]]
    if flip_x:
        pos[:, 0] *= -1
    pos[:, 2] += 0.0125
    if pcl.shape[1] == 6:
        vel = pcl[:, [5, 3, 4]]
        if flip_x:
            vel[:, 0] *= -1
        return np.column_stack([pos, vel])
    return pos


# --------------------------------------------------------------------------- a3
def compute_color(pcl, mode=0, const_rgb=(0.3, 0.3, 0.3), vel_norm=10.0, user_rgb=None):
    """Colour hook.  mode 0 = the reference (compute_color, example_renderer.py:89-92:
    constant grey).  mode 1 feeds the per-point normalisation the reference computes and
    then ignores (example_renderer.py:115-124, (p-min)/(range+1e-8)) to the PointFlow
    colormap (clip to [0.001,1], scale to unit length).  mode 2 ramps on
    min(|v|/vel_norm,1), the normalisation of traj_ball_renderer.py:134.  mode 3 = user.
    All f32.  Returns (N,4) f32: r,g,b,|v|."""
    pcl = np.asarray(pcl, dtype=np.float32)
    n = pcl.shape[0]
    out = np.zeros((n, 4), np.float32)
    if pcl.shape[1] == 6:
        v = pcl[:, 3:6]
        out[:, 3] = np.sqrt((v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]) + v[:, 2] * v[:, 2], dtype=np.float32)
    if mode == 0:
        out[:, :3] = np.asarray(const_rgb, np.float32)
    elif mode == 1:
        pos = pcl[:, :3]
        pmin = pos.min(axis=0)
        rng = (pos.max(axis=0) - pmin) + np.float32(1e-8)
        q = np.clip((pos - pmin) / rng, np.float32(0.001), np.float32(1.0)).astype(np.float32)
        nrm = np.sqrt((q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1]) + q[:, 2] * q[:, 2], dtype=np.float32)
        out[:, :3] = q / nrm[:, None]
    elif mode == 2:
        s = np.minimum(out[:, 3] / np.float32(vel_norm), np.float32(1.0))
        out[:, :3] = velocity_ramp(s)
    elif mode == 3:
        out[:, :3] = np.asarray(user_rgb, np.float32).reshape(n, 3)
    else:
        raise ValueError(mode)
    return out


_RAMP = np.array([[0.10, 0.25, 0.85], [0.95, 0.85, 0.25], [0.90, 0.15, 0.10]], np.float32)


def velocity_ramp(s):
    """3-stop piecewise-linear ramp slow(blue) -> mid(yellow) -> fast(red), f32."""
    s = np.asarray(s, np.float32)
    t = s * np.float32(2.0)
    lo = t < np.float32(1.0)
    f = np.where(lo, t, t - np.float32(1.0)).astype(np.float32)
    a = np.where(lo[:, None], _RAMP[0], _RAMP[1])
    b = np.where(lo[:, None], _RAMP[1], _RAMP[2])
    return (a + (b - a) * f[:, None]).astype(np.float32)


# --------------------------------------------------------------------------- a4
def _lerp3(a, b, p):
    return tuple(a[k] + (b[k] - a[k]) * p for k in range(3))


def camera_position(preset, frame_index=0, total_frames=220, last_motion_frame=199, fade_frames=20):
    """Per-frame eye position (python floats, f64) of every compute_camera_position:
    example_renderer.py:20 (fixed) ; traj_renderer.py:519-527 ; traj_ball_renderer.py:281-307
    = traj_vel_renderer.py:381-407 ; traj_original.py:62-66 ; traj_b0.py:84-115 ;
    traj_b1.py:84-115.  last_motion_frame / fade_frames are literals (199 / 20) in the
    reference; they are parameters here only so longer synthetic trajectories can reuse
    the schedule."""
    if preset == "example":
        return (2.2, 2.2, 4.2)
    if preset == "traj":
        progress = frame_index / max(total_frames - 1, 1)
        return (2.8 - 2.0 * progress, 2.8 - 2.0 * progress, 3.0 - 2.0 * progress)
    if preset == "traj_original":
        return (-1.8, -1.8, 1.8)
    keys = {
        "traj_ball": ((2.8, 2.8, 3.0), (1.8, 1.8, 1.8), (1.6, 1.6, 1.6)),
        "traj_vel": ((2.8, 2.8, 3.0), (1.8, 1.8, 1.8), (1.6, 1.6, 1.6)),
        "traj_b0": ((-2.2, -3.3, 2.0), (-1.3, -2.5, 0.8), (-1.0, -2.0, 0.7)),
        "traj_b1": ((-3.5, -2.5, 2.8), (-2.3, -1.5, 1.2), (-2.0, -1.2, 1.0)),
    }[preset]
    if frame_index <= last_motion_frame:
        return _lerp3(keys[0], keys[1], frame_index / max(last_motion_frame, 1))
    return _lerp3(keys[1], keys[2], (frame_index - last_motion_frame) / max(fade_frames, 1))


# --------------------------------------------------------------------------- 8f-1 trails
TRAIL_RADIUS = 0.0007                      # traj_ball_renderer.py:160
TRAIL_RGB = (0.2, 1.0, 0.4)                # traj_ball_renderer.py:179
TRAIL_LEN = (0.07, 0.3)                    # base / max trail length, traj_ball_renderer.py:132-133


def trail_length_scale(preset, frame_index):
    """length_scale of _add_velocity_trail: traj_ball_renderer.py:119-124 (ramp-in over frames 0-19),
    traj_vel_renderer.py:215-224 (ramp-in, then fade-out over 200-219), traj_original.py:78 /
    traj_b0.py:127 / traj_b1.py:127 (always 1).  Python floats, f64."""
    if preset in ("traj_original", "traj_b0", "traj_b1"):
        return 1.0
    if frame_index <= 19:
        return frame_index / 19.0
    if preset == "traj_vel" and frame_index > 199:
        return 1.0 - (frame_index - 199) / 20
    return 1.0


def round6(x, exact=False):
    """The `{:.6f}` text round trip of the reference's curve files (traj_ball_renderer.py:176).
    exact=True formats every value like python does; the default is rint(x*1e6)/1e6 in f64, which is
    what the CUDA path computes and agrees with the text round trip except on astronomically rare
    near-ties (tests assert agreement on their data)."""
    x = np.asarray(x, np.float64)
    if exact:
        return np.array([float(f"{v:.6f}") for v in x.ravel()], np.float64).reshape(x.shape)
    return np.rint(x * 1e6) / 1e6


def velocity_trails(pcl6, length_scale, exact_text=False):
    """_add_velocity_trail (traj_ball_renderer.py:98-188) for a transformed (N,6) f32 array: the
    straight trail from  position + (-v/|v|) * L  to  position,  L = (0.07 + 0.23*min(|v|/10,1)) *
    length_scale, all in f64, both ends then pass through the 6-decimal text file and are read as
    f32.  Returns tail (N,3) f32, head (N,3) f32, valid (N,) bool (|v| >= 1e-6 and length_scale > 0).
    The 19 interior control points are collinear (to 5e-7) and are not modelled."""
    pcl6 = np.asarray(pcl6, np.float32)
    pos = pcl6[:, :3].astype(np.float64)
    vel = pcl6[:, 3:6].astype(np.float64)
    vn = np.sqrt((vel[:, 0] * vel[:, 0] + vel[:, 1] * vel[:, 1]) + vel[:, 2] * vel[:, 2])
    valid = (vn >= 1e-6) & (length_scale > 0)
    safe = np.where(valid, vn, 1.0)
    length = (TRAIL_LEN[0] + (TRAIL_LEN[1] - TRAIL_LEN[0]) * np.minimum(vn / 10.0, 1.0)) * length_scale
    direction = -vel / safe[:, None]
    tail = pos + direction * length[:, None] * 1.0
    return (round6(tail, exact_text).astype(np.float32), round6(pos, exact_text).astype(np.float32), valid)


# scene constants per script: XMLTemplates HEAD/TAIL (all [R])
PRESETS = {
    "example": dict(target=(0.0, 0.0, 0.0), fov=30.0, flip_x=True, floor_z=-0.2,
                    floor_min=(-10.0, -10.0), floor_max=(10.0, 10.0)),          # example_renderer.py:20-22,56-62
    "traj": dict(target=(0.0, 0.0, -0.05), fov=36.0, flip_x=True, floor_z=-0.5,
                 floor_min=(-10.0, -10.0), floor_max=(10.0, 10.0)),             # traj_renderer.py:24-26,66-72
    "traj_ball": dict(target=(0.0, 0.0, -0.05), fov=36.0, flip_x=True, floor_z=-0.5,
                      floor_min=(-10.0, -10.0), floor_max=(10.0, 10.0)),        # traj_ball_renderer.py:17-19,59-65
    "traj_vel": dict(target=(0.0, 0.0, -0.05), fov=36.0, flip_x=True, floor_z=-0.5,
                     floor_min=(-10.0, -10.0), floor_max=(10.0, 10.0)),         # traj_vel_renderer.py:17-19,59-65
    "traj_original": dict(target=(0.0, 0.0, -0.05), fov=36.0, flip_x=False, floor_z=-0.5,
                          floor_min=(-10.0, -10.0), floor_max=(10.0, 10.0)),    # traj_original.py:19-21 (+ inherited TAIL)
    "traj_b0": dict(target=(-0.02, 0.15, -0.05), fov=36.0, flip_x=False, floor_z=-0.8,
                    floor_min=(-10.0, -10.0), floor_max=(30.0, 30.0)),          # traj_b0.py:19-21,42-48
    "traj_b1": dict(target=(0.0, -0.02, 0.0), fov=36.0, flip_x=False, floor_z=-0.8,
                    floor_min=(-10.0, -10.0), floor_max=(30.0, 30.0)),          # traj_b1.py:19-21,42-48
}


# ------------------------------------------------------------------- raycast.c binding
class Frame(ctypes.Structure):
    _fields_ = [("L", ctypes.c_float * 3), ("U", ctypes.c_float * 3), ("D", ctypes.c_float * 3),
                ("O", ctypes.c_float * 3), ("T", ctypes.c_float), ("Th", ctypes.c_float),
                ("TW", ctypes.c_float), ("near_clip", ctypes.c_float), ("far_clip", ctypes.c_float),
                ("W", ctypes.c_int32), ("H", ctypes.c_int32)]


class Scene(ctypes.Structure):
    _fields_ = [("has_floor", ctypes.c_int32), ("floor_z", ctypes.c_float),
                ("floor_min", ctypes.c_float * 2), ("floor_max", ctypes.c_float * 2),
                ("floor_albedo", ctypes.c_float), ("light_z", ctypes.c_float),
                ("light_half", ctypes.c_float), ("radiance", ctypes.c_float),
                ("bounce", ctypes.c_float)]


_lib = None


def build(force=False):
    """make -C oracle (gcc, a second or two)."""
    if force or not os.path.isfile(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "raycast.c")):
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        fp = ctypes.POINTER(ctypes.c_float)
        L.orc_camera_frame.argtypes = [fp, fp, fp, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                       ctypes.c_int, ctypes.c_int, ctypes.POINTER(Frame)]
        L.orc_camera_frame.restype = None
        L.orc_visibility.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint32, ctypes.POINTER(Frame),
                                     ctypes.POINTER(Scene), ctypes.c_void_p, ctypes.c_int]
        L.orc_visibility.restype = None
        L.orc_shade.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint32,
                                ctypes.c_int, ctypes.POINTER(Frame), ctypes.POINTER(Scene), ctypes.c_void_p]
        L.orc_shade.restype = None
        L.orc_visibility_caps.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint32, ctypes.POINTER(Frame),
                                          ctypes.c_void_p, ctypes.c_int]
        L.orc_visibility_caps.restype = None
        L.orc_shade_caps.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint32,
                                     ctypes.POINTER(ctypes.c_float), ctypes.POINTER(Frame), ctypes.POINTER(Scene), ctypes.c_void_p]
        L.orc_shade_caps.restype = None
        L.orc_num_threads.restype = ctypes.c_int
        L.orc_set_num_threads.argtypes = [ctypes.c_int]
        L.orc_set_num_threads.restype = None
        _lib = L
    return _lib


def camera_frame(origin, target, up, fov_x_deg, near_clip, far_clip, W, H):
    f = Frame()
    a = lambda v: (ctypes.c_float * 3)(*[float(np.float32(x)) for x in v])
    lib().orc_camera_frame(a(origin), a(target), a(up), float(fov_x_deg), float(near_clip), float(far_clip),
                           int(W), int(H), ctypes.byref(f))
    return f


def make_scene(has_floor=True, floor_z=-0.2, floor_min=(-10.0, -10.0), floor_max=(10.0, 10.0),
               floor_albedo=1.0, light_z=15.0, light_half=8.0, radiance=4.0, bounce=1.0):
    s = Scene()
    s.has_floor = int(has_floor)
    s.floor_z = floor_z
    s.floor_min = (ctypes.c_float * 2)(*floor_min)
    s.floor_max = (ctypes.c_float * 2)(*floor_max)
    s.floor_albedo = floor_albedo
    s.light_z = light_z
    s.light_half = light_half
    s.radiance = radiance
    s.bounce = bounce
    return s


def visibility(pos4, frame, scene, id_base=0, brute_force=False):
    """(H,W) uint64 keys = (f32 bits of camera-space depth << 32) | id."""
    pos4 = np.ascontiguousarray(pos4, np.float32).reshape(-1, 4)
    vis = np.empty((frame.H, frame.W), np.uint64)
    lib().orc_visibility(pos4.ctypes.data, pos4.shape[0], int(id_base), ctypes.byref(frame), ctypes.byref(scene),
                         vis.ctypes.data, 0 if brute_force else 1)
    return vis


def shade(vis, pos4, attr4, frame, scene, id_base=0, owner_only=False):
    pos4 = np.ascontiguousarray(pos4, np.float32).reshape(-1, 4)
    attr4 = np.ascontiguousarray(attr4, np.float32).reshape(-1, 4)
    vis = np.ascontiguousarray(vis, np.uint64)
    rgba = np.empty((frame.H, frame.W, 4), np.uint8)
    lib().orc_shade(vis.ctypes.data, pos4.ctypes.data, attr4.ctypes.data, pos4.shape[0], int(id_base),
                    int(owner_only), ctypes.byref(frame), ctypes.byref(scene), rgba.ctypes.data)
    return rgba


def _caps(tail, head, valid, radius):
    tail = np.asarray(tail, np.float32).reshape(-1, 3)
    head = np.asarray(head, np.float32).reshape(-1, 3)
    a4 = np.ascontiguousarray(np.concatenate([tail, np.full((len(tail), 1), radius, np.float32)], axis=1))
    b4 = np.ascontiguousarray(np.concatenate([head, np.asarray(valid, np.float32).reshape(-1, 1)], axis=1))
    return a4, b4


def add_trails(vis, tail, head, valid, frame, cap_id_base, radius=TRAIL_RADIUS, brute_force=False):
    """Merge the trails (capsules tail->head) into the (H,W) uint64 key buffer `vis`; trail j has
    id cap_id_base + j.  Returns a new array."""
    a4, b4 = _caps(tail, head, valid, radius)
    out = np.ascontiguousarray(vis, np.uint64).copy()
    lib().orc_visibility_caps(a4.ctypes.data, b4.ctypes.data, a4.shape[0], int(cap_id_base), ctypes.byref(frame),
                              out.ctypes.data, 0 if brute_force else 1)
    return out


def shade_trails(rgba, vis, tail, head, valid, frame, scene, cap_id_base, radius=TRAIL_RADIUS, rgb=TRAIL_RGB):
    a4, b4 = _caps(tail, head, valid, radius)
    out = np.ascontiguousarray(rgba, np.uint8).copy()
    vis = np.ascontiguousarray(vis, np.uint64)
    c = (ctypes.c_float * 3)(*rgb)
    lib().orc_shade_caps(vis.ctypes.data, a4.ctypes.data, b4.ctypes.data, a4.shape[0], int(cap_id_base), c,
                         ctypes.byref(frame), ctypes.byref(scene), out.ctypes.data)
    return out


def set_num_threads(n=None):
    """Use n OpenMP threads (default: every host CPU) regardless of OMP_NUM_THREADS."""
    lib().orc_set_num_threads(int(n or os.cpu_count() or 1))
    return num_threads()


def num_threads():
    return int(lib().orc_num_threads())


# --------------------------------------------------- independent f64 cross-check (small cases)
def visibility_f64(pos4, origin, target, up, fov_x_deg, near_clip, far_clip, W, H, scene=None):
    """Textbook double-precision ray-sphere nearest hit with NORMALISED world-space rays —
    a different formulation from raycast.c on purpose.  Returns (ids (H,W) uint32,
    depth (H,W) f64 camera-space z, edge (H,W) = min over spheres of |disc|/r^2 (distance
    from a silhouette), gap (H,W) = relative depth gap between the two nearest hits), so
    tests can exclude knife-edge pixels.  O(W*H*N) numpy: small cases only."""
    pos4 = np.asarray(pos4, np.float64).reshape(-1, 4)
    o = np.asarray(origin, np.float64)
    d = np.asarray(target, np.float64) - o
    d /= np.linalg.norm(d)
    left = np.cross(np.asarray(up, np.float64), d)
    left /= np.linalg.norm(left)
    newup = np.cross(d, left)
    T = np.tan(np.deg2rad(fov_x_deg) / 2.0)
    ii = (np.arange(W) + 0.5) / W
    jj = (np.arange(H) + 0.5) / H
    u = (1.0 - 2.0 * ii) * T
    w = (1.0 - 2.0 * jj) * T * H / W
    dirs = u[None, :, None] * left + w[:, None, None] * newup + d  # (H,W,3), z_cam = 1
    vlen = np.linalg.norm(dirs, axis=2)
    nd = dirs / vlen[..., None]
    ids = np.full((H, W), ID_MISS, np.uint32)
    depth = np.full((H, W), np.inf)
    margin = np.full((H, W), np.inf)
    if scene is not None and scene.has_floor:
        tz = (scene.floor_z - o[2]) / dirs[..., 2]
        hx = o[0] + tz * dirs[..., 0]
        hy = o[1] + tz * dirs[..., 1]
        ok = (tz >= near_clip) & (tz <= far_clip) & (hx >= scene.floor_min[0]) & (hx <= scene.floor_max[0]) \
            & (hy >= scene.floor_min[1]) & (hy <= scene.floor_max[1])
        depth = np.where(ok, tz, depth)
        ids = np.where(ok, np.uint32(ID_FLOOR), ids)
    second = np.full((H, W), np.inf)
    for k in range(pos4.shape[0]):
        c = pos4[k, :3] - o
        r = pos4[k, 3]
        b = nd @ c
        q2 = c @ c - b * b
        disc = r * r - q2
        hit = disc >= 0
        t = b - np.sqrt(np.where(hit, disc, 0.0))
        zc = t / vlen  # camera-space depth since dirs has z_cam = 1
        hit &= (zc >= near_clip) & (zc <= far_clip)
        margin = np.minimum(margin, np.abs(disc) / (r * r))
        closer = hit & (zc < depth)
        second = np.where(closer, depth, np.where(hit, np.minimum(second, zc), second))
        depth = np.where(closer, zc, depth)
        ids = np.where(closer, np.uint32(k), ids)
    gap = (second - depth) / np.maximum(depth, 1e-30)
    return ids, depth, margin, gap
