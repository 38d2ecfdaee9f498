"""ctypes binding of libpcr.so — the C ABI declared in include/pcr.h.

There is no CPU or PyTorch fallback: if the extension is missing or no CUDA device is
present, the calls raise.  torch is used only for device memory and streams.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PCR_LIBPCR") or os.path.join(_HERE, "libpcr.so")      # PCR_LIBPCR: a diagnostics build (A/B runs)

ID_FLOOR = 0xFFFFFFFE
ID_MISS = 0xFFFFFFFF
KEY_MISS = 0x7F800000FFFFFFFF

COLOR_CONST, COLOR_POSITION, COLOR_VELOCITY, COLOR_USER = 0, 1, 2, 3
MEAN_AUTO, MEAN_SEQUENTIAL, MEAN_F64 = 0, 1, 2     # pcr_style.mean_mode (AUTO = the reference's sequential mean for whole frames)

# every symbol include/pcr.h declares (tests check the library exports all of them)
SYMBOLS = (
    "pcr_abi_version", "pcr_create", "pcr_destroy", "pcr_last_error", "pcr_camera_frame",
    "pcr_standardize", "pcr_render", "pcr_shade", "pcr_render_frames", "pcr_render_frames_host",
    "pcr_zmin", "pcr_zmerge_nccl", "pcr_stats_partial", "pcr_standardize_with_stats", "pcr_counters",
    "pcr_transform_coordinates", "pcr_profile", "pcr_profile_read", "pcr_kernel_name", "pcr_set_occlusion", "pcr_finalize_stats", "pcr_velocity_trails", "pcr_render_shard", "pcr_shade_shard",
    "pcr_set_droplet_mesh", "pcr_droplet_transforms", "pcr_history_trails", "pcr_render_droplet_frames",
    "pcr_peer_alloc", "pcr_ipc_export", "pcr_ipc_open", "pcr_ipc_close", "pcr_peer_set", "pcr_peer_begin_frame",
    "pcr_render_shard_peer", "pcr_shade_shard_peer", "pcr_selftest_scale_div", "pcr_render_transformed", "pcr_prefetch_frames",
    "pcr_render_frames_host_submit", "pcr_host_wait",
)
HISTORY_FRAMES, MAX_CTRL = 20, 21                   # PCR_HISTORY_FRAMES, PCR_MAX_CTRL
TRAILS_NONE, TRAILS_VELOCITY, TRAILS_HISTORY = 0, 1, 2


class Camera(ctypes.Structure):
    """pcr_camera — the <sensor> block of XMLTemplates.HEAD (example_renderer.py:16-31)."""
    _fields_ = [("origin", ctypes.c_float * 3), ("target", ctypes.c_float * 3), ("up", ctypes.c_float * 3),
                ("fov_x_deg", ctypes.c_float), ("near_clip", ctypes.c_float), ("far_clip", ctypes.c_float),
                ("width", ctypes.c_int32), ("height", ctypes.c_int32), ("trail_scale", ctypes.c_double)]


class Style(ctypes.Structure):
    """pcr_style — BALL_SEGMENT / TAIL constants (example_renderer.py:41-72) + transform flavour."""
    _fields_ = [("color_mode", ctypes.c_int32), ("const_rgb", ctypes.c_float * 3), ("radius", ctypes.c_float),
                ("flip_x", ctypes.c_int32), ("z_lift", ctypes.c_float), ("vel_norm", ctypes.c_float),
                ("has_floor", ctypes.c_int32), ("floor_z", ctypes.c_float), ("floor_min", ctypes.c_float * 2),
                ("floor_max", ctypes.c_float * 2), ("floor_albedo", ctypes.c_float), ("light_z", ctypes.c_float),
                ("light_half", ctypes.c_float), ("radiance", ctypes.c_float), ("bounce", ctypes.c_float),
                ("xform", ctypes.c_int32), ("mean_mode", ctypes.c_int32), ("trails", ctypes.c_int32),
                ("trail_radius", ctypes.c_float), ("trail_rgb", ctypes.c_float * 3),
                ("trail_len_min", ctypes.c_double), ("trail_len_max", ctypes.c_double)]


class Frame(ctypes.Structure):
    """pcr_frame — derived f32 camera frame."""
    _fields_ = [("L", ctypes.c_float * 3), ("U", ctypes.c_float * 3), ("D", ctypes.c_float * 3),
                ("O", ctypes.c_float * 3), ("T", ctypes.c_float), ("Th", ctypes.c_float), ("TW", ctypes.c_float),
                ("near_clip", ctypes.c_float), ("far_clip", ctypes.c_float), ("W", ctypes.c_int32), ("H", ctypes.c_int32)]


_lib = None


def load_library():
    """dlopen libpcr.so and declare the prototypes.  Raises if the extension was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -m pointcloud_render_b200.build` "
            "(there is no CPU fallback for the render path)")
    L = ctypes.CDLL(LIB_PATH)
    vp, i64, i32, u32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_uint32
    camp, styp = ctypes.POINTER(Camera), ctypes.POINTER(Style)
    L.pcr_abi_version.restype = i32
    L.pcr_create.argtypes = [ctypes.POINTER(vp), i32, i64, i32, i32, i32, i64]
    L.pcr_destroy.argtypes = [vp]
    L.pcr_destroy.restype = None
    L.pcr_last_error.argtypes = [vp]
    L.pcr_last_error.restype = ctypes.c_char_p
    L.pcr_camera_frame.argtypes = [camp, ctypes.POINTER(Frame)]
    L.pcr_standardize.argtypes = [vp, vp, i32, i64, i32, vp, vp, styp, vp, vp, vp, vp, vp]
    L.pcr_render.argtypes = [vp, vp, vp, i64, u32, camp, styp, vp, vp, vp]
    L.pcr_shade.argtypes = [vp, vp, vp, vp, i64, u32, i32, camp, styp, vp, vp]
    L.pcr_render_transformed.argtypes = [vp, vp, i64, i32, vp, vp, camp, styp, vp, vp, vp]
    L.pcr_render_frames.argtypes = [vp, vp, i32, i64, i32, i32, vp, vp, camp, styp, vp, vp, vp]
    L.pcr_render_frames_host_submit.argtypes = [vp, vp, i32, i64, i32, i32, vp, vp, camp, styp, vp, vp, ctypes.POINTER(i64)]
    L.pcr_host_wait.argtypes = [vp, i64]
    L.pcr_prefetch_frames.argtypes = [vp, vp, i32, i64, i32, i32, styp, vp]
    L.pcr_render_frames_host.argtypes = [vp, vp, i32, i64, i32, i32, vp, vp, camp, styp, vp, vp]
    L.pcr_zmin.argtypes = [vp, vp, vp, i64, vp]
    L.pcr_zmerge_nccl.argtypes = [vp, vp, i64, vp, vp]
    L.pcr_stats_partial.argtypes = [vp, vp, i32, i64, i32, vp, vp]
    L.pcr_standardize_with_stats.argtypes = [vp, vp, i32, i64, i32, vp, vp, styp, vp, vp, vp, vp, vp]
    L.pcr_counters.argtypes = [vp, ctypes.POINTER(i64), vp]
    L.pcr_selftest_scale_div.argtypes = [vp, ctypes.POINTER(ctypes.c_float), i32, ctypes.POINTER(ctypes.c_uint64)]
    L.pcr_transform_coordinates.argtypes = [vp, vp, i64, i32, i32, ctypes.c_float, vp, vp]
    L.pcr_set_occlusion.argtypes = [vp, i32, i32, i64]
    L.pcr_finalize_stats.argtypes = [vp, vp, i32, i64, i32, vp, vp]
    L.pcr_render_shard.argtypes = [vp, vp, i32, i64, i32, vp, vp, vp, u32, camp, styp, vp, vp]
    L.pcr_shade_shard.argtypes = [vp, vp, vp, i32, i64, i32, vp, vp, vp, u32, i32, camp, styp, vp, vp]
    L.pcr_velocity_trails.argtypes = [vp, vp, i64, styp, ctypes.c_double, vp, vp, vp, vp]
    L.pcr_set_droplet_mesh.argtypes = [vp, vp, i32, i32]
    L.pcr_droplet_transforms.argtypes = [vp, vp, i64, i32, vp, vp, vp]
    L.pcr_history_trails.argtypes = [vp, vp, i32, vp, i64, vp, vp, vp]
    L.pcr_render_droplet_frames.argtypes = [vp, vp, i32, i64, i32, i32, i32, vp, camp, styp, vp, vp, vp]
    L.pcr_peer_alloc.argtypes = [vp, i32, i32, ctypes.POINTER(vp), ctypes.POINTER(vp)]
    L.pcr_ipc_export.argtypes = [vp, vp, vp]
    L.pcr_ipc_open.argtypes = [vp, vp, ctypes.POINTER(vp)]
    L.pcr_ipc_close.argtypes = [vp, vp]
    L.pcr_peer_set.argtypes = [vp, i32, i32, i32, ctypes.POINTER(vp), ctypes.POINTER(vp)]
    L.pcr_peer_begin_frame.argtypes = [vp, camp, styp, vp]
    L.pcr_render_shard_peer.argtypes = [vp, vp, i32, i64, i32, vp, vp, vp, u32, camp, styp, vp, vp]
    L.pcr_shade_shard_peer.argtypes = [vp, vp, vp, i32, i64, i32, vp, vp, vp, u32, camp, styp, vp]
    L.pcr_profile.argtypes = [vp, i32]
    L.pcr_profile_read.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(i64), i32]
    L.pcr_kernel_name.argtypes = [i32]
    L.pcr_kernel_name.restype = ctypes.c_char_p
    for name in SYMBOLS:
        fn = getattr(L, name)
        if name not in ("pcr_destroy", "pcr_last_error", "pcr_kernel_name"):
            fn.restype = i32
    _lib = L
    return L


def make_camera(origin, target, up=(0.0, 0.0, 1.0), fov_x_deg=30.0, near_clip=0.1, far_clip=100.0,
                width=1920, height=1080, trail_scale=0.0):
    c = Camera()
    c.origin = (ctypes.c_float * 3)(*[float(x) for x in origin])
    c.target = (ctypes.c_float * 3)(*[float(x) for x in target])
    c.up = (ctypes.c_float * 3)(*[float(x) for x in up])
    c.fov_x_deg, c.near_clip, c.far_clip = float(fov_x_deg), float(near_clip), float(far_clip)
    c.width, c.height = int(width), int(height)
    c.trail_scale = float(trail_scale)
    return c


def make_style(color_mode=COLOR_CONST, const_rgb=(0.3, 0.3, 0.3), radius=0.01, flip_x=True, z_lift=0.0125,
               vel_norm=10.0, has_floor=True, floor_z=-0.2, floor_min=(-10.0, -10.0), floor_max=(10.0, 10.0),
               floor_albedo=1.0, light_z=15.0, light_half=8.0, radiance=4.0, bounce=1.0, xform=0, mean_mode=MEAN_AUTO,
               trails=False, trail_radius=0.0007, trail_rgb=(0.2, 1.0, 0.4), trail_len_min=0.07, trail_len_max=0.3):
    s = Style()
    s.color_mode = int(color_mode)
    s.const_rgb = (ctypes.c_float * 3)(*const_rgb)
    s.radius, s.flip_x, s.z_lift, s.vel_norm = float(radius), int(bool(flip_x)), float(z_lift), float(vel_norm)
    s.has_floor, s.floor_z = int(bool(has_floor)), float(floor_z)
    s.floor_min = (ctypes.c_float * 2)(*floor_min)
    s.floor_max = (ctypes.c_float * 2)(*floor_max)
    s.floor_albedo, s.light_z, s.light_half = float(floor_albedo), float(light_z), float(light_half)
    s.radiance, s.bounce, s.xform, s.mean_mode = float(radiance), float(bounce), int(xform), int(mean_mode)
    s.trails, s.trail_radius = int(trails), float(trail_radius)     # True -> 1 (velocity trail), 2 = history trail
    s.trail_rgb = (ctypes.c_float * 3)(*trail_rgb)
    s.trail_len_min, s.trail_len_max = float(trail_len_min), float(trail_len_max)
    return s


def camera_frame(cam):
    f = Frame()
    rc = load_library().pcr_camera_frame(ctypes.byref(cam), ctypes.byref(f))
    if rc != 0:
        raise ValueError("degenerate camera")
    return f


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _check_points(t, cuda=True, dims=2, f32_only=False, what="points"):
    """Every wrapper hands raw pointers to C: a wrong dtype, a strided view (cloud[:, :3] of an (N,6) tensor) or
    a tensor on the wrong device would be silently reinterpreted / read out of bounds.  Raises instead."""
    import torch
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{what}: expected a torch tensor, got {type(t).__name__}")
    ok = (torch.float32,) if f32_only else (torch.float32, torch.float64)
    if t.dtype not in ok:
        raise TypeError(f"{what}: dtype must be {' or '.join(str(d) for d in ok)}, got {t.dtype}")
    if t.is_cuda != bool(cuda):
        raise ValueError(f"{what}: must be a {'CUDA' if cuda else 'CPU'} tensor")
    if t.dim() != dims or t.shape[-1] not in (3, 6):
        raise ValueError(f"{what}: shape must be {'(F,N,3|6)' if dims == 3 else '(N,3|6)'}, got {tuple(t.shape)}")
    if not t.is_contiguous():
        raise ValueError(f"{what}: must be contiguous (got strides {t.stride()}); call .contiguous()")
    return t


def _check_vector(t, numel, cuda=True, dtype=None, what="array"):
    """Optional per-point arrays (radius [n], rgb [n][3], stats [10] ...): dtype, size, device, contiguity."""
    import torch
    if t is None:
        return None
    dtype = dtype or torch.float32
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{what}: expected a torch tensor, got {type(t).__name__}")
    if t.dtype != dtype or t.numel() != int(numel) or t.is_cuda != bool(cuda) or not t.is_contiguous():
        raise ValueError(f"{what}: expected a contiguous {'CUDA' if cuda else 'CPU'} {dtype} tensor of {int(numel)} elements, "
                         f"got {t.dtype} x {t.numel()} on {t.device}")
    return t


def _stream_ptr(stream):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return ctypes.c_void_p(s.cuda_stream)


class Context:
    """Owns one pcr_ctx (scratch for one GPU).  Arguments are torch CUDA tensors."""

    def __init__(self, device=0, max_points=1 << 20, max_w=1920, max_h=1080, max_batch=1, pair_capacity=0):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("pcr needs a CUDA device (no CPU fallback)")
        self.lib = load_library()
        self.device = int(device)
        self.max_points, self.max_w, self.max_h, self.max_batch = int(max_points), int(max_w), int(max_h), int(max_batch)
        h = ctypes.c_void_p()
        rc = self.lib.pcr_create(ctypes.byref(h), self.device, self.max_points, self.max_w, self.max_h, self.max_batch,
                                 int(pair_capacity))
        if rc != 0:
            raise RuntimeError(f"pcr_create failed with status {rc}")
        self.handle = h

    def close(self):
        if getattr(self, "handle", None):
            self.lib.pcr_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            msg = self.lib.pcr_last_error(self.handle)
            raise RuntimeError(f"pcr error {rc}: {msg.decode() if msg else ''}")

    # ---- K0 + K1 -------------------------------------------------------------------------
    def standardize(self, pts, style, radius=None, rgb=None, want_vel=False, want_stats=False, stream=None):
        """pts: (N,3|6) float32/float64 CUDA tensor -> pos4 (N,4), attr4 (N,4)[, vel4][, stats(10)]."""
        import torch
        _check_points(pts)
        n, cols = pts.shape
        is64 = pts.dtype == torch.float64
        _check_vector(radius, n, what="radius"); _check_vector(rgb, 3 * n, what="rgb")
        dev = pts.device
        pos = torch.empty((n, 4), dtype=torch.float32, device=dev)
        attr = torch.empty((n, 4), dtype=torch.float32, device=dev)
        vel = torch.empty((n, 4), dtype=torch.float32, device=dev) if (want_vel and cols == 6) else None
        stats = torch.empty(10, dtype=torch.float64, device=dev) if want_stats else None
        self._check(self.lib.pcr_standardize(self.handle, _ptr(pts), int(is64), n, cols, _ptr(radius), _ptr(rgb),
                                             ctypes.byref(style), _ptr(pos), _ptr(attr), _ptr(vel), _ptr(stats),
                                             _stream_ptr(stream)))
        out = [pos, attr]
        if want_vel:
            out.append(vel)
        if want_stats:
            out.append(stats)
        return tuple(out)

    def transform_coordinates(self, pcl, flip_x=True, z_lift=0.0125, stream=None):
        """(N,3|6) float32 CUDA tensor -> transformed copy (transform_coordinates alone)."""
        import torch
        _check_points(pcl, f32_only=True)
        n, cols = pcl.shape
        out = torch.empty_like(pcl)
        self._check(self.lib.pcr_transform_coordinates(self.handle, _ptr(pcl), n, cols, int(bool(flip_x)), float(z_lift),
                                                       _ptr(out), _stream_ptr(stream)))
        return out

    def velocity_trails(self, pcl6, style, trail_scale, stream=None):
        """(N,6) float32 CUDA tensor (transformed) -> tail (N,3), head (N,3) float32, valid (N,) uint8."""
        import torch
        _check_points(pcl6, f32_only=True)
        if pcl6.shape[1] != 6:
            raise ValueError("velocity_trails needs (N,6) points")
        n = pcl6.shape[0]
        tail = torch.empty((n, 3), dtype=torch.float32, device=pcl6.device)
        head = torch.empty((n, 3), dtype=torch.float32, device=pcl6.device)
        valid = torch.empty(n, dtype=torch.uint8, device=pcl6.device)
        self._check(self.lib.pcr_velocity_trails(self.handle, _ptr(pcl6), n, ctypes.byref(style), float(trail_scale), _ptr(tail),
                                                 _ptr(head), _ptr(valid), _stream_ptr(stream)))
        return tail, head, valid

    def stats_partial(self, pts, stream=None):
        import torch
        _check_points(pts)
        n, cols = pts.shape
        out = torch.empty(10, dtype=torch.float64, device=pts.device)
        self._check(self.lib.pcr_stats_partial(self.handle, _ptr(pts), int(pts.dtype == torch.float64), n, cols,
                                               _ptr(out), _stream_ptr(stream)))
        return out[:9]

    def finalize_stats(self, partials, n_total, is_f64=False, stream=None):
        """partials: (k,9) float64 CUDA tensor of shard totals -> stats (10,) on the device, no host sync."""
        import torch
        assert partials.is_cuda and partials.dtype == torch.float64 and partials.is_contiguous()
        out = torch.empty(10, dtype=torch.float64, device=partials.device)
        self._check(self.lib.pcr_finalize_stats(self.handle, _ptr(partials), int(partials.numel() // 9), int(n_total), int(bool(is_f64)),
                                                _ptr(out), _stream_ptr(stream)))
        return out

    def standardize_with_stats(self, pts, style, stats10, radius=None, rgb=None, stream=None, out=None):
        """out = (pos4, attr4) reuses caller buffers (no allocation per call)."""
        import torch
        _check_points(pts)
        n, cols = pts.shape
        _check_vector(stats10, 10, dtype=torch.float64, what="stats10")
        _check_vector(radius, n, what="radius"); _check_vector(rgb, 3 * n, what="rgb")
        pos = out[0] if out is not None else torch.empty((n, 4), dtype=torch.float32, device=pts.device)
        attr = out[1] if out is not None else torch.empty((n, 4), dtype=torch.float32, device=pts.device)
        self._check(self.lib.pcr_standardize_with_stats(self.handle, _ptr(pts), int(pts.dtype == torch.float64), n, cols,
                                                        _ptr(radius), _ptr(rgb), ctypes.byref(style), _ptr(stats10),
                                                        _ptr(pos), _ptr(attr), None, _stream_ptr(stream)))
        return pos, attr

    # ---- K2..K4 ----------------------------------------------------------------------------
    def render(self, pos4, attr4, cam, style, id_base=0, shade=True, stream=None, out_vis=None, out_rgba=None):
        """pos4/attr4: (N,4) float32 CUDA -> vis (H,W) int64 view of the uint64 keys, rgba (H,W,4) uint8."""
        import torch
        n = pos4.shape[0]
        _check_vector(pos4, 4 * n, what="pos4"); _check_vector(attr4, 4 * n, what="attr4")
        dev = pos4.device
        vis = out_vis if out_vis is not None else torch.empty((cam.height, cam.width), dtype=torch.int64, device=dev)
        rgba = (out_rgba if out_rgba is not None else torch.empty((cam.height, cam.width, 4), dtype=torch.uint8, device=dev)) if shade else None
        self._check(self.lib.pcr_render(self.handle, _ptr(pos4) if n else None, _ptr(attr4) if n else None, n, int(id_base),
                                        ctypes.byref(cam), ctypes.byref(style), _ptr(vis), _ptr(rgba), _stream_ptr(stream)))
        return vis, rgba

    def render_transformed(self, pcl, cam, style, radius=None, rgb=None, want_vis=True, stream=None):
        """pcl: (N,3|6) float32 CUDA tensor, already standardised + transformed (what generate_xml_content gets) ->
        vis (H,W) int64, rgba (H,W,4) uint8.  6 columns + style.trails: spheres AND their velocity trails."""
        import torch
        _check_points(pcl, f32_only=True)
        n, cols = pcl.shape
        _check_vector(radius, n, what="radius"); _check_vector(rgb, 3 * n, what="rgb")
        vis = torch.empty((cam.height, cam.width), dtype=torch.int64, device=pcl.device) if want_vis else None
        rgba = torch.empty((cam.height, cam.width, 4), dtype=torch.uint8, device=pcl.device)
        self._check(self.lib.pcr_render_transformed(self.handle, _ptr(pcl) if n else None, n, cols, _ptr(radius), _ptr(rgb),
                                                    ctypes.byref(cam), ctypes.byref(style), _ptr(vis), _ptr(rgba), _stream_ptr(stream)))
        return vis, rgba

    def shade(self, vis, pos4, attr4, cam, style, id_base=0, owner_only=False, stream=None, out_rgba=None):
        import torch
        rgba = out_rgba if out_rgba is not None else torch.empty((cam.height, cam.width, 4), dtype=torch.uint8, device=vis.device)
        n = pos4.shape[0]
        _check_vector(pos4, 4 * n, what="pos4"); _check_vector(attr4, 4 * n, what="attr4")
        _check_vector(vis, cam.height * cam.width, dtype=torch.int64, what="vis")
        self._check(self.lib.pcr_shade(self.handle, _ptr(vis), _ptr(pos4) if n else None, _ptr(attr4) if n else None, n,
                                       int(id_base), int(owner_only), ctypes.byref(cam), ctypes.byref(style), _ptr(rgba),
                                       _stream_ptr(stream)))
        return rgba

    def render_shard(self, pts, stats10, cam, style, id_base=0, radius=None, rgb=None, out_vis=None, stream=None):
        """One shard of a point-sharded cloud, fused path: raw (n, 3|6) points + global stats -> vis keys."""
        import torch
        _check_points(pts)
        n, cols = pts.shape
        _check_vector(stats10, 10, dtype=torch.float64, what="stats10")
        _check_vector(radius, n, what="radius"); _check_vector(rgb, 3 * n, what="rgb")
        vis = out_vis if out_vis is not None else torch.empty((cam.height, cam.width), dtype=torch.int64, device=pts.device)
        self._check(self.lib.pcr_render_shard(self.handle, _ptr(pts) if n else None, int(pts.dtype == torch.float64), n, cols, _ptr(radius),
                                              _ptr(rgb), _ptr(stats10), int(id_base), ctypes.byref(cam), ctypes.byref(style), _ptr(vis),
                                              _stream_ptr(stream)))
        return vis

    def shade_shard(self, vis, pts, stats10, cam, style, id_base=0, owner_only=True, radius=None, rgb=None, out_rgba=None, stream=None):
        import torch
        _check_points(pts)
        n, cols = pts.shape
        _check_vector(stats10, 10, dtype=torch.float64, what="stats10")
        _check_vector(vis, cam.height * cam.width, dtype=torch.int64, what="vis")
        _check_vector(radius, n, what="radius"); _check_vector(rgb, 3 * n, what="rgb")
        rgba = out_rgba if out_rgba is not None else torch.empty((cam.height, cam.width, 4), dtype=torch.uint8, device=vis.device)
        self._check(self.lib.pcr_shade_shard(self.handle, _ptr(vis), _ptr(pts) if n else None, int(pts.dtype == torch.float64), n, cols,
                                             _ptr(radius), _ptr(rgb), _ptr(stats10), int(id_base), int(owner_only), ctypes.byref(cam),
                                             ctypes.byref(style), _ptr(rgba), _stream_ptr(stream)))
        return rgba

    def render_frames(self, traj, cams, style, radius=None, rgb=None, want_vis=False, out_rgba=None, out_vis=None,
                      stream=None):
        """traj: (F,N,3|6) CUDA tensor; cams: list of Camera -> rgba (F,H,W,4) uint8 [, vis (F,H,W) int64]."""
        import torch
        _check_points(traj, dims=3, what="traj")
        F, n, cols = traj.shape
        if len(cams) != F:
            raise ValueError("one camera per frame")
        _check_vector(radius, n, what="radius"); _check_vector(rgb, 3 * n, what="rgb")
        if F == 0:
            return (torch.empty((0, 0, 0, 4), dtype=torch.uint8, device=traj.device), None) if want_vis else \
                torch.empty((0, 0, 0, 4), dtype=torch.uint8, device=traj.device)
        W, H = cams[0].width, cams[0].height
        cam_arr = (Camera * F)(*cams)
        rgba = out_rgba if out_rgba is not None else torch.empty((F, H, W, 4), dtype=torch.uint8, device=traj.device)
        vis = out_vis if out_vis is not None else (
            torch.empty((F, H, W), dtype=torch.int64, device=traj.device) if want_vis else None)
        self._check(self.lib.pcr_render_frames(self.handle, _ptr(traj), int(traj.dtype == torch.float64), n, cols, F,
                                               _ptr(radius), _ptr(rgb), cam_arr, ctypes.byref(style), _ptr(vis),
                                               _ptr(rgba), _stream_ptr(stream)))
        return (rgba, vis) if (want_vis or out_vis is not None) else rgba

    def prefetch_frames(self, traj, style, stream=None):
        """Hint (pcr_prefetch_frames): these (F,N,3|6) CUDA frames will be passed to render_frames later, unchanged —
        their standardisation statistics (incl. the serial reference-exact mean) are computed now on side streams."""
        import torch
        _check_points(traj, dims=3, what="traj")
        F, n, cols = traj.shape
        self._check(self.lib.pcr_prefetch_frames(self.handle, _ptr(traj) if F else None, int(traj.dtype == torch.float64), n, cols, F,
                                                 ctypes.byref(style), _stream_ptr(stream)))

    def render_frames_host(self, traj_host, cams, style, radius_host=None, rgb_host=None, out_rgba=None, out_vis=None):
        """Host-buffer entry (what a reference script would call): traj_host is a CPU tensor or
        numpy array (F,N,3|6), ideally pinned; returns rgba as a CPU tensor (F,H,W,4).  Synchronous."""
        import torch
        t = torch.as_tensor(traj_host)
        ticket, rgba = self.render_frames_host_submit(t, cams, style, radius_host, rgb_host, out_rgba, out_vis)
        self.host_wait(ticket)
        self.host_wait(-1)
        return rgba

    def render_frames_host_submit(self, traj_host, cams, style, radius_host=None, rgb_host=None, out_rgba=None, out_vis=None):
        """Asynchronous host-buffer entry (pcr_render_frames_host_submit): enqueues the copies and kernels and returns
        (ticket, rgba); the images are in `rgba` once host_wait(ticket) has returned.  The buffers (and this call's
        arguments) are kept alive until then."""
        import torch
        t = _check_points(torch.as_tensor(traj_host), cuda=False, dims=3, what="traj_host")
        F, n, cols = t.shape
        if len(cams) != F:
            raise ValueError("one camera per frame")
        radius_host = None if radius_host is None else _check_vector(torch.as_tensor(radius_host), n, cuda=False, what="radius_host")
        rgb_host = None if rgb_host is None else _check_vector(torch.as_tensor(rgb_host), 3 * n, cuda=False, what="rgb_host")
        W, H = (cams[0].width, cams[0].height) if F else (0, 0)
        cam_arr = (Camera * max(F, 1))(*cams)
        rgba = out_rgba if out_rgba is not None else torch.empty((F, H, W, 4), dtype=torch.uint8).pin_memory()
        hp = lambda x: None if x is None else ctypes.c_void_p(torch.as_tensor(x).data_ptr())
        ticket = ctypes.c_int64(-1)
        self._check(self.lib.pcr_render_frames_host_submit(self.handle, hp(t), int(t.dtype == torch.float64), n, cols, F,
                                                           hp(radius_host), hp(rgb_host), cam_arr, ctypes.byref(style),
                                                           hp(out_vis), hp(rgba), ctypes.byref(ticket)))
        if not hasattr(self, "_inflight"):
            self._inflight = {}
        self._inflight[int(ticket.value)] = (t, radius_host, rgb_host, rgba, out_vis, cam_arr, style)
        return int(ticket.value), rgba

    def host_wait(self, ticket=-1):
        """Block until the host-buffer call `ticket` (or, with -1, every submitted call) has delivered its images."""
        self._check(self.lib.pcr_host_wait(self.handle, int(ticket)))
        keep = getattr(self, "_inflight", {})
        for k in [k for k in keep if ticket < 0 or k <= ticket]:
            del keep[k]

    # ---- droplet scene (traj_renderer.py / traj_vel_renderer.py) ---------------------------
    def set_droplet_mesh(self, verts, n_rings, n_segments):
        """verts: ((n_rings+1)*n_segments, 3) float32 numpy array, ring-major (the OBJ's vertices)."""
        v = np.ascontiguousarray(verts, np.float32)
        assert v.shape == ((n_rings + 1) * n_segments, 3)
        self._check(self.lib.pcr_set_droplet_mesh(self.handle, v.ctypes.data, int(n_rings), int(n_segments)))
        self.droplet_mesh = (int(n_rings), int(n_segments))

    def droplet_transforms(self, pcl, rot=None, stream=None):
        """(N,3|6) float32 CUDA tensor (transformed) [+ rot (N,9) for 3 columns] -> (N,12) rows [R | position]."""
        import torch
        _check_points(pcl, f32_only=True)
        n, cols = pcl.shape
        _check_vector(rot, 9 * n, what="rot")
        xf = torch.empty((n, 12), dtype=torch.float32, device=pcl.device)
        self._check(self.lib.pcr_droplet_transforms(self.handle, _ptr(pcl), n, cols, _ptr(rot), _ptr(xf), _stream_ptr(stream)))
        return xf

    def history_trails(self, hist, pos, stream=None):
        """hist (h,N,3), pos (N,3) float32 CUDA tensors (transformed) -> ctrl (N,21,3) float32, count (N,) int32."""
        import torch
        _check_points(pos, f32_only=True, what="pos")
        n = pos.shape[0]
        if hist.dim() != 3 or hist.shape[1:] != (n, 3) or hist.dtype != torch.float32 or not hist.is_contiguous() or (hist.shape[0] and not hist.is_cuda):
            raise ValueError("hist: expected a contiguous float32 CUDA tensor (h, N, 3)")
        ctrl = torch.zeros((n, MAX_CTRL, 3), dtype=torch.float32, device=pos.device)
        count = torch.empty(n, dtype=torch.int32, device=pos.device)
        self._check(self.lib.pcr_history_trails(self.handle, _ptr(hist) if hist.shape[0] else None, int(hist.shape[0]), _ptr(pos), n,
                                                _ptr(ctrl), _ptr(count), _stream_ptr(stream)))
        return ctrl, count

    def render_droplet_frames(self, traj, cams, style, n_history=0, rot=None, want_vis=False, out_rgba=None, out_vis=None, stream=None):
        """traj: (n_history + F, N, 3|6) CUDA tensor — the F frames to render preceded by their history halo;
        cams: F cameras -> rgba (F,H,W,4) uint8 [, vis (F,H,W) int64]."""
        import torch
        _check_points(traj, dims=3, what="traj")
        total, n, cols = traj.shape
        F = total - int(n_history)
        if F < 0 or len(cams) != F:
            raise ValueError("one camera per rendered frame (frames after the history halo)")
        _check_vector(rot, 9 * n, what="rot")
        W, H = (cams[0].width, cams[0].height) if F else (0, 0)
        rgba = out_rgba if out_rgba is not None else torch.empty((F, H, W, 4), dtype=torch.uint8, device=traj.device)
        vis = out_vis if out_vis is not None else (torch.empty((F, H, W), dtype=torch.int64, device=traj.device) if want_vis else None)
        if F:
            cam_arr = (Camera * F)(*cams)
            self._check(self.lib.pcr_render_droplet_frames(self.handle, _ptr(traj), int(traj.dtype == torch.float64), n, cols, F,
                                                           int(n_history), _ptr(rot), cam_arr, ctypes.byref(style), _ptr(vis), _ptr(rgba),
                                                           _stream_ptr(stream)))
        return (rgba, vis) if (want_vis or out_vis is not None) else rgba

    # ---- fused z-merge over peer memory (point-sharded clouds) -----------------------------
    def peer_alloc(self, width, height):
        """This rank's merged z-buffer and image (library-owned cudaMalloc blocks): raw device pointers."""
        m, im = ctypes.c_void_p(), ctypes.c_void_p()
        self._check(self.lib.pcr_peer_alloc(self.handle, int(width), int(height), ctypes.byref(m), ctypes.byref(im)))
        self.peer_shape = (int(height), int(width))
        return m.value, im.value

    def ipc_export(self, ptr):
        h = (ctypes.c_uint8 * 64)()
        self._check(self.lib.pcr_ipc_export(self.handle, ctypes.c_void_p(ptr), h))
        return bytes(h)

    def ipc_open(self, handle):
        out = ctypes.c_void_p()
        buf = (ctypes.c_uint8 * 64).from_buffer_copy(handle)
        self._check(self.lib.pcr_ipc_open(self.handle, buf, ctypes.byref(out)))
        return out.value

    def ipc_close(self, ptr):
        self._check(self.lib.pcr_ipc_close(self.handle, ctypes.c_void_p(ptr)))

    def peer_set(self, rank, world, merged_ptrs, image_ptrs, dst_rank=0):
        arr_m = (ctypes.c_void_p * max(world, 1))(*[ctypes.c_void_p(p) for p in merged_ptrs])
        arr_i = (ctypes.c_void_p * max(world, 1))(*[ctypes.c_void_p(p) for p in image_ptrs])
        self._check(self.lib.pcr_peer_set(self.handle, int(rank), int(world), int(dst_rank), arr_m, arr_i))

    def peer_begin_frame(self, cam, style, stream=None):
        self._check(self.lib.pcr_peer_begin_frame(self.handle, ctypes.byref(cam), ctypes.byref(style), _stream_ptr(stream)))

    def render_shard_peer(self, pts, stats10, cam, style, id_base=0, radius=None, rgb=None, out_vis=None, stream=None):
        import torch
        _check_points(pts)
        n, cols = pts.shape
        _check_vector(stats10, 10, dtype=torch.float64, what="stats10")
        _check_vector(radius, n, what="radius"); _check_vector(rgb, 3 * n, what="rgb")
        vis = out_vis if out_vis is not None else torch.empty((cam.height, cam.width), dtype=torch.int64, device=pts.device)
        self._check(self.lib.pcr_render_shard_peer(self.handle, _ptr(pts) if n else None, int(pts.dtype == torch.float64), n, cols,
                                                   _ptr(radius), _ptr(rgb), _ptr(stats10), int(id_base), ctypes.byref(cam),
                                                   ctypes.byref(style), _ptr(vis), _stream_ptr(stream)))
        return vis

    def shade_shard_peer(self, vis, pts, stats10, cam, style, id_base=0, radius=None, rgb=None, stream=None):
        import torch
        _check_points(pts)
        n, cols = pts.shape
        _check_vector(stats10, 10, dtype=torch.float64, what="stats10")
        _check_vector(radius, n, what="radius"); _check_vector(rgb, 3 * n, what="rgb")
        self._check(self.lib.pcr_shade_shard_peer(self.handle, _ptr(vis), _ptr(pts) if n else None, int(pts.dtype == torch.float64), n, cols,
                                                  _ptr(radius), _ptr(rgb), _ptr(stats10), int(id_base), ctypes.byref(cam),
                                                  ctypes.byref(style), _stream_ptr(stream)))

    def peer_buffers(self):
        """torch views (no copy) of this rank's merged keys (H,W) int64 and image (H,W,4) uint8."""
        H, W = self.peer_shape
        m, im = self.peer_alloc(W, H)
        return _device_view(m, (H, W), "int64", self.device), _device_view(im, (H, W, 4), "uint8", self.device)

    # ---- merge ---------------------------------------------------------------------------
    def zmin_(self, dst, src, stream=None):
        self._check(self.lib.pcr_zmin(self.handle, _ptr(dst), _ptr(src), dst.numel(), _stream_ptr(stream)))
        return dst

    def zmerge_nccl_(self, vis, comm, stream=None):
        """C1 through the C ABI: in-place ncclAllReduce(uint64, min) of the (H,W) keys over `comm` (sharding.NcclComm
        or a raw ncclComm_t address)."""
        import torch
        _check_vector(vis, vis.numel(), dtype=torch.int64, what="vis")
        handle = getattr(comm, "comm", comm)
        handle = handle if isinstance(handle, ctypes.c_void_p) else ctypes.c_void_p(int(handle))
        self._check(self.lib.pcr_zmerge_nccl(self.handle, _ptr(vis), vis.numel(), handle, _stream_ptr(stream)))
        return vis

    def set_occlusion(self, mode=-1, step=0, min_points=0):
        """Occlusion pre-pass: mode -1 auto, 0 off, 1 always (see pcr_set_occlusion)."""
        self._check(self.lib.pcr_set_occlusion(self.handle, int(mode), int(step), int(min_points)))

    def profile(self, enable=True):
        self._check(self.lib.pcr_profile(self.handle, int(bool(enable))))

    def profile_read(self):
        """{kernel name: (total ms, launches)} since the last read (waits for the recorded events)."""
        ms = (ctypes.c_double * 32)()
        cnt = (ctypes.c_int64 * 32)()
        k = self.lib.pcr_profile_read(self.handle, ms, cnt, 32)
        if k < 0:
            self._check(k)
        return {self.lib.pcr_kernel_name(i).decode(): (ms[i], cnt[i]) for i in range(k) if cnt[i] > 0}

    def counters(self, stream=None):
        out = (ctypes.c_int64 * 4)()
        self._check(self.lib.pcr_counters(self.handle, out, _stream_ptr(stream)))
        return {"launches": out[0], "pairs_last_frame": out[1], "overflow_frames": out[2]}

    def selftest_scale_div(self, divisors):
        """Mismatching quotient bit patterns between the kernels' hoisted-reciprocal division and __fdiv_rn over
        every binary32 dividend, for each given divisor (expected 0)."""
        d = (ctypes.c_float * len(divisors))(*[float(x) for x in divisors])
        bad = ctypes.c_uint64(0)
        self._check(self.lib.pcr_selftest_scale_div(self.handle, d, len(divisors), ctypes.byref(bad)))
        return int(bad.value)


def _device_view(ptr, shape, dtype, device):
    """A torch tensor aliasing raw device memory (library-owned; the caller keeps the context alive)."""
    import torch

    class _Mem:
        __cuda_array_interface__ = {"shape": tuple(shape), "typestr": {"int64": "<i8", "uint8": "|u1"}[dtype], "data": (int(ptr), False),
                                    "version": 3, "strides": None}
    return torch.as_tensor(_Mem(), device=f"cuda:{device}")


def keys_to_ids(vis):
    """Low 32 bits of the keys as uint32 numpy array."""
    a = vis.cpu().numpy() if hasattr(vis, "cpu") else np.asarray(vis)
    return (a.view(np.uint64) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
