"""Host-side mirror of the reference's renderer classes, backed by the pcr C ABI.

Same class and method names, argument meaning and error behaviour as the reference scripts:

  PointCloudRenderer        example_renderer.py:77-199
  TrajectoryBallRenderer    traj_ball_renderer.py:80-416
  FixedFrame199Renderer     traj_original.py:6-142   (no x flip, fixed camera)
  TrajB0Renderer            traj_b0.py:6-191         (class is also called FixedFrame199Renderer there)
  TrajB1Renderer            traj_b1.py:6-191
  TrajectoryRenderer        traj_renderer.py:87-676      (droplet meshes + Catmull-Rom history trails)
  TrajectoryVelRenderer     traj_vel_renderer.py:80-530  (droplet meshes + straight velocity trails)

What changes is the seam: `render_scene` takes the transformed point array
instead of an XML path, and nothing is written to disk between the stages — centres, radii and
colours stay in HBM from K1 to K4.  There is no CPU fallback: every array method runs through
libpcr.so on a CUDA device and raises if either is missing.
"""
import os

import numpy as np

from . import _native
from .io import load_point_cloud as _load_point_cloud, write_png
from .presets import PRESETS

_ENGINES = {}


def _engine(device, n, w, h, batch=1):
    """One grow-only pcr context per device (scratch for n points, w x h pixels)."""
    key = int(device)
    eng = _ENGINES.get(key)
    if eng is None or eng.max_points < n or eng.max_w < w or eng.max_h < h or eng.max_batch < batch:
        if eng is not None:
            n, w, h, batch = max(n, eng.max_points), max(w, eng.max_w), max(h, eng.max_h), max(batch, eng.max_batch)
            eng.close()
        eng = _native.Context(device=key, max_points=max(int(n), 1), max_w=int(w), max_h=int(h), max_batch=int(batch))
        _ENGINES[key] = eng
    return eng


def release_engines():
    for eng in _ENGINES.values():
        eng.close()
    _ENGINES.clear()


def _to_device(pcl, device):
    """numpy / torch (any device) -> contiguous CUDA tensor f32|f64, plus 'was numpy' flag."""
    import torch
    if isinstance(pcl, torch.Tensor):
        t, was_np = pcl, False
    else:
        a = np.asarray(pcl)
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64)          # numpy promotes ints to f64 in mean/divide too
        t, was_np = torch.from_numpy(np.ascontiguousarray(a)), True
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float64)
    return t.to(f"cuda:{device}", non_blocking=True).contiguous(), was_np


def _back(t, was_np):
    return t.cpu().numpy() if was_np else t


class RenderedScene:
    """What render_scene returns (stands in for the Mitsuba image tensor): sRGB8 image and the
    visibility keys, both still on the device."""

    def __init__(self, rgba, vis):
        self.rgba = rgba      # (H,W,4) uint8 CUDA tensor
        self.vis = vis        # (H,W) int64 CUDA tensor holding the uint64 keys

    def numpy(self):
        return self.rgba.cpu().numpy()

    def point_ids(self):
        return _native.keys_to_ids(self.vis)


class PointCloudRenderer:
    """example_renderer.py:77 — one cloud -> one image, fixed camera."""
    PRESET = "example"
    DEVICE = 0
    WITH_VELOCITY = False            # load_point_cloud reads x,y,z only (example_renderer.py:108-109)
    TRAILS_DEFAULT = False           # example_renderer.py draws no trails

    def __init__(self, file_path, output_folder=None, width=None, height=None, color_mode=_native.COLOR_CONST,
                 radius=None, user_rgb=None, trails=None):
        self.file_path = file_path
        self.folder, full_filename = os.path.split(file_path) if file_path else ("", "")
        self.folder = self.folder or '.'
        self.filename, _ = os.path.splitext(full_filename)
        self.output_folder = output_folder
        self.config = PRESETS[self.PRESET]
        self.width = int(width or self.config.width)
        self.height = int(height or self.config.height)
        self.color_mode = int(color_mode)
        self.radius = radius          # None -> the BALL_SEGMENT literal; float; or per-point array (extension)
        self.user_rgb = user_rgb
        # draw the script's velocity trails for 6-column frames: None = what the script does (the trajectory scripts
        # always call _add_velocity_trail for an (N,6) cloud, traj_ball_renderer.py:319-324; example_renderer.py never)
        self.trails = self.TRAILS_DEFAULT if trails is None else bool(trails)

    # ---- hooks the reference exposes ----------------------------------------------------
    @staticmethod
    def compute_color(x=0.0, y=0.0, z=0.0, noise_seed=0):
        """example_renderer.py:89-92: constant grey, arguments ignored (colour modes 1-3 of
        pcr_style are the extensions north_star names; mode 0 is this)."""
        g = 0.3
        return np.array([g, g, g])

    @classmethod
    def standardize_point_cloud(cls, pcl):
        """example_renderer.py:94-98 / traj_ball_renderer.py:190-202 on the device (K0 + K1 with
        the axis transform switched off).  numpy in -> numpy out, CUDA tensor in -> CUDA tensor out."""
        import torch
        t, was_np = _to_device(pcl, cls.DEVICE)
        if t.dim() != 2 or t.shape[1] not in (3, 6):
            raise ValueError("point cloud must be (N,3) or (N,6)")
        n, cols = t.shape
        eng = _engine(cls.DEVICE, n, 16, 16)
        style = PRESETS[cls.PRESET].style(xform=1)
        if cols == 6:
            pos, _, vel = eng.standardize(t, style, want_vel=True)
            out = torch.cat([pos[:, :3], vel[:, :3]], dim=1).contiguous()
        else:
            pos, _ = eng.standardize(t, style)
            out = pos[:, :3].contiguous()
        return _back(out, was_np)

    @classmethod
    def transform_coordinates(cls, pcl):
        """traj_ball_renderer.py:204-221 (and the inline example_renderer.py:171-173)."""
        import torch
        t, was_np = _to_device(pcl, cls.DEVICE)
        t = t.to(torch.float32)
        cfg = PRESETS[cls.PRESET]
        eng = _engine(cls.DEVICE, t.shape[0], 16, 16)
        return _back(eng.transform_coordinates(t, flip_x=cfg.flip_x, z_lift=cfg.z_lift), was_np)

    @classmethod
    def compute_camera_position(cls, frame_index=0, total_frames=220):
        return PRESETS[cls.PRESET].camera_position(frame_index, total_frames)

    def load_point_cloud(self):
        return _load_point_cloud(self.file_path, with_velocity=self.WITH_VELOCITY)

    @classmethod
    def init_mitsuba_variant(cls):
        """Name kept from example_renderer.py:137-151.  There is no variant chain any more: the
        only back end is the sm_100a extension, so this loads it and fails loudly otherwise."""
        import torch
        _native.load_library()
        if not torch.cuda.is_available():
            raise RuntimeError("pcr needs a CUDA device (no CPU fallback)")
        print(f'Using CUDA GPU (pcr / {torch.cuda.get_device_name(cls.DEVICE)})')
        return True

    init_device = init_mitsuba_variant

    # ---- the seam: render_scene / save_scene ---------------------------------------------
    def _style(self):
        return self.config.style(color_mode=self.color_mode, trails=self.trails)

    def _per_point(self, n, device):
        import torch
        radius = rgb = None
        if self.radius is not None and np.ndim(self.radius) > 0:
            radius = torch.as_tensor(np.asarray(self.radius, np.float32)).to(device)
            if radius.numel() != n:
                raise ValueError("per-point radius must have one entry per point")
        if self.user_rgb is not None:
            rgb = torch.as_tensor(np.asarray(self.user_rgb, np.float32)).to(device).contiguous()
            if rgb.numel() != 3 * n:
                raise ValueError("user_rgb must be (N,3)")
        return radius, rgb

    def render_scene(self, pcl, frame_index=0, total_frames=220):
        """Replaces generate_xml_content + save_xml + mi.load_file + mi.render
        (example_renderer.py:113-157): `pcl` is the standardised, transformed (N,3|6) array."""
        import torch
        t, _ = _to_device(pcl, self.DEVICE)
        t = t.to(torch.float32)
        if t.dim() != 2 or t.shape[1] not in (3, 6):
            raise ValueError("point cloud must be (N,3) or (N,6)")
        n = t.shape[0]
        eng = _engine(self.DEVICE, n, self.width, self.height)
        style = self._style()
        if self.radius is not None and np.ndim(self.radius) == 0:
            style.radius = float(self.radius)
        radius, rgb = self._per_point(n, t.device)
        cam = self.config.camera(frame_index, total_frames, self.width, self.height)
        if style.trails == 1 and t.shape[1] == 6:
            # spheres AND the velocity trail of every point, like generate_xml_content (traj_ball_renderer.py:319-330):
            # the fused path with the frame taken as it is (pcr_render_transformed)
            vis, rgba = eng.render_transformed(t, cam, style, radius=radius, rgb=rgb)
            return RenderedScene(rgba, vis)
        # colour hook + float4 packing without touching the coordinates again (xform = 1 keeps
        # positions as they are; centre 0 / scale 1 are injected so nothing is re-standardised)
        stats = torch.zeros(10, dtype=torch.float64, device=t.device)
        stats[9] = 1.0
        if n > 0:
            stats[3:9] = eng.stats_partial(t)[3:9]      # min / max of the cloud for the position colormap
        st1 = self.config.style(color_mode=self.color_mode, xform=1)
        st1.radius = style.radius
        pos4, attr4 = eng.standardize_with_stats(t, st1, stats, radius=radius, rgb=rgb)
        vis, rgba = eng.render(pos4, attr4, cam, style)
        return RenderedScene(rgba, vis)

    @staticmethod
    def save_scene(output_file_path, rendered_scene, writer=None):
        """example_renderer.py:159-161: the sRGB transfer already happened in K4; this is the PNG.
        With an output.AsyncImageWriter the file is written in the background, like Mitsuba's
        write_bitmap(write_async=True); call writer.drain() before reading it back."""
        if writer is not None:
            return writer.submit(output_file_path, rendered_scene.numpy())
        write_png(f'{output_file_path}.png', rendered_scene.numpy())

    def _output_path(self, output_filename):
        if self.output_folder:
            os.makedirs(self.output_folder, exist_ok=True)
            return os.path.join(self.output_folder, output_filename)
        return os.path.join(self.folder, output_filename)

    def process(self):
        """example_renderer.py:163-199.  A 3-D input renders every frame to the same file name
        (last frame wins), exactly as the reference does (:165-175)."""
        pcl_data = self.load_point_cloud()
        if len(pcl_data.shape) < 3:
            pcl_data = pcl_data[np.newaxis, :, :]
        total_frames = len(pcl_data)
        for index, pcl in enumerate(pcl_data):
            pcl, _ = _to_device(pcl, self.DEVICE)
            pcl = self.standardize_point_cloud(pcl)
            pcl = self.transform_coordinates(pcl)
            output_file_path = self._output_path(f'{self.filename}')
            if total_frames > 1:
                print(f'  Frame {index+1}/{total_frames}: Rendering...', end=' ', flush=True)
            else:
                print('  Rendering...', end=' ', flush=True)
            rendered_scene = self.render_scene(pcl)
            print('Saving...', end=' ', flush=True)
            self.save_scene(output_file_path, rendered_scene)
            print('Done!')


class TrajectoryBallRenderer(PointCloudRenderer):
    """traj_ball_renderer.py:80 — per-frame sphere render with an animated camera."""
    PRESET = "traj_ball"
    WITH_VELOCITY = True
    TRAILS_DEFAULT = True            # _add_velocity_trail for every point of an (N,6) cloud (traj_ball_renderer.py:319-324)

    @staticmethod
    def compute_color():
        return np.array([0.3, 0.3, 0.3])

    def process(self, frame_index=0, total_frames=220):
        """traj_ball_renderer.py:365-398.  An (N,6) cloud gets its velocity trails (render_scene -> pcr_render_transformed),
        as generate_xml_content draws them (:319-330); trails=False in the constructor turns them off."""
        pcl = self.load_point_cloud()
        if len(pcl.shape) == 3:
            pcl = pcl[0]
        pcl, _ = _to_device(pcl, self.DEVICE)
        pcl = self.standardize_point_cloud(pcl)
        pcl = self.transform_coordinates(pcl)
        # the '_b0' suffix for fade frames is inherited by every subclass (traj_ball_renderer.py:376)
        output_filename = f'frame_{frame_index:04d}_b0' if frame_index > 199 else self.filename
        output_file_path = self._output_path(output_filename)
        print('  Rendering...', end=' ', flush=True)
        rendered_scene = self.render_scene(pcl, frame_index=frame_index, total_frames=total_frames)
        print('Saving...', end=' ', flush=True)
        self.save_scene(output_file_path, rendered_scene)
        print('Done!')

    # ---- batched trajectory path: the whole hot path in one C-ABI call ------------------------
    def render_trajectory(self, traj, first_frame=0, total_frames=None, want_vis=False, max_batch=16, stretch=True, out_rgba=None):
        """traj: (F,N,3|6) array (numpy/CPU tensor -> pcr_render_frames_host with copies overlapped;
        CUDA tensor -> pcr_render_frames).  Frame f uses compute_camera_position(first_frame+f).
        stretch=True rescales the 220-frame camera schedule to total_frames."""
        import torch
        F, n, cols = traj.shape
        total = int(total_frames or F)
        cfg = self.config.for_trajectory(total) if stretch else self.config
        cams = [cfg.camera(first_frame + f, total, self.width, self.height) for f in range(F)]
        style = self._style()
        if self.radius is not None and np.ndim(self.radius) == 0:
            style.radius = float(self.radius)
        eng = _engine(self.DEVICE, n, self.width, self.height, batch=min(max_batch, max(F, 1)))
        if isinstance(traj, torch.Tensor) and traj.is_cuda:
            radius, rgb = self._per_point(n, traj.device)
            return eng.render_frames(traj.contiguous(), cams, style, radius=radius, rgb=rgb, want_vis=want_vis, out_rgba=out_rgba)
        t = torch.as_tensor(traj)
        radius = None if self.radius is None or np.ndim(self.radius) == 0 else np.ascontiguousarray(self.radius, np.float32)
        rgb = None if self.user_rgb is None else np.ascontiguousarray(self.user_rgb, np.float32)
        out_vis = torch.empty((F, self.height, self.width), dtype=torch.int64).pin_memory() if want_vis else None
        rgba = eng.render_frames_host(t.contiguous(), cams, style, radius_host=radius, rgb_host=rgb, out_vis=out_vis, out_rgba=out_rgba)
        return (rgba, out_vis) if want_vis else rgba


    def render_trajectory_to_files(self, traj, output_folder=None, first_frame=0, total_frames=None, stem=None, chunk=16, writer=None):
        """Render a trajectory and write one PNG per frame with the reference's naming rule; encoding
        runs on a thread pool while the GPU renders the next chunk (output.render_to_files)."""
        from .output import render_to_files
        return render_to_files(self, traj, output_folder or self.output_folder or self.folder, first_frame, total_frames, stem, chunk, writer)


class FixedFrame199Renderer(TrajectoryBallRenderer):
    """traj_original.py:6 — no x flip (:40-60), fixed camera (-1.8,-1.8,1.8) (:62-66)."""
    PRESET = "traj_original"


class TrajB0Renderer(TrajectoryBallRenderer):
    """traj_b0.py:6 — dataset batch_0: no x flip, shifted look-at, 3-keyframe camera, big floor."""
    PRESET = "traj_b0"


class TrajB1Renderer(TrajectoryBallRenderer):
    """traj_b1.py:6 — dataset batch_1."""
    PRESET = "traj_b1"


class _DropletMixin:
    """What traj_renderer.py and traj_vel_renderer.py share: every point is an instance of the droplet mesh
    (_create_droplet_mesh, traj_renderer.py:102-153) oriented along its velocity, plus one trail curve.
    TRAILS: 1 = straight velocity trail (traj_vel_renderer.py:194-288), 2 = Catmull-Rom history trail
    (traj_renderer.py:204-396).  droplets=False in the constructor keeps the sphere path of the parent."""
    TRAILS = _native.TRAILS_NONE

    def __init__(self, file_path, output_folder=None, droplet_mesh_path=None, droplets=True, **kw):
        # droplet_mesh_path (traj_renderer.py:93-99): a caller-supplied OBJ replaces the built-in droplet.  It must be
        # a ring-structured surface of revolution (droplets.load_ring_mesh_obj raises otherwise — never silently ignored)
        super().__init__(file_path, output_folder=output_folder, **kw)
        self.droplets = bool(droplets)
        self.droplet_mesh_path = droplet_mesh_path
        self._mesh = None
        if droplet_mesh_path is not None:
            from . import droplets as _d
            self._mesh = _d.load_ring_mesh_obj(droplet_mesh_path)
        self.curve_files = []           # the reference's temp curve files: none are written any more

    @staticmethod
    def _create_droplet_mesh():
        """traj_renderer.py:102-153 without the OBJ file: ((17*20), 3) float32 vertices, (640, 3) faces."""
        from . import droplets
        return droplets.droplet_vertices(), droplets.droplet_faces()

    @classmethod
    def generate_rotation_matrix_from_velocity(cls, velocity, translation):
        """traj_renderer.py:159-202 on the device: flattened 4x4 (float32 values, what a loader reads
        from the matrix the reference prints)."""
        import torch
        row = torch.tensor([[*map(float, translation), *map(float, velocity)]], dtype=torch.float32, device=f"cuda:{cls.DEVICE}")
        xf = _engine(cls.DEVICE, 1, 16, 16).droplet_transforms(row).cpu().numpy().reshape(3, 4).astype(np.float64)
        return np.concatenate([xf, [[0.0, 0.0, 0.0, 1.0]]]).flatten()

    @staticmethod
    def generate_random_rotation_matrix(seed, translation):
        """traj_renderer.py:398-418 (numpy's legacy generator stays on the host)."""
        from . import droplets
        m = np.eye(4)
        m[:3, :3] = droplets.random_rotations(int(seed) + 1)[int(seed)].reshape(3, 3)
        m[:3, 3] = translation
        return m.flatten()

    def _droplet_engine(self, n, batch=1):
        from . import droplets
        eng = _engine(self.DEVICE, n, self.width, self.height, batch=batch)
        verts, rings, segs = self._mesh if self._mesh is not None else (droplets.droplet_vertices(), droplets.N_RINGS, droplets.N_SEGMENTS)
        key = (rings, segs, hash(verts.tobytes()))
        if getattr(eng, "droplet_mesh_key", None) != key:
            eng.set_droplet_mesh(verts, rings, segs)
            eng.droplet_mesh_key = key
        return eng

    def _rotations(self, n, cols, device):
        """Rotations of points without velocity: random per index for traj_renderer.py (:566),
        identity for traj_vel_renderer.py (:427-430)."""
        import torch
        if cols == 6 or self.TRAILS != _native.TRAILS_HISTORY:
            return None
        from . import droplets
        return torch.from_numpy(np.ascontiguousarray(droplets.random_rotations(n))).to(device)

    def render_scene(self, pcl, frame_index=0, total_frames=220, history_pcls=None):
        """generate_xml_content + mi.load_file + mi.render of the droplet scripts (traj_renderer.py:529-602):
        `pcl` and `history_pcls` are standardised, transformed (N,3|6) arrays, history oldest first."""
        import torch
        if not self.droplets:
            return super().render_scene(pcl, frame_index=frame_index, total_frames=total_frames)
        t, _ = _to_device(pcl, self.DEVICE)
        t = t.to(torch.float32)
        n, cols = t.shape
        hist = []
        if self.TRAILS == _native.TRAILS_HISTORY and cols == 6 and history_pcls:
            for h in list(history_pcls)[-_native.HISTORY_FRAMES:]:
                h, _ = _to_device(h, self.DEVICE)
                h = h.to(torch.float32)
                if h.shape[0] < n:                       # the reference skips points a history frame lacks
                    raise ValueError("history frames must hold every point of the current frame")
                hh = torch.zeros((n, cols), dtype=torch.float32, device=t.device)
                hh[:, :3] = h[:n, :3]
                hist.append(hh)
        buf = torch.stack(hist + [t]).contiguous()
        eng = self._droplet_engine(n)
        style = self.config.style(color_mode=self.color_mode, xform=2, trails=self.TRAILS if self.trails else 0)
        cam = self.config.camera(frame_index, total_frames, self.width, self.height)
        rgba, vis = eng.render_droplet_frames(buf, [cam], style, n_history=len(hist), rot=self._rotations(n, cols, t.device), want_vis=True)
        return RenderedScene(rgba[0], vis[0])

    def render_trajectory(self, traj, first_frame=0, total_frames=None, want_vis=False, max_batch=16, stretch=True, out_rgba=None,
                          n_history=0):
        """traj: (n_history + F, N, 3|6) raw frames — the F frames to render preceded by n_history frames of
        history (a frame-sharded caller passes a 20-frame halo; rank 0 passes 0).  Whole path in one call:
        pcr_render_droplet_frames.  Frame f of the F uses compute_camera_position(first_frame + f)."""
        import torch
        if not self.droplets:
            return super().render_trajectory(traj, first_frame, total_frames, want_vis, max_batch, stretch, out_rgba)
        t = traj if isinstance(traj, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(traj))
        was_cuda = t.is_cuda
        total, n, cols = t.shape
        F = total - int(n_history)
        tot = int(total_frames or (first_frame + F))
        cfg = self.config.for_trajectory(tot) if stretch else self.config
        cams = [cfg.camera(first_frame + f, tot, self.width, self.height) for f in range(F)]
        style = self.config.style(color_mode=self.color_mode, trails=self.TRAILS if self.trails else 0)
        eng = self._droplet_engine(n, batch=min(max_batch, max(F, 1)))
        dev = t.to(f"cuda:{self.DEVICE}", non_blocking=True).contiguous()
        out = eng.render_droplet_frames(dev, cams, style, n_history=int(n_history), rot=self._rotations(n, cols, dev.device),
                                        want_vis=want_vis, out_rgba=out_rgba if (out_rgba is not None and out_rgba.is_cuda) else None)
        rgba, vis = out if want_vis else (out, None)
        if not was_cuda:
            if out_rgba is not None and not out_rgba.is_cuda:
                out_rgba.copy_(rgba)
                rgba = out_rgba
            else:
                rgba = rgba.cpu()
            vis = None if vis is None else vis.cpu()
        return (rgba, vis) if want_vis else rgba


class TrajectoryRenderer(_DropletMixin, TrajectoryBallRenderer):
    """traj_renderer.py:87 — droplets + Catmull-Rom history trails, linear dolly camera (:519-527)."""
    PRESET = "traj"
    TRAILS = _native.TRAILS_HISTORY

    def __init__(self, file_path, output_folder=None, droplet_mesh_path=None, droplets=True, trails=True, **kw):
        super().__init__(file_path, output_folder=output_folder, droplet_mesh_path=droplet_mesh_path, droplets=droplets, trails=trails, **kw)

    def process(self, frame_index=0, history_pcls=None, total_frames=220):
        """traj_renderer.py:604-646: history_pcls = the previous (<= 20) frames, standardised and transformed."""
        pcl = self.load_point_cloud()
        if len(pcl.shape) == 3:
            pcl = pcl[0]
        pcl, _ = _to_device(pcl, self.DEVICE)
        pcl = self.standardize_point_cloud(pcl)
        pcl = self.transform_coordinates(pcl)
        output_filename = f'frame_{frame_index:04d}_b0' if frame_index > 199 else self.filename
        output_file_path = self._output_path(output_filename)
        print('  Rendering...', end=' ', flush=True)
        rendered_scene = self.render_scene(pcl, frame_index=frame_index, total_frames=total_frames, history_pcls=history_pcls)
        print('Saving...', end=' ', flush=True)
        self.save_scene(output_file_path, rendered_scene)
        print('Done!')

    @staticmethod
    def cleanup_temp_meshes():
        """traj_renderer.py:648-653: nothing to clean — no temp_meshes/ directory is created."""

    def cleanup_temp_curves(self):
        self.curve_files = []

    @staticmethod
    def cleanup_temp_curves_dir():
        """traj_renderer.py:666-671: nothing to clean — no temp_curves/ directory is created."""


class TrajectoryVelRenderer(_DropletMixin, TrajectoryBallRenderer):
    """traj_vel_renderer.py:80 — droplets + straight velocity trails (:194-288), camera as traj_ball (:381-407);
    velocity attribute available to the colour hook (PCR_COLOR_VELOCITY)."""
    PRESET = "traj_vel"
    TRAILS = _native.TRAILS_VELOCITY

    def __init__(self, file_path, output_folder=None, droplet_mesh_path=None, droplets=True, trails=True, **kw):
        super().__init__(file_path, output_folder=output_folder, droplet_mesh_path=droplet_mesh_path, droplets=droplets, trails=trails, **kw)

    cleanup_temp_meshes = TrajectoryRenderer.cleanup_temp_meshes
    cleanup_temp_curves = TrajectoryRenderer.cleanup_temp_curves
    cleanup_temp_curves_dir = TrajectoryRenderer.cleanup_temp_curves_dir
