"""Per-script scene presets: every literal the reference hard-codes in its XML templates and
camera schedules, as one RenderConfig per script.

The reference has no configuration system — customisation is "subclass and override the
class attributes" (traj_b0.py:10-60).  Here the same literals live in data.
"""
from dataclasses import dataclass, replace
from typing import Optional, Tuple

from . import _native

Vec3 = Tuple[float, float, float]


@dataclass(frozen=True)
class RenderConfig:
    name: str
    target: Vec3                      # <lookat target=...>
    fov: float                        # <float name="fov">, horizontal
    flip_x: bool                      # transform_coordinates flavour
    floor_z: float
    floor_min: Tuple[float, float]
    floor_max: Tuple[float, float]
    spp: int                          # sampleCount of the reference (informational: we cast 1 ray/pixel)
    # camera schedule: None = fixed eye, else (start, mid, end) keyframes or "dolly"
    eye: Optional[Vec3] = None
    keys: Optional[Tuple[Vec3, Vec3, Vec3]] = None
    dolly: bool = False
    last_motion_frame: int = 199      # literals of traj_ball_renderer.py:287-288
    fade_frames: int = 20
    up: Vec3 = (0.0, 0.0, 1.0)
    near_clip: float = 0.1            # nearClip / farClip of every HEAD
    far_clip: float = 100.0
    width: int = 1920                 # film size of every HEAD
    height: int = 1080
    radius: float = 0.01              # BALL_SEGMENT radius literal
    const_rgb: Vec3 = (0.3, 0.3, 0.3)  # compute_color
    z_lift: float = 0.0125
    vel_norm: float = 10.0            # traj_ball_renderer.py:134
    light_z: float = 15.0             # TAIL emitter
    light_half: float = 8.0
    radiance: float = 4.0
    floor_albedo: float = 1.0         # diffuseReflectance of surfaceMaterial
    bounce: float = 1.0
    # velocity trails (_add_velocity_trail): 'ramp' = traj_ball_renderer.py:119-124, 'ramp_fade' =
    # traj_vel_renderer.py:215-224, 'const' = traj_original.py:78 / traj_b0.py:127, None = the script draws none
    trail_schedule: Optional[str] = None
    trail_radius: float = 0.0007
    trail_rgb: Vec3 = (0.2, 1.0, 0.4)
    trail_len_min: float = 0.07
    trail_len_max: float = 0.3

    def camera_position(self, frame_index=0, total_frames=220):
        """compute_camera_position of the script this preset mirrors (python floats, f64):
        example_renderer.py:20 ; traj_renderer.py:519-527 ; traj_ball_renderer.py:281-307 ;
        traj_vel_renderer.py:381-407 ; traj_original.py:62-66 ; traj_b0.py:84-115 ; traj_b1.py:84-115."""
        if self.eye is not None:
            return self.eye
        if self.dolly:
            progress = frame_index / max(total_frames - 1, 1)
            return (2.8 - 2.0 * progress, 2.8 - 2.0 * progress, 3.0 - 2.0 * progress)
        start, mid, end = self.keys
        if frame_index <= self.last_motion_frame:
            p = frame_index / max(self.last_motion_frame, 1)
            a, b = start, mid
        else:
            p = (frame_index - self.last_motion_frame) / max(self.fade_frames, 1)
            a, b = mid, end
        return tuple(a[k] + (b[k] - a[k]) * p for k in range(3))

    def trail_length_scale(self, frame_index):
        """length_scale of the script's _add_velocity_trail (python floats, f64); 0 when it draws none.
        The 0-19 ramp and the 199/20 fade are literals of the reference (not stretched)."""
        if self.trail_schedule is None:
            return 0.0
        if self.trail_schedule == "const":
            return 1.0
        if frame_index <= 19:
            return frame_index / 19.0
        if self.trail_schedule == "ramp_fade" and frame_index > 199:
            return 1.0 - (frame_index - 199) / 20
        return 1.0

    def camera(self, frame_index=0, total_frames=220, width=None, height=None):
        return _native.make_camera(self.camera_position(frame_index, total_frames), self.target, self.up, self.fov,
                                   self.near_clip, self.far_clip, width or self.width, height or self.height,
                                   trail_scale=self.trail_length_scale(frame_index))

    def style(self, color_mode=_native.COLOR_CONST, xform=0, mean_mode=_native.MEAN_AUTO, trails=False):
        """trails: False / True = the script's velocity trails (when it draws any), 2 = the Catmull-Rom history
        trails of traj_renderer.py (droplet path only)."""
        return _native.make_style(color_mode=color_mode, const_rgb=self.const_rgb, radius=self.radius,
                                  flip_x=self.flip_x, z_lift=self.z_lift, vel_norm=self.vel_norm, has_floor=True,
                                  floor_z=self.floor_z, floor_min=self.floor_min, floor_max=self.floor_max,
                                  floor_albedo=self.floor_albedo, light_z=self.light_z, light_half=self.light_half,
                                  radiance=self.radiance, bounce=self.bounce, xform=xform, mean_mode=mean_mode,
                                  trails=(2 if trails == 2 else int(bool(trails) and self.trail_schedule is not None)),
                                  trail_radius=self.trail_radius,
                                  trail_rgb=self.trail_rgb, trail_len_min=self.trail_len_min, trail_len_max=self.trail_len_max)

    def for_trajectory(self, n_frames):
        """Stretch the 220-frame schedule (199 motion + 20 fade) over an n_frames trajectory
        (deviation from the literals): the fade keeps 20 frames."""
        if self.keys is None or n_frames <= 0:
            return self
        fade = min(self.fade_frames, max(n_frames - 2, 1))
        return replace(self, last_motion_frame=max(n_frames - 1 - fade, 1), fade_frames=fade)


_BALL_KEYS = ((2.8, 2.8, 3.0), (1.8, 1.8, 1.8), (1.6, 1.6, 1.6))

PRESETS = {
    # example_renderer.py:16-31,55-62
    "example": RenderConfig("example", (0.0, 0.0, 0.0), 30.0, True, -0.2, (-10.0, -10.0), (10.0, 10.0), 256,
                            eye=(2.2, 2.2, 4.2)),
    # traj_renderer.py:20-35,66-72,519-527
    "traj": RenderConfig("traj", (0.0, 0.0, -0.05), 36.0, True, -0.5, (-10.0, -10.0), (10.0, 10.0), 256, dolly=True),
    # traj_ball_renderer.py:13-28,59-65,281-307
    "traj_ball": RenderConfig("traj_ball", (0.0, 0.0, -0.05), 36.0, True, -0.5, (-10.0, -10.0), (10.0, 10.0), 128,
                              keys=_BALL_KEYS, trail_schedule="ramp"),
    # traj_vel_renderer.py:13-28,59-65,381-407
    "traj_vel": RenderConfig("traj_vel", (0.0, 0.0, -0.05), 36.0, True, -0.5, (-10.0, -10.0), (10.0, 10.0), 128,
                             keys=_BALL_KEYS, trail_schedule="ramp_fade"),
    # traj_original.py:10-38,40-66 (TAIL inherited from traj_ball_renderer.py:59-65)
    "traj_original": RenderConfig("traj_original", (0.0, 0.0, -0.05), 36.0, False, -0.5, (-10.0, -10.0), (10.0, 10.0),
                                  128, eye=(-1.8, -1.8, 1.8), trail_schedule="const"),
    # traj_b0.py:10-60,84-115 — floor = [-1,1]^2 scaled by 20 then translated by (10,10,-0.8)
    "traj_b0": RenderConfig("traj_b0", (-0.02, 0.15, -0.05), 36.0, False, -0.8, (-10.0, -10.0), (30.0, 30.0), 128,
                            keys=((-2.2, -3.3, 2.0), (-1.3, -2.5, 0.8), (-1.0, -2.0, 0.7)), trail_schedule="const"),
    # traj_b1.py:10-60,84-115
    "traj_b1": RenderConfig("traj_b1", (0.0, -0.02, 0.0), 36.0, False, -0.8, (-10.0, -10.0), (30.0, 30.0), 128,
                            keys=((-3.5, -2.5, 2.8), (-2.3, -1.5, 1.2), (-2.0, -1.2, 1.0)), trail_schedule="const"),
}
