"""pointcloud_render_b200 — B200-native sphere-splat renderer for the hot path of
EvaShenLu/PointCloud_Render (example_renderer.py / traj_*_renderer.py): standardise ->
axis transform -> colour hook -> projection + tile binning -> sphere visibility -> shading.

The compute lives in libpcr.so (hand-written sm_100a CUDA behind the C ABI of include/pcr.h);
this package is the host-side mirror of the reference's renderer classes.  No CPU fallback.
"""
from . import _native, presets, synthetic  # noqa: F401
from ._native import (COLOR_CONST, COLOR_POSITION, COLOR_USER, COLOR_VELOCITY, ID_FLOOR, ID_MISS, KEY_MISS,  # noqa: F401
                      Context, load_library, make_camera, make_style)
from .output import AsyncImageWriter, trajectory_frame_name  # noqa: F401
from .presets import PRESETS, RenderConfig  # noqa: F401
from .renderers import (FixedFrame199Renderer, PointCloudRenderer, RenderedScene, TrajB0Renderer, TrajB1Renderer,  # noqa: F401
                        TrajectoryBallRenderer, TrajectoryRenderer, TrajectoryVelRenderer, release_engines)

__version__ = "0.1.0"
