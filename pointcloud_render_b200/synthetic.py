"""Seeded synthetic point clouds and trajectories.

The reference ships no sample data (its `ply/`, `trajectory_ply/`, `batch_0/` folders are
referenced at example_renderer.py:206-207 / traj_ball_renderer.py:423-424 but absent), so
tests and bench.py use these shapes.  numpy only; nothing here touches the GPU.
"""
import numpy as np

# named workloads of BASELINE.json:configs (+ H, the headline of BASELINE.json:metric)
CONFIGS = {
    # name: (frames, points, cols, W, H, preset, color_mode, per-point radius)
    "C1": dict(frames=1, points=2048, cols=3, width=800, height=600, preset="example", color_mode=0, radii=False),
    "C2": dict(frames=100, points=2048, cols=3, width=1024, height=1024, preset="traj", color_mode=1, radii=False),
    "C3": dict(frames=500, points=16384, cols=6, width=1024, height=1024, preset="traj_b0", color_mode=0, radii=True),
    "C4": dict(frames=1000, points=100000, cols=6, width=1920, height=1080, preset="traj_vel", color_mode=2, radii=False),
    "C5": dict(frames=1, points=50_000_000, cols=3, width=4096, height=4096, preset="example", color_mode=0, radii=False),
    # the droplet scenes of the same two scripts: what traj_renderer.py / traj_vel_renderer.py really draw
    "C2D": dict(frames=100, points=2048, cols=6, width=1024, height=1024, preset="traj", color_mode=0, radii=False, droplet_trails=2),
    "C4D": dict(frames=1000, points=100000, cols=6, width=1920, height=1080, preset="traj_vel", color_mode=0, radii=False, droplet_trails=1),
    "H": dict(frames=100, points=1_000_000, cols=3, width=1024, height=1024, preset="traj_ball", color_mode=0, radii=False),
}


def _cloud(rng, n, shape):
    if shape == "gauss":
        p = rng.standard_normal((n, 3))
    elif shape == "cube":
        p = rng.random((n, 3))
    elif shape == "shell":
        g = rng.standard_normal((n, 3))
        g /= np.linalg.norm(g, axis=1, keepdims=True)
        p = g * (1.0 + 0.02 * rng.standard_normal((n, 1)))
    else:
        raise ValueError(shape)
    return p


def cloud(n, shape="gauss", seed=0, dtype=np.float32):
    """(n,3) cloud: 'gauss' = standard normal, 'cube' = uniform [0,1)^3, 'shell' = thin sphere surface."""
    return np.ascontiguousarray(_cloud(np.random.default_rng(seed), n, shape), dtype=dtype)


def trajectory(frames, n, cols=6, shape="gauss", seed=0, dt=0.01, dtype=np.float32, frame_indices=None):
    """(frames, n, cols) ballistic trajectory: P_f = P_0 + f dt V + 0.5 (f dt)^2 g, V = 3 N(0,1),
    g = (0,-1,0) in input axes; velocity columns = V + f dt g (cols == 6).  frame_indices: the physics frame f of
    every returned frame (default 0..frames-1)."""
    rng = np.random.default_rng(seed)
    p0 = _cloud(rng, n, shape)                    # P0 (== cloud(n, shape, seed)) and V come from ONE generator,
    v = 3.0 * rng.standard_normal((n, 3))         # so V is independent of P0
    g = np.array([0.0, -1.0, 0.0])
    idx = list(range(frames)) if frame_indices is None else [int(f) for f in frame_indices]
    assert len(idx) == frames
    out = np.empty((frames, n, cols), dtype=dtype)
    for k, f in enumerate(idx):
        t = f * dt
        out[k, :, :3] = p0 + t * v + 0.5 * t * t * g
        if cols == 6:
            out[k, :, 3:6] = v + t * g
    return out


def radii(n, seed=0, base=0.01):
    """Per-point ball radius r = base * U(0.5,1.5) — an extension; the reference's radius is the
    literal 0.01 (traj_ball_renderer.py:39)."""
    rng = np.random.default_rng(seed + 7919)
    return (base * rng.uniform(0.5, 1.5, n)).astype(np.float32)
