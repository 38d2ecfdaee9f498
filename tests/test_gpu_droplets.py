"""SURVEY.md §8f-2 on the GPU, through the C ABI: droplet matrices and Catmull-Rom history trails bit-identical
to what the reference prints / writes (tests/golden/droplets.npz), visibility keys of the droplet scene
(mesh instances = VA-3, polylines = VA-2) BIT-EXACT against oracle/raycast.c, images within tolerance.  `-m gpu`.

Image tolerance for this scene: <= 1 code value on 99.9 % of the pixels and PSNR >= 45 dB against the oracle's
f64 evaluation of the same shading model (the smooth normal is interpolated across ring bands: a hit that
falls within float error of a band edge may take the neighbouring band's slope)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from pointcloud_render_b200 import _native, droplets, renderers, synthetic  # noqa: E402
from pointcloud_render_b200.presets import PRESETS  # noqa: E402


@pytest.fixture(scope="module")
def ctx(lib):
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    c = _native.Context(device=0, max_points=1 << 16, max_w=1920, max_h=1080, max_batch=3)
    c.set_droplet_mesh(droplets.droplet_vertices(), droplets.N_RINGS, droplets.N_SEGMENTS)
    yield c
    c.close()


@pytest.fixture(scope="module")
def do():
    from oracle import droplet_oracle
    return droplet_oracle


@pytest.fixture(scope="module")
def g(golden):
    return golden("droplets.npz")


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def keys(vis):
    return vis.cpu().numpy().view(np.uint64)


def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


def check_image(got, want):
    d = np.abs(got.astype(int) - want.astype(int)).max(axis=-1)
    assert (d > 1).mean() <= 1e-3, f"{(d > 1).sum()} pixels differ by more than one code value (max {d.max()})"
    assert psnr(got, want) >= 45.0


def test_droplet_matrices_match_the_reference(ctx, do, g):
    """pcr_droplet_transforms against the matrices generate_rotation_matrix_from_velocity printed."""
    pcl6 = g["pcl6"]
    xf = ctx.droplet_transforms(dev(pcl6)).cpu().numpy()
    np.testing.assert_array_equal(bits(xf), bits(g["xf_velocity"]))
    rot = dev(droplets.random_rotations(64))
    xr = ctx.droplet_transforms(dev(pcl6[:64, :3]), rot=rot).cpu().numpy()
    np.testing.assert_array_equal(bits(xr), bits(g["xf_random"]))
    ident = ctx.droplet_transforms(dev(pcl6[:8, :3])).cpu().numpy().reshape(-1, 3, 4)
    np.testing.assert_array_equal(ident[:, :, :3], np.broadcast_to(np.eye(3, dtype=np.float32), (8, 3, 3)))
    # a larger random set against the oracle (itself pinned to the reference): the float64 libm of the device
    # (acos / cos / sin) may differ from numpy's in the last place, which survives the rounding to float32 in
    # about one element per 10^8 — allow one float32 ulp on at most 1e-5 of the elements
    rng = np.random.default_rng(9)
    big = (rng.standard_normal((200_000, 6)) * [0.3, 0.3, 0.3, 4, 4, 4]).astype(np.float32)
    got = ctx.droplet_transforms(dev(big)).cpu().numpy()
    want = do.to_world_f32(do.rotation_from_velocity(big[:, 3:6]), big[:, :3])
    diff = bits(got).astype(np.int64) - bits(want).astype(np.int64)
    assert np.abs(diff).max() <= 1 and (diff != 0).mean() <= 1e-5


def test_history_trails_match_the_reference_curve_files(ctx, do, g):
    """pcr_history_trails against the control points the reference wrote to its curve files, every history
    length the sampling plan distinguishes (and 25 > 20: only the last 20 frames count)."""
    from oracle.gen_golden import droplet_inputs
    for h in g["history_lengths"]:
        hist, pos = g[f"hist_{h}"], g[f"pos_{h}"]
        ctrl, cnt = ctx.history_trails(dev(hist), dev(pos))
        cnt = cnt.cpu().numpy()
        np.testing.assert_array_equal(cnt, g[f"count_{h}"], err_msg=f"h={h}")
        valid = np.arange(21)[None, :] < cnt[:, None]
        np.testing.assert_array_equal(bits(ctrl.cpu().numpy())[valid], bits(g[f"ctrl_{h}"])[valid], err_msg=f"h={h}")
    # larger random histories against the oracle
    for h in (2, 3, 9, 20):
        hist, pos = droplet_inputs(h, n=5000, seed=3)
        ctrl, cnt = ctx.history_trails(dev(hist), dev(pos))
        wc, wn = do.history_trails(hist, pos)
        np.testing.assert_array_equal(cnt.cpu().numpy(), wn)
        valid = np.arange(21)[None, :] < wn[:, None]
        np.testing.assert_array_equal(bits(ctrl.cpu().numpy())[valid], bits(wc)[valid])


def moving_trajectory(frames, n, cols, seed, dt=0.01, dtype=np.float32):
    """Like synthetic.trajectory but with velocities independent of the positions: synthetic.trajectory draws V
    from the stream that made P0 (V = 3 P0, a pure expansion), which the per-frame standardisation removes —
    the standardised points would stand still and the reference would draw no history trail at all."""
    rng = np.random.default_rng(seed)
    p0 = rng.standard_normal((n, 3))
    v = 3.0 * rng.standard_normal((n, 3))
    v[:4] = 0.0                                          # a few points that never move (no trail, identity rotation)
    g = np.array([0.0, -1.0, 0.0])
    out = np.empty((frames, n, cols), dtype)
    for f in range(frames):
        t = f * dt
        out[f, :, :3] = p0 + t * v + 0.5 * t * t * g
        if cols == 6:
            out[f, :, 3:6] = np.where(np.any(v != 0, axis=1, keepdims=True), v + t * g, 0.0)
    return out


def oracle_scene(orc, do, cfg, traj, f, n_hist_avail, cam_frame, total, W, H, trails, rot=None):
    """The scene the reference would emit for buffer frame f (history = the up to 20 frames before it), rendered
    by the oracle: keys and image."""
    cols = traj.shape[2]
    pcl = orc.transform_coordinates(orc.standardize_point_cloud(traj[f]), cfg.flip_x)
    n = pcl.shape[0]
    if cols == 6:
        R = do.rotation_from_velocity(pcl[:, 3:6])
    elif rot is not None:
        R = rot.reshape(-1, 3, 3).astype(np.float64)
    else:
        R = np.broadcast_to(np.eye(3), (n, 3, 3))
    xf = do.to_world_f32(R, pcl[:, :3])
    ctrl = count = None
    if cols == 6 and trails == 2:
        h0 = max(0, f - 20)
        hist = np.stack([orc.transform_coordinates(orc.standardize_point_cloud(traj[k]), cfg.flip_x)[:, :3] for k in range(h0, f)]) \
            if f > h0 else np.zeros((0, n, 3), np.float32)
        ctrl, count = do.history_trails(hist, pcl[:, :3])
    elif cols == 6 and trails == 1:
        tail, head, valid = orc.velocity_trails(pcl, cfg.trail_length_scale(cam_frame))
        ctrl = np.zeros((n, 21, 3), np.float32)
        ctrl[:, 0], ctrl[:, 1] = tail, head
        count = np.where(valid, 2, 0).astype(np.int32)
    fr = orc.camera_frame(cfg.camera_position(cam_frame, total), cfg.target, cfg.up, cfg.fov, cfg.near_clip, cfg.far_clip, W, H)
    sc = orc.make_scene(True, cfg.floor_z, cfg.floor_min, cfg.floor_max, cfg.floor_albedo, cfg.light_z, cfg.light_half, cfg.radiance, cfg.bounce)
    vis = orc.visibility(np.zeros((0, 4), np.float32), fr, sc)
    if ctrl is not None:
        vis = do.add_polylines(vis, ctrl, count, fr, n, radius=cfg.trail_radius)
    vis = do.add_droplets(vis, xf, fr)
    img = do.shade_droplet_scene(vis, xf, ctrl, count, fr, sc, rgb=cfg.const_rgb, trail_rgb=cfg.trail_rgb)
    return vis, img


@pytest.mark.parametrize("preset,trails,cols,dtype,n,W,H,n_hist,first", [
    ("traj", 2, 6, np.float32, 2048, 1024, 1024, 0, 0),       # C2's shape: frames 0..4, history grows from nothing
    ("traj", 2, 6, np.float32, 700, 640, 360, 22, 150),       # full 20-frame history behind a halo, camera close
    ("traj", 2, 6, np.float64, 500, 333, 211, 7, 40),         # float64 input, ragged film, short halo
    ("traj_vel", 1, 6, np.float32, 1500, 800, 600, 0, 10),    # traj_vel_renderer: straight velocity trails (ramp-in)
    ("traj_vel", 1, 6, np.float32, 900, 640, 480, 0, 205),    # ... fade-out
    ("traj", 2, 3, np.float32, 600, 512, 512, 3, 30),         # no velocity: random rotations, no trails
    ("traj_vel", 1, 3, np.float32, 600, 512, 512, 0, 30),     # no velocity in the vel script: identity, no trails
])
def test_droplet_scene_matches_oracle(ctx, orc, do, preset, trails, cols, dtype, n, W, H, n_hist, first):
    import dataclasses
    cfg = PRESETS[preset]
    if trails == 2:
        cfg = dataclasses.replace(cfg, trail_radius=0.0015)      # the reference's 0.0007 covers few pixel centres
    F = 5
    traj = moving_trajectory(n_hist + F, n, cols, seed=n, dtype=dtype)
    cams = [cfg.camera(first + k, 220, W, H) for k in range(F)]
    style = cfg.style(trails=trails if trails == 2 else True)
    assert style.trails == trails
    rot = droplets.random_rotations(n) if (cols == 3 and trails == 2) else None
    rgba, vis = ctx.render_droplet_frames(dev(traj), cams, style, n_history=n_hist, rot=None if rot is None else dev(rot), want_vis=True)
    saw_trail = saw_droplet = False
    for k in range(F):
        want, img = oracle_scene(orc, do, cfg, traj, n_hist + k, n_hist, first + k, 220, W, H, trails, rot)
        got = keys(vis[k])
        bad = np.argwhere(got != want)
        assert len(bad) == 0, f"frame {k}: {len(bad)} pixels differ, first {bad[:3].tolist()}: got {got[tuple(bad[0])]:#x} want {want[tuple(bad[0])]:#x}"
        ids = (want & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        saw_droplet |= bool(np.any(ids < n))
        saw_trail |= bool(np.any((ids >= n) & (ids < 2 * n)))
        check_image(rgba[k].cpu().numpy(), img)
    assert saw_droplet
    if cols == 3:
        assert not saw_trail
    else:
        assert saw_trail


def test_droplet_occlusion_cull_never_changes_a_key(lib, orc, do):
    """Dense droplet scene (C4D's shape, shrunk): the occluder pre-pass + Hi-Z cull of pcr_render_droplet_frames only skips
    buried droplets and trail segments — keys and image identical with the cull forced on (every 4th / 16th droplet as
    occluders) and off, and equal to the oracle's."""
    n, W, H, F = 12_000, 640, 360, 2
    cfg = PRESETS["traj_vel"]
    traj = moving_trajectory(F, n, 6, seed=77)
    traj[:, :, :3] *= 0.25                                   # (per-frame standardisation undoes this; kept for a non-unit scale)
    cams = [cfg.camera(140 + k, 220, W, H) for k in range(F)]
    style = cfg.style(trails=True)
    outs = []
    for mode, step in ((0, 0), (1, 4), (1, 16)):
        c = _native.Context(device=0, max_points=n, max_w=W, max_h=H, max_batch=2)
        try:
            c.set_droplet_mesh(droplets.droplet_vertices(), droplets.N_RINGS, droplets.N_SEGMENTS)
            c.set_occlusion(mode=mode, step=step)
            rgba, vis = c.render_droplet_frames(dev(traj), cams, style, want_vis=True)
            outs.append((keys(vis).copy(), rgba.cpu().numpy()))
        finally:
            c.close()
    for k, im in outs[1:]:
        np.testing.assert_array_equal(k, outs[0][0])
        np.testing.assert_array_equal(im, outs[0][1])
    want, img = oracle_scene(orc, do, cfg, traj, 1, 0, 141, 220, W, H, 1)
    np.testing.assert_array_equal(outs[2][0][1], want)
    check_image(outs[2][1][1], img)
    ids = (want & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    visible = len(np.unique(ids[ids < n]))
    assert visible < 0.6 * n, f"the scene should bury a good part of its droplets (visible: {visible} of {n})"


def test_facade_droplet_renderers(tmp_path, orc, do):
    """TrajectoryRenderer / TrajectoryVelRenderer: process() with history (the reference's per-frame entry) gives
    the same image as the batched whole path, files are named like the reference's, nothing else is written."""
    import os
    n, W, H = 400, 320, 240
    traj = moving_trajectory(6, n, 6, seed=5)
    r = renderers.TrajectoryRenderer(str(tmp_path / "frame_0005_b0.npy"), output_folder=str(tmp_path / "render"), width=W, height=H)
    np.save(tmp_path / "frame_0005_b0.npy", traj[5])
    hist = [r.transform_coordinates(r.standardize_point_cloud(traj[k])) for k in range(5)]
    r.process(frame_index=5, history_pcls=hist, total_frames=220)
    from PIL import Image
    img = np.asarray(Image.open(tmp_path / "render" / "frame_0005_b0.png").convert("RGBA"))
    batch, vis = r.render_trajectory(traj, first_frame=0, total_frames=220, stretch=False, want_vis=True)
    np.testing.assert_array_equal(img, batch[5].numpy())
    ids = _native.keys_to_ids(vis[5])
    assert np.any(ids < n) and np.any((ids >= n) & (ids < 2 * n))
    assert sorted(os.listdir(tmp_path)) == ["frame_0005_b0.npy", "render"]           # no temp_meshes / temp_curves
    # frame-sharded: rank 1 renders frames 3..5 with a 3-frame halo -> identical images
    part = r.render_trajectory(traj, first_frame=3, total_frames=220, stretch=False, n_history=3)
    np.testing.assert_array_equal(part.numpy(), batch[3:].numpy())
    m = r.generate_rotation_matrix_from_velocity([0.0, 0.0, -3.0], [1.0, 2.0, 3.0]).reshape(4, 4)
    np.testing.assert_array_equal(m, [[1, 0, 0, 1], [0, 1, 0, 2], [0, 0, 1, 3], [0, 0, 0, 1]])
    v = renderers.TrajectoryVelRenderer(None, width=W, height=H)
    a = v.render_trajectory(traj[:2], first_frame=100, total_frames=220, stretch=False)
    b = renderers.TrajectoryVelRenderer(None, width=W, height=H, droplets=False).render_trajectory(
        torch.from_numpy(traj[:2]).cuda(), first_frame=100, total_frames=220, stretch=False)
    assert a.shape == (2, H, W, 4) and not np.array_equal(a.numpy(), b.cpu().numpy())     # droplets are not spheres


def test_droplet_entry_rejects_bad_arguments(lib):
    c = _native.Context(device=0, max_points=1024, max_w=64, max_h=64, max_batch=2)
    cfg = PRESETS["traj"]
    x = torch.zeros((1, 16, 6), dtype=torch.float32, device="cuda")
    with pytest.raises(RuntimeError, match="pcr_set_droplet_mesh"):
        c.render_droplet_frames(x, [cfg.camera(0, 220, 64, 48)], cfg.style(trails=2))
    with pytest.raises(RuntimeError):
        c.set_droplet_mesh(np.zeros((340, 3), np.float32), 16, 20)                  # ring z does not decrease
    c.close()
