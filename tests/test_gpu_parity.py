"""Parity of the CUDA path (through the C ABI, libpcr.so) with the CPU oracle.  `-m gpu`.

Bars: visibility keys (depth bits | point id) BIT-EXACT; standardised positions BIT-EXACT against
the reference's numpy arithmetic in mean_mode SEQUENTIAL (= AUTO, the default, at every size),
and in mean_mode F64 bit-exact against the order-independent definition (mean summed in
f64, rounded once) and within ref_atol() of the reference; velocities / min / max / scale exact;
sRGB8 images within 1 code value and PSNR >= 50 dB of the oracle's f64 evaluation of the same
shading model."""

import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from pointcloud_render_b200 import _native, synthetic  # noqa: E402
from pointcloud_render_b200.presets import PRESETS  # noqa: E402


def ref_atol(x):
    """Tolerance against the reference's own float arithmetic.  Its np.mean(axis=0) is a sequential
    sum in the input dtype (oracle/pcr_oracle.py:standardize_point_cloud), whose rounding error is
    at most (N-1) * eps * sum|x| / N per axis — in practice ~sqrt(N) * eps * max|x|.  The CUDA path
    sums in f64 and rounds once, so it differs from the reference by the REFERENCE's summation
    error divided by the scale, plus one final f32 rounding (2^-24 at |value| <= 1)."""
    x = np.asarray(x)[:, :3]
    eps = np.finfo(x.dtype).eps
    scale = float(np.amax(x - np.amin(x, axis=0)))
    return 2.0 ** -23 + 2.0 * np.sqrt(len(x)) * eps * float(np.abs(x).max()) / scale


@pytest.fixture(scope="module", params=["auto", "occlusion-always"])
def ctx(lib, request):
    """Every test runs twice: default settings (occlusion pre-pass only for n >= 131072) and with
    the pre-pass forced on for every cloud size (step 4 so that even tiny clouds exercise it)."""
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    c = _native.Context(device=0, max_points=1 << 20, max_w=1920, max_h=1080, max_batch=4)
    if request.param == "occlusion-always":
        c.set_occlusion(mode=1, step=4)
    yield c
    c.close()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def keys(vis):
    return vis.cpu().numpy().view(np.uint64)


def orc_scene(orc, cfg):
    return orc.make_scene(True, cfg.floor_z, cfg.floor_min, cfg.floor_max, cfg.floor_albedo, cfg.light_z, cfg.light_half,
                          cfg.radiance, cfg.bounce)


def orc_frame(orc, cfg, frame_index, total, W, H):
    return orc.camera_frame(cfg.camera_position(frame_index, total), cfg.target, cfg.up, cfg.fov, cfg.near_clip,
                            cfg.far_clip, W, H)


def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


def check_image(got, want):
    d = np.abs(got.astype(int) - want.astype(int))
    assert d.max() <= 1, f"max abs diff {d.max()}"
    assert psnr(got, want) >= 50.0


# ------------------------------------------------------------------------------- K0 + K1
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("cols", [3, 6])
@pytest.mark.parametrize("n", [2, 257, 100_003])
@pytest.mark.parametrize("preset", ["traj_ball", "traj_b0"])
def test_standardize_transform(ctx, orc, dtype, cols, n, preset):
    rng = np.random.default_rng(n + cols)
    x = np.ascontiguousarray(rng.standard_normal((n, cols)) * [1, 0.6, 1.7, 3, 3, 3][:cols] + [0.3, -2, 5, 0, 0, 0][:cols], dtype=dtype)
    cfg = PRESETS[preset]
    want = orc.transform_coordinates(orc.standardize_point_cloud(x), flip_x=cfg.flip_x)
    exact = orc.transform_coordinates(orc.standardize_point_cloud(x, exact_mean=True), flip_x=cfg.flip_x)
    # the reference's own arithmetic (sequential sum in the input dtype): bit-exact, also the default here (n <= 131072)
    for mode in (_native.MEAN_SEQUENTIAL, _native.MEAN_AUTO):
        got = ctx.standardize(dev(x), cfg.style(mean_mode=mode))[0].cpu().numpy()
        np.testing.assert_array_equal(got[:, :3], want[:, :3])
    pos4, attr4, vel4, stats = ctx.standardize(dev(x), cfg.style(mean_mode=_native.MEAN_F64), want_vel=True, want_stats=True)
    pos4, attr4, stats = pos4.cpu().numpy(), attr4.cpu().numpy(), stats.cpu().numpy()
    np.testing.assert_array_equal(pos4[:, :3], exact[:, :3])            # order-independent definition: bit-exact
    np.testing.assert_allclose(pos4[:, :3], want[:, :3], rtol=0, atol=ref_atol(x))   # vs the reference's sequential mean
    assert np.all(pos4[:, 3] == np.float32(0.01))
    np.testing.assert_array_equal(attr4[:, :3], np.float32(0.3))
    np.testing.assert_array_equal(stats[3:6], x[:, :3].min(0).astype(np.float64))
    np.testing.assert_array_equal(stats[6:9], x[:, :3].max(0).astype(np.float64))
    assert stats[9] == float(np.amax(x[:, :3] - np.amin(x[:, :3], axis=0)))
    if cols == 6:
        np.testing.assert_array_equal(vel4.cpu().numpy()[:, :3], want[:, 3:6])
        np.testing.assert_array_equal(attr4[:, 3], orc.compute_color(want)[:, 3])


@pytest.mark.parametrize("dtype,cols", [(np.float32, 3), (np.float32, 6), (np.float64, 3), (np.float64, 6)])
def test_sequential_mean_every_alignment_and_lane(lib, orc, dtype, cols):
    """k_mean_sequential (three axis warps, lane = frame, bulk copies of the 16-byte aligned interior of every stage): the
    reference's np.mean bit for bit for frames that start at any offset inside a 16-byte granule, for point counts around
    the stage and register-set boundaries, and for every lane of a block (frames 0..10 = a full block of 8 + a ragged one).
    Per frame: stats through pcr_standardize on the frame's slice (one lane); all frames at once: pcr_render_frames keys
    equal to the keys of the per-frame two-step renders."""
    cfg = PRESETS["traj_ball"]
    W, H, F = 96, 64, 11
    stage_pts = 6144 // (cols * np.dtype(dtype).itemsize)           # MEAN_STAGE_BYTES
    c = _native.Context(device=0, max_points=8192, max_w=W, max_h=H, max_batch=16)
    try:
        for n in sorted({1, 2, 3, 5, 15, 16, 17, 31, 33, stage_pts - 1, stage_pts, stage_pts + 1, 2 * stage_pts + 7, 4 * stage_pts,
                         5 * stage_pts + 31, 1001, 2048, 4095} - {0}):
            if n > 8192:
                continue
            rng = np.random.default_rng(n * 7 + cols)
            traj = np.ascontiguousarray(rng.standard_normal((F, n, cols)) * 0.7 + [3.0, -1.0, 0.25, 0, 0, 0][:cols], dtype=dtype)
            d = dev(traj)
            style = cfg.style(mean_mode=_native.MEAN_SEQUENTIAL)
            for f in range(F):
                if n > 64 and f not in (0, 1, 2, 3, 7, 8, 10):
                    continue
                stats = c.standardize(d[f], style, want_stats=True)[-1].cpu().numpy()
                want = traj[f][:, :3].mean(axis=0)                      # numpy: sequential sum in the input dtype, one division
                np.testing.assert_array_equal(stats[:3], want.astype(np.float64), err_msg=f"n={n} frame {f}")
            if n >= 2:
                cams = [cfg.camera(3 * f, 40, W, H) for f in range(F)]
                vis = c.render_frames(d, cams, style, want_vis=True)[1]
                for f in (0, 5, 7, 8, 10):
                    pos4, attr4 = c.standardize(d[f], style)
                    np.testing.assert_array_equal(pos4.cpu().numpy()[:, :3], orc.transform_coordinates(orc.standardize_point_cloud(traj[f]), cfg.flip_x)[:, :3])
                    assert torch.equal(c.render(pos4, attr4, cams[f], style)[0], vis[f]), f"n={n} frame {f}"
    finally:
        c.close()


def test_standardize_golden_inputs(ctx, orc, golden):
    """The reference's own outputs (tests/golden/standardize.npz)."""
    g = golden("standardize.npz")
    for tag in ("f32_3", "f64_3", "f32_6", "f64_6"):
        x = g[f"in_{tag}"]
        for short, preset in (("ball", "traj_ball"), ("b0", "traj_b0"), ("orig", "traj_original")):
            cfg = PRESETS[preset]
            out = ctx.standardize(dev(x), cfg.style(), want_vel=True)
            want = g[f"xf_{short}_{tag}"]
            np.testing.assert_array_equal(out[0].cpu().numpy()[:, :3], want[:, :3])      # the reference's output, bit for bit
            if x.shape[1] == 6:
                np.testing.assert_array_equal(out[2].cpu().numpy()[:, :3], want[:, 3:6])
            std = ctx.standardize(dev(x), cfg.style(xform=1))[0].cpu().numpy()[:, :3]
            np.testing.assert_array_equal(std, g[f"std_{short}_{tag}"][:, :3])
            std64 = ctx.standardize(dev(x), cfg.style(xform=1, mean_mode=_native.MEAN_F64))[0].cpu().numpy()[:, :3]
            np.testing.assert_array_equal(std64, orc.standardize_point_cloud(x, exact_mean=True)[:, :3])
            np.testing.assert_allclose(std64, std, rtol=0, atol=ref_atol(x))
            # transform_coordinates alone is an exact permutation + one f32 add
            xf = ctx.transform_coordinates(dev(g[f"std_{short}_{tag}"]), flip_x=cfg.flip_x).cpu().numpy()
            np.testing.assert_array_equal(xf, want)


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_colour_hook_extensions(ctx, orc, mode):
    rng = np.random.default_rng(mode)
    x = rng.standard_normal((5000, 6)).astype(np.float32) * np.float32(4)
    cfg = PRESETS["traj_vel"]
    rgb = rng.random((5000, 3)).astype(np.float32)
    pos4, attr4 = ctx.standardize(dev(x), cfg.style(color_mode=mode), rgb=dev(rgb) if mode == 3 else None)
    pos = pos4.cpu().numpy()
    # the hook is evaluated on the device's own transformed positions: feed those to the oracle
    vel = orc.transform_coordinates(orc.standardize_point_cloud(x), True)[:, 3:6]
    want = orc.compute_color(np.concatenate([pos[:, :3], vel], axis=1), mode=mode, user_rgb=rgb)
    got = attr4.cpu().numpy()
    if mode == 1:
        np.testing.assert_allclose(got[:, :3], want[:, :3], rtol=0, atol=3e-6)   # min/max come from the raw stats path
    else:
        np.testing.assert_array_equal(got, want)


def test_per_point_radius(ctx):
    x = synthetic.cloud(1000)
    r = synthetic.radii(1000)
    pos4, _ = ctx.standardize(dev(x), PRESETS["traj_b0"].style(), radius=dev(r))
    np.testing.assert_array_equal(pos4.cpu().numpy()[:, 3], r)


# ------------------------------------------------------------------------------- K2 + K3
def render_case(ctx, orc, pos4, cfg, frame_index, total, W, H, id_base=0, c=None):
    c = c or ctx
    cam = cfg.camera(frame_index, total, W, H)
    attr4 = np.concatenate([np.random.default_rng(1).random((len(pos4), 3), dtype=np.float32),
                            np.zeros((len(pos4), 1), np.float32)], axis=1)
    vis, rgba = c.render(dev(pos4), dev(attr4), cam, cfg.style(), id_base=id_base)
    fr, sc = orc_frame(orc, cfg, frame_index, total, W, H), orc_scene(orc, cfg)
    want = orc.visibility(pos4, fr, sc, id_base=id_base)
    got = keys(vis)
    bad = np.argwhere(got != want)
    assert len(bad) == 0, f"{len(bad)} pixels differ, first {bad[:3].tolist()}: got {got[tuple(bad[0])]:#x} want {want[tuple(bad[0])]:#x}"
    check_image(rgba.cpu().numpy(), orc.shade(want, pos4, attr4, fr, sc, id_base=id_base))
    return got


def test_visibility_golden_example_scene_c1(ctx, orc, golden):
    """Config C1: the scene the reference's generate_xml_content emitted (centres parsed back out
    of its XML), 800x600, bit-exact keys; plus the native 1920x1080 film."""
    import hashlib
    g = golden("scene_example.npz")
    pos4 = np.concatenate([g["centers"], g["radius"][:, None]], axis=1).astype(np.float32)
    got = render_case(ctx, orc, pos4, PRESETS["example"], 0, 1, 800, 600)
    assert hashlib.sha256(got.tobytes()).hexdigest() == str(golden("vis_example.npz")["sha256_800x600"])
    render_case(ctx, orc, pos4, PRESETS["example"], 0, 1, 1920, 1080)


@pytest.mark.parametrize("name", ["traj_ball", "traj_original", "traj_b0", "traj_b1"])
def test_visibility_golden_traj_scenes(ctx, orc, golden, name):
    g = golden(f"scene_{name}.npz")
    pos4 = np.concatenate([g["centers"], g["radius"][:, None]], axis=1).astype(np.float32)
    render_case(ctx, orc, pos4, PRESETS[name], int(g["frame"]), 220, 1920, 1080)
    render_case(ctx, orc, pos4, PRESETS[name], int(g["frame"]), 220, 1024, 1024)


@pytest.mark.parametrize("preset,frame_index,W,H,n,shape", [
    ("traj", 0, 1024, 1024, 2048, "gauss"),          # C2 far
    ("traj", 99, 1024, 1024, 2048, "gauss"),         # C2 near (camera dollies into the cloud)
    ("traj_b0", 250, 1024, 1024, 16384, "gauss"),    # C3
    ("traj_b1", 499, 1024, 1024, 16384, "shell"),
    ("traj_vel", 700, 1920, 1080, 100_000, "gauss"),  # C4
    ("traj_ball", 50, 1024, 1024, 300_000, "cube"),
    ("example", 0, 333, 77, 5000, "gauss"),          # ragged W,H (not multiples of the tile)
    ("traj_original", 0, 17, 1080, 3000, "shell"),
    ("traj_ball", 10, 1920, 16, 3000, "gauss"),
])
def test_visibility_matches_oracle(ctx, orc, preset, frame_index, W, H, n, shape):
    cfg = PRESETS[preset]
    total = {"traj": 100, "traj_b0": 500, "traj_b1": 500, "traj_vel": 1000}.get(preset, 220)
    cfg = cfg.for_trajectory(total)
    x = synthetic.cloud(n, shape, seed=n % 97)
    p = orc.transform_coordinates(orc.standardize_point_cloud(x), cfg.flip_x)
    r = synthetic.radii(n, seed=3) if preset in ("traj_b0", "traj_b1") else np.full(n, 0.01, np.float32)
    render_case(ctx, orc, np.concatenate([p, r[:, None]], axis=1), cfg, frame_index, total, W, H)


def test_visibility_edge_cases(ctx, orc):
    cfg = PRESETS["traj_ball"]
    W, H = 320, 200
    # empty cloud: floor / miss only
    render_case(ctx, orc, np.zeros((0, 4), np.float32), cfg, 0, 220, W, H)
    # one point; duplicates (ties -> lower id); id_base offset
    render_case(ctx, orc, np.array([[0, 0, 0, 0.05]], np.float32), cfg, 0, 220, W, H)
    got = render_case(ctx, orc, np.array([[0, 0, 0, 0.2]] * 5, np.float32), cfg, 0, 220, W, H, id_base=1000)
    ids = (got & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    assert 1000 in ids and not np.any((ids > 1000) & (ids < 0xFFFFFFFE))
    # spheres behind the eye, straddling the near plane, enclosing the eye, below the floor, huge, sub-pixel
    eye = np.array(cfg.camera_position(0, 220), np.float32)
    d = (np.array(cfg.target, np.float32) - eye)
    d /= np.linalg.norm(d)
    pts = [list(eye - d * 1.0) + [0.3], list(eye + d * 0.1) + [0.05], list(eye + d * 0.12) + [0.05], list(eye) + [0.5],
           [0, 0, -2.0, 0.3], [0, 0, 0.2, 1.5], [0.3, 0.1, 0, 1e-4], [0.31, 0.1, 0, 0.0], [5, 5, 0, 0.5], [-30, 40, 0, 3.0]]
    render_case(ctx, orc, np.array(pts, np.float32), cfg, 0, 220, W, H)
    # non-finite centres are skipped by the binner and by the oracle alike
    pts = np.array([[np.nan, 0, 0, 0.1], [0, np.inf, 0, 0.1], [0, 0, 0, 0.1]], np.float32)
    render_case(ctx, orc, pts, cfg, 0, 220, W, H)


def test_pair_capacity_overflow_takes_the_unbinned_raster(orc, lib):
    """More (tile,sphere) pairs than pair_capacity: the frame falls back to the per-sphere
    atomicMin raster; the keys are the same."""
    c = _native.Context(device=0, max_points=20000, max_w=640, max_h=480, max_batch=1, pair_capacity=1000)
    try:
        cfg = PRESETS["traj_ball"]
        p = orc.transform_coordinates(orc.standardize_point_cloud(synthetic.cloud(20000)), True)
        pos4 = np.concatenate([p, np.full((20000, 1), 0.01, np.float32)], axis=1)
        render_case(c, orc, pos4, cfg, 100, 220, 640, 480, c=c)
        assert c.counters()["overflow_frames"] == 1
        render_case(c, orc, pos4[:100], cfg, 100, 220, 640, 480, c=c)       # and recovers on the next frame
        assert c.counters()["overflow_frames"] == 0
    finally:
        c.close()


def test_capacity_and_argument_errors(ctx):
    cfg = PRESETS["example"]
    with pytest.raises(RuntimeError, match="larger than the context"):
        ctx.render(torch.zeros((4, 4), device="cuda"), torch.zeros((4, 4), device="cuda"), cfg.camera(0, 1, 4096, 4096), cfg.style())
    big = torch.zeros(((1 << 20) + 1, 3), device="cuda")
    with pytest.raises(RuntimeError, match="max_points"):
        ctx.standardize(big, cfg.style())
    with pytest.raises(ValueError, match="shape"):                      # the wrapper refuses it ...
        ctx.standardize(torch.zeros((8, 4), device="cuda"), cfg.style())
    import ctypes                                                       # ... and so does the C entry behind it
    z, st = torch.zeros((8, 4), device="cuda"), cfg.style()
    out = torch.zeros((8, 4), device="cuda")
    rc = ctx.lib.pcr_standardize(ctx.handle, ctypes.c_void_p(z.data_ptr()), 0, 8, 4, None, None, ctypes.byref(st),
                                 ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(out.data_ptr()), None, None, None)
    assert rc == -1 and b"cols" in ctx.lib.pcr_last_error(ctx.handle)
    cam = _native.make_camera((1, 1, 1), (1, 1, 1), width=64, height=64)
    with pytest.raises(RuntimeError, match="degenerate camera"):
        ctx.render(torch.zeros((4, 4), device="cuda"), torch.zeros((4, 4), device="cuda"), cam, cfg.style())


# ------------------------------------------------------------------------------- whole path
@pytest.mark.parametrize("cfgname", ["C2", "C3", "C4"])
def test_trajectory_frames_match_oracle(ctx, orc, cfgname):
    """pcr_render_frames (batched K0..K4) on a few frames of each trajectory config, device and
    host-buffer entries, against the oracle frame by frame."""
    c = synthetic.CONFIGS[cfgname]
    F, n, cols, W, H = 6, min(c["points"], 30000), c["cols"], c["width"], c["height"]
    traj = synthetic.trajectory(F, n, cols, seed=2)
    radius = synthetic.radii(n) if c["radii"] else None
    cfg = PRESETS[c["preset"]].for_trajectory(c["frames"])
    first = c["frames"] - F                       # the last frames: camera closest, fade branch included
    cams = [cfg.camera(first + f, c["frames"], W, H) for f in range(F)]
    style = cfg.style(color_mode=c["color_mode"])
    rgba, vis = ctx.render_frames(dev(traj), cams, style, radius=None if radius is None else dev(radius), want_vis=True)
    sc = orc_scene(orc, cfg)
    rgba, vis = rgba.cpu().numpy(), keys(vis)
    for f in range(F):
        pos4, attr4 = ctx.standardize(dev(traj[f]), style, radius=None if radius is None else dev(radius))
        pos4, attr4 = pos4.cpu().numpy(), attr4.cpu().numpy()
        want_pos = orc.transform_coordinates(orc.standardize_point_cloud(traj[f]), cfg.flip_x)
        np.testing.assert_array_equal(pos4[:, :3], want_pos[:, :3])      # n <= 131072: the reference's sequential mean
        fr = orc_frame(orc, cfg, first + f, c["frames"], W, H)
        want = orc.visibility(pos4, fr, sc)      # same f32 centres on both sides -> bit-exact keys
        np.testing.assert_array_equal(vis[f], want)
        check_image(rgba[f], orc.shade(want, pos4, attr4, fr, sc))
    host_vis = torch.empty((F, H, W), dtype=torch.int64).pin_memory()
    host_rgba = ctx.render_frames_host(torch.from_numpy(traj).pin_memory(), cams, style, radius_host=radius, out_vis=host_vis)
    np.testing.assert_array_equal(host_vis.numpy().view(np.uint64), vis)
    np.testing.assert_array_equal(host_rgba.numpy(), rgba)


def test_point_sharded_merge_equals_unsharded(ctx, orc):
    """C5's path on one device: k point shards, each rendered with its id_base into its own
    z-buffer, merged with pcr_zmin (the local half of the uint64 min all-reduce), owner-only
    shading assembled with a byte MAX — identical to the unsharded render, bit for bit."""
    from pointcloud_render_b200 import sharding
    n, W, H, k = 200_000, 1024, 1024, 4
    cfg = PRESETS["example"]
    x = synthetic.cloud(n, "gauss", 1)
    # (the point-sharded entries take the parallel float64 mean: a float32 fold cannot be split across shards)
    cam, style = cfg.camera(0, 1, W, H), cfg.style(color_mode=1, mean_mode=_native.MEAN_F64)
    pos4, attr4, stats = ctx.standardize(dev(x), style, want_stats=True)
    vis_full, rgba_full = ctx.render(pos4, attr4, cam, style)
    merged, parts = None, []
    total = np.zeros(9)
    shards = [sharding.point_shard(n, r, k) for r in range(k)]
    partials = [ctx.stats_partial(dev(x[a:b])).cpu().numpy() for a, b in shards]
    total[:3] = np.sum([p[:3] for p in partials], axis=0)
    total[3:6] = np.min([p[3:6] for p in partials], axis=0)
    total[6:9] = np.max([p[6:9] for p in partials], axis=0)
    gstats = sharding.finalize_stats(total, n, np.float32)
    np.testing.assert_array_equal(gstats[3:], stats.cpu().numpy()[3:])
    np.testing.assert_allclose(gstats[:3], stats.cpu().numpy()[:3], rtol=1e-6)
    gstats_d = stats                                             # use the device's own mean so centres are identical
    for (a, b) in shards:
        p4, a4 = ctx.standardize_with_stats(dev(x[a:b]), style, gstats_d)
        np.testing.assert_array_equal(p4.cpu().numpy(), pos4[a:b].cpu().numpy())
        v, _ = ctx.render(p4, a4, cam, style, id_base=a, shade=False)
        # the fused shard entry (K1 inlined in K2a, nothing materialised) gives the same keys
        assert torch.equal(ctx.render_shard(dev(x[a:b]), gstats_d, cam, style, id_base=a), v)
        parts.append((a, p4, a4))
        merged = v if merged is None else ctx.zmin_(merged, v)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(keys(merged), keys(vis_full))
    img = None
    for (a, p4, a4) in parts:
        part = ctx.shade(merged, p4, a4, cam, style, id_base=a, owner_only=True)
        assert torch.equal(ctx.shade_shard(merged, dev(x[a:a + p4.shape[0]]), gstats_d, cam, style, id_base=a, owner_only=True), part)
        img = part if img is None else torch.maximum(img, part)
    np.testing.assert_array_equal(img.cpu().numpy(), rgba_full.cpu().numpy())


def test_headline_size_properties(ctx, orc):
    """H (1 M points, 1024^2): oracle comparison on the full size (the bbox-accelerated oracle takes
    well under a second) plus size-independent properties: idempotence, every id valid, each
    winner really covers its pixel."""
    n, W, H = 1_000_000, 1024, 1024
    cfg = PRESETS["traj_ball"].for_trajectory(100)
    x = synthetic.trajectory(1, n, 3, seed=0)[0]
    style = cfg.style()
    # 1 M points: the default (AUTO) is the reference's own sequential float32 mean, bit-identical to numpy; the
    # parallel f64 mean is the order-independent definition, within the REFERENCE's summation error of it
    pos4, attr4 = ctx.standardize(dev(x), style)
    seq = pos4.cpu().numpy()
    np.testing.assert_array_equal(seq[:, :3], orc.transform_coordinates(orc.standardize_point_cloud(x), cfg.flip_x))
    f64 = ctx.standardize(dev(x), cfg.style(mean_mode=_native.MEAN_F64))[0].cpu().numpy()
    np.testing.assert_array_equal(f64[:, :3], orc.transform_coordinates(orc.standardize_point_cloud(x, exact_mean=True), cfg.flip_x))
    np.testing.assert_allclose(f64[:, :3], seq[:, :3], rtol=0, atol=ref_atol(x))
    for frame_index in (0, 99):
        cam = cfg.camera(frame_index, 100, W, H)
        vis, rgba = ctx.render(pos4, attr4, cam, style)
        vis2, rgba2 = ctx.render(pos4, attr4, cam, style)
        assert torch.equal(vis, vis2) and torch.equal(rgba, rgba2)
        want = orc.visibility(pos4.cpu().numpy(), orc_frame(orc, cfg, frame_index, 100, W, H), orc_scene(orc, cfg))
        np.testing.assert_array_equal(keys(vis), want)
        ids = _native.keys_to_ids(vis)
        assert np.all((ids < n) | (ids >= 0xFFFFFFFE))
        assert ctx.counters()["overflow_frames"] == 0


def test_fused_headline_batches_match_oracle_on_reference_centres(lib, orc):
    """The configuration the headline number is quoted on, through the entry bench.py times: pcr_render_frames, 1 M
    points, 1024^2, 32 frames per launch, default style (mean_mode AUTO = the reference's sequential float32 mean).
    40 frames = one full batch + a ragged one, the second prepared by the look-ahead on a side stream; the same frames
    again after pcr_prefetch_frames, and through the host-buffer entry.  Keys bit-exact against the oracle fed the
    centres numpy computes (orc.standardize_point_cloud is the reference verbatim — NOT the device's own stats)."""
    n, W, H, F, B = 1_000_000, 1024, 1024, 40, 32
    cfg = PRESETS["traj_ball"].for_trajectory(100)
    traj = synthetic.trajectory(F, n, 3, seed=4)
    cam_idx = [(37 * f) % 100 for f in range(F)]                       # far, middle and nearest cameras of the schedule
    cams = [cfg.camera(i, 100, W, H) for i in cam_idx]
    style = cfg.style()
    assert style.mean_mode == _native.MEAN_AUTO
    c = _native.Context(device=0, max_points=n, max_w=W, max_h=H, max_batch=B)
    try:
        d = dev(traj)
        rgba, vis = c.render_frames(d, cams, style, want_vis=True)
        assert c.counters()["overflow_frames"] == 0
        sc = orc_scene(orc, cfg)
        for f in (0, 17, 31, 32, 39):
            p = orc.transform_coordinates(orc.standardize_point_cloud(traj[f]), cfg.flip_x)      # numpy, reference arithmetic
            pos4 = np.concatenate([p, np.full((n, 1), cfg.radius, np.float32)], axis=1)
            want = orc.visibility(pos4, orc_frame(orc, cfg, cam_idx[f], 100, W, H), sc)
            bad = np.argwhere(keys(vis[f]) != want)
            assert len(bad) == 0, f"frame {f}: {len(bad)} pixels differ, first {bad[:3].tolist()}"
            if f in (0, 39):
                check_image(rgba[f].cpu().numpy(), orc.shade(want, pos4, orc.compute_color(p, mode=0), orc_frame(orc, cfg, cam_idx[f], 100, W, H), sc))
        # hints change when K0 runs, never what it computes
        c.prefetch_frames(d[:B], style)
        c.prefetch_frames(d[B:], style)
        rgba2, vis2 = c.render_frames(d[:B], cams[:B], style, want_vis=True)
        rgba3, vis3 = c.render_frames(d[B:], cams[B:], style, want_vis=True)
        assert torch.equal(vis2, vis[:B]) and torch.equal(rgba2, rgba[:B]) and torch.equal(vis3, vis[B:]) and torch.equal(rgba3, rgba[B:])
        c.prefetch_frames(d[:B], style)                                  # an unconsumed hint, then dropped
        c.prefetch_frames(d[:0], style)
        host_vis = torch.empty((F, H, W), dtype=torch.int64).pin_memory()
        host_rgba = c.render_frames_host(torch.from_numpy(traj).pin_memory(), cams, style, out_vis=host_vis)
        assert torch.equal(host_vis, vis.cpu()) and torch.equal(host_rgba, rgba.cpu())
        # the f64 mean is a different (more accurate) centre: some keys legitimately differ at this size — which is
        # why the default must be the reference's arithmetic
        vis64 = c.render_frames(d[:1], cams[:1], cfg.style(mean_mode=_native.MEAN_F64), want_vis=True)[1]
        p64 = orc.transform_coordinates(orc.standardize_point_cloud(traj[0], exact_mean=True), cfg.flip_x)
        want64 = orc.visibility(np.concatenate([p64, np.full((n, 1), cfg.radius, np.float32)], axis=1), orc_frame(orc, cfg, cam_idx[0], 100, W, H), sc)
        np.testing.assert_array_equal(keys(vis64[0]), want64)
    finally:
        c.close()


def test_c5_film_ten_million_points_through_render_shard(lib, orc):
    """C5's film (4096^2) at 10 M points through the point-sharded entries on one device: two shards of 5 M points,
    pcr_stats_partial -> pcr_finalize_stats -> pcr_render_shard -> pcr_zmin merge, keys bit-exact against the oracle on
    the whole cloud (standardised with the shard totals folded in rank order: the f64 mean of the point-sharded path);
    the merged buffer also goes through pcr_zmerge_nccl on a one-rank communicator (the NCCL call path of C1)."""
    from pointcloud_render_b200 import sharding
    n, W, H, world = 10_000_000, 4096, 4096, 2
    cfg = PRESETS["example"]
    x = synthetic.cloud(n, "gauss", 31)
    style, cam = cfg.style(), cfg.camera(0, 1, W, H)
    shards = [sharding.point_shard(n, r, world) for r in range(world)]
    c = _native.Context(device=0, max_points=shards[0][1] - shards[0][0], max_w=W, max_h=H, max_batch=1)
    try:
        d = dev(x)
        parts = torch.stack([c.stats_partial(d[a:b]) for a, b in shards]).contiguous()
        stats = c.finalize_stats(parts, n)
        merged = None
        for a, b in shards:
            v = c.render_shard(d[a:b], stats, cam, style, id_base=a).clone()
            merged = v if merged is None else c.zmin_(merged, v)
        assert c.counters()["overflow_frames"] == 0
        p = orc.transform_coordinates(orc.standardize_point_cloud(x, exact_mean=True), True)
        pos4 = np.concatenate([p, np.full((n, 1), cfg.radius, np.float32)], axis=1)
        want = orc.visibility(pos4, orc_frame(orc, cfg, 0, 1, W, H), orc_scene(orc, cfg))
        bad = np.argwhere(keys(merged) != want)
        assert len(bad) == 0, f"{len(bad)} pixels differ, first {bad[:3].tolist()}"
        comm = sharding.NcclComm(0, 1)
        try:
            again = merged.clone()
            c.zmerge_nccl_(again, comm)
            torch.cuda.synchronize()
            assert torch.equal(again, merged)
        finally:
            comm.close()
    finally:
        c.close()


@pytest.mark.parametrize("step", [2, 16, 64])
def test_occlusion_prepass_never_changes_a_key(lib, orc, step):
    """The pre-pass only skips work: keys and images are identical with it on, off, and for any
    subsampling step — also when the pre-pass or the main pass overflows pair_capacity."""
    n, W, H = 400_000, 1024, 768
    cfg = PRESETS["traj_ball"]
    p = orc.transform_coordinates(orc.standardize_point_cloud(synthetic.cloud(n, "gauss", 9)), True)
    pos4 = dev(np.concatenate([p, np.full((n, 1), 0.01, np.float32)], axis=1))
    attr4 = dev(np.full((n, 4), 0.3, np.float32))
    cam, style = cfg.camera(120, 220, W, H), cfg.style()
    results = []
    for mode, cap in ((0, 0), (1, 0), (1, 60_000)):
        c = _native.Context(device=0, max_points=n, max_w=W, max_h=H, max_batch=1, pair_capacity=cap)
        try:
            c.set_occlusion(mode=mode, step=step)
            vis, rgba = c.render(pos4, attr4, cam, style)
            results.append((vis.clone(), rgba.clone(), c.counters()))
        finally:
            c.close()
    assert torch.equal(results[0][0], results[1][0]) and torch.equal(results[0][1], results[1][1])
    assert torch.equal(results[0][0], results[2][0]) and torch.equal(results[0][1], results[2][1])
    # it did skip work (a very sparse pre-pass, thinned out further behind the centre plane, skips less)
    assert results[1][2]["pairs_last_frame"] < (0.5 if step <= 16 else 0.8) * results[0][2]["pairs_last_frame"]
    want = orc.visibility(pos4.cpu().numpy(), orc_frame(orc, cfg, 120, 220, W, H), orc_scene(orc, cfg))
    np.testing.assert_array_equal(keys(results[1][0]), want)


def test_facade_end_to_end(tmp_path, orc):
    """The reference-facing classes: process() on a .npy file writes the PNG the reference names."""
    from PIL import Image
    from pointcloud_render_b200 import renderers
    x = synthetic.cloud(2048, "gauss", 0)
    np.save(tmp_path / "pts_0.npy", x)
    r = renderers.PointCloudRenderer(str(tmp_path / "pts_0.npy"), output_folder=str(tmp_path / "render"), width=800, height=600)
    assert r.init_mitsuba_variant() is True
    r.process()
    img = np.asarray(Image.open(tmp_path / "render" / "pts_0.png"))
    cfg = PRESETS["example"]
    std = r.standardize_point_cloud(x)
    np.testing.assert_array_equal(std, orc.standardize_point_cloud(x))
    p = r.transform_coordinates(std)
    np.testing.assert_array_equal(p, orc.transform_coordinates(std, True))
    pos4 = np.concatenate([p, np.full((2048, 1), 0.01, np.float32)], axis=1)
    fr, sc = orc_frame(orc, cfg, 0, 1, 800, 600), orc_scene(orc, cfg)
    want_vis = orc.visibility(pos4, fr, sc)
    scene = r.render_scene(p)
    np.testing.assert_array_equal(keys(scene.vis), want_vis)
    check_image(img, orc.shade(want_vis, pos4, np.full((2048, 4), 0.3, np.float32), fr, sc)[..., :3])
    # trajectory facade: frame > 199 is written as frame_XXXX_b0 by every subclass (traj_ball_renderer.py:376)
    np.save(tmp_path / "frame_0199_b1.npy", synthetic.trajectory(1, 500, 6)[0])
    rb = renderers.TrajB1Renderer(str(tmp_path / "frame_0199_b1.npy"), output_folder=str(tmp_path / "render"), width=320, height=180)
    rb.process(frame_index=205, total_frames=220)
    assert (tmp_path / "render" / "frame_0205_b0.png").exists()
    rgba = rb.render_trajectory(synthetic.trajectory(3, 500, 6), total_frames=3)
    assert tuple(rgba.shape) == (3, 180, 320, 4)
    renderers.release_engines()


def test_large_film_uses_global_atomic_binning(lib, orc):
    """4096 x 4096 (65 536 tiles — too many for the shared-memory tile histogram, so K2 takes the
    per-pair global-atomic path) with large projected spheres, pre-pass on and off."""
    n, W, H = 300_000, 4096, 4096
    cfg = PRESETS["example"]
    p = orc.transform_coordinates(orc.standardize_point_cloud(synthetic.cloud(n, "gauss", 21)), True)
    pos4 = np.concatenate([p, np.full((n, 1), 0.01, np.float32)], axis=1)
    want = orc.visibility(pos4, orc_frame(orc, cfg, 0, 1, W, H), orc_scene(orc, cfg))
    c = _native.Context(device=0, max_points=n, max_w=W, max_h=H, max_batch=1)
    try:
        for mode in (0, 1):
            c.set_occlusion(mode=mode)
            vis, _ = c.render(dev(pos4), dev(np.full((n, 4), 0.3, np.float32)), cfg.camera(0, 1, W, H), cfg.style(), shade=False)
            np.testing.assert_array_equal(keys(vis), want)
            assert c.counters()["overflow_frames"] == 0
    finally:
        c.close()


def test_trajectory_to_files_output_stage(tmp_path, orc):
    """§8f-3: render_trajectory_to_files = host-buffer render + thread-pool PNG encode; the files
    decode to exactly the frames the device path produces, with the reference's naming rule."""
    from PIL import Image
    from pointcloud_render_b200 import renderers
    F, n = 7, 3000
    traj = synthetic.trajectory(F, n, 6, seed=8)
    r = renderers.TrajB1Renderer(None, output_folder=str(tmp_path / "render"), width=320, height=200)
    paths = r.render_trajectory_to_files(traj, first_frame=196, total_frames=220, chunk=3)
    want = r.render_trajectory(torch.from_numpy(traj).cuda(), first_frame=196, total_frames=220).cpu().numpy()
    names = [os.path.basename(p) for p in paths]
    assert names == ["frame_0196.png", "frame_0197.png", "frame_0198.png", "frame_0199.png",
                     "frame_0200_b0.png", "frame_0201_b0.png", "frame_0202_b0.png"]
    for k, p in enumerate(paths):
        np.testing.assert_array_equal(np.asarray(Image.open(p)), want[k][..., :3])
    renderers.release_engines()


# ------------------------------------------------------------------------------- 8f-1 velocity trails
@pytest.mark.parametrize("key,preset", [("traj_ball_7", "traj_ball"), ("traj_vel_211", "traj_vel"), ("traj_b0_4", "traj_b0"),
                                        ("traj_ball_150", "traj_ball")])
def test_trail_geometry_matches_reference_curve_files(ctx, orc, golden, key, preset):
    """pcr_velocity_trails against the control points the reference itself wrote to its curve files
    (tests/golden/trails.npz): tail, head and 'draws a trail' bit for bit."""
    g = golden("trails.npz")
    frame = int(key.rsplit("_", 1)[1])
    cfg = PRESETS[preset]
    assert cfg.trail_length_scale(frame) == orc.trail_length_scale(preset, frame)
    tail, head, valid = ctx.velocity_trails(dev(g[f"pcl_{key}"]), cfg.style(trails=True), cfg.trail_length_scale(frame))
    v = g[f"valid_{key}"]
    np.testing.assert_array_equal(valid.cpu().numpy().astype(bool), v)
    np.testing.assert_array_equal(tail.cpu().numpy()[v], g[f"tail_{key}"][v])
    np.testing.assert_array_equal(head.cpu().numpy()[v], g[f"head_{key}"][v])
    # and a larger random set against the oracle (itself pinned to the reference)
    rng = np.random.default_rng(5)
    pcl = (rng.standard_normal((50_000, 6)) * [0.2, 0.2, 0.2, 5, 5, 5]).astype(np.float32)
    for scale in (1.0, 7 / 19.0):
        t2, h2, v2 = ctx.velocity_trails(dev(pcl), cfg.style(trails=True), scale)
        wt, wh, wv = orc.velocity_trails(pcl, scale)
        np.testing.assert_array_equal(v2.cpu().numpy().astype(bool), wv)
        np.testing.assert_array_equal(t2.cpu().numpy()[wv], wt[wv])
        np.testing.assert_array_equal(h2.cpu().numpy()[wv], wh[wv])


@pytest.mark.parametrize("preset,frame_index,W,H,n,radius", [
    ("traj_ball", 150, 1920, 1080, 4000, None),        # the reference's 0.0007 radius at its native film size
    ("traj_ball", 7, 1024, 1024, 20_000, None),        # ramp-in (7/19 of the full length)
    ("traj_vel", 211, 1024, 768, 3000, 0.004),         # fade-out; thicker trails cover many pixels
    ("traj_b0", 60, 1024, 1024, 16384, 0.002),         # C3-like, per-point sphere radius
    ("traj_original", 199, 333, 211, 2000, 0.01),      # ragged film, fat capsules
])
def test_trails_visibility_and_image_match_oracle(ctx, orc, preset, frame_index, W, H, n, radius):
    """Whole path with velocity trails: keys (spheres, trails = ids n+i, floor) bit-exact against the
    oracle's sphere + capsule caster, image within one code value."""
    import dataclasses
    cfg = PRESETS[preset]
    if radius is not None:
        cfg = dataclasses.replace(cfg, trail_radius=radius)
    traj = synthetic.trajectory(2, n, 6, seed=frame_index)
    rad = synthetic.radii(n) if preset == "traj_b0" else None
    cams = [cfg.camera(frame_index + k, 220, W, H) for k in range(2)]
    style = cfg.style(trails=True)
    rgba, vis = ctx.render_frames(dev(traj), cams, style, radius=None if rad is None else dev(rad), want_vis=True)
    sc = orc_scene(orc, cfg)
    for k in range(2):
        pcl = orc.transform_coordinates(orc.standardize_point_cloud(traj[k]), cfg.flip_x)
        r = rad if rad is not None else np.full(n, cfg.radius, np.float32)
        pos4 = np.concatenate([pcl[:, :3], r[:, None]], axis=1)
        fr = orc_frame(orc, cfg, frame_index + k, 220, W, H)
        tail, head, valid = orc.velocity_trails(pcl, cfg.trail_length_scale(frame_index + k))
        want = orc.add_trails(orc.visibility(pos4, fr, sc), tail, head, valid, fr, n, radius=cfg.trail_radius)
        got = keys(vis[k])
        bad = np.argwhere(got != want)
        assert len(bad) == 0, f"frame {k}: {len(bad)} pixels differ, first {bad[:3].tolist()}"
        ids = (want & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        assert np.any((ids >= n) & (ids < 2 * n)), "the scene should show some trail pixels"
        attr4 = orc.compute_color(pcl, mode=0)
        img = orc.shade_trails(orc.shade(want, pos4, attr4, fr, sc), want, tail, head, valid, fr, sc, n, radius=cfg.trail_radius, rgb=cfg.trail_rgb)
        check_image(rgba[k].cpu().numpy(), img)
    # without the flag (or with 3-column frames) nothing changes: no id >= n appears
    rgba0, vis0 = ctx.render_frames(dev(traj), cams, cfg.style(), radius=None if rad is None else dev(rad), want_vis=True)
    ids0 = _native.keys_to_ids(vis0)
    assert not np.any((ids0 >= n) & (ids0 < 0xFFFFFFFE))


def test_trails_with_occlusion_prepass_and_overflow(lib, orc):
    """Trails through the Hi-Z pre-pass (they are culled by it but are never occluders) and through
    the pair_capacity overflow path."""
    import dataclasses
    n, W, H = 150_000, 1024, 768
    cfg = dataclasses.replace(PRESETS["traj_ball"], trail_radius=0.002)
    traj = synthetic.trajectory(1, n, 6, seed=3)
    cams, style = [cfg.camera(120, 220, W, H)], cfg.style(trails=True, mean_mode=_native.MEAN_F64)
    results = []
    for mode, cap in ((0, 0), (1, 0), (1, 50_000)):
        c = _native.Context(device=0, max_points=n, max_w=W, max_h=H, max_batch=1, pair_capacity=cap)
        try:
            c.set_occlusion(mode=mode)
            rgba, vis = c.render_frames(dev(traj), cams, style, want_vis=True)
            results.append((vis.clone(), rgba.clone(), c.counters()["overflow_frames"]))
        finally:
            c.close()
    assert results[2][2] == 1 and results[0][2] == 0
    for r in results[1:]:
        assert torch.equal(results[0][0], r[0]) and torch.equal(results[0][1], r[1])
    pcl = orc.transform_coordinates(orc.standardize_point_cloud(traj[0], exact_mean=True), True)
    pos4 = np.concatenate([pcl[:, :3], np.full((n, 1), 0.01, np.float32)], axis=1)
    fr = orc_frame(orc, cfg, 120, 220, W, H)
    tail, head, valid = orc.velocity_trails(pcl, 1.0)
    want = orc.add_trails(orc.visibility(pos4, fr, orc_scene(orc, cfg)), tail, head, valid, fr, n, radius=0.002)
    np.testing.assert_array_equal(keys(results[0][0][0]), want)


def test_float64_frames_with_trails_device_and_host_entries(ctx, orc):
    """float64 input through the fused path (K1 in f64, the reference's f64 sequential mean), with
    trails, through both the device-buffer and the host-buffer entry; 9 frames over batches of 4
    exercise the stats-ahead stream."""
    import dataclasses
    F, n, W, H = 9, 5000, 640, 480
    cfg = dataclasses.replace(PRESETS["traj_vel"], trail_radius=0.003)
    traj = synthetic.trajectory(F, n, 6, seed=12, dtype=np.float64) * 3.0 + 1.5
    cams = [cfg.camera(195 + k, 220, W, H) for k in range(F)]           # crosses the fade-out of the trail length
    style = cfg.style(trails=True, color_mode=2)
    rgba, vis = ctx.render_frames(dev(traj), cams, style, want_vis=True)
    host_vis = torch.empty((F, H, W), dtype=torch.int64).pin_memory()
    host_rgba = ctx.render_frames_host(torch.from_numpy(traj).pin_memory(), cams, style, out_vis=host_vis)
    assert torch.equal(host_vis, vis.cpu()) and torch.equal(host_rgba, rgba.cpu())
    sc = orc_scene(orc, cfg)
    for k in (0, 4, 8):
        pcl = orc.transform_coordinates(orc.standardize_point_cloud(traj[k]), cfg.flip_x)
        pos4 = np.concatenate([pcl[:, :3], np.full((n, 1), cfg.radius, np.float32)], axis=1)
        fr = orc_frame(orc, cfg, 195 + k, 220, W, H)
        tail, head, valid = orc.velocity_trails(pcl, cfg.trail_length_scale(195 + k))
        want = orc.add_trails(orc.visibility(pos4, fr, sc), tail, head, valid, fr, n, radius=cfg.trail_radius)
        np.testing.assert_array_equal(keys(vis[k]), want)
        attr4 = orc.compute_color(pcl, mode=2)
        img = orc.shade_trails(orc.shade(want, pos4, attr4, fr, sc), want, tail, head, valid, fr, sc, n, radius=cfg.trail_radius, rgb=cfg.trail_rgb)
        check_image(rgba[k].cpu().numpy(), img)


def test_process_draws_the_velocity_trails_the_reference_draws(tmp_path, lib, orc):
    """TrajectoryBallRenderer.process() (and the Original/B0/B1 subclasses) must draw what generate_xml_content
    emits for an (N,6) cloud: one sphere AND one velocity trail per point (traj_ball_renderer.py:319-330,
    traj_b0.py:117-191).  process() -> render_scene -> pcr_render_transformed; keys bit-exact against the oracle's
    sphere + capsule caster on the facade's own standardised / transformed array, image within one code value."""
    from PIL import Image
    from pointcloud_render_b200 import renderers
    n, W, H = 3000, 480, 270
    for cls, preset, frame in ((renderers.TrajectoryBallRenderer, "traj_ball", 57), (renderers.TrajB0Renderer, "traj_b0", 120),
                               (renderers.FixedFrame199Renderer, "traj_original", 199)):
        cfg = PRESETS[preset]
        raw = synthetic.trajectory(1, n, 6, seed=frame)[0]
        raw[:, 3:] *= 2.0
        path = tmp_path / f"frame_{frame:04d}_{preset}.npy"
        np.save(path, raw)
        r = cls(str(path), output_folder=str(tmp_path / "render"), width=W, height=H)
        assert r.trails, "the trajectory scripts draw trails by default"
        pcl = r.transform_coordinates(r.standardize_point_cloud(raw))
        np.testing.assert_array_equal(pcl, orc.transform_coordinates(orc.standardize_point_cloud(raw), cfg.flip_x))
        scene = r.render_scene(pcl, frame_index=frame, total_frames=220)
        pos4 = np.concatenate([pcl[:, :3], np.full((n, 1), cfg.radius, np.float32)], axis=1)
        fr, sc = orc_frame(orc, cfg, frame, 220, W, H), orc_scene(orc, cfg)
        tail, head, valid = orc.velocity_trails(pcl, cfg.trail_length_scale(frame))
        want = orc.add_trails(orc.visibility(pos4, fr, sc), tail, head, valid, fr, n, radius=cfg.trail_radius)
        np.testing.assert_array_equal(keys(scene.vis), want)
        ids = scene.point_ids()
        assert np.any((ids >= n) & (ids < 2 * n)), "process() must show trail ids (n + point)"
        img = orc.shade_trails(orc.shade(want, pos4, orc.compute_color(pcl, mode=0), fr, sc), want, tail, head, valid, fr, sc, n,
                               radius=cfg.trail_radius, rgb=cfg.trail_rgb)
        check_image(scene.numpy(), img)
        # the file process() writes is that image
        r.process(frame_index=frame, total_frames=220)
        np.testing.assert_array_equal(np.asarray(Image.open(tmp_path / "render" / f"frame_{frame:04d}_{preset}.png")), scene.numpy()[..., :3])
        # trails=False (or a 3-column cloud) draws spheres only
        off = cls(str(path), width=W, height=H, trails=False).render_scene(pcl, frame_index=frame, total_frames=220)
        np.testing.assert_array_equal(keys(off.vis), orc.visibility(pos4, fr, sc))
        three = r.render_scene(pcl[:, :3].copy(), frame_index=frame, total_frames=220)
        np.testing.assert_array_equal(keys(three.vis), orc.visibility(pos4, fr, sc))
    renderers.release_engines()


def test_wrappers_reject_wrong_dtype_layout_and_device(ctx):
    """The ctypes wrappers hand raw pointers to C: a strided view, a half / int tensor or a CPU tensor must raise,
    not be reinterpreted."""
    cfg = PRESETS["traj_ball"]
    style, cam = cfg.style(), cfg.camera(0, 220, 64, 48)
    six = torch.zeros((100, 6), dtype=torch.float32, device="cuda")
    with pytest.raises(ValueError):
        ctx.standardize(six[:, :3], style)                                  # non-contiguous slice of an (N,6) tensor
    with pytest.raises(TypeError):
        ctx.standardize(six.half(), style)
    with pytest.raises(TypeError):
        ctx.render_frames(torch.zeros((1, 100, 3), dtype=torch.int32, device="cuda"), [cam], style)
    with pytest.raises(ValueError):
        ctx.render_frames(torch.zeros((1, 100, 3), dtype=torch.float32), [cam], style)      # CPU tensor in the device entry
    with pytest.raises(ValueError):
        ctx.render_frames_host(torch.zeros((1, 100, 3), dtype=torch.float32, device="cuda"), [cam], style)
    with pytest.raises(ValueError):
        ctx.render_frames(torch.zeros((1, 100, 3), dtype=torch.float32, device="cuda"), [cam], style,
                          radius=torch.zeros(99, dtype=torch.float32, device="cuda"))
    with pytest.raises(ValueError):
        ctx.stats_partial(torch.zeros((100, 4), dtype=torch.float32, device="cuda"))
    with pytest.raises(ValueError):
        ctx.render_shard(six, torch.zeros(9, dtype=torch.float64, device="cuda"), cam, style)


def test_host_entry_asynchronous_calls_in_flight(ctx, orc):
    """pcr_render_frames_host_submit / pcr_host_wait: three calls submitted back to back (more chunks than staging
    slots, so slots, prepared-stats slots and camera ring are all recycled while earlier chunks are still in flight)
    deliver exactly what the synchronous entry and the device entry deliver."""
    F, n, W, H = 11, 20_000, 400, 304
    cfg = PRESETS["traj_ball"].for_trajectory(100)
    style = cfg.style()
    calls = []
    for c in range(3):
        traj = synthetic.trajectory(F, n, 3, seed=40 + c)
        cams = [cfg.camera((9 * f + c) % 100, 100, W, H) for f in range(F)]
        calls.append((traj, cams, torch.from_numpy(traj).pin_memory(), torch.empty((F, H, W), dtype=torch.int64).pin_memory()))
    tickets = [ctx.render_frames_host_submit(host, cams, style, out_vis=hv) for _, cams, host, hv in calls]
    ctx.host_wait(tickets[1][0])
    ctx.host_wait(-1)
    for (traj, cams, host, hv), (_, rgba) in zip(calls, tickets):
        want_rgba, want_vis = ctx.render_frames(dev(traj), cams, style, want_vis=True)
        assert torch.equal(hv, want_vis.cpu()) and torch.equal(rgba, want_rgba.cpu())
        sync = ctx.render_frames_host(host, cams, style)
        assert torch.equal(sync, rgba)


def test_zero_frames_and_tiny_inputs(ctx):
    cfg = PRESETS["traj_ball"]
    empty = torch.empty((0, 100, 3), dtype=torch.float32, device="cuda")
    out = ctx.render_frames(empty, [], cfg.style())
    assert tuple(out.shape) == (0, 1080, 1920, 4) or out.numel() == 0
    one = torch.zeros((1, 1, 3), dtype=torch.float32, device="cuda")      # a single point: extent 0 -> NaN positions, like numpy
    rgba, vis = ctx.render_frames(one, [cfg.camera(0, 220, 64, 48)], cfg.style(), want_vis=True)
    ids = _native.keys_to_ids(vis)
    assert np.all(ids >= 0xFFFFFFFE)                                      # nothing but floor / miss


def test_hoisted_scale_division_is_ieee_exact(ctx):
    """K1 divides every coordinate of a frame by one scale; the kernels reuse the scale's refined reciprocal
    (scale_div, pcr_kernels.cuh).  Exhaustive check: every binary32 dividend x a spread of divisors (typical cloud
    extents, powers of two, all-ones mantissas, both ends of the fast range and beyond it) must give the bit
    pattern of __fdiv_rn."""
    rng = np.random.default_rng(7)
    divisors = list(np.exp(rng.uniform(np.log(1e-3), np.log(1e4), 40)).astype(np.float32))
    divisors += [1.0, 2.0, 0.5, 3.0, 1.9999999, 1.0000001, 0.99999994, 7.9999995, 2.0 ** -40, 2.0 ** 40, 2.0 ** -41, 2.0 ** 41,
                 1e-30, 1e30, 1.1754944e-38, 3.4028235e38]
    assert ctx.selftest_scale_div(divisors) == 0


@pytest.mark.parametrize("shape", ["cube", "shell", "slab", "two-blobs"])
def test_prepass_cut_and_cull4_on_other_shapes(lib, orc, shape):
    """The occluder pre-pass leaves out most of the points behind the cloud's centre plane, and k_project_cull4's first test
    works on approximate camera-space centres: both are tuned on a Gaussian ball and must stay exact on anything — a uniform
    cube, a hollow shell (everything visible is near the front, the back is seen through nothing), a thin slab facing the camera
    (every point within the cut's margin of the centre plane) and two separate blobs (the centre plane lies in the gap between
    them).  n % 4 == 0 and 3 float columns: the fused entry takes k_project_cull4.  Keys against the oracle fed numpy's
    standardisation; two cameras of the schedule."""
    n, W, H = 240_000, 640, 480
    cfg = PRESETS["traj_ball"].for_trajectory(100)
    rng = np.random.default_rng(5)
    if shape == "slab":
        x = np.column_stack([rng.random(n), 0.02 * rng.random(n), rng.random(n)])        # thin along the axis that becomes depth-ish
    elif shape == "two-blobs":
        x = rng.standard_normal((n, 3)) * 0.2 + np.where(rng.random((n, 1)) < 0.5, -1.0, 1.0) * np.array([0.0, 0.0, 1.0])
    else:
        x = synthetic.cloud(n, shape, seed=5)
    x = np.ascontiguousarray(x, dtype=np.float32)
    traj = np.stack([x, x[::-1].copy()])
    cams = [cfg.camera(10, 100, W, H), cfg.camera(90, 100, W, H)]
    c = _native.Context(device=0, max_points=n, max_w=W, max_h=H, max_batch=2)
    try:
        rgba, vis = c.render_frames(dev(traj), cams, cfg.style(), want_vis=True)
        assert c.counters()["overflow_frames"] == 0
        sc = orc_scene(orc, cfg)
        for f, ci in ((0, 10), (1, 90)):
            p = orc.transform_coordinates(orc.standardize_point_cloud(traj[f]), cfg.flip_x)
            pos4 = np.concatenate([p, np.full((n, 1), cfg.radius, np.float32)], axis=1)
            want = orc.visibility(pos4, orc_frame(orc, cfg, ci, 100, W, H), sc)
            bad = np.argwhere(keys(vis[f]) != want)
            assert len(bad) == 0, f"{shape} frame {f}: {len(bad)} pixels differ, first {bad[:3].tolist()}"
    finally:
        c.close()


@pytest.mark.parametrize("switch", ["PCR_LAZY_FILL", "PCR_SAMPLE_PREPASS", "PCR_TWO_PHASE"])
def test_work_saving_devices_never_change_a_result(lib, orc, switch, monkeypatch):
    """Lazy floor fill, the compact pre-pass sample written by K0 and the two-phase cull of K2a only move or skip work:
    with each of them switched off (environment switch read by pcr_create) the fused whole-path entry returns the same
    keys and the same image, on a dense cloud with the occlusion pre-pass and on a small one without; the dense
    frame's keys are also checked against the oracle."""
    cfg = PRESETS["traj_ball"].for_trajectory(100)
    outs = []
    for off in (False, True):
        if off:
            monkeypatch.setenv(switch, "0")
        c = _native.Context(device=0, max_points=300_000, max_w=800, max_h=608, max_batch=2)
        try:
            res = []
            for n, W, H in ((300_000, 800, 608), (5_000, 640, 360)):
                traj = synthetic.trajectory(3, n, 3, seed=11)
                cams = [cfg.camera(97 + f, 100, W, H) for f in range(3)]
                rgba, vis = c.render_frames(dev(traj), cams, cfg.style(), want_vis=True)
                res.append((rgba.cpu().numpy(), keys(vis)))
            outs.append(res)
        finally:
            c.close()
    for (rgba_on, vis_on), (rgba_off, vis_off) in zip(*outs):
        np.testing.assert_array_equal(vis_on, vis_off)
        np.testing.assert_array_equal(rgba_on, rgba_off)
    # the dense case against the oracle fed the REFERENCE's numpy standardisation
    traj = synthetic.trajectory(3, 300_000, 3, seed=11)
    p = orc.transform_coordinates(orc.standardize_point_cloud(traj[2]), cfg.flip_x)
    pos4 = np.concatenate([p, np.full((len(p), 1), cfg.radius, np.float32)], axis=1)
    want = orc.visibility(pos4, orc_frame(orc, cfg, 99, 100, 800, 608), orc_scene(orc, cfg))
    np.testing.assert_array_equal(outs[0][0][1][2], want)


@pytest.mark.parametrize("n", [511, 512, 513, 4095, 4096, 4097, 9001])
def test_raster_item_and_chunk_boundaries(ctx, orc, n):
    """K3 streams a work item (one tile, at most 4096 pairs) through its shared-memory ring in chunks of 512 pairs
    with 16-byte-granular bulk copies; a tile with more pairs is split into several items that merge with atomicMin.
    n tiny spheres whose pixel boxes all lie inside ONE 16x16 tile put exactly n pairs into that tile: the counts
    straddle every boundary (last chunk of 1 or 511 pairs, exactly one full item, one item + 1 pair, three items).
    The film is sized so that the look-at target projects onto the centre of a tile."""
    cfg = PRESETS["traj_ball"]
    W, H = 1008, 624                                   # centre pixel (504, 312): both = 8 mod 16
    rng = np.random.default_rng(n)
    target = np.array(cfg.target, np.float32)
    pos = target[None, :] + rng.uniform(-0.002, 0.002, (n, 3)).astype(np.float32)
    pos4 = np.concatenate([pos, np.full((n, 1), 0.004, np.float32)], axis=1)
    got = render_case(ctx, orc, pos4, cfg, 199, 220, W, H)
    ids = (got & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    hit = np.argwhere(ids < n)
    assert len(hit) > 10
    assert hit[:, 0].min() // 16 == hit[:, 0].max() // 16 and hit[:, 1].min() // 16 == hit[:, 1].max() // 16     # one tile


def test_visibility_buffer_on_an_8_byte_boundary(lib, orc):
    """The raster prefetches a tile's current keys with 16-byte-granular bulk copies; a caller's visibility buffer
    that is only 8-byte aligned must fall back to plain loads (dense cloud with the pre-pass on, so that the seeded
    main pass and split tiles are exercised)."""
    n, W, H = 200_000, 640, 480
    cfg = PRESETS["traj_ball"]
    p = orc.transform_coordinates(orc.standardize_point_cloud(synthetic.cloud(n, "gauss", 5)), True)
    pos4 = dev(np.concatenate([p, np.full((n, 1), 0.01, np.float32)], axis=1))
    attr4 = dev(np.full((n, 4), 0.3, np.float32))
    cam, style = cfg.camera(150, 220, W, H), cfg.style()
    c = _native.Context(device=0, max_points=n, max_w=W, max_h=H, max_batch=1)
    try:
        c.set_occlusion(mode=1, step=8)
        vis_a, rgba_a = c.render(pos4, attr4, cam, style)
        backing = torch.empty(W * H + 1, dtype=torch.int64, device="cuda")
        odd = backing[1:].view(H, W)
        assert odd.data_ptr() % 16 == 8
        vis_b, rgba_b = c.render(pos4, attr4, cam, style, out_vis=odd)
        assert torch.equal(vis_a, vis_b) and torch.equal(rgba_a, rgba_b)
    finally:
        c.close()
