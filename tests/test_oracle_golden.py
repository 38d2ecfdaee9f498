"""The CPU oracle against the committed outputs of the reference (tests/golden/, written by
oracle/gen_golden.py) and against hand-checkable known answers.  Runs everywhere (no GPU, no
/root/reference)."""
import hashlib

import numpy as np
import pytest

NAMES = {"ball": "traj_ball", "traj": "traj", "vel": "traj_vel", "orig": "traj_original", "b0": "traj_b0", "b1": "traj_b1"}


def test_standardize_and_transform_golden(golden, orc):
    g = golden("standardize.npz")
    for tag in ("f32_3", "f64_3", "f32_6", "f64_6"):
        x = g[f"in_{tag}"]
        std = orc.standardize_point_cloud(x.copy())
        for short, preset in NAMES.items():
            np.testing.assert_array_equal(std, g[f"std_{short}_{tag}"])
            xf = orc.transform_coordinates(std.copy(), flip_x=orc.PRESETS[preset]["flip_x"])
            np.testing.assert_array_equal(xf, g[f"xf_{short}_{tag}"])
        if tag.endswith("_3"):
            np.testing.assert_array_equal(std, g[f"std_example_{tag}"])
            np.testing.assert_array_equal(orc.transform_coordinates(std.copy(), True), g[f"xf_example_{tag}"])


def test_camera_golden(golden, orc):
    g = golden("camera.npz")
    for preset in ("traj", "traj_ball", "traj_vel", "traj_original", "traj_b0", "traj_b1"):
        got = np.array([orc.camera_position(preset, int(f), 220) for f in g["frames"]])
        np.testing.assert_array_equal(got, g[preset])
    # the key-frames the reference's code (not its comments) uses: traj_ball_renderer.py:292-301
    assert orc.camera_position("traj_ball", 0) == (2.8, 2.8, 3.0)
    assert orc.camera_position("traj_ball", 199) == (1.8, 1.8, 1.8)
    np.testing.assert_allclose(orc.camera_position("traj_ball", 219), (1.6, 1.6, 1.6), atol=1e-12)
    np.testing.assert_allclose(orc.camera_position("traj", 219, 220), (0.8, 0.8, 1.0), atol=1e-12)


@pytest.mark.parametrize("name", ["example", "traj_ball", "traj_original", "traj_b0", "traj_b1"])
def test_scene_golden(golden, orc, name):
    g = golden(f"scene_{name}.npz")
    pr = orc.PRESETS[name]
    p = orc.transform_coordinates(orc.standardize_point_cloud(g["input"]), flip_x=pr["flip_x"])
    np.testing.assert_array_equal(p, g["centers"])
    assert np.all(g["radius"] == np.float32(0.01)) and np.all(g["reflectance"] == np.float32(0.3))
    assert tuple(g["origin"]) == orc.camera_position(name, int(g["frame"]), 220)
    assert tuple(g["target"]) == pr["target"] and float(g["fov"]) == pr["fov"]
    assert float(g["floor_z"]) == pr["floor_z"]
    assert tuple(g["floor_min"]) == pr["floor_min"] and tuple(g["floor_max"]) == pr["floor_max"]
    assert (float(g["near_clip"]), float(g["far_clip"])) == (0.1, 100.0)
    assert (int(g["width"]), int(g["height"])) == (1920, 1080)


def _example_scene(golden, orc, W, H):
    g = golden("scene_example.npz")
    pos4 = np.concatenate([g["centers"], g["radius"][:, None]], axis=1).astype(np.float32)
    scene = orc.make_scene(True, float(g["floor_z"]), tuple(g["floor_min"]), tuple(g["floor_max"]), 1.0,
                           float(g["light_z"]), float(g["light_half"]), float(g["radiance"]), 1.0)
    frame = orc.camera_frame(g["origin"], g["target"], g["up"], float(g["fov"]), 0.1, 100.0, W, H)
    return g, pos4, scene, frame


def test_visibility_self_pin(golden, orc):
    """Oracle output has not drifted (unpinned by the reference — Mitsuba absent)."""
    v = golden("vis_example.npz")
    g, pos4, scene, frame = _example_scene(golden, orc, 200, 150)
    vis = orc.visibility(pos4, frame, scene, brute_force=True)
    np.testing.assert_array_equal(vis, v["keys_200x150"])
    np.testing.assert_array_equal(orc.visibility(pos4, frame, scene), vis)        # bbox mode == definition
    attr4 = np.concatenate([g["reflectance"], np.zeros((len(pos4), 1), np.float32)], axis=1)
    img = orc.shade(vis, pos4, attr4, frame, scene)
    assert np.abs(img.astype(int) - v["rgba_200x150"].astype(int)).max() <= 1      # libm pow/acos may differ by an ulp
    _, _, _, full = _example_scene(golden, orc, 800, 600)
    assert hashlib.sha256(orc.visibility(pos4, full, scene).tobytes()).hexdigest() == str(v["sha256_800x600"])


@pytest.mark.parametrize("preset,frame_idx,W,H", [("example", 0, 160, 120), ("traj_ball", 150, 128, 128),
                                                  ("traj_b0", 210, 131, 77), ("traj", 219, 96, 96)])
def test_visibility_against_f64_formulation(orc, preset, frame_idx, W, H):
    """The f32 VA-1 arithmetic agrees with an independent textbook f64 ray-sphere caster
    everywhere except within rounding distance of a silhouette or of a depth tie."""
    rng = np.random.default_rng(4)
    pr = orc.PRESETS[preset]
    p = orc.transform_coordinates(orc.standardize_point_cloud(rng.standard_normal((400, 3)).astype(np.float32)), pr["flip_x"])
    pos4 = np.concatenate([p, rng.uniform(0.005, 0.03, (400, 1)).astype(np.float32)], axis=1)
    eye = orc.camera_position(preset, frame_idx, 220)
    scene = orc.make_scene(True, pr["floor_z"], pr["floor_min"], pr["floor_max"])
    frame = orc.camera_frame(eye, pr["target"], (0, 0, 1), pr["fov"], 0.1, 100.0, W, H)
    vis = orc.visibility(pos4, frame, scene, brute_force=True)
    ids = (vis & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    depth = (vis >> np.uint64(32)).astype(np.uint32).view(np.float32)
    ids64, depth64, margin, gap = orc.visibility_f64(pos4, np.float32(eye), np.float32(pr["target"]), (0, 0, 1), pr["fov"],
                                                     0.1, 100.0, W, H, scene)
    safe = (margin > 1e-3) & (gap > 1e-5)
    assert safe.mean() > 0.9
    np.testing.assert_array_equal(ids[safe], ids64[safe])
    hit = safe & (ids64 != orc.ID_MISS)
    np.testing.assert_allclose(depth[hit], depth64[hit], rtol=2e-5)


def test_single_sphere_known_answer(orc):
    """One sphere at the origin, `example` camera: it projects to the image centre with pixel
    radius r * (W/2) / tan(15 deg) / |eye|."""
    W, H = 801, 601
    eye = (2.2, 2.2, 4.2)
    frame = orc.camera_frame(eye, (0, 0, 0), (0, 0, 1), 30.0, 0.1, 100.0, W, H)
    scene = orc.make_scene(has_floor=False)
    r = 0.05
    vis = orc.visibility(np.array([[0, 0, 0, r]], np.float32), frame, scene, brute_force=True)
    ids = (vis & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    ys, xs = np.nonzero(ids == 0)
    assert ids[H // 2, W // 2] == 0 and np.all(ids[ids != 0] == orc.ID_MISS)
    assert abs(xs.mean() - W // 2) < 0.01 and abs(ys.mean() - H // 2) < 0.01
    dist = np.linalg.norm(eye)
    r_px = r * (W / 2) / np.tan(np.deg2rad(15.0)) / np.sqrt(dist * dist - r * r)
    assert abs(np.sqrt(len(xs) / np.pi) - r_px) < 0.15
    d = (vis[H // 2, W // 2] >> np.uint64(32)).astype(np.uint32).view(np.float32)
    assert abs(float(d) - (dist - r)) < 1e-5


def test_ties_keep_the_lower_id(orc):
    frame = orc.camera_frame((0, -3, 0.5), (0, 0, 0.5), (0, 0, 1), 40.0, 0.1, 100.0, 64, 64)
    pos4 = np.array([[0, 0, 0.5, 0.2]] * 3, np.float32)
    vis = orc.visibility(pos4, frame, orc.make_scene(has_floor=False), id_base=10)
    ids = (vis & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    assert set(np.unique(ids)) == {10, orc.ID_MISS}


def test_floor_occludes_and_is_finite(orc):
    """example floor z=-0.2 cuts the lower part of the cloud; rays that leave
    the floor rectangle are MISS."""
    pr = orc.PRESETS["example"]
    frame = orc.camera_frame((2.2, 2.2, 4.2), (0, 0, 0), (0, 0, 1), 30.0, 0.1, 100.0, 64, 48)
    scene = orc.make_scene(True, pr["floor_z"], pr["floor_min"], pr["floor_max"])
    pos4 = np.array([[0, 0, -0.4, 0.1], [0, 0, 0.3, 0.1]], np.float32)
    ids = (orc.visibility(pos4, frame, scene, brute_force=True) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    assert 1 in ids and 0 not in ids and np.all((ids == 1) | (ids == orc.ID_FLOOR))
    tiny = orc.make_scene(True, -0.2, (-0.5, -0.5), (0.5, 0.5))
    ids = (orc.visibility(pos4[:0], frame, tiny, brute_force=True) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    assert orc.ID_MISS in ids and orc.ID_FLOOR in ids


def test_emitter_form_factor_known_value(orc):
    """F = 0.26354 for the 16x16 emitter 15 above an up-facing point (SURVEY.md §8a K4 note):
    an up-facing floor pixel under the light centre shades to radiance * F, clamped to white."""
    frame = orc.camera_frame((0, -1e-3, 5), (0, 0, 0), (0, 0, 1), 20.0, 0.1, 100.0, 5, 5)
    for radiance, expect in ((1.0, 0.26354), (0.5, 0.13177)):
        scene = orc.make_scene(True, 0.0, (-10, -10), (10, 10), 1.0, 15.0, 8.0, radiance, 1.0)
        vis = orc.visibility(np.zeros((0, 4), np.float32), frame, scene)
        img = orc.shade(vis, np.zeros((0, 4), np.float32), np.zeros((0, 4), np.float32), frame, scene)
        lin = expect
        srgb = 1.055 * lin ** (1 / 2.4) - 0.055
        assert abs(int(img[2, 2, 0]) - round(srgb * 255)) <= 1 and img[2, 2, 3] == 255


@pytest.mark.parametrize("key,preset", [("traj_ball_7", "traj_ball"), ("traj_vel_211", "traj_vel"), ("traj_b0_4", "traj_b0"),
                                        ("traj_ball_150", "traj_ball")])
def test_velocity_trails_golden(golden, orc, key, preset):
    """Trail end points against the reference's own curve files (tests/golden/trails.npz)."""
    g = golden("trails.npz")
    frame = int(key.rsplit("_", 1)[1])
    pcl = orc.transform_coordinates(orc.standardize_point_cloud(g[f"raw_{key}"]), orc.PRESETS[preset]["flip_x"])
    np.testing.assert_array_equal(pcl, g[f"pcl_{key}"])
    for exact in (True, False):
        tail, head, valid = orc.velocity_trails(pcl, orc.trail_length_scale(preset, frame), exact_text=exact)
        v = g[f"valid_{key}"]
        np.testing.assert_array_equal(valid, v)
        np.testing.assert_array_equal(tail[v], g[f"tail_{key}"][v])
        np.testing.assert_array_equal(head[v], g[f"head_{key}"][v])


def test_trail_length_scale_schedules(orc):
    assert [orc.trail_length_scale("traj_ball", f) for f in (0, 19, 20, 199, 219)] == [0.0, 1.0, 1.0, 1.0, 1.0]
    assert orc.trail_length_scale("traj_vel", 209) == 0.5 and orc.trail_length_scale("traj_vel", 219) == 0.0
    assert orc.trail_length_scale("traj_vel", 10) == 10 / 19.0
    assert all(orc.trail_length_scale(p, f) == 1.0 for p in ("traj_original", "traj_b0", "traj_b1") for f in (0, 5, 205))


def test_capsule_visibility_modes_agree_and_known_answers(orc):
    """The capsule (trail) caster: bbox mode == brute force; a thick capsule across the view has the
    silhouette of a stadium of the right width; thin reference-size trails hit pixels along the line."""
    W, H = 160, 120
    frame = orc.camera_frame((0, -4, 0), (0, 0, 0), (0, 0, 1), 40.0, 0.1, 100.0, W, H)
    scene = orc.make_scene(has_floor=False)
    base = orc.visibility(np.zeros((0, 4), np.float32), frame, scene)
    rng = np.random.default_rng(0)
    tail = rng.uniform(-1, 1, (40, 3)).astype(np.float32)
    head = (tail + rng.uniform(-0.4, 0.4, (40, 3))).astype(np.float32)
    valid = np.ones(40, bool)
    valid[3] = False
    a = orc.add_trails(base, tail, head, valid, frame, 100, radius=0.02, brute_force=True)
    b = orc.add_trails(base, tail, head, valid, frame, 100, radius=0.02)
    np.testing.assert_array_equal(a, b)
    ids = (a & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    assert 103 not in ids and 100 in ids
    # horizontal capsule of radius 0.25 and half-length 1 in the plane y = 0 (depth 4): height 2r, width 2(1+r)
    one = orc.add_trails(base, np.float32([[-1, 0, 0]]), np.float32([[1, 0, 0]]), [True], frame, 7, radius=0.25, brute_force=True)
    m = (one & np.uint64(0xFFFFFFFF)) == 7
    px_per_unit = (W / 2) / np.tan(np.deg2rad(20.0)) / 4.0
    rows, cols = np.nonzero(m.any(axis=1))[0], np.nonzero(m.any(axis=0))[0]
    assert abs((rows.max() - rows.min() + 1) - 2 * 0.25 * px_per_unit) <= 2.5
    assert abs((cols.max() - cols.min() + 1) - 2 * 1.25 * px_per_unit) <= 3.5
    d = (one[H // 2, W // 2] >> np.uint64(32)).astype(np.uint32).view(np.float32)
    assert abs(float(d) - 3.75) < 1e-3          # even W,H: the centre pixel ray is half a pixel off the axis
    # an end-on capsule (ray parallel to the axis) is seen through its end sphere
    end_on = orc.add_trails(base, np.float32([[0, 0, 0]]), np.float32([[0, 2, 0]]), [True], frame, 9, radius=0.3, brute_force=True)
    assert (end_on[H // 2, W // 2] & np.uint64(0xFFFFFFFF)) == 9
    # reference-size trails (r = 0.0007) are sub-pixel: at 1200 x 900 the 0.58-pixel-wide line is hit by
    # roughly every second pixel-centre ray along it, and every hit lies on the projected segment
    W2, H2 = 1200, 900
    big = orc.camera_frame((0, -4, 0), (0, 0, 0), (0, 0, 1), 40.0, 0.1, 100.0, W2, H2)
    base2 = orc.visibility(np.zeros((0, 4), np.float32), big, scene)
    thin = orc.add_trails(base2, np.float32([[-1, 0, -0.5]]), np.float32([[1, 0, 0.5]]), [True], big, 11)
    ys, xs = np.nonzero((thin & np.uint64(0xFFFFFFFF)) == 11)
    ppu = (W2 / 2) / np.tan(np.deg2rad(20.0)) / 4.0
    assert 0.3 * 2 * ppu < len(xs) <= 2 * ppu + 2
    x_w = (xs + 0.5 - W2 / 2) / ppu
    z_w = (H2 / 2 - 0.5 - ys) / ppu
    assert np.abs(np.abs(z_w) - 0.5 * np.abs(x_w)).max() < 1.0 / ppu
