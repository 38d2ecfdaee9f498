"""Pins oracle/pcr_oracle.py against the UNMODIFIED reference (imported from /root/reference
with mitsuba/plyfile stubbed).  Runs only where the reference tree exists; the same
comparisons against committed outputs of the reference are in test_oracle_golden.py."""
import numpy as np
import pytest

from oracle import scene_from_xml


def _rand(n, cols, seed, dtype):
    rng = np.random.default_rng(seed)
    return np.ascontiguousarray(rng.standard_normal((n, cols)) * 2.5 + 1.0, dtype=dtype)


CLASSES = {
    "traj_ball": ("traj_ball_renderer", "TrajectoryBallRenderer"),
    "traj": ("traj_renderer", "TrajectoryRenderer"),
    "traj_vel": ("traj_vel_renderer", "TrajectoryVelRenderer"),
    "traj_original": ("traj_original", "FixedFrame199Renderer"),
    "traj_b0": ("traj_b0", "FixedFrame199Renderer"),
    "traj_b1": ("traj_b1", "FixedFrame199Renderer"),
}


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("cols", [3, 6])
@pytest.mark.parametrize("n", [1, 2, 1000, 65537])
def test_standardize_matches_reference(reference, orc, dtype, cols, n):
    x = _rand(n, cols, n + cols, dtype)
    if n == 1:
        with np.errstate(all="ignore"):
            ours = orc.standardize_point_cloud(x.copy())
            ref = reference["traj_ball_renderer"].TrajectoryBallRenderer.standardize_point_cloud(x.copy())
        np.testing.assert_array_equal(np.isnan(ours), np.isnan(ref))
        return
    ours = orc.standardize_point_cloud(x.copy())
    for mod, cls in CLASSES.values():
        ref = getattr(reference[mod], cls).standardize_point_cloud(x.copy())
        assert ref.dtype == np.float32 and ours.dtype == np.float32
        np.testing.assert_array_equal(ours, ref)
    if cols == 3:
        np.testing.assert_array_equal(ours, reference["example_renderer"].PointCloudRenderer.standardize_point_cloud(x.copy()))


@pytest.mark.parametrize("cols", [3, 6])
def test_transform_matches_reference(reference, orc, cols):
    x = orc.standardize_point_cloud(_rand(777, cols, 3, np.float64))
    for name, (mod, cls) in CLASSES.items():
        ref = getattr(reference[mod], cls).transform_coordinates(x.copy())
        ours = orc.transform_coordinates(x.copy(), flip_x=orc.PRESETS[name]["flip_x"])
        np.testing.assert_array_equal(ours, ref)
    if cols == 3:
        p = x.copy()[:, [2, 0, 1]]
        p[:, 0] *= -1
        p[:, 2] += 0.0125                       # example_renderer.py:171-173
        np.testing.assert_array_equal(orc.transform_coordinates(x.copy(), flip_x=True), p)


def test_camera_schedules_match_reference(reference, orc):
    for name, (mod, cls) in CLASSES.items():
        fn = getattr(reference[mod], cls).compute_camera_position
        for f in list(range(0, 220, 7)) + [19, 199, 200, 219]:
            assert tuple(fn(f, 220)) == orc.camera_position(name, f, 220), (name, f)


def test_compute_color_is_constant_grey(reference, orc):
    c = reference["example_renderer"].PointCloudRenderer.compute_color(0.1, 0.9, 0.5, noise_seed=3)
    assert c.tolist() == [0.3, 0.3, 0.3]
    assert reference["traj_ball_renderer"].TrajectoryBallRenderer.compute_color().tolist() == [0.3, 0.3, 0.3]
    out = orc.compute_color(np.zeros((4, 3), np.float32), mode=0)
    np.testing.assert_array_equal(out[:, :3], np.float32(0.3))


def test_emitted_scene_round_trips(reference, orc):
    """Centres / radius / camera parsed back out of the reference's XML equal what the oracle
    (and therefore the kernels) are handed: 'same centres, radii and camera'."""
    x = _rand(300, 3, 9, np.float32)
    r = reference["example_renderer"].PointCloudRenderer("a.npy")
    p = orc.transform_coordinates(orc.standardize_point_cloud(x))
    sc = scene_from_xml.parse_scene(r.generate_xml_content(p))
    np.testing.assert_array_equal(sc["centers"], p)
    assert np.all(sc["radius"] == np.float32(0.01))
    np.testing.assert_array_equal(sc["reflectance"], np.full((300, 3), 0.3, np.float32))
    pr = orc.PRESETS["example"]
    assert tuple(sc["origin"]) == orc.camera_position("example") and tuple(sc["target"]) == pr["target"]
    assert sc["fov"] == pr["fov"] and (sc["width"], sc["height"], sc["spp"]) == (1920, 1080, 256)
    assert sc["floor_z"] == pr["floor_z"] and sc["floor_min"] == pr["floor_min"] and sc["floor_max"] == pr["floor_max"]
    assert (sc["light_z"], sc["light_half"], sc["radiance"]) == (15.0, 8.0, 4.0)

    for name in ("traj_ball", "traj_original", "traj_b0", "traj_b1"):
        mod, cls = CLASSES[name]
        rr = getattr(reference[mod], cls)("f.npy")
        q = orc.transform_coordinates(orc.standardize_point_cloud(x), flip_x=orc.PRESETS[name]["flip_x"])
        sc = scene_from_xml.parse_scene(rr.generate_xml_content(q, frame_index=211, total_frames=220))
        np.testing.assert_array_equal(sc["centers"], q)
        assert tuple(sc["origin"]) == orc.camera_position(name, 211, 220)
        pr = orc.PRESETS[name]
        assert tuple(sc["target"]) == pr["target"] and sc["fov"] == pr["fov"]
        assert sc["floor_z"] == pr["floor_z"] and sc["floor_min"] == pr["floor_min"] and sc["floor_max"] == pr["floor_max"]


def _reference_trails(cls, pcl6, frame_index, tmp_path, monkeypatch):
    """Run the reference's own _add_velocity_trail for every point and read back the curve files
    it writes (first and last control point = tail and head, as Mitsuba would load them)."""
    monkeypatch.chdir(tmp_path)
    r = cls("f.npy")
    tails, heads, valid = [], [], []
    for idx, pt in enumerate(pcl6):
        segs = []
        r._add_velocity_trail(segs, pt[:3], pt[3:6], point_index=idx, frame_index=frame_index)
        if not segs:
            valid.append(False); tails.append([0, 0, 0]); heads.append([0, 0, 0])
            continue
        rows = np.loadtxt(r.curve_files[-1])
        assert rows.shape == (21, 4) and np.all(rows[:, 3] == 0.0007)
        # interior control points are collinear with the ends to the file's 1e-6 resolution
        tt = np.linspace(0, 1, 20)[:, None]
        np.testing.assert_allclose(rows[:20, :3], rows[0, :3] + (rows[20, :3] - rows[0, :3]) * tt, atol=1.1e-6)
        tails.append(rows[0, :3]); heads.append(rows[20, :3]); valid.append(True)
    return np.float32(tails), np.float32(heads), np.array(valid)


@pytest.mark.parametrize("name,frame_index", [("traj_ball", 0), ("traj_ball", 7), ("traj_ball", 150), ("traj_vel", 3),
                                              ("traj_vel", 211), ("traj_vel", 219), ("traj_original", 199), ("traj_b0", 4), ("traj_b1", 205)])
def test_velocity_trails_match_reference_curve_files(reference, orc, tmp_path, monkeypatch, name, frame_index):
    """trail end points equal the control points the reference writes to its
    temp_curves/*.txt files (what Mitsuba's linearcurve loader would read), bit for bit as f32."""
    mod, cls = CLASSES[name]
    rng = np.random.default_rng(frame_index)
    x = (rng.standard_normal((300, 6)) * [1, 1, 1, 4, 4, 4]).astype(np.float32)
    x[5, 3:] = 0                                   # no velocity -> no trail
    x[6, 3:] = [30, -40, 5]                        # faster than the /10 normaliser -> clamped length
    pcl = orc.transform_coordinates(orc.standardize_point_cloud(x), orc.PRESETS[name]["flip_x"])
    t_ref, h_ref, v_ref = _reference_trails(getattr(reference[mod], cls), pcl, frame_index, tmp_path, monkeypatch)
    scale = orc.trail_length_scale(name, frame_index)
    for exact in (True, False):
        tail, head, valid = orc.velocity_trails(pcl, scale, exact_text=exact)
        np.testing.assert_array_equal(valid, v_ref)
        np.testing.assert_array_equal(tail[valid], t_ref[valid])
        np.testing.assert_array_equal(head[valid], h_ref[valid])
    assert not v_ref[5] and (v_ref.sum() == (299 if scale > 0 else 0))
