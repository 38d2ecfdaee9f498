"""Host-side logic of the package: presets / camera schedules against the reference's golden
outputs, file I/O, synthetic workloads, sharding arithmetic."""
import os

import numpy as np
import pytest

from pointcloud_render_b200 import io as pio
from pointcloud_render_b200 import sharding, synthetic
from pointcloud_render_b200.presets import PRESETS
from pointcloud_render_b200 import renderers


def test_preset_cameras_match_reference_golden(golden):
    g = golden("camera.npz")
    for name in ("traj", "traj_ball", "traj_vel", "traj_original", "traj_b0", "traj_b1"):
        got = np.array([PRESETS[name].camera_position(int(f), 220) for f in g["frames"]])
        np.testing.assert_array_equal(got, g[name])
    assert PRESETS["example"].camera_position(5) == (2.2, 2.2, 4.2)


@pytest.mark.parametrize("name", ["example", "traj_ball", "traj_original", "traj_b0", "traj_b1"])
def test_preset_scene_constants_match_emitted_xml(golden, name):
    g = golden(f"scene_{name}.npz")
    cfg = PRESETS[name]
    cam = cfg.camera(int(g["frame"]), 220)
    np.testing.assert_array_equal(np.float32(g["origin"]), np.array(list(cam.origin), np.float32))
    np.testing.assert_array_equal(np.float32(g["target"]), np.array(list(cam.target), np.float32))
    assert list(cam.up) == [0.0, 0.0, 1.0] and cam.fov_x_deg == float(g["fov"])
    assert (cam.near_clip, cam.far_clip) == (np.float32(g["near_clip"]), np.float32(g["far_clip"]))
    assert (cam.width, cam.height, cfg.spp) == (int(g["width"]), int(g["height"]), int(g["spp"]))
    st = cfg.style()
    assert st.radius == np.float32(g["radius"][0]) and list(st.const_rgb) == [np.float32(0.3)] * 3
    assert st.floor_z == np.float32(g["floor_z"])
    assert list(st.floor_min) == list(np.float32(g["floor_min"])) and list(st.floor_max) == list(np.float32(g["floor_max"]))
    assert (st.light_z, st.light_half, st.radiance) == (float(g["light_z"]), float(g["light_half"]), float(g["radiance"]))
    assert st.flip_x == int(cfg.flip_x) and st.z_lift == np.float32(0.0125)


def test_schedule_stretch_keeps_keyframes():
    cfg = PRESETS["traj_b0"].for_trajectory(500)
    assert cfg.last_motion_frame == 479 and cfg.fade_frames == 20
    assert cfg.camera_position(0, 500) == (-2.2, -3.3, 2.0)
    np.testing.assert_allclose(cfg.camera_position(479, 500), (-1.3, -2.5, 0.8), atol=1e-12)
    np.testing.assert_allclose(cfg.camera_position(499, 500), (-1.0, -2.0, 0.7), atol=1e-12)
    assert PRESETS["example"].for_trajectory(500) is PRESETS["example"]


def _write_ply(path, arr, names, fmt):
    with open(path, "wb") as f:
        f.write(b"ply\nformat %s 1.0\ncomment made by test\nelement vertex %d\n" % (fmt.encode(), len(arr)))
        for n in names:
            f.write(b"property float %s\n" % n.encode())
        f.write(b"element face 0\nproperty list uchar int vertex_indices\nend_header\n")
        if fmt == "ascii":
            for row in arr:
                f.write((" ".join(repr(float(v)) for v in row) + "\n").encode())
        else:
            f.write(arr.astype("<f4" if fmt == "binary_little_endian" else ">f4").tobytes())


@pytest.mark.parametrize("fmt", ["ascii", "binary_little_endian", "binary_big_endian"])
def test_ply_loader_variants(tmp_path, fmt):
    """load_point_cloud of traj_ball_renderer.py:223-279: vx,vy,vz else nx,ny,nz else xyz."""
    rng = np.random.default_rng(1)
    a = rng.standard_normal((50, 6)).astype(np.float32)
    p = str(tmp_path / "v.ply")
    _write_ply(p, a, ["x", "y", "z", "vx", "vy", "vz"], fmt)
    np.testing.assert_array_equal(pio.load_point_cloud(p), a)
    np.testing.assert_array_equal(pio.load_point_cloud(p, with_velocity=False), a[:, :3])   # example_renderer.py:108-109
    _write_ply(p, a, ["x", "y", "z", "nx", "ny", "nz"], fmt)
    np.testing.assert_array_equal(pio.load_point_cloud(p), a)
    _write_ply(p, a[:, :3], ["x", "y", "z"], fmt)
    np.testing.assert_array_equal(pio.load_point_cloud(p), a[:, :3])


def test_npy_npz_and_errors(tmp_path):
    a = np.arange(30, dtype=np.float64).reshape(10, 3)
    np.save(tmp_path / "a.npy", a)
    np.savez(tmp_path / "a.npz", pred=a)
    np.testing.assert_array_equal(pio.load_point_cloud(str(tmp_path / "a.npy")), a)
    np.testing.assert_array_equal(pio.load_point_cloud(str(tmp_path / "a.npz")), a)
    with pytest.raises(ValueError, match="Unsupported file format"):
        pio.load_point_cloud(str(tmp_path / "a.txt"))
    (tmp_path / "bad.ply").write_bytes(b"not a ply\n")
    with pytest.raises(ValueError):
        pio.load_point_cloud(str(tmp_path / "bad.ply"))


def test_png_writer(tmp_path):
    from PIL import Image
    img = np.random.default_rng(0).integers(0, 255, (12, 20, 4), dtype=np.uint8)
    pio.write_png(str(tmp_path / "o.png"), img)
    np.testing.assert_array_equal(np.asarray(Image.open(tmp_path / "o.png")), img[..., :3])


def test_renderer_naming_rules():
    r = renderers.TrajB1Renderer("batch_1/frame_0005_b1.ply", output_folder=None)
    assert (r.folder, r.filename) == ("batch_1", "frame_0005_b1")
    assert renderers.TrajB1Renderer.compute_camera_position(0) == (-3.5, -2.5, 2.8)
    assert renderers.PointCloudRenderer.compute_color(1, 2, 3, noise_seed=4).tolist() == [0.3, 0.3, 0.3]
    assert renderers.TrajectoryBallRenderer.compute_color().tolist() == [0.3, 0.3, 0.3]
    assert renderers.FixedFrame199Renderer.compute_camera_position(17) == (-1.8, -1.8, 1.8)


def test_synthetic_is_seeded_and_shaped():
    a, b = synthetic.cloud(100, "gauss", 3), synthetic.cloud(100, "gauss", 3)
    np.testing.assert_array_equal(a, b)
    assert synthetic.cloud(10, "cube").min() >= 0 and abs(np.linalg.norm(synthetic.cloud(500, "shell"), axis=1).mean() - 1) < 0.01
    t = synthetic.trajectory(4, 50, 6)
    assert t.shape == (4, 50, 6) and t.dtype == np.float32
    np.testing.assert_allclose(t[3, :, 4] - t[0, :, 4], -0.03, atol=1e-5)      # gravity on the input y axis
    # P0 and V come from one generator: frame 0 is cloud(n, seed) and V is NOT a multiple of P0 (a radial
    # expansion would be removed by the per-frame standardisation: every frame the same cloud, only radial trails)
    big = synthetic.trajectory(2, 4000, 6, seed=5)
    np.testing.assert_array_equal(big[0, :, :3], synthetic.cloud(4000, "gauss", 5))
    p0, v = big[0, :, :3].astype(np.float64), big[0, :, 3:].astype(np.float64)
    cos = np.abs((p0 * v).sum(1)) / (np.linalg.norm(p0, axis=1) * np.linalg.norm(v, axis=1))
    assert cos.mean() < 0.6 and abs(np.corrcoef(p0[:, 0], v[:, 0])[0, 1]) < 0.1
    r = synthetic.radii(1000)
    assert r.dtype == np.float32 and 0.005 <= r.min() and r.max() <= 0.015
    assert synthetic.CONFIGS["H"]["points"] == 1_000_000 and synthetic.CONFIGS["H"]["width"] == 1024


def test_shard_ranges_partition():
    for n in (0, 1, 7, 100, 1_000_003):
        for world in (1, 2, 3, 8):
            r = [sharding.frame_shard(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_finalize_stats_follows_reference_roundings(orc, dtype):
    x = (np.random.default_rng(2).standard_normal((1000, 3)) * 3 + 7).astype(dtype)
    tot = np.concatenate([x.astype(np.float64).sum(0), x.min(0), x.max(0)])
    st = sharding.finalize_stats(tot, len(x), dtype)
    assert st[9] == float(np.amax(x - np.amin(x, axis=0)))                      # example_renderer.py:97
    np.testing.assert_allclose(st[:3], np.mean(x, axis=0), rtol=1e-6 if dtype == np.float32 else 1e-14)
    out = ((x - st[:3].astype(dtype)) / dtype(st[9])).astype(np.float32)
    np.testing.assert_allclose(out, orc.standardize_point_cloud(x), atol=2e-7)


def test_async_writer_and_naming(tmp_path):
    """Output stage: frames > 199 are named frame_XXXX_b0 by every trajectory
    class (traj_ball_renderer.py:376); the writer pool produces the same bytes as a direct save."""
    from PIL import Image
    from pointcloud_render_b200 import output
    assert output.trajectory_frame_name("frame_0042_b1", 42) == "frame_0042_b1"
    assert output.trajectory_frame_name("frame_0199_b1", 205) == "frame_0205_b0"
    rng = np.random.default_rng(0)
    frames = rng.integers(0, 255, (6, 40, 64, 4), dtype=np.uint8)
    with output.AsyncImageWriter(workers=3) as w:
        futs = [w.submit(str(tmp_path / "out" / f"f{k}"), frames[k]) for k in range(6)]
        assert len(w.drain()) == 6 and all(f.done() for f in futs)
    for k in range(6):
        np.testing.assert_array_equal(np.asarray(Image.open(tmp_path / "out" / f"f{k}.png")), frames[k][..., :3])
    with output.AsyncImageWriter(fmt="npy") as w:
        w.submit(str(tmp_path / "raw"), frames[0])
    np.testing.assert_array_equal(np.load(tmp_path / "raw.npy"), frames[0][..., :3])
    with pytest.raises(ValueError):
        output.AsyncImageWriter(fmt="bmp")


def test_bench_workloads_and_ring_are_consistent():
    """bench.py's host-side definitions (no GPU): every workload's ring holds at least five steps (a four-step look-ahead never
    meets the slice being rendered) of inputs larger than the L2, its slots visit the whole camera schedule, the algorithmic
    bytes follow SURVEY.md 8d, and the config dict — the same function serves both arms — names what is rendered."""
    import importlib.util
    import os
    spec_ = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec_)
    spec_.loader.exec_module(bench)
    for name in ("H", "C2", "C3", "C4"):
        spec = bench.workload_spec(name, 0)
        B = spec["frames_per_step"]
        assert 1 <= B <= 64
        ring = bench.ring_frames(spec)
        assert ring % B == 0 and ring >= 5 * B
        assert ring * spec["input_bytes_per_frame"] >= 126e6                      # inputs larger than the L2
        idx = bench.ring_indices(spec, ring)
        assert len(idx) == ring and set(idx) == set(range(min(spec["frames"], ring))) or len(set(idx)) == min(spec["frames"], ring)
        b_in = 4 * spec["cols"] + (4 if spec["radii"] else 0)
        assert spec["algorithmic_bytes_per_frame"] == spec["points"] * b_in + spec["width"] * spec["height"] * 12
        cfg = bench.config_dict(name, spec, ring, 1)
        assert cfg["frames_per_step_per_gpu"] == B and cfg["points"] == spec["points"] and name in cfg["workload"]
    assert bench.workload_spec("H", 0)["points"] == 1_000_000 and bench.workload_spec("H", 16)["frames_per_step"] == 16
