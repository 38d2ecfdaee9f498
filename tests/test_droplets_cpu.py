"""SURVEY.md §8f-2 on the CPU: the droplet / history-trail oracle against the UNMODIFIED reference (live where
/root/reference exists, and through tests/golden/droplets.npz everywhere), the host-side mesh / legacy-RNG
tables of the product, and the mesh / polyline caster of oracle/raycast.c against closed-form answers."""
import contextlib
import io
import os

import numpy as np
import pytest

from oracle import droplet_oracle as do
from oracle.gen_golden import droplet_inputs
from pointcloud_render_b200 import droplets


@pytest.fixture(scope="module")
def g(golden):
    return golden("droplets.npz")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_mesh_matches_the_reference_obj(g):
    v, f = do.droplet_mesh()
    np.testing.assert_array_equal(bits(v), bits(g["mesh_verts"]))
    np.testing.assert_array_equal(f, g["mesh_faces"])
    # the product's own construction (no file is written) is the same mesh
    np.testing.assert_array_equal(bits(droplets.droplet_vertices()), bits(g["mesh_verts"]))
    np.testing.assert_array_equal(droplets.droplet_faces(), g["mesh_faces"])
    assert v.shape == (340, 3) and f.shape == (640, 3)
    # known answers: pole at (0,0,r), tip at (0,0,-0.8*length), every ring's z strictly below the previous one
    np.testing.assert_allclose(v[0], [0, 0, 0.008], atol=1e-7)
    np.testing.assert_allclose(v[-1], [0, 0, -0.028], atol=1e-6)
    assert np.all(np.diff(v.reshape(17, 20, 3)[:, 0, 2]) < 0)


def test_rotations_match_the_reference_matrices(g):
    pcl6 = g["pcl6"]
    xf = do.to_world_f32(do.rotation_from_velocity(pcl6[:, 3:6]), pcl6[:, :3])
    np.testing.assert_array_equal(bits(xf), bits(g["xf_velocity"]))
    xr = do.to_world_f32(do.random_rotation(range(64)), pcl6[:64, :3])
    np.testing.assert_array_equal(bits(xr), bits(g["xf_random"]))
    np.testing.assert_array_equal(bits(droplets.random_rotations(64)), bits(g["xf_random"].reshape(-1, 3, 4)[:, :, :3].reshape(-1, 9)))
    # the tip axis (0,0,-1) lands on the velocity direction; rotations are orthonormal
    R = do.rotation_from_velocity(pcl6[6:, 3:6])
    tip = R @ np.array([0.0, 0.0, -1.0])
    v = pcl6[6:, 3:6].astype(np.float64)
    np.testing.assert_allclose(tip, v / np.linalg.norm(v, axis=1, keepdims=True), atol=1e-12)
    np.testing.assert_allclose(R @ R.transpose(0, 2, 1), np.broadcast_to(np.eye(3), R.shape), atol=1e-12)
    # degenerate inputs the reference branches on: zero velocity and velocity along the default axis -> identity;
    # opposite to it -> half turn
    np.testing.assert_array_equal(do.rotation_from_velocity(pcl6[:2, 3:6]), np.broadcast_to(np.eye(3), (2, 3, 3)))
    np.testing.assert_allclose(do.rotation_from_velocity(pcl6[2:3, 3:6])[0] @ [0, 0, -1.0], [0, 0, 1.0], atol=1e-12)


def test_history_trails_match_the_reference_curve_files(g):
    for h in g["history_lengths"]:
        hist, pos = droplet_inputs(int(h))
        np.testing.assert_array_equal(bits(hist), bits(g[f"hist_{h}"]))
        ctrl, cnt = do.history_trails(hist, pos)
        np.testing.assert_array_equal(cnt, g[f"count_{h}"])
        np.testing.assert_array_equal(bits(ctrl), bits(g[f"ctrl_{h}"]))
        if h < 2:
            assert not cnt.any()
        else:
            assert 3 <= cnt.max() <= 21 and (cnt == 0).sum() >= 3      # stationary points draw nothing


def test_oracle_against_the_live_reference(reference):
    """Same pins with fresh random inputs, run through the reference itself (skipped on the GPU box)."""
    traj = reference["traj_renderer"].TrajectoryRenderer
    vel = reference["traj_vel_renderer"].TrajectoryVelRenderer
    rng = np.random.default_rng(2024)
    pcl6 = (rng.standard_normal((300, 6)) * [0.3, 0.3, 0.3, 5, 5, 5]).astype(np.float32)
    for cls in (traj, vel):
        want = np.array([cls.generate_rotation_matrix_from_velocity(r[3:6], r[:3]) for r in pcl6]).reshape(-1, 4, 4)
        got = do.to_world_f32(do.rotation_from_velocity(pcl6[:, 3:6]), pcl6[:, :3])
        np.testing.assert_array_equal(bits(got), bits(want[:, :3, :].reshape(-1, 12).astype(np.float32)))
    import tempfile
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            r = traj("x.npy", droplet_mesh_path="m.obj")
            for h in (2, 6, 14, 20, 23):
                hist, pos = droplet_inputs(h, n=24, seed=7)
                ctrl, cnt = do.history_trails(hist, pos)
                for i in range(pos.shape[0]):
                    segs, r.curve_files = [], []
                    with contextlib.redirect_stdout(io.StringIO()):
                        r._add_trail_lines(segs, pos[i], np.zeros(3), [hist[k, i] for k in range(h)], point_index=i)
                    if not segs:
                        assert cnt[i] == 0
                        continue
                    rows = np.loadtxt(r.curve_files[-1], ndmin=2)
                    assert np.all(rows[:, 3] == 0.0007)
                    assert len(rows) == cnt[i]
                    np.testing.assert_array_equal(bits(rows[:, :3].astype(np.float32)), bits(ctrl[i, :cnt[i]]))
        finally:
            os.chdir(cwd)


def test_mesh_caster_known_answers(orc):
    """oracle/raycast.c VA-3: brute-force and boxed modes agree; a droplet looked at along its axis shows a
    disc of the cap's radius at the cap's depth; polylines behave like the capsules they are made of."""
    fr = orc.camera_frame((0.0, 0.0, 2.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 2.0, 0.1, 100.0, 201, 201)
    sc = orc.make_scene(has_floor=False)
    base = orc.visibility(np.zeros((0, 4), np.float32), fr, sc)
    xf = np.array([[1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0]], np.float32)        # identity at the origin: pole towards the eye
    a = do.add_droplets(base, xf, fr, brute_force=True)
    b = do.add_droplets(base, xf, fr)
    np.testing.assert_array_equal(a, b)
    ids = (a & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    depth = (a >> np.uint64(32)).astype(np.uint32).view(np.float32)
    assert ids[100, 100] == 0 and abs(depth[100, 100] - (2.0 - 0.008)) < 2e-6        # the pole, 8 mm in front of the origin
    # silhouette = the 20-gon of the widest ring (the cap stops at theta = pi/3, so that is ring 5, not the equator)
    v = do.droplet_mesh()[0]
    k = np.argmax(np.hypot(v[:, 0], v[:, 1]))
    r_px = np.hypot(v[k, 0], v[k, 1]) / (2.0 - v[k, 2]) / np.tan(np.radians(1.0)) * 100.5
    covered = (ids == 0).sum()
    assert np.pi * (r_px * np.cos(np.pi / 20)) ** 2 * 0.97 < covered < np.pi * r_px ** 2 * 1.03
    # a polyline of two collinear segments = one capsule
    ctrl = np.zeros((1, 21, 3), np.float32)
    ctrl[0, :3] = [[-0.02, 0.01, 0], [0.0, 0.01, 0], [0.02, 0.01, 0]]
    p = do.add_polylines(base, ctrl, np.array([3], np.int32), fr, 7, radius=0.004)
    q = orc.add_trails(base, ctrl[:, 0], ctrl[:, 2], [True], fr, 7, radius=0.004)
    # same pixels, same depth up to a few ulp (each segment's test is rooted at its own end point)
    np.testing.assert_array_equal(p & np.uint64(0xFFFFFFFF), q & np.uint64(0xFFFFFFFF))
    dp, dq = ((x >> np.uint64(32)).astype(np.uint32).view(np.float32) for x in (p, q))
    hit = (p & np.uint64(0xFFFFFFFF)) == 7
    np.testing.assert_allclose(dp[hit], dq[hit], rtol=2e-6)
    np.testing.assert_array_equal(p, do.add_polylines(base, ctrl, np.array([3], np.int32), fr, 7, radius=0.004, brute_force=True))
    assert ((p & np.uint64(0xFFFFFFFF)) == 7).sum() > 50


def test_ring_profile_normals():
    prof = do.ring_profile(do.droplet_mesh()[0])
    assert prof.shape == (17, 4)
    np.testing.assert_allclose(np.hypot(prof[:, 2], prof[:, 3]), 1.0, atol=1e-6)
    np.testing.assert_allclose(prof[0, 2:], [0, 1], atol=1e-7)
    np.testing.assert_allclose(prof[-1, 2:], [0, -1], atol=1e-7)
    # on the spherical cap the smooth normal is radial: (sin theta, cos theta)
    th = np.pi * np.arange(1, 5) / 16
    np.testing.assert_allclose(prof[1:5, 2:], np.stack([np.sin(th), np.cos(th)], 1), atol=2e-3)


def test_droplet_mesh_path_is_loaded_or_rejected(tmp_path):
    """droplet_mesh_path (traj_renderer.py:93-99): an OBJ with the ring structure of _create_droplet_mesh is loaded
    for pcr_set_droplet_mesh; anything else raises — it is never silently ignored."""
    from pointcloud_render_b200 import droplets
    verts, faces = droplets.droplet_vertices(), droplets.droplet_faces()
    good = tmp_path / "droplet.obj"
    droplets.write_obj(good, verts, faces)
    v, rings, segs = droplets.load_ring_mesh_obj(good)
    assert (rings, segs) == (droplets.N_RINGS, droplets.N_SEGMENTS)
    np.testing.assert_array_equal(v, verts)
    # a different surface of revolution with the same topology: a 12 x 8 ellipsoid
    rows = []
    for i in range(13):
        th = np.pi * i / 12
        for j in range(8):
            ph = 2 * np.pi * j / 8
            rows.append([0.01 * np.sin(th) * np.cos(ph), 0.01 * np.sin(th) * np.sin(ph), 0.02 * np.cos(th)])
    f = []
    for i in range(12):
        for j in range(8):
            v0, v1 = i * 8 + j, i * 8 + (j + 1) % 8
            f += [[v0, v0 + 8, v1], [v1, v0 + 8, v1 + 8]]
    other = tmp_path / "ellipsoid.obj"
    droplets.write_obj(other, rows, f)
    v2, r2, s2 = droplets.load_ring_mesh_obj(other)
    assert (r2, s2, v2.shape) == (12, 8, (104, 3)) and v2.dtype == np.float32
    # not ring structured / not a surface of revolution -> ValueError
    bad = tmp_path / "bad.obj"
    bad.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n")
    with pytest.raises(ValueError):
        droplets.load_ring_mesh_obj(bad)
    skew = np.array(rows)
    skew[20, 0] += 0.001
    droplets.write_obj(bad, skew, f)
    with pytest.raises(ValueError):
        droplets.load_ring_mesh_obj(bad)
    from pointcloud_render_b200 import renderers
    with pytest.raises(ValueError):
        renderers.TrajectoryRenderer(None, droplet_mesh_path=str(bad))
    assert renderers.TrajectoryRenderer(None, droplet_mesh_path=str(other))._mesh[1:] == (12, 8)
