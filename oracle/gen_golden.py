"""Generate tests/golden/*.npz by running the UNMODIFIED reference (stub-imported from
/root/reference, oracle/ref_import.py) on seeded synthetic inputs — TEST INFRASTRUCTURE.

Run here (the reference cannot travel to the GPU box):  python -m oracle.gen_golden
What gets pinned:
  standardize.npz   standardize_point_cloud / transform_coordinates outputs of every flavour
  camera.npz        compute_camera_position of every script at characteristic frames
  trails.npz        tail / head control points of the curve files _add_velocity_trail writes
  scene_*.npz       the scene generate_xml_content emits (centres, radius, reflectance, sensor,
                    floor, emitter), parsed back by oracle/scene_from_xml.py
  droplets.npz      §8f-2: the droplet OBJ's vertices / faces, the matrices generate_rotation_matrix_from_velocity
                    and generate_random_rotation_matrix print (as float32), and the control points of the
                    curve files _add_trail_lines writes for histories of 1..25 frames
  vis_example.npz   visibility ids of the C oracle for the example scene at 200x150 and the
                    sha256 of the full C1 (800x600) key buffer — NOT pinned by the reference
                    (Mitsuba absent): guards the oracle against silent change only.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import pcr_oracle as orc  # noqa: E402
from oracle import ref_import, scene_from_xml  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
CAMERA_FRAMES = np.array([0, 1, 19, 20, 57, 100, 198, 199, 200, 205, 210, 219])


def synth(n, cols, seed, dtype):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((n, cols)) * np.array([1.0, 0.6, 1.7, 3, 3, 3][:cols]) + np.array([0.3, -2.0, 5.0, 0, 0, 0][:cols])
    return np.ascontiguousarray(a, dtype=dtype)


def droplet_inputs(h, n=48, seed=100):
    """Seeded history of h transformed frames + current positions, with the degenerate cases the reference
    branches on: points that never move, points that move once and stop, a trail that returns to its start."""
    rng = np.random.default_rng(seed + h)
    base = (rng.standard_normal((n, 3)) * 0.3).astype(np.float32)
    vel = (rng.standard_normal((n, 3)) * 0.02).astype(np.float32)
    hist = np.stack([(base + vel * np.float32(k) + np.float32(0.001 * k * k)).astype(np.float32) for k in range(h)])
    pos = (base + vel * np.float32(h) + np.float32(0.001 * h * h)).astype(np.float32)
    hist[:, :4] = hist[0:1, :4]
    pos[:3] = hist[0, :3]
    if h > 2:
        hist[1:, 5] = hist[1, 5]
        pos[6] = hist[0, 6]
    if h > 3:
        hist[2, 7] = np.nan
    return hist, pos


def gen_droplets(ref):
    import contextlib
    import io
    import tempfile
    traj = ref["traj_renderer"]
    g = {}
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                path = traj.TrajectoryRenderer._create_droplet_mesh()
            V, F = [], []
            for line in open(path):
                p = line.split()
                if p and p[0] == "v":
                    V.append([float(x) for x in p[1:4]])
                elif p and p[0] == "f":
                    F.append([int(x) - 1 for x in p[1:4]])
            g["mesh_verts"] = np.array(V, np.float64).astype(np.float32)
            g["mesh_faces"] = np.array(F, np.int32)
            rng = np.random.default_rng(42)
            pcl6 = (rng.standard_normal((400, 6)) * [0.3, 0.3, 0.3, 3, 3, 3]).astype(np.float32)
            pcl6[0, 3:] = 0
            pcl6[1, 3:] = [0, 0, -2]
            pcl6[2, 3:] = [0, 0, 3]
            pcl6[3, 3:] = [1e-9, 0, 5]
            pcl6[4, 3:] = [0, 1e-7, 0]
            pcl6[5, 3:] = [1e-7, 0, 0]
            g["pcl6"] = pcl6
            g["xf_velocity"] = np.array([traj.TrajectoryRenderer.generate_rotation_matrix_from_velocity(r[3:6], r[:3])
                                         for r in pcl6]).reshape(-1, 4, 4)[:, :3, :].reshape(-1, 12).astype(np.float32)
            g["xf_random"] = np.array([traj.TrajectoryRenderer.generate_random_rotation_matrix(i, pcl6[i, :3])
                                       for i in range(64)]).reshape(-1, 4, 4)[:, :3, :].reshape(-1, 12).astype(np.float32)
            r = traj.TrajectoryRenderer("x.npy", droplet_mesh_path="m.obj")
            hs = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 15, 17, 19, 20, 25]
            g["history_lengths"] = np.array(hs)
            for h in hs:
                hist, pos = droplet_inputs(h)
                n = pos.shape[0]
                ctrl, cnt = np.zeros((n, 21, 3), np.float32), np.zeros(n, np.int32)
                for i in range(n):
                    segs = []
                    r.curve_files = []
                    r._add_trail_lines(segs, pos[i], np.zeros(3), [hist[k, i] for k in range(h)], point_index=i)
                    if segs:
                        rows = np.loadtxt(r.curve_files[-1], ndmin=2)[:, :3].astype(np.float32)
                        cnt[i] = len(rows)
                        ctrl[i, :len(rows)] = rows
                g[f"hist_{h}"], g[f"pos_{h}"], g[f"ctrl_{h}"], g[f"count_{h}"] = hist, pos, ctrl, cnt
        finally:
            os.chdir(cwd)
    np.savez_compressed(os.path.join(OUT, "droplets.npz"), **g)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_import.load()
    if "droplets" in sys.argv[1:]:
        gen_droplets(ref)
        return
    ex, ball, orig, b0, b1, traj, vel = (ref[k] for k in ("example_renderer", "traj_ball_renderer", "traj_original",
                                                           "traj_b0", "traj_b1", "traj_renderer", "traj_vel_renderer"))

    # ---- a1 / a2 -----------------------------------------------------------------------------
    g = {}
    for tag, n, cols, dtype in (("f32_3", 257, 3, np.float32), ("f64_3", 257, 3, np.float64),
                                ("f32_6", 193, 6, np.float32), ("f64_6", 193, 6, np.float64)):
        x = synth(n, cols, 11 + cols, dtype)
        g[f"in_{tag}"] = x
        if cols == 3:
            s = ex.PointCloudRenderer.standardize_point_cloud(x.copy())
            g[f"std_example_{tag}"] = s
            p = s[:, [2, 0, 1]]
            p[:, 0] *= -1
            p[:, 2] += 0.0125                                     # example_renderer.py:171-173
            g[f"xf_example_{tag}"] = p
        for name, cls in (("ball", ball.TrajectoryBallRenderer), ("traj", traj.TrajectoryRenderer),
                          ("vel", vel.TrajectoryVelRenderer), ("orig", orig.FixedFrame199Renderer),
                          ("b0", b0.FixedFrame199Renderer), ("b1", b1.FixedFrame199Renderer)):
            s = cls.standardize_point_cloud(x.copy())
            g[f"std_{name}_{tag}"] = s
            g[f"xf_{name}_{tag}"] = cls.transform_coordinates(s.copy())
    np.savez_compressed(os.path.join(OUT, "standardize.npz"), **g)

    # ---- a4 ----------------------------------------------------------------------------------
    cams = {"frames": CAMERA_FRAMES}
    for name, cls in (("traj_ball", ball.TrajectoryBallRenderer), ("traj_vel", vel.TrajectoryVelRenderer),
                      ("traj_original", orig.FixedFrame199Renderer), ("traj_b0", b0.FixedFrame199Renderer),
                      ("traj_b1", b1.FixedFrame199Renderer)):
        cams[name] = np.array([cls.compute_camera_position(int(f), 220) for f in CAMERA_FRAMES], np.float64)
    cams["traj"] = np.array([traj.TrajectoryRenderer.compute_camera_position(int(f), 220) for f in CAMERA_FRAMES], np.float64)
    np.savez_compressed(os.path.join(OUT, "camera.npz"), **cams)

    # ---- a6: emitted scenes -----------------------------------------------------------------
    def dump(name, xml, pcl_in, frame):
        sc = scene_from_xml.parse_scene(xml)
        np.savez_compressed(os.path.join(OUT, f"scene_{name}.npz"), input=pcl_in, frame=frame,
                            centers=sc["centers"], radius=sc["radius"], reflectance=sc["reflectance"],
                            origin=np.array(sc["origin"]), target=np.array(sc["target"]), up=np.array(sc["up"]),
                            fov=sc["fov"], near_clip=sc["near_clip"], far_clip=sc["far_clip"],
                            width=sc["width"], height=sc["height"], spp=sc["spp"],
                            floor_z=sc["floor_z"], floor_min=np.array(sc["floor_min"]), floor_max=np.array(sc["floor_max"]),
                            light_z=sc["light_z"], light_half=sc["light_half"], radiance=sc["radiance"])
        return sc

    rng = np.random.default_rng(0)
    c1 = rng.standard_normal((2048, 3)).astype(np.float32)                       # config C1's cloud
    r = ex.PointCloudRenderer("c1.npy")
    p = r.standardize_point_cloud(c1.copy())
    p = p[:, [2, 0, 1]]
    p[:, 0] *= -1
    p[:, 2] += 0.0125
    sc_ex = dump("example", r.generate_xml_content(p), c1, 0)

    x3 = synth(160, 3, 5, np.float32)       # position-only input: no trail files are written
    for name, cls, frame in (("traj_ball", ball.TrajectoryBallRenderer, 57), ("traj_original", orig.FixedFrame199Renderer, 199),
                             ("traj_b0", b0.FixedFrame199Renderer, 205), ("traj_b1", b1.FixedFrame199Renderer, 100)):
        rr = cls("f.npy")
        q = rr.transform_coordinates(rr.standardize_point_cloud(x3.copy()))
        dump(name, rr.generate_xml_content(q, frame_index=frame, total_frames=220), x3, frame)

    # ---- 8f-1: velocity trails = the control points the reference writes to temp_curves/*.txt ---
    import tempfile
    tg = {}
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            for name, cls, frame in (("traj_ball", ball.TrajectoryBallRenderer, 7), ("traj_vel", vel.TrajectoryVelRenderer, 211),
                                     ("traj_b0", b0.FixedFrame199Renderer, 4), ("traj_ball", ball.TrajectoryBallRenderer, 150)):
                rng = np.random.default_rng(frame)
                x = (rng.standard_normal((200, 6)) * [1, 1, 1, 4, 4, 4]).astype(np.float32)
                x[5, 3:] = 0
                x[6, 3:] = [30, -40, 5]
                rr = cls("f.npy")
                pcl = rr.transform_coordinates(rr.standardize_point_cloud(x.copy()))
                tails, heads, valid = np.zeros((200, 3), np.float32), np.zeros((200, 3), np.float32), np.zeros(200, bool)
                for idx, pt in enumerate(pcl):
                    segs = []
                    rr._add_velocity_trail(segs, pt[:3], pt[3:6], point_index=idx, frame_index=frame)
                    if segs:
                        rows = np.loadtxt(rr.curve_files[-1])
                        tails[idx], heads[idx], valid[idx] = rows[0, :3], rows[-1, :3], True
                key = f"{name}_{frame}"
                tg[f"raw_{key}"], tg[f"pcl_{key}"], tg[f"tail_{key}"], tg[f"head_{key}"], tg[f"valid_{key}"] = x, pcl, tails, heads, valid
        finally:
            os.chdir(cwd)
    np.savez_compressed(os.path.join(OUT, "trails.npz"), **tg)

    gen_droplets(ref)

    # ---- oracle self-pin (unpinned by the reference) -----------------------------------------
    pos4 = np.concatenate([sc_ex["centers"], sc_ex["radius"][:, None]], axis=1).astype(np.float32)
    scene = orc.make_scene(True, sc_ex["floor_z"], sc_ex["floor_min"], sc_ex["floor_max"], 1.0, sc_ex["light_z"],
                           sc_ex["light_half"], sc_ex["radiance"], 1.0)
    small = orc.camera_frame(sc_ex["origin"], sc_ex["target"], sc_ex["up"], sc_ex["fov"], sc_ex["near_clip"],
                             sc_ex["far_clip"], 200, 150)
    vis_small = orc.visibility(pos4, small, scene, brute_force=True)
    full = orc.camera_frame(sc_ex["origin"], sc_ex["target"], sc_ex["up"], sc_ex["fov"], sc_ex["near_clip"],
                            sc_ex["far_clip"], 800, 600)
    vis_full = orc.visibility(pos4, full, scene)
    attr4 = np.concatenate([sc_ex["reflectance"], np.zeros((len(pos4), 1), np.float32)], axis=1)
    img_small = orc.shade(vis_small, pos4, attr4, small, scene)
    np.savez_compressed(os.path.join(OUT, "vis_example.npz"), keys_200x150=vis_small, rgba_200x150=img_small,
                        sha256_800x600=hashlib.sha256(vis_full.tobytes()).hexdigest())
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
