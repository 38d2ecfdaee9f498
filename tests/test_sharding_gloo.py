"""N>1 host logic on CPU: world_size-2 gloo process group running the two multi-GPU modes of
pointcloud_render_b200/sharding.py.  The compute between the collectives is the CPU oracle here
(the GPU box runs the same collectives over NCCL with the pcr kernels in between)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import pcr_oracle as orc
    from pointcloud_render_b200 import sharding, synthetic
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---------------- point-sharded cloud: C0 + C1 ----------------
        n, W, H = 5000, 160, 120
        cloud = synthetic.cloud(n, "gauss", 5, np.float32)
        a, b = sharding.point_shard(n, rank, world)
        mine = cloud[a:b]
        part = np.concatenate([mine.astype(np.float64).sum(0), mine.min(0), mine.max(0)])
        stats = sharding.allreduce_stats(torch.from_numpy(part), b - a, np.float32).numpy()
        std = ((mine - stats[:3].astype(np.float32)) / np.float32(stats[9])).astype(np.float32)
        pos = orc.transform_coordinates(std, True)
        pos4 = np.concatenate([pos, np.full((b - a, 1), 0.02, np.float32)], axis=1)
        pr = orc.PRESETS["traj_ball"]
        frame = orc.camera_frame(orc.camera_position("traj_ball", 150), pr["target"], (0, 0, 1), pr["fov"], 0.1, 100.0, W, H)
        scene = orc.make_scene(True, pr["floor_z"], pr["floor_min"], pr["floor_max"])
        vis = torch.from_numpy(orc.visibility(pos4, frame, scene, id_base=a).view(np.int64))
        sharding.zmerge_(vis)
        attr4 = np.full((b - a, 4), 0.3, np.float32)
        rgba = torch.from_numpy(orc.shade(vis.numpy().view(np.uint64), pos4, attr4, frame, scene, id_base=a, owner_only=True))
        sharding.assemble_image_(rgba)
        # ---------------- frame-sharded trajectory: no collective ----------------
        F = 7
        f0, f1 = sharding.frame_shard(F, rank, world)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), vis=vis.numpy(), rgba=rgba.numpy(), stats=stats,
                 frames=np.arange(f0, f1))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_point_and_frame_sharding(tmp_path, orc):
    import torch.multiprocessing as mp
    from pointcloud_render_b200 import synthetic
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (np.load(tmp_path / f"rank{k}.npz") for k in range(world))
    # every rank ends with the same merged buffers
    np.testing.assert_array_equal(r0["vis"], r1["vis"])
    np.testing.assert_array_equal(r0["rgba"], r1["rgba"])
    np.testing.assert_array_equal(r0["stats"], r1["stats"])
    # ... and they equal the unsharded render (min is exact and order independent)
    n, W, H = 5000, 160, 120
    cloud = synthetic.cloud(n, "gauss", 5, np.float32)
    stats = r0["stats"]
    assert stats[9] == float(np.amax(cloud - np.amin(cloud, axis=0)))
    std = ((cloud - stats[:3].astype(np.float32)) / np.float32(stats[9])).astype(np.float32)
    np.testing.assert_allclose(std, orc.standardize_point_cloud(cloud), atol=2e-7)
    pos4 = np.concatenate([orc.transform_coordinates(std, True), np.full((n, 1), 0.02, np.float32)], axis=1)
    pr = orc.PRESETS["traj_ball"]
    frame = orc.camera_frame(orc.camera_position("traj_ball", 150), pr["target"], (0, 0, 1), pr["fov"], 0.1, 100.0, W, H)
    scene = orc.make_scene(True, pr["floor_z"], pr["floor_min"], pr["floor_max"])
    full = orc.visibility(pos4, frame, scene)
    np.testing.assert_array_equal(r0["vis"].view(np.uint64), full)
    img = orc.shade(full, pos4, np.full((n, 4), 0.3, np.float32), frame, scene)
    np.testing.assert_array_equal(r0["rgba"], img)
    # frame sharding covers every frame exactly once
    assert sorted(np.concatenate([r0["frames"], r1["frames"]]).tolist()) == list(range(7))
