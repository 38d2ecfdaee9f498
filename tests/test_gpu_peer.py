"""Fused z-merge over peer memory (pcr_peer_*) emulated on ONE device: G contexts play G ranks,
their "peer" pointers are plain device pointers of the same process, the cross-rank barriers are stream
synchronisations.  The merged rows and the image must be byte-identical to a single-GPU render of the whole cloud
(min is exact and order independent).  The real thing — CUDA IPC + atomicMin over NVLink, one process per GPU —
runs in tools/multi_gpu_check.py (tests/test_multi_gpu.py) where the box has >= 2 GPUs.  `-m gpu`."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from pointcloud_render_b200 import _native, sharding, synthetic  # noqa: E402
from pointcloud_render_b200.presets import PRESETS  # noqa: E402


@pytest.mark.parametrize("world,n,W,H,occlusion,pair_cap", [
    (2, 200_000, 1024, 768, -1, 0),        # pre-pass on (n >= 131072 per shard is NOT reached: 100 k each -> off) ...
    (3, 600_000, 1000, 701, -1, 0),        # ... here on (200 k per shard); rows do not divide evenly
    (4, 400_000, 640, 480, 1, 0),          # pre-pass forced on
    (2, 300_000, 800, 600, 0, 60_000),     # no lists: the per-sphere overflow raster pushes too
    (1, 50_000, 320, 240, -1, 0),          # a mesh of one
])
def test_fused_merge_equals_single_gpu(lib, world, n, W, H, occlusion, pair_cap):
    assert torch.cuda.is_available()
    cfg = PRESETS["example"]
    style, cam = cfg.style(color_mode=1), cfg.camera(0, 1, W, H)
    cloud = torch.from_numpy(synthetic.cloud(n, "gauss", seed=world)).cuda()
    shards = [sharding.point_shard(n, r, world) for r in range(world)]
    ctxs = [_native.Context(device=0, max_points=n, max_w=W, max_h=H, max_batch=1, pair_capacity=pair_cap) for _ in range(world)]
    try:
        for c in ctxs:
            c.set_occlusion(mode=occlusion)
        ptrs = [c.peer_alloc(W, H) for c in ctxs]
        for r, c in enumerate(ctxs):
            c.peer_set(r, world, [p[0] for p in ptrs], [p[1] for p in ptrs], dst_rank=world - 1)
        # poison the buffers: every pixel must be rewritten by the frame
        for c in ctxs:
            m, im = c.peer_buffers()
            m.fill_(-1)
            im.fill_(7)
        parts = torch.stack([ctxs[r].stats_partial(cloud[a:b]) for r, (a, b) in enumerate(shards)]).contiguous()
        stats = ctxs[0].finalize_stats(parts, n)
        for c in ctxs:
            c.peer_begin_frame(cam, style)
        torch.cuda.synchronize()                                   # "all-gather": every rank's rows are initialised
        vis = [ctxs[r].render_shard_peer(cloud[a:b], stats, cam, style, id_base=a) for r, (a, b) in enumerate(shards)]
        torch.cuda.synchronize()                                   # barrier: all pushes have landed
        for r, (a, b) in enumerate(shards):
            ctxs[r].shade_shard_peer(vis[r], cloud[a:b], stats, cam, style, id_base=a)
        torch.cuda.synchronize()
        # single-GPU reference through the same fused entries
        vis1 = ctxs[0].render_shard(cloud, stats, cam, style, id_base=0)
        rgba1 = ctxs[0].shade_shard(vis1, cloud, stats, cam, style, id_base=0, owner_only=False)
        merged = torch.empty_like(vis1)
        for r, c in enumerate(ctxs):
            y0, y1 = sharding.frame_shard(H, r, world)
            merged[y0:y1] = c.peer_buffers()[0][y0:y1]
        assert torch.equal(merged, vis1)
        assert torch.equal(ctxs[world - 1].peer_buffers()[1], rgba1)
        ids = _native.keys_to_ids(vis1)
        assert (ids < n).sum() > 1000
        if pair_cap:
            assert ctxs[0].counters()["overflow_frames"] == 1
        # a second frame with another camera reuses the buffers
        cam2 = PRESETS["traj_ball"].camera(120, 220, W, H)
        st2 = PRESETS["traj_ball"].style()
        for c in ctxs:
            c.peer_begin_frame(cam2, st2)
        torch.cuda.synchronize()
        vis = [ctxs[r].render_shard_peer(cloud[a:b], stats, cam2, st2, id_base=a) for r, (a, b) in enumerate(shards)]
        torch.cuda.synchronize()
        for r, (a, b) in enumerate(shards):
            ctxs[r].shade_shard_peer(vis[r], cloud[a:b], stats, cam2, st2, id_base=a)
        torch.cuda.synchronize()
        v2 = ctxs[0].render_shard(cloud, stats, cam2, st2, id_base=0)
        assert torch.equal(ctxs[world - 1].peer_buffers()[1], ctxs[0].shade_shard(v2, cloud, stats, cam2, st2, id_base=0, owner_only=False))
    finally:
        for c in ctxs:
            c.close()


def test_peer_entries_reject_bad_arguments(lib):
    c = _native.Context(device=0, max_points=1024, max_w=64, max_h=64, max_batch=1)
    cfg = PRESETS["example"]
    pts = torch.zeros((16, 3), dtype=torch.float32, device="cuda")
    stats = torch.zeros(10, dtype=torch.float64, device="cuda")
    with pytest.raises(RuntimeError, match="pcr_peer_set"):
        c.render_shard_peer(pts, stats, cfg.camera(0, 1, 64, 64), cfg.style())
    with pytest.raises(RuntimeError, match="pcr_peer_alloc"):
        c.peer_set(0, 1, [1], [1])
    m, im = c.peer_alloc(64, 48)
    with pytest.raises(RuntimeError):
        c.peer_set(0, 2, [m, 0], [im, 0])                      # NULL peer
    with pytest.raises(RuntimeError):
        c.peer_set(1, 2, [m, m], [im, 12345])                  # entry [rank] is not this context's buffer
    c.peer_set(0, 1, [m], [im])
    with pytest.raises(RuntimeError, match="does not match"):
        c.peer_begin_frame(cfg.camera(0, 1, 64, 64), cfg.style())
    c.close()
