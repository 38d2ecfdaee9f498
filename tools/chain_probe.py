#!/usr/bin/env python
"""Do the serial-mean chains of several prepared batches run side by side?  (diagnostics)"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointcloud_render_b200 import _native, synthetic
from pointcloud_render_b200.presets import PRESETS

n, W, H = 1_000_000, 1024, 1024
cfg = PRESETS["traj_ball"].for_trajectory(100)
style = cfg.style()
x = torch.from_numpy(synthetic.trajectory(8, n, 3, seed=0)).cuda()
chunks = [x.clone() for _ in range(6)]
ctx = _native.Context(device=0, max_points=n, max_w=W, max_h=H, max_batch=32)
cams = [cfg.camera(f, 100, W, H) for f in range(8)]
for nchunks in (1, 2, 4, 6):
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for c in chunks[:nchunks]:
            ctx.prefetch_frames(c, style)
        for c in chunks[:nchunks]:
            ctx.render_frames(c, cams, style)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"{nchunks} chunks of 8 frames prefetched together: {dt * 1e3:.2f} ms total")
ctx.close()
