import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from pointcloud_render_b200 import _native, synthetic
from pointcloud_render_b200.presets import PRESETS
n, W, H = 1_000_000, 1024, 1024
cfg = PRESETS["traj_ball"].for_trajectory(100)
x = torch.from_numpy(synthetic.trajectory(1, n, 3, seed=0)[0]).cuda()
ctx = _native.Context(0, n, W, H, 1)
style = cfg.style()
pos4, attr4 = ctx.standardize(x, style)
for fi in (0, 50, 99):
    ctx.counters()
    ctx.render(pos4, attr4, cfg.camera(fi, 100, W, H), style)
    print(fi, ctx.counters())
