#!/usr/bin/env python
"""Host-side timeline of the asynchronous host-buffer entry on workload H: how long each submit and each wait takes with
K calls in flight (diagnostics for bench.py's e2e number)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from pointcloud_render_b200 import _native

def main():
    depth = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    mean = sys.argv[2] if len(sys.argv) > 2 else "auto"
    spec = bench.workload_spec("H", 0)
    B, W, H, n = spec["frames_per_step"], spec["width"], spec["height"], spec["points"]
    F = int(sys.argv[3]) if len(sys.argv) > 3 else B          # frames per call
    ring = 2 * F
    host = torch.from_numpy(bench.make_ring(spec, ring, seed=0)).pin_memory()
    cams, cfg = bench.cameras_for(spec, ring)
    style = cfg.style(mean_mode=_native.MEAN_F64 if mean == "f64" else _native.MEAN_AUTO)
    ctx = _native.Context(device=0, max_points=n, max_w=W, max_h=H, max_batch=B)
    outs = [torch.empty((F, H, W, 4), dtype=torch.uint8).pin_memory() for _ in range(depth + 1)]
    inflight, log = [], []
    def step(s):
        k = (s % 2) * F
        t0 = time.perf_counter()
        if len(inflight) >= depth:
            ctx.host_wait(inflight.pop(0))
        t1 = time.perf_counter()
        inflight.append(ctx.render_frames_host_submit(host[k:k + F], cams[k:k + F], style, out_rgba=outs[s % (depth + 1)])[0])
        t2 = time.perf_counter()
        log.append((t1 - t0, t2 - t1))
    for s in range(4):
        step(s)
    ctx.host_wait(-1); inflight.clear(); log.clear()
    t0 = time.perf_counter()
    N = max(4, 512 // F)
    for s in range(N):
        step(s)
    ctx.host_wait(-1)
    dt = time.perf_counter() - t0
    print(f"depth {depth} mean {mean}: {F} frames/call, {N * F / dt:.0f} frames/s, {dt / N * 1e3:.2f} ms/call; wait ms {np.mean([a for a, _ in log]) * 1e3:.2f}, submit ms {np.mean([b for _, b in log]) * 1e3:.2f} (max {max(b for _, b in log) * 1e3:.2f})")
    ctx.close()

if __name__ == "__main__":
    main()
