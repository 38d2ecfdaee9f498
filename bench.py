#!/usr/bin/env python
"""bench.py — frames/s of the pcr hot path on BASELINE.json's headline workload.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload H|C2|C3|C4]

Workload H (default): 1 M-point trajectory frames at 1024x1024, `traj_ball` camera schedule.
A *step* is one pass of the whole hot path (K0 stats, K1 standardise/transform/colour, K2 project +
tile binning, K3 sphere raster, K4 shade) over one batch of `frames_per_step` frames per GPU.

  value   frames/s over all ranks, inputs resident in HBM (a ring of frames larger than L2)
  e2e     the same through pcr_render_frames_host: pinned HOST trajectory in, HOST images out,
          H2D/D2H copies inside the timed region (overlapped with the kernels by the library)
  roofline  HBM roofline of the dominant kernel (per-kernel CUDA events recorded by the library on
          its launching stream, over the timed region)
  cpu_baseline  the CPU oracle (numpy standardise/transform + OpenMP C ray caster + shading) on a
          bounded sample of the same frames, all host cores.  The reference's real renderer is
          Mitsuba (absent): the oracle casts 1 ray/pixel where the reference traces 128-256
          multi-bounce paths, so it is a strict lower bound on the reference's CPU time.

Multi-GPU (torchrun, one rank per GPU): trajectory frames shard across ranks with no collective
(weak scaling: every rank renders its own frames_per_step frames per step).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="H", choices=["H", "C2", "C3", "C4", "C5", "C2D", "C4D"])
    ap.add_argument("--points", type=int, default=0, help="override the workload's point count (C5 scaling studies)")
    ap.add_argument("--frames-per-step", type=int, default=0, help="frames per step per GPU (0 = workload default)")
    ap.add_argument("--max-batch", type=int, default=64, help="frames per internal launch (pcr_create max_batch, <= 64)")
    ap.add_argument("--ring", type=int, default=0, help="resident frames per GPU (0 = enough to exceed L2)")
    ap.add_argument("--trails", action="store_true", help="also draw the reference's velocity trails (6-column workloads C3/C4)")
    ap.add_argument("--merge", default="fused", choices=["fused", "nccl"],
                    help="C5 on several GPUs: z-merge fused into the raster over peer memory (default) or NCCL min all-reduce")
    ap.add_argument("--mean", default="auto", choices=["auto", "f64"],
                    help="standardisation centre: auto = the reference's sequential np.mean (default), f64 = parallel float64 sums")
    ap.add_argument("--lookahead", type=int, default=4, help="steps of K0 look-ahead (the serial mean's latency is ~3 steps at H)")
    ap.add_argument("--no-lookahead", action="store_true", help="do not hint the next steps' frames (pcr_prefetch_frames)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-profile", action="store_true", help="do not bracket kernels with events in the timed region")
    ap.add_argument("--trace-steps", action="store_true", help="diagnostics: an event after every step of the headline region, per-step ms in the line")
    ap.add_argument("--profile-one-step", action="store_true",
                    help="for ncu --profile-from-start off: warm up, then ONE device-resident step between cudaProfilerStart/Stop, no timing")
    return ap.parse_args()


def workload_spec(name, frames_per_step):
    from pointcloud_render_b200 import synthetic
    c = dict(synthetic.CONFIGS[name])
    default_fps = {"H": 64, "C4": 64, "C3": 64, "C2": 64, "C5": 1, "C2D": 16, "C4D": 8}[name]
    c["frames_per_step"] = frames_per_step or default_fps
    b_in = 4 * c["cols"] + (4 if c["radii"] else 0)
    # input read once, u64 visibility written once, RGBA8 written once
    c["algorithmic_bytes_per_frame"] = c["points"] * b_in + c["width"] * c["height"] * (8 + 4)
    c["input_bytes_per_frame"] = c["points"] * 4 * c["cols"]
    return c


SCHEDULE_STRIDE = 37          # coprime with the schedule lengths: consecutive ring slots jump across the camera schedule


def config_dict(name, spec, ring, n_gpus, trails=False, mean="auto"):
    """The SAME dict from both arms (--impl ours / reference): what is rendered, not how."""
    return {"workload": f"{name}: {spec['points']} points/frame, {spec['width']}x{spec['height']}, preset {spec['preset']}, "
                        f"{spec['cols']} cols f32, colour mode {spec['color_mode']}" + (", per-point radius" if spec["radii"] else ""),
            "points": spec["points"], "width": spec["width"], "height": spec["height"],
            "frames_per_step_per_gpu": spec["frames_per_step"], "resident_ring_frames": ring,
            "camera_schedule": f"ring slot i = trajectory frame (and camera index) ({SCHEDULE_STRIDE}*i) mod {spec['frames']} of the "
                               f"{spec['frames']}-frame schedule: every step mixes far, middle and nearest cameras, all indices are visited",
            "mean_mode": "sequential float32 np.mean of the reference (PCR_MEAN_AUTO)" if mean == "auto" else "parallel float64 sums (PCR_MEAN_F64)",
            "trails": bool(trails),
            "l2_policy": "inputs larger than L2: each step reads a different slice of a resident frame ring "
                         f"({ring * spec['input_bytes_per_frame'] / 1e6:.0f} MB) and rewrites {spec['frames_per_step'] * spec['width'] * spec['height'] * 12 / 1e6:.0f} MB of outputs",
            "parallelism": "single GPU" if n_gpus == 1 else f"frames sharded over {n_gpus} GPUs, no collective"}


def ring_frames(spec, requested=0):
    """Resident frames per GPU: a multiple of frames_per_step, at least 5 steps (so that the whole camera schedule is
    visited and a four-step look-ahead never meets the slice being rendered) whose inputs exceed the 126 MB L2."""
    B = spec["frames_per_step"]
    ring = requested or max(5 * B, B * int(np.ceil(160e6 / spec["input_bytes_per_frame"] / B)))
    return (ring + B - 1) // B * B


def ring_indices(spec, ring, rank=0):
    """Trajectory frame (= camera index, as in the reference's main loop: frame f is rendered with camera f) of every
    ring slot: a stride permutation of the schedule, so every step of 32 consecutive slots spans the whole schedule —
    the far, cheap cameras and the nearest, most expensive ones alike."""
    return [(SCHEDULE_STRIDE * (i + 1000 * rank)) % spec["frames"] for i in range(ring)]


def make_ring(spec, ring, seed, rank=0):
    """`ring` frames of the workload's trajectory, slot i = physics frame ring_indices()[i]."""
    from pointcloud_render_b200 import synthetic
    return synthetic.trajectory(ring, spec["points"], spec["cols"], "gauss", seed=seed, frame_indices=ring_indices(spec, ring, rank))


def cameras_for(spec, ring, rank=0):
    from pointcloud_render_b200.presets import PRESETS
    total = spec["frames"]
    cfg = PRESETS[spec["preset"]].for_trajectory(total)
    return [cfg.camera(f, total, spec["width"], spec["height"]) for f in ring_indices(spec, ring, rank)], cfg


# --------------------------------------------------------------------------------------------------
# host placement
def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if part:
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def smi_topology():
    """`nvidia-smi topo -m` (text) and, per GPU index, the CPU / NUMA affinity columns it prints — the only place the
    GPU's socket shows up when the container hides /sys/bus/pci/devices/*/numa_node (it reads -1 there)."""
    import re
    try:
        text = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
    except (OSError, subprocess.TimeoutExpired):
        return "", {}
    text = re.sub(r"\x1b\[[0-9;]*m", "", text)
    aff = {}
    for line in text.splitlines():
        m = re.match(r"^GPU(\d+)\s", line)
        if not m:
            continue
        toks = line.split()
        lists = [t for t in toks[1:] if re.fullmatch(r"\d+(-\d+)?(,\d+(-\d+)?)*", t)]
        # after the link matrix: CPU affinity (a range list), NUMA affinity, GPU NUMA id
        cpu = next((t for t in lists if "-" in t or "," in t), None)
        rest = lists[lists.index(cpu) + 1:] if cpu in lists else []
        aff[int(m.group(1))] = {"cpu_affinity": cpu, "numa_affinity": rest[0] if rest else None}
    return text, aff


def bind_to_gpu_numa(local_rank):
    """Run this rank (and first-touch its pinned staging memory: cudaHostAlloc pages are touched by the calling thread)
    on the NUMA node its GPU hangs off.  The node comes from sysfs, or — when the container reports -1 there — from the
    CPU-affinity column of `nvidia-smi topo -m`.  Pure placement: returns a small dict for the JSON line."""
    info = {"bound": False}
    try:
        import torch
        p = torch.cuda.get_device_properties(local_rank)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = -1
        try:
            node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        except OSError:
            pass
        info.update(pci=bdf, numa_node=node)
        cpus, source = set(), None
        if node >= 0:
            cpus, source = _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read()), "sysfs"
        else:
            _, aff = smi_topology()
            # torch's device index follows CUDA_VISIBLE_DEVICES; nvidia-smi lists physical GPUs: match by PCI bus id
            idx = local_rank
            try:
                q = subprocess.run(["nvidia-smi", "--query-gpu=index,pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True, timeout=20).stdout
                for line in q.splitlines():
                    i, bus = [x.strip() for x in line.split(",")]
                    if bus.lower().endswith(bdf[5:].lower()):
                        idx = int(i)
            except (OSError, ValueError, subprocess.TimeoutExpired):
                pass
            a = aff.get(idx)
            if a and a["cpu_affinity"]:
                cpus, source = _parse_cpulist(a["cpu_affinity"]), "nvidia-smi topo -m"
                info.update(smi_numa_affinity=a["numa_affinity"], smi_cpu_affinity=a["cpu_affinity"])
        # memory first: prefer the GPU's node for everything this process allocates from here on (the pinned trajectory
        # ring and image buffers) — works even when the container's CPU set is not on that node
        mem_node = node if node >= 0 else (int(info["smi_numa_affinity"]) if str(info.get("smi_numa_affinity", "")).isdigit() else -1)
        if mem_node >= 0:
            import ctypes
            libc = ctypes.CDLL(None, use_errno=True)
            mask = (ctypes.c_ulong * 16)()
            mask[mem_node // 64] = 1 << (mem_node % 64)
            MPOL_PREFERRED, SYS_set_mempolicy = 1, 238               # x86_64
            rc = libc.syscall(SYS_set_mempolicy, MPOL_PREFERRED, mask, 16 * 64)
            info.update(mempolicy=f"preferred node {mem_node}" if rc == 0 else f"set_mempolicy failed (errno {ctypes.get_errno()})")
        allowed = cpus & os.sched_getaffinity(0)
        if allowed and len(allowed) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, allowed)
            info.update(bound=True, cpus=len(allowed), source=source)
        elif allowed:
            info.update(cpus=len(allowed), source=source, note="the GPU's CPU set is every CPU this process may use: nothing to bind")
    except Exception as e:          # placement is best effort
        info["note"] = f"{type(e).__name__}: {e}"[:160]
    return info


# --------------------------------------------------------------------------------------------------
# clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.samples, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        self.mark0 = self.mark1 = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def begin(self):
        self.mark0 = time.perf_counter()

    def end(self):
        self.mark1 = time.perf_counter()

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        if not self.proc or self.mark0 is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "nvidia-smi not available"}
        rows = [s for t, s in self.samples if self.mark0 - 0.05 <= t <= (self.mark1 or t) + 0.05] or [s for _, s in self.samples[-3:]]
        sm, mx, pw, reasons = [], [], [], set()
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "no samples"}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU side (oracle port) — used for cpu_baseline and for --impl reference only
def cpu_frame(orc, frame_pts, cfg, cam_index, spec, radius):
    """One frame of the path on the CPU: numpy standardise + transform + colour hook (reference
    arithmetic), OpenMP C visibility + shading.  Returns seconds."""
    t0 = time.perf_counter()
    p = orc.transform_coordinates(orc.standardize_point_cloud(frame_pts), cfg.flip_x)
    attr4 = orc.compute_color(p, mode=spec["color_mode"], const_rgb=cfg.const_rgb, vel_norm=cfg.vel_norm)
    r = radius if radius is not None else np.full(len(p), cfg.radius, np.float32)
    pos4 = np.concatenate([p[:, :3], r[:, None]], axis=1)
    fr = orc.camera_frame(cfg.camera_position(cam_index, spec["frames"]), cfg.target, cfg.up, cfg.fov, cfg.near_clip, cfg.far_clip,
                          spec["width"], spec["height"])
    sc = orc.make_scene(True, cfg.floor_z, cfg.floor_min, cfg.floor_max, cfg.floor_albedo, cfg.light_z, cfg.light_half,
                        cfg.radiance, cfg.bounce)
    vis = orc.visibility(pos4, fr, sc)
    orc.shade(vis, pos4, attr4, fr, sc)
    return time.perf_counter() - t0


def xml_emit_seconds_per_point(n=20000):
    """Cost of the reference's per-point XML string loop (example_renderer.py:113-128), restated:
    reported beside the baseline, NOT included in it."""
    seg = ('<shape type="sphere"><float name="radius" value="0.01"/><transform name="toWorld"><translate x="{}" y="{}" z="{}"/>'
           '</transform><bsdf type="diffuse"><rgb name="reflectance" value="{},{},{}"/></bsdf></shape>')
    pcl = np.random.default_rng(0).standard_normal((n, 3)).astype(np.float32)
    lo, rng = pcl.min(0), pcl.max(0) - pcl.min(0)
    t0 = time.perf_counter()
    out = []
    for idx, point in enumerate(pcl):
        q = (point - lo) / (rng + 1e-8)
        color = np.array([0.3, 0.3, 0.3]) if q[0] > -1 else None
        out.append(seg.format(point[0], point[1], point[2], *color))
    "".join(out)
    return (time.perf_counter() - t0) / n


def cpu_baseline(spec, ring_host, radius, budget_s=12.0, max_frames=40):
    from oracle import pcr_oracle as orc
    from pointcloud_render_b200.presets import PRESETS
    cfg = PRESETS[spec["preset"]].for_trajectory(spec["frames"])
    idx = ring_indices(spec, len(ring_host))
    orc.set_num_threads()                                       # all host cores, whatever OMP_NUM_THREADS says
    cpu_frame(orc, ring_host[0], cfg, idx[0], spec, radius)     # warm-up (page in, OpenMP pool)
    total, frames = 0.0, 0
    while frames < max_frames and total < budget_s:
        total += cpu_frame(orc, ring_host[frames % len(ring_host)], cfg, idx[frames % len(ring_host)], spec, radius)
        frames += 1
    per_pt = xml_emit_seconds_per_point()
    return {"value": frames / total, "unit": "frames/s", "cores": orc.num_threads(), "kind": "port",
            "sample": f"{frames} frames of the same workload ({total:.1f} s): numpy standardise+transform, OpenMP C ray caster 1 ray/pixel + shading",
            "host_cpus": os.cpu_count(),
            "note": "lower bound on the reference: Mitsuba (absent) traces 128-256 spp multi-bounce paths; the per-point XML emit loop "
                    f"of the reference (example_renderer.py:113-128) would add ~{per_pt * spec['points']:.1f} s/frame at this size "
                    f"({per_pt * 1e6:.1f} us/point measured on 20000 points) and is NOT included"}


def run_reference(args, spec, rank, world):
    """--impl reference: the CPU path, all host threads, each step a bounded sample of the workload."""
    if rank != 0:
        return
    from oracle import pcr_oracle as orc
    from pointcloud_render_b200.presets import PRESETS
    orc.build()
    orc.set_num_threads()                                       # torchrun exports OMP_NUM_THREADS=1: use every host core anyway
    cfg = PRESETS[spec["preset"]].for_trajectory(spec["frames"])
    frames_per_step = 2 if spec["points"] >= 500_000 else 4
    ring = 16                                                    # the first 16 slots of the bench ring: same frames, same cameras
    host = make_ring(spec, ring, seed=0)
    idx = ring_indices(spec, ring)
    from pointcloud_render_b200 import synthetic
    radius = synthetic.radii(spec["points"]) if spec["radii"] else None
    k = 0
    for _ in range(args.warmup):
        for _ in range(frames_per_step):
            cpu_frame(orc, host[k % ring], cfg, idx[k % ring], spec, radius); k += 1
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for _ in range(frames_per_step):
            cpu_frame(orc, host[k % ring], cfg, idx[k % ring], spec, radius); k += 1
    dt = time.perf_counter() - t0
    fps = args.steps * frames_per_step / dt
    sample = (f"each step = {frames_per_step} frames of the same workload (a bounded sample of the {spec['frames_per_step']}-frame step; the first "
              f"{ring} slots of the same ring, same cameras), oracle port (numpy + OpenMP C ray caster, 1 ray/pixel)")
    line = {"impl": "reference", "metric": "frames_per_s", "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "mpoints_per_s": fps * spec["points"] / 1e6,
            "config": config_dict(args.workload, spec, ring_frames(spec, args.ring), args.gpus, trails=bool(args.trails and spec["cols"] == 6), mean=args.mean),
            "reference_frames_per_step": frames_per_step,
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": orc.num_threads(), "kind": "port", "sample": sample,
                             "host_cpus": os.cpu_count()},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
def run_point_sharded(args, spec, rank, world, local_rank):
    """C5: ONE huge cloud, point-sharded over the ranks (strong scaling): K0 partials -> C0 all-gather
    -> K1 -> K2/K3 into a full-frame z-buffer per rank -> C1 int64-min all-reduce over NVLink ->
    owner-only K4 -> byte-MAX all-reduce of the RGBA8 image.  A step = one frame of the cloud."""
    import torch
    import torch.distributed as dist
    from pointcloud_render_b200 import _native, sharding, synthetic
    from pointcloud_render_b200.presets import PRESETS
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n, W, H = spec["points"], spec["width"], spec["height"]
    a, b = sharding.point_shard(n, rank, world)
    rng = np.random.default_rng(1000 + rank)
    local = torch.from_numpy(rng.standard_normal((b - a, 3)).astype(np.float32)).cuda()   # a Gaussian cloud, shard by shard
    cfg = PRESETS[spec["preset"]]
    # one cloud split over the ranks: the point-sharded entries take the parallel float64 mean (a float32 fold cannot be
    # split across shards); the single-GPU run uses the same centre so that 1 vs N GPUs render the same scene
    style, cam = cfg.style(color_mode=spec["color_mode"], mean_mode=_native.MEAN_F64), cfg.camera(0, 1, W, H)
    ctx = _native.Context(device=local_rank, max_points=b - a, max_w=W, max_h=H, max_batch=1)

    bufs = sharding.point_sharded_buffers(b - a, cam, local.device)          # allocated once, reused every frame
    frames1 = local.unsqueeze(0)
    fused = world > 1 and args.merge == "fused"
    mesh = sharding.PeerMesh(ctx, cam) if fused else None

    def step():
        if fused:
            return sharding.render_point_sharded_fused(ctx, mesh, local, a, n, cam, style, buffers=bufs)
        if world > 1:
            return sharding.render_point_sharded(ctx, local, a, n, cam, style, buffers=bufs)
        # one GPU: the whole-path entry (K1 fused into K2a / K4, nothing materialised)
        rgba1, vis1 = ctx.render_frames(frames1, [cam], style, out_rgba=bufs["rgba"].unsqueeze(0), out_vis=bufs["vis"].unsqueeze(0))
        return vis1[0], rgba1[0]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(args.warmup):
        step()
    barrier()
    ctx.profile_read()
    ctx.profile(True)
    launches0 = ctx.counters()["launches"]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    if sampler:
        sampler.begin()
    barrier()
    ev[0].record()
    for _ in range(args.steps):
        vis, rgba = step()
    ev[1].record()
    barrier()
    if sampler:
        sampler.end()
        sampler.stop()
    ms = ev[0].elapsed_time(ev[1])
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ctx.profile(False)
    prof = ctx.profile_read()
    counters = ctx.counters()
    if rank == 0:
        fps = args.steps / (ms * 1e-3)
        kernels = {k: {"ms_total": round(v[0], 4), "launches": int(v[1]), "us_per_launch": round(v[0] / v[1] * 1e3, 3)}
                   for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}
        kernel_ms = sum(v[0] for v in prof.values())
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peak = float(json.load(open(peaks_path))["hbm_gbs"]) if os.path.isfile(peaks_path) else 6650.0
        per_gpu_bytes = (b - a) * 12 + W * H * 8 + W * H * 4 // world       # SURVEY.md 8(d), point-sharded definition
        top = max(prof.items(), key=lambda kv: kv[1][0])
        line = {"metric": "frames_per_s", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "mpoints_per_s": fps * n / 1e6,
                "config": {"workload": f"C5: one {n}-point Gaussian cloud at {W}x{H}, preset {spec['preset']}, point-sharded over {world} GPU(s)",
                           "points": n, "width": W, "height": H, "points_per_gpu": b - a,
                           "mean_mode": "parallel float64 sums (PCR_MEAN_F64: the point-sharded path)",
                           "l2_policy": f"inputs larger than L2 ({(b - a) * 12 / 1e6:.0f} MB of points + {W * H * 8 / 1e6:.0f} MB z-buffer per GPU)",
                           "parallelism": "single GPU" if world == 1 else (
                               f"points sharded over {world} GPUs; z-merge fused into the raster: winners pushed into the row owner's z-buffer with "
                               "64-bit atomicMin over NVLink (CUDA IPC peer memory), image pixels stored straight into rank 0's buffer; "
                               "collectives left: 72 B/rank stats all-gather + two 1-element all-reduces as barriers" if fused else
                               f"points sharded over {world} GPUs; C0 all-gather (72 B/rank) + C1 "
                               f"ncclAllReduce(int64,min) of {W * H * 8 / 1e6:.0f} MB + byte-MAX all-reduce of {W * H * 4 / 1e6:.0f} MB RGBA8")},
                "e2e": None, "gpu_launches": int(counters["launches"] - launches0), "kernels": kernels,
                "kernel_ms_per_step": kernel_ms / args.steps, "collective_and_gap_ms_per_step": ms / args.steps - kernel_ms / args.steps,
                "roofline": {"bound": "hbm", "kernel": top[0], "achieved": per_gpu_bytes / (top[1][0] / top[1][1] * 1e-3) / 1e9, "peak": peak,
                             "unit": "GB/s", "frac": per_gpu_bytes / (top[1][0] / top[1][1] * 1e-3) / 1e9 / peak, "traffic": None,
                             "algorithmic_bytes_per_launch": per_gpu_bytes,
                             "note": "per-GPU algorithmic bytes = shard points*12 + W*H*8 (own z-buffer) + W*H*4/world (image slice)"},
                "clocks": sampler.summary() if sampler else None, "pairs_last_frame": counters["pairs_last_frame"],
                "overflow_frames": counters["overflow_frames"],
                "sphere_pixels": int((_native.keys_to_ids(ctx.peer_buffers()[0] if fused else vis) < n).sum()) if not fused else None,
                "merge": ("fused-peer-memory" if fused else "nccl-allreduce") if world > 1 else None}
        print(json.dumps(line), flush=True)
    if mesh:
        mesh.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def moving_trajectory(frames, n, seed, dt=0.01):
    """(frames, n, 6) ballistic trajectory whose velocities are independent of the positions (synthetic.trajectory
    draws V = 3 P0, a pure expansion that the per-frame standardisation removes: no history trail would show)."""
    rng = np.random.default_rng(seed)
    p0, v, g = rng.standard_normal((n, 3)), 3.0 * rng.standard_normal((n, 3)), np.array([0.0, -1.0, 0.0])
    out = np.empty((frames, n, 6), np.float32)
    for f in range(frames):
        t = f * dt
        out[f, :, :3] = p0 + t * v + 0.5 * t * t * g
        out[f, :, 3:] = v + t * g
    return out


def run_droplets(args, spec, rank, world, local_rank):
    """C2D / C4D: the droplet scene of traj_renderer.py (history trails) / traj_vel_renderer.py (velocity trails)
    through pcr_render_droplet_frames.  A step renders frames_per_step frames that have a full 20-frame history."""
    import torch
    from oracle import droplet_oracle as do, pcr_oracle as orc
    from pointcloud_render_b200 import _native, droplets
    from pointcloud_render_b200.presets import PRESETS
    torch.cuda.set_device(local_rank)
    B, n, W, H = spec["frames_per_step"], spec["points"], spec["width"], spec["height"]
    trails = spec["droplet_trails"]
    cfg = PRESETS[spec["preset"]].for_trajectory(spec["frames"])
    halo = 20 if trails == 2 else 0
    host_np = moving_trajectory(halo + 2 * B, n, seed=rank)
    host = torch.from_numpy(host_np).pin_memory()
    resident = host.cuda()
    ctx = _native.Context(device=local_rank, max_points=n, max_w=W, max_h=H, max_batch=min(B, args.max_batch))
    ctx.set_droplet_mesh(droplets.droplet_vertices(), droplets.N_RINGS, droplets.N_SEGMENTS)
    style = cfg.style(color_mode=0, trails=trails if trails == 2 else True)
    first = spec["frames"] // 2
    cams = [cfg.camera(first + k, spec["frames"], W, H) for k in range(2 * B)]
    rgba = torch.empty((B, H, W, 4), dtype=torch.uint8, device="cuda")
    host_rgba = torch.empty((B, H, W, 4), dtype=torch.uint8).pin_memory()

    def step_device(s):
        k = (s % 2) * B
        ctx.render_droplet_frames(resident[k:k + halo + B], cams[k:k + B], style, n_history=halo, out_rgba=rgba)

    def step_host(s):
        k = (s % 2) * B
        dev = host[k:k + halo + B].cuda(non_blocking=True)
        ctx.render_droplet_frames(dev, cams[k:k + B], style, n_history=halo, out_rgba=rgba)
        host_rgba.copy_(rgba, non_blocking=True)
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    for s in range(args.warmup):
        step_device(s)
    torch.cuda.synchronize()
    ctx.profile_read()
    ctx.profile(True)
    launches0 = ctx.counters()["launches"]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.begin()
    ev0.record()
    for s in range(args.steps):
        step_device(args.warmup + s)
    ev1.record()
    torch.cuda.synchronize()
    dev_ms = ev0.elapsed_time(ev1)
    ctx.profile(False)
    prof = ctx.profile_read()
    launches = ctx.counters()["launches"] - launches0
    fps = B * args.steps / (dev_ms * 1e-3)
    for s in range(2):
        step_host(s)
    t0 = time.perf_counter()
    for s in range(args.steps):
        step_host(s)
    e2e_s = time.perf_counter() - t0
    sampler.end()
    sampler.stop()
    total_ms = sum(v[0] for v in prof.values())
    kernels = {k: {"ms_total": round(v[0], 4), "launches": int(v[1]), "us_per_launch": round(v[0] / v[1] * 1e3, 3), "share": round(v[0] / total_ms, 4)}
               for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}
    top = max(prof.items(), key=lambda kv: kv[1][0])
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks_path))["hbm_gbs"]) if os.path.isfile(peaks_path) else 6650.0
    frames_per_launch = min(B, ctx.max_batch)
    bytes_per_launch = spec["algorithmic_bytes_per_frame"] * frames_per_launch
    achieved = bytes_per_launch / (top[1][0] / top[1][1] * 1e-3) / 1e9
    # CPU side: the oracle's droplet scene for a few frames (numpy geometry + OpenMP C mesh / polyline caster + shading)
    orc.set_num_threads()
    t_cpu, frames_cpu = 0.0, 0
    sc = orc.make_scene(True, cfg.floor_z, cfg.floor_min, cfg.floor_max, cfg.floor_albedo, cfg.light_z, cfg.light_half, cfg.radiance, cfg.bounce)
    while frames_cpu < 6 and t_cpu < 20.0 and not args.no_cpu_baseline:
        f = halo + frames_cpu
        t0 = time.perf_counter()
        std = [orc.transform_coordinates(orc.standardize_point_cloud(host_np[k]), cfg.flip_x) for k in range(f - halo, f + 1)]
        pcl = std[-1]
        xf = do.to_world_f32(do.rotation_from_velocity(pcl[:, 3:6]), pcl[:, :3])
        if trails == 2:
            ctrl, count = do.history_trails(np.stack([x[:, :3] for x in std[:-1]]), pcl[:, :3])
        else:
            tail, head, valid = orc.velocity_trails(pcl, cfg.trail_length_scale(first + frames_cpu))
            ctrl = np.zeros((n, 21, 3), np.float32)
            ctrl[:, 0], ctrl[:, 1] = tail, head
            count = np.where(valid, 2, 0).astype(np.int32)
        fr = orc.camera_frame(cfg.camera_position(first + frames_cpu, spec["frames"]), cfg.target, cfg.up, cfg.fov, cfg.near_clip, cfg.far_clip, W, H)
        vis = do.add_droplets(do.add_polylines(orc.visibility(np.zeros((0, 4), np.float32), fr, sc), ctrl, count, fr, n, radius=cfg.trail_radius), xf, fr)
        do.shade_droplet_scene(vis, xf, ctrl, count, fr, sc)
        t_cpu += time.perf_counter() - t0
        frames_cpu += 1
    line = {"metric": "frames_per_s", "value": fps, "unit": "frames/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "mpoints_per_s": fps * n / 1e6,
            "config": {"workload": f"{args.workload}: droplet scene, {n} points/frame (6 cols f32), {W}x{H}, preset {spec['preset']}, "
                                   + ("Catmull-Rom history trails over a 20-frame halo" if trails == 2 else "straight velocity trails"),
                       "points": n, "width": W, "height": H, "frames_per_step_per_gpu": B,
                       "l2_policy": f"outputs larger than L2: every step rewrites {B * W * H * 12 / 1e6:.0f} MB of keys + images; the inputs "
                                    f"({(halo + B) * n * 24 / 1e6:.1f} MB per step) are far smaller than L2 and alternate between two windows",
                       "parallelism": "single GPU"},
            "e2e": {"value": B * args.steps / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": int((halo + B) * n * 24),
                    "d2h_bytes_per_step": int(B * W * H * 4), "ms_per_step": e2e_s / args.steps * 1e3,
                    "api": "pinned host trajectory window -> device, pcr_render_droplet_frames, RGBA8 -> pinned host"},
            "gpu_launches": int(launches), "kernels": kernels,
            "roofline": {"bound": "hbm", "kernel": top[0], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "algorithmic_bytes_per_launch": bytes_per_launch, "frames_per_launch": frames_per_launch,
                         "us_per_launch": top[1][0] / top[1][1] * 1e3,
                         "note": "algorithmic bytes = N*24 + W*H*(8+4) per frame; the droplet raster is bound by ray-triangle tests, not HBM"},
            "clocks": sampler.summary()}
    if frames_cpu:
        line["cpu_baseline"] = {"value": frames_cpu / t_cpu, "unit": "frames/s", "cores": orc.num_threads(), "kind": "port",
                                "sample": f"{frames_cpu} frames of the same workload ({t_cpu:.1f} s): numpy standardise / rotations / Catmull-Rom "
                                          "(per-point python loop for the trail de-duplication), OpenMP C mesh + polyline caster, shading",
                                "host_cpus": os.cpu_count()}
    print(json.dumps(line), flush=True)
    ctx.close()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    spec = workload_spec(args.workload, args.frames_per_step)
    if args.points:
        spec["points"] = args.points
    if args.workload in ("C2D", "C4D"):
        if args.impl == "reference" or world > 1:
            raise SystemExit("the droplet workloads are single-GPU device arms (--impl ours, --gpus 1)")
        run_droplets(args, spec, rank, world, local_rank)
        return
    if args.workload == "C5" and args.impl == "ours":
        run_point_sharded(args, spec, rank, world, local_rank)
        return

    if args.impl == "reference":
        run_reference(args, spec, rank, world)
        return

    import torch
    import torch.distributed as dist
    from pointcloud_render_b200 import _native, synthetic
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa(local_rank)                  # before the pinned trajectory is allocated
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    B = spec["frames_per_step"]
    ring = ring_frames(spec, args.ring)
    W, H, n = spec["width"], spec["height"], spec["points"]

    host_np = make_ring(spec, ring, seed=rank, rank=rank)          # every rank renders its own frames
    host = torch.from_numpy(host_np).pin_memory()
    resident = host.cuda(non_blocking=True)
    radius_np = synthetic.radii(n) if spec["radii"] else None
    radius = torch.from_numpy(radius_np).cuda() if radius_np is not None else None
    ctx = _native.Context(device=local_rank, max_points=n, max_w=W, max_h=H, max_batch=min(B, args.max_batch))
    cams_all, cfg = cameras_for(spec, ring, rank)
    mean_mode = _native.MEAN_F64 if args.mean == "f64" else _native.MEAN_AUTO
    style = cfg.style(color_mode=spec["color_mode"], trails=args.trails and spec["cols"] == 6, mean_mode=mean_mode)
    rgba = torch.empty((B, H, W, 4), dtype=torch.uint8, device="cuda")
    host_rgba = [torch.empty((B, H, W, 4), dtype=torch.uint8).pin_memory() for _ in range(2)]
    slots = ring // B
    # hinted steps ahead: bounded by the ring and by the library's prepared-batch slots (6, one kept free; a step is
    # ceil(B / max_batch) batches)
    batches_per_step = -(-B // ctx.max_batch)
    lookahead = 0 if args.no_lookahead else max(1, min(args.lookahead, slots - 1, 5 // batches_per_step))

    hinted = set()

    def hint(s):
        ka = (s % slots) * B
        ctx.prefetch_frames(resident[ka:ka + B], style)
        hinted.add(s)

    def step_device(s, cams=None):
        # K0 of step s + lookahead (statistics incl. the serial reference-exact mean, pre-pass sample) is started now
        # on side streams — the hint pcr_prefetch_frames; every step issues exactly one, so K steps = K x K0 + K renders
        k = (s % slots) * B
        if lookahead:
            hint(s + lookahead)
        hinted.discard(s)
        ctx.render_frames(resident[k:k + B], cams if cams is not None else cams_all[k:k + B], style, radius=radius, out_rgba=rgba)

    def prime(first):
        # steady state at the start of a timed region whatever --warmup is: the hints that the `lookahead` steps BEFORE the
        # region would have issued (steps first .. first + lookahead - 1), if the warm-up was too short to have issued them
        for s in range(first, first + lookahead):
            if s not in hinted:
                hint(s)

    inflight = []

    def step_host(s):
        # the asynchronous form of the host-buffer entry, two calls in flight (the output buffers alternate): the end of
        # call s (its last kernels, the serial mean's latency, the last D2H copy) overlaps the input copy of call s + 1
        k = (s % slots) * B
        if len(inflight) >= 2:
            ctx.host_wait(inflight.pop(0))
        inflight.append(ctx.render_frames_host_submit(host[k:k + B], cams_all[k:k + B], style, radius_host=radius_np, out_rgba=host_rgba[s % 2])[0])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    trace = []

    def timed(n_steps, first, cams_of=None):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        prime(first)
        barrier()
        ev0.record()
        marks = []
        for s in range(n_steps):
            step_device(first + s, cams_of(first + s) if cams_of else None)
            if args.trace_steps and cams_of is None:
                marks.append(torch.cuda.Event(enable_timing=True)); marks[-1].record()
        ev1.record()
        barrier()
        if marks:
            trace.append([round(a.elapsed_time(b), 3) for a, b in zip([ev0] + marks[:-1], marks)])
        return max_over_ranks(ev0.elapsed_time(ev1))

    sampler = ClockSampler(local_rank) if rank == 0 else None

    # ---------------- device-resident throughput ----------------
    for s in range(args.warmup):
        step_device(s)
    barrier()
    if args.profile_one_step:
        prime(args.warmup)                       # the step's own K0 was hinted earlier, like in the steady state
        barrier()
        torch.cuda.profiler.start()
        step_device(args.warmup)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({"profiled": "one device-resident step", "workload": args.workload, "frames": B}), flush=True)
        ctx.close()
        return
    # the headline region runs WITHOUT the per-kernel event brackets (two event records around every launch keep launches
    # from queueing back to back); the same steps are then repeated with the brackets on for the per-kernel table
    ctx.profile_read()
    launches0 = ctx.counters()["launches"]
    if sampler:
        sampler.begin()
    dev_ms = timed(args.steps, args.warmup)
    counters = ctx.counters()
    launches = counters["launches"] - launches0
    frames_total = world * B * args.steps
    fps = frames_total / (dev_ms * 1e-3)
    prof, prof_ms = {}, None
    if not args.no_profile:
        ctx.profile(True)
        prof_ms = timed(args.steps, args.warmup + args.steps)
        ctx.profile(False)
        prof = ctx.profile_read()

    # the same step with the cameras of one third of the schedule only (far / middle / nearest): what each costs
    thirds = None
    if spec["frames"] >= 6 and cfg.eye is None and world == 1:
        thirds = {}
        total = spec["frames"]
        bounds = {"far": (0, total // 3), "middle": (total // 3, 2 * total // 3), "nearest": (2 * total // 3, total)}
        k3 = max(4, args.steps // 2)
        for name, (lo, hi) in bounds.items():
            cams3 = [cfg.camera(lo + (j * 7) % (hi - lo), total, W, H) for j in range(B)]
            timed(lookahead + 1, 0, lambda s: cams3)                     # fills the look-ahead pipeline again
            thirds[name] = {"camera_indices": [lo, hi - 1], "frames_per_s": B * k3 / (timed(k3, lookahead + 1, lambda s: cams3) * 1e-3)}

    # ---------------- end to end through the host-buffer entry ----------------
    e2e = None
    if not args.no_e2e:
        for s in range(max(args.warmup, 1)):
            step_host(s)
        ctx.host_wait(-1); inflight.clear()
        barrier()
        t0 = time.perf_counter()
        for s in range(args.steps):
            step_host(args.warmup + s)
        ctx.host_wait(-1); inflight.clear()                  # every image of every step is in host memory
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        # context for the end-to-end number: what a bare pinned H2D copy of one step's input achieves on this box
        cp0, cp1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dst = torch.empty_like(resident[:B])
        dst.copy_(host[:B], non_blocking=True)
        torch.cuda.synchronize()
        cp0.record()
        for _ in range(3):
            dst.copy_(host[:B], non_blocking=True)
        cp1.record()
        torch.cuda.synchronize()
        h2d_gbs = 3 * B * spec["input_bytes_per_frame"] / (cp0.elapsed_time(cp1) * 1e-3) / 1e9
        # the host fabric under the traffic pattern of the pipeline: every rank copies a step's input H2D and a step's images
        # D2H at the same time, all ranks together (what the box's PCIe switches / host memory give to N GPUs at once)
        img_dev = torch.empty((B, H, W, 4), dtype=torch.uint8, device="cuda")
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        barrier()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        s_in.wait_event(b0)
        s_out.wait_event(b0)
        for _ in range(3):
            with torch.cuda.stream(s_in):
                dst.copy_(host[:B], non_blocking=True)
            with torch.cuda.stream(s_out):
                host_rgba[0].copy_(img_dev, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s_in)
        torch.cuda.current_stream().wait_stream(s_out)
        b1.record()
        torch.cuda.synchronize()
        bidir_s = b0.elapsed_time(b1) * 1e-3
        bidir = [3 * B * spec["input_bytes_per_frame"] / bidir_s / 1e9, 3 * B * W * H * 4 / bidir_s / 1e9]
        h2d_all, bidir_all = [h2d_gbs], [bidir]
        if world > 1:
            t = torch.tensor([h2d_gbs] + bidir, dtype=torch.float64, device="cuda")
            g = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(g, t)
            h2d_all = [float(x[0].item()) for x in g]
            bidir_all = [[float(x[1].item()), float(x[2].item())] for x in g]
        e2e = {"value": frames_total / e2e_s, "unit": "frames/s", "pcie_h2d_gbs_measured": h2d_gbs, "pcie_h2d_gbs_per_rank": h2d_all,
               "h2d_copy_bound_frames_per_s": sum(h2d_all) * 1e9 / spec["input_bytes_per_frame"],
               "host_fabric_probe": {"what": "all ranks at once: a step's input H2D and a step's images D2H concurrently, pinned memory, 3 repeats",
                                     "h2d_gbs_per_rank": [x[0] for x in bidir_all], "d2h_gbs_per_rank": [x[1] for x in bidir_all],
                                     "h2d_gbs_sum": sum(x[0] for x in bidir_all), "d2h_gbs_sum": sum(x[1] for x in bidir_all),
                                     "frames_per_s_bound": sum(x[0] for x in bidir_all) * 1e9 / spec["input_bytes_per_frame"]},
               "h2d_bytes_per_step": int(B * spec["input_bytes_per_frame"] + (n * 4 if spec["radii"] else 0)),
               "d2h_bytes_per_step": int(B * W * H * 4), "ms_per_step": e2e_s / args.steps * 1e3,
               "api": "pcr_render_frames_host_submit / pcr_host_wait (pinned host trajectory in, pinned host RGBA8 out; copies overlap "
                      "kernels; two calls in flight, every step's H2D and D2H inside the timed region)"}
    if sampler:
        sampler.end()           # the clock record covers both timed regions (device-resident and end-to-end)
        sampler.stop()
    numa_all = [numa]
    if world > 1:
        numa_all = [None] * world
        dist.all_gather_object(numa_all, numa)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- rooflines ----------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    clocks = sampler.summary() if sampler else None
    roofline = None
    kernels = {}
    SIDE = ("k_stats", "k_mean_sequential")          # launched on side streams (K0 look-ahead): they overlap the render kernels
    if prof:
        side_active = lookahead > 0
        inline = {k: v for k, v in prof.items() if not (side_active and k in SIDE)}
        total_ms = sum(v[0] for v in inline.values())
        for name, (ms, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
            kernels[name] = {"ms_total": round(ms, 4), "launches": int(cnt), "us_per_launch": round(ms / cnt * 1e3, 3)}
            if name in inline:
                kernels[name]["share"] = round(ms / total_ms, 4)
            else:
                kernels[name]["overlapped"] = "side stream (look-ahead): runs concurrently with the render kernels, not part of the shares"
        top = max(inline.items(), key=lambda kv: kv[1][0])
        top_ms, top_cnt = top[1]
        launches_per_step = top_cnt / args.steps
        frames_per_launch = min(B, ctx.max_batch)            # every launch covers one whole internal batch
        step_bytes = spec["algorithmic_bytes_per_frame"] * B
        # the kernel's average launch handles step_bytes / launches_per_step of the step's algorithmic bytes (its
        # launches share the step: pre-pass + main pass), so achieved = step bytes / (the kernel's time per step)
        bytes_per_launch = step_bytes / launches_per_step
        us_per_launch = top_ms / top_cnt * 1e3
        achieved = bytes_per_launch / (us_per_launch * 1e-6) / 1e9
        traffic = issue = None
        cpath = os.path.join(ROOT, "profiles", "ncu_counters.json")
        if os.path.isfile(cpath):
            # committed ncu capture of the same command (profiles/): DRAM bytes and warp instructions per frame and launch
            cj = json.load(open(cpath)).get(args.workload, {})
            # (k_project_cull4 is the main-pass variant of K2a: the library's profile counts it as k_project_count)
            names = ("k_project_count", "k_project_cull4") if top[0] == "k_project_count" else (top[0],)
            ks = [cj.get("kernels", {}).get(nm) for nm in names]
            ks = [k for k in ks if k]
            traffic = (sum(k["dram_bytes"] for k in ks) / sum(k["launches"] for k in ks) / cj["frames_per_step"] * frames_per_launch) if ks else None
            wi = cj.get("warp_instructions_per_frame")
            if wi and clocks and clocks.get("sm_mhz"):
                sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
                issue_peak = sms * 4 * clocks["sm_mhz"] * 1e6                # warp instructions / s: 4 schedulers per SM, 1 per clock
                issue = {"warp_instructions_per_step": wi * B, "issue_peak_per_s": issue_peak,
                         "achieved_per_s": wi * B * args.steps / (dev_ms * 1e-3),
                         "frac": wi * B * args.steps / (dev_ms * 1e-3) / issue_peak,
                         "source": f"profiles/ncu_counters.json ({cj.get('capture', 'ncu smsp__inst_executed.sum per kernel')}); peak = SMs x 4 schedulers x measured SM clock"}
        roofline = {"bound": "hbm", "kernel": top[0], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": bytes_per_launch, "algorithmic_bytes_per_step": step_bytes,
                    "frames_per_launch": frames_per_launch, "us_per_launch": us_per_launch, "launches_per_step": launches_per_step,
                    "whole_step_achieved_gbs": step_bytes * args.steps / (dev_ms * 1e-3) / 1e9,
                    "whole_step_frac": step_bytes * args.steps / (dev_ms * 1e-3) / 1e9 / peak,
                    "issue": issue,
                    "note": "algorithmic bytes = N*b_in + W*H*(8+4) per frame (SURVEY.md 8d). frac = the step's algorithmic bytes / the top kernel's "
                            "time per step (all its launches) / peak; whole_step_frac = the same bytes / the whole step. The path is bound by "
                            "sphere-pixel tests (instruction issue: roofline.issue), not by HBM bytes"}

    line = {"metric": "frames_per_s", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "mpoints_per_s": fps * n / 1e6,
            "config": config_dict(args.workload, spec, ring, world, trails=bool(args.trails and spec["cols"] == 6), mean=args.mean),
            "stats_lookahead_steps": lookahead, "ms_per_step_with_kernel_events": (prof_ms / args.steps) if prof_ms else None,
            "e2e": e2e, "gpu_launches": int(launches), "kernels": kernels, "roofline": roofline,
            "per_schedule_third": thirds,
            "clocks": clocks, "host_placement": numa_all if world > 1 else numa,
            "nvidia_smi_topo": smi_topology()[0][:4000] if world > 1 else None,
            "pairs_last_frame": counters["pairs_last_frame"], "overflow_frames": counters["overflow_frames"]}
    if trace:
        line["step_ms"] = trace[0]
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(spec, host_np, radius_np)
    print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
