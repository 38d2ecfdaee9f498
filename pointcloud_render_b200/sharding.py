"""Multi-GPU modes of the hot path, one process per GPU over torch.distributed.

The reference renders frames strictly sequentially (traj_ball_renderer.py:461-470) and has no
collective anywhere; both modes are new:

  frames  — every frame is standardised and rendered independently (traj_ball_renderer.py:373),
            so a trajectory shards over ranks in contiguous blocks with NO collective.
  points  — one huge cloud is split by point range.  Two exchange steps exist:
            C0  all-gather of every shard's {sum xyz | min xyz | max xyz} (9 doubles), folded in
                rank order and finalised on the device so every rank standardises with the same
                global mean / extent (example_renderer.py:96-97);
            C1  min all-reduce of the packed (depth|id) z-buffer — keys are < 2^63 (the depth is a
                positive float), so int64 MIN orders them exactly like uint64 MIN: NCCL's
                ncclAllReduce(ncclInt64, ncclMin) over NVLink, exact and order independent.
            Shading then runs owner-only and the RGBA8 slices are assembled with a byte MAX.

            Fused alternative (render_point_sharded_fused): no big collective at all — every rank owns a band
            of image rows of the merged z-buffer, the raster pushes its winners into the owner's rows with
            64-bit atomicMin over NVLink while it is still rasterising, and the shade kernel stores each pixel
            it won directly into rank 0's image.  Peer memory is mapped with CUDA IPC (PeerMesh).

The collectives run on whatever backend the process group has (nccl on the GPU box, gloo in the
CPU tests); the compute between them is the pcr C ABI.
"""
import numpy as np


def frame_shard(n_frames, rank, world):
    """Contiguous block [start, stop) of frames for `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(n_frames), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def point_shard(n_points, rank, world):
    """Contiguous point range [start, stop) of `rank`; start is also the rank's id_base."""
    return frame_shard(n_points, rank, world)


def finalize_stats(total9, n_total, dtype):
    """double[9] totals (sum xyz, min xyz, max xyz) -> double[10] (mean xyz, min xyz, max xyz, scale)
    with the roundings of standardize_point_cloud: mean and extent in the INPUT dtype
    (example_renderer.py:96-97: np.mean / np.amax(pcl - np.amin(pcl, 0)))."""
    T = np.dtype(dtype).type
    total9 = np.asarray(total9, np.float64)
    out = np.empty(10, np.float64)
    out[0:3] = [float(T(total9[k] / float(n_total))) for k in range(3)]
    out[3:9] = total9[3:9]
    out[9] = max(float(T(T(total9[6 + k]) - T(total9[3 + k]))) for k in range(3))
    return out


def allreduce_stats(partial9, n_local, dtype, group=None):
    """C0.  partial9: float64 tensor[9] of this rank's shard (pcr_stats_partial).  Returns the
    global float64 tensor[10] on the same device.  Three tiny all-reduces (sum / min / max)."""
    import torch
    import torch.distributed as dist
    s = partial9[0:3].clone()
    lo = partial9[3:6].clone()
    hi = partial9[6:9].clone()
    n = torch.tensor([float(n_local)], dtype=torch.float64, device=partial9.device)
    dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(n, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    total = torch.cat([s, lo, hi]).cpu().numpy()
    stats = finalize_stats(total, int(round(float(n.item()))), dtype)
    return torch.from_numpy(stats).to(partial9.device)


def allgather_stats_device(ctx, partial9, n_total, is_f64, group=None):
    """C0 without a host round trip: ONE all-gather of the 9 shard totals, folded in rank order and
    finalised on the device (pcr_finalize_stats).  Every rank computes bit-identical stats."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    gathered = torch.empty((world, 9), dtype=torch.float64, device=partial9.device)
    dist.all_gather_into_tensor(gathered, partial9.contiguous(), group=group)
    return ctx.finalize_stats(gathered, n_total, is_f64)


def zmerge_(vis, group=None):
    """C1.  In-place min all-reduce of the (H,W) int64 view of the uint64 keys."""
    import torch
    import torch.distributed as dist
    assert vis.dtype == torch.int64
    dist.all_reduce(vis, op=dist.ReduceOp.MIN, group=group)
    return vis


class NcclComm:
    """An ncclComm_t of our own, for the C entry pcr_zmerge_nccl (`ncclAllReduce(ncclUint64, ncclMin)` on the raw keys;
    torch.distributed does not hand out its communicator).  Created with the NCCL already loaded in this process
    (torch's): rank 0 draws the ncclUniqueId, `bcast` (a callable taking / returning 128 bytes — e.g. a
    torch.distributed broadcast on any backend) distributes it, every rank calls ncclCommInitRank.  One rank per GPU."""

    def __init__(self, rank=0, world=1, bcast=None):
        import ctypes
        import glob
        import os
        import torch
        cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "nccl", "lib", "libnccl.so*")) + ["libnccl.so.2"]
        self.lib = None
        for c in cands:
            try:
                self.lib = ctypes.CDLL(c, mode=ctypes.RTLD_GLOBAL)        # global: pcr_zmerge_nccl finds ncclAllReduce with dlsym
                break
            except OSError:
                continue
        if self.lib is None:
            raise RuntimeError("libnccl not found")

        class UniqueId(ctypes.Structure):
            _fields_ = [("internal", ctypes.c_char * 128)]
        uid = UniqueId()
        if rank == 0:
            self._ck(self.lib.ncclGetUniqueId(ctypes.byref(uid)), "ncclGetUniqueId")
        raw = bytes(ctypes.string_at(ctypes.byref(uid), 128))
        if world > 1:
            raw = bcast(raw)
            ctypes.memmove(ctypes.byref(uid), raw, 128)
        self.comm = ctypes.c_void_p()
        self.lib.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, UniqueId, ctypes.c_int]
        self._ck(self.lib.ncclCommInitRank(ctypes.byref(self.comm), int(world), uid, int(rank)), "ncclCommInitRank")
        self.rank, self.world = int(rank), int(world)

    @staticmethod
    def _ck(rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed with ncclResult {rc}")

    @classmethod
    def from_process_group(cls, group=None):
        """One communicator over the ranks of a torch.distributed group (the id travels as a 128-byte broadcast)."""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)

        def bcast(raw):
            dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
            t = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
            dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            return bytes(t.cpu().numpy().tobytes())
        return cls(rank, world, bcast)

    def close(self):
        import ctypes
        if getattr(self, "comm", None) and self.comm.value:
            self.lib.ncclCommDestroy.argtypes = [ctypes.c_void_p]
            self.lib.ncclCommDestroy(self.comm)
            self.comm = None


def assemble_image_(rgba, group=None):
    """Byte-wise MAX of owner-only shaded images (pcr_shade owner_only=1 writes 0 where the
    winner is not local, and floor/miss pixels only on rank 0)."""
    import torch.distributed as dist
    dist.all_reduce(rgba, op=dist.ReduceOp.MAX, group=group)
    return rgba


def point_sharded_buffers(n_local, cam, device):
    """Reusable work buffers for render_point_sharded (avoids per-frame allocations)."""
    import torch
    return {"vis": torch.empty((cam.height, cam.width), dtype=torch.int64, device=device),
            "rgba": torch.empty((cam.height, cam.width, 4), dtype=torch.uint8, device=device)}


def render_point_sharded(ctx, pts_local, id_base, n_total, cam, style, radius=None, rgb=None, group=None, shade=True, buffers=None,
                         nccl_comm=None):
    """The whole point-sharded path on one rank: K0 partials -> C0 -> K2/K3 (K1 inlined) -> C1 -> K4(owner)
    -> byte MAX.  pts_local: (n_local, 3|6) CUDA tensor, this rank's slice of the n_total-point cloud.
    Stream-ordered, no host synchronisation.  nccl_comm (an NcclComm): C1 through the C entry pcr_zmerge_nccl
    (ncclAllReduce on the uint64 keys) instead of torch.distributed's int64 all-reduce — same bits."""
    import torch
    b = buffers or {}
    part = ctx.stats_partial(pts_local)
    stats = allgather_stats_device(ctx, part, n_total, pts_local.dtype == torch.float64, group)
    # fused: K1 is evaluated inside K2a (binning) and K4 (winners only); nothing is materialised
    vis = ctx.render_shard(pts_local, stats, cam, style, id_base=id_base, radius=radius, rgb=rgb, out_vis=b.get("vis"))
    if nccl_comm is not None:
        ctx.zmerge_nccl_(vis, nccl_comm)
    else:
        zmerge_(vis, group)
    if not shade:
        return vis, None
    rgba = ctx.shade_shard(vis, pts_local, stats, cam, style, id_base=id_base, owner_only=True, radius=radius, rgb=rgb,
                           out_rgba=b.get("rgba"))
    assemble_image_(rgba, group)
    return vis, rgba


class PeerMesh:
    """Peer-memory set-up of the fused merge: every rank allocates its merged z-buffer + image inside libpcr
    (cudaMalloc), the 64-byte CUDA IPC handles are all-gathered, every rank maps every other rank's buffers
    over NVLink and hands the pointer table to its context (pcr_peer_set).  One process per GPU."""

    def __init__(self, ctx, cam, group=None, dst_rank=0):
        import torch
        import torch.distributed as dist
        self.ctx, self.group = ctx, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        dev = torch.device("cuda", ctx.device)
        m, im = ctx.peer_alloc(cam.width, cam.height)
        mine = torch.frombuffer(bytearray(ctx.ipc_export(m) + ctx.ipc_export(im)), dtype=torch.uint8).to(dev)
        allh = torch.empty((self.world, 128), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allh, mine, group=group)
        allh = allh.cpu().numpy()
        self.opened, merged, image = [], [], []
        for r in range(self.world):
            if r == self.rank:
                merged.append(m)
                image.append(im)
            else:
                pm, pi = ctx.ipc_open(allh[r, :64].tobytes()), ctx.ipc_open(allh[r, 64:].tobytes())
                self.opened += [pm, pi]
                merged.append(pm)
                image.append(pi)
        ctx.peer_set(self.rank, self.world, merged, image, dst_rank=dst_rank)
        self.flag = torch.zeros(1, dtype=torch.int32, device=dev)
        self.dst_rank = dst_rank

    def barrier(self):
        """Stream-ordered cross-rank barrier: a 1-element all-reduce (every rank's earlier kernels — and the
        remote reductions / stores they issued — complete before any rank's later kernels start)."""
        import torch.distributed as dist
        dist.all_reduce(self.flag, group=self.group)

    def close(self):
        import torch
        torch.cuda.synchronize()
        self.barrier()
        torch.cuda.synchronize()
        self.ctx.peer_set(0, 0, [], [])
        for p in self.opened:
            self.ctx.ipc_close(p)
        self.opened = []


def render_point_sharded_fused(ctx, mesh, pts_local, id_base, n_total, cam, style, radius=None, rgb=None, buffers=None):
    """The point-sharded path with the z-merge fused into the raster (see PeerMesh): returns (local keys,
    image) — the image is complete on rank mesh.dst_rank only (a view of its peer image buffer), None elsewhere.
    Stream-ordered, no host synchronisation; three tiny collectives per frame."""
    import torch
    b = buffers or {}
    ctx.peer_begin_frame(cam, style)
    part = ctx.stats_partial(pts_local)
    stats = allgather_stats_device(ctx, part, n_total, pts_local.dtype == torch.float64, mesh.group)   # also: every rank's rows are initialised
    vis = ctx.render_shard_peer(pts_local, stats, cam, style, id_base=id_base, radius=radius, rgb=rgb, out_vis=b.get("vis"))
    mesh.barrier()                                                   # all pushes have landed
    ctx.shade_shard_peer(vis, pts_local, stats, cam, style, id_base=id_base, radius=radius, rgb=rgb)
    mesh.barrier()                                                   # the image is complete; rows may be reused
    return vis, (ctx.peer_buffers()[1] if mesh.rank == mesh.dst_rank else None)
