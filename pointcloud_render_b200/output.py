"""Output stage: the file half of save_scene, off the render thread.

The reference's save_scene is mi.util.write_bitmap(path + '.png', image) (example_renderer.py:159-161),
which Mitsuba runs asynchronously by default (write_async=True).  At thousands of frames per second the
PNG encoder, not the renderer, is the bottleneck (12-20 ms per 1024^2 frame and core), so frames are
encoded by a pool of worker threads (the encoders release the GIL) while the GPU renders the next
chunk.  Naming rules of the reference's process() are kept by `trajectory_frame_name`.
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np


def trajectory_frame_name(filename, frame_index):
    """Output stem of TrajectoryBallRenderer.process (traj_ball_renderer.py:376): the input file's
    stem for frames <= 199, 'frame_XXXX_b0' for the fade frames — also for the b1 / original
    subclasses, which inherit the rule."""
    return f'frame_{frame_index:04d}_b0' if frame_index > 199 else filename


def _encode_png(path, rgb, compress_level):
    try:
        import cv2
        ok, buf = cv2.imencode(".png", rgb[..., ::-1], [cv2.IMWRITE_PNG_COMPRESSION, int(compress_level)])
        if not ok:
            raise RuntimeError("cv2.imencode failed")
        with open(path, "wb") as f:
            f.write(buf.tobytes())
    except ImportError:
        from PIL import Image
        Image.fromarray(rgb).save(path, format="PNG", compress_level=int(compress_level))
    return path


class AsyncImageWriter:
    """Thread-pool image writer.  submit() copies nothing: the caller must keep the array alive and
    unmodified until the returned future is done (render_to_files double-buffers for that)."""

    def __init__(self, workers=None, compress_level=1, fmt="png"):
        if fmt not in ("png", "npy"):
            raise ValueError("fmt must be 'png' or 'npy'")
        self.fmt, self.compress_level = fmt, compress_level
        self.pool = ThreadPoolExecutor(max_workers=workers or min(32, (os.cpu_count() or 4)))
        self.pending = []

    def submit(self, output_file_path, rgba):
        """Write `rgba` ((H,W,4|3) uint8, sRGB) to output_file_path + '.png' (or '.npy')."""
        a = np.asarray(rgba)
        rgb = a[..., :3] if a.shape[-1] == 4 else a
        os.makedirs(os.path.dirname(output_file_path) or ".", exist_ok=True)
        if self.fmt == "npy":
            fut = self.pool.submit(lambda p=output_file_path + ".npy", x=rgb: (np.save(p, x), p)[1])
        else:
            fut = self.pool.submit(_encode_png, output_file_path + ".png", np.ascontiguousarray(rgb), self.compress_level)
        self.pending.append(fut)
        return fut

    def drain(self):
        """Wait for everything submitted so far; re-raises the first encoder error."""
        done, self.pending = self.pending, []
        return [f.result() for f in done]

    def close(self):
        self.drain()
        self.pool.shutdown(wait=True)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def render_to_files(renderer, traj, output_folder, first_frame=0, total_frames=None, stem=None, chunk=16, writer=None):
    """Whole output path for a trajectory: render chunks of frames through the host-buffer entry
    (pcr_render_frames_host), hand each chunk to the writer pool and keep rendering — two pinned
    image buffers alternate so the encoders of chunk k read while chunk k+1 is produced.
    Returns the list of written paths."""
    import torch
    F = traj.shape[0]
    total = int(total_frames or F)
    own = writer is None
    writer = writer or AsyncImageWriter()
    bufs = [torch.empty((chunk, renderer.height, renderer.width, 4), dtype=torch.uint8).pin_memory() for _ in range(2)]
    in_flight = [[], []]
    paths = []
    try:
        for k, f0 in enumerate(range(0, F, chunk)):
            nb = min(chunk, F - f0)
            slot = k & 1
            for fut in in_flight[slot]:
                fut.result()                                   # encoders of chunk k-2 are done with this buffer
            out = bufs[slot][:nb]
            renderer.render_trajectory(traj[f0:f0 + nb], first_frame=first_frame + f0, total_frames=total, out_rgba=out)
            in_flight[slot] = []
            for j in range(nb):
                idx = first_frame + f0 + j
                name = trajectory_frame_name(f'{stem or "frame"}_{idx:04d}', idx)
                fut = writer.submit(os.path.join(output_folder, name), out[j].numpy())
                in_flight[slot].append(fut)
                paths.append(os.path.join(output_folder, name) + "." + writer.fmt)
        writer.drain()
    finally:
        if own:
            writer.close()
    return paths
