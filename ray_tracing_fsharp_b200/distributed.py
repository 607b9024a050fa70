"""One frame shared out over ranks (one process per GPU): the sample-split of SURVEY.md §8e.

    phase 1  rt_device_probe   rank r probes the tiles t with t mod world == r      (Scene.fs:172-188)
             all_reduce(flags, MAX)                                                  [only if world > 1]
    phase 2  rt_device_main    rank r adds samples n_probe + r + j*world of every flagged pixel (:191-192)
             all_reduce(stats, SUM)   int32 {sumR, sumG, sumB, count} per pixel      [only if world > 1]
    tail     rt_device_finalize  PixelStats.mean (+ gamma) -> RGB8                   (Pixel.fs:103-108)

Integer sums keyed by sample index make the result independent of `world`.

The PRODUCT form of this sequence is one library call, `rt_comm_render` (csrc/rtfs_comm.cu): kernels and NCCL
collectives enqueued by the library on one stream (`native.CommHandle`; `comm_from_torch_distributed` below only
carries the 128-byte NCCL id from rank 0 to the others).  `render_split_frame` spells the same steps out phase by
phase over a backend: tests/ drive it with a CPU backend over gloo to check the decomposition, and with
`DeviceBackend` (rt_device_probe / rt_device_main on torch tensors) to emulate several ranks on one GPU.
"""
import ctypes as C

from . import abi, native


def render_split_frame(backend, rank: int, world: int, all_reduce_max=None, all_reduce_sum=None):
    """Runs the three steps above on `backend`; returns (stats, flags) after the reductions.
    `all_reduce_max(flags)` / `all_reduce_sum(stats)` are in-place collectives (ignored when world == 1)."""
    stats, flags = backend.alloc()
    backend.probe(rank, world, stats, flags)
    if world > 1:
        all_reduce_max(flags)
    backend.main(rank, world, stats, flags)
    if world > 1:
        all_reduce_sum(stats)
    return stats, flags


def comm_from_torch_distributed(device_index: int, stream=None) -> native.CommHandle:
    """This process's rank of the library's communicator, bootstrapped over an initialised torch.distributed group
    (any backend): rank 0 draws the NCCL id, broadcast_object_list carries it, every rank joins.  Without a process
    group (or world 1) the communicator is a single rank and NCCL is never loaded."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return native.CommHandle(None, 0, 1, device_index, stream)
    rank, world = dist.get_rank(), dist.get_world_size()
    box = [native.CommHandle.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return native.CommHandle(box[0], rank, world, device_index, stream)


class DeviceBackend:
    """rt_device_probe / rt_device_main / rt_device_finalize on torch CUDA tensors (current stream)."""

    def __init__(self, scene: native.SceneHandle, camera: abi.RtCamera, max_w: int, max_h: int, seed: int = 0, adaptive: bool = True,
                 flags: int = 0):
        import torch
        self.torch = torch
        self.scene, self.camera, self.max_w, self.max_h = scene, camera, max_w, max_h
        self.opts = abi.RtRenderOpts(seed, int(adaptive), abi.RT_MODE_MEGAKERNEL, 0, flags)
        self.n_pixels = (2 * max_w + 1) * (2 * max_h + 1)
        self.device = torch.device("cuda", scene.device)
        self.launches = 0
        self._stats = None
        self._flags = None
        self._rgb = None

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def alloc(self):
        torch = self.torch
        if self._stats is None:
            self._stats = torch.empty((self.n_pixels, 4), dtype=torch.int32, device=self.device)
            self._flags = torch.empty((self.n_pixels,), dtype=torch.uint8, device=self.device)
            self._rgb = torch.empty((self.n_pixels, 3), dtype=torch.uint8, device=self.device)
        self._stats.zero_()
        self._flags.zero_()
        return self._stats, self._flags

    def probe(self, rank, world, stats, flags):
        native.check(native.lib().rt_device_probe(self.scene.ptr, C.byref(self.camera), self.max_w, self.max_h, C.byref(self.opts), rank,
                                                  world, C.c_void_p(stats.data_ptr()), C.c_void_p(flags.data_ptr()), self._stream(), None))
        self.launches += 2 if self.opts.adaptive else 0  # probe + probe_flags

    def main(self, rank, world, stats, flags):
        native.check(native.lib().rt_device_main(self.scene.ptr, C.byref(self.camera), self.max_w, self.max_h, C.byref(self.opts), rank,
                                                 world, C.c_void_p(stats.data_ptr()), C.c_void_p(flags.data_ptr()), self._stream(), None))
        self.launches += 2  # compact + main

    def finalize(self, stats, gamma=False):
        native.check(native.lib().rt_device_finalize(self.scene.device, C.c_void_p(stats.data_ptr()), self.n_pixels, int(gamma),
                                                     C.c_void_p(self._rgb.data_ptr()), self._stream()))
        self.launches += 1
        return self._rgb

    _host_frames = {}  # n_pixels -> two pinned host frames, used alternately

    def finalize_to_host(self, stats, gamma=False):
        """finalize + device->host copy of the RGB8 frame into a recycled PINNED host buffer (a fresh pageable tensor costs
        the copy a page fault per 4 KiB: 1.3 ms of a 12 ms 8-GPU frame).  Returns a uint8 numpy view [n_pixels, 3] that
        stays valid until the call after next."""
        torch = self.torch
        ring = DeviceBackend._host_frames.setdefault(self.n_pixels, [[], 0])
        if len(ring[0]) < 2:
            ring[0].append(torch.empty((self.n_pixels, 3), dtype=torch.uint8, pin_memory=True))
        host = ring[0][ring[1] % len(ring[0])]
        ring[1] += 1
        host.copy_(self.finalize(stats, gamma), non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return host.numpy()

    def counters(self) -> abi.RtStats:
        """Work counters (paths, rays, ...) of this rank since the last probe; synchronises the stream."""
        out = abi.RtStats()
        native.check(native.lib().rt_device_counters(self.scene.ptr, self._stream(), C.byref(out)))
        return out
