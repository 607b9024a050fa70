"""Host-side mirror of the reference's render surface: Camera, Scene, Image, ImageOutput.

Same names, argument order and meaning as the F# modules, so that callers (and tests) read like the
reference's.  Everything that computes goes through the C ABI (native.py); nothing here traces rays.

  Camera.make_basic     RayTracing/Camera.fs:34-59
  Scene.make            RayTracing/Scene.fs:15-28
  Scene.render          RayTracing/Scene.fs:196-236
  Image / Image.render  RayTracing/Domain.fs:9-31
  ImageOutput.write_ppm RayTracing/ImageOutput.fs:163-197, PixelOutput.correct :11-18
"""
import os
import secrets
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import numpy as np

from . import abi, native
from .domain import marshal


class Camera:
    @staticmethod
    def make_basic(samples_per_pixel: int, focal_length: float, aspect_ratio: float, origin, view_direction, view_up) -> abi.RtCamera:
        """Camera.makeBasic.  BounceDepth is 150 as in the reference (Camera.fs:58); override the field
        like `{ camera with BounceDepth = 50 }`: `cam.bounce_depth = 50`."""
        return native.camera_make_basic(samples_per_pixel, focal_length, aspect_ratio, origin, view_direction, view_up)


@dataclass
class Image:
    """Domain.fs:9-15: `Rows` is a sequence of deferred rows; here each is a thunk returning uint8 [cols, 3]."""
    Rows: Sequence[Callable[[], np.ndarray]]
    RowCount: int
    ColCount: int
    # set by Scene.render: forces every row at once (one frame, one progress call per row) without 2·maxH+1 array copies
    _all_rows: Optional[Callable[[], np.ndarray]] = None

    @staticmethod
    def row_count(i: "Image") -> int:
        return i.RowCount

    @staticmethod
    def col_count(i: "Image") -> int:
        return i.ColCount

    @staticmethod
    def render(i: "Image") -> np.ndarray:
        """Image.render (Domain.fs:23-24): force every row; returns uint8 [rows, cols, 3]."""
        if i._all_rows is not None:
            return i._all_rows()
        return np.stack([row() for row in i.Rows]) if i.RowCount else np.zeros((0, i.ColCount, 3), np.uint8)


class Scene:
    """Scene.make / Scene.render.  The record is opaque, as in the reference (Scene.fs:5-9)."""

    def __init__(self, handle, hittables, textures, keep, devices):
        self._handle = handle
        self._hittables, self._textures, self._keep = hittables, textures, keep
        self._devices = devices
        self.last_stats: Optional[abi.RtStats] = None

    @staticmethod
    def make(objects, device: int = 0, devices: Optional[List[int]] = None) -> "Scene":
        """Scene.make: partitions bounded / unbounded objects and builds the BVH (host), uploads it.
        `devices=[...]` (more than one) splits every frame over those GPUs of this process."""
        hittables, textures, keep = marshal(objects)
        if devices is not None and len(devices) > 1:
            handle = native.MultiHandle(hittables, textures, devices, keepalive=keep)
            return Scene(handle, hittables, textures, keep, list(devices))
        dev = devices[0] if devices else device
        handle = native.SceneHandle(hittables, textures, dev, keepalive=keep)
        return Scene(handle, hittables, textures, keep, [dev])

    @property
    def handle(self):
        return self._handle

    @staticmethod
    def render(progress_increment: Callable[[float], None], print_fn: Callable[[str], None], max_width_coord: int, max_height_coord: int,
               camera: abi.RtCamera, s: "Scene", seed: Optional[int] = None, adaptive: bool = True, mode: int = abi.RT_MODE_MEGAKERNEL,
               flags: int = 0, frames: Optional["native.FrameRing"] = None):
        """Scene.render: returns (total progress, Image).  Like the reference it is lazy: nothing is
        traced until a row is forced; the first forced row renders the whole frame on the GPU.
        `seed`: key of the counter RNG; None draws one from the OS as `FloatProducer (Random ())` does
        (Scene.fs:205).  `frames`: a native.FrameRing to take the output array from (explicit reuse by a host that renders
        frame after frame); by default every frame is a fresh array.  `print_fn` is accepted and ignored exactly as the reference ignores it (Scene.fs:158)."""
        rows, cols = 2 * max_height_coord + 1, 2 * max_width_coord + 1  # Scene.fs:208-209
        if seed is None:
            seed = secrets.randbits(64)
        state = {}

        def frame():
            if "rgb" not in state:
                out = frames.next() if frames is not None else None
                if isinstance(s._handle, native.MultiHandle):
                    rgb, _, stats = s._handle.render(camera, max_width_coord, max_height_coord, seed=seed, adaptive=adaptive, flags=flags,
                                                     rgb_out=out)
                else:
                    rgb, _, stats = s._handle.render(camera, max_width_coord, max_height_coord, seed=seed, adaptive=adaptive, mode=mode,
                                                     flags=flags, rgb_out=out)
                s.last_stats = stats
                state["rgb"] = rgb
            return state["rgb"]

        def row_thunk(r):
            def force():
                out = frame()[r]
                progress_increment(1.0)  # Scene.fs:232
                return out
            return force

        def all_rows():
            out = frame()
            for _ in range(rows):
                progress_increment(1.0)
            return out

        return float(rows), Image([row_thunk(r) for r in range(rows)], rows, cols, all_rows)


class PixelOutput:
    @staticmethod
    def correct(b: int) -> int:
        """PixelOutput.correct (ImageOutput.fs:11-18)."""
        return native.gamma_correct(b)


class ImageOutput:
    @staticmethod
    def write_ppm(gamma_correct: bool, pixels: np.ndarray, path: Optional[os.PathLike] = None) -> bytes:
        """ImageOutput.writePpm (ImageOutput.fs:163-197): P3 text, byte-identical format."""
        data = native.ppm_format(pixels, gamma_correct)
        if path is not None:
            native.ppm_write_file(pixels, path, gamma_correct)
        return data
