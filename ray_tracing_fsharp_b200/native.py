"""ctypes binding of librtfs_b200.so (include/rtfs_b200.h).

This is the only way the package reaches the compute path, and it fails loudly: a missing library
raises ImportError-like RuntimeError with the build command, a missing GPU surfaces as
RtError(RT_ERR_NO_DEVICE) from the library itself.  There is no CPU fallback anywhere.
"""
import ctypes as C
import os

import numpy as np

from . import abi
from .abi import RtCamera, RtHittable, RtRenderOpts, RtStats, RtTexture

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librtfs_b200.so")


class RtError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"librtfs_b200 error {code}: {message}")
        self.code = code


_lib = None

_vp = C.c_void_p
_i32 = C.c_int32


def _declare(lib):
    sig = {
        "rt_abi_version": (C.c_int, []),
        "rt_last_error": (C.c_char_p, []),
        "rt_device_count": (C.c_int, []),
        "rt_host_pin": (C.c_int, [_vp, C.c_size_t]),
        "rt_host_unpin": (C.c_int, [_vp]),
        "rt_camera_make_basic": (C.c_int, [_i32, C.c_double, C.c_double, _vp, _vp, _vp, C.POINTER(RtCamera)]),
        "rt_ppm_format": (C.c_int, [_vp, _i32, _i32, _i32, _vp, C.c_size_t, C.POINTER(C.c_size_t)]),
        "rt_ppm_write_file": (C.c_int, [_vp, _i32, _i32, _i32, C.c_char_p]),
        "rt_gamma_correct": (C.c_uint8, [C.c_uint8]),
        "rt_scene_create": (C.c_int, [_vp, _i32, _vp, _i32, _i32, C.POINTER(_vp)]),
        "rt_scene_destroy": (None, [_vp]),
        "rt_scene_bvh_node_count": (C.c_int, [_vp, _i32]),
        "rt_scene_bvh_nodes": (C.c_int, [_vp, _i32, _vp, _vp, _vp]),
        "rt_scene_wide_bvh_check": (C.c_int, [_vp, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(C.c_double)]),
        "rt_scene_device_bvh_check": (C.c_int, [_vp, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32)]),
        "rt_scene_device_bytes": (C.c_size_t, [_vp]),
        "rt_scene_shared_memory_bytes": (C.c_size_t, [_vp]),
        "rt_render": (C.c_int, [_vp, C.POINTER(RtCamera), _i32, _i32, C.POINTER(RtRenderOpts), _vp, _vp, C.POINTER(RtStats)]),
        "rt_render_multi": (C.c_int, [_vp, _i32, _vp, _i32, _vp, _i32, C.POINTER(RtCamera), _i32, _i32, C.POINTER(RtRenderOpts), _vp,
                                      _vp, C.POINTER(RtStats)]),
        "rt_multi_create": (C.c_int, [_vp, _i32, _vp, _i32, _vp, _i32, C.POINTER(_vp)]),
        "rt_multi_render": (C.c_int, [_vp, C.POINTER(RtCamera), _i32, _i32, C.POINTER(RtRenderOpts), _vp, _vp, C.POINTER(RtStats)]),
        "rt_multi_destroy": (None, [_vp]),
        "rt_device_probe": (C.c_int, [_vp, C.POINTER(RtCamera), _i32, _i32, C.POINTER(RtRenderOpts), _i32, _i32, _vp, _vp, _vp,
                                      C.POINTER(RtStats)]),
        "rt_device_main": (C.c_int, [_vp, C.POINTER(RtCamera), _i32, _i32, C.POINTER(RtRenderOpts), _i32, _i32, _vp, _vp, _vp,
                                     C.POINTER(RtStats)]),
        "rt_device_counters": (C.c_int, [_vp, _vp, C.POINTER(RtStats)]),
        "rt_device_finalize": (C.c_int, [_i32, _vp, _i32, _i32, _vp, _vp]),
        "rt_comm_unique_id": (C.c_int, [_vp]),
        "rt_comm_create": (C.c_int, [_vp, _i32, _i32, _i32, _vp, C.POINTER(_vp)]),
        "rt_comm_destroy": (None, [_vp]),
        "rt_comm_uses_peer_memory": (C.c_int, [_vp]),
        "rt_comm_nccl_version": (C.c_int, [C.POINTER(_i32), C.c_char_p, C.c_size_t]),
        "rt_comm_render": (C.c_int, [_vp, _vp, C.POINTER(RtCamera), _i32, _i32, C.POINTER(RtRenderOpts), _vp, _vp, C.POINTER(RtStats)]),
        "rt_comm_last_stats": (C.c_int, [_vp, _vp, C.POINTER(RtStats)]),
        "rt_comm_frame": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp)]),
        "rt_test_sphere_hit": (C.c_int, [_i32, _i32, _vp, _vp, _vp, _vp, _vp]),
        "rt_test_plane_hit": (C.c_int, [_i32, _i32, _vp, _vp, _vp, _vp, _vp]),
        "rt_test_aabb_hit": (C.c_int, [_i32, _i32, _vp, _vp, _vp, _vp, _vp]),
        "rt_test_hit_object": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
        "rt_test_reflection": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
        "rt_test_camera_rays": (C.c_int, [_i32, C.POINTER(RtCamera), _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
        "rt_test_texture": (C.c_int, [_vp, _i32, _vp, _vp, _vp]),
        "rt_test_combine_darken": (C.c_int, [_i32, _i32, _vp, _vp, _vp, _vp]),
        "rt_test_rng": (C.c_int, [_i32, C.c_uint64, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
        "rt_test_trace_samples": (C.c_int, [_vp, C.POINTER(RtCamera), _i32, _i32, C.c_uint64, _i32, _vp, _vp, _vp, _vp, _vp]),
        "rt_measure_fp32_peak": (C.c_int, [_i32, C.POINTER(C.c_double)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    return sorted(sig)


EXPORTS = None


def lib():
    """The loaded library.  Raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib, EXPORTS
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `make -C ray_tracing_fsharp_b200/csrc` "
                "(or __graft_entry__.build()).  There is no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        EXPORTS = _declare(handle)
        if handle.rt_abi_version() != abi.RT_ABI_VERSION:
            raise RuntimeError("librtfs_b200.so was built from a different include/rtfs_b200.h (ABI version mismatch)")
        _lib = handle
    return _lib


def check(rc):
    if rc != abi.RT_OK:
        raise RtError(rc, lib().rt_last_error().decode("utf-8", "replace"))
    return rc


def ptr(a):
    """void* of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data)


def f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a.reshape(shape) if shape is not None else a


class FrameRing:
    """Explicit recycling of output frames, for a host that renders frame after frame: hands out `count` uint8 arrays of
    one shape in turn (`next()`), each already touched.  Opt-in — pass `rgb_out=ring.next()` (or `frames=ring` to
    Scene.render): the caller states that it is done with the frame it got `count` calls ago.  Without it every render
    returns a fresh array that is never handed out again.  (A fresh np.empty is untouched memory, and the device->host copy
    into it pays a page fault per 4 KiB — 1.3 ms for the 2.9 MB C2 frame; a .NET caller's zero-initialised, reused array
    does not.)"""

    def __init__(self, shape, count=2):
        if count < 1:
            raise ValueError("FrameRing needs at least one frame")
        self.shape = tuple(shape)
        self._frames = [np.zeros(self.shape, np.uint8) for _ in range(count)]
        self._pinned = []
        for f in self._frames:
            f.fill(0)  # touch the pages now
            # page-lock it where a device is present (rt_host_pin): the device->host copy then runs at PCIe speed
            if lib().rt_host_pin(C.c_void_p(f.ctypes.data), f.nbytes) == abi.RT_OK:
                self._pinned.append(f)
        self._at = 0

    def __del__(self):
        try:
            for f in self._pinned:
                lib().rt_host_unpin(C.c_void_p(f.ctypes.data))
            self._pinned = []
        except Exception:
            pass

    def next(self):
        f = self._frames[self._at % len(self._frames)]
        self._at += 1
        return f


def _frame_buffer(shape):
    """A fresh output frame (never recycled behind the caller's back; see FrameRing for the opt-in)."""
    return np.empty(shape, np.uint8)


def _as_array(ctype, items):
    """A ctypes array of `items` for the ABI: marshal()'s array is passed through as it is (no per-element copy)."""
    if isinstance(items, C.Array) and items._type_ is ctype and len(items) > 0:
        return items
    return (ctype * max(1, len(items)))(*items)


def device_count():
    return int(lib().rt_device_count())


class SceneHandle:
    """Owner of an RtScene* (rt_scene_create / rt_scene_destroy)."""

    def __init__(self, hittables, textures=(), device=0, keepalive=None):
        self.n_objects = len(hittables)
        self._h = _as_array(RtHittable, hittables)
        self._t = _as_array(RtTexture, textures)
        self._keep = keepalive
        self.device = device
        out = C.c_void_p()
        check(lib().rt_scene_create(self._h, len(hittables), self._t, len(textures), device, C.byref(out)))
        self.ptr = out

    def close(self):
        if getattr(self, "ptr", None):
            lib().rt_scene_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def bvh_nodes(self, which=abi.RT_BVH_SAH):
        n = lib().rt_scene_bvh_node_count(self.ptr, which)
        bounds = np.empty((n, 6))
        right = np.empty(n, np.int32)
        prim = np.empty(n, np.int32)
        if n:
            check(lib().rt_scene_bvh_nodes(self.ptr, which, ptr(bounds), ptr(right), ptr(prim)))
        return bounds, right, prim

    def wide_bvh_check(self):
        """Builds (if need be) and verifies the 8-wide compressed tree on the host; returns its figures."""
        n, depth, ns, mc = _i32(), _i32(), _i32(), C.c_double()
        check(lib().rt_scene_wide_bvh_check(self.ptr, C.byref(n), C.byref(depth), C.byref(ns), C.byref(mc)))
        return {"nodes": n.value, "depth": depth.value, "spheres": ns.value, "mean_children": mc.value}

    def device_bvh_check(self):
        """Verifies on the host the binary tree in the form the render kernels read it; returns its figures."""
        n, depth, ns = _i32(), _i32(), _i32()
        check(lib().rt_scene_device_bvh_check(self.ptr, C.byref(n), C.byref(depth), C.byref(ns)))
        return {"nodes": n.value, "depth": depth.value, "spheres": ns.value}

    def device_bytes(self):
        return int(lib().rt_scene_device_bytes(self.ptr))

    def shared_memory_bytes(self):
        return int(lib().rt_scene_shared_memory_bytes(self.ptr))

    # ---- render ----
    def render(self, camera, max_w, max_h, seed=0, adaptive=True, mode=abi.RT_MODE_MEGAKERNEL, gamma=False, flags=0,
               want_sums=False, rgb_out=None):
        rows, cols = 2 * max_h + 1, 2 * max_w + 1
        rgb = rgb_out if rgb_out is not None else _frame_buffer((rows, cols, 3))
        sums = np.empty((rows, cols, 4), np.int32) if want_sums else None
        opts = RtRenderOpts(seed, int(adaptive), mode, int(gamma), flags)
        stats = RtStats()
        check(lib().rt_render(self.ptr, C.byref(camera), max_w, max_h, C.byref(opts), ptr(rgb), ptr(sums), C.byref(stats)))
        return rgb, sums, stats

    # ---- conformance ----
    def hit_object(self, o, d, traversal=0):
        o, d = f64(o, (-1, 3)), f64(d, (-1, 3))
        n = len(o)
        prim = np.empty(n, np.int32)
        t = np.empty(n)
        strike = np.empty((n, 3))
        check(lib().rt_test_hit_object(self.ptr, traversal, n, ptr(o), ptr(d), ptr(prim), ptr(t), ptr(strike)))
        return prim, t, strike

    def reflection(self, prim, o, d, strike, colour_in, uniforms):
        prim = np.ascontiguousarray(prim, np.int32)
        o, d, strike = f64(o, (-1, 3)), f64(d, (-1, 3)), f64(strike, (-1, 3))
        colour_in = np.ascontiguousarray(colour_in, np.uint8).reshape(-1, 3)
        uniforms = f64(uniforms, (-1, 4))
        n = len(prim)
        absorbed = np.empty(n, np.uint8)
        colour = np.empty((n, 3), np.uint8)
        oo = np.empty((n, 3))
        do = np.empty((n, 3))
        inside = np.empty(n, np.uint8)
        check(lib().rt_test_reflection(self.ptr, n, ptr(prim), ptr(o), ptr(d), ptr(strike), ptr(colour_in), ptr(uniforms), ptr(absorbed),
                                       ptr(colour), ptr(oo), ptr(do), ptr(inside)))
        return absorbed, colour, oo, do, inside

    def texture(self, prim, point):
        prim = np.ascontiguousarray(prim, np.int32)
        point = f64(point, (-1, 3))
        out = np.empty((len(prim), 3), np.uint8)
        check(lib().rt_test_texture(self.ptr, len(prim), ptr(prim), ptr(point), ptr(out)))
        return out

    def trace_samples(self, camera, max_w, max_h, seed, row_idx, col_idx, sample):
        row_idx, col_idx, sample = [np.ascontiguousarray(a, np.int32) for a in (row_idx, col_idx, sample)]
        n = len(row_idx)
        colour = np.empty((n, 3), np.uint8)
        rays = np.empty(n, np.int32)
        check(lib().rt_test_trace_samples(self.ptr, C.byref(camera), max_w, max_h, seed, n, ptr(row_idx), ptr(col_idx), ptr(sample),
                                          ptr(colour), ptr(rays)))
        return colour, rays


class MultiHandle:
    """Owner of an RtMulti* (rt_multi_create / rt_multi_destroy): one frame split over several GPUs of this process."""

    def __init__(self, hittables, textures=(), devices=(0,), keepalive=None):
        self._h = _as_array(RtHittable, hittables)
        self._t = _as_array(RtTexture, textures)
        self._keep = keepalive
        self.devices = list(devices)
        dv = (C.c_int32 * len(self.devices))(*self.devices)
        out = C.c_void_p()
        check(lib().rt_multi_create(self._h, len(hittables), self._t, len(textures), dv, len(self.devices), C.byref(out)))
        self.ptr = out

    def close(self):
        if getattr(self, "ptr", None):
            lib().rt_multi_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def render(self, camera, max_w, max_h, seed=0, adaptive=True, gamma=False, flags=0, want_sums=False, rgb_out=None):
        rows, cols = 2 * max_h + 1, 2 * max_w + 1
        rgb = rgb_out if rgb_out is not None else _frame_buffer((rows, cols, 3))
        sums = np.empty((rows, cols, 4), np.int32) if want_sums else None
        opts = RtRenderOpts(seed, int(adaptive), abi.RT_MODE_MEGAKERNEL, int(gamma), flags)
        stats = RtStats()
        check(lib().rt_multi_render(self.ptr, C.byref(camera), max_w, max_h, C.byref(opts), ptr(rgb), ptr(sums), C.byref(stats)))
        return rgb, sums, stats


class CommHandle:
    """Owner of an RtComm* (rt_comm_create / rt_comm_destroy): this process's rank of a one-process-per-GPU job whose
    collectives the library issues itself (NCCL, bound at run time).  Creation is collective over the ranks."""

    @staticmethod
    def _torch_first():
        """The library binds whatever libnccl.so.2 the process already holds, else the system's.  A process that will
        ALSO import PyTorch must import it first: PyTorch needs the (newer) NCCL build it bundles, and the dynamic loader
        keeps a single library per soname — binding the system's copy first makes `import torch` fail afterwards."""
        import importlib.util
        import sys
        if "torch" not in sys.modules and importlib.util.find_spec("torch") is not None:
            import torch  # noqa: F401

    @staticmethod
    def unique_id() -> bytes:
        CommHandle._torch_first()
        buf = (C.c_uint8 * abi.RT_COMM_ID_BYTES)()
        check(lib().rt_comm_unique_id(buf))
        return bytes(buf)

    @staticmethod
    def nccl_version():
        CommHandle._torch_first()
        v = _i32()
        path = C.create_string_buffer(512)
        check(lib().rt_comm_nccl_version(C.byref(v), path, 512))
        return int(v.value), path.value.decode("utf-8", "replace")

    def __init__(self, unique_id: bytes, rank: int, world: int, device: int, stream=None):
        if unique_id is None and world == 1:
            unique_id = bytes(abi.RT_COMM_ID_BYTES)
        if len(unique_id) != abi.RT_COMM_ID_BYTES:
            raise ValueError("unique_id must be RT_COMM_ID_BYTES long")
        if world > 1:
            CommHandle._torch_first()
        self.rank, self.world, self.device = rank, world, device
        idb = (C.c_uint8 * abi.RT_COMM_ID_BYTES).from_buffer_copy(unique_id)  # ignored by the library when world == 1
        out = C.c_void_p()
        check(lib().rt_comm_create(idb, rank, world, device, C.c_void_p(stream) if stream else None, C.byref(out)))
        self.ptr = out

    def close(self):
        if getattr(self, "ptr", None):
            lib().rt_comm_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def render(self, scene: "SceneHandle", camera, max_w, max_h, seed=0, adaptive=True, gamma=False, flags=0, want_rgb=True,
               want_sums=False, want_stats=True, rgb_out=None):
        """One frame, collectively.  want_rgb / want_sums / want_stats all False: enqueue only (device-resident)."""
        rows, cols = 2 * max_h + 1, 2 * max_w + 1
        rgb = (rgb_out if rgb_out is not None else _frame_buffer((rows, cols, 3))) if want_rgb else None
        sums = np.empty((rows, cols, 4), np.int32) if want_sums else None
        opts = RtRenderOpts(seed, int(adaptive), abi.RT_MODE_MEGAKERNEL, int(gamma), flags)
        stats = RtStats() if (want_stats or want_rgb or want_sums) else None
        check(lib().rt_comm_render(self.ptr, scene.ptr, C.byref(camera), max_w, max_h, C.byref(opts), ptr(rgb), ptr(sums),
                                   C.byref(stats) if stats is not None else None))
        return rgb, sums, stats

    def uses_peer_memory(self) -> bool:
        return bool(lib().rt_comm_uses_peer_memory(self.ptr))

    def last_stats(self, scene: "SceneHandle"):
        """Timing / counters of the last enqueue-only render (call after synchronising the stream)."""
        stats = RtStats()
        check(lib().rt_comm_last_stats(self.ptr, scene.ptr, C.byref(stats)))
        return stats

    def frame_pointers(self):
        rgb, stats = C.c_void_p(), C.c_void_p()
        check(lib().rt_comm_frame(self.ptr, C.byref(rgb), C.byref(stats)))
        return rgb.value, stats.value


# ---- free-standing conformance wrappers ---------------------------------------------------------
def sphere_hit(o, d, c, r, device=0):
    o, d, c, r = f64(o, (-1, 3)), f64(d, (-1, 3)), f64(c, (-1, 3)), f64(r, (-1,))
    t = np.empty(len(o))
    check(lib().rt_test_sphere_hit(device, len(o), ptr(o), ptr(d), ptr(c), ptr(r), ptr(t)))
    return t


def plane_hit(o, d, p, n, device=0):
    o, d, p, n = [f64(a, (-1, 3)) for a in (o, d, p, n)]
    t = np.empty(len(o))
    check(lib().rt_test_plane_hit(device, len(o), ptr(o), ptr(d), ptr(p), ptr(n), ptr(t)))
    return t


def aabb_hit(o, d, bmin, bmax, device=0):
    o, d, bmin, bmax = [f64(a, (-1, 3)) for a in (o, d, bmin, bmax)]
    hit = np.empty(len(o), np.uint8)
    check(lib().rt_test_aabb_hit(device, len(o), ptr(o), ptr(d), ptr(bmin), ptr(bmax), ptr(hit)))
    return hit.astype(bool)


def camera_rays(camera, max_w, max_h, row, col, r1, r2, device=0):
    row, col = np.ascontiguousarray(row, np.int32), np.ascontiguousarray(col, np.int32)
    r1, r2 = f64(r1), f64(r2)
    n = len(row)
    o = np.empty((n, 3))
    d = np.empty((n, 3))
    check(lib().rt_test_camera_rays(device, C.byref(camera), max_w, max_h, n, ptr(row), ptr(col), ptr(r1), ptr(r2), ptr(o), ptr(d)))
    return o, d


def combine_darken(a, b, albedo, device=0):
    a = np.ascontiguousarray(a, np.uint8).reshape(-1, 3)
    b = np.ascontiguousarray(b, np.uint8).reshape(-1, 3)
    albedo = np.ascontiguousarray(np.broadcast_to(f64(albedo), (len(a),)))
    out = np.empty_like(a)
    check(lib().rt_test_combine_darken(device, len(a), ptr(a), ptr(b), ptr(albedo), ptr(out)))
    return out


def rng(seed, pixel, sample, bounce, retry, device=0):
    pixel, sample, bounce, retry = [np.ascontiguousarray(a, np.uint32) for a in (pixel, sample, bounce, retry)]
    n = len(pixel)
    words = np.empty((n, 4), np.uint32)
    u = np.empty((n, 4))
    check(lib().rt_test_rng(device, seed, n, ptr(pixel), ptr(sample), ptr(bounce), ptr(retry), ptr(words), ptr(u)))
    return words, u


def measure_fp32_peak(device=0):
    out = C.c_double()
    check(lib().rt_measure_fp32_peak(device, C.byref(out)))
    return out.value


def camera_make_basic(spp, focal, aspect, origin, view_dir, view_up):
    cam = RtCamera()
    origin, view_dir, view_up = f64(origin), f64(view_dir), f64(view_up)
    check(lib().rt_camera_make_basic(spp, focal, aspect, ptr(origin), ptr(view_dir), ptr(view_up), C.byref(cam)))
    return cam


def ppm_format(rgb, gamma=False):
    rgb = np.ascontiguousarray(rgb, np.uint8)
    rows, cols = rgb.shape[0], rgb.shape[1]
    n = C.c_size_t()
    check(lib().rt_ppm_format(ptr(rgb), rows, cols, int(gamma), None, 0, C.byref(n)))
    buf = C.create_string_buffer(n.value)
    check(lib().rt_ppm_format(ptr(rgb), rows, cols, int(gamma), C.cast(buf, C.c_void_p), n.value, C.byref(n)))
    return buf.raw[:n.value]


def ppm_write_file(rgb, path, gamma=False):
    rgb = np.ascontiguousarray(rgb, np.uint8)
    check(lib().rt_ppm_write_file(ptr(rgb), rgb.shape[0], rgb.shape[1], int(gamma), os.fsencode(path)))


def gamma_correct(b):
    return int(lib().rt_gamma_correct(int(b)))
