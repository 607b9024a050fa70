"""ctypes front-end of the CPU oracle (oracle/oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package never does (tests/test_layout.py greps for that).
"""
import ctypes as C
import os
import subprocess

import numpy as np

from ray_tracing_fsharp_b200.abi import RtCamera, RtHittable, RtTexture

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")


def build(force=False):
    """Compile oracle.cpp with g++ (oracle/Makefile)."""
    src = os.path.join(_HERE, "oracle.cpp")
    hdr = os.path.join(_HERE, "..", "include", "rtfs_b200.h")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-s", "-B" if force else "-s"])
    return _LIB_PATH


class OrcCounters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("paths", "rays", "box_tests", "sphere_tests", "plane_tests", "candidates")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_scene_create.restype = C.c_void_p
        _lib.orc_ppm_format.restype = C.c_size_t
        _lib.orc_gamma_correct.restype = C.c_uint8
        _lib.orc_gamma_correct.argtypes = [C.c_uint8]
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


def _vp(a):
    return C.c_void_p(a.ctypes.data)


# ---- RNG -------------------------------------------------------------------------------------
def xorshift_words(state, n):
    st = np.array(state, dtype=np.uint32)
    raw = np.empty(n, np.uint32)
    u = np.empty(n, np.float64)
    lib().orc_xorshift_words(_vp(st), C.c_int(n), _vp(raw), _vp(u))
    return raw, u, st


def philox4x32_10(ctr, key):
    c = np.array(ctr, dtype=np.uint32)
    k = np.array(key, dtype=np.uint32)
    out = np.empty(4, np.uint32)
    lib().orc_philox4x32_10(_vp(c), _vp(k), _vp(out))
    return out


def counter_uniforms(seed, pixel, sample, bounce, retry):
    pixel, sample, bounce, retry = [np.ascontiguousarray(a, dtype=np.uint32) for a in (pixel, sample, bounce, retry)]
    n = len(pixel)
    words = np.empty((n, 4), np.uint32)
    u = np.empty((n, 4), np.float64)
    lib().orc_counter_uniforms(C.c_uint64(seed), C.c_int(n), _vp(pixel), _vp(sample), _vp(bounce), _vp(retry),
                               _vp(words), _vp(u))
    return words, u


# ---- vectors ---------------------------------------------------------------------------------
def unitise(v):
    v = _d(v)
    out = np.empty(3)
    ok = lib().orc_unitise(_vp(v), _vp(out))
    return out if ok else None


def unit_random_explicit(u3):
    u3 = _d(u3).reshape(-1, 3)
    out = np.empty_like(u3)
    lib().orc_unit_random_explicit(C.c_int(len(u3)), _vp(u3), _vp(out))
    return out


def walk_along(o, d, m):
    o, d = _d(o), _d(d)
    out = np.empty(3)
    lib().orc_walk_along(_vp(o), _vp(d), C.c_double(m), _vp(out))
    return out


def plane_orthonormal_basis(origin, v1, v2, up):
    origin, v1, v2, up = _d(origin), _d(v1), _d(v2), _d(up)
    x = np.empty(3)
    y = np.empty(3)
    ok = lib().orc_plane_orthonormal_basis(_vp(origin), _vp(v1), _vp(v2), _vp(up), _vp(x), _vp(y))
    return (x, y) if ok else None


# ---- primitives ------------------------------------------------------------------------------
def sphere_hit(o, d, c, r):
    o, d, c, r = _d(o).reshape(-1, 3), _d(d).reshape(-1, 3), _d(c).reshape(-1, 3), _d(r).reshape(-1)
    t = np.empty(len(o))
    lib().orc_sphere_hit(C.c_int(len(o)), _vp(o), _vp(d), _vp(c), _vp(r), _vp(t))
    return t


def plane_hit(o, d, p, n):
    o, d, p, n = [_d(a).reshape(-1, 3) for a in (o, d, p, n)]
    t = np.empty(len(o))
    lib().orc_plane_hit(C.c_int(len(o)), _vp(o), _vp(d), _vp(p), _vp(n), _vp(t))
    return t


def aabb_hit(o, d, bmin, bmax):
    o, d, bmin, bmax = [_d(a).reshape(-1, 3) for a in (o, d, bmin, bmax)]
    hit = np.empty(len(o), np.uint8)
    lib().orc_aabb_hit(C.c_int(len(o)), _vp(o), _vp(d), _vp(bmin), _vp(bmax), _vp(hit))
    return hit.astype(bool)


def plane_map(radius, centre, phi, theta):
    centre = _d(centre)
    out = np.empty(3)
    lib().orc_plane_map(C.c_double(radius), _vp(centre), C.c_double(phi), C.c_double(theta), _vp(out))
    return out


def plane_map_inverse(radius, centre, p):
    centre, p = _d(centre), _d(p)
    uv = np.empty(2)
    lib().orc_plane_map_inverse(C.c_double(radius), _vp(centre), _vp(p), _vp(uv))
    return uv


def combine(a, b):
    a = np.ascontiguousarray(a, np.uint8).reshape(-1, 3)
    b = np.ascontiguousarray(b, np.uint8).reshape(-1, 3)
    out = np.empty_like(a)
    lib().orc_combine(C.c_int(len(a)), _vp(a), _vp(b), _vp(out))
    return out


def darken(albedo, a):
    a = np.ascontiguousarray(a, np.uint8).reshape(-1, 3)
    albedo = np.ascontiguousarray(np.broadcast_to(_d(albedo), (len(a),)))
    out = np.empty_like(a)
    lib().orc_darken(C.c_int(len(a)), _vp(albedo), _vp(a), _vp(out))
    return out


def gamma_correct(b):
    return int(lib().orc_gamma_correct(C.c_uint8(int(b))))


def stats_mean(stats4):
    s = np.ascontiguousarray(stats4, np.int32)
    out = np.empty(3, np.uint8)
    lib().orc_stats_mean(_vp(s), _vp(out))
    return out


def ppm_format(rgb, gamma=False):
    rgb = np.ascontiguousarray(rgb, np.uint8)
    rows, cols = rgb.shape[0], rgb.shape[1]
    n = lib().orc_ppm_format(_vp(rgb), C.c_int(rows), C.c_int(cols), C.c_int(int(gamma)), None, C.c_size_t(0))
    buf = C.create_string_buffer(n)
    lib().orc_ppm_format(_vp(rgb), C.c_int(rows), C.c_int(cols), C.c_int(int(gamma)), buf, C.c_size_t(n))
    return buf.raw[:n]


# ---- camera ----------------------------------------------------------------------------------
def camera_make_basic(spp, focal, aspect, origin, view_dir, view_up):
    cam = RtCamera()
    origin, view_dir, view_up = _d(origin), _d(view_dir), _d(view_up)
    ok = lib().orc_camera_make_basic(C.c_int(spp), C.c_double(focal), C.c_double(aspect), _vp(origin), _vp(view_dir),
                                     _vp(view_up), C.byref(cam))
    if not ok:
        raise ValueError("Camera.makeBasic: degenerate basis (the reference would throw)")
    return cam


def camera_rays(cam, max_w, max_h, row, col, r1, r2):
    row = np.ascontiguousarray(row, np.int32)
    col = np.ascontiguousarray(col, np.int32)
    r1, r2 = _d(r1), _d(r2)
    n = len(row)
    o = np.empty((n, 3))
    d = np.empty((n, 3))
    lib().orc_camera_rays(C.byref(cam), C.c_int(max_w), C.c_int(max_h), C.c_int(n), _vp(row), _vp(col), _vp(r1), _vp(r2),
                          _vp(o), _vp(d))
    return o, d


# ---- scene -----------------------------------------------------------------------------------
class Scene:
    """Scene.make over an array of RtHittable (+ RtTexture)."""

    def __init__(self, hittables, textures=()):
        self.n = len(hittables)
        self._h = (RtHittable * max(1, len(hittables)))(*hittables)
        self._t = (RtTexture * max(1, len(textures)))(*textures)
        self._keep = list(textures)
        self._ptr = C.c_void_p(lib().orc_scene_create(self._h, C.c_int(len(hittables)), self._t, C.c_int(len(textures))))

    def __del__(self):
        try:
            if self._ptr:
                lib().orc_scene_destroy(self._ptr)
                self._ptr = None
        except Exception:
            pass

    def bvh_nodes(self):
        n = lib().orc_scene_bvh_node_count(self._ptr)
        bounds = np.empty((n, 6))
        right = np.empty(n, np.int32)
        prim = np.empty(n, np.int32)
        if n:
            lib().orc_scene_bvh_nodes(self._ptr, _vp(bounds), _vp(right), _vp(prim))
        return bounds, right, prim

    def hit_object(self, o, d):
        o, d = _d(o).reshape(-1, 3), _d(d).reshape(-1, 3)
        n = len(o)
        prim = np.empty(n, np.int32)
        t = np.empty(n)
        strike = np.empty((n, 3))
        cn = OrcCounters()
        lib().orc_hit_object(self._ptr, C.c_int(n), _vp(o), _vp(d), _vp(prim), _vp(t), _vp(strike), C.byref(cn))
        return prim, t, strike, cn.as_dict()

    def all_hits(self, o, d):
        o, d = _d(o), _d(d)
        t = np.empty(self.n)
        lib().orc_all_hits(self._ptr, _vp(o), _vp(d), _vp(t))
        return t

    def reflection(self, prim, o, d, strike, colour_in, uniforms):
        prim = np.ascontiguousarray(prim, np.int32)
        o, d, strike = [_d(a).reshape(-1, 3) for a in (o, d, strike)]
        colour_in = np.ascontiguousarray(colour_in, np.uint8).reshape(-1, 3)
        uniforms = _d(uniforms).reshape(-1, 4)
        n = len(prim)
        absorbed = np.empty(n, np.uint8)
        colour = np.empty((n, 3), np.uint8)
        oo = np.empty((n, 3))
        do = np.empty((n, 3))
        inside = np.empty(n, np.uint8)
        lib().orc_reflection(self._ptr, C.c_int(n), _vp(prim), _vp(o), _vp(d), _vp(strike), _vp(colour_in), _vp(uniforms),
                             _vp(absorbed), _vp(colour), _vp(oo), _vp(do), _vp(inside))
        return absorbed, colour, oo, do, inside

    def texture(self, prim, point):
        prim = np.ascontiguousarray(prim, np.int32)
        point = _d(point).reshape(-1, 3)
        out = np.empty((len(prim), 3), np.uint8)
        lib().orc_texture(self._ptr, C.c_int(len(prim)), _vp(prim), _vp(point), _vp(out))
        return out

    def trace_samples(self, cam, max_w, max_h, seed, row_idx, col_idx, sample):
        row_idx, col_idx, sample = [np.ascontiguousarray(a, np.int32) for a in (row_idx, col_idx, sample)]
        n = len(row_idx)
        colour = np.empty((n, 3), np.uint8)
        rays = np.empty(n, np.int32)
        lib().orc_trace_samples(self._ptr, C.byref(cam), C.c_int(max_w), C.c_int(max_h), C.c_uint64(seed), C.c_int(n),
                                _vp(row_idx), _vp(col_idx), _vp(sample), _vp(colour), _vp(rays))
        return colour, rays

    def render(self, cam, max_w, max_h, seed=0, rng_mode=1, adaptive=True, threads=None, row_begin=0, row_step=1):
        """Scene.render + Image.render.  Returns (rgb[rows,cols,3], stats[rows,cols,4], counters, rows_rendered)."""
        rows, cols = 2 * max_h + 1, 2 * max_w + 1
        rgb = np.zeros((rows, cols, 3), np.uint8)
        stats = np.zeros((rows, cols, 4), np.int32)
        cn = OrcCounters()
        threads = threads or os.cpu_count() or 1
        done = lib().orc_render(self._ptr, C.byref(cam), C.c_int(max_w), C.c_int(max_h), C.c_uint64(seed),
                                C.c_int(rng_mode), C.c_int(int(adaptive)), C.c_int(threads), C.c_int(row_begin),
                                C.c_int(row_step), _vp(rgb), _vp(stats), C.byref(cn))
        return rgb, stats, cn.as_dict(), done


def render_split(scene, cam, max_w, max_h, seed, adaptive, phase, rank, world, stats, flags):
    """One rank's share of one phase of the sample-split frame (orc_render_split); stats/flags are updated in place."""
    assert stats.dtype == np.int32 and flags.dtype == np.uint8 and stats.flags["C_CONTIGUOUS"] and flags.flags["C_CONTIGUOUS"]
    lib().orc_render_split(scene._ptr, C.byref(cam), C.c_int(max_w), C.c_int(max_h), C.c_uint64(seed), C.c_int(int(adaptive)),
                           C.c_int(phase), C.c_int(rank), C.c_int(world), _vp(stats), _vp(flags))


def sphere_reflection_direct(style, albedo, tex_colour, ior, prob, fuzz, centre, radius, o, d, strike, colour_in,
                             uniforms):
    """Sphere.reflection called with explicit parameters, as TestSphere.fs:52-152 calls it."""
    tex_colour = np.ascontiguousarray(tex_colour, np.uint8)
    colour_in = np.ascontiguousarray(colour_in, np.uint8)
    centre, o, d, strike, uniforms = [_d(a) for a in (centre, o, d, strike, uniforms)]
    colour = np.empty(3, np.uint8)
    oo = np.empty(3)
    do = np.empty(3)
    absorbed = lib().orc_sphere_reflection_direct(
        C.c_int(style), C.c_double(albedo), _vp(tex_colour), C.c_double(ior), C.c_double(prob), C.c_double(fuzz),
        _vp(centre), C.c_double(radius), _vp(o), _vp(d), _vp(strike), _vp(colour_in), _vp(uniforms), _vp(colour),
        _vp(oo), _vp(do))
    return bool(absorbed), colour, oo, do
