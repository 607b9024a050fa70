"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol include/rtfs_b200.h declares, the
ctypes mirror matches the header's struct layouts, the host logic (Camera.makeBasic, P3 writer, gamma, the two
BVH builders, argument validation) agrees with the oracle, and the product package never touches the oracle."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

import oracle
from helpers import small_random_spheres
from ray_tracing_fsharp_b200 import abi, native, sample_images
from ray_tracing_fsharp_b200.domain import (Colour, Hittable, InfinitePlane, InfinitePlaneStyle, ParameterisedTexture, Pixel, Sphere,
                                            SphereStyle, Texture, marshal)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rtfs_b200.h")


def test_library_exports_every_declared_symbol():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", text))
    assert len(declared) >= 30
    lib = native.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared in include/rtfs_b200.h but not exported by librtfs_b200.so"
    assert declared == set(native.EXPORTS), declared ^ set(native.EXPORTS)
    assert lib.rt_abi_version() == abi.RT_ABI_VERSION


def test_ctypes_mirror_matches_header_layout():
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "rtfs_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu\n", sizeof(RtTexture), sizeof(RtHittable), sizeof(RtCamera), sizeof(RtRenderOpts), sizeof(RtStats));
  printf("%zu %zu %zu %zu\n", offsetof(RtTexture, rgb8), offsetof(RtTexture, map_radius), offsetof(RtHittable, texture), offsetof(RtCamera, samples_per_pixel));
  printf("%zu %zu %zu %zu\n", offsetof(RtStats, kernel_ms), offsetof(RtStats, launches), offsetof(RtStats, main_ms), offsetof(RtStats, degenerate_paths));
  return 0; }
'''
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        out = subprocess.check_output([exe], text=True).split()
    got = [int(x) for x in out]
    want = [C.sizeof(abi.RtTexture), C.sizeof(abi.RtHittable), C.sizeof(abi.RtCamera), C.sizeof(abi.RtRenderOpts), C.sizeof(abi.RtStats),
            abi.RtTexture.rgb8.offset, abi.RtTexture.map_radius.offset, abi.RtHittable.texture.offset, abi.RtCamera.samples_per_pixel.offset,
            abi.RtStats.kernel_ms.offset, abi.RtStats.launches.offset, abi.RtStats.main_ms.offset, abi.RtStats.degenerate_paths.offset]
    assert got == want


@pytest.mark.parametrize("config", ["C1", "C2", "C3", "C4"])
def test_camera_make_basic_matches_oracle(config):
    spec = sample_images.CONFIGS[config]()
    a = native.camera_make_basic(spec.spp, spec.focal_length, spec.aspect_ratio, spec.origin, spec.view_direction, spec.view_up)
    b = oracle.camera_make_basic(spec.spp, spec.focal_length, spec.aspect_ratio, spec.origin, spec.view_direction, spec.view_up)
    assert bytes(a) == bytes(b)
    assert a.bounce_depth == 150
    x, y, v = np.array(a.xaxis_dir), np.array(a.yaxis_dir), np.array(a.view_dir)
    assert np.allclose(np.cross(x, y), v, atol=1e-12)  # X x Y = view


def test_camera_make_basic_survey_check_values():
    a = native.camera_make_basic(1, 10.0, 1.5, (13.0, 2.0, -3.0), tuple(-np.array([13.0, 2.0, -3.0]) / np.sqrt(182.0)), (0.0, 1.0, 0.0))
    assert np.allclose(a.xaxis_dir, (0.224860, 0.0, 0.974391), atol=1e-6)
    assert np.allclose(a.yaxis_dir, (-0.144453, 0.988950, 0.033335), atol=1e-6)
    with pytest.raises(native.RtError):  # viewUp parallel to the view direction: the reference throws
        native.camera_make_basic(1, 1.0, 1.0, (0, 0, 0), (0.0, 1.0, 0.0), (0.0, 1.0, 0.0))
    with pytest.raises(native.RtError):
        native.camera_make_basic(1, 1.0, 1.0, (0, 0, 0), (0.0, 2.0, 0.0), (0.0, 0.0, 1.0))


def test_ppm_writer_matches_reference_golden_and_oracle():
    # RayTracing.Test/PpmOutputExample.txt via TestPpmOutput.fs:12-46 (the Wikipedia P3 example)
    px = np.array([[[255, 0, 0], [0, 255, 0], [0, 0, 255]], [[255, 255, 0], [255, 255, 255], [0, 0, 0]]], np.uint8)
    golden = open(os.path.join(ROOT, "tests", "golden", "PpmOutputExample.txt"), "rb").read()
    assert native.ppm_format(px, False) == golden
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (37, 53, 3)).astype(np.uint8)
    for gamma in (False, True):
        assert native.ppm_format(img, gamma) == oracle.ppm_format(img, gamma)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "o.ppm")
        native.ppm_write_file(img, path, True)
        assert open(path, "rb").read() == oracle.ppm_format(img, True)


def test_gamma_correct_all_bytes():
    for b in range(256):
        assert native.gamma_correct(b) == oracle.gamma_correct(b)
    assert native.gamma_correct(0) == 0 and native.gamma_correct(255) == 255 and native.gamma_correct(64) == 128


def _host_scene(spec):
    hs, ts, keep = marshal(spec.objects)
    return hs, ts, native.SceneHandle(hs, ts, device=-1, keepalive=keep)


@pytest.mark.parametrize("which", ["C1", "C2", "C4", "reduced"])
def test_reference_tree_equals_oracle_tree(which):
    spec = small_random_spheres() if which == "reduced" else sample_images.CONFIGS[which]()
    hs, ts, h = _host_scene(spec)
    b1, r1, p1 = h.bvh_nodes(abi.RT_BVH_REFERENCE)
    b2, r2, p2 = oracle.Scene(hs, ts).bvh_nodes()
    assert np.array_equal(r1, r2) and np.array_equal(p1, p2) and np.array_equal(b1, b2)
    n_bounded = sum(1 for x in hs if x.shape == abi.RT_SHAPE_SPHERE)
    assert len(r1) == max(0, 2 * n_bounded - 1)  # BoundingBoxTree: one object per leaf


@pytest.mark.parametrize("which", ["C2", "C4", "reduced", "C5small"])
def test_sah_tree_is_a_valid_bvh(which):
    if which == "C5small":
        spec = sample_images.many_spheres(n=3000)
    else:
        spec = small_random_spheres() if which == "reduced" else sample_images.CONFIGS[which]()
    hs, ts, h = _host_scene(spec)
    bounds, right, prim = h.bvh_nodes(abi.RT_BVH_SAH)
    want = sorted(i for i, x in enumerate(hs) if x.shape == abi.RT_SHAPE_SPHERE and x.radius >= 0)  # F16: r < 0 is never hit
    leaves = sorted(int(p) for p in prim if p >= 0)
    assert leaves == want
    # every node's box encloses its subtree (DFS pre-order: left child = i + 1)
    def check(i):
        if right[i] < 0:
            x = hs[prim[i]]
            # the device intersects the FP32-rounded sphere, so that is what the box must enclose
            c, r = np.array(x.p, np.float32).astype(np.float64), float(np.float32(x.radius))
            assert (bounds[i, :3] <= c - r).all() and (bounds[i, 3:] >= c + r).all()
            return bounds[i, :3], bounds[i, 3:]
        lmn, lmx = check(i + 1)
        rmn, rmx = check(right[i])
        assert (bounds[i, :3] <= np.minimum(lmn, rmn)).all() and (bounds[i, 3:] >= np.maximum(lmx, rmx)).all()
        return bounds[i, :3], bounds[i, 3:]
    import sys
    sys.setrecursionlimit(10000)
    if len(right):
        check(0)


@pytest.mark.parametrize("which", ["reduced", "C1", "C2", "C4", "mid"])
def test_wide_bvh_is_valid_on_the_host(which):
    """rt_scene_wide_bvh_check: the 8-wide compressed tree (RT_FLAG_WIDE_BVH) holds every bounded sphere of non-negative
    radius exactly once and every decoded (quantised) box contains what lies below it; at most 8 children per node."""
    if which == "reduced":
        spec = small_random_spheres()
    elif which == "mid":
        spec = sample_images.many_spheres(n=3000)
    else:
        spec = sample_images.CONFIGS[which]()
    hs, ts, keep = marshal(spec.objects)
    h = native.SceneHandle(hs, ts, -1, keepalive=keep)
    info = h.wide_bvh_check()
    n_spheres = sum(1 for o in spec.objects if isinstance(o, Hittable.Sphere) and o.sphere.Radius >= 0)
    assert info["spheres"] == n_spheres
    assert 1 <= info["nodes"] <= max(1, n_spheres) and 1.0 <= info["mean_children"] <= 8.0
    assert info["depth"] <= 12
    again = h.wide_bvh_check()  # idempotent
    assert again == info


@pytest.mark.parametrize("which", ["reduced", "C1", "C2", "C3", "C4", "mid", "one-sphere"])
def test_device_bvh_is_valid_on_the_host(which):
    """rt_scene_device_bvh_check: the binary tree as the render kernels read it (centre / half-extent boxes, the two children's
    values side by side for the packed slab test): no converted box is smaller than the box it was made from, every bounded
    sphere of non-negative radius is the leaf of exactly one node and inside that leaf's box, depth within the walk stacks."""
    if which == "reduced":
        spec = small_random_spheres()
    elif which == "mid":
        spec = sample_images.many_spheres(n=3000)
    elif which == "one-sphere":
        spec = sample_images.CONFIGS["C3"]()
    else:
        spec = sample_images.CONFIGS[which]()
    hs, ts, keep = marshal(spec.objects)
    h = native.SceneHandle(hs, ts, -1, keepalive=keep)
    info = h.device_bvh_check()
    n_spheres = sum(1 for o in spec.objects if isinstance(o, Hittable.Sphere) and o.sphere.Radius >= 0)
    assert info["spheres"] == n_spheres
    assert info["nodes"] == max(1, n_spheres - 1) if n_spheres else info["nodes"] == 0
    assert (info["depth"] <= 64) and (info["depth"] >= 1 if n_spheres else info["depth"] == 0)


def test_scene_create_validates_like_the_type_system_would():
    lib = native.lib()

    def create(hs, ts=()):
        H = (abi.RtHittable * max(1, len(hs)))(*hs)
        T = (abi.RtTexture * max(1, len(ts)))(*ts)
        out = C.c_void_p()
        rc = lib.rt_scene_create(H, len(hs), T, len(ts), -1, C.byref(out))
        if rc == 0:
            lib.rt_scene_destroy(out)
        return rc, lib.rt_last_error().decode()

    ok = abi.RtHittable()
    ok.shape, ok.style, ok.radius, ok.texture = abi.RT_SHAPE_SPHERE, abi.RT_STYLE_GLASS, 1.0, -1
    assert create([ok])[0] == abi.RT_OK
    bad = abi.RtHittable.from_buffer_copy(ok)
    bad.shape = 9
    assert create([bad])[0] == abi.RT_ERR_INVALID_ARGUMENT
    bad = abi.RtHittable.from_buffer_copy(ok)
    bad.shape, bad.n[1] = abi.RT_SHAPE_INFINITE_PLANE, 1.0  # InfinitePlaneStyle has no Glass case (F13)
    rc, msg = create([bad])
    assert rc == abi.RT_ERR_INVALID_ARGUMENT and "InfinitePlaneStyle" in msg
    bad = abi.RtHittable.from_buffer_copy(ok)
    bad.texture = 3
    assert create([bad])[0] == abi.RT_ERR_INVALID_ARGUMENT
    bad = abi.RtHittable.from_buffer_copy(ok)
    bad.p[0] = float("nan")
    assert create([bad])[0] == abi.RT_ERR_INVALID_ARGUMENT
    t = abi.RtTexture()
    t.kind = 7
    assert create([ok], [t])[0] == abi.RT_ERR_UNSUPPORTED
    assert lib.rt_scene_create(None, 0, None, 0, -1, None) == abi.RT_ERR_INVALID_ARGUMENT

    # checkered textures: cycles (A -> B -> A) and nests deeper than the device follows (8 levels) are refused
    def checker(even, odd):
        c = abi.RtTexture()
        c.kind, c.even, c.odd, c.grid_size, c.map_radius = abi.RT_TEX_CHECKERED, even, odd, 4.0, 1.0
        return c

    def colour():
        c = abi.RtTexture()
        c.kind = abi.RT_TEX_COLOUR
        return c

    rc, msg = create([ok], [checker(1, 2), checker(0, 2), colour()])
    assert rc == abi.RT_ERR_INVALID_ARGUMENT and "cycle" in msg
    chain = [checker(i + 1, 9) for i in range(8)] + [colour(), colour()]  # 8 levels: the deepest the device follows
    assert create([ok], chain)[0] == abi.RT_OK
    chain = [checker(i + 1, 10) for i in range(9)] + [colour(), colour()]  # 9 levels
    rc, msg = create([ok], chain)
    assert rc == abi.RT_ERR_UNSUPPORTED and "nested deeper" in msg


def test_closures_cannot_cross_the_abi():
    s = Hittable.Sphere(Sphere.make(SphereStyle.LambertReflection(1.0, Texture.Arbitrary(lambda p: Colour.Red)), (0, 0, 0), 1.0))
    with pytest.raises(NotImplementedError):
        marshal([s])


def test_bake_samples_a_closure_onto_the_texel_grid_of_the_image_lookup():
    """a ParameterisedTexture.Arbitrary closure becomes an Image whose texel (x, y) holds the
    closure's value at the centre of the (u, v) cell that Texture.fs:63-67 maps to (x, y)."""
    interpret = Sphere.plane_map_inverse(2.0, (1.0, 2.0, 3.0))
    closure = ParameterisedTexture.Arbitrary(lambda u, v: Texture.Colour(Pixel(int(255 * u), int(255 * v), 7)))
    with pytest.raises(NotImplementedError):
        marshal([Hittable.Sphere(Sphere.make(SphereStyle.LambertReflection(1.0, ParameterisedTexture.to_texture(interpret, closure)), (1, 2, 3), 2.0))])
    w, h = 16, 8
    baked = ParameterisedTexture.bake(closure, interpret, w, h)
    assert baked.img.shape == (h, w, 3)
    rng = np.random.default_rng(3)
    for u, v in rng.random((200, 2)):
        x, y = int((1.0 - u) * (w - 1)), int(v * (h - 1))  # the lookup of Texture.fs:63-67
        got = ParameterisedTexture.colour_at(interpret, baked, u, v).as_tuple()
        assert got == tuple(int(c) for c in baked.img[y, x])
        want = closure.f(u, v).pixel.as_tuple()
        assert abs(got[0] - want[0]) <= 255 / (w - 1) + 1 and abs(got[1] - want[1]) <= 255 / (h - 1) + 1 and got[2] == 7
    # a closure that wants the point itself (Texture.Arbitrary): planeMap supplies it, and planeMapInverse inverts planeMap
    by_point = ParameterisedTexture.Arbitrary(lambda u, v: Texture.Arbitrary(lambda p: Pixel(*[int(min(255, abs(c) * 40)) for c in p])))
    p = Sphere.plane_map(2.0, (1.0, 2.0, 3.0), 0.3, 0.6)
    assert np.allclose(oracle.plane_map(2.0, (1.0, 2.0, 3.0), 0.3, 0.6), p)
    assert np.allclose(oracle.plane_map_inverse(2.0, (1.0, 2.0, 3.0), p), (0.3, 0.6))
    assert ParameterisedTexture.colour_at(interpret, by_point, 0.3, 0.6).as_tuple() == tuple(int(min(255, abs(c) * 40)) for c in p)
    # a baked checker is the checker (away from its edges), and the result crosses the ABI
    chk = ParameterisedTexture.Checkered(ParameterisedTexture.Colour(Colour.Red), ParameterisedTexture.Colour(Colour.Blue), 10.0)
    assert ParameterisedTexture.colour_at(interpret, chk, 0.05, 0.05) == Colour.Blue  # sin(.5) sin(.5) > 0: odd
    hs, ts, _ = marshal([Hittable.Sphere(Sphere.make(SphereStyle.LambertReflection(1.0, ParameterisedTexture.to_texture(interpret, baked)), (1, 2, 3), 2.0))])
    assert ts[hs[0].texture].kind == abi.RT_TEX_IMAGE and (ts[hs[0].texture].width, ts[hs[0].texture].height) == (w, h)


@pytest.mark.parametrize("name", sorted(sample_images.REFERENCE_SAMPLES))
def test_reference_sample_scenes_marshal_and_render_on_the_oracle(name):
    """The reference's own sample scenes as scene specs: they cross the ABI's data layout and the oracle renders them
    (a lit, deterministic frame); the GPU parity tests compare against exactly these renders."""
    spec = sample_images.REFERENCE_SAMPLES[name](0.02)
    assert spec.bounce_depth == 150 and spec.spp == 50  # Camera.makeBasic 50 ..., Camera.fs:43
    hs, ts, _keep = marshal(spec.objects)
    assert len(hs) == len(spec.objects)
    cam = oracle.camera_make_basic(spec.spp, spec.focal_length, spec.aspect_ratio, spec.origin, spec.view_direction, spec.view_up)
    cam.bounce_depth = spec.bounce_depth
    a, sa, ca, _ = oracle.Scene(hs, ts).render(cam, spec.max_width_coord, spec.max_height_coord, seed=3, rng_mode=1, adaptive=True)
    b, sb, _, _ = oracle.Scene(hs, ts).render(cam, spec.max_width_coord, spec.max_height_coord, seed=3, rng_mode=1, adaptive=True)
    assert a.shape == (spec.rows, spec.cols, 3) and np.array_equal(sa, sb) and a.max() > 0 and ca["rays"] >= ca["paths"] > 0
    if name == "moved-camera":  # F16: the bounded inner shell (radius -0.45) has an inverted box and is never the closest hit
        n = 4000
        rng = np.random.default_rng(5)
        o = np.tile(np.array(spec.origin, float), (n, 1))
        target = np.array([-1.0, 0.0, 1.0]) + 0.4 * rng.normal(size=(n, 3)) / 3
        d = target - o
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        prim, _, _, _ = oracle.Scene(hs, ts).hit_object(o, d)
        assert (prim == 3).sum() > 1000 and not (prim == 4).any()


def test_output_frames_are_never_recycled_unless_the_caller_opts_in():
    """Every render gets a fresh output array by default (a caller holding only a raw pointer to an earlier frame must not
    see it change); native.FrameRing is the explicit opt-in: `count` arrays handed out in turn."""
    shape = (7, 9, 3)
    a = native._frame_buffer(shape)
    addr = a.ctypes.data
    b = native._frame_buffer(shape)
    assert a is not b and b.ctypes.data != addr and a.shape == shape and a.dtype == np.uint8
    ring = native.FrameRing(shape, count=2)
    f0, f1, f2 = ring.next(), ring.next(), ring.next()
    assert f0 is not f1 and f2 is f0 and f0.shape == shape and f0.dtype == np.uint8
    with pytest.raises(ValueError):
        native.FrameRing(shape, count=0)


def test_marshal_preserves_texture_structure():
    interpret = Sphere.plane_map_inverse(2.0, (1.0, 2.0, 3.0))
    img = np.arange(4 * 6 * 3, dtype=np.uint8).reshape(4, 6, 3)
    tex = ParameterisedTexture.Checkered(ParameterisedTexture.of_image(img), ParameterisedTexture.Colour(Colour.Blue), 5.0)
    objs = [Hittable.Sphere(Sphere.make(SphereStyle.LambertReflection(0.5, ParameterisedTexture.to_texture(interpret, tex)), (1, 2, 3), 2.0)),
            Hittable.InfinitePlane(InfinitePlane.make(InfinitePlaneStyle.PureReflection(0.9, Pixel(1, 2, 3)), (0, 0, 0), (0, 1, 0)))]
    hs, ts, keep = marshal(objs)
    assert len(ts) == 3 and ts[hs[0].texture].kind == abi.RT_TEX_CHECKERED
    chk = ts[hs[0].texture]
    assert ts[chk.even].kind == abi.RT_TEX_IMAGE and ts[chk.odd].kind == abi.RT_TEX_COLOUR
    assert (ts[chk.even].width, ts[chk.even].height) == (6, 4)
    # ofImage flips rows (Texture.fs:34): img[0] is the bottom row of the bitmap
    assert ts[chk.even].rgb8[0] == img[3, 0, 0]
    assert list(chk.map_centre) == [1.0, 2.0, 3.0] and chk.map_radius == 2.0
    assert hs[1].shape == abi.RT_SHAPE_INFINITE_PLANE and list(hs[1].colour) == [1, 2, 3] and hs[1].texture == -1


def test_no_cpu_fallback_without_a_device():
    if native.device_count() > 0:
        pytest.skip("a GPU is present")
    spec = sample_images.few_spheres()
    hs, ts, keep = marshal(spec.objects)
    with pytest.raises(native.RtError) as e:
        native.SceneHandle(hs, ts, 0, keepalive=keep)
    assert e.value.code == abi.RT_ERR_NO_DEVICE
    h = native.SceneHandle(hs, ts, -1, keepalive=keep)  # host-only handle: inspection works, compute does not
    cam = native.camera_make_basic(4, 1.0, 1.0, (0, 0, 0), (0.0, 0.0, 1.0), (0.0, 1.0, 0.0))
    with pytest.raises(native.RtError) as e:
        h.render(cam, 4, 4)
    assert e.value.code == abi.RT_ERR_NO_DEVICE
    with pytest.raises(native.RtError) as e:
        native.sphere_hit([0, 0, 0], [0, 0, 1], [0, 0, 5], [1.0])
    assert e.value.code == abi.RT_ERR_NO_DEVICE
    with pytest.raises(native.RtError) as e:
        native.MultiHandle(hs, ts, [0, 1])
    assert e.value.code == abi.RT_ERR_NO_DEVICE


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "ray_tracing_fsharp_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text and "oracle.cpp" not in text, f
    # and the shared library does not link it
    out = subprocess.check_output(["ldd", native.LIB_PATH], text=True)
    assert "oracle" not in out


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm: oracle with all host threads on a bounded sample) — schema check."""
    import json
    import sys
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                                   "--cpu-seconds", "0.5", "--config", "C1"], text=True, timeout=300)
    lines = [l for l in out.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "sample" in d["config"]
