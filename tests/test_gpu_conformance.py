"""Level-1 parity: each device function of the render path against the CPU oracle on the
same seeded inputs, through the C ABI.  Bars: integer / byte / index results bit-exact; hit t, normals and
scattered directions within 1e-5 relative; box and closest-primitive decisions identical on vectors whose
double-precision margin exceeds FP32 rounding (the number filtered out is asserted to be small).
Inputs are rounded to FP32 first so that both sides see identical numbers."""
import numpy as np
import pytest

import oracle
from ray_tracing_fsharp_b200 import abi, native, sample_images
from helpers import camera_sample_rays, f32, fp32_safe_closest_hit, random_unit_vectors, scene_pair, small_random_spheres, unit
from ray_tracing_fsharp_b200.domain import marshal

pytestmark = pytest.mark.gpu

REL = 1e-5  # north_star: hit t, normal, scattered direction within 1e-5 relative


def test_library_sees_a_device():
    assert native.device_count() >= 1


# ---- counter RNG ------------------------------------------------------------------------------------
def test_rng_words_bit_exact_and_uniforms():
    rng = np.random.default_rng(1)
    n = 50_000
    pixel, sample, bounce, retry = [rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32) for _ in range(4)]
    for seed in (0, 1, 0xDEADBEEFCAFEF00D):
        words, u = native.rng(seed, pixel, sample, bounce, retry)
        ow, ou = oracle.counter_uniforms(seed, pixel, sample, bounce, retry)
        assert np.array_equal(words, ow)
        assert np.abs(u - ou).max() <= 1.2e-7  # float(w) * 2^-32 against double w / (2^32 - 1)
        assert u.min() >= 0.0 and u.max() <= 1.0


def test_philox_known_answer():
    # Random123 known-answer vectors for philox4x32-10
    words, _ = native.rng(0, [0], [0], [0], [0])
    assert [hex(x) for x in words[0]] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    seed = 0xFFFFFFFFFFFFFFFF
    words, _ = native.rng(seed, [0xFFFFFFFF], [0xFFFFFFFF], [0xFFFFFFFF], [0xFFFFFFFF])
    assert [hex(x) for x in words[0]] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]


# ---- colour -----------------------------------------------------------------------------------------
def test_combine_darken_exhaustive_bit_exact():
    rng = np.random.default_rng(2)
    albedos = np.concatenate([[0.0, 0.5, 1.0, 0.25, 0.75, 0.9, 0.95, 0.1, 1.0 / 3.0, 0.3, 0.7], rng.random(21), rng.random(8) * rng.random(8)])
    a, b = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    pa = np.stack([a.ravel(), b.ravel(), a.ravel()], 1)
    pb = np.stack([b.ravel(), a.ravel()[::-1], np.full(a.size, 255, np.uint8)], 1)
    for al in albedos:
        got = native.combine_darken(pa, pb, al)
        want = oracle.darken(al, oracle.combine(pa, pb))
        assert np.array_equal(got, want), f"albedo {al}"


# ---- Sphere.firstIntersection --------------------------------------------------------------------------
def test_sphere_hit_against_oracle():
    rng = np.random.default_rng(3)
    n = 200_000
    c = f32(rng.uniform(-20, 20, (n, 3)))
    r = f32(rng.uniform(0.05, 5.0, n) * rng.choice([1.0, 1.0, 1.0, -1.0], n))
    o = f32(rng.uniform(-25, 25, (n, 3)))
    aim = c + rng.normal(size=(n, 3)) * np.abs(r)[:, None] * 0.9  # most rays pass near the sphere
    d = unit(aim - o)  # unit in double for the oracle (its formulas assume |d| = 1); the library rounds it to FP32
    want = oracle.sphere_hit(o, d, c, r)
    got = native.sphere_hit(o, d, c, r)
    oc = o - c
    b = (d * oc).sum(1)
    disc = b * b - ((oc * oc).sum(1) - r * r)
    dist = np.sqrt((oc * oc).sum(1))
    # FP32-safe margin: not grazing, origin not within 1 % of the surface (the t ~ 0 cancellation regime)
    safe = (np.abs(disc) > 1e-3 * r * r) & (np.abs(dist - np.abs(r)) > 1e-2 * np.abs(r))
    assert safe.mean() > 0.9
    hit_w, hit_g = ~np.isnan(want), ~np.isnan(got)
    assert np.array_equal(hit_w[safe], hit_g[safe])
    both = safe & hit_w
    assert both.sum() > 50_000
    rel = np.abs(got[both] - want[both]) / want[both]
    assert rel.max() <= REL, rel.max()
    # unfiltered: decisions may differ only where the margin is tiny
    assert (hit_w != hit_g).mean() < 1e-3


def test_sphere_hit_reference_regression_case():
    # TestSphereIntersection.fs:37-57
    o = f32([1.462205539, -4.888279676, 7.123293244])
    d = unit([-9.549697616, 4.400018428, 10.41024923])
    c = f32([-5.688391601, -5.360125644, 9.074300761])
    r = f32([8.199747973])
    want = oracle.sphere_hit(o, d, c, r)
    got = native.sphere_hit(o, d, c, r)
    assert not np.isnan(want[0]) and abs(got[0] - want[0]) / want[0] <= REL
    p = o + got[0] * d
    assert abs(np.sqrt(((p - c) ** 2).sum()) - r[0]) < 1e-5 * r[0]


# ---- InfinitePlane.intersection ----------------------------------------------------------------------------
def test_plane_hit_against_oracle():
    rng = np.random.default_rng(4)
    n = 100_000
    o, p = f32(rng.uniform(-50, 50, (n, 3))), rng.uniform(-50, 50, (n, 3))
    d, nrm = f32(random_unit_vectors(rng, n)), f32(random_unit_vectors(rng, n))
    want = oracle.plane_hit(o, d, p, nrm)
    got = native.plane_hit(o, d, p, nrm)
    assert np.array_equal(np.isnan(want), np.isnan(got))
    hit = ~np.isnan(want)
    assert hit.sum() > 30_000
    # Two evaluations on the device (rtfs_internal.h DUnbounded, classify_unbounded): planes with |n.p0| <= 16 take the
    # FP32 expanded form k - n.o, the others the FP64 numerator and denominator; the quotient is FP32 in both.
    k = (p * nrm).sum(1)
    wide = hit & (np.abs(k) > 16.0)
    assert wide.sum() > 20_000
    assert (np.abs(got[wide] - want[wide]) / want[wide]).max() <= 1e-6
    # FP32 class: k - n.o rounds at 6e-8 of |k| + |n.o| and n.d at 6e-8, so the bar of 1e-5 holds wherever the ray neither
    # starts within 3 % of the plane (relative to those magnitudes) nor runs within 1e-2 of parallel to it; the filtered
    # share is asserted small
    near = hit & ~wide
    height = np.abs(((p - o) * nrm).sum(1))
    safe = near & (height > 3e-2 * (np.abs(k) + np.abs((o * nrm).sum(1)))) & (np.abs((d * nrm).sum(1)) > 1e-2)
    assert near.sum() > 5_000 and safe.sum() > 0.9 * near.sum(), (near.sum(), safe.sum())
    assert (np.abs(got[safe] - want[safe]) / want[safe]).max() <= REL
    assert np.quantile(np.abs(got[near] - want[near]) / want[near], 0.99) <= REL
    # parallel ray and ray pointing away
    assert np.isnan(native.plane_hit([0, 1, 0], [1, 0, 0], [0, 0, 0], [0, 1, 0])[0])
    assert np.isnan(native.plane_hit([0, 1, 0], [0, 1, 0], [0, 0, 0], [0, 1, 0])[0])
    assert native.plane_hit([0, 1, 0], [0, -1, 0], [0, 0, 0], [0, 1, 0])[0] == 1.0


# ---- BoundingBox.hits -------------------------------------------------------------------------------------
def _slab_interval(o, d, bmin, bmax):
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / d
        t0, t1 = (bmin - o) * inv, (bmax - o) * inv
    lo, hi = np.minimum(t0, t1), np.maximum(t0, t1)
    return np.nanmax(lo, 1), np.nanmin(hi, 1)


def test_aabb_decisions_identical_where_fp32_safe():
    rng = np.random.default_rng(5)
    n = 300_000
    centre = rng.uniform(-20, 20, (n, 3))
    half = rng.uniform(0.05, 6.0, (n, 3))
    bmin, bmax = f32(centre - half), f32(centre + half)
    o = f32(rng.uniform(-30, 30, (n, 3)))
    aim = centre + rng.normal(size=(n, 3)) * half * 1.2
    d = f32(unit(aim - o))
    want = oracle.aabb_hit(o, d, bmin, bmax)
    got = native.aabb_hit(o, d, bmin, bmax)
    tmin, tmax = _slab_interval(o, d, bmin, bmax)
    scale = np.abs(o).max(1) + np.abs(bmax).max(1) + 1.0
    safe = (np.abs(tmax - tmin) > 1e-4 * scale) & (np.abs(tmax) > 1e-4 * scale)
    assert safe.mean() > 0.97
    assert 0.2 < want.mean() < 0.95
    assert np.array_equal(want[safe], got[safe])
    assert (want != got).mean() < 1e-4


DELTA = 0.00000001


def _sort(x1, x2):
    return min(x1, x2), (x1 + DELTA / 2.0 if x1 == x2 else max(x1, x2))


@pytest.mark.parametrize("axis", [0, 1, 2])
@pytest.mark.parametrize("negate", [True, False])
def test_bounding_box_behind_ray_is_not_hit(axis, negate):  # TestBoundingBox.fs:16-44, :46-73, :86-114
    rng = np.random.default_rng(6 + axis)
    n = 3000
    vals = rng.normal(size=(6, n)) * 100.0
    o = np.zeros(3)
    o[axis] = -DELTA if negate else DELTA
    d = np.zeros(3)
    d[axis] = -1.0 if negate else 1.0
    bmin, bmax = np.zeros((n, 3)), np.zeros((n, 3))
    for i in range(n):
        lo, hi = _sort(abs(vals[0, i]) if negate else -abs(vals[0, i]), abs(vals[1, i]) if negate else -abs(vals[1, i]))
        bmin[i, axis], bmax[i, axis] = lo, hi
        k = 2
        for ax in range(3):
            if ax != axis:
                bmin[i, ax], bmax[i, ax] = _sort(vals[k, i], vals[k + 1, i])
                k += 2
    got = native.aabb_hit(np.tile(o, (n, 1)), np.tile(d, (n, 1)), bmin, bmax)
    assert not got.any()
    assert not oracle.aabb_hit(np.tile(o, (n, 1)), np.tile(d, (n, 1)), bmin, bmax).any()


def test_bounding_box_reference_fixed_cases():
    # TestBoundingBox.fs:75-84: degenerate box touching the origin plane, 0 * inf = NaN on two axes
    z1, z2 = _sort(-abs(0.0), -abs(0.0))
    x1, x2 = _sort(0.0, 0.0)
    y1, y2 = _sort(0.0, 1.0)
    assert not native.aabb_hit([0.0, 0.0, DELTA], [0.0, 0.0, 1.0], [x1, y1, z1], [x2, y2, z2])[0]
    # :116-123
    assert native.aabb_hit([0, 0, 0], [0, 0, 1], [-1, -1, -1], [1, 1, 1])[0]
    # axis-parallel rays (inverse direction +-inf) inside and outside the other slabs
    assert native.aabb_hit([0.5, 0.5, -3], [0, 0, 1], [0, 0, 0], [1, 1, 1])[0]
    assert not native.aabb_hit([1.5, 0.5, -3], [0, 0, 1], [0, 0, 0], [1, 1, 1])[0]
    assert not native.aabb_hit([0.5, 0.5, 3], [0, 0, 1], [0, 0, 0], [1, 1, 1])[0]


# ---- camera ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("config", ["C1", "C2", "C3", "C4"])
def test_camera_rays(config):
    spec = sample_images.CONFIGS[config]()
    cam = native.camera_make_basic(spec.spp, spec.focal_length, spec.aspect_ratio, spec.origin, spec.view_direction, spec.view_up)
    rng = np.random.default_rng(7)
    n = 50_000
    row = rng.integers(-spec.max_height_coord - 1, spec.max_height_coord, n).astype(np.int32)
    col = rng.integers(-spec.max_width_coord, spec.max_width_coord + 1, n).astype(np.int32)
    r1, r2 = f32(rng.random(n)), f32(rng.random(n))
    wo, wd = oracle.camera_rays(cam, spec.max_width_coord, spec.max_height_coord, row, col, r1, r2)
    go, gd = native.camera_rays(cam, spec.max_width_coord, spec.max_height_coord, row, col, r1, r2)
    assert np.abs(go - wo).max() <= 1e-6 * (1 + np.abs(wo).max())
    assert np.abs(gd - wd).max() <= REL
    assert np.abs((gd * gd).sum(1) - 1).max() < 1e-6


# ---- Scene.hitObject -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("traversal", [0, 1, 2])
@pytest.mark.parametrize("which", ["C2", "C4", "reduced"])
def test_hit_object_matches_oracle(which, traversal):
    """traversal 0: the binary SAH tree, ordered and culled; 1: the reference's own exhaustive DFS; 2: the 8-wide compressed tree.
    Scene.hitObject (Scene.fs:62-91): the closest primitive is IDENTICAL to the oracle's on every ray whose
    double-precision margin exceeds FP32 rounding (helpers.fp32_safe_closest_hit: nothing grazed in front of the
    winner, no origin on a surface, winner leading the runner-up by > 1e-4); the filtered share is asserted small."""
    spec = small_random_spheres() if which == "reduced" else sample_images.CONFIGS[which]()
    osc, dsc, cam = scene_pair(spec)
    hs, _ts, _keep = marshal(spec.objects)
    rng = np.random.default_rng(8)
    n = 40_000
    o1, d1 = camera_sample_rays(spec, cam, rng, n // 2)
    # secondary-like rays: from points near the ground, random directions
    o2 = f32(np.stack([rng.uniform(-8, 8, n // 2), rng.uniform(0.45, 2.5, n // 2), rng.uniform(-8, 8, n // 2)], 1))
    d2 = random_unit_vectors(rng, n // 2)
    o, d = np.concatenate([o1, o2]), np.concatenate([d1, d2])
    wp, wt, ws, counters = osc.hit_object(o, d)
    gp, gt, gs = dsc.hit_object(o, d, traversal=traversal)
    safe = fp32_safe_closest_hit(hs, o, d, wp, wt)
    filtered = 1.0 - safe.mean()
    print(f"hit_object {which} traversal={traversal}: {int((~safe).sum())} of {n} rays filtered as not FP32-safe ({filtered:.2%})")
    assert filtered < 0.12, filtered
    same = wp == gp
    assert np.array_equal(wp[safe], gp[safe]), f"{int((~same & safe).sum())} FP32-safe rays disagree with the oracle"
    # unfiltered: decisions may differ only where the margin is tiny
    assert same.mean() > 0.999, same.mean()
    hit = safe & (wp >= 0)
    assert hit.sum() > 0.5 * n
    rel = np.abs(gt[hit] - wt[hit]) / wt[hit]
    assert rel.max() <= REL, rel.max()  # t within 1e-5 relative on every FP32-safe ray
    assert np.median(rel) < 1e-6
    err = np.abs(gs[hit] - ws[hit]).max(1) / (1.0 + np.abs(ws[hit]).max(1))
    assert err.max() <= REL


def test_negative_radius_bounded_sphere_is_never_hit():
    """F16: Sphere.make builds an inverted box for r < 0, so the reference never intersects it."""
    spec = sample_images.mixed_planes()
    osc, dsc, cam = scene_pair(spec)
    idx = [i for i, ob in enumerate(spec.objects) if getattr(ob, "sphere", None) is not None and ob.sphere.Radius < 0
           and type(ob).__name__ == "Sphere"]
    assert len(idx) == 1
    c = np.array(spec.objects[idx[0]].sphere.Centre)
    rng = np.random.default_rng(9)
    o = f32(c + unit(rng.normal(size=(2000, 3))) * 3.0)
    d = unit(c - o + rng.normal(size=(2000, 3)) * 0.05)
    wp, _, _, _ = osc.hit_object(o, d)
    for traversal in (0, 1):
        gp, _, _ = dsc.hit_object(o, d, traversal=traversal)
        assert not (gp == idx[0]).any() and not (wp == idx[0]).any()
        assert (gp == wp).mean() > 0.999


# ---- Hittable.Reflection -----------------------------------------------------------------------------------------
def _reflection_vectors(spec, osc, cam, rng, n):
    """Hit points of primary and random rays with the primitive they hit, as inputs for the scatter test."""
    o1, d1 = camera_sample_rays(spec, cam, rng, n)
    o2 = f32(np.stack([rng.uniform(-6, 6, n), rng.uniform(0.5, 3, n), rng.uniform(-6, 6, n)], 1))
    d2 = random_unit_vectors(rng, n)
    o, d = np.concatenate([o1, o2]), np.concatenate([d1, d2])
    prim, t, strike, _ = osc.hit_object(o, d)
    keep = prim >= 0
    return o[keep], d[keep], prim[keep], f32(strike[keep])


@pytest.mark.parametrize("which", ["C1", "C2", "C3", "C4", "reduced"])
def test_reflection_matches_oracle(which):
    spec = small_random_spheres() if which == "reduced" else sample_images.CONFIGS[which]()
    osc, dsc, cam = scene_pair(spec)
    rng = np.random.default_rng(10)
    o, d, prim, strike = _reflection_vectors(spec, osc, cam, rng, 20_000)
    n = len(o)
    colour_in = rng.integers(0, 256, (n, 3)).astype(np.uint8)
    uniforms = f32(rng.random((n, 4)))
    wa, wc, wo, wd, wi = osc.reflection(prim, o, d, strike, colour_in, uniforms)
    ga, gc, go, gd, gi = dsc.reflection(prim, o, d, strike, colour_in, uniforms)
    assert np.array_equal(wa, ga)              # absorbed / continue / error
    assert np.array_equal(wi, gi)              # inside flag (F10)
    cont = wa == 0
    # a stochastic branch (Dielectric u > prob, Glass u < Schlick) can flip when u is within FP32 rounding of the
    # threshold, and the texel index can flip on a texel edge: allow a tiny fraction, everything else is exact
    colour_same = (wc == gc).all(1)
    assert colour_same.mean() > 0.9995, colour_same.mean()
    derr = np.abs(gd[cont] - wd[cont]).max(1)
    assert (derr <= REL).mean() > 0.9995, (derr <= REL).mean()
    assert np.abs(go[cont] - wo[cont]).max() <= 1e-6 * (1 + np.abs(wo).max())
    assert np.abs((gd[cont] ** 2).sum(1) - 1).max() < 1e-6


def test_glass_and_dielectric_known_answers():
    """TestSphere.fs:52-152 on the device: grazing glass reflects (direction unchanged), head-on passes straight through."""
    from ray_tracing_fsharp_b200.domain import Colour, Hittable, Sphere, SphereStyle, Texture, marshal
    objs = [Hittable.Sphere(Sphere.make(SphereStyle.Glass(1.0, Texture.Colour(Colour.Green), 1.5), (0.0, 0.0, 0.0), 1.0)),
            Hittable.Sphere(Sphere.make(SphereStyle.Dielectric(1.0, Texture.Colour(Colour.Green), 1.5, 1.0), (10.0, 0.0, 0.0), 1.0))]
    hs, ts, keep = marshal(objs)
    dsc = native.SceneHandle(hs, ts, 0, keepalive=keep)
    white = [[255, 255, 255]]
    # grazing: ray along +x touching the top of the sphere at (0, 1, 0)
    a, c, o, d, inside = dsc.reflection([0], [[-2.0, 1.0, 0.0]], [[1.0, 0.0, 0.0]], [[0.0, 1.0, 0.0]], white, [[0.5, 0.5, 0.5, 0.5]])
    assert a[0] == 0 and np.allclose(d[0], [1, 0, 0], atol=1e-6) and np.allclose(o[0], [0, 1, 0]) and list(c[0]) == [0, 255, 0]
    # head-on through the centre: straight through for u above R0 = 0.04
    a, c, o, d, inside = dsc.reflection([0], [[-2.0, 0.0, 0.0]], [[1.0, 0.0, 0.0]], [[-1.0, 0.0, 0.0]], white, [[0.5, 0.5, 0.5, 0.5]])
    assert a[0] == 0 and np.allclose(d[0], [1, 0, 0], atol=1e-6) and np.allclose(o[0], [-1, 0, 0]) and list(c[0]) == [0, 255, 0]
    # ... and reflects for u below it (the 4 % case that makes the reference's own test flaky)
    a, c, o, d, inside = dsc.reflection([0], [[-2.0, 0.0, 0.0]], [[1.0, 0.0, 0.0]], [[-1.0, 0.0, 0.0]], white, [[0.01, 0.5, 0.5, 0.5]])
    assert np.allclose(d[0], [-1, 0, 0], atol=1e-6)
    # dielectric with refraction probability 1: straight through
    a, c, o, d, inside = dsc.reflection([1], [[8.0, 0.0, 0.0]], [[1.0, 0.0, 0.0]], [[9.0, 0.0, 0.0]], white, [[0.5, 0.5, 0.5, 0.5]])
    assert a[0] == 0 and np.allclose(d[0], [1, 0, 0], atol=1e-6) and list(c[0]) == [0, 255, 0]


# ---- textures ---------------------------------------------------------------------------------------------------
def test_image_texture_lookup():
    spec = sample_images.earth()
    osc, dsc, cam = scene_pair(spec)
    rng = np.random.default_rng(11)
    n = 100_000
    p = f32(random_unit_vectors(rng, n))
    prim = np.zeros(n, np.int32)
    want = osc.texture(prim, p)
    got = dsc.texture(prim, p)
    same = (want == got).all(1)
    # a point within FP32 rounding of a texel edge may land in the neighbouring texel
    assert same.mean() > 0.998, same.mean()


def test_plane_map_known_answers_via_checker():
    """The twelve fixed (u, v) <-> point pairs of TestSphere.fs:196-214 pin the parameterisation; on the device it is
    observable through a checker texture whose cells are aligned with u and v."""
    from ray_tracing_fsharp_b200.domain import (Colour, Hittable, ParameterisedTexture, Pixel, Sphere, SphereStyle, marshal)
    interpret = Sphere.plane_map_inverse(1.0, (0.0, 0.0, 0.0))
    chk = ParameterisedTexture.Checkered(ParameterisedTexture.Colour(Colour.Red), ParameterisedTexture.Colour(Colour.Blue), 10.0)
    objs = [Hittable.Sphere(Sphere.make(SphereStyle.LambertReflection(1.0, ParameterisedTexture.to_texture(interpret, chk)), (0, 0, 0), 1.0))]
    hs, ts, keep = marshal(objs)
    dsc = native.SceneHandle(hs, ts, 0, keepalive=keep)
    osc = oracle.Scene(hs, ts)
    rng = np.random.default_rng(12)
    p = f32(random_unit_vectors(rng, 50_000))
    prim = np.zeros(len(p), np.int32)
    want, got = osc.texture(prim, p), dsc.texture(prim, p)
    assert ((want == got).all(1)).mean() > 0.999
    assert 0.2 < (got[:, 0] == 255).mean() < 0.8


def test_baked_closure_texture_on_the_device():
    """a ParameterisedTexture.Arbitrary closure sampled into an Image (ParameterisedTexture.bake) is
    looked up on the device exactly as the oracle looks up that Image, and agrees with the closure itself to within
    the texel grid."""
    from ray_tracing_fsharp_b200.domain import Hittable, ParameterisedTexture, Pixel, Sphere, SphereStyle, Texture, marshal
    interpret = Sphere.plane_map_inverse(1.5, (0.5, -1.0, 2.0))
    closure = ParameterisedTexture.Arbitrary(lambda u, v: Texture.Colour(Pixel(int(255 * u), int(255 * v), 99)))
    baked = ParameterisedTexture.bake(closure, interpret, 256, 128)
    objs = [Hittable.Sphere(Sphere.make(SphereStyle.LambertReflection(1.0, ParameterisedTexture.to_texture(interpret, baked)), (0.5, -1.0, 2.0), 1.5))]
    hs, ts, keep = marshal(objs)
    dsc = native.SceneHandle(hs, ts, 0, keepalive=keep)
    osc = oracle.Scene(hs, ts)
    rng = np.random.default_rng(13)
    p = f32(1.5 * random_unit_vectors(rng, 50_000) + np.array([0.5, -1.0, 2.0]))
    prim = np.zeros(len(p), np.int32)
    want, got = osc.texture(prim, p), dsc.texture(prim, p)
    assert ((want == got).all(1)).mean() > 0.995
    uv = np.array([oracle.plane_map_inverse(1.5, (0.5, -1.0, 2.0), q) for q in p[:2000]])
    direct = np.stack([np.floor(255 * uv[:, 0]), np.floor(255 * uv[:, 1]), np.full(len(uv), 99)], 1)
    err = np.abs(got[:2000].astype(int) - direct.astype(int))
    assert err[:, 0].max() <= 3 and err[:, 1].max() <= 4 and err[:, 2].max() == 0, err.max(0)


# ---- one path ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("config", ["C1", "C2", "C3", "C4", "C5"])
def test_trace_samples_match_oracle_sample_for_sample(config):
    spec = sample_images.CONFIGS[config]()
    osc, dsc, cam = scene_pair(spec)
    rng = np.random.default_rng(13)
    n = 30_000 if config != "C5" else 8_000  # the oracle walks the 100 k-sphere reference tree exhaustively
    row = rng.integers(0, spec.rows, n).astype(np.int32)
    col = rng.integers(0, spec.cols, n).astype(np.int32)
    smp = rng.integers(0, spec.spp, n).astype(np.int32)
    wc, wr = osc.trace_samples(cam, spec.max_width_coord, spec.max_height_coord, 99, row, col, smp)
    gc, gr = dsc.trace_samples(cam, spec.max_width_coord, spec.max_height_coord, 99, row, col, smp)
    same = (wc == gc).all(1)
    if config != "C5":
        assert same.mean() > 0.998, same.mean()         # FP32 flips a decision on a few paths; they then diverge
        assert (wr == gr).mean() > 0.998
    else:
        # Spheres of radius 0.05-0.3 at |x| ~ 50-90: an FP32 strike point is off by ~4e-6, i.e. 1e-4 of such a radius in the
        # normal, and every bounce off so small a sphere magnifies a direction error by ~distance / radius ~ 100.  After
        # two or three bounces the FP32 and FP64 trajectories part (measured: 9 % of paths end with another ray count),
        # by which time the integer colour is mostly black already (measured: 99.25 % of paths give identical bytes).
        # Unbiased all the same: the means agree to 0.015 of 255 here and the 4096-spp report covers the scene.
        assert same.mean() > 0.985, same.mean()
        assert (wr == gr).mean() > 0.85
    assert np.abs(gc.astype(float).mean(0) - wc.astype(float).mean(0)).max() < 0.5


def test_fp32_peak_microbenchmark_is_plausible():
    tf = native.measure_fp32_peak(0)
    assert 20.0 < tf < 120.0, tf
