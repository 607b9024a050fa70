"""Independent checks of the oracle on the rows the reference's own tests do not pin (SURVEY.md §8c "not pinned"):
closed-form optics and arithmetic written here from first principles in numpy, not from the oracle's code.  They do not
replace running the F# reference (impossible in this image), but an oracle that broke Snell's law, the reflection law,
the cube-normalised lobes (F4), round-half-even (F1), truncating means (F2) or the tie rule (F12) would fail them."""
import numpy as np
import pytest

import oracle
from helpers import unit
from ray_tracing_fsharp_b200 import abi
from ray_tracing_fsharp_b200.domain import (Colour, Hittable, InfinitePlane, InfinitePlaneStyle, Pixel, Sphere, SphereStyle, Texture, marshal)

WHITE = (255, 255, 255)
C = np.array([0.0, 0.0, 0.0])


def _hit_point(theta, phi=0.3):
    """A point on the unit sphere and an incoming unit ray that strikes it at angle `theta` to the outward normal."""
    n = np.array([np.sin(phi), 0.0, -np.cos(phi)])
    t = unit(np.cross(n, [0.0, 1.0, 0.0]))
    d = -np.cos(theta) * n + np.sin(theta) * t        # travelling inwards
    return n, t, d, n - 3.0 * d                       # strike point = n (radius 1), origin 3 back along the ray


def _scatter(style, n, d, o, u, albedo=1.0, ior=1.5, prob=1.0, fuzz=0.0, radius=1.0, colour_in=WHITE):
    return oracle.sphere_reflection_direct(style, albedo, (255, 255, 255), ior, prob, fuzz, centre=C, radius=radius, o=o, d=d, strike=n,
                                           colour_in=colour_in, uniforms=u)


@pytest.mark.parametrize("theta", [0.05, 0.4, 0.9, 1.3, 1.5])
def test_pure_reflection_obeys_the_reflection_law(theta):
    n, t, d, o = _hit_point(theta)
    absorbed, colour, oo, dd = _scatter(abi.RT_STYLE_PURE_REFLECTION, n, d, o, [0, 0, 0, 0])
    assert not absorbed and np.allclose(oo, n, atol=1e-12)
    assert np.allclose(dd, d - 2.0 * (d @ n) * n, atol=1e-10)      # mirror image
    assert abs(np.linalg.norm(dd) - 1.0) < 1e-12


@pytest.mark.parametrize("theta", [0.05, 0.4, 0.9, 1.3, 1.5])
def test_refraction_entering_obeys_snell(theta):
    n, t, d, o = _hit_point(theta)
    ior = 1.5
    absorbed, colour, oo, dd = _scatter(abi.RT_STYLE_DIELECTRIC, n, d, o, [0.0, 0, 0, 0], ior=ior, prob=1.0)
    sin_t = np.sin(theta) / ior
    want = -np.sqrt(1 - sin_t ** 2) * n + sin_t * t                # in the plane of incidence, bent towards the normal
    assert np.allclose(dd, want, atol=1e-10)
    assert abs(dd @ np.cross(n, d)) < 1e-10                        # stays in the plane of incidence


@pytest.mark.parametrize("theta,expect_tir", [(0.2, False), (0.6, False), (0.75, True), (1.2, True)])
def test_refraction_leaving_and_total_internal_reflection(theta, expect_tir):
    # inside the unit sphere, travelling outwards: origin inside => the normal is flipped inwards (F10)
    n_out = np.array([0.0, 0.0, 1.0])
    t = np.array([1.0, 0.0, 0.0])
    d = np.cos(theta) * n_out + np.sin(theta) * t                   # travelling outwards
    o = n_out - 0.5 * d                                             # origin inside the sphere
    ior = 1.5
    absorbed, colour, oo, dd = oracle.sphere_reflection_direct(abi.RT_STYLE_DIELECTRIC, 1.0, WHITE, ior, 1.0, 0.0, centre=C, radius=1.0, o=o,
                                                               d=d, strike=n_out, colour_in=WHITE, uniforms=[0.0, 0, 0, 0])
    sin_t = np.sin(theta) * ior
    if expect_tir:
        assert sin_t > 1.0
        assert np.allclose(dd, d - 2.0 * (d @ n_out) * n_out, atol=1e-10)   # reflected back inside
    else:
        want = np.sqrt(1 - sin_t ** 2) * n_out + sin_t * t
        assert np.allclose(dd, want, atol=1e-10)


def test_glass_uses_schlick_threshold():
    theta = 1.0
    n, t, d, o = _hit_point(theta)
    r0 = ((1 - 1.5) / (1 + 1.5)) ** 2
    r = r0 + (1 - r0) * (1 - np.cos(theta)) ** 5
    below = _scatter(abi.RT_STYLE_GLASS, n, d, o, [r - 1e-6, 0, 0, 0])[3]
    above = _scatter(abi.RT_STYLE_GLASS, n, d, o, [r + 1e-6, 0, 0, 0])[3]
    assert np.allclose(below, d - 2.0 * (d @ n) * n, atol=1e-10)    # u < R reflects (Sphere.fs:288-296)
    assert above @ n < 0                                            # refracted into the sphere


def test_lambert_and_fuzz_lobes_are_cube_normalised():
    """F4: UnitVector.random normalises a point of the cube [-1,1]^3 — not a uniform direction on the sphere."""
    rng = np.random.default_rng(1)
    n, t, d, o = _hit_point(0.7)
    for _ in range(50):
        u = rng.random(3)
        offset = unit(2 * u - 1)
        dd = _scatter(abi.RT_STYLE_LAMBERT_REFLECTION, n, d, o, [*u, 0.5])[3]
        assert np.allclose(dd, unit(n + offset), atol=1e-10)        # Sphere.fs:211-220
        refl = d - 2.0 * (d @ n) * n
        dd = _scatter(abi.RT_STYLE_FUZZED_REFLECTION, n, d, o, [*u, 0.5], fuzz=0.35)[3]
        assert np.allclose(dd, unit(refl + 0.35 * offset), atol=1e-10)  # Sphere.fs:89-104: kept even when it points inwards (F11)


def test_colour_arithmetic():
    # Pixel.combine truncates, Pixel.darken rounds half to even (F1)
    assert oracle.combine([[200, 100, 7]], [[100, 255, 128]]).tolist() == [[78, 100, 3]]       # 20000/255=78.4, 896/255=3.5
    assert oracle.darken(0.5, [[1, 3, 5]]).tolist() == [[0, 2, 2]]                             # 0.5->0, 1.5->2, 2.5->2
    assert oracle.darken(0.5, [[7, 255, 0]]).tolist() == [[4, 128, 0]]                         # 3.5->4, 127.5->128
    # PixelStats.mean truncates (F2)
    assert oracle.stats_mean([10, 11, 12, 3]).tolist() == [3, 3, 4]
    # gamma: round(255 sqrt(b/255)), half to even
    assert [oracle.gamma_correct(b) for b in (0, 1, 64, 128, 255)] == [0, 16, 128, 181, 255]


def _scene(objs):
    hs, ts, _keep = marshal(objs)
    return oracle.Scene(hs, ts)


def test_hit_object_tie_rule_and_order():
    """F12: an unbounded object must be Less by more than 1e-8 in t^2 to replace the tree's candidate."""
    lam = SphereStyle.LambertReflection(1.0, Texture.Colour(Colour.White))
    sc = _scene([Hittable.UnboundedSphere(Sphere.make(lam, (0, 0, 5), 1.0)), Hittable.Sphere(Sphere.make(lam, (0, 0, 5), 1.0))])
    prim, t, strike, _ = sc.hit_object([[0, 0, 0]], [[0, 0, 1]])
    assert prim[0] == 1 and t[0] == 4.0 and np.allclose(strike[0], [0, 0, 4])                  # exact tie: the bounded one stays
    sc = _scene([Hittable.UnboundedSphere(Sphere.make(lam, (0, 0, 5), 1.001)), Hittable.Sphere(Sphere.make(lam, (0, 0, 5), 1.0))])
    prim, t, _, _ = sc.hit_object([[0, 0, 0]], [[0, 0, 1]])
    assert prim[0] == 0 and abs(t[0] - 3.999) < 1e-12                                          # clearly nearer: it wins
    # behind the ray / no object: nothing
    prim, t, _, _ = sc.hit_object([[0, 0, 0]], [[0, 0, -1]])
    assert prim[0] == -1 and np.isnan(t[0])


def test_infinite_plane_semantics():
    """InfinitePlane.fs: t = n.(p0 - o) / n.d; Lambert uses the plane's normal unflipped, so a ray that hits the back
    scatters THROUGH the plane (F11); PureReflection mirrors."""
    floor = InfinitePlane.make(InfinitePlaneStyle.LambertReflection(0.5, Pixel(200, 100, 50)), (0.0, -1.0, 0.0), (0.0, 1.0, 0.0))
    mirror = InfinitePlane.make(InfinitePlaneStyle.PureReflection(1.0, Pixel(255, 255, 255)), (0.0, 0.0, 10.0), (0.0, 0.0, -1.0))
    sc = _scene([Hittable.InfinitePlane(floor), Hittable.InfinitePlane(mirror)])
    d = unit([0.0, -1.0, 1.0])
    prim, t, strike, _ = sc.hit_object([[0, 0, 0]], [d])
    assert prim[0] == 0 and abs(t[0] - np.sqrt(2.0)) < 1e-12 and np.allclose(strike[0], [0, -1, 1])
    u = [0.9, 0.2, 0.6, 0.0]
    a, col, oo, dd, _ = sc.reflection([0], [[0, 0, 0]], [d], strike, [WHITE], [u])
    assert a[0] == 0 and col[0].tolist() == [100, 50, 25] and np.allclose(dd[0], unit(np.array([0, 1.0, 0]) + unit(2 * np.array(u[:3]) - 1)))
    # from below: same formula, the direction still leans towards +y, i.e. through the plane
    d_up = unit([0.0, 1.0, 1.0])
    a, col, oo, dd, _ = sc.reflection([0], [[0, -3, 0]], [d_up], [[0, -1, 2]], [WHITE], [u])
    assert np.allclose(dd[0], unit(np.array([0, 1.0, 0]) + unit(2 * np.array(u[:3]) - 1)))
    a, col, oo, dd, _ = sc.reflection([1], [[0, 0, 0]], [unit([1.0, 0, 1.0])], [[10, 0, 10]], [WHITE], [u])
    assert np.allclose(dd[0], unit([1.0, 0, -1.0]))


def test_camera_ray_generation_formula():
    """Scene.fs:129-144 written out: P = C + X (col + r1) VW / maxW + Y (row + r2) VH / maxH, ray = unit(P - origin)."""
    cam = oracle.camera_make_basic(4, 2.0, 1.5, (1.0, 2.0, 3.0), tuple(unit([0.0, 0.0, 1.0])), (0.0, 1.0, 0.0))
    assert np.allclose(cam.xaxis_dir, [1, 0, 0]) and np.allclose(cam.yaxis_dir, [0, 1, 0]) and np.allclose(cam.xaxis_origin, [1, 2, 5])
    assert cam.viewport_height == 2.0 and cam.viewport_width == 3.0
    o, d = oracle.camera_rays(cam, 30, 20, [7, -21], [-30, 30], [0.25, 1.0], [0.5, 0.0])
    want0 = unit(np.array([1, 2, 5]) + np.array([1, 0, 0]) * ((-30 + 0.25) * 3.0 / 30) + np.array([0, 1, 0]) * ((7 + 0.5) * 2.0 / 20) - np.array([1, 2, 3]))
    want1 = unit(np.array([1, 2, 5]) + np.array([1, 0, 0]) * ((30 + 1.0) * 3.0 / 30) + np.array([0, 1, 0]) * ((-21 + 0.0) * 2.0 / 20) - np.array([1, 2, 3]))
    assert np.allclose(o, [[1, 2, 3], [1, 2, 3]]) and np.allclose(d[0], want0, atol=1e-12) and np.allclose(d[1], want1, atol=1e-12)


def test_trace_ray_terminal_colours():
    """F3: no sky — a miss is Black; exceeding the bounce budget is HotPink; an emitter returns combine(colour, light)."""
    lam = SphereStyle.LambertReflection(1.0, Texture.Colour(Pixel(255, 255, 255)))
    light = SphereStyle.LightSource(Texture.Colour(Pixel(200, 100, 50)))
    cam = oracle.camera_make_basic(1, 1.0, 1.0, (0.0, 0.0, 0.0), (0.0, 0.0, 1.0), (0.0, 1.0, 0.0))
    cam.bounce_depth = 3
    rows = np.array([5], np.int32)
    col, rays = _scene([Hittable.UnboundedSphere(Sphere.make(light, (0, 0, 0), 50.0))]).trace_samples(cam, 5, 5, 1, rows, rows, [0])
    assert col[0].tolist() == [200, 100, 50] and rays[0] == 1
    col, rays = _scene([]).trace_samples(cam, 5, 5, 1, rows, rows, [0])
    assert col[0].tolist() == [0, 0, 0] and rays[0] == 1
    col, rays = _scene([Hittable.UnboundedSphere(Sphere.make(lam, (0, 0, 0), 50.0))]).trace_samples(cam, 5, 5, 1, rows, rows, [0])
    assert col[0].tolist() == [205, 105, 180] and rays[0] == 4     # maxCount + 1 interactions


def test_earth_map_row_flip_and_lookup_on_the_real_bitmap():
    """The `earth` path on the real decoded picture (SampleImages.fs:962-968): ParameterisedTexture.ofImage stores row
    img.Height - y - 1 of the bitmap as img.[y] (Texture.fs:30-48), and colourAt (Texture.fs:63-67) reads
    img.[int (v (H - 1))].[int ((1 - u) (W - 1))] with (u, v) = Sphere.planeMapInverse (Sphere.fs:55-61).  Written here
    from those lines in numpy, independently of the oracle, and compared with the oracle's lookup on the committed
    decode of earthmap.jpg — so that the flip has met a real bitmap: the picture is far from symmetric top to bottom."""
    import oracle
    from ray_tracing_fsharp_b200 import sample_images
    from ray_tracing_fsharp_b200.domain import ParameterisedTexture, marshal
    bitmap, source = sample_images.load_earthmap()
    assert source.startswith("earthmap.jpg"), source  # the fixture is committed; the synthetic stand-in must not be what runs
    assert bitmap.shape == (512, 1024, 3)
    assert np.abs(bitmap[:40].astype(int).mean() - bitmap[-40:].astype(int).mean()) > 1.0  # top and bottom differ
    img = ParameterisedTexture.of_image(bitmap).img
    assert np.array_equal(img[0], bitmap[511]) and np.array_equal(img[511], bitmap[0]) and np.array_equal(img[100, 7], bitmap[411, 7])
    spec = sample_images.earth()
    hs, ts, _keep = marshal(spec.objects)
    osc = oracle.Scene(hs, ts)
    rng = np.random.default_rng(21)
    p = rng.normal(size=(20000, 3))
    p /= np.sqrt((p * p).sum(1, keepdims=True))
    got = osc.texture(np.zeros(len(p), np.int32), p)
    theta = np.arccos(np.clip(-p[:, 1], -1.0, 1.0))
    phi = np.arctan2(-p[:, 2], p[:, 0]) + np.pi
    u, v = phi / (2 * np.pi), theta / np.pi
    x = ((1.0 - u) * 1023).astype(int)
    y = (v * 511).astype(int)
    want = bitmap[511 - y, x]  # img.[y] is bitmap row H - 1 - y
    assert (got == want).all(1).mean() > 0.9999
