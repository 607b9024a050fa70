"""Scene specifications for the five BASELINE.json configs and for the reference's own ray-traced
sample scenes (REFERENCE_SAMPLES; `random-spheres` and `earth` are C2 and C3), built with the mirrored
F# surface in domain.py.  They follow RayTracing.App/SampleImages.fs where a sample exists
(randomSpheres :812-960, earth :962-1010, glassSphere :506-597, movedCamera :710-810) but draw their
random parameters from a fixed-seed numpy generator instead of `System.Random ()`, so the oracle
and the GPU are fed identical scenes.  No file under /root/reference is read (the earth map comes from a committed
fixture, see load_earthmap).

Each function returns a `SceneSpec`: objects, camera arguments (for Camera.makeBasic), the two
half-extents handed to Scene.render (F6: image is (2*max_w+1) x (2*max_h+1)), spp and depth.
"""
import functools
import os
from dataclasses import dataclass
from typing import Any, List, Tuple

import numpy as np

from .domain import (Colour, Hittable, InfinitePlane, InfinitePlaneStyle, ParameterisedTexture, Pixel, Sphere,
                     SphereStyle, Texture)


@dataclass
class SceneSpec:
    name: str
    objects: List[Any]
    spp: int
    focal_length: float
    aspect_ratio: float
    origin: Tuple[float, float, float]
    look_at: Tuple[float, float, float]
    view_up: Tuple[float, float, float]
    max_width_coord: int
    max_height_coord: int
    bounce_depth: int

    @property
    def view_direction(self):
        v = np.asarray(self.look_at, float) - np.asarray(self.origin, float)
        return tuple(v / np.sqrt(v @ v))

    @property
    def rows(self):
        return 2 * self.max_height_coord + 1

    @property
    def cols(self):
        return 2 * self.max_width_coord + 1


def few_spheres(max_w=200, max_h=112, spp=16, depth=50) -> SceneSpec:
    """C1: few-sphere Lambertian scene on a ground plane, 401x225, 16 spp, depth 50."""
    objs = [
        Hittable.InfinitePlane(InfinitePlane.make(
            InfinitePlaneStyle.LambertReflection(0.5, Pixel(204, 204, 0)), (0.0, -0.5, 0.0), (0.0, 1.0, 0.0))),
        Hittable.Sphere(Sphere.make(SphereStyle.LambertReflection(1.0, Texture.Colour(Pixel(100, 150, 200))),
                                    (1.0, 0.0, 1.0), 0.5)),
        Hittable.Sphere(Sphere.make(SphereStyle.LambertReflection(1.0, Texture.Colour(Pixel(25, 50, 120))),
                                    (0.0, 0.0, 1.0), 0.5)),
        Hittable.Sphere(Sphere.make(SphereStyle.LambertReflection(0.9, Texture.Colour(Colour.White)),
                                    (-1.0, 0.0, 1.0), 0.5)),
        Hittable.UnboundedSphere(Sphere.make(SphereStyle.LightSource(Texture.Colour(Pixel(200, 200, 200))),
                                             (0.0, 0.0, 0.0), 200.0)),
    ]
    return SceneSpec("C1 few-sphere Lambertian on ground plane", objs, spp, 1.0, 16.0 / 9.0, (0.0, 0.0, 0.0),
                     (0.0, 0.0, 1.0), (0.0, 1.0, 0.0), max_w, max_h, depth)


def random_spheres(max_w=600, max_h=400, spp=500, depth=50, seed=1) -> SceneSpec:
    """C2: the RTOW final scene exactly as randomSpheres (SampleImages.fs:812-960)."""
    rng = np.random.default_rng(seed)
    objs = []
    for a in range(-11, 11):
        for b in range(-11, 11):
            material_choice = rng.random()
            centre = (a + 0.9 * rng.random(), 0.2, b + 0.9 * rng.random())
            d = np.asarray(centre) - np.asarray((4.0, 0.2, 0.0))
            if d @ d > 0.9 * 0.9:
                if material_choice < 0.8 - 1e-8:  # Float.compare materialChoice 0.8 = Less
                    albedo = rng.random() * rng.random()
                    style = SphereStyle.LambertReflection(albedo, Texture.Colour(Colour.random(rng)))
                elif material_choice < 0.95 - 1e-8:
                    albedo = rng.random() / 2.0 + 0.5
                    fuzz = rng.random() / 2.0
                    style = SphereStyle.FuzzedReflection(albedo, Texture.Colour(Colour.random(rng)), fuzz)
                else:
                    style = SphereStyle.Glass(1.0, Texture.Colour(Colour.White), 1.5)
                objs.append(Hittable.Sphere(Sphere.make(style, centre, 0.2)))
    objs.append(Hittable.Sphere(Sphere.make(SphereStyle.Glass(1.0, Texture.Colour(Colour.White), 1.5),
                                            (0.0, 1.0, 0.0), 1.0)))
    objs.append(Hittable.Sphere(Sphere.make(SphereStyle.LambertReflection(1.0, Texture.Colour(Pixel(80, 40, 20))),
                                            (-4.0, 1.0, 0.0), 1.0)))
    objs.append(Hittable.Sphere(Sphere.make(SphereStyle.PureReflection(1.0, Texture.Colour(Pixel(180, 150, 128))),
                                            (4.0, 1.0, 0.0), 1.0)))
    # Ceiling
    objs.append(Hittable.UnboundedSphere(Sphere.make(SphereStyle.LightSource(Texture.Colour(Pixel(200, 200, 255))),
                                                     (0.0, 0.0, 0.0), 2000.0)))
    # Floor
    objs.append(Hittable.UnboundedSphere(Sphere.make(SphereStyle.LambertReflection(0.5, Texture.Colour(Colour.White)),
                                                     (0.0, -1000.0, 0.0), 1000.0)))
    return SceneSpec("C2 RTOW final scene (random-spheres)", objs, spp, 10.0, 3.0 / 2.0, (13.0, 2.0, -3.0),
                     (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), max_w, max_h, depth)


def synthetic_earthmap(width=1024, height=512, seed=7) -> np.ndarray:
    """A 1024x512 RGB8 lat-long picture standing in for earthmap.jpg (same size and kind of content:
    blue oceans, green/brown land, white caps).  The real JPEG is a reference asset and needs Skia's
    decoder (parity unpinned at that boundary), so benchmarks use this deterministic
    stand-in; row 0 is the top of the picture, as SKBitmap presents it."""
    rng = np.random.default_rng(seed)
    lat = np.linspace(np.pi / 2, -np.pi / 2, height)[:, None]
    lon = np.linspace(-np.pi, np.pi, width, endpoint=False)[None, :]
    x, y, z = np.cos(lat) * np.cos(lon), np.sin(lat) + 0 * lon, np.cos(lat) * np.sin(lon)
    h = np.zeros((height, width))
    for k in range(1, 7):
        for _ in range(4):
            w = rng.normal(size=3) * (2.0 ** k) * 0.6
            ph = rng.uniform(0, 2 * np.pi)
            h += np.sin(w[0] * x + w[1] * y + w[2] * z + ph) / (1.7 ** k)
    h = (h - h.min()) / (h.max() - h.min())
    land = h > 0.55
    img = np.zeros((height, width, 3))
    img[..., 0] = np.where(land, 60 + 150 * (h - 0.55) / 0.45, 10 + 30 * h)
    img[..., 1] = np.where(land, 120 + 60 * (1 - h), 40 + 80 * h)
    img[..., 2] = np.where(land, 40 + 40 * h, 120 + 120 * h)
    cap = np.abs(lat) > 1.25
    img[np.broadcast_to(cap, h.shape)] = 235
    return np.clip(img, 0, 255).astype(np.uint8)


# earthmap.jpg (the embedded resource LoadImage.fromResource decodes, RayTracing.App/LoadImage.fs:9-18) decoded once with
# PIL / libjpeg-turbo by tests/golden/make_golden.py and committed losslessly: nothing here reads /root/reference
_EARTHMAP_FIXTURE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "earthmap_rgb8.png")


@functools.lru_cache(maxsize=1)
def load_earthmap():
    """The bitmap `earth` textures its sphere with (SampleImages.fs:962-968), as [H, W, 3] uint8 with row 0 the top row of
    the picture — what SKBitmap.GetPixel(x, y) indexes — plus a word saying what it is:
      "earthmap.jpg decoded with PIL (committed fixture)"   the reference's own JPEG, decoded in the build container and
                                           committed as a PNG, so that this container and the GPU box render the same texels;
      "synthetic 1024x512 lat-long map"    the deterministic stand-in, when the fixture (or PIL) is missing.
    Decoded texels are INPUT to the library (the decoder stays on the host side of the ABI): Skia's decoder
    may differ from libjpeg-turbo's by a level or two per texel, which is outside the contract."""
    try:
        from PIL import Image as PilImage
        if os.path.exists(_EARTHMAP_FIXTURE):
            with PilImage.open(_EARTHMAP_FIXTURE) as im:
                return np.ascontiguousarray(np.asarray(im.convert("RGB"), dtype=np.uint8)), "earthmap.jpg decoded with PIL (committed fixture)"
    except Exception:  # PIL missing or the file unreadable: fall through to the stand-in
        pass
    return synthetic_earthmap(), "synthetic 1024x512 lat-long map"


def earth(max_w=960, max_h=540, spp=256, depth=50, bitmap=None) -> SceneSpec:
    """C3: `earth` (SampleImages.fs:962-1010) plus the two InfinitePlanes the config asks for.  `bitmap`: the decoded
    picture, row 0 on top (default: load_earthmap(), i.e. the reference's earthmap.jpg wherever it can be read)."""
    source = "caller's bitmap"
    if bitmap is None:
        bitmap, source = load_earthmap()
    texture = ParameterisedTexture.of_image(bitmap)
    interpret = Sphere.plane_map_inverse(1.0, (0.0, 0.0, 0.0))
    inv_sqrt2 = 1.0 / np.sqrt(2.0)
    objs = [
        Hittable.Sphere(Sphere.make(
            SphereStyle.LambertReflection(1.0, ParameterisedTexture.to_texture(interpret, texture)),
            (0.0, 0.0, 0.0), 1.0)),
        Hittable.UnboundedSphere(Sphere.make(SphereStyle.LightSource(Texture.Colour(Pixel(130, 130, 200))),
                                             (0.0, 0.0, 0.0), 200.0)),
        Hittable.InfinitePlane(InfinitePlane.make(InfinitePlaneStyle.LambertReflection(0.6, Pixel(200, 200, 180)),
                                                  (0.0, -1.0, 0.0), (0.0, 1.0, 0.0))),
        Hittable.InfinitePlane(InfinitePlane.make(InfinitePlaneStyle.PureReflection(0.9, Pixel(230, 230, 240)),
                                                  (-3.0, 0.0, 3.0), (inv_sqrt2, 0.0, -inv_sqrt2))),
    ]
    return SceneSpec(f"C3 earthmap image-textured sphere + planes [{source}]", objs, spp, 12.0, 16.0 / 9.0, (13.0, 2.0, -3.0),
                     (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), max_w, max_h, depth)


def mixed_planes(max_w=960, max_h=540, spp=1024, depth=100) -> SceneSpec:
    """C4: all four InfinitePlane styles + Glass / Dielectric / hollow-shell spheres (F13: the
    reference has no finite Plane primitive and no dielectric plane; refraction comes from spheres)."""
    s3 = 1.0 / np.sqrt(3.0)
    objs = [
        # Lambert floor
        Hittable.InfinitePlane(InfinitePlane.make(InfinitePlaneStyle.LambertReflection(0.5, Pixel(204, 204, 0)),
                                                  (0.0, -0.5, 0.0), (0.0, 1.0, 0.0))),
        # mirror wall on the right
        Hittable.InfinitePlane(InfinitePlane.make(InfinitePlaneStyle.PureReflection(0.9, Pixel(220, 220, 255)),
                                                  (3.0, 0.0, 0.0), (-1.0, 0.0, 0.0))),
        # fuzzed wall at the back
        Hittable.InfinitePlane(InfinitePlane.make(InfinitePlaneStyle.FuzzedReflection(0.8, Pixel(255, 200, 200), 0.3),
                                                  (0.0, 0.0, 6.0), (0.0, 0.0, -1.0))),
        # emitting ceiling
        Hittable.InfinitePlane(InfinitePlane.make(InfinitePlaneStyle.LightSource(Texture.Colour(Pixel(255, 255, 255))),
                                                  (0.0, 4.0, 0.0), (0.0, -1.0, 0.0))),
        Hittable.Sphere(Sphere.make(SphereStyle.Glass(0.9, Texture.Colour(Colour.White), 1.5), (-1.1, 0.0, 2.0), 0.5)),
        Hittable.Sphere(Sphere.make(SphereStyle.Dielectric(0.95, Texture.Colour(Pixel(200, 255, 200)), 1.5, 0.9),
                                    (0.0, 0.0, 2.5), 0.5)),
        # hollow shell: outer bounded, inner UnboundedSphere with negative radius so that it IS intersected (F16)
        Hittable.Sphere(Sphere.make(SphereStyle.Glass(1.0, Texture.Colour(Colour.White), 1.5), (1.1, 0.0, 2.0), 0.5)),
        Hittable.UnboundedSphere(Sphere.make(SphereStyle.Glass(1.0, Texture.Colour(Colour.White), 1.0 / 1.5),
                                             (1.1, 0.0, 2.0), -0.45)),
        # a bounded negative-radius sphere: never hit in the reference (F16) — pins that quirk
        Hittable.Sphere(Sphere.make(SphereStyle.LambertReflection(1.0, Texture.Colour(Colour.Red)), (0.0, 1.2, 3.0), -0.4)),
        Hittable.Sphere(Sphere.make(SphereStyle.LambertReflection(0.8, Texture.Colour(Pixel(25, 50, 120))),
                                    (-2.0, 0.1, 3.5), 0.6)),
        Hittable.Sphere(Sphere.make(SphereStyle.FuzzedReflection(0.9, Texture.Colour(Pixel(255, 215, 0)), 0.1),
                                    (2.0, 0.1, 4.0), 0.6)),
        # light leaking in from behind the camera as well, so the scene is enclosed
        Hittable.UnboundedSphere(Sphere.make(SphereStyle.LightSource(Texture.Colour(Pixel(130, 130, 200))),
                                             (0.0, 0.0, 0.0), 200.0)),
    ]
    return SceneSpec("C4 InfinitePlane mixed-primitive scene with dielectric refraction", objs, spp, 1.0, 16.0 / 9.0,
                     (0.0, 0.6, -1.5), (0.0, 0.2, 2.0), (0.0, 1.0, 0.0), max_w, max_h, depth)


def many_spheres(n=100_000, max_w=1920, max_h=1080, spp=4096, depth=50, seed=5) -> SceneSpec:
    """C5: synthetic 100k random spheres, materials 80/15/5 % Lambert/Fuzzed/Glass as in C2."""
    rng = np.random.default_rng(seed)
    cx = rng.uniform(-50, 50, n)
    cy = rng.uniform(0.2, 10, n)
    cz = rng.uniform(-50, 50, n)
    rad = rng.uniform(0.05, 0.3, n)
    choice = rng.random(n)
    u1, u2 = rng.random(n), rng.random(n)
    cols = rng.integers(0, 256, size=(n, 3))
    objs = []
    for i in range(n):
        col = Texture.Colour(Pixel(int(cols[i, 0]), int(cols[i, 1]), int(cols[i, 2])))
        if choice[i] < 0.8:
            style = SphereStyle.LambertReflection(float(u1[i] * u2[i]), col)
        elif choice[i] < 0.95:
            style = SphereStyle.FuzzedReflection(float(u1[i] / 2 + 0.5), col, float(u2[i] / 2))
        else:
            style = SphereStyle.Glass(1.0, Texture.Colour(Colour.White), 1.5)
        objs.append(Hittable.Sphere(Sphere.make(style, (float(cx[i]), float(cy[i]), float(cz[i])), float(rad[i]))))
    objs.append(Hittable.UnboundedSphere(Sphere.make(SphereStyle.LightSource(Texture.Colour(Pixel(200, 200, 255))),
                                                     (0.0, 0.0, 0.0), 2000.0)))
    objs.append(Hittable.UnboundedSphere(Sphere.make(SphereStyle.LambertReflection(0.5, Texture.Colour(Colour.White)),
                                                     (0.0, -1000.0, 0.0), 1000.0)))
    return SceneSpec(f"C5 synthetic {n} random spheres", objs, spp, 4.0, 16.0 / 9.0, (60.0, 20.0, -60.0),
                     (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), max_w, max_h, depth)


# ---- the reference's own sample scenes (RayTracing.App/SampleImages.fs), as parity cases ---------------------
# Object lists, cameras and half-extents as the sample functions give them; bounce depth 150 (Camera.makeBasic,
# Camera.fs:43).  `scale` shrinks the half-extents (the samples render up to 4267x2401).  The materials' FloatProducer
# arguments have no counterpart (the device RNG is keyed per pixel).
def _sample(name, objs, spp, focal, origin, look_at, pixels, scale):
    aspect = 16.0 / 9.0
    max_w, max_h = int(aspect * float(pixels)), pixels  # `aspectRatio * (float pixels) |> int`, e.g. SampleImages.fs:96
    return SceneSpec(name, objs, spp, focal, aspect, origin, look_at, (0.0, 1.0, 0.0), max(1, int(max_w * scale)),
                     max(1, int(max_h * scale)), 150)


def _sph(style, centre, radius, unbounded=False):
    s = Sphere.make(style, centre, radius)
    return Hittable.UnboundedSphere(s) if unbounded else Hittable.Sphere(s)


def _col(r, g, b):
    return Texture.Colour(Pixel(r, g, b))


def shiny_floor(scale=1.0) -> SceneSpec:
    """`shiny-floor`, shinyPlane (SampleImages.fs:57-96): an emitting sphere over a mirror InfinitePlane."""
    objs = [_sph(SphereStyle.LightSource(_col(0, 255, 255)), (1.5, 0.5, 8.0), 0.5),
            Hittable.InfinitePlane(InfinitePlane.make(InfinitePlaneStyle.PureReflection(0.5, Colour.White), (0.0, -1.0, 0.0), (0.0, 1.0, 0.0)))]
    return _sample("shiny-floor", objs, 50, 2.0, (0.0, 0.0, 0.0), (0.0, 0.0, 1.0), 400, scale)


def fuzzy_floor(scale=1.0) -> SceneSpec:
    """`fuzzy-floor`, fuzzyPlane (SampleImages.fs:98-136): the same over a fuzzed InfinitePlane (fuzz 0.75)."""
    objs = [_sph(SphereStyle.LightSource(_col(0, 255, 255)), (1.5, 0.5, 8.0), 0.5),
            Hittable.InfinitePlane(InfinitePlane.make(InfinitePlaneStyle.FuzzedReflection(1.0, Colour.White, 0.75), (0.0, -1.0, 0.0), (0.0, 1.0, 0.0)))]
    return _sample("fuzzy-floor", objs, 50, 2.0, (0.0, 0.0, 0.0), (0.0, 0.0, 1.0), 400, scale)


def spheres(scale=1.0) -> SceneSpec:
    """`spheres` (SampleImages.fs:138-264): four spheres between two mirror planes, a fuzzed floor and a dim emitting
    plane behind the camera."""
    r2 = 1.0 / np.sqrt(2.0)
    objs = [
        _sph(SphereStyle.LambertReflection(0.95, _col(255, 255, 0)), (0.0, 0.0, 9.0), 1.0),
        _sph(SphereStyle.PureReflection(1.0, _col(0, 255, 255)), (1.5, 0.5, 8.0), 0.5),
        _sph(SphereStyle.LightSource(_col(200, 220, 255)), (-1.5, 1.0, 8.0), 0.5),
        _sph(SphereStyle.FuzzedReflection(1.0, _col(255, 100, 0), 0.2), (-0.4, 1.5, 10.0), 0.25),
        Hittable.InfinitePlane(InfinitePlane.make(InfinitePlaneStyle.PureReflection(0.8, Colour.White), (0.0, 0.0, 12.0), (r2, 0.0, -r2))),
        Hittable.InfinitePlane(InfinitePlane.make(InfinitePlaneStyle.FuzzedReflection(0.85, Pixel(255, 100, 100), 0.8), (0.0, -1.0, 0.0), (0.0, 1.0, 0.0))),
        Hittable.InfinitePlane(InfinitePlane.make(InfinitePlaneStyle.PureReflection(0.95, Colour.White), (0.0, 0.0, 12.0), (-r2, 0.0, -r2))),
        Hittable.InfinitePlane(InfinitePlane.make(InfinitePlaneStyle.LightSource(_col(15, 15, 15)), (0.0, 1.0, -1.0), (0.0, 0.0, 1.0))),
    ]
    return _sample("spheres", objs, 50, 7.0, (0.0, 0.0, 0.0), (0.0, 0.0, 1.0), 200, scale)


def inside_sphere(scale=1.0) -> SceneSpec:
    """`inside-sphere` (SampleImages.fs:266-411): the camera sits inside a bounded r = 100 fuzzed sphere, on top of a
    bounded r = 75 one; light comes from a bounded r = 9 sphere and a plane behind the camera."""
    objs = [
        _sph(SphereStyle.LambertReflection(0.95, _col(255, 255, 0)), (0.0, 0.0, 9.0), 1.0),
        _sph(SphereStyle.PureReflection(1.0, _col(0, 255, 255)), (1.5, 0.5, 8.0), 0.5),
        _sph(SphereStyle.PureReflection(1.0, _col(255, 20, 20)), (-1.8, 0.8, 8.0), 0.5),
        _sph(SphereStyle.LightSource(Texture.Colour(Colour.White)), (-10.0, 8.0, 0.0), 9.0),
        _sph(SphereStyle.FuzzedReflection(1.0, _col(255, 100, 0), 0.2), (1.4, 1.5, 10.0), 0.25),
        _sph(SphereStyle.PureReflection(0.9, _col(255, 255, 255)), (0.0, 10.0, 20.0), 8.0),
        _sph(SphereStyle.FuzzedReflection(0.6, _col(200, 50, 255), 0.4), (0.0, -76.0, 9.0), 75.0),
        _sph(SphereStyle.FuzzedReflection(0.4, _col(200, 200, 200), 0.0), (0.0, 0.0, 20.0), 100.0),
        Hittable.InfinitePlane(InfinitePlane.make(InfinitePlaneStyle.LightSource(_col(80, 80, 150)), (0.0, 0.0, -5.0), (0.0, 0.0, 1.0))),
    ]
    return _sample("inside-sphere", objs, 50, 7.0, (0.0, 0.0, 0.0), (0.0, 0.0, 1.0), 1200, scale)


def _three_on_a_floor(left, right_colour, light, floor_unbounded, light_unbounded):
    return [
        _sph(SphereStyle.LambertReflection(0.5, _col(204, 204, 0)), (0.0, -100.5, 1.0), 100.0, floor_unbounded),
        _sph(SphereStyle.PureReflection(1.0, right_colour), (1.0, 0.0, 1.0), 0.5),
        _sph(SphereStyle.LambertReflection(1.0, _col(25, 50, 120)), (0.0, 0.0, 1.0), 0.5),
        _sph(left, (-1.0, 0.0, 1.0), 0.5),
        _sph(SphereStyle.LightSource(_col(*light)), (0.0, 0.0, 0.0), 200.0, light_unbounded),
    ]


def total_refraction(scale=1.0) -> SceneSpec:
    """`total-refraction` (SampleImages.fs:413-504): Dielectric 1.5 with refraction probability 1 on a BOUNDED
    r = 100 floor inside a BOUNDED r = 200 light."""
    objs = _three_on_a_floor(SphereStyle.Dielectric(1.0, Texture.Colour(Colour.White), 1.5, 1.0), _col(204, 153, 51), (80, 80, 150), False, False)
    return _sample("total-refraction", objs, 50, 1.0, (0.0, 0.0, 0.0), (0.0, 0.0, 1.0), 300, scale)


def glass_sphere(scale=1.0) -> SceneSpec:
    """`glass` (SampleImages.fs:506-597): Glass 1.5 (albedo 0.9), unbounded floor and light."""
    objs = _three_on_a_floor(SphereStyle.Glass(0.9, Texture.Colour(Colour.White), 1.5), _col(100, 150, 200), (200, 200, 200), True, True)
    return _sample("glass", objs, 50, 1.0, (0.0, 0.0, 0.0), (0.0, 0.0, 1.0), 200, scale)


def textured_sphere(scale=1.0, bake=(512, 256)) -> SceneSpec:
    """`textured-sphere` (SampleImages.fs:599-708): as `glass`, the mirror sphere carrying a checker (grid 50) of two
    closures of (u, v).  The closures are sampled into Images (ParameterisedTexture.bake); the checker stays a checker."""
    interpret = Sphere.plane_map_inverse(0.5, (1.0, 0.0, 1.0))
    even = ParameterisedTexture.Arbitrary(lambda x, y: Texture.Colour(Pixel(int(x * 255.0) & 255, 0, int(y * 255.0) & 255)))
    odd = ParameterisedTexture.Arbitrary(lambda x, y: Texture.Colour(Pixel(100, int(x * 255.0) & 255, int(y * 255.0) & 255)))
    texture = ParameterisedTexture.Checkered(ParameterisedTexture.bake(even, interpret, *bake), ParameterisedTexture.bake(odd, interpret, *bake), 50.0)
    objs = _three_on_a_floor(SphereStyle.Glass(0.9, Texture.Colour(Colour.White), 1.5), ParameterisedTexture.to_texture(interpret, texture),
                             (200, 200, 200), True, True)
    return _sample("textured-sphere", objs, 50, 1.0, (0.0, 0.0, 0.0), (0.0, 0.0, 1.0), 200, scale)


def moved_camera(scale=1.0) -> SceneSpec:
    """`moved-camera` (SampleImages.fs:710-810): camera at (-2, 2, -1) looking at the glass sphere; the inner shell is a
    BOUNDED sphere of radius -0.45, which the reference can never hit (F16)."""
    objs = _three_on_a_floor(SphereStyle.Glass(1.0, Texture.Colour(Colour.White), 1.5), _col(204, 153, 51), (130, 130, 200), False, False)
    objs.insert(4, _sph(SphereStyle.Glass(1.0, Texture.Colour(Colour.White), 1.0 / 1.5), (-1.0, 0.0, 1.0), -0.45))
    return _sample("moved-camera", objs, 50, 10.0, (-2.0, 2.0, -1.0), (-1.0, 0.0, 1.0), 300, scale)


REFERENCE_SAMPLES = {
    "shiny-floor": shiny_floor,
    "fuzzy-floor": fuzzy_floor,
    "spheres": spheres,
    "inside-sphere": inside_sphere,
    "total-refraction": total_refraction,
    "glass": glass_sphere,
    "textured-sphere": textured_sphere,
    "moved-camera": moved_camera,
}

CONFIGS = {
    "C1": few_spheres,
    "C2": random_spheres,
    "C3": earth,
    "C4": mixed_planes,
    "C5": many_spheres,
}
