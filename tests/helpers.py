"""Shared fixtures for the parity tests: seeded scenes, cameras and ray generators."""
import numpy as np

import oracle
from ray_tracing_fsharp_b200 import abi, sample_images
from ray_tracing_fsharp_b200.domain import marshal


def unit(v):
    v = np.asarray(v, float)
    return v / np.sqrt((v * v).sum(-1, keepdims=True))


def random_unit_vectors(rng, n):
    return unit(rng.normal(size=(n, 3)))


def f32(a):
    """Round to FP32-representable doubles, so that the oracle and the device see the same inputs."""
    return np.asarray(a, np.float32).astype(np.float64)


def oracle_camera(spec):
    cam = oracle.camera_make_basic(spec.spp, spec.focal_length, spec.aspect_ratio, spec.origin, spec.view_direction, spec.view_up)
    cam.bounce_depth = spec.bounce_depth
    return cam


def scene_pair(spec, device=0):
    """(oracle scene, device scene handle, camera) for a SceneSpec."""
    from ray_tracing_fsharp_b200 import native
    hs, ts, keep = marshal(spec.objects)
    osc = oracle.Scene(hs, ts)
    dsc = native.SceneHandle(hs, ts, device, keepalive=keep)
    return osc, dsc, oracle_camera(spec)


def small_random_spheres(n_side=4, seed=3):
    """A reduced RTOW-style scene (same material mix, fewer spheres) for cases the oracle must finish quickly."""
    rng = np.random.default_rng(seed)
    from ray_tracing_fsharp_b200.domain import Colour, Hittable, Pixel, Sphere, SphereStyle, Texture
    objs = []
    for a in range(-n_side, n_side):
        for b in range(-n_side, n_side):
            choice = rng.random()
            centre = (a + 0.9 * rng.random(), 0.2, b + 0.9 * rng.random())
            if choice < 0.6:
                style = SphereStyle.LambertReflection(rng.random() * rng.random(), Texture.Colour(Colour.random(rng)))
            elif choice < 0.8:
                style = SphereStyle.FuzzedReflection(rng.random() / 2 + 0.5, Texture.Colour(Colour.random(rng)), rng.random() / 2)
            elif choice < 0.9:
                style = SphereStyle.Glass(1.0, Texture.Colour(Colour.White), 1.5)
            else:
                style = SphereStyle.Dielectric(0.95, Texture.Colour(Colour.random(rng)), 1.4, 0.8)
            objs.append(Hittable.Sphere(Sphere.make(style, centre, 0.2)))
    objs.append(Hittable.Sphere(Sphere.make(SphereStyle.Glass(1.0, Texture.Colour(Colour.White), 1.5), (0.0, 1.0, 0.0), 1.0)))
    objs.append(Hittable.Sphere(Sphere.make(SphereStyle.PureReflection(1.0, Texture.Colour(Pixel(180, 150, 128))), (2.5, 1.0, 0.0), 1.0)))
    objs.append(Hittable.UnboundedSphere(Sphere.make(SphereStyle.LightSource(Texture.Colour(Pixel(200, 200, 255))), (0.0, 0.0, 0.0), 2000.0)))
    objs.append(Hittable.UnboundedSphere(Sphere.make(SphereStyle.LambertReflection(0.5, Texture.Colour(Colour.White)), (0.0, -1000.0, 0.0),
                                                     1000.0)))
    return sample_images.SceneSpec("reduced RTOW scene", objs, 32, 6.0, 1.5, (8.0, 2.0, -3.0), (0.0, 0.5, 0.0), (0.0, 1.0, 0.0), 60, 40, 50)


def camera_sample_rays(spec, cam, rng, n):
    """Primary rays of random pixels as the oracle generates them: FP32-representable origins, unit directions in
    double (the oracle's formulas assume |d| = 1 to double precision, as the reference's UnitVector guarantees; the
    device receives the same direction rounded to FP32)."""
    row = rng.integers(-spec.max_height_coord - 1, spec.max_height_coord, n).astype(np.int32)
    col = rng.integers(-spec.max_width_coord, spec.max_width_coord + 1, n).astype(np.int32)
    o, d = oracle.camera_rays(cam, spec.max_width_coord, spec.max_height_coord, row, col, rng.random(n), rng.random(n))
    return f32(o), d
