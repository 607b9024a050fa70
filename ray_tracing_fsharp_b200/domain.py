"""Host-side mirror of the reference's scene-description API (names follow the F# modules).

The reference's toolchain (.NET / F#) is absent from this image, so the host surface that would
stay in F# (`Scene/Hittable/Camera/Texture`) is mirrored here in Python, one class
per F# type, with the same names and argument meaning, so that tests read like the reference's.
Everything here is plain data; `marshal()` flattens a Hittable list into the C-ABI arrays
(include/rtfs_b200.h) exactly as the F# shim (shim/RayTracing.Gpu.fs) would.

Reference types mirrored:
  Pixel, Colour                     RayTracing/Pixel.fs:9-76
  Texture, ParameterisedTexture     RayTracing/Texture.fs:5-72
  SphereStyle, Sphere               RayTracing/Sphere.fs:10-37, :302-337
  InfinitePlaneStyle, InfinitePlane RayTracing/InfinitePlane.fs:3-13, :101-119
  Hittable                          RayTracing/Hittable.fs:3-6
"""
import ctypes
import math
import struct
from dataclasses import dataclass, field
from typing import Any, List, Optional, Sequence, Tuple

import numpy as np

from . import abi


# ---- Pixel.fs -------------------------------------------------------------------------------
@dataclass(frozen=True)
class Pixel:
    Red: int
    Green: int
    Blue: int

    def __post_init__(self):
        for c in (self.Red, self.Green, self.Blue):
            if not (0 <= int(c) <= 255):
                raise ValueError("Pixel channels are bytes")

    def as_tuple(self):
        return (int(self.Red), int(self.Green), int(self.Blue))


class Colour:
    Black = Pixel(0, 0, 0)
    White = Pixel(255, 255, 255)
    Red = Pixel(255, 0, 0)
    Green = Pixel(0, 255, 0)
    Blue = Pixel(0, 0, 255)
    Yellow = Pixel(255, 255, 0)
    HotPink = Pixel(205, 105, 180)

    @staticmethod
    def random(rand: np.random.Generator) -> Pixel:
        b = rand.integers(0, 256, size=3)
        return Pixel(int(b[0]), int(b[1]), int(b[2]))


# ---- Texture.fs -----------------------------------------------------------------------------
@dataclass(frozen=True)
class PlaneMapInverse:
    """The `interpret` closure `Sphere.planeMapInverse radius centre` (Sphere.fs:55-61) as data."""
    radius: float
    centre: Tuple[float, float, float]


class ParameterisedTexture:
    @dataclass(frozen=True)
    class Colour:
        pixel: Pixel

    @dataclass(frozen=True)
    class Checkered:
        even: Any
        odd: Any
        grid_size: float

    @dataclass(frozen=True, eq=False)
    class Image:
        """`Pixel[][]` as ParameterisedTexture.Image holds it: img[y][x], uint8 array [H, W, 3]."""
        img: np.ndarray

    @dataclass(frozen=True, eq=False)
    class Arbitrary:
        """ParameterisedTexture.Arbitrary of (float -> float -> Texture) (Texture.fs:24): a host closure of the
        surface coordinates (u, v) in [0, 1]^2.  It cannot cross the ABI; `bake` samples it into an Image."""
        f: Any

    @staticmethod
    def bake(texture, interpret: PlaneMapInverse, width: int = 1024, height: int = 512) -> "ParameterisedTexture.Image":
        """Samples a ParameterisedTexture (typically an Arbitrary closure) into a `width` x `height` Image that
        the device can look up.  Texel (x, y) of an Image answers the lookups with
        int((1 - u)(W - 1)) = x and int(v (H - 1)) = y (Texture.fs:63-67); it is given the closure's value at
        the centre of that cell.  An approximation by construction (a closure with detail finer than a texel
        is smoothed to the texel grid), hence explicit: Scene.make never bakes on its own."""
        if width < 2 or height < 2:
            raise ValueError("bake: width and height must be at least 2")
        img = np.empty((height, width, 3), np.uint8)
        for y in range(height):
            v = min(1.0, (y + 0.5) / (height - 1))
            for x in range(width):
                u = max(0.0, 1.0 - (x + 0.5) / (width - 1))
                img[y, x] = ParameterisedTexture.colour_at(interpret, texture, u, v).as_tuple()
        return ParameterisedTexture.Image(img)

    @staticmethod
    def colour_at(interpret: PlaneMapInverse, t, u: float, v: float) -> Pixel:
        """ParameterisedTexture.colourAt (Texture.fs:50-67) at surface coordinates (u, v), on the host."""
        if isinstance(t, ParameterisedTexture.Colour):
            return t.pixel
        if isinstance(t, ParameterisedTexture.Arbitrary):
            r = t.f(u, v)
            if isinstance(r, Pixel):
                return r
            if isinstance(r, Texture.Colour):
                return r.pixel
            if isinstance(r, Texture.Arbitrary):  # Texture.colourAt p (f x y): the closure wants the point itself
                return r.f(Sphere.plane_map(interpret.radius, interpret.centre, u, v))
            raise TypeError("ParameterisedTexture.Arbitrary must return a Texture")
        if isinstance(t, ParameterisedTexture.Checkered):
            sine = math.sin(t.grid_size * u) * math.sin(t.grid_size * v)
            less = abs(sine) >= 1e-8 and sine < 0.0  # Float.compare sine 0.0 = Less (Float.fs:88-96)
            return ParameterisedTexture.colour_at(interpret, t.even if less else t.odd, u, v)
        if isinstance(t, ParameterisedTexture.Image):
            h, w = t.img.shape[0], t.img.shape[1]
            x = int((1.0 - u) * float(w - 1))
            y = int(v * float(h - 1))
            return Pixel(*[int(c) for c in t.img[y, x]])
        raise TypeError(f"not a ParameterisedTexture: {t!r}")

    @staticmethod
    def of_image(bitmap: np.ndarray) -> "ParameterisedTexture.Image":
        """ParameterisedTexture.ofImage (Texture.fs:30-48): `bitmap` is [H, W, 3] with row 0 the
        top row of the picture (what SKBitmap.GetPixel(x, y) indexes); rows are flipped."""
        b = np.ascontiguousarray(bitmap, dtype=np.uint8)
        if b.ndim != 3 or b.shape[2] != 3:
            raise ValueError("bitmap must be [H, W, 3] uint8")
        return ParameterisedTexture.Image(np.ascontiguousarray(b[::-1]))

    @staticmethod
    def to_texture(interpret: PlaneMapInverse, texture) -> "Texture":
        """ParameterisedTexture.toTexture (Texture.fs:69-72) — keeps the structure instead of
        erasing it to a closure (F14), because a closure cannot cross the ABI."""
        if isinstance(texture, ParameterisedTexture.Colour):
            return Texture.Colour(texture.pixel)
        if not isinstance(interpret, PlaneMapInverse):
            raise TypeError("interpret must be Sphere.plane_map_inverse(radius, centre)")
        return Texture.Parameterised(interpret, texture)


class Texture:
    @dataclass(frozen=True)
    class Colour:
        pixel: Pixel

    @dataclass(frozen=True, eq=False)
    class Parameterised:
        interpret: PlaneMapInverse
        texture: Any

    @dataclass(frozen=True, eq=False)
    class Arbitrary:
        """Texture.Arbitrary (Point -> Pixel): a host closure.  Not executable on the device;
        Scene.make raises (RT_ERR_UNSUPPORTED semantics) rather than mis-rendering."""
        f: Any


# ---- Sphere.fs ------------------------------------------------------------------------------
class SphereStyle:
    @dataclass(frozen=True)
    class LightSource:
        texture: Any

    @dataclass(frozen=True)
    class LightSourceCap:
        colour: Pixel

    @dataclass(frozen=True)
    class PureReflection:
        albedo: float
        texture: Any

    @dataclass(frozen=True)
    class FuzzedReflection:
        albedo: float
        texture: Any
        fuzz: float
        rand: Any = None  # FloatProducer in the reference; the device RNG is keyed per pixel instead

    @dataclass(frozen=True)
    class LambertReflection:
        albedo: float
        texture: Any
        rand: Any = None

    @dataclass(frozen=True)
    class Dielectric:
        albedo: float
        texture: Any
        boundaryRefractance: float
        refraction: float
        rand: Any = None

    @dataclass(frozen=True)
    class Glass:
        albedo: float
        texture: Any
        ior: float
        rand: Any = None


@dataclass(frozen=True)
class Sphere:
    Style: Any
    Centre: Tuple[float, float, float]
    Radius: float

    @staticmethod
    def make(style, centre, radius) -> "Sphere":
        c = tuple(float(x) for x in centre)
        if len(c) != 3:
            raise ValueError("centre must have 3 coordinates")
        return Sphere(style, c, float(radius))

    @staticmethod
    def plane_map(radius, centre, phi: float, theta: float) -> Tuple[float, float, float]:
        """Sphere.planeMap (Sphere.fs:46-52): the point of the sphere at surface coordinates (phi, theta) in [0, 1]^2."""
        t = theta * math.pi
        ph = phi * math.pi * 2.0 - math.pi
        return (centre[0] + radius * math.cos(ph) * math.sin(t), centre[1] - radius * math.cos(t),
                centre[2] - radius * math.sin(ph) * math.sin(t))

    @staticmethod
    def plane_map_inverse(radius, centre) -> PlaneMapInverse:
        return PlaneMapInverse(float(radius), tuple(float(x) for x in centre))


# ---- InfinitePlane.fs -------------------------------------------------------------------------
class InfinitePlaneStyle:
    @dataclass(frozen=True)
    class LightSource:
        texture: Any

    @dataclass(frozen=True)
    class PureReflection:
        albedo: float
        colour: Pixel

    @dataclass(frozen=True)
    class LambertReflection:
        albedo: float
        colour: Pixel
        rand: Any = None

    @dataclass(frozen=True)
    class FuzzedReflection:
        albedo: float
        colour: Pixel
        fuzz: float
        rand: Any = None


@dataclass(frozen=True)
class InfinitePlane:
    Style: Any
    Point: Tuple[float, float, float]
    Normal: Tuple[float, float, float]

    @staticmethod
    def make(style, point_on_plane, normal) -> "InfinitePlane":
        """InfinitePlane.make (InfinitePlane.fs:114-119).  `normal` must be a UnitVector."""
        n = np.asarray(normal, dtype=np.float64)
        if n.shape != (3,) or abs(float(n @ n) - 1.0) > 1e-6:
            raise ValueError("normal must be a unit vector (UnitVector in the reference)")
        return InfinitePlane(style, tuple(float(x) for x in point_on_plane), tuple(float(x) for x in n))


# ---- Hittable.fs ------------------------------------------------------------------------------
class Hittable:
    @dataclass(frozen=True)
    class Sphere:
        sphere: Sphere

    @dataclass(frozen=True)
    class UnboundedSphere:
        sphere: Sphere

    @dataclass(frozen=True)
    class InfinitePlane:
        plane: InfinitePlane


# ---- marshalling into the C ABI ----------------------------------------------------------------
class _TextureTable:
    def __init__(self):
        self.entries: List[abi.RtTexture] = []
        self.keep: List[np.ndarray] = []

    def add_param(self, interpret: PlaneMapInverse, t) -> int:
        e = abi.RtTexture()
        e.even = e.odd = -1
        e.map_centre[:] = interpret.centre
        e.map_radius = interpret.radius
        if isinstance(t, ParameterisedTexture.Colour):
            e.kind = abi.RT_TEX_COLOUR
            e.colour[:] = t.pixel.as_tuple()
        elif isinstance(t, ParameterisedTexture.Image):
            img = np.ascontiguousarray(t.img, dtype=np.uint8)
            self.keep.append(img)
            e.kind = abi.RT_TEX_IMAGE
            e.height, e.width = img.shape[0], img.shape[1]
            import ctypes as C
            e.rgb8 = img.ctypes.data_as(C.POINTER(C.c_uint8))
        elif isinstance(t, ParameterisedTexture.Checkered):
            ev = self.add_param(interpret, t.even)
            od = self.add_param(interpret, t.odd)
            e.kind = abi.RT_TEX_CHECKERED
            e.even, e.odd = ev, od
            e.grid_size = float(t.grid_size)
        else:
            raise NotImplementedError(
                "ParameterisedTexture.Arbitrary is a host closure and cannot be evaluated on the device "
                "(RT_ERR_UNSUPPORTED); sample it into an Image first with ParameterisedTexture.bake")
        self.entries.append(e)
        return len(self.entries) - 1


_HITTABLE = struct.Struct("<ii3d3d5di3Bx")  # RtHittable, include/rtfs_b200.h (layout checked in tests/test_host_cpu.py)
assert _HITTABLE.size == ctypes.sizeof(abi.RtHittable)
_TEXTURE_OFFSET = abi.RtHittable.texture.offset


def _texture_of(tex):
    """(constant colour, Texture.Parameterised or None) of a style's texture field."""
    if isinstance(tex, Texture.Colour):
        return tex.pixel.as_tuple(), None
    if isinstance(tex, Pixel):
        return tex.as_tuple(), None
    if isinstance(tex, Texture.Parameterised):
        return (0, 0, 0), tex
    raise NotImplementedError("Texture.Arbitrary is a host closure and cannot be evaluated on the device (RT_ERR_UNSUPPORTED)")


def _pack(obj):
    """The RtHittable record of one Hittable (texture index -1) and its Texture.Parameterised, if it has one.  The record
    is kept on the (immutable) object: Scene.make on a scene built once costs one memcpy per object afterwards, which is
    what the F# shim's marshalling loop costs."""
    albedo = fuzz = ior = prob = 0.0
    colour, ptex = (0, 0, 0), None
    n = (0.0, 0.0, 0.0)
    if isinstance(obj, (Hittable.Sphere, Hittable.UnboundedSphere)):
        s = obj.sphere
        shape = abi.RT_SHAPE_SPHERE if isinstance(obj, Hittable.Sphere) else abi.RT_SHAPE_UNBOUNDED_SPHERE
        p, radius, st = s.Centre, s.Radius, s.Style
        if isinstance(st, SphereStyle.LightSource):
            style = abi.RT_STYLE_LIGHT_SOURCE
            colour, ptex = _texture_of(st.texture)
        elif isinstance(st, SphereStyle.LightSourceCap):
            style, colour = abi.RT_STYLE_LIGHT_SOURCE_CAP, st.colour.as_tuple()
        elif isinstance(st, SphereStyle.PureReflection):
            style, albedo = abi.RT_STYLE_PURE_REFLECTION, st.albedo
            colour, ptex = _texture_of(st.texture)
        elif isinstance(st, SphereStyle.FuzzedReflection):
            style, albedo, fuzz = abi.RT_STYLE_FUZZED_REFLECTION, st.albedo, st.fuzz
            colour, ptex = _texture_of(st.texture)
        elif isinstance(st, SphereStyle.LambertReflection):
            style, albedo = abi.RT_STYLE_LAMBERT_REFLECTION, st.albedo
            colour, ptex = _texture_of(st.texture)
        elif isinstance(st, SphereStyle.Dielectric):
            style, albedo, ior, prob = abi.RT_STYLE_DIELECTRIC, st.albedo, st.boundaryRefractance, st.refraction
            colour, ptex = _texture_of(st.texture)
        elif isinstance(st, SphereStyle.Glass):
            style, albedo, ior = abi.RT_STYLE_GLASS, st.albedo, st.ior
            colour, ptex = _texture_of(st.texture)
        else:
            raise TypeError(f"unknown SphereStyle {st!r}")
    elif isinstance(obj, Hittable.InfinitePlane):
        pl = obj.plane
        shape, p, n, radius, st = abi.RT_SHAPE_INFINITE_PLANE, pl.Point, pl.Normal, 0.0, pl.Style
        if isinstance(st, InfinitePlaneStyle.LightSource):
            style = abi.RT_STYLE_LIGHT_SOURCE
            colour, ptex = _texture_of(st.texture)
        elif isinstance(st, InfinitePlaneStyle.PureReflection):
            style, albedo, colour = abi.RT_STYLE_PURE_REFLECTION, st.albedo, st.colour.as_tuple()
        elif isinstance(st, InfinitePlaneStyle.LambertReflection):
            style, albedo, colour = abi.RT_STYLE_LAMBERT_REFLECTION, st.albedo, st.colour.as_tuple()
        elif isinstance(st, InfinitePlaneStyle.FuzzedReflection):
            style, albedo, fuzz, colour = abi.RT_STYLE_FUZZED_REFLECTION, st.albedo, st.fuzz, st.colour.as_tuple()
        else:
            raise TypeError(f"unknown InfinitePlaneStyle {st!r}")
    else:
        raise TypeError(f"not a Hittable: {obj!r}")
    rec = (_HITTABLE.pack(shape, style, p[0], p[1], p[2], n[0], n[1], n[2], radius, albedo, fuzz, ior, prob, -1, colour[0], colour[1], colour[2]), ptex)
    object.__setattr__(obj, "_abi", rec)  # frozen dataclass: plain attribute, not a field
    return rec


def marshal(objects: Sequence[Any]):
    """Hittable array -> (RtHittable array, list[RtTexture], keepalive): what the F# shim's marshalling loop produces.
    The first result is a ctypes array (indexable and iterable like a list of RtHittable) over one contiguous buffer."""
    table = _TextureTable()
    recs, textured = [], []
    for i, obj in enumerate(objects):
        rec = getattr(obj, "_abi", None) or _pack(obj)
        recs.append(rec[0])
        if rec[1] is not None:
            textured.append((i, rec[1]))
    buf = bytearray(b"".join(recs))
    for i, tex in textured:
        struct.pack_into("<i", buf, i * _HITTABLE.size + _TEXTURE_OFFSET, table.add_param(tex.interpret, tex.texture))
    n = len(recs)
    out = (abi.RtHittable * n).from_buffer(buf) if n else (abi.RtHittable * 0)()
    return out, table.entries, table.keep
