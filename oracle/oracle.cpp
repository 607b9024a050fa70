// oracle.cpp — CPU restatement of the reference's per-pixel render path.  TEST INFRASTRUCTURE ONLY.
//
// This file is the parity oracle for the CUDA path in ray_tracing_fsharp_b200/csrc.  It is a
// line-by-line restatement, in IEEE double, of the F# functions of Smaug123/ray-tracing-fsharp that
// lie on the render path (Scene.render -> renderPixel -> traceOnce -> traceRay -> hitObject and the
// intersection / scatter / texture / colour functions they call).  Every function cites the
// reference file:line it follows (paths relative to /root/reference).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
// this library.  Nothing under ray_tracing_fsharp_b200/ links, imports or calls it.
//
// PARITY PINNING.  The reference is F#/.NET and cannot be executed in this environment (no dotnet,
// mono or fsi), so the oracle is pinned against every known-answer vector the reference's own test
// project holds for this path (tests/test_oracle_reference_kats.py):
//   PpmOutputExample.txt (TestPpmOutput.fs:12-46), the sphere regression case
//   (TestSphereIntersection.fs:37-57), the three Glass/Dielectric scatter cases (TestSphere.fs:52-152),
//   the twelve planeMap / planeMapInverse pairs (TestSphere.fs:196-214), the AABB decision cases
//   (TestBoundingBox.fs:16-123), combine-with-white/black (TestPixel.fs:156-183) and the RNG range /
//   spread properties (TestRandom.fs:11-71); plus the reference's FsCheck properties re-run with
//   seeded generators.  Functions the reference never tests (hitObject, BoundingBoxTree.make, traceRay,
//   renderPixel, camera ray generation, InfinitePlane.*, darken, PixelStats.mean, gamma, image textures)
//   are restated from the source alone: for those rows parity is UNPINNED beyond code review, and
//   DESIGN.md says so.
//
// .NET semantics honoured: Math.Round = round-half-to-even (nearbyint in the default rounding mode);
// float -> int/byte conversion truncates; int division truncates; `**` is pow; Array.minBy returns the
// first minimum; Array.sortBy is unstable in .NET (we use a stable sort: the topology of the reference
// tree is therefore not reproduced bit-for-bit and does not need to be).
#include "../include/rtfs_b200.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <thread>
#include <vector>

namespace {

constexpr double PI = 3.14159265358979323846; // System.Math.PI

// ---------------------------------------------------------------------------------------------
// Float.fs:78-96 — tolerance predicates
// ---------------------------------------------------------------------------------------------
constexpr double TOL = 0.00000001;
enum Comparison { Greater, Equal, Less };
inline bool feq(double a, double b) { return std::fabs(a - b) < TOL; }  // Float.equal   :82
inline bool fpos(double a) { return a > TOL; }                          // Float.positive :86
inline Comparison fcmp(double a, double b) {                            // Float.compare :88-96
    if (std::fabs(a - b) < TOL) return Equal;
    if (a < b) return Less;
    return Greater;
}

// ---------------------------------------------------------------------------------------------
// Float.fs:13-76 — FloatProducer (xorshift128, byte-swapped, divided by UInt32.MaxValue)
// ---------------------------------------------------------------------------------------------
inline uint32_t xorshift_generate(uint32_t s[4]) { // generateInt32 :14-20
    uint32_t &x = s[0], &y = s[1], &z = s[2], &w = s[3];
    uint32_t t = x ^ (x << 11);
    x = y;
    y = z;
    z = w;
    w = w ^ (w >> 19) ^ (t ^ (t >> 8));
    return w;
}
inline uint32_t to_int(uint32_t w) { // toInt :22-27 (byte swap)
    uint32_t highest = (w & 0xFFu), second = ((w >> 8) & 0xFFu), third = ((w >> 16) & 0xFFu),
             lowest = ((w >> 24) & 0xFFu);
    return (highest << 24) ^ (second << 16) ^ (third << 8) ^ lowest;
}
inline double to_double(uint32_t i) { return double(i) / double(0xFFFFFFFFu); } // toDouble :29

// Philox4x32-10 (Salmon et al., SC'11; Random123).  NOT part of the reference: it replaces the
// reference's time-seeded shared xorshift streams (Scene.fs:205, SampleImages.fs:833-835) so that
// the oracle and the GPU draw identical uniforms for a given (seed, pixel, sample, bounce, retry).
inline void philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = uint64_t(0xD2511F53u) * c0;
        uint64_t p1 = uint64_t(0xCD9E8D57u) * c2;
        uint32_t n0 = uint32_t(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = uint32_t(p1);
        uint32_t n2 = uint32_t(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = uint32_t(p0);
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// The three ways the oracle can be fed uniforms.  All of them present FloatProducer's interface
// (Get / GetTwo / GetThree, Float.fs:38-76).
struct Rng {
    enum Kind { XORSHIFT, COUNTER, EXPLICIT } kind = XORSHIFT;
    // XORSHIFT
    uint32_t s[4] = {1, 2, 3, 4};
    // COUNTER: words = philox(ctr = {pixel, sample, bounce, retry}, key = {seed lo, seed hi})
    uint64_t seed = 0;
    uint32_t pixel = 0, sample = 0, bounce = 0, retry = 0;
    // EXPLICIT: 4 uniforms, rotated by one on every further draw
    double u[4] = {0, 0, 0, 0};
    int rot = 0;

    void block(double out[4]) {
        if (kind == XORSHIFT) {
            // caller picks how many it consumes; xorshift is sequential so draw lazily instead
            out[0] = out[1] = out[2] = out[3] = 0;
        } else if (kind == COUNTER) {
            uint32_t ctr[4] = {pixel, sample, bounce, retry}, key[2] = {uint32_t(seed), uint32_t(seed >> 32)}, w[4];
            philox4x32_10(ctr, key, w);
            for (int i = 0; i < 4; ++i) out[i] = to_double(w[i]);
            ++retry;
        } else {
            for (int i = 0; i < 4; ++i) out[i] = u[(i + rot) & 3];
            ++rot;
        }
    }
    double get() { // Get :38-47
        if (kind == XORSHIFT) return to_double(to_int(xorshift_generate(s)));
        double b[4];
        block(b);
        return b[0];
    }
    void get_two(double &a, double &b2) { // GetTwo :49-60
        if (kind == XORSHIFT) {
            uint32_t one = xorshift_generate(s), two = xorshift_generate(s);
            a = to_double(to_int(one));
            b2 = to_double(to_int(two));
            return;
        }
        double b[4];
        block(b);
        a = b[0];
        b2 = b[1];
    }
    void get_three(double &a, double &b2, double &c) { // GetThree :62-76
        if (kind == XORSHIFT) {
            uint32_t one = xorshift_generate(s), two = xorshift_generate(s), three = xorshift_generate(s);
            a = to_double(to_int(one));
            b2 = to_double(to_int(two));
            c = to_double(to_int(three));
            return;
        }
        double b[4];
        block(b);
        a = b[0];
        b2 = b[1];
        c = b[2];
    }
};

// ---------------------------------------------------------------------------------------------
// Point.fs — Point / Vector / UnitVector
// ---------------------------------------------------------------------------------------------
struct V3 {
    double x, y, z;
};
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }           // Vector.dot :18
inline V3 vsum(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }              // Vector.sum :20
inline V3 vscale(double s, V3 v) { return {s * v.x, s * v.y, s * v.z}; }              // Vector.scale :22-24
inline V3 vdiff(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }             // Vector.difference :26 / Point.differenceToThenFrom :91
inline V3 cross(V3 a, V3 b) {                                                         // Vector.cross :45-46
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - b.x * a.y};
}
inline bool unitise(V3 v, V3 &out) { // Vector.unitise :28-35
    double d = dot(v, v);
    if (feq(d, 0.0)) return false;
    double factor = 1.0 / std::sqrt(d);
    out = vscale(factor, v);
    return true;
}
inline double coord(V3 p, int i) { return i == 0 ? p.x : (i == 1 ? p.y : p.z); } // Point.coordinate :82-87

inline V3 unit_random(Rng &rng) { // UnitVector.random :49-59 (recursion = loop)
    for (;;) {
        double r1, r2, r3;
        rng.get_three(r1, r2, r3);
        double x = (2.0 * r1) - 1.0, y = (2.0 * r2) - 1.0, z = (2.0 * r3) - 1.0;
        V3 out;
        if (unitise(V3{x, y, z}, out)) return out;
    }
}

// ---------------------------------------------------------------------------------------------
// Ray.fs
// ---------------------------------------------------------------------------------------------
struct Ray {
    V3 o{0, 0, 0}, d{0, 0, 0};
};
inline bool overwrite_with_make(V3 origin, V3 vec, Ray &ray) { // Ray.overwriteWithMake :11-24
    double d = dot(vec, vec);
    if (feq(d, 0.0)) return false;
    ray.o = origin;
    double factor = 1.0 / std::sqrt(d);
    ray.d = vscale(factor, vec);
    return true;
}
inline bool ray_make_opt(V3 origin, V3 vec, Ray &out) { // Ray.make' :26-34
    V3 u;
    if (!unitise(vec, u)) return false;
    out = Ray{origin, u};
    return true;
}
inline V3 walk_along_ray(V3 o, V3 v, double m) { return {o.x + (v.x * m), o.y + (v.y * m), o.z + (v.z * m)}; } // :42-43
inline V3 walk_along(const Ray &r, double m) { return walk_along_ray(r.o, r.d, m); }                            // :45-46

// ---------------------------------------------------------------------------------------------
// Plane.fs
// ---------------------------------------------------------------------------------------------
struct OrthoPlane {
    V3 v1, v2, point;
};
inline bool make_normal_to(V3 point, V3 v, OrthoPlane &out) { // Plane.makeNormalTo :22-36
    V3 v1 = feq(v.z, 0.0) ? V3{0.0, 0.0, 1.0} : V3{1.0, 1.0, (-v.x - v.y) / v.z};
    V3 v2u, v1u;
    if (!unitise(cross(v, v1), v2u)) return false; // ValueOption.get would throw
    if (!unitise(v1, v1u)) return false;
    out = OrthoPlane{v1u, v2u, point};
    return true;
}
inline bool make_orthonormal_spanned_by(const Ray &r1, const Ray &r2, OrthoPlane &out) { // :64-79
    double coefficient = dot(r1.d, r2.d);
    V3 v2;
    if (!unitise(vdiff(r2.d, vscale(coefficient, r1.d)), v2)) return false;
    out = OrthoPlane{r1.d, v2, r1.o};
    return true;
}
inline bool plane_basis(V3 view_up, const OrthoPlane &plane, Ray &x_axis, Ray &y_axis) { // Plane.basis :82-97
    V3 up;
    if (!unitise(view_up, up)) return false;
    double v1c = dot(plane.v1, up), v2c = dot(plane.v2, up);
    V3 v2, v1;
    if (!unitise(vsum(vscale(v1c, plane.v1), vscale(v2c, plane.v2)), v2)) return false;
    if (!unitise(vsum(vscale(v2c, plane.v1), vscale(-v1c, plane.v2)), v1)) return false;
    x_axis = Ray{plane.point, v1};
    y_axis = Ray{plane.point, v2};
    return true;
}

// ---------------------------------------------------------------------------------------------
// Pixel.fs
// ---------------------------------------------------------------------------------------------
struct Pixel {
    uint8_t r, g, b;
};
constexpr Pixel BLACK{0, 0, 0}, WHITE{255, 255, 255}, HOTPINK{205, 105, 180}; // Pixel.fs:18-66
inline Pixel combine(Pixel a, Pixel b) {                                       // Pixel.combine :136-141
    return {uint8_t((int(a.r) * int(b.r)) / 255), uint8_t((int(a.g) * int(b.g)) / 255),
            uint8_t((int(a.b) * int(b.b)) / 255)};
}
inline uint8_t round_to_byte(double v) { return uint8_t(int64_t(std::nearbyint(v))); } // Math.Round |> byte
inline Pixel darken(double albedo, Pixel p) {                                          // Pixel.darken :144-151
    return {round_to_byte(double(p.r) * albedo), round_to_byte(double(p.g) * albedo),
            round_to_byte(double(p.b) * albedo)};
}
struct PixelStats { // Pixel.fs:68-108
    int count = 0, r = 0, g = 0, b = 0;
    void add(Pixel p) {
        count += 1;
        r += p.r;
        g += p.g;
        b += p.b;
    }
    Pixel mean() const { return {uint8_t(r / count), uint8_t(g / count), uint8_t(b / count)}; }
};
inline int pixel_difference(Pixel a, Pixel b) { // Pixel.difference :113-116
    return std::abs(int(a.r) - int(b.r)) + std::abs(int(a.g) - int(b.g)) + std::abs(int(a.b) - int(b.b));
}
inline uint8_t gamma_correct(uint8_t b) { // PixelOutput.correct ImageOutput.fs:11-18
    int i = int(std::nearbyint(std::sqrt(double(b) / 255.0) * 255.0));
    if (i == 256) i = 255;
    return uint8_t(i);
}

// ---------------------------------------------------------------------------------------------
// Texture.fs + Sphere.planeMap/planeMapInverse (Sphere.fs:47-61)
// ---------------------------------------------------------------------------------------------
inline V3 plane_map(double radius, V3 centre, double phi, double theta) { // Sphere.planeMap :47-52
    theta = theta * PI;
    phi = phi * PI * 2.0 - PI;
    V3 p{radius * std::cos(phi) * std::sin(theta), -radius * std::cos(theta), -radius * std::sin(phi) * std::sin(theta)};
    return vsum(p, centre);
}
inline void plane_map_inverse(double radius, V3 centre, V3 p, double &u, double &v) { // :55-61
    V3 q = vscale(1.0 / radius, vdiff(p, centre));
    double theta = std::acos(-q.y);
    double phi = std::atan2(-q.z, q.x) + PI;
    u = phi / (2.0 * PI);
    v = theta / PI;
}

struct Tex {
    int kind = RT_TEX_COLOUR;
    Pixel colour{0, 0, 0};
    int w = 0, h = 0;
    std::vector<uint8_t> img; // img[y][x] row-major
    int even = -1, odd = -1;
    double grid = 0;
    V3 map_centre{0, 0, 0};
    double map_radius = 1;
};

struct SceneData;
Pixel param_colour_at(const SceneData &sc, int tex, V3 centre, double radius, V3 p);

// ---------------------------------------------------------------------------------------------
// BoundingBox.fs
// ---------------------------------------------------------------------------------------------
struct Box {
    V3 mn, mx;
};
inline double box_volume(const Box &b) { // volume :13-16
    return (b.mx.x - b.mn.x) * (b.mx.y - b.mn.y) * (b.mx.z - b.mn.z);
}
inline V3 inverse_directions(const Ray &r) { return {1.0 / r.d.x, 1.0 / r.d.y, 1.0 / r.d.z}; } // :25-28
inline bool box_hits(V3 inv, const Ray &ray, const Box &box) { // BoundingBox.hits :30-94
    double t_min = -std::numeric_limits<double>::infinity();
    double t_max = std::numeric_limits<double>::infinity();
    {
        double t0 = (box.mn.x - ray.o.x) * inv.x, t1 = (box.mx.x - ray.o.x) * inv.x;
        if (inv.x < 0.0) std::swap(t0, t1);
        t_min = (t0 > t_min) ? t0 : t_min;
        t_max = (t1 < t_max) ? t1 : t_max;
        if (t_max < t_min || 0.0 >= t_max) return false;
    }
    {
        double t0 = (box.mn.y - ray.o.y) * inv.y, t1 = (box.mx.y - ray.o.y) * inv.y;
        if (inv.y < 0.0) std::swap(t0, t1);
        t_min = (t0 > t_min) ? t0 : t_min;
        t_max = (t1 < t_max) ? t1 : t_max;
        if (t_max < t_min || 0.0 >= t_max) return false;
    }
    double t0 = (box.mn.z - ray.o.z) * inv.z, t1 = (box.mx.z - ray.o.z) * inv.z;
    if (inv.z < 0.0) std::swap(t0, t1);
    t_min = (t0 > t_min) ? t0 : t_min;
    t_max = (t1 < t_max) ? t1 : t_max;
    return t_max >= t_min && t_max >= 0.0;
}
inline Box merge_two(const Box &i, const Box &j) { // mergeTwo :96-108 (F# min/max: NaN-propagating, irrelevant here)
    return {{std::min(i.mn.x, j.mn.x), std::min(i.mn.y, j.mn.y), std::min(i.mn.z, j.mn.z)},
            {std::max(i.mx.x, j.mx.x), std::max(i.mx.y, j.mx.y), std::max(i.mx.z, j.mx.z)}};
}

// ---------------------------------------------------------------------------------------------
// Scene objects: Sphere.fs:302-337, InfinitePlane.fs:101-119, Hittable.fs
// ---------------------------------------------------------------------------------------------
struct Obj {
    int shape, style;
    V3 p, n;
    double radius, radius_sq;
    double albedo, fuzz, ior, prob;
    int texture;
    Pixel colour;
    Box box; // Sphere.make :333-336 (Min/Max swapped for negative radius; not repaired)
};

struct Counters { // algorithmic work of the reference traversal
    uint64_t paths = 0, rays = 0, box_tests = 0, sphere_tests = 0, plane_tests = 0, candidates = 0;
    void add(const Counters &o) {
        paths += o.paths; rays += o.rays; box_tests += o.box_tests; sphere_tests += o.sphere_tests;
        plane_tests += o.plane_tests; candidates += o.candidates;
    }
};

struct LightRay {
    Ray ray;
    Pixel colour;
};

// Sphere.firstIntersection :349-386
inline bool sphere_first_intersection(V3 centre, double radius_sq, const Ray &ray, double &t_out) {
    V3 difference = vdiff(ray.o, centre);
    double b = dot(ray.d, difference);
    double c = dot(difference, difference) - radius_sq;
    double disc = (b * b - c);
    bool have = false;
    double ip = 0;
    switch (fcmp(disc, 0.0)) {
    case Equal:
        have = true;
        ip = -b;
        break;
    case Less:
        break;
    case Greater: {
        double intermediate = std::sqrt(disc);
        double i1 = intermediate - b;
        double i2 = -(b + intermediate);
        bool i1p = fpos(i1), i2p = fpos(i2);
        if (i1p && i2p) {
            switch (fcmp(i1, i2)) {
            case Less: ip = i1; break;
            case Greater: ip = i2; break;
            case Equal: ip = i1; break;
            }
            have = true;
        } else if (i1p) {
            ip = i1;
            have = true;
        } else if (i2p) {
            ip = i2;
            have = true;
        }
    } break;
    }
    if (!have) return false;
    if (fpos(ip)) {
        t_out = ip;
        return true;
    }
    return false;
}

// InfinitePlane.intersection :125-136
inline bool plane_intersection(V3 point, V3 normal, const Ray &ray, double &t_out) {
    double denominator = dot(normal, ray.d);
    if (feq(denominator, 0.0)) return false;
    double t = dot(normal, vdiff(point, ray.o)) / denominator;
    if (fpos(t)) {
        t_out = t;
        return true;
    }
    return false;
}

// Sphere.reflectWithoutFuzz :68-87
inline void reflect_without_fuzz(const Ray &normal, V3 strike, LightRay &in) {
    OrthoPlane plane;
    if (!make_orthonormal_spanned_by(normal, in.ray, plane)) {
        in.ray.d = vscale(-1.0, in.ray.d); // Ray.flipInPlace
        in.ray.o = strike;                 // Ray.translateToIntersect
    } else {
        double normal_component = -dot(plane.v1, in.ray.d);
        double tangent_component = dot(plane.v2, in.ray.d);
        V3 dest = walk_along_ray(walk_along_ray(plane.point, plane.v1, normal_component), plane.v2, tangent_component);
        overwrite_with_make(strike, vdiff(dest, strike), in.ray);
    }
}
// Sphere.addFuzz :89-104
inline void add_fuzz(double fuzz, Rng &rng, V3 strike, LightRay &reflected) {
    bool done = false;
    while (!done) {
        V3 offset = unit_random(rng);
        V3 sphere_centre = walk_along(reflected.ray, 1.0);
        V3 target = walk_along_ray(sphere_centre, offset, fuzz);
        done = overwrite_with_make(strike, vdiff(target, strike), reflected.ray);
    }
}
// Sphere.refract :108-146
inline void refract(bool inside, const Ray &normal, V3 strike, double incoming_cos, double index, LightRay &in) {
    index = inside ? 1.0 / index : index / 1.0;
    OrthoPlane plane;
    if (!make_orthonormal_spanned_by(normal, in.ray, plane)) {
        overwrite_with_make(strike, in.ray.d, in.ray);
        return;
    }
    double incoming_sin = std::sqrt(1.0 - incoming_cos * incoming_cos);
    double outgoing_sin = incoming_sin / index;
    if (fcmp(outgoing_sin, 1.0) == Greater) {
        reflect_without_fuzz(normal, strike, in);
        return;
    }
    double outgoing_cos = std::sqrt(1.0 - outgoing_sin * outgoing_sin);
    Ray tmp{walk_along(normal, -outgoing_cos), plane.v2};
    V3 outgoing_point = walk_along(tmp, outgoing_sin);
    overwrite_with_make(strike, vdiff(outgoing_point, strike), in.ray);
}

struct SceneData {
    std::vector<Obj> objs;
    std::vector<Tex> texs;
    std::vector<int> bounded, unbounded; // Scene.make partition, order preserving (Scene.fs:16-22)
    // BoundingBoxTree (BoundingBoxTree.fs:3-5) flattened in DFS pre-order: left child = i+1
    struct Node {
        Box box;
        int right; // -1 => leaf
        int prim;  // leaf: object index
    };
    std::vector<Node> nodes;
};

// Texture.colourAt (Texture.fs:12-15) over the structure the shim preserves
inline Pixel texture_colour_at(const SceneData &sc, const Obj &o, V3 point) {
    if (o.texture < 0) return o.colour;
    const Tex &t = sc.texs[o.texture];
    return param_colour_at(sc, o.texture, t.map_centre, t.map_radius, point);
}
// ParameterisedTexture.colourAt (Texture.fs:50-67); interpret = Sphere.planeMapInverse radius centre
Pixel param_colour_at(const SceneData &sc, int tex, V3 centre, double radius, V3 p) {
    const Tex &t = sc.texs[tex];
    switch (t.kind) {
    case RT_TEX_COLOUR:
        return t.colour;
    case RT_TEX_CHECKERED: {
        double x, y;
        plane_map_inverse(radius, centre, p, x, y);
        double sine = std::sin(t.grid * x) * std::sin(t.grid * y);
        if (fcmp(sine, 0.0) == Less) return param_colour_at(sc, t.even, centre, radius, p);
        return param_colour_at(sc, t.odd, centre, radius, p);
    }
    case RT_TEX_IMAGE: {
        double x, y;
        plane_map_inverse(radius, centre, p, x, y);
        int xi = int((1.0 - x) * double(t.w - 1));
        int yi = int(y * double(t.h - 1));
        const uint8_t *px = &t.img[(size_t(yi) * t.w + xi) * 3];
        return {px[0], px[1], px[2]};
    }
    }
    return BLACK;
}

// Sphere.reflection :150-300.  Returns true (absorbed) with `out` set, or false with `in` mutated.
inline bool sphere_reflection(const SceneData &sc, const Obj &s, Rng &rng, LightRay &in, V3 strike, Pixel &out,
                              bool *inside_out = nullptr) {
    bool flipped = (fcmp(s.radius, 0.0) == Less); // Sphere.fs:321
    bool inside = false;
    Ray normal;
    ray_make_opt(strike, vdiff(strike, s.p), normal); // Sphere.normal :65-66
    switch (fcmp(dot(vdiff(s.p, in.ray.o), vdiff(s.p, in.ray.o)), s.radius_sq)) { // :165-179
    case Equal:
    case Less:
        if (!flipped) {
            inside = true;
            normal.d = vscale(-1.0, normal.d);
        }
        break;
    case Greater:
        if (flipped) {
            inside = true;
            normal.d = vscale(-1.0, normal.d);
        }
        break;
    }
    if (inside_out) *inside_out = inside;

    switch (s.style) {
    case RT_STYLE_LIGHT_SOURCE: // :185-189
        out = combine(in.colour, texture_colour_at(sc, s, strike));
        return true;
    case RT_STYLE_LIGHT_SOURCE_CAP: { // :190-200
        double centre_coord = coord(s.p, 0);
        double lower = centre_coord + (s.radius - (s.radius / 4.0));
        double strike_coord = coord(strike, 0);
        out = (fcmp(strike_coord, lower) == Greater) ? combine(s.colour, in.colour) : BLACK;
        return true;
    }
    case RT_STYLE_LAMBERT_REFLECTION: { // :202-222
        in.colour = darken(s.albedo, combine(in.colour, texture_colour_at(sc, s, strike)));
        V3 sphere_centre = walk_along(normal, 1.0);
        bool done = false;
        while (!done) {
            V3 offset = unit_random(rng);
            V3 target = walk_along_ray(sphere_centre, offset, 1.0);
            done = overwrite_with_make(strike, vdiff(target, strike), in.ray);
        }
        return false;
    }
    case RT_STYLE_PURE_REFLECTION: { // :224-233
        Pixel darkened = darken(s.albedo, combine(in.colour, texture_colour_at(sc, s, strike)));
        reflect_without_fuzz(normal, strike, in);
        in.colour = darkened;
        return false;
    }
    case RT_STYLE_FUZZED_REFLECTION: { // :235-246
        in.colour = darken(s.albedo, combine(in.colour, texture_colour_at(sc, s, strike)));
        reflect_without_fuzz(normal, strike, in);
        add_fuzz(s.fuzz, rng, strike, in);
        return false;
    }
    case RT_STYLE_DIELECTRIC: { // :248-267
        Pixel nc = darken(s.albedo, combine(in.colour, texture_colour_at(sc, s, strike)));
        double rand = rng.get();
        if (rand > s.prob) {
            in.colour = nc;
            reflect_without_fuzz(normal, strike, in);
        } else {
            double incoming_cos = dot(in.ray.d, normal.d);
            refract(inside, normal, strike, incoming_cos, s.ior, in);
            in.colour = nc;
        }
        return false;
    }
    case RT_STYLE_GLASS: { // :269-300
        Pixel nc = darken(s.albedo, combine(in.colour, texture_colour_at(sc, s, strike)));
        double incoming_cos = dot(vscale(-1.0, in.ray.d), normal.d);
        double rand = rng.get();
        double refr = inside ? 1.0 / s.ior : s.ior;
        double param = (1.0 - refr) / (1.0 + refr);
        param = param * param;
        double reflection_prob = param + (1.0 - param) * std::pow(1.0 - incoming_cos, 5.0);
        if (rand < reflection_prob) {
            reflect_without_fuzz(normal, strike, in);
            in.colour = nc;
        } else {
            refract(inside, normal, strike, incoming_cos, s.ior, in);
            in.colour = nc;
        }
        return false;
    }
    }
    return false;
}

// InfinitePlane.pureOutgoing :18-38
inline bool plane_pure_outgoing(V3 strike, V3 normal, const Ray &incoming, Ray &out) {
    OrthoPlane plane;
    Ray nr{strike, normal};
    if (!make_orthonormal_spanned_by(nr, incoming, plane)) {
        out = Ray{strike, vscale(-1.0, incoming.d)}; // Ray.flip |> Ray.parallelTo strikePoint
        return true;
    }
    double normal_component = -dot(plane.v1, incoming.d);
    double tangent_component = dot(plane.v2, incoming.d);
    Ray tmp{walk_along(Ray{plane.point, plane.v1}, normal_component), plane.v2};
    V3 s = walk_along(tmp, tangent_component);
    return ray_make_opt(strike, vdiff(s, strike), out);
}
// InfinitePlane.reflection :43-99.  error=true where the reference would throw (ValueOption.get :86).
inline bool plane_reflection(const SceneData &sc, const Obj &pl, Rng &rng, LightRay &in, V3 strike, Pixel &out,
                             bool &error) {
    error = false;
    switch (pl.style) {
    case RT_STYLE_LIGHT_SOURCE: // :52-56
        out = combine(in.colour, texture_colour_at(sc, pl, strike));
        return true;
    case RT_STYLE_FUZZED_REFLECTION: { // :58-76
        Pixel nc = darken(pl.albedo, combine(in.colour, pl.colour));
        Ray pure;
        plane_pure_outgoing(strike, pl.n, in.ray, pure);
        Ray outgoing;
        bool have = false;
        while (!have) {
            V3 offset = unit_random(rng);
            V3 sphere_centre = walk_along(pure, 1.0);
            V3 target = walk_along(Ray{sphere_centre, offset}, pl.fuzz);
            have = ray_make_opt(strike, vdiff(target, strike), outgoing);
        }
        in.colour = nc;
        in.ray = outgoing;
        return false;
    }
    case RT_STYLE_LAMBERT_REFLECTION: { // :78-93
        V3 sphere_centre = walk_along(Ray{strike, pl.n}, 1.0);
        V3 offset = unit_random(rng);
        V3 target = walk_along(Ray{sphere_centre, offset}, 1.0);
        Ray outgoing;
        if (!ray_make_opt(strike, vdiff(target, strike), outgoing)) {
            error = true;
            return false;
        }
        in.colour = darken(pl.albedo, combine(in.colour, pl.colour));
        in.ray = outgoing;
        return false;
    }
    case RT_STYLE_PURE_REFLECTION: { // :95-99
        in.colour = darken(pl.albedo, combine(in.colour, pl.colour));
        Ray outgoing;
        plane_pure_outgoing(strike, pl.n, in.ray, outgoing);
        in.ray = outgoing;
        return false;
    }
    }
    error = true;
    return false;
}

// Hittable.hits :27-31
inline bool hittable_hits(const Obj &o, const Ray &ray, double &t, Counters &cn) {
    if (o.shape == RT_SHAPE_INFINITE_PLANE) {
        cn.plane_tests++;
        return plane_intersection(o.p, o.n, ray, t);
    }
    cn.sphere_tests++;
    return sphere_first_intersection(o.p, o.radius_sq, ray, t);
}

// BoundingBoxTree.make :9-43
int build_tree(SceneData &sc, std::vector<int> boxes) {
    Box bound_all = sc.objs[boxes[0]].box;
    for (size_t i = 1; i < boxes.size(); ++i) bound_all = merge_two(bound_all, sc.objs[boxes[i]].box); // Array.reduce mergeTwo
    int me = int(sc.nodes.size());
    sc.nodes.push_back({bound_all, -1, -1});
    if (boxes.size() == 1) {
        sc.nodes[me].box = sc.objs[boxes[0]].box; // Leaf boxes.[0]: the object's own box
        sc.nodes[me].prim = boxes[0];
        return me;
    }
    if (boxes.size() == 2) {
        build_tree(sc, {boxes[0]});
        int r = build_tree(sc, {boxes[1]});
        sc.nodes[me].right = r;
        return me;
    }
    std::vector<int> best_l, best_r;
    double best_v = 0;
    for (int axis = 0; axis < 3; ++axis) {
        std::vector<int> sorted = boxes;
        std::stable_sort(sorted.begin(), sorted.end(),
                         [&](int a, int b) { return coord(sc.objs[a].box.mn, axis) < coord(sc.objs[b].box.mn, axis); });
        size_t half = sorted.size() / 2;
        std::vector<int> l(sorted.begin(), sorted.begin() + half + 1), r(sorted.begin() + half + 1, sorted.end());
        Box lb = sc.objs[l[0]].box, rb = sc.objs[r[0]].box;
        for (size_t i = 1; i < l.size(); ++i) lb = merge_two(lb, sc.objs[l[i]].box);
        for (size_t i = 1; i < r.size(); ++i) rb = merge_two(rb, sc.objs[r[i]].box);
        double v = box_volume(lb) + box_volume(rb);
        if (axis == 0 || v < best_v) { // Array.minBy: first minimum
            best_v = v;
            best_l = l;
            best_r = r;
        }
    }
    build_tree(sc, best_l);
    int r = build_tree(sc, best_r);
    sc.nodes[me].right = r;
    return me;
}

struct Best {
    double best_float; // t^2
    int best_object;
    double best_length;
};
// Scene.bestCandidate :30-60 (recursive, exhaustive, left then right)
void best_candidate(const SceneData &sc, V3 inv, const Ray &ray, Best &b, int node, Counters &cn) {
    const SceneData::Node &nd = sc.nodes[node];
    cn.box_tests++;
    if (nd.right < 0) {
        if (box_hits(inv, ray, nd.box)) {
            double point;
            if (hittable_hits(sc.objs[nd.prim], ray, point, cn)) {
                cn.candidates++;
                double a = point * point;
                if (a < b.best_float) b = Best{a, nd.prim, point};
            }
        }
    } else if (box_hits(inv, ray, nd.box)) {
        best_candidate(sc, inv, ray, b, node + 1, cn);
        best_candidate(sc, inv, ray, b, nd.right, cn);
    }
}
// Scene.hitObject :62-91
inline bool hit_object(const SceneData &sc, const Ray &ray, int &obj, double &length, V3 &strike, Counters &cn) {
    Best b{std::numeric_limits<double>::infinity(), -1, std::numeric_limits<double>::quiet_NaN()};
    cn.rays++;
    if (!sc.nodes.empty()) best_candidate(sc, inverse_directions(ray), ray, b, 0, cn);
    for (int i : sc.unbounded) {
        double point;
        if (hittable_hits(sc.objs[i], ray, point, cn)) {
            cn.candidates++;
            double a = point * point;
            if (fcmp(a, b.best_float) == Less) b = Best{a, i, point};
        }
    }
    if (std::isnan(b.best_length)) return false;
    obj = b.best_object;
    length = b.best_length;
    strike = walk_along(ray, b.best_length);
    return true;
}

// Hittable.Reflection :8-12
inline bool hittable_reflection(const SceneData &sc, int obj, Rng &rng, LightRay &lr, V3 strike, Pixel &out, bool &error) {
    const Obj &o = sc.objs[obj];
    error = false;
    if (o.shape == RT_SHAPE_INFINITE_PLANE) return plane_reflection(sc, o, rng, lr, strike, out, error);
    return sphere_reflection(sc, o, rng, lr, strike, out);
}

// Scene.traceRay :93-114.  `rng_for_bounce` lets the counter RNG re-key per bounce.
inline Pixel trace_ray(const SceneData &sc, int max_count, LightRay &ray, Rng &rng, Counters &cn) {
    int bounces = 0;
    Pixel result = BLACK;
    bool done = false;
    while (bounces <= max_count && !done) {
        int obj;
        double len;
        V3 strike;
        if (!hit_object(sc, ray.ray, obj, len, strike, cn)) {
            done = true;
        } else {
            if (rng.kind == Rng::COUNTER) {
                rng.bounce = uint32_t(bounces + 1);
                rng.retry = 0;
            }
            Pixel colour;
            bool error;
            if (hittable_reflection(sc, obj, rng, ray, strike, colour, error)) {
                done = true;
                result = colour;
            } else {
                if (error) return BLACK; // reference throws; unreachable on non-degenerate input
                bounces += 1;
            }
        }
    }
    return done ? result : HOTPINK;
}

struct Cam {
    RtCamera c;
};

// traceOnce's ray generation, Scene.fs:129-144
inline bool camera_ray(const RtCamera &c, int max_w, int max_h, int row, int col, double rand1, double rand2, Ray &out) {
    double landing = ((double(col) + rand1) * c.viewport_width) / double(max_w);
    Ray xaxis{{c.xaxis_origin[0], c.xaxis_origin[1], c.xaxis_origin[2]}, {c.xaxis_dir[0], c.xaxis_dir[1], c.xaxis_dir[2]}};
    V3 point_on_x = walk_along(xaxis, landing);
    double walk = ((double(row) + rand2) * c.viewport_height) / double(max_h);
    V3 ydir{c.yaxis_dir[0], c.yaxis_dir[1], c.yaxis_dir[2]};
    V3 end_point = walk_along_ray(point_on_x, ydir, walk);
    V3 vo{c.view_origin[0], c.view_origin[1], c.view_origin[2]};
    return ray_make_opt(vo, vdiff(end_point, vo), out);
}

// Scene.traceOnce :118-155
inline void trace_once(const SceneData &sc, Rng &rng, const RtCamera &cam, int max_w, int max_h, int row, int col,
                       PixelStats &stats, Counters &cn) {
    double rand1, rand2;
    if (rng.kind == Rng::COUNTER) {
        rng.bounce = 0;
        rng.retry = 0;
    }
    rng.get_two(rand1, rand2);
    Ray ray;
    camera_ray(cam, max_w, max_h, row, col, rand1, rand2, ray);
    LightRay lr{ray, WHITE};
    cn.paths++;
    Pixel result = trace_ray(sc, cam.bounce_depth, lr, rng, cn);
    stats.add(result);
}

// Scene.renderPixel :157-194.  `adaptive = false` is NOT the reference: it always takes spp samples.
inline Pixel render_pixel(const SceneData &sc, Rng &rng, const RtCamera &cam, int max_w, int max_h, int row, int col,
                          bool adaptive, PixelStats &stats, Counters &cn) {
    uint32_t sample = 0;
    auto once = [&]() {
        if (rng.kind == Rng::COUNTER) rng.sample = sample;
        ++sample;
        trace_once(sc, rng, cam, max_w, max_h, row, col, stats, cn);
    };
    if (!adaptive) {
        for (int i = 0; i < cam.samples_per_pixel; ++i) once();
        return stats.mean();
    }
    int first_trial = std::min(5, cam.samples_per_pixel / 2);
    for (int i = 0; i <= first_trial; ++i) once();
    Pixel old_mean = stats.mean();
    for (int i = 1; i <= first_trial; ++i) once();
    Pixel new_mean = stats.mean();
    if (pixel_difference(new_mean, old_mean) == 0) return new_mean;
    for (int i = 1; i <= (cam.samples_per_pixel - 2 * first_trial - 1); ++i) once();
    return stats.mean();
}

std::string ppm_format(const uint8_t *rgb, int rows, int cols, bool gamma) { // ImageOutput.writePpm :163-197
    std::string s;
    s.reserve(size_t(rows) * cols * 12 + 32);
    char buf[64];
    s += "P3\n";
    snprintf(buf, sizeof buf, "%d %d\n", cols, rows);
    s += buf;
    s += "255\n";
    for (int r = 0; r < rows; ++r) {
        for (int c = 0; c < cols; ++c) {
            const uint8_t *p = rgb + (size_t(r) * cols + c) * 3;
            int R = p[0], G = p[1], B = p[2];
            if (gamma) {
                R = gamma_correct(p[0]);
                G = gamma_correct(p[1]);
                B = gamma_correct(p[2]);
            }
            snprintf(buf, sizeof buf, "%d %d %d", R, G, B);
            s += buf;
            if (c != cols - 1) s += " ";
        }
        if (r != rows - 1) s += "\n";
    }
    return s;
}

inline V3 v3(const double *p) { return {p[0], p[1], p[2]}; }
inline void put(double *p, V3 v) {
    p[0] = v.x;
    p[1] = v.y;
    p[2] = v.z;
}

} // namespace

// =================================================================================================
// C interface (loaded with ctypes by oracle/__init__.py)
// =================================================================================================
extern "C" {

struct OrcScene {
    SceneData sc;
};
struct OrcCounters {
    uint64_t paths, rays, box_tests, sphere_tests, plane_tests, candidates;
};

// ---- RNG ----
void orc_xorshift_words(uint32_t state[4], int n, uint32_t *raw_out, double *u_out) {
    for (int i = 0; i < n; ++i) {
        uint32_t w = xorshift_generate(state);
        if (raw_out) raw_out[i] = w;
        if (u_out) u_out[i] = to_double(to_int(w));
    }
}
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox4x32_10(ctr, key, out); }
void orc_counter_uniforms(uint64_t seed, int n, const uint32_t *pixel, const uint32_t *sample, const uint32_t *bounce,
                          const uint32_t *retry, uint32_t *words, double *u) {
    for (int i = 0; i < n; ++i) {
        uint32_t ctr[4] = {pixel[i], sample[i], bounce[i], retry[i]}, key[2] = {uint32_t(seed), uint32_t(seed >> 32)};
        philox4x32_10(ctr, key, words + 4 * i);
        for (int k = 0; k < 4; ++k) u[4 * i + k] = to_double(words[4 * i + k]);
    }
}

// ---- vectors / rays ----
int orc_unitise(const double v[3], double out[3]) {
    V3 o;
    if (!unitise(v3(v), o)) return 0;
    put(out, o);
    return 1;
}
void orc_unit_random_explicit(int n, const double *u3, double *out) { // UnitVector.random on explicit draws
    for (int i = 0; i < n; ++i) {
        Rng r;
        r.kind = Rng::EXPLICIT;
        r.u[0] = u3[3 * i]; r.u[1] = u3[3 * i + 1]; r.u[2] = u3[3 * i + 2]; r.u[3] = 0.123;
        put(out + 3 * i, unit_random(r));
    }
}
void orc_walk_along(const double o[3], const double d[3], double m, double out[3]) { put(out, walk_along_ray(v3(o), v3(d), m)); }
int orc_plane_orthonormal_basis(const double origin[3], const double v1[3], const double v2[3], const double up[3],
                                double x_out[3], double y_out[3]) {
    // Plane.orthonormalise (Plane.fs:40-55) then Plane.basis — TestPlane.fs:11-26
    V3 a = v3(v1), b = v3(v2);
    double coefficient = dot(a, b);
    V3 vec2;
    if (!unitise(vdiff(b, vscale(coefficient, a)), vec2)) return 0;
    OrthoPlane pl{a, vec2, v3(origin)};
    Ray x, y;
    if (!plane_basis(v3(up), pl, x, y)) return 0;
    put(x_out, x.d);
    put(y_out, y.d);
    return 1;
}

// ---- primitives ----
void orc_sphere_hit(int n, const double *o, const double *d, const double *c, const double *r, double *t) {
    for (int i = 0; i < n; ++i) {
        Ray ray{v3(o + 3 * i), v3(d + 3 * i)};
        double tt;
        t[i] = sphere_first_intersection(v3(c + 3 * i), r[i] * r[i], ray, tt) ? tt : std::numeric_limits<double>::quiet_NaN();
    }
}
void orc_plane_hit(int n, const double *o, const double *d, const double *p, const double *nrm, double *t) {
    for (int i = 0; i < n; ++i) {
        Ray ray{v3(o + 3 * i), v3(d + 3 * i)};
        double tt;
        t[i] = plane_intersection(v3(p + 3 * i), v3(nrm + 3 * i), ray, tt) ? tt : std::numeric_limits<double>::quiet_NaN();
    }
}
void orc_aabb_hit(int n, const double *o, const double *d, const double *bmin, const double *bmax, uint8_t *hit) {
    for (int i = 0; i < n; ++i) {
        Ray ray{v3(o + 3 * i), v3(d + 3 * i)};
        hit[i] = box_hits(inverse_directions(ray), ray, Box{v3(bmin + 3 * i), v3(bmax + 3 * i)}) ? 1 : 0;
    }
}
void orc_plane_map(double radius, const double centre[3], double phi, double theta, double out[3]) {
    put(out, plane_map(radius, v3(centre), phi, theta));
}
void orc_plane_map_inverse(double radius, const double centre[3], const double p[3], double uv[2]) {
    plane_map_inverse(radius, v3(centre), v3(p), uv[0], uv[1]);
}
void orc_combine(int n, const uint8_t *a, const uint8_t *b, uint8_t *out) {
    for (int i = 0; i < n; ++i) {
        Pixel p = combine({a[3 * i], a[3 * i + 1], a[3 * i + 2]}, {b[3 * i], b[3 * i + 1], b[3 * i + 2]});
        out[3 * i] = p.r; out[3 * i + 1] = p.g; out[3 * i + 2] = p.b;
    }
}
void orc_darken(int n, const double *albedo, const uint8_t *a, uint8_t *out) {
    for (int i = 0; i < n; ++i) {
        Pixel p = darken(albedo[i], {a[3 * i], a[3 * i + 1], a[3 * i + 2]});
        out[3 * i] = p.r; out[3 * i + 1] = p.g; out[3 * i + 2] = p.b;
    }
}
uint8_t orc_gamma_correct(uint8_t b) { return gamma_correct(b); }
void orc_stats_mean(const int32_t stats[4], uint8_t out[3]) { // PixelStats.mean
    PixelStats s;
    s.r = stats[0]; s.g = stats[1]; s.b = stats[2]; s.count = stats[3];
    Pixel p = s.mean();
    out[0] = p.r; out[1] = p.g; out[2] = p.b;
}
size_t orc_ppm_format(const uint8_t *rgb, int rows, int cols, int gamma, char *out, size_t cap) {
    std::string s = ppm_format(rgb, rows, cols, gamma != 0);
    if (out && cap) memcpy(out, s.data(), std::min(cap, s.size()));
    return s.size();
}

// ---- camera ----
int orc_camera_make_basic(int spp, double focal, double aspect, const double origin[3], const double view_dir[3],
                          const double view_up[3], RtCamera *out) { // Camera.makeBasic Camera.fs:34-59
    double height = 2.0;
    Ray view{v3(origin), v3(view_dir)};
    V3 corner = walk_along(view, focal);
    OrthoPlane view_plane;
    if (!make_normal_to(corner, v3(view_dir), view_plane)) return 0;
    Ray x_axis, y_axis;
    if (!plane_basis(v3(view_up), view_plane, x_axis, y_axis)) return 0;
    memset(out, 0, sizeof *out);
    put(out->view_origin, view.o);
    put(out->view_dir, view.d);
    put(out->xaxis_origin, x_axis.o);
    put(out->xaxis_dir, x_axis.d);
    put(out->yaxis_dir, y_axis.d);
    out->viewport_height = height;
    out->viewport_width = aspect * height;
    out->focal_length = focal;
    out->samples_per_pixel = spp;
    out->bounce_depth = 150;
    return 1;
}
void orc_camera_rays(const RtCamera *cam, int max_w, int max_h, int n, const int32_t *row, const int32_t *col,
                     const double *r1, const double *r2, double *o_out, double *d_out) {
    for (int i = 0; i < n; ++i) {
        Ray ray;
        camera_ray(*cam, max_w, max_h, row[i], col[i], r1[i], r2[i], ray);
        put(o_out + 3 * i, ray.o);
        put(d_out + 3 * i, ray.d);
    }
}

// ---- scene ----
OrcScene *orc_scene_create(const RtHittable *objs, int n, const RtTexture *texs, int ntex) { // Scene.make :15-28
    auto *s = new OrcScene();
    SceneData &sc = s->sc;
    for (int i = 0; i < ntex; ++i) {
        Tex t;
        t.kind = texs[i].kind;
        t.colour = {texs[i].colour[0], texs[i].colour[1], texs[i].colour[2]};
        t.w = texs[i].width;
        t.h = texs[i].height;
        if (t.kind == RT_TEX_IMAGE) t.img.assign(texs[i].rgb8, texs[i].rgb8 + size_t(t.w) * t.h * 3);
        t.even = texs[i].even;
        t.odd = texs[i].odd;
        t.grid = texs[i].grid_size;
        t.map_centre = v3(texs[i].map_centre);
        t.map_radius = texs[i].map_radius;
        sc.texs.push_back(std::move(t));
    }
    for (int i = 0; i < n; ++i) {
        const RtHittable &h = objs[i];
        Obj o{};
        o.shape = h.shape;
        o.style = h.style;
        o.p = v3(h.p);
        o.n = v3(h.n);
        o.radius = h.radius;
        o.radius_sq = h.radius * h.radius; // Sphere.make :326
        o.albedo = h.albedo;
        o.fuzz = h.fuzz;
        o.ior = h.ior;
        o.prob = h.prob;
        o.texture = h.texture;
        o.colour = {h.colour[0], h.colour[1], h.colour[2]};
        o.box = Box{vsum(o.p, V3{-h.radius, -h.radius, -h.radius}), vsum(o.p, V3{h.radius, h.radius, h.radius})};
        sc.objs.push_back(o);
        if (h.shape == RT_SHAPE_SPHERE) sc.bounded.push_back(i); // Hittable.BoundingBox :14-18
        else sc.unbounded.push_back(i);
    }
    if (!sc.bounded.empty()) build_tree(sc, sc.bounded);
    return s;
}
void orc_scene_destroy(OrcScene *s) { delete s; }
int orc_scene_bvh_node_count(const OrcScene *s) { return int(s->sc.nodes.size()); }
void orc_scene_bvh_nodes(const OrcScene *s, double *bounds, int32_t *right, int32_t *prim) {
    for (size_t i = 0; i < s->sc.nodes.size(); ++i) {
        const auto &nd = s->sc.nodes[i];
        put(bounds + 6 * i, nd.box.mn);
        put(bounds + 6 * i + 3, nd.box.mx);
        right[i] = nd.right;
        prim[i] = nd.prim;
    }
}

void orc_hit_object(const OrcScene *s, int n, const double *o, const double *d, int32_t *prim_out, double *t_out,
                    double *strike_out, OrcCounters *counters) {
    Counters cn;
    for (int i = 0; i < n; ++i) {
        Ray ray{v3(o + 3 * i), v3(d + 3 * i)};
        int obj;
        double len;
        V3 strike;
        if (hit_object(s->sc, ray, obj, len, strike, cn)) {
            prim_out[i] = obj;
            t_out[i] = len;
            if (strike_out) put(strike_out + 3 * i, strike);
        } else {
            prim_out[i] = -1;
            t_out[i] = std::numeric_limits<double>::quiet_NaN();
            if (strike_out) put(strike_out + 3 * i, V3{0, 0, 0});
        }
    }
    if (counters) *counters = OrcCounters{cn.paths, cn.rays, cn.box_tests, cn.sphere_tests, cn.plane_tests, cn.candidates};
}
// second-best margin helper for tests: t of every object hit (brute force over all objects, ignoring boxes)
void orc_all_hits(const OrcScene *s, const double o[3], const double d[3], double *t_per_object) {
    Counters cn;
    Ray ray{v3(o), v3(d)};
    for (size_t i = 0; i < s->sc.objs.size(); ++i) {
        double t;
        t_per_object[i] = hittable_hits(s->sc.objs[i], ray, t, cn) ? t : std::numeric_limits<double>::quiet_NaN();
    }
}

void orc_reflection(const OrcScene *s, int n, const int32_t *prim, const double *o, const double *d, const double *strike,
                    const uint8_t *colour_in, const double *uniforms, uint8_t *absorbed, uint8_t *colour_out,
                    double *o_out, double *d_out, uint8_t *inside_out) {
    for (int i = 0; i < n; ++i) {
        Rng rng;
        rng.kind = Rng::EXPLICIT;
        for (int k = 0; k < 4; ++k) rng.u[k] = uniforms[4 * i + k];
        LightRay lr{Ray{v3(o + 3 * i), v3(d + 3 * i)}, Pixel{colour_in[3 * i], colour_in[3 * i + 1], colour_in[3 * i + 2]}};
        Pixel out{0, 0, 0};
        const Obj &ob = s->sc.objs[prim[i]];
        bool inside = false, error = false, abs_;
        if (ob.shape == RT_SHAPE_INFINITE_PLANE) abs_ = plane_reflection(s->sc, ob, rng, lr, v3(strike + 3 * i), out, error);
        else abs_ = sphere_reflection(s->sc, ob, rng, lr, v3(strike + 3 * i), out, &inside);
        absorbed[i] = abs_ ? 1 : (error ? 2 : 0);
        Pixel c = abs_ ? out : lr.colour;
        colour_out[3 * i] = c.r; colour_out[3 * i + 1] = c.g; colour_out[3 * i + 2] = c.b;
        put(o_out + 3 * i, lr.ray.o);
        put(d_out + 3 * i, lr.ray.d);
        if (inside_out) inside_out[i] = inside ? 1 : 0;
    }
}
// Sphere.reflection called directly with explicit parameters (the form TestSphere.fs:52-152 uses)
int orc_sphere_reflection_direct(int style, double albedo, const uint8_t tex_colour[3], double ior, double prob, double fuzz,
                                 const double centre[3], double radius, const double o[3], const double d[3],
                                 const double strike[3], const uint8_t colour_in[3], const double uniforms[4],
                                 uint8_t colour_out[3], double o_out[3], double d_out[3]) {
    SceneData sc;
    Obj ob{};
    ob.shape = RT_SHAPE_SPHERE;
    ob.style = style;
    ob.p = v3(centre);
    ob.radius = radius;
    ob.radius_sq = radius * radius;
    ob.albedo = albedo;
    ob.fuzz = fuzz;
    ob.ior = ior;
    ob.prob = prob;
    ob.texture = -1;
    ob.colour = {tex_colour[0], tex_colour[1], tex_colour[2]};
    Rng rng;
    rng.kind = Rng::EXPLICIT;
    for (int k = 0; k < 4; ++k) rng.u[k] = uniforms[k];
    LightRay lr{Ray{v3(o), v3(d)}, Pixel{colour_in[0], colour_in[1], colour_in[2]}};
    Pixel out{0, 0, 0};
    bool absorbed = sphere_reflection(sc, ob, rng, lr, v3(strike), out);
    Pixel c = absorbed ? out : lr.colour;
    colour_out[0] = c.r; colour_out[1] = c.g; colour_out[2] = c.b;
    put(o_out, lr.ray.o);
    put(d_out, lr.ray.d);
    return absorbed ? 1 : 0;
}
void orc_texture(const OrcScene *s, int n, const int32_t *prim, const double *point, uint8_t *colour_out) {
    for (int i = 0; i < n; ++i) {
        Pixel p = texture_colour_at(s->sc, s->sc.objs[prim[i]], v3(point + 3 * i));
        colour_out[3 * i] = p.r; colour_out[3 * i + 1] = p.g; colour_out[3 * i + 2] = p.b;
    }
}

// traceOnce for explicit (row index, col index, sample) with the counter RNG
void orc_trace_samples(const OrcScene *s, const RtCamera *cam, int max_w, int max_h, uint64_t seed, int n,
                       const int32_t *row_idx, const int32_t *col_idx, const int32_t *sample, uint8_t *colour_out,
                       int32_t *rays_out) {
    int cols = 2 * max_w + 1;
    for (int i = 0; i < n; ++i) {
        Rng rng;
        rng.kind = Rng::COUNTER;
        rng.seed = seed;
        rng.pixel = uint32_t(row_idx[i] * cols + col_idx[i]);
        rng.sample = uint32_t(sample[i]);
        PixelStats st;
        Counters cn;
        trace_once(s->sc, rng, *cam, max_w, max_h, max_h - row_idx[i] - 1, col_idx[i] - max_w, st, cn);
        colour_out[3 * i] = uint8_t(st.r); colour_out[3 * i + 1] = uint8_t(st.g); colour_out[3 * i + 2] = uint8_t(st.b);
        if (rays_out) rays_out[i] = int32_t(cn.rays);
    }
}

// Scene.render :196-236 + Image.render (Domain.fs:23-24): one task per row, `threads` workers.
// rng_mode: 0 = xorshift128 per worker thread seeded from `seed` (the reference's generator without its
//               cross-thread lock; statistically equivalent, not sample-identical to any F# run),
//           1 = counter RNG keyed (seed, pixel, sample, bounce, retry) — sample-identical to the GPU.
// Rows row_begin, row_begin+row_step, ... only (row_step = 1: whole image) — used for bounded CPU timing.
int orc_render(const OrcScene *s, const RtCamera *cam, int max_w, int max_h, uint64_t seed, int rng_mode, int adaptive,
               int threads, int row_begin, int row_step, uint8_t *rgb_out, int32_t *stats_out, OrcCounters *counters) {
    int rows = 2 * max_h + 1, cols = 2 * max_w + 1; // :208-209
    if (threads < 1) threads = 1;
    if (row_step < 1) row_step = 1;
    std::atomic<int> next{0};
    std::vector<int> row_list;
    for (int r = row_begin; r < rows; r += row_step) row_list.push_back(r);
    std::vector<Counters> per_thread(threads);
    auto worker = [&](int tid) {
        Rng rng;
        if (rng_mode == 0) {
            rng.kind = Rng::XORSHIFT;
            // System.Random.Next() gives four non-negative 31-bit ints (Float.fs:33-36); derive them from the seed
            uint32_t ctr[4] = {uint32_t(tid), 0x5eed, 0, 0}, key[2] = {uint32_t(seed), uint32_t(seed >> 32)}, w[4];
            philox4x32_10(ctr, key, w);
            for (int i = 0; i < 4; ++i) rng.s[i] = w[i] & 0x7FFFFFFFu;
            if ((rng.s[0] | rng.s[1] | rng.s[2] | rng.s[3]) == 0) rng.s[0] = 1;
        } else {
            rng.kind = Rng::COUNTER;
            rng.seed = seed;
        }
        Counters &cn = per_thread[tid];
        for (;;) {
            int k = next.fetch_add(1);
            if (k >= int(row_list.size())) break;
            int row_idx = row_list[k];
            int row = max_h - row_idx - 1; // :219
            for (int col_idx = 0; col_idx < cols; ++col_idx) {
                int col = col_idx - max_w; // :226
                if (rng.kind == Rng::COUNTER) rng.pixel = uint32_t(row_idx * cols + col_idx);
                PixelStats st;
                Pixel p = render_pixel(s->sc, rng, *cam, max_w, max_h, row, col, adaptive != 0, st, cn);
                size_t idx = size_t(row_idx) * cols + col_idx;
                if (rgb_out) {
                    rgb_out[idx * 3] = p.r; rgb_out[idx * 3 + 1] = p.g; rgb_out[idx * 3 + 2] = p.b;
                }
                if (stats_out) {
                    stats_out[idx * 4] = st.r; stats_out[idx * 4 + 1] = st.g; stats_out[idx * 4 + 2] = st.b;
                    stats_out[idx * 4 + 3] = st.count;
                }
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(worker, t);
    worker(0);
    for (auto &t : pool) t.join();
    Counters total;
    for (auto &c : per_thread) total.add(c);
    if (counters)
        *counters = OrcCounters{total.paths, total.rays, total.box_tests, total.sphere_tests, total.plane_tests, total.candidates};
    return int(row_list.size());
}

// The sample-split decomposition of Scene.renderPixel that the multi-GPU path uses,
// restated on the CPU so that "sum over ranks == unsplit render" can be checked without GPUs.
//   phase 1: rank r probes the 8x4 tiles t with t mod world == r: the first 2*firstTrial+1 samples of
//            renderPixel (Scene.fs:172-182) go into stats, flags[p] = 1 where the two means differ (:183-188);
//   phase 2: for every flagged pixel, rank r adds the samples n_probe + r + j*world < spp (:191-192).
// With adaptive = 0 phase 1 only raises flags (rank 0) and phase 2 shares out all spp samples.
// The caller sums stats over ranks and takes the maximum of flags between the phases.
void orc_render_split(const OrcScene *s, const RtCamera *cam, int max_w, int max_h, uint64_t seed, int adaptive, int phase,
                      int rank, int world, int32_t *stats, uint8_t *flags) {
    const int rows = 2 * max_h + 1, cols = 2 * max_w + 1;
    const int tiles_x = (cols + 7) / 8;
    const int spp = cam->samples_per_pixel;
    const int first_trial = std::min(5, spp / 2);
    const int n_probe = adaptive ? 2 * first_trial + 1 : 0;
    const int sample_end = adaptive ? std::max(n_probe, spp) : spp;
    Counters cn;
    Rng rng;
    rng.kind = Rng::COUNTER;
    rng.seed = seed;
    auto trace = [&](int row_idx, int col_idx, int sample, PixelStats &st) {
        rng.pixel = uint32_t(row_idx * cols + col_idx);
        rng.sample = uint32_t(sample);
        trace_once(s->sc, rng, *cam, max_w, max_h, max_h - row_idx - 1, col_idx - max_w, st, cn);
    };
    for (int row_idx = 0; row_idx < rows; ++row_idx) {
        for (int col_idx = 0; col_idx < cols; ++col_idx) {
            const size_t p = size_t(row_idx) * cols + col_idx;
            const int tile = (row_idx / 4) * tiles_x + col_idx / 8;
            if (phase == 1) {
                if (!adaptive) {
                    if (rank == 0) flags[p] = 1;
                    continue;
                }
                if (tile % world != rank) continue;
                PixelStats st;
                for (int i = 0; i <= first_trial; ++i) trace(row_idx, col_idx, i, st);
                Pixel old_mean = st.mean();
                for (int i = 1; i <= first_trial; ++i) trace(row_idx, col_idx, first_trial + i, st);
                Pixel new_mean = st.mean();
                stats[4 * p] = st.r; stats[4 * p + 1] = st.g; stats[4 * p + 2] = st.b; stats[4 * p + 3] = st.count;
                flags[p] = (pixel_difference(new_mean, old_mean) != 0 && sample_end > n_probe) ? 1 : 0;
            } else {
                if (!flags[p]) continue;
                PixelStats st;
                for (int smp = n_probe + rank; smp < sample_end; smp += world) trace(row_idx, col_idx, smp, st);
                stats[4 * p] += st.r; stats[4 * p + 1] += st.g; stats[4 * p + 2] += st.b; stats[4 * p + 3] += st.count;
            }
        }
    }
}

} // extern "C"
