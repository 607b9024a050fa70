"""Worker of tests/test_host_debug_asan.py: runs under LD_PRELOAD=libasan with the host-compiled device functions
(csrc/build/librtfs_host_debug.so, ASan + UBSan) and compares them with the oracle.  Exit code 0 = all good."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from helpers import camera_sample_rays, f32, fp32_safe_closest_hit, oracle_camera, random_unit_vectors, small_random_spheres  # noqa: E402
from ray_tracing_fsharp_b200 import abi, sample_images  # noqa: E402
from ray_tracing_fsharp_b200.domain import marshal  # noqa: E402

LIB = os.path.join(ROOT, "ray_tracing_fsharp_b200", "csrc", "build", "librtfs_host_debug.so")


def vp(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


def arrays(hs, ts):
    h = hs if isinstance(hs, C.Array) and len(hs) else (abi.RtHittable * max(1, len(hs)))(*hs)
    t = (abi.RtTexture * max(1, len(ts)))(*ts)
    return h, t


def main():
    mode = sys.argv[1]
    lib = C.CDLL(LIB)
    if mode == "overflow":
        lib.dbg_stack_overflow(int(sys.argv[2]))  # must not return when the argument exceeds the stack
        print("returned")
        return 0
    if mode == "frames":
        cases = [("reduced", small_random_spheres(), 30, 20, 12), ("C1", sample_images.CONFIGS["C1"](), 40, 22, 16),
                 ("C2", sample_images.CONFIGS["C2"](), 24, 16, 12), ("C3", sample_images.CONFIGS["C3"](), 32, 18, 12),
                 ("C4", sample_images.CONFIGS["C4"](), 24, 13, 16), ("mid", sample_images.many_spheres(n=3000), 24, 13, 12)]
        for name, spec, mw, mh, spp in cases:
            spec.max_width_coord, spec.max_height_coord, spec.spp = mw, mh, spp
            hs, ts, keep = marshal(spec.objects)
            h, t = arrays(hs, ts)
            cam = oracle_camera(spec)
            ref, ref_stats, counters, _ = oracle.Scene(hs, ts).render(cam, mw, mh, seed=5, rng_mode=1, adaptive=True, threads=1)
            rows, cols = 2 * mh + 1, 2 * mw + 1
            prev = None
            for wide in (0, 1):
                rgb = np.zeros((rows, cols, 3), np.uint8)
                sums = np.zeros((rows, cols, 4), np.int32)
                rays = C.c_uint64()
                rc = lib.dbg_render(h, len(hs), t, len(ts), C.byref(cam), mw, mh, C.c_uint64(5), 1, wide, vp(rgb), vp(sums), C.byref(rays))
                assert rc == 0, rc
                same = (rgb == ref).all(2).mean()
                print(f"{name} wide={wide}: {same:.4f} of pixels byte-identical to the oracle, rays {rays.value} vs {counters['rays']}", flush=True)
                assert same > 0.97, (name, wide, same)
                assert abs(rays.value - counters["rays"]) <= 0.02 * counters["rays"]
                if prev is not None:
                    assert np.array_equal(prev, sums), f"{name}: the wide and the binary tree disagree"
                prev = sums
        return 0
    if mode == "hits":
        for name, spec in [("reduced", small_random_spheres()), ("C2", sample_images.CONFIGS["C2"]()), ("mid", sample_images.many_spheres(n=3000))]:
            hs, ts, keep = marshal(spec.objects)
            h, _ = arrays(hs, ts)
            cam = oracle_camera(spec)
            rng = np.random.default_rng(8)
            n = 6000
            o1, d1 = camera_sample_rays(spec, cam, rng, n // 2)
            o2 = f32(np.stack([rng.uniform(-8, 8, n // 2), rng.uniform(0.45, 2.5, n // 2), rng.uniform(-8, 8, n // 2)], 1))
            d2 = random_unit_vectors(rng, n // 2)
            o, d = np.concatenate([o1, o2]), np.concatenate([d1, d2])
            wp, wt, _, _ = oracle.Scene(hs, ts).hit_object(o, d)
            safe = fp32_safe_closest_hit(hs, o, d, wp, wt)
            of, df = np.ascontiguousarray(o, np.float32), np.ascontiguousarray(d, np.float32)
            for wide in (0, 1):
                gp = np.zeros(n, np.int32)
                gt = np.zeros(n, np.float32)
                rc = lib.dbg_hit_object(h, len(hs), wide, n, vp(of), vp(df), vp(gp), vp(gt))
                assert rc == 0
                bad = int(((wp != gp) & safe).sum())
                print(f"{name} wide={wide}: {bad} FP32-safe rays of {int(safe.sum())} disagree with the oracle", flush=True)
                assert bad == 0
        return 0
    raise SystemExit("unknown mode")


if __name__ == "__main__":
    sys.exit(main())
