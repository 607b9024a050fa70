"""How much of a warp's walk time is lost to the longest walk among its 32 lanes, and what other schedules would recover.

Runs the host build of the device functions (csrc/host_debug, the same traversal code) over a small frame with a log of the
slab tests of every ray, then replays the log through three schedules of a 32-lane warp:
  lockstep   what render_kernel does: every lane walks ONE ray per pass, the pass lasts as long as its longest walk
             (lanes regenerate a path as soon as theirs ends, so 32 rays are in flight in every pass);
  pool P     P rays are collected, the 32 lanes pull walks from the pool until it is dry (list scheduling), then P
             scatters are done at full width;
  quanta K   as pool, but a walk is cut into pieces of at most K node visits that go back to the pool.
Prints the lane efficiency of the walk under each.  Cost unit: one node visit (two slab tests) or one leaf visit."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import oracle_camera  # noqa: E402
from ray_tracing_fsharp_b200 import abi, sample_images  # noqa: E402
from ray_tracing_fsharp_b200.domain import marshal  # noqa: E402

lib = C.CDLL(os.path.join(ROOT, "ray_tracing_fsharp_b200", "csrc", "build", "librtfs_host_debug.so"))


def ray_log(spec, mw, mh, spp):
    spec.max_width_coord, spec.max_height_coord, spec.spp = mw, mh, spp
    hs, ts, keep = marshal(spec.objects)
    t = (abi.RtTexture * max(1, len(ts)))(*ts)
    cam = oracle_camera(spec)
    rows, cols = 2 * mh + 1, 2 * mw + 1
    rgb = np.zeros((rows, cols, 3), np.uint8)
    sums = np.zeros((rows, cols, 4), np.int32)
    cap = 40_000_000
    log = np.zeros(cap, np.uint32)
    n = C.c_uint64()
    rc = lib.dbg_render_logged(hs, len(hs), t, len(ts), C.byref(cam), mw, mh, C.c_uint64(5), 0, 0, C.c_void_p(rgb.ctypes.data),
                               C.c_void_p(sums.ctypes.data), C.c_void_p(log.ctypes.data), C.c_uint64(cap), C.byref(n))
    assert rc == 0
    return log[:min(cap, n.value)]


def list_schedule(work, lanes=32):
    t = np.zeros(lanes, np.int64)
    for v in work:
        t[t.argmin()] += v
    return int(t.max())


def simulate(log):
    visits = ((log & 0x7FFFFFFF) >> 8).astype(np.int64) // 2 + (log & 255).astype(np.int64)  # node visits + leaf / unbounded tests
    last = (log >> 31).astype(bool)
    ends = np.nonzero(last)[0]
    starts = np.concatenate([[0], ends[:-1] + 1])
    paths = [visits[a:b + 1] for a, b in zip(starts, ends)]
    total = int(visits.sum())
    # lockstep with regeneration: 32 lanes, each runs its paths back to back, one ray per pass
    lanes = [[] for _ in range(32)]
    for k in range(len(paths)):
        lanes[k % 32].extend(paths[k].tolist())
    m = max(len(x) for x in lanes)
    arr = np.zeros((32, m), np.int64)
    for i, x in enumerate(lanes):
        arr[i, :len(x)] = x
    lock = int(arr.max(0).sum())
    out = {"rays": len(visits), "mean visits": visits.mean(), "p99": np.percentile(visits, 99), "max": visits.max(),
           "lockstep efficiency": total / (32 * lock)}
    flat = visits
    for P in (64, 96, 128, 256):
        cost = sum(list_schedule(flat[a:a + P]) for a in range(0, len(flat) - P + 1, P))
        out[f"pool {P}"] = flat[:len(flat) // P * P].sum() / (32 * cost)
    for K in (4, 8, 16):
        P = 64
        cost = 0
        for a in range(0, len(flat) - P + 1, P):
            pieces = []
            for v in flat[a:a + P]:
                while v > K:
                    pieces.append(K)
                    v -= K
                pieces.append(v)
            cost += list_schedule(pieces)  # a piece can only follow its predecessor; ignoring that is slightly optimistic
        out[f"pool 64, quanta {K}"] = flat[:len(flat) // P * P].sum() / (32 * cost)
    return out


if __name__ == "__main__":
    for name, spec, mw, mh, spp in [("C2", sample_images.CONFIGS["C2"](), 60, 40, 8), ("C5 (100k spheres)", sample_images.CONFIGS["C5"](), 48, 27, 4),
                                    ("C4", sample_images.CONFIGS["C4"](), 48, 27, 8)]:
        res = simulate(ray_log(spec, mw, mh, spp))
        print(name, {k: (round(float(v), 3)) for k, v in res.items()}, flush=True)
