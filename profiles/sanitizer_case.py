"""A small exercise of every kernel for compute-sanitizer: megakernel (shared-memory and global-memory variants,
adaptive and not, with counters), wavefront, conformance entry points.  Checked against the oracle where cheap."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from ray_tracing_fsharp_b200 import abi, native, sample_images  # noqa: E402
from ray_tracing_fsharp_b200.domain import marshal  # noqa: E402
from ray_tracing_fsharp_b200.scene import Camera  # noqa: E402

for name, mw, mh, spp in [("C1", 30, 17, 16), ("C2", 24, 16, 14), ("C3", 20, 11, 12), ("C4", 20, 11, 12)]:
    spec = sample_images.CONFIGS[name]()
    cam = Camera.make_basic(spp, spec.focal_length, spec.aspect_ratio, spec.origin, spec.view_direction, spec.view_up)
    cam.bounce_depth = spec.bounce_depth
    hs, ts, keep = marshal(spec.objects)
    h = native.SceneHandle(hs, ts, 0, keepalive=keep)
    base = None
    for flags in (0, abi.RT_FLAG_NO_SMEM, abi.RT_FLAG_COUNTERS):
        for adaptive in (True, False):
            _, sums, st = h.render(cam, mw, mh, seed=1, adaptive=adaptive, flags=flags, want_sums=True)
            if adaptive:
                if base is None:
                    base = sums
                assert np.array_equal(base, sums)
    _, sums_w, _ = h.render(cam, mw, mh, seed=1, adaptive=True, mode=abi.RT_MODE_WAVEFRONT, want_sums=True)
    assert np.array_equal(base, sums_w)
    rng = np.random.default_rng(0)
    n = 2000
    o = rng.uniform(-5, 5, (n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    for trav in (0, 1):
        h.hit_object(o, d, traversal=trav)
    h.trace_samples(cam, mw, mh, 3, rng.integers(0, 2 * mh + 1, n), rng.integers(0, 2 * mw + 1, n), rng.integers(0, spp, n))
    h.close()
    print(name, "ok", flush=True)
print("fp32 peak", native.measure_fp32_peak(0))
