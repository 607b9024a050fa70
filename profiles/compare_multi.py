"""One process, N devices (rt_multi_*: peer-memory flags + fused reduce/finalize, no collective library): frame time of
C2 for N = 1, 2, 4, 8 (as many as the box has), and a bit-identity check of the images."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from ray_tracing_fsharp_b200 import native, sample_images  # noqa: E402
from ray_tracing_fsharp_b200.domain import marshal  # noqa: E402
from ray_tracing_fsharp_b200.scene import Camera  # noqa: E402

spec = sample_images.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C2"]()
cam = Camera.make_basic(spec.spp, spec.focal_length, spec.aspect_ratio, spec.origin, spec.view_direction, spec.view_up)
cam.bounce_depth = spec.bounce_depth
hs, ts, keep = marshal(spec.objects)
n_dev = native.device_count()
ref = None
base = None
for n in [k for k in (1, 2, 4, 8) if k <= n_dev]:
    m = native.MultiHandle(hs, ts, list(range(n)), keepalive=keep)
    best_k, best_t, best_w = 1e9, 1e9, 1e9
    for rep in range(4):
        t0 = time.perf_counter()
        rgb, _, st = m.render(cam, spec.max_width_coord, spec.max_height_coord, seed=9)
        w = (time.perf_counter() - t0) * 1e3
        best_k, best_t, best_w = min(best_k, st.kernel_ms), min(best_t, st.total_ms), min(best_w, w)
    if ref is None:
        ref, base = rgb.copy(), best_t
    print(f"{n} device(s): device time {best_k:8.2f} ms  call {best_t:8.2f} ms  wall {best_w:8.2f} ms  {st.rays / best_t / 1e3:9.0f} Mrays/s  "
          f"efficiency {base / n / best_t:.3f}  identical image: {bool(np.array_equal(rgb, ref))}  launches {st.launches}", flush=True)
    m.close()
