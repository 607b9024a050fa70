"""Where the end-to-end time of bench.py's `e2e` goes: the mirrored public API (Scene.make, Scene.render, Image.render)
stage by stage, next to the bare ABI calls (marshal, rt_scene_create, rt_render into a caller-owned buffer)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from ray_tracing_fsharp_b200 import native, sample_images  # noqa: E402
from ray_tracing_fsharp_b200.domain import marshal  # noqa: E402
from ray_tracing_fsharp_b200.scene import Camera, Image, Scene  # noqa: E402

spec = sample_images.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C2"]()
cam = Camera.make_basic(spec.spp, spec.focal_length, spec.aspect_ratio, spec.origin, spec.view_direction, spec.view_up)
cam.bounce_depth = spec.bounce_depth
mw, mh = spec.max_width_coord, spec.max_height_coord
native.lib()
native.device_count()
rgb = np.empty((spec.rows, spec.cols, 3), np.uint8)
T = time.perf_counter
for i in range(4):
    t0 = T(); hs, ts, keep = marshal(spec.objects)
    t1 = T(); h = native.SceneHandle(hs, ts, 0, keepalive=keep)
    t2 = T(); _, _, st = h.render(cam, mw, mh, seed=i, rgb_out=rgb)
    t3 = T(); h.close()
    t4 = T()
    print(f"ABI:    marshal {1e3 * (t1 - t0):5.2f}  create {1e3 * (t2 - t1):5.2f}  rt_render {1e3 * (t3 - t2):6.2f} (kernels {st.kernel_ms:.2f}, stream total {st.total_ms:.2f})  "
          f"destroy {1e3 * (t4 - t3):5.2f}  sum {1e3 * (t4 - t0):6.2f} ms")
for i in range(4):
    t0 = T(); sc = Scene.make(spec.objects, device=0)
    t1 = T(); _, image = Scene.render(lambda _p: None, lambda _s: None, mw, mh, cam, sc, seed=10 + i)
    t2 = T(); px = Image.render(image)
    t3 = T(); sc.handle.close()
    t4 = T()
    st = sc.last_stats
    print(f"mirror: Scene.make {1e3 * (t1 - t0):5.2f}  Scene.render {1e3 * (t2 - t1):5.2f}  Image.render {1e3 * (t3 - t2):6.2f} (kernels {st.kernel_ms:.2f}, stream total {st.total_ms:.2f})  "
          f"close {1e3 * (t4 - t3):5.2f}  sum {1e3 * (t4 - t0):6.2f} ms")
