"""Times one rank's share of the C2 frame on ONE GPU for world = 1, 2, 4, 8 (no collectives): what each phase costs
per rank and how far it is from the ideal 1/world of the single-GPU time."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ray_tracing_fsharp_b200 import native, sample_images  # noqa: E402
from ray_tracing_fsharp_b200.distributed import DeviceBackend  # noqa: E402
from ray_tracing_fsharp_b200.domain import marshal  # noqa: E402
from ray_tracing_fsharp_b200.scene import Camera  # noqa: E402

spec = sample_images.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C2"]()
cam = Camera.make_basic(spec.spp, spec.focal_length, spec.aspect_ratio, spec.origin, spec.view_direction, spec.view_up)
cam.bounce_depth = int(sys.argv[2]) if len(sys.argv) > 2 else spec.bounce_depth  # a lower budget shortens the longest paths (the tail)
hs, ts, keep = marshal(spec.objects)
scene = native.SceneHandle(hs, ts, 0, keepalive=keep)
mw, mh = spec.max_width_coord, spec.max_height_coord
base = None
for world in (1, 2, 4, 8):
    be = DeviceBackend(scene, cam, mw, mh, seed=1, adaptive=True)
    # flags of the whole frame (all ranks' probes), so that the main phase sees what it would see after the all-reduce
    full = DeviceBackend(scene, cam, mw, mh, seed=1, adaptive=True)
    fs, ff = full.alloc()
    full.probe(0, 1, fs, ff)
    torch.cuda.synchronize()
    best = [1e9, 1e9, 1e9]
    for rep in range(4):
        stats, flags = be.alloc()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        be.probe(0, world, stats, flags)
        ev[1].record()
        flags.copy_(ff)
        ev[2].record()
        be.main(0, world, stats, flags)
        ev[3].record()
        torch.cuda.synchronize()
        t = [ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3])]
        best = [min(best[0], t[0]), min(best[1], t[1]), min(best[2], t[0] + t[1])]
    if world == 1:
        base = best
    print(f"world {world}: probe {best[0]:7.3f} ms (ideal {base[0] / world:7.3f})   main {best[1]:7.3f} ms (ideal {base[1] / world:7.3f})   "
          f"sum {best[2]:7.3f} ms  efficiency {base[2] / world / best[2]:.3f}")
