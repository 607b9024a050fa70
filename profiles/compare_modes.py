"""Renders the same frame with the megakernel and with the wavefront variant through rt_render and prints the
library's own timings and work counters.  Run it plain for times; run it under
  ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,\\
sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,lts__t_bytes.sum,dram__bytes.sum
for the counters DESIGN.md compares (profiles/summarise_modes.py aggregates the CSV)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ray_tracing_fsharp_b200 import abi, native, sample_images  # noqa: E402
from ray_tracing_fsharp_b200.domain import marshal  # noqa: E402
from ray_tracing_fsharp_b200.scene import Camera  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C2")
ap.add_argument("--spp", type=int, default=0)
ap.add_argument("--modes", default="megakernel,wavefront", help="comma-separated: megakernel (the default lockstep schedule), flow, wide, wavefront")
ap.add_argument("--quantum", default="")
ap.add_argument("--repeat", type=int, default=1)
args = ap.parse_args()
spec = sample_images.CONFIGS[args.config]()
if args.spp:
    spec.spp = args.spp
cam = Camera.make_basic(spec.spp, spec.focal_length, spec.aspect_ratio, spec.origin, spec.view_direction, spec.view_up)
cam.bounce_depth = spec.bounce_depth
hs, ts, keep = marshal(spec.objects)
scene = native.SceneHandle(hs, ts, 0, keepalive=keep)
for mode in args.modes.split(","):
    m = abi.RT_MODE_WAVEFRONT if mode == "wavefront" else abi.RT_MODE_MEGAKERNEL
    flags = {"flow": abi.RT_FLAG_FLOW, "wide": abi.RT_FLAG_WIDE_BVH}.get(mode, 0)
    for _ in range(args.repeat):
        rgb, _, st = scene.render(cam, spec.max_width_coord, spec.max_height_coord, seed=5, mode=m, flags=flags)
        print(f"{mode:10s} {spec.cols}x{spec.rows} {spec.spp}spp: kernels {st.kernel_ms:9.2f} ms  total {st.total_ms:9.2f} ms  launches {st.launches:5d}  "
              f"paths {st.paths}  rays {st.rays}  => {st.rays / st.kernel_ms / 1e3:9.1f} Mrays/s", flush=True)
