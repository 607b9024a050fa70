"""Level-2 parity numbers for DESIGN.md: GPU (Philox) against the CPU oracle (the reference's xorshift128), independent
streams, 4096 spp, adaptive early-out on both sides as the reference renders.  Prints, per scene: per-channel MAD of
the pre-gamma means, sigma-bar (RMS standard error of a difference of two independent estimates, from two GPU seeds),
the share of early-out pixels on both sides, and the histogram of |difference| of the gamma-corrected P3 bytes."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np  # noqa: E402

import oracle  # noqa: E402  (this script is measurement tooling, like bench.py's CPU leg)
from helpers import scene_pair  # noqa: E402
from ray_tracing_fsharp_b200 import sample_images  # noqa: E402
from ray_tracing_fsharp_b200.scene import ImageOutput  # noqa: E402

# the BASELINE configs at reduced half-extents, then the reference's other sample scenes (SampleImages.fs) likewise
CASES = [(name, sample_images.CONFIGS[name](), mw, mh) for name, mw, mh in [("C1", 200, 112), ("C2", 150, 100), ("C3", 120, 67), ("C4", 120, 67),
                                                                            ("C5", 60, 34)]]
for name, scale in [("shiny-floor", 0.1), ("fuzzy-floor", 0.1), ("spheres", 0.25), ("inside-sphere", 0.04), ("total-refraction", 0.15), ("glass", 0.25),
                    ("textured-sphere", 0.25), ("moved-camera", 0.15)]:
    sp = sample_images.REFERENCE_SAMPLES[name](scale)
    CASES.append((name, sp, sp.max_width_coord, sp.max_height_coord))
if len(sys.argv) > 1:
    CASES = [c for c in CASES if c[0] in sys.argv[1:]]

for name, spec, mw, mh in CASES:
    spec.spp = 4096
    osc, dsc, cam = scene_pair(spec)
    t0 = time.perf_counter()
    ref, ref_stats, counters, _ = osc.render(cam, mw, mh, seed=1234, rng_mode=0, adaptive=True)
    t_cpu = time.perf_counter() - t0
    a, sa, st = dsc.render(cam, mw, mh, seed=1, adaptive=True, want_sums=True)
    b, sb, _ = dsc.render(cam, mw, mh, seed=2, adaptive=True, want_sums=True)
    mean_a, mean_b, mean_o = sa[..., :3] / sa[..., 3:4], sb[..., :3] / sb[..., 3:4], ref_stats[..., :3] / ref_stats[..., 3:4]
    sigma_bar = np.sqrt(((mean_a - mean_b) ** 2).mean(axis=(0, 1)))
    mad = np.abs(mean_a - mean_o).mean(axis=(0, 1))
    mad_gpu = np.abs(mean_a - mean_b).mean(axis=(0, 1))
    ppm_g = np.array(ImageOutput.write_ppm(True, a).split()[4:], dtype=np.int32)
    ppm_o = np.array(oracle.ppm_format(ref, True).split()[4:], dtype=np.int32)
    d = np.abs(ppm_g - ppm_o)
    hist = [float(np.mean(d == k)) for k in range(4)] + [float(np.mean(d >= 4))]
    print(f"{name} {2 * mw + 1}x{2 * mh + 1} @4096 spp: MAD(gpu, oracle) = {np.round(mad, 3)}  MAD(gpu, gpu') = {np.round(mad_gpu, 3)}  "
          f"sigma-bar = {np.round(sigma_bar, 3)}  =>  MAD / sigma-bar = {np.round(mad / sigma_bar, 2)}")
    print(f"      early-out pixels: oracle {np.mean(ref_stats[..., 3] == 11):.4f}  gpu {np.mean(sa[..., 3] == 11):.4f};  "
          f"P3 byte |diff| 0/1/2/3/>=4: {[round(h, 4) for h in hist]}  mean {d.mean():.3f};  oracle {t_cpu:.1f} s, gpu {st.total_ms:.1f} ms", flush=True)
