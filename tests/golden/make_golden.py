"""Regenerates the golden fixtures from the reference checkout (run in the build container only;
/root/reference does not exist on the GPU box and nothing at test time reads it).

  PpmOutputExample.txt  <- RayTracing.Test/PpmOutputExample.txt, the P3 golden file of TestPpmOutput.fs:12-46
                           (the only golden-output fixture the reference holds for this path)
  earthmap_rgb8.png     <- RayTracing.App/earthmap.jpg (the resource LoadImage.fromResource decodes, LoadImage.fs:9-18) decoded
                           with PIL / libjpeg-turbo and stored losslessly: the texels the `earth` scene (C3) is rendered
                           with, here and on the GPU box.  Decoded texels are input to the library, not reference code.
  oracle_frames.npz     <- small frames of the four scene families rendered by the CPU oracle (regression fixtures of
                           the oracle, not outputs of the F# reference)
"""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/RayTracing.Test"

def copy_reference_fixtures():
    shutil.copyfile(os.path.join(REF, "PpmOutputExample.txt"), os.path.join(HERE, "PpmOutputExample.txt"))
    print("wrote PpmOutputExample.txt")


def make_earthmap():
    import numpy as np
    from PIL import Image
    with Image.open("/root/reference/RayTracing.App/earthmap.jpg") as im:
        rgb = np.asarray(im.convert("RGB"), dtype=np.uint8)
    Image.fromarray(rgb).save(os.path.join(HERE, "earthmap_rgb8.png"), "PNG", optimize=True)
    with Image.open(os.path.join(HERE, "earthmap_rgb8.png")) as back:
        assert np.array_equal(np.asarray(back.convert("RGB")), rgb)
    print(f"wrote earthmap_rgb8.png {rgb.shape}")


def make_oracle_frames():
    """Small frames rendered by the CPU oracle with the counter RNG (the RNG the device shares), kept as regression
    fixtures: the oracle must keep reproducing them bit for bit, and the GPU path is compared with them as well.
    These are NOT outputs of the F# reference (which cannot run here and is non-deterministic by construction, F9)."""
    import sys
    import numpy as np
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    import oracle
    from ray_tracing_fsharp_b200 import sample_images
    from ray_tracing_fsharp_b200.domain import marshal
    out = {}
    for name, max_w, max_h, spp in [("C1", 30, 17, 16), ("C2", 24, 16, 16), ("C3", 24, 13, 12), ("C4", 24, 13, 16)]:
        spec = sample_images.CONFIGS[name]()
        hs, ts, _keep = marshal(spec.objects)
        cam = oracle.camera_make_basic(spp, spec.focal_length, spec.aspect_ratio, spec.origin, spec.view_direction, spec.view_up)
        cam.bounce_depth = spec.bounce_depth
        rgb, stats, counters, _ = oracle.Scene(hs, ts).render(cam, max_w, max_h, seed=2024, rng_mode=1, adaptive=True, threads=1)
        out[f"{name}_rgb"] = rgb
        out[f"{name}_stats"] = stats
        out[f"{name}_work"] = np.array([counters["paths"], counters["rays"]], np.int64)
        out[f"{name}_shape"] = np.array([max_w, max_h, spp], np.int32)
    np.savez_compressed(os.path.join(HERE, "oracle_frames.npz"), **out)
    print("wrote oracle_frames.npz")


if __name__ == "__main__":
    copy_reference_fixtures()
    make_earthmap()  # before the frames: C3 is rendered with it
    make_oracle_frames()
