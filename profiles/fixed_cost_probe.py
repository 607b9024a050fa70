import sys, os
sys.path.insert(0,'/root/repo')
from ray_tracing_fsharp_b200 import native, sample_images, abi
from ray_tracing_fsharp_b200.domain import marshal
from ray_tracing_fsharp_b200.scene import Camera
import numpy as np
def run(name, spp, flags, mw=None, mh=None):
    spec=sample_images.CONFIGS[name]()
    if mw: spec.max_width_coord, spec.max_height_coord = mw, mh
    cam=Camera.make_basic(spp,spec.focal_length,spec.aspect_ratio,spec.origin,spec.view_direction,spec.view_up); cam.bounce_depth=spec.bounce_depth
    hs,ts,keep=marshal(spec.objects); h=native.SceneHandle(hs,ts,0,keepalive=keep)
    for i in range(3):
        _,_,st=h.render(cam,spec.max_width_coord,spec.max_height_coord,seed=i,flags=flags)
    print(f"{name} spp={spp} flags={flags} {spec.cols}x{spec.rows}: kernel_ms {st.kernel_ms:.3f} rays {st.rays} paths {st.paths}", flush=True)
run('C2',12,0)
run('C2',12,abi.RT_FLAG_NO_SMEM)
run('C2',12,0,60,40)
run('C2',12,abi.RT_FLAG_NO_SMEM,60,40)
run('C1',12,0,60,40)
