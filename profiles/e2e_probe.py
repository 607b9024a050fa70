import time, sys
sys.path.insert(0,'/root/repo')
import numpy as np
from ray_tracing_fsharp_b200 import sample_images, native, abi
from ray_tracing_fsharp_b200.domain import marshal
from ray_tracing_fsharp_b200.scene import Camera
import sys
spec=sample_images.CONFIGS[sys.argv[1] if len(sys.argv)>1 else 'C2']()
cam=Camera.make_basic(spec.spp,spec.focal_length,spec.aspect_ratio,spec.origin,spec.view_direction,spec.view_up); cam.bounce_depth=spec.bounce_depth
native.lib(); native.device_count()
rgb=np.empty((spec.rows,spec.cols,3),np.uint8)
for i in range(5):
    t0=time.perf_counter(); hs,ts,keep=marshal(spec.objects); t1=time.perf_counter()
    h=native.SceneHandle(hs,ts,0,keepalive=keep); t2=time.perf_counter()
    _,_,st=h.render(cam,spec.max_width_coord,spec.max_height_coord,seed=i,rgb_out=rgb); t3=time.perf_counter()
    _,_,st2=h.render(cam,spec.max_width_coord,spec.max_height_coord,seed=i+10,rgb_out=rgb); t4=time.perf_counter()
    h.close(); t5=time.perf_counter()
    print(f"marshal {1e3*(t1-t0):.1f}  create {1e3*(t2-t1):.1f}  render1 {1e3*(t3-t2):.1f} (kernel {st.kernel_ms:.1f} total {st.total_ms:.1f})  render2 {1e3*(t4-t3):.1f} (kernel {st2.kernel_ms:.1f} total {st2.total_ms:.1f})  close {1e3*(t5-t4):.1f}")
