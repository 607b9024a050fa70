"""Worker of tests/test_gpu_comm.py: one rank of a torchrun job that renders one frame through rt_comm_render (the
collectives inside librtfs_b200.so) and, on rank 0, writes the frame and the sums to the .npz given on the command line."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from helpers import small_random_spheres  # noqa: E402
from ray_tracing_fsharp_b200 import native, sample_images  # noqa: E402
from ray_tracing_fsharp_b200.distributed import comm_from_torch_distributed  # noqa: E402
from ray_tracing_fsharp_b200.domain import marshal  # noqa: E402
from ray_tracing_fsharp_b200.scene import Camera  # noqa: E402


def main():
    out_path, which = sys.argv[1], sys.argv[2]
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank()
    comm = comm_from_torch_distributed(local)  # its own stream
    if which == "reduced":
        spec = small_random_spheres()
    else:
        spec = sample_images.CONFIGS[which]()
        spec.max_width_coord, spec.max_height_coord, spec.spp = 75, 50, 64
    cam = Camera.make_basic(spec.spp, spec.focal_length, spec.aspect_ratio, spec.origin, spec.view_direction, spec.view_up)
    cam.bounce_depth = spec.bounce_depth
    hs, ts, keep = marshal(spec.objects)
    scene = native.SceneHandle(hs, ts, local, keepalive=keep)
    mw, mh = spec.max_width_coord, spec.max_height_coord
    results = {}
    for adaptive in (True, False):
        rgb_a, sums, st = comm.render(scene, cam, mw, mh, seed=31, adaptive=adaptive, want_rgb=True, want_sums=True)
        rgb_b, _, _ = comm.render(scene, cam, mw, mh, seed=31, adaptive=adaptive, want_rgb=True, want_sums=False)
        rgb_g, _, _ = comm.render(scene, cam, mw, mh, seed=31, adaptive=adaptive, gamma=True, want_rgb=True)
        rays = torch.tensor([st.rays, st.paths], dtype=torch.int64, device="cuda")
        dist.all_reduce(rays)
        k = "a" if adaptive else "f"
        results.update({f"rgb_{k}": rgb_a.copy(), f"rgb_rs_{k}": rgb_b.copy(), f"rgb_gamma_{k}": rgb_g.copy(), f"sums_{k}": sums.copy(),
                        f"work_{k}": rays.cpu().numpy()})
    # every rank holds the complete frame: compare rank r's copy with rank 0's through a broadcast
    mine = torch.from_numpy(results["rgb_rs_a"].copy()).cuda()
    ref = mine.clone()
    dist.broadcast(ref, src=0)
    same = torch.tensor([int(torch.equal(mine, ref))], device="cuda")
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    results["all_ranks_hold_the_frame"] = np.array([int(same.item())])
    results["peer_memory"] = np.array([int(comm.uses_peer_memory())])
    if rank == 0:
        np.savez(out_path, **results)
    scene.close()
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
