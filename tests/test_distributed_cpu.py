"""The N > 1 path on CPU: the same orchestration the GPU ranks run (distributed.render_split_frame: probe ->
all_reduce(flags, MAX) -> main -> all_reduce(stats, SUM)) driven over gloo with world_size 2 and 3, with the
oracle's restatement of the sample split as the backend.  The reduced frame must equal the unsplit
Scene.render of the oracle bit for bit: that is the property that makes 1/2/4/8-GPU images identical."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from helpers import oracle_camera, small_random_spheres
from ray_tracing_fsharp_b200 import sample_images
from ray_tracing_fsharp_b200.distributed import render_split_frame
from ray_tracing_fsharp_b200.domain import marshal


class OracleBackend:
    """CPU stand-in for DeviceBackend: same interface, oracle/orc_render_split underneath (test infrastructure)."""

    def __init__(self, scene, cam, max_w, max_h, seed, adaptive):
        self.scene, self.cam, self.max_w, self.max_h, self.seed, self.adaptive = scene, cam, max_w, max_h, seed, adaptive
        self.n_pixels = (2 * max_w + 1) * (2 * max_h + 1)

    def alloc(self):
        return torch.zeros((self.n_pixels, 4), dtype=torch.int32), torch.zeros((self.n_pixels,), dtype=torch.uint8)

    def probe(self, rank, world, stats, flags):
        oracle.render_split(self.scene, self.cam, self.max_w, self.max_h, self.seed, self.adaptive, 1, rank, world, stats.numpy(), flags.numpy())

    def main(self, rank, world, stats, flags):
        oracle.render_split(self.scene, self.cam, self.max_w, self.max_h, self.seed, self.adaptive, 2, rank, world, stats.numpy(), flags.numpy())


def _spec(which):
    if which == "reduced":
        spec = small_random_spheres()
        spec.max_width_coord, spec.max_height_coord, spec.spp = 21, 13, 29
    else:
        spec = sample_images.few_spheres(max_w=19, max_h=11, spp=16)
    return spec


def _worker(rank, world, port, which, adaptive, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    spec = _spec(which)
    hs, ts, _keep = marshal(spec.objects)
    scene = oracle.Scene(hs, ts)
    cam = oracle_camera(spec)
    backend = OracleBackend(scene, cam, spec.max_width_coord, spec.max_height_coord, 17, adaptive)
    stats, flags = render_split_frame(backend, rank, world, lambda t: dist.all_reduce(t, op=dist.ReduceOp.MAX),
                                      lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM))
    np.save(os.path.join(out_dir, f"stats_{rank}.npy"), stats.numpy())
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,which,adaptive", [(2, "reduced", True), (3, "C1", True), (2, "C1", False)])
def test_split_frame_over_gloo_equals_unsplit_render(tmp_path, world, which, adaptive):
    mp.spawn(_worker, args=(world, _free_port(), which, adaptive, str(tmp_path)), nprocs=world, join=True)
    spec = _spec(which)
    hs, ts, _keep = marshal(spec.objects)
    cam = oracle_camera(spec)
    _, want, _, _ = oracle.Scene(hs, ts).render(cam, spec.max_width_coord, spec.max_height_coord, seed=17, rng_mode=1, adaptive=adaptive)
    for r in range(world):
        got = np.load(tmp_path / f"stats_{r}.npy").reshape(want.shape)
        assert np.array_equal(got, want), f"rank {r}"  # every rank holds the full reduced frame


def test_single_rank_needs_no_collective():
    spec = _spec("C1")
    hs, ts, _keep = marshal(spec.objects)
    cam = oracle_camera(spec)
    scene = oracle.Scene(hs, ts)
    backend = OracleBackend(scene, cam, spec.max_width_coord, spec.max_height_coord, 3, True)
    stats, flags = render_split_frame(backend, 0, 1)  # no all_reduce callables: must not be called
    _, want, _, _ = scene.render(cam, spec.max_width_coord, spec.max_height_coord, seed=3, rng_mode=1, adaptive=True)
    assert np.array_equal(stats.numpy().reshape(want.shape), want)
    assert set(np.unique(flags.numpy()).tolist()) <= {0, 1}
