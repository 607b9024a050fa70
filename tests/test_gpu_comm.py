"""rt_comm_render: the one-process-per-GPU frame whose collectives (NCCL) the library issues itself (csrc/rtfs_comm.cu).
Bar: bit-identical to rt_render for every world size — integer sums keyed by sample index."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from helpers import scene_pair, small_random_spheres
from ray_tracing_fsharp_b200 import native, sample_images

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _spec(which):
    if which == "reduced":
        return small_random_spheres()
    spec = sample_images.CONFIGS[which]()
    spec.max_width_coord, spec.max_height_coord, spec.spp = 75, 50, 64
    return spec


@pytest.mark.parametrize("which", ["reduced", "C3"])
def test_single_rank_communicator_equals_rt_render(which):
    """world = 1: no NCCL is loaded, the frame takes the same kernels as rt_render — on the caller's stream or its own."""
    import torch
    spec = _spec(which)
    _osc, dsc, cam = scene_pair(spec)
    mw, mh = spec.max_width_coord, spec.max_height_coord
    for stream in (None, torch.cuda.current_stream().cuda_stream or None):
        comm = native.CommHandle(None, 0, 1, 0, stream)
        for adaptive in (True, False):
            want_rgb, want_sums, want_st = dsc.render(cam, mw, mh, seed=31, adaptive=adaptive, want_sums=True)
            want_rgb, want_sums = want_rgb.copy(), want_sums.copy()
            rgb, sums, st = comm.render(dsc, cam, mw, mh, seed=31, adaptive=adaptive, want_rgb=True, want_sums=True)
            assert np.array_equal(rgb, want_rgb) and np.array_equal(sums, want_sums)
            assert int(st.rays) == int(want_st.rays) and int(st.paths) == int(want_st.paths)
            assert 0 < st.main_rays <= st.rays and st.main_ms > 0 and st.degenerate_paths == 0
            # enqueue-only use (device-resident): nothing synchronises until the caller does; the figures come afterwards
            comm.render(dsc, cam, mw, mh, seed=31, adaptive=adaptive, want_rgb=False, want_sums=False, want_stats=False)
            torch.cuda.synchronize()
            st2 = comm.last_stats(dsc)
            assert int(st2.rays) == int(want_st.rays) and st2.total_ms > 0
            d_rgb, _ = comm.frame_pointers()
            assert d_rgb
        comm.close()


@pytest.mark.parametrize("world,tail", [(2, "peer"), (2, "nccl"), (4, "peer"), (8, "peer")])
def test_ranks_over_nccl_equal_rt_render(world, tail):
    """One process per GPU under torchrun; the library all-reduces the flags and finishes the frame either with ONE kernel
    over NVLink peer memory (the ranks' buffers mapped with CUDA IPC) or, with RTFS_COMM_NO_PEER=1, with ncclReduceScatter +
    finalize + ncclAllGather.  Both must reproduce rt_render bit for bit."""
    if native.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ)
    if tail == "nccl":
        env["RTFS_COMM_NO_PEER"] = "1"
    else:
        env.pop("RTFS_COMM_NO_PEER", None)
    for which in ("reduced", "C3"):
        spec = _spec(which)
        _osc, dsc, cam = scene_pair(spec)
        mw, mh = spec.max_width_coord, spec.max_height_coord
        with tempfile.TemporaryDirectory() as td:
            out = os.path.join(td, "frame.npz")
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                   "--master-port", str(29500 + world), os.path.join(HERE, "comm_worker.py"), out, which]
            res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
            assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
            got = np.load(out)
            for k, adaptive in (("a", True), ("f", False)):
                rgb, sums, st = dsc.render(cam, mw, mh, seed=31, adaptive=adaptive, want_sums=True)
                assert np.array_equal(got[f"rgb_{k}"], rgb) and np.array_equal(got[f"sums_{k}"], sums)
                assert np.array_equal(got[f"rgb_rs_{k}"], rgb)  # reduce-scatter + all-gather path
                rgb_g, _, _ = dsc.render(cam, mw, mh, seed=31, adaptive=adaptive, gamma=True)
                assert np.array_equal(got[f"rgb_gamma_{k}"], rgb_g)
                assert int(got[f"work_{k}"][0]) == int(st.rays) and int(got[f"work_{k}"][1]) == int(st.paths)
            assert int(got["all_ranks_hold_the_frame"][0]) == 1
            if tail == "nccl":
                assert int(got["peer_memory"][0]) == 0
            else:
                print(f"world {world}: peer memory {'mapped' if int(got['peer_memory'][0]) else 'REFUSED (NCCL tail used)'}")
