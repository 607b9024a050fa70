"""What would batching the LEAF visits of a warp buy the lockstep walk?

Today every step of a warp's walk issues the node-visit code for the lanes at an internal node AND the leaf code (sphere
test) for the lanes at a leaf, the second at 1-4 lanes.  Batched: a lane that reaches a leaf waits until `threshold` lanes
of the warp have one (or no lane has a node left), then the leaf code is issued once for all of them.  This replays the
visit sequences of real rays (host build of the traversal code, csrc/host_debug) through both rules with the measured
instruction costs of the two code paths and prints the issued warp-instructions per ray."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import oracle_camera  # noqa: E402
from ray_tracing_fsharp_b200 import abi, sample_images  # noqa: E402
from ray_tracing_fsharp_b200.domain import marshal  # noqa: E402

lib = C.CDLL(os.path.join(ROOT, "ray_tracing_fsharp_b200", "csrc", "build", "librtfs_host_debug.so"))
COST_NODE, COST_LEAF, COST_VOTE = 52, 34, 3  # issued instructions of the two code paths of bvh_visit (SASS), and of a ballot + branch


def paths_of(spec, mw, mh, spp):
    spec.max_width_coord, spec.max_height_coord, spec.spp = mw, mh, spp
    hs, ts, keep = marshal(spec.objects)
    t = (abi.RtTexture * max(1, len(ts)))(*ts)
    cam = oracle_camera(spec)
    cap = 200_000_000
    buf = C.create_string_buffer(cap)
    n = C.c_uint64()
    assert lib.dbg_visit_trace(hs, len(hs), t, len(ts), C.byref(cam), mw, mh, C.c_uint64(5), spp, buf, C.c_uint64(cap), C.byref(n)) == 0
    text = buf.raw[:min(cap, n.value)].decode()
    return [[ray for ray in p.split(".") if True][:-1] for p in text.split("/") if p]


def simulate(paths, threshold):
    """32 lanes, each running its paths back to back, one ray per pass (path regeneration); returns issued instructions."""
    lanes = [[] for _ in range(32)]
    for k, p in enumerate(paths):
        lanes[k % 32].extend(p)
    n_pass = max(len(x) for x in lanes)
    issued = rays = 0
    for i in range(n_pass):
        seqs = [x[i] for x in lanes if i < len(x)]
        rays += len(seqs)
        pos = [0] * len(seqs)
        while True:
            at_node = [k for k, s in enumerate(seqs) if pos[k] < len(s) and s[pos[k]] == "I"]
            at_leaf = [k for k, s in enumerate(seqs) if pos[k] < len(s) and s[pos[k]] == "L"]
            if not at_node and not at_leaf:
                break
            if threshold == 0:  # today: both code paths whenever some lane needs them
                issued += (COST_NODE if at_node else 0) + (COST_LEAF if at_leaf else 0)
                for k in at_node + at_leaf:
                    pos[k] += 1
            else:
                issued += COST_VOTE
                if at_leaf and (len(at_leaf) >= threshold or not at_node):
                    issued += COST_LEAF
                    for k in at_leaf:
                        pos[k] += 1
                else:
                    issued += COST_NODE
                    for k in at_node:
                        pos[k] += 1
    return issued / rays


if __name__ == "__main__":
    for name, spec, mw, mh, spp in [("C2", sample_images.CONFIGS["C2"](), 40, 27, 6), ("C5 (100 k spheres)", sample_images.CONFIGS["C5"](), 32, 18, 3)]:
        paths = paths_of(spec, mw, mh, spp)
        base = simulate(paths, 0)
        print(name, f"rays {sum(len(p) for p in paths)}: today {base:.1f} walk instructions issued per ray;",
              ", ".join(f"threshold {t}: {simulate(paths, t):.1f} ({100 * (simulate(paths, t) / base - 1):+.1f} %)" for t in (2, 4, 6, 8, 12)), flush=True)
