"""A lane-by-lane model of render_flow_kernel's control flow (csrc/rtfs_device.cu): the ring of 64 ray slots, the READY /
WALKED / EMPTY stacks, the item slots with their in-flight counters, path regeneration, parking — everything but the
tracing itself, which is replaced by random walk lengths and random path endings.  It checks the state machine (every
path of every item is traced exactly once, every item is flushed exactly once and only when none of its paths is in
flight, the loop ends) for many work shapes, and reports the lane efficiency of the walk quanta and of the passes.

    python profiles/flow_schedule_model.py            # invariants over random work shapes + efficiency for C2- / C5-like walks
"""
import sys

import numpy as np

RING, ITEMS, LANES = 64, 3, 32


def run(rng, n_items, pool_sizes, walk_len, p_end, quantum=4, check=True):
    """pool_sizes[i]: paths of item i; walk_len(): node visits of one ray; p_end: chance that a scatter ends the path."""
    ready, walked, empty = [], [], list(range(RING))
    ring_item = [None] * RING      # (item slot index, item number) of the path in a ring slot
    inflight = [0] * ITEMS
    slot_item = [None] * ITEMS     # item number loaded in each item slot
    cursor = [0] * ITEMS
    holds = [False] * ITEMS
    flushed, started, finished = [], [0] * n_items, [0] * n_items
    next_item = [0]

    def fetch():
        i = next_item[0]
        next_item[0] += 1
        return i

    state = {"cur": 0, "finishing": False, "prefetched": None}
    first = fetch()
    if first >= n_items:
        return None
    slot_item[0], cursor[0], holds[0] = first, 0, True
    state["prefetched"] = fetch()

    def flush(sidx):
        it = slot_item[sidx]
        assert inflight[sidx] == 0, "flushed an item with paths in flight"
        assert it not in flushed, "item flushed twice"
        flushed.append(it)

    def fill(wants):  # wants: list of ring slots whose path just ended (one per lane at most)
        lanes_want = list(wants)
        free_lanes = LANES - len(lanes_want)
        for _ in range(min(free_lanes, len(empty))):
            lanes_want.append(empty.pop())
        while not state["finishing"] and lanes_want:
            dry = False
            still = []
            for s in lanes_want:
                cur = state["cur"]
                q = cursor[cur]
                cursor[cur] += 1
                if q >= pool_sizes[slot_item[cur]]:
                    dry = True
                    still.append(s)
                else:
                    started[slot_item[cur]] += 1
                    inflight[cur] += 1
                    ring_item[s] = (cur, slot_item[cur])
                    ready.append(s)
            lanes_want = still
            if dry:
                free_slot = -1
                for sidx in range(ITEMS):
                    if sidx == state["cur"]:
                        continue
                    if holds[sidx]:
                        if inflight[sidx] != 0:
                            continue
                        flush(sidx)
                        holds[sidx] = False
                    if free_slot < 0:
                        free_slot = sidx
                if free_slot < 0:
                    break
                item = state["prefetched"]
                if item < n_items:
                    slot_item[free_slot], cursor[free_slot], holds[free_slot] = item, 0, True
                    state["prefetched"] = fetch()
                    state["cur"] = free_slot
                else:
                    state["finishing"] = True
        empty.extend(lanes_want)

    def shade_pass():
        n = min(LANES, len(walked))
        mine = [walked.pop() for _ in range(n)]
        wants = []
        for s in mine:
            if rng.random() < p_end:
                sidx, it = ring_item[s]
                assert slot_item[sidx] == it, "the item slot was reloaded while one of its paths was in flight"
                inflight[sidx] -= 1
                finished[it] += 1
                wants.append(s)
            else:
                ready.append(s)
        fill(wants)
        return n

    have = [False] * LANES
    slot = [0] * LANES
    left = [0] * LANES
    visits_done = visits_slots = passes = pass_lanes = 0
    fill([])
    fill([])
    guard = 0
    while True:
        guard += 1
        assert guard < 10_000_000, "the schedule does not terminate"
        for l in range(LANES):
            if not have[l] and ready:
                slot[l] = ready.pop()
                left[l] = walk_len()
                have[l] = True
        walking = sum(have)
        if len(walked) >= 32 or (walking != LANES and not ready and walked):
            pass_lanes += shade_pass()
            passes += 1
            continue
        if walking == 0:
            break
        for _ in range(quantum):
            active = [l for l in range(LANES) if have[l] and left[l] > 0]
            if active:  # a quantum step is issued for the whole warp if any lane walks
                visits_slots += LANES
                visits_done += len(active)
            for l in active:
                left[l] -= 1
        for l in range(LANES):
            if have[l] and left[l] <= 0:
                walked.append(slot[l])
                have[l] = False
    for sidx in range(ITEMS):
        if holds[sidx]:
            flush(sidx)
    if check:
        assert not ready and not walked and len(empty) == RING, (len(ready), len(walked), len(empty))
        assert sorted(flushed) == list(range(n_items)), "not every item was flushed exactly once"
        for i in range(n_items):
            assert started[i] == pool_sizes[i] == finished[i], (i, started[i], finished[i], pool_sizes[i])
    return {"walk lane efficiency": visits_done / max(1, visits_slots), "pass lane efficiency": pass_lanes / max(1, 32 * passes), "passes": passes}


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    # invariants over many work shapes: tiny and empty-ish pools, single items, long and short walks, paths that never end early
    for trial in range(400):
        n_items = int(rng.integers(1, 12))
        pools = [int(rng.choice([1, 2, 7, 32, 33, 64, 200, 1024])) for _ in range(n_items)]
        mean = float(rng.choice([1, 3, 12, 40]))
        p_end = float(rng.choice([0.05, 0.3, 1.0]))
        run(rng, n_items, pools, lambda: int(rng.geometric(1.0 / mean)), p_end, quantum=int(rng.choice([1, 4, 8])))
    print("state machine: 400 random work shapes traced every path exactly once, flushed every item exactly once, terminated")
    for name, mean, shape in [("C2-like walks (mean 12.6 visits)", 12.6, 3.0), ("C5-like walks (mean 33.6 visits, heavy tail)", 33.6, 1.2)]:
        for quantum in (1, 2, 4, 8):
            walk = lambda: max(1, int(rng.gamma(shape, mean / shape)))  # noqa: E731
            res = run(rng, 40, [1024] * 40, walk, 0.3, quantum=quantum, check=True)
            print(f"{name}, quantum {quantum}: " + ", ".join(f"{k} {v:.3f}" if isinstance(v, float) else f"{k} {v}" for k, v in res.items()))
