// rtfs_flow.cuh — the flow schedule of the render kernel (RT_FLAG_FLOW): included by rtfs_device.cu, inside namespace rtfs,
// after the lockstep kernel whose frame parameters, scene staging, tile geometry and counters it shares.
// Measured slower than lockstep on B200 (92.5 against 67.0 ms on the RTOW frame); kept selectable
// because it is bit-identical and the next architecture may price its bookkeeping differently.
#pragma once

// ---------------------------------------------------------------------------------------------------------------
// The flow kernel: the same frame, with the walk decoupled from the scatter inside every warp.
//
// render_kernel (above) moves its 32 lanes in lockstep: one ray per lane per pass, so a pass lasts as long as the
// longest of 32 walks — measured 16-18 of 32 lanes active in the walk on the RTOW scene, 9 on the 100 k-sphere scene
// (profiles/warp_schedule_sim.py reproduces both from the walk lengths alone: 0.58 and 0.34).  Here every warp keeps a
// RING of 64 rays in shared memory.  A ray is READY (waiting for its walk), WALKING (one per lane, its walk state in that
// lane's registers and stack), WALKED (closest sphere known, waiting for the rest of hitObject and for its scatter) or
// the slot is EMPTY.  The warp alternates:
//   * walk quanta: every lane that has a ray performs up to `quantum` node visits; lanes whose walk ended hand the
//     result to the ring and take the next READY ray at the next schedule point, so a long walk delays nobody;
//   * a shading pass as soon as 32 rays are WALKED: unbounded objects, strike point, scatter, bounce count
//     (Scene.fs:77-114) for 32 rays at full width; a path that ends is accumulated into its item (PixelStats.add) and its
//     slot is given the next camera sample of the warp's work items (path regeneration), so the pass hands 32 READY
//     rays back.  Walks in progress are parked in shared memory over the pass (five words per lane).
// With 64 slots the ring never runs dry in steady state: 32 walking + READY + WALKED = 64, so READY reaches 0 exactly
// when WALKED reaches 32.  Same RNG keys, same integer sums: bit-identical frames to render_kernel's (tested).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kRing = 64;          // ray slots per warp
constexpr int kFlowItems = 3;      // work items a warp can have in flight
// node visits between two schedule points (FrameParams.quantum): a lane whose walk ends waits for the next schedule point
// to take another ray, and a schedule point costs ~50 issue slots when some lane does; 4 suits walks of ~13 visits (the RTOW
// scene), 8 walks of ~34 (the 100 k-sphere scene).  RTFS_FLOW_QUANTUM overrides both (tuning runs).
constexpr int kFlowQuantumSmall = 4, kFlowQuantumBig = 8;
constexpr int kMaxFlowDepth = 253; // the bounce count shares a word with the colour
enum RingField { RF_OX = 0, RF_OY, RF_OZ, RF_DX, RF_DY, RF_DZ, RF_COLOUR, RF_LAST, RF_WHERE, RF_T, RF_REF, RF_FIELDS };
struct FlowItem {
    int cursor, j_begin, j_len, to_b;
    int pix[32]; // row << 16 | col of each of the 32 pixels, -1: none
    int acc[3][32];
};
struct FlowWarp {
    uint32_t ring[RF_FIELDS][kRing];
    uint8_t ready[kRing], walked[kRing], empty[kRing];
    int park_node[32], park_sp[32], park_best[32], park_slot[32];
    float park_t[32];
    int inflight[4];
    int pad[12];
    FlowItem item[kFlowItems];
};
static_assert(sizeof(FlowWarp) % 16 == 0, "FlowWarp must be a multiple of 16 bytes");

template <bool PROBE, bool SMEM, bool COUNT>
__global__ void __launch_bounds__(kBlockThreads, 1) render_flow_kernel(const FrameParams fp) {
    const SceneAccess<SMEM> sc = stage_scene<SMEM>(fp);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned full = 0xffffffffu, lt = (1u << lane) - 1u;
    LocalStack stack;
    FlowWarp *fw = reinterpret_cast<FlowWarp *>(rtfs_smem + fp.s_warp) + warp;
    unsigned long long *work = fp.counters + (PROBE ? CN_WORK_PROBE : CN_WORK_MAIN);
    uint32_t n_paths = 0, n_rays = 0;
    TraversalCounters cn{0, 0};

    const unsigned n_list = PROBE ? 0u : (unsigned)fp.counters[CN_LIST];
    const unsigned n_units = PROBE ? unsigned((fp.tiles_x * fp.tiles_y - fp.rank + fp.world - 1) / fp.world) : (n_list + 31u) / 32u;
    const unsigned long long n_items = (unsigned long long)n_units * (unsigned long long)fp.n_chunks;

    auto load_item = [&](FlowItem *it, unsigned long long item, int &n_entries) -> int {
        const int k = int(item / n_units);
        const unsigned unit = unsigned(item - (unsigned long long)k * n_units);
        const int j_begin = int(fp.chunk_begin[k]), j_len = int(fp.chunk_len[k]);
        int my_pixel = -1;
        n_entries = 32;
        if constexpr (PROBE) {
            int tile = int(unit) * fp.world + fp.rank;
            int ty = tile / fp.tiles_x, tx = tile - ty * fp.tiles_x;
            int r = ty * kTileH + (lane >> 3), c = tx * kTileW + (lane & 7);
            if (r < fp.cam.rows && c < fp.cam.cols) my_pixel = r * fp.cam.cols + c;
        } else {
            unsigned e = unit * 32u + lane;
            if (e < n_list) my_pixel = int(fp.list[e]);
            n_entries = int(min(32u, n_list - unit * 32u));
        }
        it->acc[0][lane] = 0; it->acc[1][lane] = 0; it->acc[2][lane] = 0;
        it->pix[lane] = my_pixel < 0 ? -1 : (((my_pixel / fp.cam.cols) << 16) | (my_pixel % fp.cam.cols));
        if (lane == 0) {
            it->cursor = 0;
            it->j_begin = j_begin;
            it->j_len = j_len;
            it->to_b = (PROBE && j_begin > fp.first_trial) ? 1 : 0;
        }
        __syncwarp();
        return n_entries * j_len;
    };
    auto flush_item = [&](FlowItem *it) { // PixelStats.add of a whole item: one RED per channel and pixel
        __syncwarp();
        int rc = it->pix[lane];
        if (rc >= 0) {
            int *st = (it->to_b ? fp.stats_b : fp.stats) + 4 * (size_t(rc >> 16) * fp.cam.cols + size_t(rc & 0xffff));
            atomicAdd(st + 0, it->acc[0][lane]);
            atomicAdd(st + 1, it->acc[1][lane]);
            atomicAdd(st + 2, it->acc[2][lane]);
            atomicAdd(st + 3, it->j_len);
        }
        __syncwarp();
    };
    auto fetch_item = [&]() -> unsigned long long { return lane == 0 ? atomicAdd(work, 1ull) : 0ull; };

    unsigned long long item = __shfl_sync(full, fetch_item(), 0);
    if (item < n_items) {
        // ---- warp-uniform state ----
        int cur = 0, n_entries, rt = 0, wt = 0, et = kRing; // current item slot; tops of the READY / WALKED / EMPTY stacks
        unsigned holds = 1u;
        bool finishing = false;
        int pool = load_item(&fw->item[0], item, n_entries);
        int j_begin = fw->item[0].j_begin;
        unsigned long long prefetched = fetch_item();
        if (lane < 4) fw->inflight[lane] = 0;
        fw->empty[lane] = uint8_t(lane);
        fw->empty[lane + 32] = uint8_t(lane + 32);
        __syncwarp();

        // Gives camera samples to EMPTY slots (and to the slots of the paths that just ended: `want`, slot `s`) while the
        // warp's work items last; the slots that get one become READY, the others go (back) on the EMPTY stack.
        auto fill = [&](bool want, int s) {
            { // lanes with nothing to refill adopt an EMPTY slot
                const unsigned freel = __ballot_sync(full, !want);
                const int rank = __popc(freel & lt);
                if (!want && rank < et) {
                    s = fw->empty[et - 1 - rank];
                    want = true;
                }
                et = max(0, et - __popc(freel));
            }
            while (!finishing && __any_sync(full, want)) {
                bool dry = false, got = false;
                if (want) {
                    FlowItem *it = &fw->item[cur];
                    int q = atomicAdd(&it->cursor, 1);
                    if (q >= pool) {
                        dry = true;
                    } else {
                        int j = (n_entries == 32) ? (q >> 5) : (q / n_entries);
                        int ls = q - j * n_entries;
                        int rc = it->pix[ls]; // -1 where the tile overhangs the image edge
                        if (rc >= 0) {
                            uint32_t sample = uint32_t(PROBE ? j_begin + j : fp.sample_begin + fp.rank + (j_begin + j) * fp.world);
                            ++n_paths;
                            PathState ps;
                            if (path_begin(ps, fp.cam, fp.k0, fp.k1, rc >> 16, rc & 0xffff, sample)) {
                                fw->ring[RF_OX][s] = __float_as_uint(ps.o.x); fw->ring[RF_OY][s] = __float_as_uint(ps.o.y); fw->ring[RF_OZ][s] = __float_as_uint(ps.o.z);
                                fw->ring[RF_DX][s] = __float_as_uint(ps.d.x); fw->ring[RF_DY][s] = __float_as_uint(ps.d.y); fw->ring[RF_DZ][s] = __float_as_uint(ps.d.z);
                                fw->ring[RF_COLOUR][s] = kWhite;
                                fw->ring[RF_LAST][s] = uint32_t(kNoRef);
                                fw->ring[RF_WHERE][s] = (uint32_t(cur) << 27) | (uint32_t(ls) << 22) | sample;
                                atomicAdd(&fw->inflight[cur], 1);
                                got = true;
                                want = false;
                            } else {
                                atomicAdd(fp.counters + CN_DEGENERATE, 1ull); // Ray.make' failed: the reference throws
                            }
                        }
                    }
                }
                { // the slots that got a path are READY
                    const unsigned m = __ballot_sync(full, got);
                    if (got) fw->ready[rt + __popc(m & lt)] = uint8_t(s);
                    rt += __popc(m);
                }
                if (__any_sync(full, dry)) { // the current item is used up: retire finished items, open the next one
                    __syncwarp();
                    int free_slot = -1;
#pragma unroll
                    for (int sidx = 0; sidx < kFlowItems; ++sidx) {
                        if (sidx == cur) continue;
                        if ((holds >> sidx) & 1u) {
                            if (fw->inflight[sidx] != 0) continue; // paths of that item are still in the ring
                            flush_item(&fw->item[sidx]);
                            holds &= ~(1u << sidx);
                        }
                        if (free_slot < 0) free_slot = sidx;
                    }
                    if (free_slot < 0) break; // every other item still has paths in flight: try again at the next pass
                    item = __shfl_sync(full, prefetched, 0);
                    if (item < n_items) {
                        pool = load_item(&fw->item[free_slot], item, n_entries);
                        j_begin = fw->item[free_slot].j_begin;
                        prefetched = fetch_item();
                        cur = free_slot;
                        holds |= 1u << free_slot;
                    } else {
                        finishing = true;
                    }
                }
            }
            { // whoever still wants a path keeps its slot EMPTY
                const unsigned m = __ballot_sync(full, want);
                if (want) fw->empty[et + __popc(m & lt)] = uint8_t(s);
                et += __popc(m);
            }
            __syncwarp();
        };

        // One shading pass: the rest of hitObject and the scatter of up to 32 WALKED rays, at full width.
        auto shade_pass = [&]() {
            const int n = min(32, wt);
            const bool mine = lane < n;
            int s = mine ? int(fw->walked[wt - 1 - lane]) : 0;
            wt -= n;
            bool want = false, cont = false;
            if (mine) {
                PathState ps;
                ps.o = f3(__uint_as_float(fw->ring[RF_OX][s]), __uint_as_float(fw->ring[RF_OY][s]), __uint_as_float(fw->ring[RF_OZ][s]));
                ps.d = f3(__uint_as_float(fw->ring[RF_DX][s]), __uint_as_float(fw->ring[RF_DY][s]), __uint_as_float(fw->ring[RF_DZ][s]));
                const uint32_t cb = fw->ring[RF_COLOUR][s], where = fw->ring[RF_WHERE][s];
                ps.colour = cb & 0x00FFFFFFu;
                ps.bounces = int(cb >> 24);
                ps.last = int(fw->ring[RF_LAST][s]);
                const int my = int(where >> 27), ls = int((where >> 22) & 31u);
                RTFS_BOUNDS(my < kFlowItems);
                const int rc = fw->item[my].pix[ls];
                ps.rng.k0 = fp.k0;
                ps.rng.k1 = fp.k1;
                ps.rng.pixel = uint32_t((rc >> 16) * fp.cam.cols + (rc & 0xffff));
                ps.rng.sample = where & 0x003FFFFFu;
                const Hit h = finish_hit<SMEM, COUNT>(sc, ps.o, ps.d, ps.last, __uint_as_float(fw->ring[RF_T][s]), int(fw->ring[RF_REF][s]), cn);
                uint32_t result;
                if (path_after_hit<SMEM>(ps, sc, h, fp.cam.depth, result)) {
                    int *acc = &fw->item[my].acc[0][0]; // PixelStats.add into the item's accumulators
                    atomicAdd(acc + ls, int((result >> 16) & 255u));
                    atomicAdd(acc + 32 + ls, int((result >> 8) & 255u));
                    atomicAdd(acc + 64 + ls, int(result & 255u));
                    if (result & kDegenerate) atomicAdd(fp.counters + CN_DEGENERATE, 1ull);
                    atomicSub(&fw->inflight[my], 1);
                    want = true;
                } else {
                    fw->ring[RF_OX][s] = __float_as_uint(ps.o.x); fw->ring[RF_OY][s] = __float_as_uint(ps.o.y); fw->ring[RF_OZ][s] = __float_as_uint(ps.o.z);
                    fw->ring[RF_DX][s] = __float_as_uint(ps.d.x); fw->ring[RF_DY][s] = __float_as_uint(ps.d.y); fw->ring[RF_DZ][s] = __float_as_uint(ps.d.z);
                    fw->ring[RF_COLOUR][s] = ps.colour | (uint32_t(ps.bounces) << 24);
                    fw->ring[RF_LAST][s] = uint32_t(ps.last);
                    cont = true;
                }
            }
            __syncwarp();
            { // the paths that go on are READY again
                const unsigned m = __ballot_sync(full, cont);
                if (cont) fw->ready[rt + __popc(m & lt)] = uint8_t(s);
                rt += __popc(m);
            }
            fill(want, s);
        };

        // ---- per-lane walk state ----
        bool have = false, done = false;
        int slot = 0, node = 0, sp = 0, best = kNoRef, last = kNoRef;
        float best_t = kNoHitT;
        float3 o = f3(0.f, 0.f, 0.f), d = f3(0.f, 0.f, 1.f);
        RaySlabs rs = make_slabs(o, d);

        fill(false, 0);
        fill(false, 0);
        for (;;) {
            // ---- schedule point: idle lanes take READY rays ----
            const unsigned idle = __ballot_sync(full, !have);
            if (idle != 0u && rt > 0) {
                const int rank = __popc(idle & lt);
                if (!have && rank < rt) {
                    slot = fw->ready[rt - 1 - rank];
                    o = f3(__uint_as_float(fw->ring[RF_OX][slot]), __uint_as_float(fw->ring[RF_OY][slot]), __uint_as_float(fw->ring[RF_OZ][slot]));
                    d = f3(__uint_as_float(fw->ring[RF_DX][slot]), __uint_as_float(fw->ring[RF_DY][slot]), __uint_as_float(fw->ring[RF_DZ][slot]));
                    last = int(fw->ring[RF_LAST][slot]);
                    rs = make_slabs(o, d);
                    node = sc.root();
                    sp = 0;
                    best_t = kNoHitT;
                    best = kNoRef;
                    have = true;
                    done = sc.g.n_bounded <= 0;
                    ++n_rays;
                }
                rt = max(0, rt - __popc(idle));
            }
            const unsigned walking = __ballot_sync(full, have);
            if (wt >= 32 || (walking != full && rt == 0 && wt > 0)) {
                // ---- shading pass; the walks in progress wait in shared memory ----
                // (every lane parks and reloads, walking or not: nothing of the walk is then live across the pass, so
                // the compiler has nothing to spill around it; an idle lane's values are never used)
                fw->park_node[lane] = node;
                fw->park_sp[lane] = sp;
                fw->park_best[lane] = best;
                fw->park_t[lane] = best_t;
                fw->park_slot[lane] = have ? slot : -1;
                shade_pass();
                slot = fw->park_slot[lane];
                have = slot >= 0;
                slot = max(slot, 0);
                node = fw->park_node[lane];
                sp = fw->park_sp[lane];
                best = fw->park_best[lane];
                best_t = fw->park_t[lane];
                o = f3(__uint_as_float(fw->ring[RF_OX][slot]), __uint_as_float(fw->ring[RF_OY][slot]), __uint_as_float(fw->ring[RF_OZ][slot]));
                d = f3(__uint_as_float(fw->ring[RF_DX][slot]), __uint_as_float(fw->ring[RF_DY][slot]), __uint_as_float(fw->ring[RF_DZ][slot]));
                last = int(fw->ring[RF_LAST][slot]);
                rs = make_slabs(o, d);
                done = false;
                continue;
            }
            if (walking == 0u) break; // nothing walking, nothing READY, nothing WALKED: the warp's work is done
            // ---- a quantum of the walk ----
#pragma unroll 1
            for (int v = 0; v < fp.quantum; ++v) {
                if (have && !done) done = bvh_visit<SMEM, COUNT>(sc, rs, o, d, last, node, sp, stack, best_t, best, cn);
            }
            // ---- finished walks go to the ring ----
            const bool fin = have && done;
            const unsigned m = __ballot_sync(full, fin);
            if (fin) {
                fw->ring[RF_T][slot] = __float_as_uint(best_t);
                fw->ring[RF_REF][slot] = uint32_t(best);
                fw->walked[wt + __popc(m & lt)] = uint8_t(slot);
                have = false;
                done = false;
            }
            wt += __popc(m);
            __syncwarp();
        }
        for (int sidx = 0; sidx < kFlowItems; ++sidx) // no path is in flight any more
            if ((holds >> sidx) & 1u) flush_item(&fw->item[sidx]);
    }
    flush_counters(fp.counters, n_paths, n_rays, cn, COUNT);
}
