// rtfs_device.cu — device half of librtfs_b200.so: scene upload, the render kernels (probe / compact /
// main / finalize), the per-primitive conformance kernels and the C-ABI entry points that launch them.
// Compiled for sm_100a only.  There is no CPU fallback: every entry point here needs a CUDA device.
#include "rtfs_core.cuh"
#include "rtfs_device.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <type_traits>
#include <vector>

namespace rtfs {

// ---------------------------------------------------------------------------------------------------
// workspace pool
// ---------------------------------------------------------------------------------------------------
static std::mutex g_pool_mutex;
static std::vector<DeviceWorkspace *> g_pool;

static void workspace_destroy(DeviceWorkspace *ws) {
    if (!ws) return;
    cudaSetDevice(ws->device);
    cudaFree(ws->d_stats);
    cudaFree(ws->d_flags);
    cudaFree(ws->d_rgb);
    cudaFree(ws->d_list);
    cudaFree(ws->d_stats_b);
    cudaFree(ws->d_counters);
    cudaFree(ws->d_blob);
    if (ws->h_blob) cudaFreeHost(ws->h_blob);
    if (ws->h_counters) cudaFreeHost(ws->h_counters);
    for (auto &e : ws->ev)
        if (e) cudaEventDestroy(e);
    if (ws->stream) cudaStreamDestroy(ws->stream);
    delete ws;
}

int workspace_acquire(int device, DeviceWorkspace **out) {
    *out = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        for (size_t i = 0; i < g_pool.size(); ++i)
            if (g_pool[i]->device == device) {
                *out = g_pool[i];
                g_pool.erase(g_pool.begin() + i);
                return RT_OK;
            }
    }
    auto *ws = new DeviceWorkspace();
    ws->device = device;
    cudaDeviceProp prop;
    bool ok = cudaGetDeviceProperties(&prop, device) == cudaSuccess;
    if (ok) {
        ws->sm_count = prop.multiProcessorCount;
        ws->smem_optin = prop.sharedMemPerBlockOptin;
        ok = cudaMalloc((void **)&ws->d_counters, CN_SLOTS * sizeof(unsigned long long)) == cudaSuccess &&
             cudaMallocHost((void **)&ws->h_counters, 2 * CN_SLOTS * sizeof(unsigned long long)) == cudaSuccess &&
             cudaStreamCreateWithFlags(&ws->stream, cudaStreamNonBlocking) == cudaSuccess;
        for (auto &e : ws->ev) ok = ok && cudaEventCreate(&e) == cudaSuccess;
    }
    if (!ok) {
        std::string why = cudaGetErrorString(cudaGetLastError());
        workspace_destroy(ws);
        return fail(RT_ERR_CUDA, "device workspace allocation failed: " + why);
    }
    *out = ws;
    return RT_OK;
}

void workspace_release(DeviceWorkspace *ws) {
    if (!ws) return;
    cudaSetDevice(ws->device);
    cudaStreamSynchronize(ws->stream);
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if (g_pool.size() < 32) {
        g_pool.push_back(ws);
        return;
    }
    workspace_destroy(ws);
}

// Image textures: cudaMallocArray / cudaCreateTextureObject / their destructors cost milliseconds to tens of
// milliseconds, so arrays are pooled by (device, width, height) like the workspaces; a reused array is overwritten.
static std::vector<PooledTexture> g_texture_pool;

static bool texture_acquire(int device, int w, int h, PooledTexture &out) {
    {
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        for (size_t i = 0; i < g_texture_pool.size(); ++i)
            if (g_texture_pool[i].device == device && g_texture_pool[i].w == w && g_texture_pool[i].h == h) {
                out = g_texture_pool[i];
                g_texture_pool.erase(g_texture_pool.begin() + i);
                return true;
            }
    }
    out = PooledTexture{device, w, h, nullptr, 0};
    cudaChannelFormatDesc desc = cudaCreateChannelDesc<uchar4>();
    if (cudaMallocArray(&out.array, &desc, w, h) != cudaSuccess) return false;
    cudaResourceDesc res;
    std::memset(&res, 0, sizeof res);
    res.resType = cudaResourceTypeArray;
    res.res.array.array = out.array;
    cudaTextureDesc td;
    std::memset(&td, 0, sizeof td);
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 0;
    if (cudaCreateTextureObject(&out.object, &res, &td, nullptr) != cudaSuccess) {
        cudaFreeArray(out.array);
        return false;
    }
    return true;
}
static void texture_release(const PooledTexture &t) {
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if (g_texture_pool.size() < 16) {
        g_texture_pool.push_back(t);
        return;
    }
    cudaSetDevice(t.device);
    cudaDestroyTextureObject(t.object);
    cudaFreeArray(t.array);
}

// ---------------------------------------------------------------------------------------------------
// device scene
// ---------------------------------------------------------------------------------------------------
void device_scene_free(RtScene *scene) {
    auto *ds = static_cast<DeviceScene *>(scene->dev);
    if (!ds) return;
    cudaSetDevice(ds->device);
    // rt_device_probe / rt_device_main / rt_comm_render launch on the caller's stream, not the workspace's: quiesce the whole
    // device before the workspace and the texture arrays go back to pools another handle may draw from at once
    cudaDeviceSynchronize();
    cudaFree(ds->ref_nodes);
    cudaFree(ds->wide_blob); // (the main blob lives in the workspace)
    if (ds->ws) workspace_release(ds->ws);
    for (const PooledTexture &t : ds->images) texture_release(t);
    delete ds;
    scene->dev = nullptr;
}

size_t device_scene_bytes(const RtScene *scene) {
    auto *ds = static_cast<const DeviceScene *>(scene->dev);
    return ds ? ds->bytes : 0;
}

int device_scene_upload(RtScene *scene) {
    int rc = require_device(scene->device);
    if (rc != RT_OK) return rc;
    auto *ds = new DeviceScene();
    scene->dev = ds;
    ds->device = scene->device;
    const HostSceneLayout &L = scene->layout;
    auto bail = [&](int code) {
        device_scene_free(scene);
        return code;
    };
    if ((rc = workspace_acquire(ds->device, &ds->ws)) != RT_OK) return bail(rc);

    // textures: image texels go into a cudaArray read through a texture object (point sampling)
    std::vector<DTexture> dt(scene->textures.size());
    for (size_t i = 0; i < scene->textures.size(); ++i) {
        const RtTexture &t = scene->textures[i];
        DTexture &o = dt[i];
        std::memset(&o, 0, sizeof o);
        o.kind = t.kind;
        o.rgb = (uint32_t(t.colour[0]) << 16) | (uint32_t(t.colour[1]) << 8) | uint32_t(t.colour[2]);
        o.w = t.width;
        o.h = t.height;
        o.even = t.even;
        o.odd = t.odd;
        o.grid = float(t.grid_size);
        o.cx = float(t.map_centre[0]);
        o.cy = float(t.map_centre[1]);
        o.cz = float(t.map_centre[2]);
        o.inv_radius = float(1.0 / (t.map_radius != 0.0 ? t.map_radius : 1.0));
        if (t.kind == RT_TEX_IMAGE) {
            std::vector<uchar4> texels(size_t(t.width) * t.height);
            for (size_t k = 0; k < texels.size(); ++k) texels[k] = make_uchar4(t.rgb8[3 * k], t.rgb8[3 * k + 1], t.rgb8[3 * k + 2], 255);
            PooledTexture pt;
            if (!texture_acquire(ds->device, t.width, t.height, pt)) return bail(fail(RT_ERR_CUDA, "texture allocation failed for an image texture"));
            ds->images.push_back(pt);
            if (cudaMemcpy2DToArray(pt.array, 0, 0, texels.data(), size_t(t.width) * sizeof(uchar4), size_t(t.width) * sizeof(uchar4), t.height,
                                    cudaMemcpyHostToDevice) != cudaSuccess)
                return bail(fail(RT_ERR_CUDA, "cudaMemcpy2DToArray failed for an image texture"));
            o.tex = pt.object;
            ds->bytes += texels.size() * sizeof(uchar4);
        }
    }

    // everything else travels as one blob: one allocation, one host->device copy
    auto align = [](size_t x) { return (x + 255) & ~size_t(255); };
    const size_t off_nodes = 0;
    const size_t off_spheres = align(off_nodes + L.nodes.size() * sizeof(DNode));
    const size_t off_mats = align(off_spheres + L.spheres.size() * sizeof(DSphere));
    const size_t off_unb = align(off_mats + L.materials.size() * sizeof(DMaterial));
    const size_t off_ref = align(off_unb + L.unbounded.size() * sizeof(DUnbounded));
    const size_t off_tex = off_ref;
    const size_t total = align(off_tex + dt.size() * sizeof(DTexture)) + 256;
    DeviceWorkspace *ws = ds->ws;
    if (ws->blob_cap < total) { // grow the pooled blob and its pinned staging copy
        cudaFree(ws->d_blob);
        if (ws->h_blob) cudaFreeHost(ws->h_blob);
        ws->d_blob = ws->h_blob = nullptr;
        ws->blob_cap = 0;
        const size_t cap = std::max<size_t>(total + total / 2, 256 << 10);
        if (cudaMalloc(&ws->d_blob, cap) != cudaSuccess || cudaMallocHost(&ws->h_blob, cap) != cudaSuccess) {
            cudaFree(ws->d_blob);
            ws->d_blob = nullptr;
            return bail(fail(RT_ERR_CUDA, "cudaMalloc failed for the scene"));
        }
        ws->blob_cap = cap;
    }
    uint8_t *host = static_cast<uint8_t *>(ws->h_blob);
    auto put = [&](size_t off, const void *src, size_t n) {
        if (n) std::memcpy(host + off, src, n);
    };
    { // the device walks centre / half-extent boxes (rtfs_internal.h)
        DNode *dn = reinterpret_cast<DNode *>(host + off_nodes);
        for (size_t i = 0; i < L.nodes.size(); ++i) dn[i] = device_node_of(L.nodes[i]);
    }
    put(off_spheres, L.spheres.data(), L.spheres.size() * sizeof(DSphere));
    put(off_mats, L.materials.data(), L.materials.size() * sizeof(DMaterial));
    put(off_unb, L.unbounded.data(), L.unbounded.size() * sizeof(DUnbounded));
    put(off_tex, dt.data(), dt.size() * sizeof(DTexture));
    ds->blob = ws->d_blob;
    if (cudaMemcpyAsync(ds->blob, host, total, cudaMemcpyHostToDevice, ws->stream) != cudaSuccess || cudaStreamSynchronize(ws->stream) != cudaSuccess)
        return bail(fail(RT_ERR_CUDA, "host->device copy of the scene failed"));
    ds->bytes += total;
    uint8_t *base = static_cast<uint8_t *>(ds->blob);
    ds->g.nodes = reinterpret_cast<const uint4 *>(base + off_nodes);
    ds->g.spheres = reinterpret_cast<const float4 *>(base + off_spheres);
    ds->g.mats = reinterpret_cast<const uint4 *>(base + off_mats);
    ds->g.unb = reinterpret_cast<const DUnbounded *>(base + off_unb);
    ds->g.tex = reinterpret_cast<const DTexture *>(base + off_tex);
    ds->g.n_nodes = int32_t(L.nodes.size());
    ds->g.n_bounded = L.n_bounded;
    ds->g.n_unbounded = int32_t(L.unbounded.size());
    ds->g.n_tex = int32_t(scene->textures.size());
    ds->max_depth = L.max_depth;
    // lean: nothing in this scene needs the FP64 evaluation of an unbounded object or a texture lookup (SceneAccess, rtfs_core.cuh)
    ds->lean = true;
    for (const DUnbounded &u : L.unbounded) ds->lean = ds->lean && u.fp32 == 1;
    for (const DMaterial &m : L.materials) ds->lean = ds->lean && m.texture < 0;
    if (!L.wide_nodes.empty()) return device_scene_ensure_wide(scene); // built at creation for big scenes
    return RT_OK;
}

// uploads the 8-wide compressed tree (building it first if the scene was too small to get one at creation)
int device_scene_ensure_wide(RtScene *scene) {
    auto *ds = static_cast<DeviceScene *>(scene->dev);
    if (!ds || ds->wide_blob || ds->g.n_bounded <= 0) return RT_OK;
    HostSceneLayout &L = scene->layout;
    build_wide_layout(L);
    if (L.wide_depth > 30) return fail(RT_ERR_UNSUPPORTED, "the wide BVH is deeper than the traversal stack");
    auto align = [](size_t x) { return (x + 255) & ~size_t(255); };
    const size_t off_nodes = 0;
    const size_t off_sph = align(L.wide_nodes.size() * sizeof(DWideNode));
    const size_t off_w2d = align(off_sph + L.wide_spheres.size() * sizeof(DSphere));
    const size_t off_d2w = align(off_w2d + L.wide_to_dev.size() * sizeof(int32_t));
    const size_t total = align(off_d2w + L.dev_to_wide.size() * sizeof(int32_t)) + 256;
    std::vector<uint8_t> host(total, 0);
    std::memcpy(host.data() + off_nodes, L.wide_nodes.data(), L.wide_nodes.size() * sizeof(DWideNode));
    std::memcpy(host.data() + off_sph, L.wide_spheres.data(), L.wide_spheres.size() * sizeof(DSphere));
    std::memcpy(host.data() + off_w2d, L.wide_to_dev.data(), L.wide_to_dev.size() * sizeof(int32_t));
    std::memcpy(host.data() + off_d2w, L.dev_to_wide.data(), L.dev_to_wide.size() * sizeof(int32_t));
    RT_CUDA(cudaSetDevice(ds->device));
    RT_CUDA(cudaMalloc(&ds->wide_blob, total));
    RT_CUDA(cudaMemcpy(ds->wide_blob, host.data(), total, cudaMemcpyHostToDevice));
    ds->bytes += total;
    uint8_t *base = static_cast<uint8_t *>(ds->wide_blob);
    ds->g.wide_nodes = reinterpret_cast<const uint4 *>(base + off_nodes);
    ds->g.wide_spheres = reinterpret_cast<const float4 *>(base + off_sph);
    ds->g.wide_to_dev = reinterpret_cast<const int32_t *>(base + off_w2d);
    ds->g.dev_to_wide = reinterpret_cast<const int32_t *>(base + off_d2w);
    ds->wide_depth = L.wide_depth;
    return RT_OK;
}

// which tree a frame walks: the binary SAH tree (staged in shared memory when the scene fits there, read from global
// memory otherwise) unless the caller asks for the 8-wide compressed one.  Measured on B200: the wide
// walk executes more instructions per ray than the binary one (eight quantised boxes decoded and tested per visit for a
// third as many visits) and these kernels are issue-bound, not bandwidth-bound, so it is not the default for any size.
bool frame_walks_wide_tree(const DeviceScene *ds, int opt_flags, bool smem_fits) {
    (void)smem_fits;
    if (ds->g.n_bounded <= 0 || (opt_flags & RT_FLAG_BVH2)) return false;
    return (opt_flags & RT_FLAG_WIDE_BVH) != 0;
}
int prepare_frame(RtScene *scene, const RtRenderOpts *opts) {
    if (opts->flags & RT_FLAG_WIDE_BVH) return device_scene_ensure_wide(scene);
    return RT_OK;
}

// uploads the reference-topology tree the first time the conformance traversal asks for it
int device_scene_ensure_reference(RtScene *scene) {
    auto *ds = static_cast<DeviceScene *>(scene->dev);
    if (!ds || ds->ref_nodes) return RT_OK;
    scene_ensure_reference(scene);
    const auto &nodes = scene->layout.ref_nodes;
    RT_CUDA(cudaSetDevice(ds->device));
    RT_CUDA(cudaMalloc((void **)&ds->ref_nodes, std::max<size_t>(nodes.size(), 1) * sizeof(DRefNode)));
    if (!nodes.empty()) RT_CUDA(cudaMemcpy(ds->ref_nodes, nodes.data(), nodes.size() * sizeof(DRefNode), cudaMemcpyHostToDevice));
    ds->n_ref_nodes = int32_t(nodes.size());
    return RT_OK;
}

// ---------------------------------------------------------------------------------------------------
// render kernels
// ---------------------------------------------------------------------------------------------------
template <int SMEM>
__device__ __forceinline__ SceneAccess<SMEM> stage_scene(const FrameParams &fp) {
    SceneAccess<SMEM> sc;
    sc.g = fp.g;
    const uint32_t window = uint32_t(__cvta_generic_to_shared(rtfs_smem));
    sc.s_nodes = window + 16u * fp.s_nodes;
    sc.s_spheres = window + 16u * fp.s_spheres;
    sc.s_mats = window + 16u * fp.s_mats;
    if (SMEM) {
        sc.stage_tree(fp.s_nodes, fp.s_spheres, fp.s_mats);
    }
    return sc;
}

// per-warp scratch in shared memory: item slots (one being issued, the others with their last paths still in flight).
// 400 bytes each: for a scene read from global memory what shared memory the slots take is L1 the walk does not get.
// A pixel's red and green sums share a word (16 bits each): a chunk is at most kMaxChunkSamples = 256 samples, 256 x 255 < 2^16.
struct ItemSlot {
    int cursor;    // next path of the item's pool
    int j_begin;   // first local sample index of the item's chunk
    int j_len;     // samples in the chunk (also what the flush adds to the pixels' count field)
    int to_b;      // probe: the chunk belongs to the second accumulator set
    int pix[32];   // row << 16 | col of each of the 32 pixels, -1: none
    unsigned acc_rg[32]; // sum of red | sum of green << 16
    unsigned acc_b[32];  // sum of blue
};
static_assert(sizeof(ItemSlot) % 16 == 0, "ItemSlot must be a multiple of 16 bytes");
constexpr size_t warp_scratch_bytes(bool staged) { return size_t(item_slots(staged)) * sizeof(ItemSlot); } // per warp

__device__ __forceinline__ void flush_counters(unsigned long long *counters, uint32_t paths, uint32_t rays, TraversalCounters cn, bool count) {
    for (int off = 16; off > 0; off >>= 1) {
        paths += __shfl_down_sync(0xffffffffu, paths, off);
        rays += __shfl_down_sync(0xffffffffu, rays, off);
        if (count) {
            cn.box_tests += __shfl_down_sync(0xffffffffu, cn.box_tests, off);
            cn.prim_tests += __shfl_down_sync(0xffffffffu, cn.prim_tests, off);
        }
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(counters + CN_PATHS, (unsigned long long)paths);
        atomicAdd(counters + CN_RAYS, (unsigned long long)rays);
        if (count) {
            atomicAdd(counters + CN_BOX, (unsigned long long)cn.box_tests);
            atomicAdd(counters + CN_PRIM, (unsigned long long)cn.prim_tests);
        }
    }
}

// The render kernel.  Persistent warps pull work items from a global cursor; an item is a unit of 32 pixels
// times a chunk of sample indices.  Inside an item the 32 lanes pull (pixel, sample) paths from a warp-local pool,
// so every lane traces until the pool is dry (path regeneration); per-sample results are added to the pixels'
// integer accumulators in shared memory (PixelStats.add, Pixel.fs:87-95) and flushed with one RED per channel.
//   PROBE = true : unit = one 8x4 tile owned by this rank; samples = the 2*firstTrial+1 probe samples of
//                  renderPixel (Scene.fs:172-182): those up to firstTrial go to `stats`, the rest to `stats_b`
//                  (no chunk straddles the two), so that probe_flags_kernel can compare the two means.
//   PROBE = false: unit = 32 consecutive entries of the flagged-pixel list; samples = this rank's share of the
//                  remaining sample indices (Scene.fs:191-192).
// A warp issues one instruction every ~7.6 cycles whatever the load, so the kernel ends one whole item after the
// work runs out: hence chunks of decreasing length, the last ones a single sample (see build_chunks).
template <bool PROBE, int SMEM, bool COUNT, bool SSTACK, bool WIDE>
__global__ void __launch_bounds__(block_threads(SMEM), blocks_per_sm(SMEM)) render_kernel(const FrameParams fp) {
    const SceneAccess<SMEM> sc = stage_scene<SMEM>(fp);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    typename std::conditional<SSTACK, SharedStack, LocalStack>::type stack;
    if constexpr (SSTACK) {
        stack.base = uint32_t(__cvta_generic_to_shared(rtfs_smem)) + 16u * fp.s_stack + 4u * threadIdx.x;
        stack.stride = 4u * blockDim.x; // words of one stack level (wide walk: two levels per entry)
        stack.levels = fp.stack_levels;
    }
    constexpr int kSlots = SMEM ? kItemSlots : kGlobalItemSlots;
    ItemSlot *const slots = reinterpret_cast<ItemSlot *>(rtfs_smem + fp.s_warp) + warp * kSlots; // this warp's item slots
    unsigned long long *work = fp.counters + (PROBE ? CN_WORK_PROBE : CN_WORK_MAIN);
    uint32_t n_paths = 0, n_rays = 0;
    TraversalCounters cn{0, 0};

    const unsigned n_list = PROBE ? 0u : (unsigned)fp.counters[CN_LIST];
    const unsigned n_units = PROBE ? unsigned((fp.tiles_x * fp.tiles_y - fp.rank + fp.world - 1) / fp.world) : (n_list + 31u) / 32u;
    const unsigned long long n_items = (unsigned long long)n_units * (unsigned long long)fp.n_chunks;

    // Loads work item `item` into a slot: chunk-major numbering; this lane's pixel; the pool.  Returns the pool size.
    auto load_item = [&](ItemSlot *sl, unsigned long long item, int &n_entries) -> int {
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
        sl->acc_rg[lane] = 0u; sl->acc_b[lane] = 0u;
        sl->pix[lane] = my_pixel < 0 ? -1 : (((my_pixel / fp.cam.cols) << 16) | (my_pixel % fp.cam.cols));
        if (lane == 0) {
            sl->cursor = 0;
            sl->j_begin = j_begin;
            sl->j_len = j_len;
            sl->to_b = (PROBE && j_begin > fp.first_trial) ? 1 : 0;
        }
        __syncwarp();
        return n_entries * j_len;
    };
    // Retires a slot whose paths have all finished: one RED per channel and pixel (PixelStats.add, Pixel.fs:87-95)
    auto flush_item = [&](ItemSlot *sl) {
        __syncwarp();
        const int rc = sl->pix[lane];
        if (rc >= 0) {
            int *st = (sl->to_b ? fp.stats_b : fp.stats) + 4 * (size_t(rc >> 16) * size_t(fp.cam.cols) + size_t(rc & 0xffff));
            const unsigned rg = sl->acc_rg[lane];
            atomicAdd(st + 0, int(rg & 0xffffu));
            atomicAdd(st + 1, int(rg >> 16));
            atomicAdd(st + 2, int(sl->acc_b[lane]));
            atomicAdd(st + 3, sl->j_len);
        }
        __syncwarp();
    };
    auto fetch_item = [&]() -> unsigned long long { // lane 0 pulls the next item number; consumed (shuffled) later
        return lane == 0 ? atomicAdd(work, 1ull) : 0ull;
    };

    unsigned long long item = __shfl_sync(0xffffffffu, fetch_item(), 0);
    if (threadIdx.x == 0 && item < n_items) atomicAdd(fp.counters + CN_BUSY_BLOCKS, 1ull); // blocks that found work: the resident ones
    if (item < n_items) {
        // kItemSlots slots per warp: lanes pull paths from slot `cur`; when its pool is dry the next item is loaded into a
        // free slot straight away, so lanes do not idle while the last paths of an item complete.  A slot is free once
        // its in-flight paths have finished and it has been flushed; with long, uneven paths (big scenes) several older
        // items can have stragglers at once, hence more than two slots.  Only the warp's final items drain.
        int cur = 0;
        unsigned holds = 1u; // slots that hold an item not yet flushed
        bool finishing = false;
        int n_entries;
        int pool = load_item(&slots[0], item, n_entries);
        int j_begin = slots[0].j_begin;
        unsigned long long prefetched = fetch_item();
        PathState ps;
        bool active = false, dry = false;
        int lane_slot = 0, my = 0; // my: slot index of this lane's path; lane_slot: its pixel within the item
        for (;;) {
            if (!active && !dry) {
                ItemSlot *sl = &slots[cur];
                int q = atomicAdd(&sl->cursor, 1);
                if (q >= pool) {
                    dry = true;
                } else {
                    int j = (n_entries == 32) ? (q >> 5) : (q / n_entries);
                    lane_slot = q - j * n_entries;
                    my = cur;
                    int rc = sl->pix[lane_slot]; // row << 16 | col, or -1 where the tile overhangs the image edge
                    if (rc >= 0) {
                        uint32_t sample = uint32_t(PROBE ? j_begin + j : fp.sample_begin + fp.rank + (j_begin + j) * fp.world);
                        ++n_paths;
                        active = path_begin(ps, fp.cam, fp.k0, fp.k1, rc >> 16, rc & 0xffff, sample); // false: Ray.make' failed (the reference throws)
                        if (!active) atomicAdd(fp.counters + CN_DEGENERATE, 1ull);
                    }
                }
            }
            // Every lane votes, every iteration: this is where the warp reconverges (lanes that regenerate a path and lanes
            // that do not would otherwise drift apart for good and the traversal would run at a fraction of the warp width).
            if (__any_sync(0xffffffffu, dry)) {
                // the current pool is exhausted: retire the older slots whose paths have all finished, then move on
                if (!finishing) {
                    int free_slot = -1;
#pragma unroll
                    for (int sidx = 0; sidx < kSlots; ++sidx) {
                        if (sidx == cur) continue;
                        if ((holds >> sidx) & 1u) {
                            if (__any_sync(0xffffffffu, active && my == sidx)) continue; // stragglers
                            flush_item(&slots[sidx]);
                            holds &= ~(1u << sidx);
                        }
                        if (free_slot < 0) free_slot = sidx;
                    }
                    if (free_slot >= 0) {
                        item = __shfl_sync(0xffffffffu, prefetched, 0);
                        if (item < n_items) {
                            ItemSlot *next = &slots[free_slot];
                            pool = load_item(next, item, n_entries);
                            j_begin = next->j_begin;
                            prefetched = fetch_item();
                            cur = free_slot;
                            holds |= 1u << free_slot;
                            dry = false;
                        } else {
                            finishing = true;
                        }
                    }
                }
                if (finishing && !__any_sync(0xffffffffu, active)) break;
            }
            const unsigned tracing = __ballot_sync(0xffffffffu, active);
            if (active) {
                uint32_t result;
                ++n_rays;
                if (path_step<SMEM, COUNT, decltype(stack), WIDE>(ps, sc, fp.cam.depth, result, cn, tracing, stack)) {
                    RTFS_BOUNDS(my >= 0 && my < kSlots && lane_slot >= 0 && lane_slot < 32);
                    ItemSlot *mine = &slots[my]; // PixelStats.add into the item's accumulators
                    atomicAdd(&mine->acc_rg[lane_slot], ((result >> 16) & 255u) | ((result << 8) & 0x00ff0000u));
                    atomicAdd(&mine->acc_b[lane_slot], result & 255u);
                    if (result & kDegenerate) atomicAdd(fp.counters + CN_DEGENERATE, 1ull);
                    active = false;
                }
            }
        }
        for (int sidx = 0; sidx < kSlots; ++sidx) // no path is in flight any more
            if ((holds >> sidx) & 1u) flush_item(&slots[sidx]);
    }
    flush_counters(fp.counters, n_paths, n_rays, cn, COUNT);
}


#include "rtfs_flow.cuh" // render_flow_kernel: the flow schedule (RT_FLAG_FLOW)

// End of the probe phase (Scene.fs:177-188) for the tiles this rank owns: merge the two accumulator sets into
// `stats`, flag the pixels whose two truncated means differ (PixelStats.mean Pixel.fs:103-108, Pixel.difference :113-116)
__global__ void probe_flags_kernel(int32_t *stats, const int32_t *stats_b, uint8_t *flags, int rows, int cols, int tiles_x, int rank, int world,
                                   int first_trial, int n_probe, int more) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= rows * cols) return;
    int r = p / cols, c = p - r * cols;
    int tile = (r / kTileH) * tiles_x + c / kTileW;
    if (tile % world != rank) return;
    int4 a = reinterpret_cast<int4 *>(stats)[p], b = reinterpret_cast<const int4 *>(stats_b)[p];
    int n_old = first_trial + 1;
    int4 s = make_int4(a.x + b.x, a.y + b.y, a.z + b.z, n_probe);
    int diff = abs(s.x / n_probe - a.x / n_old) + abs(s.y / n_probe - a.y / n_old) + abs(s.z / n_probe - a.z / n_old);
    reinterpret_cast<int4 *>(stats)[p] = s;
    flags[p] = (diff != 0 && more) ? 1 : 0;
}

// flags -> list of flagged pixel ids, one warp per 8x4 tile so that list neighbours are image neighbours
// (with several devices in one process the flag of a tile is read from its owner's buffer over NVLink)
__global__ void compact_kernel(const FlagsView flags, uint32_t *list, unsigned long long *counters, int rows, int cols, int tiles_x,
                               int tiles_y) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; tile < tiles_x * tiles_y; tile += warps) {
        int ty = tile / tiles_x, tx = tile - ty * tiles_x;
        int r = ty * kTileH + (lane >> 3), c = tx * kTileW + (lane & 7);
        const uint8_t *owner = flags.by_rank[flags.world > 1 ? tile % flags.world : 0];
        bool f = (r < rows && c < cols) && owner[r * cols + c] != 0;
        unsigned m = __ballot_sync(0xffffffffu, f);
        if (m == 0) continue;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(counters + CN_LIST, (unsigned long long)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (f) list[base + __popc(m & ((1u << lane) - 1u))] = uint32_t(r * cols + c);
    }
}

// PixelStats.mean (Pixel.fs:103-108) and, optionally, PixelOutput.correct (ImageOutput.fs:11-18)
__global__ void finalize_kernel(const int32_t *stats, int n_pixels, int gamma, uint8_t *rgb) {
    __shared__ uint8_t lut[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        int v = i;
        if (gamma) {
            v = __double2int_rn(sqrt(double(i) / 255.0) * 255.0); // Math.Round: half to even
            if (v == 256) v = 255;
        }
        lut[i] = uint8_t(v);
    }
    __syncthreads();
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pixels) return;
    int4 s = reinterpret_cast<const int4 *>(stats)[p];
    int n = s.w > 0 ? s.w : 1;
    rgb[3 * size_t(p) + 0] = lut[(s.x / n) & 255];
    rgb[3 * size_t(p) + 1] = lut[(s.y / n) & 255];
    rgb[3 * size_t(p) + 2] = lut[(s.z / n) & 255];
}

// ---------------------------------------------------------------------------------------------------
// frame driver
// ---------------------------------------------------------------------------------------------------
int check_frame_args(const RtScene *scene, const RtCamera *camera, int max_w, int max_h, const RtRenderOpts *opts) {
    if (!scene || !camera || !opts) return fail(RT_ERR_INVALID_ARGUMENT, "render: null argument");
    if (!scene->dev) return fail(RT_ERR_NO_DEVICE, "render: the scene was created without a device (device = -1); there is no CPU fallback");
    if (max_w <= 0 || max_h <= 0) return fail(RT_ERR_INVALID_ARGUMENT, "render: maxWidthCoord and maxHeightCoord must be positive");
    if (max_w > 32767 || max_h > 16383) return fail(RT_ERR_INVALID_ARGUMENT, "render: image too large (at most 65535 x 32767)");
    if (camera->samples_per_pixel < 1 || camera->samples_per_pixel > (1 << 22))
        return fail(RT_ERR_INVALID_ARGUMENT, "render: samples_per_pixel must be in [1, 2^22] (integer sums are 32-bit)");
    if (camera->bounce_depth < 0) return fail(RT_ERR_INVALID_ARGUMENT, "render: bounce_depth must be non-negative");
    return RT_OK;
}

struct LaunchPlan {
    bool smem;
    size_t smem_bytes;
    int blocks, threads;
};

typedef void (*RenderKernelFn)(const FrameParams);
template <bool PROBE, int SMEM, bool SSTACK, bool WIDE>
static RenderKernelFn pick_count(bool count) {
    return count ? render_kernel<PROBE, SMEM, true, SSTACK, WIDE> : render_kernel<PROBE, SMEM, false, SSTACK, WIDE>;
}
template <bool PROBE>
static RenderKernelFn pick_global(bool sstack, bool wide, bool count) { // the scene is read from global memory
    if (wide) return sstack ? pick_count<PROBE, 0, true, true>(count) : pick_count<PROBE, 0, false, true>(count);
    return sstack ? pick_count<PROBE, 0, true, false>(count) : pick_count<PROBE, 0, false, false>(count);
}
template <bool PROBE>
static RenderKernelFn pick_flow(bool smem, bool count) {
    if (smem) return count ? render_flow_kernel<PROBE, true, true> : render_flow_kernel<PROBE, true, false>;
    return count ? render_flow_kernel<PROBE, false, true> : render_flow_kernel<PROBE, false, false>;
}
static RenderKernelFn pick_kernel(bool probe, bool smem, bool lean, bool count, bool sstack, bool wide) {
    if (smem && lean) return probe ? pick_count<true, 2, false, false>(count) : pick_count<false, 2, false, false>(count);
    if (probe) return smem ? pick_count<true, 1, false, false>(count) : pick_global<true>(sstack, wide, count);
    return smem ? pick_count<false, 1, false, false>(count) : pick_global<false>(sstack, wide, count);
}

// does a block with `bytes` of dynamic shared memory leave room for `per_sm` of its kind on an SM (1 KB reserved per block)
static bool smem_fits(const DeviceScene *ds, size_t bytes, int per_sm) {
    return (bytes + 1024) * size_t(per_sm) <= ds->ws->smem_optin + (per_sm > 1 ? 1024 : 0);
}

// lays out shared memory and sizes the persistent grid
static int plan_launch(DeviceScene *ds, FrameParams &fp, bool probe, bool count, bool no_smem, LaunchPlan &plan, RenderKernelFn &fn) {
    const bool wide_asked = frame_walks_wide_tree(ds, fp.opt_flags, true);
    // the flow schedule (ring of rays per warp) on request only — measured slower than lockstep on B200 —
    // and not with the wide tree (no flow variant) nor when the bounce budget does not fit the byte it shares with the colour
    const bool flow = (fp.opt_flags & RT_FLAG_FLOW) && !(fp.opt_flags & RT_FLAG_LOCKSTEP) && !wide_asked && fp.cam.depth <= kMaxFlowDepth;
    size_t warp_q = (kBlockThreads / 32) * (flow ? sizeof(FlowWarp) : warp_scratch_bytes(true)) / 16;
    const size_t nodes_q = size_t(ds->g.n_nodes) * kStagedNodeQuads, sph_q = size_t(ds->g.n_bounded), mat_q = size_t(ds->g.n_bounded + ds->g.n_unbounded) * 2;
    const size_t scene_q = nodes_q + sph_q + mat_q;
    // stage the scene in shared memory when it fits beside the per-warp scratch (one block per SM)
    bool smem = smem_fits(ds, (scene_q + warp_q) * 16, kBlocksPerSm) && ds->g.n_bounded > 0;
    const bool wide = frame_walks_wide_tree(ds, fp.opt_flags, smem);
    if (no_smem || wide) smem = false;
    // the launch shape follows from where the scene is read (rtfs_device.h); the flow kernels keep one block of 1024
    plan.threads = flow ? kBlockThreads : block_threads(smem);
    const int want_per_sm = flow ? 1 : blocks_per_sm(smem);
    warp_q = size_t(plan.threads / 32) * (flow ? sizeof(FlowWarp) : warp_scratch_bytes(smem)) / 16;
    fp.s_nodes = 0;
    fp.s_spheres = uint32_t(nodes_q);
    fp.s_mats = uint32_t(nodes_q + sph_q);
    fp.s_warp = smem ? uint32_t(scene_q) : 0u;
    plan.smem = smem;
    // A scene read from L2 leaves shared memory free: the walk stacks go there when tree depth x block size words fit
    // (measured on the 100 k-sphere scene at 16 spp: 124.3 against 128.2 ms; with the scene itself in shared memory the
    // same change is within noise, 67.1 against 67.4 ms on C2, 137.7 against 136.7 on C4, so those kernels keep a local stack)
    const size_t used_q = (smem ? scene_q : 0) + warp_q;
    const size_t stack_q = size_t(wide ? 2 * (ds->wide_depth + 1) : ds->max_depth + 1) * size_t(plan.threads) * 4 / 16;
    const bool sstack = !flow && !smem && ds->g.n_bounded > 0 && smem_fits(ds, (used_q + stack_q) * 16, want_per_sm);
    fp.s_stack = uint32_t(used_q);
    {
        static const int forced = [] {
            const char *e = std::getenv("RTFS_FLOW_QUANTUM");
            return e ? std::max(1, std::min(64, std::atoi(e))) : 0;
        }();
        fp.quantum = forced ? forced : (smem ? kFlowQuantumSmall : kFlowQuantumBig);
    }
    fp.stack_levels = int32_t(stack_q * 16 / (size_t(plan.threads) * 4));
    plan.smem_bytes = (used_q + (sstack ? stack_q : 0)) * 16;
    fn = flow ? (probe ? pick_flow<true>(smem, count) : pick_flow<false>(smem, count)) : pick_kernel(probe, smem, ds->lean && !(fp.opt_flags & RT_FLAG_NO_LEAN), count, sstack, wide);
    { // The dynamic shared-memory limit of a kernel is a per-function, per-device setting: it is only ever RAISED here
      // (to the largest size any scene has asked for), and the occupancy of a (kernel, size, device) triple is asked once.
        struct Known {
            const void *fn;
            size_t smem;
            int device, per_sm;
        };
        static std::mutex mutex;
        static std::vector<Known> known; // smem = the size the entry's occupancy was computed for
        static std::vector<Known> limits; // smem = the limit currently set for (fn, device)
        std::lock_guard<std::mutex> lock(mutex);
        Known *limit = nullptr;
        for (Known &k : limits)
            if (k.fn == (const void *)fn && k.device == ds->device) limit = &k;
        if (!limit || limit->smem < plan.smem_bytes) {
            RT_CUDA(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(plan.smem_bytes)));
            // Two blocks per SM must both be resident (a persistent kernel's second wave finds no work left).  The occupancy
            // query below answers 2 whenever the SM COULD hold them, but the carve-out the driver picks by default at launch
            // is not always large enough for both (measured: with 61 KB of dynamic shared memory per block only 148 of the 296
            // blocks ever ran, 100 k spheres 279 ms instead of 202), so the largest carve-out is asked for explicitly.
            if (!smem && !flow) RT_CUDA(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            if (limit) limit->smem = plan.smem_bytes;
            else limits.push_back(Known{(const void *)fn, plan.smem_bytes, ds->device, 0});
        }
        int per_sm = -1;
        for (const Known &k : known)
            if (k.fn == (const void *)fn && k.smem == plan.smem_bytes && k.device == ds->device) per_sm = k.per_sm;
        if (per_sm < 0) {
            RT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)fn, plan.threads, plan.smem_bytes));
            known.push_back(Known{(const void *)fn, plan.smem_bytes, ds->device, per_sm});
        }
        if (per_sm < 1) return fail(RT_ERR_CUDA, "render kernel does not fit on an SM");
        plan.blocks = per_sm * ds->ws->sm_count;
    }
    return RT_OK;
}

static int ensure_list(DeviceScene *ds, size_t n_pixels) {
    if (ds->ws->list_pixels >= n_pixels) return RT_OK;
    cudaFree(ds->ws->d_list);
    ds->ws->d_list = nullptr;
    ds->ws->list_pixels = 0;
    RT_CUDA(cudaMalloc((void **)&ds->ws->d_list, n_pixels * sizeof(uint32_t)));
    ds->ws->list_pixels = n_pixels;
    return RT_OK;
}

void fill_frame(FrameParams &fp, DeviceScene *ds, const RtCamera &cam, int max_w, int max_h, const RtRenderOpts &opts, int rank, int world) {
    std::memset(&fp, 0, sizeof fp);
    fp.g = ds->g;
    fp.cam = make_dev_camera(cam, max_w, max_h);
    fp.k0 = uint32_t(opts.seed);
    fp.k1 = uint32_t(opts.seed >> 32);
    fp.rank = rank;
    fp.world = world;
    fp.tiles_x = (fp.cam.cols + kTileW - 1) / kTileW;
    fp.tiles_y = (fp.cam.rows + kTileH - 1) / kTileH;
    fp.adaptive = opts.adaptive ? 1 : 0;
    if (fp.adaptive) {
        fp.first_trial = std::min(5, cam.samples_per_pixel / 2); // Scene.fs:172
        fp.n_probe = 2 * fp.first_trial + 1;
        fp.sample_begin = fp.n_probe;
        fp.sample_end = std::max(fp.n_probe, cam.samples_per_pixel); // the third loop runs spp - 2 firstTrial - 1 times, Scene.fs:191
    } else {
        fp.first_trial = 0;
        fp.n_probe = 0;
        fp.sample_begin = 0;
        fp.sample_end = cam.samples_per_pixel;
    }
    fp.counters = ds->ws->d_counters;
    fp.opt_flags = opts.flags;
}

// Cuts this rank's local sample range [0, n_local) into chunks: `bulk` samples per chunk, then a taper
// bulk/2, bulk/4, ..., 2, 1, 1 so that the last layers of items are short (a warp cannot be sped up, so the
// kernel ends one item after the work runs out).  No chunk straddles `boundary` (probe: firstTrial + 1, the first
// sample of the second accumulator set).  Chunks are stored by decreasing length: items are numbered chunk-major.
static int build_chunks(FrameParams &fp, int n_local, int boundary, size_t n_units, int resident_warps) {
    fp.n_chunks = 0;
    if (n_local <= 0) return RT_OK;
    // bulk chunk: every warp should see ~16 bulk items; at most 32 samples (1024 paths) per item
    long long bulk = (long long)(n_units * size_t(n_local)) / (16LL * std::max(1, resident_warps));
    bulk = std::max(1LL, std::min(32LL, bulk));
    bulk = std::max(bulk, (long long)((n_local + 159) / 160)); // the table has kMaxChunks entries; n_local <= kMaxLaunchSamples keeps bulk <= kMaxChunkSamples
    std::vector<int> sizes;
    for (int t = int(bulk) / 2; t >= 1; t /= 2) sizes.push_back(t);
    if (bulk > 1) sizes.push_back(1);
    int taper = 0;
    for (int t : sizes) taper += t;
    while (!sizes.empty() && taper > n_local) { // short range: drop the longest taper chunks
        taper -= sizes.front();
        sizes.erase(sizes.begin());
    }
    int rest = n_local - taper;
    std::vector<int> all;
    while (rest > 0) {
        int c = int(std::min<long long>(bulk, rest));
        all.push_back(c);
        rest -= c;
    }
    all.insert(all.end(), sizes.begin(), sizes.end());
    std::vector<std::pair<int, int>> chunks;
    int at = 0;
    for (int c : all) {
        if (boundary > at && boundary < at + c) { // split at the boundary
            chunks.push_back({at, boundary - at});
            chunks.push_back({boundary, at + c - boundary});
        } else {
            chunks.push_back({at, c});
        }
        at += c;
    }
    std::stable_sort(chunks.begin(), chunks.end(), [](const std::pair<int, int> &a, const std::pair<int, int> &b) { return a.second > b.second; });
    if (chunks.size() > size_t(kMaxChunks)) return fail(RT_ERR_UNSUPPORTED, "render: the sample range needs more chunks than the work table holds");
    fp.n_chunks = int(chunks.size());
    for (int k = 0; k < fp.n_chunks; ++k) {
        fp.chunk_begin[k] = uint32_t(chunks[k].first);
        fp.chunk_len[k] = uint32_t(chunks[k].second);
    }
    return RT_OK;
}

int launch_probe(DeviceScene *ds, FrameParams fp, bool count, bool no_smem, cudaStream_t st, int *launches) {
    DeviceWorkspace *ws = ds->ws;
    const size_t n_pixels = size_t(fp.cam.rows) * fp.cam.cols;
    if (ws->probe_pixels < n_pixels) {
        cudaFree(ws->d_stats_b);
        ws->d_stats_b = nullptr;
        ws->probe_pixels = 0;
        RT_CUDA(cudaMalloc((void **)&ws->d_stats_b, n_pixels * 4 * sizeof(int32_t)));
        ws->probe_pixels = n_pixels;
    }
    RT_CUDA(cudaMemsetAsync(ws->d_stats_b, 0, n_pixels * 4 * sizeof(int32_t), st));
    RT_CUDA(cudaMemsetAsync(ws->d_counters + CN_WORK_PROBE, 0, sizeof(unsigned long long), st));
    fp.stats_b = ws->d_stats_b;
    LaunchPlan plan;
    RenderKernelFn fn;
    int rc = plan_launch(ds, fp, true, count, no_smem, plan, fn);
    if (rc != RT_OK) return rc;
    const size_t n_units = size_t((fp.tiles_x * fp.tiles_y - fp.rank + fp.world - 1) / fp.world);
    rc = build_chunks(fp, fp.n_probe, fp.first_trial + 1, n_units, plan.blocks * (plan.threads / 32));
    if (rc != RT_OK) return rc;
    fn<<<plan.blocks, plan.threads, plan.smem_bytes, st>>>(fp);
    RT_CUDA(cudaGetLastError());
    ++*launches;
    probe_flags_kernel<<<unsigned((n_pixels + 255) / 256), 256, 0, st>>>(fp.stats, ws->d_stats_b, fp.flags, fp.cam.rows, fp.cam.cols, fp.tiles_x,
                                                                          fp.rank, fp.world, fp.first_trial, fp.n_probe,
                                                                          fp.sample_end > fp.sample_begin ? 1 : 0);
    RT_CUDA(cudaGetLastError());
    ++*launches;
    return RT_OK;
}

int launch_main(DeviceScene *ds, FrameParams fp, const FlagsView &flags, bool count, bool no_smem, cudaStream_t st, int *launches) {
    const size_t n_pixels = size_t(fp.cam.rows) * fp.cam.cols;
    int rc = ensure_list(ds, n_pixels);
    if (rc != RT_OK) return rc;
    RT_CUDA(cudaMemsetAsync(ds->ws->d_counters + CN_LIST, 0, sizeof(unsigned long long), st));
    RT_CUDA(cudaMemsetAsync(ds->ws->d_counters + CN_WORK_MAIN, 0, sizeof(unsigned long long), st));
    compact_kernel<<<ds->ws->sm_count * 4, 256, 0, st>>>(flags, ds->ws->d_list, ds->ws->d_counters, fp.cam.rows, fp.cam.cols, fp.tiles_x, fp.tiles_y);
    RT_CUDA(cudaGetLastError());
    ++*launches;
    fp.list = ds->ws->d_list;
    LaunchPlan plan;
    RenderKernelFn fn;
    rc = plan_launch(ds, fp, false, count, no_smem, plan, fn);
    if (rc != RT_OK) return rc;
    int n_span = fp.sample_end - fp.sample_begin - fp.rank;
    int n_local = n_span > 0 ? (n_span + fp.world - 1) / fp.world : 0;
    // the list length is only known on the device; size the chunks for the whole frame (an upper bound on the units)
    // one launch covers at most kMaxLaunchSamples of this rank's sample indices (a chunk stays within kMaxChunkSamples: the
    // 16-bit halves of an item's red | green word); a frame with more is several launches
    for (int base = 0; base < n_local; base += kMaxLaunchSamples) {
        rc = build_chunks(fp, std::min(kMaxLaunchSamples, n_local - base), -1, (n_pixels + 31) / 32, plan.blocks * (plan.threads / 32));
        if (rc != RT_OK) return rc;
        for (int k = 0; k < fp.n_chunks; ++k) fp.chunk_begin[k] += uint32_t(base);
        if (base > 0) RT_CUDA(cudaMemsetAsync(ds->ws->d_counters + CN_WORK_MAIN, 0, sizeof(unsigned long long), st));
        fn<<<plan.blocks, plan.threads, plan.smem_bytes, st>>>(fp);
        RT_CUDA(cudaGetLastError());
        ++*launches;
    }
    return RT_OK;
}

void read_counters(DeviceScene *ds, RtStats *stats, size_t n_pixels, bool adaptive) {
    stats->paths = ds->ws->h_counters[CN_PATHS];
    stats->rays = ds->ws->h_counters[CN_RAYS];
    stats->box_tests = ds->ws->h_counters[CN_BOX];
    stats->prim_tests = ds->ws->h_counters[CN_PRIM];
    stats->pixels_early_out = adaptive ? (unsigned long long)n_pixels - ds->ws->h_counters[CN_LIST] : 0;
    stats->degenerate_paths = ds->ws->h_counters[CN_DEGENERATE];
    if (std::getenv("RTFS_DEBUG_BLOCKS")) std::fprintf(stderr, "rtfs: blocks that found work (probe + main): %llu\n", ds->ws->h_counters[CN_BUSY_BLOCKS]);
}

// implemented in rtfs_wavefront.cu
int render_wavefront(RtScene *scene, const RtCamera *camera, int32_t max_w, int32_t max_h, const RtRenderOpts *opts, uint8_t *rgb_out,
                     int32_t *sums_out, RtStats *stats);

} // namespace rtfs

using namespace rtfs;

// =================================================================================================
// C ABI — device entry points
// =================================================================================================
extern "C" {

int rt_host_pin(void *ptr, size_t bytes) {
    if (!ptr || !bytes) return fail(RT_ERR_INVALID_ARGUMENT, "rt_host_pin: null or empty buffer");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(RT_ERR_NO_DEVICE, "rt_host_pin: no CUDA device is visible");
    }
    RT_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return RT_OK;
}
int rt_host_unpin(void *ptr) {
    if (!ptr) return fail(RT_ERR_INVALID_ARGUMENT, "rt_host_unpin: null buffer");
    RT_CUDA(cudaHostUnregister(ptr));
    return RT_OK;
}

int rt_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// bytes of the scene each persistent block stages in shared memory (0: the BVH is read from global memory / L2)
size_t rt_scene_shared_memory_bytes(const RtScene *scene) {
    if (!scene || !scene->dev) return 0;
    auto *ds = static_cast<const DeviceScene *>(scene->dev);
    const size_t warp_q = (kBlockThreads / 32) * warp_scratch_bytes(true) / 16; // the default (lockstep) schedule's per-warp scratch
    const size_t scene_q = size_t(ds->g.n_nodes) * kStagedNodeQuads + size_t(ds->g.n_bounded) + size_t(ds->g.n_bounded + ds->g.n_unbounded) * 2;
    bool smem = smem_fits(ds, (scene_q + warp_q) * 16, kBlocksPerSm) && ds->g.n_bounded > 0;
    return smem ? scene_q * 16 : 0;
}

int rt_device_probe(RtScene *scene, const RtCamera *camera, int32_t max_w, int32_t max_h, const RtRenderOpts *opts, int32_t rank,
                    int32_t world, int32_t *d_stats, uint8_t *d_flags, void *stream, RtStats *stats) {
    int rc = check_frame_args(scene, camera, max_w, max_h, opts);
    if (rc != RT_OK) return rc;
    if (world < 1 || rank < 0 || rank >= world || !d_stats || !d_flags) return fail(RT_ERR_INVALID_ARGUMENT, "rt_device_probe: bad rank/world/buffers");
    if ((rc = prepare_frame(scene, opts)) != RT_OK) return rc;
    auto *ds = static_cast<DeviceScene *>(scene->dev);
    RT_CUDA(cudaSetDevice(ds->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FrameParams fp;
    fill_frame(fp, ds, *camera, max_w, max_h, *opts, rank, world);
    fp.stats = d_stats;
    fp.flags = d_flags;
    const size_t n_pixels = size_t(fp.cam.rows) * fp.cam.cols;
    int launches = 0;
    RT_CUDA(cudaMemsetAsync(ds->ws->d_counters, 0, CN_SLOTS * sizeof(unsigned long long), st));
    if (!fp.adaptive) {
        // no probe phase: every pixel takes all its samples in the main phase (rank 0 raises the flags once)
        if (rank == 0) RT_CUDA(cudaMemsetAsync(d_flags, 1, n_pixels, st));
    } else {
        rc = launch_probe(ds, fp, (opts->flags & RT_FLAG_COUNTERS) != 0, (opts->flags & RT_FLAG_NO_SMEM) != 0, st, &launches);
        if (rc != RT_OK) return rc;
    }
    if (stats) {
        RT_CUDA(cudaMemcpyAsync(ds->ws->h_counters, ds->ws->d_counters, CN_SLOTS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        RT_CUDA(cudaStreamSynchronize(st));
        std::memset(stats, 0, sizeof *stats);
        read_counters(ds, stats, n_pixels, false);
        stats->launches = launches;
    }
    return RT_OK;
}

int rt_device_main(RtScene *scene, const RtCamera *camera, int32_t max_w, int32_t max_h, const RtRenderOpts *opts, int32_t rank,
                   int32_t world, int32_t *d_stats, const uint8_t *d_flags, void *stream, RtStats *stats) {
    int rc = check_frame_args(scene, camera, max_w, max_h, opts);
    if (rc != RT_OK) return rc;
    if (world < 1 || rank < 0 || rank >= world || !d_stats || !d_flags) return fail(RT_ERR_INVALID_ARGUMENT, "rt_device_main: bad rank/world/buffers");
    if ((rc = prepare_frame(scene, opts)) != RT_OK) return rc;
    auto *ds = static_cast<DeviceScene *>(scene->dev);
    RT_CUDA(cudaSetDevice(ds->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FrameParams fp;
    fill_frame(fp, ds, *camera, max_w, max_h, *opts, rank, world);
    fp.stats = d_stats;
    fp.flags = const_cast<uint8_t *>(d_flags);
    const size_t n_pixels = size_t(fp.cam.rows) * fp.cam.cols;
    int launches = 0;
    rc = launch_main(ds, fp, single_flags(d_flags), (opts->flags & RT_FLAG_COUNTERS) != 0, (opts->flags & RT_FLAG_NO_SMEM) != 0, st, &launches);
    if (rc != RT_OK) return rc;
    if (stats) {
        RT_CUDA(cudaMemcpyAsync(ds->ws->h_counters, ds->ws->d_counters, CN_SLOTS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        RT_CUDA(cudaStreamSynchronize(st));
        std::memset(stats, 0, sizeof *stats);
        read_counters(ds, stats, n_pixels, fp.adaptive != 0);
        stats->launches = launches;
    }
    return RT_OK;
}

// paths / rays / tests accumulated on this scene since the last rt_device_probe (one D2H copy + a stream sync)
int rt_device_counters(RtScene *scene, void *stream, RtStats *stats) {
    if (!scene || !stats) return fail(RT_ERR_INVALID_ARGUMENT, "rt_device_counters: null argument");
    if (!scene->dev) return fail(RT_ERR_NO_DEVICE, "rt_device_counters: the scene has no device");
    auto *ds = static_cast<DeviceScene *>(scene->dev);
    RT_CUDA(cudaSetDevice(ds->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    RT_CUDA(cudaMemcpyAsync(ds->ws->h_counters, ds->ws->d_counters, CN_SLOTS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    RT_CUDA(cudaStreamSynchronize(st));
    std::memset(stats, 0, sizeof *stats);
    stats->paths = ds->ws->h_counters[CN_PATHS];
    stats->rays = ds->ws->h_counters[CN_RAYS];
    stats->box_tests = ds->ws->h_counters[CN_BOX];
    stats->prim_tests = ds->ws->h_counters[CN_PRIM];
    stats->pixels_early_out = ds->ws->h_counters[CN_LIST]; // NOTE: here the number of pixels that went on to phase 2
    return RT_OK;
}

int rt_device_finalize(int32_t device, const int32_t *d_stats, int32_t n_pixels, int32_t gamma, uint8_t *d_rgb, void *stream) {
    int rc = require_device(device);
    if (rc != RT_OK) return rc;
    if (!d_stats || !d_rgb || n_pixels <= 0) return fail(RT_ERR_INVALID_ARGUMENT, "rt_device_finalize: bad argument");
    finalize_kernel<<<(n_pixels + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_stats, n_pixels, gamma, d_rgb);
    RT_CUDA(cudaGetLastError());
    return RT_OK;
}

int rt_render(RtScene *scene, const RtCamera *camera, int32_t max_w, int32_t max_h, const RtRenderOpts *opts, uint8_t *rgb_out,
              int32_t *sums_out, RtStats *stats) {
    int rc = check_frame_args(scene, camera, max_w, max_h, opts);
    if (rc != RT_OK) return rc;
    if (!rgb_out) return fail(RT_ERR_INVALID_ARGUMENT, "rt_render: rgb_out is null");
    if (opts->mode == RT_MODE_WAVEFRONT) return render_wavefront(scene, camera, max_w, max_h, opts, rgb_out, sums_out, stats);
    if (opts->mode != RT_MODE_MEGAKERNEL) return fail(RT_ERR_INVALID_ARGUMENT, "rt_render: unknown mode");
    if ((rc = prepare_frame(scene, opts)) != RT_OK) return rc;
    auto *ds = static_cast<DeviceScene *>(scene->dev);
    RT_CUDA(cudaSetDevice(ds->device));
    const size_t n_pixels = size_t(2 * max_w + 1) * size_t(2 * max_h + 1);
    if (ds->ws->ws_pixels < n_pixels) {
        cudaFree(ds->ws->d_stats);
        cudaFree(ds->ws->d_flags);
        cudaFree(ds->ws->d_rgb);
        ds->ws->d_stats = nullptr;
        ds->ws->d_flags = nullptr;
        ds->ws->d_rgb = nullptr;
        ds->ws->ws_pixels = 0;
        RT_CUDA(cudaMalloc((void **)&ds->ws->d_stats, n_pixels * 4 * sizeof(int32_t)));
        RT_CUDA(cudaMalloc((void **)&ds->ws->d_flags, n_pixels));
        RT_CUDA(cudaMalloc((void **)&ds->ws->d_rgb, n_pixels * 3));
        ds->ws->ws_pixels = n_pixels;
    }
    cudaStream_t st = ds->ws->stream;
    FrameParams fp;
    fill_frame(fp, ds, *camera, max_w, max_h, *opts, 0, 1);
    fp.stats = ds->ws->d_stats;
    fp.flags = ds->ws->d_flags;
    const bool count = (opts->flags & RT_FLAG_COUNTERS) != 0, no_smem = (opts->flags & RT_FLAG_NO_SMEM) != 0;
    int launches = 0;
    RT_CUDA(cudaEventRecord(ds->ws->ev[0], st));
    RT_CUDA(cudaMemsetAsync(ds->ws->d_counters, 0, CN_SLOTS * sizeof(unsigned long long), st));
    RT_CUDA(cudaMemsetAsync(ds->ws->d_stats, 0, n_pixels * 4 * sizeof(int32_t), st));
    RT_CUDA(cudaEventRecord(ds->ws->ev[1], st));
    if (fp.adaptive) {
        rc = launch_probe(ds, fp, count, no_smem, st, &launches);
        if (rc != RT_OK) return rc;
    } else {
        RT_CUDA(cudaMemsetAsync(ds->ws->d_flags, 1, n_pixels, st));
    }
    RT_CUDA(cudaMemcpyAsync(ds->ws->h_counters + CN_SLOTS, ds->ws->d_counters, CN_SLOTS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    RT_CUDA(cudaEventRecord(ds->ws->ev[4], st));
    rc = launch_main(ds, fp, single_flags(ds->ws->d_flags), count, no_smem, st, &launches);
    if (rc != RT_OK) return rc;
    RT_CUDA(cudaEventRecord(ds->ws->ev[5], st));
    RT_CUDA(cudaEventRecord(ds->ws->ev[2], st));
    finalize_kernel<<<unsigned((n_pixels + 255) / 256), 256, 0, st>>>(ds->ws->d_stats, int(n_pixels), opts->gamma, ds->ws->d_rgb);
    RT_CUDA(cudaGetLastError());
    ++launches;
    RT_CUDA(cudaMemcpyAsync(rgb_out, ds->ws->d_rgb, n_pixels * 3, cudaMemcpyDeviceToHost, st));
    if (sums_out) RT_CUDA(cudaMemcpyAsync(sums_out, ds->ws->d_stats, n_pixels * 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    RT_CUDA(cudaMemcpyAsync(ds->ws->h_counters, ds->ws->d_counters, CN_SLOTS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    RT_CUDA(cudaEventRecord(ds->ws->ev[3], st));
    RT_CUDA(cudaStreamSynchronize(st));
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        read_counters(ds, stats, n_pixels, fp.adaptive != 0);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ds->ws->ev[1], ds->ws->ev[2]);
        stats->kernel_ms = ms;
        cudaEventElapsedTime(&ms, ds->ws->ev[0], ds->ws->ev[3]);
        stats->total_ms = ms;
        cudaEventElapsedTime(&ms, ds->ws->ev[4], ds->ws->ev[5]);
        stats->main_ms = ms;
        stats->main_rays = ds->ws->h_counters[CN_RAYS] - ds->ws->h_counters[CN_SLOTS + CN_RAYS];
        stats->launches = launches;
    }
    return RT_OK;
}

} // extern "C"
