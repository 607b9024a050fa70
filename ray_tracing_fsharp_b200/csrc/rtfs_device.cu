// rtfs_device.cu — device half of librtfs_b200.so: scene upload, the render kernels (probe / compact /
// main / finalize), the per-primitive conformance kernels and the C-ABI entry points that launch them.
// Compiled for sm_100a only.  There is no CPU fallback: every entry point here needs a CUDA device.
#include "rtfs_core.cuh"
#include "rtfs_device.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
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
    cudaFree(ws->d_counters);
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
             cudaMallocHost((void **)&ws->h_counters, CN_SLOTS * sizeof(unsigned long long)) == cudaSuccess &&
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

// ---------------------------------------------------------------------------------------------------
// device scene
// ---------------------------------------------------------------------------------------------------
void device_scene_free(RtScene *scene) {
    auto *ds = static_cast<DeviceScene *>(scene->dev);
    if (!ds) return;
    cudaSetDevice(ds->device);
    if (ds->ws) workspace_release(ds->ws); // synchronises the stream first: nothing is still reading the scene
    for (auto t : ds->texobjs) cudaDestroyTextureObject(t);
    for (auto a : ds->arrays) cudaFreeArray(a);
    cudaFree(ds->blob);
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
            cudaChannelFormatDesc desc = cudaCreateChannelDesc<uchar4>();
            cudaArray_t arr = nullptr;
            if (cudaMallocArray(&arr, &desc, t.width, t.height) != cudaSuccess)
                return bail(fail(RT_ERR_CUDA, "cudaMallocArray failed for an image texture"));
            ds->arrays.push_back(arr);
            if (cudaMemcpy2DToArray(arr, 0, 0, texels.data(), size_t(t.width) * sizeof(uchar4), size_t(t.width) * sizeof(uchar4), t.height,
                                    cudaMemcpyHostToDevice) != cudaSuccess)
                return bail(fail(RT_ERR_CUDA, "cudaMemcpy2DToArray failed for an image texture"));
            cudaResourceDesc res;
            std::memset(&res, 0, sizeof res);
            res.resType = cudaResourceTypeArray;
            res.res.array.array = arr;
            cudaTextureDesc td;
            std::memset(&td, 0, sizeof td);
            td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
            td.filterMode = cudaFilterModePoint;
            td.readMode = cudaReadModeElementType;
            td.normalizedCoords = 0;
            cudaTextureObject_t obj = 0;
            if (cudaCreateTextureObject(&obj, &res, &td, nullptr) != cudaSuccess)
                return bail(fail(RT_ERR_CUDA, "cudaCreateTextureObject failed"));
            ds->texobjs.push_back(obj);
            o.tex = obj;
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
    const size_t off_tex = align(off_ref + L.ref_nodes.size() * sizeof(DRefNode));
    const size_t total = align(off_tex + dt.size() * sizeof(DTexture)) + 256;
    std::vector<uint8_t> host(total, 0);
    auto put = [&](size_t off, const void *src, size_t n) {
        if (n) std::memcpy(host.data() + off, src, n);
    };
    put(off_nodes, L.nodes.data(), L.nodes.size() * sizeof(DNode));
    put(off_spheres, L.spheres.data(), L.spheres.size() * sizeof(DSphere));
    put(off_mats, L.materials.data(), L.materials.size() * sizeof(DMaterial));
    put(off_unb, L.unbounded.data(), L.unbounded.size() * sizeof(DUnbounded));
    put(off_ref, L.ref_nodes.data(), L.ref_nodes.size() * sizeof(DRefNode));
    put(off_tex, dt.data(), dt.size() * sizeof(DTexture));
    if (cudaMalloc(&ds->blob, total) != cudaSuccess) return bail(fail(RT_ERR_CUDA, "cudaMalloc failed for the scene"));
    if (cudaMemcpy(ds->blob, host.data(), total, cudaMemcpyHostToDevice) != cudaSuccess)
        return bail(fail(RT_ERR_CUDA, "host->device copy of the scene failed"));
    ds->bytes += total;
    uint8_t *base = static_cast<uint8_t *>(ds->blob);
    ds->g.nodes = reinterpret_cast<const uint4 *>(base + off_nodes);
    ds->g.spheres = reinterpret_cast<const float4 *>(base + off_spheres);
    ds->g.mats = reinterpret_cast<const uint4 *>(base + off_mats);
    ds->g.unb = reinterpret_cast<const DUnbounded *>(base + off_unb);
    ds->ref_nodes = reinterpret_cast<DRefNode *>(base + off_ref);
    ds->g.tex = reinterpret_cast<const DTexture *>(base + off_tex);
    ds->n_ref_nodes = int32_t(L.ref_nodes.size());
    ds->g.n_nodes = int32_t(L.nodes.size());
    ds->g.n_bounded = L.n_bounded;
    ds->g.n_unbounded = int32_t(L.unbounded.size());
    ds->g.n_tex = int32_t(scene->textures.size());
    return RT_OK;
}

// ---------------------------------------------------------------------------------------------------
// render kernels
// ---------------------------------------------------------------------------------------------------
template <bool SMEM>
__device__ __forceinline__ SceneAccess<SMEM> stage_scene(const FrameParams &fp) {
    SceneAccess<SMEM> sc;
    sc.g = fp.g;
    const uint32_t window = uint32_t(__cvta_generic_to_shared(rtfs_smem));
    sc.s_nodes = window + 16u * fp.s_nodes;
    sc.s_spheres = window + 16u * fp.s_spheres;
    sc.s_mats = window + 16u * fp.s_mats;
    if (SMEM) {
        const int n_nodes_q = fp.g.n_nodes * 4, n_sph_q = fp.g.n_bounded, n_mat_q = (fp.g.n_bounded + fp.g.n_unbounded) * 2;
        for (int i = threadIdx.x; i < n_nodes_q; i += blockDim.x) rtfs_smem[fp.s_nodes + i] = __ldg(fp.g.nodes + i);
        for (int i = threadIdx.x; i < n_sph_q; i += blockDim.x) rtfs_smem[fp.s_spheres + i] = __ldg(reinterpret_cast<const uint4 *>(fp.g.spheres) + i);
        for (int i = threadIdx.x; i < n_mat_q; i += blockDim.x) rtfs_smem[fp.s_mats + i] = __ldg(fp.g.mats + i);
        __syncthreads();
    }
    return sc;
}

// per-warp scratch in shared memory: a pool cursor and two sets of 32x3 integer accumulators
struct WarpScratch {
    int cursor;
    int pad[3];
    int pix[32];
    int acc[2][3][32];
};
static_assert(sizeof(WarpScratch) % 16 == 0, "WarpScratch must be a multiple of 16 bytes");

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

// The render kernel.  Persistent warps pull work items from a global cursor; an item is 32 pixels times a
// run of sample indices.  Inside an item the 32 lanes pull (pixel, sample) paths from a warp-local pool, so
// every lane traces until the pool is dry (path regeneration) and per-sample results are added to the
// pixel's integer accumulators in shared memory (PixelStats.add, Pixel.fs:87-95).
//   PROBE = true : item = one 8x4 tile owned by this rank; paths = the 2*firstTrial+1 probe samples of
//                  renderPixel (Scene.fs:172-182); writes the sums, and flags pixels whose two truncated
//                  means differ (Scene.fs:183-188).
//   PROBE = false: item = 32 consecutive entries of the flagged-pixel list x one chunk of this rank's
//                  share of the remaining sample indices (Scene.fs:191-192); sums are added atomically.
template <bool PROBE, bool SMEM, bool COUNT>
__global__ void __launch_bounds__(kBlockThreads, 1) render_kernel(const FrameParams fp) {
    const SceneAccess<SMEM> sc = stage_scene<SMEM>(fp);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpScratch *ws = reinterpret_cast<WarpScratch *>(rtfs_smem + fp.s_warp) + warp;
    unsigned long long *work = fp.counters + (PROBE ? CN_WORK_PROBE : CN_WORK_MAIN);
    uint32_t n_paths = 0, n_rays = 0;
    TraversalCounters cn{0, 0};

    // main-phase geometry of the sample split: this rank owns sample_begin + rank + j * world
    const int n_span = fp.sample_end - fp.sample_begin - fp.rank;
    const int n_local = PROBE ? fp.n_probe : (n_span > 0 ? (n_span + fp.world - 1) / fp.world : 0);
    const int n_chunks = PROBE ? 1 : (n_local + fp.chunk - 1) / fp.chunk;
    const unsigned n_list = PROBE ? 0u : (unsigned)fp.counters[CN_LIST];
    const unsigned long long n_items =
        PROBE ? (unsigned long long)((fp.tiles_x * fp.tiles_y - fp.rank + fp.world - 1) / fp.world)
              : (unsigned long long)((n_list + 31u) / 32u) * (unsigned long long)n_chunks;

    for (;;) {
        unsigned long long item = 0;
        if (lane == 0) item = atomicAdd(work, 1ull);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_items) break;

        // ---- decode the item: this lane's pixel, and the pool ----
        int my_pixel = -1; // row_idx * cols + col_idx of the pixel this lane finalises
        int n_entries, j_begin, j_len;
        if (PROBE) {
            int tile = int(item) * fp.world + fp.rank;
            int ty = tile / fp.tiles_x, tx = tile - ty * fp.tiles_x;
            int r = ty * kTileH + (lane >> 3), c = tx * kTileW + (lane & 7);
            if (r < fp.cam.rows && c < fp.cam.cols) my_pixel = r * fp.cam.cols + c;
            n_entries = 32;
            j_begin = 0;
            j_len = fp.n_probe;
        } else {
            unsigned seg = unsigned(item / (unsigned long long)n_chunks);
            int chunk = int(item - (unsigned long long)seg * n_chunks);
            unsigned e = seg * 32u + lane;
            if (e < n_list) my_pixel = int(fp.list[e]);
            n_entries = int(min(32u, n_list - seg * 32u));
            j_begin = chunk * fp.chunk;
            j_len = min(fp.chunk, n_local - j_begin);
        }
        ws->acc[0][0][lane] = 0; ws->acc[0][1][lane] = 0; ws->acc[0][2][lane] = 0;
        if (PROBE) { ws->acc[1][0][lane] = 0; ws->acc[1][1][lane] = 0; ws->acc[1][2][lane] = 0; }
        ws->pix[lane] = my_pixel < 0 ? -1 : (((my_pixel / fp.cam.cols) << 16) | (my_pixel % fp.cam.cols));
        if (lane == 0) ws->cursor = 0;
        __syncwarp();
        const int pool = n_entries * j_len;

        // ---- drain the pool ----
        // No lane leaves this loop before the whole warp is done: the full-mask vote below is the point where the
        // warp reconverges every iteration (lanes that regenerate a path and lanes that do not would otherwise drift
        // apart for good, and the traversal would run at a fraction of the warp width).
        PathState ps;
        bool active = false, dry = false;
        int slot = 0, set = 0;
        for (;;) {
            if (!active && !dry) {
                int q = atomicAdd(&ws->cursor, 1);
                if (q >= pool) {
                    dry = true;
                } else {
                    int j = q / n_entries;
                    slot = q - j * n_entries;
                    int rc = ws->pix[slot]; // row << 16 | col, or -1 where the tile overhangs the image edge
                    if (rc >= 0) {
                        uint32_t sample = uint32_t(PROBE ? j : fp.sample_begin + fp.rank + (j_begin + j) * fp.world);
                        set = (PROBE && j > fp.first_trial) ? 1 : 0;
                        ++n_paths;
                        active = path_begin(ps, fp.cam, fp.k0, fp.k1, rc >> 16, rc & 0xffff, sample); // false: Ray.make' failed (the reference throws)
                    }
                }
            }
            if (!__any_sync(0xffffffffu, active || !dry)) break;
            if (active) {
                uint32_t result;
                ++n_rays;
                if (path_step<SMEM, COUNT>(ps, sc, fp.cam.depth, result, cn)) {
                    atomicAdd(&ws->acc[set][0][slot], int((result >> 16) & 255u));
                    atomicAdd(&ws->acc[set][1][slot], int((result >> 8) & 255u));
                    atomicAdd(&ws->acc[set][2][slot], int(result & 255u));
                    active = false;
                }
            }
        }
        __syncwarp();

        // ---- retire the item ----
        if (my_pixel >= 0) {
            if (PROBE) {
                int a0 = ws->acc[0][0][lane], a1 = ws->acc[0][1][lane], a2 = ws->acc[0][2][lane];
                int b0 = a0 + ws->acc[1][0][lane], b1 = a1 + ws->acc[1][1][lane], b2 = a2 + ws->acc[1][2][lane];
                int n_old = fp.first_trial + 1, n_new = fp.n_probe;
                // PixelStats.mean (Pixel.fs:103-108): truncating integer division; Pixel.difference :113-116
                int diff = abs(b0 / n_new - a0 / n_old) + abs(b1 / n_new - a1 / n_old) + abs(b2 / n_new - a2 / n_old);
                reinterpret_cast<int4 *>(fp.stats)[my_pixel] = make_int4(b0, b1, b2, n_new);
                fp.flags[my_pixel] = (diff != 0 && fp.sample_end > fp.sample_begin) ? 1 : 0;
            } else {
                int *st = fp.stats + 4 * size_t(my_pixel);
                atomicAdd(st + 0, ws->acc[0][0][lane]);
                atomicAdd(st + 1, ws->acc[0][1][lane]);
                atomicAdd(st + 2, ws->acc[0][2][lane]);
                atomicAdd(st + 3, j_len);
            }
        }
        __syncwarp();
    }
    flush_counters(fp.counters, n_paths, n_rays, cn, COUNT);
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
    int blocks;
};

typedef void (*RenderKernelFn)(const FrameParams);
static RenderKernelFn pick_kernel(bool probe, bool smem, bool count) {
    if (probe) {
        if (smem) return count ? render_kernel<true, true, true> : render_kernel<true, true, false>;
        return count ? render_kernel<true, false, true> : render_kernel<true, false, false>;
    }
    if (smem) return count ? render_kernel<false, true, true> : render_kernel<false, true, false>;
    return count ? render_kernel<false, false, true> : render_kernel<false, false, false>;
}

// lays out shared memory and sizes the persistent grid
static int plan_launch(DeviceScene *ds, FrameParams &fp, bool probe, bool count, bool no_smem, LaunchPlan &plan, RenderKernelFn &fn) {
    const size_t warp_q = (kBlockThreads / 32) * sizeof(WarpScratch) / 16;
    const size_t nodes_q = size_t(ds->g.n_nodes) * 4, sph_q = size_t(ds->g.n_bounded), mat_q = size_t(ds->g.n_bounded + ds->g.n_unbounded) * 2;
    const size_t scene_q = nodes_q + sph_q + mat_q;
    // stage the scene in shared memory when it fits beside the per-warp scratch (one block per SM)
    bool smem = (scene_q + warp_q) * 16 + 1024 <= ds->ws->smem_optin && ds->g.n_bounded > 0;
    if (no_smem) smem = false;
    fp.s_nodes = 0;
    fp.s_spheres = uint32_t(nodes_q);
    fp.s_mats = uint32_t(nodes_q + sph_q);
    fp.s_warp = smem ? uint32_t(scene_q) : 0u;
    plan.smem = smem;
    plan.smem_bytes = ((smem ? scene_q : 0) + warp_q) * 16;
    fn = pick_kernel(probe, smem, count);
    RT_CUDA(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(plan.smem_bytes)));
    int per_sm = 0;
    RT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)fn, kBlockThreads, plan.smem_bytes));
    if (per_sm < 1) return fail(RT_ERR_CUDA, "render kernel does not fit on an SM");
    plan.blocks = per_sm * ds->ws->sm_count;
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
}

// picks the samples-per-item of the main phase: enough items to balance the persistent warps, long
// enough runs that the tail of a pool (lanes idling on the last paths) stays small
static int pick_chunk(int n_local, size_t n_pixels, int resident_warps) {
    if (n_local <= 0) return 1;
    size_t segs = (n_pixels + 31) / 32;
    int chunk = 64;
    while (chunk > 8 && segs * size_t((n_local + chunk - 1) / chunk) < size_t(resident_warps) * 24) chunk /= 2;
    return std::min(chunk, std::max(1, n_local));
}

int launch_probe(DeviceScene *ds, FrameParams fp, bool count, bool no_smem, cudaStream_t st, int *launches) {
    RT_CUDA(cudaMemsetAsync(ds->ws->d_counters + CN_WORK_PROBE, 0, sizeof(unsigned long long), st));
    LaunchPlan plan;
    RenderKernelFn fn;
    int rc = plan_launch(ds, fp, true, count, no_smem, plan, fn);
    if (rc != RT_OK) return rc;
    fn<<<plan.blocks, kBlockThreads, plan.smem_bytes, st>>>(fp);
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
    fp.chunk = pick_chunk(n_local, n_pixels, plan.blocks * (kBlockThreads / 32));
    if (n_local > 0) {
        fn<<<plan.blocks, kBlockThreads, plan.smem_bytes, st>>>(fp);
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
    const size_t warp_q = (kBlockThreads / 32) * sizeof(WarpScratch) / 16;
    const size_t scene_q = size_t(ds->g.n_nodes) * 4 + size_t(ds->g.n_bounded) + size_t(ds->g.n_bounded + ds->g.n_unbounded) * 2;
    bool smem = (scene_q + warp_q) * 16 + 1024 <= ds->ws->smem_optin && ds->g.n_bounded > 0;
    return smem ? scene_q * 16 : 0;
}

int rt_device_probe(RtScene *scene, const RtCamera *camera, int32_t max_w, int32_t max_h, const RtRenderOpts *opts, int32_t rank,
                    int32_t world, int32_t *d_stats, uint8_t *d_flags, void *stream, RtStats *stats) {
    int rc = check_frame_args(scene, camera, max_w, max_h, opts);
    if (rc != RT_OK) return rc;
    if (world < 1 || rank < 0 || rank >= world || !d_stats || !d_flags) return fail(RT_ERR_INVALID_ARGUMENT, "rt_device_probe: bad rank/world/buffers");
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
    rc = launch_main(ds, fp, single_flags(ds->ws->d_flags), count, no_smem, st, &launches);
    if (rc != RT_OK) return rc;
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
        stats->launches = launches;
    }
    return RT_OK;
}

} // extern "C"
