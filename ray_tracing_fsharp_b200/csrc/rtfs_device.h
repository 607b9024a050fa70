// rtfs_device.h — declarations shared by the device translation units of librtfs_b200.so.
#pragma once
#include "rtfs_core.cuh"

#include <string>
#include <vector>

namespace rtfs {

// ---------------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------------
#define RT_CUDA(expr)                                                                                         \
    do {                                                                                                      \
        cudaError_t e__ = (expr);                                                                             \
        if (e__ != cudaSuccess)                                                                               \
            return fail(RT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));                    \
    } while (0)

inline int require_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(RT_ERR_NO_DEVICE, "no CUDA device is visible: librtfs_b200 has no CPU fallback");
    }
    if (device < 0 || device >= n) return fail(RT_ERR_INVALID_ARGUMENT, "CUDA device ordinal out of range");
    RT_CUDA(cudaSetDevice(device));
    return RT_OK;
}

// ---------------------------------------------------------------------------------------------------
// device scene
// ---------------------------------------------------------------------------------------------------
enum CounterSlot { CN_PATHS = 0, CN_RAYS = 1, CN_BOX = 2, CN_PRIM = 3, CN_LIST = 4, CN_WORK_PROBE = 5, CN_WORK_MAIN = 6, CN_DEGENERATE = 7, CN_BUSY_BLOCKS = 8, CN_SLOTS = 16 };
// (the last slot of the pinned copy is a scratch word of the wavefront driver; the pinned buffer holds a second set of
// CN_SLOTS words: the snapshot taken between the probe and the main phase)

// Frame buffers, scratch, stream and events of one device.  Allocating these costs milliseconds, so they are
// pooled: a scene handle borrows one for its lifetime and returns it on destroy (handles stay independent of
// one another; nothing is shared between two live handles).
struct DeviceWorkspace {
    int device = 0;
    int sm_count = 0;
    size_t smem_optin = 0;
    size_t probe_pixels = 0; // second accumulator set of the probe phase
    int32_t *d_stats_b = nullptr;
    size_t ws_pixels = 0; // frame buffers of rt_render
    int32_t *d_stats = nullptr;
    uint8_t *d_flags = nullptr;
    uint8_t *d_rgb = nullptr;
    size_t list_pixels = 0; // flagged-pixel list of the main phase
    uint32_t *d_list = nullptr;
    // the scene blob of the handle that holds this workspace, and its pinned staging copy: pooled with the workspace so
    // that creating a scene costs one copy, not a cudaMalloc / cudaFree pair (grow-only)
    void *d_blob = nullptr;
    void *h_blob = nullptr;
    size_t blob_cap = 0;
    unsigned long long *d_counters = nullptr;
    unsigned long long *h_counters = nullptr; // pinned
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; // begin, zeroed, traced, end, main begin, main end
};
int workspace_acquire(int device, DeviceWorkspace **out);
void workspace_release(DeviceWorkspace *ws);

struct PooledTexture { // an image texture's array + texture object, pooled by (device, width, height)
    int device, w, h;
    cudaArray_t array;
    cudaTextureObject_t object;
};
struct DeviceScene {
    int device = 0;
    SceneGlobal g{};
    DRefNode *ref_nodes = nullptr;
    int32_t n_ref_nodes = 0;
    void *blob = nullptr; // nodes | spheres | materials | unbounded | textures, in the workspace's pooled allocation (not owned)
    void *wide_blob = nullptr; // the 8-wide compressed tree: nodes | spheres in its order | the two index maps
    int32_t wide_depth = 0;
    std::vector<PooledTexture> images; // image textures borrowed from the pool
    size_t bytes = 0;
    int32_t max_depth = 0; // depth of the tree: the walk's stack never holds more entries
    bool lean = false;     // no FP64 unbounded object, no texture index: the staged kernels have a variant without that code
    DeviceWorkspace *ws = nullptr;
};

constexpr int kTileW = 8, kTileH = 4; // a warp's 32 pixels
constexpr int kMaxChunks = 192;
constexpr int kMaxChunkSamples = 256;                    // samples of one pixel in a work item (16-bit colour sums: 256 x 255 < 2^16)
constexpr int kMaxLaunchSamples = 160 * kMaxChunkSamples; // sample indices of one rank a single main-phase launch covers (build_chunks)
// Launch shape of the render kernels (measured on B200).
//   scene staged in shared memory: ONE persistent block of 1024 threads per SM at 64 registers per thread.  That kernel is bound
//     by instruction issue and its walk loop wants the registers (896 threads at 72 registers and 2 x 640 at 48 both lost);
//   scene read from global memory / L2 (big scenes): TWO blocks of 768 threads per SM at 40 registers per thread.  That kernel
//     waits on L2 (long-scoreboard stalls), so 48 warps per SM hide more latency than 32: 100 k spheres at 32 spp 219 -> 203 ms;
//     the spills land outside the walk loop.  (2 x 640 at 48 registers: 211 ms; 2 x 1024 at 32: 253 ms, spills inside the loop.)
// The macros exist for A/B builds.
#ifndef RTFS_BLOCK_THREADS
#define RTFS_BLOCK_THREADS 1024
#endif
#ifndef RTFS_BLOCKS_PER_SM
#define RTFS_BLOCKS_PER_SM 1
#endif
#ifndef RTFS_GLOBAL_BLOCK_THREADS
#define RTFS_GLOBAL_BLOCK_THREADS 768
#endif
#ifndef RTFS_GLOBAL_BLOCKS_PER_SM
#define RTFS_GLOBAL_BLOCKS_PER_SM 2
#endif
#ifndef RTFS_ITEM_SLOTS
#define RTFS_ITEM_SLOTS 4
#endif
#ifndef RTFS_GLOBAL_ITEM_SLOTS
#define RTFS_GLOBAL_ITEM_SLOTS 4
#endif
constexpr int kBlockThreads = RTFS_BLOCK_THREADS;
constexpr int kBlocksPerSm = RTFS_BLOCKS_PER_SM;
constexpr int kGlobalBlockThreads = RTFS_GLOBAL_BLOCK_THREADS;
constexpr int kGlobalBlocksPerSm = RTFS_GLOBAL_BLOCKS_PER_SM;
// Work-item slots per warp (rtfs_device.cu, 400 bytes each).  Four keep a warp fed best: on the 100 k-sphere scene (two blocks of
// 768 threads per SM) four slots 199.5 ms, three 202, two 222 (warps run dry).
constexpr int kItemSlots = RTFS_ITEM_SLOTS, kGlobalItemSlots = RTFS_GLOBAL_ITEM_SLOTS;
constexpr int item_slots(bool staged) { return staged ? kItemSlots : kGlobalItemSlots; }
constexpr int block_threads(bool staged) { return staged ? kBlockThreads : kGlobalBlockThreads; }
constexpr int blocks_per_sm(bool staged) { return staged ? kBlocksPerSm : kGlobalBlocksPerSm; }

struct FrameParams {
    SceneGlobal g;
    DevCamera cam;
    uint32_t k0, k1;
    int32_t rank, world;
    int32_t tiles_x, tiles_y;
    int32_t first_trial;  // min 5 (spp / 2), Scene.fs:172
    int32_t n_probe;      // 2 * first_trial + 1
    int32_t sample_begin; // first sample index of the main phase
    int32_t sample_end;   // one past the last
    int32_t adaptive;
    // work items: (unit, chunk k) where a unit is 32 pixels and chunk k covers this rank's local sample indices
    // [chunk_begin[k], chunk_begin[k] + chunk_len[k]); chunks are sorted by decreasing length and items are
    // numbered chunk-major, so the persistent warps meet the long items first and the short ones last
    int32_t n_chunks;
    uint32_t chunk_begin[kMaxChunks];
    uint32_t chunk_len[kMaxChunks];
    int32_t *stats;       // rows*cols*4 {sumR, sumG, sumB, count}
    int32_t *stats_b;     // probe phase: sums of the samples after the first firstTrial + 1 (same layout)
    uint8_t *flags;       // rows*cols
    const uint32_t *list; // flagged pixel ids (main phase)
    unsigned long long *counters;
    // shared-memory staging
    uint32_t s_nodes, s_spheres, s_mats, s_warp; // offsets in uint4 units
    uint32_t s_stack; // SSTACK kernels: the walk stacks, one column of `stack_levels` words per thread (uint4 units)
    int32_t opt_flags; // RtRenderOpts.flags
    int32_t stack_levels; // words per thread of the shared-memory walk stacks
    int32_t quantum;      // flow kernel: node visits between two schedule points
};


constexpr int kMaxDevices = 16;
// where the main phase finds the probe flags: one buffer (single device, or already combined by the
// caller's all-reduce), or one buffer per rank, each holding the flags of the tiles that rank probed
struct FlagsView {
    const uint8_t *by_rank[kMaxDevices];
    int32_t world;
};
inline FlagsView single_flags(const uint8_t *p) {
    FlagsView v{};
    v.by_rank[0] = p;
    v.world = 1;
    return v;
}

int check_frame_args(const RtScene *scene, const RtCamera *camera, int max_w, int max_h, const RtRenderOpts *opts);
void fill_frame(FrameParams &fp, DeviceScene *ds, const RtCamera &cam, int max_w, int max_h, const RtRenderOpts &opts, int rank, int world);
int launch_probe(DeviceScene *ds, FrameParams fp, bool count, bool no_smem, cudaStream_t st, int *launches);
int launch_main(DeviceScene *ds, FrameParams fp, const FlagsView &flags, bool count, bool no_smem, cudaStream_t st, int *launches);
// makes sure the tree the frame will walk is on the device (the wide tree is built on first use for small scenes)
int prepare_frame(RtScene *scene, const RtRenderOpts *opts);
bool frame_walks_wide_tree(const DeviceScene *ds, int opt_flags, bool smem_fits);
void read_counters(DeviceScene *ds, RtStats *stats, size_t n_pixels, bool adaptive);

inline DevCamera make_dev_camera(const RtCamera &c, int max_w, int max_h) {
    DevCamera d{};
    d.ox = float(c.view_origin[0]); d.oy = float(c.view_origin[1]); d.oz = float(c.view_origin[2]);
    d.cx = float(c.xaxis_origin[0] - c.view_origin[0]);
    d.cy = float(c.xaxis_origin[1] - c.view_origin[1]);
    d.cz = float(c.xaxis_origin[2] - c.view_origin[2]);
    d.xx = float(c.xaxis_dir[0]); d.xy = float(c.xaxis_dir[1]); d.xz = float(c.xaxis_dir[2]);
    d.yx = float(c.yaxis_dir[0]); d.yy = float(c.yaxis_dir[1]); d.yz = float(c.yaxis_dir[2]);
    d.sx = float(c.viewport_width / double(max_w));
    d.sy = float(c.viewport_height / double(max_h));
    d.max_w = max_w;
    d.max_h = max_h;
    d.rows = 2 * max_h + 1; // Scene.fs:208-209
    d.cols = 2 * max_w + 1;
    d.spp = c.samples_per_pixel;
    d.depth = c.bounce_depth;
    return d;
}


} // namespace rtfs
