/* rtfs_b200.h — C ABI of the B200-native path-tracing core for Smaug123/ray-tracing-fsharp.
 *
 * This is the drop-in boundary for the reference's per-pixel render loop.  The reference has no
 * FFI of its own; the seam it offers is the F# function
 *
 *   Scene.render : (float<progress> -> unit) -> (string -> unit) -> maxWidthCoord:int ->
 *                  maxHeightCoord:int -> Camera -> Scene -> float<progress> * Image
 *                                                            (RayTracing/Scene.fs:196-204)
 *   Scene.make   : Hittable array -> Scene                   (RayTracing/Scene.fs:15-28)
 *
 * so every entry point below cites the reference function it replaces.  The F#-side P/Invoke
 * binding a maintainer would add is in INTEGRATION.md and shim/RayTracing.Gpu.fs.
 *
 * Conventions
 *   - plain C types only; doubles on the ABI because the host types are double
 *     (RayTracing/Point.fs:8-11); the device computes in FP32 (see DESIGN.md for the three
 *     places where FP64 is used on the device for exactness).
 *   - every function returns RT_OK (0) or a negative RT_ERR_* code; nothing throws or aborts across
 *     the boundary; rt_last_error() gives a thread-local message for the last failure.
 *   - the caller owns every buffer it passes; the library copies what it needs.  Opaque handles
 *     (RtScene*) own device memory and are released with rt_scene_destroy.  A handle is not
 *     thread-safe; distinct handles are independent.
 *   - there is NO CPU fallback: entry points that compute fail with RT_ERR_NO_DEVICE when no
 *     CUDA device is present.  Host-side helpers (camera basis, BVH build, P3 writer) run anywhere.
 */
#ifndef RTFS_B200_H
#define RTFS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_ABI_VERSION 2

/* ---- error codes --------------------------------------------------------------------------- */
#define RT_OK 0
#define RT_ERR_INVALID_ARGUMENT (-1)
#define RT_ERR_CUDA (-2)
#define RT_ERR_NO_DEVICE (-3)
#define RT_ERR_UNSUPPORTED (-4) /* e.g. a texture the device cannot evaluate (F# closures)       */
#define RT_ERR_DEGENERATE (-5)  /* where the reference would throw from ValueOption.get          */
#define RT_ERR_IO (-6)

/* ---- scene description --------------------------------------------------------------------- */

/* Hittable cases, RayTracing/Hittable.fs:3-6. */
enum RtShape { RT_SHAPE_SPHERE = 0, RT_SHAPE_UNBOUNDED_SPHERE = 1, RT_SHAPE_INFINITE_PLANE = 2 };

/* SphereStyle cases, RayTracing/Sphere.fs:10-37.  InfinitePlaneStyle (InfinitePlane.fs:3-13) uses
 * the subset LIGHT_SOURCE, PURE_REFLECTION, LAMBERT_REFLECTION, FUZZED_REFLECTION. */
enum RtStyle {
    RT_STYLE_LIGHT_SOURCE = 0,
    RT_STYLE_LIGHT_SOURCE_CAP = 1,
    RT_STYLE_PURE_REFLECTION = 2,
    RT_STYLE_FUZZED_REFLECTION = 3,
    RT_STYLE_LAMBERT_REFLECTION = 4,
    RT_STYLE_DIELECTRIC = 5,
    RT_STYLE_GLASS = 6
};

/* Texture / ParameterisedTexture cases, RayTracing/Texture.fs:5-22.  F# closures
 * (Texture.Arbitrary, ParameterisedTexture.Arbitrary) cannot cross the ABI; the shim must pass the
 * structure (image, checker) before ParameterisedTexture.toTexture erases it (Texture.fs:69-72). */
enum RtTextureKind { RT_TEX_COLOUR = 0, RT_TEX_IMAGE = 1, RT_TEX_CHECKERED = 2 };

typedef struct RtTexture {
    int32_t kind;        /* RtTextureKind                                                        */
    uint8_t colour[3];   /* RT_TEX_COLOUR                                                        */
    uint8_t _pad0;
    int32_t width;       /* RT_TEX_IMAGE: img.[0].Length                                         */
    int32_t height;      /* RT_TEX_IMAGE: img.Length                                             */
    const uint8_t *rgb8; /* RT_TEX_IMAGE: the Pixel[][] of ParameterisedTexture.Image, row-major
                            img.[y].[x], 3 bytes per texel, i.e. AFTER ofImage's row flip
                            (Texture.fs:30-48)                                                   */
    int32_t even;        /* RT_TEX_CHECKERED: texture index used when sin(g u) sin(g v) < 0      */
    int32_t odd;         /* RT_TEX_CHECKERED: texture index used otherwise (Texture.fs:56-62)    */
    double grid_size;    /* RT_TEX_CHECKERED                                                     */
    /* the `interpret` closure is always Sphere.planeMapInverse radius centre (Sphere.fs:55-61) */
    double map_centre[3];
    double map_radius;
} RtTexture;

/* One element of the Hittable array handed to Scene.make (Scene.fs:15).  Array order is
 * significant: it is the reference's tie-break order (Scene.fs:47, :77-86). */
typedef struct RtHittable {
    int32_t shape;     /* RtShape                                                                */
    int32_t style;     /* RtStyle                                                                */
    double p[3];       /* Sphere.Centre (Sphere.fs:305) or InfinitePlane.Point (InfinitePlane.fs:105) */
    double n[3];       /* InfinitePlane.Normal (unit); ignored for spheres                       */
    double radius;     /* Sphere.Radius; negative radius is legal (Sphere.fs:160-182)            */
    double albedo;     /* float<albedo>                                                          */
    double fuzz;       /* float<fuzz>,  FUZZED_REFLECTION                                        */
    double ior;        /* float<ior>,   DIELECTRIC / GLASS                                       */
    double prob;       /* float<prob>,  DIELECTRIC: probability of refraction                    */
    int32_t texture;   /* index into textures[], or -1: constant `colour`                        */
    uint8_t colour[3]; /* Texture.Colour / LightSourceCap colour / plane colour                  */
    uint8_t _pad0;
} RtHittable;

/* Camera record, RayTracing/Camera.fs:3-28.  ViewportYAxis' origin is never read by the render
 * loop (Scene.fs:139-140 uses only its Vector) so only the direction crosses the ABI. */
typedef struct RtCamera {
    double view_origin[3];  /* View.Origin                                                      */
    double view_dir[3];     /* View.Vector                                                      */
    double xaxis_origin[3]; /* ViewportXAxis.Origin                                             */
    double xaxis_dir[3];    /* ViewportXAxis.Vector                                             */
    double yaxis_dir[3];    /* ViewportYAxis.Vector                                             */
    double viewport_width;
    double viewport_height;
    double focal_length;
    int32_t samples_per_pixel;
    int32_t bounce_depth;
} RtCamera;

enum RtMode { RT_MODE_MEGAKERNEL = 0, RT_MODE_WAVEFRONT = 1 };
enum RtBvhKind { RT_BVH_SAH = 0, RT_BVH_REFERENCE = 1 };

typedef struct RtRenderOpts {
    uint64_t seed;    /* key of the counter-based RNG; replaces `FloatProducer (Random ())` (Scene.fs:205) */
    int32_t adaptive; /* 1: the reference's early-out rule (Scene.fs:172-194); 0: always spp samples */
    int32_t mode;     /* RtMode                                                                  */
    int32_t gamma;    /* rt_render only: 1 = apply PixelOutput.correct (ImageOutput.fs:11-18) on the device */
    int32_t flags;    /* RT_FLAG_* bits                                                           */
} RtRenderOpts;

/* RtRenderOpts.flags */
#define RT_FLAG_COUNTERS 1 /* also count the device traversal's box / primitive tests (slower kernel variant) */
#define RT_FLAG_NO_SMEM 2  /* read the BVH from global memory even when it would fit in shared memory      */
#define RT_FLAG_WIDE_BVH 4 /* walk the 8-wide compressed tree (read from global memory) instead of the binary one:
                              measured slower on B200 for every scene tried, kept selectable          */
#define RT_FLAG_BVH2 8     /* walk the binary tree (the default; overrides RT_FLAG_WIDE_BVH)                      */
#define RT_FLAG_LOCKSTEP 16 /* the default schedule, spelt out: the 32 lanes of a warp trace one ray each per pass      */
#define RT_FLAG_FLOW 32     /* the flow schedule: every warp keeps a ring of 64 rays in shared memory, lanes pull walks
                              from it in quanta, scatters run in full-width passes.  Bit-identical frames; measured
                              slower than lockstep on B200, kept selectable                          */
#define RT_FLAG_NO_LEAN 64  /* do not pick the kernel specialised for scenes without FP64 unbounded objects and without
                              texture lookups even where the scene allows it (A/B tests; frames are bit-identical)      */

typedef struct RtStats {
    uint64_t paths;     /* traceOnce calls (Scene.fs:118)                                        */
    uint64_t rays;      /* hitObject calls (Scene.fs:99)                                         */
    uint64_t box_tests; /* slab tests executed by the device traversal (its own, not the reference's) */
    uint64_t prim_tests;
    double kernel_ms;   /* device time of the trace kernels (CUDA events)                        */
    double total_ms;    /* device time of the whole call incl. copies                            */
    uint64_t pixels_early_out;
    int32_t launches;   /* kernels launched by this call                                         */
    int32_t _pad0;
    double main_ms;     /* device time of the main-phase kernel alone (compact + render_kernel<PROBE=0>) */
    uint64_t main_rays; /* hitObject calls of that launch (this rank's)                          */
    uint64_t degenerate_paths; /* paths ended where the reference would throw (Ray.make' / ValueOption.get failing):
                                  rendered Black, counted here; 0 on non-degenerate scenes        */
} RtStats;

typedef struct RtScene RtScene;

/* ---- library ------------------------------------------------------------------------------- */
int rt_abi_version(void);
const char *rt_last_error(void);
/* number of CUDA devices visible (0 on a CPU-only host; never fails) */
int rt_device_count(void);

/* ---- host-side helpers (no GPU needed) ------------------------------------------------------ */

/* Camera.makeBasic (RayTracing/Camera.fs:34-59) incl. Plane.makeNormalTo' and Plane.basis
 * (RayTracing/Plane.fs:22-38, :82-97).  bounce_depth is set to 150 as in the reference
 * (Camera.fs:58); callers override the field afterwards like `{ camera with BounceDepth = 50 }`. */
int rt_camera_make_basic(int32_t samples_per_pixel, double focal_length, double aspect_ratio,
                         const double origin[3], const double view_direction[3],
                         const double view_up[3], RtCamera *out);

/* ImageOutput.writePpm (RayTracing/ImageOutput.fs:163-197): P3 text, byte-identical format.
 * `rgb` is rows*cols*3, row 0 first.  gamma_correct applies PixelOutput.correct.  Writes at most
 * `cap` bytes into `out` and returns the full length in *len (call with cap = 0 to size it). */
int rt_ppm_format(const uint8_t *rgb, int32_t rows, int32_t cols, int32_t gamma_correct, char *out,
                  size_t cap, size_t *len);
int rt_ppm_write_file(const uint8_t *rgb, int32_t rows, int32_t cols, int32_t gamma_correct,
                      const char *path);
/* PixelOutput.correct (ImageOutput.fs:11-18) for one byte (host). */
uint8_t rt_gamma_correct(uint8_t b);

/* ---- scene --------------------------------------------------------------------------------- */

/* Scene.make (Scene.fs:15-28) + BoundingBoxTree.make (BoundingBoxTree.fs:9-43), rebuilt as a
 * flattened BVH in device memory.  `device` is a CUDA ordinal.  Pass device = -1 to build the
 * host side only (no GPU required; for rt_scene_bvh_* inspection). */
int rt_scene_create(const RtHittable *objects, int32_t n_objects, const RtTexture *textures,
                    int32_t n_textures, int32_t device, RtScene **out);
void rt_scene_destroy(RtScene *scene);

/* Inspection of the two host-built trees (CPU tests use these).
 * Node layout (flattened, DFS pre-order, left child = i+1):
 *   bounds[6*i..] = min xyz, max xyz ; right[i] = index of right child, or -1 for a leaf ;
 *   prim[i] = index into the caller's Hittable array for a leaf, else -1. */
int rt_scene_bvh_node_count(const RtScene *scene, int32_t which /* RtBvhKind */);
int rt_scene_bvh_nodes(const RtScene *scene, int32_t which, double *bounds, int32_t *right,
                       int32_t *prim);
/* The third tree: the 8-wide compressed BVH the render kernels walk for scenes read from global memory (quantised
 * child boxes, 80-byte nodes; replaces BoundingBoxTree.fs:9-43 / Scene.fs:30-60 for those scenes).  Builds it on the
 * host if need be and verifies it: every bounded sphere in exactly one leaf slot, every decoded box containing what
 * lies below it (RT_ERR_DEGENERATE otherwise).  Outputs are optional. */
int rt_scene_wide_bvh_check(RtScene *scene, int32_t *n_nodes, int32_t *depth, int32_t *n_spheres,
                            double *mean_children);
/* The tree as the render kernels read it (binary SAH tree, child boxes as centre / half-extent with the left and right child's
 * values side by side for the packed FP32 slab test; replaces BoundingBoxTree.fs:9-43 / Scene.fs:30-60): verifies on the host
 * that no converted box is smaller than the box it was made from, that every bounded sphere of non-negative radius is the leaf
 * of exactly one node and lies inside that leaf's box, and that the tree is no deeper than the walk stacks assume
 * (RT_ERR_DEGENERATE otherwise).  Outputs are optional. */
int rt_scene_device_bvh_check(RtScene *scene, int32_t *n_nodes, int32_t *depth, int32_t *n_spheres);
/* bytes of scene data resident on the device (nodes + primitives + materials + textures) */
size_t rt_scene_device_bytes(const RtScene *scene);
/* bytes of BVH + spheres + materials that every persistent block stages in shared memory; 0 when the scene
 * is too large for that and is read from global memory (L1 / L2) instead */
size_t rt_scene_shared_memory_bytes(const RtScene *scene);

/* ---- host buffers ----------------------------------------------------------------------------- */
/* Page-locks a host buffer the caller owns (an output frame it renders into again and again) so that the device->host
 * copy of rt_render / rt_multi_render / rt_comm_render runs at PCIe speed and asynchronously instead of through the
 * driver's staging buffer (2.9 MB C2 frame: ~0.25 ms -> ~0.1 ms).  Optional: every entry point accepts pageable memory.
 * The F# host: GCHandle.Alloc(array, GCHandleType.Pinned) + rt_host_pin(AddrOfPinnedObject, length).  Unpin before freeing. */
int rt_host_pin(void *ptr, size_t bytes);
int rt_host_unpin(void *ptr);

/* ---- render (host buffers) ------------------------------------------------------------------ */

/* Scene.render + Image.render (Scene.fs:196-236, Domain.fs:23-24) on one GPU.
 * rgb_out: rows*cols*3 bytes, rows = 2*max_height_coord+1, cols = 2*max_width_coord+1
 * (Scene.fs:208-209), row 0 = top row — exactly the Pixel[][] that Image.render yields
 * (pre-gamma truncated means, Pixel.fs:103-108) unless opts->gamma is set.
 * sums_out (optional): rows*cols*4 int32 {sumR,sumG,sumB,count} as PixelStats holds them
 * (Pixel.fs:78-85).  stats optional. */
int rt_render(RtScene *scene, const RtCamera *camera, int32_t max_width_coord,
              int32_t max_height_coord, const RtRenderOpts *opts, uint8_t *rgb_out,
              int32_t *sums_out, RtStats *stats);

/* The same frame split over the GPUs of this process (the F# host is one process): replicated scene,
 * sample-index split (Scene.fs:191-192 shared out by sample index), probe tiles shared out by tile.
 * The devices exchange flags and sums through NVLink peer memory inside the kernels that consume them
 * (no separate collective): the last kernel sums every device's PixelStats buffer for its slice of the
 * pixels with peer loads, divides (Pixel.fs:103-108), applies gamma and stores RGB8 to pinned host
 * memory.  Results are bit-identical for every device count (integer sums keyed by sample index).
 * All devices must have peer access to one another (RT_ERR_UNSUPPORTED otherwise). */
typedef struct RtMulti RtMulti;
int rt_multi_create(const RtHittable *objects, int32_t n_objects, const RtTexture *textures,
                    int32_t n_textures, const int32_t *devices, int32_t n_devices, RtMulti **out);
int rt_multi_render(RtMulti *multi, const RtCamera *camera, int32_t max_width_coord,
                    int32_t max_height_coord, const RtRenderOpts *opts, uint8_t *rgb_out,
                    int32_t *sums_out, RtStats *stats);
void rt_multi_destroy(RtMulti *multi);
/* create + render + destroy in one call */
int rt_render_multi(const RtHittable *objects, int32_t n_objects, const RtTexture *textures,
                    int32_t n_textures, const int32_t *devices, int32_t n_devices,
                    const RtCamera *camera, int32_t max_width_coord, int32_t max_height_coord,
                    const RtRenderOpts *opts, uint8_t *rgb_out, int32_t *sums_out, RtStats *stats);

/* ---- render (device buffers; one rank of a sample-split job) -------------------------------- */
/* All d_* pointers are device pointers on the scene's device; `stream` is a cudaStream_t.
 * d_stats: rows*cols*4 int32 {sumR,sumG,sumB,count}; d_flags: rows*cols uint8.
 *
 * Phase 1 — renderPixel's first two loops (Scene.fs:172-188) for the pixel tiles owned by
 * rank/world: adds the 2*firstTrial+1 probe samples into d_stats and sets d_flags[p] = 1 where the
 * two truncated means differ (the pixel continues).  With adaptive = 0 this only sets flags.
 * Buffers must be zeroed by the caller before phase 1. */
int rt_device_probe(RtScene *scene, const RtCamera *camera, int32_t max_width_coord,
                    int32_t max_height_coord, const RtRenderOpts *opts, int32_t rank, int32_t world,
                    int32_t *d_stats, uint8_t *d_flags, void *stream, RtStats *stats);
/* Phase 2 — the third loop (Scene.fs:191-192): for every flagged pixel, this rank's share of the
 * remaining sample indices (index mod world == rank) is added into d_stats.  d_flags must hold
 * the combined flags of all ranks (all-reduce max/sum between the phases when world > 1). */
int rt_device_main(RtScene *scene, const RtCamera *camera, int32_t max_width_coord,
                   int32_t max_height_coord, const RtRenderOpts *opts, int32_t rank, int32_t world,
                   int32_t *d_stats, const uint8_t *d_flags, void *stream, RtStats *stats);
/* Work counters of this scene accumulated since its last rt_device_probe call: paths, rays, (with
 * RT_FLAG_COUNTERS) box_tests / prim_tests; pixels_early_out holds the number of pixels that went ON to
 * phase 2.  One small device->host copy and a synchronisation of `stream`; call it outside timed regions. */
int rt_device_counters(RtScene *scene, void *stream, RtStats *stats);
/* PixelStats.mean (Pixel.fs:103-108) + optional PixelOutput.correct: d_stats (after the sum
 * all-reduce) -> d_rgb rows*cols*3. */
int rt_device_finalize(int32_t device, const int32_t *d_stats, int32_t n_pixels, int32_t gamma,
                       uint8_t *d_rgb, void *stream);

/* ---- render (one process per GPU; the collectives are issued by the library) -------------------- */
/* The sample-split frame of SURVEY.md 8e for a job of `world` processes, one per GPU, each holding a replica of
 * the scene: rank r probes its share of the tiles (Scene.fs:172-188), the flags are all-reduced (MAX), rank r
 * adds its share of the remaining sample indices (Scene.fs:191-192), the PixelStats sums are reduce-scattered
 * (int32 SUM) so that every rank divides (Pixel.fs:103-108) and gamma-corrects (ImageOutput.fs:11-18) one slice
 * of the pixels, and the RGB8 slices are all-gathered (as ONE kernel over NVLink peer memory where the ranks' buffers can
 * be mapped into one another, see rt_comm_uses_peer_memory; else with ncclReduceScatter / ncclAllGather).  All of it — kernels and NCCL calls — is enqueued by
 * rt_comm_render on one stream; the image is bit-identical to rt_render's for every world size.
 * NCCL (libnccl.so.2) is bound at run time on first use: RT_ERR_UNSUPPORTED if it cannot be loaded.
 *
 * Bootstrap as with NCCL itself: one rank calls rt_comm_unique_id, the host program hands the 128 bytes to every
 * rank (the F# host: a file, a pipe or MPI; bench.py: torch.distributed's store), every rank calls rt_comm_create
 * (collective).  `stream`: the cudaStream_t to enqueue on, or NULL for a stream owned by the communicator. */
#define RT_COMM_ID_BYTES 128
typedef struct RtComm RtComm;
int rt_comm_unique_id(uint8_t *id_out /* RT_COMM_ID_BYTES */);
int rt_comm_create(const uint8_t *id, int32_t rank, int32_t world, int32_t device, void *stream,
                   RtComm **out);
void rt_comm_destroy(RtComm *comm); /* collective, like rt_comm_create */
/* 1 when the ranks' sum and frame buffers are mapped into one another (CUDA IPC over NVLink peer access) and the tail of a
 * frame — sum over ranks, divide, gamma, gather — is ONE kernel of peer loads and peer stores between two 4-byte barriers;
 * 0 when that mapping was refused on some rank (or RTFS_COMM_NO_PEER=1) and the tail is ncclReduceScatter + finalize +
 * ncclAllGather.  Decided collectively whenever the communicator (re)allocates its buffers. */
int rt_comm_uses_peer_memory(const RtComm *comm);
/* version (e.g. 22809) and file name of the NCCL build the library bound */
int rt_comm_nccl_version(int32_t *version_out, char *path_out, size_t path_cap);
/* One frame (collective: every rank calls it with the same camera, extents and opts, on its own replica of the
 * scene).  rgb_out / sums_out: host buffers as for rt_render, or NULL on ranks that do not want them (the
 * frame is complete in device memory on every rank either way; passing sums_out makes the ranks all-reduce the
 * sums instead of reduce-scattering them).  stats: THIS rank's paths and rays.  With rgb_out, sums_out and
 * stats all NULL the call only enqueues work and returns without synchronising (device-resident use). */
int rt_comm_render(RtComm *comm, RtScene *scene, const RtCamera *camera, int32_t max_width_coord,
                   int32_t max_height_coord, const RtRenderOpts *opts, uint8_t *rgb_out,
                   int32_t *sums_out, RtStats *stats);
/* Timing and work counters of the last rt_comm_render that was enqueued with rgb_out, sums_out and stats all
 * NULL, once the caller has synchronised the stream: kernel_ms, main_ms, total_ms from the events the call
 * recorded; paths, rays, main_rays from counter snapshots it copied to pinned memory. */
int rt_comm_last_stats(RtComm *comm, RtScene *scene, RtStats *stats);
/* device pointers of the last frame: RGB8 rows*cols*3, and the sum buffer (complete only after a call with
 * sums_out or world == 1; otherwise only this rank's slice holds totals).  Valid until the next call. */
int rt_comm_frame(RtComm *comm, const uint8_t **d_rgb_out, const int32_t **d_stats_out);

/* ---- per-primitive conformance entry points (device) ---------------------------------------- */
/* Each runs one thread per vector through exactly the __device__ function the render kernels call.
 * Inputs/outputs are host pointers; doubles are rounded to FP32 on upload where the device
 * function takes FP32. */

/* Sphere.firstIntersection (Sphere.fs:349-386).  t_out[i] = NaN when there is no hit. */
int rt_test_sphere_hit(int32_t device, int32_t n, const double *origin, const double *dir,
                       const double *centre, const double *radius, double *t_out);
/* InfinitePlane.intersection (InfinitePlane.fs:125-136). */
int rt_test_plane_hit(int32_t device, int32_t n, const double *origin, const double *dir,
                      const double *point, const double *normal, double *t_out);
/* BoundingBox.hits with inverseDirections (BoundingBox.fs:25-94): the decision-exact slab test that the
 * reference-order traversal (rt_test_hit_object traversal = 1) runs.  The render traversal uses a
 * conservative variant of it (padded, culled by the best hit); that one is covered by
 * rt_test_hit_object traversal = 0. */
int rt_test_aabb_hit(int32_t device, int32_t n, const double *origin, const double *dir,
                     const double *box_min, const double *box_max, uint8_t *hit_out);
/* Scene.hitObject (Scene.fs:62-91).  prim_out = index into the caller's Hittable array or -1;
 * traversal: 0 = the render kernels' ordered/culled traversal of the binary SAH tree,
 *            1 = exhaustive left-then-right DFS of the reference-topology tree (F12),
 *            2 = the render kernels' traversal of the 8-wide compressed tree (big scenes, RT_FLAG_WIDE_BVH). */
int rt_test_hit_object(RtScene *scene, int32_t traversal, int32_t n, const double *origin,
                       const double *dir, int32_t *prim_out, double *t_out, double *strike_out);
/* Hittable.Reflection (Hittable.fs:8-12 -> Sphere.fs:150-300 / InfinitePlane.fs:43-99) with
 * explicit uniforms: uniforms[4*i..] are the FloatProducer draws for vector i (GetThree uses
 * [0..2], Get uses [0]); retries re-use the same draws rotated by one (they never happen on
 * non-degenerate inputs).  absorbed_out[i] = 1 when the reference returns ValueSome colour. */
int rt_test_reflection(RtScene *scene, int32_t n, const int32_t *prim, const double *origin,
                       const double *dir, const double *strike, const uint8_t *colour_in,
                       const double *uniforms, uint8_t *absorbed_out, uint8_t *colour_out,
                       double *origin_out, double *dir_out, uint8_t *inside_out);
/* traceOnce's ray generation (Scene.fs:129-144): one ray per (row, col, rand1, rand2), where
 * row/col are the signed coordinates renderPixel receives. */
int rt_test_camera_rays(int32_t device, const RtCamera *camera, int32_t max_width_coord,
                        int32_t max_height_coord, int32_t n, const int32_t *row, const int32_t *col,
                        const double *rand1, const double *rand2, double *origin_out, double *dir_out);
/* Texture.colourAt on the primitive's texture (Texture.fs:12-15, :50-67). */
int rt_test_texture(RtScene *scene, int32_t n, const int32_t *prim, const double *point,
                    uint8_t *colour_out);
/* Pixel.combine + Pixel.darken (Pixel.fs:136-151) on the device, explicit albedo. */
int rt_test_combine_darken(int32_t device, int32_t n, const uint8_t *a, const uint8_t *b,
                           const double *albedo, uint8_t *out);
/* The counter-based RNG: uniforms of (seed, pixel, sample, bounce, retry) as the device draws them. */
int rt_test_rng(int32_t device, uint64_t seed, int32_t n, const uint32_t *pixel,
                const uint32_t *sample, const uint32_t *bounce, const uint32_t *retry,
                uint32_t *words_out /* 4n */, double *uniforms_out /* 4n */);
/* traceOnce for explicit (pixel row/col, sample index): the final Pixel of that one path. */
int rt_test_trace_samples(RtScene *scene, const RtCamera *camera, int32_t max_width_coord,
                          int32_t max_height_coord, uint64_t seed, int32_t n, const int32_t *row_idx,
                          const int32_t *col_idx, const int32_t *sample, uint8_t *colour_out,
                          int32_t *rays_out);

/* FP32 FMA microbenchmark used as the roofline denominator: returns
 * measured TFLOP/s of a dependent-chain FFMA kernel filling the device. */
int rt_measure_fp32_peak(int32_t device, double *tflops_out);

#ifdef __cplusplus
}
#endif
#endif /* RTFS_B200_H */
