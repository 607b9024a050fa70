"""Level-2 parity: rendered frames of the CUDA path against the CPU oracle, through the C ABI and
the mirrored Scene.render surface.

 (i)   same counter RNG on both sides: near-identical bytes (differences only where FP32 flips a decision);
 (ii)  independent RNGs (oracle: the reference's xorshift128; GPU: Philox) at high spp: per-channel mean absolute
       difference of the pre-gamma means under 3 sigma of the per-pixel estimator;
 (iii) size-independent properties at full size: determinism, invariance under the sample split, integer
       accumulator invariants, the early-out rule's count field, gamma == host LUT, P3 bytes.
"""
import ctypes as C

import numpy as np
import pytest

import oracle
from helpers import oracle_camera, scene_pair, small_random_spheres
from ray_tracing_fsharp_b200 import abi, native, sample_images
from ray_tracing_fsharp_b200.domain import marshal
from ray_tracing_fsharp_b200.scene import Camera, Image, ImageOutput, Scene

pytestmark = pytest.mark.gpu


def _small(config, max_w, max_h, spp):
    spec = sample_images.CONFIGS[config]()
    spec.max_width_coord, spec.max_height_coord, spec.spp = max_w, max_h, spp
    return spec


@pytest.mark.parametrize("config,max_w,max_h,spp", [("C1", 100, 56, 16), ("C2", 60, 40, 24), ("C3", 64, 36, 16), ("C4", 64, 36, 32)])
def test_frame_matches_oracle_with_shared_rng(config, max_w, max_h, spp):
    spec = _small(config, max_w, max_h, spp)
    osc, dsc, cam = scene_pair(spec)
    ref, ref_stats, counters, _ = osc.render(cam, max_w, max_h, seed=5, rng_mode=1, adaptive=True)
    rgb, sums, stats = dsc.render(cam, max_w, max_h, seed=5, adaptive=True, want_sums=True)
    assert rgb.shape == ref.shape == (2 * max_h + 1, 2 * max_w + 1, 3)  # F6
    same_px = (rgb == ref).all(2)
    assert same_px.mean() > 0.985, same_px.mean()
    # the integer accumulators themselves: identical for almost every pixel, never far off
    same_sums = (sums == ref_stats).all(2)
    assert same_sums.mean() > 0.97, same_sums.mean()
    assert np.array_equal(sums[..., 3] == spp, ref_stats[..., 3] == spp) or (sums[..., 3] != ref_stats[..., 3]).mean() < 0.01
    assert abs(int(stats.paths) - counters["paths"]) <= 0.01 * counters["paths"]
    assert abs(int(stats.rays) - counters["rays"]) <= 0.01 * counters["rays"]
    assert np.abs(rgb.astype(int) - ref.astype(int)).mean() < 0.3


SAMPLE_SCALES = {"shiny-floor": 0.1, "fuzzy-floor": 0.1, "spheres": 0.25, "inside-sphere": 0.04, "total-refraction": 0.15, "glass": 0.25,
                 "textured-sphere": 0.25, "moved-camera": 0.15}


@pytest.mark.parametrize("name", sorted(sample_images.REFERENCE_SAMPLES))
def test_reference_sample_scene_matches_oracle_with_shared_rng(name):
    """The reference's own sample scenes (RayTracing.App/SampleImages.fs; `random-spheres` and `earth` are C2 and C3 above),
    at reduced half-extents, 50 spp, depth 150 as Camera.makeBasic sets it: bounded spheres of radius 75-200 in the tree,
    a camera inside a bounded sphere (F10), Dielectric with refraction probability 1, a bounded negative-radius shell
    (F16), every InfinitePlane style, a checker of two baked closures."""
    spec = sample_images.REFERENCE_SAMPLES[name](SAMPLE_SCALES[name])
    osc, dsc, cam = scene_pair(spec)
    mw, mh = spec.max_width_coord, spec.max_height_coord
    ref, ref_stats, counters, _ = osc.render(cam, mw, mh, seed=11, rng_mode=1, adaptive=True)
    rgb, sums, stats = dsc.render(cam, mw, mh, seed=11, adaptive=True, want_sums=True)
    same_px = (rgb == ref).all(2)
    same_sums = (sums == ref_stats).all(2)
    print(f"{name}: {rgb.shape[1]}x{rgb.shape[0]}, {int(stats.paths)} paths, {int(stats.rays)} rays; identical pixels {same_px.mean():.4f}, "
          f"identical sums {same_sums.mean():.4f}, mean |diff| {np.abs(rgb.astype(int) - ref.astype(int)).mean():.4f}")
    assert same_px.mean() > 0.995, same_px.mean()  # measured 0.9997-1.0000
    assert same_sums.mean() > 0.99, same_sums.mean()  # measured 0.9984-1.0000
    assert abs(int(stats.rays) - counters["rays"]) <= 0.01 * counters["rays"]
    assert np.abs(rgb.astype(int) - ref.astype(int)).mean() < 0.1
    assert rgb.max() > 0  # something is lit


def test_adaptive_rule_and_counts():
    """F5: count is 2*firstTrial+1 where the two truncated means agree, else spp; spp < 11 gives 2*(spp/2)+1 samples."""
    spec = _small("C1", 80, 45, 16)
    osc, dsc, cam = scene_pair(spec)
    rgb, sums, stats = dsc.render(cam, 80, 45, seed=1, adaptive=True, want_sums=True)
    counts = np.unique(sums[..., 3])
    assert set(counts.tolist()) <= {11, 16} and 11 in counts and 16 in counts
    assert int(stats.pixels_early_out) == int((sums[..., 3] == 11).sum())
    assert int(stats.paths) == int(sums[..., 3].sum())
    # mean = truncating division of the integer sums (Pixel.fs:103-108)
    assert np.array_equal(rgb, (sums[..., :3] // sums[..., 3:4]).astype(np.uint8))
    for spp, expect in [(1, {1}), (4, {5}), (7, {7}), (10, {11}), (11, {11})]:
        cam.samples_per_pixel = spp
        _, s2, st2 = dsc.render(cam, 20, 11, seed=2, adaptive=True, want_sums=True)
        assert set(np.unique(s2[..., 3]).tolist()) == expect, (spp, np.unique(s2[..., 3]))
        ref, rs, _, _ = osc.render(cam, 20, 11, seed=2, rng_mode=1, adaptive=True)
        assert np.array_equal(rs[..., 3], s2[..., 3])
    cam.samples_per_pixel = 16
    _, s3, st3 = dsc.render(cam, 80, 45, seed=1, adaptive=False, want_sums=True)
    assert (s3[..., 3] == 16).all() and int(st3.paths) == 16 * s3.shape[0] * s3.shape[1] and st3.pixels_early_out == 0


def test_shared_memory_and_global_memory_kernels_agree_bit_for_bit():
    spec = small_random_spheres()
    osc, dsc, cam = scene_pair(spec)
    a, sa, _ = dsc.render(cam, spec.max_width_coord, spec.max_height_coord, seed=9, want_sums=True)
    b, sb, _ = dsc.render(cam, spec.max_width_coord, spec.max_height_coord, seed=9, want_sums=True, flags=abi.RT_FLAG_NO_SMEM)
    c, sc, stc = dsc.render(cam, spec.max_width_coord, spec.max_height_coord, seed=9, want_sums=True, flags=abi.RT_FLAG_COUNTERS)
    assert np.array_equal(sa, sb) and np.array_equal(a, b)
    assert np.array_equal(sa, sc) and stc.box_tests > 0 and stc.prim_tests > 0
    d, sd, _ = dsc.render(cam, spec.max_width_coord, spec.max_height_coord, seed=10, want_sums=True)
    assert not np.array_equal(sa, sd)  # the seed matters


@pytest.mark.parametrize("which", ["reduced", "C1", "C3", "C4"])
def test_wavefront_and_megakernel_agree_bit_for_bit(which):
    """RT_MODE_WAVEFRONT runs the same device functions on a different schedule (queues in HBM, one launch pair per
    bounce): identical RNG keys and integer sums => identical accumulators, adaptive and not."""
    if which == "reduced":
        spec = small_random_spheres()
    else:
        spec = _small(which, 48, 27, 24)
    osc, dsc, cam = scene_pair(spec)
    mw, mh = spec.max_width_coord, spec.max_height_coord
    for adaptive in (True, False):
        a, sa, sta = dsc.render(cam, mw, mh, seed=41, adaptive=adaptive, want_sums=True)
        b, sb, stb = dsc.render(cam, mw, mh, seed=41, adaptive=adaptive, want_sums=True, mode=abi.RT_MODE_WAVEFRONT)
        assert np.array_equal(sa, sb) and np.array_equal(a, b)
        assert int(sta.rays) == int(stb.rays) and int(sta.paths) == int(stb.paths)
        assert int(sta.pixels_early_out) == int(stb.pixels_early_out)
        assert stb.launches > sta.launches


def _mid_size_scene(n=6000):
    spec = sample_images.many_spheres(n=n)
    spec.max_width_coord, spec.max_height_coord, spec.spp = 64, 36, 24
    return spec


@pytest.mark.parametrize("which", ["reduced", "C2", "C4", "mid"])
def test_wide_bvh_and_binary_bvh_agree_bit_for_bit(which):
    """The 8-wide compressed tree (csrc/rtfs_core.cuh wide_closest; the default for scenes read from global memory) against
    the binary SAH tree: the closest hit does not depend on the tree, so the same RNG keys and integer sums give identical
    accumulators — with the binary tree staged in shared memory (small scenes) and read from global memory (a scene of
    6000 spheres, too big for shared memory)."""
    if which == "reduced":
        spec = small_random_spheres()
    elif which == "mid":
        spec = _mid_size_scene()
    else:
        spec = _small(which, 48, 27, 24)
    osc, dsc, cam = scene_pair(spec)
    mw, mh = spec.max_width_coord, spec.max_height_coord
    info = dsc.wide_bvh_check()
    assert info["spheres"] == sum(1 for o in spec.objects if type(o).__name__ == "Sphere" and o.sphere.Radius >= 0)
    for adaptive in (True, False):
        a, sa, sta = dsc.render(cam, mw, mh, seed=43, adaptive=adaptive, want_sums=True, flags=abi.RT_FLAG_BVH2)
        sa = sa.copy()
        b, sb, stb = dsc.render(cam, mw, mh, seed=43, adaptive=adaptive, want_sums=True, flags=abi.RT_FLAG_WIDE_BVH)
        assert np.array_equal(sa, sb)
        assert int(sta.rays) == int(stb.rays) and int(sta.paths) == int(stb.paths)
        c, sc_, stc = dsc.render(cam, mw, mh, seed=43, adaptive=adaptive, want_sums=True, flags=abi.RT_FLAG_WIDE_BVH | abi.RT_FLAG_COUNTERS)
        assert np.array_equal(sa, sc_) and stc.box_tests > 0 and stc.prim_tests > 0
        d, sd, std_ = dsc.render(cam, mw, mh, seed=43, adaptive=adaptive, want_sums=True)  # the default: binary, shared memory if it fits
        assert np.array_equal(sa, sd)
        e, se, _ = dsc.render(cam, mw, mh, seed=43, adaptive=adaptive, want_sums=True, mode=abi.RT_MODE_WAVEFRONT)
        assert np.array_equal(sa, se)
    if which == "mid":  # and against the oracle, sample for sample (shared counter RNG)
        ref, ref_stats, counters, _ = osc.render(cam, mw, mh, seed=43, rng_mode=1, adaptive=True)
        rgb, sums, _ = dsc.render(cam, mw, mh, seed=43, adaptive=True, want_sums=True)
        assert (rgb == ref).all(2).mean() > 0.97, (rgb == ref).all(2).mean()


@pytest.mark.parametrize("which", ["reduced", "C1", "C2", "C3", "C4", "mid", "empty-ish"])
def test_flow_and_lockstep_schedules_agree_bit_for_bit(which):
    """The flow schedule (RT_FLAG_FLOW, render_flow_kernel: a ring of 64 rays per warp, walks pulled by the lanes, scatters
    done in full-width passes) against the default lockstep kernel: same RNG keys, same integer sums, so the
    accumulators, the path / ray counts and the early-out decisions must be identical — scene in shared memory, scene in
    global memory, with the counters on, adaptive and not, probe-only frames, and a frame smaller than one warp's ring."""
    if which == "reduced":
        spec = small_random_spheres()
    elif which == "mid":
        spec = _mid_size_scene()
    elif which == "empty-ish":
        spec = _small("C4", 2, 1, 7)  # 5 x 3 pixels: fewer rays than ring slots
    else:
        spec = _small(which, 48, 27, 24)
    osc, dsc, cam = scene_pair(spec)
    mw, mh = spec.max_width_coord, spec.max_height_coord
    for adaptive in (True, False):
        for extra in (0, abi.RT_FLAG_NO_SMEM, abi.RT_FLAG_COUNTERS):
            a, sa, sta = dsc.render(cam, mw, mh, seed=47, adaptive=adaptive, want_sums=True, flags=extra)
            sa = sa.copy()
            b, sb, stb = dsc.render(cam, mw, mh, seed=47, adaptive=adaptive, want_sums=True, flags=extra | abi.RT_FLAG_FLOW)
            assert np.array_equal(sa, sb), (which, adaptive, extra, int((sa != sb).any(2).sum()))
            assert int(sta.rays) == int(stb.rays) and int(sta.paths) == int(stb.paths)
            assert int(sta.pixels_early_out) == int(stb.pixels_early_out)
            if extra == abi.RT_FLAG_COUNTERS:
                assert int(sta.box_tests) == int(stb.box_tests) and int(sta.prim_tests) == int(stb.prim_tests)
    # spp values that exercise the chunk table's tail (1-sample items) and a probe-only frame
    for spp in (1, 5, 11, 12, 33):
        cam.samples_per_pixel = spp
        _, s1, _ = dsc.render(cam, mw, mh, seed=48, adaptive=True, want_sums=True)
        s1 = s1.copy()
        _, s2, _ = dsc.render(cam, mw, mh, seed=48, adaptive=True, want_sums=True, flags=abi.RT_FLAG_FLOW)
        assert np.array_equal(s1, s2), (which, spp)


@pytest.mark.parametrize("which", ["reduced", "C1", "C2"])
def test_lean_kernel_agrees_bit_for_bit(which):
    """Scenes whose unbounded objects are all cleared for FP32 and whose materials carry no texture index run a kernel compiled
    without the FP64 evaluations and the texture lookup (SceneAccess<2>, DeviceScene::lean); RT_FLAG_NO_LEAN keeps the full
    kernel.  Same sums, same counts, adaptive and not, with the counters on."""
    spec = small_random_spheres() if which == "reduced" else _small(which, 48, 27, 24)
    osc, dsc, cam = scene_pair(spec)
    mw, mh = spec.max_width_coord, spec.max_height_coord
    for adaptive in (True, False):
        for extra in (0, abi.RT_FLAG_COUNTERS):
            _, sa, sta = dsc.render(cam, mw, mh, seed=53, adaptive=adaptive, want_sums=True, flags=extra)
            sa = sa.copy()
            _, sb, stb = dsc.render(cam, mw, mh, seed=53, adaptive=adaptive, want_sums=True, flags=extra | abi.RT_FLAG_NO_LEAN)
            assert np.array_equal(sa, sb), (which, adaptive, extra, int((sa != sb).any(2).sum()))
            assert int(sta.rays) == int(stb.rays) and int(sta.paths) == int(stb.paths)
            if extra:
                assert int(sta.box_tests) == int(stb.box_tests) and int(sta.prim_tests) == int(stb.prim_tests)


def test_more_samples_than_one_main_launch_covers():
    """A work item holds at most 256 samples of a pixel (its red and green sums share a word) and one main-phase launch at
    most 160 x 256 sample indices of a rank: a frame with more is several launches.  3 x 3 pixels at 41 500 spp (adaptive off
    and on), against the flow kernel (accumulators of its own, one launch) and against the known totals."""
    spec = _small("C1", 1, 1, 41500)
    osc, dsc, cam = scene_pair(spec)
    for adaptive in (False, True):
        _, s1, st1 = dsc.render(cam, 1, 1, seed=5, adaptive=adaptive, want_sums=True)
        s1 = s1.copy()
        _, s2, st2 = dsc.render(cam, 1, 1, seed=5, adaptive=adaptive, want_sums=True, flags=abi.RT_FLAG_FLOW)
        assert np.array_equal(s1, s2) and int(st1.paths) == int(st2.paths) and int(st1.rays) == int(st2.rays)
        assert set(np.unique(s1[..., 3]).tolist()) <= {11, 41500}
        if not adaptive:
            assert (s1[..., 3] == 41500).all() and int(st1.paths) == 9 * 41500
        assert (s1[..., :3] <= 255 * s1[..., 3:4]).all() and (s1[..., :3] >= 0).all()


def test_stats_of_a_frame_and_page_locked_output_frames():
    """RtStats of rt_render (the main-phase kernel's own time and rays, no degenerate paths on a well-formed scene), and
    native.FrameRing: explicitly recycled, page-locked output arrays (rt_host_pin) receive the same frame as a fresh one."""
    spec = _small("C2", 60, 40, 32)
    osc, dsc, cam = scene_pair(spec)
    mw, mh = spec.max_width_coord, spec.max_height_coord
    a, _, st = dsc.render(cam, mw, mh, seed=21)
    assert 0 < st.main_rays < st.rays and 0.0 < st.main_ms <= st.kernel_ms <= st.total_ms and st.degenerate_paths == 0
    ring = native.FrameRing((spec.rows, spec.cols, 3), count=2)
    f0 = ring.next()
    b, _, _ = dsc.render(cam, mw, mh, seed=21, rgb_out=f0)
    assert b is f0 and np.array_equal(a, b)
    f1 = ring.next()
    c, _, _ = dsc.render(cam, mw, mh, seed=22, rgb_out=f1)
    assert c is f1 and c is not b and not np.array_equal(b, c) and np.array_equal(a, f0)  # the earlier frame is untouched
    assert ring.next() is f0
    assert native.lib().rt_host_unpin(None) == abi.RT_ERR_INVALID_ARGUMENT


def test_hot_pink_when_the_bounce_budget_runs_out():
    """F3: a path that is still alive after maxCount + 1 interactions is HotPink, a miss is Black."""
    from ray_tracing_fsharp_b200.domain import Colour, Hittable, Pixel, Sphere, SphereStyle, Texture
    # camera inside a mirror ball: no emitter can ever be reached
    objs = [Hittable.Sphere(Sphere.make(SphereStyle.PureReflection(1.0, Texture.Colour(Colour.White)), (0.0, 0.0, 0.0), 5.0))]
    hs, ts, keep = marshal(objs)
    dsc = native.SceneHandle(hs, ts, 0, keepalive=keep)
    cam = Camera.make_basic(4, 1.0, 1.0, (0.0, 0.0, 0.0), (0.0, 0.0, 1.0), (0.0, 1.0, 0.0))
    cam.bounce_depth = 7
    rgb, sums, stats = dsc.render(cam, 8, 8, seed=3, adaptive=False, want_sums=True)
    assert (rgb == np.array([205, 105, 180], np.uint8)).all()
    assert int(stats.rays) == int(stats.paths) * 8  # maxCount + 1 interactions
    # empty scene: every ray escapes => Black
    dsc2 = native.SceneHandle([], [], 0)
    rgb2, _, st2 = dsc2.render(cam, 8, 8, seed=3, adaptive=False)
    assert (rgb2 == 0).all() and int(st2.rays) == int(st2.paths)


def test_statistical_parity_at_4096_spp():
    """Independent RNGs: oracle with the reference's xorshift128, GPU with Philox; 4096 spp, adaptive off on both so
    that every pixel is a 4096-sample estimator.  Per channel, MAD of the (pre-gamma, truncated) means must be under
    3 * sigma-bar, sigma-bar = mean over pixels of sqrt(2 s^2 / N) (difference of two independent means)."""
    for config, max_w, max_h in [("C1", 24, 13), ("C2", 24, 16), ("C4", 24, 13)]:
        spec = _small(config, max_w, max_h, 4096)
        osc, dsc, cam = scene_pair(spec)
        ref, ref_stats, _, _ = osc.render(cam, max_w, max_h, seed=77, rng_mode=0, adaptive=False)
        rgb, sums, _ = dsc.render(cam, max_w, max_h, seed=78, adaptive=False, want_sums=True)
        n = 4096
        mean_g = sums[..., :3] / n
        mean_o = ref_stats[..., :3] / n
        # per-sample variance bound from the mean alone: a byte-valued sample x in [0, 255] has s^2 <= m (255 - m);
        # use the empirical variance of the GPU estimator from two half-runs instead, which is tighter
        _, s_a, _ = dsc.render(cam, max_w, max_h, seed=79, adaptive=False, want_sums=True)
        var_of_mean = ((sums[..., :3] - s_a[..., :3]) / n) ** 2 / 2.0  # unbiased estimate of Var(mean) per pixel
        sigma_bar = np.sqrt(2.0 * var_of_mean.mean(axis=(0, 1)))        # difference of two independent means
        mad = np.abs(mean_g - mean_o).mean(axis=(0, 1))
        assert (mad < 3.0 * sigma_bar + 0.5).all(), (config, mad, sigma_bar)  # +0.5: both sides truncate to bytes
        assert np.abs(rgb.astype(int) - ref.astype(int)).mean() < 1.5, config


def test_scene_render_surface_and_ppm_bytes():
    spec = _small("C1", 40, 22, 16)
    cam = Camera.make_basic(spec.spp, spec.focal_length, spec.aspect_ratio, spec.origin, spec.view_direction, spec.view_up)
    assert cam.bounce_depth == 150  # F7
    cam.bounce_depth = spec.bounce_depth
    scene = Scene.make(spec.objects, device=0)
    ticks = []
    total, image = Scene.render(ticks.append, lambda s: None, 40, 22, cam, scene, seed=11)
    assert total == 45.0 and Image.row_count(image) == 45 and Image.col_count(image) == 81 and ticks == []  # lazy
    pixels = Image.render(image)
    assert len(ticks) == 45 and pixels.shape == (45, 81, 3)
    # rows can also be forced one at a time, as `Image.Rows` consumers do (ImageOutput.fs:142): one tick per row
    ticks2 = []
    _, image2 = Scene.render(ticks2.append, lambda s: None, 40, 22, cam, scene, seed=11)
    assert np.array_equal(image2.Rows[3](), pixels[3]) and np.array_equal(image2.Rows[44](), pixels[44]) and len(ticks2) == 2
    # PPM bytes: the device result through the library's writer == the oracle's writer on the same pixels
    assert ImageOutput.write_ppm(False, pixels) == oracle.ppm_format(pixels, False)
    assert ImageOutput.write_ppm(True, pixels) == oracle.ppm_format(pixels, True)
    # gamma on the device == PixelOutput.correct applied on the host
    rgb_g, _, _ = scene.handle.render(cam, 40, 22, seed=11, gamma=True)
    lut = np.array([oracle.gamma_correct(i) for i in range(256)], np.uint8)
    assert np.array_equal(rgb_g, lut[pixels])
    # and against the oracle's frame with the shared RNG
    hs, ts, _ = marshal(spec.objects)
    ref, _, _, _ = oracle.Scene(hs, ts).render(cam, 40, 22, seed=11, rng_mode=1, adaptive=True)
    assert (pixels == ref).all(2).mean() > 0.985


def test_sample_split_is_invariant_emulated_on_one_gpu():
    """The multi-GPU decomposition run rank by rank on one device: summing the ranks' buffers reproduces the
    single-rank frame bit for bit (integer sums keyed by sample index), for world = 1, 2, 3, 8."""
    import torch
    from ray_tracing_fsharp_b200.distributed import DeviceBackend, render_split_frame
    spec = small_random_spheres()
    osc, dsc, cam = scene_pair(spec)
    max_w, max_h = spec.max_width_coord, spec.max_height_coord
    _, base, _ = dsc.render(cam, max_w, max_h, seed=21, adaptive=True, want_sums=True)
    n_pixels = base.shape[0] * base.shape[1]
    for world in (1, 2, 3, 8):
        backends = [DeviceBackend(dsc, cam, max_w, max_h, seed=21, adaptive=True) for _ in range(world)]
        bufs = [b.alloc() for b in backends]
        for r in range(world):
            backends[r].probe(r, world, *bufs[r])
        flags = torch.stack([f for _, f in bufs]).max(0).values.contiguous()  # all_reduce(MAX)
        for r in range(world):
            bufs[r][1].copy_(flags)
            backends[r].main(r, world, *bufs[r])
        total = torch.stack([s for s, _ in bufs]).sum(0).to(torch.int32).contiguous()  # all_reduce(SUM)
        torch.cuda.synchronize()
        assert np.array_equal(total.cpu().numpy().reshape(base.shape), base), world
        rgb = backends[0].finalize(total).cpu().numpy().reshape(base.shape[0], base.shape[1], 3)
        assert np.array_equal(rgb, (base[..., :3] // base[..., 3:4]).astype(np.uint8))
        assert np.array_equal(backends[0].finalize_to_host(total).reshape(rgb.shape), rgb)  # the pinned, recycled host frame
    # non-adaptive split
    _, base2, _ = dsc.render(cam, max_w, max_h, seed=22, adaptive=False, want_sums=True)
    world = 4
    backends = [DeviceBackend(dsc, cam, max_w, max_h, seed=22, adaptive=False) for _ in range(world)]
    bufs = [b.alloc() for b in backends]
    for r in range(world):
        backends[r].probe(r, world, *bufs[r])
    flags = torch.stack([f for _, f in bufs]).max(0).values.contiguous()
    for r in range(world):
        bufs[r][1].copy_(flags)
        backends[r].main(r, world, *bufs[r])
    total = torch.stack([s for s, _ in bufs]).sum(0).to(torch.int32)
    assert np.array_equal(total.cpu().numpy().reshape(base2.shape), base2)


def test_multi_device_frame_is_bit_identical():
    n = native.device_count()
    if n < 2:
        pytest.skip("needs two GPUs with peer access")
    spec = small_random_spheres()
    osc, dsc, cam = scene_pair(spec)
    hs, ts, keep = marshal(spec.objects)
    rgb1, sums1, _ = dsc.render(cam, spec.max_width_coord, spec.max_height_coord, seed=31, want_sums=True)
    for world in sorted({2, min(n, 4), n}):
        m = native.MultiHandle(hs, ts, list(range(world)), keepalive=keep)
        rgb, sums, stats = m.render(cam, spec.max_width_coord, spec.max_height_coord, seed=31, want_sums=True)
        assert np.array_equal(sums, sums1) and np.array_equal(rgb, rgb1)
        m.close()


def test_full_size_properties_c2():
    """BASELINE.json's full frame size (1201x801) at reduced spp: determinism, split invariance of the early-out
    accounting, and the accumulator invariants.  (The 500 spp frame itself is what bench.py renders.)"""
    spec = sample_images.random_spheres(spp=24)
    hs, ts, keep = marshal(spec.objects)
    dsc = native.SceneHandle(hs, ts, 0, keepalive=keep)
    cam = oracle_camera(spec)
    a, sa, sta = dsc.render(cam, 600, 400, seed=1, want_sums=True)
    b, sb, stb = dsc.render(cam, 600, 400, seed=1, want_sums=True)
    assert a.shape == (801, 1201, 3)
    assert np.array_equal(sa, sb) and int(sta.rays) == int(stb.rays)                 # deterministic
    assert set(np.unique(sa[..., 3]).tolist()) <= {11, 24}
    assert int(sta.paths) == int(sa[..., 3].sum())
    assert (sa[..., :3] <= 255 * sa[..., 3:4]).all() and (sa[..., :3] >= 0).all()
    assert np.array_equal(a, (sa[..., :3] // sa[..., 3:4]).astype(np.uint8))
    # the sky (light dome seen directly) is exactly its emitted colour and early-outs
    assert (a[0, 0] == np.array([200, 200, 255], np.uint8)).all() and sa[0, 0, 3] == 11


@pytest.mark.parametrize("config,spp", [("C3", None), ("C4", None), ("C5", 12)])
def test_full_size_properties_other_configs(config, spp):
    """BASELINE.json's other configs at their full frame sizes (C3, C4 at their full spp; the 100 k-sphere C5, whose tree
    is read from L2, at 12 spp): determinism, the early-out accounting, the accumulator invariants, and the two-rank
    sample split summing to the single-rank frame bit for bit."""
    import torch
    from ray_tracing_fsharp_b200.distributed import DeviceBackend
    spec = sample_images.CONFIGS[config]()
    if spp:
        spec.spp = spp
    hs, ts, keep = marshal(spec.objects)
    dsc = native.SceneHandle(hs, ts, 0, keepalive=keep)
    cam = oracle_camera(spec)
    mw, mh = spec.max_width_coord, spec.max_height_coord
    a, sa, sta = dsc.render(cam, mw, mh, seed=3, want_sums=True)
    b, sb, stb = dsc.render(cam, mw, mh, seed=3, want_sums=True)
    assert a.shape == (spec.rows, spec.cols, 3) and (dsc.shared_memory_bytes() == 0) == (config == "C5")
    assert np.array_equal(sa, sb) and int(sta.rays) == int(stb.rays)
    assert set(np.unique(sa[..., 3]).tolist()) <= {11, spec.spp}
    assert int(sta.paths) == int(sa[..., 3].sum()) and int(sta.pixels_early_out) == int((sa[..., 3] == 11).sum())
    assert (sa[..., :3] <= 255 * sa[..., 3:4]).all() and (sa[..., :3] >= 0).all()
    assert np.array_equal(a, (sa[..., :3] // sa[..., 3:4]).astype(np.uint8))
    backends = [DeviceBackend(dsc, cam, mw, mh, seed=3, adaptive=True) for _ in range(2)]
    bufs = [be.alloc() for be in backends]
    for r in range(2):
        backends[r].probe(r, 2, *bufs[r])
    flags = torch.maximum(bufs[0][1], bufs[1][1])  # all_reduce(MAX)
    for r in range(2):
        bufs[r][1].copy_(flags)
        backends[r].main(r, 2, *bufs[r])
    total = (bufs[0][0] + bufs[1][0]).cpu().numpy().reshape(sa.shape)  # all_reduce(SUM)
    assert np.array_equal(total, sa)


def test_errors_are_codes_not_crashes():
    lib = native.lib()
    cam = Camera.make_basic(4, 1.0, 1.0, (0.0, 0.0, 0.0), (0.0, 0.0, 1.0), (0.0, 1.0, 0.0))
    dsc = native.SceneHandle([], [], 0)
    opts = abi.RtRenderOpts(0, 1, 0, 0, 0)
    out = np.zeros((3, 3, 3), np.uint8)
    assert lib.rt_render(dsc.ptr, C.byref(cam), 0, 1, C.byref(opts), native.ptr(out), None, None) == abi.RT_ERR_INVALID_ARGUMENT
    assert lib.rt_render(dsc.ptr, C.byref(cam), 1, 1, C.byref(opts), None, None, None) == abi.RT_ERR_INVALID_ARGUMENT
    opts.mode = 7
    assert lib.rt_render(dsc.ptr, C.byref(cam), 1, 1, C.byref(opts), native.ptr(out), None, None) == abi.RT_ERR_INVALID_ARGUMENT
    assert b"mode" in lib.rt_last_error()
    with pytest.raises(native.RtError):
        native.SceneHandle([], [], 99)


def test_edge_sizes_and_budgets():
    """Smallest frame (3x3), one sample, zero bounce budget, a frame that is not a multiple of the 8x4 tile, and a
    frame with fewer pixels than one warp — each against the oracle with the shared RNG."""
    spec = _small("C1", 1, 1, 1)
    osc, dsc, cam = scene_pair(spec)
    for max_w, max_h, spp, depth, adaptive in [(1, 1, 1, 50, True), (1, 1, 3, 0, True), (5, 3, 12, 2, True), (13, 2, 7, 50, False),
                                               (2, 9, 16, 1, True)]:
        cam.samples_per_pixel, cam.bounce_depth = spp, depth
        ref, ref_stats, counters, _ = osc.render(cam, max_w, max_h, seed=3, rng_mode=1, adaptive=adaptive)
        rgb, sums, stats = dsc.render(cam, max_w, max_h, seed=3, adaptive=adaptive, want_sums=True)
        assert rgb.shape == (2 * max_h + 1, 2 * max_w + 1, 3)
        assert np.array_equal(sums[..., 3], ref_stats[..., 3])
        assert (sums == ref_stats).all(2).mean() > 0.95 and np.abs(rgb.astype(int) - ref.astype(int)).max() <= 40
        assert int(stats.paths) == counters["paths"]
        b, sb, _ = dsc.render(cam, max_w, max_h, seed=3, adaptive=adaptive, want_sums=True, mode=abi.RT_MODE_WAVEFRONT)
        assert np.array_equal(sb, sums)
    cam.bounce_depth = 0  # maxCount = 0: one interaction; anything that does not hit an emitter straight away is HotPink
    cam.samples_per_pixel = 4
    rgb, sums, stats = dsc.render(cam, 10, 6, seed=1, adaptive=False, want_sums=True)
    assert int(stats.rays) == int(stats.paths)
    hot = (rgb == np.array([205, 105, 180], np.uint8)).all(2)
    assert hot.any() and not hot.all()


@pytest.mark.parametrize("name", ["C1", "C2", "C3", "C4"])
def test_frames_against_committed_golden_fixtures(name):
    """tests/golden/oracle_frames.npz: frames the oracle rendered with the shared counter RNG when the fixtures were
    made (it must still reproduce them bit for bit: test_oracle_reference_kats.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_frames.npz"))
    max_w, max_h, spp = [int(x) for x in g[f"{name}_shape"]]
    spec = _small(name, max_w, max_h, spp)
    osc, dsc, cam = scene_pair(spec)
    rgb, sums, stats = dsc.render(cam, max_w, max_h, seed=2024, adaptive=True, want_sums=True)
    want_rgb, want_stats = g[f"{name}_rgb"], g[f"{name}_stats"]
    assert (rgb == want_rgb).all(2).mean() > 0.985
    assert (sums == want_stats).all(2).mean() > 0.97
    assert abs(int(stats.paths) - int(g[f"{name}_work"][0])) <= 0.01 * int(g[f"{name}_work"][0])


def test_large_sample_counts_keep_every_sample():
    """The work table cuts a rank's samples into chunks; every sample index must be traced exactly once whatever
    spp is (here a prime well above the table's bulk size, and the 4096 of BASELINE's C5)."""
    spec = _small("C1", 6, 4, 2999)
    osc, dsc, cam = scene_pair(spec)
    for spp in (2999, 4096):
        cam.samples_per_pixel = spp
        _, sums, st = dsc.render(cam, 6, 4, seed=5, adaptive=False, want_sums=True)
        assert (sums[..., 3] == spp).all() and int(st.paths) == spp * sums.shape[0] * sums.shape[1]
        _, sums_a, st_a = dsc.render(cam, 6, 4, seed=5, adaptive=True, want_sums=True)
        assert set(np.unique(sums_a[..., 3]).tolist()) <= {11, spp}
    cam.samples_per_pixel = 2999
    ref, ref_stats, _, _ = osc.render(cam, 6, 4, seed=5, rng_mode=1, adaptive=False)
    _, sums, _ = dsc.render(cam, 6, 4, seed=5, adaptive=False, want_sums=True)
    assert np.abs(sums[..., :3] - ref_stats[..., :3]).max() <= 0.002 * 255 * 2999  # a few paths of 2999 may differ (FP32)


@pytest.mark.parametrize("config,max_w,max_h", [("C2", 90, 60), ("C3", 120, 67), ("C4", 80, 45), ("C5", 44, 25)])
def test_4096_spp_with_the_reference_early_out_against_the_oracle_render(config, max_w, max_h):
    """BASELINE's level-2 bar for configs[1..4] with the reference's rule ON (renderPixel's early-out, Scene.fs:157-194),
    on crops (reduced half-extents, same camera and scene) sized so that the oracle finishes in well under a minute:
    C2 181x121, C3 241x135 (the earth-map texture path), C4 161x91 (planes + glass, depth 100), C5 89x51 (100 000
    spheres, BVH read from L2).  Oracle: the reference's xorshift128 stream; GPU: Philox — independent streams.
    Per channel, MAD of the pre-gamma means < 3 sigma-bar (sigma-bar from two GPU seeds); the GPU differs from the oracle
    no more than from itself under another seed; the share of early-out pixels agrees; P3 bytes are diffed."""
    spec = _small(config, max_w, max_h, 4096)
    osc, dsc, cam = scene_pair(spec)
    ref, ref_stats, counters, _ = osc.render(cam, max_w, max_h, seed=4321, rng_mode=0, adaptive=True)
    a, sa, _ = dsc.render(cam, max_w, max_h, seed=11, adaptive=True, want_sums=True)
    b, sb, _ = dsc.render(cam, max_w, max_h, seed=12, adaptive=True, want_sums=True)
    assert set(np.unique(sa[..., 3]).tolist()) <= {11, 4096} and set(np.unique(ref_stats[..., 3]).tolist()) <= {11, 4096}
    mean_a, mean_b, mean_o = sa[..., :3] / sa[..., 3:4], sb[..., :3] / sb[..., 3:4], ref_stats[..., :3] / ref_stats[..., 3:4]
    sigma_bar = np.sqrt(((mean_a - mean_b) ** 2).mean(axis=(0, 1)))
    mad = np.abs(mean_a - mean_o).mean(axis=(0, 1))
    mad_gpu = np.abs(mean_a - mean_b).mean(axis=(0, 1))
    frac_o, frac_g = (ref_stats[..., 3] == 11).mean(), (sa[..., 3] == 11).mean()
    ppm_g = np.array(ImageOutput.write_ppm(True, a).split()[4:], dtype=np.int32)
    ppm_o = np.array(oracle.ppm_format(ref, True).split()[4:], dtype=np.int32)
    diff = np.abs(ppm_g - ppm_o)
    print(f"{config} {2 * max_w + 1}x{2 * max_h + 1} @4096 spp adaptive: MAD {np.round(mad, 3)} MAD(gpu,gpu') {np.round(mad_gpu, 3)} sigma-bar "
          f"{np.round(sigma_bar, 3)} early-out oracle {frac_o:.4f} gpu {frac_g:.4f} P3 |diff| <=1: {np.mean(diff <= 1):.3f} mean {diff.mean():.3f}")
    assert (mad < 3.0 * sigma_bar + 0.25).all(), (mad, sigma_bar)
    assert (mad < 1.5 * mad_gpu + 0.25).all(), (mad, mad_gpu)
    assert abs(frac_o - frac_g) < 0.03, (frac_o, frac_g)
    assert diff.size == 3 * (2 * max_w + 1) * (2 * max_h + 1)
    assert np.mean(diff <= 3) > 0.8 and np.mean(diff) < 3.0, (np.mean(diff <= 3), np.mean(diff))


def test_c1_full_size_4096_spp_against_the_oracle_render():
    """BASELINE's level-2 bar on configs[0] at its real size (401x225), 4096 spp, adaptive early-out on both sides as
    the reference renders: the oracle with the reference's xorshift128 generator, the GPU with Philox — independent
    streams.  Per channel, the mean absolute difference of the pre-gamma means must be under 3 sigma-bar, sigma-bar
    being the RMS over pixels of the standard error of the difference of two independent estimates (taken from two
    GPU renders with different seeds); the gamma-corrected P3 bytes are diffed."""
    spec = sample_images.few_spheres()
    spec.spp = 4096
    osc, dsc, cam = scene_pair(spec)
    mw, mh = spec.max_width_coord, spec.max_height_coord
    ref, ref_stats, counters, _ = osc.render(cam, mw, mh, seed=1234, rng_mode=0, adaptive=True)
    a, sa, _ = dsc.render(cam, mw, mh, seed=1, adaptive=True, want_sums=True)
    b, sb, _ = dsc.render(cam, mw, mh, seed=2, adaptive=True, want_sums=True)
    mean_a = sa[..., :3] / sa[..., 3:4]
    mean_b = sb[..., :3] / sb[..., 3:4]
    mean_o = ref_stats[..., :3] / ref_stats[..., 3:4]
    var_diff = (mean_a - mean_b) ** 2                       # E = Var(a) + Var(b) = variance of a difference of two estimates
    sigma_bar = np.sqrt(var_diff.mean(axis=(0, 1)))
    mad = np.abs(mean_a - mean_o).mean(axis=(0, 1))
    assert (mad < 3.0 * sigma_bar + 0.25).all(), (mad, sigma_bar)
    # and the two GPU renders are as far from each other as each is from the oracle (same distribution)
    mad_gpu = np.abs(mean_a - mean_b).mean(axis=(0, 1))
    assert (mad < 1.5 * mad_gpu + 0.25).all(), (mad, mad_gpu)
    # early-out behaves alike: the fraction of pixels that stop after 11 samples
    frac_o, frac_g = (ref_stats[..., 3] == 11).mean(), (sa[..., 3] == 11).mean()
    assert abs(frac_o - frac_g) < 0.02, (frac_o, frac_g)
    # P3 bytes (gamma-corrected), GPU writer vs oracle writer
    ppm_g = np.array(ImageOutput.write_ppm(True, a).split()[4:], dtype=np.int32)
    ppm_o = np.array(oracle.ppm_format(ref, True).split()[4:], dtype=np.int32)
    diff = np.abs(ppm_g - ppm_o)
    assert diff.size == 3 * 401 * 225
    assert np.mean(diff <= 2) > 0.9 and np.mean(diff) < 1.5, (np.mean(diff <= 2), np.mean(diff), np.percentile(diff, 99))
