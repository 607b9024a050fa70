#!/usr/bin/env bash
# Runs BOTH arms of the comparison inside the reference itself, on a machine that has the .NET SDK (7 or 8) and a B200
# (or any sm_100a GPU) with librtfs_b200.so built:
#   CPU arm  the UNMODIFIED renderer of Smaug123/ray-tracing-fsharp (Scene.make |> Scene.render)
#   GPU arm  the same sample functions with Scene.make / Scene.render swapped for GpuScene.make / SceneGpu.render
#            (RayTracing/Gpu.fs = shim/RayTracing.Gpu.fs of this repository, P/Invoke into librtfs_b200.so)
# and compares the two images statistically (compare_arms.py).
#
# UNVERIFIED in this repository's build image, which has no .NET: every CPU number bench.py reports is the C++
# restatement of the F# algorithm under oracle/, labelled as such.  This script is the route to the real drop-in and to the
# real F# baseline the north star names; fill expected_results.md with what it prints.
#
#   usage: time_reference.sh /path/to/ray-tracing-fsharp /path/to/this/repo [sample] [baseline|native]
#     sample    random-spheres (default; the RTOW final scene, C2) or earth (C3's texture path)
#     baseline  BASELINE.json's C2 configuration: pixels = 400 -> 1201x801, BounceDepth = 50        (default)
#     native    the sample as its author wrote it: pixels = 800 -> 2401x1601, BounceDepth = 150 (Camera.fs:58)
set -euo pipefail
ref="${1:?path to a checkout of Smaug123/ray-tracing-fsharp}"
here="${2:?path to this repository (for shim/, baseline/fsharp/ and librtfs_b200.so)}"
sample="${3:-random-spheres}"
config="${4:-baseline}"
work="$(mktemp -d)"
cp -r "$ref" "$work/src"
cd "$work/src"
# 1. the binding: one new file in the RayTracing assembly + the patch (fsproj compile order, two call sites per sample)
cp "$here/shim/RayTracing.Gpu.fs" RayTracing/Gpu.fs
patch -p1 < "$here/baseline/fsharp/gpu-backend.patch"
if [ "$config" = "baseline" ] && [ "$sample" = "random-spheres" ]; then
    # BASELINE.json configs[1]: 1201x801, depth 50 (the author's sample is 2401x1601 at Camera.makeBasic's depth 150)
    sed -i '827s/let pixels = 800/let pixels = 400/' RayTracing.App/SampleImages.fs
    sed -i '/^    let randomSpheres/,/^    let earth/ s/^\(        let pixels = 400\)$/\1\n        let camera = { camera with BounceDepth = 50 }/' RayTracing.App/SampleImages.fs
fi
dotnet build -c Release RayTracing.App >/dev/null
export LD_LIBRARY_PATH="$here/ray_tracing_fsharp_b200:${LD_LIBRARY_PATH:-}"
ln -sf "$here/ray_tracing_fsharp_b200/librtfs_b200.so" "$work/src/librtfs_b200.so"
echo "host cores: $(nproc)"
echo "== CPU arm (reference renderer, unmodified code path) =="
/usr/bin/time -f "cpu_arm_wall_s %e" dotnet run -c Release --no-build --project RayTracing.App -- "$sample" "$work/cpu.png" 2>&1 | tail -3
echo "== GPU arm (same sample, Scene.make / Scene.render -> GpuScene.make / SceneGpu.render) =="
RTFS_GPU=1 /usr/bin/time -f "gpu_arm_wall_s %e" dotnet run -c Release --no-build --project RayTracing.App -- "$sample" "$work/gpu.png" 2>&1 | tail -3
RTFS_GPU=1 /usr/bin/time -f "gpu_arm_second_run_wall_s %e" dotnet run -c Release --no-build --project RayTracing.App -- "$sample" "$work/gpu2.png" 2>&1 | tail -1
python3 "$here/baseline/fsharp/compare_arms.py" "$work/cpu.png" "$work/gpu.png" "$work/gpu2.png"
echo "images kept in $work"
