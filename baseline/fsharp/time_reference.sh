#!/usr/bin/env bash
# Times the UNMODIFIED reference renderer (Smaug123/ray-tracing-fsharp) on a machine that has the .NET 8 SDK.
# UNVERIFIED in this repository's build image, which has no .NET: every CPU number reported by bench.py is the
# C++ restatement of the F# algorithm under oracle/, labelled as such.
#
#   usage: time_reference.sh /path/to/ray-tracing-fsharp [sample] [output.png]
# `random-spheres` is the RTOW final scene; as written by its author it renders 2401x1601 at 500 spp, depth 150
# (RayTracing.App/SampleImages.fs:812-960; RayTracing/Camera.fs:58).  To time BASELINE.json's 1201x801 / depth 50
# configuration, change `pixels = 800` to 400 at SampleImages.fs:827 and add `BounceDepth = 50` to the camera.
set -euo pipefail
repo="${1:?path to a checkout of Smaug123/ray-tracing-fsharp}"
sample="${2:-random-spheres}"
out="${3:-/tmp/reference-${sample}.png}"
cd "$repo"
dotnet build -c Release RayTracing.App >/dev/null
echo "cores: $(nproc)"
/usr/bin/time -v dotnet run -c Release --no-build --project RayTracing.App -- "$sample" "$out"
