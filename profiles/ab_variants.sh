#!/usr/bin/env bash
# A/B of library builds (run under gpurun): every ray_tracing_fsharp_b200/variants/*.so in turn takes the place of
# librtfs_b200.so for one short C2 bench (with the frame hash: a variant must reproduce it) and the 100 k-sphere scene at 32 spp;
# the default build runs first and last.  Variants are built by hand with -DRTFS_... (see rtfs_device.h) and are not committed.
L=ray_tracing_fsharp_b200/librtfs_b200.so
cp $L /tmp/default.so
run() {
    timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --other-configs ${OTHER:-C4} > gpurun_out/abv_$1.log 2> gpurun_out/abv_$1.err || tail -c 300 gpurun_out/abv_$1.err
    python -c "
import json; d=json.loads(open('gpurun_out/abv_$1.log').read().strip().splitlines()[-1]); print('$1', 'C2', round(d['ms_per_step'],3), 'main', round(d['roofline']['kernel_ms_per_launch'],3), 'e2e', round(d['e2e']['ms_per_step'],3), {k:round(v['ms_per_step'],2) for k,v in d['other_configs'].items()}, str(d['frame_sha256'])[:80])"
    if [ -n "$WITH_C5" ]; then
    timeout 300 python bench.py --config C5 --spp 32 --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs --no-hash > gpurun_out/abv_$1_c5.log 2>gpurun_out/abv_$1_c5.err
    python -c "
import json; d=json.loads(open('gpurun_out/abv_$1_c5.log').read().strip().splitlines()[-1]); print('$1', 'C5/32spp', round(d['ms_per_step'],2), 'main', round(d['roofline']['kernel_ms_per_launch'],2))"
    fi
}
run default
for v in ray_tracing_fsharp_b200/variants/*.so; do
    cp $v $L
    run $(basename $v .so)
done
cp /tmp/default.so $L
run default_again
