#!/usr/bin/env bash
# quick check of a kernel change (run under gpurun): parity tests that exercise the walk, then C2 full size, C4 and the 100 k-sphere scene at 32 spp
timeout 900 python -m pytest tests -m gpu -x -q -k "sphere or plane or reflection or hit_object or frame_matches or shared_rng or bit_for_bit or trace_samples or sample_scene" 2>&1 | tail -4
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --other-configs C3,C4 --no-hash > gpurun_out/ab_c2.log 2>gpurun_out/ab_c2.err; tail -c 300 gpurun_out/ab_c2.err
python -c "
import json; d=json.loads(open('gpurun_out/ab_c2.log').read().strip().splitlines()[-1]); print('C2', d['ms_per_step'], d['value'], 'main', d['roofline']['kernel_ms_per_launch'], 'e2e', d['e2e']['ms_per_step'], {k:round(v['ms_per_step'],2) for k,v in d['other_configs'].items()}, d['traversal'])"
timeout 300 python bench.py --config C5 --spp 32 --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs --no-hash > gpurun_out/ab_c5.log 2>gpurun_out/ab_c5.err; tail -c 300 gpurun_out/ab_c5.err
python -c "
import json; d=json.loads(open('gpurun_out/ab_c5.log').read().strip().splitlines()[-1]); print('C5/32spp', d['ms_per_step'], d['value'], 'main', d['roofline']['kernel_ms_per_launch'], d['traversal'])"
