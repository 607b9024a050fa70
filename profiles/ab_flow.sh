#!/usr/bin/env bash
# A/B of the two schedules (run under gpurun): bit-identity tests, then C2 at full size and the 100 k-sphere scene at 32 spp
timeout 600 python -m pytest tests -m gpu -x -q -k "flow_and_lockstep" 2>&1 | tail -15 > gpurun_out/r2_pytest5.log; cat gpurun_out/r2_pytest5.log
for s in lockstep flow; do
  timeout 300 python bench.py --schedule $s --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r2_flow_c2_$s.log 2>gpurun_out/r2_flow_c2_$s.err; tail -c 300 gpurun_out/r2_flow_c2_$s.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2_flow_c2_$s.log').read().strip().splitlines()[-1]); print('C2 $s', d['ms_per_step'], d['value'], d['roofline']['kernel_ms_per_launch'], d['frame_sha256']['rgb8'][:12], d['frame_sha256']['all_paths_equal'])"
  timeout 300 python bench.py --config C5 --spp 32 --schedule $s --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs --no-hash > gpurun_out/r2_flow_c5_$s.log 2>gpurun_out/r2_flow_c5_$s.err; tail -c 300 gpurun_out/r2_flow_c5_$s.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2_flow_c5_$s.log').read().strip().splitlines()[-1]); print('C5/32spp $s', d['ms_per_step'], d['value'], d['roofline']['kernel_ms_per_launch'])"
done
