#!/usr/bin/env bash
# bench.py under torchrun at the GPU counts given as arguments (run under gpurun --gpus N)
for n in "$@"; do
  if [ "$n" = "1" ]; then timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/r02_scale_n$n.log 2> gpurun_out/r02_scale_n$n.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r02_scale_n$n.log 2> gpurun_out/r02_scale_n$n.err; fi
  tail -c 600 gpurun_out/r02_scale_n$n.err | grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$"
  python -c "
import json; d=json.loads(open('gpurun_out/r02_scale_n$n.log').read().strip().splitlines()[-1]); print('N', d['n_gpus'], 'ms', round(d['ms_per_step'],3), 'Mrays/s', round(d['value']), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'e2e Mrays/s', round(d['e2e']['value']), 'hash', d['frame_sha256']['rgb8'][:12], d['frame_sha256']['all_paths_equal'], d['frame_sha256'].get('rt_multi_render',{}).get('wall_ms'), {k:(round(v['ms_per_step'],2), round(v['mrays_per_s'])) for k,v in d['other_configs'].items()})"
done
