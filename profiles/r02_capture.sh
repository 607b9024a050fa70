#!/usr/bin/env bash
# Round-2 evidence run (one GPU, under gpurun): the GPU test suite, the default bench line, the ncu launch list of a short
# bench run and one `--set full` capture of the dominant kernel (main-phase render_kernel, C2).
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02_pytest_gpu.log; cat gpurun_out/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
python bench.py > gpurun_out/r02_bench_n1.log 2> gpurun_out/r02_bench_n1.err; tail -c 400 gpurun_out/r02_bench_n1.err; head -c 400 gpurun_out/r02_bench_n1.log; echo
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-configs --no-hash"
$CMD > gpurun_out/r02_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 5 -c 1 -o gpurun_out/prof_r02_main $CMD > gpurun_out/r02_ncu_full.log 2>&1
tail -2 gpurun_out/r02_ncu_full.log; ls -la gpurun_out/prof_r02_main.ncu-rep gpurun_out/r02_launches.csv
