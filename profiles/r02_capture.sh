#!/usr/bin/env bash
# Round-2 evidence run (one GPU, under gpurun): the GPU test suite, smoke(), the default bench line, the reference arm, the ncu
# launch list of a short bench run and one `--set full` capture of the dominant kernel (main-phase render_kernel, C2).
# TAG names the outputs (default r02): gpurun_out/${TAG}_*.
TAG=${TAG:-r02}
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
python bench.py > gpurun_out/${TAG}_bench_n1.log 2> gpurun_out/${TAG}_bench_n1.err; tail -c 400 gpurun_out/${TAG}_bench_n1.err; head -c 400 gpurun_out/${TAG}_bench_n1.log; echo
if [ -n "$WITH_REFERENCE" ]; then python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.log 2> gpurun_out/${TAG}_bench_reference.err; head -c 600 gpurun_out/${TAG}_bench_reference.log; echo; fi
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-configs --no-hash"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 5 -c 1 -o gpurun_out/prof_${TAG}_main $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log; ls -la gpurun_out/prof_${TAG}_main.ncu-rep gpurun_out/${TAG}_launches.csv
