#!/usr/bin/env bash
# ncu --set full with source counters of ONE launch of the main-phase kernel of a schedule (C2 at 60 spp)
#   usage: ncu_flow_source.sh <mode: megakernel|lockstep> <kernel regex> <output name>
mode=$1; regex=$2; out=$3
python profiles/compare_modes.py --config C2 --spp 60 --modes $mode > gpurun_out/${out}_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:$regex -s 1 -c 1 -o gpurun_out/$out python profiles/compare_modes.py --config C2 --spp 60 --modes $mode > gpurun_out/${out}_ncu.log 2>&1
tail -2 gpurun_out/${out}_ncu.log
ls -la gpurun_out/$out.ncu-rep
