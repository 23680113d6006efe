#!/bin/bash
# full ncu capture of one kernel: scripts/gpu_prof.sh <kernel-regex> <skip> <pairs> <name> [extra bench args]
K=$1; S=$2; P=$3; N=$4; shift 4
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -o gpurun_out/$N -f python bench.py --pairs $P --steps 1 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/$N.log 2>&1
ls -la gpurun_out/$N.ncu-rep
