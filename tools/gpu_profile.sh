#!/bin/bash
# ncu: one full capture of the fit kernel (1 GPU) + optional launch list.  Usage: bash tools/gpu_profile.sh <tag> [list]
TAG=${1:-prof}
mkdir -p gpurun_out
if [ "$2" == "list" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_$TAG.log 2>&1
fi
ncu --set full --clock-control none --import-source on -k regex:fit_kernel -s 3 -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu2_$TAG.log 2>&1
ls -la gpurun_out/ | grep $TAG
