#!/bin/bash
# ncu: launch list of a short bench run + one full capture of the headline fit kernel and of the wave kernel (1 GPU).
# Usage (repo root, under gpurun): bash tools/gpu_profile.sh <tag>
TAG=${1:-prof}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fit_ -s 3 -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu2_$TAG.log 2>&1
B200LM_WAVE_CFG=1 ncu --set full --clock-control none --import-source on -k regex:fit_wave -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_wave \
    python tools/wave_prof.py 160000 > gpurun_out/prof_${TAG}_wave.log 2>&1
ls -la gpurun_out/ | grep $TAG
