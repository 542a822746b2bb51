#!/bin/bash
# ncu: launch list of a short bench run + one full capture of the headline fit kernel and of the wave kernel (1 GPU).
# The reports are summarised ON the GPU box (stall breakdown, per-function and per-line samples) and deleted: only the
# text comes back (gpurun returns at most 64 MiB).   Usage (repo root, under gpurun): bash tools/gpu_profile.sh <tag>
TAG=${1:-prof}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fit_ -s 3 -c 1 -f -o /tmp/prof_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu2_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fit_wave -s 1 -c 1 -f -o /tmp/prof_${TAG}_wave \
    python tools/wave_prof.py 160000 > gpurun_out/prof_${TAG}_wave.log 2>&1
for t in $TAG ${TAG}_wave; do
  { python tools/ncu_stalls.py /tmp/prof_$t.ncu-rep; python tools/ncu_funcs.py /tmp/prof_$t.ncu-rep 2>/dev/null | head -25; python tools/ncu_lines.py /tmp/prof_$t.ncu-rep 25; } > gpurun_out/ncu_$t.txt 2>&1
done
ls -la gpurun_out/ | grep $TAG
