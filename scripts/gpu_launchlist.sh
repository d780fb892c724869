#!/bin/bash
# ncu launch list (gpu__time_duration per launch) of a short eager chain. usage: gpu_launchlist.sh <engine> <tag>
E=${1:-strict}; TAG=${2:-r1}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 3000 --csv \
  --log-file gpurun_out/launches_${E}_${TAG}.csv \
  python bench.py --engine $E --steps 1 --warmup 1 --no-cpu-baseline --steps-per-graph 1 --no-graph --timesteps 12 \
  > gpurun_out/ncu_bench_${E}.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/ncu_bench_${E}.log | cut -c1-200
