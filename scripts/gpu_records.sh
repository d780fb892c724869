#!/bin/bash
# Reproducible records for profiles/: bench lines of every workload + ncu evidence.  usage: bash scripts/gpu_records.sh [tag]
TAG=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total,driver_version --format=csv > gpurun_out/gpu_$TAG.txt 2>&1
echo "== bench configs1 strict (with cpu_baseline + gpu_reference)"; timeout 900 python bench.py --engine strict --steps 3 --warmup 3 > gpurun_out/bench_configs1_strict_$TAG.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_configs1_strict_$TAG.log | cut -c1-300
echo "== bench configs1 fast"; timeout 600 python bench.py --engine fast --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_configs1_fast_$TAG.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_configs1_fast_$TAG.log | cut -c1-200
echo "== bench geom9 strict"; timeout 900 python bench.py --engine strict --workload geom9 --steps 3 --warmup 3 --cpu-steps 3 > gpurun_out/bench_geom9_strict_$TAG.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_geom9_strict_$TAG.log | cut -c1-300
echo "== bench sweep strict"; timeout 1200 python bench.py --engine strict --workload sweep --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_sweep_strict_$TAG.log 2>&1; echo "rc=$?"; grep -c metric gpurun_out/bench_sweep_strict_$TAG.log
echo "== bench configs1 geom sizes"; timeout 600 python bench.py --engine strict --sizes geom --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_configs1_geomsizes_strict_$TAG.log 2>&1; echo "rc=$?"
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_$TAG.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_reference_$TAG.log | cut -c1-200
bash scripts/gpu_profile.sh strict $TAG
