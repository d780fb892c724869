#!/bin/bash
# ncu evidence for profiles/ (B200_PROFILING.md recipe): (1) launch list of a short eager chain, (2) --set full
# captures of the two dominant kernels at the C2 shape.  usage: bash scripts/gpu_profile.sh <engine> <tag>
E=${1:-strict}; TAG=${2:-r2}
mkdir -p gpurun_out
echo "== ncu launch list ($E)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
  --log-file gpurun_out/launches_${E}_${TAG}.csv \
  python bench.py --engine $E --steps 1 --warmup 1 --no-cpu-baseline --steps-per-graph 1 --no-graph --timesteps 12 \
  > gpurun_out/ncu_bench_${E}.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/ncu_bench_${E}.log | cut -c1-200
echo "== ncu full: edge kernel ($E)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:edge_tc_k --launch-skip 24 -c 2 \
  -o gpurun_out/edge_${E}_${TAG} -f python scripts/profile_forward.py $E 4 > gpurun_out/ncu_edge_${E}.log 2>&1; echo "rc=$?"
echo "== ncu full: node GEMM ($E)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:linear_tc_k --launch-skip 56 -c 3 \
  -o gpurun_out/node_${E}_${TAG} -f python scripts/profile_forward.py $E 4 > gpurun_out/ncu_node_${E}.log 2>&1; echo "rc=$?"
