#!/bin/bash
# Full GPU round: parity tests (incl. the same-device reference comparisons), default bench (with gpu_reference and the
# CPU reference leg), ncu launch list. Every stage has its own timeout.  usage: bash scripts/gpu_round.sh [tag]
TAG=${1:-r2}
mkdir -p gpurun_out
rm -f gpurun_out/parity_reference.json
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_gpu_$TAG.log
cat gpurun_out/parity_reference.json gpurun_out/parity_t1000_fixture_*.json 2>/dev/null
echo "== bench strict" ; timeout 900 python bench.py --engine strict --steps 3 --warmup 3 > gpurun_out/bench_strict_$TAG.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_strict_$TAG.log
echo "== bench fast" ; timeout 600 python bench.py --engine fast --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_fast_$TAG.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_fast_$TAG.log | cut -c1-400
