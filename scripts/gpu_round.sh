#!/bin/bash
# Full GPU round: parity tests, bench (strict / fast), ncu launch list. Every stage has its own timeout.
mkdir -p gpurun_out
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_gpu.log
for e in strict fast; do
echo "== bench $e" ; timeout 600 python bench.py --engine $e --steps 2 --warmup 3 > gpurun_out/bench_$e.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_$e.log
done
echo "== ncu launch list (strict)"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_strict.csv python bench.py --engine strict --steps 1 --warmup 1 --no-cpu-baseline --steps-per-graph 1 --no-graph > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_bench.log
