#!/bin/bash
# First GPU bring-up: fp32-engine parity, tensor-core engine bring-up, a short bench. Every stage has its own timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== fp32 parity" ; timeout 600 python -m pytest tests -m gpu -x -q -k "not strict and not fast and not tensor_core" > gpurun_out/pytest_fp32.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_fp32.log
echo "== tc bringup" ; timeout 180 python scripts/tc_debug.py > gpurun_out/tc_debug.log 2>&1; echo "rc=$?"; tail -20 gpurun_out/tc_debug.log
echo "== bench fp32" ; timeout 600 python bench.py --engine fp32 --steps 1 --warmup 1 > gpurun_out/bench_fp32.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/bench_fp32.log
