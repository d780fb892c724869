#!/bin/bash
# What the kernel families cost INSIDE the graph-replayed step (PDL overlap included): builds of the library that skip
# the node-GEMM launches / the edge-kernel launches (wrong numerics, timing only) against the product build.
# usage (build container): bash scripts/step_ablation.sh build ; (GPU box): bash scripts/step_ablation.sh run
set -e
SRC="hierdiff_b200/csrc"; FILES="$SRC/hd_layout.cu $SRC/hd_fp32.cu $SRC/hd_tc.cu $SRC/hd_node.cu $SRC/hd_api.cu $SRC/hd_egcl.cu $SRC/hd_loss.cu"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -shared -Xcompiler -fvisibility=hidden"
if [ "$1" = build ]; then
  mkdir -p build
  nvcc $FLAGS -DHD_EXP_SKIP_NODE -o build/libhd_skip_node.so $FILES &
  nvcc $FLAGS -DHD_EXP_SKIP_EDGE -o build/libhd_skip_edge.so $FILES &
  wait; ls -la build/libhd_skip_*.so
else
  for v in "" build/libhd_skip_node.so build/libhd_skip_edge.so; do
    echo "== ${v:-product}"
    HD_LIB_PATH=${v:+$PWD/$v} python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-reference 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('ms per reverse step', d['ms_per_step']/1.0, 'launches/step', d['launches_per_reverse_step'])"
  done
fi
