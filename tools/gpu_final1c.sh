#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
bash tools/gpu_final1b.sh
grep -E "folp_create\]|folp host" gpurun_out/bench_c2.err | sed -n 8,14p
