#!/bin/bash
# the GPU test tier as the driver runs it (on an N-GPU box the partitioned tests run too)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
