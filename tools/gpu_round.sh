#!/bin/bash
# One gpurun call of the development loop: GPU parity tests, kernel probes of the built
# variants, the default bench line and an ncu launch list. Outputs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
FOLP_TIMING=1 timeout 400 python tools/probe_kernels.py --workload c2 --iters 2000 \
  firstorderlp.jl_b200/libfolp_b200.so firstorderlp.jl_b200/libfolp_b200.so:FOLP_NO_ROW_SORT=1 \
  firstorderlp.jl_b200/libfolp_b200.so:FOLP_TR_MULTIKERNEL=1 \
  $(ls scratch/*.so 2>/dev/null) > gpurun_out/probe_c2.log 2>&1
grep -E '^\{|folp_create\] (TOTAL|pack|transpose|index|vectors)|folp_destroy' gpurun_out/probe_c2.log | cut -c1-600
python tools/host_prepare_probe.py 2>/dev/null | tee gpurun_out/host_prepare.log
FOLP_NO_THP=1 python tools/host_prepare_probe.py 2>/dev/null | tee -a gpurun_out/host_prepare.log
FOLP_TIMING=1 timeout 500 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
echo "bench rc=$?"
cut -c1-1500 gpurun_out/bench_c2.json
grep -E 'folp_create\]|folp_destroy' gpurun_out/bench_c2.err | tail -24
