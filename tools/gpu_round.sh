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
FOLP_THP=1 python tools/host_prepare_probe.py 2>/dev/null | tee -a gpurun_out/host_prepare.log
FOLP_TIMING=1 timeout 500 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
echo "bench rc=$?"
cut -c1-1500 gpurun_out/bench_c2.json
grep -E 'folp_create\]|folp_destroy' gpurun_out/bench_c2.err | tail -24
if [ -n "$NCU" ]; then
  # launch list of the default bench workload (cold-cache, serialised per-launch times: shares only)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 700 --csv \
    --log-file gpurun_out/launches_c2.csv python bench.py --steps 1 --warmup 3 --skip-e2e --skip-cpu \
    > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launch list rc=$?"
  # full capture of the three per-attempt kernels, two launches each
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_spmv|k_primal" -s 3000 -c 6 \
    -f -o gpurun_out/full_c2 python bench.py --steps 1 --warmup 3 --skip-e2e --skip-cpu \
    > gpurun_out/ncu_full.log 2>&1
  echo "ncu full rc=$?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"EpiPlain" -s 10 -c 2 \
    -f -o gpurun_out/full_plain_c2 python bench.py --steps 1 --warmup 3 --skip-e2e --skip-cpu \
    > gpurun_out/ncu_full_plain.log 2>&1
  echo "ncu plain rc=$?"
  ls -la gpurun_out/
fi
