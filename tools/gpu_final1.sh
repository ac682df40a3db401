#!/bin/bash
# round 2 final, 1 GPU: test tier, default bench line, other workloads, ncu launch list + full capture
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
FOLP_TIMING=1 timeout 900 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_c2.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_c2_reference.json 2> gpurun_out/bench_c2_reference.err
echo "reference rc=$?"; cut -c1-300 gpurun_out/bench_c2_reference.json
for w in netlib_small netlib pagerank; do
  timeout 600 python bench.py --workload $w --cpu-iters 80 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  echo "bench $w rc=$?"; cut -c1-200 gpurun_out/bench_$w.json
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 600 --csv \
  --log-file gpurun_out/launches_c2.csv python bench.py --steps 1 --warmup 3 --skip-e2e --skip-cpu --skip-target \
  > gpurun_out/ncu_launches.log 2>&1
echo "ncu launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_spmv|k_primal|k_tr_multi" -s 2000 -c 7 \
  -f -o gpurun_out/full_c2 python bench.py --steps 1 --warmup 3 --skip-e2e --skip-cpu --skip-target \
  > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/full_c2.ncu-rep
