#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/pytest_gpu_dist_x$N.log 2>&1
echo "pytest dist rc=$?"; tail -3 gpurun_out/pytest_gpu_dist_x$N.log
FOLP_TIMING=1 timeout 1200 python bench.py --gpus $N --single-process --skip-target --skip-cpu > gpurun_out/bench_c2_x${N}_single_process.json 2> gpurun_out/bench_c2_x${N}_single_process.err
echo "bench single-process x$N rc=$?"
python - $N <<'PY'
import json, sys
d = json.load(open("gpurun_out/bench_c2_x%s_single_process.json" % sys.argv[1]))
dd = d["detail"]
print("single-process value %.0f pure %.0f create %.2f e2e %.0f %s long %s" % (
    d["value"], dd["pure_step_iterations_per_s"], dd["folp_create_seconds"], d["e2e"]["value"],
    [round(x, 3) for x in d["e2e"]["seconds_create_solve_destroy"]], d["e2e"]["long_solve"]))
PY
grep "folp_create\]" gpurun_out/bench_c2_x${N}_single_process.err | tail -14
