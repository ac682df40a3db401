#!/bin/bash
# round 2: full GPU test tier on an N-GPU box (includes the world-size-2 partitioned parity worker), then the
# partitioned probe and bench line
N=${1:-2}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_x$N.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_x$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
B=firstorderlp.jl_b200/libfolp_b200.so
for w in ${WORKLOADS:-c2}; do
  timeout 600 $TR tools/probe_kernels.py --workload $w --iters 2000 $B $B:FOLP_EVAL_SYNC=1 > gpurun_out/probe_${w}_x$N.log 2> gpurun_out/probe_${w}_x$N.err
  echo "probe $w x$N rc=$?"; grep '^{' gpurun_out/probe_${w}_x$N.log | grep '"rank": 0' | cut -c1-400
done
FOLP_TIMING=1 timeout 900 $TR bench.py --gpus $N > gpurun_out/bench_c2_x$N.json 2> gpurun_out/bench_c2_x$N.err
echo "bench x$N rc=$?"; cut -c1-600 gpurun_out/bench_c2_x$N.json; grep -E "PARITY|folp_create\] (TOTAL|device)" gpurun_out/bench_c2_x$N.err | head
python - $N <<'PY'
import json, sys
d = json.load(open("gpurun_out/bench_c2_x%s.json" % sys.argv[1]))
print("value", d["value"], "e2e", d["e2e"], "parity", d["detail"]["parity"], "create", d["detail"]["folp_create_seconds"], "pure", d["detail"]["pure_step_iterations_per_s"])
PY
