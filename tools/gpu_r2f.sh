#!/bin/bash
# round 2: N-GPU box (N = $1): parity in both launch forms, then the bench line (parity + target sub-run inside)
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tests/dist_worker.py > gpurun_out/dist_worker_x$N.log 2>&1
echo "dist_worker x$N rc=$?"; grep -E "ALL OK|FAILED" gpurun_out/dist_worker_x$N.log | head -5
timeout 600 python tests/dist_worker.py --single-process $N > gpurun_out/multi_worker_x$N.log 2>&1
echo "single-process x$N rc=$?"; grep -E "ALL OK|FAILED" gpurun_out/multi_worker_x$N.log | head -5; tail -3 gpurun_out/multi_worker_x$N.log
FOLP_TIMING=1 timeout 900 $TR bench.py --gpus $N > gpurun_out/bench_c2_x$N.json 2> gpurun_out/bench_c2_x$N.err
echo "bench x$N rc=$?"
python - $N <<'PY'
import json, sys
d = json.load(open("gpurun_out/bench_c2_x%s.json" % sys.argv[1]))
print("value", d["value"], "pure", d["detail"]["pure_step_iterations_per_s"], "create", d["detail"]["folp_create_seconds"])
print("e2e", d["e2e"])
print("parity", {k: v for k, v in d["detail"]["parity"].items() if k != "rule"})
print("target", d["detail"]["target"])
print("phases", d["roofline"]["per_kernel"])
print("clocks", d["clocks"])
PY
grep -E "folp_create\] TOTAL" gpurun_out/bench_c2_x$N.err | head -4
