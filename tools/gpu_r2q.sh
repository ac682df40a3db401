#!/bin/bash
# N >= 4 GPUs, defaults (multicast region, cached across handles): partitioned parity under torchrun, bench line
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523"
FOLP_EXPECT_EXCHANGE=multicast timeout 600 $TR tests/dist_worker.py > gpurun_out/dist_worker_mc_x$N.log 2>&1
echo "torchrun default rc=$?"; grep -E "exchange mode|ALL OK|FAILED" gpurun_out/dist_worker_mc_x$N.log | head -4
FOLP_TIMING=1 timeout 1200 $TR bench.py --gpus $N $BENCH_ARGS > gpurun_out/bench_c2_x${N}_mc.json 2> gpurun_out/bench_c2_x${N}_mc.err
echo "bench x$N rc=$?"
python - $N <<'PY'
import json, sys
d = json.load(open("gpurun_out/bench_c2_x%s_mc.json" % sys.argv[1])); dd = d["detail"]
print("value %.0f pure %.0f create %.2f e2e %.0f %s long %s" % (d["value"], dd["pure_step_iterations_per_s"], dd["folp_create_seconds"], d["e2e"]["value"], [round(x, 3) for x in d["e2e"]["seconds_create_solve_destroy"]], d["e2e"].get("long_solve")))
print("   parity", {k: v for k, v in dd["parity"].items() if k != "rule"})
print("   target", dd.get("target"))
print("   phases", {k[:20]: v for k, v in d["roofline"]["per_kernel"].items() if k != "what"})
print("   exchange", dd.get("exchange"))
PY
grep -E "no multicast|folp_create\] (vectors|TOTAL)" gpurun_out/bench_c2_x${N}_mc.err | sed -n 1,24p
