#!/bin/bash
# round 2 final, N GPUs: GPU test tier (runs the partitioned tests the box has GPUs for), bench line under torchrun,
# bench line of the single-process form
N=${1:-2}
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
  timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_x$N.log 2>&1
  echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_x$N.log
fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
FOLP_TIMING=1 timeout 1200 $TR bench.py --gpus $N > gpurun_out/bench_c2_x$N.json 2> gpurun_out/bench_c2_x$N.err
echo "bench x$N rc=$?"
timeout 1200 python bench.py --gpus $N --single-process --skip-target > gpurun_out/bench_c2_x${N}_single_process.json 2> gpurun_out/bench_c2_x${N}_single_process.err
echo "bench single-process x$N rc=$?"
python - $N <<'PY'
import json, sys
for tag in ("", "_single_process"):
    try:
        d = json.load(open("gpurun_out/bench_c2_x%s%s.json" % (sys.argv[1], tag)))
    except Exception as e:
        print(tag, "missing", e); continue
    dd = d["detail"]
    print(tag or "torchrun", "value %.0f pure %.0f create %.2f e2e %.0f %s long %s" % (
        d["value"], dd["pure_step_iterations_per_s"], dd["folp_create_seconds"], d["e2e"]["value"],
        [round(x, 3) for x in d["e2e"]["seconds_create_solve_destroy"]], d["e2e"]["long_solve"]))
    print("   parity", {k: v for k, v in dd["parity"].items() if k != "rule"})
    print("   target", dd["target"])
    print("   phases", {k[:20]: v for k, v in d["roofline"]["per_kernel"].items() if k != "what"})
PY
