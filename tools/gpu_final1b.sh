#!/bin/bash
mkdir -p gpurun_out
FOLP_TIMING=1 timeout 900 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_c2.json")); dd = d["detail"]
print("value %.0f pure %.0f create %.2f e2e %.0f %s long %s" % (d["value"], dd["pure_step_iterations_per_s"], dd["folp_create_seconds"], d["e2e"]["value"], [round(x, 3) for x in d["e2e"]["seconds_create_solve_destroy"]], d["e2e"]["long_solve"]))
print("roofline frac %.3f achieved %.0f traffic %s in_loop %s" % (d["roofline"]["frac"], d["roofline"]["achieved"], d["roofline"]["traffic"], d["roofline"]["in_loop"]))
print("parity", {k: v for k, v in dd["parity"].items() if k != "rule"}); print("target", dd["target"]); print("cpu", d["cpu_baseline"]); print("clocks", d["clocks"])
PY
