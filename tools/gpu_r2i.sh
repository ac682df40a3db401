#!/bin/bash
# variable reordering by column length: parity tests, then A/B per-kernel times on three workloads
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
B=firstorderlp.jl_b200/libfolp_b200.so
for w in ${WORKLOADS:-c2 pagerank netlib}; do
  timeout 600 python tools/probe_kernels.py --workload $w --iters 2000 $B $B:FOLP_NO_VAR_SORT=1 $B > gpurun_out/probe_${w}_varsort.log 2> gpurun_out/probe_${w}_varsort.err
  echo "probe $w rc=$?"
  python - $w <<'PY'
import json, sys
for l in open("gpurun_out/probe_%s_varsort.log" % sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l)
        print("%-10s %-26s K1 %6.2f K2 %6.2f K3 %6.2f iter %7.2f plainAt %6.2f run %8.0f pure %8.0f" % (
            sys.argv[1], d["env"], d["k_primal_us"], d["k_dual_us"], d["k_trans_us"], d["iter_us"], d["plain_At_us"],
            d.get("run_it_per_s", 0), d.get("pure_step_it_per_s", 0)))
PY
done
