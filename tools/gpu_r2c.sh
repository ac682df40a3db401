#!/bin/bash
# persistent kernel iteration: parity tests, then per-phase times against the kernel-per-phase form
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_persistent.log 2>&1
echo "pytest (persistent) rc=$?"; tail -4 gpurun_out/pytest_gpu_persistent.log
B=firstorderlp.jl_b200/libfolp_b200.so
for w in ${WORKLOADS:-c2 netlib netlib_small}; do
  timeout 600 python tools/probe_kernels.py --workload $w --iters 2000 $B $B:FOLP_NO_PERSISTENT=1 $EXTRA > gpurun_out/probe_${w}_persistent.log 2> gpurun_out/probe_${w}_persistent.err
  echo "probe $w rc=$?"
  python - $w <<'PY'
import json, sys
for l in open("gpurun_out/probe_%s_persistent.log" % sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l)
        print("%-10s %-18s %-26s K1 %6.2f K2 %6.2f K3 %6.2f iter %7.2f run %8.0f pure %8.0f" % (
            sys.argv[1], d["lib"], d["env"], d["k_primal_us"], d["k_dual_us"], d["k_trans_us"], d["iter_us"],
            d.get("run_it_per_s", 0), d.get("pure_step_it_per_s", 0)))
PY
done
