#!/bin/bash
# round 2, call B: persistent take_step kernel -- parity tests on both forms, then per-phase times
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_persistent.log 2>&1
echo "pytest (persistent) rc=$?"; tail -15 gpurun_out/pytest_gpu_persistent.log
FOLP_NO_PERSISTENT=1 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_legacy.log 2>&1
echo "pytest (kernel per phase) rc=$?"; tail -3 gpurun_out/pytest_gpu_legacy.log
B=firstorderlp.jl_b200/libfolp_b200.so
for w in c2 netlib netlib_small; do
  timeout 600 python tools/probe_kernels.py --workload $w --iters 2000 $B $B:FOLP_NO_PERSISTENT=1 > gpurun_out/probe_${w}_persistent.log 2> gpurun_out/probe_${w}_persistent.err
  echo "probe $w rc=$?"
  python - $w <<'PY'
import json, sys
for l in open("gpurun_out/probe_%s_persistent.log" % sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l)
        print("%-10s %-26s K1 %6.2f K2 %6.2f K3 %6.2f iter %7.2f run %8.0f pure %8.0f" % (
            sys.argv[1], d["env"], d["k_primal_us"], d["k_dual_us"], d["k_trans_us"], d["iter_us"],
            d.get("run_it_per_s", 0), d.get("pure_step_it_per_s", 0)))
PY
done
tail -5 gpurun_out/probe_c2_persistent.err
