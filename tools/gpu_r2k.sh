#!/bin/bash
mkdir -p gpurun_out
B=firstorderlp.jl_b200/libfolp_b200.so
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
SPECS="$B"
for f in scratch/libfolp_*.so; do SPECS="$SPECS $f"; done
SPECS="$SPECS $B:FOLP_VAR_SORT_WINDOW=512 $B:FOLP_VAR_SORT_WINDOW=1024 $B:FOLP_VAR_SORT_WINDOW=4096 $B"
for w in ${WORKLOADS:-c2}; do
  timeout 900 python tools/probe_kernels.py --workload $w --iters 2000 $SPECS > gpurun_out/probe_${w}_varsort3.log 2> gpurun_out/probe_${w}_varsort3.err
  echo "probe $w rc=$?"
  python - $w <<'PY'
import json, sys
for l in open("gpurun_out/probe_%s_varsort3.log" % sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l)
        print("%-10s %-18s %-34s K1 %5.2f K2 %6.2f K3 %6.2f iter %7.2f run %8.0f pure %8.0f" % (
            sys.argv[1], d["lib"], d["env"], d["k_primal_us"], d["k_dual_us"], d["k_trans_us"], d["iter_us"],
            d.get("run_it_per_s", 0), d.get("pure_step_it_per_s", 0)))
PY
done
