#!/bin/bash
mkdir -p gpurun_out
B=firstorderlp.jl_b200/libfolp_b200.so
for w in ${WORKLOADS:-c2}; do
  timeout 900 python tools/probe_kernels.py --workload $w --iters 2000 $B $B:FOLP_BALANCE_TILES=1 $B:FOLP_VAR_SORT_WINDOW=8192 $B:FOLP_VAR_SORT_WINDOW=8192,FOLP_BALANCE_TILES=1 $B:FOLP_VAR_SORT_WINDOW=512 $B:FOLP_VAR_SORT_WINDOW=32768 $B:FOLP_VAR_SORT_WINDOW=1000000000,FOLP_BALANCE_TILES=1 $B > gpurun_out/probe_${w}_varsort2.log 2> gpurun_out/probe_${w}_varsort2.err
  echo "probe $w rc=$?"
  python - $w <<'PY'
import json, sys
for l in open("gpurun_out/probe_%s_varsort2.log" % sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l)
        print("%-10s %-62s K2 %6.2f K3 %6.2f iter %7.2f plainAt %6.2f run %8.0f pure %8.0f" % (
            sys.argv[1], d["env"], d["k_dual_us"], d["k_trans_us"], d["iter_us"], d["plain_At_us"],
            d.get("run_it_per_s", 0), d.get("pure_step_it_per_s", 0)))
PY
done
