#!/bin/bash
# kernel / cache-hint / grid variants on the default workload (one process, one problem)
mkdir -p gpurun_out
B=firstorderlp.jl_b200/libfolp_b200.so
SPECS="$B"
for f in scratch/libfolp_*.so; do SPECS="$SPECS $f"; done
SPECS="$SPECS $B"
timeout 900 python tools/probe_kernels.py --workload ${W:-c2} --iters 4000 $SPECS > gpurun_out/probe_${W:-c2}_variants2.log 2> gpurun_out/probe_${W:-c2}_variants2.err
echo "probe rc=$?"
python - <<PY
import json
for l in open("gpurun_out/probe_${W:-c2}_variants2.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("%-22s K1 %5.1f K2 %5.1f K3 %5.1f iter %6.1f plainA %5.1f plainAt %5.1f run %7.0f pure %7.0f" % (
            d["lib"], d["k_primal_us"], d["k_dual_us"], d["k_trans_us"], d["iter_us"], d["plain_A_us"],
            d["plain_At_us"], d.get("run_it_per_s", 0), d.get("pure_step_it_per_s", 0)))
PY
