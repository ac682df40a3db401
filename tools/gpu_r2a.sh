#!/bin/bash
# round 2, call A: baseline tests, kernel variants x balanced order, small-instance regime
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
B=firstorderlp.jl_b200/libfolp_b200.so
SPECS="$B $B:FOLP_BALANCE_TILES=1"
for f in scratch/libfolp_*.so; do SPECS="$SPECS $f $f:FOLP_BALANCE_TILES=1"; done
timeout 900 python tools/probe_kernels.py --workload c2 --iters 2000 $SPECS > gpurun_out/probe_c2_variants.log 2> gpurun_out/probe_c2_variants.err
echo "probe rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/probe_c2_variants.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("%-22s %-24s K1 %5.1f K2 %5.1f K3 %5.1f iter %6.1f plainA %5.1f plainAt %5.1f run %7.0f pure %7.0f" % (
            d["lib"], d["env"], d["k_primal_us"], d["k_dual_us"], d["k_trans_us"], d["iter_us"], d["plain_A_us"],
            d["plain_At_us"], d.get("run_it_per_s", 0), d.get("pure_step_it_per_s", 0)))
PY
for w in netlib_small netlib; do
  timeout 300 python bench.py --workload $w --cpu-iters 4000 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  echo "bench $w rc=$?"; cut -c1-900 gpurun_out/bench_$w.json
done
